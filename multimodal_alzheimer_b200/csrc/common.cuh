// Shared device/host helpers for the adni_b200 kernels (sm_100a only).
//
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA / TMEM),
// plus the host-side error plumbing used by every C-ABI entry point.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../../include/adni_b200.h"

namespace adni {

// ------------------------------------------------------------------------------------------------
// Host-side error handling: every C-ABI call returns 0 or a negative ADNI_E* code and stores a
// message retrievable through adni_last_error_string().
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define ADNI_CUDA_OK(expr)                                         \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return ::adni::cuda_fail(_e, #expr);    \
  } while (0)

#define ADNI_REQUIRE(cond, code, ...)                              \
  do {                                                             \
    if (!(cond)) {                                                 \
      ::adni::set_error(__VA_ARGS__);                              \
      return (code);                                               \
    }                                                              \
  } while (0)

#define ADNI_LAUNCH_CHECK(name)                                    \
  do {                                                             \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return ::adni::cuda_fail(_e, name);     \
  } while (0)

int num_sms();

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();

// Encode a bf16 tensor map of rank `rank` (dims innermost first, strides in elements for dims 1..).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_elems, const uint32_t* box, bool swizzle128);

// Programmatic dependent launch is opt-in (ADNI_PDL=1, runtime.cu): measured 0-4 % SLOWER inside the step's CUDA graph.
bool pdl_enabled();

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A training step is ~400 launches, many of them 5-50 us long at the 8-GPU
// per-rank batch, so the launch-to-launch gap (grid drain, launch latency, barrier / TMEM set-up of the next kernel) is
// a double-digit share of the step.  With ADNI_PDL=1 every kernel of this library is launched with the programmatic-
// stream-serialization attribute; every kernel signals `launch_dependents` first thing and executes `griddepcontrol.wait` before its first
// global-memory access: the next kernel's CTAs are scheduled onto SMs as the current kernel's CTAs retire, run their
// prologue, and block in `wait` until the current grid has completed and flushed - the stream's data dependencies are
// unchanged.  (A dependent grid can only start once ALL CTAs of its primary have started, so a waiting grid never
// keeps an earlier one from being scheduled.)  Stream capture records the attribute as a programmatic graph edge.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_trigger();
  pdl_wait();
}

template <typename... KArgs>
struct PdlLaunch {
  void (*kern)(KArgs...);
  dim3 grid, block;
  size_t smem;
  cudaStream_t stream;
  template <typename... Args>
  void operator()(Args&&... args) const {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);   // errors surface through ADNI_LAUNCH_CHECK
  }
};

// kernel<<<grid, block, smem, stream>>>(args...)  ==  pdl_launch(kernel, grid, block, smem, stream)(args...)
template <typename... KArgs>
PdlLaunch<KArgs...> pdl_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
  return PdlLaunch<KArgs...>{kern, grid, block, smem, stream};
}

// ------------------------------------------------------------------------------------------------
// Device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 %%rx;\n"
      ".reg .pred %%px;\n"
      "     elect.sync %%rx|%%px, %1;\n"
      "@%%px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred;
}

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a dead pipeline traps (-> cudaErrorLaunchFailure) instead of hanging the GPU box.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // NOTE (measured, tools/ubench + wgrad shape profile): the %globaltimer read after the first failed poll acts as a
  // back-off.  Polling try_wait in a tight loop from the first failure on made every tcgen05 kernel slower (wgrad
  // -35 %, fprop -5 %): the polls compete with the TMA complete_tx / tcgen05.commit updates of the same barrier.
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0 && global_timer_ns() - t0 > 4000000000ull) {  // 4 s: the pipeline is dead
      printf("adni_b200: mbarrier timeout block %d thread %d parity %u\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

// Tight polling variant.  Measured better for the producer / MMA-issuer threads of the igemm and halo conv kernels
// (short stages, double-buffered accumulators: fprop+dgrad 27.4 -> 26.3 ms per step) and worse for wgrad2 / stem.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity, int tag = 0) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 4095u) == 0) {
      const uint64_t t = global_timer_ns();
      if (t0 == 0) {
        t0 = t;
      } else if (t - t0 > 4000000000ull) {  // 4 s: the pipeline is dead
        printf("adni_b200: mbarrier timeout (spin) kernel-tag %d block %d thread %d parity %u\n", tag, blockIdx.x, threadIdx.x, parity);
        __trap();
      }
    }
  }
}

// Whole-warp wait for warps that are NOT on the critical path (epilogue warps waiting for an accumulator): lane 0
// polls with a back-off sleep, the other lanes park on the warp barrier.  128 threads spinning on try_wait for a whole
// main loop take issue slots and shared-memory bandwidth from the single TMA / MMA issuing threads (wgrad: -35 %).
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
      __nanosleep(64);
      if ((++spins & 4095u) == 0) {
        const uint64_t t = global_timer_ns();
        if (t0 == 0) {
          t0 = t;
        } else if (t - t0 > 4000000000ull) {
          printf("adni_b200: mbarrier timeout (parked) block %d thread %d parity %u\n", blockIdx.x, threadIdx.x, parity);
          __trap();
        }
      }
    }
  }
  __syncwarp();
}

// ---- TMA ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// TMA store (shared -> global, bulk async-group completion): the issuing thread commits a group and, before the
// staged tile is overwritten, waits until the engine has finished READING it.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d_u32(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_u32(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1);
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d_u32(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---- tcgen05 / TMEM -----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// tcgen05.commit: arrive on an mbarrier once all previously issued MMAs of this thread retire.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32 (cute/arch/mma_sm100_desc.hpp bit layout).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                               // c_format = F32
         | (1u << 7)                             // a_format = BF16
         | (1u << 10)                            // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)        // a_major
         | ((b_mn_major ? 1u : 0u) << 16)        // b_major
         | (static_cast<uint32_t>(N >> 3) << 17) // n_dim
         | (static_cast<uint32_t>(M >> 4) << 24);  // m_dim
}

// Shared-memory matrix descriptor for 128-byte-swizzled operand tiles (rows of 128 B, 1024-B atoms).
//   K-major : SBO = byte distance between 8-row groups (1024 for dense 128-B rows); LBO unused.
//   MN-major: LBO = byte distance between 64-element MN groups; SBO = distance between 8-k groups.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (sm_100)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}
// No-swizzle ("interleave") descriptor: core matrices of 8 rows x 16 B stored contiguously (128 B).
//   K-major : LBO = distance between core matrices adjacent in K, SBO = adjacent in M/N.
//   MN-major: LBO = distance between 8-k groups, SBO = distance between 8-element MN groups.
__device__ __forceinline__ uint64_t umma_smem_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Transpose-reduce: every lane holds v[0..31]; afterwards lane l returns sum over lanes of v[l].
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; i++) {
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}
#endif  // __CUDACC__

}  // namespace adni

// One-shot all-reduce over NVLink peer memory for the tiny fp64 messages of synchronised BatchNorm (sum x, sum x^2
// forward; sum g, sum g*xhat backward; the loss normaliser): 2 x C doubles, 84 + 84 times per two-encoder step.
// A NCCL all-reduce of 1-8 KB is pure launch + protocol latency; here every rank stores its vector straight into a
// slot of every peer's symmetric buffer and sums the world's slots in RANK ORDER (so all ranks obtain bit-identical
// sums).  One CTA, no host involvement, graph-capturable.
//
// Low-latency cells (round 2): a double travels as two 8-byte packets {32 data bits | 32-bit call number}; an 8-byte
// store is single-copy atomic over NVLink, so every packet validates itself and the receiver simply polls the cells it
// needs.  The first version wrote plain doubles, then `__threadfence_system()` (a full NVLink round trip until every
// remote store is acknowledged), then a flag per peer - two traversals plus the fence on the critical path of each of
// the 168 exchanges of a step, which at 8 ranks had become the larger part of the data-parallel overhead.
//
// The reference is single-GPU; this is the exchange that makes N ranks on shards of a
// batch reproduce its full-batch BatchNorm.
//
// Symmetric buffer layout per rank (allocated and exchanged by the host: torch symmetric memory = plumbing):
//   cells  ulonglong2 [2 parities][world][max_n]      (zero-initialised: call numbers start at 1)
// Call k (k = 1, 2, ...) uses parity k & 1.  A peer can be at most one call ahead of a rank that is still reading its
// cells (it needs that rank's packets of call k+1 to finish k+1), so two parities are enough and nothing is ever reset.
#include "common.cuh"

namespace adni {
extern void count_launch();
namespace {

constexpr int kPeerThreads = 256;
constexpr int kPeerMaxWorld = 64;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_cell(ulonglong2* p, unsigned long long lo, unsigned long long hi) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_cell(const ulonglong2* p) {
  ulonglong2 v;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_f64_kernel(double* __restrict__ data, int n,
                                                                          const unsigned long long* __restrict__ peers,
                                                                          unsigned long long* __restrict__ call_counter,
                                                                          int rank, int world, int max_n) {
  pdl_enter();
  const unsigned long long call = *call_counter + 1;
  const unsigned long long tag = (call & 0xffffffffull) << 32;
  const size_t par_off = static_cast<size_t>(call & 1ull) * world * max_n;
  // 1. my vector into cell block [par][rank] of every rank's buffer (peer stores travel over NVLink); every thread
  //    handles the same elements i in both phases, so data[i] is read before it is overwritten
  for (int i = threadIdx.x; i < n; i += kPeerThreads) {
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(data[i]));
    const unsigned long long lo = (bits & 0xffffffffull) | tag, hi = (bits >> 32) | tag;
    for (int r = 0; r < world; r++)
      st_cell(reinterpret_cast<ulonglong2*>(peers[r]) + par_off + static_cast<size_t>(rank) * max_n + i, lo, hi);
  }
  // 2. sum the world's cells in rank order as they arrive
  const ulonglong2* cells = reinterpret_cast<const ulonglong2*>(peers[rank]) + par_off;
  const uint64_t t0 = global_timer_ns();
  for (int i = threadIdx.x; i < n; i += kPeerThreads) {
    double s = 0.0;
    for (int r = 0; r < world; r++) {
      const ulonglong2* c = cells + static_cast<size_t>(r) * max_n + i;
      ulonglong2 v = ld_cell(c);
      uint32_t spins = 0;
      while ((v.x & 0xffffffff00000000ull) != tag || (v.y & 0xffffffff00000000ull) != tag) {
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 8000000000ull) {  // 8 s: a peer is gone
          printf("adni_b200: peer all-reduce timeout rank %d waiting for rank %d call %llu\n", rank, r, call);
          __trap();
        }
        v = ld_cell(c);
      }
      s += __longlong_as_double(static_cast<long long>((v.y << 32) | (v.x & 0xffffffffull)));
    }
    data[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) *call_counter = call;
}

// ------------------------------------------------------------------------------------------------------------------
// Two-shot all-reduce of the fp32 gradient buckets over NVLink peer memory (the reference
// is single-GPU).  Every rank's bucket lives in symmetric memory at the same offset.  Rank r owns the r-th slice of the
// bucket: it loads that slice from every rank (remote loads over NVLink), sums in RANK ORDER and stores the sum back
// into every rank's bucket (remote stores) - reduce-scatter and all-gather of the classic two-shot scheme in one kernel,
// 2 x (W-1)/W x bytes over each GPU's links, both directions busy at once.  A slice is summed by exactly one rank, so all
// ranks end up with bit-identical gradients, and the result is deterministic.
//   phase 0  every rank signals "my gradients are final" (its backward kernels precede this kernel in stream order)
//            and waits for the world's signals before touching peer memory
//   phase 1  grid-stride over the own slice, float4 loads / stores, kPeerUnroll independent elements per thread
//   phase 2  the last CTA of the grid to finish signals "I am done with everybody's memory" and waits for the world's:
//            only then may this rank's stream go on to read (Adam) or overwrite (next backward) its bucket
// No shared memory beyond a pointer table and 48 registers per thread: the CTAs co-reside with the persistent conv
// grids (192 threads + 220 KB of shared memory per SM), which NCCL's CTAs cannot - an all-reduce issued on a side stream
// while backward is still running does not push conv CTAs into a second wave.
// Flag buffer per rank (symmetric): ready[64], done[64] (monotonic call numbers, compared with >=).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kGradThreads = 256;
constexpr int kGradUnroll = 4;

__device__ __forceinline__ float4 ld_cg_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void wait_flags(const unsigned long long* mine, unsigned long long call, int rank, const char* what) {
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (ld_acquire_sys(mine) < call) {
    if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 8000000000ull) {  // 8 s: a peer is gone
      printf("adni_b200: peer gradient all-reduce timeout (%s) rank %d waiting for rank %d call %llu\n", what, rank,
             static_cast<int>(threadIdx.x), call);
      __trap();
    }
  }
}

__global__ void __launch_bounds__(kGradThreads) peer_allreduce_f32_kernel(long long elem_off, long long n4,
                                                                          const unsigned long long* __restrict__ data_peers,
                                                                          const unsigned long long* __restrict__ flag_peers,
                                                                          unsigned long long* __restrict__ call_counter,
                                                                          unsigned int* __restrict__ arrive, int rank, int world) {
  pdl_enter();
  __shared__ float4* base[kPeerMaxWorld];
  __shared__ int is_last;
  // every CTA reads the call number before anything else: the counter is only advanced by the last CTA to FINISH
  const unsigned long long call = *reinterpret_cast<volatile unsigned long long*>(call_counter) + 1;
  if (threadIdx.x < world) base[threadIdx.x] = reinterpret_cast<float4*>(data_peers[threadIdx.x]) + elem_off / 4;
  unsigned long long* my_flags = reinterpret_cast<unsigned long long*>(flag_peers[rank]);
  // ---- phase 0
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long*>(flag_peers[threadIdx.x]) + rank, call);
  }
  if (threadIdx.x < world) wait_flags(my_flags + threadIdx.x, call, rank, "ready");
  __syncthreads();
  // ---- phase 1: my slice of the bucket
  const long long per = (n4 + world - 1) / world;
  const long long begin = per * rank, end = min(n4, begin + per);
  const long long stride = static_cast<long long>(gridDim.x) * kGradThreads;
  for (long long i0 = begin + static_cast<long long>(blockIdx.x) * kGradThreads + threadIdx.x; i0 < end; i0 += stride * kGradUnroll) {
    float4 acc[kGradUnroll];
#pragma unroll
    for (int u = 0; u < kGradUnroll; u++) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; r++) {          // rank order: the sum is the same whoever computes it
      const float4* src = base[r];
#pragma unroll
      for (int u = 0; u < kGradUnroll; u++) {
        const long long i = i0 + u * stride;
        if (i < end) {
          const float4 v = ld_cg_f4(src + i);
          acc[u].x += v.x;
          acc[u].y += v.y;
          acc[u].z += v.z;
          acc[u].w += v.w;
        }
      }
    }
    for (int r = 0; r < world; r++) {
      float4* dst = base[r];
#pragma unroll
      for (int u = 0; u < kGradUnroll; u++) {
        const long long i = i0 + u * stride;
        if (i < end) __stcg(dst + i, acc[u]);
      }
    }
  }
  // ---- phase 2
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(arrive, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (is_last) {
    if (threadIdx.x < world) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned long long*>(flag_peers[threadIdx.x]) + kPeerMaxWorld + rank, call);
      wait_flags(my_flags + kPeerMaxWorld + threadIdx.x, call, rank, "done");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      *arrive = 0u;
      *call_counter = call;
    }
  }
}

}  // namespace
}  // namespace adni

using namespace adni;

extern "C" {

size_t adni_peer_buffer_bytes(int world, int max_n) {
  return sizeof(ulonglong2) * 2 * static_cast<size_t>(world) * static_cast<size_t>(max_n);
}

int adni_peer_allreduce_f64(double* data, int n, const void* peers, void* call_counter, int rank, int world, int max_n,
                            void* stream) {
  ADNI_REQUIRE(data && peers && call_counter && n > 0, ADNI_EINVAL, "peer_allreduce_f64: bad arguments");
  ADNI_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, ADNI_EINVAL,
               "peer_allreduce_f64: rank %d / world %d out of range", rank, world);
  ADNI_REQUIRE(n <= max_n, ADNI_ENOTSUP, "peer_allreduce_f64: %d elements exceed the slot size %d", n, max_n);
  pdl_launch(peer_allreduce_f64_kernel, 1, kPeerThreads, 0, static_cast<cudaStream_t>(stream))(
      data, n, static_cast<const unsigned long long*>(peers), static_cast<unsigned long long*>(call_counter), rank, world,
      max_n);
  count_launch();
  ADNI_LAUNCH_CHECK("peer_allreduce_f64_kernel");
  return ADNI_OK;
}

size_t adni_peer_grad_flag_bytes(void) { return 2 * kPeerMaxWorld * sizeof(unsigned long long); }

int adni_peer_allreduce_f32(long long elem_offset, long long n, const void* data_peers, const void* flag_peers,
                            void* call_counter, void* arrive_counter, int rank, int world, int ctas, void* stream) {
  ADNI_REQUIRE(data_peers && flag_peers && call_counter && arrive_counter && n > 0, ADNI_EINVAL, "peer_allreduce_f32: bad arguments");
  ADNI_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, ADNI_EINVAL,
               "peer_allreduce_f32: rank %d / world %d out of range", rank, world);
  ADNI_REQUIRE(elem_offset >= 0 && elem_offset % 4 == 0 && n % 4 == 0, ADNI_EINVAL,
               "peer_allreduce_f32: offset %lld / length %lld must be multiples of 4 elements", elem_offset, n);
  ADNI_REQUIRE(ctas >= 1 && ctas <= 1024, ADNI_EINVAL, "peer_allreduce_f32: %d CTAs", ctas);
  pdl_launch(peer_allreduce_f32_kernel, ctas, kGradThreads, 0, static_cast<cudaStream_t>(stream))(
      elem_offset, n / 4, static_cast<const unsigned long long*>(data_peers), static_cast<const unsigned long long*>(flag_peers),
      static_cast<unsigned long long*>(call_counter), static_cast<unsigned int*>(arrive_counter), rank, world);
  count_launch();
  ADNI_LAUNCH_CHECK("peer_allreduce_f32_kernel");
  return ADNI_OK;
}

}  // extern "C"

// Micro-benchmark: cost per pipeline stage of the producer/consumer mbarrier ring used by the conv kernels.
// Variants isolate tcgen05.fence::after_thread_sync, warp-wide waits + elect, tcgen05.commit vs plain arrive,
// and MMAs per stage.
#include <cstdio>
#include <vector>
#include "../../multimodal_alzheimer_b200/csrc/common.cuh"
using namespace adni;
namespace adni { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -3; } }

constexpr int STAGES = 8;

template <int N, int MMAS, int FENCE, int WARPWIDE, int COMMIT>
__global__ void __launch_bounds__(128, 1) pipe_kernel(int iters, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[STAGES], empty[STAGES], done;
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_end;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    reinterpret_cast<uint32_t*>(smem)[i] = (0x3C00u | (h & 0x3FFu)) | ((0x3C00u | ((h >> 10) & 0x3FFu)) << 16);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; it++) {
        mbar_wait(&empty[st], ph ^ 1);
        mbar_arrive(&full[st]);
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, false);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    const uint64_t hi = umma_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFF00000000ull;
    const uint32_t lo0 = (uint32_t)(umma_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFFull);
    if (WARPWIDE || lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; it++) {
        mbar_wait(&full[st], ph);
        if (FENCE) tc_fence_after();
        const uint32_t a = lo0 + (((a0 + (it & 3) * 16384) & 0x3FFFF) >> 4), b = lo0 + (((b0 + (it & 1) * 8192) & 0x3FFFF) >> 4);
        if (!WARPWIDE || elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < MMAS; k++) umma_bf16(tmem + (it & 1) * 256, hi | (a + k * 2), hi | (b + k * 2), idesc, (it > 1 || k > 0) ? 1u : 0u);
          if (COMMIT) umma_commit(&empty[st]); else mbar_arrive(&empty[st]);
        }
        if (WARPWIDE) __syncwarp();
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
      if (!WARPWIDE || elect_one_sync()) { umma_commit(&done); }
      if (WARPWIDE) __syncwarp();
      mbar_wait(&done, 0);
      if (lane == 0) t_end = clock64() - t0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t_end;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int MMAS, int FENCE, int WARPWIDE, int COMMIT>
void run(const char* name, int iters) {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  auto k = pipe_kernel<N, MMAS, FENCE, WARPWIDE, COMMIT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<148, 128, 200 * 1024>>>(iters, d);
  k<<<148, 128, 200 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148); cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (auto v : h) avg += v; avg /= 148;
  printf("%-58s N=%3d mmas/stage=%d  cycles per stage = %7.1f  (MMA ideal %5.1f)  (%s)\n", name, N, MMAS, avg / iters, MMAS * N / 2.0, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  //   N  MMAS FENCE WARPWIDE COMMIT
  run<64, 0, 0, 0, 0>("no MMA, plain arrive, lane0 only, no fence", 20000);
  run<64, 0, 1, 0, 0>("no MMA, plain arrive, lane0 only, fence", 20000);
  run<64, 0, 0, 0, 1>("no MMA, tcgen05.commit, lane0 only, no fence", 20000);
  run<64, 0, 1, 0, 1>("no MMA, tcgen05.commit, lane0 only, fence", 20000);
  run<64, 0, 1, 1, 1>("no MMA, tcgen05.commit, warp-wide+elect, fence", 20000);
  run<64, 4, 0, 0, 1>("4 MMA, tcgen05.commit, lane0 only, no fence", 20000);
  run<64, 4, 1, 0, 1>("4 MMA, tcgen05.commit, lane0 only, fence", 20000);
  run<64, 4, 1, 1, 1>("4 MMA, tcgen05.commit, warp-wide+elect, fence", 20000);
  run<64, 4, 0, 1, 1>("4 MMA, tcgen05.commit, warp-wide+elect, no fence", 20000);
  run<128, 4, 1, 1, 1>("4 MMA, tcgen05.commit, warp-wide+elect, fence", 20000);
  run<128, 4, 0, 1, 1>("4 MMA, tcgen05.commit, warp-wide+elect, no fence", 20000);
  run<256, 4, 1, 1, 1>("4 MMA, tcgen05.commit, warp-wide+elect, fence", 20000);
  run<256, 4, 0, 1, 1>("4 MMA, tcgen05.commit, warp-wide+elect, no fence", 20000);
  run<64, 8, 1, 1, 1>("8 MMA, tcgen05.commit, warp-wide+elect, fence", 10000);
  return 0;
}

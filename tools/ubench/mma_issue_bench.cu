// Micro-benchmark: does tcgen05.mma issue throughput scale with the number of issuing threads (N=64 tiles)?
#include <cstdio>
#include <vector>
#include "../../multimodal_alzheimer_b200/csrc/common.cuh"
using namespace adni;
namespace adni { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -3; } }

template <int N, int ISSUERS, int COMMIT_EVERY>
__global__ void __launch_bounds__(128, 1) bench_kernel(int iters, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint64_t ring[4][8];
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_end[4];
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    reinterpret_cast<uint32_t*>(smem)[i] = (0x3C00u | (h & 0x3FFu)) | ((0x3C00u | ((h >> 10) & 0x3FFu)) << 16);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; i++) { mbar_init(&bar[i], 1); for (int j = 0; j < 8; j++) mbar_init(&ring[i][j], 1); }
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
  if (lane == 0 && warp < ISSUERS) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, false, false);
    const uint32_t a0 = smem_u32(smem) + warp * 16384, b0 = smem_u32(smem + 96 * 1024);
    const uint64_t hi = umma_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFF00000000ull;
    const uint32_t lo0 = (uint32_t)(umma_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFFull);
    int slot = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; it++) {
      const uint32_t a = lo0 + (((a0 + (it & 1) * 65536) & 0x3FFFF) >> 4), b = lo0 + (((b0 + (it & 1) * 8192) & 0x3FFFF) >> 4);
#pragma unroll
      for (int k = 0; k < 4; k++) umma_bf16(tmem + warp * 128 + (it & 1) * 64, hi | (a + k * 2), hi | (b + k * 2), idesc, (it > 1 || k > 0) ? 1u : 0u);
      if (COMMIT_EVERY && (it % COMMIT_EVERY) == COMMIT_EVERY - 1) {
        umma_commit(&ring[warp][slot]);
        if (COMMIT_EVERY == 2) {  // also wait for the commit issued 4 rounds ago (pipeline-like)
        }
        if (++slot == 8) { slot = 0; ph ^= 1; }
      }
    }
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    t_end[warp] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x == 0) { long long m = 0; for (int i = 0; i < ISSUERS; i++) m = m > t_end[i] ? m : t_end[i]; out_cycles[blockIdx.x] = m; }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int ISSUERS, int CE>
void run(const char* name, int iters) {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  auto k = bench_kernel<N, ISSUERS, CE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<148, 128, 200 * 1024>>>(iters, d);
  k<<<148, 128, 200 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148); cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (auto v : h) avg += v; avg /= 148;
  double per = avg / (iters * 4.0 * ISSUERS);
  printf("%-44s N=%3d issuers=%d  cycles per MMA (SM-wide) = %6.1f  ideal %5.1f  (%s)\n", name, N, ISSUERS, per, N / 2.0, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 1, 0>("1 issuer, no commits", 20000);
  run<64, 2, 0>("2 issuers, no commits", 20000);
  run<64, 4, 0>("4 issuers, no commits", 20000);
  run<64, 1, 1>("1 issuer, commit every 4 MMAs", 20000);
  run<64, 2, 1>("2 issuers, commit every 4 MMAs", 20000);
  run<128, 1, 1>("1 issuer, commit every 4 MMAs", 20000);
  run<128, 2, 1>("2 issuers, commit every 4 MMAs", 20000);
  run<256, 1, 1>("1 issuer, commit every 4 MMAs", 20000);
  return 0;
}

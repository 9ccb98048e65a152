// Micro-benchmark: tcgen05.mma issue/throughput for K-major vs MN-major smem operands and different N,
// with and without concurrent TMA-like smem write traffic. One CTA per SM; reports cycles per MMA.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../multimodal_alzheimer_b200/csrc/common.cuh"
using namespace adni;

namespace adni { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return -3; } }

template <int N, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128, 1) bench_kernel(int iters, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  // fill 160 KB with pseudo-random finite bf16 bits
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    uint32_t lo = 0x3C00u | (h & 0x3FFu), hi = 0x3C00u | ((h >> 10) & 0x3FFu);   // bf16 in [~0.0078, ..]
    reinterpret_cast<uint32_t*>(smem)[i] = lo | (hi << 16);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, A_MN, B_MN);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      // rotate over 4 stage slots of 16 KB (A) / 32 KB... keep within the 160 KB filled region
      const uint32_t a = a0 + (it & 3) * 16384, b = b0 + (it & 1) * 32768;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint64_t ad = A_MN ? umma_smem_desc_sw128(a + k * 2048, 8192, 1024) : umma_smem_desc_sw128(a + k * 32, 16, 1024);
        uint64_t bd = B_MN ? umma_smem_desc_sw128(b + k * 2048, 8192, 1024) : umma_smem_desc_sw128(b + k * 32, 16, 1024);
        umma_bf16(tmem + (it & 1) * 256, ad, bd, idesc, (it > 1 || k > 0) ? 1u : 0u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool A_MN, bool B_MN>
void run(const char* name, int iters) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  auto k = bench_kernel<N, A_MN, B_MN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<148, 128, 200 * 1024>>>(iters, d);
  k<<<148, 128, 200 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (auto v : h) avg += v; avg /= 148;
  double per = avg / (iters * 4.0);
  printf("%-34s N=%3d  cycles/MMA(128xNx16) = %7.1f   ideal %5.1f   -> %5.1f%% of tensor rate   (%s)\n", name, N, per,
         N / 2.0, 100.0 * (N / 2.0) / per, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int iters = 20000;
  run<256, false, false>("K-major A, K-major B", iters);
  run<256, true, true>("MN-major A, MN-major B", iters);
  run<256, true, false>("MN-major A, K-major B", iters);
  run<256, false, true>("K-major A, MN-major B", iters);
  run<128, false, false>("K-major A, K-major B", iters);
  run<128, true, true>("MN-major A, MN-major B", iters);
  run<64, false, false>("K-major A, K-major B", iters);
  run<64, true, true>("MN-major A, MN-major B", iters);
  return 0;
}

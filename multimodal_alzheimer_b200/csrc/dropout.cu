// Training-mode nn.Dropout on the CUDA path (Small_PET_CNN / PET_MRI_EF / PET_MRI_FMF stem stacks and dense tails;
// reference pkg/models/pet_models/pet_cnn.py:26-27,38-40, fusion_models/early_fusion.py:42-43).
//
// The keep mask is never stored: it is a pure function of (seed, *offset, element index) through Philox4x32-10, so
// the backward pass regenerates it from the same three numbers.  `offset` is read from DEVICE memory (a one-element
// int64 counter the host wrapper bumps after every forward): a captured CUDA graph therefore draws a fresh mask on
// every replay.  HBM-bound: one read + one write per element, 8 elements (one Philox pair) per thread.
#include "common.cuh"

namespace adni {
extern void count_launch();

namespace {

struct U4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                            uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  return U4{c0, c1, c2, c3};
}

// keep decision for 8 consecutive elements of vector `vec`: bit j of the result = element j kept.
// uniform in [0,1) from the top 24 bits; kept iff u >= p  (P[keep] = 1 - p to 2^-24).
__device__ __forceinline__ uint32_t keep_bits8(unsigned long long seed, unsigned long long offset, long long vec,
                                               uint32_t thresh24) {
  const uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  const uint32_t c2 = static_cast<uint32_t>(offset), c3 = static_cast<uint32_t>(offset >> 32);
  const unsigned long long ctr = static_cast<unsigned long long>(vec) * 2ull;
  const U4 a = philox4x32_10(static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), c2, c3, k0, k1);
  const U4 b = philox4x32_10(static_cast<uint32_t>(ctr + 1), static_cast<uint32_t>((ctr + 1) >> 32), c2, c3, k0, k1);
  const uint32_t r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) bits |= ((r[j] >> 8) >= thresh24 ? 1u : 0u) << j;
  return bits;
}

__global__ void __launch_bounds__(256) dropout_bf16_kernel(const __nv_bfloat16* __restrict__ x,
                                                           __nv_bfloat16* __restrict__ y, long long n,
                                                           uint32_t thresh24, float scale, unsigned long long seed,
                                                           const long long* __restrict__ offset_dev) {
  pdl_enter();
  const unsigned long long offset = static_cast<unsigned long long>(*offset_dev);
  const long long nvec = (n + 7) / 8;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t bits = keep_bits8(seed, offset, v, thresh24);
    const long long base = v * 8;
    if (base + 8 <= n) {
      const uint4 in = *reinterpret_cast<const uint4*>(x + base);
      const uint32_t w[4] = {in.x, in.y, in.z, in.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const float lo = (bits >> (2 * e)) & 1u ? __uint_as_float(w[e] << 16) * scale : 0.f;
        const float hi = (bits >> (2 * e + 1)) & 1u ? __uint_as_float(w[e] & 0xFFFF0000u) * scale : 0.f;
        const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
        o[e] = *reinterpret_cast<const uint32_t*>(&p);
      }
      *reinterpret_cast<uint4*>(y + base) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
      for (int j = 0; base + j < n; j++)
        y[base + j] = __float2bfloat16_rn((bits >> j) & 1u ? __bfloat162float(x[base + j]) * scale : 0.f);
    }
  }
}

__global__ void __launch_bounds__(256) dropout_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                          uint32_t thresh24, float scale, unsigned long long seed,
                                                          const long long* __restrict__ offset_dev) {
  pdl_enter();
  const unsigned long long offset = static_cast<unsigned long long>(*offset_dev);
  const long long nvec = (n + 7) / 8;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t bits = keep_bits8(seed, offset, v, thresh24);
    const long long base = v * 8;
    for (int j = 0; j < 8 && base + j < n; j++) y[base + j] = (bits >> j) & 1u ? x[base + j] * scale : 0.f;
  }
}

}  // namespace
}  // namespace adni

using namespace adni;

extern "C" int adni_dropout(const void* x, void* y, long long n, int is_f32, double p, unsigned long long seed,
                            const long long* offset_dev, void* stream) {
  ADNI_REQUIRE(x && y && offset_dev && n >= 0, ADNI_EINVAL, "dropout: null pointer");
  ADNI_REQUIRE(p >= 0.0 && p < 1.0, ADNI_EINVAL, "dropout: p must be in [0, 1) (got %f)", p);
  if (n == 0) return ADNI_OK;
  ADNI_REQUIRE(is_f32 || (reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0),
               ADNI_EINVAL, "dropout: bf16 tensors must be 16-byte aligned");
  const uint32_t thresh24 = static_cast<uint32_t>(p * 16777216.0 + 0.5);
  const float scale = static_cast<float>(1.0 / (1.0 - p));
  const long long nvec = (n + 7) / 8;
  const long long want = (nvec + 255) / 256;
  const int grid = static_cast<int>(want < 148LL * 16 ? want : 148LL * 16);
  auto st = static_cast<cudaStream_t>(stream);
  if (is_f32)
    pdl_launch(dropout_f32_kernel, grid, 256, 0, st)(static_cast<const float*>(x), static_cast<float*>(y), n, thresh24, scale,
                                             seed, offset_dev);
  else
    pdl_launch(dropout_bf16_kernel, grid, 256, 0, st)(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), n,
                                              thresh24, scale, seed, offset_dev);
  count_launch();
  ADNI_LAUNCH_CHECK("dropout_kernel");
  return ADNI_OK;
}

// Non-overlapping max pooling (kernel == stride, no padding) for the small-CNN stacks, with the ReLU that precedes it
// fused in:  Conv3d -> ReLU -> MaxPool3d(2)  (pkg/models/pet_models/pet_cnn.py:21-25, fusion_models/early_fusion.py:37-41,
// fusion_models/anat_pet_featuremapfusion.py:43-62) and  ... -> BatchNorm3d -> ReLU -> MaxPool3d(2) (the generic
// adni_maxpool3d_bwd dispatches here when the windows do not overlap).
//
//   forward   p = relu(max_window y) = max_window relu(y);  arg-max = first maximum of y in (d, h, w) scan order (the
//             element torch's MaxPool3d picks on relu(y) whenever the maximum is positive; when it is zero the
//             gradient is killed by relu'(0) = 0 whichever element is chosen)
//   backward  dy[i] = dp[window(i)] if i is the window's arg-max and (no ReLU or p[window] > 0), else 0
// The stored ReLU output (1 GB per batch of 32 128^3 volumes at 8 channels) never exists, and every input voxel is
// written exactly once by one thread: HBM-bound, one 16-byte vector (8 channels) per thread.
#include <algorithm>

#include "common.cuh"

namespace adni {
extern void count_launch();

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  v[0] = bf16_lo(r.x), v[1] = bf16_hi(r.x), v[2] = bf16_lo(r.y), v[3] = bf16_hi(r.y);
  v[4] = bf16_lo(r.z), v[5] = bf16_hi(r.z), v[6] = bf16_lo(r.w), v[7] = bf16_hi(r.w);
}

template <int K>
__global__ void __launch_bounds__(kThreads) pool_nonoverlap_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int D,
                                                                       int H, int W, int vpr, int Do, int Ho, int Wo,
                                                                       int relu, __nv_bfloat16* __restrict__ y,
                                                                       uint8_t* __restrict__ am) {
  pdl_enter();
  const long long total = static_cast<long long>(N) * Do * Ho * Wo * vpr;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int cv = static_cast<int>(i % vpr);
    long long r = i / vpr;
    const int ow = static_cast<int>(r % Wo);
    r /= Wo;
    const int oh = static_cast<int>(r % Ho);
    r /= Ho;
    const int od = static_cast<int>(r % Do);
    const int n = static_cast<int>(r / Do);
    float best[8];
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      best[j] = -INFINITY;
      arg[j] = 0;
    }
    uint4 v[K * K * K];
#pragma unroll
    for (int kd = 0; kd < K; kd++)
#pragma unroll
      for (int kh = 0; kh < K; kh++)
#pragma unroll
        for (int kw = 0; kw < K; kw++) {
          const long long off =
              ((((static_cast<long long>(n) * D + od * K + kd) * H + oh * K + kh) * W + ow * K + kw) * vpr + cv) * 8;
          v[(kd * K + kh) * K + kw] = __ldg(reinterpret_cast<const uint4*>(x + off));
        }
#pragma unroll
    for (int s = 0; s < K * K * K; s++) {
      float f[8];
      unpack8(v[s], f);
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (f[j] > best[j] || (s == 0 && !(f[j] < best[j]))) {   // strict >: the first maximum wins; slot 0 also takes NaN / -inf
          best[j] = f[j];
          arg[j] = s;
        }
    }
    uint4 o;
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; j++) best[j] = fmaxf(best[j], 0.f);
    }
    o.x = pack_bf16x2(best[0], best[1]);
    o.y = pack_bf16x2(best[2], best[3]);
    o.z = pack_bf16x2(best[4], best[5]);
    o.w = pack_bf16x2(best[6], best[7]);
    *reinterpret_cast<uint4*>(y + i * 8) = o;
    uint2 a;
    a.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
    a.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
    *reinterpret_cast<uint2*>(am + i * 8) = a;
  }
}

__global__ void __launch_bounds__(kThreads) pool_nonoverlap_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                       const uint8_t* __restrict__ am,
                                                                       const __nv_bfloat16* __restrict__ pooled, int N,
                                                                       int D, int H, int W, int vpr, int k, int Do, int Ho,
                                                                       int Wo, __nv_bfloat16* __restrict__ dx) {
  pdl_enter();
  const long long total = static_cast<long long>(N) * D * H * W * vpr;
  for (long long i = blockIdx.x * static_cast<long long>(kThreads) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * kThreads) {
    const int cv = static_cast<int>(i % vpr);
    long long r = i / vpr;
    const int iw = static_cast<int>(r % W);
    r /= W;
    const int ih = static_cast<int>(r % H);
    r /= H;
    const int id = static_cast<int>(r % D);
    const int n = static_cast<int>(r / D);
    const int od = id / k, oh = ih / k, ow = iw / k;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (od < Do && oh < Ho && ow < Wo) {   // floor mode: the ragged tail of an axis belongs to no window
      const int slot = ((id - od * k) * k + (ih - oh * k)) * k + (iw - ow * k);
      const long long wi = (((static_cast<long long>(n) * Do + od) * Ho + oh) * Wo + ow) * vpr + cv;
      const uint2 a = __ldg(reinterpret_cast<const uint2*>(am + wi * 8));
      const uint4 g = __ldg(reinterpret_cast<const uint4*>(dy + wi * 8));
      uint4 pv = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);   // 1.0: no ReLU mask
      if (pooled != nullptr) pv = __ldg(reinterpret_cast<const uint4*>(pooled + wi * 8));
      const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
      const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
      uint32_t ow4[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const uint32_t aw = e < 2 ? a.x : a.y;
        const int a0 = (aw >> ((2 * e & 3) * 8)) & 255, a1 = (aw >> (((2 * e + 1) & 3) * 8)) & 255;
        const bool k0 = a0 == slot && bf16_lo(pw[e]) > 0.f, k1 = a1 == slot && bf16_hi(pw[e]) > 0.f;
        ow4[e] = (k0 ? (gw[e] & 0x0000FFFFu) : 0u) | (k1 ? (gw[e] & 0xFFFF0000u) : 0u);
      }
      o = make_uint4(ow4[0], ow4[1], ow4[2], ow4[3]);
    }
    *reinterpret_cast<uint4*>(dx + i * 8) = o;
  }
}

int grid_for(long long total) {
  const long long want = (total + kThreads - 1) / kThreads;
  return static_cast<int>(std::min<long long>(want, 148LL * 32));
}

}  // namespace

int launch_pool_nonoverlap_bwd(const __nv_bfloat16* dy, const uint8_t* am, const __nv_bfloat16* pooled, int N, int D, int H,
                               int W, int C, int k, __nv_bfloat16* dx, cudaStream_t st) {
  const int Do = D / k, Ho = H / k, Wo = W / k;
  const long long total = static_cast<long long>(N) * D * H * W * (C / 8);
  pdl_launch(pool_nonoverlap_bwd_kernel, grid_for(total), kThreads, 0, st)(dy, am, pooled, N, D, H, W, C / 8, k, Do, Ho, Wo, dx);
  count_launch();
  ADNI_LAUNCH_CHECK("pool_nonoverlap_bwd_kernel");
  return ADNI_OK;
}

}  // namespace adni

using namespace adni;

extern "C" {

int adni_relu_maxpool_fwd(const adni_bf16* x, int N, int D, int H, int W, int C, int k, int relu, adni_bf16* y,
                          uint8_t* argmax, void* stream) {
  ADNI_REQUIRE(x && y && argmax && N > 0, ADNI_EINVAL, "relu_maxpool_fwd: null pointer");
  ADNI_REQUIRE(C % 8 == 0 && (k == 2 || k == 3) && D >= k && H >= k && W >= k, ADNI_ENOTSUP,
               "relu_maxpool_fwd: unsupported C=%d k=%d (%dx%dx%d)", C, k, D, H, W);
  const int Do = D / k, Ho = H / k, Wo = W / k;
  const long long total = static_cast<long long>(N) * Do * Ho * Wo * (C / 8);
  auto xs = reinterpret_cast<const __nv_bfloat16*>(x);
  auto ys = reinterpret_cast<__nv_bfloat16*>(y);
  auto st = static_cast<cudaStream_t>(stream);
  if (k == 2)
    pdl_launch(pool_nonoverlap_fwd_kernel<2>, grid_for(total), kThreads, 0, st)(xs, N, D, H, W, C / 8, Do, Ho, Wo, relu, ys, argmax);
  else
    pdl_launch(pool_nonoverlap_fwd_kernel<3>, grid_for(total), kThreads, 0, st)(xs, N, D, H, W, C / 8, Do, Ho, Wo, relu, ys, argmax);
  count_launch();
  ADNI_LAUNCH_CHECK("pool_nonoverlap_fwd_kernel");
  return ADNI_OK;
}

int adni_relu_maxpool_bwd(const adni_bf16* dy, const uint8_t* argmax, const adni_bf16* pooled, int N, int D, int H, int W,
                          int C, int k, adni_bf16* dx, void* stream) {
  ADNI_REQUIRE(dy && argmax && dx && N > 0, ADNI_EINVAL, "relu_maxpool_bwd: null pointer");
  ADNI_REQUIRE(C % 8 == 0 && k >= 1 && k * k * k <= 255 && D >= k && H >= k && W >= k, ADNI_ENOTSUP,
               "relu_maxpool_bwd: unsupported C=%d k=%d", C, k);
  return launch_pool_nonoverlap_bwd(reinterpret_cast<const __nv_bfloat16*>(dy), argmax,
                                    reinterpret_cast<const __nv_bfloat16*>(pooled), N, D, H, W, C, k,
                                    reinterpret_cast<__nv_bfloat16*>(dx), static_cast<cudaStream_t>(stream));
}

}  // extern "C"

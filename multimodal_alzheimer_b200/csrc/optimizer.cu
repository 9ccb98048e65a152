// Multi-tensor Adam (+ L2) with per-tensor learning rate: the step that follows the backward pass in every
// configure_optimizers of the path (pkg/models/mri_models/anat_cnn.py:111-136, fusion_models/anat_pet_fusion.py:94-127,
// fusion_models/all_modalities_fusion.py:98-137 - one Adam param group per tensor, encoder tensors at lr_pretrained).
//
// HBM bound: 16 B read (p, g, m, v) + 12 B written (p, m, v) per parameter; a two-ResNet-18 fusion model has 66.4 M
// parameters in ~130 tensors = 1.86 GB per step (0.28 ms at the measured 6.5 TB/s).  The tensor table travels BY VALUE
// in the kernel parameters (no device-side job table to refresh: gradient tensors are re-allocated every step, and a
// parameter block is baked into a CUDA-graph node at capture, where the addresses are stable).
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace adni {
extern void count_launch();
namespace {

constexpr int kAdamMaxTensors = 60;   // per launch: 60 x 58 B + 62 x 4 B + 16 B = 3.7 KB of kernel parameters (limit 4 KB)
constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = kAdamThreads * 4 * 8;  // elements per block: 8 float4 per thread

struct AdamBatch {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  float* step[kAdamMaxTensors];        // per-tensor step counter (fp32, as torch.optim.Adam keeps it)
  long long n[kAdamMaxTensors];
  float lr[kAdamMaxTensors];
  float wd[kAdamMaxTensors];
  int chunk_begin[kAdamMaxTensors + 1];  // first block of tensor i; [count] = total blocks
  int count;
  // optional device-side hyper-parameter table fp32 [2][hyper_n] (row 0 lr, row 1 weight decay), entry hyper_index[i]
  // for tensor i: read at run time, so a CUDA graph replays with the learning rate the host last uploaded
  // (ReduceLROnPlateau, anat_cnn.py:131-135).  Null -> the by-value lr / wd above.
  const float* hyper;
  int hyper_n;
  short hyper_index[kAdamMaxTensors];
};

struct AdamScalars {
  float b1, b2, one_minus_b1, one_minus_b2, eps;
  double beta1, beta2;
};

// torch.optim.Adam (single-tensor path, amsgrad = False, maximize = False), same operation order:
//   g' = g + wd p;  m += (g' - m)(1 - b1);  v = b2 v + (1 - b2) g' g';
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float wd, const AdamScalars& s,
                                          float step_size, float bc2_sqrt) {
  g = fmaf(wd, p, g);
  m = fmaf(g - m, s.one_minus_b1, m);
  v = fmaf(s.one_minus_b2 * g, g, s.b2 * v);
  const float denom = sqrtf(v) / bc2_sqrt + s.eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(kAdamThreads) adam_multi_kernel(const __grid_constant__ AdamBatch tb,
                                                                  const AdamScalars s) {
  pdl_enter();
  // block -> tensor: binary search over the (<= 65 entry) chunk table in the constant bank
  int lo = 0, hi = tb.count - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((int)blockIdx.x >= tb.chunk_begin[mid]) lo = mid; else hi = mid - 1;
  }
  const int i = lo;
  const long long n = tb.n[i];
  const long long begin = (long long)((int)blockIdx.x - tb.chunk_begin[i]) * kAdamChunk;
  const long long end = begin + kAdamChunk < n ? begin + kAdamChunk : n;
  float* __restrict__ p = tb.p[i];
  const float* __restrict__ g = tb.g[i];
  float* __restrict__ m = tb.m[i];
  float* __restrict__ v = tb.v[i];
  const float wd = tb.hyper ? __ldg(tb.hyper + tb.hyper_n + tb.hyper_index[i]) : tb.wd[i];

  __shared__ float sh[2];
  if (threadIdx.x == 0) {
    // the step counter is bumped by adam_bump_kernel AFTER every block of this launch has read it (stream order)
    const double t = (double)__ldg(tb.step[i]) + 1.0;
    const double bc1 = 1.0 - pow(s.beta1, t), bc2 = 1.0 - pow(s.beta2, t);
    const float lr = tb.hyper ? __ldg(tb.hyper + tb.hyper_index[i]) : tb.lr[i];
    sh[0] = (float)((double)lr / bc1);
    sh[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = sh[0], bc2_sqrt = sh[1];

  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    const long long end4 = begin + ((end - begin) & ~3LL);
    for (long long e = begin + threadIdx.x * 4; e < end4; e += kAdamThreads * 4) {
      float4 P = *reinterpret_cast<const float4*>(p + e);
      const float4 G = __ldcs(reinterpret_cast<const float4*>(g + e));  // the gradient is dead after this step
      float4 M = *reinterpret_cast<const float4*>(m + e);
      float4 V = *reinterpret_cast<const float4*>(v + e);
      adam_elem(P.x, G.x, M.x, V.x, wd, s, step_size, bc2_sqrt);
      adam_elem(P.y, G.y, M.y, V.y, wd, s, step_size, bc2_sqrt);
      adam_elem(P.z, G.z, M.z, V.z, wd, s, step_size, bc2_sqrt);
      adam_elem(P.w, G.w, M.w, V.w, wd, s, step_size, bc2_sqrt);
      *reinterpret_cast<float4*>(p + e) = P;
      *reinterpret_cast<float4*>(m + e) = M;
      *reinterpret_cast<float4*>(v + e) = V;
    }
    for (long long e = end4 + threadIdx.x; e < end; e += kAdamThreads) {
      float P = p[e], M = m[e], V = v[e];
      adam_elem(P, g[e], M, V, wd, s, step_size, bc2_sqrt);
      p[e] = P, m[e] = M, v[e] = V;
    }
  } else {
    for (long long e = begin + threadIdx.x; e < end; e += kAdamThreads) {
      float P = p[e], M = m[e], V = v[e];
      adam_elem(P, g[e], M, V, wd, s, step_size, bc2_sqrt);
      p[e] = P, m[e] = M, v[e] = V;
    }
  }
}

__global__ void adam_bump_kernel(const __grid_constant__ AdamBatch tb) {
  pdl_enter();
  const int i = threadIdx.x;
  if (i < tb.count) *tb.step[i] += 1.0f;
}

}  // namespace
}  // namespace adni

using namespace adni;

extern "C" {

int adni_adam_max_tensors_per_launch(void) { return kAdamMaxTensors; }

int adni_adam_step_multi(int n_tensors, void* const* params, const void* const* grads, void* const* exp_avg,
                         void* const* exp_avg_sq, void* const* steps, const long long* numel, const float* lr,
                         const float* weight_decay, const float* hyper_dev, double beta1, double beta2, double eps,
                         void* stream) {
  ADNI_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && steps && numel && lr &&
                                                     weight_decay)),
               ADNI_EINVAL, "adam_step_multi: null table");
  ADNI_REQUIRE(hyper_dev == nullptr || n_tensors <= 32767, ADNI_ENOTSUP, "adam_step_multi: device hyper table holds <= 32767 tensors");
  ADNI_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, ADNI_EINVAL,
               "adam_step_multi: betas must lie in [0, 1) and eps must be >= 0 (torch.optim.Adam raises ValueError)");
  AdamScalars s;
  s.beta1 = beta1, s.beta2 = beta2;
  s.b1 = (float)beta1, s.b2 = (float)beta2;
  s.one_minus_b1 = (float)(1.0 - beta1), s.one_minus_b2 = (float)(1.0 - beta2);
  s.eps = (float)eps;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int j = 0;
  while (j < n_tensors) {
    AdamBatch tb;
    memset(&tb, 0, sizeof(tb));
    int cnt = 0, blocks = 0;
    for (; j < n_tensors && cnt < kAdamMaxTensors; j++) {
      ADNI_REQUIRE(numel[j] >= 0, ADNI_EINVAL, "adam_step_multi: negative numel for tensor %d", j);
      if (numel[j] == 0) continue;
      ADNI_REQUIRE(params[j] && grads[j] && exp_avg[j] && exp_avg_sq[j] && steps[j], ADNI_EINVAL,
                   "adam_step_multi: null pointer for tensor %d", j);
      const long long chunks = (numel[j] + kAdamChunk - 1) / kAdamChunk;
      ADNI_REQUIRE(chunks + blocks <= 0x7fffffffLL, ADNI_ENOTSUP, "adam_step_multi: too many elements in one launch");
      tb.p[cnt] = static_cast<float*>(params[j]);
      tb.g[cnt] = static_cast<const float*>(grads[j]);
      tb.m[cnt] = static_cast<float*>(exp_avg[j]);
      tb.v[cnt] = static_cast<float*>(exp_avg_sq[j]);
      tb.step[cnt] = static_cast<float*>(steps[j]);
      tb.n[cnt] = numel[j];
      tb.lr[cnt] = lr[j];
      tb.wd[cnt] = weight_decay[j];
      tb.hyper_index[cnt] = static_cast<short>(j);
      tb.chunk_begin[cnt] = blocks;
      blocks += (int)chunks;
      cnt++;
    }
    if (cnt == 0) continue;
    tb.chunk_begin[cnt] = blocks;
    tb.count = cnt;
    tb.hyper = hyper_dev;
    tb.hyper_n = n_tensors;
    pdl_launch(adam_multi_kernel, blocks, kAdamThreads, 0, st)(tb, s);
    count_launch();
    ADNI_LAUNCH_CHECK("adam_multi_kernel");
    pdl_launch(adam_bump_kernel, 1, kAdamMaxTensors, 0, st)(tb);
    count_launch();
    ADNI_LAUNCH_CHECK("adam_bump_kernel");
  }
  return ADNI_OK;
}

}  // extern "C"

// Latency-bound tail of the step: Linear (+ReLU) heads, BatchNorm1d, and the fp64 focal / weighted-CE loss
// with its gradient.  Shapes are (B <= a few hundred) x (<= 2048) — one or two small
// launches each, warp-shuffle and shared-memory reductions, no library GEMM.
#include "common.cuh"

namespace adni {
extern void count_launch();
namespace {

// one warp per (b, o)
__global__ void linear_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W,
                                  const float* __restrict__ bias, float* __restrict__ y, int ldy, int B, int in,
                                  int out, int relu) {
  pdl_enter();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * out) return;
  const int b = warp / out, o = warp % out;
  const float* xr = x + (long long)b * ldx;
  const float* wr = W + (long long)o * in;
  float s = 0.f;
  for (int i = lane; i < in; i += 32) s = fmaf(xr[i], wr[i], s);
  s = warp_sum(s);
  if (lane == 0) {
    if (bias) s += bias[o];
    if (relu) s = fmaxf(s, 0.f);
    y[(long long)b * ldy + o] = s;
  }
}

__device__ __forceinline__ float masked_dy(const float* dy, int lddy, const float* y, int ldy, int b, int o, int relu) {
  const float g = dy[(long long)b * lddy + o];
  return (relu && !(y[(long long)b * ldy + o] > 0.f)) ? 0.f : g;
}

// dx[b][i] = sum_o g[b][o] W[o][i]
__global__ void linear_bwd_dx_kernel(const float* __restrict__ W, const float* __restrict__ y, int ldy,
                                     const float* __restrict__ dy, int lddy, float* __restrict__ dx, int lddx,
                                     int accumulate, int B, int in, int out, int relu) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * in) return;
  const int b = idx / in, i = idx % in;
  float s = 0.f;
  for (int o = 0; o < out; o++) s = fmaf(masked_dy(dy, lddy, y, ldy, b, o, relu), W[(long long)o * in + i], s);
  float* d = dx + (long long)b * lddx + i;
  *d = accumulate ? *d + s : s;
}

// dW[o][i] += sum_b g[b][o] x[b][i];  db[o] += sum_b g[b][o]   (thread i == in handles the bias column)
__global__ void linear_bwd_dw_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ y, int ldy,
                                     const float* __restrict__ dy, int lddy, float* __restrict__ dW,
                                     float* __restrict__ db, int B, int in, int out, int relu) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)out * (in + 1)) return;
  const int o = (int)(idx / (in + 1)), i = (int)(idx % (in + 1));
  float s = 0.f;
  if (i < in) {
    for (int b = 0; b < B; b++) s = fmaf(masked_dy(dy, lddy, y, ldy, b, o, relu), x[(long long)b * ldx + i], s);
    if (dW) dW[(long long)o * in + i] += s;
  } else {
    for (int b = 0; b < B; b++) s += masked_dy(dy, lddy, y, ldy, b, o, relu);
    if (db) db[o] += s;
  }
}

__global__ void relu_f32_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out,
                                long long n) {
  pdl_enter();
  // dy == nullptr: out = max(x, 0);  else: out = dy * (x > 0)   (x is the forward OUTPUT in the backward case)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = dy ? (x[i] > 0.f ? dy[i] : 0.f) : fmaxf(x[i], 0.f);
}

__global__ void rows_stats_kernel(const float* __restrict__ x, int ldx, int B, int C, double* __restrict__ stats) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0, q = 0;
  for (int b = 0; b < B; b++) {
    const double v = x[(long long)b * ldx + c];
    s += v;
    q += v * v;
  }
  atomicAdd(stats + c, s);
  atomicAdd(stats + C + c, q);
}

__global__ void bn1d_apply_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ scale,
                                  const float* __restrict__ shift, float* __restrict__ y, int ldy, int B, int C,
                                  int relu) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int b = idx / C, c = idx % C;
  float v = fmaf(x[(long long)b * ldx + c], scale[c], shift[c]);
  if (relu) v = fmaxf(v, 0.f);
  y[(long long)b * ldy + c] = v;
}

__global__ void bn1d_bwd_reduce_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
                                       const float* __restrict__ x, int ldx, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, int B, int C, int relu,
                                       double* __restrict__ red) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0, q = 0;
  for (int b = 0; b < B; b++) {
    const float g = masked_dy(dy, lddy, y, ldy, b, c, relu);
    const float xh = (x[(long long)b * ldx + c] - mean[c]) * invstd[c];
    s += g;
    q += (double)g * xh;
  }
  atomicAdd(red + c, s);
  atomicAdd(red + C + c, q);
}

__global__ void bn1d_bwd_apply_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
                                      const float* __restrict__ x, int ldx, const float* __restrict__ mean,
                                      const float* __restrict__ invstd, const float* __restrict__ gamma,
                                      const double* __restrict__ red, double inv_count, int B, int C, int relu,
                                      float* __restrict__ dx, int lddx, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int b = idx / C, c = idx % C;
  const float g = masked_dy(dy, lddy, y, ldy, b, c, relu);
  const float xh = (x[(long long)b * ldx + c] - mean[c]) * invstd[c];
  const float gm = gamma ? gamma[c] : 1.f;
  dx[(long long)b * lddx + c] =
      gm * invstd[c] * (g - (float)(red[c] * inv_count) - xh * (float)(red[C + c] * inv_count));
  if (b == 0) {
    if (dbeta) dbeta[c] = (float)red[c];
    if (dgamma) dgamma[c] = (float)red[C + c];
  }
}

// ---------------------------------------------------------------------------------------------
// Loss (fp64).  One thread per sample, block reduction, one fp64 atomic pair per block.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxClasses = 16;

__device__ __forceinline__ double pow_gamma(double base, double gamma) {
  if (gamma == 1.0) return base;
  if (gamma == 2.0) return base * base;
  return pow(base, gamma);
}

template <typename T>
__global__ void loss_fwd_kernel(const T* __restrict__ logits, int ld, const int64_t* __restrict__ target, int B,
                                int C, double gamma, const double* __restrict__ cw, double* __restrict__ partial,
                                double* __restrict__ coeff) {
  pdl_enter();
  __shared__ double sm[2][32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double num = 0, nrm = 0;
  if (i < B) {
    const T* z = logits + (long long)i * ld;
    double m = (double)z[0];
    for (int c = 1; c < C; c++) m = fmax(m, (double)z[c]);
    double se = 0;
    for (int c = 0; c < C; c++) se += exp((double)z[c] - m);
    const int t = (int)target[i];
    const double logpt = (double)z[t] - m - log(se);
    if (gamma > 0) {
      const double pt = exp(logpt);
      const double cf = pow_gamma(1.0 - pt, gamma);
      num = -1.0 * cf * logpt;
      nrm = 1.0;
      coeff[i] = cf;
    } else {
      const double w = cw ? cw[t] : 1.0;
      num = -w * logpt;
      nrm = w;
      coeff[i] = w;
    }
  }
  num = warp_sum(num);
  nrm = warp_sum(nrm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sm[0][warp] = num;
    sm[1][warp] = nrm;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < (blockDim.x + 31) / 32; w++) {
      a += sm[0][w];
      b += sm[1][w];
    }
    atomicAdd(partial + 0, a);
    atomicAdd(partial + 1, b);
  }
}

template <typename T>
__global__ void loss_bwd_kernel(const T* __restrict__ logits, int ld, const int64_t* __restrict__ target, int B,
                                int C, const double* __restrict__ coeff, const double* __restrict__ denom,
                                const double* __restrict__ upstream, T* __restrict__ dl, int lddl) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const T* z = logits + (long long)i * ld;
  double m = (double)z[0];
  for (int c = 1; c < C; c++) m = fmax(m, (double)z[c]);
  double se = 0;
  for (int c = 0; c < C; c++) se += exp((double)z[c] - m);
  const int t = (int)target[i];
  const double k = upstream[0] * coeff[i] / denom[0];
  for (int c = 0; c < C; c++) {
    const double p = exp((double)z[c] - m) / se;
    dl[(long long)i * lddl + c] = (T)(k * (p - (c == t ? 1.0 : 0.0)));
  }
}

}  // namespace
}  // namespace adni

using namespace adni;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

int adni_linear_fwd(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int B, int in, int out,
                    int relu, void* stream) {
  ADNI_REQUIRE(x && W && y && B > 0 && in > 0 && out > 0 && ldx >= in && ldy >= out, ADNI_EINVAL,
               "linear_fwd: bad arguments");
  const long long threads = (long long)B * out * 32;
  pdl_launch(linear_fwd_kernel, (unsigned)((threads + 255) / 256), 256, 0, ST(stream))(x, ldx, W, b, y, ldy, B, in, out, relu);
  count_launch();
  ADNI_LAUNCH_CHECK("linear_fwd_kernel");
  return ADNI_OK;
}

int adni_linear_bwd(const float* x, int ldx, const float* W, const float* y, int ldy, const float* dy, int lddy,
                    float* dx, int lddx, int accumulate_dx, float* dW, float* db, int B, int in, int out, int relu,
                    void* stream) {
  ADNI_REQUIRE(x && W && dy && B > 0 && in > 0 && out > 0, ADNI_EINVAL, "linear_bwd: bad arguments");
  ADNI_REQUIRE(!relu || y, ADNI_EINVAL, "linear_bwd: relu mask needs the forward output");
  if (dx) {
    pdl_launch(linear_bwd_dx_kernel, (B * in + 255) / 256, 256, 0, ST(stream))(W, y, ldy, dy, lddy, dx, lddx, accumulate_dx, B,
                                                                        in, out, relu);
    count_launch();
    ADNI_LAUNCH_CHECK("linear_bwd_dx_kernel");
  }
  if (dW || db) {
    const long long n = (long long)out * (in + 1);
    pdl_launch(linear_bwd_dw_kernel, (unsigned)((n + 255) / 256), 256, 0, ST(stream))(x, ldx, y, ldy, dy, lddy, dW, db, B, in,
                                                                               out, relu);
    count_launch();
    ADNI_LAUNCH_CHECK("linear_bwd_dw_kernel");
  }
  return ADNI_OK;
}

int adni_relu_f32(const float* x, const float* dy, float* out, long long n, void* stream) {
  ADNI_REQUIRE(x && out && n > 0, ADNI_EINVAL, "relu_f32: bad arguments");
  pdl_launch(relu_f32_kernel, (unsigned)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256), 256, 0, ST(stream))(x, dy, out, n);
  count_launch();
  ADNI_LAUNCH_CHECK("relu_f32_kernel");
  return ADNI_OK;
}

int adni_rows_stats_f32(const float* x, int ldx, int B, int C, double* stats, void* stream) {
  ADNI_REQUIRE(x && stats && B > 0 && C > 0, ADNI_EINVAL, "rows_stats_f32: bad arguments");
  pdl_launch(rows_stats_kernel, (C + 127) / 128, 128, 0, ST(stream))(x, ldx, B, C, stats);
  count_launch();
  ADNI_LAUNCH_CHECK("rows_stats_kernel");
  return ADNI_OK;
}

int adni_bn1d_apply(const float* x, int ldx, const float* scale, const float* shift, float* y, int ldy, int B, int C,
                    int relu, void* stream) {
  ADNI_REQUIRE(x && scale && shift && y && B > 0 && C > 0, ADNI_EINVAL, "bn1d_apply: bad arguments");
  pdl_launch(bn1d_apply_kernel, (B * C + 255) / 256, 256, 0, ST(stream))(x, ldx, scale, shift, y, ldy, B, C, relu);
  count_launch();
  ADNI_LAUNCH_CHECK("bn1d_apply_kernel");
  return ADNI_OK;
}

int adni_bn1d_bwd_reduce(const float* dy, int lddy, const float* y, int ldy, const float* x, int ldx, const float* mean,
                         const float* invstd, int B, int C, int relu, double* red, void* stream) {
  ADNI_REQUIRE(dy && x && mean && invstd && red && B > 0 && C > 0, ADNI_EINVAL, "bn1d_bwd_reduce: bad arguments");
  ADNI_REQUIRE(!relu || y, ADNI_EINVAL, "bn1d_bwd_reduce: relu mask needs the forward output");
  pdl_launch(bn1d_bwd_reduce_kernel, (C + 127) / 128, 128, 0, ST(stream))(dy, lddy, y, ldy, x, ldx, mean, invstd, B, C, relu,
                                                                   red);
  count_launch();
  ADNI_LAUNCH_CHECK("bn1d_bwd_reduce_kernel");
  return ADNI_OK;
}

int adni_bn1d_bwd_apply(const float* dy, int lddy, const float* y, int ldy, const float* x, int ldx, const float* mean,
                        const float* invstd, const float* gamma, const double* red, double count, int B, int C,
                        int relu, float* dx, int lddx, float* dgamma, float* dbeta, void* stream) {
  ADNI_REQUIRE(dy && x && mean && invstd && red && dx && B > 0 && C > 0 && count > 0, ADNI_EINVAL,
               "bn1d_bwd_apply: bad arguments");
  ADNI_REQUIRE(!relu || y, ADNI_EINVAL, "bn1d_bwd_apply: relu mask needs the forward output");
  pdl_launch(bn1d_bwd_apply_kernel, (B * C + 255) / 256, 256, 0, ST(stream))(dy, lddy, y, ldy, x, ldx, mean, invstd, gamma, red,
                                                                      1.0 / count, B, C, relu, dx, lddx, dgamma, dbeta);
  count_launch();
  ADNI_LAUNCH_CHECK("bn1d_bwd_apply_kernel");
  return ADNI_OK;
}

int adni_loss_fwd(const void* logits, int logits_f64, int ld, const int64_t* target, int B, int C, double gamma,
                  const double* class_weights, double* partial, double* per_sample_coeff, void* stream) {
  ADNI_REQUIRE(logits && target && partial && per_sample_coeff && B > 0, ADNI_EINVAL, "loss_fwd: bad arguments");
  ADNI_REQUIRE(C >= 2 && C <= kMaxClasses && ld >= C, ADNI_ENOTSUP, "loss_fwd: C=%d outside [2,%d]", C, kMaxClasses);
  ADNI_REQUIRE(gamma >= 0, ADNI_EINVAL, "loss_fwd: negative focal gamma");
  if (logits_f64)
    pdl_launch(loss_fwd_kernel<double>, (B + 127) / 128, 128, 0, ST(stream))(static_cast<const double*>(logits), ld, target, B,
                                                                    C, gamma, class_weights, partial,
                                                                    per_sample_coeff);
  else
    pdl_launch(loss_fwd_kernel<float>, (B + 127) / 128, 128, 0, ST(stream))(static_cast<const float*>(logits), ld, target, B,
                                                                   C, gamma, class_weights, partial, per_sample_coeff);
  count_launch();
  ADNI_LAUNCH_CHECK("loss_fwd_kernel");
  return ADNI_OK;
}

int adni_loss_bwd(const void* logits, int logits_f64, int ld, const int64_t* target, int B, int C,
                  const double* per_sample_coeff, const double* denom, const double* upstream, void* dlogits, int lddl,
                  void* stream) {
  ADNI_REQUIRE(logits && target && per_sample_coeff && denom && upstream && dlogits && B > 0, ADNI_EINVAL,
               "loss_bwd: bad arguments");
  ADNI_REQUIRE(C >= 2 && C <= kMaxClasses, ADNI_ENOTSUP, "loss_bwd: C=%d outside [2,%d]", C, kMaxClasses);
  if (logits_f64)
    pdl_launch(loss_bwd_kernel<double>, (B + 127) / 128, 128, 0, ST(stream))(static_cast<const double*>(logits), ld, target, B,
                                                                    C, per_sample_coeff, denom, upstream,
                                                                    static_cast<double*>(dlogits), lddl);
  else
    pdl_launch(loss_bwd_kernel<float>, (B + 127) / 128, 128, 0, ST(stream))(static_cast<const float*>(logits), ld, target, B, C,
                                                                   per_sample_coeff, denom, upstream,
                                                                   static_cast<float*>(dlogits), lddl);
  count_launch();
  ADNI_LAUNCH_CHECK("loss_bwd_kernel");
  return ADNI_OK;
}

}  // extern "C"

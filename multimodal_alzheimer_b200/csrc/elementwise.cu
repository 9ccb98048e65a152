// HBM-bound kernels around the convolutions: BatchNorm finalize / apply / backward, ReLU, residual add,
// per-channel statistics, max / global-average pooling, casts and weight-layout transposes.
// All activation tensors are [rows][C] views of NDHWC bf16 with C % 8 == 0: every thread moves 16-byte
// vectors (8 channels), grids are sized in multiples of the SM count.
#include <algorithm>

#include "common.cuh"

namespace adni {

extern void count_launch();

namespace {

constexpr int kEwThreads = 256;

inline int ew_grid(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

struct Vec8 {
  float v[8];
};
__device__ __forceinline__ Vec8 load8(const __nv_bfloat16* p) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  Vec8 o;
  o.v[0] = bf16_lo(r.x);
  o.v[1] = bf16_hi(r.x);
  o.v[2] = bf16_lo(r.y);
  o.v[3] = bf16_hi(r.y);
  o.v[4] = bf16_lo(r.z);
  o.v[5] = bf16_hi(r.z);
  o.v[6] = bf16_lo(r.w);
  o.v[7] = bf16_hi(r.w);
  return o;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const Vec8& a) {
  uint4 o;
  o.x = pack_bf16x2(a.v[0], a.v[1]);
  o.y = pack_bf16x2(a.v[2], a.v[3]);
  o.z = pack_bf16x2(a.v[4], a.v[5]);
  o.w = pack_bf16x2(a.v[6], a.v[7]);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ Vec8 loadf8(const float* p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  Vec8 o;
  o.v[0] = a.x;
  o.v[1] = a.y;
  o.v[2] = a.z;
  o.v[3] = a.w;
  o.v[4] = b.x;
  o.v[5] = b.y;
  o.v[6] = b.z;
  o.v[7] = b.w;
  return o;
}

// ---------------------------------------------------------------------------------------------
struct BnChannel {
  float mean, invstd, scale, shift;
  double var;
};
// One channel's BatchNorm parameters from the fp64 sums; bn_finalize_kernel and the fused bn_train_apply_kernel
// evaluate exactly this expression, so every block of the fused kernel derives bit-identical scale / shift.
__device__ __forceinline__ BnChannel bn_channel(const double* ssum, const double* ssq, double count, const float* gamma,
                                                const float* beta, float eps, int c) {
  BnChannel o;
  // mean and variance in fp64 (sum x^2 / n - mean^2 cancels), the reciprocal square root in fp32 like torch's own
  // fp32 BatchNorm: the fp64 divide + sqrt cost every thread of bn_train_apply ~5 us of prologue per launch
  const double inv = 1.0 / count;
  const double m = ssum[c] * inv;
  double var = ssq[c] * inv - m * m;
  if (var < 0) var = 0;
  o.var = var;
  o.invstd = 1.0f / sqrtf((float)var + eps);
  const float g = gamma ? gamma[c] : 1.f;
  const float b = beta ? beta[c] : 0.f;
  o.mean = (float)m;
  o.scale = g * o.invstd;
  o.shift = b - (float)m * o.scale;
  return o;
}

__global__ void bn_finalize_kernel(const double* ssum, const double* ssq, double count, int C, const float* gamma,
                                   const float* beta, float eps, float momentum, float* rmean, float* rvar,
                                   float* mean, float* invstd, float* scale, float* shift) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const BnChannel ch = bn_channel(ssum, ssq, count, gamma, beta, eps, c);
  if (mean) mean[c] = ch.mean;
  if (invstd) invstd[c] = ch.invstd;
  if (scale) scale[c] = ch.scale;
  if (shift) shift[c] = ch.shift;
  if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * ch.mean;
  if (rvar) {
    const double unb = count > 1 ? ch.var * count / (count - 1) : ch.var;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
  }
}

__device__ __forceinline__ Vec8 unpack_vec8(const uint4& r) {
  Vec8 o;
  o.v[0] = bf16_lo(r.x);
  o.v[1] = bf16_hi(r.x);
  o.v[2] = bf16_lo(r.y);
  o.v[3] = bf16_hi(r.y);
  o.v[4] = bf16_lo(r.z);
  o.v[5] = bf16_hi(r.z);
  o.v[6] = bf16_lo(r.w);
  o.v[7] = bf16_hi(r.w);
  return o;
}
constexpr int kEwUnroll = 4;  // independent 16-byte loads in flight per thread and tensor

__global__ void __launch_bounds__(kEwThreads) bn_apply_kernel(const __nv_bfloat16* __restrict__ y,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              const __nv_bfloat16* __restrict__ res,
                                                              __nv_bfloat16* __restrict__ out, long long nvec, int vpr,
                                                              int relu) {
  pdl_enter();
  const long long stride = (long long)gridDim.x * blockDim.x;
  // blockDim % vpr == 0  =>  i % vpr is the same for every vector a thread touches: its 8 scale / shift values
  // stay in registers instead of being re-fetched (4 x 16-byte L1 loads per 16-byte data vector otherwise)
  const bool fixed_c = (kEwThreads % vpr) == 0;
  Vec8 fsc, fsh;
  if (fixed_c) {
    const int c0 = (int)(threadIdx.x % vpr) * 8;
    fsc = loadf8(scale + c0);
    fsh = loadf8(shift + c0);
  }
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * kEwUnroll) {
    uint4 ry[kEwUnroll], rr[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; u++) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        ry[u] = __ldg(reinterpret_cast<const uint4*>(y + i * 8));
        if (res) rr[u] = __ldg(reinterpret_cast<const uint4*>(res + i * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; u++) {
      const long long i = i0 + u * stride;
      if (i >= nvec) break;
      Vec8 a = unpack_vec8(ry[u]);
      Vec8 sc = fsc, sh = fsh;
      if (!fixed_c) {
        const int c0 = (int)(i % vpr) * 8;
        sc = loadf8(scale + c0);
        sh = loadf8(shift + c0);
      }
#pragma unroll
      for (int j = 0; j < 8; j++) a.v[j] = fmaf(a.v[j], sc.v[j], sh.v[j]);
      if (res) {
        const Vec8 r = unpack_vec8(rr[u]);
#pragma unroll
        for (int j = 0; j < 8; j++) a.v[j] += r.v[j];
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; j++) a.v[j] = fmaxf(a.v[j], 0.f);
      }
      store8(out + i * 8, a);
    }
  }
}

// Training-mode BatchNorm forward in one launch: every thread derives scale / shift of ITS 8 channels from the fp64
// batch sums (blockDim % vpr == 0, so they never change), block 0 additionally publishes mean / invstd / scale /
// shift for the backward pass and updates the running statistics.  Saves the separate bn_finalize launch and its
// dependency bubble per BatchNorm layer (21 per encoder and step).
__global__ void __launch_bounds__(kEwThreads)
    bn_train_apply_kernel(const __nv_bfloat16* __restrict__ y, const double* __restrict__ ssum,
                          const double* __restrict__ ssq, double count, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float eps, float momentum, float* __restrict__ rmean,
                          float* __restrict__ rvar, float* __restrict__ bnp, const __nv_bfloat16* __restrict__ res,
                          __nv_bfloat16* __restrict__ out, long long nvec, int vpr, int C, int relu) {
  pdl_enter();
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const BnChannel ch = bn_channel(ssum, ssq, count, gamma, beta, eps, c);
      bnp[c] = ch.mean;
      bnp[C + c] = ch.invstd;
      bnp[2 * C + c] = ch.scale;
      bnp[3 * C + c] = ch.shift;
      if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * ch.mean;
      if (rvar) {
        const double unb = count > 1 ? ch.var * count / (count - 1) : ch.var;
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
      }
    }
  }
  Vec8 sc, sh;
  {
    const int c0 = (int)(threadIdx.x % vpr) * 8;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const BnChannel ch = bn_channel(ssum, ssq, count, gamma, beta, eps, c0 + j);
      sc.v[j] = ch.scale;
      sh.v[j] = ch.shift;
    }
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * kEwUnroll) {
    uint4 ry[kEwUnroll], rr[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; u++) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        ry[u] = __ldg(reinterpret_cast<const uint4*>(y + i * 8));
        if (res) rr[u] = __ldg(reinterpret_cast<const uint4*>(res + i * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; u++) {
      const long long i = i0 + u * stride;
      if (i >= nvec) break;
      Vec8 a = unpack_vec8(ry[u]);
#pragma unroll
      for (int j = 0; j < 8; j++) a.v[j] = fmaf(a.v[j], sc.v[j], sh.v[j]);
      if (res) {
        const Vec8 r = unpack_vec8(rr[u]);
#pragma unroll
        for (int j = 0; j < 8; j++) a.v[j] += r.v[j];
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; j++) a.v[j] = fmaxf(a.v[j], 0.f);
      }
      store8(out + i * 8, a);
    }
  }
}

// Per-channel reductions over rows.  Block = 256 threads = (vpr channel vectors) x (rpp row lanes); a
// block walks a slab of rows, reduces its row lanes through shared memory and issues one fp64 atomic per
// channel.  MODE 0: sum x, sum x^2.  MODE 1: BatchNorm backward sums  sum g, sum g*xhat.
// rows per block: sized on the host so that the grid is >= 4 blocks per SM (small-spatial layers have few rows).
inline int red_slab_rows(long long rows, int C) {
  const int rpp = kEwThreads / (C / 8);
  long long slab = (rows + (long long)num_sms() * 4 - 1) / ((long long)num_sms() * 4);
  const long long min_slab = (long long)rpp * 8;  // >= 8 rows per thread keeps the atomic tail small
  if (slab < min_slab) slab = min_slab;
  if (slab > 4096) slab = 4096;
  return (int)slab;
}

template <int MODE>
__global__ void __launch_bounds__(kEwThreads) channel_reduce_kernel(const __nv_bfloat16* __restrict__ a,    // x | dout
                                                                    const __nv_bfloat16* __restrict__ outp,  // - | out (relu mask)
                                                                    const __nv_bfloat16* __restrict__ yraw,  // - | y
                                                                    const float* __restrict__ mean,
                                                                    const float* __restrict__ invstd,
                                                                    const float* __restrict__ scale,
                                                                    const float* __restrict__ shift, long long rows,
                                                                    int C, int relu, int slab_rows,
                                                                    double* __restrict__ r0,
                                                                    double* __restrict__ r1) {
  pdl_enter();
  __shared__ float sm[2][kEwThreads][8];
  const int vpr = C / 8;
  const int rpp = kEwThreads / vpr;  // row lanes per block (vpr <= 256)
  const int cv = threadIdx.x % vpr;
  const int rl = threadIdx.x / vpr;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; j++) s0[j] = s1[j] = 0.f;
  const long long row_begin = (long long)blockIdx.x * slab_rows;
  const long long row_end = min(rows, row_begin + slab_rows);
  if (rl < rpp) {
    Vec8 mu, is, sc, sh;
    if (MODE == 1) {
      mu = loadf8(mean + cv * 8);
      is = loadf8(invstd + cv * 8);
      if (relu == 2) {  // ReLU mask recomputed from y (no residual): out > 0  <=>  y*scale + shift > 0
        sc = loadf8(scale + cv * 8);
        sh = loadf8(shift + cv * 8);
      }
    }
    for (long long r0 = row_begin + rl; r0 < row_end; r0 += (long long)rpp * kEwUnroll) {
      uint4 ra[kEwUnroll], ry[kEwUnroll], ro[kEwUnroll];
#pragma unroll
      for (int u = 0; u < kEwUnroll; u++) {  // all loads of the batch are issued before the first use
        const long long r = r0 + (long long)u * rpp;
        if (r < row_end) {
          const long long off = r * C + cv * 8;
          ra[u] = __ldg(reinterpret_cast<const uint4*>(a + off));
          if (MODE == 1) {
            ry[u] = __ldg(reinterpret_cast<const uint4*>(yraw + off));
            if (relu == 1) ro[u] = __ldg(reinterpret_cast<const uint4*>(outp + off));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kEwUnroll; u++) {
        const long long r = r0 + (long long)u * rpp;
        if (r >= row_end) break;
        const Vec8 x = unpack_vec8(ra[u]);
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            s0[j] += x.v[j];
            s1[j] = fmaf(x.v[j], x.v[j], s1[j]);
          }
        } else {
          Vec8 g = x;
          const Vec8 yv = unpack_vec8(ry[u]);
          if (relu == 1) {
            const Vec8 o = unpack_vec8(ro[u]);
#pragma unroll
            for (int j = 0; j < 8; j++) g.v[j] = o.v[j] > 0.f ? g.v[j] : 0.f;
          } else if (relu == 2) {
#pragma unroll
            for (int j = 0; j < 8; j++) g.v[j] = fmaf(yv.v[j], sc.v[j], sh.v[j]) > 0.f ? g.v[j] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float xh = (yv.v[j] - mu.v[j]) * is.v[j];
            s0[j] += g.v[j];
            s1[j] = fmaf(g.v[j], xh, s1[j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sm[0][threadIdx.x][j] = s0[j];
    sm[1][threadIdx.x][j] = s1[j];
  }
  __syncthreads();
  // thread t < vpr*8 reduces channel (t/8 vector, t%8 lane) across the row lanes
  for (int t = threadIdx.x; t < vpr * 8; t += kEwThreads) {
    const int v = t / 8, j = t % 8;
    float a0 = 0.f, a1 = 0.f;
    for (int l = 0; l < rpp; l++) {
      a0 += sm[0][l * vpr + v][j];
      a1 += sm[1][l * vpr + v][j];
    }
    atomicAdd(r0 + v * 8 + j, (double)a0);
    atomicAdd(r1 + v * 8 + j, (double)a1);
  }
}

// dy = gamma*invstd*(g - mean_g - xhat*mean_gx) rewritten as  dy = A[c]*g + B[c]*y + K[c]:  the per-channel
// coefficients are computed once per block into shared memory (fp64 sums -> fp32), so the streaming loop issues only
// the tensor loads plus 6 shared-memory vector loads instead of ~28 scalar parameter loads per 8 elements.
__global__ void __launch_bounds__(kEwThreads)
    bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ outp,
                        const __nv_bfloat16* __restrict__ yraw, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const double* __restrict__ red, double inv_count, long long nvec, int vpr, int C, int relu,
                        __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dres, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, double pg_scale, int red_form) {
  pdl_enter();
  // red_form 0: red[C + c] = sum g*xhat;  1: red[C + c] = sum g*y (a conv epilogue produced it, adni_conv3d_dgrad_bnred):
  // sum g*xhat = invstd * (sum g*y - mean * sum g), evaluated in fp64
  auto sum_gx = [&](int c) -> double {
    return red_form ? (double)invstd[c] * (red[C + c] - (double)mean[c] * red[c]) : red[C + c];
  };
  if (blockIdx.x == 0 && (dgamma || dbeta)) {  // dgamma = sum g*xhat, dbeta = sum g (x pg_scale, see the C-ABI comment)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = (float)(red[c] * pg_scale);
      if (dgamma) dgamma[c] = (float)(sum_gx(c) * pg_scale);
    }
  }
  extern __shared__ float coef[];  // [5][C]: A, B, K, scale, shift (only used when the channel vector is not fixed)
  // dy = A*g + B*y + K with A = gamma*invstd, B = -A*invstd*mean(g*xhat), K = -A*mean(g) - B*mu
  auto coefs = [&](int c, float& a, float& b, float& k) {
    const float gm = gamma ? gamma[c] : 1.f;
    const float is = invstd[c], mu = mean[c];
    const float mg = (float)(red[c] * inv_count);
    const float mgx = (float)(sum_gx(c) * inv_count);
    a = gm * is;
    b = -a * is * mgx;
    k = -a * mg - b * mu;
  };
  // blockDim % vpr == 0  =>  a thread always works on the same 8 channels: their coefficients live in registers.
  // (The shared-memory table costs 40 scalar LDS per 16-byte vector with 8-way bank conflicts at C = 512: the kernel
  // ran at 45 % of HBM bandwidth because of it.)
  const bool fixed_c = (kEwThreads % vpr) == 0;
  float fa[8], fb[8], fk[8], fsc[8], fsh[8];
  if (fixed_c) {
    const int c0 = (int)(threadIdx.x % vpr) * 8;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      coefs(c0 + j, fa[j], fb[j], fk[j]);
      fsc[j] = relu == 2 ? scale[c0 + j] : 0.f;
      fsh[j] = relu == 2 ? shift[c0 + j] : 0.f;
    }
  } else {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a, b, k;
      coefs(c, a, b, k);
      coef[c] = a;
      coef[C + c] = b;
      coef[2 * C + c] = k;
      coef[3 * C + c] = relu == 2 ? scale[c] : 0.f;
      coef[4 * C + c] = relu == 2 ? shift[c] : 0.f;
    }
    __syncthreads();
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * kEwUnroll) {
    uint4 rg[kEwUnroll], ry[kEwUnroll], ro[kEwUnroll];
#pragma unroll
    for (int u = 0; u < kEwUnroll; u++) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        rg[u] = __ldg(reinterpret_cast<const uint4*>(dout + i * 8));
        ry[u] = __ldg(reinterpret_cast<const uint4*>(yraw + i * 8));
        if (relu == 1) ro[u] = __ldg(reinterpret_cast<const uint4*>(outp + i * 8));
      }
    }
#pragma unroll
    for (int u = 0; u < kEwUnroll; u++) {
      const long long i = i0 + u * stride;
      if (i >= nvec) break;
      if (!fixed_c) {
        const int c0 = (int)(i % vpr) * 8;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          fa[j] = coef[c0 + j];
          fb[j] = coef[C + c0 + j];
          fk[j] = coef[2 * C + c0 + j];
          fsc[j] = coef[3 * C + c0 + j];
          fsh[j] = coef[4 * C + c0 + j];
        }
      }
      Vec8 g = unpack_vec8(rg[u]);
      const Vec8 yv = unpack_vec8(ry[u]);
      if (relu == 1) {
        const Vec8 o = unpack_vec8(ro[u]);
#pragma unroll
        for (int j = 0; j < 8; j++) g.v[j] = o.v[j] > 0.f ? g.v[j] : 0.f;
      } else if (relu == 2) {
#pragma unroll
        for (int j = 0; j < 8; j++) g.v[j] = fmaf(yv.v[j], fsc[j], fsh[j]) > 0.f ? g.v[j] : 0.f;
      }
      if (dres) store8(dres + i * 8, g);
      Vec8 r;
#pragma unroll
      for (int j = 0; j < 8; j++) r.v[j] = fmaf(fa[j], g.v[j], fmaf(fb[j], yv.v[j], fk[j]));
      store8(dy + i * 8, r);
    }
  }
}

__global__ void bn_eval_params_kernel(const float* rm, const float* rv, const float* gamma, const float* beta, float eps,
                                      int C, float* scale, float* shift) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = (gamma ? gamma[c] : 1.f) / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = (beta ? beta[c] : 0.f) - rm[c] * sc;
}

__global__ void __launch_bounds__(kEwThreads) relu_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                              __nv_bfloat16* __restrict__ y, long long nvec) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    Vec8 a = load8(x + i * 8);
#pragma unroll
    for (int j = 0; j < 8; j++) a.v[j] = fmaxf(a.v[j], 0.f);
    store8(y + i * 8, a);
  }
}
__global__ void __launch_bounds__(kEwThreads) relu_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                              const __nv_bfloat16* __restrict__ y,
                                                              __nv_bfloat16* __restrict__ dx, long long nvec) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    Vec8 g = load8(dy + i * 8);
    const Vec8 o = load8(y + i * 8);
#pragma unroll
    for (int j = 0; j < 8; j++) g.v[j] = o.v[j] > 0.f ? g.v[j] : 0.f;
    store8(dx + i * 8, g);
  }
}

__global__ void bn_param_grads_kernel(const double* red, int C, float* dgamma, float* dbeta) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] = (float)red[c];
  if (dgamma) dgamma[c] = (float)red[C + c];
}

// ---------------------------------------------------------------------------------------------
// Pooling
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
    maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, int N, int D, int H, int W, int C, int k, int s, int pad,
                       int Do, int Ho, int Wo, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ amax) {
  pdl_enter();
  const int vpr = C / 8;
  const long long total = (long long)N * Do * Ho * Wo * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vpr);
    long long r = i / vpr;
    const int ow = (int)(r % Wo);
    r /= Wo;
    const int oh = (int)(r % Ho);
    r /= Ho;
    const int od = (int)(r % Do);
    const int n = (int)(r / Do);
    float best[8];
    int bidx[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      best[j] = -INFINITY;
      bidx[j] = -1;
    }
    for (int kd = 0; kd < k; kd++) {
      const int id = od * s - pad + kd;
      if (id < 0 || id >= D) continue;
      for (int kh = 0; kh < k; kh++) {
        const int ih = oh * s - pad + kh;
        if (ih < 0 || ih >= H) continue;
        for (int kw = 0; kw < k; kw++) {
          const int iw = ow * s - pad + kw;
          if (iw < 0 || iw >= W) continue;
          const Vec8 v = load8(x + ((((long long)n * D + id) * H + ih) * W + iw) * C + cv * 8);
          const int slot = (kd * k + kh) * k + kw;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (bidx[j] < 0 || v.v[j] > best[j]) {  // strict '>' : first maximum in scan order wins
              best[j] = v.v[j];
              bidx[j] = slot;
            }
          }
        }
      }
    }
    Vec8 o;
#pragma unroll
    for (int j = 0; j < 8; j++) o.v[j] = best[j];
    store8(y + i * 8, o);
    uint2 pk;
    pk.x = (uint32_t)(bidx[0] & 255) | ((uint32_t)(bidx[1] & 255) << 8) | ((uint32_t)(bidx[2] & 255) << 16) |
           ((uint32_t)(bidx[3] & 255) << 24);
    pk.y = (uint32_t)(bidx[4] & 255) | ((uint32_t)(bidx[5] & 255) << 8) | ((uint32_t)(bidx[6] & 255) << 16) |
           ((uint32_t)(bidx[7] & 255) << 24);
    *reinterpret_cast<uint2*>(amax + i * 8) = pk;
  }
}

__global__ void __launch_bounds__(kEwThreads)
    maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ amax, int N, int D, int H,
                       int W, int C, int k, int s, int pad, int Do, int Ho, int Wo, __nv_bfloat16* __restrict__ dx) {
  pdl_enter();
  const int vpr = C / 8;
  const long long total = (long long)N * D * H * W * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vpr);
    long long r = i / vpr;
    const int iw = (int)(r % W);
    r /= W;
    const int ih = (int)(r % H);
    r /= H;
    const int id = (int)(r % D);
    const int n = (int)(r / D);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    // windows containing this input voxel: o*s - pad + kk == i  with 0 <= kk < k
    const int od_lo = max(0, (id + pad - k + s) / s), od_hi = min(Do - 1, (id + pad) / s);
    const int oh_lo = max(0, (ih + pad - k + s) / s), oh_hi = min(Ho - 1, (ih + pad) / s);
    const int ow_lo = max(0, (iw + pad - k + s) / s), ow_hi = min(Wo - 1, (iw + pad) / s);
    for (int od = od_lo; od <= od_hi; od++) {
      const int kd = id + pad - od * s;
      if (kd < 0 || kd >= k) continue;
      for (int oh = oh_lo; oh <= oh_hi; oh++) {
        const int kh = ih + pad - oh * s;
        if (kh < 0 || kh >= k) continue;
        for (int ow = ow_lo; ow <= ow_hi; ow++) {
          const int kw = iw + pad - ow * s;
          if (kw < 0 || kw >= k) continue;
          const int slot = (kd * k + kh) * k + kw;
          const long long o = ((((long long)n * Do + od) * Ho + oh) * Wo + ow) * vpr + cv;
          const uint2 pk = __ldg(reinterpret_cast<const uint2*>(amax + o * 8));
          const Vec8 g = load8(dy + o * 8);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const uint32_t word = j < 4 ? pk.x : pk.y;
            const int a = (word >> ((j & 3) * 8)) & 255;
            if (a == slot) acc[j] += g.v[j];
          }
        }
      }
    }
    Vec8 o;
#pragma unroll
    for (int j = 0; j < 8; j++) o.v[j] = acc[j];
    store8(dx + i * 8, o);
  }
}

// Tiled max-pool backward: a block owns an 8x8x8 input tile (x 64 channels) and first stages the <= 6^3 pooled
// windows that can point into it (dy + arg-max) in shared memory, so global traffic is ~1x instead of the up to
// 8 window probes per voxel of the straightforward gather above.
constexpr int kMpT = 8;
constexpr int kMpWMax = 6;
__global__ void __launch_bounds__(kEwThreads)
    maxpool_bwd_tiled_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ amax, int N, int D, int H,
                             int W, int C, int k, int s, int pad, int Do, int Ho, int Wo, int tiles_d, int tiles_h,
                             int tiles_w, __nv_bfloat16* __restrict__ dx) {
  pdl_enter();
  __shared__ uint4 s_dy[kMpWMax * kMpWMax * kMpWMax * 8];   // [window][cv] 8 channels bf16
  __shared__ uint2 s_am[kMpWMax * kMpWMax * kMpWMax * 8];   // [window][cv] 8 arg-max bytes
  const int vpr = C / 8;
  const int cv0 = blockIdx.y * 8;  // 64-channel chunk
  int t = blockIdx.x;
  const int tw = t % tiles_w;
  t /= tiles_w;
  const int th = t % tiles_h;
  t /= tiles_h;
  const int td = t % tiles_d;
  const int n = t / tiles_d;
  const int i0d = td * kMpT, i0h = th * kMpT, i0w = tw * kMpT;
  auto lo = [&](int i0) {
    const int num = i0 + pad - k + 1;
    return num <= 0 ? 0 : (num + s - 1) / s;
  };
  const int od0 = lo(i0d), oh0 = lo(i0h), ow0 = lo(i0w);
  const int nd = min(Do - 1, (i0d + kMpT - 1 + pad) / s) - od0 + 1;
  const int nh = min(Ho - 1, (i0h + kMpT - 1 + pad) / s) - oh0 + 1;
  const int nw = min(Wo - 1, (i0w + kMpT - 1 + pad) / s) - ow0 + 1;
  const int nwin = nd * nh * nw;
  for (int i = threadIdx.x; i < nwin * 8; i += kEwThreads) {
    const int cv = i & 7, wdx = i >> 3;
    const int ww = wdx % nw, wh = (wdx / nw) % nh, wd = wdx / (nw * nh);
    if (cv0 + cv < vpr) {
      const long long o = ((((long long)n * Do + od0 + wd) * Ho + oh0 + wh) * Wo + ow0 + ww) * vpr + cv0 + cv;
      s_dy[i] = __ldg(reinterpret_cast<const uint4*>(dy + o * 8));
      s_am[i] = __ldg(reinterpret_cast<const uint2*>(amax + o * 8));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMpT * kMpT * kMpT * 8; i += kEwThreads) {
    const int cv = i & 7, v = i >> 3;
    const int iw = i0w + (v % kMpT), ih = i0h + ((v / kMpT) % kMpT), id = i0d + v / (kMpT * kMpT);
    if (iw >= W || ih >= H || id >= D || cv0 + cv >= vpr) continue;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    const int a_lo = max(od0, lo(id)), a_hi = min(od0 + nd - 1, (id + pad) / s);
    const int b_lo = max(oh0, lo(ih)), b_hi = min(oh0 + nh - 1, (ih + pad) / s);
    const int c_lo = max(ow0, lo(iw)), c_hi = min(ow0 + nw - 1, (iw + pad) / s);
    for (int od = a_lo; od <= a_hi; od++) {
      const int kd = id + pad - od * s;
      for (int oh = b_lo; oh <= b_hi; oh++) {
        const int kh = ih + pad - oh * s;
        for (int ow = c_lo; ow <= c_hi; ow++) {
          const int kw = iw + pad - ow * s;
          const int slot = (kd * k + kh) * k + kw;
          const int wdx = ((od - od0) * nh + (oh - oh0)) * nw + (ow - ow0);
          const uint2 pk = s_am[wdx * 8 + cv];
          const uint4 g = s_dy[wdx * 8 + cv];
          const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const uint32_t word = j < 4 ? pk.x : pk.y;
            const int a = (word >> ((j & 3) * 8)) & 255;
            const float gv = (j & 1) ? bf16_hi(gw[j >> 1]) : bf16_lo(gw[j >> 1]);
            if (a == slot) acc[j] += gv;
          }
        }
      }
    }
    Vec8 o;
#pragma unroll
    for (int j = 0; j < 8; j++) o.v[j] = acc[j];
    store8(dx + (((((long long)n * D + id) * H + ih) * W + iw) * vpr + cv0 + cv) * 8, o);
  }
}

constexpr int kGapSplit = 8;
__global__ void __launch_bounds__(kEwThreads)
    gap_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long P, int C, float inv_p, float* __restrict__ feat) {
  pdl_enter();
  // grid: (channel-vector blocks of 32, P splits, N); block = 32 channel vectors x 8 row lanes
  __shared__ float sm[8][32][8];
  const int vpr = C / 8;
  const int cv = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const int n = blockIdx.z;
  const long long per = (P + gridDim.y - 1) / gridDim.y;
  const long long p0 = (long long)blockIdx.y * per, p1 = min(P, p0 + per);
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; j++) s[j] = 0.f;
  if (cv < vpr) {
    for (long long pp = p0 + rl; pp < p1; pp += 8) {
      const Vec8 v = load8(x + ((long long)n * P + pp) * C + cv * 8);
#pragma unroll
      for (int j = 0; j < 8; j++) s[j] += v.v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) sm[rl][threadIdx.x & 31][j] = s[j];
  __syncthreads();
  const int t = threadIdx.x;  // 256 threads -> 32 vectors x 8 lanes
  const int v = t >> 3, j = t & 7;
  const int cvv = blockIdx.x * 32 + v;
  if (cvv < vpr) {
    float a = 0.f;
#pragma unroll
    for (int l = 0; l < 8; l++) a += sm[l][v][j];
    atomicAdd(feat + (long long)n * C + cvv * 8 + j, a * inv_p);
  }
}

__global__ void __launch_bounds__(kEwThreads)
    gap_bwd_kernel(const float* __restrict__ dfeat, int N, long long P, int C, float inv_p,
                   __nv_bfloat16* __restrict__ dx) {
  pdl_enter();
  const int vpr = C / 8;
  const long long total = (long long)N * P * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vpr);
    const int n = (int)(i / ((long long)P * vpr));
    Vec8 g = loadf8(dfeat + (long long)n * C + cv * 8);
#pragma unroll
    for (int j = 0; j < 8; j++) g.v[j] *= inv_p;
    store8(dx + i * 8, g);
  }
}

// ---------------------------------------------------------------------------------------------
// Casts and batched transposes (weight layouts)
// ---------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void cast_to_bf16_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn((float)x[i]);
}
__global__ void cast_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long n) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}

__device__ __forceinline__ void store_t(__nv_bfloat16* p, float v, bool) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store_t(float* p, float v, bool acc) { *p = acc ? *p + v : v; }

// in [batch][R][Cc] (fp32) -> out [batch][Cc][R]
template <typename TOut>
__global__ void transpose_kernel(const float* __restrict__ in, TOut* __restrict__ out, int R, int Cc, bool accumulate) {
  pdl_enter();
  __shared__ float tile[32][33];
  const long long boff = (long long)blockIdx.z * R * Cc;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < Cc) tile[j][threadIdx.x] = in[boff + (long long)r * Cc + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < Cc) store_t(out + boff + (long long)c * R + r, tile[threadIdx.x][j], accumulate);
  }
}

template <typename TOut>
int launch_transpose(const float* in, TOut* out, int batch, int R, int Cc, bool accumulate, cudaStream_t stream) {
  dim3 block(32, 8);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = std::min(batch - b0, 65535);
    dim3 grid((Cc + 31) / 32, (R + 31) / 32, nb);
    pdl_launch(transpose_kernel<TOut>, grid, block, 0, stream)(in + (long long)b0 * R * Cc, out + (long long)b0 * R * Cc, R, Cc,
                                                       accumulate);
    count_launch();
  }
  ADNI_LAUNCH_CHECK("transpose_kernel");
  return ADNI_OK;
}


// ------------------------------------------------------------------------------------------------
// Multi-tensor weight conversion: every conv weight of an encoder in ONE launch.
// fp32 parameter [Cout][Cin][taps]  ->  bf16 OTI [Cout][taps][Cin] (fprop B operand) and ITO [Cin][taps][Cout] (dgrad).
// A ResNet-18 encoder has 19 tensor-core convs: 38 + 38 transpose launches per step otherwise, a fixed cost that
// does not shrink with the per-GPU batch (11 % of the 8-GPU step).
struct WeightJob {
  const float* src;
  __nv_bfloat16* oti;
  __nv_bfloat16* ito;
  int cout, cin, taps;
  int tile_begin;  // first block of this tensor
  int tiles_ci;    // 16-channel tiles along Cin
};
constexpr int kWTile = 16;
constexpr int kWMaxTaps = 27;

__global__ void __launch_bounds__(256) weights_multi_kernel(const WeightJob* __restrict__ jobs, int n_jobs) {
  pdl_enter();
  __shared__ float tile[kWTile][kWTile + 1][kWMaxTaps];  // +1: the ITO pass reads with the co index fastest
  int j = 0;
  while (j + 1 < n_jobs && (int)blockIdx.x >= jobs[j + 1].tile_begin) j++;
  const WeightJob job = jobs[j];
  const int t = blockIdx.x - job.tile_begin;
  const int co0 = (t / job.tiles_ci) * kWTile, ci0 = (t % job.tiles_ci) * kWTile;
  const int taps = job.taps;
  // thread (a, b) of the 16 x 16 tile walks the taps: no divisions anywhere (the index arithmetic of a generic
  // element-per-thread mapping cost more than the memory traffic)
  const int a = threadIdx.x >> 4, b = threadIdx.x & 15;
  {
    const int co = a, ci = b;  // a thread reads the `taps` contiguous floats of one (co, ci)
    if (co0 + co < job.cout && ci0 + ci < job.cin) {
      const float* src = job.src + ((long long)(co0 + co) * job.cin + ci0 + ci) * taps;
      for (int tp = 0; tp < taps; tp++) tile[co][ci][tp] = __ldg(src + tp);
    }
  }
  __syncthreads();
  if (job.oti) {
    const int co = a, ci = b;  // 16 lanes write 16 consecutive Cin (32 bytes)
    if (co0 + co < job.cout && ci0 + ci < job.cin) {
      __nv_bfloat16* dst = job.oti + (long long)(co0 + co) * taps * job.cin + ci0 + ci;
      for (int tp = 0; tp < taps; tp++) dst[(long long)tp * job.cin] = __float2bfloat16_rn(tile[co][ci][tp]);
    }
  }
  if (job.ito) {
    const int ci = a, co = b;  // 16 lanes write 16 consecutive Cout
    if (co0 + co < job.cout && ci0 + ci < job.cin) {
      __nv_bfloat16* dst = job.ito + (long long)(ci0 + ci) * taps * job.cout + co0 + co;
      for (int tp = 0; tp < taps; tp++) dst[(long long)tp * job.cout] = __float2bfloat16_rn(tile[co][ci][tp]);
    }
  }
}

}  // namespace

// pool_small.cu
int launch_pool_nonoverlap_bwd(const __nv_bfloat16* dy, const uint8_t* am, const __nv_bfloat16* pooled, int N, int D, int H,
                               int W, int C, int k, __nv_bfloat16* dx, cudaStream_t st);
}  // namespace adni

using namespace adni;
typedef __nv_bfloat16 bf16;
#define BF(p) reinterpret_cast<bf16*>(p)
#define CBF(p) reinterpret_cast<const bf16*>(p)
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

int adni_bn_finalize(const double* stat_sum, const double* stat_sqsum, double count, int C, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                     float* mean, float* invstd, float* scale, float* shift, void* stream) {
  ADNI_REQUIRE(stat_sum && stat_sqsum && C > 0 && count > 0, ADNI_EINVAL, "bn_finalize: bad arguments");
  pdl_launch(bn_finalize_kernel, (C + 127) / 128, 128, 0, ST(stream))(stat_sum, stat_sqsum, count, C, gamma, beta, eps,
                                                               momentum, running_mean, running_var, mean, invstd,
                                                               scale, shift);
  count_launch();
  ADNI_LAUNCH_CHECK("bn_finalize_kernel");
  return ADNI_OK;
}

int adni_bn_param_grads(const double* red, int C, float* dgamma, float* dbeta, void* stream) {
  ADNI_REQUIRE(red && C > 0 && (dgamma || dbeta), ADNI_EINVAL, "bn_param_grads: bad arguments");
  pdl_launch(bn_param_grads_kernel, (C + 127) / 128, 128, 0, ST(stream))(red, C, dgamma, dbeta);
  count_launch();
  ADNI_LAUNCH_CHECK("bn_param_grads_kernel");
  return ADNI_OK;
}

int adni_bn_eval_params(const float* running_mean, const float* running_var, const float* gamma, const float* beta,
                        float eps, int C, float* scale, float* shift, void* stream) {
  ADNI_REQUIRE(running_mean && running_var && scale && shift && C > 0, ADNI_EINVAL, "bn_eval_params: bad arguments");
  pdl_launch(bn_eval_params_kernel, (C + 127) / 128, 128, 0, ST(stream))(running_mean, running_var, gamma, beta, eps, C, scale,
                                                                  shift);
  count_launch();
  ADNI_LAUNCH_CHECK("bn_eval_params_kernel");
  return ADNI_OK;
}

int adni_relu_fwd(const adni_bf16* x, adni_bf16* y, long long n, void* stream) {
  ADNI_REQUIRE(x && y && n > 0 && n % 8 == 0, ADNI_EINVAL, "relu_fwd: bad arguments");
  pdl_launch(relu_fwd_kernel, ew_grid(n / 8, kEwThreads * 4), kEwThreads, 0, ST(stream))(CBF(x), BF(y), n / 8);
  count_launch();
  ADNI_LAUNCH_CHECK("relu_fwd_kernel");
  return ADNI_OK;
}
int adni_relu_bwd(const adni_bf16* dy, const adni_bf16* y, adni_bf16* dx, long long n, void* stream) {
  ADNI_REQUIRE(dy && y && dx && n > 0 && n % 8 == 0, ADNI_EINVAL, "relu_bwd: bad arguments");
  pdl_launch(relu_bwd_kernel, ew_grid(n / 8, kEwThreads * 4), kEwThreads, 0, ST(stream))(CBF(dy), CBF(y), BF(dx), n / 8);
  count_launch();
  ADNI_LAUNCH_CHECK("relu_bwd_kernel");
  return ADNI_OK;
}

int adni_channel_stats(const adni_bf16* x, long long rows, int C, double* sum, double* sqsum, void* stream) {
  ADNI_REQUIRE(x && sum && sqsum && rows > 0, ADNI_EINVAL, "channel_stats: bad arguments");
  ADNI_REQUIRE(C % 8 == 0 && C >= 8 && C <= 2048, ADNI_ENOTSUP, "channel_stats: C=%d must be a multiple of 8 in [8,2048]",
               C);
  const int slab = red_slab_rows(rows, C);
  const int grid = (int)((rows + slab - 1) / slab);
  pdl_launch(channel_reduce_kernel<0>, grid, kEwThreads, 0, ST(stream))(CBF(x), nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                 nullptr, rows, C, 0, slab, sum, sqsum);
  count_launch();
  ADNI_LAUNCH_CHECK("channel_reduce_kernel<0>");
  return ADNI_OK;
}

int adni_bn_apply(const adni_bf16* y, const float* scale, const float* shift, const adni_bf16* residual, adni_bf16* out,
                  long long rows, int C, int relu, double* out_sum, double* out_sqsum, void* stream) {
  ADNI_REQUIRE(y && scale && shift && out && rows > 0, ADNI_EINVAL, "bn_apply: bad arguments");
  ADNI_REQUIRE(C % 8 == 0 && C >= 8, ADNI_ENOTSUP, "bn_apply: C=%d must be a multiple of 8", C);
  const long long nvec = rows * (C / 8);
  pdl_launch(bn_apply_kernel, ew_grid(nvec, kEwThreads * 4), kEwThreads, 0, ST(stream))(CBF(y), scale, shift, CBF(residual),
                                                                                 BF(out), nvec, C / 8, relu);
  count_launch();
  ADNI_LAUNCH_CHECK("bn_apply_kernel");
  if (out_sum && out_sqsum) return adni_channel_stats(out, rows, C, out_sum, out_sqsum, stream);
  return ADNI_OK;
}

int adni_bn_bwd_reduce(const adni_bf16* dout, const adni_bf16* out, const adni_bf16* y, const float* mean,
                       const float* invstd, const float* scale, const float* shift, long long rows, int C, int relu,
                       double* red, void* stream) {
  ADNI_REQUIRE(dout && y && mean && invstd && red && rows > 0, ADNI_EINVAL, "bn_bwd_reduce: bad arguments");
  ADNI_REQUIRE(!relu || out || (scale && shift), ADNI_EINVAL,
               "bn_bwd_reduce: relu mask needs the forward output or scale/shift");
  if (relu) relu = out ? 1 : 2;
  ADNI_REQUIRE(C % 8 == 0 && C >= 8 && C <= 2048, ADNI_ENOTSUP, "bn_bwd_reduce: C=%d must be a multiple of 8 in [8,2048]",
               C);
  const int slab = red_slab_rows(rows, C);
  const int grid = (int)((rows + slab - 1) / slab);
  pdl_launch(channel_reduce_kernel<1>, grid, kEwThreads, 0, ST(stream))(CBF(dout), CBF(out), CBF(y), mean, invstd, scale, shift,
                                                                 rows, C, relu, slab, red, red + C);
  count_launch();
  ADNI_LAUNCH_CHECK("channel_reduce_kernel<1>");
  return ADNI_OK;
}

int adni_bn_bwd_apply(const adni_bf16* dout, const adni_bf16* out, const adni_bf16* y, const float* mean,
                      const float* invstd, const float* gamma, const float* scale, const float* shift,
                      const double* red, double count, long long rows, int C, int relu, adni_bf16* dy, adni_bf16* dres,
                      float* dgamma, float* dbeta, double param_grad_scale, int red_form, void* stream) {
  ADNI_REQUIRE(dout && y && mean && invstd && red && dy && rows > 0 && count > 0 && (red_form == 0 || red_form == 1),
               ADNI_EINVAL, "bn_bwd_apply: bad arguments");
  ADNI_REQUIRE(!relu || out || (scale && shift), ADNI_EINVAL,
               "bn_bwd_apply: relu mask needs the forward output or scale/shift");
  if (relu) relu = out ? 1 : 2;
  ADNI_REQUIRE(C % 8 == 0 && C >= 8, ADNI_ENOTSUP, "bn_bwd_apply: C=%d must be a multiple of 8", C);
  const long long nvec = rows * (C / 8);
  ADNI_REQUIRE(C <= 2048, ADNI_ENOTSUP, "bn_bwd_apply: C=%d > 2048", C);
  pdl_launch(bn_bwd_apply_kernel, ew_grid(nvec, kEwThreads * 4), kEwThreads, 5 * C * sizeof(float), ST(stream))(
      CBF(dout), CBF(out), CBF(y), mean, invstd, gamma, scale, shift, red, 1.0 / count, nvec, C / 8, C, relu, BF(dy),
      BF(dres), dgamma, dbeta, param_grad_scale, red_form);
  count_launch();
  ADNI_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return ADNI_OK;
}

int adni_bn_train_apply(const adni_bf16* y, const double* stat_sum, const double* stat_sqsum, double count,
                        const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                        float* running_var, float* bnp, const adni_bf16* residual, adni_bf16* out, long long rows, int C,
                        int relu, void* stream) {
  ADNI_REQUIRE(y && stat_sum && stat_sqsum && bnp && out && rows > 0 && count > 0, ADNI_EINVAL,
               "bn_train_apply: bad arguments");
  ADNI_REQUIRE(C % 8 == 0 && C >= 8, ADNI_ENOTSUP, "bn_train_apply: C=%d must be a multiple of 8", C);
  const int vpr = C / 8;
  if (kEwThreads % vpr != 0) {  // a thread's channels change from vector to vector: finalize first, then apply
    int rc = adni_bn_finalize(stat_sum, stat_sqsum, count, C, gamma, beta, eps, momentum, running_mean, running_var, bnp,
                              bnp + C, bnp + 2 * C, bnp + 3 * C, stream);
    if (rc) return rc;
    return adni_bn_apply(y, bnp + 2 * C, bnp + 3 * C, residual, out, rows, C, relu, nullptr, nullptr, stream);
  }
  const long long nvec = rows * vpr;
  pdl_launch(bn_train_apply_kernel, ew_grid(nvec, kEwThreads * 4), kEwThreads, 0, ST(stream))(
      CBF(y), stat_sum, stat_sqsum, count, gamma, beta, eps, momentum, running_mean, running_var, bnp, CBF(residual),
      BF(out), nvec, vpr, C, relu);
  count_launch();
  ADNI_LAUNCH_CHECK("bn_train_apply_kernel");
  return ADNI_OK;
}

int adni_maxpool3d_fwd(const adni_bf16* x, int N, int D, int H, int W, int C, int k, int stride, int pad, adni_bf16* y,
                       uint8_t* argmax, void* stream) {
  ADNI_REQUIRE(x && y && argmax, ADNI_EINVAL, "maxpool3d_fwd: null pointer");
  ADNI_REQUIRE(C % 8 == 0 && k >= 1 && k * k * k <= 255 && stride >= 1 && 2 * pad <= k, ADNI_ENOTSUP,
               "maxpool3d_fwd: unsupported C=%d k=%d stride=%d pad=%d", C, k, stride, pad);
  const int Do = (D + 2 * pad - k) / stride + 1, Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  ADNI_REQUIRE(Do > 0 && Ho > 0 && Wo > 0, ADNI_EINVAL, "maxpool3d_fwd: empty output");
  const long long total = (long long)N * Do * Ho * Wo * (C / 8);
  pdl_launch(maxpool_fwd_kernel, ew_grid(total, kEwThreads), kEwThreads, 0, ST(stream))(CBF(x), N, D, H, W, C, k, stride, pad,
                                                                                 Do, Ho, Wo, BF(y), argmax);
  count_launch();
  ADNI_LAUNCH_CHECK("maxpool_fwd_kernel");
  return ADNI_OK;
}

int adni_maxpool3d_bwd(const adni_bf16* dy, const uint8_t* argmax, int N, int D, int H, int W, int C, int k, int stride,
                       int pad, adni_bf16* dx, void* stream) {
  ADNI_REQUIRE(dy && dx && argmax, ADNI_EINVAL, "maxpool3d_bwd: null pointer");
  ADNI_REQUIRE(C % 8 == 0 && k >= 1 && k * k * k <= 255 && stride >= 1 && 2 * pad <= k, ADNI_ENOTSUP,
               "maxpool3d_bwd: unsupported C=%d k=%d stride=%d pad=%d", C, k, stride, pad);
  if (stride == k && pad == 0)   // non-overlapping windows (MaxPool3d(2)): one thread per input vector, pool_small.cu
    return launch_pool_nonoverlap_bwd(CBF(dy), argmax, nullptr, N, D, H, W, C, k, BF(dx), ST(stream));
  const int Do = (D + 2 * pad - k) / stride + 1, Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const long long total = (long long)N * D * H * W * (C / 8);
  // windows that can touch an 8-voxel tile edge: (8 + k - 2)/s + 1 (+1 for unaligned tiles when s does not divide 8)
  const int wmax = (kMpT + k - 2) / stride + 1 + ((kMpT % stride) ? 1 : 0);
  if (wmax <= kMpWMax) {
    const int td = (D + kMpT - 1) / kMpT, th = (H + kMpT - 1) / kMpT, tw = (W + kMpT - 1) / kMpT;
    dim3 grid((unsigned)((long long)N * td * th * tw), (unsigned)((C / 8 + 7) / 8));
    pdl_launch(maxpool_bwd_tiled_kernel, grid, kEwThreads, 0, ST(stream))(CBF(dy), argmax, N, D, H, W, C, k, stride, pad, Do, Ho,
                                                                  Wo, td, th, tw, BF(dx));
  } else {
    pdl_launch(maxpool_bwd_kernel, ew_grid(total, kEwThreads), kEwThreads, 0, ST(stream))(CBF(dy), argmax, N, D, H, W, C, k,
                                                                                   stride, pad, Do, Ho, Wo, BF(dx));
  }
  count_launch();
  ADNI_LAUNCH_CHECK("maxpool_bwd_kernel");
  return ADNI_OK;
}

int adni_gap_fwd(const adni_bf16* x, int N, long long P, int C, float* feat, void* stream) {
  ADNI_REQUIRE(x && feat && N > 0 && P > 0, ADNI_EINVAL, "gap_fwd: bad arguments");
  ADNI_REQUIRE(C % 8 == 0 && N <= 65535, ADNI_ENOTSUP, "gap_fwd: C=%d must be a multiple of 8", C);
  ADNI_CUDA_OK(cudaMemsetAsync(feat, 0, sizeof(float) * (size_t)N * C, ST(stream)));
  const int split = (int)std::min<long long>(kGapSplit, (P + 63) / 64);
  dim3 grid((C / 8 + 31) / 32, split, N);
  pdl_launch(gap_fwd_kernel, grid, kEwThreads, 0, ST(stream))(CBF(x), P, C, 1.0f / (float)P, feat);
  count_launch();
  ADNI_LAUNCH_CHECK("gap_fwd_kernel");
  return ADNI_OK;
}

int adni_gap_bwd(const float* dfeat, int N, long long P, int C, adni_bf16* dx, void* stream) {
  ADNI_REQUIRE(dfeat && dx && N > 0 && P > 0, ADNI_EINVAL, "gap_bwd: bad arguments");
  ADNI_REQUIRE(C % 8 == 0, ADNI_ENOTSUP, "gap_bwd: C=%d must be a multiple of 8", C);
  const long long total = (long long)N * P * (C / 8);
  pdl_launch(gap_bwd_kernel, ew_grid(total, kEwThreads * 4), kEwThreads, 0, ST(stream))(dfeat, N, P, C, 1.0f / (float)P,
                                                                                 BF(dx));
  count_launch();
  ADNI_LAUNCH_CHECK("gap_bwd_kernel");
  return ADNI_OK;
}

int adni_cast_f32_to_bf16(const float* x, adni_bf16* y, long long n, void* stream) {
  ADNI_REQUIRE(x && y && n > 0, ADNI_EINVAL, "cast: bad arguments");
  pdl_launch(cast_to_bf16_kernel<float>, ew_grid(n, kEwThreads * 8), kEwThreads, 0, ST(stream))(x, BF(y), n);
  count_launch();
  ADNI_LAUNCH_CHECK("cast_to_bf16_kernel");
  return ADNI_OK;
}
int adni_cast_f64_to_bf16(const double* x, adni_bf16* y, long long n, void* stream) {
  ADNI_REQUIRE(x && y && n > 0, ADNI_EINVAL, "cast: bad arguments");
  pdl_launch(cast_to_bf16_kernel<double>, ew_grid(n, kEwThreads * 8), kEwThreads, 0, ST(stream))(x, BF(y), n);
  count_launch();
  ADNI_LAUNCH_CHECK("cast_to_bf16_kernel");
  return ADNI_OK;
}
int adni_cast_bf16_to_f32(const adni_bf16* x, float* y, long long n, void* stream) {
  ADNI_REQUIRE(x && y && n > 0, ADNI_EINVAL, "cast: bad arguments");
  pdl_launch(cast_bf16_to_f32_kernel, ew_grid(n, kEwThreads * 8), kEwThreads, 0, ST(stream))(CBF(x), y, n);
  count_launch();
  ADNI_LAUNCH_CHECK("cast_bf16_to_f32_kernel");
  return ADNI_OK;
}

int adni_weights_to_kernel_layout(const float* w_ncdhw, int Cout, int Cin, int taps, adni_bf16* w_oti, adni_bf16* w_ito,
                                  void* stream) {
  ADNI_REQUIRE(w_ncdhw && Cout > 0 && Cin > 0 && taps > 0, ADNI_EINVAL, "weights_to_kernel_layout: bad arguments");
  int rc = ADNI_OK;
  // OTI: per output channel, [Cin][taps] -> [taps][Cin]
  if (w_oti) rc = launch_transpose<bf16>(w_ncdhw, BF(w_oti), Cout, Cin, taps, false, ST(stream));
  // ITO: [Cout][Cin*taps] -> [Cin*taps][Cout]
  if (rc == ADNI_OK && w_ito) rc = launch_transpose<bf16>(w_ncdhw, BF(w_ito), 1, Cout, Cin * taps, false, ST(stream));
  return rc;
}

int adni_weights_multi_job_bytes(void) { return (int)sizeof(WeightJob); }

/* jobs: device array of n_jobs records {src, oti, ito, cout, cin, taps, tile_begin, tiles_ci} (see
 * adni_weights_multi_job_bytes and multimodal_alzheimer_b200/kernels.py: WeightArena); total_tiles = sum of tiles. */
int adni_weights_to_kernel_layout_multi(const void* jobs, int n_jobs, int total_tiles, void* stream) {
  ADNI_REQUIRE(jobs && n_jobs > 0 && total_tiles > 0, ADNI_EINVAL, "weights_to_kernel_layout_multi: bad arguments");
  pdl_launch(weights_multi_kernel, total_tiles, 256, 0, ST(stream))(static_cast<const WeightJob*>(jobs), n_jobs);
  count_launch();
  ADNI_LAUNCH_CHECK("weights_multi_kernel");
  return ADNI_OK;
}

int adni_wgrad_to_param_layout(const float* dw_oti, int Cout, int Cin, int taps, float* grad_ncdhw, int accumulate,
                               void* stream) {
  ADNI_REQUIRE(dw_oti && grad_ncdhw && Cout > 0 && Cin > 0 && taps > 0, ADNI_EINVAL,
               "wgrad_to_param_layout: bad arguments");
  // per output channel, [taps][Cin] -> [Cin][taps]
  return launch_transpose<float>(dw_oti, grad_ncdhw, Cout, taps, Cin, accumulate != 0, ST(stream));
}

}  // extern "C"

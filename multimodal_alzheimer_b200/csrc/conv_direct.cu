// CUDA-core direct Conv3d engine for small channel counts: the Small_PET_CNN stack
// (pkg/models/pet_models/pet_cnn.py:18-28: Conv3d 'same' + bias, Cin in {1,8,16,32}, Cout in {8..64}, k in
// {3,5,7}) and the 1-channel 7x7x7 stride-2 stem of MedicalNet's ResNet.  These layers are far below the
// tensor-core tile granularity (K per tap < 64), so they run as register-tiled direct convolutions: bf16
// NDHWC activations, fp32 accumulation, weights staged in shared memory.
#include <algorithm>

#include "common.cuh"

namespace adni {

extern void count_launch();

namespace {

constexpr int kCoTile = 16;   // output channels per thread
constexpr int kThreads = 128;

struct DirectParams {
  const __nv_bfloat16* in;   // [N][Di][Hi][Wi][Ci]
  const __nv_bfloat16* w;    // [Co][taps][Ci]
  const float* bias;         // [Co] or null
  const __nv_bfloat16* addend;  // same shape as out or null
  __nv_bfloat16* out;        // [N][Do][Ho][Wo][Co]
  double* ssum;
  double* ssq;
  int N, Di, Hi, Wi, Ci;
  int Do, Ho, Wo, Co;
  int k, stride, pad, dil;
  int ci_chunk;  // input channels staged per smem pass
};

__device__ __forceinline__ float ldbf(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// TRANSPOSED = false: out[o] = sum_k in[o*stride + k*dil - pad] * w[k]         (fprop)
// TRANSPOSED = true : out[i] = sum_k in[(i + pad - k*dil)/stride] * w[k]      (dgrad; `in` is dy, w is ITO)
template <bool TRANSPOSED>
__global__ void __launch_bounds__(kThreads) direct_conv_kernel(const DirectParams p) {
  pdl_enter();
  extern __shared__ float w_s[];  // [taps][ci_chunk][kCoTile]
  const int taps = p.k * p.k * p.k;
  const int co0 = blockIdx.y * kCoTile;
  const long long npos = (long long)p.N * p.Do * p.Ho * p.Wo;
  const long long pos = (long long)blockIdx.x * kThreads + threadIdx.x;
  const bool active = pos < npos;
  int n = 0, od = 0, oh = 0, ow = 0;
  if (active) {
    long long r = pos;
    ow = int(r % p.Wo);
    r /= p.Wo;
    oh = int(r % p.Ho);
    r /= p.Ho;
    od = int(r % p.Do);
    n = int(r / p.Do);
  }
  float acc[kCoTile];
#pragma unroll
  for (int j = 0; j < kCoTile; j++) acc[j] = 0.f;

  for (int c0 = 0; c0 < p.Ci; c0 += p.ci_chunk) {
    const int cc = min(p.ci_chunk, p.Ci - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < taps * cc * kCoTile; i += kThreads) {
      const int j = i % kCoTile;
      const int ci = (i / kCoTile) % cc;
      const int t = i / (kCoTile * cc);
      const int co = co0 + j;
      w_s[i] = co < p.Co ? ldbf(p.w + ((long long)co * taps + t) * p.Ci + c0 + ci) : 0.f;
    }
    __syncthreads();
    if (active) {
      int t = 0;
      for (int kd = 0; kd < p.k; kd++) {
        int id;
        bool okd;
        if (!TRANSPOSED) {
          id = od * p.stride + kd * p.dil - p.pad;
          okd = id >= 0 && id < p.Di;
        } else {
          const int num = od + p.pad - kd * p.dil;
          id = num / p.stride;
          okd = num >= 0 && num % p.stride == 0 && id < p.Di;
        }
        for (int kh = 0; kh < p.k; kh++) {
          int ih;
          bool okh;
          if (!TRANSPOSED) {
            ih = oh * p.stride + kh * p.dil - p.pad;
            okh = ih >= 0 && ih < p.Hi;
          } else {
            const int num = oh + p.pad - kh * p.dil;
            ih = num / p.stride;
            okh = num >= 0 && num % p.stride == 0 && ih < p.Hi;
          }
          for (int kw = 0; kw < p.k; kw++, t++) {
            int iw;
            bool okw;
            if (!TRANSPOSED) {
              iw = ow * p.stride + kw * p.dil - p.pad;
              okw = iw >= 0 && iw < p.Wi;
            } else {
              const int num = ow + p.pad - kw * p.dil;
              iw = num / p.stride;
              okw = num >= 0 && num % p.stride == 0 && iw < p.Wi;
            }
            if (!(okd && okh && okw)) continue;
            const __nv_bfloat16* ip = p.in + ((((long long)n * p.Di + id) * p.Hi + ih) * p.Wi + iw) * p.Ci + c0;
            const float* wp = w_s + (long long)t * cc * kCoTile;
            if ((cc & 7) == 0 && (p.Ci & 7) == 0) {
              for (int ci = 0; ci < cc; ci += 8) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(ip + ci));
                const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int e = 0; e < 8; e++) {
                  const float xv = (e & 1) ? bf16_hi(rw[e >> 1]) : bf16_lo(rw[e >> 1]);
                  const float4* w4 = reinterpret_cast<const float4*>(wp + (ci + e) * kCoTile);
#pragma unroll
                  for (int j4 = 0; j4 < kCoTile / 4; j4++) {
                    const float4 wv = w4[j4];
                    acc[j4 * 4 + 0] = fmaf(xv, wv.x, acc[j4 * 4 + 0]);
                    acc[j4 * 4 + 1] = fmaf(xv, wv.y, acc[j4 * 4 + 1]);
                    acc[j4 * 4 + 2] = fmaf(xv, wv.z, acc[j4 * 4 + 2]);
                    acc[j4 * 4 + 3] = fmaf(xv, wv.w, acc[j4 * 4 + 3]);
                  }
                }
              }
            } else {
              for (int ci = 0; ci < cc; ci++) {
                const float xv = ldbf(ip + ci);
                const float4* w4 = reinterpret_cast<const float4*>(wp + ci * kCoTile);
#pragma unroll
                for (int j4 = 0; j4 < kCoTile / 4; j4++) {
                  const float4 wv = w4[j4];
                  acc[j4 * 4 + 0] = fmaf(xv, wv.x, acc[j4 * 4 + 0]);
                  acc[j4 * 4 + 1] = fmaf(xv, wv.y, acc[j4 * 4 + 1]);
                  acc[j4 * 4 + 2] = fmaf(xv, wv.z, acc[j4 * 4 + 2]);
                  acc[j4 * 4 + 3] = fmaf(xv, wv.w, acc[j4 * 4 + 3]);
                }
              }
            }
          }
        }
      }
    }
  }

  // epilogue: bias, residual-gradient addend, store, per-channel statistics
  const long long obase = pos * p.Co + co0;
#pragma unroll
  for (int j = 0; j < kCoTile; j++) {
    const int co = co0 + j;
    if (co < p.Co) {
      if (p.bias) acc[j] += p.bias[co];
      if (active && p.addend) acc[j] += ldbf(p.addend + obase + j);
    }
  }
  if (active) {
    if ((p.Co & 7) == 0 && co0 + kCoTile <= p.Co) {
      uint4 o0, o1;
      o0.x = pack_bf16x2(acc[0], acc[1]);
      o0.y = pack_bf16x2(acc[2], acc[3]);
      o0.z = pack_bf16x2(acc[4], acc[5]);
      o0.w = pack_bf16x2(acc[6], acc[7]);
      o1.x = pack_bf16x2(acc[8], acc[9]);
      o1.y = pack_bf16x2(acc[10], acc[11]);
      o1.z = pack_bf16x2(acc[12], acc[13]);
      o1.w = pack_bf16x2(acc[14], acc[15]);
      uint4* op = reinterpret_cast<uint4*>(p.out + obase);
      op[0] = o0;
      op[1] = o1;
    } else {
#pragma unroll
      for (int j = 0; j < kCoTile; j++)
        if (co0 + j < p.Co) p.out[obase + j] = __float2bfloat16_rn(acc[j]);
    }
  }
  if (p.ssum != nullptr) {
    __shared__ float red[2][kThreads / 32][kCoTile];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < kCoTile; j++) {
      const float v = active ? acc[j] : 0.f;
      const float s1 = warp_sum(v);
      const float s2 = warp_sum(v * v);
      if (lane == 0) {
        red[0][warp][j] = s1;
        red[1][warp][j] = s2;
      }
    }
    __syncthreads();
    if (threadIdx.x < kCoTile && co0 + threadIdx.x < p.Co) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < kThreads / 32; w++) {
        a += red[0][w][threadIdx.x];
        b += red[1][w][threadIdx.x];
      }
      atomicAdd(p.ssum + co0 + threadIdx.x, (double)a);
      atomicAdd(p.ssq + co0 + threadIdx.x, (double)b);
    }
  }
}

// wgrad: dw[co][t][ci] += sum_o dy[o][co] * x[o*stride + off_t][ci].  A block owns a slab of output
// positions; thread j owns column (t, ci) and keeps Co_tile accumulators; dy slab is staged in smem.
constexpr int kWgSlab = 256;
constexpr int kWgThreads = 256;
constexpr int kWgCoTile = 32;

struct DirectWgradParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* dy;
  float* dw;
  float* dbias;
  int N, Di, Hi, Wi, Ci;
  int Do, Ho, Wo, Co;
  int k, stride, pad, dil;
};

__global__ void __launch_bounds__(kWgThreads) direct_wgrad_kernel(const DirectWgradParams p) {
  pdl_enter();
  __shared__ float dy_s[kWgSlab][kWgCoTile];
  __shared__ int pos_s[kWgSlab][4];  // n, id0, ih0, iw0 (input origin of the receptive field) ; n = -1 if inactive
  const int taps = p.k * p.k * p.k;
  const int ncols = taps * p.Ci;
  const int co0 = blockIdx.y * kWgCoTile;
  const long long npos = (long long)p.N * p.Do * p.Ho * p.Wo;
  const long long slab0 = (long long)blockIdx.x * kWgSlab;

  for (int i = threadIdx.x; i < kWgSlab; i += kWgThreads) {
    const long long pos = slab0 + i;
    if (pos < npos) {
      long long r = pos;
      const int ow = int(r % p.Wo);
      r /= p.Wo;
      const int oh = int(r % p.Ho);
      r /= p.Ho;
      const int od = int(r % p.Do);
      pos_s[i][0] = int(r / p.Do);
      pos_s[i][1] = od * p.stride - p.pad;
      pos_s[i][2] = oh * p.stride - p.pad;
      pos_s[i][3] = ow * p.stride - p.pad;
    } else {
      pos_s[i][0] = -1;
    }
  }
  for (int i = threadIdx.x; i < kWgSlab * kWgCoTile; i += kWgThreads) {
    const int o = i / kWgCoTile, j = i % kWgCoTile;
    const long long pos = slab0 + o;
    const int co = co0 + j;
    dy_s[o][j] = (pos < npos && co < p.Co) ? ldbf(p.dy + pos * p.Co + co) : 0.f;
  }
  __syncthreads();

  if (p.dbias != nullptr && blockIdx.z == 0 && threadIdx.x < kWgCoTile && co0 + threadIdx.x < p.Co) {
    float s = 0.f;
    for (int o = 0; o < kWgSlab; o++) s += dy_s[o][threadIdx.x];
    atomicAdd(p.dbias + co0 + threadIdx.x, s);
  }

  for (int col = blockIdx.z * kWgThreads + threadIdx.x; col < ncols; col += gridDim.z * kWgThreads) {
    const int t = col / p.Ci, ci = col % p.Ci;
    const int kw = t % p.k, kh = (t / p.k) % p.k, kd = t / (p.k * p.k);
    const int offd = kd * p.dil, offh = kh * p.dil, offw = kw * p.dil;
    float acc[kWgCoTile];
#pragma unroll
    for (int j = 0; j < kWgCoTile; j++) acc[j] = 0.f;
    for (int o = 0; o < kWgSlab; o++) {
      const int n = pos_s[o][0];
      if (n < 0) break;
      const int id = pos_s[o][1] + offd, ih = pos_s[o][2] + offh, iw = pos_s[o][3] + offw;
      if (id < 0 || id >= p.Di || ih < 0 || ih >= p.Hi || iw < 0 || iw >= p.Wi) continue;
      const float xv = ldbf(p.x + ((((long long)n * p.Di + id) * p.Hi + ih) * p.Wi + iw) * p.Ci + ci);
      const float4* d4 = reinterpret_cast<const float4*>(&dy_s[o][0]);
#pragma unroll
      for (int j4 = 0; j4 < kWgCoTile / 4; j4++) {
        const float4 dv = d4[j4];
        acc[j4 * 4 + 0] = fmaf(xv, dv.x, acc[j4 * 4 + 0]);
        acc[j4 * 4 + 1] = fmaf(xv, dv.y, acc[j4 * 4 + 1]);
        acc[j4 * 4 + 2] = fmaf(xv, dv.z, acc[j4 * 4 + 2]);
        acc[j4 * 4 + 3] = fmaf(xv, dv.w, acc[j4 * 4 + 3]);
      }
    }
#pragma unroll
    for (int j = 0; j < kWgCoTile; j++)
      if (co0 + j < p.Co) atomicAdd(p.dw + (long long)(co0 + j) * ncols + col, acc[j]);
  }
}

int launch_direct(const DirectParams& p0, bool transposed, cudaStream_t stream) {
  DirectParams p = p0;
  const int taps = p.k * p.k * p.k;
  // stage as many input channels as fit in 40 KB of fp32 weights
  int chunk = (40 * 1024) / (taps * kCoTile * 4);
  if (chunk < 1) {
    set_error("direct conv: %d taps do not fit the shared-memory weight stage", taps);
    return ADNI_ENOTSUP;
  }
  if (chunk >= 8) chunk &= ~7;
  p.ci_chunk = std::min(chunk, p.Ci);
  const size_t smem = (size_t)taps * p.ci_chunk * kCoTile * 4;
  const long long npos = (long long)p.N * p.Do * p.Ho * p.Wo;
  dim3 grid((unsigned)((npos + kThreads - 1) / kThreads), (unsigned)((p.Co + kCoTile - 1) / kCoTile));
  if (transposed)
    pdl_launch(direct_conv_kernel<true>, grid, kThreads, smem, stream)(p);
  else
    pdl_launch(direct_conv_kernel<false>, grid, kThreads, smem, stream)(p);
  count_launch();
  ADNI_LAUNCH_CHECK("direct_conv_kernel");
  return ADNI_OK;
}

inline int oext(int in, int k, int s, int pad, int dil) { return (in + 2 * pad - dil * (k - 1) - 1) / s + 1; }

}  // namespace

int direct_conv_fprop(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* w_oti,
                      const float* bias, __nv_bfloat16* y, double* ssum, double* ssq, cudaStream_t stream) {
  DirectParams p{};
  p.in = x;
  p.w = w_oti;
  p.bias = bias;
  p.addend = nullptr;
  p.out = y;
  p.ssum = ssum;
  p.ssq = ssq;
  p.N = g.N;
  p.Di = g.D;
  p.Hi = g.H;
  p.Wi = g.W;
  p.Ci = g.Cin;
  p.Do = oext(g.D, g.k, g.stride, g.pad, g.dil);
  p.Ho = oext(g.H, g.k, g.stride, g.pad, g.dil);
  p.Wo = oext(g.W, g.k, g.stride, g.pad, g.dil);
  p.Co = g.Cout;
  p.k = g.k;
  p.stride = g.stride;
  p.pad = g.pad;
  p.dil = g.dil;
  return launch_direct(p, false, stream);
}

int direct_conv_dgrad(const adni_conv3d_geom& g, const __nv_bfloat16* dy, const __nv_bfloat16* w_ito,
                      const __nv_bfloat16* addend, __nv_bfloat16* dx, cudaStream_t stream) {
  DirectParams p{};
  p.in = dy;
  p.w = w_ito;
  p.bias = nullptr;
  p.addend = addend;
  p.out = dx;
  p.N = g.N;
  p.Di = oext(g.D, g.k, g.stride, g.pad, g.dil);
  p.Hi = oext(g.H, g.k, g.stride, g.pad, g.dil);
  p.Wi = oext(g.W, g.k, g.stride, g.pad, g.dil);
  p.Ci = g.Cout;
  p.Do = g.D;
  p.Ho = g.H;
  p.Wo = g.W;
  p.Co = g.Cin;
  p.k = g.k;
  p.stride = g.stride;
  p.pad = g.pad;
  p.dil = g.dil;
  return launch_direct(p, true, stream);
}

int direct_conv_wgrad(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw,
                      float* dbias, cudaStream_t stream) {
  DirectWgradParams p{};
  p.x = x;
  p.dy = dy;
  p.dw = dw;
  p.dbias = dbias;
  p.N = g.N;
  p.Di = g.D;
  p.Hi = g.H;
  p.Wi = g.W;
  p.Ci = g.Cin;
  p.Do = oext(g.D, g.k, g.stride, g.pad, g.dil);
  p.Ho = oext(g.H, g.k, g.stride, g.pad, g.dil);
  p.Wo = oext(g.W, g.k, g.stride, g.pad, g.dil);
  p.Co = g.Cout;
  p.k = g.k;
  p.stride = g.stride;
  p.pad = g.pad;
  p.dil = g.dil;
  const long long npos = (long long)p.N * p.Do * p.Ho * p.Wo;
  const int ncols = g.k * g.k * g.k * g.Cin;
  const unsigned zc = (unsigned)std::min((ncols + kWgThreads - 1) / kWgThreads, 8);
  dim3 grid((unsigned)((npos + kWgSlab - 1) / kWgSlab), (unsigned)((p.Co + kWgCoTile - 1) / kWgCoTile), zc);
  pdl_launch(direct_wgrad_kernel, grid, kWgThreads, 0, stream)(p);
  count_launch();
  ADNI_LAUNCH_CHECK("direct_wgrad_kernel");
  return ADNI_OK;
}

}  // namespace adni

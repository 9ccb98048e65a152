// Fused element-wise path of the ResNet stem: bn1 -> ReLU -> MaxPool3d(3,2,1) forward, and the matching backward
// (max-pool scatter + ReLU mask + BatchNorm backward) WITHOUT ever materialising the 64x64^3 activated tensor or
// its gradient (33.5 MB per volume each).  HBM traffic per volume: forward reads y once; backward reads y twice
// (reduction pass, apply pass) and writes dy once.
//   MedicalNet ResNet.forward: x = maxpool(relu(bn1(conv1(x))))   (call site pkg/models/mri_models/anat_cnn.py:95)
#include <algorithm>

#include "common.cuh"

namespace adni {
extern void count_launch();
namespace {

constexpr int kThreads = 256;
constexpr int kK = 3, kS = 2, kPad = 1;  // MaxPool3d(kernel_size=3, stride=2, padding=1)

__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]) {
  v[0] = bf16_lo(r.x);
  v[1] = bf16_hi(r.x);
  v[2] = bf16_lo(r.y);
  v[3] = bf16_hi(r.y);
  v[4] = bf16_lo(r.z);
  v[5] = bf16_hi(r.z);
  v[6] = bf16_lo(r.w);
  v[7] = bf16_hi(r.w);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]);
  o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]);
  o.w = pack_bf16x2(v[6], v[7]);
  return o;
}

// ------------------------------------------------------------------------------------------------ forward
// block: 4x4x4 pooled outputs x 32 channels (4 vectors of 8); input tile 9x9x9 voxels staged activated in smem
constexpr int kFT = 4;
constexpr int kFI = kFT * kS + 1;  // 9
constexpr int kFCv = 4;

__global__ void __launch_bounds__(kThreads)
    bn_relu_pool_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                            const float* __restrict__ shift, int N, int D, int H, int W, int C, int Do, int Ho, int Wo,
                            int tiles_d, int tiles_h, int tiles_w, __nv_bfloat16* __restrict__ p,
                            uint8_t* __restrict__ amax) {
  __shared__ uint4 s_a[kFI * kFI * kFI * kFCv];
  const int vpr = C / 8;
  // channel chunks are the fastest-varying block index: the blocks sharing a voxel tile (and its 128-byte lines)
  // run at the same time, so the second half of every line is an L2 hit instead of a second DRAM read
  const int n_chunks = (vpr + kFCv - 1) / kFCv;
  const int cv0 = (blockIdx.x % n_chunks) * kFCv;
  int t = blockIdx.x / n_chunks;
  const int tw = t % tiles_w;
  t /= tiles_w;
  const int th = t % tiles_h;
  t /= tiles_h;
  const int td = t % tiles_d;
  const int n = t / tiles_d;
  const int od0 = td * kFT, oh0 = th * kFT, ow0 = tw * kFT;
  const int id0 = od0 * kS - kPad, ih0 = oh0 * kS - kPad, iw0 = ow0 * kS - kPad;
  const uint32_t ninf2 = 0xFF80FF80u;  // bf16 -inf pair: the padding value of max-pool
#pragma unroll 4
  for (int i = threadIdx.x; i < kFI * kFI * kFI * kFCv; i += kThreads) {
    const int cv = i % kFCv, v = i / kFCv;
    const int iw = iw0 + v % kFI, ih = ih0 + (v / kFI) % kFI, id = id0 + v / (kFI * kFI);
    uint4 out = make_uint4(ninf2, ninf2, ninf2, ninf2);
    if (id >= 0 && id < D && ih >= 0 && ih < H && iw >= 0 && iw < W && cv0 + cv < vpr) {
      const int c0 = (cv0 + cv) * 8;
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(y + ((((long long)n * D + id) * H + ih) * W + iw) * C + c0));
      float f[8];
      unpack8(raw, f);
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c0)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c0) + 1);
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c0)), h1 = __ldg(reinterpret_cast<const float4*>(shift + c0) + 1);
      const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int j = 0; j < 8; j++) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
      out = pack8(f);  // bf16-rounded activation: identical to what bn_apply would have stored
    }
    s_a[i] = out;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kFT * kFT * kFT * kFCv; i += kThreads) {
    const int cv = i % kFCv, o = i / kFCv;
    const int lw = o % kFT, lh = (o / kFT) % kFT, ld = o / (kFT * kFT);
    const int od = od0 + ld, oh = oh0 + lh, ow = ow0 + lw;
    if (od >= Do || oh >= Ho || ow >= Wo || cv0 + cv >= vpr) continue;
    float best[8];
    int bidx[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      best[j] = -INFINITY;
      bidx[j] = -1;
    }
    for (int kd = 0; kd < kK; kd++) {
      const int id = od * kS - kPad + kd;
      if (id < 0 || id >= D) continue;
      for (int kh = 0; kh < kK; kh++) {
        const int ih = oh * kS - kPad + kh;
        if (ih < 0 || ih >= H) continue;
        for (int kw = 0; kw < kK; kw++) {
          const int iw = ow * kS - kPad + kw;
          if (iw < 0 || iw >= W) continue;
          float f[8];
          unpack8(s_a[(((ld * kS + kd) * kFI + (lh * kS + kh)) * kFI + (lw * kS + kw)) * kFCv + cv], f);
          const int slot = (kd * kK + kh) * kK + kw;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (bidx[j] < 0 || f[j] > best[j]) {  // first maximum in (d,h,w) scan order wins
              best[j] = f[j];
              bidx[j] = slot;
            }
          }
        }
      }
    }
    const long long oi = ((((long long)n * Do + od) * Ho + oh) * Wo + ow) * vpr + cv0 + cv;
    *reinterpret_cast<uint4*>(p + oi * 8) = pack8(best);
    uint2 pk;
    pk.x = (uint32_t)(bidx[0] & 255) | ((uint32_t)(bidx[1] & 255) << 8) | ((uint32_t)(bidx[2] & 255) << 16) |
           ((uint32_t)(bidx[3] & 255) << 24);
    pk.y = (uint32_t)(bidx[4] & 255) | ((uint32_t)(bidx[5] & 255) << 8) | ((uint32_t)(bidx[6] & 255) << 16) |
           ((uint32_t)(bidx[7] & 255) << 24);
    *reinterpret_cast<uint2*>(amax + oi * 8) = pk;
  }
}

// ------------------------------------------------------------------------------------------------ backward
// block: 8x8x8 input voxels x 32 channels.  Phase 1 scatters the gradients of the <= 5^3 pooled windows that can
// select a voxel of the tile into a shared-memory fp32 tile (each window contributes to exactly one voxel per
// channel: 8 shared atomics per (window, 8-channel vector) instead of probing up to 8 windows per voxel).
// Phase 2 streams y once: ReLU mask, then MODE 0: sum g, sum g*xhat (fp64 atomics) / MODE 1: dy = A*g + B*y + K.
constexpr int kBT = 8;
constexpr int kBCv = 4;  // 32 channels per block
constexpr int kBSmemFloats = kBT * kBT * kBT * kBCv * 8;

template <int MODE>
__global__ void __launch_bounds__(kThreads, 3)
    pool_bn_bwd_kernel(const __nv_bfloat16* __restrict__ dp, const uint8_t* __restrict__ amax,
                       const __nv_bfloat16* __restrict__ y, const float* __restrict__ bnp /* [4][C] */,
                       const float* __restrict__ gamma, const double* __restrict__ red_in, double inv_count, int N,
                       int D, int H, int W, int C, int Do, int Ho, int Wo, int tiles_d, int tiles_h, int tiles_w,
                       double* __restrict__ red_out, __nv_bfloat16* __restrict__ dy) {
  extern __shared__ float s_g[];  // [512 voxels][kBCv][8]
  __shared__ float s_red[2][kThreads / kBCv][kBCv * 8 + 1];
  const int vpr = C / 8;
  const int n_chunks = (vpr + kBCv - 1) / kBCv;  // fastest-varying: see bn_relu_pool_fwd_kernel
  const int cv0 = (blockIdx.x % n_chunks) * kBCv;
  int t = blockIdx.x / n_chunks;
  const int tw = t % tiles_w;
  t /= tiles_w;
  const int th = t % tiles_h;
  t /= tiles_h;
  const int td = t % tiles_d;
  const int n = t / tiles_d;
  const int i0d = td * kBT, i0h = th * kBT, i0w = tw * kBT;
  auto lo = [&](int i0) {
    const int num = i0 + kPad - kK + 1;
    return num <= 0 ? 0 : (num + kS - 1) / kS;
  };
  const int od0 = lo(i0d), oh0 = lo(i0h), ow0 = lo(i0w);
  const int nd = min(Do - 1, (i0d + kBT - 1 + kPad) / kS) - od0 + 1;
  const int nh = min(Ho - 1, (i0h + kBT - 1 + kPad) / kS) - oh0 + 1;
  const int nw = min(Wo - 1, (i0w + kBT - 1 + kPad) / kS) - ow0 + 1;
  // per-channel constants of this block's 32 channels in shared memory: mu, is, sc, sh, A, B, K
  __shared__ float s_c[7][kBCv * 8];
  const int cv = threadIdx.x % kBCv;  // the channel vector of a thread is fixed: kThreads % kBCv == 0
  const bool cv_ok = cv0 + cv < vpr;
  if (threadIdx.x < kBCv * 8) {
    const int c = min(cv0 * 8 + (int)threadIdx.x, C - 1);
    const float mu = bnp[c], is = bnp[C + c];
    s_c[0][threadIdx.x] = mu;
    s_c[1][threadIdx.x] = is;
    s_c[2][threadIdx.x] = bnp[2 * C + c];
    s_c[3][threadIdx.x] = bnp[3 * C + c];
    if (MODE == 1) {
      const float gm = gamma ? gamma[c] : 1.f;
      const float mg = (float)(red_in[c] * inv_count), mgx = (float)(red_in[C + c] * inv_count);
      const float a = gm * is, b = -a * is * mgx;
      s_c[4][threadIdx.x] = a;
      s_c[5][threadIdx.x] = b;
      s_c[6][threadIdx.x] = -a * mg - b * mu;
    }
  }
  // y loads are issued first: their latency overlaps the zeroing and the scatter phase
  constexpr int kIters = kBT * kBT * kBT * kBCv / kThreads;  // 8
  uint4 yraw[kIters];
  int offs[kIters];  // element offset inside the sample (fits 32 bits), -1 = outside the volume
  const long long sample = (long long)n * D * H * W * C;
#pragma unroll
  for (int it = 0; it < kIters; it++) {
    const int i = threadIdx.x + it * kThreads;
    const int v = i / kBCv;
    const int iw = i0w + (v % kBT), ih = i0h + ((v / kBT) % kBT), id = i0d + v / (kBT * kBT);
    const bool ok = iw < W && ih < H && id < D && cv_ok;
    offs[it] = ok ? ((id * H + ih) * W + iw) * C + (cv0 + cv) * 8 : -1;
    yraw[it] = ok ? __ldg(reinterpret_cast<const uint4*>(y + sample + offs[it])) : make_uint4(0, 0, 0, 0);
  }
  for (int i = threadIdx.x; i < kBSmemFloats / 4; i += kThreads) reinterpret_cast<float4*>(s_g)[i] = make_float4(0, 0, 0, 0);
  __syncthreads();
  // phase 1: scatter
  for (int i = threadIdx.x; i < nd * nh * nw * kBCv; i += kThreads) {
    const int wcv = i % kBCv, wdx = i / kBCv;
    if (cv0 + wcv >= vpr) continue;
    const int ww = wdx % nw, wh = (wdx / nw) % nh, wd = wdx / (nw * nh);
    const int od = od0 + wd, oh = oh0 + wh, ow = ow0 + ww;
    const long long o = ((((long long)n * Do + od) * Ho + oh) * Wo + ow) * vpr + cv0 + wcv;
    float gv[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dp + o * 8)), gv);
    const uint2 pk = __ldg(reinterpret_cast<const uint2*>(amax + o * 8));
    const int bd = od * kS - kPad - i0d, bh = oh * kS - kPad - i0h, bw = ow * kS - kPad - i0w;  // window origin in tile
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint32_t word = j < 4 ? pk.x : pk.y;
      const int slot = (int)((word >> ((j & 3) * 8)) & 255);
      const int kd = slot / (kK * kK), kh = (slot / kK) % kK, kw = slot % kK;
      const int vd = bd + kd, vh = bh + kh, vw = bw + kw;
      if (vd >= 0 && vd < kBT && vh >= 0 && vh < kBT && vw >= 0 && vw < kBT)
        atomicAdd(&s_g[(((vd * kBT + vh) * kBT + vw) * kBCv + wcv) * 8 + j], gv[j]);
    }
  }
  __syncthreads();
  // phase 2: stream y (all 8 loads of a thread are issued before the first use: memory-level parallelism)
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; j++) a0[j] = a1[j] = 0.f;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = s_c[2][cv * 8 + j];
    sh[j] = s_c[3][cv * 8 + j];
  }
#pragma unroll
  for (int it = 0; it < kIters; it++) {
    if (offs[it] < 0) continue;
    const int i = threadIdx.x + it * kThreads;
    float yv[8];
    unpack8(yraw[it], yv);
    const float4 g0 = reinterpret_cast<const float4*>(s_g)[i * 2], g1 = reinterpret_cast<const float4*>(s_g)[i * 2 + 1];
    float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < 8; j++) g[j] = fmaf(yv[j], sc[j], sh[j]) > 0.f ? g[j] : 0.f;  // ReLU mask
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        a0[j] += g[j];
        a1[j] = fmaf(g[j], (yv[j] - s_c[0][cv * 8 + j]) * s_c[1][cv * 8 + j], a1[j]);
      }
    } else {
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; j++)
        r[j] = fmaf(s_c[4][cv * 8 + j], g[j], fmaf(s_c[5][cv * 8 + j], yv[j], s_c[6][cv * 8 + j]));
      *reinterpret_cast<uint4*>(dy + sample + offs[it]) = pack8(r);
    }
  }
  if (MODE == 0) {
    const int grp = threadIdx.x / kBCv;  // threads sharing one channel vector
#pragma unroll
    for (int j = 0; j < 8; j++) {
      s_red[0][grp][cv * 8 + j] = a0[j];
      s_red[1][grp][cv * 8 + j] = a1[j];
    }
    __syncthreads();
    if (threadIdx.x < 2 * kBCv * 8) {
      const int which = threadIdx.x / (kBCv * 8), c = threadIdx.x % (kBCv * 8);
      float s = 0.f;
      for (int l = 0; l < kThreads / kBCv; l++) s += s_red[which][l][c];
      const int ch = cv0 * 8 + c;
      if (ch < C) atomicAdd(red_out + which * C + ch, (double)s);
    }
  }
}

int check_pool(int C, int k, int stride, int pad) {
  ADNI_REQUIRE(k == 3 && stride == 2 && pad == 1, ADNI_ENOTSUP,
               "fused stem pooling supports MaxPool3d(3, 2, 1) only (k=%d s=%d p=%d)", k, stride, pad);
  ADNI_REQUIRE(C % 8 == 0 && C >= 8, ADNI_ENOTSUP, "fused stem pooling: C=%d must be a multiple of 8", C);
  return ADNI_OK;
}

}  // namespace
}  // namespace adni

using namespace adni;
#define ST(s) static_cast<cudaStream_t>(s)
typedef __nv_bfloat16 bf16;

extern "C" {

int adni_bn_relu_maxpool_fwd(const adni_bf16* y, const float* scale, const float* shift, int N, int D, int H, int W,
                             int C, int k, int stride, int pad, adni_bf16* p, uint8_t* argmax, void* stream) {
  ADNI_REQUIRE(y && scale && shift && p && argmax && N > 0, ADNI_EINVAL, "bn_relu_maxpool_fwd: bad arguments");
  int rc = check_pool(C, k, stride, pad);
  if (rc) return rc;
  const int Do = (D + 2 - 3) / 2 + 1, Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int td = (Do + kFT - 1) / kFT, th = (Ho + kFT - 1) / kFT, tw = (Wo + kFT - 1) / kFT;
  dim3 grid((unsigned)((long long)N * td * th * tw * ((C / 8 + kFCv - 1) / kFCv)));
  bn_relu_pool_fwd_kernel<<<grid, kThreads, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(y), scale, shift, N, D, H, W,
                                                             C, Do, Ho, Wo, td, th, tw, reinterpret_cast<bf16*>(p),
                                                             argmax);
  count_launch();
  ADNI_LAUNCH_CHECK("bn_relu_pool_fwd_kernel");
  return ADNI_OK;
}

int adni_maxpool_bn_bwd_reduce(const adni_bf16* dp, const uint8_t* argmax, const adni_bf16* y, const float* bnp, int N,
                               int D, int H, int W, int C, int k, int stride, int pad, double* red, void* stream) {
  ADNI_REQUIRE(dp && argmax && y && bnp && red && N > 0, ADNI_EINVAL, "maxpool_bn_bwd_reduce: bad arguments");
  int rc = check_pool(C, k, stride, pad);
  if (rc) return rc;
  const int Do = (D + 2 - 3) / 2 + 1, Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int td = (D + kBT - 1) / kBT, th = (H + kBT - 1) / kBT, tw = (W + kBT - 1) / kBT;
  dim3 grid((unsigned)((long long)N * td * th * tw * ((C / 8 + kBCv - 1) / kBCv)));
  static bool attr0 = false;
  if (!attr0) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(pool_bn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kBSmemFloats * 4));
    attr0 = true;
  }
  pool_bn_bwd_kernel<0><<<grid, kThreads, kBSmemFloats * 4, ST(stream)>>>(reinterpret_cast<const bf16*>(dp), argmax,
                                                           reinterpret_cast<const bf16*>(y), bnp, nullptr, nullptr, 0.0,
                                                           N, D, H, W, C, Do, Ho, Wo, td, th, tw, red, nullptr);
  count_launch();
  ADNI_LAUNCH_CHECK("pool_bn_bwd_kernel<0>");
  return ADNI_OK;
}

int adni_maxpool_bn_bwd_apply(const adni_bf16* dp, const uint8_t* argmax, const adni_bf16* y, const float* bnp,
                              const float* gamma, const double* red, double count, int N, int D, int H, int W, int C,
                              int k, int stride, int pad, adni_bf16* dy, void* stream) {
  ADNI_REQUIRE(dp && argmax && y && bnp && red && dy && N > 0 && count > 0, ADNI_EINVAL,
               "maxpool_bn_bwd_apply: bad arguments");
  int rc = check_pool(C, k, stride, pad);
  if (rc) return rc;
  const int Do = (D + 2 - 3) / 2 + 1, Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const int td = (D + kBT - 1) / kBT, th = (H + kBT - 1) / kBT, tw = (W + kBT - 1) / kBT;
  dim3 grid((unsigned)((long long)N * td * th * tw * ((C / 8 + kBCv - 1) / kBCv)));
  static bool attr1 = false;
  if (!attr1) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(pool_bn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kBSmemFloats * 4));
    attr1 = true;
  }
  pool_bn_bwd_kernel<1><<<grid, kThreads, kBSmemFloats * 4, ST(stream)>>>(reinterpret_cast<const bf16*>(dp), argmax,
                                                           reinterpret_cast<const bf16*>(y), bnp, gamma, red,
                                                           1.0 / count, N, D, H, W, C, Do, Ho, Wo, td, th, tw, nullptr,
                                                           reinterpret_cast<bf16*>(dy));
  count_launch();
  ADNI_LAUNCH_CHECK("pool_bn_bwd_kernel<1>");
  return ADNI_OK;
}

}  // extern "C"

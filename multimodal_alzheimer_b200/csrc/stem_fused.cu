// Fused element-wise path of the ResNet stem: bn1 -> ReLU -> MaxPool3d(3,2,1) forward, and the matching backward
// (max-pool scatter + ReLU mask + BatchNorm backward) WITHOUT ever materialising the 64x64^3 activated tensor or
// its gradient (33.5 MB per volume each).  HBM traffic per volume: forward reads y once; backward reads y twice
// (reduction pass, apply pass) and writes dy once.
//   MedicalNet ResNet.forward: x = maxpool(relu(bn1(conv1(x))))   (call site pkg/models/mri_models/anat_cnn.py:95)
#include <stdlib.h>

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
  pdl_enter();
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
  pdl_enter();
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


// ================================================================================================
// Streaming variants (default).  No shared-memory staging, no block-wide phases: every thread keeps its loads in
// flight while it computes, and every 128-byte line of y is requested as a whole line by 8 neighbouring lanes.
// ================================================================================================

// Forward.  Thread = one pooled (oh, ow) column x 8 channels, marching along od.  Max-pooling is separable with the
// first-maximum tie rule intact: the in-plane winner is the first maximum in (kh, kw) order, planes are compared in
// kd order with a strict '>'.  The odd input plane 2*od+1 is shared by outputs od and od+1: its in-plane result is
// carried in registers, so every input plane is read (and BN+ReLU'd) once per thread.
constexpr int kSFh = 2, kSFw = 8, kSFThreads = kSFh * kSFw * 8;  // 2 x 8 pooled columns x 8 channel vectors

struct PlaneMax {
  float v[8];
  uint32_t idx;     // 8 nibbles: kh*3 + kw of the winner of every channel
  uint32_t raw[4];  // the winner's RAW conv output (bf16 pairs): BatchNorm backward needs xhat at the arg-max only
};
// lane (j & 1) of word `dst` <- the same lane of `src`
__device__ __forceinline__ uint32_t put_lane(uint32_t dst, uint32_t src, int j) {
  return (j & 1) ? ((dst & 0x0000FFFFu) | (src & 0xFFFF0000u)) : ((dst & 0xFFFF0000u) | (src & 0x0000FFFFu));
}

// The nine 16-byte vectors of one input plane under a thread's window.  Loading and reducing are separate steps so
// that the loads of the NEXT plane are in flight while the current one is reduced: with both in one routine a warp
// alternated between waiting for 9 loads and ~550 dependent ALU instructions, and at 4 resident blocks per SM the
// kernel ran at a quarter of the HBM bandwidth with neither the memory nor the issue slots busy.
struct PlaneRaw {
  uint4 r[9];
};

// okmask bit (kh*3 + kw): the tap lies inside the plane (a property of the thread's (oh, ow), the same for every plane)
__device__ __forceinline__ void plane_load(PlaneRaw& pr, const __nv_bfloat16* __restrict__ ytap0, long long row_stride,
                                           int C, uint32_t okmask) {
#pragma unroll
  for (int kh = 0; kh < 3; kh++) {
#pragma unroll
    for (int kw = 0; kw < 3; kw++) {
      const int t = kh * 3 + kw;
      pr.r[t] = (okmask >> t) & 1u ? __ldg(reinterpret_cast<const uint4*>(ytap0 + kh * row_stride + kw * C))
                                   : make_uint4(0, 0, 0, 0);
    }
  }
}

__device__ __forceinline__ PlaneMax plane_reduce(const PlaneRaw& pr, uint32_t okmask, const float (&sc)[8],
                                                 const float (&sh)[8]) {
  PlaneMax m;
#pragma unroll
  for (int j = 0; j < 8; j++) m.v[j] = -INFINITY;
  m.idx = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) m.raw[k] = 0u;
#pragma unroll
  for (int t = 0; t < 9; t++) {
    // branch-free: a tap outside the plane (zero vector, border threads only) is computed and then ignored through the
    // predicate of the compare - a per-tap branch cost ~20 reconvergence instructions per tap on every thread
    const bool ok = (okmask >> t) & 1u;
    float f[8];
    unpack8(pr.r[t], f);
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
    float r[8];
    unpack8(pack8(f), r);  // compare what bn_apply would have stored: the bf16-rounded activation
    const uint32_t rw[4] = {pr.r[t].x, pr.r[t].y, pr.r[t].z, pr.r[t].w};
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (ok && r[j] > m.v[j]) {  // first maximum in (kh, kw) order wins
        m.v[j] = r[j];
        m.idx = (m.idx & ~(0xFu << (4 * j))) | (static_cast<uint32_t>(t) << (4 * j));
        m.raw[j >> 1] = put_lane(m.raw[j >> 1], rw[j >> 1], j);
      }
    }
  }
  return m;
}

__global__ void __launch_bounds__(kSFThreads, 3)
    bn_relu_pool_fwd_stream_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                   const float* __restrict__ shift, int D, int H, int W, int C, int Do, int Ho, int Wo,
                                   int tiles_h, int tiles_w, int dsplit, __nv_bfloat16* __restrict__ p,
                                   uint8_t* __restrict__ amax, __nv_bfloat16* __restrict__ yraw) {
  pdl_enter();
  const int vpr = C / 8;
  __shared__ float s_sc[64], s_sh[64];  // the block's 64 channels
  if (threadIdx.x < 64) {
    const int c = min(blockIdx.y * 64 + (int)threadIdx.x, C - 1);
    s_sc[threadIdx.x] = scale[c];
    s_sh[threadIdx.x] = shift[c];
  }
  __syncthreads();
  const int cv = blockIdx.y * 8 + (threadIdx.x & 7);
  const int lw = (threadIdx.x >> 3) % kSFw, lh = threadIdx.x / (8 * kSFw);
  int t = blockIdx.x;
  const int ds = t % dsplit;
  t /= dsplit;
  const int tw = t % tiles_w;
  t /= tiles_w;
  const int th = t % tiles_h;
  const int n = t / tiles_h;
  const int oh = th * kSFh + lh, ow = tw * kSFw + lw;
  if (oh >= Ho || ow >= Wo || cv >= vpr) return;
  const int per = (Do + dsplit - 1) / dsplit;
  const int od_begin = ds * per, od_end = min(Do, od_begin + per);
  if (od_begin >= od_end) return;
  const int coff = cv * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = s_sc[(threadIdx.x & 7) * 8 + j];
    sh[j] = s_sh[(threadIdx.x & 7) * 8 + j];
  }
  uint32_t okmask = 0;
#pragma unroll
  for (int kh = 0; kh < 3; kh++) {
#pragma unroll
    for (int kw = 0; kw < 3; kw++) {
      const int ih = oh * kS - kPad + kh, iw = ow * kS - kPad + kw;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) okmask |= 1u << (kh * 3 + kw);
    }
  }
  const long long row_stride = (long long)W * C;
  const long long plane = (long long)H * row_stride;
  // address of tap (0, 0) in plane 0 (possibly outside the tensor: masked taps are never dereferenced)
  const __nv_bfloat16* ys =
      y + (long long)n * D * plane + ((long long)(oh * kS - kPad) * W + (ow * kS - kPad)) * C + coff;

  PlaneRaw bufA, bufB;  // A: even planes 2*od (window centre), B: odd planes 2*od +- 1 (shared by two outputs)
  PlaneMax carry;
  bool have_carry = od_begin * kS - kPad >= 0;  // the plane below the first output of this d-range
  if (have_carry) plane_load(bufB, ys + (long long)(od_begin * kS - kPad) * plane, row_stride, C, okmask);
  plane_load(bufA, ys + (long long)(od_begin * kS) * plane, row_stride, C, okmask);
  if (have_carry) carry = plane_reduce(bufB, okmask, sc, sh);
  for (int od = od_begin; od < od_end; od++) {
    const bool has_odd = od * kS + 1 < D;
    if (has_odd) plane_load(bufB, ys + (long long)(od * kS + 1) * plane, row_stride, C, okmask);
    float best[8];
    uint32_t bidx[8], braw[4];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      best[j] = have_carry ? carry.v[j] : -INFINITY;
      bidx[j] = have_carry ? ((carry.idx >> (4 * j)) & 0xFu) : 0xFFu;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) braw[k] = have_carry ? carry.raw[k] : 0u;
    const PlaneMax mid = plane_reduce(bufA, okmask, sc, sh);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (mid.v[j] > best[j]) {
        best[j] = mid.v[j];
        bidx[j] = 9u + ((mid.idx >> (4 * j)) & 0xFu);
        braw[j >> 1] = put_lane(braw[j >> 1], mid.raw[j >> 1], j);
      }
    }
    if (od + 1 < od_end) plane_load(bufA, ys + (long long)((od + 1) * kS) * plane, row_stride, C, okmask);
    have_carry = has_odd;
    if (have_carry) {
      carry = plane_reduce(bufB, okmask, sc, sh);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (carry.v[j] > best[j]) {
          best[j] = carry.v[j];
          bidx[j] = 18u + ((carry.idx >> (4 * j)) & 0xFu);
          braw[j >> 1] = put_lane(braw[j >> 1], carry.raw[j >> 1], j);
        }
      }
    }
    const long long oi = ((((long long)n * Do + od) * Ho + oh) * Wo + ow) * vpr + cv;
    *reinterpret_cast<uint4*>(p + oi * 8) = pack8(best);
    if (yraw != nullptr) *reinterpret_cast<uint4*>(yraw + oi * 8) = make_uint4(braw[0], braw[1], braw[2], braw[3]);
    uint2 pk;
    pk.x = (bidx[0] & 255u) | ((bidx[1] & 255u) << 8) | ((bidx[2] & 255u) << 16) | ((bidx[3] & 255u) << 24);
    pk.y = (bidx[4] & 255u) | ((bidx[5] & 255u) << 8) | ((bidx[6] & 255u) << 16) | ((bidx[7] & 255u) << 24);
    *reinterpret_cast<uint2*>(amax + oi * 8) = pk;
  }
}

// ------------------------------------------------------------------------------------------------
// EXPERIMENTAL packed-compare variant of the streaming forward (ADNI_POOL_STREAM=2; same results bit for bit, meant
// to be validated by the existing tests with that setting before it becomes the default).  The forward kernel is
// ALU-issue bound: after BN + ReLU + rounding the default path unpacks the rounded activation again and spends a
// compare + two selects per (tap, channel).  Here the running maximum and its tap index stay in the packed bf16x2
// domain - one HSET2 mask and two LOP3 selects per tap and channel PAIR; the winner's bits are stored as they are.
struct PlaneMaxP {
  uint32_t v[4];    // running maxima, two bf16 lanes per word (channels 2k, 2k + 1)
  uint32_t idx[4];  // tap index kh*3 + kw of the winner in each 16-bit lane
  uint32_t raw[4];  // the winner's raw conv output
};
__device__ __forceinline__ uint32_t gt2_mask(uint32_t a, uint32_t b) {
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
}
constexpr uint32_t kNegInf2 = 0xFF80FF80u;  // (-inf, -inf) in bf16

__device__ __forceinline__ PlaneMaxP plane_reduce_packed(const PlaneRaw& pr, uint32_t okmask, const float (&sc)[8],
                                                         const float (&sh)[8]) {
  PlaneMaxP m;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    m.v[k] = kNegInf2;
    m.idx[k] = 0;
    m.raw[k] = 0;
  }
#pragma unroll
  for (int t = 0; t < 9; t++) {
    const uint32_t okm = ((okmask >> t) & 1u) ? 0xFFFFFFFFu : 0u;  // a tap outside the plane never wins
    float f[8];
    unpack8(pr.r[t], f);
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
    const uint4 yb = pack8(f);  // what bn_apply would have stored: the bf16-rounded activation
    const uint32_t w[4] = {yb.x, yb.y, yb.z, yb.w};
    const uint32_t rw[4] = {pr.r[t].x, pr.r[t].y, pr.r[t].z, pr.r[t].w};
    const uint32_t tt = static_cast<uint32_t>(t) * 0x00010001u;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t g = gt2_mask(w[k], m.v[k]) & okm;  // strict '>': the first maximum in (kh, kw) order wins
      m.v[k] = (w[k] & g) | (m.v[k] & ~g);
      m.idx[k] = (tt & g) | (m.idx[k] & ~g);
      m.raw[k] = (rw[k] & g) | (m.raw[k] & ~g);
    }
  }
  return m;
}

__global__ void __launch_bounds__(kSFThreads, 3)
    bn_relu_pool_fwd_stream_packed_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                          const float* __restrict__ shift, int D, int H, int W, int C, int Do, int Ho,
                                          int Wo, int tiles_h, int tiles_w, int dsplit, __nv_bfloat16* __restrict__ p,
                                          uint8_t* __restrict__ amax, __nv_bfloat16* __restrict__ yraw) {
  pdl_enter();
  const int vpr = C / 8;
  __shared__ float s_sc[64], s_sh[64];
  if (threadIdx.x < 64) {
    const int c = min(blockIdx.y * 64 + (int)threadIdx.x, C - 1);
    s_sc[threadIdx.x] = scale[c];
    s_sh[threadIdx.x] = shift[c];
  }
  __syncthreads();
  const int cv = blockIdx.y * 8 + (threadIdx.x & 7);
  const int lw = (threadIdx.x >> 3) % kSFw, lh = threadIdx.x / (8 * kSFw);
  int t = blockIdx.x;
  const int ds = t % dsplit;
  t /= dsplit;
  const int tw = t % tiles_w;
  t /= tiles_w;
  const int th = t % tiles_h;
  const int n = t / tiles_h;
  const int oh = th * kSFh + lh, ow = tw * kSFw + lw;
  if (oh >= Ho || ow >= Wo || cv >= vpr) return;
  const int per = (Do + dsplit - 1) / dsplit;
  const int od_begin = ds * per, od_end = min(Do, od_begin + per);
  if (od_begin >= od_end) return;
  const int coff = cv * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    sc[j] = s_sc[(threadIdx.x & 7) * 8 + j];
    sh[j] = s_sh[(threadIdx.x & 7) * 8 + j];
  }
  uint32_t okmask = 0;
#pragma unroll
  for (int kh = 0; kh < 3; kh++) {
#pragma unroll
    for (int kw = 0; kw < 3; kw++) {
      const int ih = oh * kS - kPad + kh, iw = ow * kS - kPad + kw;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) okmask |= 1u << (kh * 3 + kw);
    }
  }
  const long long row_stride = (long long)W * C;
  const long long plane = (long long)H * row_stride;
  const __nv_bfloat16* ys =
      y + (long long)n * D * plane + ((long long)(oh * kS - kPad) * W + (ow * kS - kPad)) * C + coff;

  PlaneRaw bufA, bufB;
  PlaneMaxP carry;
  bool have_carry = od_begin * kS - kPad >= 0;
  if (have_carry) plane_load(bufB, ys + (long long)(od_begin * kS - kPad) * plane, row_stride, C, okmask);
  plane_load(bufA, ys + (long long)(od_begin * kS) * plane, row_stride, C, okmask);
  if (have_carry) carry = plane_reduce_packed(bufB, okmask, sc, sh);
  for (int od = od_begin; od < od_end; od++) {
    const bool has_odd = od * kS + 1 < D;
    if (has_odd) plane_load(bufB, ys + (long long)(od * kS + 1) * plane, row_stride, C, okmask);
    uint32_t best[4], bidx[4], braw[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      best[k] = have_carry ? carry.v[k] : kNegInf2;
      bidx[k] = have_carry ? carry.idx[k] : 0x00FF00FFu;
      braw[k] = have_carry ? carry.raw[k] : 0u;
    }
    const PlaneMaxP mid = plane_reduce_packed(bufA, okmask, sc, sh);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t g = gt2_mask(mid.v[k], best[k]);  // planes in kd order, strict '>'
      best[k] = (mid.v[k] & g) | (best[k] & ~g);
      bidx[k] = ((mid.idx[k] + 9u * 0x00010001u) & g) | (bidx[k] & ~g);
      braw[k] = (mid.raw[k] & g) | (braw[k] & ~g);
    }
    if (od + 1 < od_end) plane_load(bufA, ys + (long long)((od + 1) * kS) * plane, row_stride, C, okmask);
    have_carry = has_odd;
    if (have_carry) {
      carry = plane_reduce_packed(bufB, okmask, sc, sh);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t g = gt2_mask(carry.v[k], best[k]);
        best[k] = (carry.v[k] & g) | (best[k] & ~g);
        bidx[k] = ((carry.idx[k] + 18u * 0x00010001u) & g) | (bidx[k] & ~g);
        braw[k] = (carry.raw[k] & g) | (braw[k] & ~g);
      }
    }
    const long long oi = ((((long long)n * Do + od) * Ho + oh) * Wo + ow) * vpr + cv;
    *reinterpret_cast<uint4*>(p + oi * 8) = make_uint4(best[0], best[1], best[2], best[3]);
    if (yraw != nullptr) *reinterpret_cast<uint4*>(yraw + oi * 8) = make_uint4(braw[0], braw[1], braw[2], braw[3]);
    uint2 pk;  // bytes 0 and 2 of every index word: channels (0,1,2,3) and (4,5,6,7)
    pk.x = __byte_perm(bidx[0], bidx[1], 0x6420);
    pk.y = __byte_perm(bidx[2], bidx[3], 0x6420);
    *reinterpret_cast<uint2*>(amax + oi * 8) = pk;
  }
}

// Backward as a GATHER over 2x2x2 input cells.  Thread = the 8 voxels (2cd+ed, 2ch+eh, 2cw+ew), e in {0,1}, of one
// cell x 8 channels.  Along every axis a voxel with an even coordinate is the centre tap of window c, an odd one the
// last tap of window c and the first tap of window c+1, so the cell only ever receives gradient from the 2x2x2
// windows (c + a), a in {0,1}: the thread loads those 8 (dp, arg-max) vectors once, and every voxel adds dp where the
// recorded arg-max slot is its own tap (fixed order: deterministic, no atomics, no shared-memory staging).  Then the
// ReLU mask recomputed from y and
//   MODE 0: sum g, sum g*xhat per channel (block reduction, fp64 atomics per block);
//   MODE 1: dy = A*g + B*y + K.
constexpr int kGThreads = 128;

__device__ __forceinline__ constexpr int pool_tap(int e, int a) { return e == 0 ? 1 : (a == 0 ? 2 : 0); }

template <int MODE>
__global__ void __launch_bounds__(kGThreads, 3)
    pool_bn_bwd_cell_kernel(const __nv_bfloat16* __restrict__ dp, const uint8_t* __restrict__ amax,
                            const __nv_bfloat16* __restrict__ y, const float* __restrict__ bnp /* [4][C] */,
                            const float* __restrict__ gamma, const double* __restrict__ red_in, double inv_count, int N,
                            int D, int H, int W, int C, int Do, int Ho, int Wo, double* __restrict__ red_out,
                            __nv_bfloat16* __restrict__ dy) {
  pdl_enter();
  const int vpr = C / 8;  // power of two <= 32 (checked by the launcher): a thread's channel vector never changes
  const int cv = threadIdx.x & (vpr - 1);
  const int coff = cv * 8;
  // per-channel constants live in shared memory (7 x 8 registers per thread otherwise): mu, is, sc, sh, A, B, K
  __shared__ float s_c[7][256];
  for (int c = threadIdx.x; c < C; c += kGThreads) {
    const float mu = bnp[c], is = bnp[C + c];
    s_c[0][c] = mu;
    s_c[1][c] = is;
    s_c[2][c] = bnp[2 * C + c];
    s_c[3][c] = bnp[3 * C + c];
    if (MODE == 1) {
      const float gm = gamma ? gamma[c] : 1.f;
      const float mg = (float)(red_in[c] * inv_count), mgx = (float)(red_in[C + c] * inv_count);
      const float a = gm * is, b = -a * is * mgx;
      s_c[4][c] = a;
      s_c[5][c] = b;
      s_c[6][c] = -a * mg - b * mu;
    }
  }
  __syncthreads();
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; j++) a0[j] = a1[j] = 0.f;

  const int Cd = (D + 1) / 2, Ch = (H + 1) / 2, Cw = (W + 1) / 2;
  const long long items = (long long)N * Cd * Ch * Cw * vpr;
  const long long stride = (long long)gridDim.x * kGThreads;
  const long long sW = C, sH = (long long)W * C, sD = (long long)H * W * C;
  for (long long it = (long long)blockIdx.x * kGThreads + threadIdx.x; it < items; it += stride) {
    long long r = it / vpr;
    const int cw = (int)(r % Cw);
    r /= Cw;
    const int ch = (int)(r % Ch);
    r /= Ch;
    const int cd = (int)(r % Cd);
    const int n = (int)(r / Cd);
    const bool vd1 = 2 * cd + 1 < D, vh1 = 2 * ch + 1 < H, vw1 = 2 * cw + 1 < W;
    const long long ybase = ((((long long)n * D + 2 * cd) * H + 2 * ch) * W + 2 * cw) * C + coff;
    // all loads of the item are issued before the first use: 8 y vectors, 8 windows
    uint4 yr[8];
#pragma unroll
    for (int v = 0; v < 8; v++) {
      const int ed = v >> 2, eh = (v >> 1) & 1, ew = v & 1;
      const bool ok = (!ed || vd1) && (!eh || vh1) && (!ew || vw1);
      yr[v] = ok ? __ldg(reinterpret_cast<const uint4*>(y + ybase + ed * sD + eh * sH + ew * sW)) : make_uint4(0, 0, 0, 0);
    }
    uint4 wd[8];
    uint2 wm[8];
#pragma unroll
    for (int w = 0; w < 8; w++) {
      const int ad = w >> 2, ah = (w >> 1) & 1, aw = w & 1;
      const bool ok = cd + ad < Do && ch + ah < Ho && cw + aw < Wo;
      const long long o = ((((long long)n * Do + cd + ad) * Ho + ch + ah) * Wo + cw + aw) * vpr + cv;
      wd[w] = ok ? __ldg(reinterpret_cast<const uint4*>(dp + o * 8)) : make_uint4(0, 0, 0, 0);
      wm[w] = ok ? __ldg(reinterpret_cast<const uint2*>(amax + o * 8)) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    }
#pragma unroll
    for (int v = 0; v < 8; v++) {
      const int ed = v >> 2, eh = (v >> 1) & 1, ew = v & 1;
      const bool ok = (!ed || vd1) && (!eh || vh1) && (!ew || vw1);
      if (!ok) continue;  // warp-uniform except at the ragged W edge
      float g[8];
#pragma unroll
      for (int j = 0; j < 8; j++) g[j] = 0.f;
#pragma unroll
      for (int w = 0; w < 8; w++) {
        const int ad = w >> 2, ah = (w >> 1) & 1, aw = w & 1;
        if (ad > ed || ah > eh || aw > ew) continue;  // compile-time: an even coordinate only sees window c
        const uint32_t slot = static_cast<uint32_t>((pool_tap(ed, ad) * 3 + pool_tap(eh, ah)) * 3 + pool_tap(ew, aw));
        // a zero byte of x = the arg-max slot of that channel is this voxel's tap (no SIMD byte compare: its
        // emulation cost more than the eight tests it feeds)
        const uint32_t x_lo = wm[w].x ^ (slot * 0x01010101u), x_hi = wm[w].y ^ (slot * 0x01010101u);
        float f[8];
        unpack8(wd[w], f);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (((j < 4 ? x_lo : x_hi) & (0xFFu << ((j & 3) * 8))) == 0u) g[j] += f[j];
        }
      }
      float yv[8];
      unpack8(yr[v], yv);
#pragma unroll
      for (int j = 0; j < 8; j++) g[j] = fmaf(yv[j], s_c[2][coff + j], s_c[3][coff + j]) > 0.f ? g[j] : 0.f;  // ReLU mask
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          a0[j] += g[j];
          a1[j] = fmaf(g[j], (yv[j] - s_c[0][coff + j]) * s_c[1][coff + j], a1[j]);
        }
      } else {
        float rr[8];
#pragma unroll
        for (int j = 0; j < 8; j++) rr[j] = fmaf(s_c[4][coff + j], g[j], fmaf(s_c[5][coff + j], yv[j], s_c[6][coff + j]));
        *reinterpret_cast<uint4*>(dy + ybase + ed * sD + eh * sH + ew * sW) = pack8(rr);
      }
    }
  }
  if (MODE == 0) {
    __shared__ float s_part[2][kGThreads / 32][32 * 8 + 8];
    // lanes with the same (lane mod vpr) hold the same channels: fold them with shuffles, then across warps in smem
#pragma unroll
    for (int j = 0; j < 8; j++) {
      for (int o = 16; o >= vpr; o >>= 1) {
        a0[j] += __shfl_xor_sync(0xffffffffu, a0[j], o);
        a1[j] += __shfl_xor_sync(0xffffffffu, a1[j], o);
      }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < vpr) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        s_part[0][warp][lane * 8 + j] = a0[j];
        s_part[1][warp][lane * 8 + j] = a1[j];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kGThreads) {
      const int which = i / C, c = i % C;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kGThreads / 32; w++) s += s_part[which][w][c];
      atomicAdd(red_out + which * C + c, (double)s);
    }
  }
}

int pool_variant() {
  // 2 (default since round 2): streaming forward with the running maximum kept in the packed bf16x2 domain - bit-identical
  // to variant 1 and measured 1.76 -> 1.07 ms per step (two launches of 32 volumes, profiles/r02_pool_ab.md)
  const char* e = getenv("ADNI_POOL_STREAM");
  return e ? atoi(e) : 2;
}
bool pow2_le32(int v) { return v >= 1 && v <= 32 && (v & (v - 1)) == 0; }

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
                             int C, int k, int stride, int pad, adni_bf16* p, uint8_t* argmax, adni_bf16* y_at_argmax,
                             void* stream) {
  ADNI_REQUIRE(y && scale && shift && p && argmax && N > 0, ADNI_EINVAL, "bn_relu_maxpool_fwd: bad arguments");
  int rc = check_pool(C, k, stride, pad);
  if (rc) return rc;
  const int Do = (D + 2 - 3) / 2 + 1, Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  if (pool_variant()) {
    const int th = (Ho + kSFh - 1) / kSFh, tw = (Wo + kSFw - 1) / kSFw;
    const int dsplit = Do >= 16 ? 2 : 1;
    dim3 grid((unsigned)((long long)N * th * tw * dsplit), (unsigned)((C / 8 + 7) / 8));
    if (pool_variant() == 2) {  // experimental packed-compare variant, same results
      pdl_launch(bn_relu_pool_fwd_stream_packed_kernel, grid, kSFThreads, 0, ST(stream))(
          reinterpret_cast<const bf16*>(y), scale, shift, D, H, W, C, Do, Ho, Wo, th, tw, dsplit,
          reinterpret_cast<bf16*>(p), argmax, reinterpret_cast<bf16*>(y_at_argmax));
      count_launch();
      ADNI_LAUNCH_CHECK("bn_relu_pool_fwd_stream_packed_kernel");
      return ADNI_OK;
    }
    pdl_launch(bn_relu_pool_fwd_stream_kernel, grid, kSFThreads, 0, ST(stream))(reinterpret_cast<const bf16*>(y), scale, shift,
                                                                        D, H, W, C, Do, Ho, Wo, th, tw, dsplit,
                                                                        reinterpret_cast<bf16*>(p), argmax,
                                                                        reinterpret_cast<bf16*>(y_at_argmax));
    count_launch();
    ADNI_LAUNCH_CHECK("bn_relu_pool_fwd_stream_kernel");
    return ADNI_OK;
  }
  ADNI_REQUIRE(y_at_argmax == nullptr, ADNI_ENOTSUP, "bn_relu_maxpool_fwd: y_at_argmax needs a streaming variant (ADNI_POOL_STREAM != 0)");
  const int td = (Do + kFT - 1) / kFT, th = (Ho + kFT - 1) / kFT, tw = (Wo + kFT - 1) / kFT;
  dim3 grid((unsigned)((long long)N * td * th * tw * ((C / 8 + kFCv - 1) / kFCv)));
  pdl_launch(bn_relu_pool_fwd_kernel, grid, kThreads, 0, ST(stream))(reinterpret_cast<const bf16*>(y), scale, shift, N, D, H, W,
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
  if (pool_variant() && pow2_le32(C / 8)) {
    pdl_launch(pool_bn_bwd_cell_kernel<0>, num_sms() * 3, kGThreads, 0, ST(stream))(
        reinterpret_cast<const bf16*>(dp), argmax, reinterpret_cast<const bf16*>(y), bnp, nullptr, nullptr, 0.0, N, D, H,
        W, C, Do, Ho, Wo, red, nullptr);
    count_launch();
    ADNI_LAUNCH_CHECK("pool_bn_bwd_cell_kernel<0>");
    return ADNI_OK;
  }
  const int td = (D + kBT - 1) / kBT, th = (H + kBT - 1) / kBT, tw = (W + kBT - 1) / kBT;
  dim3 grid((unsigned)((long long)N * td * th * tw * ((C / 8 + kBCv - 1) / kBCv)));
  static bool attr0 = false;
  if (!attr0) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(pool_bn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kBSmemFloats * 4));
    attr0 = true;
  }
  pdl_launch(pool_bn_bwd_kernel<0>, grid, kThreads, kBSmemFloats * 4, ST(stream))(reinterpret_cast<const bf16*>(dp), argmax,
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
  if (pool_variant() && pow2_le32(C / 8)) {
    pdl_launch(pool_bn_bwd_cell_kernel<1>, num_sms() * 3, kGThreads, 0, ST(stream))(
        reinterpret_cast<const bf16*>(dp), argmax, reinterpret_cast<const bf16*>(y), bnp, gamma, red, 1.0 / count, N, D,
        H, W, C, Do, Ho, Wo, nullptr, reinterpret_cast<bf16*>(dy));
    count_launch();
    ADNI_LAUNCH_CHECK("pool_bn_bwd_cell_kernel<1>");
    return ADNI_OK;
  }
  const int td = (D + kBT - 1) / kBT, th = (H + kBT - 1) / kBT, tw = (W + kBT - 1) / kBT;
  dim3 grid((unsigned)((long long)N * td * th * tw * ((C / 8 + kBCv - 1) / kBCv)));
  static bool attr1 = false;
  if (!attr1) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(pool_bn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kBSmemFloats * 4));
    attr1 = true;
  }
  pdl_launch(pool_bn_bwd_kernel<1>, grid, kThreads, kBSmemFloats * 4, ST(stream))(reinterpret_cast<const bf16*>(dp), argmax,
                                                           reinterpret_cast<const bf16*>(y), bnp, gamma, red,
                                                           1.0 / count, N, D, H, W, C, Do, Ho, Wo, td, th, tw, nullptr,
                                                           reinterpret_cast<bf16*>(dy));
  count_launch();
  ADNI_LAUNCH_CHECK("pool_bn_bwd_kernel<1>");
  return ADNI_OK;
}

}  // extern "C"

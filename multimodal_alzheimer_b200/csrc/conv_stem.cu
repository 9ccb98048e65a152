// Tensor-core path for the 1-channel ResNet stem  conv1 = Conv3d(1, 64, k=7, stride=2, pad=3, bias=False)
// (MedicalNet ResNet.__init__; reference call site pkg/models/mri_models/anat_cnn.py:18-31).
//
// A 1-channel NDHWC tensor cannot be fed to TMA directly (inner extent 2 B < 16 B), and K = 343 per output
// voxel is far too small per tap.  Trick: the W-axis filter window becomes the "channel" dimension.  A tiny
// expansion kernel writes X8[n][d][h][w'][j] = x[n][d][h][2w'-3+j] (j = 0..7, zero outside) — 16-byte pixels
// TMA can fetch.  For every (kd, kh) tap the A operand of the implicit GEMM is then one tiled-TMA box of X8
// (through 4 (d,h)-parity views with doubled strides), landing in shared memory as 128 rows x 16 B: exactly the
// no-swizzle K-major UMMA core-matrix layout.  Two taps form one K=16 tcgen05.mma; the 64x(56x8) weight matrix
// stays resident in shared memory for the life of the CTA.
//   fprop : D[128 positions][64 cout]  = sum_{kd,kh} X8box(kd,kh)[128][8] * W[(kd,kh)][64][8]^T
//   wgrad : D[(kd,kh,j)][64 cout]     += sum_pos X8box(kd,kh)[pos][j] * dY[pos][cout]   (A, B both MN-major)
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace adni {
extern void count_launch();
namespace {

constexpr int kStemThreads = 192;
constexpr int kK = 7;        // filter size
constexpr int kKP = 8;       // filter row padded to 8 taps / 8 window elements
constexpr int kCout = 64;
constexpr int kTapsP = kK * kKP;  // 56 (kd, kh padded) taps

struct StemParams {
  CUtensorMap x_maps[4];  // (pd, ph) parity views of X8: dims (8, W', H2, D2, N)
  CUtensorMap dy_map;     // wgrad only: (64, Wo, Ho, Do, N) box (64, bw, bh, bd, 1), 128-B swizzle
  CUtensorMap x_plane;    // plane-resident fprop: (W'*8, H, D, N) view of X8, box (64, 38, 1, 1)
  int D, H;               // input extents (plane-resident fprop)
  int total;              // plane-resident fprop: N * tiles_h * tiles_w * Do plane pieces
  int debug;              // diagnostics (ADNI_STEM_DEBUG bits): 1 no BN sums, 2 no stores, 4 no MMA issue
  int ext_d[2], ext_h[2];
  const __nv_bfloat16* w2g;  // [56][64][8] bf16, zero padded
  int N, Do, Ho, Wo;
  int bd, bh, bw;
  int tiles_d, tiles_h, tiles_w;
  __nv_bfloat16* out;  // [N][Do][Ho][Wo][64]
  double* stat_sum;
  double* stat_sq;
  float* dw;  // wgrad: [64][343] fp32 (+=)
};

struct AxisOff {
  int par, off;
};
__device__ __forceinline__ AxisOff axis_off(int kk) {  // input coord = 2*o + kk - 3  ->  parity view, o + off
  const int o = kk - 3;
  const int par = o & 1;
  return {par, (o - par) / 2};
}

// ------------------------------------------------------------------------------------------------
__global__ void stem_expand_kernel(const __nv_bfloat16* __restrict__ x, int N, int D, int H, int W, int Wo,
                                   __nv_bfloat16* __restrict__ x8) {
  pdl_enter();
  const long long total = (long long)N * D * H * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(i % Wo);
    const long long row = i / Wo;  // (n, d, h)
    const __nv_bfloat16* xr = x + row * W;
    const int w0 = 2 * wo - 3;
    uint32_t pk[4];
#pragma unroll
    for (int j2 = 0; j2 < 4; j2++) {
      const int wa = w0 + 2 * j2, wb = wa + 1;
      const uint16_t a = (wa >= 0 && wa < W) ? __bfloat16_as_ushort(xr[wa]) : (uint16_t)0;
      const uint16_t b = (wb >= 0 && wb < W && (2 * j2 + 1) < kK) ? __bfloat16_as_ushort(xr[wb]) : (uint16_t)0;
      pk[j2] = (uint32_t)a | ((uint32_t)b << 16);
    }
    *reinterpret_cast<uint4*>(x8 + i * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// w_ncdhw fp32 [64][1][7][7][7] -> w2g bf16 [kd*8+kh][co][8] (zero padded kh = 7 and j = 7)
__global__ void stem_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w2g) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kTapsP * kCout * 8) return;
  const int j = i & 7, co = (i >> 3) % kCout, t = i / (8 * kCout);
  const int kd = t / kKP, kh = t % kKP;
  float v = 0.f;
  if (kh < kK && j < kK) v = w[((co * kK + kd) * kK + kh) * kK + j];
  w2g[i] = __float2bfloat16_rn(v);
}

// dw2 fp32 [(kd*8+kh)*8+j][64] -> grad fp32 [64][7][7][7]
__global__ void stem_wgrad_unpack_kernel(const float* __restrict__ dw2, float* __restrict__ grad) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kCout * kK * kK * kK) return;
  const int kw = i % kK, kh = (i / kK) % kK, kd = (i / (kK * kK)) % kK, co = i / (kK * kK * kK);
  grad[i] = dw2[(((kd * kKP + kh) * 8) + kw) * kCout + co];
}

// ------------------------------------------------------------------------------------------------
struct StemTile {
  int n, d0, h0, w0;
};
__device__ __forceinline__ StemTile stem_decode(const StemParams& p, int tile) {
  StemTile c;
  const int tw = tile % p.tiles_w;
  tile /= p.tiles_w;
  const int th = tile % p.tiles_h;
  tile /= p.tiles_h;
  const int td = tile % p.tiles_d;
  c.n = tile / p.tiles_d;
  c.d0 = td * p.bd;
  c.h0 = th * p.bh;
  c.w0 = tw * p.bw;
  return c;
}
__device__ __forceinline__ bool stem_kd_valid(const StemParams& p, const StemTile& c, int kd) {
  const AxisOff a = axis_off(kd);
  const int d = c.d0 + a.off;
  return d + p.bd > 0 && d < p.ext_d[a.par];
}

constexpr int kFStages = 6;
constexpr int kFTapBytes = 128 * 16;             // one tap: 128 positions x 16 B
constexpr int kFStageBytes = kKP * kFTapBytes;   // one kd row: 8 taps
constexpr int kWBytes = kTapsP * kCout * 16;     // resident weights: 57344 B
constexpr int kFBarOff = kWBytes + kFStages * kFStageBytes;
constexpr int kFStatOff = kFBarOff + 256;
constexpr int kFSmem = kFStatOff + 4 * 2 * kCout * 4 + 1024;

__global__ void __launch_bounds__(kStemThreads, 1) stem_fprop_kernel(const __grid_constant__ StemParams p) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + kWBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kFBarOff);
  uint64_t* empty = full + kFStages;
  uint64_t* tfull = empty + kFStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* stat_smem = reinterpret_cast<float*>(smem + kFStatOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.N * p.tiles_d * p.tiles_h * p.tiles_w;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kFStages; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  // resident weights: [tap][co][8] is already the no-swizzle K-major core-matrix order (8 rows x 16 B)
  for (int i = threadIdx.x; i < kWBytes / 16; i += kStemThreads)
    reinterpret_cast<uint4*>(smem_w)[i] = __ldg(reinterpret_cast<const uint4*>(p.w2g) + i);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // lanes 0..7 each issue one (kd, kh) tap box of the stage
    const uint32_t tap_bytes = static_cast<uint32_t>(p.bw * p.bh * p.bd) * 16u;
    const int kh = lane & 7;
    const AxisOff ah = axis_off(kh < kK ? kh : kK - 1);  // padded tap: finite duplicate data, zero weights
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const StemTile c = stem_decode(p, tile);
      for (int kd = 0; kd < kK; kd++) {
        if (!stem_kd_valid(p, c, kd)) continue;
        const AxisOff ad = axis_off(kd);
        if (lane == 0) {
          mbar_wait(&empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&full[st], tap_bytes * kKP);
        }
        __syncwarp();
        if (lane < kKP)
          tma_load_5d(smem_a + st * kFStageBytes + kh * kFTapBytes, &p.x_maps[ad.par * 2 + ah.par], &full[st], 0, c.w0,
                      c.h0 + ah.off, c.d0 + ad.off, c.n);
        if (++st == kFStages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kCout, false, false);
      const uint32_t w_addr = smem_u32(smem_w);
      int st = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t accph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const StemTile c = stem_decode(p, tile);
        mbar_wait(&tempty[acc], accph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kCout);
        uint32_t accum = 0;
        for (int kd = 0; kd < kK; kd++) {
          if (!stem_kd_valid(p, c, kd)) continue;
          mbar_wait(&full[st], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + st * kFStageBytes);
#pragma unroll
          for (int i = 0; i < kKP / 2; i++) {
            // K = 16 = taps (2i, 2i+1): K-direction core-matrix stride (LBO) = one tap block; M/N-direction
            // core-matrix stride (SBO) = 128 B (8 rows x 16 B)
            const uint64_t adesc = umma_smem_desc_nosw(a_addr + (2 * i) * kFTapBytes, kFTapBytes, 128);
            const uint64_t bdesc = umma_smem_desc_nosw(w_addr + (kd * kKP + 2 * i) * (kCout * 16), kCout * 16, 128);
            umma_bf16(d_tmem, adesc, bdesc, idesc, accum | static_cast<uint32_t>(i));
          }
          accum = 1;
          umma_commit(&empty[st]);
          if (++st == kFStages) {
            st = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
        if (++acc == 2) {
          acc = 0;
          accph ^= 1;
        }
      }
    }
  } else {
    const int q = warp & 3, ew = warp - 2, et = threadIdx.x - 64;
    const int row = q * 32 + lane;
    const bool do_stats = p.stat_sum != nullptr;
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const StemTile c = stem_decode(p, tile);
      bool has_k = false;
      for (int kd = 0; kd < kK; kd++) has_k |= stem_kd_valid(p, c, kd);
      const int rw = row % p.bw, rh = (row / p.bw) % p.bh, rd = row / (p.bw * p.bh);
      const int od = c.d0 + rd, oh = c.h0 + rh, ow = c.w0 + rw;
      const bool valid = rd < p.bd && od < p.Do && oh < p.Ho && ow < p.Wo;
      const long long off = ((((long long)c.n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * kCout;
      mbar_wait(&tfull[acc], accph);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < kCout / 32; chunk++) {
        uint32_t v[32];
        if (has_k) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * kCout + chunk * 32),
                        v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = 0u;
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]);
        if (do_stats) {
          float s1[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const float x = valid ? f[j] : 0.f;
            s1[j] = x;
            s2[j] = x * x;
          }
          const float cs1 = warp_column_sums(s1, lane);
          const float cs2 = warp_column_sums(s2, lane);
          stat_smem[(ew * 2 + 0) * kCout + chunk * 32 + lane] = cs1;
          stat_smem[(ew * 2 + 1) * kCout + chunk * 32 + lane] = cs2;
        }
        if (valid) {
          uint4* op = reinterpret_cast<uint4*>(p.out + off + chunk * 32);
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            uint4 o;
            o.x = pack_bf16x2(f[j4 * 8 + 0], f[j4 * 8 + 1]);
            o.y = pack_bf16x2(f[j4 * 8 + 2], f[j4 * 8 + 3]);
            o.z = pack_bf16x2(f[j4 * 8 + 4], f[j4 * 8 + 5]);
            o.w = pack_bf16x2(f[j4 * 8 + 6], f[j4 * 8 + 7]);
            op[j4] = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) {
        acc = 0;
        accph ^= 1;
      }
      if (do_stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et < kCout) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int w4 = 0; w4 < 4; w4++) {
            a += stat_smem[(w4 * 2 + 0) * kCout + et];
            b += stat_smem[(w4 * 2 + 1) * kCout + et];
          }
          atomicAdd(p.stat_sum + et, static_cast<double>(a));
          atomicAdd(p.stat_sq + et, static_cast<double>(b));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}


// ------------------------------------------------------------------------------------------------
// Plane-resident fprop (default).  The tap-box kernel above fetches 49 boxes (100 KB) of window pixels per
// 128-position tile: 784 B out of L2 per output voxel, and TMA moves them as 16-byte rows - it runs at 1270
// cycles per kd row against 240 cycles of MMA work.  Here the M tile is one output plane piece of 16 (oh) x 8 (ow)
// positions and the CTA walks a column of pieces along od, keeping the X8 input planes 2*od-3 .. 2*od+3 resident in a
// shared-memory ring: one TMA box (8 pixels x 38 rows, 4.75 KB) per input plane, two new planes per output piece
// (74 B per output voxel).  With rows of 8 pixels = 128 B = one no-swizzle core matrix, the A operand of the taps
// (kd, kh), (kd, kh+1) is a VIEW of plane kd: start = plane + kh * 128 B, LBO (next K core matrix = next input row)
// = 128 B, SBO (next 8 positions = next oh = two input rows down, the conv stride) = 256 B.  The weights are
// resident, so the MMA issuer only ever waits for two plane barriers per piece.
constexpr int kPRows = 2 * 16 + 6;       // input rows feeding 16 output rows
constexpr int kPlaneBytes = 5120;        // 38 rows x 128 B = 4864, padded
constexpr int kPRing = 14;               // 7 planes in use + prefetch
constexpr int kPBarOff = kWBytes + kPRing * kPlaneBytes;
constexpr int kPStatOff = kPBarOff + 512;
constexpr int kPSmem = kPStatOff + 4 * 2 * kCout * 4 + 1024;

struct StemSeg {
  int n, h0, w0, dA, dB, pf, pl;
};
__device__ __forceinline__ StemSeg stem_segment(const StemParams& p, int i, int end) {
  StemSeg s;
  int col = i / p.Do;
  s.dA = i - col * p.Do;
  s.dB = min(p.Do, s.dA + (end - i));
  const int tw = col % p.tiles_w;
  col /= p.tiles_w;
  const int th = col % p.tiles_h;
  s.n = col / p.tiles_h;
  s.h0 = th * 16;
  s.w0 = tw * 8;
  s.pf = max(2 * s.dA - 3, 0);
  s.pl = min(2 * (s.dB - 1) + 3, p.D - 1);
  return s;
}

__global__ void __launch_bounds__(kStemThreads, 1) stem_fprop_plane_kernel(const __grid_constant__ StemParams p) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_p = smem + kWBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kPBarOff);
  uint64_t* empty = full + kPRing;
  uint64_t* tfull = empty + kPRing;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* stat_smem = reinterpret_cast<float*>(smem + kPStatOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int begin = static_cast<int>(static_cast<long long>(p.total) * blockIdx.x / gridDim.x);
  const int end = static_cast<int>(static_cast<long long>(p.total) * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kPRing; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kWBytes / 16; i += kStemThreads)
    reinterpret_cast<uint4*>(smem_w)[i] = __ldg(reinterpret_cast<const uint4*>(p.w2g) + i);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== plane producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&p.x_plane);
      uint32_t seq = 0;
      for (int i = begin; i < end;) {
        const StemSeg s = stem_segment(p, i, end);
        for (int pz = s.pf; pz <= s.pl; pz++, seq++) {
          const uint32_t slot = seq % kPRing, par = (seq / kPRing) & 1u;
          mbar_wait_spin(&empty[slot], par ^ 1u, 3406);
          mbar_arrive_expect_tx(&full[slot], 8u * kPRows * 16u);
          tma_load_4d(smem_p + slot * kPlaneBytes, &p.x_plane, &full[slot], s.w0 * 8, 2 * s.h0 - 3, pz, s.n);
        }
        i += s.dB - s.dA;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    // Everything per piece is scalar ring arithmetic (no divisions, no parameter loads): the issuing thread's own
    // instruction stream was the bound of this kernel (ncu: no stall reason dominant, ~60 instructions per kd).
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kCout, false, false);
      // no-swizzle descriptors, split into 32-bit halves so that the per-MMA work is one 32-bit add per operand:
      //   low word  = start address >> 4 | (LBO >> 4) << 16,   high word = SBO >> 4 | version bit
      const uint32_t a_hi = (256u >> 4) | (1u << 14), b_hi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo0 = ((128u >> 4) << 16) + ((smem_u32(smem_p) & 0x3FFFFu) >> 4);
      const uint32_t b_lo0 = (((kCout * 16u) >> 4) << 16) + ((smem_u32(smem_w) & 0x3FFFFu) >> 4);
      const int D = p.D;
      const bool issue = !(p.debug & 4);
      int acc = 0;
      uint32_t accph = 0;
      // ring cursors (slot, parity) of the next plane to wait for / to hand back, and the running plane count
      int w_slot = 0, r_slot = 0;
      uint32_t w_par = 0;
      int n_waited = 0, n_released = 0, seq0 = 0;
      for (int i = begin; i < end;) {
        const StemSeg s = stem_segment(p, i, end);
        int rel0 = 2 * s.dA - 3 - s.pf;                   // plane 2*od-3 relative to the first loaded plane (<= 0)
        int sl0 = (seq0 + rel0 + 4 * kPRing) % kPRing;   // its ring slot (virtual for planes above the volume)
        for (int od = s.dA; od < s.dB; od++, rel0 += 2) {
          const int kd_lo = max(0, 3 - 2 * od), kd_hi = min(kK - 1, D + 2 - 2 * od);
          mbar_wait_spin(&tempty[acc], accph ^ 1u, 3438);
          const int need = seq0 + rel0 + kd_hi;  // last plane of this piece, as a running count
          while (n_waited <= need) {
            mbar_wait_spin(&full[w_slot], w_par, 3441);
            n_waited++;
            if (++w_slot == kPRing) {
              w_slot = 0;
              w_par ^= 1u;
            }
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kCout);
          // input planes below 2*(od+1)-3 are not needed by any later piece of the column
          const int last_free = seq0 + ((od == s.dB - 1) ? s.pl - s.pf : min(rel0 + 1, s.pl - s.pf));
          if (elect_one_sync()) {
            if (issue) {
#pragma unroll
              for (int kd = 0; kd < kK; kd++) {
                if (kd < kd_lo || kd > kd_hi) continue;
                int slot = sl0 + kd;
                if (slot >= kPRing) slot -= kPRing;
                const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(slot) * (kPlaneBytes >> 4);
                const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(kd) * (kKP * kCout);  // 8 taps x 64 co x 16 B >> 4
#pragma unroll
                for (int i2 = 0; i2 < kKP / 2; i2++) {
                  const uint64_t adesc = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + i2 * 16);           // + 2 rows
                  const uint64_t bdesc = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + i2 * (2 * kCout));  // + 2 taps
                  umma_bf16(d_tmem, adesc, bdesc, idesc, (kd == kd_lo && i2 == 0) ? 0u : 1u);
                }
              }
            }
            umma_commit(&tfull[acc]);
            int rs = r_slot;
            for (int r = n_released; r <= last_free; r++) {
              umma_commit(&empty[rs]);
              if (++rs == kPRing) rs = 0;
            }
          }
          __syncwarp();
          while (n_released <= last_free) {
            n_released++;
            if (++r_slot == kPRing) r_slot = 0;
          }
          sl0 += 2;
          if (sl0 >= kPRing) sl0 -= kPRing;
          if (++acc == 2) {
            acc = 0;
            accph ^= 1u;
          }
        }
        seq0 += s.pl - s.pf + 1;
        i += s.dB - s.dA;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3, ew = warp - 2, et = threadIdx.x - 64;
    const int row = q * 32 + lane;
    const int rh = row >> 3, rw = row & 7;
    const bool do_stats = p.stat_sum != nullptr && !(p.debug & 1);
    const bool do_store = !(p.debug & 2);
    // BatchNorm sums: every thread owns one tile row and keeps fp32 partial sums of its 64 channels in registers
    // for the whole CTA (one row per piece, a few hundred pieces): 2 FMA-class instructions per element, and ONE
    // cross-lane transpose-reduce + fp64 atomic per channel at the very end.  (Reducing per piece cost two
    // 62-instruction shuffle trees, two named barriers and, before that, 128 same-address atomics per piece.)
    float cs1[kCout], cs2[kCout];
#pragma unroll
    for (int j = 0; j < kCout; j++) cs1[j] = cs2[j] = 0.f;
    // the fp32 partials are folded into per-CTA fp64 sums every kFlushPieces pieces (bounded fp32 chains: sum(x) stays
    // accurate when it nearly cancels)
    double cta_sum = 0.0, cta_sq = 0.0;
    constexpr int kFlushPieces = 16;
    int since_flush = 0;
    auto flush_reg_stats = [&]() {
#pragma unroll
      for (int chunk = 0; chunk < kCout / 32; chunk++) {
        float t1[32], t2[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
          t1[j] = cs1[chunk * 32 + j];
          t2[j] = cs2[chunk * 32 + j];
          cs1[chunk * 32 + j] = 0.f;
          cs2[chunk * 32 + j] = 0.f;
        }
        const float c1 = warp_column_sums(t1, lane);
        const float c2 = warp_column_sums(t2, lane);
        stat_smem[(ew * 2 + 0) * kCout + chunk * 32 + lane] = c1;
        stat_smem[(ew * 2 + 1) * kCout + chunk * 32 + lane] = c2;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < kCout) {
#pragma unroll
        for (int w4 = 0; w4 < 4; w4++) {
          cta_sum += static_cast<double>(stat_smem[(w4 * 2 + 0) * kCout + et]);
          cta_sq += static_cast<double>(stat_smem[(w4 * 2 + 1) * kCout + et]);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      since_flush = 0;
    };
    int acc = 0;
    uint32_t accph = 0;
    for (int i = begin; i < end;) {
      const StemSeg s = stem_segment(p, i, end);
      const int oh = s.h0 + rh, ow = s.w0 + rw;
      const bool valid = oh < p.Ho && ow < p.Wo;
      for (int od = s.dA; od < s.dB; od++) {
        const long long off = ((((long long)s.n * p.Do + od) * p.Ho + oh) * p.Wo + ow) * kCout;
        if (do_stats && ++since_flush > kFlushPieces) flush_reg_stats();
        mbar_wait_spin(&tfull[acc], accph, 3547);
        tc_fence_after();
#pragma unroll
        for (int chunk = 0; chunk < kCout / 32; chunk++) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * kCout + chunk * 32),
                        v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]);
          if (do_stats && valid) {
#pragma unroll
            for (int j = 0; j < 32; j++) {
              cs1[chunk * 32 + j] += f[j];
              cs2[chunk * 32 + j] = fmaf(f[j], f[j], cs2[chunk * 32 + j]);
            }
          }
          if (valid && do_store) {
            uint4* op = reinterpret_cast<uint4*>(p.out + off + chunk * 32);
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) {
              uint4 o;
              o.x = pack_bf16x2(f[j4 * 8 + 0], f[j4 * 8 + 1]);
              o.y = pack_bf16x2(f[j4 * 8 + 2], f[j4 * 8 + 3]);
              o.z = pack_bf16x2(f[j4 * 8 + 4], f[j4 * 8 + 5]);
              o.w = pack_bf16x2(f[j4 * 8 + 6], f[j4 * 8 + 7]);
              op[j4] = o;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (++acc == 2) {
          acc = 0;
          accph ^= 1u;
        }
      }
      i += s.dB - s.dA;
    }
    if (do_stats) {
      flush_reg_stats();
      if (et < kCout && begin < end) {
        atomicAdd(p.stat_sum + et, cta_sum);
        atomicAdd(p.stat_sq + et, cta_sq);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: every CTA accumulates D[4 M-tiles x 128 (tap,j) rows][64 cout] over its share of the position boxes
// (64 positions per K-block) and adds the result to dw2 at the end.
constexpr int kWTapBytes = 64 * 16;                       // one tap: 64 positions x 16 B
constexpr int kWAStageBytes = 64 * kWTapBytes;            // 64 tap slots (56 used)
constexpr int kWBStageBytes = 64 * 128;                   // dY box: 64 positions x 64 channels, swizzled
constexpr int kWStageBytes = kWAStageBytes + kWBStageBytes;
constexpr int kWStages = 3;
constexpr int kWBarOff = kWStages * kWStageBytes;
constexpr int kWSmem = kWBarOff + 256 + 1024;

__global__ void __launch_bounds__(kStemThreads, 1) stem_wgrad_kernel(const __grid_constant__ StemParams p) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWBarOff);
  uint64_t* empty = full + kWStages;
  uint64_t* tfull = empty + kWStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_boxes = p.N * p.tiles_d * p.tiles_h * p.tiles_w;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWStages; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  // unused tap slots (56..63) feed accumulator rows that are never stored, but keep them finite
  for (int i = threadIdx.x; i < kWStages * kWStageBytes / 16; i += kStemThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // the whole warp issues: lane l loads taps l and l+32 (56 slots, 49 real), lane 31 additionally the dY box
    int st = 0;
    uint32_t ph = 0;
    for (int box = blockIdx.x; box < total_boxes; box += gridDim.x) {
      const StemTile c = stem_decode(p, box);
      if (lane == 0) {
        mbar_wait(&empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&full[st], kK * kK * kWTapBytes + kWBStageBytes);
      }
      __syncwarp();
      uint8_t* a_dst = smem + st * kWStageBytes;
#pragma unroll
      for (int r = 0; r < 2; r++) {
        const int t = lane + 32 * r;  // slot kd*8 + kh
        const int kd = t >> 3, kh = t & 7;
        if (t < kTapsP && kh < kK) {
          const AxisOff ad = axis_off(kd), ah = axis_off(kh);
          tma_load_5d(a_dst + t * kWTapBytes, &p.x_maps[ad.par * 2 + ah.par], &full[st], 0, c.w0, c.h0 + ah.off,
                      c.d0 + ad.off, c.n);
        }
      }
      if (lane == 31) tma_load_5d(a_dst + kWAStageBytes, &p.dy_map, &full[st], 0, c.w0, c.h0, c.d0, c.n);
      if (++st == kWStages) {
        st = 0;
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kCout, true, true);
      int st = 0;
      uint32_t ph = 0;
      uint32_t accum = 0;
      for (int box = blockIdx.x; box < total_boxes; box += gridDim.x) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + st * kWStageBytes);
        const uint32_t b_addr = a_addr + kWAStageBytes;
#pragma unroll
        for (int k = 0; k < 4; k++) {     // 16 positions per MMA
#pragma unroll
          for (int mt = 0; mt < 4; mt++) {  // 16 taps (128 rows) per M tile
            // A (no swizzle, MN-major): 8-element MN groups (taps) kWTapBytes apart (SBO), 8-position K groups
            // 128 B apart (LBO).  B (128-B swizzle, MN-major): 8-position K groups 1024 B apart (SBO).
            const uint64_t adesc = umma_smem_desc_nosw(a_addr + mt * 16 * kWTapBytes + k * 256, 128, kWTapBytes);
            const uint64_t bdesc = umma_smem_desc_sw128(b_addr + k * 2048, 8192, 1024);
            umma_bf16(tmem_base + static_cast<uint32_t>(mt * kCout), adesc, bdesc, idesc, accum | static_cast<uint32_t>(k));
          }
        }
        accum = 1;
        umma_commit(&empty[st]);
        if (++st == kWStages) {
          st = 0;
          ph ^= 1;
        }
      }
      umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool any = blockIdx.x < total_boxes;
    mbar_wait(tfull, 0);
    tc_fence_after();
    if (any) {
#pragma unroll 1
      for (int mt = 0; mt < 4; mt++) {
        const int m = mt * 128 + row;  // (tap, j) row
        const int t = m >> 3, j = m & 7;
        const bool keep = t < kTapsP && (t % kKP) < kK && j < kK;
#pragma unroll 1
        for (int chunk = 0; chunk < kCout / 32; chunk++) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mt * kCout + chunk * 32), v);
          tmem_ld_wait();
          if (keep) {
            float* dst = p.dw + (long long)m * kCout + chunk * 32;
#pragma unroll
            for (int j4 = 0; j4 < 8; j4++) {
              float4 val;
              val.x = __uint_as_float(v[j4 * 4 + 0]);
              val.y = __uint_as_float(v[j4 * 4 + 1]);
              val.z = __uint_as_float(v[j4 * 4 + 2]);
              val.w = __uint_as_float(v[j4 * 4 + 3]);
              atomicAdd(reinterpret_cast<float4*>(dst + j4 * 4), val);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}


// ------------------------------------------------------------------------------------------------
// Plane-resident wgrad (default).  Same resident X8 planes as stem_fprop_plane_kernel; the dY piece (16 oh x 8 ow
// positions x 64 channels, one swizzled TMA box) streams through a 4-stage ring.  For every input plane kd of a
// piece:  D_kd[(kh, j)][cout] += sum_pos plane_kd[2*oh + kh][ow][j] * dY[pos][cout]
//   A (no swizzle, MN-major): the 8 j of a pixel are the 16-byte MN run, the 8 ow of an input row the 8-position K
//     group (128 B); next K group = next oh = two input rows down (LBO 256 B); next MN group = next kh = next input
//     row (SBO 128 B).  M = 128 is issued, rows 64..127 (kh 8..15) read whatever follows and are never stored.
//   B (128-byte swizzle, MN-major): dY rows, 8-position K groups 1024 B apart.
// The 7 accumulators D_kd (64 columns each) live in TMEM for the whole CTA and are added to dw2 once at the end.
constexpr int kWPStages = 4;
constexpr int kWPDyBytes = 128 * 128;
constexpr int kWPDyOff = kPRing * kPlaneBytes;  // 71680: 1024-aligned
constexpr int kWPBarOff = kWPDyOff + kWPStages * kWPDyBytes;
constexpr int kWPSmem = kWPBarOff + 512 + 1024;
constexpr int kWPThreads = 224;

__global__ void __launch_bounds__(kWPThreads, 1) stem_wgrad_plane_kernel(const __grid_constant__ StemParams p) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_p = smem;
  uint8_t* smem_dy = smem + kWPDyOff;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWPBarOff);
  uint64_t* empty = full + kPRing;
  uint64_t* dfull = empty + kPRing;
  uint64_t* dempty = dfull + kWPStages;
  uint64_t* tfull = dempty + kWPStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int begin = static_cast<int>(static_cast<long long>(p.total) * blockIdx.x / gridDim.x);
  const int end = static_cast<int>(static_cast<long long>(p.total) * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kPRing; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kWPStages; i++) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 6) {
    // ===================== plane producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&p.x_plane);
      uint32_t seq = 0;
      for (int i = begin; i < end;) {
        const StemSeg s = stem_segment(p, i, end);
        for (int pz = s.pf; pz <= s.pl; pz++, seq++) {
          const uint32_t slot = seq % kPRing, par = (seq / kPRing) & 1u;
          mbar_wait_spin(&empty[slot], par ^ 1u, 3809);
          mbar_arrive_expect_tx(&full[slot], 8u * kPRows * 16u);
          tma_load_4d(smem_p + slot * kPlaneBytes, &p.x_plane, &full[slot], s.w0 * 8, 2 * s.h0 - 3, pz, s.n);
        }
        i += s.dB - s.dA;
      }
    }
  } else if (warp == 0) {
    // ===================== dY producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&p.dy_map);
      int st = 0;
      uint32_t ph = 0;
      for (int i = begin; i < end;) {
        const StemSeg s = stem_segment(p, i, end);
        for (int od = s.dA; od < s.dB; od++) {
          mbar_wait_spin(&dempty[st], ph ^ 1u, 3825);
          mbar_arrive_expect_tx(&dfull[st], kWPDyBytes);
          tma_load_5d(smem_dy + st * kWPDyBytes, &p.dy_map, &dfull[st], 0, s.w0, s.h0, od, s.n);
          if (++st == kWPStages) {
            st = 0;
            ph ^= 1u;
          }
        }
        i += s.dB - s.dA;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    // p.debug bit 8 (default; ADNI_STEM_M64=0 clears it): issue M = 64 MMAs - only the 64 real (kh, j) rows, half the
    // tensor-pipe time of the padded M = 128 tile (measured 0.895 -> 0.691 ms per 32 volumes, same results)
    const uint32_t idesc = (p.debug & 8) ? umma_idesc_bf16(64, kCout, true, true) : umma_idesc_bf16(128, kCout, true, true);
    const uint32_t a_hi = (128u >> 4) | (1u << 14);                 // SBO = next kh row
    const uint32_t a_lo0 = ((256u >> 4) << 16) + ((smem_u32(smem_p) & 0x3FFFFu) >> 4);  // LBO = next oh
    const uint64_t b_desc0 = umma_smem_desc_sw128(0, 8192, 1024);
    const uint32_t b_hi = static_cast<uint32_t>(b_desc0 >> 32);
    const uint32_t b_lo0 = static_cast<uint32_t>(b_desc0 & 0xFFFFFFFFull) + ((smem_u32(smem_dy) & 0x3FFFFu) >> 4);
    const int D = p.D;
    int st = 0;
    uint32_t ph = 0;
    // ring cursors as in stem_fprop_plane_kernel
    int w_slot = 0, r_slot = 0;
    uint32_t w_par = 0;
    int n_waited = 0, n_released = 0, seq0 = 0;
    uint32_t started = 0;
    for (int i = begin; i < end;) {
      const StemSeg s = stem_segment(p, i, end);
      int rel0 = 2 * s.dA - 3 - s.pf;
      int sl0 = (seq0 + rel0 + 4 * kPRing) % kPRing;
      for (int od = s.dA; od < s.dB; od++, rel0 += 2) {
        const int kd_lo = max(0, 3 - 2 * od), kd_hi = min(kK - 1, D + 2 - 2 * od);
        mbar_wait_spin(&dfull[st], ph, 3858);
        const int need = seq0 + rel0 + kd_hi;
        while (n_waited <= need) {
          mbar_wait_spin(&full[w_slot], w_par, 3861);
          n_waited++;
          if (++w_slot == kPRing) {
            w_slot = 0;
            w_par ^= 1u;
          }
        }
        tc_fence_after();
        const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(st) * (kWPDyBytes >> 4);
        const int last_free = seq0 + ((od == s.dB - 1) ? s.pl - s.pf : min(rel0 + 1, s.pl - s.pf));
        if (elect_one_sync()) {
#pragma unroll
          for (int kd = 0; kd < kK; kd++) {
            if (kd < kd_lo || kd > kd_hi) continue;
            int slot = sl0 + kd;
            if (slot >= kPRing) slot -= kPRing;
            const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(slot) * (kPlaneBytes >> 4);
            const uint32_t acc_on = (started >> kd) & 1u;
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {  // 16 positions (two oh rows) per MMA
              const uint64_t adesc = (static_cast<uint64_t>(a_hi) << 32) | (a_lo + ks * (512 >> 4));
              const uint64_t bdesc = (static_cast<uint64_t>(b_hi) << 32) | (b_lo + ks * (2048 >> 4));
              umma_bf16(tmem_base + static_cast<uint32_t>(kd * kCout), adesc, bdesc, idesc, acc_on | static_cast<uint32_t>(ks));
            }
          }
          umma_commit(&dempty[st]);
          int rs = r_slot;
          for (int r = n_released; r <= last_free; r++) {
            umma_commit(&empty[rs]);
            if (++rs == kPRing) rs = 0;
          }
        }
        __syncwarp();
#pragma unroll
        for (int kd = 0; kd < kK; kd++)
          if (kd >= kd_lo && kd <= kd_hi) started |= 1u << kd;
        while (n_released <= last_free) {
          n_released++;
          if (++r_slot == kPRing) r_slot = 0;
        }
        sl0 += 2;
        if (sl0 >= kPRing) sl0 -= kPRing;
        if (++st == kWPStages) {
          st = 0;
          ph ^= 1u;
        }
      }
      seq0 += s.pl - s.pf + 1;
      i += s.dB - s.dA;
    }
    if (elect_one_sync()) umma_commit(tfull);
    __syncwarp();
  } else if (warp < 6) {
    // ===================== epilogue (warps 2..5): D_kd -> red.add into dw2, once per CTA =====================
    const int q = warp & 3;
    // M = 128: TMEM lane = row, rows 0..63 = (kh, j), rows 64..127 are the padding half of the tile.
    // M = 64: row i sits in lane (i % 16) + 32 * (i / 16) - 16 rows per 32-lane quadrant.
    const int row = (p.debug & 8) ? (lane < 16 ? q * 16 + lane : 127) : q * 32 + lane;
    // which kd accumulators received anything (a CTA range that never sees a valid input plane for kd leaves it unset)
    uint32_t started = 0;
    for (int i = begin; i < end;) {
      const StemSeg s = stem_segment(p, i, end);
      for (int od = s.dA; od < s.dB; od++)
        for (int kd = 0; kd < kK; kd++) {
          const int pz = 2 * od - 3 + kd;
          if (pz >= 0 && pz < p.D) started |= 1u << kd;
        }
      i += s.dB - s.dA;
    }
    mbar_wait_spin(tfull, 0, 3928);
    tc_fence_after();
    const int kh = row >> 3, j = row & 7;
    const bool keep = row < 64 && kh < kK && j < kK;
#pragma unroll 1
    for (int kd = 0; kd < kK; kd++) {
      if (!((started >> kd) & 1u)) continue;
#pragma unroll 1
      for (int chunk = 0; chunk < kCout / 32; chunk++) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(kd * kCout + chunk * 32), v);
        tmem_ld_wait();
        if (keep) {
          float* dst = p.dw + (long long)(kd * 64 + row) * kCout + chunk * 32;
#pragma unroll
          for (int j4 = 0; j4 < 8; j4++) {
            float4 val;
            val.x = __uint_as_float(v[j4 * 4 + 0]);
            val.y = __uint_as_float(v[j4 * 4 + 1]);
            val.z = __uint_as_float(v[j4 * 4 + 2]);
            val.w = __uint_as_float(v[j4 * 4 + 3]);
            atomicAdd(reinterpret_cast<float4*>(dst + j4 * 4), val);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
struct StemGeom {
  int N, D, H, W, Do, Ho, Wo;
};

int make_x8_maps(StemParams& p, const __nv_bfloat16* x8, const StemGeom& g, int bd, int bh, int bw) {
  const long long sw = 8, sh = (long long)g.Wo * 8, sd = (long long)g.H * g.Wo * 8, sn = (long long)g.D * g.H * g.Wo * 8;
  for (int pd = 0; pd < 2; pd++)
    for (int ph = 0; ph < 2; ph++) {
      const int D2 = (g.D - pd + 1) / 2, H2 = (g.H - ph + 1) / 2;
      const uint64_t dims[5] = {8, (uint64_t)g.Wo, (uint64_t)H2, (uint64_t)D2, (uint64_t)g.N};
      const uint64_t strides[5] = {1, (uint64_t)sw, (uint64_t)(2 * sh), (uint64_t)(2 * sd), (uint64_t)sn};
      const uint32_t box[5] = {8, (uint32_t)bw, (uint32_t)bh, (uint32_t)bd, 1};
      int rc = make_tmap_bf16(&p.x_maps[pd * 2 + ph], x8 + pd * sd + ph * sh, 5, dims, strides, box, false);
      if (rc) return rc;
      p.ext_d[pd] = D2;
      p.ext_h[ph] = H2;
    }
  return ADNI_OK;
}

void pick_box(int rows, int Do, int Ho, int Wo, int* bd, int* bh, int* bw) {
  // widest W run first (contiguous 16-B pixels), then H, then D; product == rows (power of two)
  int w = 1;
  while (w * 2 <= rows && w * 2 <= Wo) w *= 2;
  int h = 1;
  while (w * h * 2 <= rows && h * 2 <= Ho) h *= 2;
  *bw = w;
  *bh = h;
  *bd = rows / (w * h);
  (void)Do;
}

int check_stem(int D, int H, int W, int k, int stride, int pad, int Cout) {
  ADNI_REQUIRE(k == 7 && stride == 2 && pad == 3 && Cout == 64, ADNI_ENOTSUP,
               "stem tensor-core path supports Conv3d(1,64,k=7,stride=2,pad=3) only (k=%d s=%d p=%d Cout=%d)", k,
               stride, pad, Cout);
  ADNI_REQUIRE(D >= 2 && H >= 2 && W >= 2, ADNI_EINVAL, "stem: volume too small");
  return ADNI_OK;
}

}  // namespace
}  // namespace adni

using namespace adni;
#define ST(s) static_cast<cudaStream_t>(s)
typedef __nv_bfloat16 bf16;

extern "C" {

long long adni_stem_x8_elems(int N, int D, int H, int W) {
  const int Wo = (W + 6 - 7) / 2 + 1;
  return (long long)N * D * H * Wo * 8;
}

int adni_stem_expand(const adni_bf16* x, int N, int D, int H, int W, adni_bf16* x8, void* stream) {
  ADNI_REQUIRE(x && x8 && N > 0, ADNI_EINVAL, "stem_expand: bad arguments");
  const int Wo = (W + 6 - 7) / 2 + 1;
  const long long total = (long long)N * D * H * Wo;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16);
  pdl_launch(stem_expand_kernel, grid, 256, 0, ST(stream))(reinterpret_cast<const bf16*>(x), N, D, H, W, Wo,
                                                    reinterpret_cast<bf16*>(x8));
  count_launch();
  ADNI_LAUNCH_CHECK("stem_expand_kernel");
  return ADNI_OK;
}

int adni_stem_weights(const float* w_ncdhw, adni_bf16* w2g, void* stream) {
  ADNI_REQUIRE(w_ncdhw && w2g, ADNI_EINVAL, "stem_weights: null pointer");
  pdl_launch(stem_weights_kernel, (kTapsP * kCout * 8 + 255) / 256, 256, 0, ST(stream))(w_ncdhw, reinterpret_cast<bf16*>(w2g));
  count_launch();
  ADNI_LAUNCH_CHECK("stem_weights_kernel");
  return ADNI_OK;
}

int adni_stem_fprop(const adni_bf16* x8, int N, int D, int H, int W, const adni_bf16* w2g, adni_bf16* y,
                    double* stat_sum, double* stat_sqsum, void* stream) {
  ADNI_REQUIRE(x8 && w2g && y && N > 0, ADNI_EINVAL, "stem_fprop: bad arguments");
  int rc = check_stem(D, H, W, 7, 2, 3, 64);
  if (rc) return rc;
  StemGeom g{N, D, H, W, (D - 1) / 2 + 1, (H - 1) / 2 + 1, (W - 1) / 2 + 1};
  StemParams p;
  memset(&p, 0, sizeof(p));
  pick_box(128, g.Do, g.Ho, g.Wo, &p.bd, &p.bh, &p.bw);
  rc = make_x8_maps(p, reinterpret_cast<const bf16*>(x8), g, p.bd, p.bh, p.bw);
  if (rc) return rc;
  p.w2g = reinterpret_cast<const bf16*>(w2g);
  p.N = N;
  p.Do = g.Do;
  p.Ho = g.Ho;
  p.Wo = g.Wo;
  p.tiles_d = (g.Do + p.bd - 1) / p.bd;
  p.tiles_h = (g.Ho + p.bh - 1) / p.bh;
  p.tiles_w = (g.Wo + p.bw - 1) / p.bw;
  p.out = reinterpret_cast<bf16*>(y);
  p.stat_sum = stat_sum;
  p.stat_sq = stat_sqsum;
  const char* env_planes = getenv("ADNI_STEM_PLANES");
  if (!env_planes || atoi(env_planes) != 0) {
    // a whole X8 row (Wo pixels x 8) is the innermost dimension: the 8 pixels of a piece are ONE 128-byte run for TMA
    const uint64_t dims[4] = {(uint64_t)g.Wo * 8, (uint64_t)H, (uint64_t)D, (uint64_t)N};
    const uint64_t strides[4] = {1, (uint64_t)g.Wo * 8, (uint64_t)H * g.Wo * 8, (uint64_t)D * H * g.Wo * 8};
    const uint32_t box[4] = {64, (uint32_t)kPRows, 1, 1};
    rc = make_tmap_bf16(&p.x_plane, x8, 4, dims, strides, box, false);
    if (rc) return rc;
    p.D = D;
    p.H = H;
    p.tiles_h = (g.Ho + 15) / 16;
    p.tiles_w = (g.Wo + 7) / 8;
    const long long total_pieces = (long long)N * p.tiles_h * p.tiles_w * g.Do;
    ADNI_REQUIRE(total_pieces <= 0x7fffffffLL, ADNI_ENOTSUP, "stem_fprop: too many plane pieces");
    p.total = (int)total_pieces;
    {
      const char* dbg = getenv("ADNI_STEM_DEBUG");
      p.debug = dbg ? atoi(dbg) : 0;
    }
    static bool attr_p = false;
    if (!attr_p) {
      ADNI_CUDA_OK(cudaFuncSetAttribute(stem_fprop_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem));
      attr_p = true;
    }
    pdl_launch(stem_fprop_plane_kernel, std::min(p.total, num_sms()), kStemThreads, kPSmem, ST(stream))(p);
    count_launch();
    ADNI_LAUNCH_CHECK("stem_fprop_plane_kernel");
    return ADNI_OK;
  }
  static bool attr = false;
  if (!attr) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(stem_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem));
    attr = true;
  }
  const int total = N * p.tiles_d * p.tiles_h * p.tiles_w;
  pdl_launch(stem_fprop_kernel, std::min(total, num_sms()), kStemThreads, kFSmem, ST(stream))(p);
  count_launch();
  ADNI_LAUNCH_CHECK("stem_fprop_kernel");
  return ADNI_OK;
}

/* grad_ncdhw fp32 [64][1][7][7][7] = sum over positions dy^T * window(x).  workspace: 512*64 floats. */
int adni_stem_wgrad(const adni_bf16* x8, const adni_bf16* dy, int N, int D, int H, int W, float* workspace,
                    float* grad_ncdhw, void* stream) {
  ADNI_REQUIRE(x8 && dy && workspace && grad_ncdhw && N > 0, ADNI_EINVAL, "stem_wgrad: bad arguments");
  int rc = check_stem(D, H, W, 7, 2, 3, 64);
  if (rc) return rc;
  StemGeom g{N, D, H, W, (D - 1) / 2 + 1, (H - 1) / 2 + 1, (W - 1) / 2 + 1};
  StemParams p;
  memset(&p, 0, sizeof(p));
  pick_box(64, g.Do, g.Ho, g.Wo, &p.bd, &p.bh, &p.bw);
  rc = make_x8_maps(p, reinterpret_cast<const bf16*>(x8), g, p.bd, p.bh, p.bw);
  if (rc) return rc;
  {
    const uint64_t dims[5] = {64, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)N};
    const uint64_t strides[5] = {1, 64, (uint64_t)g.Wo * 64, (uint64_t)g.Ho * g.Wo * 64,
                                 (uint64_t)g.Do * g.Ho * g.Wo * 64};
    const uint32_t box[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, 1};
    rc = make_tmap_bf16(&p.dy_map, dy, 5, dims, strides, box, true);
    if (rc) return rc;
  }
  p.N = N;
  p.Do = g.Do;
  p.Ho = g.Ho;
  p.Wo = g.Wo;
  p.tiles_d = (g.Do + p.bd - 1) / p.bd;
  p.tiles_h = (g.Ho + p.bh - 1) / p.bh;
  p.tiles_w = (g.Wo + p.bw - 1) / p.bw;
  p.dw = workspace;
  ADNI_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(float) * 512 * kCout, ST(stream)));
  const char* env_planes = getenv("ADNI_STEM_PLANES");
  if (!env_planes || atoi(env_planes) != 0) {
    {
      const uint64_t dims[4] = {(uint64_t)g.Wo * 8, (uint64_t)H, (uint64_t)D, (uint64_t)N};
      const uint64_t strides[4] = {1, (uint64_t)g.Wo * 8, (uint64_t)H * g.Wo * 8, (uint64_t)D * H * g.Wo * 8};
      const uint32_t box[4] = {64, (uint32_t)kPRows, 1, 1};
      rc = make_tmap_bf16(&p.x_plane, x8, 4, dims, strides, box, false);
      if (rc) return rc;
    }
    {
      const uint64_t dims[5] = {64, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)g.Do, (uint64_t)N};
      const uint64_t strides[5] = {1, 64, (uint64_t)g.Wo * 64, (uint64_t)g.Ho * g.Wo * 64,
                                   (uint64_t)g.Do * g.Ho * g.Wo * 64};
      const uint32_t box[5] = {64, 8, 16, 1, 1};
      rc = make_tmap_bf16(&p.dy_map, dy, 5, dims, strides, box, true);
      if (rc) return rc;
    }
    p.D = D;
    p.H = H;
    p.tiles_h = (g.Ho + 15) / 16;
    p.tiles_w = (g.Wo + 7) / 8;
    const long long total_pieces = (long long)N * p.tiles_h * p.tiles_w * g.Do;
    ADNI_REQUIRE(total_pieces <= 0x7fffffffLL, ADNI_ENOTSUP, "stem_wgrad: too many plane pieces");
    p.total = (int)total_pieces;
    {
      const char* m64 = getenv("ADNI_STEM_M64");
      p.debug = (m64 && atoi(m64) == 0) ? 0 : 8;
    }
    static bool attr_p = false;
    if (!attr_p) {
      ADNI_CUDA_OK(cudaFuncSetAttribute(stem_wgrad_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWPSmem));
      attr_p = true;
    }
    pdl_launch(stem_wgrad_plane_kernel, std::min(p.total, num_sms()), kWPThreads, kWPSmem, ST(stream))(p);
    count_launch();
    ADNI_LAUNCH_CHECK("stem_wgrad_plane_kernel");
    pdl_launch(stem_wgrad_unpack_kernel, (kCout * 343 + 255) / 256, 256, 0, ST(stream))(workspace, grad_ncdhw);
    count_launch();
    ADNI_LAUNCH_CHECK("stem_wgrad_unpack_kernel");
    return ADNI_OK;
  }
  static bool attr = false;
  if (!attr) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem));
    attr = true;
  }
  const int total = N * p.tiles_d * p.tiles_h * p.tiles_w;
  pdl_launch(stem_wgrad_kernel, std::min(total, num_sms()), kStemThreads, kWSmem, ST(stream))(p);
  count_launch();
  ADNI_LAUNCH_CHECK("stem_wgrad_kernel");
  pdl_launch(stem_wgrad_unpack_kernel, (kCout * 343 + 255) / 256, 256, 0, ST(stream))(workspace, grad_ncdhw);
  count_launch();
  ADNI_LAUNCH_CHECK("stem_wgrad_unpack_kernel");
  return ADNI_OK;
}

}  // extern "C"

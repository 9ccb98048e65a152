// Small-channel Conv3d engine on the warp-level tensor-core path (mma.sync m16n8k16, bf16 x bf16 -> fp32):
// the Small_PET_CNN / early-fusion / feature-map-fusion stem stacks of the reference
// (pkg/models/pet_models/pet_cnn.py:18-28, fusion_models/early_fusion.py:34-44,
// fusion_models/anat_pet_featuremapfusion.py:40-64): Conv3d 'same' + bias with Cin in {1, 8, 16, 32, 64},
// Cout in {8 .. 64}, k in {3, 5, 7}.  Their channel counts are far below the 64-channel K slices / 128-row tiles
// of the tcgen05 engines (padding to them would execute 8-32x the algorithmic FLOPs), and the old CUDA-core direct
// kernel ran them at 7-12 TFLOP/s, which made the reference-faithful PET branch 85 % of a training step.
//
// Formulation (fprop, and dgrad = the same kernel over dy with flipped taps and the ITO weight copy):
//   a CTA stages a halo tile of the NDHWC input in shared memory ((R+k-1) x (8+k-1) x (16+k-1) positions, 16-byte
//   channel chunks XOR-swizzled by position so that ldmatrix rows never share a bank group) and the whole weight
//   tensor as ready-made B fragments; a warp owns R rows of 16 consecutive w positions (M = 16 per mma) and walks
//   K = taps x Cin in k16 steps whose two k8 halves are (tap, 8-channel chunk) pairs - each half has its own row
//   addresses in ldmatrix.x4, so any tap pairs with any other.
//   Cin = 1 (the first layer): the tile loader expands the volume in shared memory to 8 "channels" = the 8 inputs
//   w-pad .. w-pad+7 of every position (the stem's X8 idea, done on chip), which turns the k^3 conv into a
//   (k, k, 1) conv over 8 channels: K = k*k*8 instead of a 1-wide K no tensor instruction can use.
// wgrad: M = Cout, N = (tap, Cin) columns, K = positions; both operands come out of the same tiles through
//   ldmatrix.trans; each warp keeps its share of the [Cout] x [taps*Cin] accumulator in registers over all the tiles
//   of its CTA and flushes once with red.global.add.
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace adni {
extern void count_launch();

namespace {

constexpr int kThreads = 256;  // 8 warps
constexpr int TH = 8, TW = 16;
constexpr int kMaxQ = 1024;    // k8 halves (taps x channel chunks) the offset table holds

struct SmallParams {
  const __nv_bfloat16* in;   // [N][Di][Hi][Wi][Ci]
  const __nv_bfloat16* w;    // [Co][taps][Ci] (OTI; ITO of the forward conv for dgrad)
  const float* bias;         // [Co] or null
  __nv_bfloat16* out;        // [N][Do][Ho][Wo][Co]
  double* ssum;              // [Co] or null
  double* ssq;
  int N, Di, Hi, Wi, Ci;     // Ci = logical input channels (1 in C1 mode)
  int Do, Ho, Wo, Co;
  int k;                     // isotropic kernel extent
  int pad;                   // low-side padding of THIS conv (dgrad: k - 1 - pad of the forward conv)
  int flip;                  // 1: taps are read mirrored (dgrad)
  int CH;                    // 16-byte channel chunks per position in smem (Cie / 8), a power of two
  int chs;                   // log2(CH)
  int sh;                    // swizzle shift: chunk ^= (p >> sh) & (CH - 1)
  int taps_e;                // effective taps: k*k*k, or k*k in C1 mode
  int Q, S;                  // k8 halves, k16 steps
  int HD, HH, HW;            // halo tile extents (positions); C1 mode: HW = TW
  int tiles_d, tiles_h, tiles_w;
  long long total_tiles;
  int RAWW;                  // C1 mode: raw row width staged before the expansion (TW + 7)
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int bytes = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// byte offset of (position p, logical chunk c) inside a swizzled tile
__device__ __forceinline__ uint32_t swz(int p, int c, int CH, int sh) {
  return static_cast<uint32_t>((p * CH + (c ^ ((p >> sh) & (CH - 1)))) << 4);
}

struct TileOrigin {
  int n, d0, h0, w0;
};
__device__ __forceinline__ TileOrigin tile_origin(const SmallParams& p, int tile, int TD) {
  // 32-bit arithmetic (the host refuses more than 2^31 tiles): four 64-bit divisions per tile and thread cost as much
  // as the tile's tensor-core work on the 1-channel first layer (131072 tiles of 512 outputs per batch)
  TileOrigin o;
  int t = tile;
  o.w0 = (t % p.tiles_w) * TW;
  t /= p.tiles_w;
  o.h0 = (t % p.tiles_h) * TH;
  t /= p.tiles_h;
  o.d0 = (t % p.tiles_d) * TD;
  o.n = t / p.tiles_d;
  return o;
}

// Stage the halo tile of `in` whose output origin is `o` into smem (swizzled 16-byte chunks, zero outside the volume).
// Vector mode (Ci % 8 == 0): one cp.async per chunk, a warp per halo row (the row is contiguous in global memory).
__device__ __forceinline__ void load_tile_vec(const SmallParams& p, const TileOrigin& o, uint8_t* tile_s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = p.HD * p.HH, row_chunks = p.HW * p.CH;
  const uint32_t base = smem_u32(tile_s);
  for (int r = warp; r < rows; r += kThreads / 32) {
    const int hd = r / p.HH, hh = r - hd * p.HH;
    const int id = o.d0 + hd - p.pad, ih = o.h0 + hh - p.pad;
    const bool row_ok = id >= 0 && id < p.Di && ih >= 0 && ih < p.Hi;
    const long long grow = ((static_cast<long long>(o.n) * p.Di + (row_ok ? id : 0)) * p.Hi + (row_ok ? ih : 0)) * p.Wi;
    for (int x = lane; x < row_chunks; x += 32) {
      const int hw = x >> p.chs, c = x & (p.CH - 1);
      const int iw = o.w0 + hw - p.pad;
      const bool ok = row_ok && iw >= 0 && iw < p.Wi;
      const __nv_bfloat16* src = p.in + (grow + (ok ? iw : 0)) * p.Ci + c * 8;
      cp_async16(base + swz(r * p.HW + hw, c, p.CH, p.sh), src, ok);
    }
  }
}

// C1 mode: raw 1-channel rows (TW + 7 inputs: w0 - pad .. w0 - pad + TW + 6) -> smem, then every position becomes the
// 16-byte "pixel" of its 8 inputs w - pad .. w - pad + 7.
// The raw rows are fetched one tile AHEAD into registers (a warp per halo row, RAWW = 23 <= 32 inputs: one coalesced
// load per row; <= kC1Rows rows per warp) while the current tile is being computed, then committed to shared memory and
// expanded: the global-load latency of a tile with ~1 us of tensor work would otherwise be fully exposed.
constexpr int kC1Rows = 20;   // rows per warp: 12 x 12 halo rows of the k = 5, depth-8 tile / 8 warps = 18 (k = 7 uses depth 4: 17.5)
struct C1Prefetch {
  uint32_t v[kC1Rows / 2];    // two bf16 per register
};

__device__ __forceinline__ void c1_issue(const SmallParams& p, const TileOrigin& o, C1Prefetch& pf) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = p.HD * p.HH;
  const int iw = o.w0 + lane - p.pad;
  const bool col_ok = lane < p.RAWW && iw >= 0 && iw < p.Wi;
  const uint16_t* in16 = reinterpret_cast<const uint16_t*>(p.in);
#pragma unroll
  for (int u = 0; u < kC1Rows; u += 2) {
    uint32_t pair = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int r = warp + (u + h) * (kThreads / 32);
      const int hd = r / p.HH, hh = r - hd * p.HH;
      const int id = o.d0 + hd - p.pad, ih = o.h0 + hh - p.pad;
      const bool ok = r < rows && col_ok && id >= 0 && id < p.Di && ih >= 0 && ih < p.Hi;
      const uint32_t val = ok ? __ldg(in16 + ((static_cast<long long>(o.n) * p.Di + id) * p.Hi + ih) * p.Wi + iw) : 0u;
      pair |= val << (16 * h);
    }
    pf.v[u / 2] = pair;
  }
}

// registers -> raw rows in smem -> (barrier) -> every position's 16-byte pixel of its 8 inputs w - pad .. w - pad + 7
__device__ __forceinline__ void c1_commit(const SmallParams& p, const C1Prefetch& pf, uint8_t* tile_s,
                                          __nv_bfloat16* raw_s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = p.HD * p.HH;
  uint16_t* raw16 = reinterpret_cast<uint16_t*>(raw_s);
  if (lane < p.RAWW) {
#pragma unroll
    for (int u = 0; u < kC1Rows; u++) {
      const int r = warp + u * (kThreads / 32);
      if (r < rows) raw16[r * p.RAWW + lane] = static_cast<uint16_t>(pf.v[u / 2] >> (16 * (u & 1)));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows * TW; i += kThreads) {
    const int r = i >> 4, hw = i & 15;
    const uint16_t* s = raw16 + r * p.RAWW + hw;
    uint4 v;
    v.x = s[0] | (static_cast<uint32_t>(s[1]) << 16);
    v.y = s[2] | (static_cast<uint32_t>(s[3]) << 16);
    v.z = s[4] | (static_cast<uint32_t>(s[5]) << 16);
    v.w = s[6] | (static_cast<uint32_t>(s[7]) << 16);
    *reinterpret_cast<uint4*>(tile_s + (static_cast<size_t>(i) << 4)) = v;   // CH = 1: no swizzle
  }
}

// offset table: q -> (position offset of the tap inside the halo tile) * 8 + chunk
__device__ __forceinline__ void fill_tap_table(const SmallParams& p, bool c1, int* tab) {
  for (int q = threadIdx.x; q < p.Q; q += kThreads) {
    const int tap = c1 ? q : q >> p.chs, c = c1 ? 0 : q & (p.CH - 1);
    int off;
    if (c1) {
      const int tkd = tap / p.k, tkh = tap - tkd * p.k;
      off = (tkd * p.HH + tkh) * p.HW;
    } else {
      const int tkd = tap / (p.k * p.k), rem = tap - tkd * p.k * p.k, tkh = rem / p.k, tkw = rem - tkh * p.k;
      off = (tkd * p.HH + tkh) * p.HW + tkw;
    }
    tab[q] = off * 8 + c;
  }
}

// =================================================================================================
// fprop / dgrad
// =================================================================================================
template <int NT, int R, bool C1>
__global__ void __launch_bounds__(kThreads) small_fprop_kernel(const SmallParams p) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t smem[];
  // layout: [B fragments S*NT*32 uint2][tap table Q ints][stats 2*NT*8 doubles][tile][raw (C1)]
  uint2* wfrag = reinterpret_cast<uint2*>(smem);
  size_t off = static_cast<size_t>(p.S) * NT * 32 * 8;
  int* tab = reinterpret_cast<int*>(smem + off);
  off += ((static_cast<size_t>(p.Q) * 4 + 15) / 16) * 16;
  double* stat_s = reinterpret_cast<double*>(smem + off);
  off += 2 * NT * 8 * 8;
  uint8_t* tile_s = smem + off;
  off += static_cast<size_t>(p.HD) * p.HH * p.HW * p.CH * 16;
  __nv_bfloat16* raw_s = reinterpret_cast<__nv_bfloat16*>(smem + off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int taps = p.k * p.k * p.k;
  const int co0 = blockIdx.y * NT * 8;   // output-channel chunk of this CTA (wide layers with long K are split over y)

  // ---- B fragments: wfrag[(s*NT + j)*32 + lane] = {B[16s+2t..+1][8j+g], B[16s+8+2t..+1][8j+g]} -----------------
  for (int i = threadIdx.x; i < p.S * NT * 32; i += kThreads) {
    const int l = i & 31, sj = i >> 5, j = sj % NT, s = sj / NT;
    const int n = co0 + 8 * j + (l >> 2), tt = l & 3;
    uint32_t regs[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int q = 2 * s + h;
      uint32_t v = 0;
      if (q < p.Q && n < p.Co) {
        if (C1) {  // k index inside the half = kw (0..7, zero beyond k); q = (kd, kh)
          const int kw0 = 2 * tt;
          const long long base = static_cast<long long>(n) * taps + q * p.k;
          const uint16_t* w16 = reinterpret_cast<const uint16_t*>(p.w);
          const uint32_t lo = kw0 < p.k ? w16[base + kw0] : 0u;
          const uint32_t hi = kw0 + 1 < p.k ? w16[base + kw0 + 1] : 0u;
          v = lo | (hi << 16);
        } else {
          const int tap = q >> p.chs, c = q & (p.CH - 1);
          const int tsel = p.flip ? taps - 1 - tap : tap;
          v = *reinterpret_cast<const uint32_t*>(p.w + (static_cast<long long>(n) * taps + tsel) * p.Ci + c * 8 + 2 * tt);
        }
      }
      regs[h] = v;
    }
    wfrag[i] = make_uint2(regs[0], regs[1]);
  }
  fill_tap_table(p, C1, tab);
  if (threadIdx.x < 2 * NT * 8) stat_s[threadIdx.x] = 0.0;
  __syncthreads();

  // this lane's row inside an ldmatrix.x4: matrices 0/1 = rows 0-7 / 8-15 of k-half 0, matrices 2/3 = of k-half 1
  const int lm = lane >> 3;
  const int lrow = (lane & 7) + 8 * (lm & 1);
  const int lhalf = lm >> 1;
  const uint32_t tile_base = smem_u32(tile_s);
  const bool do_stats = p.ssum != nullptr;

  const int total_tiles = static_cast<int>(p.total_tiles);
  C1Prefetch pf;
  if (C1 && static_cast<int>(blockIdx.x) < total_tiles) c1_issue(p, tile_origin(p, blockIdx.x, R), pf);
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const TileOrigin o = tile_origin(p, tile, R);
    __syncthreads();  // the previous tile's fragments have been read
    if (C1) {
      c1_commit(p, pf, tile_s, raw_s);
    } else {
      load_tile_vec(p, o, tile_s);
      cp_async_wait_all();
    }
    __syncthreads();
    if (C1 && tile + static_cast<int>(gridDim.x) < total_tiles)
      c1_issue(p, tile_origin(p, tile + gridDim.x, R), pf);   // the next tile's rows are in flight under this tile's MMAs

    float acc[R][NT][4];
#pragma unroll
    for (int i = 0; i < R; i++)
#pragma unroll
      for (int j = 0; j < NT; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[i][j][e] = 0.f;
    int prow[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int r = warp * R + i;           // row of the tile: td = r / TH, th = r % TH
      prow[i] = ((r >> 3) * p.HH + (r & 7)) * p.HW + lrow;
    }

    for (int s = 0; s < p.S; s++) {
      const int q = min(2 * s + lhalf, p.Q - 1);   // an odd tail half re-reads a valid address; its weights are zero
      const int e = tab[q];
      const int toff = e >> 3, c = e & 7;
      uint32_t a[R][4];
#pragma unroll
      for (int i = 0; i < R; i++) ldsm_x4(tile_base + swz(prow[i] + toff, c, p.CH, p.sh), a[i]);
#pragma unroll
      for (int j = 0; j < NT; j++) {
        const uint2 b = wfrag[(s * NT + j) * 32 + lane];
#pragma unroll
        for (int i = 0; i < R; i++) mma_bf16(acc[i][j], a[i], b.x, b.y);
      }
    }

    // ---- epilogue: bias, bf16 store, per-channel sums ----
    float s1[NT][2], s2[NT][2];
#pragma unroll
    for (int j = 0; j < NT; j++) s1[j][0] = s1[j][1] = s2[j][0] = s2[j][1] = 0.f;
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int r = warp * R + i;
      const int od = o.d0 + (r >> 3), oh = o.h0 + (r & 7);
      const bool row_ok = od < p.Do && oh < p.Ho;
      const long long rbase = ((static_cast<long long>(o.n) * p.Do + od) * p.Ho + oh) * p.Wo;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const int ow = o.w0 + g + 8 * half;
        const bool ok = row_ok && ow < p.Wo;
#pragma unroll
        for (int j = 0; j < NT; j++) {
          const int ch = co0 + 8 * j + 2 * t;
          float v0 = acc[i][j][2 * half], v1 = acc[i][j][2 * half + 1];
          if (p.bias != nullptr && ch < p.Co) {
            v0 += __ldg(p.bias + ch);
            v1 += __ldg(p.bias + ch + 1);
          }
          if (ok && ch < p.Co) {
            *reinterpret_cast<uint32_t*>(p.out + (rbase + ow) * p.Co + ch) = pack_bf16x2(v0, v1);
            s1[j][0] += v0;
            s1[j][1] += v1;
            s2[j][0] += v0 * v0;
            s2[j][1] += v1 * v1;
          }
        }
      }
    }
    if (do_stats) {
#pragma unroll
      for (int j = 0; j < NT; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          float a1 = s1[j][e], a2 = s2[j][e];
#pragma unroll
          for (int m = 4; m <= 16; m <<= 1) {
            a1 += __shfl_xor_sync(0xffffffffu, a1, m);
            a2 += __shfl_xor_sync(0xffffffffu, a2, m);
          }
          if (g == 0) {
            atomicAdd(&stat_s[(8 * j + 2 * t + e) * 2 + 0], static_cast<double>(a1));
            atomicAdd(&stat_s[(8 * j + 2 * t + e) * 2 + 1], static_cast<double>(a2));
          }
        }
    }
  }
  if (do_stats) {
    __syncthreads();
    if (threadIdx.x < NT * 8 && co0 + threadIdx.x < p.Co) {
      atomicAdd(p.ssum + co0 + threadIdx.x, stat_s[threadIdx.x * 2 + 0]);
      atomicAdd(p.ssq + co0 + threadIdx.x, stat_s[threadIdx.x * 2 + 1]);
    }
  }
}

// =================================================================================================
// wgrad: dw[co][tap][ci] += sum_pos dy[pos][co] * x[pos + tap][ci]
// =================================================================================================
struct SmallWgradParams {
  SmallParams x;             // geometry + the x tensor (`in`); `out` unused
  const __nv_bfloat16* dy;   // [N][Do][Ho][Wo][Co]
  float* dw;                 // [Co][taps][Ci] fp32, accumulated
  int CHo, sho, chos;        // dy tile: 16-byte chunks per position, swizzle shift, log2(CHo)
  int n_tiles_total;         // N-tiles (tap, chunk) = Q
  int nt_per_cta;            // N-tiles a CTA (blockIdx.y) covers
};

template <int MT, int NW, bool C1, int TD>
__global__ void __launch_bounds__(kThreads) small_wgrad_kernel(const SmallWgradParams wp) {
  pdl_enter();
  const SmallParams& p = wp.x;
  extern __shared__ __align__(16) uint8_t smem[];
  int* tab = reinterpret_cast<int*>(smem);
  size_t off = ((static_cast<size_t>(p.Q) * 4 + 15) / 16) * 16;
  uint8_t* tile_s = smem + off;
  off += static_cast<size_t>(p.HD) * p.HH * p.HW * p.CH * 16;
  uint8_t* dy_s = smem + off;
  off += static_cast<size_t>(TD) * TH * TW * wp.CHo * 16;
  __nv_bfloat16* raw_s = reinterpret_cast<__nv_bfloat16*>(smem + off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int taps = p.k * p.k * p.k;
  fill_tap_table(p, C1, tab);

  // this warp's N-tiles: q = q0 + warp + 8*i, i < NW
  const int q0 = blockIdx.y * wp.nt_per_cta;
  const int q_end = min(q0 + wp.nt_per_cta, wp.n_tiles_total);
  float acc[MT][NW][4];
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int i = 0; i < NW; i++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[m][i][e] = 0.f;

  const uint32_t tile_base = smem_u32(tile_s), dy_base = smem_u32(dy_s);
  // ldmatrix row roles.  A (dy^T, x4.trans): matrices 0..3 = (k 0-7, m 0-7), (k 0-7, m 8-15), (k 8-15, m 0-7),
  // (k 8-15, m 8-15): lane -> position row (lane & 7) + 8 * (lane >> 4), chunk bit (lane >> 3) & 1.
  const int a_row = (lane & 7) + 8 * (lane >> 4);
  const int a_cb = (lane >> 3) & 1;
  // B (x, x2.trans): matrices 0/1 = k 0-7 / k 8-15: lanes 0-15 give the row addresses
  const int b_row = lane & 15;
  __syncthreads();

  const int total_tiles = static_cast<int>(p.total_tiles);
  C1Prefetch pf;
  if (C1 && static_cast<int>(blockIdx.x) < total_tiles) c1_issue(p, tile_origin(p, blockIdx.x, TD), pf);
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const TileOrigin o = tile_origin(p, tile, TD);
    __syncthreads();
    if (C1) {
      c1_commit(p, pf, tile_s, raw_s);
    } else {
      load_tile_vec(p, o, tile_s);
    }
    {  // dy tile: TD x TH rows of TW positions x Co channels, zero outside the output volume
      const int row_chunks = TW * wp.CHo;
      for (int r = warp; r < TD * TH; r += kThreads / 32) {
        const int od = o.d0 + (r >> 3), oh = o.h0 + (r & 7);
        const bool row_ok = od < p.Do && oh < p.Ho;
        const long long grow = ((static_cast<long long>(o.n) * p.Do + (row_ok ? od : 0)) * p.Ho + (row_ok ? oh : 0)) * p.Wo;
        for (int x = lane; x < row_chunks; x += 32) {
          const int pw = x >> wp.chos, c = x & (wp.CHo - 1);
          const int ow = o.w0 + pw;
          const bool ok = row_ok && ow < p.Wo;
          cp_async16(dy_base + swz(r * TW + pw, c, wp.CHo, wp.sho), wp.dy + (grow + (ok ? ow : 0)) * p.Co + c * 8, ok);
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();
    if (C1 && tile + static_cast<int>(gridDim.x) < total_tiles) c1_issue(p, tile_origin(p, tile + gridDim.x, TD), pf);

    for (int r = 0; r < TD * TH; r++) {
      const int pr = ((r >> 3) * p.HH + (r & 7)) * p.HW;   // first position of the row inside the halo tile
      uint32_t a[MT][4];
#pragma unroll
      for (int m = 0; m < MT; m++) ldsm_x4_t(dy_base + swz(r * TW + a_row, 2 * m + a_cb, wp.CHo, wp.sho), a[m]);
#pragma unroll
      for (int i = 0; i < NW; i++) {
        const int q = q0 + warp + 8 * i;
        if (q < q_end) {
          const int e = tab[q];
          uint32_t b[2];
          ldsm_x2_t(tile_base + swz(pr + (e >> 3) + b_row, e & 7, p.CH, p.sh), b);
#pragma unroll
          for (int m = 0; m < MT; m++) mma_bf16(acc[m][i], a[m], b[0], b[1]);
        }
      }
    }
  }

  // ---- flush: acc[m][i] = (co = 16m + g (+8), column n = 2t (+1) of N-tile q) ----
#pragma unroll
  for (int i = 0; i < NW; i++) {
    const int q = q0 + warp + 8 * i;
    if (q >= q_end) continue;
    const int tap_e = C1 ? q : q >> p.chs, c = C1 ? 0 : q & (p.CH - 1);
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int co = 16 * m + g + 8 * (e >> 1);
        const int nn = 2 * t + (e & 1);
        if (co >= p.Co) continue;
        long long idx;
        if (C1) {
          if (nn >= p.k) continue;           // columns k .. 7 of the expanded window carry no weight
          idx = static_cast<long long>(co) * taps + tap_e * p.k + nn;
        } else {
          idx = (static_cast<long long>(co) * taps + tap_e) * p.Ci + c * 8 + nn;
        }
        atomicAdd(wp.dw + idx, acc[m][i][e]);
      }
  }
}

int ilog2(int v) {
  int s = 0;
  while ((1 << s) < v) s++;
  return s;
}

// Geometry shared by fprop / dgrad / wgrad.  `Ci` = channels of the tensor the taps slide over.
bool plan_small(SmallParams& p, int N, int Di, int Hi, int Wi, int Ci, int Do, int Ho, int Wo, int Co, int k, int pad,
                int flip, int TD) {
  const bool c1 = Ci == 1;
  p.N = N, p.Di = Di, p.Hi = Hi, p.Wi = Wi, p.Ci = Ci, p.Do = Do, p.Ho = Ho, p.Wo = Wo, p.Co = Co;
  p.k = k, p.pad = pad, p.flip = flip;
  p.CH = c1 ? 1 : Ci / 8;
  p.chs = ilog2(p.CH);
  p.sh = p.CH > 1 ? ilog2(8 / p.CH) : 0;
  p.taps_e = c1 ? k * k : k * k * k;
  p.Q = p.taps_e * p.CH;
  p.S = (p.Q + 1) / 2;
  p.HD = TD + k - 1, p.HH = TH + k - 1, p.HW = c1 ? TW : TW + k - 1;
  p.RAWW = TW + 7;
  p.tiles_d = (Do + TD - 1) / TD, p.tiles_h = (Ho + TH - 1) / TH, p.tiles_w = (Wo + TW - 1) / TW;
  p.total_tiles = static_cast<long long>(N) * p.tiles_d * p.tiles_h * p.tiles_w;
  if (c1 && p.HD * p.HH > kC1Rows * (kThreads / 32)) return false;   // the tile loader's register prefetch holds kC1Rows rows per warp
  return p.Q <= kMaxQ && p.total_tiles < (1LL << 31);
}

size_t fprop_smem(const SmallParams& p, int NT) {
  size_t b = static_cast<size_t>(p.S) * NT * 32 * 8;
  b += ((static_cast<size_t>(p.Q) * 4 + 15) / 16) * 16;
  b += 2 * NT * 8 * 8;
  b += static_cast<size_t>(p.HD) * p.HH * p.HW * p.CH * 16;
  if (p.Ci == 1) b += static_cast<size_t>(p.HD) * p.HH * p.RAWW * 2 + 16;
  return b;
}

constexpr size_t kSmemLimit = 220 * 1024;

template <int NT, int R, bool C1>
int launch_fprop_t(const SmallParams& p, size_t smem, int ysplit, cudaStream_t st) {
  auto kern = small_fprop_kernel<NT, R, C1>;
  static size_t attr = 0;
  if (smem > attr) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  const int per_sm = std::max<size_t>(1, std::min<size_t>(4, kSmemLimit / (smem + 1024)));
  const long long want = std::max<long long>(1, static_cast<long long>(num_sms()) * per_sm / ysplit);
  dim3 grid(static_cast<unsigned>(std::min<long long>(p.total_tiles, want)), static_cast<unsigned>(ysplit));
  pdl_launch(kern, grid, kThreads, smem, st)(p);
  count_launch();
  ADNI_LAUNCH_CHECK("small_fprop_kernel");
  return ADNI_OK;
}

// Tile depth (R = 4 rows per warp when it fits, else 2) and output-channel chunk (NT n-tiles of 8 per CTA; the rest
// of a wide layer goes to blockIdx.y) such that B fragments + halo tile fit shared memory.
bool pick_small_cfg(const SmallParams& p4, const SmallParams& p2, int nt_total, int* nt, bool* r4) {
  // (the 1-channel, 8-output first layer additionally has an R = 8 variant, chosen in small_conv_fprop)
  for (int n = nt_total; n >= 1; n >>= 1) {
    if (n <= 4 && fprop_smem(p4, n) <= kSmemLimit) {
      *nt = n, *r4 = true;
      return true;
    }
    if (fprop_smem(p2, n) <= kSmemLimit) {
      *nt = n, *r4 = false;
      return true;
    }
  }
  return false;
}

template <bool C1>
int launch_fprop_nt(const SmallParams& p4, const SmallParams& p2, int nt_total, cudaStream_t st) {
  int NT = 0;
  bool r4 = false;
  if (!pick_small_cfg(p4, p2, nt_total, &NT, &r4)) {
    set_error("small conv: tile + weights exceed shared memory (Cin=%d Cout=%d k=%d)", p4.Ci, p4.Co, p4.k);
    return ADNI_ENOTSUP;
  }
  const SmallParams& p = r4 ? p4 : p2;
  const size_t smem = fprop_smem(p, NT);
  const int ysplit = nt_total / NT;
  switch (NT) {
    case 1:
      return r4 ? launch_fprop_t<1, 4, C1>(p, smem, ysplit, st) : launch_fprop_t<1, 2, C1>(p, smem, ysplit, st);
    case 2:
      return r4 ? launch_fprop_t<2, 4, C1>(p, smem, ysplit, st) : launch_fprop_t<2, 2, C1>(p, smem, ysplit, st);
    case 4:
      return r4 ? launch_fprop_t<4, 4, C1>(p, smem, ysplit, st) : launch_fprop_t<4, 2, C1>(p, smem, ysplit, st);
    default:
      return launch_fprop_t<8, 2, C1>(p, smem, ysplit, st);
  }
}

int nt_for(int Co) { return Co <= 8 ? 1 : (Co <= 16 ? 2 : (Co <= 32 ? 4 : 8)); }

inline int oext(int in, int k, int pad) { return in + 2 * pad - (k - 1); }

template <int MT, int NW, bool C1, int TD = 4>
int launch_wgrad_t(SmallWgradParams& wp, cudaStream_t st) {
  const SmallParams& p = wp.x;
  auto kern = small_wgrad_kernel<MT, NW, C1, TD>;
  size_t smem = ((static_cast<size_t>(p.Q) * 4 + 15) / 16) * 16 + static_cast<size_t>(p.HD) * p.HH * p.HW * p.CH * 16 +
                static_cast<size_t>(TD) * TH * TW * wp.CHo * 16 + 32;   // + slack: Cout = 8 reads one chunk past the dy tile
  if (C1) smem += static_cast<size_t>(p.HD) * p.HH * p.RAWW * 2 + 16;
  if (smem > kSmemLimit) {
    set_error("small conv wgrad: %zu bytes of shared memory needed (Cin=%d Cout=%d k=%d)", smem, p.Ci, p.Co, p.k);
    return ADNI_ENOTSUP;
  }
  static size_t attr = 0;
  if (smem > attr) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  wp.nt_per_cta = 8 * NW;
  const int ysplit = (wp.n_tiles_total + wp.nt_per_cta - 1) / wp.nt_per_cta;
  const int per_sm = std::max<size_t>(1, std::min<size_t>(NW <= 4 ? 4 : 2, kSmemLimit / (smem + 1024)));
  const long long want = std::max<long long>(1, static_cast<long long>(num_sms()) * per_sm / ysplit);
  dim3 grid(static_cast<unsigned>(std::min<long long>(p.total_tiles, want)), static_cast<unsigned>(ysplit));
  pdl_launch(kern, grid, kThreads, smem, st)(wp);
  count_launch();
  ADNI_LAUNCH_CHECK("small_wgrad_kernel");
  return ADNI_OK;
}

}  // namespace

// Which geometries this engine takes (stride 1, dilation 1; the taps slide over `Ci_slide` channels and produce
// `Co_out`): Ci_slide in {1, 8, 16, 32, 64}, Co_out a multiple of 8 up to 64 (Cin = 1 never needs a dgrad).
bool small_conv_supported(const adni_conv3d_geom& g, int pass) {
  auto pow2 = [](int v, int lo, int hi) { return v >= lo && v <= hi && (v & (v - 1)) == 0; };
  if (g.stride != 1 || g.dil != 1 || g.k < 1 || g.k > 7 || g.pad > g.k - 1) return false;
  if (!pow2(g.Cout, 8, 64)) return false;
  if (g.Cin != 1 && !pow2(g.Cin, 8, 64)) return false;
  if (g.Cin == 1 && pass == 1) return false;               // the first layer (on-chip window expansion) has no dgrad
  const int slide = pass == 1 ? g.Cout : g.Cin;            // channels of the tensor the taps slide over
  const int outc = pass == 1 ? g.Cin : g.Cout;
  if (slide != 1 && g.k * g.k * g.k * (slide / 8) > kMaxQ) return false;
  if (pass == 2) return true;                              // wgrad tiles always fit (<= 64 channels on both sides)
  // fprop / dgrad keep every B fragment in shared memory next to the halo tile: must fit with the shallow tile (R = 2)
  SmallParams p4, p2;
  memset(&p4, 0, sizeof(p4));
  memset(&p2, 0, sizeof(p2));
  plan_small(p4, 1, 8, 8, 16, slide, 8, 8, 16, outc, g.k, 0, 0, 4);
  plan_small(p2, 1, 8, 8, 16, slide, 8, 8, 16, outc, g.k, 0, 0, 2);
  int nt = 0;
  bool r4 = false;
  return pick_small_cfg(p4, p2, nt_for(outc), &nt, &r4);
}

int small_conv_fprop(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* w_oti, const float* bias,
                     __nv_bfloat16* y, double* ssum, double* ssq, cudaStream_t stream) {
  const int Do = oext(g.D, g.k, g.pad), Ho = oext(g.H, g.k, g.pad), Wo = oext(g.W, g.k, g.pad);
  SmallParams p4, p2;
  memset(&p4, 0, sizeof(p4));
  memset(&p2, 0, sizeof(p2));
  plan_small(p4, g.N, g.D, g.H, g.W, g.Cin, Do, Ho, Wo, g.Cout, g.k, g.pad, 0, 4);
  plan_small(p2, g.N, g.D, g.H, g.W, g.Cin, Do, Ho, Wo, g.Cout, g.k, g.pad, 0, 2);
  for (SmallParams* p : {&p4, &p2}) {
    p->in = x, p->w = w_oti, p->bias = bias, p->out = y, p->ssum = ssum, p->ssq = ssq;
  }
  const int NT = nt_for(g.Cout);
  if (g.Cin == 1 && NT == 1) {   // most tiles, least work per tile: 8 rows per warp (tile depth 8, halo 2.25x instead of 3x)
    SmallParams p8;
    memset(&p8, 0, sizeof(p8));
    const bool fits = plan_small(p8, g.N, g.D, g.H, g.W, g.Cin, Do, Ho, Wo, g.Cout, g.k, g.pad, 0, 8);
    p8.in = x, p8.w = w_oti, p8.bias = bias, p8.out = y, p8.ssum = ssum, p8.ssq = ssq;
    const size_t smem = fprop_smem(p8, 1);
    if (fits && smem <= kSmemLimit) return launch_fprop_t<1, 8, true>(p8, smem, 1, stream);
  }
  return g.Cin == 1 ? launch_fprop_nt<true>(p4, p2, NT, stream) : launch_fprop_nt<false>(p4, p2, NT, stream);
}

// dx[N,D,H,W,Cin] = sum_k dy[i + pad - k] w[k]^T: a forward conv over dy with mirrored taps and pad' = k - 1 - pad.
int small_conv_dgrad(const adni_conv3d_geom& g, const __nv_bfloat16* dy, const __nv_bfloat16* w_ito, __nv_bfloat16* dx,
                     cudaStream_t stream) {
  const int Do = oext(g.D, g.k, g.pad), Ho = oext(g.H, g.k, g.pad), Wo = oext(g.W, g.k, g.pad);
  SmallParams p4, p2;
  memset(&p4, 0, sizeof(p4));
  memset(&p2, 0, sizeof(p2));
  plan_small(p4, g.N, Do, Ho, Wo, g.Cout, g.D, g.H, g.W, g.Cin, g.k, g.k - 1 - g.pad, 1, 4);
  plan_small(p2, g.N, Do, Ho, Wo, g.Cout, g.D, g.H, g.W, g.Cin, g.k, g.k - 1 - g.pad, 1, 2);
  for (SmallParams* p : {&p4, &p2}) {
    p->in = dy, p->w = w_ito, p->bias = nullptr, p->out = dx, p->ssum = nullptr, p->ssq = nullptr;
  }
  return launch_fprop_nt<false>(p4, p2, nt_for(g.Cin), stream);
}

int small_conv_wgrad(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw,
                     cudaStream_t stream) {
  const int Do = oext(g.D, g.k, g.pad), Ho = oext(g.H, g.k, g.pad), Wo = oext(g.W, g.k, g.pad);
  SmallWgradParams wp;
  memset(&wp, 0, sizeof(wp));
  // the 1-channel first layer has the most tiles and the least work per tile: deeper tiles (8 planes) halve its
  // per-tile overheads and cut the halo from 3x to 2.25x
  int wg_td = (g.Cin == 1 && g.Cout == 8) ? 8 : 4;
  if (wg_td == 8 && !plan_small(wp.x, g.N, g.D, g.H, g.W, g.Cin, Do, Ho, Wo, g.Cout, g.k, g.pad, 0, 8)) wg_td = 4;   // k = 7
  if (!plan_small(wp.x, g.N, g.D, g.H, g.W, g.Cin, Do, Ho, Wo, g.Cout, g.k, g.pad, 0, wg_td)) {
    set_error("small conv wgrad: geometry not supported (Cin=%d Cout=%d k=%d)", g.Cin, g.Cout, g.k);
    return ADNI_ENOTSUP;
  }
  wp.x.in = x;
  wp.dy = dy;
  wp.dw = dw;
  // Cout = 8: the second half of the 16-row M tile reads the next position's chunk; those accumulator rows (co >= 8)
  // are never flushed.  Other channel counts are powers of two >= 16 (small_conv_supported).
  wp.CHo = g.Cout / 8;
  wp.chos = ilog2(wp.CHo);
  wp.sho = wp.CHo > 1 ? ilog2(8 / wp.CHo) : 0;
  wp.n_tiles_total = wp.x.Q;
  const bool c1 = g.Cin == 1;
  const int MT = (g.Cout + 15) / 16;
  // the first layer has k*k <= 49 N-tiles: 4 (k = 5) or 7 per warp keep the kernel at <= 64 registers, 4 CTAs per SM
  if (MT == 1 && c1 && wg_td == 8)
    return wp.x.Q <= 32 ? launch_wgrad_t<1, 4, true, 8>(wp, stream) : launch_wgrad_t<1, 8, true, 8>(wp, stream);
  if (MT == 1 && c1) return wp.x.Q <= 32 ? launch_wgrad_t<1, 4, true>(wp, stream) : launch_wgrad_t<1, 8, true>(wp, stream);
  if (MT == 1) return launch_wgrad_t<1, 16, false>(wp, stream);
  if (MT == 2) return c1 ? launch_wgrad_t<2, 8, true>(wp, stream) : launch_wgrad_t<2, 8, false>(wp, stream);
  return c1 ? launch_wgrad_t<4, 4, true>(wp, stream) : launch_wgrad_t<4, 4, false>(wp, stream);
}

}  // namespace adni

// tcgen05 / TMEM implicit-GEMM Conv3d kernels (fprop, dgrad; wgrad lives in conv_wgrad2.cu) fed by TMA box loads.
//
// Replaces the cuDNN Conv3d forward/backward dispatched by MedicalNet's ResNet (SURVEY.md K1-K3;
// reference call sites pkg/models/mri_models/anat_cnn.py:18-31,95).
//
// Formulation.  Activations are NDHWC bf16.  An output tile is a spatial BOX of <=128 output positions
// (bw x bh x bd) of one sample times BLOCK_N output channels.  For every filter tap the A operand of the
// GEMM is the same box shifted by (tap*dilation - pad): ONE tiled 5-D TMA load with the hardware
// zero-filling the out-of-bounds halo.  The box lands in shared memory as 128 rows x 128 B (64 channels),
// 128-byte swizzled: exactly the canonical K-major UMMA operand layout, so no thread ever touches operand
// data.  Stride-2 convs address the input through 8 parity-class tensor maps (doubled strides), dgrad is
// the same kernel on dY with the [Cin][tap][Cout] weight copy and mirrored tap offsets, and taps whose
// shifted box lies entirely in the padding (up to 31 % of layer4's dilation-4 taps) are skipped.
//
// Warp roles (192 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer
// + TMEM owner, warps 2-5 = epilogue (TMEM -> registers -> bf16 global, fused bias / residual-gradient
// add / BatchNorm sum & sum-of-squares).  Two TMEM accumulators let the epilogue of tile i overlap the
// mainloop of tile i+1.
#include <algorithm>

#include "conv_igemm.cuh"

namespace adni {

constexpr int kIgemmThreads = 192;

// explicit shared-space accesses for the epilogue staging block (pointers derived from the manually aligned dynamic
// shared-memory base are generic to the compiler: LD / ST instead of LDS / STS)
__device__ __forceinline__ void sts_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

struct TileCoord {
  int n, d0, h0, w0, n0;
};

template <int BLOCK_N>
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile) {
  TileCoord c;
  const int nt = tile % p.n_tiles;
  int m = tile / p.n_tiles;
  const int tw = m % p.tiles_w;
  m /= p.tiles_w;
  const int th = m % p.tiles_h;
  m /= p.tiles_h;
  const int td = m % p.tiles_d;
  c.n = m / p.tiles_d;
  c.d0 = td * p.bd;
  c.h0 = th * p.bh;
  c.w0 = tw * p.bw;
  c.n0 = nt * BLOCK_N;
  return c;
}

__device__ __forceinline__ bool box_in_range(const int* ext, int d, int h, int w, int bd, int bh, int bw) {
  return d + bd > 0 && d < ext[0] && h + bh > 0 && h < ext[1] && w + bw > 0 && w < ext[2];
}

template <int BLOCK_N, int STAGES>
struct IgemmCfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int STAT_OFF = BAR_OFF + 256;
  static constexpr int STAT_BYTES = 4 * 2 * BLOCK_N * 4;
  // epilogue staging: one 32-row x 32-column fp32 chunk per epilogue warp.  Pitch 36 floats keeps the three access
  // patterns conflict-free: STS.128 of a thread's own row, LDS.32 down a column (BatchNorm sums), LDS.128 of the
  // 8-column pieces of the coalesced store mapping.
  static constexpr int STG_OFF = STAT_OFF + STAT_BYTES;
  static constexpr int STG_PITCH = 36;
  static constexpr int STG_BYTES = 4 * 32 * STG_PITCH * 4;
  static constexpr int SMEM_BYTES = STG_OFF + STG_BYTES + 1024;  // + slack for manual 1024-B alignment
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
};

template <int BLOCK_N, int STAGES>
__device__ __forceinline__ void igemm_kmajor_body(const IgemmParams& p) {
  pdl_trigger();   // PDL (common.cuh): the next kernel of the stream may be scheduled once every CTA of this grid has started
  using Cfg = IgemmCfg<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* stat_smem = reinterpret_cast<float*>(smem + Cfg::STAT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.N * p.tiles_d * p.tiles_h * p.tiles_w * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) {
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
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // barriers, TMEM and the role split are set up under the previous kernel's tail; global memory only from here on

  // 64-bit mask of the taps whose shifted box intersects the input for this tile (lane t tests taps t, t+32)
  auto tap_mask = [&](const TileCoord& c) -> unsigned long long {
    bool v0 = false, v1 = false;
    if (lane < p.ntaps) {
      const ConvTap tap = p.taps[lane];
      v0 = box_in_range(p.a_ext[tap.map], c.d0 + tap.dd, c.h0 + tap.dh, c.w0 + tap.dw, p.bd, p.bh, p.bw);
    }
    if (lane + 32 < p.ntaps) {
      const ConvTap tap = p.taps[lane + 32];
      v1 = box_in_range(p.a_ext[tap.map], c.d0 + tap.dd, c.h0 + tap.dh, c.w0 + tap.dw, p.bd, p.bh, p.bw);
    }
    const unsigned long long lo = __ballot_sync(0xffffffffu, v0), hi = __ballot_sync(0xffffffffu, v1);
    return lo | (hi << 32);
  };

  if (warp == 0) {
    // ===================== TMA producer (lane 0: A box, lane 1: weight tile) =====================
    if (lane == 0) {
      tma_prefetch_desc(&p.a_maps[0]);
      tma_prefetch_desc(&p.b_map);
    }
    const uint32_t tx_bytes = static_cast<uint32_t>(p.bw * p.bh * p.bd) * 128u + Cfg::B_BYTES;
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord c = decode_tile<BLOCK_N>(p, tile);
      unsigned long long mask = tap_mask(c);
      while (mask) {
        const int t = __ffsll(static_cast<long long>(mask)) - 1;
        mask &= mask - 1;
        const ConvTap tap = p.taps[t];
        const int d = c.d0 + tap.dd, h = c.h0 + tap.dh, w = c.w0 + tap.dw;
        const CUtensorMap* amap = &p.a_maps[tap.map];
        for (int kb = 0; kb < p.kc_blocks; kb++) {
          if (lane == 0) {
            mbar_wait_spin(&empty[st], ph ^ 1, 2136);
            if (p.debug == 2)
              mbar_arrive(&full[st]);
            else
              mbar_arrive_expect_tx(&full[st], tx_bytes);
          }
          __syncwarp();
          if (p.debug != 2) {
            if (lane == 0) tma_load_5d(smem_a + st * Cfg::A_BYTES, amap, &full[st], kb * 64, w, h, d, c.n);
            if (lane == 1) tma_load_2d(smem_b + st * Cfg::B_BYTES, &p.b_map, &full[st], tap.kofs + kb * 64, c.n0);
          }
          if (++st == STAGES) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (lane 0 issues; the warp computes the tap mask) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, false, false);
    const uint64_t desc_hi = umma_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo0 = static_cast<uint32_t>(umma_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFFull);
    int st = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord c = decode_tile<BLOCK_N>(p, tile);
      const int nkb = __popcll(tap_mask(c)) * p.kc_blocks;
      // The whole warp walks the loop with warp-uniform state (so ptxas keeps stage / descriptor arithmetic on the
      // uniform datapath instead of ELECT + R2UR per operand); one elected lane issues the tcgen05 instructions.
      mbar_wait_spin(&tempty[acc], accph ^ 1, 2168);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
      for (int i = 0; i < nkb; i++) {
        mbar_wait_spin(&full[st], ph, 2172);
        tc_fence_after();
        const uint32_t a_lo = desc_lo0 + ((smem_u32(smem_a + st * Cfg::A_BYTES) & 0x3FFFFu) >> 4);
        const uint32_t b_lo = desc_lo0 + ((smem_u32(smem_b + st * Cfg::B_BYTES) & 0x3FFFFu) >> 4);
        if (elect_one_sync()) {
          if (p.debug != 1) {
#pragma unroll
            for (int k = 0; k < 4; k++)
              umma_bf16(d_tmem, desc_hi | (a_lo + k * 2), desc_hi | (b_lo + k * 2), idesc,
                        static_cast<uint32_t>(i | k));
          }
          umma_commit(&empty[st]);  // frees the smem slot once these MMAs have read it
        }
        __syncwarp();
        if (++st == STAGES) {
          st = 0;
          ph ^= 1;
        }
      }
      if (elect_one_sync()) umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == 2) {
        acc = 0;
        accph ^= 1;
      }
    }
  } else {
    // ===================== Epilogue (warps 2..5) =====================
    // A thread owns one accumulator row (TMEM lane).  Per 32-column chunk the warp stages its 32 x 32 fp32 block in
    // shared memory; from there (a) lane l sums column l (the fused BatchNorm statistics: 32 LDS instead of a 62-shuffle
    // transposing butterfly) and (b) the block leaves as bf16 with FOUR lanes per row (8 rows x 64 contiguous bytes per
    // store instruction instead of 32 rows x 16 bytes).  The per-channel sums are carried per CTA in registers and
    // flushed with one fp64 atomic per channel when the CTA's channel tile changes (once per CTA for the usual tile
    // counts): 1x1x1 convs have main loops of 1-16 K blocks per tile, so this epilogue - not the tensor pipe - is what
    // their tile rate is made of (ResNet-50, profiles/r02_shapes_r50_*.json).
    constexpr int NCH = BLOCK_N / 32;
    constexpr int NCOL = (BLOCK_N + 127) / 128;
    constexpr int PITCH = Cfg::STG_PITCH;
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;  // 0..127
    const int row = q * 32 + lane;
    const bool do_red = p.red_y != nullptr;                       // BatchNorm-backward sums (dgrad calls)
    const bool do_stats = p.stat_sum != nullptr && !do_red;       // BatchNorm-forward sums (fprop calls)
    const uint32_t stg = smem_u32(smem + Cfg::STG_OFF) + static_cast<uint32_t>(ew * (32 * PITCH) * 4);   // byte address
    const uint32_t stat_u32 = smem_u32(stat_smem);
    // box coordinates of this thread's accumulator row and of the four rows it stores (tile independent)
    const int rw = row % p.bw;
    const int rh = (row / p.bw) % p.bh;
    const int rd = row / (p.bw * p.bh);
    const int srow = lane >> 2, scol = (lane & 3) * 8;
    int sw_[4], sh_[4], sd_[4];
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int rr = q * 32 + it * 8 + srow;
      sw_[it] = rr % p.bw;
      sh_[it] = (rr / p.bw) % p.bh;
      sd_[it] = rr / (p.bw * p.bh);
    }
    double acc_s[NCOL], acc_q[NCOL];
#pragma unroll
    for (int i = 0; i < NCOL; i++) acc_s[i] = acc_q[i] = 0.0;
    int acc_n0 = -1;
    auto flush_sums = [&]() {
      if (acc_n0 < 0) return;
#pragma unroll
      for (int i = 0; i < NCOL; i++) {
        const int col = et + i * 128;
        if (col < BLOCK_N) {
          atomicAdd(p.stat_sum + acc_n0 + col, acc_s[i]);
          atomicAdd(p.stat_sq + acc_n0 + col, acc_q[i]);
        }
        acc_s[i] = acc_q[i] = 0.0;
      }
    };
    int acc = 0;
    uint32_t accph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord c = decode_tile<BLOCK_N>(p, tile);
      const bool has_k = tap_mask(c) != 0ull;
      const int od = c.d0 + rd, oh = c.h0 + rh, ow = c.w0 + rw;
      const bool valid = rd < p.bd && od < p.Do && oh < p.Ho && ow < p.Wo;
      const long long off = c.n * p.out_sn + od * p.out_sd + oh * p.out_sh + ow * p.out_sw + c.n0;
      if ((do_stats || do_red) && c.n0 != acc_n0) {
        flush_sums();
        acc_n0 = c.n0;
      }

      if (do_red) {
        // fused BatchNorm-backward sums: the y / mask rows of chunk c+1 are in flight while chunk c is processed, and
        // those of chunk 0 while this thread still waits for the accumulator (they do not depend on the MMA result)
        uint4 y_nxt[4], m_nxt[4];
        const bool red_row = valid;
        const bool red_mask = red_row && p.red_mask != nullptr;
        auto red_prefetch = [&](int chunk) {
          if (red_row) {
            const uint4* yp = reinterpret_cast<const uint4*>(p.red_y + off + chunk * 32);
#pragma unroll
            for (int j4 = 0; j4 < 4; j4++) y_nxt[j4] = __ldg(yp + j4);
            if (red_mask) {
              const uint4* mp = reinterpret_cast<const uint4*>(p.red_mask + off + chunk * 32);
#pragma unroll
              for (int j4 = 0; j4 < 4; j4++) m_nxt[j4] = __ldg(mp + j4);
            }
          }
        };
        red_prefetch(0);
        mbar_wait_spin(&tfull[acc], accph, 2217);
        tc_fence_after();
#pragma unroll 1
        for (int chunk = 0; chunk < BLOCK_N / 32; chunk++) {
          uint4 y_cur[4], m_cur[4];
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            y_cur[j4] = y_nxt[j4];
            m_cur[j4] = m_nxt[j4];
          }
          if (chunk + 1 < BLOCK_N / 32) red_prefetch(chunk + 1);
          uint32_t v[32];
          if (has_k) {
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                              static_cast<uint32_t>(acc * BLOCK_N + chunk * 32),
                          v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = 0u;
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] += __ldg(p.bias + c.n0 + chunk * 32 + j);
          }
          uint32_t packed[16];
          if (valid) {
            if (p.addend != nullptr) {
              const uint4* ap = reinterpret_cast<const uint4*>(p.addend + off + chunk * 32);
#pragma unroll
              for (int j4 = 0; j4 < 4; j4++) {
                const uint4 a = __ldg(ap + j4);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  f[j4 * 8 + e * 2 + 0] += bf16_lo(aw[e]);
                  f[j4 * 8 + e * 2 + 1] += bf16_hi(aw[e]);
                }
              }
            }
#pragma unroll
            for (int j2 = 0; j2 < 16; j2++) packed[j2] = pack_bf16x2(f[2 * j2], f[2 * j2 + 1]);
            if (p.debug != 3) {
              uint4* op = reinterpret_cast<uint4*>(p.out + off + chunk * 32);
#pragma unroll
              for (int j4 = 0; j4 < 4; j4++) op[j4] = make_uint4(packed[4 * j4], packed[4 * j4 + 1], packed[4 * j4 + 2], packed[4 * j4 + 3]);
            }
          }
          {
            // sums of the STORED gradient (bf16), masked by the preceding layer's ReLU
            float s1[32], s2[32];
            if (valid) {
#pragma unroll
              for (int j4 = 0; j4 < 4; j4++) {
                const uint32_t yw[4] = {y_cur[j4].x, y_cur[j4].y, y_cur[j4].z, y_cur[j4].w};
                const uint32_t mw[4] = {m_cur[j4].x, m_cur[j4].y, m_cur[j4].z, m_cur[j4].w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
#pragma unroll
                  for (int h = 0; h < 2; h++) {
                    const int j = j4 * 8 + e * 2 + h;
                    const float y = h ? bf16_hi(yw[e]) : bf16_lo(yw[e]);
                    float g = h ? bf16_hi(packed[j >> 1]) : bf16_lo(packed[j >> 1]);
                    if (red_mask) {
                      g = (h ? bf16_hi(mw[e]) : bf16_lo(mw[e])) > 0.f ? g : 0.f;
                    } else if (p.red_scale != nullptr) {
                      const int ch = c.n0 + chunk * 32 + j;
                      g = fmaf(y, __ldg(p.red_scale + ch), __ldg(p.red_shift + ch)) > 0.f ? g : 0.f;
                    }
                    s1[j] = g;
                    s2[j] = g * y;
                  }
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++) s1[j] = s2[j] = 0.f;
            }
            const float cs1 = warp_column_sums(s1, lane);
            const float cs2 = warp_column_sums(s2, lane);
            stat_smem[(ew * 2 + 0) * BLOCK_N + chunk * 32 + lane] = cs1;
            stat_smem[(ew * 2 + 1) * BLOCK_N + chunk * 32 + lane] = cs2;
          }
        }
        // accumulator drained -> hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      } else {
        long long soff[4];
        bool sval[4];
#pragma unroll
        for (int it = 0; it < 4; it++) {
          const int d = c.d0 + sd_[it], h = c.h0 + sh_[it], w = c.w0 + sw_[it];
          sval[it] = sd_[it] < p.bd && d < p.Do && h < p.Ho && w < p.Wo;
          soff[it] = c.n * p.out_sn + d * p.out_sd + h * p.out_sh + w * p.out_sw + c.n0 + scol;
        }
        // residual-gradient rows (dgrad addend) in the store mapping: chunk c + 1 is in flight while chunk c is processed,
        // chunk 0 while this thread still waits for the accumulator
        uint4 ad_nxt[4];
        auto addend_prefetch = [&](int chunk) {
          if (p.addend != nullptr) {
#pragma unroll
            for (int it = 0; it < 4; it++)
              if (sval[it]) ad_nxt[it] = __ldg(reinterpret_cast<const uint4*>(p.addend + soff[it] + chunk * 32));
          }
        };
        addend_prefetch(0);
        mbar_wait_spin(&tfull[acc], accph, 2217);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
        uint32_t v[32];
        if (has_k) tmem_ld_32x32(t_row, v);   // chunk c + 1 is in flight while chunk c is processed
#pragma unroll 1
        for (int chunk = 0; chunk < NCH; chunk++) {
          float f[32];
          if (has_k) {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]);
            if (chunk + 1 < NCH) {
              tmem_ld_32x32(t_row + static_cast<uint32_t>((chunk + 1) * 32), v);
            } else {
              // accumulator drained -> hand the TMEM buffer back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[acc]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] = 0.f;
            if (chunk + 1 == NCH) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[acc]);
            }
          }
          uint4 ad_cur[4];
#pragma unroll
          for (int it = 0; it < 4; it++) ad_cur[it] = ad_nxt[it];
          if (chunk + 1 < NCH) addend_prefetch(chunk + 1);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] += __ldg(p.bias + c.n0 + chunk * 32 + j);
          }
          // rows outside the output are staged as zeros: they must not count in the statistics (and are never stored)
#pragma unroll
          for (int j4 = 0; j4 < 8; j4++) {
            const uint32_t a = stg + static_cast<uint32_t>((lane * PITCH + j4 * 4) * 4);
            if (valid) sts_f4(a, f[j4 * 4], f[j4 * 4 + 1], f[j4 * 4 + 2], f[j4 * 4 + 3]);
            else sts_f4(a, 0.f, 0.f, 0.f, 0.f);
          }
          __syncwarp();
          if (do_stats) {
            float cs1 = 0.f, cs2 = 0.f, cs1b = 0.f, cs2b = 0.f;   // two chains: the 32 loads are independent
#pragma unroll
            for (int r = 0; r < 32; r += 2) {
              const float x0 = lds_f1(stg + static_cast<uint32_t>((r * PITCH + lane) * 4));
              const float x1 = lds_f1(stg + static_cast<uint32_t>(((r + 1) * PITCH + lane) * 4));
              cs1 += x0;
              cs2 = fmaf(x0, x0, cs2);
              cs1b += x1;
              cs2b = fmaf(x1, x1, cs2b);
            }
            sts_f1(stat_u32 + static_cast<uint32_t>(((ew * 2 + 0) * BLOCK_N + chunk * 32 + lane) * 4), cs1 + cs1b);
            sts_f1(stat_u32 + static_cast<uint32_t>(((ew * 2 + 1) * BLOCK_N + chunk * 32 + lane) * 4), cs2 + cs2b);
          }
#pragma unroll
          for (int it = 0; it < 4; it++) {
            if (sval[it]) {
              const uint32_t sp = stg + static_cast<uint32_t>(((it * 8 + srow) * PITCH + scol) * 4);
              const float4 a = lds_f4(sp);
              const float4 b = lds_f4(sp + 16);
              float o[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
              if (p.addend != nullptr) {
                const uint32_t aw[4] = {ad_cur[it].x, ad_cur[it].y, ad_cur[it].z, ad_cur[it].w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  o[e * 2 + 0] += bf16_lo(aw[e]);
                  o[e * 2 + 1] += bf16_hi(aw[e]);
                }
              }
              if (p.debug != 3)
                *reinterpret_cast<uint4*>(p.out + soff[it] + chunk * 32) =
                    make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
            }
          }
          __syncwarp();   // the staging block is rewritten by the next chunk
        }
      }
      if (++acc == 2) {
        acc = 0;
        accph ^= 1;
      }
      if (do_stats || do_red) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int i = 0; i < NCOL; i++) {
          const int col = et + i * 128;
          if (col < BLOCK_N) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int w4 = 0; w4 < 4; w4++) {
              a += stat_smem[(w4 * 2 + 0) * BLOCK_N + col];
              b += stat_smem[(w4 * 2 + 1) * BLOCK_N + col];
            }
            acc_s[i] += static_cast<double>(a);
            acc_q[i] += static_cast<double>(b);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
    if (do_stats || do_red) flush_sums();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kIgemmThreads, 1) igemm_kmajor_kernel(const __grid_constant__ IgemmParams p) {
  igemm_kmajor_body<BLOCK_N, STAGES>(p);
}

// Several independent problems of one shape family in ONE launch: blockIdx.y selects the problem, the CTAs of a row
// walk that problem's tiles.  A stride-2 dgrad is 8 such problems (the parity classes of dx, each with its own tap
// subset and output view); as 8 launches of ~15 us each they cost 0.25 ms per layer2.0-type block at the 8-GPU per-rank
// batch whatever the batch size (profiles/r02_b4_launch_summary.md).
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kIgemmThreads, 1) igemm_kmajor_multi_kernel(const __grid_constant__ IgemmMulti pm) {
  igemm_kmajor_body<BLOCK_N, STAGES>(pm.cls[blockIdx.y]);
}

// =================================================================================================
// Launchers
// =================================================================================================
extern void count_launch();

template <int BLOCK_N, int STAGES>
static int launch_igemm_t(const IgemmParams& p, cudaStream_t stream) {
  using Cfg = IgemmCfg<BLOCK_N, STAGES>;
  auto kern = igemm_kmajor_kernel<BLOCK_N, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int total_tiles = p.N * p.tiles_d * p.tiles_h * p.tiles_w * p.n_tiles;
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  pdl_launch(kern, grid, kIgemmThreads, Cfg::SMEM_BYTES, stream)(p);
  count_launch();
  ADNI_LAUNCH_CHECK("igemm_kmajor_kernel");
  return ADNI_OK;
}

template <int BLOCK_N, int STAGES>
static int launch_igemm_multi_t(const IgemmMulti& pm, cudaStream_t stream) {
  using Cfg = IgemmCfg<BLOCK_N, STAGES>;
  auto kern = igemm_kmajor_multi_kernel<BLOCK_N, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  int max_tiles = 1;
  for (int c = 0; c < pm.ncls; c++) {
    const IgemmParams& p = pm.cls[c];
    max_tiles = std::max(max_tiles, p.N * p.tiles_d * p.tiles_h * p.tiles_w * p.n_tiles);
  }
  const int per_cls = std::max(1, std::min(max_tiles, num_sms() / pm.ncls));   // one resident wave over all problems
  pdl_launch(kern, dim3(per_cls, pm.ncls, 1), kIgemmThreads, Cfg::SMEM_BYTES, stream)(pm);
  count_launch();
  ADNI_LAUNCH_CHECK("igemm_kmajor_multi_kernel");
  return ADNI_OK;
}

int launch_igemm_multi(const IgemmMulti& pm, int block_n, cudaStream_t stream) {
  switch (block_n) {
    case 64:
      return launch_igemm_multi_t<64, 8>(pm, stream);
    case 128:
      return launch_igemm_multi_t<128, 6>(pm, stream);
    case 256:
      return launch_igemm_multi_t<256, 4>(pm, stream);
    default:
      set_error("igemm: unsupported BLOCK_N %d", block_n);
      return ADNI_ENOTSUP;
  }
}

int launch_igemm(const IgemmParams& p, int block_n, cudaStream_t stream) {
  switch (block_n) {
    case 64:
      return launch_igemm_t<64, 8>(p, stream);
    case 128:
      return launch_igemm_t<128, 6>(p, stream);
    case 256:
      return launch_igemm_t<256, 4>(p, stream);
    default:
      set_error("igemm: unsupported BLOCK_N %d", block_n);
      return ADNI_ENOTSUP;
  }
}

}  // namespace adni

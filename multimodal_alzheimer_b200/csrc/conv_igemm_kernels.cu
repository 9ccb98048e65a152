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

// n / d with the host-side multiplier floor(2^32 / d) + 1 (0 for d == 1): exact while n * d < 2^32 (IgemmParams::fd_mul)
__device__ __forceinline__ int fdiv(int n, uint32_t mul) {
  return mul ? static_cast<int>(__umulhi(static_cast<uint32_t>(n), mul)) : n;
}

template <int BLOCK_N>
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile) {
  TileCoord c;
  int m, nt, tw, th, td;
  if (p.fd_ok) {   // four divisions by run-time constants per tile and warp are ~80 instructions as real divisions
    m = fdiv(tile, p.fd_mul[0]);
    nt = tile - m * p.n_tiles;
    int q = fdiv(m, p.fd_mul[1]);
    tw = m - q * p.tiles_w;
    m = q;
    q = fdiv(m, p.fd_mul[2]);
    th = m - q * p.tiles_h;
    m = q;
    q = fdiv(m, p.fd_mul[3]);
    td = m - q * p.tiles_d;
    c.n = q;
  } else {
    nt = tile % p.n_tiles;
    m = tile / p.n_tiles;
    tw = m % p.tiles_w;
    m /= p.tiles_w;
    th = m % p.tiles_h;
    m /= p.tiles_h;
    td = m % p.tiles_d;
    c.n = m / p.tiles_d;
  }
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
  // epilogue staging right after the stages (1024-byte aligned: the TMA-store blocks of the flat path are 128-byte
  // swizzled); the fp32 staging of the box path and the four 4 KB bf16 blocks of the flat path share it
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  static constexpr int STG_PITCH = 36;
  static constexpr int STG_BYTES = 19 * 1024;   // >= 4 * 32 * STG_PITCH * 4 = 18432 and >= 4 * 4096
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int STAT_OFF = BAR_OFF + 256;
  static constexpr int STAT_BYTES = 4 * 2 * BLOCK_N * 4;
  // epilogue staging: one 32-row x 32-column fp32 chunk per epilogue warp.  Pitch 36 floats keeps the three access
  // patterns conflict-free: STS.128 of a thread's own row, LDS.32 down a column (BatchNorm sums), LDS.128 of the
  // 8-column pieces of the coalesced store mapping.
  static constexpr int SMEM_BYTES = STAT_OFF + STAT_BYTES + 1024;  // + slack for manual 1024-B alignment
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
};

__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// One 32-column chunk of the staged epilogue, common case (accumulator present, no bias): `v` = this thread's row.
template <int BLOCK_N, int PITCH, bool STATS, bool ADDEND>
__device__ __forceinline__ void epi_chunk_fast(const uint32_t (&v)[32], const uint4 (&ad)[4], int chunk, uint32_t stg, uint32_t stat0,
                                               int lane, bool all_valid, bool row_valid, unsigned sval_mask,
                                               __nv_bfloat16* const (&optr)[4]) {
  const uint32_t own = stg + static_cast<uint32_t>(lane * PITCH * 4);
  if (all_valid) {
#pragma unroll
    for (int j4 = 0; j4 < 8; j4++) sts_u4(own + j4 * 16, v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
  } else {   // rows outside the output are staged as zeros: they must not count in the statistics (and are never stored)
#pragma unroll
    for (int j4 = 0; j4 < 8; j4++)
      sts_u4(own + j4 * 16, row_valid ? v[j4 * 4] : 0u, row_valid ? v[j4 * 4 + 1] : 0u, row_valid ? v[j4 * 4 + 2] : 0u,
             row_valid ? v[j4 * 4 + 3] : 0u);
  }
  __syncwarp();
  if (STATS) {
    float cs1 = 0.f, cs2 = 0.f, cs1b = 0.f, cs2b = 0.f;   // two chains: the 32 loads are independent
    const uint32_t colp = stg + static_cast<uint32_t>(lane * 4);
#pragma unroll
    for (int r = 0; r < 32; r += 2) {
      const float x0 = lds_f1(colp + r * PITCH * 4);
      const float x1 = lds_f1(colp + (r + 1) * PITCH * 4);
      cs1 += x0;
      cs2 = fmaf(x0, x0, cs2);
      cs1b += x1;
      cs2b = fmaf(x1, x1, cs2b);
    }
    sts_f1(stat0 + static_cast<uint32_t>(chunk * 128), cs1 + cs1b);
    sts_f1(stat0 + static_cast<uint32_t>(BLOCK_N * 4 + chunk * 128), cs2 + cs2b);
  }
  const uint32_t rowp = stg + static_cast<uint32_t>(((lane >> 2) * PITCH + (lane & 3) * 8) * 4);
#pragma unroll
  for (int it = 0; it < 4; it++) {
    if (sval_mask & (1u << it)) {
      const float4 a = lds_f4(rowp + it * 8 * PITCH * 4);
      const float4 b = lds_f4(rowp + it * 8 * PITCH * 4 + 16);
      float o[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      if (ADDEND) {
        const uint32_t aw[4] = {ad[it].x, ad[it].y, ad[it].z, ad[it].w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          o[e * 2 + 0] += bf16_lo(aw[e]);
          o[e * 2 + 1] += bf16_hi(aw[e]);
        }
      }
      *reinterpret_cast<uint4*>(optr[it] + chunk * 32) =
          make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
  }
  __syncwarp();   // the staging block is rewritten by the next chunk
}

// One accumulator tile through the staged epilogue: two register sets alternate (the TMEM load of chunk c + 1 and its
// addend rows are in flight while chunk c is processed), the TMEM buffer is handed back after the last load.
template <int BLOCK_N, int PITCH, bool STATS, bool ADDEND>
__device__ __forceinline__ void epi_tile_fast(uint32_t t_row, uint32_t stg, uint32_t stat0, int lane, bool all_valid, bool row_valid,
                                              unsigned sval_mask, __nv_bfloat16* const (&optr)[4], const __nv_bfloat16* const (&aptr)[4],
                                              uint64_t* tfull_bar, uint32_t tfull_phase, uint64_t* tempty_bar) {
  constexpr int NCH = BLOCK_N / 32;
  static_assert(NCH % 2 == 0, "chunks are processed in pairs");
  uint4 ad_a[4], ad_b[4];
  auto addend_prefetch = [&](uint4 (&ad)[4], int chunk) {
    if (ADDEND) {
#pragma unroll
      for (int it = 0; it < 4; it++)
        if (sval_mask & (1u << it)) ad[it] = __ldg(reinterpret_cast<const uint4*>(aptr[it] + chunk * 32));
    }
  };
  addend_prefetch(ad_a, 0);   // does not depend on the accumulator
  mbar_wait_spin(tfull_bar, tfull_phase, 2217);
  tc_fence_after();
  uint32_t va[32], vb[32];
  tmem_ld_32x32(t_row, va);
#pragma unroll 1
  for (int chunk = 0; chunk < NCH; chunk += 2) {
    tmem_ld_wait();
    tmem_ld_32x32(t_row + static_cast<uint32_t>((chunk + 1) * 32), vb);
    addend_prefetch(ad_b, chunk + 1);
    epi_chunk_fast<BLOCK_N, PITCH, STATS, ADDEND>(va, ad_a, chunk, stg, stat0, lane, all_valid, row_valid, sval_mask, optr);
    tmem_ld_wait();
    if (chunk + 2 < NCH) {
      tmem_ld_32x32(t_row + static_cast<uint32_t>((chunk + 2) * 32), va);
      addend_prefetch(ad_a, chunk + 2);
    } else {   // accumulator drained -> hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
    epi_chunk_fast<BLOCK_N, PITCH, STATS, ADDEND>(vb, ad_b, chunk + 1, stg, stat0, lane, all_valid, row_valid, sval_mask, optr);
  }
}

__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u1(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// Flat (1x1x1, stride 1) epilogue of one accumulator tile, one warp = 32 consecutive positions.  Per 64-channel group:
// [addend block by TMA into the warp's staging block] -> two TMEM loads -> (+ addend) -> bf16 -> swizzled STS of the own
// row -> TMA store of the 32 x 64 block; the BatchNorm sums are column sums of the STAGED (stored) values, two columns
// per lane.  TMEM reads (64 B / clk / SM) and the ~1000 shared-memory wavefronts per tile are what is left of the
// ~4000 LSU wavefronts per tile of the register -> fp32 staging -> 16-byte global stores path, which held the
// ResNet-50 1x1x1 convs at 10-13 k cycles per tile whatever the MMA / TMA / store traffic (tools/k1_probe.py).
template <int BLOCK_N, bool STATS, bool ADDEND>
__device__ __forceinline__ void epi_tile_flat(const IgemmParams& p, uint32_t t_row, uint32_t buf, uint32_t stat0, int lane, int row0,
                                              int n0, uint64_t* tfull_bar, uint32_t tfull_phase, uint64_t* tempty_bar,
                                              uint64_t* add_bar, uint32_t& add_phase) {
  constexpr int NG = BLOCK_N / 64;
  const uint32_t own = buf + static_cast<uint32_t>(lane * 128);
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
  mbar_wait_spin(tfull_bar, tfull_phase, 2217);
  tc_fence_after();
#pragma unroll 1
  for (int g = 0; g < NG; g++) {
    const int col0 = n0 + g * 64;
    if (lane == 0) tma_store_wait_read();   // the previous block has left the staging buffer
    __syncwarp();
    if (ADDEND) {
      if (lane == 0) {
        mbar_arrive_expect_tx(add_bar, 32 * 128);
        tma_load_2d_u32(buf, &p.add_map, add_bar, col0, row0);
      }
    }
    uint32_t v0[32], v1[32];
    tmem_ld_32x32(t_row + static_cast<uint32_t>(g * 64), v0);
    tmem_ld_32x32(t_row + static_cast<uint32_t>(g * 64 + 32), v1);
    tmem_ld_wait();
    if (g + 1 == NG) {   // accumulator drained -> hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
    if (ADDEND) {
      mbar_wait_spin(add_bar, add_phase, 2218);
      add_phase ^= 1u;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint32_t a = own + ((static_cast<uint32_t>(j) ^ sw) << 4);
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; e++) f[e] = __uint_as_float(j < 4 ? v0[j * 8 + e] : v1[(j - 4) * 8 + e]);
      if (ADDEND) {
        const uint4 ad = lds_u4(a);
        const uint32_t aw[4] = {ad.x, ad.y, ad.z, ad.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          f[e * 2 + 0] += bf16_lo(aw[e]);
          f[e * 2 + 1] += bf16_hi(aw[e]);
        }
      }
      sts_u4(a, pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && p.debug != 3) {
      tma_store_2d_u32(&p.out_map, buf, col0, row0);
      tma_store_commit();
    }
    if (STATS) {
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
      const uint32_t colw = buf + static_cast<uint32_t>((lane & 3) * 4);
      const uint32_t ch = static_cast<uint32_t>(lane >> 2);
#pragma unroll
      for (int r = 0; r < 32; r++) {
        const uint32_t w = lds_u1(colw + static_cast<uint32_t>(r * 128) + ((ch ^ static_cast<uint32_t>(r & 7)) << 4));
        const float lo = bf16_lo(w), hi = bf16_hi(w);
        s0 += lo;
        q0 = fmaf(lo, lo, q0);
        s1 += hi;
        q1 = fmaf(hi, hi, q1);
      }
      // columns g*64 + 2*lane, + 1 of this warp's partial sums
      const uint32_t sa = stat0 + static_cast<uint32_t>((g * 64 + lane) * 4);   // stat0 already carries + lane * 4
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sa), "f"(s0), "f"(s1) : "memory");
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sa + static_cast<uint32_t>(BLOCK_N * 4)), "f"(q0), "f"(q1) : "memory");
    }
  }
}

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
  uint64_t* abar = tempty + 2;   // per epilogue warp: addend block landed (flat path)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(abar + 4);
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
    for (int i = 0; i < 4; i++) mbar_init(&abar[i], 1);
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
    long long srel[4];   // offset of the four stored rows relative to the tile origin (tile independent)
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int rr = q * 32 + it * 8 + srow;
      sw_[it] = rr % p.bw;
      sh_[it] = (rr / p.bw) % p.bh;
      sd_[it] = rr / (p.bw * p.bh);
      srel[it] = sd_[it] * p.out_sd + sh_[it] * p.out_sh + sw_[it] * p.out_sw + scol;
    }
    const bool fast_ok = p.bias == nullptr && p.debug != 3;
    const uint32_t flat_buf = smem_u32(smem + Cfg::STG_OFF) + static_cast<uint32_t>(ew * 4096);
    uint32_t add_phase = 0;
    const uint32_t stat0 = stat_u32 + static_cast<uint32_t>((ew * 2 * BLOCK_N + lane) * 4);
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

      if (p.flat) {
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
        const int row0 = c.w0 + q * 32;
        if (do_stats) {
          epi_tile_flat<BLOCK_N, true, false>(p, t_row, flat_buf, stat0, lane, row0, c.n0, &tfull[acc], accph, &tempty[acc], &abar[ew],
                                              add_phase);
        } else if (p.addend != nullptr) {
          epi_tile_flat<BLOCK_N, false, true>(p, t_row, flat_buf, stat0, lane, row0, c.n0, &tfull[acc], accph, &tempty[acc], &abar[ew],
                                              add_phase);
        } else {
          epi_tile_flat<BLOCK_N, false, false>(p, t_row, flat_buf, stat0, lane, row0, c.n0, &tfull[acc], accph, &tempty[acc], &abar[ew],
                                               add_phase);
        }
      } else if (do_red) {
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
      } else if (fast_ok && has_k) {
        // common case: accumulator present, no bias -> the specialised staged epilogue (no run-time switches inside)
        const long long tile_off = c.n * p.out_sn + c.d0 * p.out_sd + c.h0 * p.out_sh + c.w0 * p.out_sw + c.n0;
        const int lim_d = min(p.bd, p.Do - c.d0), lim_h = p.Ho - c.h0, lim_w = p.Wo - c.w0;
        unsigned sval_mask = 0;
        __nv_bfloat16* optr[4];
        const __nv_bfloat16* aptr[4];
#pragma unroll
        for (int it = 0; it < 4; it++) {
          if (sd_[it] < lim_d && sh_[it] < lim_h && sw_[it] < lim_w) sval_mask |= 1u << it;
          optr[it] = p.out + tile_off + srel[it];
          aptr[it] = p.addend + tile_off + srel[it];
        }
        const bool all_valid = __all_sync(0xffffffffu, valid);
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
        if (do_stats) {
          epi_tile_fast<BLOCK_N, PITCH, true, false>(t_row, stg, stat0, lane, all_valid, valid, sval_mask, optr, aptr, &tfull[acc],
                                                     accph, &tempty[acc]);
        } else if (p.addend != nullptr) {
          epi_tile_fast<BLOCK_N, PITCH, false, true>(t_row, stg, stat0, lane, all_valid, valid, sval_mask, optr, aptr, &tfull[acc],
                                                     accph, &tempty[acc]);
        } else {
          epi_tile_fast<BLOCK_N, PITCH, false, false>(t_row, stg, stat0, lane, all_valid, valid, sval_mask, optr, aptr, &tfull[acc],
                                                      accph, &tempty[acc]);
        }
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
              a += lds_f1(stat_u32 + static_cast<uint32_t>(((w4 * 2 + 0) * BLOCK_N + col) * 4));
              b += lds_f1(stat_u32 + static_cast<uint32_t>(((w4 * 2 + 1) * BLOCK_N + col) * 4));
            }
            acc_s[i] += static_cast<double>(a);
            acc_q[i] += static_cast<double>(b);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
    if (do_stats || do_red) flush_sums();
    if (p.flat && lane == 0) tma_store_wait_all();   // the staged blocks must outlive the bulk stores that read them
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

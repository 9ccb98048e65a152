// wgrad on tcgen05, second generation: one CTA owns the FULL TMEM (512 fp32 columns) as MT accumulators of
// 128 (Cout rows) x 512/MT (K_total columns), so that every dY box (the operand all CTAs read at the same time,
// i.e. an L2 broadcast) is reused for 512/MT columns and every X box (the operand that is unique per CTA and
// therefore bound by the ~5.5 TB/s of distinct L2->SM traffic measured on B200) is reused for ALL MT*128 output
// channels.  Measured motivation (profiles/r01_notes.md): the first-generation 128x256 tile moved 32 KB of unique
// X data per 128x256x64 MAC unit and ran at 2600 cycles per unit; this layout moves 8 KB (Cout = 512).
//
//   dW[Cout][K_total] += sum over position boxes (32 positions each)  dY[pos][Cout]^T * X[pos + tap][Cin slice]
// Operands are MN-major for this GEMM (channels contiguous): TMA boxes of 32 positions x 64 channels with the
// 128-byte swizzle are consumed through MN-major UMMA descriptors.
#include <type_traits>

#include "conv_igemm.cuh"

namespace adni {
extern void count_launch();

namespace {

constexpr int kThreads = 192;
// positions per TMA box = K of one pipeline stage.  MT = 2 (Cout >= 256, 8 boxes per stage) takes 64-position boxes:
// 8 MMAs per stage instead of 4 halves the barrier / commit overhead per MMA; the 10-box stages of MT = 1 / 4 keep 32
// (a 64-position stage would be 80 KB: only two stages deep).
constexpr int wgrad2_box_rows_c(int mt) { return mt == 2 ? 64 : 32; }

__device__ __forceinline__ bool box_hits(const int* ext, int d, int h, int w, int bd, int bh, int bw) {
  return d + bd > 0 && d < ext[0] && h + bh > 0 && h < ext[1] && w + bw > 0 && w < ext[2];
}

template <int MT>
struct W2Cfg {
  static constexpr int ROWS = wgrad2_box_rows_c(MT);
  static constexpr int BOX_BYTES = ROWS * 128;
  static constexpr int BLOCK_N = 512 / MT;
  static constexpr int NG = BLOCK_N / 64;                       // 64-column groups per tile
  static constexpr int NSUB = BLOCK_N > 256 ? 256 : BLOCK_N;    // N of one tcgen05.mma
  static constexpr int N_SUBS = BLOCK_N / NSUB;
  static constexpr int A_BOXES = 2 * MT;
  static constexpr int STAGE_BOXES = A_BOXES + NG;
  static constexpr int STAGE_BYTES = STAGE_BOXES * BOX_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
};

template <int NG>
struct ItemCtx {
  int map[NG], dd[NG], dh[NG], dw[NG], c0[NG];
  int ng, b_begin, b_end, mt_outer, g0;
};

// Work walk of a CTA: items blockIdx.x, + gridDim.x, ... of the static split-K schedule, or the tiles of its stream-K
// range (W2Sched).  `cur` is the item / tile index; ctx() resolves it to (N tile, M tile, position-box range).
template <bool SK, typename SchedT>
__device__ __forceinline__ void w2_range(const WgradParams& p, const SchedT& sk, int& cur, int& stop, int& step) {
  if constexpr (SK) {
    cur = sk.tile_begin[blockIdx.x];
    stop = sk.tile_last[blockIdx.x] + 1;
    step = 1;
  } else {
    cur = blockIdx.x;
    stop = p.m_tiles * p.n_tiles * p.splits;
    step = gridDim.x;
  }
}

template <int NG, bool SK, typename SchedT>
__device__ __forceinline__ ItemCtx<NG> make_ctx(const WgradParams& p, const SchedT& sk, int item) {
  ItemCtx<NG> x;
  const int nt = item % p.n_tiles;
  const int r = item / p.n_tiles;
  x.mt_outer = r % p.m_tiles;
  x.g0 = nt * NG;
  x.ng = min(NG, p.n_groups - x.g0);
  if constexpr (SK) {
    const int chunk0 = (r / p.m_tiles) * sk.chunk_boxes;   // virtual tile -> position chunk (W2Sched)
    x.b_begin = item == sk.tile_begin[blockIdx.x] ? sk.box_begin[blockIdx.x] : chunk0;
    x.b_end = item == sk.tile_last[blockIdx.x] ? sk.box_end[blockIdx.x] : min(chunk0 + sk.chunk_boxes, p.pos_boxes);
  } else {
    const int ks = r / p.m_tiles;
    x.b_begin = ks * p.boxes_per_split;
    x.b_end = min(x.b_begin + p.boxes_per_split, p.pos_boxes);
  }
#pragma unroll
  for (int g = 0; g < NG; g++) {
    const int gg = min(x.g0 + g, p.n_groups - 1);
    const int t = gg / p.cin_blocks;
    const ConvTap tap = p.taps[t];
    x.map[g] = tap.map;
    x.dd[g] = tap.dd;
    x.dh[g] = tap.dh;
    x.dw[g] = tap.dw;
    x.c0[g] = (gg - t * p.cin_blocks) * 64;
  }
  return x;
}

struct BoxWalk {
  int n, td, th, tw;
  __device__ __forceinline__ void init(const WgradParams& p, int b) {
    tw = b % p.tiles_w;
    b /= p.tiles_w;
    th = b % p.tiles_h;
    b /= p.tiles_h;
    td = b % p.tiles_d;
    n = b / p.tiles_d;
  }
  __device__ __forceinline__ void next(const WgradParams& p) {
    if (++tw == p.tiles_w) {
      tw = 0;
      if (++th == p.tiles_h) {
        th = 0;
        if (++td == p.tiles_d) {
          td = 0;
          ++n;
        }
      }
    }
  }
};

template <int NG>
__device__ __forceinline__ bool box_active(const WgradParams& p, const ItemCtx<NG>& x, int d0, int h0, int w0) {
  bool any = false;
#pragma unroll
  for (int g = 0; g < NG; g++)
    any |= (g < x.ng) && box_hits(p.x_ext[x.map[g]], d0 + x.dd[g], h0 + x.dh[g], w0 + x.dw[g], p.bd, p.bh, p.bw);
  return any;
}

template <int MT, bool SK>
__global__ void __launch_bounds__(kThreads, 1)
    wgrad2_kernel(const __grid_constant__ WgradParams p,
                  const __grid_constant__ typename std::conditional<SK, W2Sched, SkNone>::type sk) {
  pdl_trigger();   // PDL (common.cuh): the next kernel of the stream may be scheduled once every CTA of this grid has started
  using Cfg = W2Cfg<MT>;
  using SchedT = typename std::conditional<SK, W2Sched, SkNone>::type;
  constexpr int NG = Cfg::NG;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int kBoxRows = Cfg::ROWS;
  constexpr int kBoxBytes = Cfg::BOX_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
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
  pdl_wait();      // barriers, TMEM and the role split are set up under the previous kernel's tail; global memory only from here on

  if (warp == 0) {
    // ===================== TMA producer: lane l issues box l of the stage =====================
    // Box activity is evaluated 32 boxes at a time (lane i <-> box b0+i, one ballot); the coordinates of an active
    // box are then broadcast with shuffles, so the per-K-block path has no divisions and no table lookups.
    int st = 0;
    uint32_t ph = 0;
    int item, item_stop, item_step;
    w2_range<SK, SchedT>(p, sk, item, item_stop, item_step);
    for (; item < item_stop; item += item_step) {
      const ItemCtx<NG> x = make_ctx<NG, SK, SchedT>(p, sk, item);
      const uint32_t tx_bytes = static_cast<uint32_t>(Cfg::A_BOXES + x.ng) * kBoxBytes;
      const int g = lane - Cfg::A_BOXES;
      int my_map = 0, my_dd = 0, my_dh = 0, my_dw = 0, my_c0 = 0;
#pragma unroll
      for (int gi = 0; gi < NG; gi++)
        if (g == gi) {
          my_map = x.map[gi];
          my_dd = x.dd[gi];
          my_dh = x.dh[gi];
          my_dw = x.dw[gi];
          my_c0 = x.c0[gi];
        }
      const int co_base = x.mt_outer * (MT * 128);
      const CUtensorMap* my_tmap = lane < Cfg::A_BOXES ? &p.dy_map : &p.x_maps[my_map];
      const int my_c = lane < Cfg::A_BOXES ? co_base + lane * 64 : my_c0;
      const bool issuer = lane < Cfg::A_BOXES || g < x.ng;
      for (int b0 = x.b_begin; b0 < x.b_end; b0 += 32) {
        BoxWalk bi;
        bi.init(p, min(b0 + lane, p.pos_boxes - 1));
        const int d0 = bi.td * p.bd, h0 = bi.th * p.bh, w0 = bi.tw * p.bw;
        const bool act = (b0 + lane < x.b_end) && box_active<NG>(p, x, d0, h0, w0);
        uint32_t mask = __ballot_sync(0xffffffffu, act);
        while (mask) {
          const int src = __ffs(mask) - 1;
          mask &= mask - 1;
          const int bn = __shfl_sync(0xffffffffu, bi.n, src);
          const int bd0 = __shfl_sync(0xffffffffu, d0, src);
          const int bh0 = __shfl_sync(0xffffffffu, h0, src);
          const int bw0 = __shfl_sync(0xffffffffu, w0, src);
          if (lane == 0) {
            mbar_wait(&empty[st], ph ^ 1);
            mbar_arrive_expect_tx(&full[st], tx_bytes);
          }
          __syncwarp();
          if (issuer)
            tma_load_5d(smem + st * Cfg::STAGE_BYTES + lane * kBoxBytes, my_tmap, &full[st], my_c, bw0 + my_dw,
                        bh0 + my_dh, bd0 + my_dd, bn);
          if (++st == STAGES) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp evaluates activity, lane 0 issues) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, Cfg::NSUB, true, true);
    // descriptor high words are loop invariant; only the 14-bit start-address field changes
    const uint64_t desc_hi = umma_smem_desc_sw128(0, kBoxBytes, 1024) & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo_base = static_cast<uint32_t>(umma_smem_desc_sw128(0, kBoxBytes, 1024) & 0xFFFFFFFFull);
    int st = 0;
    uint32_t ph = 0, accph = 0;
    int item, item_stop, item_step;
    w2_range<SK, SchedT>(p, sk, item, item_stop, item_step);
    for (; item < item_stop; item += item_step) {
      const ItemCtx<NG> x = make_ctx<NG, SK, SchedT>(p, sk, item);
      if (lane == 0) {
        mbar_wait(tempty, accph ^ 1);
        tc_fence_after();
      }
      __syncwarp();
      uint32_t accum = 0;
      for (int b0 = x.b_begin; b0 < x.b_end; b0 += 32) {
        BoxWalk bi;
        bi.init(p, min(b0 + lane, p.pos_boxes - 1));
        const bool act = (b0 + lane < x.b_end) && box_active<NG>(p, x, bi.td * p.bd, bi.th * p.bh, bi.tw * p.bw);
        const int nact = __popc(__ballot_sync(0xffffffffu, act));
        if (lane == 0) {
          for (int i = 0; i < nact; i++) {
            mbar_wait(&full[st], ph);
            tc_fence_after();
            const uint32_t a_lo = desc_lo_base + ((smem_u32(smem + st * Cfg::STAGE_BYTES) & 0x3FFFFu) >> 4);
            const uint32_t b_lo = a_lo + ((Cfg::A_BOXES * kBoxBytes) >> 4);
#pragma unroll
            for (int k = 0; k < kBoxRows / 16; k++) {
#pragma unroll
              for (int mt = 0; mt < MT; mt++) {
                // A: the two 64-channel groups of M tile mt are kBoxBytes apart (LBO); 8-position K groups 1024 B (SBO)
                const uint64_t adesc = desc_hi | (a_lo + ((mt * 2 * kBoxBytes + k * 2048) >> 4));
#pragma unroll
                for (int ns = 0; ns < Cfg::N_SUBS; ns++) {
                  const uint64_t bdesc = desc_hi | (b_lo + ((ns * (Cfg::NSUB / 64) * kBoxBytes + k * 2048) >> 4));
                  umma_bf16(tmem_base + static_cast<uint32_t>(mt * Cfg::BLOCK_N + ns * Cfg::NSUB), adesc, bdesc, idesc,
                            accum | static_cast<uint32_t>(k));
                }
              }
            }
            accum = 1;
            umma_commit(&empty[st]);
            if (++st == STAGES) {
              st = 0;
              ph ^= 1;
            }
          }
        }
        st = __shfl_sync(0xffffffffu, st, 0);
        ph = __shfl_sync(0xffffffffu, ph, 0);
        accum = __shfl_sync(0xffffffffu, accum, 0);
      }
      if (lane == 0) umma_commit(tfull);
      accph ^= 1;
    }
  } else {
    // ===================== Epilogue: TMEM -> red.global.add =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t accph = 0;
    int item, item_stop, item_step;
    w2_range<SK, SchedT>(p, sk, item, item_stop, item_step);
    for (; item < item_stop; item += item_step) {
      const ItemCtx<NG> x = make_ctx<NG, SK, SchedT>(p, sk, item);
      bool has_k = false;
      {
        BoxWalk bi;
        bi.init(p, x.b_begin);
        for (int b = x.b_begin; b < x.b_end && !has_k; b++, bi.next(p))
          has_k = box_active<NG>(p, x, bi.td * p.bd, bi.th * p.bh, bi.tw * p.bw);
      }
      mbar_wait(tfull, accph);
      tc_fence_after();
      if (has_k) {
#pragma unroll 1
        for (int mt = 0; mt < MT; mt++) {
          const int co = x.mt_outer * (MT * 128) + mt * 128 + row;
          float* dst = p.dw + static_cast<long long>(co) * p.k_total + static_cast<long long>(x.g0) * 64;
#pragma unroll 1
          for (int chunk = 0; chunk < x.ng * 2; chunk++) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                              static_cast<uint32_t>(mt * Cfg::BLOCK_N + chunk * 32),
                          v);
            tmem_ld_wait();
            if (co < p.cout) {
#pragma unroll
              for (int j4 = 0; j4 < 8; j4++) {
                float4 val;
                val.x = __uint_as_float(v[j4 * 4 + 0]);
                val.y = __uint_as_float(v[j4 * 4 + 1]);
                val.z = __uint_as_float(v[j4 * 4 + 2]);
                val.w = __uint_as_float(v[j4 * 4 + 3]);
                atomicAdd(reinterpret_cast<float4*>(dst + chunk * 32 + j4 * 4), val);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      accph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int MT>
int launch_t(const WgradParams& p, const W2Sched* sk, cudaStream_t stream) {
  using Cfg = W2Cfg<MT>;
  if (sk != nullptr) {
    auto kern = wgrad2_kernel<MT, true>;
    static bool attr_set = false;
    if (!attr_set) {
      ADNI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
      attr_set = true;
    }
    pdl_launch(kern, sk->ctas, kThreads, Cfg::SMEM_BYTES, stream)(p, *sk);
  } else {
    auto kern = wgrad2_kernel<MT, false>;
    static bool attr_set = false;
    if (!attr_set) {
      ADNI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
      attr_set = true;
    }
    const int total = p.m_tiles * p.n_tiles * p.splits;
    const int grid = total < num_sms() ? total : num_sms();
    pdl_launch(kern, grid, kThreads, Cfg::SMEM_BYTES, stream)(p, SkNone{0});
  }
  count_launch();
  ADNI_LAUNCH_CHECK("wgrad2_kernel");
  return ADNI_OK;
}

}  // namespace

// mt_cfg: accumulator rows per CTA / 128 (1, 2 or 4)
int wgrad2_groups_per_tile(int mt_cfg) { return 8 / mt_cfg; }
int wgrad2_box_rows(int mt_cfg) { return wgrad2_box_rows_c(mt_cfg); }

int launch_wgrad2(const WgradParams& p, int mt_cfg, cudaStream_t stream, const W2Sched* sk) {
  switch (mt_cfg) {
    case 1:
      return launch_t<1>(p, sk, stream);
    case 2:
      return launch_t<2>(p, sk, stream);
    case 4:
      return launch_t<4>(p, sk, stream);
    default:
      set_error("wgrad2: unsupported M-tile count %d", mt_cfg);
      return ADNI_ENOTSUP;
  }
}

}  // namespace adni

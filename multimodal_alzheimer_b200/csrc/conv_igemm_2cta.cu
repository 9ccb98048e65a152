// EXPERIMENTAL (opt-in: ADNI_IGEMM_2CTA=1; not on the default path.  On a B200: outputs bit-identical to the 1-CTA
// kernel, conv parity tests green, 0-3.4 % faster on layers 3-4 - profiles/r01_2cta_ab.json):
// CTA-pair variant of igemm_kmajor_kernel for the N = 256 convs of layers 3-4 (reference call sites
// pkg/models/mri_models/anat_cnn.py:18-31,95 -> MedicalNet layer3/layer4 3x3x3 dilated convs, cuDNN there).
//
// Why: the 1-CTA kernel moves 48 KB of operands per 64-deep K block and SM (16 KB activation box + 32 KB = 256 weight
// rows) and is bound by distinct L2->SM traffic at ~0.73 of the tensor peak in executed FLOPs.  With
// tcgen05.mma.cta_group::2 the two CTAs of a cluster (one TPC) compute ONE 256 x 256 tile: each CTA stages its own
// 128-position box (A) and only HALF of the weight tile (128 of the 256 B rows, 16 KB); the tensor cores of both SMs
// read both halves.  32 KB per K block and SM: -33 % operand traffic for the same math.
//
// Structure (per CTA, rank r = %cluster_ctarank; 192 threads as in the 1-CTA kernel):
//   warp 0  TMA producer: A box of M tile (2*pair + r) and B rows [n0 + 128 r, n0 + 128 r + 128) into its own smem,
//           cp.async.bulk.tensor...cta_group::2 completing transaction bytes on the LEADER's (rank 0) full barrier.
//   warp 1  rank 0: MMA issuer - waits its full barrier (expects both CTAs' bytes), issues tcgen05.mma.cta_group::2
//           M = 256, N = 256, K = 16 (operand descriptors = the same smem offsets in both CTAs), frees the stage in
//           BOTH CTAs with tcgen05.commit...multicast::cluster; rank 1: only owns the TMEM allocation of its SM.
//   warps 2-5  epilogue of the CTA's own 128 rows (own TMEM lanes), identical to the 1-CTA kernel; hands the
//           accumulator back by arriving on the leader's tempty barrier (8 arrivals: 4 warps x 2 CTAs).
// The two M tiles of a pair are the same spatial box of two consecutive samples: identical tap masks, so both
// producers push the same sequence of stages without executing taps that are all padding for one of them.
#include "conv_igemm.cuh"

namespace adni {
extern void count_launch();

namespace {

constexpr int k2Threads = 192;
constexpr int k2BlockN = 256;
constexpr int k2HalfN = 128;
constexpr int k2Stages = 6;

struct Cfg2 {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = k2HalfN * 128;
  static constexpr int BAR_OFF = k2Stages * (A_BYTES + B_BYTES);
  static constexpr int STAT_OFF = BAR_OFF + 256;
  static constexpr int STAT_BYTES = 4 * 2 * k2BlockN * 4;
  static constexpr int SMEM_BYTES = STAT_OFF + STAT_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * k2BlockN;  // two accumulators of 256 fp32 columns, in both SMs
};
static_assert(Cfg2::SMEM_BYTES <= 227 * 1024, "2-CTA igemm: shared memory budget");

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t n_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t d;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(saddr), "r"(rank));
  return d;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data into the executing CTA's smem, transaction bytes onto the barrier at `bar_cluster`
// (a shared::cluster address: the leader's full barrier).
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem2_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive (once the MMAs issued so far retire) on the barrier at this smem offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma2_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Tile2 {
  int n, d0, h0, w0;
};
// M tile index -> (sample, box origin); the tile order is the 1-CTA kernel's with the channel tile factored out
__device__ __forceinline__ Tile2 decode_mtile(const IgemmParams& p, int m) {
  Tile2 c;
  const int tw = m % p.tiles_w;
  m /= p.tiles_w;
  const int th = m % p.tiles_h;
  m /= p.tiles_h;
  const int td = m % p.tiles_d;
  c.n = m / p.tiles_d;
  c.d0 = td * p.bd;
  c.h0 = th * p.bh;
  c.w0 = tw * p.bw;
  return c;
}
__device__ __forceinline__ bool box_hits(const int* ext, int d, int h, int w, int bd, int bh, int bw) {
  return d + bd > 0 && d < ext[0] && h + bh > 0 && h < ext[1] && w + bw > 0 && w < ext[2];
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
    igemm_kmajor_2cta_kernel(const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + k2Stages * Cfg2::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg2::BAR_OFF);  // used in the leader only
  uint64_t* empty = full + k2Stages;                                    // one per CTA (multicast commit)
  uint64_t* tfull = empty + k2Stages;                                   // one per CTA (multicast commit)
  uint64_t* tempty = tfull + 2;                                         // used in the leader only (8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* stat_smem = reinterpret_cast<float*>(smem + Cfg2::STAT_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int sp_tiles = p.tiles_d * p.tiles_h * p.tiles_w;  // spatial tiles of one sample
  const int total_pairs = (p.N / 2) * sp_tiles * p.n_tiles;  // N is even (checked on the host)
  const int pair0 = static_cast<int>(cluster_id_x());
  const int pair_step = static_cast<int>(n_clusters_x());

  if (threadIdx.x == 0) {
    for (int i = 0; i < k2Stages; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {  // the same logical warp of both CTAs allocates the pair's TMEM columns
    tmem2_alloc(tmem_slot, Cfg2::TMEM_COLS);
    tmem2_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // taps whose shifted box intersects the input for EITHER M tile of the pair (lane t tests taps t, t + 32)
  auto tap_mask = [&](const Tile2& a, const Tile2& b) -> unsigned long long {
    bool v0 = false, v1 = false;
    if (lane < p.ntaps) {
      const ConvTap tap = p.taps[lane];
      const int* ext = p.a_ext[tap.map];
      v0 = box_hits(ext, a.d0 + tap.dd, a.h0 + tap.dh, a.w0 + tap.dw, p.bd, p.bh, p.bw) ||
           box_hits(ext, b.d0 + tap.dd, b.h0 + tap.dh, b.w0 + tap.dw, p.bd, p.bh, p.bw);
    }
    if (lane + 32 < p.ntaps) {
      const ConvTap tap = p.taps[lane + 32];
      const int* ext = p.a_ext[tap.map];
      v1 = box_hits(ext, a.d0 + tap.dd, a.h0 + tap.dh, a.w0 + tap.dw, p.bd, p.bh, p.bw) ||
           box_hits(ext, b.d0 + tap.dd, b.h0 + tap.dh, b.w0 + tap.dw, p.bd, p.bh, p.bw);
    }
    const unsigned long long lo = __ballot_sync(0xffffffffu, v0), hi = __ballot_sync(0xffffffffu, v1);
    return lo | (hi << 32);
  };
  // pair index -> channel tile (fastest, like the 1-CTA tile order) and the two M tiles: the SAME spatial box of two
  // consecutive samples, so that both have the same tap mask (pairing neighbouring boxes of one sample made the pair
  // execute the union of two different masks: +13 % MMAs on layer4's dilation-4 convs, measured slower than 1 CTA)
  auto decode_pair = [&](int pair, int& n0, Tile2& t0, Tile2& t1) {
    const int nt = pair % p.n_tiles;
    const int mp = pair / p.n_tiles;
    const int sp = mp % sp_tiles;
    const int np = mp / sp_tiles;
    n0 = nt * k2BlockN;
    t0 = decode_mtile(p, (2 * np) * sp_tiles + sp);
    t1 = decode_mtile(p, (2 * np + 1) * sp_tiles + sp);
  };

  if (warp == 0) {
    // ===================== TMA producer (lane 0: own A box, lane 1: own half of the weight tile) ================
    if (lane == 0) {
      tma_prefetch_desc(&p.a_maps[0]);
      tma_prefetch_desc(&p.b_map);
    }
    const uint32_t own_bytes = static_cast<uint32_t>(p.bw * p.bh * p.bd) * 128u + Cfg2::B_BYTES;
    int st = 0;
    uint32_t ph = 0;
    for (int pair = pair0; pair < total_pairs; pair += pair_step) {
      int n0;
      Tile2 t0, t1;
      decode_pair(pair, n0, t0, t1);
      const Tile2 mine = leader ? t0 : t1;
      unsigned long long mask = tap_mask(t0, t1);
      while (mask) {
        const int t = __ffsll(static_cast<long long>(mask)) - 1;
        mask &= mask - 1;
        const ConvTap tap = p.taps[t];
        const int d = mine.d0 + tap.dd, h = mine.h0 + tap.dh, w = mine.w0 + tap.dw;
        const CUtensorMap* amap = &p.a_maps[tap.map];
        for (int kb = 0; kb < p.kc_blocks; kb++) {
          if (lane == 0) {
            mbar_wait_spin(&empty[st], ph ^ 1, 3136);  // freed in both CTAs by the leader's multicast commit
            if (leader) mbar_arrive_expect_tx(&full[st], 2u * own_bytes);
          }
          __syncwarp();
          const uint32_t bar = mapa_shared(smem_u32(&full[st]), 0);
          if (lane == 0) tma2_load_5d(smem_a + st * Cfg2::A_BYTES, amap, bar, kb * 64, w, h, d, mine.n);
          if (lane == 1)
            tma2_load_2d(smem_b + st * Cfg2::B_BYTES, &p.b_map, bar, tap.kofs + kb * 64, n0 + static_cast<int>(rank) * k2HalfN);
          if (++st == k2Stages) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      constexpr uint32_t idesc = umma_idesc_bf16(256, k2BlockN, false, false);
      const uint64_t desc_hi = umma_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFF00000000ull;
      const uint32_t desc_lo0 = static_cast<uint32_t>(umma_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFFull);
      int st = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t accph = 0;
      for (int pair = pair0; pair < total_pairs; pair += pair_step) {
        int n0;
        Tile2 t0, t1;
        decode_pair(pair, n0, t0, t1);
        const int nkb = __popcll(tap_mask(t0, t1)) * p.kc_blocks;
        mbar_wait_spin(&tempty[acc], accph ^ 1, 3168);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * k2BlockN);
        for (int i = 0; i < nkb; i++) {
          mbar_wait_spin(&full[st], ph, 3172);  // both CTAs' boxes and weight halves have landed
          tc_fence_after();
          const uint32_t a_lo = desc_lo0 + ((smem_u32(smem_a + st * Cfg2::A_BYTES) & 0x3FFFFu) >> 4);
          const uint32_t b_lo = desc_lo0 + ((smem_u32(smem_b + st * Cfg2::B_BYTES) & 0x3FFFFu) >> 4);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; k++)
              umma2_bf16(d_tmem, desc_hi | (a_lo + k * 2), desc_hi | (b_lo + k * 2), idesc, static_cast<uint32_t>(i | k));
            umma2_commit_multicast(&empty[st], 0b11);  // frees the stage in both CTAs
          }
          __syncwarp();
          if (++st == k2Stages) {
            st = 0;
            ph ^= 1;
          }
        }
        if (elect_one_sync()) umma2_commit_multicast(&tfull[acc], 0b11);  // accumulator complete -> both epilogues
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          accph ^= 1;
        }
      }
    }
  } else {
    // ===================== Epilogue (warps 2..5): this CTA's 128 rows =====================
    const int q = warp & 3;
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;
    const int row = q * 32 + lane;
    const bool do_stats = p.stat_sum != nullptr;
    const uint32_t tempty_leader0 = mapa_shared(smem_u32(&tempty[0]), 0);
    const uint32_t tempty_leader1 = mapa_shared(smem_u32(&tempty[1]), 0);
    int acc = 0;
    uint32_t accph = 0;
    for (int pair = pair0; pair < total_pairs; pair += pair_step) {
      int n0;
      Tile2 t0, t1;
      decode_pair(pair, n0, t0, t1);
      const Tile2 c = leader ? t0 : t1;
      const bool has_k = tap_mask(t0, t1) != 0ull;
      const int rw = row % p.bw;
      const int rh = (row / p.bw) % p.bh;
      const int rd = row / (p.bw * p.bh);
      const int od = c.d0 + rd, oh = c.h0 + rh, ow = c.w0 + rw;
      const bool valid = rd < p.bd && od < p.Do && oh < p.Ho && ow < p.Wo;
      const long long off = c.n * p.out_sn + od * p.out_sd + oh * p.out_sh + ow * p.out_sw + n0;

      mbar_wait_spin(&tfull[acc], accph, 3217);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < k2BlockN / 32; chunk++) {
        uint32_t v[32];
        if (has_k) {
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * k2BlockN + chunk * 32), v);
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
          for (int j = 0; j < 32; j++) f[j] += __ldg(p.bias + n0 + chunk * 32 + j);
        }
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
          stat_smem[(ew * 2 + 0) * k2BlockN + chunk * 32 + lane] = cs1;
          stat_smem[(ew * 2 + 1) * k2BlockN + chunk * 32 + lane] = cs2;
        }
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
      // accumulator drained in this CTA -> one of the 8 arrivals the leader's MMA warp waits for
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
      if (++acc == 2) {
        acc = 0;
        accph ^= 1;
      }
      if (do_stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int col = et; col < k2BlockN; col += 128) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int w4 = 0; w4 < 4; w4++) {
            a += stat_smem[(w4 * 2 + 0) * k2BlockN + col];
            b += stat_smem[(w4 * 2 + 1) * k2BlockN + col];
          }
          atomicAdd(p.stat_sum + n0 + col, static_cast<double>(a));
          atomicAdd(p.stat_sq + n0 + col, static_cast<double>(b));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }

  // Neither CTA may retire (or free its TMEM) while its peer can still multicast onto its barriers, read its
  // operand stages or arrive on its tempty barriers.
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem2_dealloc(tmem_base, Cfg2::TMEM_COLS);
  }
}

}  // namespace

// Whether the CTA-pair engine can take this plan: N tile of 256, an even number of samples, >= 2 SMs.
bool igemm_2cta_supported(const IgemmParams& p, int block_n) {
  return block_n == k2BlockN && p.N >= 2 && (p.N % 2) == 0 && num_sms() >= 2;
}

// `p.b_map` must have been encoded with a (64, 128) box: each CTA of a pair loads half of the 256 weight rows.
int launch_igemm_2cta(const IgemmParams& p, cudaStream_t stream) {
  auto kern = igemm_kmajor_2cta_kernel;
  static bool attr_set = false;
  if (!attr_set) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES));
    attr_set = true;
  }
  const int total_pairs = (p.N / 2) * p.tiles_d * p.tiles_h * p.tiles_w * p.n_tiles;
  const int max_clusters = num_sms() / 2;
  const int clusters = total_pairs < max_clusters ? total_pairs : max_clusters;
  kern<<<2 * clusters, k2Threads, Cfg2::SMEM_BYTES, stream>>>(p);
  count_launch();
  ADNI_LAUNCH_CHECK("igemm_kmajor_2cta_kernel");
  return ADNI_OK;
}

}  // namespace adni

// Halo-resident tcgen05 implicit-GEMM Conv3d (fprop + dgrad) for the 3x3x3 / stride 1 / dilation 1 convs of
// ResNet layer1 (64 -> 64 at 32^3) and layer2 (128 -> 128 at 16^3); reference call sites as in
// conv_igemm_kernels.cu (MedicalNet BasicBlock convs reached from pkg/models/mri_models/anat_cnn.py:95).
//
// Why a second engine: the tap-per-TMA-box kernel re-fetches the activation box once per filter tap (27x).
// With only 64 / 128 output channels to amortise it over, that A-operand stream (measured ~5.2 TB/s out of L2 on
// every layer) is the bound: layer1 runs at 0.33 PFLOP/s, 16 % tensor-pipe active.  Here the CTA owns a COLUMN
// of output plane pieces (8 w x 16 h positions, walking along d) and keeps the zero-padded input planes
// ((8+2) x (16+2) positions x 64 channels each, one TMA box per plane, hardware zero fill at the borders) in a
// shared-memory ring.  Every tap's A operand is a shifted VIEW of a resident plane: the UMMA shared-memory
// descriptor starts at row (kh*10 + kw) of the plane and strides 10 rows between the 16 eight-row groups of the
// M = 128 tile, so the 128-byte swizzle phase (address bits 7..9) stays the one TMA wrote.  L2 -> SM traffic for
// activations drops from 27x to ~1.4x the tensor; only the weight tiles (shared by all CTAs) are streamed.
//
// What bounds it after that, and the three answers (measured with the diagnostics described in DESIGN.md):
//  (1) the issuing thread: one pipeline stage costs ~190 cycles of mbarrier wait / tcgen05.commit plus ~48 cycles
//      per tcgen05.mma, and the tensor pipe only queues 1-2 MMAs ahead (tools/ubench/pipe_bench.cu).  TWO MMA
//      warps take alternate weight stages and accumulate into their own TMEM accumulators (summed in the epilogue in
//      a fixed order: deterministic) - the ring depth is EVEN so that every ring slot has one owner; a weight stage
//      carries TPS taps (3 for N = 64: one kh row);
//  (2) the weight stream itself (24 KB per 12 MMAs = 33 B/clk/SM of the same lines for all 148 CTAs): with G = 2 a
//      weight stage is applied to TWO consecutive output pieces of the column (planes d-1 .. d+2 resident);
//  (3) instruction overhead: everything is a template constant or scalar ring arithmetic - no divisions, no
//      parameter-table lookups inside the loops.
//
// Warp roles (256 threads, 1 CTA / SM): warp 0 = weight-stage TMA producer, warps 1 and 7 = MMA issuers (warp 1 owns
// TMEM), warps 2-5 = epilogue (TMEM -> bias / residual-gradient add / BatchNorm sums -> bf16 stores), warp 6 =
// halo-plane TMA producer.  Accumulators: [issuer 2][buffer 2][piece G] x BLOCK_N columns = all 512 TMEM columns;
// the two buffers overlap the epilogue of group i with the MMAs of group i+1.
#include "conv_igemm.cuh"

namespace adni {

extern void count_launch();

namespace {

constexpr int kHaloThreads = 256;
constexpr int kHaloBoxH = 16, kHaloBoxW = 8;
constexpr int kHaloRows = kHaloBoxH + 2;
constexpr int kHaloPitch = kHaloBoxW + 2;                                         // positions per halo row
constexpr int kHaloPlaneKbBytes = (kHaloPitch * kHaloRows * 128 + 1023) & ~1023;  // one 64-channel plane: 23552

template <int BLOCK_N, int KB, int TPS, int G>
struct HaloCfg {
  static constexpr int B_TILE = BLOCK_N * 128;      // one tap x 64 channels of weights
  static constexpr int STAGE_BYTES = TPS * B_TILE;  // one weight stage
  static constexpr int PLANE_BYTES = KB * kHaloPlaneKbBytes;
  static constexpr int STAT_BYTES = 4 * 2 * BLOCK_N * 4;
  static constexpr int AVAIL = 232448 - 1024 - 512 - STAT_BYTES;
  static constexpr int RING = (G + 2) + (KB == 1 ? 1 : 0);  // planes in use + one prefetch slot where it fits
  static constexpr int NBST_FIT = (AVAIL - RING * PLANE_BYTES) / STAGE_BYTES;
  // EVEN ring depth: stage i is issued by MMA warp (i & 1), so every ring slot has ONE owner.  With an odd depth the
  // uses of a slot alternate between the two issuers; an issuer waiting only for its own uses is then two barrier
  // phases further on its next visit, which a parity wait cannot tell from zero: it could run ahead onto a stage
  // whose TMA load is still in flight and corrupt the empty barrier's arrival count (a rare hang, found the hard way).
  static constexpr int NBST = (NBST_FIT > 8 ? 8 : NBST_FIT) & ~1;
  static constexpr int BAR_OFF = RING * PLANE_BYTES + NBST * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 512 + STAT_BYTES + 1024;
  static constexpr int TMEM_COLS = 4 * G * BLOCK_N;
  static_assert(NBST >= 4 && (NBST & 1) == 0, "weight pipeline: need an even depth >= 4");
  static_assert(TMEM_COLS <= 512, "accumulators exceed TMEM");
  static_assert(2 * (RING + NBST) + 4 <= 60, "barrier block too small");
};

struct HaloSeg {
  int n, h0, w0;  // sample and origin of the plane piece
  int dA, dB;     // output planes [dA, dB) of this column handled by this CTA
  int pf, pl;     // input planes [pf, pl] loaded for them
};

__device__ __forceinline__ HaloSeg halo_segment(const HaloParams& p, int i, int end) {
  HaloSeg s;
  int col = i / p.D;
  s.dA = i - col * p.D;
  s.dB = min(p.D, s.dA + (end - i));
  const int tw = col % p.tiles_w;
  col /= p.tiles_w;
  const int th = col % p.tiles_h;
  s.n = col / p.tiles_h;
  s.h0 = th * kHaloBoxH;
  s.w0 = tw * kHaloBoxW;
  s.pf = max(s.dA - 1, 0);
  s.pl = min(s.dB, p.D - 1);
  return s;
}

template <int BLOCK_N, int KB, int TPS, int G>
__global__ void __launch_bounds__(kHaloThreads, 1) igemm_halo_kernel(const __grid_constant__ HaloParams p) {
  pdl_trigger();   // PDL (common.cuh): the next kernel of the stream may be scheduled once every CTA of this grid has started
  using Cfg = HaloCfg<BLOCK_N, KB, TPS, G>;
  constexpr int RING = Cfg::RING, NBST = Cfg::NBST;
  constexpr int B_TILE = Cfg::B_TILE, STAGE_BYTES = Cfg::STAGE_BYTES, PLANE_BYTES = Cfg::PLANE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_p = smem;
  uint8_t* smem_b = smem + RING * PLANE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* full_p = bars;
  uint64_t* empty_p = full_p + RING;
  uint64_t* full_b = empty_p + RING;
  uint64_t* empty_b = full_b + NBST;
  uint64_t* tfull = empty_b + NBST;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* stat_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int begin = static_cast<int>(static_cast<long long>(p.total) * blockIdx.x / gridDim.x);
  const int end = static_cast<int>(static_cast<long long>(p.total) * (blockIdx.x + 1) / gridDim.x);
  const int D = p.D;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; i++) {
      mbar_init(&full_p[i], 1);
      mbar_init(&empty_p[i], 2);  // both MMA issuers release a plane
    }
    for (int i = 0; i < NBST; i++) {
      mbar_init(&full_b[i], 1);
      mbar_init(&empty_b[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&tfull[i], 2);
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

  if (warp == 6) {
    // ===================== halo-plane producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&p.a_map);
      constexpr uint32_t tx = static_cast<uint32_t>(KB * kHaloPitch * kHaloRows) * 128u;
      int slot = 0;
      uint32_t par = 0;
      for (int i = begin; i < end;) {
        const HaloSeg s = halo_segment(p, i, end);
        for (int pz = s.pf; pz <= s.pl; pz++) {
          mbar_wait_spin(&empty_p[slot], par ^ 1u, 1142);
          mbar_arrive_expect_tx(&full_p[slot], tx);
#pragma unroll
          for (int kb = 0; kb < KB; kb++)
            tma_load_5d(smem_p + slot * PLANE_BYTES + kb * kHaloPlaneKbBytes, &p.a_map, &full_p[slot], kb * 64, s.w0 - 1,
                        s.h0 - 1, pz, s.n);
          if (++slot == RING) {
            slot = 0;
            par ^= 1u;
          }
        }
        i += s.dB - s.dA;
      }
    }
  } else if (warp == 0) {
    // ===================== weight-stage producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&p.b_map);
      constexpr int C = KB * 64;
      const int mirror = p.mirror;
      int st = 0;
      uint32_t ph = 0;
      for (int i = begin; i < end;) {
        const HaloSeg s = halo_segment(p, i, end);
        for (int d = s.dA; d < s.dB; d += G) {
          const int g = min(G, s.dB - d);
#pragma unroll 1
          for (int kd = 0; kd < 3; kd++) {
            if (d + kd + g - 2 < 0 || d + kd - 1 > D - 1) continue;  // no piece of the group has this input plane
#pragma unroll 1
            for (int t9 = 0; t9 < 9; t9 += TPS) {
#pragma unroll
              for (int kb = 0; kb < KB; kb++) {
                mbar_wait_spin(&empty_b[st], ph ^ 1u, 1175);
                mbar_arrive_expect_tx(&full_b[st], STAGE_BYTES);
#pragma unroll
                for (int tp = 0; tp < TPS; tp++) {
                  const int idx = kd * 9 + t9 + tp;  // halo offset (od, oh, ow) -> weight tap (mirrored for dgrad)
                  tma_load_2d(smem_b + st * STAGE_BYTES + tp * B_TILE, &p.b_map, &full_b[st],
                              (mirror ? 26 - idx : idx) * C + kb * 64, 0);
                }
                if (++st == NBST) {
                  st = 0;
                  ph ^= 1u;
                }
              }
            }
          }
        }
        i += s.dB - s.dA;
      }
    }
  } else if (warp == 1 || warp == 7) {
    // ===================== MMA issuers (warp-uniform control flow, one elected lane issues) =====================
    const uint32_t mw = warp == 1 ? 0u : 1u;  // this warp issues the weight stages whose running index has parity mw
    constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N, false, false);
    // descriptors as 32-bit halves: low word = start address >> 4 (LBO unused for swizzled K-major),
    // high word = SBO >> 4 | version | SWIZZLE_128B
    const uint32_t a_hi = static_cast<uint32_t>(umma_smem_desc_sw128(0, 16, kHaloPitch * 128u) >> 32);
    const uint32_t b_hi = static_cast<uint32_t>(umma_smem_desc_sw128(0, 16, 1024) >> 32);
    const uint32_t lo_c = static_cast<uint32_t>(umma_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFFull);
    const uint32_t a_lo0 = lo_c + ((smem_u32(smem_p) & 0x3FFFFu) >> 4);
    const uint32_t b_lo0 = lo_c + ((smem_u32(smem_b) & 0x3FFFFu) >> 4);
    int st = 0;
    uint32_t ph = 0;
    int buf = 0;
    uint32_t bufph = 0;
    int w_slot = 0, r_slot = 0;  // ring cursors of the next plane to wait for / to hand back
    uint32_t w_par = 0;
    int n_waited = 0, n_released = 0, seq0 = 0;
    for (int i = begin; i < end;) {
      const HaloSeg s = halo_segment(p, i, end);
      const int rel0 = s.dA - 1 - s.pf;           // plane dA-1 relative to the first loaded plane (-1 or 0)
      int sl0 = (seq0 + rel0 + 2 * RING) % RING;  // ring slot of plane d-1 (virtual for the plane above the volume)
      for (int d = s.dA; d < s.dB; d += G) {
        const int g = min(G, s.dB - d);
        mbar_wait_spin(&tempty[buf], bufph ^ 1u, 1219);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (mw * 2u + static_cast<uint32_t>(buf)) * (G * BLOCK_N);
        uint32_t started = 0;  // bit q: this issuer's accumulator of piece q has been written in this group
#pragma unroll 1
        for (int kd = 0; kd < 3; kd++) {
          if (d + kd + g - 2 < 0 || d + kd - 1 > D - 1) continue;
          // planes are waited for when their kd comes up (the last one is only needed by the final third of the
          // group, so its load overlaps the first two thirds); both issuers observe every plane, in load order
          const int need = seq0 + min(d + g + kd - 2, D - 1) - s.pf;
          while (n_waited <= need) {
            mbar_wait_spin(&full_p[w_slot], w_par, 1222);
            n_waited++;
            if (++w_slot == RING) {
              w_slot = 0;
              w_par ^= 1u;
            }
          }
          tc_fence_after();
          // descriptor base and validity of the input plane of every piece for this kd
          uint32_t a_plane[G];
          bool q_ok[G];
#pragma unroll
          for (int q = 0; q < G; q++) {
            const int z = d + q + kd - 1;
            q_ok[q] = q < g && z >= 0 && z < D;
            int sl = sl0 + q + kd;
            if (sl >= RING) sl -= RING;
            a_plane[q] = a_lo0 + static_cast<uint32_t>(sl) * (PLANE_BYTES >> 4);
          }
#pragma unroll 1
          for (int t9 = 0; t9 < 9; t9 += TPS) {
            const int kh = t9 / 3, kw0 = t9 - kh * 3;
            const uint32_t row_off = static_cast<uint32_t>(kh * kHaloPitch + kw0) * (128u >> 4);
#pragma unroll
            for (int kb = 0; kb < KB; kb++) {
              if ((static_cast<uint32_t>(st) & 1u) == mw) {  // ring slot st belongs to issuer (st & 1), see HaloCfg::NBST
                mbar_wait_spin(&full_b[st], ph, 1253);
                tc_fence_after();
                const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(st) * (STAGE_BYTES >> 4);
                if (elect_one_sync()) {
#pragma unroll
                  for (int q = 0; q < G; q++) {
                    if (!q_ok[q]) continue;
                    const uint32_t a_lo = a_plane[q] + kb * (kHaloPlaneKbBytes >> 4) + row_off;
                    const uint32_t first = (started >> q) & 1u;
#pragma unroll
                    for (int tp = 0; tp < TPS; tp++) {
#pragma unroll
                      for (int k = 0; k < 4; k++)
                        umma_bf16(d_tmem + q * BLOCK_N, (static_cast<uint64_t>(a_hi) << 32) | (a_lo + tp * 8 + k * 2),
                                  (static_cast<uint64_t>(b_hi) << 32) | (b_lo + tp * (B_TILE >> 4) + k * 2), idesc,
                                  first | static_cast<uint32_t>(tp | k));
                    }
                  }
                  umma_commit(&empty_b[st]);  // frees the weight stage once these MMAs have read it
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < G; q++)
                  if (q_ok[q]) started |= 1u << q;
              }
              if (++st == NBST) {
                st = 0;
                ph ^= 1u;
              }
            }
          }
        }
        // planes no later group of this column needs go back to the producer once these MMAs retire
        const int last_free = seq0 + ((d + g >= s.dB) ? s.pl - s.pf : (d + g - 2 - s.pf));
        if (elect_one_sync()) {
          umma_commit(&tfull[buf]);
          int rs = r_slot;
          for (int r = n_released; r <= last_free; r++) {
            umma_commit(&empty_p[rs]);
            if (++rs == RING) rs = 0;
          }
        }
        __syncwarp();
        while (n_released <= last_free) {
          n_released++;
          if (++r_slot == RING) r_slot = 0;
        }
        sl0 += g;
        if (sl0 >= RING) sl0 -= RING;
        if (++buf == 2) {
          buf = 0;
          bufph ^= 1u;
        }
      }
      seq0 += s.pl - s.pf + 1;
      i += s.dB - s.dA;
    }
  } else {
    // ===================== Epilogue (warps 2..5) =====================
    const int q4 = warp & 3;  // TMEM lane quadrant this warp may access
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;  // 0..127
    const int row = q4 * 32 + lane;
    const int rh = row / kHaloBoxW, rw = row % kHaloBoxW;
    const bool do_red = p.red_y != nullptr;                  // BatchNorm-BACKWARD sums fused into a dgrad call
    const bool do_fwd_stats = p.stat_sum != nullptr && !do_red;
    const bool do_stats = do_fwd_stats || do_red;            // either way: two per-channel sums into stat_sum / stat_sq
    // BatchNorm sums.  BLOCK_N = 64: every thread owns one tile row and keeps fp32 partial sums of its 64 channels in
    // registers for the whole CTA; one transpose-reduce and one fp64 atomic per channel at the end.  BLOCK_N = 128:
    // per-piece transpose-reduce into per-CTA fp64 sums (the register file does not hold 256 accumulators).
    constexpr bool REG_STATS = BLOCK_N == 64;
    constexpr int NREG = REG_STATS ? BLOCK_N : 1;
    float rs1[NREG], rs2[NREG];
#pragma unroll
    for (int j = 0; j < NREG; j++) rs1[j] = rs2[j] = 0.f;
    double cta_sum = 0.0, cta_sq = 0.0;
    // fp32 register partials are folded into the fp64 per-CTA sums every kFlushGroups groups: a thread never chains
    // more than 2 * kFlushGroups fp32 additions per channel, which keeps sum(x) accurate when it nearly cancels
    constexpr int kFlushGroups = 8;
    int since_flush = 0;
    auto flush_reg_stats = [&]() {
#pragma unroll
      for (int chunk = 0; chunk < (REG_STATS ? BLOCK_N / 32 : 0); chunk++) {
        float t1[32], t2[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
          t1[j] = rs1[chunk * 32 + j];
          t2[j] = rs2[chunk * 32 + j];
          rs1[chunk * 32 + j] = 0.f;
          rs2[chunk * 32 + j] = 0.f;
        }
        const float c1 = warp_column_sums(t1, lane);
        const float c2 = warp_column_sums(t2, lane);
        stat_smem[(ew * 2 + 0) * BLOCK_N + chunk * 32 + lane] = c1;
        stat_smem[(ew * 2 + 1) * BLOCK_N + chunk * 32 + lane] = c2;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < BLOCK_N) {
#pragma unroll
        for (int w4 = 0; w4 < 4; w4++) {
          cta_sum += static_cast<double>(stat_smem[(w4 * 2 + 0) * BLOCK_N + et]);
          cta_sq += static_cast<double>(stat_smem[(w4 * 2 + 1) * BLOCK_N + et]);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      since_flush = 0;
    };
    int buf = 0;
    uint32_t bufph = 0;
    for (int i = begin; i < end;) {
      const HaloSeg s = halo_segment(p, i, end);
      const int oh = s.h0 + rh, ow = s.w0 + rw;
      const bool valid = oh < p.H && ow < p.W;
      for (int d = s.dA; d < s.dB; d += G) {
        const int g = min(G, s.dB - d);
        if (REG_STATS && do_stats && ++since_flush > kFlushGroups) flush_reg_stats();
        mbar_wait_spin(&tfull[buf], bufph, 1367);
        tc_fence_after();
#pragma unroll 1
        for (int q = 0; q < g; q++) {
          const long long off = s.n * p.out_sn + (d + q) * p.out_sd + oh * p.out_sh + ow * p.out_sw;
#pragma unroll
          for (int chunk = 0; chunk < BLOCK_N / 32; chunk++) {
            uint32_t v[32], v2[32];
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) +
                                   static_cast<uint32_t>((buf * G + q) * BLOCK_N + chunk * 32);
            tmem_ld_32x32(taddr, v);                     // issuer 0's partial sums
            tmem_ld_32x32(taddr + 2 * G * BLOCK_N, v2);  // issuer 1's
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j++) f[j] += __ldg(p.bias + chunk * 32 + j);
            }
            if (do_fwd_stats) {   // forward sums: y, y^2 (before any addend)
              if constexpr (REG_STATS) {
                if (valid) {
#pragma unroll
                  for (int j = 0; j < 32; j++) {
                    rs1[chunk * 32 + j] += f[j];
                    rs2[chunk * 32 + j] = fmaf(f[j], f[j], rs2[chunk * 32 + j]);
                  }
                }
              } else {
                float s1[32], s2[32];
#pragma unroll
                for (int j = 0; j < 32; j++) {
                  const float x = valid ? f[j] : 0.f;
                  s1[j] = x;
                  s2[j] = x * x;
                }
                const float cs1 = warp_column_sums(s1, lane);
                const float cs2 = warp_column_sums(s2, lane);
                stat_smem[(ew * 2 + 0) * BLOCK_N + chunk * 32 + lane] = cs1;
                stat_smem[(ew * 2 + 1) * BLOCK_N + chunk * 32 + lane] = cs2;
              }
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
              uint4* op = reinterpret_cast<uint4*>(p.out + off + chunk * 32);
#pragma unroll
              for (int j4 = 0; j4 < 4; j4++)
                op[j4] = make_uint4(packed[4 * j4], packed[4 * j4 + 1], packed[4 * j4 + 2], packed[4 * j4 + 3]);
            }
            if (do_red) {  // backward sums of the STORED gradient, masked by the preceding layer's ReLU: g, g*y
              float s1[REG_STATS ? 1 : 32], s2[REG_STATS ? 1 : 32];
              if constexpr (!REG_STATS) {
#pragma unroll
                for (int j = 0; j < 32; j++) s1[j] = s2[j] = 0.f;
              }
              if (valid) {
                const uint4* yp = reinterpret_cast<const uint4*>(p.red_y + off + chunk * 32);
                const uint4* mp =
                    p.red_mask != nullptr ? reinterpret_cast<const uint4*>(p.red_mask + off + chunk * 32) : nullptr;
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) {
                  const uint4 yv = __ldg(yp + j4);
                  const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
                  uint32_t mw[4] = {0u, 0u, 0u, 0u};
                  if (mp != nullptr) {
                    const uint4 mv = __ldg(mp + j4);
                    mw[0] = mv.x, mw[1] = mv.y, mw[2] = mv.z, mw[3] = mv.w;
                  }
#pragma unroll
                  for (int e = 0; e < 4; e++) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                      const int j = j4 * 8 + e * 2 + h;
                      const float y = h ? bf16_hi(yw[e]) : bf16_lo(yw[e]);
                      float gq = h ? bf16_hi(packed[j >> 1]) : bf16_lo(packed[j >> 1]);
                      if (mp != nullptr) {
                        gq = (h ? bf16_hi(mw[e]) : bf16_lo(mw[e])) > 0.f ? gq : 0.f;
                      } else if (p.red_scale != nullptr) {
                        gq = fmaf(y, __ldg(p.red_scale + chunk * 32 + j), __ldg(p.red_shift + chunk * 32 + j)) > 0.f ? gq : 0.f;
                      }
                      if constexpr (REG_STATS) {
                        rs1[chunk * 32 + j] += gq;
                        rs2[chunk * 32 + j] = fmaf(gq, y, rs2[chunk * 32 + j]);
                      } else {
                        s1[j] = gq;
                        s2[j] = gq * y;
                      }
                    }
                  }
                }
              }
              if constexpr (!REG_STATS) {
                const float cs1 = warp_column_sums(s1, lane);
                const float cs2 = warp_column_sums(s2, lane);
                stat_smem[(ew * 2 + 0) * BLOCK_N + chunk * 32 + lane] = cs1;
                stat_smem[(ew * 2 + 1) * BLOCK_N + chunk * 32 + lane] = cs2;
              }
            }
          }
          if (do_stats && !REG_STATS) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et < BLOCK_N) {  // BLOCK_N <= 128: thread et owns channel et
              float a = 0.f, b = 0.f;
#pragma unroll
              for (int w4 = 0; w4 < 4; w4++) {
                a += stat_smem[(w4 * 2 + 0) * BLOCK_N + et];
                b += stat_smem[(w4 * 2 + 1) * BLOCK_N + et];
              }
              cta_sum += static_cast<double>(a);
              cta_sq += static_cast<double>(b);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
          }
        }
        // accumulators drained -> hand the TMEM buffer back to the MMA warps
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        if (++buf == 2) {
          buf = 0;
          bufph ^= 1u;
        }
      }
      i += s.dB - s.dA;
    }
    if (do_stats && begin < end) {
      if constexpr (REG_STATS) flush_reg_stats();
      // flushed ONCE per CTA: same-address fp64 atomics retire at ~1 per 27 cycles in L2, one per (piece, channel)
      // from 148 CTAs is a serial bottleneck of its own
      if (et < BLOCK_N) {
        atomicAdd(p.stat_sum + et, cta_sum);
        atomicAdd(p.stat_sq + et, cta_sq);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BLOCK_N, int KB, int TPS, int G>
int launch_halo_t(const HaloParams& p, cudaStream_t stream) {
  using Cfg = HaloCfg<BLOCK_N, KB, TPS, G>;
  auto kern = igemm_halo_kernel<BLOCK_N, KB, TPS, G>;
  static bool attr_set = false;
  if (!attr_set) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int grid = p.total < num_sms() ? p.total : num_sms();
  pdl_launch(kern, grid, kHaloThreads, Cfg::SMEM_BYTES, stream)(p);
  count_launch();
  ADNI_LAUNCH_CHECK("igemm_halo_kernel");
  return ADNI_OK;
}

}  // namespace

int halo_pitch() { return kHaloPitch; }
int halo_rows() { return kHaloRows; }

// channels = Cin = Cout (64 or 128)
int launch_igemm_halo(const HaloParams& p, int channels, cudaStream_t stream) {
  switch (channels) {
    case 64:
      return launch_halo_t<64, 1, 3, 2>(p, stream);
    case 128:
      return launch_halo_t<128, 2, 1, 1>(p, stream);
    default:
      set_error("igemm_halo: unsupported channel count %d", channels);
      return ADNI_ENOTSUP;
  }
}

}  // namespace adni

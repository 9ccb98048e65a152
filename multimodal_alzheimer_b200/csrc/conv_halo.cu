// Halo-resident tcgen05 implicit-GEMM Conv3d (fprop + dgrad) for the 3x3x3 / stride 1 / dilation 1 convs of
// ResNet layer1 (64 -> 64 at 32^3) and layer2 (128 -> 128 at 16^3); reference call sites as in
// conv_igemm_kernels.cu (MedicalNet BasicBlock convs reached from pkg/models/mri_models/anat_cnn.py:95).
//
// Why a second engine: the tap-per-TMA-box kernel re-fetches the activation box once per filter tap (27x).
// With only 64 / 128 output channels to amortise it over, that A-operand stream (measured ~5.2 TB/s out of L2 on
// every layer) is the bound: layer1 runs at 0.33 PFLOP/s, 16 % tensor-pipe active.  Here the CTA owns a COLUMN
// of output plane pieces (8 w x 16 h positions, walking along d) and keeps the zero-padded input planes
// d-1, d, d+1 ((8+2) x (16+2) positions x 64 channels each, one TMA box per plane, hardware zero fill at the
// borders) in a shared-memory ring.  Every tap's A operand is a shifted VIEW of a resident plane: the UMMA
// shared-memory descriptor starts at row (kh*pitch + kw) of the plane and strides `pitch` rows between the
// 16 eight-row groups of the M = 128 tile, so the 128-byte swizzle phase (address bits 7..9) stays the one
// TMA wrote.  L2 -> SM traffic for activations drops from 27x to ~1.4x the tensor; only the weight tiles
// (shared by all CTAs) are streamed per tap.
//
// Issue-side bound.  tools/ubench/pipe_bench.cu: one pipeline stage costs the issuing thread ~190 cycles of
// mbarrier wait / tcgen05.commit plus ~48 cycles per tcgen05.mma, and the tensor pipe only queues 1-2 MMAs ahead,
// so with N = 64 / 128 (60 / 64-cycle MMAs) a single issuer leaves the pipe idle for most of that overhead.
// Two remedies here: (1) TWO MMA-issuing warps take alternate weight stages and accumulate into their own TMEM
// accumulators (summed in the epilogue in a fixed order, so results stay deterministic) - one warp's barrier
// overhead overlaps the other's MMAs; (2) a weight stage carries `tps` taps (3 for N = 64: one kh row, 12 MMAs).
//
// Warp roles (256 threads, 1 CTA / SM): warp 0 = weight-tile TMA producer, warps 1 and 7 = MMA issuers (warp 1 owns
// TMEM), warps 2-5 = epilogue (TMEM -> bias / residual-gradient add / BatchNorm sums -> bf16 stores), warp 6 =
// halo-plane TMA producer.  Two accumulator buffers per issuer overlap the epilogue of piece i with the MMAs of
// piece i+1 (4 x BLOCK_N TMEM columns in total).
#include "conv_igemm.cuh"

namespace adni {

extern void count_launch();

namespace {

constexpr int kHaloThreads = 256;
constexpr int kHaloBoxH = 16, kHaloBoxW = 8;
constexpr int kHaloRows = kHaloBoxH + 2;
constexpr int kMaxRing = 4, kMaxBStages = 16;

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

template <int BLOCK_N>
__global__ void __launch_bounds__(kHaloThreads, 1) igemm_halo_kernel(const __grid_constant__ HaloParams p) {
  constexpr int B_TILE = BLOCK_N * 128;  // one tap x 64 channels of weights
  constexpr int TMEM_COLS = 4 * BLOCK_N;
  const int stage_bytes = p.tps * B_TILE;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int plane_bytes = p.kb * p.plane_kb_bytes;
  uint8_t* smem_p = smem;
  uint8_t* smem_b = smem + p.ring * plane_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + p.b_stages * stage_bytes);
  uint64_t* full_p = bars;
  uint64_t* empty_p = bars + kMaxRing;
  uint64_t* full_b = bars + 2 * kMaxRing;
  uint64_t* empty_b = full_b + kMaxBStages;
  uint64_t* tfull = empty_b + kMaxBStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* stat_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int begin = static_cast<int>(static_cast<long long>(p.total) * blockIdx.x / gridDim.x);
  const int end = static_cast<int>(static_cast<long long>(p.total) * (blockIdx.x + 1) / gridDim.x);
  const int ring = p.ring, nbst = p.b_stages;
  // diagnostics (ADNI_HALO_DEBUG bits): 1 no MMA issue, 2 no weight TMA, 4 no plane TMA, 8 no epilogue stores

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxRing; i++) {
      mbar_init(&full_p[i], 1);
      mbar_init(&empty_p[i], 2);
    }
    for (int i = 0; i < kMaxBStages; i++) {
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
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 6) {
    // ===================== halo-plane producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&p.a_map);
      const uint32_t tx = static_cast<uint32_t>(p.kb * p.pitch * kHaloRows) * 128u;
      uint32_t seq = 0;
      for (int i = begin; i < end;) {
        const HaloSeg s = halo_segment(p, i, end);
        for (int pz = s.pf; pz <= s.pl; pz++, seq++) {
          const uint32_t slot = seq % ring, par = (seq / ring) & 1u;
          mbar_wait_spin(&empty_p[slot], par ^ 1u);
          if (p.debug & 4) {
            mbar_arrive(&full_p[slot]);
            continue;
          }
          mbar_arrive_expect_tx(&full_p[slot], tx);
          for (int kb = 0; kb < p.kb; kb++)
            tma_load_5d(smem_p + slot * plane_bytes + kb * p.plane_kb_bytes, &p.a_map, &full_p[slot], kb * 64,
                        s.w0 - 1, s.h0 - 1, pz, s.n);
        }
        i += s.dB - s.dA;
      }
    }
  } else if (warp == 0) {
    // ===================== weight-tile producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&p.b_map);
      int st = 0;
      uint32_t ph = 0;
      for (int i = begin; i < end;) {
        const HaloSeg s = halo_segment(p, i, end);
        for (int d = s.dA; d < s.dB; d++) {
          for (int kd = 0; kd < 3; kd++) {
            const int pz = d + kd - 1;
            if (pz < 0 || pz >= p.D) continue;
            for (int kh = 0; kh < 3; kh++) {
              for (int kw0 = 0; kw0 < 3; kw0 += p.tps) {
                for (int kb = 0; kb < p.kb; kb++) {
                  mbar_wait_spin(&empty_b[st], ph ^ 1u);
                  if (p.debug & 2) {
                    mbar_arrive(&full_b[st]);
                  } else {
                    mbar_arrive_expect_tx(&full_b[st], static_cast<uint32_t>(stage_bytes));
                    for (int tp = 0; tp < p.tps; tp++)
                      tma_load_2d(smem_b + st * stage_bytes + tp * B_TILE, &p.b_map, &full_b[st],
                                  p.kofs[kd * 9 + kh * 3 + kw0 + tp] + kb * 64, 0);
                  }
                  if (++st == nbst) {
                    st = 0;
                    ph ^= 1u;
                  }
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
    const uint64_t a_hi = umma_smem_desc_sw128(0, 16, static_cast<uint32_t>(p.pitch) * 128u) & 0xFFFFFFFF00000000ull;
    const uint64_t b_hi = umma_smem_desc_sw128(0, 16, 1024) & 0xFFFFFFFF00000000ull;
    const uint32_t smem_p_u32 = smem_u32(smem_p), smem_b_u32 = smem_u32(smem_b);
    int st = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t accph = 0;
    uint32_t seq0 = 0, waited = 0, gs = 0;
    for (int i = begin; i < end;) {
      const HaloSeg s = halo_segment(p, i, end);
      for (int d = s.dA; d < s.dB; d++) {
        mbar_wait_spin(&tempty[acc], accph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (mw * 2u + static_cast<uint32_t>(acc)) * BLOCK_N;
        uint32_t accum = 0;
        for (int kd = 0; kd < 3; kd++) {
          const int pz = d + kd - 1;
          if (pz < 0 || pz >= p.D) continue;
          const uint32_t sq = seq0 + static_cast<uint32_t>(pz - s.pf);
          const uint32_t slot = sq % ring;
          while (waited <= sq) {  // planes become visible in load order; each is waited for exactly once
            mbar_wait_spin(&full_p[waited % ring], (waited / ring) & 1u);
            waited++;
          }
          const uint32_t plane_addr = smem_p_u32 + slot * plane_bytes;
          for (int kh = 0; kh < 3; kh++) {
            for (int kw0 = 0; kw0 < 3; kw0 += p.tps) {
              for (int kb = 0; kb < p.kb; kb++, gs++) {
                if ((gs & 1u) == mw) {
                  mbar_wait_spin(&full_b[st], ph);
                  tc_fence_after();
                  const uint32_t a_lo =
                      ((plane_addr + kb * p.plane_kb_bytes + static_cast<uint32_t>(kh * p.pitch + kw0) * 128u) & 0x3FFFFu) >> 4;
                  const uint32_t b_lo = ((smem_b_u32 + st * stage_bytes) & 0x3FFFFu) >> 4;
                  if (elect_one_sync()) {
                    if (!(p.debug & 1)) {
                      for (int tp = 0; tp < p.tps; tp++) {
#pragma unroll
                        for (int k = 0; k < 4; k++)
                          umma_bf16(d_tmem, a_hi | (a_lo + tp * 8 + k * 2), b_hi | (b_lo + tp * (B_TILE >> 4) + k * 2), idesc,
                                    accum | static_cast<uint32_t>(tp | k));
                      }
                    }
                    umma_commit(&empty_b[st]);  // frees the weight stage once these MMAs have read it
                  }
                  __syncwarp();
                  accum = 1;
                }
                if (++st == nbst) {
                  st = 0;
                  ph ^= 1u;
                }
              }
            }
          }
        }
        if (elect_one_sync()) {
          umma_commit(&tfull[acc]);
          // planes no later piece of this column needs go back to the producer once these MMAs retire
          if (d - 1 >= s.pf) umma_commit(&empty_p[(seq0 + static_cast<uint32_t>(d - 1 - s.pf)) % ring]);
          if (d == s.dB - 1) {
            umma_commit(&empty_p[(seq0 + static_cast<uint32_t>(d - s.pf)) % ring]);
            if (d + 1 <= s.pl) umma_commit(&empty_p[(seq0 + static_cast<uint32_t>(d + 1 - s.pf)) % ring]);
          }
        }
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          accph ^= 1u;
        }
      }
      seq0 += static_cast<uint32_t>(s.pl - s.pf + 1);
      i += s.dB - s.dA;
    }
  } else {
    // ===================== Epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;  // 0..127
    const int row = q * 32 + lane;
    const int rh = row / kHaloBoxW, rw = row % kHaloBoxW;
    const bool do_stats = p.stat_sum != nullptr;
    // BatchNorm sums are carried per CTA in fp64 registers and flushed ONCE: same-address fp64 atomics retire at
    // ~1 per 27 cycles in L2, so one atomic per (piece, channel) from 148 CTAs is a serial bottleneck of its own
    // (stem: 65536 pieces x 27 cycles = the whole kernel time).
    double cta_sum = 0.0, cta_sq = 0.0;
    int acc = 0;
    uint32_t accph = 0;
    for (int i = begin; i < end;) {
      const HaloSeg s = halo_segment(p, i, end);
      const int oh = s.h0 + rh, ow = s.w0 + rw;
      const bool valid = oh < p.H && ow < p.W;
      for (int d = s.dA; d < s.dB; d++) {
        const long long off = s.n * p.out_sn + d * p.out_sd + oh * p.out_sh + ow * p.out_sw;
        mbar_wait_spin(&tfull[acc], accph);
        tc_fence_after();
#pragma unroll 1
        for (int chunk = 0; chunk < BLOCK_N / 32; chunk++) {
          uint32_t v[32], v2[32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 static_cast<uint32_t>(acc * BLOCK_N + chunk * 32);
          tmem_ld_32x32(taddr, v);                 // issuer 0's partial sums
          tmem_ld_32x32(taddr + 2 * BLOCK_N, v2);  // issuer 1's
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j++) f[j] += __ldg(p.bias + chunk * 32 + j);
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
            stat_smem[(ew * 2 + 0) * BLOCK_N + chunk * 32 + lane] = cs1;
            stat_smem[(ew * 2 + 1) * BLOCK_N + chunk * 32 + lane] = cs2;
          }
          if (valid && !(p.debug & 8)) {
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
        // accumulator drained -> hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        if (++acc == 2) {
          acc = 0;
          accph ^= 1u;
        }
        if (do_stats) {
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
      i += s.dB - s.dA;
    }
    if (do_stats && et < BLOCK_N && begin < end) {
      atomicAdd(p.stat_sum + et, cta_sum);
      atomicAdd(p.stat_sq + et, cta_sq);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BLOCK_N>
int launch_halo_t(const HaloParams& p, int smem_bytes, cudaStream_t stream) {
  auto kern = igemm_halo_kernel<BLOCK_N>;
  static int attr_bytes = 0;
  if (attr_bytes < smem_bytes) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_bytes = smem_bytes;
  }
  const int grid = p.total < num_sms() ? p.total : num_sms();
  kern<<<grid, kHaloThreads, smem_bytes, stream>>>(p);
  count_launch();
  ADNI_LAUNCH_CHECK("igemm_halo_kernel");
  return ADNI_OK;
}

}  // namespace

// Fills ring / b_stages / plane_kb_bytes for the given (block_n, kb, pitch); returns the dynamic shared-memory size
// or 0 when the configuration does not fit one SM.
int halo_plan_smem(HaloParams* p, int block_n) {
  const int stat_bytes = 4 * 2 * block_n * 4;
  const int avail = 232448 - 1024 - 512 - stat_bytes;
  p->plane_kb_bytes = (p->pitch * kHaloRows * 128 + 1023) & ~1023;
  const int plane_bytes = p->kb * p->plane_kb_bytes;
  p->tps = block_n == 64 ? 3 : 1;  // taps per weight stage (one kh row for N = 64)
  const int b_stage = p->tps * block_n * 128;
  for (int ring = kMaxRing; ring >= 3; ring--) {
    const int left = avail - ring * plane_bytes;
    if (left < 4 * b_stage) continue;
    p->ring = ring;
    p->b_stages = left / b_stage < 12 ? left / b_stage : 12;
    return ring * plane_bytes + p->b_stages * b_stage + 512 + stat_bytes + 1024;
  }
  return 0;
}

int launch_igemm_halo(const HaloParams& p, int block_n, int smem_bytes, cudaStream_t stream) {
  switch (block_n) {
    case 64:
      return launch_halo_t<64>(p, smem_bytes, stream);
    case 128:
      return launch_halo_t<128>(p, smem_bytes, stream);
    default:
      set_error("igemm_halo: unsupported BLOCK_N %d", block_n);
      return ADNI_ENOTSUP;
  }
}

}  // namespace adni

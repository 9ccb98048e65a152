// Halo-plane wgrad for the 64 -> 64, 3x3x3, stride 1, dilation 1 convs of ResNet layer1 (MedicalNet BasicBlock convs,
// reference call site pkg/models/mri_models/anat_cnn.py:95; autograd backward of SURVEY.md row A16).
//
// wgrad2_kernel puts Cout on the M axis: with Cout = 64 half of every M = 128 MMA is empty and the X boxes are
// re-fetched once per tap (450 TFLOP/s).  Here the roles are swapped and the taps are paired on the M axis:
//   D_pair[(tap a | tap b) x 64 cin][64 cout] += sum over the 128 positions of a piece  X(tap)[pos][cin]^T * dY[pos][cout]
// * the unit of work is (output piece of 8 w x 16 h positions, kd): ONE zero-padded input plane (10 x 18 positions x
//   64 channels, one TMA box) and the dY piece (one swizzled TMA box) per pipeline stage, 40 MMAs per stage;
// * A operand (MN-major, 128-byte swizzle) = VIEW of the resident plane: start row = tap a's (kh*10 + kw), SBO (next
//   8 positions = next h row) = 10 rows, LBO (second 64-row M group = tap b) = the row distance between the two
//   taps (1 or 8 rows) - both taps of a pair read the same plane;
// * a CTA serves a single kd (blockIdx % 3), so its 5 accumulators (taps 0|1, 2|3, 4|5, 6|7, 8|-) x 64 columns live
//   in TMEM for the whole kernel and are added once to dW in [tap][cin][cout] order (coalesced 16-byte red.add);
//   a small transpose then produces the [cout][tap][cin] layout the other wgrad engines emit.
#include <stdlib.h>
#include <string.h>

#include "conv_igemm.cuh"

namespace adni {
extern void count_launch();
namespace {

constexpr int kWHThreads = 192;
constexpr int kWHPitch = 10, kWHRows = 18;
constexpr int kWHPlaneBytes = (kWHPitch * kWHRows * 128 + 1023) & ~1023;  // 23552
constexpr int kWHDyBytes = 128 * 128;                                       // 16384
constexpr int kWHStageBytes = kWHPlaneBytes + kWHDyBytes;                    // 39936
constexpr int kWHStages = 5;
constexpr int kWHBarOff = kWHStages * kWHStageBytes;
constexpr int kWHSmem = kWHBarOff + 256 + 1024;

struct WgradHaloParams {
  CUtensorMap x_map;   // 5-D (64, W, H, D, N) view of x, box (64, 10, 18, 1, 1), 128-B swizzle
  CUtensorMap dy_map;  // 5-D (64, W, H, D, N) view of dy, box (64, 8, 16, 1, 1), 128-B swizzle
  int N, D, H, W;
  int tiles_h, tiles_w;
  int total;    // N * tiles_h * tiles_w * D pieces
  float* dw_tic;  // [27][64 cin][64 cout] fp32, accumulated with red.add
};

struct WHUnit {
  int n, h0, w0, d;
};
__device__ __forceinline__ WHUnit wh_decode(const WgradHaloParams& p, int piece) {
  WHUnit u;
  int col = piece / p.D;
  u.d = piece - col * p.D;
  const int tw = col % p.tiles_w;
  col /= p.tiles_w;
  u.h0 = (col % p.tiles_h) * 16;
  u.n = col / p.tiles_h;
  u.w0 = tw * 8;
  return u;
}

// in-plane tap t9 = kh*3 + kw  ->  first halo row of its view
__device__ __forceinline__ constexpr int wh_row(int t9) { return (t9 / 3) * kWHPitch + (t9 % 3); }

__global__ void __launch_bounds__(kWHThreads, 1) wgrad_halo_kernel(const __grid_constant__ WgradHaloParams p) {
  pdl_trigger();   // PDL (common.cuh): the next kernel of the stream may be scheduled once every CTA of this grid has started
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWHBarOff);
  uint64_t* empty = full + kWHStages;
  uint64_t* tfull = empty + kWHStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA class = kd; the CTAs of a class split the pieces evenly
  const int kd = blockIdx.x % 3;
  const int cls = blockIdx.x / 3, ncls = (gridDim.x - kd + 2) / 3;
  const int begin = static_cast<int>(static_cast<long long>(p.total) * cls / ncls);
  const int end = static_cast<int>(static_cast<long long>(p.total) * (cls + 1) / ncls);
  const int D = p.D;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kWHStages; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
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
  pdl_wait();      // barriers, TMEM and the role split are set up under the previous kernel's tail; global memory only from here on

  if (warp == 0) {
    // ===================== TMA producer: lane 0 = input plane, lane 1 = dY piece =====================
    if (lane < 2) {
      tma_prefetch_desc(lane == 0 ? &p.x_map : &p.dy_map);
    }
    int st = 0;
    uint32_t ph = 0;
    for (int piece = begin; piece < end; piece++) {
      const WHUnit u = wh_decode(p, piece);
      const int z = u.d + kd - 1;
      if (z < 0 || z >= D) continue;
      if (lane == 0) {
        mbar_wait_spin(&empty[st], ph ^ 1u, 5105);
        mbar_arrive_expect_tx(&full[st], kWHPitch * kWHRows * 128 + kWHDyBytes);
      }
      __syncwarp();
      if (lane == 0) tma_load_5d(smem + st * kWHStageBytes, &p.x_map, &full[st], 0, u.w0 - 1, u.h0 - 1, z, u.n);
      if (lane == 1) tma_load_5d(smem + st * kWHStageBytes + kWHPlaneBytes, &p.dy_map, &full[st], 0, u.w0, u.h0, u.d, u.n);
      if (++st == kWHStages) {
        st = 0;
        ph ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, true, true);
    // MN-major swizzled descriptors as 32-bit halves: low = start >> 4 | (LBO >> 4) << 16, high = SBO >> 4 | version | sw128
    const uint32_t a_hi = static_cast<uint32_t>(umma_smem_desc_sw128(0, 0, kWHPitch * 128u) >> 32);
    const uint32_t b_hi = static_cast<uint32_t>(umma_smem_desc_sw128(0, 0, 1024) >> 32);
    const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
    int st = 0;
    uint32_t ph = 0;
    uint32_t accum = 0;
    for (int piece = begin; piece < end; piece++) {
      const WHUnit u = wh_decode(p, piece);
      const int z = u.d + kd - 1;
      if (z < 0 || z >= D) continue;
      mbar_wait_spin(&full[st], ph, 5130);
      tc_fence_after();
      const uint32_t a_base = smem_lo + static_cast<uint32_t>(st) * (kWHStageBytes >> 4);
      const uint32_t b_base = a_base + (kWHPlaneBytes >> 4);
      if (elect_one_sync()) {
#pragma unroll
        for (int mt = 0; mt < 5; mt++) {
          const int ta = 2 * mt, tb = 2 * mt + 1;  // tb == 9 does not exist: that half of the tile is never stored
          const uint32_t lbo = static_cast<uint32_t>((tb < 9 ? wh_row(tb) - wh_row(ta) : 1) * 128) >> 4;
          const uint32_t a_lo = a_base + (static_cast<uint32_t>(wh_row(ta)) * 128u >> 4) + (lbo << 16);
#pragma unroll
          for (int ks = 0; ks < 8; ks++) {  // 16 positions = two h rows per MMA
            umma_bf16(tmem_base + static_cast<uint32_t>(mt * 64),
                      (static_cast<uint64_t>(a_hi) << 32) | (a_lo + ks * ((2 * kWHPitch * 128) >> 4)),
                      (static_cast<uint64_t>(b_hi) << 32) | (b_base + ks * (2048 >> 4)), idesc, accum | static_cast<uint32_t>(ks));
          }
        }
        umma_commit(&empty[st]);
      }
      __syncwarp();
      accum = 1;
      if (++st == kWHStages) {
        st = 0;
        ph ^= 1u;
      }
    }
    if (elect_one_sync()) umma_commit(tfull);
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5): 5 accumulators -> red.add into dw_tic, once per CTA =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;  // rows 0..63: first tap of the pair, 64..127: second tap
    bool any = false;               // did this CTA issue anything (a kd = 0 / 2 class of a one-plane volume may not)
    for (int piece = begin; piece < end && !any; piece++) {
      const int z = wh_decode(p, piece).d + kd - 1;
      any = z >= 0 && z < D;
    }
    mbar_wait(tfull, 0);
    tc_fence_after();
    if (any) {
#pragma unroll 1
      for (int mt = 0; mt < 5; mt++) {
        const int t9 = 2 * mt + (row >> 6);
        const bool keep = t9 < 9;
        float* dst = p.dw_tic + (static_cast<long long>(kd * 9 + t9) * 64 + (row & 63)) * 64;
#pragma unroll 1
        for (int chunk = 0; chunk < 2; chunk++) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mt * 64 + chunk * 32), v);
          tmem_ld_wait();
          if (keep) {
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// [27][64 cin][64 cout] fp32 -> [64 cout][27][64 cin] fp32
__global__ void wgrad_tic_to_oti_kernel(const float* __restrict__ tic, float* __restrict__ oti) {
  pdl_enter();
  __shared__ float tile[64][65];
  const int t = blockIdx.x;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) tile[i >> 6][i & 63] = tic[static_cast<long long>(t) * 4096 + i];
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int co = i >> 6, ci = i & 63;
    oti[(static_cast<long long>(co) * 27 + t) * 64 + ci] = tile[ci][co];
  }
}

}  // namespace

bool wgrad_halo_supported(const adni_conv3d_geom& g) {
  const char* e = getenv("ADNI_WGRAD_HALO");
  const char* det = getenv("ADNI_WGRAD_DETERMINISTIC");   // its CTAs share the scratch accumulator through red.add
  return g.k == 3 && g.stride == 1 && g.dil == 1 && g.pad == 1 && g.Cin == 64 && g.Cout == 64 && !(e && atoi(e) == 0) &&
         !(det && atoi(det) != 0);
}

// dw_tic: zeroed fp32 scratch [27][64][64]; dw_oti: output [64][27][64] (overwritten)
int launch_wgrad_halo(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw_tic,
                      float* dw_oti, cudaStream_t stream) {
  WgradHaloParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t dims[5] = {64, uint64_t(g.W), uint64_t(g.H), uint64_t(g.D), uint64_t(g.N)};
  const uint64_t strides[5] = {1, 64, uint64_t(g.W) * 64, uint64_t(g.H) * g.W * 64, uint64_t(g.D) * g.H * g.W * 64};
  {
    const uint32_t box[5] = {64, kWHPitch, kWHRows, 1, 1};
    int rc = make_tmap_bf16(&p.x_map, x, 5, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const uint32_t box[5] = {64, 8, 16, 1, 1};
    int rc = make_tmap_bf16(&p.dy_map, dy, 5, dims, strides, box, true);
    if (rc) return rc;
  }
  p.N = g.N;
  p.D = g.D;
  p.H = g.H;
  p.W = g.W;
  p.tiles_h = (g.H + 15) / 16;
  p.tiles_w = (g.W + 7) / 8;
  const long long total = (long long)g.N * p.tiles_h * p.tiles_w * g.D;
  if (total > 0x7fffffffLL) return ADNI_ENOTSUP;
  p.total = int(total);
  p.dw_tic = dw_tic;
  static bool attr = false;
  if (!attr) {
    ADNI_CUDA_OK(cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWHSmem));
    attr = true;
  }
  int grid = num_sms();
  if (grid > 3 * p.total) grid = 3 * p.total;
  if (grid < 3) grid = 3;
  pdl_launch(wgrad_halo_kernel, grid, kWHThreads, kWHSmem, stream)(p);
  count_launch();
  ADNI_LAUNCH_CHECK("wgrad_halo_kernel");
  pdl_launch(wgrad_tic_to_oti_kernel, 27, 256, 0, stream)(dw_tic, dw_oti);
  count_launch();
  ADNI_LAUNCH_CHECK("wgrad_tic_to_oti_kernel");
  return ADNI_OK;
}

}  // namespace adni

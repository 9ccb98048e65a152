// Volume-level fusion operators of the early- and feature-map-fusion models:
//   * multi-channel module input (B, C, D, H, W) fp32/fp64 NCDHW -> bf16 NDHWC
//     (pkg/models/fusion_models/early_fusion.py:84-88: torch.stack((x_pet, x_mri), dim=1).to(float32))
//   * voxel-wise maximum of two feature maps and its gradient routing
//     (pkg/models/fusion_models/anat_pet_featuremapfusion.py:121-123: torch.max(torch.stack((pet, mri)), dim=0))
//   * channel concatenation of two feature maps and the split of its gradient (:118-119, torch.cat(dim=1))
// All HBM-bound, one pass, 16-byte accesses where the channel counts allow.
#include "common.cuh"

namespace adni {
extern void count_launch();
namespace {

constexpr int kThreads = 256;

inline int grid_for(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// x [N][C][vox] -> out [N][vox][C]; one thread per (n, voxel): C coalesced plane reads, one contiguous 2C-byte write
template <typename T, int C>
__global__ void __launch_bounds__(kThreads) to_ndhwc_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                            int N, long long vox) {
  pdl_enter();
  const long long total = (long long)N * vox;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / vox, v = i - n * vox;
    const T* src = x + n * C * vox + v;
    __nv_bfloat16 o[C];
#pragma unroll
    for (int c = 0; c < C; c++) o[c] = __float2bfloat16_rn((float)__ldcs(src + c * vox));
    __nv_bfloat16* dst = out + i * C;
    if constexpr (C == 2) {
      *reinterpret_cast<__nv_bfloat162*>(dst) = __nv_bfloat162(o[0], o[1]);
    } else if constexpr (C == 4) {
      uint2 u;
      u.x = (uint32_t)__bfloat16_as_ushort(o[0]) | ((uint32_t)__bfloat16_as_ushort(o[1]) << 16);
      u.y = (uint32_t)__bfloat16_as_ushort(o[2]) | ((uint32_t)__bfloat16_as_ushort(o[3]) << 16);
      *reinterpret_cast<uint2*>(dst) = u;
    } else {
#pragma unroll
      for (int c = 0; c < C; c++) dst[c] = o[c];
    }
  }
}

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// out = a >= b ? a : b  (ties -> a: torch.max over a stacked dim returns the FIRST maximal index, and so routes the
// gradient of a tie - e.g. two post-ReLU zeros - to the first operand, the PET branch)
__global__ void __launch_bounds__(kThreads) maxout_fwd_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                              uint4* __restrict__ out, long long nvec) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 A = __ldcs(a + i), B = __ldcs(b + i);
    uint4 O;
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(&A);
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(&B);
    uint32_t* po = reinterpret_cast<uint32_t*>(&O);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t lo = bf_lo(pa[j]) >= bf_lo(pb[j]) ? (pa[j] & 0xffffu) : (pb[j] & 0xffffu);
      const uint32_t hi = bf_hi(pa[j]) >= bf_hi(pb[j]) ? (pa[j] & 0xffff0000u) : (pb[j] & 0xffff0000u);
      po[j] = lo | hi;
    }
    out[i] = O;
  }
}

__global__ void __launch_bounds__(kThreads) maxout_bwd_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ a,
                                                              const uint4* __restrict__ b, uint4* __restrict__ da,
                                                              uint4* __restrict__ db, long long nvec) {
  pdl_enter();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 G = __ldcs(dout + i), A = __ldcs(a + i), B = __ldcs(b + i);
    uint4 DA, DB;
    const uint32_t* pg = reinterpret_cast<const uint32_t*>(&G);
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(&A);
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(&B);
    uint32_t* qa = reinterpret_cast<uint32_t*>(&DA);
    uint32_t* qb = reinterpret_cast<uint32_t*>(&DB);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t mlo = bf_lo(pa[j]) >= bf_lo(pb[j]) ? 0xffffu : 0u;
      const uint32_t mhi = bf_hi(pa[j]) >= bf_hi(pb[j]) ? 0xffff0000u : 0u;
      const uint32_t m = mlo | mhi;
      qa[j] = pg[j] & m;
      qb[j] = pg[j] & ~m;
    }
    if (da) da[i] = DA;
    if (db) db[i] = DB;
  }
}

// out[row] = [a[row] (Ca) | b[row] (Cb)], 8-channel (16-byte) groups; SPLIT: the inverse for the gradient
template <bool SPLIT>
__global__ void __launch_bounds__(kThreads) concat_kernel(uint4* __restrict__ a, uint4* __restrict__ b,
                                                          uint4* __restrict__ cat, long long rows, int ga, int gb) {
  pdl_enter();
  const int gc = ga + gb;
  const long long total = rows * gc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / gc;
    const int g = (int)(i - r * gc);
    uint4* side = g < ga ? (a ? a + r * ga + g : nullptr) : (b ? b + r * gb + (g - ga) : nullptr);
    if (!side) continue;
    if (SPLIT) *side = __ldcs(cat + i); else cat[i] = __ldcs(side);
  }
}

// Zero padding on the HIGH side of D, H, W (torch's padding='same' with an even kernel pads total = dil*(k-1) as
// lo = total/2, hi = total - lo; the conv engines take the symmetric part, this kernel supplies the extra voxel):
// big = [N][D+ed][H+eh][W+ew][C], small = [N][D][H][W][C].  CROP = false: big <- small with a zero border (forward);
// CROP = true: small <- leading box of big (gradient).  VEC = channels per thread access (8 = 16 bytes, or 1).
template <bool CROP, int VEC>
__global__ void __launch_bounds__(kThreads) pad_high_kernel(const __nv_bfloat16* __restrict__ src,
                                                            __nv_bfloat16* __restrict__ dst, int N, int D, int H, int W,
                                                            int C, int ed, int eh, int ew) {
  pdl_enter();
  const int cv = C / VEC;
  const int Dd = CROP ? D : D + ed, Hd = CROP ? H : H + eh, Wd = CROP ? W : W + ew;   // destination extents
  const int Ds = CROP ? D + ed : D, Hs = CROP ? H + eh : H, Ws = CROP ? W + ew : W;   // source extents
  const long long total = (long long)N * Dd * Hd * Wd * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int c = (int)(r % cv);
    r /= cv;
    const int w = (int)(r % Wd);
    r /= Wd;
    const int h = (int)(r % Hd);
    r /= Hd;
    const int d = (int)(r % Dd);
    const int n = (int)(r / Dd);
    const bool inside = d < Ds && h < Hs && w < Ws;  // always true for CROP
    const long long so = ((((long long)n * Ds + d) * Hs + h) * Ws + w) * cv + c;
    if (VEC == 8) {
      const uint4 v = inside ? __ldg(reinterpret_cast<const uint4*>(src) + so) : make_uint4(0, 0, 0, 0);
      reinterpret_cast<uint4*>(dst)[i] = v;
    } else {
      dst[i] = inside ? src[so] : __float2bfloat16_rn(0.f);
    }
  }
}

}  // namespace
}  // namespace adni

using namespace adni;
#define ST(s) static_cast<cudaStream_t>(s)

template <bool CROP>
static int launch_pad_high(const __nv_bfloat16* src, __nv_bfloat16* dst, int N, int D, int H, int W, int C, int ed, int eh,
                           int ew, cudaStream_t st) {
  const long long elems = (long long)N * (CROP ? D : D + ed) * (CROP ? H : H + eh) * (CROP ? W : W + ew) * C;
  if (C % 8 == 0) {
    pdl_launch(pad_high_kernel<CROP, 8>, grid_for(elems / 8, kThreads * 2), kThreads, 0, st)(src, dst, N, D, H, W, C, ed, eh, ew);
  } else {
    pdl_launch(pad_high_kernel<CROP, 1>, grid_for(elems, kThreads * 4), kThreads, 0, st)(src, dst, N, D, H, W, C, ed, eh, ew);
  }
  count_launch();
  ADNI_LAUNCH_CHECK("pad_high_kernel");
  return ADNI_OK;
}

template <typename T>
static int launch_to_ndhwc(const T* x, int N, int C, long long vox, __nv_bfloat16* out, cudaStream_t st) {
  const int grid = grid_for((long long)N * vox, kThreads * 4);
  switch (C) {
    case 2: pdl_launch(to_ndhwc_kernel<T, 2>, grid, kThreads, 0, st)(x, out, N, vox); break;
    case 3: pdl_launch(to_ndhwc_kernel<T, 3>, grid, kThreads, 0, st)(x, out, N, vox); break;
    case 4: pdl_launch(to_ndhwc_kernel<T, 4>, grid, kThreads, 0, st)(x, out, N, vox); break;
    default:
      set_error("volumes_to_ndhwc: C=%d is not supported (2, 3 or 4 input modalities)", C);
      return ADNI_ENOTSUP;
  }
  count_launch();
  ADNI_LAUNCH_CHECK("to_ndhwc_kernel");
  return ADNI_OK;
}

extern "C" {

int adni_volumes_to_ndhwc_bf16(const void* x, int x_is_f64, int N, int C, long long vox, adni_bf16* out, void* stream) {
  ADNI_REQUIRE(x && out && N > 0 && C > 0 && vox > 0, ADNI_EINVAL, "volumes_to_ndhwc: bad arguments");
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  return x_is_f64 ? launch_to_ndhwc(static_cast<const double*>(x), N, C, vox, o, ST(stream))
                  : launch_to_ndhwc(static_cast<const float*>(x), N, C, vox, o, ST(stream));
}

int adni_pad_volume_high(const adni_bf16* x, int N, int D, int H, int W, int C, int ed, int eh, int ew, adni_bf16* out,
                         void* stream) {
  ADNI_REQUIRE(x && out && N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && ed >= 0 && eh >= 0 && ew >= 0, ADNI_EINVAL,
               "pad_volume_high: bad arguments");
  return launch_pad_high<false>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(out), N, D, H,
                                W, C, ed, eh, ew, ST(stream));
}

int adni_crop_volume_high(const adni_bf16* x_padded, int N, int D, int H, int W, int C, int ed, int eh, int ew,
                          adni_bf16* out, void* stream) {
  ADNI_REQUIRE(x_padded && out && N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && ed >= 0 && eh >= 0 && ew >= 0, ADNI_EINVAL,
               "crop_volume_high: bad arguments");
  return launch_pad_high<true>(reinterpret_cast<const __nv_bfloat16*>(x_padded), reinterpret_cast<__nv_bfloat16*>(out), N,
                               D, H, W, C, ed, eh, ew, ST(stream));
}

int adni_maxout_fwd(const adni_bf16* a, const adni_bf16* b, adni_bf16* out, long long n, void* stream) {
  ADNI_REQUIRE(a && b && out && n > 0, ADNI_EINVAL, "maxout_fwd: bad arguments");
  ADNI_REQUIRE(n % 8 == 0, ADNI_ENOTSUP, "maxout_fwd: n=%lld must be a multiple of 8 (channels are)", n);
  pdl_launch(maxout_fwd_kernel, grid_for(n / 8, kThreads * 2), kThreads, 0, ST(stream))(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(out), n / 8);
  count_launch();
  ADNI_LAUNCH_CHECK("maxout_fwd_kernel");
  return ADNI_OK;
}

int adni_maxout_bwd(const adni_bf16* dout, const adni_bf16* a, const adni_bf16* b, adni_bf16* da, adni_bf16* db,
                    long long n, void* stream) {
  ADNI_REQUIRE(dout && a && b && (da || db) && n > 0, ADNI_EINVAL, "maxout_bwd: bad arguments");
  ADNI_REQUIRE(n % 8 == 0, ADNI_ENOTSUP, "maxout_bwd: n=%lld must be a multiple of 8 (channels are)", n);
  pdl_launch(maxout_bwd_kernel, grid_for(n / 8, kThreads * 2), kThreads, 0, ST(stream))(
      reinterpret_cast<const uint4*>(dout), reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b),
      reinterpret_cast<uint4*>(da), reinterpret_cast<uint4*>(db), n / 8);
  count_launch();
  ADNI_LAUNCH_CHECK("maxout_bwd_kernel");
  return ADNI_OK;
}

int adni_concat_channels(const adni_bf16* a, int Ca, const adni_bf16* b, int Cb, long long rows, adni_bf16* out,
                         void* stream) {
  ADNI_REQUIRE(a && b && out && rows > 0 && Ca > 0 && Cb > 0, ADNI_EINVAL, "concat_channels: bad arguments");
  ADNI_REQUIRE(Ca % 8 == 0 && Cb % 8 == 0, ADNI_ENOTSUP, "concat_channels: Ca=%d, Cb=%d must be multiples of 8", Ca, Cb);
  pdl_launch(concat_kernel<false>, grid_for(rows * ((Ca + Cb) / 8), kThreads * 2), kThreads, 0, ST(stream))(
      reinterpret_cast<uint4*>(const_cast<adni_bf16*>(a)), reinterpret_cast<uint4*>(const_cast<adni_bf16*>(b)),
      reinterpret_cast<uint4*>(out), rows, Ca / 8, Cb / 8);
  count_launch();
  ADNI_LAUNCH_CHECK("concat_kernel");
  return ADNI_OK;
}

int adni_split_channels(const adni_bf16* dout, int Ca, int Cb, long long rows, adni_bf16* da, adni_bf16* db,
                        void* stream) {
  ADNI_REQUIRE(dout && (da || db) && rows > 0 && Ca > 0 && Cb > 0, ADNI_EINVAL, "split_channels: bad arguments");
  ADNI_REQUIRE(Ca % 8 == 0 && Cb % 8 == 0, ADNI_ENOTSUP, "split_channels: Ca=%d, Cb=%d must be multiples of 8", Ca, Cb);
  pdl_launch(concat_kernel<true>, grid_for(rows * ((Ca + Cb) / 8), kThreads * 2), kThreads, 0, ST(stream))(
      reinterpret_cast<uint4*>(da), reinterpret_cast<uint4*>(db),
      reinterpret_cast<uint4*>(const_cast<adni_bf16*>(dout)), rows, Ca / 8, Cb / 8);
  count_launch();
  ADNI_LAUNCH_CHECK("concat_kernel");
  return ADNI_OK;
}

}  // extern "C"

// Input normalisation on the device.
//
// Per-scan quantile min-max (pkg/utils/dataloader.py:239-249,261-270): the reference sorts the non-zero
// masked voxels twice on the CPU in fp64 (torch.quantile).  Here the four order statistics
// (floor/ceil rank for q and for 1-q) are found EXACTLY by a 3-pass MSD radix select (11+11+10 bits) on
// the monotone integer image of the fp32 intensities, all ranks and interpolation weights in fp64, so the
// rank indices are bit-identical to the reference and Qmin/Qmax reproduce torch's lerp formula.
// Also: PET standardisation (dataloader.py:213-215), per-scan z-score statistics (:252-256) and the
// split moments of pkg/utils/standardization.py:34-55.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace adni {
extern void count_launch();
namespace {

constexpr int kBins = 2048;
constexpr int kHistThreads = 256;

struct ScanState {
  long long n;
  long long rank[4];   // lo(q), hi(q), lo(1-q), hi(1-q)
  long long resid[4];  // rank inside the bucket selected so far
  unsigned int prefix[4];
  unsigned int pad_[4];
  double w[2];   // interpolation weights for q and 1-q
  double qv[2];  // Qmax, Qmin
};
constexpr size_t kStateBytes = 256;
constexpr size_t kScanWsBytes = kStateBytes + 4 * kBins * sizeof(unsigned int);
static_assert(sizeof(ScanState) <= kStateBytes, "ScanState too large");

__device__ __forceinline__ unsigned int f2key(float v) {
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
__device__ __forceinline__ ScanState* state_of(void* ws, int scan) {
  return reinterpret_cast<ScanState*>(static_cast<char*>(ws) + (size_t)scan * kScanWsBytes);
}
__device__ __forceinline__ unsigned int* hist_of(void* ws, int scan) {
  return reinterpret_cast<unsigned int*>(static_cast<char*>(ws) + (size_t)scan * kScanWsBytes + kStateBytes);
}

// Histogram pass of the radix select.  PASS 0 (11 leading bits, every non-zero masked voxel counts): shared-memory
// histogram per block.  PASS 1 / 2 (only voxels inside the bucket chosen so far count - typically < 1 %): no shared
// histogram to clear and flush (that fixed cost made these passes 3x slower than pass 0); the few matches go to the
// global histogram directly, one atomic per distinct bin and warp (__match_any), so a constant image does not
// serialise on one address.  Loads are 16-byte (4 voxels + 4 mask bytes per thread and iteration).
template <int PASS>
__global__ void __launch_bounds__(kHistThreads) q_hist_kernel(const float* __restrict__ x,
                                                             const uint8_t* __restrict__ mask, long long nvox,
                                                             void* ws) {
  pdl_enter();
  __shared__ unsigned int h[PASS == 0 ? kBins : 1];
  const int scan = blockIdx.y;
  if (PASS == 0)
    for (int i = threadIdx.x; i < kBins; i += kHistThreads) h[i] = 0;
  unsigned int pre[4] = {0, 0, 0, 0};
  if (PASS > 0) {
    const ScanState* st = state_of(ws, scan);
#pragma unroll
    for (int j = 0; j < 4; j++) pre[j] = st->prefix[j];
  }
  __syncthreads();
  unsigned int* gh = hist_of(ws, scan);
  const float* xs = x + (long long)scan * nvox;
  const uint8_t* ms = mask + (long long)scan * nvox;
  auto count = [&](float xv, uint8_t mv) {
    const float v = mv ? xv : 0.f;
    if (v == 0.f) return;
    const unsigned int key = f2key(v);
    if (PASS == 0) {
      atomicAdd(&h[key >> 21], 1u);
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const bool hit = PASS == 1 ? (key >> 21) == (pre[j] >> 21) : (key >> 10) == (pre[j] >> 10);
        if (hit) {
          const unsigned int bin = PASS == 1 ? ((key >> 10) & 2047u) : (key & 1023u);
          const unsigned int act = __activemask();
          const unsigned int grp = __match_any_sync(act, bin);
          if ((threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&gh[j * kBins + bin], (unsigned int)__popc(grp));
        }
      }
    }
  };
  const long long stride = (long long)gridDim.x * kHistThreads;
  const long long tid = (long long)blockIdx.x * kHistThreads + threadIdx.x;
  if ((nvox & 3) == 0 && (reinterpret_cast<uintptr_t>(xs) & 15) == 0 && (reinterpret_cast<uintptr_t>(ms) & 3) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(xs);
    const uchar4* m4 = reinterpret_cast<const uchar4*>(ms);
    for (long long i = tid; i < (nvox >> 2); i += stride) {
      const float4 xv = __ldg(x4 + i);
      const uchar4 mv = __ldg(m4 + i);
      count(xv.x, mv.x);
      count(xv.y, mv.y);
      count(xv.z, mv.z);
      count(xv.w, mv.w);
    }
  } else {
    for (long long i = tid; i < nvox; i += stride) count(xs[i], ms[i]);
  }
  if (PASS == 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += kHistThreads)
      if (h[i]) atomicAdd(&gh[i], h[i]);
  }
}

// one block of 4 warps per scan; warp j resolves rank j
template <int PASS>
__global__ void __launch_bounds__(128) q_select_kernel(void* ws, double q) {
  pdl_enter();
  const int scan = blockIdx.x;
  ScanState* st = state_of(ws, scan);
  unsigned int* gh = hist_of(ws, scan);
  const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NB = PASS == 2 ? 1024 : kBins;
  constexpr int SHIFT = PASS == 0 ? 21 : (PASS == 1 ? 10 : 0);
  const unsigned int* h = gh + (PASS == 0 ? 0 : j * kBins);
  constexpr int SEG = NB / 32;
  unsigned long long seg = 0;
  for (int b = 0; b < SEG; b++) seg += h[lane * SEG + b];
  unsigned long long incl = seg;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  long long r;
  if (PASS == 0) {
    const long long n = (long long)__shfl_sync(0xffffffffu, incl, 31);
    // rank arithmetic exactly as torch.quantile(..., interpolation='linear') on an fp64 input:
    //   pos = q * (n - 1); lo = floor(pos); hi = ceil(pos); w = pos - lo
    const double qq = (j < 2) ? q : (1.0 - q);
    const double pos = qq * (double)(n - 1);
    const double lo = floor(pos), hi = ceil(pos);
    r = (long long)((j & 1) ? hi : lo);
    if (lane == 0) {
      st->rank[j] = r;
      if ((j & 1) == 0) st->w[j >> 1] = pos - lo;
      if (j == 0) st->n = n;
    }
    if (n <= 0) r = 0;
  } else {
    r = st->resid[j];
  }
  const unsigned long long excl = incl - seg;
  const bool mine = (unsigned long long)r >= excl && (unsigned long long)r < incl;
  const unsigned int ball = __ballot_sync(0xffffffffu, mine);
  if (mine) {
    unsigned long long cum = excl;
    int b = 0;
    for (; b < SEG; b++) {
      const unsigned long long c = h[lane * SEG + b];
      if ((unsigned long long)r < cum + c) break;
      cum += c;
    }
    const unsigned int bin = (unsigned int)(lane * SEG + b);
    const unsigned int base = PASS == 0 ? 0u : st->prefix[j];
    st->prefix[j] = base | (bin << SHIFT);
    st->resid[j] = r - (long long)cum;
  } else if (ball == 0 && lane == 0) {
    st->prefix[j] = 0x7FC00000u;  // empty selection (n == 0): decodes to a NaN-producing key below
    st->resid[j] = 0;
  }
  __syncthreads();
  // reset histograms for the next pass / next call
  for (int i = threadIdx.x; i < 4 * kBins; i += 128) gh[i] = 0;
  if (PASS == 2 && threadIdx.x == 0) {
    if (st->n <= 0) {
      st->qv[0] = nan("");
      st->qv[1] = nan("");
    } else {
      for (int s = 0; s < 2; s++) {
        const double a = (double)key2f(st->prefix[2 * s]);
        const double b = (double)key2f(st->prefix[2 * s + 1]);
        const double w = st->w[s];
        // at::lerp for real types: w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
        st->qv[s] = (w < 0.5) ? (a + w * (b - a)) : (b - (b - a) * (1.0 - w));
      }
    }
  }
}

__global__ void __launch_bounds__(256) q_apply_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask,
                                                      long long nvox, const void* ws, float* __restrict__ of,
                                                      __nv_bfloat16* __restrict__ ob) {
  pdl_enter();
  const int scan = blockIdx.y;
  const ScanState* st = state_of(const_cast<void*>(ws), scan);
  const double qmax = st->qv[0], qmin = st->qv[1];
  const double range = qmax - qmin;
  const long long base = (long long)scan * nvox;
  auto norm = [&](float xv, uint8_t mv) -> float {  // the reference's fp64 sequence, rounded to fp32 once
    double v = ((double)xv - qmin) / range;
    if (v > 1.0) v = 1.0;
    if (v < 0.0) v = 0.0;
    v *= mv ? 1.0 : 0.0;
    return (float)v;
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = (nvox & 3) == 0 && (reinterpret_cast<uintptr_t>(x + base) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(mask + base) & 3) == 0 &&
                   (!of || (reinterpret_cast<uintptr_t>(of + base) & 15) == 0) &&
                   (!ob || (reinterpret_cast<uintptr_t>(ob + base) & 7) == 0);
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(x + base);
    const uchar4* m4 = reinterpret_cast<const uchar4*>(mask + base);
    for (long long i = tid; i < (nvox >> 2); i += stride) {
      const float4 xv = __ldg(x4 + i);
      const uchar4 mv = __ldg(m4 + i);
      float4 f;
      f.x = norm(xv.x, mv.x);
      f.y = norm(xv.y, mv.y);
      f.z = norm(xv.z, mv.z);
      f.w = norm(xv.w, mv.w);
      if (of) reinterpret_cast<float4*>(of + base)[i] = f;
      if (ob) {
        uint2 o;
        o.x = pack_bf16x2(f.x, f.y);
        o.y = pack_bf16x2(f.z, f.w);
        reinterpret_cast<uint2*>(ob + base)[i] = o;
      }
    }
  } else {
    for (long long i = tid; i < nvox; i += stride) {
      const float f = norm(x[base + i], mask[base + i]);
      if (of) of[base + i] = f;
      if (ob) ob[base + i] = __float2bfloat16_rn(f);
    }
  }
}

__global__ void q_export_kernel(const void* ws, int nscans, long long* info, double* qvals) {
  pdl_enter();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nscans) return;
  const ScanState* st = state_of(const_cast<void*>(ws), s);
  if (info) {
    info[8 * s + 0] = st->n;
    for (int j = 0; j < 4; j++) info[8 * s + 1 + j] = st->rank[j];
    info[8 * s + 5] = info[8 * s + 6] = info[8 * s + 7] = 0;
  }
  if (qvals) {
    qvals[2 * s + 0] = st->qv[0];
    qvals[2 * s + 1] = st->qv[1];
  }
}

__global__ void __launch_bounds__(256) standardize_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask,
                                                          long long n, double mean, double stdv,
                                                          float* __restrict__ of, __nv_bfloat16* __restrict__ ob) {
  pdl_enter();
  auto norm = [&](float xv, uint8_t mv) -> float {  // the reference's fp64 sequence, rounded to fp32 once
    double v = ((double)xv - mean) / stdv;
    v *= mv ? 1.0 : 0.0;
    return (float)v;
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (!mask || (reinterpret_cast<uintptr_t>(mask) & 3) == 0) &&
                   (!of || (reinterpret_cast<uintptr_t>(of) & 15) == 0) &&
                   (!ob || (reinterpret_cast<uintptr_t>(ob) & 7) == 0);
  long long done = 0;
  if (vec) {  // 16-byte loads, four independent fp64 divisions in flight per thread
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const uchar4* m4 = reinterpret_cast<const uchar4*>(mask);
    const long long n4 = n >> 2;
    for (long long i = tid; i < n4; i += stride) {
      const float4 xv = __ldcs(x4 + i);
      const uchar4 mv = mask ? __ldcs(m4 + i) : make_uchar4(1, 1, 1, 1);
      float4 f;
      f.x = norm(xv.x, mv.x);
      f.y = norm(xv.y, mv.y);
      f.z = norm(xv.z, mv.z);
      f.w = norm(xv.w, mv.w);
      if (of) reinterpret_cast<float4*>(of)[i] = f;
      if (ob) {
        uint2 o;
        o.x = pack_bf16x2(f.x, f.y);
        o.y = pack_bf16x2(f.z, f.w);
        reinterpret_cast<uint2*>(ob)[i] = o;
      }
    }
    done = n4 << 2;
  }
  for (long long i = done + tid; i < n; i += stride) {
    const float f = norm(x[i], mask ? mask[i] : (uint8_t)1);
    if (of) of[i] = f;
    if (ob) ob[i] = __float2bfloat16_rn(f);
  }
}

// block-level fp64 reduction helper: returns the block total in thread 0
__device__ __forceinline__ double block_sum(double v, double* sm) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double t = 0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (blockDim.x + 31) / 32; w++) t += sm[w];
  return t;
}

// moments[2s] += sum x / nvox ; moments[2s+1] += sum x^2 / nvox
__global__ void __launch_bounds__(256) scan_moments_kernel(const float* __restrict__ x, long long nvox,
                                                           double* __restrict__ moments) {
  pdl_enter();
  __shared__ double sm[8];
  const int scan = blockIdx.y;
  const float* xs = x + (long long)scan * nvox;
  double s = 0, q = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = xs[i];
    s += v;
    q += v * v;
  }
  const double ts = block_sum(s, sm);
  const double tq = block_sum(q, sm);
  if (threadIdx.x == 0) {
    atomicAdd(moments + 2 * scan, ts / (double)nvox);
    atomicAdd(moments + 2 * scan + 1, tq / (double)nvox);
  }
}

// pass 0: out[3s] += count, out[3s+1] += sum  (over non-zero masked voxels)
// pass 1: out[3s+2] += sum (x - mean)^2 with mean = out[3s+1]/out[3s]
template <int PASS>
__global__ void __launch_bounds__(256) masked_moments_kernel(const float* __restrict__ x,
                                                             const uint8_t* __restrict__ mask, long long nvox,
                                                             double* __restrict__ out) {
  pdl_enter();
  __shared__ double sm[8];
  const int scan = blockIdx.y;
  const float* xs = x + (long long)scan * nvox;
  const uint8_t* ms = mask + (long long)scan * nvox;
  const double mean = PASS == 1 ? out[3 * scan + 1] / out[3 * scan] : 0.0;
  double a = 0, b = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = ms[i] ? xs[i] : 0.f;
    if (v != 0.f) {
      if (PASS == 0) {
        a += 1.0;
        b += (double)v;
      } else {
        const double d = (double)v - mean;
        a += d * d;
      }
    }
  }
  const double ta = block_sum(a, sm);
  const double tb = PASS == 0 ? block_sum(b, sm) : 0.0;
  if (threadIdx.x == 0) {
    if (PASS == 0) {
      atomicAdd(out + 3 * scan, ta);
      atomicAdd(out + 3 * scan + 1, tb);
    } else {
      atomicAdd(out + 3 * scan + 2, ta);
    }
  }
}

__global__ void masked_finalize_kernel(double* out, int nscans) {
  pdl_enter();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nscans) return;
  const double n = out[3 * s];
  out[3 * s + 1] = out[3 * s + 1] / n;
  out[3 * s + 2] = sqrt(out[3 * s + 2] / (n - 1.0));
}

inline int scan_blocks(long long nvox, int nscans) {
  long long b = (nvox + 256 * 16 - 1) / (256 * 16);
  const long long cap = std::max(1, (num_sms() * 8) / std::max(nscans, 1));
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace
}  // namespace adni

using namespace adni;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

size_t adni_quantile_workspace_bytes(int nscans) { return (size_t)(nscans > 0 ? nscans : 0) * kScanWsBytes; }

int adni_quantile_minmax_normalize(const float* x, const uint8_t* mask, int nscans, long long nvox, double q,
                                   float* out_f32, adni_bf16* out_bf16, long long* info, double* qvals, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  ADNI_REQUIRE(x && mask && workspace && nscans > 0 && nvox > 0, ADNI_EINVAL, "quantile_minmax_normalize: bad arguments");
  ADNI_REQUIRE(q >= 0.0 && q <= 1.0, ADNI_EINVAL, "quantile_minmax_normalize: q=%f outside [0,1]", q);
  ADNI_REQUIRE(nscans <= 65535, ADNI_ENOTSUP, "quantile_minmax_normalize: more than 65535 scans per call");
  ADNI_REQUIRE(workspace_bytes >= adni_quantile_workspace_bytes(nscans), ADNI_ENOMEM,
               "quantile_minmax_normalize: workspace too small");
  cudaStream_t st = ST(stream);
  ADNI_CUDA_OK(cudaMemsetAsync(workspace, 0, adni_quantile_workspace_bytes(nscans), st));
  dim3 grid(scan_blocks(nvox, nscans), nscans);
  pdl_launch(q_hist_kernel<0>, grid, kHistThreads, 0, st)(x, mask, nvox, workspace);
  pdl_launch(q_select_kernel<0>, nscans, 128, 0, st)(workspace, q);
  pdl_launch(q_hist_kernel<1>, grid, kHistThreads, 0, st)(x, mask, nvox, workspace);
  pdl_launch(q_select_kernel<1>, nscans, 128, 0, st)(workspace, q);
  pdl_launch(q_hist_kernel<2>, grid, kHistThreads, 0, st)(x, mask, nvox, workspace);
  pdl_launch(q_select_kernel<2>, nscans, 128, 0, st)(workspace, q);
  for (int i = 0; i < 6; i++) count_launch();
  ADNI_LAUNCH_CHECK("quantile select");
  if (out_f32 || out_bf16) {
    pdl_launch(q_apply_kernel, grid, 256, 0, st)(x, mask, nvox, workspace, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16));
    count_launch();
  }
  if (info || qvals) {
    pdl_launch(q_export_kernel, (nscans + 127) / 128, 128, 0, st)(workspace, nscans, info, qvals);
    count_launch();
  }
  ADNI_LAUNCH_CHECK("quantile apply");
  return ADNI_OK;
}

int adni_standardize(const float* x, const uint8_t* mask, long long n, double mean, double std, float* out_f32,
                     adni_bf16* out_bf16, void* stream) {
  ADNI_REQUIRE(x && n > 0 && (out_f32 || out_bf16), ADNI_EINVAL, "standardize: bad arguments");
  const int grid = (int)std::min<long long>((n + 256 * 16 - 1) / (256 * 16), (long long)num_sms() * 8);
  pdl_launch(standardize_kernel, grid, 256, 0, ST(stream))(x, mask, n, mean, std, out_f32,
                                                   reinterpret_cast<__nv_bfloat16*>(out_bf16));
  count_launch();
  ADNI_LAUNCH_CHECK("standardize_kernel");
  return ADNI_OK;
}

int adni_scan_moments(const float* x, int nscans, long long nvox, double* moments, void* stream) {
  ADNI_REQUIRE(x && moments && nscans > 0 && nscans <= 65535 && nvox > 0, ADNI_EINVAL, "scan_moments: bad arguments");
  dim3 grid(scan_blocks(nvox, nscans), nscans);
  pdl_launch(scan_moments_kernel, grid, 256, 0, ST(stream))(x, nvox, moments);
  count_launch();
  ADNI_LAUNCH_CHECK("scan_moments_kernel");
  return ADNI_OK;
}

int adni_masked_std_mean(const float* x, const uint8_t* mask, int nscans, long long nvox, double* out, void* stream) {
  ADNI_REQUIRE(x && mask && out && nscans > 0 && nscans <= 65535 && nvox > 0, ADNI_EINVAL,
               "masked_std_mean: bad arguments");
  cudaStream_t st = ST(stream);
  ADNI_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double) * 3 * (size_t)nscans, st));
  dim3 grid(scan_blocks(nvox, nscans), nscans);
  pdl_launch(masked_moments_kernel<0>, grid, 256, 0, st)(x, mask, nvox, out);
  pdl_launch(masked_moments_kernel<1>, grid, 256, 0, st)(x, mask, nvox, out);
  pdl_launch(masked_finalize_kernel, (nscans + 127) / 128, 128, 0, st)(out, nscans);
  for (int i = 0; i < 3; i++) count_launch();
  ADNI_LAUNCH_CHECK("masked_std_mean");
  return ADNI_OK;
}

}  // extern "C"

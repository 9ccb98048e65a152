// Host-side planning + C-ABI entry points for Conv3d fprop / dgrad / wgrad.
//
// The planner turns a conv geometry into (tensor-map views, tap table, output view, box shape) for the
// tcgen05 kernels in conv_igemm_kernels.cu, or dispatches to the CUDA-core direct engine
// (conv_direct.cu) for the small-channel convolutions of Small_PET_CNN and the 1-channel stem.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "conv_igemm.cuh"

namespace adni {

int launch_igemm(const IgemmParams& p, int block_n, cudaStream_t stream);
int launch_igemm_multi(const IgemmMulti& pm, int block_n, cudaStream_t stream);
int launch_wgrad2(const WgradParams& p, int mt_cfg, cudaStream_t stream, const W2Sched* sk = nullptr);
int wgrad2_box_rows(int mt_cfg);
bool wgrad_halo_supported(const adni_conv3d_geom& g);
int launch_wgrad_halo(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw_tic, float* dw_oti,
                      cudaStream_t stream);
int halo_pitch();
int halo_rows();
int launch_igemm_halo(const HaloParams& p, int channels, cudaStream_t stream);

// conv_small.cu: warp-level tensor-core engine (mma.sync) for the small-channel stacks (Small_PET_CNN & co.)
bool small_conv_supported(const adni_conv3d_geom& g, int pass);
int small_conv_fprop(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* w_oti, const float* bias,
                     __nv_bfloat16* y, double* ssum, double* ssq, cudaStream_t stream);
int small_conv_dgrad(const adni_conv3d_geom& g, const __nv_bfloat16* dy, const __nv_bfloat16* w_ito, __nv_bfloat16* dx,
                     cudaStream_t stream);
int small_conv_wgrad(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw,
                     cudaStream_t stream);

int direct_conv_fprop(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* w_oti,
                      const float* bias, __nv_bfloat16* y, double* ssum, double* ssq, cudaStream_t stream);
int direct_conv_dgrad(const adni_conv3d_geom& g, const __nv_bfloat16* dy, const __nv_bfloat16* w_ito,
                      const __nv_bfloat16* addend, __nv_bfloat16* dx, cudaStream_t stream);
int direct_conv_wgrad(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw,
                      float* dbias, cudaStream_t stream);

void finish_igemm_params(IgemmParams& p) {
  const long long total = static_cast<long long>(p.N) * p.tiles_d * p.tiles_h * p.tiles_w * p.n_tiles;
  const int d[4] = {p.n_tiles, p.tiles_w, p.tiles_h, p.tiles_d};
  p.fd_ok = 1;
  for (int i = 0; i < 4; i++) {
    if (d[i] <= 0 || total * d[i] >= (1LL << 32)) p.fd_ok = 0;
    p.fd_mul[i] = d[i] <= 1 ? 0u : static_cast<uint32_t>((1ULL << 32) / static_cast<unsigned long long>(d[i])) + 1u;
  }
}

namespace {

struct View {  // a strided 5-D (N, D, H, W, C) view of an NDHWC bf16 tensor
  const __nv_bfloat16* base;
  int D, H, W;
  long long sn, sd, sh, sw;
};

struct AxisTap {  // one kernel index along one axis: which parity view, box offset
  int parity;
  int off;
};

inline int out_extent(int in, int k, int stride, int pad, int dil) {
  return (in + 2 * pad - dil * (k - 1) - 1) / stride + 1;
}

// forward-conv axis taps: input coordinate = o*stride + kk*dil - pad  -> (parity view, o + off)
std::vector<AxisTap> fwd_axis_taps(int k, int stride, int pad, int dil) {
  std::vector<AxisTap> v;
  for (int kk = 0; kk < k; kk++) {
    const int off = kk * dil - pad;
    const int par = ((off % stride) + stride) % stride;
    v.push_back({par, (off - par) / stride});
  }
  return v;
}

// number of (tile, tap) pairs along one axis whose shifted box intersects [0, ext_of_view)
long long axis_score(int b, int out_ext, const std::vector<AxisTap>& taps, const std::vector<int>& view_ext) {
  const int tiles = (out_ext + b - 1) / b;
  long long s = 0;
  for (int t = 0; t < tiles; t++)
    for (const AxisTap& a : taps) {
      const int lo = t * b + a.off;
      if (lo + b > 0 && lo < view_ext[a.parity]) s++;
    }
  return s;
}

struct Box {
  int bd, bh, bw;
};

// Choose the tile box. max_rows = 128 (fprop/dgrad, product <= 128) or exact 64 (wgrad K-block, product == 64).
Box plan_box(int Do, int Ho, int Wo, int max_rows, bool exact, const std::vector<AxisTap>& td,
             const std::vector<AxisTap>& th, const std::vector<AxisTap>& tw, const std::vector<int>& ext_d,
             const std::vector<int>& ext_h, const std::vector<int>& ext_w) {
  Box best{1, 1, 1};
  double best_score = -1;
  auto consider = [&](int bd, int bh, int bw) {
    if (bd > 256 || bh > 256 || bw > 256) return;
    const double s = double(axis_score(bd, Do, td, ext_d)) * double(axis_score(bh, Ho, th, ext_h)) *
                     double(axis_score(bw, Wo, tw, ext_w));
    // fewer MMA blocks first; then wider rows (longer contiguous TMA runs)
    if (best_score < 0 || s < best_score ||
        (s == best_score && (bw > best.bw || (bw == best.bw && bh > best.bh)))) {
      best_score = s;
      best = {bd, bh, bw};
    }
  };
  if (exact) {
    for (int bw = 1; bw <= max_rows; bw *= 2)
      for (int bh = 1; bw * bh <= max_rows; bh *= 2) consider(max_rows / (bw * bh), bh, bw);
  } else {
    for (int bw = 1; bw <= std::min(Wo, max_rows); bw++)
      for (int bh = 1; bh <= std::min(Ho, max_rows / bw); bh++) {
        const int bd = std::min(Do, max_rows / (bw * bh));
        if (bd >= 1) consider(bd, bh, bw);
      }
  }
  return best;
}

int encode_view(CUtensorMap* m, const View& v, int C, const Box& b, int N) {
  const uint64_t dims[5] = {uint64_t(C), uint64_t(v.W), uint64_t(v.H), uint64_t(v.D), uint64_t(N)};
  const uint64_t strides[5] = {1, uint64_t(v.sw), uint64_t(v.sh), uint64_t(v.sd), uint64_t(v.sn)};
  const uint32_t box[5] = {64, uint32_t(b.bw), uint32_t(b.bh), uint32_t(b.bd), 1};
  return make_tmap_bf16(m, v.base, 5, dims, strides, box, true);
}

// Parity-class views of a dense NDHWC tensor for a given stride (stride^3 views).
std::vector<View> parity_views(const __nv_bfloat16* base, int D, int H, int W, int C, int stride) {
  std::vector<View> v;
  const long long sw = C, sh = (long long)W * C, sd = (long long)H * W * C, sn = (long long)D * H * W * C;
  for (int pd = 0; pd < stride; pd++)
    for (int ph = 0; ph < stride; ph++)
      for (int pw = 0; pw < stride; pw++) {
        View x;
        x.base = base + pd * sd + ph * sh + pw * sw;
        x.D = (D - pd + stride - 1) / stride;
        x.H = (H - ph + stride - 1) / stride;
        x.W = (W - pw + stride - 1) / stride;
        x.sn = sn;
        x.sd = sd * stride;
        x.sh = sh * stride;
        x.sw = sw * stride;
        v.push_back(x);
      }
  return v;
}

bool tc_supported(const adni_conv3d_geom& g) {
  return g.Cin % 64 == 0 && g.Cout % 64 == 0 && (g.stride == 1 || g.stride == 2) && g.k * g.k * g.k <= kMaxTaps &&
         g.pad <= 120 && g.dil * (g.k - 1) <= 120;
}

int pick_block_n(int n_total) { return n_total % 256 == 0 ? 256 : (n_total % 128 == 0 ? 128 : 64); }

// ---------------------------------------------------------------------------------------------
// Halo-resident engine (conv_halo.cu): 3x3x3 / stride 1 / dilation 1 / pad 1, Cin == Cout in {64, 128}.
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

bool halo_supported(const adni_conv3d_geom& g) {
  return g.k == 3 && g.stride == 1 && g.dil == 1 && g.pad == 1 && g.Cin == g.Cout && (g.Cin == 64 || g.Cin == 128) &&
         env_int("ADNI_HALO", 1) != 0;
}

// `a` is the tensor the taps slide over (x for fprop, dy for dgrad), `w` the matching K-major weight copy
// (OTI / ITO), `mirror` selects dgrad's flipped tap order.  Returns ADNI_ENOTSUP when the tile set does not fit.
int tc_halo(const adni_conv3d_geom& g, const __nv_bfloat16* a, const __nv_bfloat16* w, bool mirror, const float* bias,
            const __nv_bfloat16* addend, __nv_bfloat16* out, double* ssum, double* ssq, cudaStream_t stream,
            const BnReduceEpilogue* red = nullptr) {
  const int C = g.Cin;
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.mirror = mirror ? 1 : 0;
  {
    const uint64_t dims[5] = {uint64_t(C), uint64_t(g.W), uint64_t(g.H), uint64_t(g.D), uint64_t(g.N)};
    const uint64_t strides[5] = {1, uint64_t(C), uint64_t(g.W) * C, uint64_t(g.H) * g.W * C,
                                 uint64_t(g.D) * g.H * g.W * C};
    const uint32_t box[5] = {64, uint32_t(halo_pitch()), uint32_t(halo_rows()), 1, 1};
    int rc = make_tmap_bf16(&p.a_map, a, 5, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {uint64_t(27) * C, uint64_t(C)};
    const uint64_t strides[2] = {1, uint64_t(27) * C};
    const uint32_t box[2] = {64, uint32_t(C)};
    int rc = make_tmap_bf16(&p.b_map, w, 2, dims, strides, box, true);
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
  p.out_sw = C;
  p.out_sh = (long long)g.W * C;
  p.out_sd = (long long)g.H * g.W * C;
  p.out_sn = (long long)g.D * g.H * g.W * C;
  p.out = out;
  p.addend = addend;
  p.bias = bias;
  p.stat_sum = ssum;
  p.stat_sq = ssq;
  if (red != nullptr) {  // dx has the layout of the tensor the BatchNorm normalised: same offsets
    p.red_y = red->y;
    p.red_mask = red->mask;
    p.red_scale = red->scale;
    p.red_shift = red->shift;
    p.stat_sum = red->sum_g;
    p.stat_sq = red->sum_gy;
  }
  return launch_igemm_halo(p, C, stream);
}

// 1x1x1 stride-1 convs are plain GEMMs over the flattened positions: out[M, n_out] = a[M, k_in] * w[n_out, k_in]^T
// (+ addend).  The flat path of the tap-per-box kernel: 128 consecutive positions per tile, TMA-store epilogue.
bool flat_1x1(const adni_conv3d_geom& g, const float* bias) {
  static const bool on = env_int("ADNI_FLAT_1X1", 1) != 0;
  const long long m = static_cast<long long>(g.N) * g.D * g.H * g.W;
  return on && g.k == 1 && g.stride == 1 && g.pad == 0 && bias == nullptr && m < (1LL << 31);
}

int tc_flat_gemm(long long m_rows, int k_in, int n_out, const __nv_bfloat16* a, const __nv_bfloat16* w, const __nv_bfloat16* addend,
                 __nv_bfloat16* out, double* ssum, double* ssq, cudaStream_t stream) {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  const int M = static_cast<int>(m_rows);
  View v;
  v.base = a;
  v.D = 1;
  v.H = 1;
  v.W = M;
  v.sw = k_in;
  v.sh = v.sd = v.sn = static_cast<long long>(M) * k_in;
  const Box b{1, 1, 128};
  int rc = encode_view(&p.a_maps[0], v, k_in, b, 1);
  if (rc) return rc;
  p.a_ext[0][0] = 1;
  p.a_ext[0][1] = 1;
  p.a_ext[0][2] = M;
  const int block_n = pick_block_n(n_out);
  {
    const uint64_t dims[2] = {uint64_t(k_in), uint64_t(n_out)};
    const uint64_t strides[2] = {1, uint64_t(k_in)};
    const uint32_t box[2] = {64, uint32_t(block_n)};
    rc = make_tmap_bf16(&p.b_map, w, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {uint64_t(n_out), uint64_t(M)};
    const uint64_t strides[2] = {1, uint64_t(n_out)};
    const uint32_t box[2] = {64, 32};
    rc = make_tmap_bf16(&p.out_map, out, 2, dims, strides, box, true);
    if (rc) return rc;
    if (addend != nullptr) {
      rc = make_tmap_bf16(&p.add_map, addend, 2, dims, strides, box, true);
      if (rc) return rc;
    }
  }
  p.taps[0].map = 0;
  p.taps[0].dd = p.taps[0].dh = p.taps[0].dw = 0;
  p.taps[0].kofs = 0;
  p.ntaps = 1;
  p.kc_blocks = k_in / 64;
  p.N = 1;
  p.Do = p.Ho = 1;
  p.Wo = M;
  p.bd = p.bh = 1;
  p.bw = 128;
  p.tiles_d = p.tiles_h = 1;
  p.tiles_w = (M + 127) / 128;
  p.n_tiles = n_out / block_n;
  p.n_total = n_out;
  p.out_sw = n_out;
  p.out_sh = p.out_sd = p.out_sn = static_cast<long long>(M) * n_out;
  p.out = out;
  p.addend = addend;
  p.stat_sum = ssum;
  p.stat_sq = ssq;
  p.flat = 1;
  {
    const char* dbg = getenv("ADNI_DEBUG_MODE");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  finish_igemm_params(p);
  return launch_igemm(p, block_n, stream);
}

int tc_fprop(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* w_oti, const float* bias,
             __nv_bfloat16* y, double* ssum, double* ssq, cudaStream_t stream) {
  if (flat_1x1(g, bias))
    return tc_flat_gemm(static_cast<long long>(g.N) * g.D * g.H * g.W, g.Cin, g.Cout, x, w_oti, nullptr, y, ssum, ssq, stream);
  if (halo_supported(g)) {
    const int rc = tc_halo(g, x, w_oti, false, bias, nullptr, y, ssum, ssq, stream);
    if (rc != ADNI_ENOTSUP) return rc;
  }
  const int Do = out_extent(g.D, g.k, g.stride, g.pad, g.dil);
  const int Ho = out_extent(g.H, g.k, g.stride, g.pad, g.dil);
  const int Wo = out_extent(g.W, g.k, g.stride, g.pad, g.dil);
  const std::vector<View> views = parity_views(x, g.D, g.H, g.W, g.Cin, g.stride);
  const std::vector<AxisTap> at = fwd_axis_taps(g.k, g.stride, g.pad, g.dil);
  std::vector<int> ext_d, ext_h, ext_w;
  for (int p = 0; p < g.stride; p++) {
    ext_d.push_back((g.D - p + g.stride - 1) / g.stride);
    ext_h.push_back((g.H - p + g.stride - 1) / g.stride);
    ext_w.push_back((g.W - p + g.stride - 1) / g.stride);
  }
  const Box b = plan_box(Do, Ho, Wo, 128, false, at, at, at, ext_d, ext_h, ext_w);

  IgemmParams p;
  memset(&p, 0, sizeof(p));
  for (size_t i = 0; i < views.size(); i++) {
    int rc = encode_view(&p.a_maps[i], views[i], g.Cin, b, g.N);
    if (rc) return rc;
    p.a_ext[i][0] = views[i].D;
    p.a_ext[i][1] = views[i].H;
    p.a_ext[i][2] = views[i].W;
  }
  const int taps = g.k * g.k * g.k;
  const int block_n = pick_block_n(g.Cout);
  {
    const uint64_t dims[2] = {uint64_t(taps) * g.Cin, uint64_t(g.Cout)};
    const uint64_t strides[2] = {1, uint64_t(taps) * g.Cin};
    const uint32_t box[2] = {64, uint32_t(block_n)};
    int rc = make_tmap_bf16(&p.b_map, w_oti, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  int t = 0;
  for (int kd = 0; kd < g.k; kd++)
    for (int kh = 0; kh < g.k; kh++)
      for (int kw = 0; kw < g.k; kw++, t++) {
        ConvTap& tp = p.taps[t];
        tp.map = int8_t((at[kd].parity * g.stride + at[kh].parity) * g.stride + at[kw].parity);
        tp.dd = int8_t(at[kd].off);
        tp.dh = int8_t(at[kh].off);
        tp.dw = int8_t(at[kw].off);
        tp.kofs = t * g.Cin;
      }
  p.ntaps = taps;
  p.kc_blocks = g.Cin / 64;
  p.N = g.N;
  p.Do = Do;
  p.Ho = Ho;
  p.Wo = Wo;
  p.bd = b.bd;
  p.bh = b.bh;
  p.bw = b.bw;
  p.tiles_d = (Do + b.bd - 1) / b.bd;
  p.tiles_h = (Ho + b.bh - 1) / b.bh;
  p.tiles_w = (Wo + b.bw - 1) / b.bw;
  p.n_tiles = g.Cout / block_n;
  p.n_total = g.Cout;
  p.out_sw = g.Cout;
  p.out_sh = (long long)Wo * g.Cout;
  p.out_sd = (long long)Ho * Wo * g.Cout;
  p.out_sn = (long long)Do * Ho * Wo * g.Cout;
  p.out = y;
  p.addend = nullptr;
  p.bias = bias;
  p.stat_sum = ssum;
  p.stat_sq = ssq;
  {
    const char* dbg = getenv("ADNI_DEBUG_MODE");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  finish_igemm_params(p);
  return launch_igemm(p, block_n, stream);
}

// dx = sum_k dy[(i + pad - k*dil)/stride] * w[k]^T, one launch per parity class of dx when stride > 1.
int tc_dgrad(const adni_conv3d_geom& g, const __nv_bfloat16* dy, const __nv_bfloat16* w_ito,
             const __nv_bfloat16* addend, __nv_bfloat16* dx, cudaStream_t stream, const BnReduceEpilogue* red = nullptr) {
  if (red == nullptr && flat_1x1(g, nullptr))   // dx[M, Cin] = dy[M, Cout] * w_ito[Cin, Cout]^T (+ addend)
    return tc_flat_gemm(static_cast<long long>(g.N) * g.D * g.H * g.W, g.Cout, g.Cin, dy, w_ito, addend, dx, nullptr, nullptr, stream);
  if (halo_supported(g)) {
    const int rc = tc_halo(g, dy, w_ito, true, nullptr, addend, dx, nullptr, nullptr, stream, red);
    if (rc != ADNI_ENOTSUP) return rc;
  }
  const int Do = out_extent(g.D, g.k, g.stride, g.pad, g.dil);
  const int Ho = out_extent(g.H, g.k, g.stride, g.pad, g.dil);
  const int Wo = out_extent(g.W, g.k, g.stride, g.pad, g.dil);
  const int taps = g.k * g.k * g.k;
  const int block_n = pick_block_n(g.Cin);
  View dyv;
  dyv.base = dy;
  dyv.D = Do;
  dyv.H = Ho;
  dyv.W = Wo;
  dyv.sw = g.Cout;
  dyv.sh = (long long)Wo * g.Cout;
  dyv.sd = (long long)Ho * Wo * g.Cout;
  dyv.sn = (long long)Do * Ho * Wo * g.Cout;
  const std::vector<View> outs = parity_views(dx, g.D, g.H, g.W, g.Cin, g.stride);
  const std::vector<View> adds =
      addend ? parity_views(addend, g.D, g.H, g.W, g.Cin, g.stride) : std::vector<View>();
  const std::vector<int> ext_d{Do}, ext_h{Ho}, ext_w{Wo};

  // stride > 1: the parity classes of dx are independent problems with the same tile shape -> ONE launch
  static const bool multi_on = env_int("ADNI_DGRAD_MULTI", 1) != 0;
  const bool multi = multi_on && g.stride > 1 && g.stride * g.stride * g.stride <= kMaxMultiProblems;
  std::vector<IgemmMulti> pm_store(multi ? 1 : 0);
  IgemmMulti* pm = multi ? pm_store.data() : nullptr;
  if (pm) pm->ncls = 0;
  int cls = 0;
  for (int pd = 0; pd < g.stride; pd++)
    for (int ph = 0; ph < g.stride; ph++)
      for (int pw = 0; pw < g.stride; pw++, cls++) {
        const View& ov = outs[cls];
        if (ov.D <= 0 || ov.H <= 0 || ov.W <= 0) continue;
        // per-axis taps contributing to this parity class
        auto axis = [&](int par) {
          std::vector<std::pair<int, AxisTap>> v;  // (kernel index, tap)
          for (int kk = 0; kk < g.k; kk++) {
            const int num = par + g.pad - kk * g.dil;
            if (((num % g.stride) + g.stride) % g.stride != 0) continue;
            v.push_back({kk, AxisTap{0, num / g.stride}});
          }
          return v;
        };
        const auto ad = axis(pd), ah = axis(ph), aw = axis(pw);
        std::vector<AxisTap> td, th, tw;
        for (auto& a : ad) td.push_back(a.second);
        for (auto& a : ah) th.push_back(a.second);
        for (auto& a : aw) tw.push_back(a.second);
        // (an empty axis makes every score 0: any box works)
        const Box b = plan_box(ov.D, ov.H, ov.W, 128, false, td, th, tw, ext_d, ext_h, ext_w);
        IgemmParams p;
        memset(&p, 0, sizeof(p));
        int rc = encode_view(&p.a_maps[0], dyv, g.Cout, b, g.N);
        if (rc) return rc;
        p.a_ext[0][0] = Do;
        p.a_ext[0][1] = Ho;
        p.a_ext[0][2] = Wo;
        {
          const uint64_t dims[2] = {uint64_t(taps) * g.Cout, uint64_t(g.Cin)};
          const uint64_t strides[2] = {1, uint64_t(taps) * g.Cout};
          const uint32_t box[2] = {64, uint32_t(block_n)};
          rc = make_tmap_bf16(&p.b_map, w_ito, 2, dims, strides, box, true);
          if (rc) return rc;
        }
        int nt = 0;
        for (auto& a : ad)
          for (auto& bq : ah)
            for (auto& cq : aw) {
              ConvTap& tp = p.taps[nt++];
              tp.map = 0;
              tp.dd = int8_t(a.second.off);
              tp.dh = int8_t(bq.second.off);
              tp.dw = int8_t(cq.second.off);
              tp.kofs = ((a.first * g.k + bq.first) * g.k + cq.first) * g.Cout;
            }
        p.ntaps = nt;
        p.kc_blocks = g.Cout / 64;
        p.N = g.N;
        p.Do = ov.D;
        p.Ho = ov.H;
        p.Wo = ov.W;
        p.bd = b.bd;
        p.bh = b.bh;
        p.bw = b.bw;
        p.tiles_d = (ov.D + b.bd - 1) / b.bd;
        p.tiles_h = (ov.H + b.bh - 1) / b.bh;
        p.tiles_w = (ov.W + b.bw - 1) / b.bw;
        p.n_tiles = g.Cin / block_n;
        p.n_total = g.Cin;
        p.out_sn = ov.sn;
        p.out_sd = ov.sd;
        p.out_sh = ov.sh;
        p.out_sw = ov.sw;
        p.out = const_cast<__nv_bfloat16*>(ov.base);
        p.addend = addend ? adds[cls].base : nullptr;
        if (red != nullptr) {  // y / mask share dx's layout: the parity class is the same element offset into them
          const long long cls_off = ov.base - dx;
          p.red_y = red->y + cls_off;
          p.red_mask = red->mask ? red->mask + cls_off : nullptr;
          p.red_scale = red->scale;
          p.red_shift = red->shift;
          p.stat_sum = red->sum_g;
          p.stat_sq = red->sum_gy;
        }
        finish_igemm_params(p);
        if (pm) {
          pm->cls[pm->ncls++] = p;
          continue;
        }
        rc = launch_igemm(p, block_n, stream);
        if (rc) return rc;
      }
  if (pm && pm->ncls > 0) return launch_igemm_multi(*pm, block_n, stream);
  return ADNI_OK;
}

// Geometry half of the wgrad plan (everything except the tensor maps and pointers): shared by tc_wgrad and the
// planner query.  Returns the M-tile configuration; fills the box and every integer field of `p`.
int plan_wgrad_geometry(const adni_conv3d_geom& g, WgradParams& p, Box& b) {
  const int Do = out_extent(g.D, g.k, g.stride, g.pad, g.dil);
  const int Ho = out_extent(g.H, g.k, g.stride, g.pad, g.dil);
  const int Wo = out_extent(g.W, g.k, g.stride, g.pad, g.dil);
  const int taps = g.k * g.k * g.k;
  const std::vector<AxisTap> at = fwd_axis_taps(g.k, g.stride, g.pad, g.dil);
  std::vector<int> ext_d, ext_h, ext_w;
  for (int q = 0; q < g.stride; q++) {
    ext_d.push_back((g.D - q + g.stride - 1) / g.stride);
    ext_h.push_back((g.H - q + g.stride - 1) / g.stride);
    ext_w.push_back((g.W - q + g.stride - 1) / g.stride);
  }
  // one CTA owns all 512 TMEM columns: mt_cfg accumulators of 128 Cout rows x (512/mt_cfg) K_total columns
  int mt_cfg = g.Cout >= 256 ? 2 : 1;  // measured best on B200 (tools/wgrad_probe.py)
  if (const char* e = getenv("ADNI_WGRAD_MT")) {  // tuning / diagnostics override
    const int v = atoi(e);
    if (v == 1 || v == 2 || v == 4) mt_cfg = v;
  }
  const int box_rows = wgrad2_box_rows(mt_cfg);
  b = plan_box(Do, Ho, Wo, box_rows, true, at, at, at, ext_d, ext_h, ext_w);
  int t = 0;
  for (int kd = 0; kd < g.k; kd++)
    for (int kh = 0; kh < g.k; kh++)
      for (int kw = 0; kw < g.k; kw++, t++) {
        ConvTap& tp = p.taps[t];
        tp.map = int8_t((at[kd].parity * g.stride + at[kh].parity) * g.stride + at[kw].parity);
        tp.dd = int8_t(at[kd].off);
        tp.dh = int8_t(at[kh].off);
        tp.dw = int8_t(at[kw].off);
        tp.kofs = t * g.Cin;
      }
  int v = 0;
  for (int pd = 0; pd < g.stride; pd++)
    for (int ph = 0; ph < g.stride; ph++)
      for (int pw = 0; pw < g.stride; pw++, v++) {
        p.x_ext[v][0] = ext_d[pd];
        p.x_ext[v][1] = ext_h[ph];
        p.x_ext[v][2] = ext_w[pw];
      }
  p.ntaps = taps;
  p.cin_blocks = g.Cin / 64;
  p.n_groups = taps * p.cin_blocks;
  const int groups = 8 / mt_cfg;
  p.N = g.N;
  p.Do = Do;
  p.Ho = Ho;
  p.Wo = Wo;
  p.bd = b.bd;
  p.bh = b.bh;
  p.bw = b.bw;
  p.tiles_d = (Do + b.bd - 1) / b.bd;
  p.tiles_h = (Ho + b.bh - 1) / b.bh;
  p.tiles_w = (Wo + b.bw - 1) / b.bw;
  p.pos_boxes = g.N * p.tiles_d * p.tiles_h * p.tiles_w;
  p.m_tiles = (g.Cout + 128 * mt_cfg - 1) / (128 * mt_cfg);
  p.n_tiles = (p.n_groups + groups - 1) / groups;
  const int base_items = p.m_tiles * p.n_tiles;
  int splits = (4 * num_sms() + base_items - 1) / base_items;
  splits = std::min(splits, std::max(1, p.pos_boxes * box_rows / 1024));
  splits = std::max(splits, 1);
  // ADNI_WGRAD_DETERMINISTIC=1: one CTA per output tile over ALL positions (no split-K, no stream-K): every element of dW
  // receives exactly one red.add onto the zeroed buffer, the MMA order inside a CTA is fixed -> bit-identical runs, at
  // the price of the schedule's balance (108 tiles on 148 SMs for layer4)
  if (env_int("ADNI_WGRAD_DETERMINISTIC", 0) != 0) splits = 1;
  p.boxes_per_split = (p.pos_boxes + splits - 1) / splits;
  p.splits = (p.pos_boxes + p.boxes_per_split - 1) / p.boxes_per_split;
  p.cout = g.Cout;
  p.k_total = p.n_groups * 64;
  return mt_cfg;
}

// Fraction of the (position box, 64-column group) MMA blocks wgrad2 issues: a box is skipped for an N tile when the
// shifted boxes of ALL the tile's taps lie in the zero padding (conv_wgrad2.cu box_active).
double wgrad_executed_fraction(const WgradParams& p, int mt_cfg) {
  const int groups = 8 / mt_cfg;
  long long issued = 0;
  for (int nt = 0; nt < p.n_tiles; nt++) {
    const int g0 = nt * groups, ng = std::min(groups, p.n_groups - g0);
    for (int td = 0; td < p.tiles_d; td++)
      for (int th = 0; th < p.tiles_h; th++)
        for (int tw = 0; tw < p.tiles_w; tw++) {
          bool any = false;
          for (int gi = 0; gi < ng && !any; gi++) {
            const ConvTap& tp = p.taps[(g0 + gi) / p.cin_blocks];
            const int* ext = p.x_ext[tp.map];
            const int d = td * p.bd + tp.dd, h = th * p.bh + tp.dh, w = tw * p.bw + tp.dw;
            any = d + p.bd > 0 && d < ext[0] && h + p.bh > 0 && h < ext[1] && w + p.bw > 0 && w < ext[2];
          }
          if (any) issued += ng;
        }
  }
  return double(issued) / (double(p.n_groups) * p.tiles_d * p.tiles_h * p.tiles_w);
}

// Stream-K plan for wgrad2 (cached per geometry): equal numbers of ACTIVE (position box, output tile) blocks per CTA.
// The static split-K schedule leaves the slowest CTA 20-35 % above the mean on the dilated layers (boxes whose taps all
// lie in the padding are skipped, and 594-648 items fall unevenly on 148 SMs); partial sums already go through
// red.global.add, so cutting tiles at arbitrary box boundaries needs no fix-up pass.
// The sequence is chunk-major (W2Sched): a plain tile-major sequence spreads the 148 CTAs over ALL positions of dY / X at
// any moment, and once the two tensors exceed the L2 the kernel streams them from HBM once per output tile
// (profiles/r02_launch_summary.md: 4.0 GB of DRAM reads per layer4 launch against 0.6 GB under the static schedule).
// Chunks of whole samples whose dY + X slices fit ADNI_WGRAD_CHUNK_MB restore the reuse, but every extra chunk costs each
// CTA one more un-overlapped accumulator flush (256 KB of red.add; the CTA owns all 512 TMEM columns), and the kernel is
// tensor-bound either way (0.98 of the burst peak in executed FLOPs): measured 53.04 ms per step with 48 MB chunks,
// 52.31 with the default of 128 MB (two chunks for layers 3 / 4 at 32 pairs, traffic unchanged), 52.52 with one chunk
// (profiles/r02_wgrad_chunk_ab.md).  The knob stays for boxes whose HBM is the scarcer resource.
struct W2Plan {
  bool use = false;
  W2Sched sched;
};

const W2Plan& plan_wgrad_stream_k(const WgradParams& p, int mt_cfg) {
  static std::mutex mu;
  static std::map<std::vector<int>, W2Plan> cache;
  const int chunk_mb = std::max(1, env_int("ADNI_WGRAD_CHUNK_MB", 128));
  std::vector<int> key = {p.ntaps, p.cin_blocks, p.N, p.Do, p.Ho, p.Wo, p.bd, p.bh, p.bw, p.m_tiles, p.n_tiles, mt_cfg, num_sms(),
                          p.cout, chunk_mb, env_int("ADNI_WGRAD_DETERMINISTIC", 0), env_int("ADNI_STREAM_K_WGRAD", 1)};
  for (int t = 0; t < p.ntaps; t++)
    key.push_back((int(p.taps[t].map) << 24) ^ ((p.taps[t].dd & 0xff) << 16) ^ ((p.taps[t].dh & 0xff) << 8) ^ (p.taps[t].dw & 0xff));
  for (int m = 0; m < kMaxMaps; m++)
    for (int a = 0; a < 3; a++) key.push_back(p.x_ext[m][a]);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  W2Plan plan;
  memset(&plan.sched, 0, sizeof(plan.sched));
  const int enabled = env_int("ADNI_STREAM_K_WGRAD", 1) != 0 && env_int("ADNI_WGRAD_DETERMINISTIC", 0) == 0;
  const int groups = 8 / mt_cfg;
  const int per_sample = p.tiles_d * p.tiles_h * p.tiles_w;
  const int G = std::min(num_sms(), kSkMaxCtas);
  if (enabled && per_sample > 0 && p.N > 0) {
    // cum[nt][r] = active boxes among the first r boxes of one sample for N tile nt (the pattern repeats per sample)
    std::vector<std::vector<int>> cum(static_cast<size_t>(p.n_tiles), std::vector<int>(static_cast<size_t>(per_sample) + 1, 0));
    for (int nt = 0; nt < p.n_tiles; nt++) {
      const int g0 = nt * groups, ng = std::min(groups, p.n_groups - g0);
      int r = 0;
      for (int td = 0; td < p.tiles_d; td++)
        for (int th = 0; th < p.tiles_h; th++)
          for (int tw = 0; tw < p.tiles_w; tw++, r++) {
            bool any = false;
            for (int gi = 0; gi < ng && !any; gi++) {
              const ConvTap& tp = p.taps[(g0 + gi) / p.cin_blocks];
              const int* ext = p.x_ext[tp.map];
              const int d = td * p.bd + tp.dd, h = th * p.bh + tp.dh, w = tw * p.bw + tp.dw;
              any = d + p.bd > 0 && d < ext[0] && h + p.bh > 0 && h < ext[1] && w + p.bw > 0 && w < ext[2];
            }
            cum[nt][r + 1] = cum[nt][r] + (any ? 1 : 0);
          }
    }
    // position chunks: whole samples whose dY + X slices fit the L2 budget
    long long x_pos = 0;
    for (int m = 0; m < kMaxMaps; m++) x_pos += static_cast<long long>(p.x_ext[m][0]) * p.x_ext[m][1] * p.x_ext[m][2];
    const long long bytes_per_sample =
        2LL * (static_cast<long long>(p.Do) * p.Ho * p.Wo * p.cout + x_pos * p.cin_blocks * 64);
    const long long fit = (static_cast<long long>(chunk_mb) << 20) / std::max(1LL, bytes_per_sample);
    const int chunk_samples = static_cast<int>(std::min<long long>(p.N, std::max(1LL, fit)));
    const int n_chunks = (p.N + chunk_samples - 1) / chunk_samples;
    const int tiles = p.m_tiles * p.n_tiles;
    const long long vtiles_ll = static_cast<long long>(tiles) * n_chunks;
    if (vtiles_ll < (1LL << 24)) {
      const int vtiles = static_cast<int>(vtiles_ll);
      auto samples_of = [&](int chunk) { return std::min(chunk_samples, p.N - chunk * chunk_samples); };
      std::vector<long long> prefix(static_cast<size_t>(vtiles) + 1, 0);   // active blocks before virtual tile v
      for (int v = 0; v < vtiles; v++)
        prefix[v + 1] = prefix[v] + static_cast<long long>(cum[(v % tiles) % p.n_tiles][per_sample]) * samples_of(v / tiles);
      const long long T = prefix[vtiles];
      if (T >= 8LL * G) {
        // raw box index (over all samples) of the a-th active box (0-based) of virtual tile v
        auto raw_of = [&](int v, long long a) -> int {
          const int nt = (v % tiles) % p.n_tiles;
          const int per = cum[nt][per_sample];
          const long long n = a / per;
          const int rem = int(a % per);
          int r = int(std::upper_bound(cum[nt].begin(), cum[nt].end(), rem) - cum[nt].begin()) - 1;   // cum[r] <= rem < cum[r+1]
          return int((static_cast<long long>(v / tiles) * chunk_samples + n) * per_sample + r);
        };
        plan.use = true;
        plan.sched.ctas = G;
        plan.sched.chunk_boxes = chunk_samples * per_sample;
        int tile = 0;
        for (int c = 0; c < G; c++) {
          const long long start = T * c / G, end = T * (c + 1) / G;   // end > start
          while (prefix[tile + 1] <= start) tile++;
          plan.sched.tile_begin[c] = tile;
          plan.sched.box_begin[c] = raw_of(tile, start - prefix[tile]);
          int last = tile;
          while (prefix[last + 1] < end) last++;
          plan.sched.tile_last[c] = last;
          plan.sched.box_end[c] = raw_of(last, end - 1 - prefix[last]) + 1;
        }
      }
    }
  }
  return cache.emplace(key, plan).first->second;
}

int tc_wgrad(const adni_conv3d_geom& g, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw,
             cudaStream_t stream) {
  WgradParams p;
  memset(&p, 0, sizeof(p));
  Box b;
  const int mt_cfg = plan_wgrad_geometry(g, p, b);
  const std::vector<View> views = parity_views(x, g.D, g.H, g.W, g.Cin, g.stride);
  View dyv;
  dyv.base = dy;
  dyv.D = p.Do;
  dyv.H = p.Ho;
  dyv.W = p.Wo;
  dyv.sw = g.Cout;
  dyv.sh = (long long)p.Wo * g.Cout;
  dyv.sd = (long long)p.Ho * p.Wo * g.Cout;
  dyv.sn = (long long)p.Do * p.Ho * p.Wo * g.Cout;
  int rc = encode_view(&p.dy_map, dyv, g.Cout, b, g.N);
  if (rc) return rc;
  for (size_t i = 0; i < views.size(); i++) {
    rc = encode_view(&p.x_maps[i], views[i], g.Cin, b, g.N);
    if (rc) return rc;
  }
  p.dw = dw;
  const W2Plan& plan = plan_wgrad_stream_k(p, mt_cfg);
  return launch_wgrad2(p, mt_cfg, stream, plan.use ? &plan.sched : nullptr);
}

int check_geom(const adni_conv3d_geom* g) {
  ADNI_REQUIRE(g != nullptr, ADNI_EINVAL, "conv3d: null geometry");
  ADNI_REQUIRE(g->N > 0 && g->D > 0 && g->H > 0 && g->W > 0 && g->Cin > 0 && g->Cout > 0, ADNI_EINVAL,
               "conv3d: non-positive extent");
  ADNI_REQUIRE(g->k > 0 && g->stride > 0 && g->dil > 0 && g->pad >= 0, ADNI_EINVAL, "conv3d: bad k/stride/dil/pad");
  ADNI_REQUIRE(out_extent(g->D, g->k, g->stride, g->pad, g->dil) > 0 &&
                   out_extent(g->H, g->k, g->stride, g->pad, g->dil) > 0 &&
                   out_extent(g->W, g->k, g->stride, g->pad, g->dil) > 0,
               ADNI_EINVAL, "conv3d: empty output");
  return ADNI_OK;
}

bool small_enabled() {
  static const bool on = env_int("ADNI_SMALL_CONV", 1) != 0;
  return on;
}

// AUTO: tcgen05 when the channel counts fill its tiles, else the mma.sync small-channel engine, else CUDA cores.
int resolve_engine(const adni_conv3d_geom& g, int pass, bool has_addend, int engine, int* out) {
  if (engine == ADNI_ENGINE_AUTO) {
    if (tc_supported(g)) engine = ADNI_ENGINE_TCGEN05;
    else if (small_enabled() && !has_addend && small_conv_supported(g, pass)) engine = ADNI_ENGINE_MMA_SYNC;
    else engine = ADNI_ENGINE_DIRECT;
  }
  if (engine == ADNI_ENGINE_TCGEN05 && !tc_supported(g)) {
    set_error("conv3d: tcgen05 engine needs Cin,Cout multiples of 64, stride 1|2, k^3 <= 64 (Cin=%d Cout=%d k=%d s=%d)",
              g.Cin, g.Cout, g.k, g.stride);
    return ADNI_ENOTSUP;
  }
  if (engine == ADNI_ENGINE_MMA_SYNC && (has_addend || !small_conv_supported(g, pass))) {
    set_error("conv3d: the small-channel engine needs stride 1, dilation 1, Cin in {1,8,16,32,64}, Cout in {8,16,32,64}, "
              "no addend (Cin=%d Cout=%d k=%d s=%d d=%d)", g.Cin, g.Cout, g.k, g.stride, g.dil);
    return ADNI_ENOTSUP;
  }
  if (engine != ADNI_ENGINE_TCGEN05 && engine != ADNI_ENGINE_DIRECT && engine != ADNI_ENGINE_MMA_SYNC) {
    set_error("conv3d: unknown engine %d", engine);
    return ADNI_EINVAL;
  }
  *out = engine;
  return ADNI_OK;
}

}  // namespace
}  // namespace adni

using namespace adni;

extern "C" {

int adni_conv3d_out_extent(int in, int k, int stride, int pad, int dil) { return out_extent(in, k, stride, pad, dil); }

int adni_conv3d_plan_info(const adni_conv3d_geom* g, int pass, int* engine_kind, double* executed_fraction) {
  int rc = check_geom(g);
  if (rc) return rc;
  ADNI_REQUIRE(engine_kind && executed_fraction && pass >= 0 && pass <= 2, ADNI_EINVAL, "conv3d_plan_info: bad arguments");
  *executed_fraction = 1.0;
  if (!tc_supported(*g)) {
    *engine_kind = (small_enabled() && small_conv_supported(*g, pass)) ? 3 : 0;
    return ADNI_OK;
  }
  *engine_kind = 1;
  if (pass == 2) {
    if (wgrad_halo_supported(*g)) {  // layer1 halo-plane wgrad: the kd = 0 / 2 taps of the first / last plane are skipped
      *executed_fraction = (3.0 * g->D - 2.0) / (3.0 * g->D);
      return ADNI_OK;
    }
    WgradParams wp;
    memset(&wp, 0, sizeof(wp));
    Box wb;
    const int mt_cfg = plan_wgrad_geometry(*g, wp, wb);
    *executed_fraction = wgrad_executed_fraction(wp, mt_cfg);
    return ADNI_OK;
  }
  if (halo_supported(*g)) {
    *engine_kind = 2;
    *executed_fraction = (3.0 * g->D - 2.0) / (3.0 * g->D);  // the kd = 0 / 2 taps of the first / last plane
    return ADNI_OK;
  }
  if (pass == 1 && g->stride != 1) return ADNI_OK;  // strided dgrad: per parity class, reported as issued
  const int Do = out_extent(g->D, g->k, g->stride, g->pad, g->dil), Ho = out_extent(g->H, g->k, g->stride, g->pad, g->dil),
            Wo = out_extent(g->W, g->k, g->stride, g->pad, g->dil);
  std::vector<AxisTap> at;
  std::vector<int> ext_d, ext_h, ext_w;
  int od = Do, oh = Ho, ow = Wo;
  if (pass == 0) {
    at = fwd_axis_taps(g->k, g->stride, g->pad, g->dil);
    for (int p = 0; p < g->stride; p++) {
      ext_d.push_back((g->D - p + g->stride - 1) / g->stride);
      ext_h.push_back((g->H - p + g->stride - 1) / g->stride);
      ext_w.push_back((g->W - p + g->stride - 1) / g->stride);
    }
  } else {  // stride-1 dgrad: the taps slide over dy, the output is dx
    for (int kk = 0; kk < g->k; kk++) at.push_back({0, g->pad - kk * g->dil});
    ext_d = {Do};
    ext_h = {Ho};
    ext_w = {Wo};
    od = g->D;
    oh = g->H;
    ow = g->W;
  }
  const Box b = plan_box(od, oh, ow, 128, false, at, at, at, ext_d, ext_h, ext_w);
  const double done = double(axis_score(b.bd, od, at, ext_d)) * double(axis_score(b.bh, oh, at, ext_h)) *
                      double(axis_score(b.bw, ow, at, ext_w));
  const double all = double((od + b.bd - 1) / b.bd) * double((oh + b.bh - 1) / b.bh) * double((ow + b.bw - 1) / b.bw) *
                     double(g->k) * g->k * g->k;
  *executed_fraction = done / all;
  return ADNI_OK;
}

int adni_conv3d_wgrad_schedule(const adni_conv3d_geom* g, int* header, int* table, int table_capacity) {
  int rc = check_geom(g);
  if (rc) return rc;
  ADNI_REQUIRE(header && table && table_capacity >= 0, ADNI_EINVAL, "conv3d_wgrad_schedule: bad arguments");
  ADNI_REQUIRE(tc_supported(*g), ADNI_ENOTSUP, "conv3d_wgrad_schedule: not a tcgen05 wgrad geometry (Cin=%d Cout=%d)", g->Cin, g->Cout);
  WgradParams wp;
  memset(&wp, 0, sizeof(wp));
  Box wb;
  const int mt_cfg = plan_wgrad_geometry(*g, wp, wb);
  const W2Plan& plan = plan_wgrad_stream_k(wp, mt_cfg);
  const int h[16] = {plan.use ? 1 : 0, plan.sched.ctas, plan.sched.chunk_boxes, wp.pos_boxes, wp.m_tiles, wp.n_tiles, 8 / mt_cfg,
                     wp.n_groups, wp.cin_blocks, wp.bd, wp.bh, wp.bw, wp.tiles_d, wp.tiles_h, wp.tiles_w, mt_cfg};
  memcpy(header, h, sizeof(h));
  if (plan.use) {
    ADNI_REQUIRE(table_capacity >= 4 * plan.sched.ctas, ADNI_EINVAL, "conv3d_wgrad_schedule: table too small");
    for (int c = 0; c < plan.sched.ctas; c++) {
      table[4 * c + 0] = plan.sched.tile_begin[c];
      table[4 * c + 1] = plan.sched.box_begin[c];
      table[4 * c + 2] = plan.sched.tile_last[c];
      table[4 * c + 3] = plan.sched.box_end[c];
    }
  }
  return ADNI_OK;
}

int adni_conv3d_fprop(const adni_conv3d_geom* g, const adni_bf16* x, const adni_bf16* w_oti, const float* bias,
                      adni_bf16* y, double* stat_sum, double* stat_sqsum, int engine, void* stream) {
  int rc = check_geom(g);
  if (rc) return rc;
  ADNI_REQUIRE(x && w_oti && y, ADNI_EINVAL, "conv3d_fprop: null pointer");
  ADNI_REQUIRE((stat_sum == nullptr) == (stat_sqsum == nullptr), ADNI_EINVAL, "conv3d_fprop: stats need both buffers");
  int eng;
  rc = resolve_engine(*g, 0, false, engine, &eng);
  if (rc) return rc;
  auto xs = reinterpret_cast<const __nv_bfloat16*>(x);
  auto ws = reinterpret_cast<const __nv_bfloat16*>(w_oti);
  auto ys = reinterpret_cast<__nv_bfloat16*>(y);
  if (eng == ADNI_ENGINE_TCGEN05)
    return tc_fprop(*g, xs, ws, bias, ys, stat_sum, stat_sqsum, static_cast<cudaStream_t>(stream));
  if (eng == ADNI_ENGINE_MMA_SYNC) {
    rc = small_conv_fprop(*g, xs, ws, bias, ys, stat_sum, stat_sqsum, static_cast<cudaStream_t>(stream));
    if (rc != ADNI_ENOTSUP || engine == ADNI_ENGINE_MMA_SYNC) return rc;   // AUTO: tile too large for shared memory
  }
  return direct_conv_fprop(*g, xs, ws, bias, ys, stat_sum, stat_sqsum, static_cast<cudaStream_t>(stream));
}

int adni_conv3d_dgrad(const adni_conv3d_geom* g, const adni_bf16* dy, const adni_bf16* w_ito, const adni_bf16* addend,
                      adni_bf16* dx, int engine, void* stream) {
  int rc = check_geom(g);
  if (rc) return rc;
  ADNI_REQUIRE(dy && w_ito && dx, ADNI_EINVAL, "conv3d_dgrad: null pointer");
  int eng;
  rc = resolve_engine(*g, 1, addend != nullptr, engine, &eng);
  if (rc) return rc;
  auto dys = reinterpret_cast<const __nv_bfloat16*>(dy);
  auto ws = reinterpret_cast<const __nv_bfloat16*>(w_ito);
  auto as = reinterpret_cast<const __nv_bfloat16*>(addend);
  auto dxs = reinterpret_cast<__nv_bfloat16*>(dx);
  if (eng == ADNI_ENGINE_TCGEN05)
    return tc_dgrad(*g, dys, ws, as, dxs, static_cast<cudaStream_t>(stream));
  if (eng == ADNI_ENGINE_MMA_SYNC) {
    rc = small_conv_dgrad(*g, dys, ws, dxs, static_cast<cudaStream_t>(stream));
    if (rc != ADNI_ENOTSUP || engine == ADNI_ENGINE_MMA_SYNC) return rc;
  }
  return direct_conv_dgrad(*g, dys, ws, as, dxs, static_cast<cudaStream_t>(stream));
}

int adni_conv3d_dgrad_bnred_profitable(const adni_conv3d_geom* g) {
  // Measured on B200 (profiles/r02_bnred_ab.md): the extra rows the epilogue reads are hidden only under a long main
  // loop - the tap-per-box engine with >= 54 K blocks per tile (3x3x3, stride 1, >= 128 channels: layers 3 / 4 and
  // the Bottleneck conv2s).  1x1x1 convs are epilogue-bound already (2.5-3x slower with the sums), and the halo
  // engine's four epilogue warps have no slack either (layer1 dgrad 0.36 -> 0.89 ms).
  if (g == nullptr || check_geom(g) != ADNI_OK || !tc_supported(*g) || halo_supported(*g)) return 0;
  return (g->k == 3 && g->stride == 1 && g->Cout >= 128) ? 1 : 0;
}

int adni_conv3d_dgrad_bnred(const adni_conv3d_geom* g, const adni_bf16* dy, const adni_bf16* w_ito, const adni_bf16* addend,
                            adni_bf16* dx, const adni_bf16* bn_y, const adni_bf16* bn_relu_out, const float* bn_scale,
                            const float* bn_shift, double* sum_g, double* sum_gy, void* stream) {
  int rc = check_geom(g);
  if (rc) return rc;
  ADNI_REQUIRE(dy && w_ito && dx && bn_y && sum_g && sum_gy, ADNI_EINVAL, "conv3d_dgrad_bnred: null pointer");
  ADNI_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), ADNI_EINVAL, "conv3d_dgrad_bnred: scale and shift go together");
  ADNI_REQUIRE(!(bn_relu_out && bn_scale), ADNI_EINVAL, "conv3d_dgrad_bnred: one ReLU mask source at most");
  ADNI_REQUIRE(tc_supported(*g), ADNI_ENOTSUP, "conv3d_dgrad_bnred: the fused reduction lives in the tcgen05 epilogues (Cin=%d Cout=%d)",
               g->Cin, g->Cout);
  BnReduceEpilogue red;
  red.y = reinterpret_cast<const __nv_bfloat16*>(bn_y);
  red.mask = reinterpret_cast<const __nv_bfloat16*>(bn_relu_out);
  red.scale = bn_scale;
  red.shift = bn_shift;
  red.sum_g = sum_g;
  red.sum_gy = sum_gy;
  return tc_dgrad(*g, reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(w_ito),
                  reinterpret_cast<const __nv_bfloat16*>(addend), reinterpret_cast<__nv_bfloat16*>(dx),
                  static_cast<cudaStream_t>(stream), &red);
}

long long adni_conv3d_wgrad_scratch_floats(const adni_conv3d_geom* g) {
  if (g == nullptr || check_geom(g) != ADNI_OK) return 0;
  return (tc_supported(*g) && wgrad_halo_supported(*g)) ? 27LL * 64 * 64 : 0;
}

int adni_conv3d_wgrad(const adni_conv3d_geom* g, const adni_bf16* x, const adni_bf16* dy, float* dw_oti, float* dbias,
                      float* scratch, int engine, void* stream) {
  int rc = check_geom(g);
  if (rc) return rc;
  ADNI_REQUIRE(x && dy && dw_oti, ADNI_EINVAL, "conv3d_wgrad: null pointer");
  int eng;
  rc = resolve_engine(*g, 2, false, engine, &eng);
  if (rc) return rc;
  auto xs = reinterpret_cast<const __nv_bfloat16*>(x);
  auto dys = reinterpret_cast<const __nv_bfloat16*>(dy);
  if (eng == ADNI_ENGINE_MMA_SYNC) {
    ADNI_REQUIRE(dbias == nullptr, ADNI_ENOTSUP, "conv3d_wgrad: the small-channel engine has no bias gradient (use channel_stats)");
    rc = small_conv_wgrad(*g, xs, dys, dw_oti, static_cast<cudaStream_t>(stream));
    if (rc != ADNI_ENOTSUP || engine == ADNI_ENGINE_MMA_SYNC) return rc;
  }
  if (eng == ADNI_ENGINE_TCGEN05) {
    ADNI_REQUIRE(dbias == nullptr, ADNI_ENOTSUP, "conv3d_wgrad: tcgen05 engine has no bias gradient (use channel_stats)");
    if (scratch != nullptr && wgrad_halo_supported(*g))
      return launch_wgrad_halo(*g, xs, dys, scratch, dw_oti, static_cast<cudaStream_t>(stream));
    return tc_wgrad(*g, xs, dys, dw_oti, static_cast<cudaStream_t>(stream));
  }
  return direct_conv_wgrad(*g, xs, dys, dw_oti, dbias, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// Parameter blocks shared by the tcgen05 implicit-GEMM conv kernels and their host-side planner.
#pragma once
#include "common.cuh"

namespace adni {

constexpr int kMaxTaps = 64;
constexpr int kMaxMaps = 8;

// One filter tap as seen by the TMA producer: which tensor map (stride-2 convs address the input through
// 8 parity-class views), the box offset relative to the output-tile origin, and the column offset of the
// tap's K-slice inside the weight matrix.
struct ConvTap {
  int8_t map, dd, dh, dw;
  int32_t kofs;
};

// fprop / dgrad: D[M = output positions, N = out channels] = sum_taps A_tap[M, Kc] * B[N, tap*Kc..]^T
struct IgemmParams {
  CUtensorMap a_maps[kMaxMaps];  // 5-D (C, W, H, D, N) views of the activation tensor, box (64, bw, bh, bd, 1)
  CUtensorMap b_map;             // 2-D (K_total, N_total) weight matrix, box (64, BLOCK_N)
  ConvTap taps[kMaxTaps];
  int a_ext[kMaxMaps][3];  // (D, H, W) extent of every A view, for the all-padding tap skip
  int ntaps;
  int kc_blocks;  // channels per tap / 64
  int N, Do, Ho, Wo;
  int bd, bh, bw;
  int tiles_d, tiles_h, tiles_w;
  int n_tiles;  // N_total / BLOCK_N
  int n_total;
  long long out_sn, out_sd, out_sh, out_sw;  // element strides of the output view (channel stride 1)
  __nv_bfloat16* out;
  const __nv_bfloat16* addend;  // same view as out, nullable
  const float* bias;            // [N_total], nullable
  double* stat_sum;             // [N_total], nullable
  double* stat_sq;
  // BatchNorm-backward sums fused into a dgrad epilogue (the gradient this kernel writes is the `dout` of a preceding
  // BatchNorm): with red_y set, stat_sum / stat_sq receive  sum g  and  sum g*y  per channel, g = the bf16 value stored
  // in `out`, masked by that layer's ReLU: red_mask > 0 (its stored output), or y*red_scale + red_shift > 0, or no mask.
  const __nv_bfloat16* red_y;     // same view as out
  const __nv_bfloat16* red_mask;  // same view as out, nullable
  const float* red_scale;         // [N_total], nullable
  const float* red_shift;
  int debug;  // diagnostics only (ADNI_DEBUG_MODE): 1 = no MMA issue, 2 = no TMA loads, 3 = no epilogue stores
  // tile-index decoding without divisions: fd_mul[i] = floor(2^32 / d_i) + 1 (0 for d_i == 1) for d = (n_tiles, tiles_w,
  // tiles_h, tiles_d); fd_ok = the products stay below 2^32 (set by finish_igemm_params)
  uint32_t fd_mul[4];
  int fd_ok;
  // 1x1x1 stride-1 problems are plain GEMMs over the flattened positions (`flat` = 1: N = D = H = 1, W = all positions,
  // box = 128 consecutive rows): the epilogue stages 32-row x 64-channel bf16 blocks per warp and hands them to TMA
  // (out_map / add_map: 2-D (channels, positions) views of out / addend, box (64, 32), 128-byte swizzle).
  int flat;
  CUtensorMap out_map;
  CUtensorMap add_map;
};

// host: fills fd_mul / fd_ok from the tile counts (call after every other field is set)
void finish_igemm_params(IgemmParams& p);

// up to 8 independent fprop / dgrad problems of one tile shape in one launch (the parity classes of a stride-2 dgrad)
constexpr int kMaxMultiProblems = 8;
struct IgemmMulti {
  IgemmParams cls[kMaxMultiProblems];
  int ncls;
};

constexpr int kSkMaxCtas = 160;   // CTAs a stream-K schedule table holds (>= the SM count)
struct SkNone {
  int unused;
};

// wgrad: D[M = Cout rows, N = K_total columns] += sum over position boxes of dY^T * X_tap
struct WgradParams {
  CUtensorMap dy_map;            // 5-D (Cout, Wo, Ho, Do, N), box (64, bw, bh, bd, 1), bw*bh*bd == 64
  CUtensorMap x_maps[kMaxMaps];  // 5-D views of x, box (64, bw, bh, bd, 1)
  ConvTap taps[kMaxTaps];
  int x_ext[kMaxMaps][3];
  int ntaps;
  int cin_blocks;  // Cin / 64 (64-column groups per tap)
  int n_groups;    // ntaps * cin_blocks
  int N, Do, Ho, Wo;
  int bd, bh, bw;
  int tiles_d, tiles_h, tiles_w;
  int pos_boxes;  // N * tiles_d * tiles_h * tiles_w
  int m_tiles;    // ceil(Cout / 128)
  int n_tiles;    // ceil(n_groups / GROUPS_PER_TILE)
  int splits;
  int boxes_per_split;
  int cout;
  int k_total;  // n_groups * 64
  float* dw;    // [Cout][K_total] fp32, accumulated with red.add
};

// Stream-K schedule of wgrad2 (conv_wgrad2.cu): the active (position box x output tile) blocks form one sequence cut
// into equal ranges.  The sequence is CHUNK-major: the position boxes are split into chunks of `chunk_boxes` (whole
// samples, sized so that the dY + X slices of one chunk stay L2-resident), and inside a chunk the order is tile-major,
// so the CTAs running at the same time read the same few samples.  Virtual tile v = chunk * (m_tiles * n_tiles) +
// mt_outer * n_tiles + nt covers the raw boxes [chunk * chunk_boxes, (chunk + 1) * chunk_boxes); CTA b walks virtual
// tiles tile_begin[b] .. tile_last[b] and, inside its first / last one, the raw position boxes from box_begin[b] / up
// to box_end[b].  Partial sums go to dw through red.global.add as in the static schedule, so an output tile may be
// shared by any number of CTAs.
struct W2Sched {
  int tile_begin[kSkMaxCtas];
  int box_begin[kSkMaxCtas];
  int tile_last[kSkMaxCtas];
  int box_end[kSkMaxCtas];
  int ctas;
  int chunk_boxes;
};

// Halo-resident fprop / dgrad for 3x3x3, stride 1, dilation 1, pad 1 convs with Cin == Cout in {64, 128}
// (ResNet layer1 / layer2), see conv_halo.cu: the M tile is one output plane piece of 8 (w) x 16 (h) positions; the
// CTA walks a column of such pieces along d and keeps the (10 x 18)-position input planes in a shared-memory ring.
struct HaloParams {
  CUtensorMap a_map;  // 5-D (C, W, H, D, N) view, box (64, 10, 18, 1, 1), 128-B swizzle
  CUtensorMap b_map;  // 2-D (K_total, N_total) weight matrix, box (64, C)
  int mirror;         // 0: halo offset (od, oh, ow) reads weight tap (od*3+oh)*3+ow (fprop); 1: tap 26 - that (dgrad)
  int N, D, H, W;     // output extents == input extents
  int tiles_h, tiles_w;
  int total;  // N * tiles_h * tiles_w * D plane pieces, split evenly (and contiguously) over the CTAs
  long long out_sn, out_sd, out_sh, out_sw;
  __nv_bfloat16* out;
  const __nv_bfloat16* addend;
  const float* bias;
  double* stat_sum;
  double* stat_sq;
  const __nv_bfloat16* red_y;     // fused BatchNorm-backward sums, as in IgemmParams
  const __nv_bfloat16* red_mask;
  const float* red_scale;
  const float* red_shift;
};

// Optional epilogue of a dgrad call: the BatchNorm-backward reduction of the layer whose output gradient it produces.
struct BnReduceEpilogue {
  const __nv_bfloat16* y;      // raw conv output that BatchNorm normalised (same shape as dx)
  const __nv_bfloat16* mask;   // the layer's ReLU output (mask = > 0) or null
  const float* scale;          // or the BatchNorm scale / shift to recompute the mask from y; null = no ReLU
  const float* shift;
  double* sum_g;               // [C] += sum g
  double* sum_gy;              // [C] += sum g*y
};

}  // namespace adni

/*
 * adni_b200 — C-ABI of the B200-native training hot path for Liz490/multimodal_alzheimer.
 *
 * The reference has no FFI: its hot path is the torch.nn module graph executed by ATen/cuDNN
 * (SURVEY.md §8b).  Each entry point below replaces one of those library dispatches; the comment on
 * every declaration cites the reference call site (file:line under the reference repo) whose
 * arithmetic it takes over.  The host-side mirror in multimodal_alzheimer_b200/ binds these symbols with
 * ctypes (INTEGRATION.md shows the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in _host.
 *   - activations: NDHWC, bf16 (uint16_t storage, "adni_bf16").  Weights for the tensor-core engine:
 *     [Cout][kd][kh][kw][Cin] bf16 ("OTI") and [Cin][kd][kh][kw][Cout] bf16 ("ITO", for dgrad).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing
 *     persistent, and returns 0 or a negative ADNI_E* code; adni_last_error_string() explains.
 *   - no CPU fallback exists: an unsupported shape is ADNI_ENOTSUP, never a silent slow path.
 */
#ifndef ADNI_B200_H
#define ADNI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADNI_OK 0
#define ADNI_EINVAL (-1)  /* bad argument                                    (reference: ValueError)      */
#define ADNI_ENOTSUP (-2) /* shape/feature not supported by any CUDA engine  (no silent fallback)         */
#define ADNI_ECUDA (-3)   /* wraps a cudaError_t                                                          */
#define ADNI_ENOMEM (-4)  /* workspace too small / allocation failure -> torch.cuda.OutOfMemoryError      */

typedef uint16_t adni_bf16;

/* Conv engine selector (tests force one; product code passes AUTO). */
#define ADNI_ENGINE_AUTO 0
#define ADNI_ENGINE_TCGEN05 1 /* implicit GEMM on tcgen05/TMEM fed by TMA            */
#define ADNI_ENGINE_DIRECT 2  /* CUDA-core direct convolution (odd shapes, strided small-channel convs) */
#define ADNI_ENGINE_MMA_SYNC 3 /* warp-level tensor-core engine for the small-channel stacks (Cin 1/8..64, Cout 8..64) */

const char* adni_last_error_string(void);
int adni_version(void);
/* Number of kernels launched by this library since load (bench.py's gpu_launches counter). */
long long adni_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Conv3d.  Replaces torch.nn.Conv3d forward / autograd backward as used by
 *   - MedicalNet ResNet (external; call sites pkg/models/mri_models/anat_cnn.py:18-31,
 *     pkg/models/pet_models/pet_resnet_cnn.py:23-35): k in {7,3,1}, stride {1,2}, dilation {1,2,4},
 *     padding = dilation*(k-1)/2, no bias;
 *   - Small_PET_CNN / Anat_CNN head convs (pkg/models/pet_models/pet_cnn.py:21,
 *     pkg/models/mri_models/anat_cnn.py:56): padding='same', bias.
 * Geometry is isotropic (same k/stride/pad/dil on D,H,W), as everywhere in the reference.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int N, D, H, W; /* input batch and spatial extent */
  int Cin, Cout;
  int k, stride, pad, dil;
} adni_conv3d_geom;

/* out spatial extent for one axis */
int adni_conv3d_out_extent(int in, int k, int stride, int pad, int dil);

/* Planner query (measurement only): which engine a geometry is routed to for pass 0 = fprop, 1 = dgrad, 2 = wgrad
 * (engine_kind 0 = direct CUDA-core, 1 = tcgen05 tap-per-box implicit GEMM, 2 = tcgen05 halo-resident, 3 = mma.sync
 * small-channel engine) and the
 * fraction of the (tile, tap) MMA blocks that are actually issued - taps whose shifted box lies entirely in the
 * zero padding are skipped, so executed FLOPs = algorithmic FLOPs x executed_fraction. */
int adni_conv3d_plan_info(const adni_conv3d_geom* g, int pass, int* engine_kind, double* executed_fraction);

/* Planner query (host only, no launch; tests / diagnostics): the stream-K schedule the tcgen05 wgrad kernel would run
 * for this geometry.  header[16] = {used, ctas, chunk_boxes, pos_boxes, m_tiles, n_tiles, groups_per_tile, n_groups,
 * cin_blocks, bd, bh, bw, tiles_d, tiles_h, tiles_w, m_tile_config}; table[4*c..] = {tile_begin, box_begin, tile_last,
 * box_end} of CTA c over the chunk-major virtual tiles (chunk * m_tiles * n_tiles + m_tile * n_tiles + n_tile).
 * Replaces nothing in the reference (cuDNN picks its own wgrad algorithm behind anat_cnn.py:95's backward). */
int adni_conv3d_wgrad_schedule(const adni_conv3d_geom* g, int* header, int* table, int table_capacity);

/* y[N,Do,Ho,Wo,Cout] = conv(x[N,D,H,W,Cin], w_oti) (+bias).  If stat_sum/stat_sqsum are non-null,
 * per-channel sum(y) and sum(y*y) (fp64, from the fp32 accumulators) are ADDED into them: the
 * BatchNorm3d batch statistics fused into the conv epilogue (MedicalNet bn1/bn2/bn3). */
int adni_conv3d_fprop(const adni_conv3d_geom* g, const adni_bf16* x, const adni_bf16* w_oti, const float* bias,
                      adni_bf16* y, double* stat_sum, double* stat_sqsum, int engine, void* stream);

/* dx[N,D,H,W,Cin] = conv_transpose(dy[N,Do,Ho,Wo,Cout], w) (+ addend, same shape as dx, may be null).
 * w_ito is the [Cin][taps][Cout] copy of the weights. */
int adni_conv3d_dgrad(const adni_conv3d_geom* g, const adni_bf16* dy, const adni_bf16* w_ito,
                      const adni_bf16* addend, adni_bf16* dx, int engine, void* stream);

/* adni_conv3d_dgrad whose epilogue also performs the BatchNorm-backward reduction of the layer that PRODUCED dx's
 * tensor (MedicalNet blocks: conv -> bn -> relu; the dx of one conv is the `dout` of the preceding BatchNorm3d):
 *   sum_g[c] += sum g,  sum_gy[c] += sum g * bn_y   with g = the bf16 value stored in dx, masked by that layer's ReLU:
 *   bn_relu_out > 0 (the layer's stored output; blocks ending in bn -> (+ residual) -> relu), or
 *   bn_y * bn_scale[c] + bn_shift[c] > 0 (recomputed; conv -> bn -> relu), or no mask (all three null).
 * bn_y / bn_relu_out have dx's shape.  sum g*xhat = invstd * (sum_gy - mean * sum_g): adni_bn_bwd_apply takes the
 * sums in this form (red_form 1).  Replaces one full read of dx and bn_y per BatchNorm layer (adni_bn_bwd_reduce).
 * tcgen05 engines only (ADNI_ENOTSUP otherwise: call adni_conv3d_dgrad + adni_bn_bwd_reduce). */
/* 1 when the fused form is the faster one for this geometry (long main loops hide the epilogue's extra reads; host-side
 * query, no launch): callers use adni_conv3d_dgrad + adni_bn_bwd_reduce otherwise. */
int adni_conv3d_dgrad_bnred_profitable(const adni_conv3d_geom* g);
int adni_conv3d_dgrad_bnred(const adni_conv3d_geom* g, const adni_bf16* dy, const adni_bf16* w_ito, const adni_bf16* addend,
                            adni_bf16* dx, const adni_bf16* bn_y, const adni_bf16* bn_relu_out, const float* bn_scale,
                            const float* bn_shift, double* sum_g, double* sum_gy, void* stream);

/* dw_oti[Cout][taps][Cin] (fp32) += sum over positions of dy^T * im2col(x).  The caller zeroes dw
 * (the split-K partial sums are accumulated with red.global.add). dbias[Cout] (fp32, may be null) +=
 * sum(dy).  `scratch` (nullable): adni_conv3d_wgrad_scratch_floats(g) ZEROED floats; when given and non-empty the
 * geometry (64 -> 64, 3x3x3, stride 1: ResNet layer1) runs on the halo-plane wgrad engine, which accumulates in
 * [tap][Cin][Cout] order in the scratch and then OVERWRITES dw_oti. */
long long adni_conv3d_wgrad_scratch_floats(const adni_conv3d_geom* g);
int adni_conv3d_wgrad(const adni_conv3d_geom* g, const adni_bf16* x, const adni_bf16* dy, float* dw_oti,
                      float* dbias, float* scratch, int engine, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor-core path of the 1-channel stem  conv1 = Conv3d(1, 64, k=7, stride=2, pad=3, bias=False)  (MedicalNet
 * ResNet; call site anat_cnn.py:18-31).  x is the bf16 volume [N][D][H][W]; adni_stem_expand writes the 16-byte
 * "filter-window pixels" X8[n][d][h][w'][j] = x[n][d][h][2w'-3+j] that TMA + tcgen05 consume.
 * ------------------------------------------------------------------------------------------- */
long long adni_stem_x8_elems(int N, int D, int H, int W);
int adni_stem_expand(const adni_bf16* x, int N, int D, int H, int W, adni_bf16* x8, void* stream);
/* w_ncdhw fp32 [64][1][7][7][7] -> w2g bf16 [56][64][8] (kd*8+kh taps, zero padded). */
int adni_stem_weights(const float* w_ncdhw, adni_bf16* w2g, void* stream);
/* y[N][Do][Ho][Wo][64] (+ fused BatchNorm sum / sum-of-squares, fp64, added into stat_*; may be null). */
int adni_stem_fprop(const adni_bf16* x8, int N, int D, int H, int W, const adni_bf16* w2g, adni_bf16* y,
                    double* stat_sum, double* stat_sqsum, void* stream);
/* grad_ncdhw fp32 [64][1][7][7][7] = dy^T * window(x).  workspace: 512*64 floats (overwritten). */
int adni_stem_wgrad(const adni_bf16* x8, const adni_bf16* dy, int N, int D, int H, int W, float* workspace,
                    float* grad_ncdhw, void* stream);

/* Weight layout conversions at the state_dict boundary (fp32 NCDHW nn.Parameter <-> kernel layouts).
 * w_ncdhw: [Cout][Cin][taps] fp32.  Either output may be null. */
int adni_weights_to_kernel_layout(const float* w_ncdhw, int Cout, int Cin, int taps, adni_bf16* w_oti,
                                  adni_bf16* w_ito, void* stream);
/* All conv weights of an encoder in one launch (same result as adni_weights_to_kernel_layout per tensor, taps <= 27).
 * `jobs` is a DEVICE array of n_jobs records laid out as
 *   struct { const float* src; adni_bf16* oti; adni_bf16* ito; int cout, cin, taps, tile_begin, tiles_ci; }
 * (adni_weights_multi_job_bytes() bytes each); tensor j owns blocks [tile_begin, tile_begin + ceil(cout/16)*tiles_ci),
 * tiles_ci = ceil(cin/16); total_tiles is the grid size.  Replaces the per-module weight conversion the reference
 * leaves to cuDNN's internal filter transforms (MedicalNet Conv3d layers, anat_cnn.py:95). */
int adni_weights_multi_job_bytes(void);
int adni_weights_to_kernel_layout_multi(const void* jobs, int n_jobs, int total_tiles, void* stream);

/* dw_oti [Cout][taps][Cin] fp32 -> grad_ncdhw [Cout][Cin][taps] fp32 (accumulate=1 adds into grad). */
int adni_wgrad_to_param_layout(const float* dw_oti, int Cout, int Cin, int taps, float* grad_ncdhw, int accumulate,
                               void* stream);

/* ---------------------------------------------------------------------------------------------
 * BatchNorm3d (train mode) + ReLU + residual add.  Replaces nn.BatchNorm3d / nn.ReLU / `out += residual`
 * inside MedicalNet BasicBlock/Bottleneck and the "batchnorm_begin" layer (anat_cnn.py:49-50).
 * ------------------------------------------------------------------------------------------- */
/* From fp64 sum/sqsum over `count` elements per channel: mean, invstd (biased var, eps), and
 * scale = gamma*invstd, shift = beta - mean*scale; running stats updated with momentum (unbiased var),
 * exactly torch semantics.  running_* may be null. */
int adni_bn_finalize(const double* stat_sum, const double* stat_sqsum, double count, int C, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                     float* mean, float* invstd, float* scale, float* shift, void* stream);

/* out = act( y*scale[c] + shift[c] + residual ), bf16 in/out, `rows` x C.  residual may be null;
 * relu != 0 applies max(.,0).  If out_sum/out_sqsum non-null, per-channel fp64 sum / sum of squares of the
 * (bf16-rounded) output are added into them (statistics for a following BatchNorm, e.g. batchnorm_begin). */
int adni_bn_apply(const adni_bf16* y, const float* scale, const float* shift, const adni_bf16* residual,
                  adni_bf16* out, long long rows, int C, int relu, double* out_sum, double* out_sqsum, void* stream);

/* Backward, pass 1: g = dout * (relu ? out > 0 : 1);  red[0:C] += sum g, red[C:2C] += sum g*xhat with
 * xhat = (y-mean)*invstd.  The ReLU mask comes from `out` (the forward output) or, when out is NULL and there was
 * no residual, is recomputed as y*scale+shift > 0 (saves one tensor read). red is fp64[2C]. */
int adni_bn_bwd_reduce(const adni_bf16* dout, const adni_bf16* out, const adni_bf16* y, const float* mean,
                       const float* invstd, const float* scale, const float* shift, long long rows, int C, int relu,
                       double* red, void* stream);

/* Backward, pass 2: dy = gamma*invstd*( g - red_g/count - xhat*red_gx/count ) (bf16);
 * dres (nullable) = g (the gradient flowing into the residual input); dgamma = red_gx * param_grad_scale, dbeta =
 * red_g * param_grad_scale are written (fp32) when non-null.  `count` is the GLOBAL element count per channel
 * (sync-BN); with all-reduced sums pass param_grad_scale = 1 / world_size so that the SUM of the ranks' parameter
 * gradients (the data-parallel gradient all-reduce) is the full-batch gradient.
 * red_form 0: red = [sum g | sum g*xhat] (adni_bn_bwd_reduce); 1: red = [sum g | sum g*y] as accumulated by the
 * epilogue of adni_conv3d_dgrad_bnred (sum g*xhat = invstd * (sum g*y - mean * sum g), evaluated in fp64). */
int adni_bn_bwd_apply(const adni_bf16* dout, const adni_bf16* out, const adni_bf16* y, const float* mean,
                      const float* invstd, const float* gamma, const float* scale, const float* shift,
                      const double* red, double count, long long rows, int C, int relu, adni_bf16* dy, adni_bf16* dres,
                      float* dgamma, float* dbeta, double param_grad_scale, int red_form, void* stream);

/* Training-mode BatchNorm forward in ONE launch (adni_bn_finalize + adni_bn_apply): scale / shift are derived from the
 * fp64 batch sums inside the apply kernel; bnp (fp32 [4][C]: mean, invstd, scale, shift) is written for the backward
 * pass and the running statistics are updated (torch semantics), both by one block. */
int adni_bn_train_apply(const adni_bf16* y, const double* stat_sum, const double* stat_sqsum, double count,
                        const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                        float* running_var, float* bnp, const adni_bf16* residual, adni_bf16* out, long long rows, int C,
                        int relu, void* stream);

/* dgamma[c] = red[C+c], dbeta[c] = red[c] (fp64 -> fp32): BatchNorm parameter gradients from the LOCAL backward
 * sums (under data parallelism the gradient all-reduce adds the ranks' contributions). */
int adni_bn_param_grads(const double* red, int C, float* dgamma, float* dbeta, void* stream);

/* Eval-mode BatchNorm (module.eval()): scale = gamma/sqrt(running_var+eps), shift = beta - running_mean*scale. */
int adni_bn_eval_params(const float* running_mean, const float* running_var, const float* gamma, const float* beta,
                        float eps, int C, float* scale, float* shift, void* stream);

/* Stand-alone nn.ReLU on bf16 activations (pet_cnn.py:24 when no BatchNorm precedes it): y = max(x,0);
 * dx = dy * (y > 0). n = number of elements (multiple of 8). */
int adni_relu_fwd(const adni_bf16* x, adni_bf16* y, long long n, void* stream);
int adni_relu_bwd(const adni_bf16* dy, const adni_bf16* y, adni_bf16* dx, long long n, void* stream);

/* Training-mode nn.Dropout(p) (pet_cnn.py:26-27,38-40; early_fusion.py:42-43; anat_pet_featuremapfusion.py:50-51):
 * y[i] = keep(i) ? x[i] / (1 - p) : 0 with keep(i) a pure function of (seed, *offset_dev, i) through Philox4x32-10
 * (uniform u in [0,1) from 24 bits, keep iff u >= p).  The mask is not stored: the backward pass calls the same
 * function on dy with the same (seed, offset).  `offset_dev` is a DEVICE int64 (the host wrapper increments it after
 * every forward), so a captured CUDA graph draws a new mask on every replay.  x, y: bf16 (is_f32 = 0, 16-byte
 * aligned) or fp32 (is_f32 = 1) tensors of n elements in any layout (the operation is elementwise). */
int adni_dropout(const void* x, void* y, long long n, int is_f32, double p, unsigned long long seed,
                 const long long* offset_dev, void* stream);

/* Per-channel fp64 sum / sum of squares of a bf16 rows x C tensor (added into sum/sqsum). */
int adni_channel_stats(const adni_bf16* x, long long rows, int C, double* sum, double* sqsum, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pooling.  MaxPool3d(3,2,1) stem (MedicalNet), MaxPool3d(2) (pet_cnn.py:25, anat_cnn.py:62),
 * AdaptiveAvgPool3d(1)+Flatten (anat_cnn.py:66-67, pet_cnn.py:32-33).
 * ------------------------------------------------------------------------------------------- */
/* floor-mode max pool, -inf padding, first maximum in (d,h,w) scan order wins (torch CPU tie rule);
 * argmax (uint8 window slot kd*k*k+kh*k+kw) saved for backward. */
int adni_maxpool3d_fwd(const adni_bf16* x, int N, int D, int H, int W, int C, int k, int stride, int pad,
                       adni_bf16* y, uint8_t* argmax, void* stream);
int adni_maxpool3d_bwd(const adni_bf16* dy, const uint8_t* argmax, int N, int D, int H, int W, int C, int k,
                       int stride, int pad, adni_bf16* dx, void* stream);

/* ReLU -> MaxPool3d(k) with non-overlapping windows (kernel == stride, no padding; k in {2, 3}) in one pass:
 * Conv3d -> ReLU -> MaxPool3d(2) of the small-CNN stacks (pet_cnn.py:21-25, early_fusion.py:37-41).
 * y = relu(max_window x) (relu = 0: plain max pool); argmax = first maximum of x in (d, h, w) order.  The backward pass
 * writes dx[i] = dy[window] where i is the arg-max and (pooled == NULL or pooled[window] > 0), else 0: MaxPool backward
 * and ReLU backward together - the ReLU output is never stored.  floor mode: a ragged axis tail gets no gradient. */
int adni_relu_maxpool_fwd(const adni_bf16* x, int N, int D, int H, int W, int C, int k, int relu, adni_bf16* y,
                          uint8_t* argmax, void* stream);
int adni_relu_maxpool_bwd(const adni_bf16* dy, const uint8_t* argmax, const adni_bf16* pooled, int N, int D, int H, int W,
                          int C, int k, adni_bf16* dx, void* stream);
/* Fused stem tail  bn1 -> ReLU -> MaxPool3d(3,2,1)  (MedicalNet ResNet.forward) and its backward, without
 * materialising the activated 64-channel tensor or its gradient:
 *   fwd   : p, argmax = maxpool(relu(y*scale + shift))            y: raw conv1 output [N][D][H][W][C]
 *   reduce: g = relu_mask(y) * maxpool_bwd(dp, argmax);  red[0:C] += sum g, red[C:2C] += sum g*xhat
 *   apply : dy = gamma*invstd*(g - red_g/count - xhat*red_gx/count)
 * bnp is the fp32 [4][C] block (mean, invstd, scale, shift) written by adni_bn_finalize.
 * y_at_argmax (nullable, pooled shape): the RAW conv output at every window's arg-max.  With it the backward sums need
 * no pass over the full-resolution tensor: each window hands its gradient to exactly one voxel, so
 *   sum g = sum_w dp[w]*mask_w,  sum g*xhat = sum_w dp[w]*mask_w*xhat(y_at_argmax[w])
 * = adni_bn_bwd_reduce(dout = dp, y = y_at_argmax, scale, shift) over the POOLED tensor (1/8 of the elements). */
int adni_bn_relu_maxpool_fwd(const adni_bf16* y, const float* scale, const float* shift, int N, int D, int H, int W,
                             int C, int k, int stride, int pad, adni_bf16* p, uint8_t* argmax, adni_bf16* y_at_argmax,
                             void* stream);
int adni_maxpool_bn_bwd_reduce(const adni_bf16* dp, const uint8_t* argmax, const adni_bf16* y, const float* bnp, int N,
                               int D, int H, int W, int C, int k, int stride, int pad, double* red, void* stream);
int adni_maxpool_bn_bwd_apply(const adni_bf16* dp, const uint8_t* argmax, const adni_bf16* y, const float* bnp,
                              const float* gamma, const double* red, double count, int N, int D, int H, int W, int C,
                              int k, int stride, int pad, adni_bf16* dy, void* stream);
/* feat[N][C] (fp32) = mean over P positions of x[N][P][C]. */
int adni_gap_fwd(const adni_bf16* x, int N, long long P, int C, float* feat, void* stream);
/* dx[N][P][C] (bf16) = dfeat[N][C] / P. */
int adni_gap_bwd(const float* dfeat, int N, long long P, int C, adni_bf16* dx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Heads.  nn.Linear (+ReLU), nn.BatchNorm1d, torch.cat — anat_cnn.py:69-77, anat_pet_fusion.py:42-51,76,
 * tabular_mri_fusion.py:33-42, pet_tabular_fusion.py:47-60, all_modalities_fusion.py:50-57.  fp32.
 * ------------------------------------------------------------------------------------------- */
/* y[B][out] = act(x[B][in] * W[out][in]^T + b).  ldx/ldy = row strides in elements (concat is expressed
 * by writing into a column slice of a wider buffer). */
int adni_linear_fwd(const float* x, int ldx, const float* W, const float* b, float* y, int ldy, int B, int in,
                    int out, int relu, void* stream);
/* Given dy (and y for the ReLU mask when relu): dx[B][in] (nullable; accumulate_dx adds),
 * dW[out][in] += , db[out] += . */
int adni_linear_bwd(const float* x, int ldx, const float* W, const float* y, int ldy, const float* dy, int lddy,
                    float* dx, int lddx, int accumulate_dx, float* dW, float* db, int B, int in, int out, int relu,
                    void* stream);
/* Stand-alone nn.ReLU on fp32 features: dy == NULL -> out = max(x,0); else out = dy * (x > 0) with x the forward
 * output (backward). */
int adni_relu_f32(const float* x, const float* dy, float* out, long long n, void* stream);
/* BatchNorm1d train-mode over B rows, fp32 (anat_cnn.py:72-73).  Statistic sums are exposed so the
 * host can all-reduce them for data parallelism: stats[0:C]=sum x, stats[C:2C]=sum x^2 (fp64). */
int adni_rows_stats_f32(const float* x, int ldx, int B, int C, double* stats, void* stream);
int adni_bn1d_apply(const float* x, int ldx, const float* scale, const float* shift, float* y, int ldy, int B, int C,
                    int relu, void* stream);
int adni_bn1d_bwd_reduce(const float* dy, int lddy, const float* y, int ldy, const float* x, int ldx,
                         const float* mean, const float* invstd, int B, int C, int relu, double* red, void* stream);
int adni_bn1d_bwd_apply(const float* dy, int lddy, const float* y, int ldy, const float* x, int ldx,
                        const float* mean, const float* invstd, const float* gamma, const double* red, double count,
                        int B, int C, int relu, float* dx, int lddx, float* dgamma, float* dbeta, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Losses, evaluated in fp64 (reference casts logits .to(double): anat_cnn.py:104).  logits_f64 selects the
 * storage type of `logits` (and of `dlogits`): 0 = fp32, 1 = fp64.
 *   gamma > 0 : FocalLoss with DETACHED modulating factor (pkg/loss_functions/focalloss.py:20-40, :30);
 *   gamma == 0: nn.CrossEntropyLoss(weight) (anat_cnn.py:84-85); class_weights may be null (all ones).
 * partial[0] += sum_i numerator_i, partial[1] += sum_i normaliser_i (w[y_i] for CE, 1 for focal) so that the
 * host can all-reduce them across ranks;  loss = partial[0]/partial[1].
 * ------------------------------------------------------------------------------------------- */
int adni_loss_fwd(const void* logits, int logits_f64, int ld, const int64_t* target, int B, int C, double gamma,
                  const double* class_weights, double* partial, double* per_sample_coeff, void* stream);
/* dlogits[B][C] = upstream * coeff_i * (softmax_i - onehot_i) / denom, where denom is the global normaliser
 * and upstream the incoming gradient (both DEVICE fp64 scalars), coeff_i = (1-pt)^gamma or w[y_i] saved by
 * adni_loss_fwd. */
int adni_loss_bwd(const void* logits, int logits_f64, int ld, const int64_t* target, int B, int C,
                  const double* per_sample_coeff, const double* denom, const double* upstream, void* dlogits, int lddl,
                  void* stream);

/* ---------------------------------------------------------------------------------------------
 * Test-epoch metrics (pkg/models/base_model.py:119-172 test_epoch_end, :212-236 bootstrap_metric; torchmetrics
 * 0.10.2 MulticlassF1Score(average='macro' | 'none'), MulticlassMatthewsCorrCoef, confusion matrix).
 * preds = argmax(logits[i][0..C)) (fp64 rows of stride ld), targets = labels[i].  For draw d the samples are
 * idx[d][0..n) (the `torch.randint(0, n, (n,))` resample of bootstrap_metric; idx == NULL: the identity, i.e. the
 * plain test-set metrics with draws == 1).  Outputs (each nullable): f1_macro[draws], f1_class[draws][C], mcc[draws]
 * (fp32, torchmetrics' reductions), confmat[draws][C][C] (int64, [target][pred]).  2 <= C <= 8.
 * ------------------------------------------------------------------------------------------- */
int adni_bootstrap_metrics(const double* logits, int ld, const int64_t* labels, const int64_t* idx, int n, int C,
                           int draws, float* f1_macro, float* f1_class, float* mcc, long long* confmat, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Input normalisation (pkg/utils/dataloader.py:213-215, 236-281; pkg/utils/standardization.py:34-55).
 * ------------------------------------------------------------------------------------------- */
/* Per-scan quantile min-max (dataloader.py:239-249,261-270) for `nscans` volumes of `nvox` voxels:
 *   m = (x*mask)[x*mask != 0]; n = len(m); rank = q*(n-1) (fp64); lo=floor, hi=ceil, w = rank-lo;
 *   Q = lerp(sorted[lo], sorted[hi], w) for q and for (1-q) computed as the fp64 expression 1-q;
 *   out = clamp((x-Qmin)/(Qmax-Qmin), 0, 1) * mask.
 * x: fp32 raw intensities, mask: uint8 (0/1).  out_f32 (nullable) fp32 and out_bf16 (nullable) bf16.
 * info (nullable): per scan 8 int64 {n, lo_max, hi_max, lo_min, hi_min, 0,0,0}; qvals (nullable): per scan
 * 2 fp64 {Qmax, Qmin}.  Order statistics are found by an exact radix select on the fp32 keys (no sort).
 * workspace: adni_quantile_workspace_bytes(nscans) bytes. */
size_t adni_quantile_workspace_bytes(int nscans);
int adni_quantile_minmax_normalize(const float* x, const uint8_t* mask, int nscans, long long nvox, double q,
                                   float* out_f32, adni_bf16* out_bf16, long long* info, double* qvals,
                                   void* workspace, size_t workspace_bytes, void* stream);
/* out = ((double)x - mean)/std [* mask]  (torchvision Normalize in fp64, dataloader.py:213-215,:252-260,:274-278). */
int adni_standardize(const float* x, const uint8_t* mask, long long n, double mean, double std, float* out_f32,
                     adni_bf16* out_bf16, void* stream);
/* Per-scan E[x], E[x^2] in fp64 (standardization.py:43-46): moments[2*s] += mean(x_s), moments[2*s+1] += mean(x_s^2). */
int adni_scan_moments(const float* x, int nscans, long long nvox, double* moments, void* stream);
/* Masked per-scan mean / unbiased std over non-zero masked voxels (dataloader.py:252-256 'normalize' variant):
 * out[3*s] = n, out[3*s+1] = mean, out[3*s+2] = std (fp64). */
int adni_masked_std_mean(const float* x, const uint8_t* mask, int nscans, long long nvox, double* out, void* stream);

/* Casts between the module boundary (fp32 NCDHW with C==1, or fp64) and the kernel layout. */
int adni_cast_f32_to_bf16(const float* x, adni_bf16* y, long long n, void* stream);
int adni_cast_f64_to_bf16(const double* x, adni_bf16* y, long long n, void* stream);
int adni_cast_bf16_to_f32(const adni_bf16* x, float* y, long long n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Volume-level fusion operators (early fusion and feature-map fusion models).
 * ------------------------------------------------------------------------------------------- */
/* Multi-modality module input: x [N][C][vox] fp32 (x_is_f64 = 0) or fp64 NCDHW -> out bf16 NDHWC [N][vox][C], C in
 * {2, 3, 4}  (pkg/models/fusion_models/early_fusion.py:84-88: torch.stack((x_pet, x_mri), dim=1).to(float32)). */
int adni_volumes_to_ndhwc_bf16(const void* x, int x_is_f64, int N, int C, long long vox, adni_bf16* out, void* stream);
/* Voxel-wise maximum of two feature maps, out = a >= b ? a : b, and its gradient routing da = (a >= b) dout,
 * db = (a < b) dout (either may be null): torch.max over the stacked pair returns the FIRST maximal index, so ties -
 * two post-ReLU zeros - go to `a`, the PET branch (anat_pet_featuremapfusion.py:121-123).  n % 8 == 0. */
int adni_maxout_fwd(const adni_bf16* a, const adni_bf16* b, adni_bf16* out, long long n, void* stream);
int adni_maxout_bwd(const adni_bf16* dout, const adni_bf16* a, const adni_bf16* b, adni_bf16* da, adni_bf16* db,
                    long long n, void* stream);
/* Channel concatenation of two NDHWC feature maps (torch.cat(dim=1) of the reference's NCDHW tensors,
 * anat_pet_featuremapfusion.py:118-119): out[row] = [a[row] | b[row]], and the split of its gradient (da / db may be
 * null).  Ca, Cb multiples of 8. */
int adni_concat_channels(const adni_bf16* a, int Ca, const adni_bf16* b, int Cb, long long rows, adni_bf16* out,
                         void* stream);
int adni_split_channels(const adni_bf16* dout, int Ca, int Cb, long long rows, adni_bf16* da, adni_bf16* db,
                        void* stream);
/* nn.Conv3d(padding='same') with an EVEN kernel (filter_size_fusion = 4, train_anat_pet_featuremapfusion.py:70) pads
 * total = dil*(k-1) voxels per axis as lo = total/2, hi = total - lo.  The conv entry points take the symmetric part
 * (pad = lo); these two supply the extra high-side voxels: out [N][D+ed][H+eh][W+ew][C] = x with a zero border, and
 * the inverse for the input gradient (out [N][D][H][W][C] = leading box of x_padded). */
int adni_pad_volume_high(const adni_bf16* x, int N, int D, int H, int W, int C, int ed, int eh, int ew, adni_bf16* out,
                         void* stream);
int adni_crop_volume_high(const adni_bf16* x_padded, int N, int D, int H, int W, int C, int ed, int eh, int ew,
                          adni_bf16* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer step: torch.optim.Adam (amsgrad=False) with L2 weight decay and one learning rate per tensor - what
 * every configure_optimizers of the path builds (pkg/models/mri_models/anat_cnn.py:111-136: one param group per
 * encoder tensor at lr_pretrained, the head at lr, weight_decay = l2_reg; fusion_models/anat_pet_fusion.py:94-127;
 * fusion_models/all_modalities_fusion.py:98-137).  SURVEY.md 8(f) N1.
 * The seven tables are HOST arrays of n_tensors entries: fp32 DEVICE pointers params / grads / exp_avg / exp_avg_sq
 * (numel[i] elements each), steps[i] = DEVICE fp32 scalar holding the number of updates tensor i has received
 * (torch keeps state['step'] the same way; read as t-1, used as t, incremented in stream order after the update),
 * lr[i], weight_decay[i].  Per element, in torch's operation order:
 *   g' = g + wd p;  m += (1-beta1)(g' - m);  v = beta2 v + (1-beta2) g'^2;
 *   p -= lr/(1-beta1^t) * m / (sqrt(v)/sqrt(1-beta2^t) + eps).
 * Tensors travel by value in the kernel parameters, adni_adam_max_tensors_per_launch() per launch (graph-capturable:
 * no host-side state).  hyper_dev, if non-null, is a DEVICE fp32 table [2][n_tensors] (row 0 learning rates, row 1
 * weight decays) that overrides lr[] / weight_decay[] and is read when the kernel RUNS: a captured CUDA graph then
 * follows ReduceLROnPlateau (anat_cnn.py:131-135) through whatever the host last uploaded into the table.
 * Entries with numel 0 are skipped. */
int adni_adam_max_tensors_per_launch(void);
int adni_adam_step_multi(int n_tensors, void* const* params, const void* const* grads, void* const* exp_avg,
                         void* const* exp_avg_sq, void* const* steps, const long long* numel, const float* lr,
                         const float* weight_decay, const float* hyper_dev, double beta1, double beta2, double eps,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: one-shot all-reduce (sum) of a small fp64 vector over NVLink peer memory - the synchronised-BatchNorm
 * statistic sums and the loss normaliser of the data-parallel step (the reference is single-GPU).
 * `peers` is a DEVICE array of `world` pointers (uint64), entry r = this process's mapping of rank r's symmetric
 * buffer of adni_peer_buffer_bytes(world, max_n) bytes, zero-initialised before the first call; `call_counter` is a
 * device uint64, zero-initialised, private to this (rank, buffer).  Every rank must issue the same sequence of calls
 * on a buffer.  In place; all ranks obtain bit-identical sums (fixed rank order). */
size_t adni_peer_buffer_bytes(int world, int max_n);
int adni_peer_allreduce_f64(double* data, int n, const void* peers, void* call_counter, int rank, int world, int max_n,
                            void* stream);

/* Two-shot all-reduce (sum, in place) of a range of the fp32 gradient bucket over NVLink peer memory: the gradient
 * exchange of the data-parallel step (the reference's optimizers see full-batch gradients, anat_cnn.py:111-128,
 * anat_pet_fusion.py:94-118; it is single-GPU itself).  The bucket of every rank is a symmetric-memory buffer;
 * `data_peers` / `flag_peers` are DEVICE arrays of `world` pointers (uint64): this process's mappings of every rank's
 * bucket and of every rank's flag buffer (adni_peer_grad_flag_bytes() bytes, zeroed before the first call).
 * [elem_offset, elem_offset + n) is the range reduced by this call (multiples of 4 elements).  `call_counter`
 * (device uint64) and `arrive_counter` (device uint32) are zero-initialised and private to this (rank, flag buffer).
 * Rank r sums the r-th slice of the range over all ranks in rank order and writes it back to all ranks: results are
 * bit-identical on every rank and deterministic.  Every rank must issue the same sequence of calls. */
size_t adni_peer_grad_flag_bytes(void);
int adni_peer_allreduce_f32(long long elem_offset, long long n, const void* data_peers, const void* flag_peers,
                            void* call_counter, void* arrive_counter, int rank, int world, int ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADNI_B200_H */

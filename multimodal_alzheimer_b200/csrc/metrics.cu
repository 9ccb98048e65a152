// Evaluation metrics of the test epoch on the device (reference pkg/models/base_model.py:119-172
// test_epoch_end and :212-236 bootstrap_metric): macro / per-class F1, Matthews correlation and the confusion matrix
// of (argmax of the logits, label), for the whole test set and for every one of the 1000 bootstrap resamples in ONE
// launch.  The reference loops 2 x 1000 times over torchmetrics objects (MulticlassF1Score /
// MulticlassMatthewsCorrCoef, torchmetrics 0.10.2 - third party, absent from the reference tree; its published
// reductions are restated here and in oracle/metrics.py):
//   confusion matrix M[target][pred]; tp = diag, fp = column sum - tp, fn = row sum - tp (int64)
//   F1_c   = safe_divide(2 tp, 2 tp + fn + fp)                 (a zero denominator counts as 1)
//   macro  = sum_c w_c F1_c / sum_c w_c, w_c = 0 for a class with tp + fp + fn == 0 else 1
//   MCC    = (c s - sum t_k p_k) / sqrt((s^2 - sum p_k^2)(s^2 - sum t_k^2)), 0 when the denominator is 0
// all in fp32 after the integer counts, like torchmetrics.
//
// One CTA per draw: threads gather (pred, label) through the draw's index row, count into a shared-memory matrix
// (C <= 8), thread 0 reduces.  Integer work; latency-bound (1000 x n gathers of 16 bytes for n ~ 10^2..10^3).
#include "common.cuh"

namespace adni {
extern void count_launch();
namespace {

constexpr int kMaxClasses = 8;
constexpr int kMetricThreads = 128;

__global__ void __launch_bounds__(kMetricThreads)
    bootstrap_metrics_kernel(const double* __restrict__ logits, int ld, const int64_t* __restrict__ labels,
                             const int64_t* __restrict__ idx /* [draws][n] or null = identity */, int n, int C,
                             float* __restrict__ f1_macro, float* __restrict__ f1_class /* [draws][C] */,
                             float* __restrict__ mcc, long long* __restrict__ confmat /* [draws][C][C] */) {
  pdl_enter();
  __shared__ unsigned int cm[kMaxClasses * kMaxClasses];
  const int draw = blockIdx.x;
  for (int i = threadIdx.x; i < C * C; i += kMetricThreads) cm[i] = 0u;
  __syncthreads();
  const int64_t* row = idx ? idx + static_cast<long long>(draw) * n : nullptr;
  for (int i = threadIdx.x; i < n; i += kMetricThreads) {
    const long long s = row ? row[i] : i;
    const double* z = logits + s * ld;
    int best = 0;  // torch.argmax: the first maximal value
    double bv = z[0];
    for (int c = 1; c < C; c++) {
      if (z[c] > bv) {
        bv = z[c];
        best = c;
      }
    }
    const int t = static_cast<int>(labels[s]);
    if (t >= 0 && t < C) atomicAdd(&cm[t * C + best], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tk[kMaxClasses], pk[kMaxClasses];
    long long tr = 0, total = 0;
    for (int c = 0; c < C; c++) tk[c] = pk[c] = 0;
    for (int t = 0; t < C; t++)
      for (int p = 0; p < C; p++) {
        const long long v = cm[t * C + p];
        if (confmat) confmat[(static_cast<long long>(draw) * C + t) * C + p] = v;
        tk[t] += v;
        pk[p] += v;
        total += v;
        if (t == p) tr += v;
      }
    // F1 (torchmetrics _fbeta_reduce, beta = 1, average = 'macro' / 'none')
    float wsum = 0.f, acc = 0.f;
    for (int c = 0; c < C; c++) {
      const long long tp = cm[c * C + c], fp = pk[c] - tp, fn = tk[c] - tp;
      const long long den = 2 * tp + fn + fp;
      const float f = __fdiv_rn(static_cast<float>(2 * tp), static_cast<float>(den == 0 ? 1 : den));
      if (f1_class) f1_class[static_cast<long long>(draw) * C + c] = f;
      const float w = (tp + fp + fn == 0) ? 0.f : 1.f;
      acc = __fadd_rn(acc, __fmul_rn(w, f));
      wsum = __fadd_rn(wsum, w);
    }
    if (f1_macro) f1_macro[draw] = __fdiv_rn(acc, wsum);
    // MCC (torchmetrics _matthews_corrcoef_reduce)
    if (mcc) {
      const float c_ = static_cast<float>(tr), s_ = static_cast<float>(total);
      float stp = 0.f, spp = 0.f, stt = 0.f;
      for (int c = 0; c < C; c++) {
        const float t_ = static_cast<float>(tk[c]), p_ = static_cast<float>(pk[c]);
        stp = __fadd_rn(stp, __fmul_rn(t_, p_));
        spp = __fadd_rn(spp, __fmul_rn(p_, p_));
        stt = __fadd_rn(stt, __fmul_rn(t_, t_));
      }
      const float cov_ytyp = __fsub_rn(__fmul_rn(c_, s_), stp);
      const float cov_ypyp = __fsub_rn(__fmul_rn(s_, s_), spp);
      const float cov_ytyt = __fsub_rn(__fmul_rn(s_, s_), stt);
      const float denom = __fmul_rn(cov_ypyp, cov_ytyt);
      mcc[draw] = denom == 0.f ? 0.f : __fdiv_rn(cov_ytyp, __fsqrt_rn(denom));
    }
  }
}

}  // namespace
}  // namespace adni

using namespace adni;

extern "C" {

int adni_bootstrap_metrics(const double* logits, int ld, const int64_t* labels, const int64_t* idx, int n, int C,
                           int draws, float* f1_macro, float* f1_class, float* mcc, long long* confmat, void* stream) {
  ADNI_REQUIRE(logits && labels && n > 0 && draws > 0 && ld >= C, ADNI_EINVAL, "bootstrap_metrics: bad arguments");
  ADNI_REQUIRE(C >= 2 && C <= kMaxClasses, ADNI_ENOTSUP, "bootstrap_metrics: %d classes (2..%d supported)", C, kMaxClasses);
  pdl_launch(bootstrap_metrics_kernel, draws, kMetricThreads, 0, static_cast<cudaStream_t>(stream))(
      logits, ld, labels, idx, n, C, f1_macro, f1_class, mcc, confmat);
  count_launch();
  ADNI_LAUNCH_CHECK("bootstrap_metrics_kernel");
  return ADNI_OK;
}

}  // extern "C"

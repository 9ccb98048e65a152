"""Oracle restatement of the test-epoch metrics (test infrastructure).

Reference: pkg/models/base_model.py:119-172 (`test_epoch_end`) and :212-236 (`bootstrap_metric`) call torchmetrics
0.10.2 (environment.yml:204) `MulticlassF1Score(average='macro' | 'none')` and `MulticlassMatthewsCorrCoef`.
torchmetrics is a third-party dependency ABSENT from the reference tree and from this image: PARITY UNPINNED against
it; its published reductions (torchmetrics/functional/classification/f_beta.py `_fbeta_reduce`,
utilities/compute.py `_safe_divide` / `_adjust_weights_safe_divide`, matthews_corrcoef.py
`_matthews_corrcoef_reduce`) are restated below in the same dtype (fp32 after the int64 counts) and cross-checked
against scikit-learn (tests/test_oracle.py), with which they coincide whenever every class occurs.
"""
import torch


def confusion_matrix(logits, labels, num_classes):
    """[target][pred] counts (int64); preds = argmax over the class axis (first maximum)."""
    preds = logits.argmax(dim=1)
    cm = torch.zeros((num_classes, num_classes), dtype=torch.int64)
    for t, p in zip(labels.tolist(), preds.tolist()):
        cm[t, p] += 1
    return cm


def f1_from_confmat(cm):
    """(macro, per-class) like MulticlassF1Score(average='macro') / (average='none')."""
    tp = cm.diag()
    fp = cm.sum(0) - tp
    fn = cm.sum(1) - tp
    num = 2 * tp
    den = 2 * tp + fn + fp
    den = torch.where(den == 0, torch.ones_like(den), den)          # _safe_divide
    score = num / den                                                 # int64 / int64 -> float32
    w = torch.ones_like(score)
    w[tp + fp + fn == 0] = 0.0                                        # _adjust_weights_safe_divide (macro)
    macro = (w * score).sum(-1) / w.sum(-1)
    return macro, score


def mcc_from_confmat(cm):
    tk = cm.sum(dim=-1).float()
    pk = cm.sum(dim=-2).float()
    c = torch.trace(cm).float()
    s = cm.sum().float()
    cov_ytyp = c * s - sum(tk * pk)
    cov_ypyp = s ** 2 - sum(pk * pk)
    cov_ytyt = s ** 2 - sum(tk * tk)
    denom = cov_ypyp * cov_ytyt
    if denom == 0:
        return torch.tensor(0.0)
    return cov_ytyp / torch.sqrt(denom)


def bootstrap_metric(metric, y_hat, y_labels, num_classes, n_drawings=1000, draws=None):
    """base_model.py:212-236.  `metric` in {'f1', 'mcc'}; `draws` (n_drawings, n) replaces the torch.randint calls
    (same values when generated in the same order from the same global seed)."""
    values = torch.zeros(n_drawings)
    n = len(y_hat)
    for i in range(n_drawings):
        mask = torch.randint(0, n, (n,)) if draws is None else draws[i]
        cm = confusion_matrix(y_hat[mask], y_labels[mask], num_classes)
        values[i] = f1_from_confmat(cm)[0] if metric == "f1" else mcc_from_confmat(cm)
    return torch.mean(values), 1.96 * torch.std(values), values

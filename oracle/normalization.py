"""Normalisation oracles (test infrastructure): the exact CPU sequences of the reference DataLoader.

quantile_minmax_oracle   pkg/utils/dataloader.py:239-249,261-270 ('per_scan_norm': 'min_max')
masked_zscore_oracle     pkg/utils/dataloader.py:252-260          ('per_scan_norm': 'normalize')
global_zscore_oracle     pkg/utils/dataloader.py:274-278          ('all_scan_norm')
pet_standardize_oracle   pkg/utils/dataloader.py:213-215          (torchvision Normalize on the fp64 tensor)
split_moments_oracle     pkg/utils/standardization.py:34-55
"""
import math

import torch


def _normalize(t, mean, std):
    # torchvision.transforms.Normalize on a (D,H,W) tensor: (t - mean) / std with scalar mean/std
    return (t - mean) / std


def quantile_minmax_oracle(mri, mask, q):
    """mri, mask: fp64 tensors (D,H,W).  Returns (normalised fp64 volume, meta)."""
    mri = mri.clone()
    data_masked = mri * mask                                            # dataloader.py:245
    data_masked = data_masked.reshape(-1)                               # :248
    data_masked = data_masked[data_masked.nonzero()]                    # :249  shape (n, 1)
    assert 0 <= q <= 1                                                  # :262
    quant_max = torch.quantile(data_masked, q, interpolation="linear")        # :263
    quant_min = torch.quantile(data_masked, 1 - q, interpolation="linear")    # :264
    out = (mri - quant_min) / (quant_max - quant_min)                   # :266
    out[out > 1] = 1                                                    # :267
    out[out < 0] = 0                                                    # :268
    out *= mask                                                         # :270
    n = data_masked.numel()
    meta = {"n": n, "qmax": float(quant_max), "qmin": float(quant_min)}
    for name, qq in (("max", q), ("min", 1 - q)):
        pos = qq * (n - 1)                      # aten quantile: rank = q * (n - 1) in the input dtype (fp64)
        meta["lo_" + name] = int(math.floor(pos))
        meta["hi_" + name] = int(math.ceil(pos))
        meta["w_" + name] = pos - math.floor(pos)
    return out, meta


def quantile_from_sorted(sorted_vals, q):
    """Re-derivation of torch.quantile('linear') from the order statistics (pins the rank/lerp formula)."""
    n = sorted_vals.numel()
    pos = q * (n - 1)
    lo, hi = int(math.floor(pos)), int(math.ceil(pos))
    w = pos - lo
    a, b = float(sorted_vals[lo]), float(sorted_vals[hi])
    return a + w * (b - a) if w < 0.5 else b - (b - a) * (1 - w)


def masked_std_mean_oracle(mri, mask):
    m = (mri * mask).reshape(-1)
    m = m[m.nonzero()]
    std, mean = torch.std_mean(m)                                       # dataloader.py:254 (unbiased)
    return m.numel(), float(mean), float(std)


def masked_zscore_oracle(mri, mask):
    _, mean, std = masked_std_mean_oracle(mri, mask)
    return _normalize(mri, mean, std) * mask                            # :256-260


def global_zscore_oracle(mri, mean, std):
    return _normalize(mri, mean, std)                                   # :277-278


def pet_standardize_oracle(pet, mean, std):
    return _normalize(pet, mean, std)                                   # :213-215


def split_moments_oracle(scans):
    """standardization.py:43-55: mean of per-scan E[x], sqrt(mean E[x^2] - mean^2). Also returns per-scan moments."""
    mean_x, mean_x2 = 0.0, 0.0
    per = []
    for x in scans:
        ex, ex2 = x.mean(), (x ** 2).mean()
        per.append([float(ex), float(ex2)])
        mean_x = mean_x + ex
        mean_x2 = mean_x2 + ex2
    mean = mean_x / len(scans)
    std = torch.sqrt(mean_x2 / len(scans) - mean ** 2)
    return float(mean), float(std), torch.tensor(per, dtype=torch.float64)

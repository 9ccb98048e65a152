"""Device-side input normalisation: the arithmetic of MultiModalDataset.__getitem__
(reference pkg/utils/dataloader.py:213-215 PET standardisation, :239-270 per-scan MRI normalisation, :274-278 global
z-score) and of pkg/utils/standardization.py:34-55, moved from 32 CPU DataLoader workers onto the GPU.

Inputs are raw fp32 volumes (B, D, H, W) (NIfTI intensities are fp32-exact) and uint8 brain masks; outputs are
fp32 (bit-identical to the reference's fp64 result cast to fp32) or bf16 (the encoder's input type).
"""
import torch

from ... import kernels as K


def _as_mask(mask):
    if mask.dtype != torch.uint8:
        mask = (mask != 0).to(torch.uint8)
    return mask.contiguous()


def normalize_mri_per_scan_min_max(mri, mask, quantile, out_dtype=torch.float32, return_info=False):
    """normalize_mri={'per_scan_norm': 'min_max'}, quantile=q (dataloader.py:261-270)."""
    assert quantile >= 0 and quantile <= 1  # dataloader.py:262
    return K.quantile_minmax_normalize(mri.contiguous(), _as_mask(mask), quantile, out_dtype=out_dtype,
                                       want_info=return_info)


def normalize_mri_per_scan_zscore(mri, mask, out_dtype=torch.float32):
    """normalize_mri={'per_scan_norm': 'normalize'} (dataloader.py:252-260): unbiased std / mean over the non-zero
    masked voxels, (x-mean)/std, re-masked.  Statistics are read back per scan (S small host syncs)."""
    mask = _as_mask(mask)
    stats = K.masked_std_mean(mri.contiguous(), mask).cpu()
    outs = [K.standardize(mri[s].contiguous(), float(stats[s, 1]), float(stats[s, 2]), mask=mask[s],
                          out_dtype=out_dtype) for s in range(mri.shape[0])]
    return torch.stack(outs)


def normalize_all_scan(x, mean, std, out_dtype=torch.float32):
    """normalize_mri={'all_scan_norm': {...}} (dataloader.py:274-278) and normalize_pet (:213-215)."""
    return K.standardize(x.contiguous(), mean, std, out_dtype=out_dtype)


normalize_pet = normalize_all_scan


def compute_std_mean(scans):
    """pkg/utils/standardization.py:34-55 over a (S, D, H, W) fp32 tensor of one split."""
    m = K.scan_moments(scans.contiguous())
    mean = m[:, 0].mean()
    std = torch.sqrt(m[:, 1].mean() - mean ** 2)
    return mean, std

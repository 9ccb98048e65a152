"""Split-level z-score statistics (reference pkg/utils/standardization.py:30-55): the mean over scans of every scan's
E[x] and E[x^2], std = sqrt(E[x^2] - mean^2) - the numbers the train scripts hard-code as `normalize_pet`
(train_pet_cnn.py:77-78) and pass as `normalize_mri={'all_scan_norm': {...}}`.

Reference: `DataLoader(dataset, batch_size=1)` over float64 CPU tensors, two reductions per scan in Python.  Here the
scans are decoded by the staging library into pinned buffers, copied in batches and reduced by one kernel
(`adni_scan_moments`, fp64 accumulation); nothing is normalised on the CPU.
"""
import torch

from ... import kernels as K
from ... import staging


class NormalizeDataset:
    @staticmethod
    def compute_std_mean(dataset, dim=None, modality="mri", device="cuda", batch_size=16, threads=8):
        """dataset: a MultiModalDataset built with normalize_mri=None / normalize_pet=None (raw intensities, as the
        reference script builds it).  Returns (mean, std) as 0-d fp64 tensors, like the reference (`mean, std`)."""
        if dim is not None:
            raise NotImplementedError("per-axis statistics (dim != None) are not used by any reference configuration")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("compute_std_mean reduces on the GPU: a CUDA device is required (no CPU fallback)")
        slot = {"mri": 1, "pet1451": 0}[modality]
        paths = [dataset._paths(i)[slot] for i in range(len(dataset))]
        if not paths or any(p is None for p in paths):
            raise ValueError(f"every sample of the dataset must have the '{modality}' modality")
        shape = staging.read_info(paths[0]).shape[:3]
        buf = torch.empty((batch_size,) + tuple(shape), dtype=torch.float32, pin_memory=True)
        moments = []
        for i in range(0, len(paths), batch_size):
            chunk = paths[i:i + batch_size]
            staging.stage_volumes(chunk, buf[:len(chunk)], threads=threads)
            x = buf[:len(chunk)].to(device, non_blocking=True)
            moments.append(K.scan_moments(x.contiguous()))          # (S, 2): E[x], E[x^2] per scan, fp64
            torch.cuda.current_stream(device).synchronize()          # the pinned buffer is reused by the next chunk
        m = torch.cat(moments)
        mean = m[:, 0].mean()
        std = torch.sqrt(m[:, 1].mean() - mean ** 2)
        return mean.cpu(), std.cpu()

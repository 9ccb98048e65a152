"""Host decode throughput of the staging library at the reference's volume size (MNI 2 mm grid,
91x109x91, .nii.gz): native threads (adni_stage_volumes) vs the numpy/gzip restatement of the reference's
`nib.load(p).get_fdata()` + `torch.tensor(...)` (oracle/nifti.py, one process - what each of the reference's
DataLoader workers does).  CPU only; prints one JSON line.  The GPU part of the pipeline (H2D + normalisation
kernels) is what bench.py's `e2e` leg times."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_alzheimer_b200 import _build, staging  # noqa: E402
from oracle import nifti as N  # noqa: E402


def main(n_files=32, shape=(91, 109, 91)):
    _build.build_stage()
    rng = np.random.default_rng(15)
    root = tempfile.mkdtemp()
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, n) for n in shape], indexing="ij")
    brain = (zz ** 2 + yy ** 2 + xx ** 2) < 0.75
    paths, masks = [], []
    for i in range(n_files):
        vol = np.where(brain, 400 * np.abs(rng.standard_normal(shape)) + 50, 0).astype(np.float32)   # skull-stripped
        p = os.path.join(root, f"mri_{i}.nii.gz")
        N.write_nifti(p, vol)
        pm = os.path.join(root, f"mask_{i}.nii.gz")
        N.write_nifti(pm, brain.astype(np.uint8))
        paths.append(p)
        masks.append(pm)
    size_mb = sum(os.path.getsize(p) for p in paths) / n_files / 1e6
    t0 = time.perf_counter()
    for p, pm in zip(paths[:8], masks[:8]):
        torch.tensor(N.read_fdata(p))
        torch.tensor(N.read_fdata(pm))
    ref = 8 / (time.perf_counter() - t0)
    out = torch.empty((n_files,) + shape, dtype=torch.float32)
    outm = torch.empty((n_files,) + shape, dtype=torch.uint8)
    res = {}
    for threads in (1, 2, 4, 8, 16, 32):
        if threads > 2 * (os.cpu_count() or 1):
            break
        best = 0.0
        for _ in range(3):
            t0 = time.perf_counter()
            staging.stage_volumes(paths, out, threads=threads)
            staging.stage_volumes(masks, outm, threads=threads)
            best = max(best, n_files / (time.perf_counter() - t0))
        res[str(threads)] = round(best, 1)
    assert torch.equal(out[3], torch.tensor(N.read_fdata(paths[3])).float())
    print(json.dumps({"metric": "MRI volume + brain mask decoded per second (91x109x91 .nii.gz -> fp32 / uint8 batch buffer)",
                      "unit": "scans/s", "host_cores": os.cpu_count(), "compressed_mb_per_volume": round(size_mb, 2),
                      "native_by_threads": res, "numpy_gzip_one_process": round(ref, 1),
                      "note": "reference = nibabel get_fdata (gzip + numpy, float64) per DataLoader worker process; "
                              "restated by oracle/nifti.py because nibabel is absent here"}))


if __name__ == "__main__":
    main()

"""A synthetic ADNI-style dataset on disk for the staging tests: the CSV schema of the reference
(pkg/utils/create_csv/data_labels.py:1-20: one row per modality and session; ID, ses, label, path_pet1451,
path_anat, path_anat_mask, AGE ... ICV) and small gzip NIfTI-1 volumes written by oracle/nifti.py.  Deterministic in
`seed`; tools/make_golden_dataset.py runs the reference's own MultiModalDataset on exactly these files."""
import os

import numpy as np
import pandas as pd

from oracle.nifti import write_nifti

SHAPE = (12, 14, 10)
TAB = ["AGE", "PTEDUCAT", "Ventricles", "Hippocampus", "WholeBrain", "Entorhinal", "Fusiform", "MidTemp", "ICV"]


def make_synthetic_adni(root, seed=15, subjects=7):
    rng = np.random.default_rng(seed)
    os.makedirs(root, exist_ok=True)
    rows = []
    labels = ["CN", "MCI", "Dementia"]
    k = 0

    def new_row(sid, ses, label):
        r = {c: np.nan for c in ["path_pet1451", "path_anat", "path_anat_mask"] + TAB}
        r.update(ID=sid, ses=ses, label=label)
        return r

    for s in range(subjects):
        sid = f"sub-{1000 + s}"
        label = labels[s % 3]
        base_day = int(rng.integers(1, 200))
        for visit in range(int(rng.integers(1, 4))):
            day0 = base_day + visit * int(rng.integers(150, 420))
            # each modality of a visit gets its own session date, up to ~7 months apart (days_threshold = 180)
            for mod in ("pet", "mri", "tab"):
                if rng.random() < 0.15:
                    continue
                day = day0 + int(rng.integers(-100, 101))
                ses = (pd.Timestamp("2010-01-01") + pd.Timedelta(days=day)).strftime("%Y-%m-%d")
                r = new_row(sid, ses, label if rng.random() > 0.1 else labels[(s + 1) % 3])
                if mod == "pet":
                    vol = np.maximum(0, rng.normal(0.5145, 0.5383, SHAPE)).astype(np.float32)
                    p = os.path.join(root, f"pet_{k}.nii.gz")
                    write_nifti(p, vol, pad_dims=0)
                    r["path_pet1451"] = p
                elif mod == "mri":
                    # int16 intensities with a scale factor, ~40 % brain, a few exact zeros inside the brain
                    vol = (400 * np.abs(rng.standard_normal(SHAPE)) + 50 * rng.random(SHAPE)).astype(np.int16)
                    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, n) for n in SHAPE], indexing="ij")
                    mask = ((zz ** 2 + yy ** 2 + xx ** 2) < 0.8).astype(np.uint8)
                    vol[rng.random(SHAPE) < 0.01] = 0
                    p = os.path.join(root, f"mri_{k}.nii.gz")
                    pm = os.path.join(root, f"mask_{k}.nii.gz")
                    write_nifti(p, vol, scl_slope=0.5 if k % 2 else float("nan"), scl_inter=0.0)
                    write_nifti(pm, mask, pad_dims=0)
                    r["path_anat"], r["path_anat_mask"] = p, pm
                else:
                    for c in TAB:
                        r[c] = float(np.round(rng.normal(0, 1), 4))
                rows.append(r)
                k += 1
    df = pd.DataFrame(rows, columns=["ID", "ses", "label", "path_pet1451", "path_anat", "path_anat_mask"] + TAB)
    path = os.path.join(root, "train_path_data_labels.csv")
    df.to_csv(path)                                   # with the index column, like data_labels.py:274
    return path

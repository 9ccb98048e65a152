"""Golden records of the reference's OWN MultiModalDataset (pkg/utils/dataloader.py, imported unmodified through
tools/reference_harness.py) on the synthetic dataset of tests/_dataset.py -> tests/golden/dataset.json.

Run in the build container (needs /root/reference).  nibabel is absent here: `nib.load(p).get_fdata()` is served by
oracle/nifti.py (the restatement of nibabel's NIfTI-1 reader; parity unpinned for the file format itself, see its
header).  Everything after the read - the pairing of modalities, label mapping, masking, torch.quantile /
std_mean / Normalize arithmetic, the tabular feature order - is the reference's code.

Recorded per configuration: the paired index (ID, label, file basenames per row), and for the first samples the
normalised volumes as the fp32 values a model sees after `x.to(torch.float32)` (anat_cnn.py:103; little-endian
float32 bytes, base64), the tabular
vector and the label.
"""
import base64
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import reference_harness as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "dataset.json")
PET_NORM = {"mean": 0.5145, "std": 0.5383}                      # train_pet_cnn.py:77-78

CONFIGS = {
    # name: (modalities, binary, normalize_mri, quantile, samples recorded in full)
    "pet": (["pet1451"], False, None, 0.99, 2),
    "mri_minmax_q97": (["t1w"], False, {"per_scan_norm": "min_max"}, 0.97, 3),
    "mri_minmax_q99_binary": (["t1w"], True, {"per_scan_norm": "min_max"}, 0.99, 1),
    "mri_zscore": (["t1w"], False, {"per_scan_norm": "normalize"}, 0.99, 2),
    "mri_allscan": (["t1w"], False, {"all_scan_norm": {"mean": 210.0, "std": 180.0}}, 0.99, 1),
    "tabular": (["tabular"], False, None, 0.99, 3),
    "pet_mri": (["pet1451", "t1w"], False, {"per_scan_norm": "min_max"}, 0.98, 2),
    "mri_tab_binary": (["t1w", "tabular"], True, {"per_scan_norm": "min_max"}, 0.98, 1),
    "pet_tab": (["pet1451", "tabular"], False, None, 0.99, 1),
    "all": (["pet1451", "t1w", "tabular"], False, {"per_scan_norm": "min_max"}, 0.98, 2),
}


def main():
    H.install()
    from oracle import nifti as N

    class _Img:
        def __init__(self, p):
            self.p = p

        def get_fdata(self):
            return N.read_fdata(self.p)

    sys.modules["nibabel"].load = lambda p: _Img(p)
    sys.modules.pop("pkg.utils.dataloader", None)
    import pkg.utils.dataloader as RD

    from tests._dataset import make_synthetic_adni
    root = tempfile.mkdtemp()
    csv = make_synthetic_adni(root, seed=15)
    base = lambda p: None if p is None else os.path.basename(p)  # noqa: E731
    records = {}
    for name, (mods, binary, nmri, q, n_full) in CONFIGS.items():
        ds = RD.MultiModalDataset(csv, binary_classification=binary, modalities=mods, normalize_pet=PET_NORM,
                                  normalize_mri=nmri, quantile=q)
        rec = {"modalities": mods, "binary": binary, "normalize_mri": nmri, "quantile": q, "len": len(ds),
               "columns": list(ds.ds.columns),
               "index": [[r["ID"], r["label"], base(r["path_pet1451"]), base(r["path_anat"]), base(r["path_anat_mask"]),
                          r["AGE"]] for _, r in ds.ds.iterrows()],
               "label_counts": ds.ds["label"].value_counts().reindex(
                   index=["CN", "Dementia"] if binary else ["CN", "MCI", "Dementia"]).tolist(),
               "samples": []}
        for i in range(min(n_full, len(ds))):
            s = ds[i]
            item = {"keys": sorted(s.keys()), "label": int(s["label"])}
            for k in ("mri", "pet1451"):
                if k in s:
                    assert s[k].dtype == torch.float64
                    item[k + "_shape"] = list(s[k].shape)
                    item[k + "_f32_b64"] = base64.b64encode(s[k].to(torch.float32).contiguous().numpy().tobytes()).decode()
            if "tabular" in s:
                item["tabular"] = s["tabular"].tolist()
                item["tabular_dtype"] = str(s["tabular"].dtype)
            rec["samples"].append(item)
        records[name] = rec
        print(f"{name:24s} len {len(ds):3d}  sample keys {rec['samples'][0]['keys'] if rec['samples'] else None}")
    # pairing only, on two more synthetic cohorts (more subjects, other session dates): index rows per configuration
    extra = {}
    for seed, subjects in ((3, 12), (2024, 9)):
        root_s = tempfile.mkdtemp()
        csv_s = make_synthetic_adni(root_s, seed=seed, subjects=subjects)
        extra[str(seed)] = {"subjects": subjects, "index": {}}
        for name in ("pet_mri", "mri_tab_binary", "pet_tab", "all"):
            mods, binary, nmri, q, _ = CONFIGS[name]
            ds = RD.MultiModalDataset(csv_s, binary_classification=binary, modalities=mods, normalize_pet=PET_NORM,
                                      normalize_mri=nmri, quantile=q)
            extra[str(seed)]["index"][name] = [[r["ID"], r["label"], base(r["path_pet1451"]), base(r["path_anat"]),
                                                base(r["path_anat_mask"]), r["AGE"]] for _, r in ds.ds.iterrows()]
            print(f"seed {seed:5d} {name:16s} len {len(ds)}")
    with open(OUT, "w") as f:
        json.dump({"extra_cohorts": extra, "source": "reference pkg/utils/dataloader.py MultiModalDataset, imported unmodified via "
                             "tools/reference_harness.py; nib.load().get_fdata() served by oracle/nifti.py",
                   "dataset": "tests/_dataset.py::make_synthetic_adni(seed=15)", "pet_norm": PET_NORM,
                   "configs": records}, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()

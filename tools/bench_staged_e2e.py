"""End-to-end training throughput FROM FILES: .nii.gz MRI + brain mask + PET on disk ->
MultiModalDataset / StagedLoader (native decode -> pinned slots -> H2D -> GPU normalisation) -> two-branch ResNet-18
fusion model training step (forward, focal loss, backward, Adam) on one B200.  Reports volumes/s for a cold epoch
(every file inflated) and a cached epoch (`enable_cache`), next to the host decode rate alone.  Needs a GPU; run as
  gpurun --timeout 600 -- 'python tools/bench_staged_e2e.py > gpurun_out/staged_e2e.json'
The volumes are synthetic on the reference's grid (MNI 2 mm, 91x109x91) unless --volume is given."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.nifti import write_nifti  # noqa: E402  (fixture writer only)


def make_dataset(root, pairs, shape, seed=15):
    import pandas as pd
    rng = np.random.default_rng(seed)
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, n) for n in shape], indexing="ij")
    brain = ((zz / 0.84) ** 2 + (yy / 0.84) ** 2 + (xx / 0.84) ** 2) <= 1
    cols = ["ID", "ses", "label", "path_pet1451", "path_anat", "path_anat_mask", "AGE", "PTEDUCAT", "Ventricles",
            "Hippocampus", "WholeBrain", "Entorhinal", "Fusiform", "MidTemp", "ICV"]
    rows = []
    mask_path = os.path.join(root, "mask.nii.gz")
    write_nifti(mask_path, brain.astype(np.uint8))
    for i in range(pairs):
        mri = np.where(brain, 400 * np.abs(rng.standard_normal(shape)) + 50 * rng.random(shape), 0).astype(np.float32)
        pet = np.maximum(0, rng.normal(0.5145, 0.5383, shape)).astype(np.float32)
        pm, pp = os.path.join(root, f"mri_{i}.nii.gz"), os.path.join(root, f"pet_{i}.nii.gz")
        write_nifti(pm, mri)
        write_nifti(pp, pet)
        label = ["CN", "MCI", "Dementia"][i % 3]
        base = {c: np.nan for c in cols}
        rows.append(dict(base, ID=f"s{i}", ses="2015-03-01", label=label, path_pet1451=pp))
        rows.append(dict(base, ID=f"s{i}", ses="2015-03-20", label=label, path_anat=pm, path_anat_mask=mask_path))
    csv = os.path.join(root, "train.csv")
    pd.DataFrame(rows, columns=cols).to_csv(csv)
    return csv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=64)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--volume", type=int, nargs=3, default=[91, 109, 91])
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--depth", type=int, default=18)
    args = ap.parse_args()
    from multimodal_alzheimer_b200 import staging
    from multimodal_alzheimer_b200 import workloads as W
    from multimodal_alzheimer_b200.pkg.utils.dataloader import MultiModalDataset, StagedLoader

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    shape = tuple(args.volume)
    root = tempfile.mkdtemp()
    t0 = time.perf_counter()
    csv = make_dataset(root, args.pairs, shape)
    t_make = time.perf_counter() - t0
    ds = MultiModalDataset(csv, modalities=["pet1451", "t1w"], normalize_pet={"mean": 0.5145, "std": 0.5383},
                           normalize_mri={"per_scan_norm": "min_max"}, quantile=0.98)
    assert len(ds) == args.pairs, (len(ds), args.pairs)

    model = W.build_model(W.product_namespace(), "pet_mri_fusion_r18", depth=args.depth)   # the bench workload's model
    model.to(dev).train()
    opt = model.configure_optimizers()

    def epoch(loader):
        torch.cuda.synchronize()
        t = time.perf_counter()
        n, loss = 0, None
        for batch in loader:
            out = model.general_step(batch, 0, "train")
            opt.zero_grad(set_to_none=True)
            out["loss"].backward()
            opt.step()
            loss = float(out["loss"].detach())          # the per-step D2H read a trainer's logging does
            n += int(batch["label"].numel())
        torch.cuda.synchronize()
        return 2 * n / (time.perf_counter() - t), loss

    # host decode alone (no GPU work): files -> pinned batch buffers
    paths = [ds._paths(i) for i in range(len(ds))]
    buf = torch.empty((len(ds),) + shape, dtype=torch.float32, pin_memory=True)
    t = time.perf_counter()
    staging.stage_volumes([p[1] for p in paths], buf, threads=args.threads)
    staging.stage_volumes([p[0] for p in paths], buf, threads=args.threads)
    decode_vps = 2 * len(ds) / (time.perf_counter() - t)
    del buf

    loader = StagedLoader(ds, batch_size=args.batch, shuffle=True, device=dev, threads=args.threads,
                          generator=torch.Generator().manual_seed(15))
    epoch(loader)                                       # warm-up: allocator, descriptor caches, kernels' first launch
    cold, _ = epoch(loader)
    ds.enable_cache(16.0)
    epoch(loader)                                       # fills the cache
    warm, loss = epoch(loader)
    print(json.dumps({"metric": "MRI+PET volumes/sec train from .nii.gz files (ResNet-18 3D fusion)", "unit": "volumes/s",
                      "volume": list(shape), "pairs": args.pairs, "batch": args.batch, "decode_threads": args.threads,
                      "host_cores": os.cpu_count(), "host_decode_only": round(decode_vps, 1),
                      "train_from_files_cold": round(cold, 1), "train_from_files_cached": round(warm, 1),
                      "loss": loss, "dataset_write_s": round(t_make, 1),
                      "note": "cold = every epoch inflates every file (what the reference's DataLoader workers do, "
                              "plus their CPU normalisation); cached = decoded scans kept in host memory"}))


if __name__ == "__main__":
    main()

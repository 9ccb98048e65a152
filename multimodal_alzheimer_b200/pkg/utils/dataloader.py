"""MultiModalDataset + StagedLoader: the input side of the path (reference pkg/utils/dataloader.py:20-446 and the
`DataLoader(trainset, batch_size, shuffle, num_workers=32)` calls of the train scripts, train_anat_cnn.py:187-198).

Reference: every `__getitem__` gunzips and reads the NIfTI files with nibabel into float64 arrays, normalises them on
the CPU (boolean indexing + two `torch.quantile` sorts per MRI scan) and ships float64 tensors through worker-process
pipes; the model casts to fp32 on the device.  Here:

  index      same constructor, same CSV schema (pkg/utils/create_csv/data_labels.py:1-20), same pairing of modalities
             within `days_threshold` (dataloader.py:100-158, 346-434) -> `self.ds`, `__len__`,
             `get_label_distribution()`.
  decode     `stage(indices, ...)`: native threads (csrc/stage/nifti_stage.cpp, C-ABI include/adni_staging.h) inflate
             the batch's files straight into pinned fp32 / uint8 batch buffers.
  normalise  on the GPU, after one async H2D copy per modality: per-scan quantile min-max / z-score over the brain
             mask, split-level standardisation (pkg/utils/normalization.py -> csrc/normalize.cu), emitting the
             encoder's input type.
  batches    `StagedLoader` yields the reference's batch dict {'mri','pet1451','tabular','label'} with device
             tensors, decoding batch i+1 on a helper thread while batch i trains (two pinned slots).

`__getitem__` returns the RAW staged sample (fp32 intensities, uint8 mask, no normalisation): normalising on the CPU
would be the reference's path, not this one; there is no CPU fallback.
"""
import threading
from datetime import datetime
from typing import Any, Dict, List

import numpy as np
import pandas as pd
import torch
from torch.utils.data import Dataset

from ... import staging
from . import normalization as norm

TABULAR_COLUMNS = ["AGE", "PTEDUCAT", "Ventricles", "Hippocampus", "PTEDUCAT", "Entorhinal", "Fusiform", "MidTemp", "ICV"]
# dataloader.py:292-304: the fifth feature is read from 'PTEDUCAT' again (named whole_brain there); kept as is


def find_corresponding_samples(df, id, label, min_time, max_time, max_days=180):
    """Rows of `df` with the same subject and label whose session lies within `max_days` of both ends of the
    interval already fused (dataloader.py:346-399)."""
    sel = df[(df["ID"] == id) & (df["label"] == label)]
    if len(sel) == 0:
        return sel
    after_min = (sel["ses"] - min_time).dt.days
    before_max = (max_time - sel["ses"]).dt.days
    return sel[(after_min <= max_days) & (before_max <= max_days)].reset_index(drop=True)


def merge_two_dfs(row, matches):
    """Fuse one already-merged sample (`row`) with every matching row of the next modality (dataloader.py:401-434):
    the time interval grows to cover the new session; columns that are empty in `matches` take `row`'s value."""
    out = matches.copy()
    ses = out["ses"]
    out["min_time"] = [s if (row["min_time"] - s).days > 0 else row["min_time"] for s in ses]
    out["max_time"] = [s if (row["max_time"] - s).days < 0 else row["max_time"] for s in ses]
    out = out.drop(columns=["ses"])
    row_nan = row.isna()
    for col in out.columns:
        if out[col].isnull().values.any() and not row_nan[col]:
            out[col] = row[col]
    return out


class MultiModalDataset(Dataset):
    """Same constructor as the reference (dataloader.py:63-75)."""

    def __init__(self, path: str, binary_classification=False, modalities: List[str] = ["pet1451", "t1w", "tabular"],
                 days_threshold: int = 180, transform_pet=None, transform_mri=None, transform_tabular=None,
                 normalize_pet: Dict[str, float] = None, normalize_mri: Dict[str, Any] = None, quantile: float = 0.99):
        self.entire_ds = pd.read_csv(path)
        if binary_classification == 2:
            binary_classification = True
        elif binary_classification == 3:
            binary_classification = False
        self.binary_classification = binary_classification
        if self.binary_classification:
            self.entire_ds = self.entire_ds[self.entire_ds["label"] != "MCI"]
            self.label_mapping = {"CN": 0, "Dementia": 1}
        else:
            self.label_mapping = {"CN": 0, "MCI": 1, "Dementia": 2}
        self.days_threshold = days_threshold
        self.modalities = modalities
        assert len(self.modalities) in range(1, 4)
        assert all([x in ["pet1451", "t1w", "tabular"] for x in self.modalities])
        assert len(set(self.modalities)) == len(self.modalities)

        key = {"pet1451": "path_pet1451", "t1w": "path_anat", "tabular": "AGE"}
        self.df_list = [self.entire_ds.dropna(subset=[key[m]]).reset_index(drop=True)
                        for m in ("pet1451", "t1w", "tabular") if m in self.modalities]   # fixed order, :112-125
        if len(self.df_list) == 1:
            self.ds = self.df_list[0]
        else:
            for df in self.df_list:
                df["ses"] = [datetime.strptime(x, "%Y-%m-%d") for x in df["ses"]]
            base = self.df_list[0].copy()
            base["min_time"] = base["ses"]
            base["max_time"] = base["ses"]
            base = base.drop(columns="ses")
            for nxt in self.df_list[1:]:
                parts = []
                for _, row in base.iterrows():
                    matches = find_corresponding_samples(nxt, row["ID"], row["label"], row["min_time"], row["max_time"],
                                                         self.days_threshold)
                    if len(matches) >= 1:
                        parts.append(merge_two_dfs(row, matches))
                base = pd.concat(parts, ignore_index=True) if parts else pd.DataFrame()
            self.ds = base
        self.ds = self.ds.astype(object).where(self.ds.notna(), None)

        if transform_pet is not None or transform_mri is not None or transform_tabular is not None:
            # numpy callables on float64 host arrays in the reference; no shipped configuration passes one
            raise NotImplementedError("host-side transforms are not part of the device staging path")
        self.transform_pet = self.transform_mri = self.transform_tabular = None
        self.normalize_pet = normalize_pet
        if self.normalize_pet:
            assert "mean" in self.normalize_pet.keys()
            assert isinstance(self.normalize_pet["mean"], float)
            assert "std" in self.normalize_pet.keys()
            assert isinstance(self.normalize_pet["std"], float)
        self.normalize_mri = normalize_mri
        if self.normalize_mri:
            assert isinstance(self.normalize_mri, dict)
            assert len(self.normalize_mri) == 1
            if "per_scan_norm" in self.normalize_mri:
                if self.normalize_mri["per_scan_norm"] not in ("normalize", "min_max"):
                    raise ValueError('If you want to normalize per scan you have to pass either "normalize" or "min_max"')
            elif "all_scan_norm" in self.normalize_mri:
                assert "mean" in self.normalize_mri["all_scan_norm"].keys()
                assert "std" in self.normalize_mri["all_scan_norm"].keys()
            else:
                raise ValueError('If you use the argument "normalize_mri" only "per_scan_norm" or "all_scan_norm" '
                                 'are allowed as keys!')
        self.quantile = quantile
        self._cache, self._cache_bytes, self.cache_limit_bytes = {}, 0, 0

    def enable_cache(self, gigabytes):
        """Keep decoded volumes (fp32 / uint8, exactly what the decoder produced) in host memory up to `gigabytes`:
        from the second epoch on a cached scan costs one memcpy into the pinned batch slot instead of a gzip inflate
        (14.6 ms per 91x109x91 volume).  The reference re-reads and re-normalises every file every epoch."""
        self.cache_limit_bytes = int(gigabytes * (1 << 30))
        return self

    def __len__(self) -> int:
        return len(self.ds)

    # ------------------------------------------------------------------ per-sample view (raw, staged)
    def _paths(self, index):
        s = self.ds.iloc[index]
        need_mask = bool(self.normalize_mri) and "per_scan_norm" in self.normalize_mri
        return (s.get("path_pet1451"), s.get("path_anat"), s.get("path_anat_mask") if need_mask else None)

    def _tabular(self, index):
        s = self.ds.iloc[index]
        if s.get("AGE") is None:
            return None
        return torch.tensor([float(s[c]) for c in TABULAR_COLUMNS], dtype=torch.float64)

    def label(self, index):
        return self.label_mapping[self.ds.iloc[index]["label"]]

    def __getitem__(self, index: int) -> Dict[str, Any]:
        """RAW sample: 'pet1451' / 'mri' fp32 intensities as stored in the files, 'mri_mask' uint8 (when per-scan
        normalisation needs it), 'tabular' (9,) fp64, 'label' int64 0-d.  Absent modalities are absent keys."""
        pet, mri, mask = self._paths(index)
        data = {}
        if pet is not None:
            data["pet1451"] = staging.read_volume(pet, torch.float32)
        if mri is not None:
            data["mri"] = staging.read_volume(mri, torch.float32)
            if mask is not None:
                data["mri_mask"] = staging.read_volume(mask, torch.uint8)
        tab = self._tabular(index)
        if tab is not None:
            data["tabular"] = tab
        data["label"] = torch.tensor(self.label(index))
        return data

    def get_label_distribution(self):
        """Absolute and normalised class frequencies in label order (dataloader.py:323-343)."""
        order = ["CN", "Dementia"] if self.binary_classification else ["CN", "MCI", "Dementia"]
        counts = self.ds["label"].value_counts().reindex(index=order)
        return torch.tensor(counts.to_numpy(dtype=np.float64)), torch.tensor((counts / len(self.ds)).to_numpy(dtype=np.float64))

    # ------------------------------------------------------------------ batch staging (host) and normalisation (GPU)
    def volume_shape(self):
        for i in range(len(self)):
            pet, mri, _ = self._paths(i)
            p = mri if mri is not None else pet
            if p is not None:
                return staging.read_info(p).shape[:3]
        return None

    def stage(self, indices, buffers, threads=8):
        """Decode the files of `indices` into the pinned batch buffers {'pet1451','mri': fp32 (B,D,H,W), 'mri_mask':
        uint8}.  Returns the set of modalities present (all samples of a dataset share them)."""
        paths = [self._paths(i) for i in indices]
        n = len(indices)
        present = []
        for slot, name in ((0, "pet1451"), (1, "mri"), (2, "mri_mask")):
            col = [p[slot] for p in paths]
            if all(c is None for c in col):
                continue
            if any(c is None for c in col):
                raise ValueError(f"batch mixes samples with and without '{name}' (the reference's collate_fn fails too)")
            dst = buffers[name][:n]
            todo = list(col)
            for i, path in enumerate(col):          # cached scans: one memcpy, no inflate
                hit = self._cache.get((name, path))
                if hit is not None:
                    dst[i].copy_(hit)
                    todo[i] = None
            if any(t is not None for t in todo):
                staging.stage_volumes(todo, dst, threads=threads)
                for i, path in enumerate(todo):
                    if path is None:
                        continue
                    nbytes = dst[i].numel() * dst[i].element_size()
                    if self._cache_bytes + nbytes <= self.cache_limit_bytes and (name, path) not in self._cache:
                        self._cache[(name, path)] = dst[i].clone()
                        self._cache_bytes += nbytes
            present.append(name)
        return present

    def normalize_on_device(self, raw, out_dtype=torch.bfloat16):
        """raw: {'pet1451','mri': fp32 (B,D,H,W) on the GPU, 'mri_mask': uint8} -> normalised volumes
        (dataloader.py:213-215, 239-281) in `out_dtype` (bf16 = encoder input; fp32 = the reference's values)."""
        out = {}
        if "pet1451" in raw:
            x = raw["pet1451"]
            if self.normalize_pet:
                x = norm.normalize_pet(x, self.normalize_pet["mean"], self.normalize_pet["std"], out_dtype=out_dtype)
            elif out_dtype != torch.float32:
                x = norm.K.cast_to_bf16(x)
            out["pet1451"] = x
        if "mri" in raw:
            x = raw["mri"]
            if self.normalize_mri and "per_scan_norm" in self.normalize_mri:
                if self.normalize_mri["per_scan_norm"] == "min_max":
                    x = norm.normalize_mri_per_scan_min_max(x, raw["mri_mask"], self.quantile, out_dtype=out_dtype)
                else:
                    x = norm.normalize_mri_per_scan_zscore(x, raw["mri_mask"], out_dtype=out_dtype)
            elif self.normalize_mri:
                st = self.normalize_mri["all_scan_norm"]
                x = norm.normalize_all_scan(x, float(st["mean"]), float(st["std"]), out_dtype=out_dtype)
            elif out_dtype != torch.float32:
                x = norm.K.cast_to_bf16(x)
            out["mri"] = x
        return out


def per_rank_samples(n, world_size, drop_last=False):
    """Samples every rank sees in one epoch: identical on all ranks, like torch's DistributedSampler (the tail is
    dropped with drop_last, else the order wraps around to pad it)."""
    return n // world_size if drop_last else (n + world_size - 1) // world_size


def epoch_batches(n, batch_size, shuffle=False, drop_last=False, generator=None, rank=0, world_size=1):
    """Index lists of one epoch: a permutation when shuffling (every rank must pass an identically seeded generator),
    strided over data-parallel ranks, cut into batches.  Every rank gets the SAME number of batches of the SAME sizes
    (each step runs collectives - sync-BN sums, the loss normaliser, the gradient buckets - and the BatchNorm element
    count is rows x world size): with world_size > 1 the order is padded by wrapping around (DistributedSampler's rule)
    or, with drop_last, truncated to a multiple of the world size before it is strided."""
    order = torch.randperm(n, generator=generator).tolist() if shuffle else list(range(n))
    if world_size > 1 and n > 0:
        total = per_rank_samples(n, world_size, drop_last) * world_size
        order = order[:total] if total <= n else order + [order[i % n] for i in range(total - n)]
    order = order[rank::world_size]
    out = [order[i:i + batch_size] for i in range(0, len(order), batch_size)]
    if drop_last and out and len(out[-1]) < batch_size:
        out.pop()
    return out


class StagedLoader:
    """Replacement for `DataLoader(dataset, batch_size=..., shuffle=..., num_workers=32)` on this path: yields the
    reference's batch dict with DEVICE tensors ('mri' / 'pet1451' normalised (B,D,H,W) in `out_dtype`, 'tabular'
    (B,9) fp64, 'label' (B,) int64).  Two pinned slots: while the trainer consumes batch i, a helper thread decodes
    batch i+1 (the native decoder releases the GIL) and its H2D copies run on a side stream."""

    def __init__(self, dataset: MultiModalDataset, batch_size: int, shuffle=False, drop_last=False, device="cuda",
                 out_dtype=torch.bfloat16, threads=8, generator=None, rank=0, world_size=1):
        self.ds, self.bs, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("StagedLoader normalises on the GPU: a CUDA device is required (no CPU fallback)")
        self.out_dtype, self.threads, self.generator = out_dtype, threads, generator
        self.rank, self.world = rank, world_size
        shape = dataset.volume_shape()
        self._slots = []
        if shape is not None:
            for _ in range(2):
                self._slots.append({"pet1451": torch.empty((self.bs,) + shape, dtype=torch.float32, pin_memory=True),
                                    "mri": torch.empty((self.bs,) + shape, dtype=torch.float32, pin_memory=True),
                                    "mri_mask": torch.empty((self.bs,) + shape, dtype=torch.uint8, pin_memory=True)})
        self._copy_stream = torch.cuda.Stream(device=self.device)

    def _batches(self):
        return epoch_batches(len(self.ds), self.bs, self.shuffle, self.drop_last, self.generator, self.rank, self.world)

    def __len__(self):
        n = per_rank_samples(len(self.ds), self.world, self.drop_last)   # no permutation drawn: len() must not advance the RNG
        return n // self.bs if self.drop_last else (n + self.bs - 1) // self.bs

    def _decode(self, idx, slot, box):
        try:
            box["present"] = self.ds.stage(idx, self._slots[slot], self.threads) if self._slots else []
        except BaseException as e:  # noqa: BLE001 - re-raised on the consumer thread
            box["error"] = e

    def __iter__(self):
        batches = self._batches()
        if not batches:
            return
        pending = None

        def launch(i):
            box = {}
            t = threading.Thread(target=self._decode, args=(batches[i], i % 2, box), daemon=True)
            t.start()
            return t, box

        pending = launch(0)
        done_events = [None, None]       # the H2D copies out of a slot must finish before it is decoded into again
        for i, idx in enumerate(batches):
            t, box = pending
            t.join()
            if "error" in box:
                raise box["error"]
            slot, n = i % 2, len(idx)
            raw = {}
            with torch.cuda.stream(self._copy_stream):
                for name in box["present"]:
                    raw[name] = self._slots[slot][name][:n].to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            done_events[slot] = ev
            if i + 1 < len(batches):
                nxt = (i + 1) % 2
                if done_events[nxt] is not None:
                    done_events[nxt].synchronize()
                pending = launch(i + 1)
            torch.cuda.current_stream(self.device).wait_event(ev)
            for v in raw.values():
                v.record_stream(torch.cuda.current_stream(self.device))
            batch = self.ds.normalize_on_device(raw, self.out_dtype)
            tabs = [self.ds._tabular(j) for j in idx]
            if all(tb is not None for tb in tabs):
                batch["tabular"] = torch.stack(tabs).to(self.device, non_blocking=True)
            batch["label"] = torch.tensor([self.ds.label(j) for j in idx], dtype=torch.int64).to(self.device)
            yield batch

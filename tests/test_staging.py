"""Input staging: NIfTI decode (C-ABI include/adni_staging.h) against the numpy restatement of
nibabel (oracle/nifti.py), the dataset index against the reference's own MultiModalDataset
(tests/golden/dataset.json, tools/make_golden_dataset.py), and - on the GPU - the batches of StagedLoader against
the reference's normalised samples."""
import base64
import ctypes
import json
import os

import numpy as np
import pytest
import torch

from oracle import nifti as N
from tests._dataset import SHAPE, make_synthetic_adni

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dataset.json")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    with open(GOLD) as f:
        return json.load(f)


def _staging():
    from multimodal_alzheimer_b200 import _build, staging
    _build.build_stage()
    return staging


@pytest.fixture(scope="module")
def adni_csv(tmp_path_factory):
    return make_synthetic_adni(str(tmp_path_factory.mktemp("adni")), seed=15)


def test_staging_library_exports_every_declared_symbol():
    S = _staging()
    lib = S.load()
    with open(os.path.join(ROOT, "include", "adni_staging.h")) as f:
        header = f.read()
    import re
    declared = set(re.findall(r"\b(adni_[a-z0-9_]+)\s*\(", header))
    assert declared == set(S.SYMBOLS), declared ^ set(S.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.adni_stage_version() == 1
    assert ctypes.sizeof(S.NiftiInfo) == 4 + 4 + 56 + 4 * 4 + 8 * 4     # ndim(+pad) dim[7] 4 x i32 2 x f64 2 x i64


CASES = [
    # dtype, scl_slope, scl_inter, big endian, gzip, trailing singleton axes, vox_offset
    (np.float32, float("nan"), float("nan"), False, True, 0, 352),
    (np.int16, 0.0125, -3.5, False, True, 1, 352),
    (np.uint8, 0.0, 0.0, True, False, 0, 352),
    (np.float64, 1.0, 0.0, True, True, 0, 400),
    (np.uint16, 2.0, 0.0, False, True, 2, 352),
    (np.int32, 1.0, 7.0, True, True, 0, 352),
    (np.int8, float("inf"), 1.0, False, True, 0, 352),
    (np.int64, 0.5, 0.25, False, False, 0, 352),
    (np.uint32, 1.0, 0.0, True, True, 0, 352),
    (np.uint64, 1.0, 0.0, False, True, 0, 352),
]


@pytest.mark.parametrize("case", CASES, ids=[np.dtype(c[0]).name for c in CASES])
def test_reader_matches_get_fdata_restatement(tmp_path, case):
    """Every NIfTI-1 datatype, both byte orders, gz and plain, scaled and unscaled: fp64 output bit-identical to
    oracle.nifti.read_fdata (= nib.load(p).get_fdata()), fp32 output = its fp32 rounding, memory order of
    torch.tensor(get_fdata())."""
    S = _staging()
    dt, slope, inter, be, gz, pad, vo = case
    rng = np.random.default_rng(7)
    shape = (7, 9, 5)
    if np.issubdtype(dt, np.floating):
        a = (rng.standard_normal(shape) * 300).astype(dt)
    else:
        ii = np.iinfo(dt)
        a = rng.integers(max(ii.min, -2 ** 40), min(ii.max, 2 ** 40), shape, dtype=dt, endpoint=True)
    p = str(tmp_path / ("v.nii" + (".gz" if gz else "")))
    N.write_nifti(p, a, slope, inter, big_endian=be, vox_offset=vo, pad_dims=pad)
    ref = torch.tensor(N.read_fdata(p))
    info = S.read_info(p)
    assert info.shape == tuple(ref.shape) and info.nvox == ref.numel() and info.swapped == int(be)
    assert torch.equal(S.read_volume(p, torch.float64).reshape(ref.shape), ref)
    assert torch.equal(S.read_volume(p, torch.float32).reshape(ref.shape), ref.float())
    out = torch.full((ref.numel() + 3,), -7.0)
    S.read_volume(p, out=out)                       # caller-owned (pinned) slot, larger than the volume
    assert torch.equal(out[:ref.numel()].reshape(ref.shape), ref.float()) and bool((out[ref.numel():] == -7).all())


def test_reader_random_shapes_and_ranks(tmp_path):
    """2-D ... 5-D images, extents of 1, extents around the 16-wide transposition tile: C-order of get_fdata()."""
    S = _staging()
    rng = np.random.default_rng(11)
    shapes = [(5, 4), (1, 7, 3), (17, 1, 2), (16, 3, 3), (33, 2, 5), (4, 3, 2, 6), (3, 2, 2, 2, 3), (91, 10, 9)]
    for i, shape in enumerate(shapes):
        a = rng.integers(-30000, 30000, shape, dtype=np.int16)
        p = str(tmp_path / f"r{i}.nii.gz")
        N.write_nifti(p, a, scl_slope=1.5, scl_inter=-2.0, big_endian=bool(i % 2))
        ref = torch.tensor(N.read_fdata(p))
        assert S.read_info(p).shape == shape
        assert torch.equal(S.read_volume(p, torch.float64), ref), shape
        assert torch.equal(S.read_volume(p, torch.float32), ref.float()), shape


def test_reader_errors(tmp_path):
    S = _staging()
    with pytest.raises(FileNotFoundError):
        S.read_info(str(tmp_path / "missing.nii.gz"))
    bad = tmp_path / "bad.nii"
    bad.write_bytes(b"\0" * 400)
    with pytest.raises(S.StagingError):
        S.read_info(str(bad))
    p = str(tmp_path / "v.nii.gz")
    N.write_nifti(p, np.arange(24, dtype=np.int16).reshape(2, 3, 4))
    with pytest.raises(ValueError):
        S.read_volume(p, out=torch.empty(10))       # buffer too small
    trunc = tmp_path / "t.nii"
    N.write_nifti(str(trunc), np.arange(24, dtype=np.int16).reshape(2, 3, 4))
    trunc.write_bytes(trunc.read_bytes()[:380])
    with pytest.raises(S.StagingError):
        S.read_volume(str(trunc))
    m = str(tmp_path / "m.nii.gz")
    N.write_nifti(m, np.array([[[0, 1, 2]]], dtype=np.uint8))
    with pytest.raises(S.StagingError):             # the reference multiplies by the mask: 2 cannot be a uint8 flag
        S.read_volume(m, torch.uint8)
    N.write_nifti(m, np.array([[[0, 1, 1]]], dtype=np.float32))
    assert S.read_volume(m, torch.uint8).flatten().tolist() == [0, 1, 1]
    # a crafted header whose 7 extents multiply past 2^63 (32767^7 ~ 2^105) must be rejected, not wrapped into a
    # small positive voxel count that passes the capacity check
    import struct
    raw = bytearray((tmp_path / "v.nii.gz").read_bytes())
    import gzip
    hdr = bytearray(gzip.decompress(bytes(raw)))
    for dims in ([7] + [32767] * 7, [7, 32767, 32767, 32767, 32767, 2, 2, 2], [4, 16384, 16384, 16384, 8, 1, 1, 1]):
        hdr[40:56] = struct.pack("<8h", *dims)
        ov = tmp_path / "overflow.nii"
        ov.write_bytes(bytes(hdr))
        with pytest.raises(S.StagingError):
            S.read_info(str(ov))
        with pytest.raises((S.StagingError, ValueError)):
            S.read_volume(str(ov), out=torch.empty(64))


def test_parallel_batch_decode(tmp_path):
    S = _staging()
    rng = np.random.default_rng(3)
    paths, refs = [], []
    for i in range(9):
        a = rng.integers(-500, 500, (6, 5, 4), dtype=np.int16)
        p = str(tmp_path / f"b{i}.nii.gz")
        N.write_nifti(p, a, scl_slope=0.25 * (i + 1), scl_inter=float(i))
        paths.append(p)
        refs.append(torch.tensor(N.read_fdata(p)).float())
    for threads in (1, 4, 16):
        out = torch.zeros((9, 6, 5, 4))
        S.stage_volumes(paths, out, threads=threads)
        assert torch.equal(out, torch.stack(refs))
    out = torch.full((3, 6, 5, 4), 5.0)
    S.stage_volumes([paths[0], None, paths[2]], out, threads=2)      # absent modality: slot untouched
    assert torch.equal(out[0], refs[0]) and bool((out[1] == 5).all()) and torch.equal(out[2], refs[2])
    odd = str(tmp_path / "odd.nii.gz")
    N.write_nifti(odd, np.zeros((2, 2, 2), dtype=np.int16))
    with pytest.raises(ValueError):
        S.stage_volumes([paths[0], odd], torch.zeros((2, 6, 5, 4)))  # every file of a batch has the batch's shape


@pytest.mark.parametrize("name", sorted(_golden()["configs"].keys()))
def test_dataset_index_matches_reference(adni_csv, name):
    """Pairing of modalities, MCI removal, row order, label counts and the raw staged samples vs the reference's own
    MultiModalDataset on the same CSV/files (tests/golden/dataset.json)."""
    from multimodal_alzheimer_b200.pkg.utils.dataloader import MultiModalDataset
    g = _golden()
    rec = g["configs"][name]
    ds = MultiModalDataset(adni_csv, binary_classification=rec["binary"], modalities=rec["modalities"],
                           normalize_pet=g["pet_norm"], normalize_mri=rec["normalize_mri"], quantile=rec["quantile"])
    assert len(ds) == rec["len"] and list(ds.ds.columns) == rec["columns"]
    base = lambda p: None if p is None else os.path.basename(p)  # noqa: E731
    index = [[r["ID"], r["label"], base(r["path_pet1451"]), base(r["path_anat"]), base(r["path_anat_mask"]), r["AGE"]]
             for _, r in ds.ds.iterrows()]
    assert index == rec["index"]
    counts, normalized = ds.get_label_distribution()
    want = torch.tensor([float("nan") if c is None else float(c) for c in rec["label_counts"]], dtype=torch.float64)
    assert torch.equal(torch.nan_to_num(counts, nan=-1.0), torch.nan_to_num(want, nan=-1.0))
    assert torch.allclose(torch.nan_to_num(normalized), torch.nan_to_num(want / rec["len"]))
    for i, item in enumerate(rec["samples"]):
        s = ds[i]
        assert sorted(k for k in s.keys() if k != "mri_mask") == item["keys"]
        assert int(s["label"]) == item["label"] and s["label"].dtype == torch.int64
        if "tabular" in item:
            assert s["tabular"].dtype == torch.float64 and s["tabular"].tolist() == item["tabular"]
        row = ds.ds.iloc[i]
        if "mri" in s:                                               # raw staged intensities == get_fdata()
            assert torch.equal(s["mri"], torch.tensor(N.read_fdata(row["path_anat"])).float())
            assert tuple(s["mri"].shape) == tuple(item["mri_shape"]) == SHAPE
        if "mri_mask" in s:
            assert torch.equal(s["mri_mask"], torch.tensor(N.read_fdata(row["path_anat_mask"]) != 0).to(torch.uint8))
        if "pet1451" in s:
            assert torch.equal(s["pet1451"], torch.tensor(N.read_fdata(row["path_pet1451"])).float())


def test_pairing_matches_reference_on_other_cohorts(tmp_path):
    """Modality pairing (dataloader.py:100-158, 346-434) on two more synthetic cohorts vs the reference's own class."""
    from multimodal_alzheimer_b200.pkg.utils.dataloader import MultiModalDataset
    g = _golden()
    base = lambda p: None if p is None else os.path.basename(p)  # noqa: E731
    for seed, rec in g["extra_cohorts"].items():
        csv = make_synthetic_adni(str(tmp_path / f"c{seed}"), seed=int(seed), subjects=rec["subjects"])
        for name, want in rec["index"].items():
            cfg = g["configs"][name]
            ds = MultiModalDataset(csv, binary_classification=cfg["binary"], modalities=cfg["modalities"],
                                   normalize_pet=g["pet_norm"], normalize_mri=cfg["normalize_mri"], quantile=cfg["quantile"])
            got = [[r["ID"], r["label"], base(r["path_pet1451"]), base(r["path_anat"]), base(r["path_anat_mask"]), r["AGE"]]
                   for _, r in ds.ds.iterrows()]
            assert got == want, (seed, name)
            assert len(want) > 0


def test_dataset_argument_errors(adni_csv):
    from multimodal_alzheimer_b200.pkg.utils.dataloader import MultiModalDataset
    with pytest.raises(AssertionError):
        MultiModalDataset(adni_csv, modalities=["pet1451", "pet1451"])
    with pytest.raises(AssertionError):
        MultiModalDataset(adni_csv, modalities=["ct"])
    with pytest.raises(ValueError):                                  # dataloader.py:272
        MultiModalDataset(adni_csv, modalities=["t1w"], normalize_mri={"per_scan_norm": "robust"})
    with pytest.raises(ValueError):                                  # dataloader.py:281
        MultiModalDataset(adni_csv, modalities=["t1w"], normalize_mri={"scanner_norm": 1})
    with pytest.raises(AssertionError):                              # dataloader.py:168
        MultiModalDataset(adni_csv, modalities=["pet1451"], normalize_pet={"mean": 1, "std": 2.0})
    with pytest.raises(NotImplementedError):
        MultiModalDataset(adni_csv, modalities=["pet1451"], transform_pet=lambda x: x)


def test_stage_batch_and_decoded_cache(adni_csv):
    """MultiModalDataset.stage: batch buffers == per-sample reads; with the cache on, the second pass is served from
    memory (identical bytes, no file access: the files are made unreadable in between)."""
    from multimodal_alzheimer_b200.pkg.utils.dataloader import MultiModalDataset
    ds = MultiModalDataset(adni_csv, modalities=["pet1451", "t1w"], normalize_mri={"per_scan_norm": "min_max"}, quantile=0.98)
    n = min(5, len(ds))
    idx = list(range(n))
    mk = lambda: {"pet1451": torch.zeros((n,) + SHAPE), "mri": torch.zeros((n,) + SHAPE),   # noqa: E731
                  "mri_mask": torch.zeros((n,) + SHAPE, dtype=torch.uint8)}
    a = mk()
    assert ds.stage(idx, a, threads=3) == ["pet1451", "mri", "mri_mask"]
    for i in idx:
        s = ds[i]
        assert torch.equal(a["pet1451"][i], s["pet1451"]) and torch.equal(a["mri"][i], s["mri"])
        assert torch.equal(a["mri_mask"][i], s["mri_mask"])
    ds.enable_cache(0.5)
    b = mk()
    ds.stage(idx, b, threads=2)                     # fills the cache
    files = {c: {ds.ds.iloc[i][c] for i in idx} for c in ("path_pet1451", "path_anat", "path_anat_mask")}
    vox = int(np.prod(SHAPE))                       # a scan paired with two partners is cached once
    assert len(ds._cache) == sum(len(v) for v in files.values())
    assert ds._cache_bytes == vox * (4 * len(files["path_pet1451"]) + 4 * len(files["path_anat"]) + len(files["path_anat_mask"]))
    moved = []
    for pth in set().union(*files.values()):        # hide the files: a cache hit must not touch them
        os.rename(pth, pth + ".hidden")
        moved.append(pth)
    try:
        c = mk()
        ds.stage(idx, c, threads=2)
    finally:
        for pth in moved:
            os.rename(pth + ".hidden", pth)
    for k in a:
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k])
    ds2 = MultiModalDataset(adni_csv, modalities=["t1w"], normalize_mri={"per_scan_norm": "min_max"}).enable_cache(1e-5)
    ds2.stage([0, 1], {"mri": torch.zeros((2,) + SHAPE), "mri_mask": torch.zeros((2,) + SHAPE, dtype=torch.uint8)})
    assert ds2._cache_bytes <= ds2.cache_limit_bytes      # budget respected (about one volume fits)


def test_epoch_batches_cover_the_dataset_once_across_ranks():
    from multimodal_alzheimer_b200.pkg.utils.dataloader import StagedLoader, epoch_batches
    assert epoch_batches(10, 4) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    assert epoch_batches(10, 4, drop_last=True) == [[0, 1, 2, 3], [4, 5, 6, 7]]
    seen = []
    for rank in range(3):
        g = torch.Generator().manual_seed(15)
        seen += [i for b in epoch_batches(11, 2, shuffle=True, generator=g, rank=rank, world_size=3) for i in b]
    assert sorted(set(seen)) == list(range(11)) and len(seen) == 12     # padded by wrapping, like DistributedSampler
    # every rank must run the same number of steps with the same batch sizes: each step has collectives
    for n, world, bs in ((119, 2, 10), (119, 2, 7), (10, 4, 3), (5, 8, 2), (64, 8, 4)):
        for drop_last in (False, True):
            shapes = [[len(b) for b in epoch_batches(n, bs, drop_last=drop_last, rank=r, world_size=world)]
                      for r in range(world)]
            assert all(s == shapes[0] for s in shapes), (n, world, bs, drop_last, shapes)
            per = (n // world) if drop_last else -(-n // world)
            assert len(shapes[0]) == (per // bs if drop_last else -(-per // bs))
    with pytest.raises(RuntimeError):                 # normalisation runs on the GPU: no CPU fallback
        StagedLoader(None, 2, device="cpu")


def _volume(item, key):
    a = np.frombuffer(base64.b64decode(item[key + "_f32_b64"]), dtype="<f4").reshape(item[key + "_shape"])
    return torch.tensor(a.copy())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pet", "mri_minmax_q97", "mri_zscore", "mri_allscan", "pet_mri", "all"])
def test_staged_loader_batches_match_reference_samples(cuda_dev, adni_csv, name):
    """StagedLoader (native decode -> pinned -> H2D -> CUDA normalisation) vs the samples the reference's own
    `__getitem__` produced (fp64 CPU arithmetic, cast to fp32 like anat_cnn.py:103).  fp32 output: rel-L2 <= 1e-6
    (the min-max path is designed bit-exact; printed); bf16 output = the fp32 result rounded."""
    from multimodal_alzheimer_b200.pkg.utils.dataloader import MultiModalDataset, StagedLoader
    from tests._util import rel_l2
    g = _golden()
    rec = g["configs"][name]
    ds = MultiModalDataset(adni_csv, binary_classification=rec["binary"], modalities=rec["modalities"],
                           normalize_pet=g["pet_norm"], normalize_mri=rec["normalize_mri"], quantile=rec["quantile"])
    bs = 4
    loader = StagedLoader(ds, batch_size=bs, device=cuda_dev, out_dtype=torch.float32, threads=4)
    assert len(loader) == (len(ds) + bs - 1) // bs
    batches = list(loader)
    assert sum(int(b["label"].numel()) for b in batches) == len(ds)
    labels = torch.cat([b["label"].cpu() for b in batches]).tolist()
    assert labels == [ds.label(i) for i in range(len(ds))]
    first = batches[0]
    for i, item in enumerate(rec["samples"]):
        for key in ("mri", "pet1451"):
            if key + "_shape" in item:
                want = _volume(item, key)
                got = first[key][i].cpu()
                assert got.dtype == torch.float32 and got.shape == want.shape
                e = rel_l2(got, want)
                print(f"{name} sample {i} {key}: rel-L2 {e:.2e}, bit-identical {torch.equal(got, want)}")
                assert e <= 1e-6
        if "tabular" in item:
            assert first["tabular"][i].cpu().tolist() == item["tabular"]
    if name == "pet_mri":                                            # the encoder's input type
        lb = next(iter(StagedLoader(ds, batch_size=bs, device=cuda_dev, out_dtype=torch.bfloat16, threads=2)))
        for key in ("mri", "pet1451"):
            assert lb[key].dtype == torch.bfloat16
            assert rel_l2(lb[key].float().cpu(), first[key].cpu()) <= 4e-3


@pytest.mark.gpu
@pytest.mark.parametrize("modality,mods", [("mri", ["t1w"]), ("pet1451", ["pet1451"])])
def test_split_statistics_match_reference_formula(cuda_dev, adni_csv, modality, mods):
    """NormalizeDataset.compute_std_mean (staged decode + adni_scan_moments) vs the reference's loop
    (standardization.py:41-53) evaluated in fp64 on the CPU over the same files: 1e-12 relative."""
    from multimodal_alzheimer_b200.pkg.utils.dataloader import MultiModalDataset
    from multimodal_alzheimer_b200.pkg.utils.standardization import NormalizeDataset
    ds = MultiModalDataset(adni_csv, modalities=mods, normalize_mri=None, normalize_pet=None)
    mean, std = NormalizeDataset.compute_std_mean(ds, modality=modality, device=cuda_dev, batch_size=4, threads=3)
    col = "path_anat" if modality == "mri" else "path_pet1451"
    mean_x, mean_x2 = 0.0, 0.0
    for i in range(len(ds)):
        x = torch.tensor(N.read_fdata(ds.ds.iloc[i][col]))
        mean_x += x.mean()
        mean_x2 += (x ** 2).mean()
    ref_mean = mean_x / len(ds)
    ref_std = torch.sqrt(mean_x2 / len(ds) - ref_mean ** 2)
    assert abs(float(mean) - float(ref_mean)) <= 1e-12 * abs(float(ref_mean)), (float(mean), float(ref_mean))
    assert abs(float(std) - float(ref_std)) <= 1e-10 * abs(float(ref_std)), (float(std), float(ref_std))
    with pytest.raises(NotImplementedError):
        NormalizeDataset.compute_std_mean(ds, dim=0, modality=modality, device=cuda_dev)

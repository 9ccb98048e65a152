"""Parity at the sizes BASELINE.json quotes (VERDICT r01 'parity holes at the BASELINE sizes').

* The bench workloads themselves (multimodal_alzheimer_b200/workloads.py - the object bench.py times): raw synthetic
  volumes -> GPU quantile / standardise kernels -> LightningModule general_step -> backward, against the CPU oracle fed
  by the reference's own fp64 normalisation sequence (dataloader.py:245-270, 213-215), 128^3 volumes, every logit,
  the loss, every parameter gradient and every BatchNorm running statistic.  Tolerances: those of
  tests/test_gpu_models.py (logits / loss 2e-2, gradients / running statistics 3e-2 rel-L2, each OR within 2x of the
  error of PyTorch's own bf16-autocast run of the oracle on the same inputs).
* Config 5's encoder at full resolution: ResNet-50 on one 160x192x160 volume.
* The quantile order statistics at 128^3 (2.1 M voxels) and 160x192x160 (4.9 M voxels): n, lo, hi, Qmin, Qmax and the
  normalised fp32 volume BIT-exact against torch.quantile in fp64 (north_star: "order-statistic indices bit-exact").
"""
import copy
import types

import pytest
import torch

from multimodal_alzheimer_b200 import workloads as W
from tests.test_gpu_models import _compare

pytestmark = pytest.mark.gpu


def _oracle_ns():
    import oracle.models as O
    from tests._models import build_model  # noqa: F401  (keeps the two oracle trunk definitions in one place)

    class ResNet_PET_Trunk(torch.nn.Module):
        def __init__(self, enc):
            super().__init__()
            self.encoder = enc
            self.encoder.model.conv_seg = self.encoder.model.conv_seg[:2]
            self.relu = torch.nn.ReLU()
            self.reduce_dim_pet = torch.nn.Sequential(torch.nn.Linear(512, 64), self.relu)

        def forward(self, x):
            out = self.encoder(x)
            return self.reduce_dim_pet(out.view(out.shape[0], -1))

    return types.SimpleNamespace(Anat_CNN=O.Anat_CNN, PET_CNN_ResNet=O.PET_CNN_ResNet, Small_PET_CNN=O.Small_PET_CNN,
                                 Anat_PET_CNN=O.Anat_PET_CNN, ResNet_PET_Trunk=ResNet_PET_Trunk,
                                 Tabular_MRT_Model=O.Tabular_MRT_Model, PET_TABULAR_CNN=O.PET_TABULAR_CNN,
                                 All_Modalities_Fusion=O.All_Modalities_Fusion, tab_key="tabular_features")


def oracle_batch(raw, tab_key="tabular_features"):
    """The reference DataLoader's CPU sequence on raw volumes (fp64 torch.quantile, Normalize)."""
    from oracle.normalization import pet_standardize_oracle, quantile_minmax_oracle
    batch = {"label": raw["label"]}
    if "mri_raw" in raw:
        batch["mri"] = torch.stack([quantile_minmax_oracle(raw["mri_raw"][i].double(), raw["mask"][i].double(), 0.98)[0]
                                    for i in range(raw["mri_raw"].shape[0])])
    if "pet_raw" in raw:
        batch["pet1451"] = pet_standardize_oracle(raw["pet_raw"].double(), W.PET_MEAN, W.PET_STD)
    if "tabular" in raw:
        batch[tab_key] = raw["tabular"]
    return batch


def _autocast_step(oracle, batch, dev):
    """PyTorch's own bf16-autocast execution of the oracle on the GPU: the measured bf16 noise floor of this input."""
    m = copy.deepcopy(oracle).to(dev).train()
    m.zero_grad(set_to_none=True)
    b = {k: v.to(dev) for k, v in batch.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m.general_step(b, 0, "train")
    out["loss"].backward()
    torch.cuda.synchronize()
    return m, out


def run_workload_parity(name, n_samples, dev, depth=None, volume=None):
    w = W.WORKLOADS[name]
    volume = tuple(volume or w["volume"])
    ns = _oracle_ns()
    oracle = W.build_model(ns, name, depth=depth)
    product = W.build_model(W.product_namespace(), name, depth=depth)
    missing = product.load_state_dict(copy.deepcopy(oracle.state_dict()), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    raw = W.synth_batch(0, n_samples, volume, w["modalities"])
    raw["label"][0], raw["label"][-1] = 0, 2
    ob = oracle_batch(raw, ns.tab_key)
    bracket, out_a = _autocast_step(oracle, ob, dev)
    out_o = oracle.general_step(ob, 0, "train")
    out_o["loss"].backward()
    product.to(dev).train()
    out_p = product.general_step(W.normalized_batch_gpu({k: v.to(dev) for k, v in raw.items()}), 0, "train")
    out_p["loss"].backward()
    torch.cuda.synchronize()
    report = _compare(oracle, product, out_o, out_p, bracket=bracket, out_a=out_a)
    print(f"{name} B={n_samples} {volume}: loss product {float(out_p['loss']):.6f} oracle {float(out_o['loss']):.6f}\n{report}")
    return out_o, out_p


@pytest.mark.parametrize("name,n_samples", [("pet_mri_fusion_r18", 2), ("mri_r18", 4), ("all_modalities", 2),
                                            ("pet_mri_fusion_faithful", 2)])
def test_bench_workload_matches_oracle_at_128(cuda_dev, name, n_samples):
    """configs[2] (the bench line), configs[1], configs[3] and the faithful configs[2] at 1x128^3 through the same
    code path bench.py times (raw inputs -> GPU normalisation -> general_step -> backward)."""
    run_workload_parity(name, n_samples, cuda_dev)


def test_config5_encoder_full_resolution(cuda_dev):
    """configs[4]: ResNet-50 Bottleneck encoder on a 160x192x160 volume (1x1x1 GEMMs up to 1024 -> 2048, dilated 3^3
    convs at 20x24x20), one sample (the CPU oracle holds ~10 GB of fp32 activations for it)."""
    run_workload_parity("mri_r50_160", 1, cuda_dev)


@pytest.mark.parametrize("shape", [(128, 128, 128), (160, 192, 160)])
@pytest.mark.parametrize("q", [0.98, 0.95])
def test_quantile_bit_exact_at_baseline_sizes(cuda_dev, shape, q):
    from multimodal_alzheimer_b200 import kernels as K
    from oracle.normalization import quantile_minmax_oracle
    raw = W.synth_batch(3, 2, shape, ("mri",))
    out, info, qv = K.quantile_minmax_normalize(raw["mri_raw"].to(cuda_dev), raw["mask"].to(cuda_dev), q, want_info=True)
    out, info, qv = out.cpu(), info.cpu(), qv.cpu()
    for s in range(2):
        ref, meta = quantile_minmax_oracle(raw["mri_raw"][s].double(), raw["mask"][s].double(), q)
        got = dict(n=int(info[s, 0]), lo_max=int(info[s, 1]), hi_max=int(info[s, 2]), lo_min=int(info[s, 3]),
                   hi_min=int(info[s, 4]))
        want = {k: meta[k] for k in got}
        assert got == want, (got, want)
        assert float(qv[s, 0]) == meta["qmax"] and float(qv[s, 1]) == meta["qmin"], (qv[s].tolist(), meta)
        assert torch.equal(out[s], ref.float()), f"normalised volume differs at {shape}, q={q}"

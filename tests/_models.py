"""Builders shared by the model-level parity tests: the same hparams build the oracle (CPU, torch.nn) and the
product (CUDA kernels); weights are copied oracle -> product through state_dict (identical keys)."""
import copy

import torch

CW3 = torch.tensor([0.4651162790697675, 0.6712473572938689, 0.8636363636363636], dtype=torch.float64)  # test_tab.py:36-40


def hp_anat(depth=10, n_classes=3, fl_gamma=None, bn_begin=False, bn_dense=False, linear_out=()):
    return dict(n_classes=n_classes, resnet_depth=depth, batchnorm_begin=bn_begin, batchnorm_dense=bn_dense,
                linear_out=list(linear_out), fl_gamma=fl_gamma, loss_class_weights=CW3[:n_classes].clone(), lr=1e-3,
                lr_pretrained=1e-4, l2_reg=0.0, reduce_factor_lr_schedule=None, norm_percentile=0.98)


def hp_pet(n_classes=3, batchnorm=False, conv_out=(8, 16, 32, 64), filter_size=(5, 5, 3, 3), linear_out=64):
    return dict(n_classes=n_classes, conv_out=list(conv_out), filter_size=list(filter_size), linear_out=linear_out,
                batchnorm=batchnorm, loss_class_weights=CW3[:n_classes].clone(), lr=1e-3,
                reduce_factor_lr_schedule=None)


def hp_fmf(n_classes=3, fusion_mode="maxout", batchnorm=True, batchnorm_fusion=True, filter_size_fusion=3, n_out_fusion=64,
           conv_out=(8, 16, 32, 64), filter_size=(5, 5, 3, 3)):
    # train_anat_pet_featuremapfusion.py:60-117 (the HPO search space)
    return dict(n_classes=n_classes, conv_out=list(conv_out), filter_size=list(filter_size), batchnorm=batchnorm,
                fusion_mode=fusion_mode, n_layers_fusion=1, filter_size_fusion=filter_size_fusion,
                n_out_fusion=n_out_fusion, batchnorm_fusion=batchnorm_fusion, loss_class_weights=CW3[:n_classes].clone(),
                lr=1e-3, l2_reg=1e-4, reduce_factor_lr_schedule=None)


def hp_fusion(n_classes=3, fl_gamma=1, simple_dim_red=False):
    return dict(n_classes=n_classes, fl_gamma=fl_gamma, loss_class_weights=CW3[:n_classes].clone(), lr=1e-3,
                lr_pretrained=1e-4, l2_reg=0.0, reduce_factor_lr_schedule=None, simple_dim_red=simple_dim_red,
                ensemble_size=4)


# Module-level parity cases, shared by tests/test_gpu_models.py (CUDA vs oracle), tests/test_oracle.py (oracle vs the
# golden outputs of the unmodified reference classes) and tools/make_golden_models.py (which writes those outputs).
CASES = [
    # kind, kwargs, batch, volume shape, modalities
    ("anat", dict(depth=10), 2, (64, 64, 64), ("mri",)),                                   # config 1 (reduced size)
    ("anat", dict(depth=18, bn_begin=True, bn_dense=True, linear_out=(32,), fl_gamma=2), 3, (48, 56, 48), ("mri",)),
    ("anat", dict(depth=18, bn_begin=True, bn_dense=True, linear_out=(32,), fl_gamma=2), 6, (40, 40, 40), ("mri",)),
    ("anat", dict(depth=50, fl_gamma=1), 2, (40, 48, 40), ("mri",)),                       # Bottleneck path
    ("pet_resnet", dict(depth=10, n_classes=2), 2, (48, 48, 48), ("pet1451",)),
    ("small_pet", dict(), 2, (32, 32, 32), ("pet1451",)),
    ("small_pet", dict(pet_batchnorm=True, n_classes=2), 3, (32, 40, 32), ("pet1451",)),
    ("anat_pet", dict(depth=10), 2, (48, 48, 48), ("mri", "pet1451")),                     # faithful config 3
    ("anat_pet_2resnet", dict(depth=10), 2, (48, 48, 48), ("mri", "pet1451")),             # north-star config 3
    ("mri_tab", dict(depth=10), 2, (48, 48, 48), ("mri", "tabular")),
    ("pet_tab", dict(simple_dim_red=True), 2, (32, 32, 32), ("pet1451", "tabular")),
    ("all", dict(depth=10), 2, (48, 48, 48), ("mri", "pet1451", "tabular")),               # config 4
    # early fusion (2-channel input) and feature-map fusion (maxout / concatenate)
    ("early_fusion", dict(), 2, (32, 32, 32), ("mri", "pet1451")),
    ("early_fusion", dict(pet_batchnorm=True, n_classes=2), 4, (32, 40, 32), ("mri", "pet1451")),
    ("fmf", dict(fusion_mode="maxout"), 4, (64, 64, 64), ("mri", "pet1451")),
    ("fmf", dict(fusion_mode="concatenate", batchnorm_fusion=False, pet_batchnorm=False, n_out_fusion=128), 2,
     (48, 64, 48), ("mri", "pet1451")),
    ("fmf", dict(fusion_mode="maxout", filter_size_fusion=5, n_classes=2), 3, (32, 32, 32), ("mri", "pet1451")),
    ("fmf", dict(fusion_mode="concatenate", filter_size_fusion=4), 4, (48, 48, 64), ("mri", "pet1451")),   # even kernel
    # better-conditioned twins of the tiny-batch cases above (the per-tensor rule, not the widened one,
    # should carry the fusion and Bottleneck paths): batch 6 keeps BatchNorm's backward away from the catastrophic
    # cancellation a batch of 2 produces in bf16 for torch's own autocast as well
    ("anat_pet_2resnet", dict(depth=10), 6, (48, 48, 48), ("mri", "pet1451")),
    ("all", dict(depth=10), 6, (48, 48, 48), ("mri", "pet1451", "tabular")),
    ("anat", dict(depth=50, fl_gamma=1), 6, (40, 48, 40), ("mri",)),
    ("pet_resnet", dict(depth=18), 6, (48, 48, 48), ("pet1451",)),
]
CASE_IDS = [f"{c[0]}-{i}" for i, c in enumerate(CASES)]

# Short training runs (training_step -> backward -> configure_optimizers().step(), a fresh synthetic batch per step):
# the reference's own classes and optimizer in tests/golden/models.json["trajectories"], the oracle on the CPU, the
# product on the GPU.  kind, kwargs, batch, volume shape, modalities, steps
TRAJECTORIES = {
    "traj-anat": ("anat", dict(depth=10, bn_begin=True, linear_out=(16,)), 4, (32, 32, 32), ("mri",), 4),
    "traj-anat_pet": ("anat_pet", dict(depth=10, fl_gamma=1), 4, (32, 32, 32), ("mri", "pet1451"), 4),
    "traj-small_pet": ("small_pet", dict(pet_batchnorm=True), 4, (32, 32, 32), ("pet1451",), 4),
}


def trajectory_batches(traj_id):
    kind, kw, B, shape, mods, steps = TRAJECTORIES[traj_id]
    return [synthetic_batch(B, shape, kw.get("n_classes", 3), seed=15 + k, modalities=mods) for k in range(steps)]


def golden_record(case_id):
    """The record tools/make_golden_models.py wrote for this case from the reference's own classes, or None (the
    two-ResNet north-star variant is not a reference configuration and has no record)."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models.json")
    if not os.path.exists(path):
        return None
    global _GOLDEN
    if _GOLDEN is None:
        with open(path) as f:
            _GOLDEN = json.load(f)
    return _GOLDEN["cases"].get(case_id)


_GOLDEN = None


def compare_with_golden(rec, out_o, out_p, tol_logits, tol_loss):
    """Product outputs vs the logits / loss the reference's own classes produced for this case (tests/golden/
    models.json).  `out_o` = the oracle run on this machine's CPU, which tests/test_oracle.py holds to the same record;
    its distance d to the record (CPU conv re-association on a different core count) is added to the tolerance, so this
    is the direct form of |product - oracle| <= tol and |oracle - reference| = d.  Returns (error, d)."""
    lg = torch.tensor(rec["outputs"], dtype=torch.float64)
    lo = out_o["outputs"].detach().double().cpu()
    lp = out_p["outputs"].detach().double().cpu()
    assert lp.shape == lg.shape, (tuple(lp.shape), tuple(lg.shape))
    nrm = float(lg.norm().clamp_min(1e-30))
    d = float((lo - lg).norm()) / nrm
    e = float((lp - lg).norm()) / nrm
    assert e <= tol_logits * float(lo.norm()) / nrm + d, f"logits vs reference record: {e:.3e} (oracle {d:.3e})"
    dl = abs(float(out_o["loss"].detach()) - rec["loss"])
    el = abs(float(out_p["loss"].detach()) - rec["loss"])
    assert el <= tol_loss + dl, f"loss vs reference record: {el:.3e} (oracle {dl:.3e})"
    return e, d


def probe(name, numel):
    """Fixed pseudo-random probe vector for a tensor called `name` (values in [-1, 1), fp64)."""
    seed = sum((i + 1) * b for i, b in enumerate(name.encode())) % (2 ** 31 - 1)
    g = torch.Generator().manual_seed(seed)
    return torch.rand(numel, generator=g, dtype=torch.float64) * 2 - 1


def fingerprint(name, t):
    """[||t||, <t, 1>/||1||, <t, probe>/||probe||]: two projections on unit vectors, so that a perturbation d of the
    tensor moves every entry by at most ||d||."""
    t = t.detach().double().flatten()
    pr = probe(name, t.numel())
    return [float(t.norm()), float(t.sum()) / t.numel() ** 0.5, float(t @ pr) / float(pr.norm())]


def build_model(kind, oracle, **kw):
    """One model of `kind` from the oracle (oracle=True; CPU torch.nn) or the product (CUDA kernels). kind in
    {'anat','pet_resnet','small_pet','anat_pet','anat_pet_2resnet','mri_tab','pet_tab','all','early_fusion','fmf'}.
    Weight initialisation consumes the global torch RNG: seed it first."""
    import oracle.models as O
    if not oracle:
        from multimodal_alzheimer_b200.pkg.models.fusion_models import anat_pet_featuremapfusion as P_fmf
        from multimodal_alzheimer_b200.pkg.models.fusion_models import early_fusion as P_ef
        from multimodal_alzheimer_b200.pkg.models.fusion_models import all_modalities_fusion as P_all
        from multimodal_alzheimer_b200.pkg.models.fusion_models import anat_pet_fusion as P_ap
        from multimodal_alzheimer_b200.pkg.models.fusion_models import pet_tabular_fusion as P_pt
        from multimodal_alzheimer_b200.pkg.models.fusion_models import tabular_mri_fusion as P_mt
        from multimodal_alzheimer_b200.pkg.models.mri_models.anat_cnn import Anat_CNN
        from multimodal_alzheimer_b200.pkg.models.pet_models.pet_cnn import Small_PET_CNN
        from multimodal_alzheimer_b200.pkg.models.pet_models.pet_resnet_cnn import PET_CNN_ResNet
    else:
        P_fmf = P_ef = P_all = P_ap = P_pt = P_mt = Anat_CNN = Small_PET_CNN = PET_CNN_ResNet = None

    depth = kw.get("depth", 10)
    nc = kw.get("n_classes", 3)

    def _thaw_or_freeze(hp):
        if kw.get("frozen"):            # lower stages frozen (train_anat_pet_fusion.py:99, anat_pet_fusion.py:35-40)
            hp["lr_pretrained"] = None
        return hp

    class OracleResNetPETTrunk(torch.nn.Module):
        def __init__(self, enc):
            super().__init__()
            self.encoder = enc
            self.encoder.model.conv_seg = self.encoder.model.conv_seg[:2]
            self.relu = torch.nn.ReLU()
            self.reduce_dim_pet = torch.nn.Sequential(torch.nn.Linear(512, 64), self.relu)

        def forward(self, x):
            out = self.encoder(x)
            return self.reduce_dim_pet(out.view(out.shape[0], -1))

    def anat(o):
        hp = _thaw_or_freeze(hp_anat(depth, nc, kw.get("fl_gamma"), kw.get("bn_begin", False), kw.get("bn_dense", False),
                                     kw.get("linear_out", ())))
        return (O.Anat_CNN if o else Anat_CNN)(hp)

    def petres(o):
        hp = _thaw_or_freeze(hp_anat(depth, nc, kw.get("fl_gamma"), kw.get("bn_begin", False), kw.get("bn_dense", False),
                                     kw.get("linear_out", ())))
        return (O.PET_CNN_ResNet if o else PET_CNN_ResNet)(hp)

    def smallpet(o):
        hp = hp_pet(nc, kw.get("pet_batchnorm", False), kw.get("conv_out", (8, 16, 32, 64)),
                    kw.get("filter_size", (5, 5, 3, 3)))
        return (O.Small_PET_CNN if o else Small_PET_CNN)(hp)

    def anat_pet(o, two_resnet):
        hp = _thaw_or_freeze(hp_fusion(nc, kw.get("fl_gamma", 1)))
        if two_resnet:
            trunk = (OracleResNetPETTrunk if o else P_ap.ResNet_PET_Trunk)(petres(o))
            return (O.Anat_PET_CNN if o else P_ap.Anat_PET_CNN)(hp, model_mri=anat(o), pet_trunk=trunk)
        return (O.Anat_PET_CNN if o else P_ap.Anat_PET_CNN)(hp, model_pet=smallpet(o), model_mri=anat(o))

    def mri_tab(o):
        return (O.Tabular_MRT_Model if o else P_mt.Tabular_MRT_Model)(_thaw_or_freeze(hp_fusion(nc, kw.get("fl_gamma", 1))),
                                                                      model_mri=anat(o))

    def pet_tab(o):
        return (O.PET_TABULAR_CNN if o else P_pt.PET_TABULAR_CNN)(
            _thaw_or_freeze(hp_fusion(nc, kw.get("fl_gamma", 1), kw.get("simple_dim_red", False))), model_pet=smallpet(o))

    def build(o):
        if kind == "anat":
            return anat(o)
        if kind == "pet_resnet":
            return petres(o)
        if kind == "small_pet":
            return smallpet(o)
        if kind == "anat_pet":
            return anat_pet(o, False)
        if kind == "anat_pet_2resnet":
            return anat_pet(o, True)
        if kind == "mri_tab":
            return mri_tab(o)
        if kind == "pet_tab":
            return pet_tab(o)
        if kind == "early_fusion":
            hp = hp_pet(nc, kw.get("pet_batchnorm", False), kw.get("conv_out", (8, 16, 32, 64)),
                        kw.get("filter_size", (5, 5, 3, 3)))
            return (O.PET_MRI_EF if o else P_ef.PET_MRI_EF)(hp)
        if kind == "fmf":
            hp = hp_fmf(nc, kw.get("fusion_mode", "maxout"), kw.get("pet_batchnorm", True), kw.get("batchnorm_fusion", True),
                        kw.get("filter_size_fusion", 3), kw.get("n_out_fusion", 64))
            return (O.PET_MRI_FMF if o else P_fmf.PET_MRI_FMF)(hp)
        if kind == "all":
            hp = _thaw_or_freeze(hp_fusion(nc, kw.get("fl_gamma", 1)))
            cls = O.All_Modalities_Fusion if o else P_all.All_Modalities_Fusion
            return cls(hp, model_anat_pet=anat_pet(o, False), model_anat_tab=mri_tab(o), model_pet_tab=pet_tab(o))
        raise ValueError(kind)

    return build(oracle)


def build_oracle(kind, seed=15, **kw):
    torch.manual_seed(seed)
    return build_model(kind, True, **kw)


def build_pair(kind, seed=15, **kw):
    """Returns (oracle_model_cpu, product_model) with identical weights."""
    oracle = build_oracle(kind, seed, **kw)
    product = build_model(kind, False, **kw)
    missing = product.load_state_dict(copy.deepcopy(oracle.state_dict()), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return oracle, product


def synthetic_batch(B, shape, n_classes=3, seed=15, modalities=("mri",)):
    g = torch.Generator().manual_seed(seed)
    batch = {}
    if "mri" in modalities:
        batch["mri"] = torch.rand((B,) + tuple(shape), generator=g, dtype=torch.float64)
    if "pet1451" in modalities:
        pet = torch.randn((B,) + tuple(shape), generator=g, dtype=torch.float64) * 0.5383 + 0.5145
        batch["pet1451"] = (pet.clamp_min(0) - 0.5145) / 0.5383
    if "tabular" in modalities:
        batch["tabular"] = torch.randn((B, 1024), generator=g, dtype=torch.float32)
    batch["label"] = torch.randint(0, n_classes, (B,), generator=g)
    batch["label"][0] = 0
    batch["label"][-1] = n_classes - 1
    return batch


def oracle_step(model, batch):
    """One fwd+bwd of the oracle on the CPU. Works for every model kind (tabular key naming differs)."""
    model.train()
    b = dict(batch)
    if "tabular" in b:
        b["tabular_features"] = b["tabular"]
    if hasattr(model, "general_step"):
        try:
            out = model.general_step(b, 0, "train")
        except (KeyError, AttributeError, TypeError):
            out = None
    else:
        out = None
    if out is None:  # stage-2 tabular models of the oracle expose forward only
        y = b["label"]
        if model.__class__.__name__ == "Tabular_MRT_Model":
            yh = model(b["tabular"], b["mri"].unsqueeze(1).float()).double()
        else:
            yh = model(b["pet1451"].unsqueeze(1).float(), b["tabular"]).double()
        out = {"loss": model.criterion(yh, y), "outputs": yh, "labels": y}
    out["loss"].backward()
    return out


def product_step(model, batch, dev):
    model.to(dev).train()
    b = {k: v.to(dev) for k, v in batch.items()}
    out = model.general_step(b, 0, "train")
    out["loss"].backward()
    torch.cuda.synchronize()
    return out


def autocast_step(model, batch, dev):
    """The oracle itself executed by PyTorch under bf16 autocast on the GPU: the noise floor of bf16 operands with
    fp32 accumulation on this input (SURVEY.md §8c 'bracketed by a torch-autocast-bf16 oracle run')."""
    model = copy.deepcopy(model).to(dev)
    model.zero_grad(set_to_none=True)
    model.train()
    b = {k: v.to(dev) for k, v in batch.items()}
    if "tabular" in b:
        b["tabular_features"] = b["tabular"]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = b["label"]
        name = model.__class__.__name__
        if name == "Tabular_MRT_Model":
            yh = model(b["tabular"], b["mri"].unsqueeze(1).float())
        elif name == "PET_TABULAR_CNN":
            yh = model(b["pet1451"].unsqueeze(1).float(), b["tabular"])
        elif name == "All_Modalities_Fusion":
            yh = model(b["pet1451"].unsqueeze(1).float(), b["mri"].unsqueeze(1).float(), b["tabular"])
        elif name in ("Anat_PET_CNN", "PET_MRI_FMF"):
            yh = model(b["pet1451"].unsqueeze(1).float(), b["mri"].unsqueeze(1).float())
        elif name == "PET_MRI_EF":
            yh = model(torch.stack((b["pet1451"], b["mri"]), dim=1).float())
        elif name in ("Small_PET_CNN", "PET_CNN_ResNet"):
            yh = model(b["pet1451"].unsqueeze(1).float())
        else:
            yh = model(b["mri"].unsqueeze(1).float())
    loss = model.criterion(yh.double(), y)
    loss.backward()
    torch.cuda.synchronize()
    return model, {"loss": loss, "outputs": yh.double(), "labels": y}

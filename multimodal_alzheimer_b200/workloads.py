"""The BASELINE.json configurations as buildable workloads: hparams, model construction, synthetic inputs, conv FLOPs.

One description serves bench.py (both arms), the BASELINE-size parity tests and the multi-rank parity check, so that
"the bench workload" and "the tested workload" are the same object.  Model classes come from a *namespace* the
caller supplies: `product_namespace()` (the CUDA classes of this package) or a namespace built by the caller over the
CPU oracle (tests / bench.py's CPU legs) - this module never imports `oracle/`.

Synthetic data follows SURVEY.md 8(d) and is generated PER GLOBAL SAMPLE INDEX on the CPU generator (seed 15, the
reference's seed, train_anat_cnn.py:165), so every rank count trains on the same global batch and the CPU arm sees the
same samples as the GPU arm.
"""
import types

import torch

CW3 = (0.4651162790697675, 0.6712473572938689, 0.8636363636363636)     # pkg/inference/test_tab.py:36-40
PET_MEAN, PET_STD = 0.5145, 0.5383                                      # pkg/models/pet_models/train_pet_cnn.py:77-78
SEED = 15                                                               # pkg/models/mri_models/train_anat_cnn.py:165

# name -> description.  `global_batch` None = weak scaling with `per_gpu_batch` samples on every rank.
WORKLOADS = {
    "mri_r10": dict(config="BASELINE.json configs[0]", kind="anat", depth=10, volume=(128, 128, 128), global_batch=2,
                    per_gpu_batch=None, modalities=("mri",), plain_head=True,
                    title="MedicalNet ResNet-10 3D MRI-only 3-class classifier, weighted CE"),
    "mri_r18": dict(config="BASELINE.json configs[1]", kind="anat", depth=18, volume=(128, 128, 128), global_batch=16,
                    per_gpu_batch=None, modalities=("mri",), plain_head=False,
                    title="ResNet-18 3D MRI classifier (BatchNorm begin/dense), weighted CE"),
    "pet_mri_fusion_r18": dict(config="BASELINE.json configs[2]", kind="fusion2", depth=18, volume=(128, 128, 128),
                               global_batch=32, per_gpu_batch=None, modalities=("mri", "pet"), plain_head=False,
                               title="PET-MRI two-branch ResNet-18 3D fusion (feature concat -> MLP head), focal loss gamma=1"),
    "pet_mri_fusion_faithful": dict(config="BASELINE.json configs[2], reference-faithful PET branch (Small_PET_CNN)",
                                    kind="fusion_faithful", depth=18, volume=(128, 128, 128), global_batch=32,
                                    per_gpu_batch=None, modalities=("mri", "pet"), plain_head=False,
                                    title="Anat_PET_CNN: ResNet-18 MRI trunk + Small_PET_CNN PET trunk, focal loss gamma=1"),
    "all_modalities": dict(config="BASELINE.json configs[3]", kind="all", depth=18, volume=(128, 128, 128), global_batch=32,
                           per_gpu_batch=None, modalities=("mri", "pet", "tab"), plain_head=False,
                           title="3-stage MRI+PET+tabular fusion (All_Modalities_Fusion, reference-faithful duplicated "
                                 "encoders: 2 x ResNet-18 MRI + 2 x Small_PET_CNN per pair), focal loss gamma=1"),
    "mri_r50_160": dict(config="BASELINE.json configs[4]", kind="anat", depth=50, volume=(160, 192, 160), global_batch=None,
                        per_gpu_batch=8, modalities=("mri",), plain_head=True,
                        title="ResNet-50 3D MedicalNet encoder on 160x192x160 MRI volumes, 8 per GPU, sync-BN, weighted CE"),
}


def class_weights(n=3):
    return torch.tensor(CW3[:n], dtype=torch.float64)


def encoder_hparams(depth, plain_head=False, n_classes=3):
    """train_anat_cnn.py:259-280 (best-run shape) or, plain_head, SURVEY.md 8(d) config 1."""
    return dict(n_classes=n_classes, resnet_depth=depth, batchnorm_begin=not plain_head, batchnorm_dense=not plain_head,
                linear_out=[], fl_gamma=None, loss_class_weights=class_weights(n_classes), lr=1e-3, lr_pretrained=1e-4,
                l2_reg=1e-4, reduce_factor_lr_schedule=None, norm_percentile=0.98)


def small_pet_hparams(n_classes=3):
    """train_pet_cnn.py:227-228."""
    return dict(n_classes=n_classes, conv_out=[8, 16, 32, 64], filter_size=[5, 5, 3, 3], linear_out=64, batchnorm=False,
                loss_class_weights=class_weights(n_classes), lr=1e-3, reduce_factor_lr_schedule=None)


def fusion_hparams(n_classes=3):
    """train_anat_pet_fusion.py:255 (fl_gamma 1); lower stages trainable (lr_pretrained set)."""
    return dict(n_classes=n_classes, fl_gamma=1, loss_class_weights=class_weights(n_classes), lr=1e-3, lr_pretrained=1e-4,
                l2_reg=1e-4, reduce_factor_lr_schedule=None, simple_dim_red=False, ensemble_size=4)


def product_namespace():
    from .pkg.models.fusion_models.all_modalities_fusion import All_Modalities_Fusion
    from .pkg.models.fusion_models.anat_pet_fusion import Anat_PET_CNN, ResNet_PET_Trunk
    from .pkg.models.fusion_models.pet_tabular_fusion import PET_TABULAR_CNN
    from .pkg.models.fusion_models.tabular_mri_fusion import Tabular_MRT_Model
    from .pkg.models.mri_models.anat_cnn import Anat_CNN
    from .pkg.models.pet_models.pet_cnn import Small_PET_CNN
    from .pkg.models.pet_models.pet_resnet_cnn import PET_CNN_ResNet
    return types.SimpleNamespace(Anat_CNN=Anat_CNN, PET_CNN_ResNet=PET_CNN_ResNet, Small_PET_CNN=Small_PET_CNN,
                                 Anat_PET_CNN=Anat_PET_CNN, ResNet_PET_Trunk=ResNet_PET_Trunk,
                                 Tabular_MRT_Model=Tabular_MRT_Model, PET_TABULAR_CNN=PET_TABULAR_CNN,
                                 All_Modalities_Fusion=All_Modalities_Fusion, tab_key="tabular")


def build_model(ns, name, depth=None, seed=SEED):
    """The model of workload `name` from the classes of `ns` (random init under `seed`, on the CPU, train mode)."""
    w = WORKLOADS[name]
    depth = depth or w["depth"]
    enc = encoder_hparams(depth, w["plain_head"])
    torch.manual_seed(seed)
    kind = w["kind"]
    if kind == "anat":
        model = ns.Anat_CNN(dict(enc))
    elif kind == "fusion2":
        model = ns.Anat_PET_CNN(fusion_hparams(), model_mri=ns.Anat_CNN(dict(enc)),
                                pet_trunk=ns.ResNet_PET_Trunk(ns.PET_CNN_ResNet(dict(enc))))
    elif kind == "fusion_faithful":
        model = ns.Anat_PET_CNN(fusion_hparams(), model_pet=ns.Small_PET_CNN(small_pet_hparams()),
                                model_mri=ns.Anat_CNN(dict(enc)))
    elif kind == "all":
        anat_pet = ns.Anat_PET_CNN(fusion_hparams(), model_pet=ns.Small_PET_CNN(small_pet_hparams()),
                                   model_mri=ns.Anat_CNN(dict(enc)))
        anat_tab = ns.Tabular_MRT_Model(fusion_hparams(), model_mri=ns.Anat_CNN(dict(enc)))
        pet_tab = ns.PET_TABULAR_CNN(fusion_hparams(), model_pet=ns.Small_PET_CNN(small_pet_hparams()))
        model = ns.All_Modalities_Fusion(fusion_hparams(), model_anat_pet=anat_pet, model_anat_tab=anat_tab,
                                         model_pet_tab=pet_tab)
    else:
        raise ValueError(kind)
    return model.train()


def per_tensor_groups(model):
    """One Adam group per trainable tensor, head tensors at lr, stage-1 encoders at lr_pretrained - the shape every
    configure_optimizers of the reference produces (anat_cnn.py:111-128, anat_pet_fusion.py:94-118).  Used for models
    whose class in the namespace has no configure_optimizers of its own (the oracle's fusion classes)."""
    hp = model.hparams
    groups = []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        stage1 = (".model." in "." + n or n.startswith("model.")) and "conv_seg" not in n
        groups.append({"params": p, "lr": hp["lr_pretrained"] if stage1 else hp["lr"]})
    return groups


# ------------------------------------------------------------------------------------------------- synthetic inputs
def _gen(index, stream):
    return torch.Generator().manual_seed(SEED * 1_000_003 + 7919 * int(index) + stream)


def brain_mask(shape):
    """Centred ellipsoid with semi-axes 0.42 x (D, H, W) (about 31 % of the voxels), SURVEY.md 8(d)."""
    axes = [((torch.arange(n, dtype=torch.float32) - (n - 1) / 2) / (0.42 * n)) ** 2 for n in shape]
    return (axes[0][:, None, None] + axes[1][None, :, None] + axes[2][None, None, :]) <= 1


def synth_sample(index, shape, modalities=("mri", "pet"), n_classes=3):
    """Raw inputs of global sample `index` (CPU tensors): MRI 400|N(0,1)|+50U(0,1) fp32 with an ellipsoid brain mask and
    0.5 % exact zeros inside it; PET max(0, N(0.5145, 0.5383)) fp32; tabular activation N(0,1) (1024,) fp32; label."""
    out = {}
    shape = tuple(shape)
    if "mri" in modalities:
        g = _gen(index, 0)
        mri = 400 * torch.randn(shape, generator=g).abs() + 50 * torch.rand(shape, generator=g)
        mask = brain_mask(shape)
        zero = torch.rand(shape, generator=g) < 0.005
        mri[zero & mask] = 0.0
        out["mri_raw"] = mri.float().contiguous()
        out["mask"] = mask.to(torch.uint8).contiguous()
    if "pet" in modalities:
        g = _gen(index, 1)
        out["pet_raw"] = (torch.randn(shape, generator=g) * PET_STD + PET_MEAN).clamp_min(0).float().contiguous()
    if "tab" in modalities:
        out["tabular"] = torch.randn((1024,), generator=_gen(index, 2), dtype=torch.float32)
    out["label"] = torch.randint(0, n_classes, (), generator=_gen(index, 3))
    return out


def synth_batch(first, count, shape, modalities=("mri", "pet"), n_classes=3):
    """Global samples [first, first + count) stacked (CPU tensors)."""
    samples = [synth_sample(first + i, shape, modalities, n_classes) for i in range(count)]
    return {k: torch.stack([s[k] for s in samples]).contiguous() for k in samples[0]}


def normalized_batch_gpu(raw, out_dtype=torch.bfloat16):
    """The reference DataLoader's normalisation on the device (dataloader.py:213-215, 261-270): raw device tensors of
    `synth_batch` -> the batch dict the LightningModules consume."""
    from .pkg.utils import normalization as norm
    batch = {"label": raw["label"]}
    if "mri_raw" in raw:
        batch["mri"] = norm.normalize_mri_per_scan_min_max(raw["mri_raw"], raw["mask"], 0.98, out_dtype=out_dtype)
    if "pet_raw" in raw:
        batch["pet1451"] = norm.normalize_pet(raw["pet_raw"], PET_MEAN, PET_STD, out_dtype=out_dtype)
    if "tabular" in raw:
        batch["tabular"] = raw["tabular"]
    return batch


# ------------------------------------------------------------------------------------------------- algorithmic FLOPs
def _ext(n, k, s, p, d=1):
    return (n + 2 * p - d * (k - 1) - 1) // s + 1


def resnet_conv_flops(depth, volume):
    """(forward FLOPs of all convs, forward FLOPs of the stem) of one MedicalNet ResNet on one `volume`:
    2 * D'H'W' * Cout * Cin * k^3 per conv."""
    basic = depth in (10, 18, 34)
    blocks = {10: [1, 1, 1, 1], 18: [2, 2, 2, 2], 34: [3, 4, 6, 3], 50: [3, 4, 6, 3]}[depth]
    exp = 1 if basic else 4
    dims = [_ext(n, 7, 2, 3) for n in volume]
    vol = lambda d: d[0] * d[1] * d[2]  # noqa: E731
    stem = 2 * vol(dims) * 64 * 343
    dims = [_ext(n, 3, 2, 1) for n in dims]
    total, inpl = stem, 64
    for li, (planes, n) in enumerate(zip([64, 128, 256, 512], blocks)):
        for b in range(n):
            stride = 2 if (li == 1 and b == 0) else 1
            out = [(n_ - 1) // stride + 1 for n_ in dims]
            if basic:
                total += 2 * vol(out) * planes * inpl * 27 + 2 * vol(out) * planes * planes * 27
            else:
                total += 2 * vol(dims) * planes * inpl                 # 1x1x1 reduce (stride on conv2)
                total += 2 * vol(out) * planes * planes * 27
                total += 2 * vol(out) * planes * exp * planes
            if stride != 1 or inpl != planes * exp:
                total += 2 * vol(out) * planes * exp * inpl
            inpl, dims = planes * exp, out
    return total, stem


def small_pet_conv_flops(volume, conv_out=(8, 16, 32, 64), filter_size=(5, 5, 3, 3)):
    dims, cin, total, first = list(volume), 1, 0, None
    for co, k in zip(conv_out, filter_size):
        f = 2 * dims[0] * dims[1] * dims[2] * co * cin * k ** 3
        first = f if first is None else first
        total += f
        dims, cin = [n // 2 for n in dims], co
    return total, first


def train_flops_per_sample(name, depth=None, volume=None):
    """Algorithmic conv FLOPs of one training step per SAMPLE of the workload (a sample = one subject: an MRI volume,
    or an (MRI, PET) pair): forward + wgrad + dgrad, no dgrad for the first conv of each encoder."""
    w = WORKLOADS[name]
    depth, volume = depth or w["depth"], tuple(volume or w["volume"])
    rf, rs = resnet_conv_flops(depth, volume)
    resnet = 3 * rf - rs
    pf, ps = small_pet_conv_flops(volume)
    small = 3 * pf - ps
    return {"anat": resnet, "fusion2": 2 * resnet, "fusion_faithful": resnet + small, "all": 2 * resnet + 2 * small}[w["kind"]]


def volumes_per_sample(name):
    return sum(1 for m in WORKLOADS[name]["modalities"] if m in ("mri", "pet"))

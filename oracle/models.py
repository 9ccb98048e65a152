"""Oracle restatement of the reference LightningModules (test infrastructure; plain torch.nn, CPU fp32).

Each class follows the reference file it names line by line; the only liberties are (a) a Lightning stand-in
(pytorch_lightning is not installed), (b) stage-N models accept already-built lower-stage modules instead of
checkpoint paths (the reference calls `load_from_checkpoint`, which needs files on the authors' cluster), and
(c) the TabPFN activation is an input tensor (B, 1024): the reference detaches it (pet_tabular_fusion.py:83).
"""
import torch
import torch.nn as nn

from .losses import make_criterion
from .medicalnet import feature_width, generate_model


class _Hparams(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class LightningStandIn(nn.Module):
    """The slice of pl.LightningModule the path uses: save_hyperparameters / hparams / log."""

    def __init__(self):
        super().__init__()
        self.logged = {}

    def save_hyperparameters(self, hparams, ignore=None):
        self._hparams = _Hparams(hparams)

    @property
    def hparams(self):
        return self._hparams

    def log(self, name, value, **kwargs):
        self.logged[name] = value

    # base_model.py:60-85 (metric updates omitted: torchmetrics is logging, not path arithmetic)
    def training_step(self, batch, batch_idx):
        return self.general_step(batch, batch_idx, "train")

    def validation_step(self, batch, batch_idx):
        return self.general_step(batch, batch_idx, "val")

    def test_step(self, batch, batch_idx):
        return self.general_step(batch, batch_idx, "test")

    def predict_step(self, batch, batch_idx):
        return self.general_step(batch, batch_idx, "pred")


def _resnet_head(hparams, n_in):
    """anat_cnn.py:33-79 / pet_resnet_cnn.py:37-81."""
    modules = nn.ModuleList()
    if "batchnorm_begin" in hparams and hparams["batchnorm_begin"]:
        modules.append(nn.BatchNorm3d(n_in))
    if "conv_out" in hparams:
        for n_out, filter_size in zip(hparams["conv_out"], hparams["filter_size"]):
            modules.append(nn.Conv3d(n_in, n_out, filter_size, padding="same"))
            if hparams["batchnorm_conv"]:
                modules.append(nn.BatchNorm3d(n_out))
            modules.append(nn.ReLU())
            modules.append(nn.MaxPool3d(2))
            n_in = n_out
    modules.append(nn.AdaptiveAvgPool3d(1))
    modules.append(nn.Flatten())
    for n_out in hparams["linear_out"]:
        modules.append(nn.Linear(n_in, n_out))
        if "batchnorm_dense" in hparams and hparams["batchnorm_dense"]:
            modules.append(nn.BatchNorm1d(n_out))
        modules.append(nn.ReLU())
        n_in = n_out
    modules.append(nn.Linear(n_in, hparams["n_classes"]))
    modules.append(nn.ReLU())                                  # final ReLU on the logits (anat_cnn.py:76-77)
    return nn.Sequential(*modules)


class Anat_CNN(LightningStandIn):
    """pkg/models/mri_models/anat_cnn.py:13-136."""
    modality = "mri"

    def __init__(self, hparams, gpu_id=None):
        super().__init__()
        self.save_hyperparameters(hparams)
        self.model = generate_model(hparams["resnet_depth"])
        self.model.conv_seg = _resnet_head(hparams, feature_width(hparams["resnet_depth"]))
        self.criterion = make_criterion(hparams)

    def forward(self, x):
        return self.model(x)

    def general_step(self, batch, batch_idx, mode):
        x = batch[self.modality].unsqueeze(1).to(dtype=torch.float32)      # anat_cnn.py:100-103
        y = batch["label"]
        y_hat = self.forward(x).to(dtype=torch.double)                     # :104
        loss = self.criterion(y_hat, y)                                    # :106
        if mode != "pred":
            self.log(mode + "_loss", loss)
        return {"loss": loss, "outputs": y_hat, "labels": y}

    def configure_optimizers(self):
        params = []                                                        # anat_cnn.py:111-128
        for name, param in self.model.named_parameters():
            if "conv_seg" in name:
                params.append({"params": param, "lr": self.hparams["lr"]})
            elif "lr_pretrained" not in self.hparams or not self.hparams["lr_pretrained"]:
                param.requires_grad = False
                params.append({"params": param})
            else:
                param.requires_grad = True
                params.append({"params": param, "lr": self.hparams["lr_pretrained"]})
        return torch.optim.Adam(params, weight_decay=self.hparams["l2_reg"])


class PET_CNN_ResNet(Anat_CNN):
    """pkg/models/pet_models/pet_resnet_cnn.py:12-166 — the same graph on batch['pet1451'] (:124-138)."""
    modality = "pet1451"


class Small_PET_CNN(LightningStandIn):
    """pkg/models/pet_models/pet_cnn.py:10-83."""

    def __init__(self, hparams, gpu_id=None):
        super().__init__()
        self.save_hyperparameters(hparams)
        modules = nn.ModuleList()
        n_in = 1
        for n_out, filter_size in zip(hparams["conv_out"], hparams["filter_size"]):
            modules.append(nn.Conv3d(n_in, n_out, filter_size, padding="same"))
            if "batchnorm" in hparams and hparams["batchnorm"]:
                modules.append(nn.BatchNorm3d(n_out))
            modules.append(nn.ReLU())
            modules.append(nn.MaxPool3d(2))
            if "dropout_conv_p" in hparams:
                modules.append(nn.Dropout(p=hparams["dropout_conv_p"]))
            n_in = n_out
        modules.append(nn.AdaptiveAvgPool3d(1))
        modules.append(nn.Flatten())
        if "linear_out" in hparams and hparams["linear_out"]:
            n_out = hparams["linear_out"]
            if "dropout_dense_p" in hparams:
                modules.append(nn.Dropout(p=hparams["dropout_dense_p"]))
            modules.append(nn.Linear(n_in, n_out))
            modules.append(nn.ReLU())
        modules.append(nn.Linear(n_out, hparams["n_classes"]))
        self.model = nn.Sequential(*modules)
        self.criterion = nn.CrossEntropyLoss(weight=hparams["loss_class_weights"])   # always CE (pet_cnn.py:47-48)

    def forward(self, x):
        return self.model(x)

    def general_step(self, batch, batch_idx, mode):
        x = batch["pet1451"].unsqueeze(1).to(dtype=torch.float32)
        y = batch["label"]
        y_hat = self.forward(x).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        if mode != "pred":
            self.log(mode + "_loss", loss)
        return {"loss": loss, "outputs": y_hat, "labels": y}

    def configure_optimizers(self):
        return torch.optim.Adam(self.model.parameters(), lr=self.hparams["lr"])


def _freeze(module):
    for _, p in module.named_parameters():
        p.requires_grad = False


def _truncate_pet(model_pet, n_classes):
    # anat_pet_fusion.py:28-31 / pet_tabular_fusion.py:28-31
    return model_pet.model[:-3] if n_classes == 2 else model_pet.model[:-1]


class Anat_PET_CNN(LightningStandIn):
    """pkg/models/fusion_models/anat_pet_fusion.py:11-127.  `model_pet` may be a Small_PET_CNN (the reference)
    or any module mapping (B,1,D,H,W) -> (B,64) (the two-ResNet north-star variant passes `pet_trunk`)."""

    def __init__(self, hparams, model_pet=None, model_mri=None, pet_trunk=None):
        super().__init__()
        self.save_hyperparameters(hparams)
        self.model_pet = pet_trunk if pet_trunk is not None else _truncate_pet(model_pet, hparams["n_classes"])
        self.model_mri = model_mri
        self.model_mri.model.conv_seg = self.model_mri.model.conv_seg[:2]            # :32
        if "lr_pretrained" not in hparams.keys() or not hparams["lr_pretrained"]:    # :35-40
            _freeze(self.model_pet)
            _freeze(self.model_mri)
        self.stage2out = nn.Linear(64 + 64, 64)
        self.cls2 = nn.Linear(64, hparams["n_classes"])
        self.relu = nn.ReLU()
        self.reduce_dim_mri = nn.Sequential(nn.Linear(512, 64), self.relu)
        self.model_fuse = nn.Sequential(self.stage2out, self.relu, self.cls2)
        self.criterion = make_criterion(hparams)

    def forward(self, x_pet, x_mri):
        bs = x_mri.shape[0]
        out_pet = self.model_pet(x_pet)
        out_mri = self.model_mri(x_mri).view(bs, -1)
        out_mri = self.reduce_dim_mri(out_mri)
        return self.model_fuse(torch.cat((out_pet, out_mri), dim=1))

    def general_step(self, batch, batch_idx, mode):
        x_pet = batch["pet1451"].unsqueeze(1).to(dtype=torch.float32)
        x_mri = batch["mri"].unsqueeze(1).to(dtype=torch.float32)
        y = batch["label"]
        y_hat = self(x_pet, x_mri).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        self.log(mode + "_loss", loss)
        return {"loss": loss, "outputs": y_hat, "labels": y}


class Tabular_MRT_Model(LightningStandIn):
    """pkg/models/fusion_models/tabular_mri_fusion.py:11-124; x_tabular = detached TabPFN activation (B,1024)."""

    def __init__(self, hparams, model_mri=None):
        super().__init__()
        self.save_hyperparameters(hparams)
        self.model_mri = model_mri
        self.model_mri.model.conv_seg = self.model_mri.model.conv_seg[:2]            # :19
        if "lr_pretrained" not in hparams.keys() or not hparams["lr_pretrained"]:
            _freeze(self.model_mri)
        self.stage2out = nn.Linear(512 + 512, 64)
        self.cls2 = nn.Linear(64, hparams["n_classes"])
        self.relu = nn.ReLU()
        self.reduce_tab = nn.Sequential(nn.Linear(1024, 512), self.relu)             # :40
        self.model_fuse = nn.Sequential(self.stage2out, self.relu, self.cls2)
        self.criterion = make_criterion(hparams)

    def forward(self, x_tabular, x_mri):
        out_tabular = self.reduce_tab(x_tabular)
        out_mri = self.model_mri(x_mri).squeeze()                                    # :77
        if out_mri.dim() == 1:
            out_mri = out_mri.unsqueeze(0)
        return self.model_fuse(torch.cat((out_tabular, out_mri), dim=1))


class PET_TABULAR_CNN(LightningStandIn):
    """pkg/models/fusion_models/pet_tabular_fusion.py:15-149."""

    def __init__(self, hparams, model_pet=None, pet_trunk=None):
        super().__init__()
        self.save_hyperparameters(hparams)
        self.model_pet = pet_trunk if pet_trunk is not None else _truncate_pet(model_pet, hparams["n_classes"])
        if "lr_pretrained" not in hparams.keys() or not hparams["lr_pretrained"]:
            _freeze(self.model_pet)
        self.stage2out = nn.Linear(64 + 64, 64)
        self.cls2 = nn.Linear(64, hparams["n_classes"])
        self.relu = nn.ReLU()
        if hparams["simple_dim_red"]:                                                # :54-57
            self.reduce_tab = nn.Sequential(nn.Linear(1024, 512), self.relu, nn.Linear(512, 64), self.relu)
        else:
            self.reduce_tab = nn.Sequential(nn.Linear(1024, 64), self.relu)
        self.model_fuse = nn.Sequential(self.stage2out, self.relu, self.cls2)
        self.criterion = make_criterion(hparams)

    def forward(self, x_pet, x_tabular):
        out_pet = self.model_pet(x_pet)
        out_tab = self.reduce_tab(x_tabular)
        return self.model_fuse(torch.cat((out_pet, out_tab), dim=1))


class All_Modalities_Fusion(LightningStandIn):
    """pkg/models/fusion_models/all_modalities_fusion.py:12-137."""

    def __init__(self, hparams, model_anat_pet, model_anat_tab, model_pet_tab):
        super().__init__()
        self.save_hyperparameters(hparams)
        self.model_anat_pet, self.model_anat_tab, self.model_pet_tab = model_anat_pet, model_anat_tab, model_pet_tab
        for m in (self.model_anat_pet, self.model_anat_tab, self.model_pet_tab):
            m.model_fuse = m.model_fuse[:-2]                                         # :29-31 -> [stage2out] only
        if "lr_pretrained" not in hparams.keys() or not hparams["lr_pretrained"]:    # :34-47
            _freeze(self.model_anat_pet.reduce_dim_mri)
            _freeze(self.model_anat_pet.model_fuse)
            _freeze(self.model_anat_tab.reduce_tab)
            _freeze(self.model_anat_tab.model_fuse)
            _freeze(self.model_pet_tab.model_fuse)
            _freeze(self.model_pet_tab.reduce_tab)
        self.stage3out = nn.Linear(64 + 64 + 64, 64)
        self.cls3 = nn.Linear(64, hparams["n_classes"])
        self.relu = nn.ReLU()
        self.model_fuse = nn.Sequential(self.stage3out, self.relu, self.cls3)
        self.criterion = make_criterion(hparams)

    def forward(self, x_pet, x_mri, x_tab):
        out_anat_pet = self.model_anat_pet(x_pet, x_mri)
        out_anat_tab = self.model_anat_tab(x_tab, x_mri)
        out_pet_tab = self.model_pet_tab(x_pet, x_tab)
        return self.model_fuse(torch.cat((out_anat_pet, out_anat_tab, out_pet_tab), dim=1))

    def general_step(self, batch, batch_idx, mode):
        x_pet = batch["pet1451"].unsqueeze(1).to(dtype=torch.float32)
        x_mri = batch["mri"].unsqueeze(1).to(dtype=torch.float32)
        x_tab = batch["tabular_features"].to(dtype=torch.float32)       # stand-in for the TabPFN activation
        y = batch["label"]
        y_hat = self(x_pet, x_mri, x_tab).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        self.log(mode + "_loss", loss)
        return {"loss": loss, "outputs": y_hat, "labels": y}


def _small_backbone(hparams, n_in):
    """The conv stack shared by pet_cnn.py:18-28, early_fusion.py:34-44 and anat_pet_featuremapfusion.py:40-64."""
    modules = nn.ModuleList()
    for n_out, filter_size in zip(hparams["conv_out"], hparams["filter_size"]):
        modules.append(nn.Conv3d(n_in, n_out, filter_size, padding="same"))
        if "batchnorm" in hparams and hparams["batchnorm"]:
            modules.append(nn.BatchNorm3d(n_out))
        modules.append(nn.ReLU())
        modules.append(nn.MaxPool3d(2))
        if "dropout_conv_p" in hparams:
            modules.append(nn.Dropout(p=hparams["dropout_conv_p"]))
        n_in = n_out
    return modules, n_in


class PET_MRI_EF(LightningStandIn):
    """pkg/models/fusion_models/early_fusion.py:19-110."""

    def __init__(self, hparams, gpu_id=None):
        super().__init__()
        self.save_hyperparameters(hparams)
        modules, n_in = _small_backbone(hparams, 2)                     # :31-44 two input channels
        modules.append(nn.AdaptiveAvgPool3d(1))
        modules.append(nn.Flatten())
        if "linear_out" in hparams and hparams["linear_out"]:           # :51-58
            n_out = hparams["linear_out"]
            if "dropout_dense_p" in hparams:
                modules.append(nn.Dropout(p=hparams["dropout_dense_p"]))
            modules.append(nn.Linear(n_in, n_out))
            modules.append(nn.ReLU())
        modules.append(nn.Linear(n_out, hparams["n_classes"]))
        self.model = nn.Sequential(*modules)
        self.criterion = nn.CrossEntropyLoss(weight=hparams["loss_class_weights"])   # :66-67

    def forward(self, x):
        return self.model(x)

    def general_step(self, batch, batch_idx, mode):
        x = torch.stack((batch["pet1451"], batch["mri"]), dim=1).to(dtype=torch.float32)   # :84-88
        y = batch["label"]
        y_hat = self.forward(x).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        if mode != "pred":
            self.log(mode + "_loss", loss)
        return {"loss": loss, "outputs": y_hat, "labels": y}

    def configure_optimizers(self):
        return torch.optim.Adam(self.model.parameters(), lr=self.hparams["lr"])        # :104-105


class PET_MRI_FMF(LightningStandIn):
    """pkg/models/fusion_models/anat_pet_featuremapfusion.py:20-172."""

    def __init__(self, hparams, gpu_id=None):
        super().__init__()
        self.save_hyperparameters(hparams)
        assert hparams["fusion_mode"] == "concatenate" or hparams["fusion_mode"] == "maxout"   # :31-32
        self.fusion_mode = hparams["fusion_mode"]
        modules_pet, n_in = _small_backbone(hparams, 1)
        modules_mri, _ = _small_backbone(hparams, 1)
        self.backbone_pet = nn.Sequential(*modules_pet)
        self.backbone_mri = nn.Sequential(*modules_mri)
        n_in_fusion = 2 * n_in if self.fusion_mode == "concatenate" else n_in             # :70-73
        modules_fused = nn.ModuleList()
        for _ in range(hparams["n_layers_fusion"]):                                       # :75-81
            modules_fused.append(nn.Conv3d(n_in_fusion, hparams["n_out_fusion"], hparams["filter_size_fusion"],
                                           padding="same"))
            if "batchnorm_fusion" in hparams and hparams["batchnorm_fusion"]:
                modules_fused.append(nn.BatchNorm3d(hparams["n_out_fusion"]))
            modules_fused.append(nn.ReLU())
            modules_fused.append(nn.MaxPool3d(2))
            n_in_fusion = n_in_fusion * 2
        modules_fused.append(nn.AdaptiveAvgPool3d(1))
        modules_fused.append(nn.Flatten())
        if "dropout_dense_p" in hparams:
            modules_fused.append(nn.Dropout(p=hparams["dropout_dense_p"]))
        modules_fused.append(nn.Linear(hparams["n_out_fusion"], 64))
        modules_fused.append(nn.ReLU())
        modules_fused.append(nn.Linear(64, hparams["n_classes"]))
        self.fuse_model = nn.Sequential(*modules_fused)
        self.criterion = nn.CrossEntropyLoss(weight=hparams["loss_class_weights"])

    def forward(self, x_pet, x_mri):
        out_pet = self.backbone_pet(x_pet)
        out_mri = self.backbone_mri(x_mri)
        if self.fusion_mode == "concatenate":
            out_fused = torch.cat((out_pet, out_mri), dim=1)                               # :118-119
        else:
            out_fused = torch.stack((out_pet, out_mri), dim=0)                             # :121-123
            out_fused, _ = torch.max(out_fused, dim=0)
        return self.fuse_model(out_fused)

    def general_step(self, batch, batch_idx, mode):
        x_pet = batch["pet1451"].unsqueeze(1).to(dtype=torch.float32)
        x_mri = batch["mri"].unsqueeze(1).to(dtype=torch.float32)
        y = batch["label"]
        y_hat = self.forward(x_pet=x_pet, x_mri=x_mri).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        if mode != "pred":
            self.log(mode + "_loss", loss)
        return {"loss": loss, "outputs": y_hat, "labels": y}

    def configure_optimizers(self):
        params = []
        for module in (self.backbone_mri, self.backbone_pet, self.fuse_model):             # :149-161
            for _, p in module.named_parameters():
                params.append({"params": p, "lr": self.hparams["lr"]})
        return torch.optim.Adam(params, weight_decay=self.hparams["l2_reg"])

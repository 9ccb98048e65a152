"""PET_MRI_FMF — feature-map fusion: two small conv backbones (PET, MRI), their feature maps fused by channel
concatenation or voxel-wise maximum, a fusion conv stack and an MLP head
(reference pkg/models/fusion_models/anat_pet_featuremapfusion.py:20-172).  SURVEY.md §8(f) N3.

As in the reference: `n_in_fusion` doubles per fusion layer while every fusion conv emits `n_out_fusion` channels
(:73-79), so only `n_layers_fusion == 1` (the only value the HPO script offers, train_anat_pet_featuremapfusion.py:69)
yields a consistent stack; more layers fail at run time in the reference too and are rejected here at construction.
`filter_size_fusion == 4` ('same' with an even kernel) pads like torch: one voxel low, two high (nn.Conv3d here)."""
import torch

from .... import autograd as A
from .... import nn as bnn
from ...loss_functions.focalloss import CrossEntropyLoss
from .._stacks import stem_stack
from ..base_model import Base_Model, adam_or_plateau, volume_input


def _backbone(hparams):
    layers, width = stem_stack(hparams, in_channels=1)
    return bnn.Sequential(*layers), width


class PET_MRI_FMF(Base_Model):
    def __init__(self, hparams, gpu_id=None):
        super().__init__(hparams, gpu_id=gpu_id)
        assert hparams["fusion_mode"] == "concatenate" or hparams["fusion_mode"] == "maxout"  # :31-32
        self.fusion_mode = hparams["fusion_mode"]
        self.backbone_pet, n_in = _backbone(self.hparams)
        self.backbone_mri, _ = _backbone(self.hparams)
        n_in_fusion = 2 * n_in if self.fusion_mode == "concatenate" else n_in
        if hparams["n_layers_fusion"] != 1:
            raise ValueError("n_layers_fusion != 1 builds an inconsistent conv stack in the reference "
                             "(anat_pet_featuremapfusion.py:73-79); only 1 is supported")
        modules_fused = []
        for _ in range(hparams["n_layers_fusion"]):
            modules_fused.append(bnn.Conv3d(n_in_fusion, hparams["n_out_fusion"], hparams["filter_size_fusion"],
                                            padding="same"))
            if "batchnorm_fusion" in self.hparams and self.hparams["batchnorm_fusion"]:
                modules_fused.append(bnn.BatchNorm3d(hparams["n_out_fusion"]))
            modules_fused.append(bnn.ReLU())
            modules_fused.append(bnn.MaxPool3d(2))
            n_in_fusion = n_in_fusion * 2
        modules_fused.append(bnn.AdaptiveAvgPool3d(1))
        modules_fused.append(bnn.Flatten())
        if "dropout_dense_p" in self.hparams:
            modules_fused.append(bnn.Dropout(p=self.hparams["dropout_dense_p"]))
        modules_fused.append(bnn.Linear(hparams["n_out_fusion"], 64))
        modules_fused.append(bnn.ReLU())
        modules_fused.append(bnn.Linear(64, self.hparams["n_classes"]))
        self.fuse_model = bnn.Sequential(*modules_fused)
        self.criterion = CrossEntropyLoss(weight=hparams["loss_class_weights"])

    def forward(self, x_pet, x_mri):
        out_pet = self.backbone_pet(x_pet)
        out_mri = self.backbone_mri(x_mri)
        if self.fusion_mode == "concatenate":
            out_fused = A.ConcatChannelsFn.apply(out_pet, out_mri)   # torch.cat(dim=1), :118-119
        else:
            out_fused = A.MaxOutFn.apply(out_pet, out_mri)           # torch.max over the stacked pair, :121-123
        return self.fuse_model(out_fused)

    def general_step(self, batch, batch_idx, mode):
        x_pet = volume_input(batch["pet1451"])
        x_mri = volume_input(batch["mri"])
        y = batch["label"]
        y_hat = self.forward(x_pet=x_pet, x_mri=x_mri).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        if mode != "pred":
            self.log(mode + "_loss", loss, on_step=True)
        return {"loss": loss, "outputs": y_hat, "labels": y}

    def configure_optimizers(self):
        parameters_optim = []
        for module in (self.backbone_mri, self.backbone_pet, self.fuse_model):      # :149-161
            for _, param in module.named_parameters():
                parameters_optim.append({"params": param, "lr": self.hparams["lr"]})
        return adam_or_plateau(self.hparams, parameters_optim, weight_decay=self.hparams["l2_reg"])

"""Anat_CNN — MRI classifier: MedicalNet ResNet-{10,18,50} encoder + configurable head
(reference pkg/models/mri_models/anat_cnn.py:13-136), on the B200-native kernels.

Differences from the reference constructor, all outside the arithmetic: the encoder is built locally
(multimodal_alzheimer_b200.medicalnet) instead of imported from an un-vendored MedicalNet checkout; the
cluster-local pretrained file (`anat_cnn.py:19`) is loaded only if hparams['pretrain_path'] names an existing
file (otherwise the upstream random init, as BASELINE.json's configs ask); CUDA_VISIBLE_DEVICES is not required.
"""
import os

import torch

from .... import nn as bnn
from ...loss_functions.focalloss import make_criterion
from ...medicalnet_glue import build_encoder
from ..base_model import Base_Model, adam_or_plateau, volume_input


def feature_width(depth):
    # anat_cnn.py:37-46
    if depth in (10, 18):
        return 512
    if depth == 50:
        return 2048
    raise ValueError("hparams['resnet_depth'] is not in [10, 18, 34, 50]")


def build_resnet_head(hparams, n_in):
    """conv_seg replacement, anat_cnn.py:33-79 (identical in pet_resnet_cnn.py:37-81)."""
    modules = []
    if "batchnorm_begin" in hparams and hparams["batchnorm_begin"]:
        modules.append(bnn.BatchNorm3d(n_in))
    if "conv_out" in hparams:
        for n_out, filter_size in zip(hparams["conv_out"], hparams["filter_size"]):
            modules.append(bnn.Conv3d(n_in, n_out, filter_size, padding="same"))
            if hparams["batchnorm_conv"]:
                modules.append(bnn.BatchNorm3d(n_out))
            modules.append(bnn.ReLU())
            modules.append(bnn.MaxPool3d(2))
            n_in = n_out
    modules.append(bnn.AdaptiveAvgPool3d(1))
    modules.append(bnn.Flatten())
    for n_out in hparams["linear_out"]:
        modules.append(bnn.Linear(n_in, n_out))
        if "batchnorm_dense" in hparams and hparams["batchnorm_dense"]:
            modules.append(bnn.BatchNorm1d(n_out))
        modules.append(bnn.ReLU())
        n_in = n_out
    modules.append(bnn.Linear(n_in, hparams["n_classes"]))
    modules.append(bnn.ReLU())  # the reference clamps the logits (anat_cnn.py:76-77)
    return bnn.Sequential(*modules)


class Anat_CNN(Base_Model):
    modality = "mri"

    def __init__(self, hparams, gpu_id=None):
        super().__init__(hparams)
        self.model = build_encoder(hparams["resnet_depth"], hparams.get("pretrain_path"))
        self.model.conv_seg = build_resnet_head(hparams, feature_width(hparams["resnet_depth"]))
        self.criterion = make_criterion(hparams)

    def forward(self, x):
        return self.model(x)

    def general_step(self, batch, batch_idx, mode):
        x = volume_input(batch[self.modality])
        y = batch["label"]
        y_hat = self.forward(x).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        if mode != "pred":
            self.log(mode + "_loss", loss, on_step=True, prog_bar=True)
        return {"loss": loss, "outputs": y_hat, "labels": y}

    def configure_optimizers(self):
        parameters_optim = []
        for name, param in self.model.named_parameters():
            if "conv_seg" in name:
                parameters_optim.append({"params": param, "lr": self.hparams["lr"]})
            elif "lr_pretrained" not in self.hparams or not self.hparams["lr_pretrained"]:
                param.requires_grad = False
                parameters_optim.append({"params": param})
            else:
                param.requires_grad = True
                parameters_optim.append({"params": param, "lr": self.hparams["lr_pretrained"]})
        return adam_or_plateau(self.hparams, parameters_optim, weight_decay=self.hparams["l2_reg"])

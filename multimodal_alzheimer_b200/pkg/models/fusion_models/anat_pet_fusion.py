"""Anat_PET_CNN — 2-stage PET-MRI fusion (reference pkg/models/fusion_models/anat_pet_fusion.py:11-127):
PET trunk (-> 64) || MRI trunk (-> 512 -> reduce_dim_mri 64) -> cat 128 -> Linear 64 -> ReLU -> Linear C.

Stage-1 models come from checkpoints in the reference (`load_from_checkpoint`, :17-23); here they may equally be
passed as module instances (`model_pet=`, `model_mri=`).  `pet_trunk=` swaps the PET branch for any module
mapping (B,1,D,H,W) -> (B,64): `ResNet_PET_Trunk` gives BASELINE.json's two-branch ResNet-18 fusion (config 3).
"""
import torch

from .... import nn as bnn
from ....branches import run_two
from ...loss_functions.focalloss import make_criterion
from ..base_model import Base_Model, adam_or_plateau, volume_input
from ..mri_models.anat_cnn import Anat_CNN
from ..pet_models.pet_cnn import Small_PET_CNN
from ..pet_models.pet_resnet_cnn import PET_CNN_ResNet


def freeze(module):
    for _, param in module.named_parameters():
        param.requires_grad = False


def truncate_pet(model_pet, n_classes):
    """anat_pet_fusion.py:28-31: 2-class -> model[:-3] (ends at Flatten), else model[:-1] (ends Linear+ReLU)."""
    return model_pet.model[:-3] if n_classes == 2 else model_pet.model[:-1]


class ResNet_PET_Trunk(torch.nn.Module):
    """PET_CNN_ResNet encoder truncated like the MRI branch (conv_seg[:2]) + Linear(512,64)+ReLU -> (B,64)."""

    def __init__(self, model_pet_resnet):
        super().__init__()
        self.encoder = model_pet_resnet
        self.encoder.model.conv_seg = self.encoder.model.conv_seg[:2]
        self.relu = bnn.ReLU()
        self.reduce_dim_pet = bnn.Sequential(bnn.Linear(512, 64), self.relu)

    def forward(self, x):
        out = self.encoder(x)
        return self.reduce_dim_pet(out.view(out.shape[0], -1))


class Anat_PET_CNN(Base_Model):
    def __init__(self, hparams, path_pet=None, path_anat=None, model_pet=None, model_mri=None, pet_trunk=None):
        super().__init__(hparams)
        if pet_trunk is None and model_pet is None:
            model_pet = Small_PET_CNN.load_from_checkpoint(path_pet or hparams["path_pet"])
        if model_mri is None:
            model_mri = Anat_CNN.load_from_checkpoint(path_anat or hparams["path_mri"])
        self.model_pet = pet_trunk if pet_trunk is not None else truncate_pet(model_pet, hparams["n_classes"])
        self.model_mri = model_mri
        self.model_mri.model.conv_seg = self.model_mri.model.conv_seg[:2]
        if "lr_pretrained" not in hparams.keys() or not self.hparams["lr_pretrained"]:
            freeze(self.model_pet)
            freeze(self.model_mri)
        self.stage2out = bnn.Linear(64 + 64, 64)
        self.cls2 = bnn.Linear(64, hparams["n_classes"])
        self.relu = bnn.ReLU()
        self.reduce_dim_mri = bnn.Sequential(bnn.Linear(512, 64), self.relu)
        self.model_fuse = bnn.Sequential(self.stage2out, self.relu, self.cls2)
        self.criterion = make_criterion(hparams)

    def forward(self, x_pet, x_mri):
        bs = x_mri.shape[0]
        # the two stage-1 trunks are independent: they run on two CUDA streams (multimodal_alzheimer_b200/branches.py)
        out_pet, out_mri = run_two(lambda: self.model_pet(x_pet), lambda: self.model_mri(x_mri), x_pet)
        out_mri = out_mri.view(bs, -1)
        out_mri = self.reduce_dim_mri(out_mri)
        out = torch.cat((out_pet, out_mri), dim=1)
        return self.model_fuse(out)

    def general_step(self, batch, batch_idx, mode):
        x_pet = volume_input(batch["pet1451"])
        x_mri = volume_input(batch["mri"])
        y = batch["label"]
        y_hat = self(x_pet, x_mri).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        self.log(mode + "_loss", loss, on_step=True, prog_bar=True)
        return {"loss": loss, "outputs": y_hat, "labels": y}

    def configure_optimizers(self):
        parameters_optim = []
        for _, param in self.model_fuse.named_parameters():
            parameters_optim.append({"params": param, "lr": self.hparams["lr"]})
        for _, param in self.reduce_dim_mri.named_parameters():
            parameters_optim.append({"params": param, "lr": self.hparams["lr"]})
        if self.hparams["lr_pretrained"]:
            for _, param in self.model_pet.named_parameters():
                parameters_optim.append({"params": param, "lr": self.hparams["lr_pretrained"]})
            for _, param in self.model_mri.named_parameters():
                parameters_optim.append({"params": param, "lr": self.hparams["lr_pretrained"]})
        return adam_or_plateau(self.hparams, parameters_optim, weight_decay=self.hparams["l2_reg"])

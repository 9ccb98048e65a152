"""Tabular_MRT_Model — MRI-tabular stage-2 fusion (reference pkg/models/fusion_models/tabular_mri_fusion.py:11-124).

The tabular branch of the reference is a frozen TabPFN 0.1.8 evaluated on the CPU whose decoder activation is
captured by a forward hook and DETACHED (:59-74); no gradient ever reaches it.  On this path the branch input is
that activation itself: `x_tabular` is a (B, 1024) feature tensor.  A callable
`tabular_features=` may be supplied to compute it from the raw (B,1,9) table.
"""
import torch

from .... import nn as bnn
from ...loss_functions.focalloss import make_criterion
from ..base_model import Base_Model, adam_or_plateau, volume_input
from ..mri_models.anat_cnn import Anat_CNN
from .anat_pet_fusion import freeze


def tabular_activation(x_tabular, extractor):
    if x_tabular.dim() == 2 and x_tabular.shape[1] == 1024:
        return x_tabular.detach().to(torch.float32)
    if extractor is None:
        raise ValueError("tabular input must be the detached (B,1024) TabPFN decoder activation, or pass "
                         "`tabular_features=` to compute it (TabPFN itself is outside the CUDA path)")
    return extractor(x_tabular).detach().to(torch.float32)


class Tabular_MRT_Model(Base_Model):
    def __init__(self, hparams, path_mri=None, model_mri=None, tabular_features=None):
        super().__init__(hparams)
        if model_mri is None:
            model_mri = Anat_CNN.load_from_checkpoint(path_mri or hparams["path_mri"])
        self.model_mri = model_mri
        self.model_mri.model.conv_seg = self.model_mri.model.conv_seg[:2]
        self.tabular_features = tabular_features
        if "lr_pretrained" not in hparams.keys() or not self.hparams["lr_pretrained"]:
            freeze(self.model_mri)
        self.stage2out = bnn.Linear(512 + 512, 64)
        self.cls2 = bnn.Linear(64, hparams["n_classes"])
        self.relu = bnn.ReLU()
        self.reduce_tab = bnn.Sequential(bnn.Linear(1024, 512), self.relu)
        self.model_fuse = bnn.Sequential(self.stage2out, self.relu, self.cls2)
        self.criterion = make_criterion(hparams)

    def forward(self, x_tabular, x_mri):
        activations = tabular_activation(x_tabular, self.tabular_features)
        out_tabular = self.reduce_tab(activations)
        out_mri = self.model_mri(x_mri)
        out_mri = out_mri.view(out_mri.shape[0], -1)   # reference: .squeeze() (:77)
        out = torch.cat((out_tabular, out_mri), dim=1)
        return self.model_fuse(out)

    def general_step(self, batch, batch_idx, mode):
        x_mri = volume_input(batch["mri"])
        y = batch["label"]
        y_hat = self(batch["tabular"], x_mri).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        self.log(mode + "_loss", loss, on_step=True, prog_bar=True)
        return {"loss": loss, "outputs": y_hat, "labels": y}

    def configure_optimizers(self):
        parameters_optim = []
        for _, param in self.model_fuse.named_parameters():
            parameters_optim.append({"params": param, "lr": self.hparams["lr"]})
        for _, param in self.reduce_tab.named_parameters():
            parameters_optim.append({"params": param, "lr": self.hparams["lr"]})
        if self.hparams["lr_pretrained"]:
            for _, param in self.model_mri.named_parameters():
                parameters_optim.append({"params": param, "lr": self.hparams["lr_pretrained"]})
        return adam_or_plateau(self.hparams, parameters_optim, weight_decay=self.hparams["l2_reg"])

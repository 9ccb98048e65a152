"""PET_TABULAR_CNN — PET-tabular stage-2 fusion (reference pkg/models/fusion_models/pet_tabular_fusion.py:15-149):
PET trunk (-> 64) || reduce_tab(TabPFN activation 1024 -> 64 | 1024 -> 512 -> 64) -> cat 128 -> 64 -> C."""
import torch

from .... import nn as bnn
from ...loss_functions.focalloss import make_criterion
from ..base_model import Base_Model, adam_or_plateau, volume_input
from ..pet_models.pet_cnn import Small_PET_CNN
from .anat_pet_fusion import freeze, truncate_pet
from .tabular_mri_fusion import tabular_activation


class PET_TABULAR_CNN(Base_Model):
    def __init__(self, hparams, path_pet=None, model_pet=None, pet_trunk=None, tabular_features=None):
        super().__init__(hparams)
        if pet_trunk is None and model_pet is None:
            model_pet = Small_PET_CNN.load_from_checkpoint(path_pet or hparams["path_pet"])
        self.model_pet = pet_trunk if pet_trunk is not None else truncate_pet(model_pet, hparams["n_classes"])
        self.tabular_features = tabular_features
        if "lr_pretrained" not in hparams.keys() or not self.hparams["lr_pretrained"]:
            freeze(self.model_pet)
        self.stage2out = bnn.Linear(64 + 64, 64)
        self.cls2 = bnn.Linear(64, hparams["n_classes"])
        self.relu = bnn.ReLU()
        if self.hparams["simple_dim_red"]:
            self.reduce_tab = bnn.Sequential(bnn.Linear(1024, 512), self.relu, bnn.Linear(512, 64), self.relu)
        else:
            self.reduce_tab = bnn.Sequential(bnn.Linear(1024, 64), self.relu)
        self.model_fuse = bnn.Sequential(self.stage2out, self.relu, self.cls2)
        self.criterion = make_criterion(hparams)

    def forward(self, x_pet, x_tabular):
        out_pet = self.model_pet(x_pet)
        activations = tabular_activation(x_tabular, self.tabular_features)
        out_tab = self.reduce_tab(activations)
        out = torch.cat((out_pet, out_tab), dim=1)
        return self.model_fuse(out)

    def general_step(self, batch, batch_idx, mode):
        x_pet = volume_input(batch["pet1451"])
        y = batch["label"]
        y_hat = self(x_pet, batch["tabular"]).to(dtype=torch.double)
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
            for _, param in self.model_pet.named_parameters():
                parameters_optim.append({"params": param, "lr": self.hparams["lr_pretrained"]})
        return adam_or_plateau(self.hparams, parameters_optim, weight_decay=self.hparams["l2_reg"])

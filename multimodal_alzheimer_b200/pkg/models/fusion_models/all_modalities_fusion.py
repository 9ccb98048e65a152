"""All_Modalities_Fusion — 3-stage fusion (reference pkg/models/fusion_models/all_modalities_fusion.py:12-137):
three stage-2 models truncated to their `stage2out` (64-d, NO ReLU, :29-31) -> cat 192 -> Linear 64 -> ReLU -> C.
As in the reference each stage-2 model owns its own copy of the stage-1 encoders (MRI and PET run twice)."""
import torch

from .... import nn as bnn
from ...loss_functions.focalloss import make_criterion
from ..base_model import Base_Model, adam_or_plateau, volume_input
from .anat_pet_fusion import Anat_PET_CNN, freeze
from .pet_tabular_fusion import PET_TABULAR_CNN
from .tabular_mri_fusion import Tabular_MRT_Model


class All_Modalities_Fusion(Base_Model):
    def __init__(self, hparams, model_anat_pet=None, model_anat_tab=None, model_pet_tab=None):
        super().__init__(hparams)
        if model_anat_pet is None:
            model_anat_pet = Anat_PET_CNN.load_from_checkpoint(hparams["path_anat_pet"], path_pet=hparams["path_pet"],
                                                               path_anat=hparams["path_anat"])
        if model_anat_tab is None:
            model_anat_tab = Tabular_MRT_Model.load_from_checkpoint(hparams["path_anat_tab"],
                                                                    path_mri=hparams["path_anat"])
        if model_pet_tab is None:
            model_pet_tab = PET_TABULAR_CNN.load_from_checkpoint(hparams["path_pet_tab"], path_pet=hparams["path_pet"])
        self.model_anat_pet, self.model_anat_tab, self.model_pet_tab = model_anat_pet, model_anat_tab, model_pet_tab
        self.model_anat_pet.model_fuse = self.model_anat_pet.model_fuse[:-2]
        self.model_anat_tab.model_fuse = self.model_anat_tab.model_fuse[:-2]
        self.model_pet_tab.model_fuse = self.model_pet_tab.model_fuse[:-2]
        if "lr_pretrained" not in hparams.keys() or not self.hparams["lr_pretrained"]:
            freeze(self.model_anat_pet.reduce_dim_mri)
            freeze(self.model_anat_pet.model_fuse)
            freeze(self.model_anat_tab.reduce_tab)
            freeze(self.model_anat_tab.model_fuse)
            freeze(self.model_pet_tab.model_fuse)
            freeze(self.model_pet_tab.reduce_tab)
        self.stage3out = bnn.Linear(64 + 64 + 64, 64)
        self.cls3 = bnn.Linear(64, hparams["n_classes"])
        self.relu = bnn.ReLU()
        self.model_fuse = bnn.Sequential(self.stage3out, self.relu, self.cls3)
        self.criterion = make_criterion(hparams)

    def forward(self, x_pet, x_mri, x_tab):
        out_anat_pet = self.model_anat_pet(x_pet, x_mri)
        out_anat_tab = self.model_anat_tab(x_tab, x_mri)
        out_pet_tab = self.model_pet_tab(x_pet, x_tab)
        out = torch.cat((out_anat_pet, out_anat_tab, out_pet_tab), dim=1)
        return self.model_fuse(out)

    def general_step(self, batch, batch_idx, mode):
        x_pet = volume_input(batch["pet1451"])
        x_mri = volume_input(batch["mri"])
        x_tab = batch["tabular"]
        y = batch["label"]
        y_hat = self(x_pet, x_mri, x_tab).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        self.log(mode + "_loss", loss, on_step=True, prog_bar=True)
        return {"loss": loss, "outputs": y_hat, "labels": y}

    def configure_optimizers(self):
        parameters_optim = []
        for _, param in self.model_fuse.named_parameters():
            parameters_optim.append({"params": param, "lr": self.hparams["lr"]})
        if self.hparams["lr_pretrained"]:
            previous = [self.model_anat_pet.model_pet, self.model_anat_pet.model_mri, self.model_anat_pet.stage2out,
                        self.model_anat_pet.reduce_dim_mri, self.model_pet_tab.model_pet, self.model_pet_tab.stage2out,
                        self.model_pet_tab.reduce_tab, self.model_anat_tab.model_mri, self.model_anat_tab.stage2out,
                        self.model_anat_tab.reduce_tab]
            for model in previous:
                for _, param in model.named_parameters():
                    parameters_optim.append({"params": param, "lr": self.hparams["lr_pretrained"]})
        return adam_or_plateau(self.hparams, parameters_optim, weight_decay=self.hparams["l2_reg"])

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


# The three stage-2 models: attribute name, class, checkpoint key in hparams, constructor arguments taken from hparams
# (all_modalities_fusion.py:17-27), and the sub-modules frozen together with them when `lr_pretrained` is unset (:34-47).
_STAGE2 = (
    ("model_anat_pet", Anat_PET_CNN, "path_anat_pet", {"path_pet": "path_pet", "path_anat": "path_anat"}, ("reduce_dim_mri",)),
    ("model_anat_tab", Tabular_MRT_Model, "path_anat_tab", {"path_mri": "path_anat"}, ("reduce_tab",)),
    ("model_pet_tab", PET_TABULAR_CNN, "path_pet_tab", {"path_pet": "path_pet"}, ("reduce_tab",)),
)
# modules handed to the optimizer at `lr_pretrained`, in the reference's order (:109-121; the TabPFN members of that
# list are not nn.Modules and are not part of this path)
_PRETRAINED_ORDER = (("model_anat_pet", ("model_pet", "model_mri", "stage2out", "reduce_dim_mri")),
                     ("model_pet_tab", ("model_pet", "stage2out", "reduce_tab")),
                     ("model_anat_tab", ("model_mri", "stage2out", "reduce_tab")))


class All_Modalities_Fusion(Base_Model):
    def __init__(self, hparams, model_anat_pet=None, model_anat_tab=None, model_pet_tab=None):
        super().__init__(hparams)
        given = {"model_anat_pet": model_anat_pet, "model_anat_tab": model_anat_tab, "model_pet_tab": model_pet_tab}
        trainable_below = "lr_pretrained" in hparams.keys() and bool(self.hparams["lr_pretrained"])
        for attr, cls, ckpt_key, ctor_args, frozen_with in _STAGE2:
            stage2 = given[attr]
            if stage2 is None:                                   # the reference's only route: checkpoints
                stage2 = cls.load_from_checkpoint(hparams[ckpt_key], **{k: hparams[v] for k, v in ctor_args.items()})
            stage2.model_fuse = stage2.model_fuse[:-2]           # keep `stage2out` only: 64 features, no ReLU (:29-31)
            if not trainable_below:
                for name in frozen_with + ("model_fuse",):
                    freeze(getattr(stage2, name))
            setattr(self, attr, stage2)
        self.stage3out = bnn.Linear(3 * 64, 64)
        self.cls3 = bnn.Linear(64, hparams["n_classes"])
        self.relu = bnn.ReLU()
        self.model_fuse = bnn.Sequential(self.stage3out, self.relu, self.cls3)
        self.criterion = make_criterion(hparams)

    def forward(self, x_pet, x_mri, x_tab):
        features = (self.model_anat_pet(x_pet, x_mri), self.model_anat_tab(x_tab, x_mri), self.model_pet_tab(x_pet, x_tab))
        return self.model_fuse(torch.cat(features, dim=1))

    def general_step(self, batch, batch_idx, mode):
        labels = batch["label"]
        logits = self(volume_input(batch["pet1451"]), volume_input(batch["mri"]), batch["tabular"]).to(dtype=torch.double)
        loss = self.criterion(logits, labels)
        self.log(mode + "_loss", loss, on_step=True, prog_bar=True)
        return {"loss": loss, "outputs": logits, "labels": labels}

    def configure_optimizers(self):
        groups = [{"params": p, "lr": self.hparams["lr"]} for p in self.model_fuse.parameters()]
        if self.hparams["lr_pretrained"]:
            for attr, members in _PRETRAINED_ORDER:
                stage2 = getattr(self, attr)
                for member in members:
                    groups += [{"params": p, "lr": self.hparams["lr_pretrained"]} for p in getattr(stage2, member).parameters()]
        return adam_or_plateau(self.hparams, groups, weight_decay=self.hparams["l2_reg"])

"""Small_PET_CNN (reference pkg/models/pet_models/pet_cnn.py:10-83): the stem stack on one input channel followed by
the dense tail (see .._stacks); always weighted cross entropy, whatever `fl_gamma` says (:47-48); a single Adam group
over all parameters at hparams['lr'] (:72-82)."""
import torch

from .... import nn as bnn
from ...loss_functions.focalloss import CrossEntropyLoss
from .._stacks import AlwaysFirstClass, dense_tail, stem_stack
from ..base_model import Base_Model, adam_or_plateau, volume_input


class Small_PET_CNN(Base_Model):
    def __init__(self, hparams, gpu_id=None):
        super().__init__(hparams, gpu_id=gpu_id)
        convs, width = stem_stack(self.hparams, in_channels=1)
        self.model = bnn.Sequential(*convs, *dense_tail(self.hparams, width))
        self.criterion = CrossEntropyLoss(weight=hparams["loss_class_weights"])

    def forward(self, x):
        return self.model(x)

    def general_step(self, batch, batch_idx, mode):
        labels = batch["label"]
        logits = self.forward(volume_input(batch["pet1451"])).to(dtype=torch.double)   # :61-65 (fp64 for the loss)
        loss = self.criterion(logits, labels)
        if mode != "pred":
            self.log(mode + "_loss", loss, on_step=True)
        return {"loss": loss, "outputs": logits, "labels": labels}

    def configure_optimizers(self):
        return adam_or_plateau(self.hparams, self.model.parameters(), lr=self.hparams["lr"])


class Random_Benchmark_All_CN(AlwaysFirstClass, Small_PET_CNN):
    """Always-CN baseline predictor (pet_cnn.py:85-90)."""

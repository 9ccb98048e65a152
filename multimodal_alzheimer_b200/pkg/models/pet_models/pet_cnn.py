"""Small_PET_CNN — n x (Conv3d 'same' + [BN] + ReLU + MaxPool3d(2) + [Dropout]) -> GAP -> [Linear+ReLU] -> Linear
(reference pkg/models/pet_models/pet_cnn.py:10-83); weighted cross entropy always (:47-48)."""
import torch

from .... import nn as bnn
from ...loss_functions.focalloss import CrossEntropyLoss
from ..base_model import Base_Model, adam_or_plateau, volume_input


class Small_PET_CNN(Base_Model):
    def __init__(self, hparams, gpu_id=None):
        super().__init__(hparams, gpu_id=gpu_id)
        modules = []
        n_in = 1
        for n_out, filter_size in zip(self.hparams["conv_out"], self.hparams["filter_size"]):
            modules.append(bnn.Conv3d(n_in, n_out, filter_size, padding="same"))
            if "batchnorm" in self.hparams and self.hparams["batchnorm"]:
                modules.append(bnn.BatchNorm3d(n_out))
            modules.append(bnn.ReLU())
            modules.append(bnn.MaxPool3d(2))
            if "dropout_conv_p" in self.hparams:
                modules.append(bnn.Dropout(p=self.hparams["dropout_conv_p"]))
            n_in = n_out
        modules.append(bnn.AdaptiveAvgPool3d(1))
        modules.append(bnn.Flatten())
        if "linear_out" in self.hparams and self.hparams["linear_out"]:
            n_out = self.hparams["linear_out"]
            if "dropout_dense_p" in self.hparams:
                modules.append(bnn.Dropout(p=self.hparams["dropout_dense_p"]))
            modules.append(bnn.Linear(n_in, n_out))
            modules.append(bnn.ReLU())
        modules.append(bnn.Linear(n_out, self.hparams["n_classes"]))
        self.model = bnn.Sequential(*modules)
        self.criterion = CrossEntropyLoss(weight=hparams["loss_class_weights"])

    def forward(self, x):
        return self.model(x)

    def general_step(self, batch, batch_idx, mode):
        x = volume_input(batch["pet1451"])
        y = batch["label"]
        y_hat = self.forward(x).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        if mode != "pred":
            self.log(mode + "_loss", loss, on_step=True)
        return {"loss": loss, "outputs": y_hat, "labels": y}

    def configure_optimizers(self):
        return adam_or_plateau(self.hparams, self.model.parameters(), lr=self.hparams["lr"])


class Random_Benchmark_All_CN(Small_PET_CNN):
    """Always-CN baseline predictor (pet_cnn.py:85-90)."""

    def forward(self, x):
        y_hat = torch.zeros_like(super().forward(x))
        y_hat[..., 0] = 1
        y_hat[..., 1:] = 0
        return y_hat

"""PET_MRI_EF — early fusion: PET and MRI stacked as the two input channels of one small CNN
(reference pkg/models/fusion_models/early_fusion.py:19-118): n x (Conv3d 'same' + [BN] + ReLU + MaxPool3d(2) +
[Dropout]) -> GAP -> [Dropout, Linear, ReLU] -> Linear; weighted cross entropy (:66-67); one Adam group over
`self.model.parameters()` at hparams['lr'] (:104-110).  SURVEY.md §8(f) N3."""
import torch

from .... import nn as bnn
from ...loss_functions.focalloss import CrossEntropyLoss
from ..base_model import Base_Model, adam_or_plateau


class PET_MRI_EF(Base_Model):
    def __init__(self, hparams, gpu_id=None):
        super().__init__(hparams, gpu_id=gpu_id)
        modules = []
        n_in = 2  # one input channel for PET and one for MRI (early_fusion.py:31-33)
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
        """x: (B, 2, D, H, W) fp32/fp64 NCDHW (channel 0 = PET, 1 = MRI), or a bf16 NDHWC (B, D, H, W, 2) volume."""
        return self.model(x)

    def general_step(self, batch, batch_idx, mode):
        x_pet, x_mri = batch["pet1451"], batch["mri"]
        y = batch["label"]
        if x_pet.dtype == torch.bfloat16:  # volumes normalised by pkg/utils/normalization.py on the device
            x = torch.stack((x_pet, x_mri), dim=-1)
        else:
            x = torch.stack((x_pet, x_mri), dim=1)  # early_fusion.py:87; the fp32 cast happens in the first kernel
        y_hat = self.forward(x).to(dtype=torch.double)
        loss = self.criterion(y_hat, y)
        if mode != "pred":
            self.log(mode + "_loss", loss, on_step=True)
        return {"loss": loss, "outputs": y_hat, "labels": y}

    def configure_optimizers(self):
        return adam_or_plateau(self.hparams, self.model.parameters(), lr=self.hparams["lr"])


class Random_Benchmark_All_CN(PET_MRI_EF):
    """Always-CN baseline predictor (early_fusion.py:112-117)."""

    def forward(self, x):
        y_hat = torch.zeros_like(super().forward(x))
        y_hat[..., 0] = 1
        y_hat[..., 1:] = 0
        return y_hat

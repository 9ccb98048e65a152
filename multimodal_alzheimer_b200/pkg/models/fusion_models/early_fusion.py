"""PET_MRI_EF — early fusion: PET and MRI stacked as the two input channels of one small CNN
(reference pkg/models/fusion_models/early_fusion.py:19-118): n x (Conv3d 'same' + [BN] + ReLU + MaxPool3d(2) +
[Dropout]) -> GAP -> [Dropout, Linear, ReLU] -> Linear; weighted cross entropy (:66-67); one Adam group over
`self.model.parameters()` at hparams['lr'] (:104-110).  SURVEY.md §8(f) N3."""
import torch

from .... import nn as bnn
from ...loss_functions.focalloss import CrossEntropyLoss
from .._stacks import AlwaysFirstClass, dense_tail, stem_stack
from ..base_model import Base_Model, adam_or_plateau


class PET_MRI_EF(Base_Model):
    def __init__(self, hparams, gpu_id=None):
        super().__init__(hparams, gpu_id=gpu_id)
        convs, width = stem_stack(self.hparams, in_channels=2)      # channel 0 = PET, 1 = MRI (early_fusion.py:31-33)
        self.model = bnn.Sequential(*convs, *dense_tail(self.hparams, width))
        self.criterion = CrossEntropyLoss(weight=hparams["loss_class_weights"])

    def forward(self, x):
        """x: (B, 2, D, H, W) fp32/fp64 NCDHW (channel 0 = PET, 1 = MRI), or a bf16 NDHWC (B, D, H, W, 2) volume."""
        return self.model(x)

    def general_step(self, batch, batch_idx, mode):
        pet, mri, labels = batch["pet1451"], batch["mri"], batch["label"]
        # early_fusion.py:87 stacks on the channel axis; volumes already normalised to bf16 on the device
        # (pkg/utils/normalization.py) are channels-last
        x = torch.stack((pet, mri), dim=-1 if pet.dtype == torch.bfloat16 else 1)
        logits = self.forward(x).to(dtype=torch.double)
        loss = self.criterion(logits, labels)
        if mode != "pred":
            self.log(mode + "_loss", loss, on_step=True)
        return {"loss": loss, "outputs": logits, "labels": labels}

    def configure_optimizers(self):
        return adam_or_plateau(self.hparams, self.model.parameters(), lr=self.hparams["lr"])


class Random_Benchmark_All_CN(AlwaysFirstClass, PET_MRI_EF):
    """Always-CN baseline predictor (early_fusion.py:112-117)."""

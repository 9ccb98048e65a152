"""FocalLoss with the reference's constructor and semantics (pkg/loss_functions/focalloss.py:12-40):
log_softmax -> gather(target) -> pt = exp(logpt) DETACHED (:30) -> -(1-pt)^gamma * logpt -> mean | sum,
evaluated in fp64 by one CUDA kernel (forward) and one (gradient)."""
import torch
import torch.nn as nn

from ... import autograd as A


class FocalLoss(nn.Module):
    def __init__(self, gamma=0, alpha=None, size_average=True):
        super().__init__()
        self.gamma = gamma
        if alpha is not None:
            raise NotImplementedError("FocalLoss(alpha=...) is never used by the reference's callers "
                                      "(anat_cnn.py:82, anat_pet_fusion.py:55) and is not on the CUDA path")
        self.alpha = None
        self.size_average = size_average

    def forward(self, input, target):
        if input.dim() > 2:
            input = input.reshape(input.size(0), input.size(1), -1).transpose(1, 2).reshape(-1, input.size(1))
        return A.LossFn.apply(input, target.reshape(-1), float(self.gamma), None, self.size_average)


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss(weight=w) as the reference uses it (anat_cnn.py:84-85): sum w[y]*nll / sum w[y]."""

    def __init__(self, weight=None):
        super().__init__()
        # a registered buffer, like torch.nn.CrossEntropyLoss (checkpoint key `criterion.weight`)
        self.register_buffer("weight", None if weight is None else torch.as_tensor(weight).clone())

    def forward(self, input, target):
        w = self.weight
        if w is not None and (w.device != input.device or w.dtype != torch.float64):
            w = w.to(device=input.device, dtype=torch.float64)
        return A.LossFn.apply(input, target.reshape(-1), 0.0, w, True)


def make_criterion(hparams):
    """Loss selection shared by every model of the path (anat_cnn.py:81-85)."""
    if "fl_gamma" in hparams and hparams["fl_gamma"]:
        return FocalLoss(gamma=hparams["fl_gamma"])
    return CrossEntropyLoss(weight=hparams["loss_class_weights"])

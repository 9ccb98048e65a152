"""Loss oracles (test infrastructure).

FocalLossOracle restates pkg/loss_functions/focalloss.py:20-40 of the reference:
  logpt = log_softmax(input, dim=1).gather(1, target)            (:27-29)
  pt    = logpt.data.exp()   -- DETACHED modulating factor        (:30)
  alpha (optional) weights logpt                                  (:32-36)
  loss  = -(1-pt)**gamma * logpt ; mean or sum                    (:38-40)
weighted CE is torch.nn.CrossEntropyLoss(weight) (pkg/models/mri_models/anat_cnn.py:84-85).
"""
import torch
import torch.nn.functional as F


class FocalLossOracle(torch.nn.Module):
    def __init__(self, gamma=0, alpha=None, size_average=True):
        super().__init__()
        self.gamma = gamma
        if isinstance(alpha, (float, int)):
            alpha = torch.tensor([alpha, 1 - alpha])
        elif isinstance(alpha, list):
            alpha = torch.tensor(alpha)
        self.alpha = alpha
        self.size_average = size_average

    def forward(self, logits, target):
        if logits.dim() > 2:
            logits = logits.reshape(logits.size(0), logits.size(1), -1).transpose(1, 2).reshape(-1, logits.size(1))
        target = target.reshape(-1, 1)
        logpt = F.log_softmax(logits, dim=1).gather(1, target).reshape(-1)
        pt = logpt.detach().exp()
        if self.alpha is not None:
            at = self.alpha.to(logits).gather(0, target.reshape(-1))
            logpt = logpt * at
        loss = -1 * (1 - pt) ** self.gamma * logpt
        return loss.mean() if self.size_average else loss.sum()


def make_criterion(hparams):
    """Loss selection of every reference model (anat_cnn.py:81-85): focal iff hparams['fl_gamma'] is truthy."""
    if "fl_gamma" in hparams and hparams["fl_gamma"]:
        return FocalLossOracle(gamma=hparams["fl_gamma"])
    return torch.nn.CrossEntropyLoss(weight=hparams["loss_class_weights"])


def focal_grad_closed_form(logits, target, gamma):
    """(1-pt)^gamma (softmax - onehot) / N — the gradient implied by the detached factor."""
    p = F.softmax(logits, dim=1)
    pt = p.gather(1, target.view(-1, 1))
    onehot = F.one_hot(target, logits.shape[1]).to(logits)
    return (1 - pt) ** gamma * (p - onehot) / logits.shape[0]

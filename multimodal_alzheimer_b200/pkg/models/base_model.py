"""Base_Model: the LightningModule surface of the path (reference pkg/models/base_model.py:11-89).

Keeps `__init__(hparams)`, `forward`, `general_step(batch, idx, mode) -> {'loss','outputs','labels'}`,
`training_step/validation_step/test_step/predict_step`, `configure_optimizers`, `hparams`, `save`.
`training_/validation_/test_epoch_end` log the reference's keys (`{train,val,test}_loss_epoch`, macro / per-class F1,
`step`: `val_loss_epoch` is what its ReduceLROnPlateau, EarlyStopping and ModelCheckpoint monitor); instead of per-step
torchmetrics objects the epoch's F1 / MCC / bootstrap intervals / confusion matrix come from one kernel launch over the
concatenated step outputs.  Confusion-matrix PLOTS are logging, not path arithmetic.  When pytorch_lightning is importable the class derives from
pl.LightningModule; otherwise from a small stand-in with the same few methods.
"""
from abc import ABC, abstractmethod

import torch
import torch.nn as nn

try:  # pragma: no cover - pytorch_lightning is not in the build image
    import pytorch_lightning as pl

    _LightningBase = pl.LightningModule
except Exception:  # noqa: BLE001
    pl = None

    class _Hparams(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

    class _LightningBase(nn.Module):
        """Minimal stand-in for pl.LightningModule (save_hyperparameters / hparams / log / load_from_checkpoint)."""

        def __init__(self):
            super().__init__()
            self._hparams = _Hparams()
            self.logged = {}

        def save_hyperparameters(self, hparams=None, ignore=None):
            hp = dict(hparams or {})
            for k in (ignore or []):
                hp.pop(k, None)
            self._hparams = _Hparams(hp)

        @property
        def hparams(self):
            return self._hparams

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, name, value, **kwargs):
            self.logged[name] = value

        def log_dict(self, d, **kwargs):
            self.logged.update(d)

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            model = cls(ckpt["hyper_parameters"], **kwargs)
            model.load_state_dict(ckpt["state_dict"])
            return model

        def save_checkpoint(self, path):
            torch.save({"state_dict": self.state_dict(), "hyper_parameters": dict(self.hparams)}, path)


class Base_Model(_LightningBase, ABC):
    def __init__(self, hparams, gpu_id=None):
        super().__init__()
        self.save_hyperparameters(hparams, ignore=["gpu_id"])
        if hparams["n_classes"] == 3:
            self.label_ind_by_names = {"CN": 0, "MCI": 1, "AD": 2}
        else:
            self.label_ind_by_names = {"CN": 0, "AD": 1}

    @abstractmethod
    def forward(self, x):
        pass

    @property
    def is_cuda(self):
        return next(self.parameters()).is_cuda

    def save(self, path):
        print("Saving model... %s" % path)
        torch.save(self, path)

    @abstractmethod
    def general_step(self, batch, batch_idx, mode) -> dict:
        pass

    def training_step(self, batch, batch_idx):
        return self.general_step(batch, batch_idx, "train")

    def validation_step(self, batch, batch_idx):
        return self.general_step(batch, batch_idx, "val")

    def test_step(self, batch, batch_idx):
        return self.general_step(batch, batch_idx, "test")

    def predict_step(self, batch, batch_idx):
        return self.general_step(batch, batch_idx, "pred")

    @abstractmethod
    def configure_optimizers(self):
        pass

    # ------------------------------------------------------------------ epoch-end hooks (base_model.py:91-172)
    def _epoch_end(self, mode, step_outputs):
        """`{mode}_loss_epoch` (the key ReduceLROnPlateau / EarlyStopping / ModelCheckpoint monitor,
        anat_cnn.py:131-135, train_anat_cnn.py:206-211), macro and per-class F1 over the epoch, `step`.  The reference
        accumulates torchmetrics states step by step; the same totals come from one confusion-matrix kernel launch over
        the epoch's concatenated outputs.  The confusion-matrix image is logging, not path arithmetic: the matrix is
        kept as `self.last_confusion_matrix[mode]`."""
        from ... import kernels as K
        avg_loss = torch.stack([x["loss"].detach() for x in step_outputs]).mean()
        y_hat = torch.cat([x["outputs"] for x in step_outputs]).detach().to(torch.float64).contiguous()
        y = torch.cat([x["labels"] for x in step_outputs]).contiguous()
        whole = K.bootstrap_metrics(y_hat, y, None, want_confmat=True)
        log = {f"{mode}_loss_epoch": avg_loss, f"{mode}_f1_epoch": whole["f1"][0],
               "step": float(getattr(self, "current_epoch", 0))}
        for i in range(self.hparams["n_classes"]):
            log[f"{mode}_f1_epoch_class_{i}"] = whole["f1_class"][0, i]
        self.log_dict(log)
        if not hasattr(self, "last_confusion_matrix"):
            self.last_confusion_matrix = {}
        self.last_confusion_matrix[mode] = whole["confmat"][0].cpu()
        return log

    def training_epoch_end(self, training_step_outputs):
        self._epoch_end("train", training_step_outputs)

    def validation_epoch_end(self, validation_step_outputs):
        self._epoch_end("val", validation_step_outputs)

    def test_epoch_end(self, outputs):
        log = self.test_epoch_metrics(outputs)
        self.last_confusion_matrix = dict(getattr(self, "last_confusion_matrix", {}), test=log.pop("confusion_matrix"))
        self.log_dict(log)

    # ------------------------------------------------------------------ test-epoch metrics
    def bootstrap_metric(self, metric, y_hat, y_labels, n_drawings=1000):
        """base_model.py:212-236: mean and 1.96 x std of `metric` ('f1' = MulticlassF1Score macro, 'mcc' =
        MulticlassMatthewsCorrCoef) over `n_drawings` resamples with replacement.  The draws are the reference's
        (`torch.randint(0, n, (n,))` per drawing, CPU generator, same order); all of them are evaluated by ONE kernel
        launch instead of 1000 torchmetrics update/compute/reset rounds."""
        from ... import kernels as K
        n = len(y_hat)
        draws = torch.stack([torch.randint(0, n, (n,)) for _ in range(n_drawings)]).to(y_hat.device)
        values = K.bootstrap_metrics(y_hat.contiguous(), y_labels.contiguous(), draws)[metric].cpu()
        return torch.mean(values), 1.96 * torch.std(values)

    def test_epoch_metrics(self, outputs):
        """The numbers `test_epoch_end` logs (base_model.py:134-172): loss, macro / per-class F1 over the test set,
        bootstrap mean and confidence interval of F1 and MCC, plus the confusion matrix the reference plots (as a
        tensor: plotting is not part of this package)."""
        from ... import kernels as K
        avg_loss = torch.stack([x["loss"] for x in outputs]).mean()
        y_hat = torch.cat([x["outputs"] for x in outputs]).detach().contiguous()
        y_labels = torch.cat([x["labels"] for x in outputs]).contiguous()
        whole = K.bootstrap_metrics(y_hat, y_labels, None, want_confmat=True)
        log = {"test_loss_epoch": avg_loss, "test_f1_epoch": whole["f1"][0].cpu(), "step": float(getattr(self, "current_epoch", 0))}
        for i in range(self.hparams["n_classes"]):
            log[f"test_f1_epoch_class_{i}"] = whole["f1_class"][0, i].cpu()
        log["test_f1_epoch_boot"], log["test_f1_epoch_ci"] = self.bootstrap_metric("f1", y_hat, y_labels)
        log["test_mcc_epoch_boot"], log["test_mcc_epoch_ci"] = self.bootstrap_metric("mcc", y_hat, y_labels)
        log["confusion_matrix"] = whole["confmat"][0].cpu()
        return log


def volume_input(x):
    """batch['mri'] / batch['pet1451'] (B, D, H, W), fp64 or fp32 -> (B, 1, D, H, W).  The reference then casts
    `.to(float32)` on the device (anat_cnn.py:100-103); here the cast (fp64 -> fp32 -> bf16) happens inside the
    first CUDA kernel of the encoder.  bf16 volumes produced by pkg/utils/normalization.py (already the encoder's
    input type) become NDHWC (B, D, H, W, 1)."""
    if x.dtype == torch.bfloat16:
        return x.unsqueeze(-1)
    return x.unsqueeze(1)


def adam_or_plateau(hparams, parameters_optim, **adam_kwargs):
    """Adam (+ optional ReduceLROnPlateau on val_loss_epoch), as every configure_optimizers of the path ends
    (anat_cnn.py:127-136).  The optimizer is a torch.optim.Adam subclass (same groups, state names, state_dict)
    whose step() is the multi-tensor CUDA kernel of csrc/optimizer.cu."""
    from ...optim import Adam
    optimizer = Adam(parameters_optim, **adam_kwargs)
    if hparams.get("reduce_factor_lr_schedule"):
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, factor=hparams["reduce_factor_lr_schedule"])
        return {"optimizer": optimizer, "lr_scheduler": scheduler, "monitor": "val_loss_epoch"}
    return optimizer

"""Layer stacks shared by the small-CNN models of the path, described once.

The reference repeats the same two stacks in three files - Small_PET_CNN (pet_models/pet_cnn.py:14-45), PET_MRI_EF
(fusion_models/early_fusion.py:29-63) and the two backbones of PET_MRI_FMF (anat_pet_featuremapfusion.py:36-64):
  stem stack   per (conv_out[i], filter_size[i]):  Conv3d 'same' (bias) -> [BatchNorm3d] -> ReLU -> MaxPool3d(2) -> [Dropout]
  dense tail   AdaptiveAvgPool3d(1) -> Flatten -> [[Dropout] -> Linear(linear_out) -> ReLU] -> Linear(n_classes)
Module ORDER is part of the checkpoint format (`model.<index>.weight` keys), so both builders emit exactly the
reference's sequence for a given hparams dict; tests/test_host_logic.py compares the resulting state_dict keys with the
reference's own classes."""
from ... import nn as bnn


def _flag(hparams, key):
    return key in hparams and bool(hparams[key])


def stem_stack(hparams, in_channels, norm_key="batchnorm"):
    """-> (list of layers, channels of the last conv)."""
    layers, width = [], in_channels
    for out_channels, kernel in zip(hparams["conv_out"], hparams["filter_size"]):
        stage = [bnn.Conv3d(width, out_channels, kernel, padding="same")]
        if _flag(hparams, norm_key):
            stage.append(bnn.BatchNorm3d(out_channels))
        stage += [bnn.ReLU(), bnn.MaxPool3d(2)]
        if "dropout_conv_p" in hparams:
            stage.append(bnn.Dropout(p=hparams["dropout_conv_p"]))
        layers += stage
        width = out_channels
    return layers, width


def dense_tail(hparams, width):
    """Global pooling and the classifier.  Without `linear_out` the last Linear reads the conv width directly."""
    layers = [bnn.AdaptiveAvgPool3d(1), bnn.Flatten()]
    if _flag(hparams, "linear_out"):
        hidden = hparams["linear_out"]
        if "dropout_dense_p" in hparams:
            layers.append(bnn.Dropout(p=hparams["dropout_dense_p"]))
        layers += [bnn.Linear(width, hidden), bnn.ReLU()]
        width = hidden
    layers.append(bnn.Linear(width, hparams["n_classes"]))
    return layers


class AlwaysFirstClass:
    """Mixin of the reference's `Random_Benchmark_All_CN` baselines (pet_cnn.py:85-90, early_fusion.py:112-117): the
    logits of the wrapped model replaced by a one-hot on class 0."""

    def forward(self, x):
        import torch
        logits = torch.zeros_like(super().forward(x))
        logits[..., 0] = 1
        return logits

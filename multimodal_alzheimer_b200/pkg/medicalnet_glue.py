"""Glue replacing `parse_opts()` + `generate_model(opts)` + `.module` of the reference (anat_cnn.py:18-31)."""
import os

import torch

from ..medicalnet import generate_model


def build_encoder(depth, pretrain_path=None):
    model = generate_model(depth)
    if pretrain_path and os.path.exists(pretrain_path):
        # upstream checkpoints hold {'state_dict': {'module.<key>': tensor}} (nn.DataParallel prefix)
        state = torch.load(pretrain_path, map_location="cpu", weights_only=False)
        state = state.get("state_dict", state)
        state = {k[len("module."):] if k.startswith("module.") else k: v for k, v in state.items()}
        own = model.state_dict()
        model.load_state_dict({k: v for k, v in state.items() if k in own and v.shape == own[k].shape}, strict=False)
    return model

"""Adam for the training path: `torch.optim.Adam`'s constructor, param groups, state names and `state_dict()`, with
`step()` executed by the multi-tensor CUDA kernel of csrc/optimizer.cu (`adni_adam_step_multi`).

Every `configure_optimizers` of the reference ends in `torch.optim.Adam(parameters_optim, weight_decay=l2_reg)` with
ONE param group per tensor (pkg/models/mri_models/anat_cnn.py:111-128, fusion_models/anat_pet_fusion.py:94-118,
fusion_models/all_modalities_fusion.py:98-128): 60-320 groups, which makes the stock per-tensor optimizer
launch-bound.  Here the whole model is updated by ceil(n_tensors / 60) launches.  Learning rate and
weight decay are read by the kernel from a small DEVICE table that every `step()` refreshes from `param_groups` through a
pinned host buffer (one asynchronous copy): eagerly `ReduceLROnPlateau` just works; under a captured CUDA graph the copy
is a graph node that re-reads the pinned buffer on every replay, so after a scheduler step call
`refresh_hyperparameters()` (host only) and the next replay trains with the new rates.

No CPU path: stepping parameters that are not CUDA fp32 tensors raises (north_star: no CPU fallback).
"""
import ctypes

import torch

from . import _lib
from . import kernels as K


class Adam(torch.optim.Adam):
    """Drop-in for `torch.optim.Adam(params, lr, betas, eps, weight_decay)` (amsgrad / maximize are not used by the
    reference and are rejected).  State per parameter: 'step' (0-d fp32 device tensor), 'exp_avg', 'exp_avg_sq' -
    the names and dtypes torch uses with capturable=True, so checkpoints interchange with torch.optim.Adam."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, **unsupported):
        for k, v in unsupported.items():
            if v:
                raise NotImplementedError(f"multimodal_alzheimer_b200.optim.Adam: {k}={v!r} is not on the reference path")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self._tables = {}  # (group key, active tensor ids) -> ctypes tables that do not change between steps

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}  # the moment tensors were replaced

    def _state_for(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        elif not torch.is_tensor(st["step"]) or st["step"].device != p.device or st["step"].dtype != torch.float32:
            # a state_dict written by stock torch.optim.Adam (host-side step numbers)
            st["step"] = torch.tensor(float(st["step"]), dtype=torch.float32, device=p.device)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # tensors that share (beta1, beta2, eps) go into one call; lr and weight decay are per tensor
        calls = {}
        for group in self.param_groups:
            if group.get("amsgrad") or group.get("maximize"):
                raise NotImplementedError("amsgrad / maximize are not on the reference path")
            lr = group["lr"]
            if torch.is_tensor(lr):
                raise NotImplementedError("tensor learning rates are not supported (lr rides in the kernel parameters)")
            key = (float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]))
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.AdniError("optim.Adam steps contiguous fp32 CUDA parameters only (no CPU fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients")
                calls.setdefault(key, []).append((p, float(lr), float(group["weight_decay"])))
        for (beta1, beta2, eps), items in calls.items():
            n = len(items)
            # keyed by address: a re-allocated parameter (.to(), load) gets a fresh table, its moments live in self.state
            tkey = (beta1, beta2, eps, tuple(p.data_ptr() for p, _, _ in items))
            tab = self._tables.get(tkey)
            if tab is None:
                states = [self._state_for(p) for p, _, _ in items]
                VP, LL, FL = ctypes.c_void_p * n, ctypes.c_longlong * n, ctypes.c_float * n
                tab = {
                    "p": VP(*[p.data_ptr() for p, _, _ in items]),
                    "m": VP(*[s["exp_avg"].data_ptr() for s in states]),
                    "v": VP(*[s["exp_avg_sq"].data_ptr() for s in states]),
                    "t": VP(*[s["step"].data_ptr() for s in states]),
                    "n": LL(*[p.numel() for p, _, _ in items]),
                    "g": VP(),
                    "lr": FL(),
                    "wd": FL(),
                    "hyper_host": torch.empty((2, n), dtype=torch.float32).pin_memory(),
                    "hyper_dev": torch.empty((2, n), dtype=torch.float32, device=items[0][0].device),
                    "hyper_np": None,
                    "bytes": 28 * sum(p.numel() for p, _, _ in items),  # 16 B read + 12 B written per parameter
                    "states": states,  # keeps the moment tensors of the table alive
                }
                if len(self._tables) > 8:
                    self._tables = {}
                self._tables[tkey] = tab
            grads = []
            if tab["hyper_np"] is None:
                tab["hyper_np"] = tab["hyper_host"].numpy()     # shares the pinned memory
            hyper = tab["hyper_np"]
            for i, (p, lr, wd) in enumerate(items):
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.to(torch.float32).contiguous()
                grads.append(g)
                tab["g"][i] = g.data_ptr()
                tab["lr"][i] = lr
                tab["wd"][i] = wd
                hyper[0, i] = lr
                hyper[1, i] = wd
            tab["hyper_dev"].copy_(tab["hyper_host"], non_blocking=True)   # captured: re-read on every graph replay
            K.call_hbm("hbm_adam", tab["bytes"], "adni_adam_step_multi", n, tab["p"], tab["g"], tab["m"], tab["v"],
                       tab["t"], tab["n"], tab["lr"], tab["wd"], _lib.ptr(tab["hyper_dev"]), beta1, beta2, eps,
                       _lib.stream_ptr())
            for p, _, _ in items:   # the kernel wrote through raw pointers: tell autograd / the weight arena
                torch.autograd.graph.increment_version(p)
            del grads
        return loss

    def refresh_hyperparameters(self):
        """Host-only: rewrite the pinned lr / weight-decay tables from `param_groups` (after a scheduler step).  A
        captured training graph picks the new values up on its next replay; eager `step()` does this itself."""
        for tkey, tab in self._tables.items():
            ptrs = {ptr_: i for i, ptr_ in enumerate(tkey[3])}
            for group in self.param_groups:
                for p in group["params"]:
                    i = ptrs.get(p.data_ptr())
                    if i is not None:
                        tab["hyper_host"][0, i] = float(group["lr"])
                        tab["hyper_host"][1, i] = float(group["weight_decay"])

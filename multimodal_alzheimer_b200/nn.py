"""nn.Module surface of the path: drop-in counterparts of the torch.nn layers the reference composes
(Conv3d, BatchNorm3d, ReLU, MaxPool3d, AdaptiveAvgPool3d, Flatten, Linear, BatchNorm1d, Dropout, Sequential),
with torch-identical parameter names/shapes (state_dict compatible) and hand-written CUDA underneath.

Activations between volume layers are bf16 NDHWC tensors [N, D, H, W, C]; a 5-D fp32/fp64 tensor entering a
volume layer is taken to be the reference's (B, 1, D, H, W) NCDHW input and is cast on the device.
`Sequential` fuses the patterns the reference builds (Conv3d -> BatchNorm3d -> ReLU, Linear -> [BatchNorm1d] ->
ReLU) into single kernels and keeps slicing semantics (`model[:-1]`, `conv_seg[:2]`,
pkg/models/fusion_models/anat_pet_fusion.py:28-32).
"""
import math

import torch
import torch.nn as tnn

from . import autograd as A

BF16 = torch.bfloat16


def as_volume(x):
    """Accept the reference's NCDHW fp32 input at the boundary; pass bf16 NDHWC tensors through."""
    if x.dtype == BF16:
        return x
    return A.InputToVolume.apply(x)


class Conv3d(tnn.Module):
    """torch.nn.Conv3d(in, out, k, stride, padding, dilation, bias) with isotropic geometry.
    padding may be an int or 'same' (pkg/models/pet_models/pet_cnn.py:21); an even kernel with 'same'
    (filter_size_fusion = 4, train_anat_pet_featuremapfusion.py:70) pads like torch: lo = total // 2, hi = total - lo."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True):
        super().__init__()

        def _iso(v, name):
            if isinstance(v, (tuple, list)):
                if len(set(v)) != 1:
                    raise ValueError(f"Conv3d: anisotropic {name} {v} is not supported")
                return int(v[0])
            return v

        k, stride, dilation = _iso(kernel_size, "kernel_size"), _iso(stride, "stride"), _iso(dilation, "dilation")
        padding = _iso(padding, "padding")
        self.pad_high_extra = 0
        if padding == "same":
            if stride != 1:
                raise ValueError("padding='same' is not supported for strided convolutions")
            total = dilation * (k - 1)            # torch: lo = total // 2, hi = total - lo (the odd voxel goes high)
            padding = total // 2
            self.pad_high_extra = total - 2 * padding
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation = (k,) * 3, (stride,) * 3, (padding,) * 3, (dilation,) * 3
        self.cfg = A.ConvCfg(k, stride, padding, dilation)
        self.weight = tnn.Parameter(torch.empty(out_channels, in_channels, k, k, k))
        self.bias = tnn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):  # torch.nn.Conv3d default init
        tnn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * self.cfg.k ** 3
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            tnn.init.uniform_(self.bias, -bound, bound)

    def forward_with_stats(self, x, want_stats=True):
        x = as_volume(x)
        if self.pad_high_extra:  # even kernel with padding='same': one more zero voxel on the high side of each axis
            x = A.PadHighFn.apply(x, self.pad_high_extra)
        return A.Conv3dFn.apply(x, self.weight, self.bias, self.cfg, want_stats)

    def forward(self, x):
        return self.forward_with_stats(x, False)[0]

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, dilation={self.dilation}, bias={self.bias is not None}")


class _BatchNorm(tnn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = tnn.Parameter(torch.ones(num_features))
        self.bias = tnn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def state(self):
        return A.BNState(self)

    def extra_repr(self):
        return f"{self.num_features}, eps={self.eps}, momentum={self.momentum}"


class BatchNorm3d(_BatchNorm):
    def forward(self, x, stats=None, residual=None, relu=False):
        return A.BatchNormActFn.apply(as_volume(x), stats, self.weight, self.bias, residual, self.state(), relu)


class BatchNorm1d(_BatchNorm):
    def forward(self, x, relu=False):
        return A.BatchNorm1dFn.apply(x, self.weight, self.bias, self.state(), relu)


class ReLU(tnn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        if x.dtype == BF16:
            return A.ReluFn.apply(x)
        # fp32 feature vectors reach a stand-alone ReLU only outside the fused Sequential patterns
        return A.ReluF32Fn.apply(x.to(torch.float32))


class MaxPool3d(tnn.Module):
    def __init__(self, kernel_size, stride=None, padding=0):
        super().__init__()
        k = kernel_size[0] if isinstance(kernel_size, (tuple, list)) else kernel_size
        s = k if stride is None else (stride[0] if isinstance(stride, (tuple, list)) else stride)
        p = padding[0] if isinstance(padding, (tuple, list)) else padding
        self.kernel_size, self.stride, self.padding = k, s, p

    def forward(self, x):
        return A.MaxPoolFn.apply(as_volume(x), self.kernel_size, self.stride, self.padding)

    def extra_repr(self):
        return f"kernel_size={self.kernel_size}, stride={self.stride}, padding={self.padding}"


class AdaptiveAvgPool3d(tnn.Module):
    def __init__(self, output_size=1):
        super().__init__()
        if output_size not in (1, (1, 1, 1)):
            raise NotImplementedError("only AdaptiveAvgPool3d(1) (global average pooling) is on the path")

    def forward(self, x):
        return A.GapFn.apply(as_volume(x))


class Flatten(tnn.Module):
    def forward(self, x):
        return x.reshape(x.shape[0], -1)


class Linear(tnn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = tnn.Parameter(torch.empty(out_features, in_features))
        self.bias = tnn.Parameter(torch.empty(out_features)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):  # torch.nn.Linear default init
        tnn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features)
            tnn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x, relu=False):
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        return A.LinearFn.apply(x, self.weight, self.bias, relu)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"


class Dropout(tnn.Module):
    """nn.Dropout(p): identity in eval mode or for p == 0; in training mode a Philox keep-mask scaled by 1/(1-p)
    (csrc/dropout.cu) that the backward pass regenerates instead of storing.  The random stream is this module's own
    (seed drawn from torch's global generator at first use, i.e. it follows `torch.manual_seed` /
    `pl.seed_everything`; a device-side call counter), not torch's CUDA generator: runs are reproducible under a seed,
    but the masks differ from stock PyTorch's (SURVEY.md App. C.9 - dropout is outside the bit-parity configs)."""

    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        if p < 0 or p > 1:
            raise ValueError(f"dropout probability has to be between 0 and 1, but got {p}")
        self.p = p
        self._seed = None
        self._counter = None

    def forward(self, x):
        if not self.training or self.p == 0:
            return x
        if self.p == 1:
            return x * 0
        if self._seed is None:
            self._seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if self._counter is None or self._counter.device != x.device:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("Dropout: run one eager training step before capturing a CUDA graph")
            self._counter = torch.zeros(1, dtype=torch.int64, device=x.device)
        if x.dtype not in (BF16, torch.float32):
            x = as_volume(x) if x.dim() == 5 else x.to(torch.float32)
        return A.DropoutFn.apply(x, self.p, self._seed, self._counter)

    def extra_repr(self):
        return f"p={self.p}"


class Sequential(tnn.Sequential):
    """nn.Sequential with kernel fusion over the layer patterns the reference builds; slicing keeps the type."""

    def forward(self, x):
        mods = list(self)
        i, n = 0, len(mods)
        while i < n:
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < n else None
            nxt2 = mods[i + 2] if i + 2 < n else None
            if isinstance(m, Conv3d) and isinstance(nxt, BatchNorm3d):
                y, st = m.forward_with_stats(x, True)
                if isinstance(nxt2, ReLU):
                    x = nxt(y, stats=st, relu=True)
                    i += 3
                else:
                    x = nxt(y, stats=st)
                    i += 2
            elif (isinstance(m, ReLU) and isinstance(nxt, MaxPool3d) and x.dtype == BF16 and x.dim() == 5
                  and A.K.relu_maxpool_supported(nxt.kernel_size, nxt.stride, nxt.padding)):
                x = A.ReluMaxPoolFn.apply(x, nxt.kernel_size)      # conv -> ReLU -> MaxPool3d(2): one pass each way
                i += 2
            elif isinstance(m, BatchNorm3d) and isinstance(nxt, ReLU):
                x = m(x, relu=True)
                i += 2
            elif isinstance(m, Linear) and isinstance(nxt, BatchNorm1d) and isinstance(nxt2, ReLU):
                x = nxt(m(x), relu=True)
                i += 3
            elif isinstance(m, Linear) and isinstance(nxt, ReLU):
                x = m(x, relu=True)
                i += 2
            else:
                x = m(x)
                i += 1
        return x

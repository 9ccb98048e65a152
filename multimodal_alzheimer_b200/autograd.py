"""torch.autograd.Function wrappers that sequence the CUDA kernels of the hot path.

Tensor convention inside the encoder: activations are bf16, shape [N, D, H, W, C] (NDHWC, contiguous).
Parameters stay torch fp32 nn.Parameters in the reference layout (Conv3d weight [Cout, Cin, kd, kh, kw]);
each Function converts them to the kernel layouts on the fly and returns gradients in parameter layout.

Residual blocks are ONE Function each (not a chain of small ones) so that the backward pass can fuse the
residual-branch gradient into the dgrad epilogue instead of letting autograd add bf16 tensors with a torch
kernel, and so that every intermediate is freed as soon as its consumer has run.

Data parallelism: when torch.distributed is initialised with world_size > 1 the BatchNorm
statistic sums (forward: sum x, sum x^2; backward: sum g, sum g*xhat) and the loss normaliser are
all-reduced so that N ranks on shards of a batch reproduce the single-process full-batch step.
"""
import os

import torch
import torch.distributed as dist

from . import data_parallel as _dp
from . import kernels as K

BF16 = torch.bfloat16

# BatchNorm-backward sums fused into the epilogue of the dgrad that produces the BatchNorm's output gradient
# (adni_conv3d_dgrad_bnred): on by default, ADNI_FUSE_BN_REDUCE=0 restores the separate reduction kernel.
_FUSE_BNRED = os.environ.get("ADNI_FUSE_BN_REDUCE", "1") != "0"
# sums a block's last dgrad computed for the PRECEDING block's final BatchNorm, keyed by the address of the gradient
# tensor it hands to that block (consumed - popped - by that block's backward within the same backward pass)
_PENDING_RED = {}


# --------------------------------------------------------------------------------------------- DP helpers
def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


# The process group used by the collectives inside the Functions.  Fusion models run their two encoder branches on
# two CUDA streams; giving each branch its own NCCL communicator keeps one branch's (compute-dependent) reductions
# from queueing behind the other's.  Forward code selects it with `dp_group(...)`; every Function remembers the
# group of its forward pass for its backward pass.
_GROUP = [None]


class dp_group:
    def __init__(self, group):
        self.group = group

    def __enter__(self):
        self.prev, _GROUP[0] = _GROUP[0], self.group

    def __exit__(self, *exc):
        _GROUP[0] = self.prev


def _allreduce_(t):
    """Sum over the data-parallel ranks: the one-shot NVLink kernel for the small fp64 statistic vectors
    (data_parallel.PeerReducer), NCCL otherwise."""
    if _world() > 1:
        if t.is_cuda:
            from . import data_parallel as dp
            red = dp.peer_reducer(_GROUP[0], t.device)
            if red is not None and red.usable(t):
                return red.all_reduce_(t)
        dist.all_reduce(t, group=_GROUP[0])
    return t


def _with_forward_group(backward):
    """Decorator for Function.backward: run the backward collectives on the group the forward used."""
    def wrapped(ctx, *grads):
        with dp_group(getattr(ctx, "dp_group", None)):
            return backward(ctx, *grads)
    return wrapped


# BatchNorm.num_batches_tracked counters: inside `defer_bn_counters()` (an encoder's forward) the +1 of every layer is
# collected and applied by one multi-tensor add at exit instead of one tiny kernel per BatchNorm layer.
_DEFERRED_COUNTERS = [None]


class defer_bn_counters:
    def __enter__(self):
        self.prev, _DEFERRED_COUNTERS[0] = _DEFERRED_COUNTERS[0], []

    def __exit__(self, *exc):
        pending, _DEFERRED_COUNTERS[0] = _DEFERRED_COUNTERS[0], self.prev
        if pending:
            torch._foreach_add_(pending, 1)


def _count_batch(bn):
    if bn.module is not None and bn.module.num_batches_tracked is not None:
        if _DEFERRED_COUNTERS[0] is not None:
            _DEFERRED_COUNTERS[0].append(bn.module.num_batches_tracked)
        else:
            bn.module.num_batches_tracked += 1


class BNState:
    """Non-tensor BatchNorm configuration handed to the Functions (running buffers are updated in place)."""

    def __init__(self, module):
        self.eps = module.eps
        self.momentum = module.momentum
        self.running_mean = module.running_mean
        self.running_var = module.running_var
        self.training = module.training
        self.module = module


def _bn_forward(y, stats, gamma, beta, bn, residual, relu):
    """Shared BN forward. Returns (out, bnp, count) with bnp = fp32 [4, C] rows (mean, invstd, scale, shift);
    stats may be None (computed with a reduction pass)."""
    C = y.shape[-1]
    rows = y.numel() // C
    if bn.training:
        if stats is None:
            stats = K.channel_stats(y.view(-1, C))
        _allreduce_(stats)
        count = rows * _world()
        _count_batch(bn)
        # finalize (scale / shift / running statistics) and apply in one launch
        out, bnp = K.bn_train_apply(y, stats, count, gamma, beta, bn.eps, bn.momentum, bn.running_mean, bn.running_var,
                                    residual=residual, relu=relu)
        return out, bnp, count
    scale, shift = K.bn_eval_params(bn.running_mean, bn.running_var, gamma, beta, bn.eps)
    out = K.bn_apply(y, scale, shift, residual=residual, relu=relu)
    return out, None, rows


def _bn_backward(dout, out, y, bnp, gamma, count, relu, want_dres, want_pg, red=None):
    """out=None with relu=True: the forward had no residual, the ReLU mask is recomputed from y*scale+shift.
    `red`: the sums [sum g | sum g*y] a dgrad epilogue already accumulated for this layer (no reduction pass then)."""
    if bnp is None:
        raise NotImplementedError("BatchNorm backward in eval mode is outside the training hot path")
    mean, invstd, scale, shift = bnp[0], bnp[1], bnp[2], bnp[3]
    if not relu or out is not None:
        scale = shift = None
    red_form = 1
    if red is None:
        red = K.bn_bwd_reduce(dout, out, y, mean, invstd, relu, scale, shift)
        red_form = 0
    _allreduce_(red)
    # dgamma / dbeta are written by the apply kernel from the all-reduced sums scaled by 1 / world_size: the gradient
    # all-reduce that follows the backward pass SUMS the ranks' parameter gradients, which restores the global sums
    # (handing every rank the full sums would count them world_size times).
    dy, dres, dgamma, dbeta = K.bn_bwd_apply(dout, out, y, mean, invstd, gamma, red, count, relu, want_dres, want_pg,
                                             scale, shift, param_grad_scale=1.0 / _world(), red_form=red_form)
    return dy, dres, dgamma, dbeta


def _dgrad_bnred(dy, w_ito, in_shape, cfg, bn_y, bnp, relu_out=None, addend=None):
    """dx = dgrad(dy) (+ addend) and, fused into its epilogue when the tcgen05 engines take the shape, the backward
    sums of the conv -> bn -> relu layer whose output dx is the gradient of: (dx, red | None).  relu_out: that layer's
    stored output (mask = > 0); None = the mask is recomputed from bn_y * scale + shift."""
    if (_FUSE_BNRED and bnp is not None and bn_y is not None
            and K.dgrad_bnred_profitable(in_shape, dy.shape[-1], cfg.k, cfg.stride, cfg.pad, cfg.dil)):
        scale, shift = (None, None) if relu_out is not None else (bnp[2], bnp[3])
        return K.conv3d_dgrad_bnred(dy, w_ito, in_shape, cfg.k, cfg.stride, cfg.pad, cfg.dil, bn_y, bn_relu_out=relu_out,
                                    bn_scale=scale, bn_shift=shift, addend=addend)
    return K.conv3d_dgrad(dy, w_ito, in_shape, cfg.k, cfg.stride, cfg.pad, cfg.dil, addend=addend), None


def _block_input_grad(ctx_tail, x, dy1, w1_ito, c1, addend):
    """The gradient a residual block returns for its input x.  When x is the output of a preceding block (its final
    bn -> + residual -> relu is described by ctx_tail = (y_last, bnp_last)), that BatchNorm's backward sums ride in the
    epilogue and are parked for the preceding block's backward."""
    tail_y, tail_p = ctx_tail
    dx, red = _dgrad_bnred(dy1, w1_ito, tuple(x.shape), c1, tail_y, tail_p, relu_out=x, addend=addend)
    if red is not None:
        _PENDING_RED[dx.data_ptr()] = (red, dx.numel())
    return dx


def _take_pending_red(dout):
    hit = _PENDING_RED.pop(dout.data_ptr(), None)
    if hit is not None and hit[1] == dout.numel():
        return hit[0]
    return None


class ConvCfg:
    def __init__(self, k, stride, pad, dil):
        self.k, self.stride, self.pad, self.dil = k, stride, pad, dil


def _conv_fwd(x, w, cfg, bias=None, stats=True, need_ito=True):
    oti, ito = K.kernel_layout(w, want_ito=need_ito)
    y, st = K.conv3d_fprop(x, oti, bias, cfg.k, cfg.stride, cfg.pad, cfg.dil, stats=stats)
    return y, st, ito


class _WShape(tuple):
    """A conv weight's shape plus the address of the parameter's storage: the backward pass asks the data-parallel
    gradient buckets (data_parallel.grad_slot) for the slot this weight's gradient should be written into."""
    ptr = None


def _ws(w):
    s = _WShape(w.shape)
    s.ptr = w.data_ptr()
    return s


def _conv_wgrad(x, dy, cfg, wshape, want_dbias=False):
    dw, db = K.conv3d_wgrad(x, dy, cfg.k, cfg.stride, cfg.pad, cfg.dil, want_dbias=want_dbias)
    ptr = getattr(wshape, "ptr", None)
    slot = _dp.grad_slot(ptr) if ptr is not None else None
    return K.wgrad_to_param_layout(dw, tuple(wshape), dst=slot), db


# --------------------------------------------------------------------------------------------- input cast
class InputToVolume(torch.autograd.Function):
    """(B,C,D,H,W) fp32/fp64 NCDHW module input -> bf16 NDHWC.  C == 1 (every encoder of the path) has the same memory
    order and is a plain cast; C in {2,3,4} is the early-fusion stack of modalities (early_fusion.py:84-88)."""

    @staticmethod
    def forward(ctx, x):
        if x.dim() != 5:
            raise NotImplementedError("volume inputs must be (B, C, D, H, W)")
        N, C, D, H, W = x.shape
        if C == 1:
            return K.cast_to_bf16(x.contiguous()).view(N, D, H, W, 1)
        return K.volumes_to_ndhwc(x)

    @staticmethod
    def backward(ctx, g):
        return None


class PadHighFn(torch.autograd.Function):
    """Extra high-side zero voxels of padding='same' with an even kernel (torch pads lo = total//2, hi = total - lo);
    the gradient is the leading box of the padded gradient."""

    @staticmethod
    def forward(ctx, x, extra):
        ctx.extra = extra
        return K.pad_volume_high(x, extra)

    @staticmethod
    def backward(ctx, g):
        return K.crop_volume_high(g.contiguous(), ctx.extra), None


class MaxOutFn(torch.autograd.Function):
    """torch.max(torch.stack((a, b), dim=0), dim=0)[0] on two feature maps (anat_pet_featuremapfusion.py:121-123);
    ties route the gradient to `a` like torch's first-index rule."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return K.maxout_fwd(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da, db = K.maxout_bwd(g.contiguous(), a, b, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return da, db


class ConcatChannelsFn(torch.autograd.Function):
    """torch.cat((a, b), dim=1) of the reference's NCDHW feature maps = concatenation of the NDHWC channel axis
    (anat_pet_featuremapfusion.py:118-119)."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.widths = (a.shape[-1], b.shape[-1])
        return K.concat_channels(a, b)

    @staticmethod
    def backward(ctx, g):
        da, db = K.split_channels(g.contiguous(), *ctx.widths, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return da, db


# --------------------------------------------------------------------------------------------- stem
class StemFn(torch.autograd.Function):
    """conv1 (7x7x7, s2) -> bn1 -> ReLU -> MaxPool3d(3,2,1)  (MedicalNet ResNet.forward, first line)."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, bn, cfg, pool):
        ctx.dp_group = _GROUP[0]
        tc = K.stem_supported(x.shape[-1], w.shape[0], cfg.k, cfg.stride, cfg.pad, cfg.dil)
        if tc:  # tcgen05 path: the W-axis filter window becomes a 16-byte pixel (csrc/conv_stem.cu)
            xs = K.stem_expand(x)
            y, st = K.stem_fprop(xs, tuple(x.shape), w)
        else:
            xs = x
            y, st, _ = _conv_fwd(x, w, cfg, need_ito=False)
        fused = bn.training and K.fused_pool_supported(*pool)
        if fused:  # bn1 -> ReLU -> max-pool in one pass over y (csrc/stem_fused.cu); `a` never exists
            C = y.shape[-1]
            _allreduce_(st)
            count = (y.numel() // C) * _world()
            bnp = K.bn_finalize(st, count, gamma, beta, bn.eps, bn.momentum, bn.running_mean, bn.running_var)
            _count_batch(bn)
            p, am, yraw = K.bn_relu_maxpool_fwd(y, bnp, *pool)
            a_shape = tuple(y.shape)
        else:
            a, bnp, count = _bn_forward(y, st, gamma, beta, bn, None, True)
            p, am = K.maxpool3d_fwd(a, *pool)
            a_shape = tuple(a.shape)
            yraw = None
        ctx.save_for_backward(xs, y, am, bnp, gamma, yraw)
        ctx.a_shape = a_shape
        ctx.cfg, ctx.pool, ctx.count, ctx.wshape, ctx.tc, ctx.xshape = cfg, pool, count, _ws(w), tc, tuple(x.shape)
        ctx.fused = fused
        return p

    @staticmethod
    @_with_forward_group
    def backward(ctx, dp):
        x, y, am, bnp, gamma, yraw = ctx.saved_tensors
        dp = dp.contiguous()
        if ctx.fused:
            if yraw is not None:
                # every window hands its gradient to ONE voxel (its arg-max): the sums run over the pooled tensor
                red = K.bn_bwd_reduce(dp, None, yraw, bnp[0], bnp[1], True, bnp[2], bnp[3])
            else:
                red = K.maxpool_bn_bwd_reduce(dp, am, y, bnp, *ctx.pool)
            dgamma, dbeta = K.bn_param_grads(red)
            _allreduce_(red)
            dy = K.maxpool_bn_bwd_apply(dp, am, y, bnp, gamma, red, ctx.count, *ctx.pool)
        else:
            da = K.maxpool3d_bwd(dp, am, ctx.a_shape, *ctx.pool)
            dy, _, dgamma, dbeta = _bn_backward(da, None, y, bnp, gamma, ctx.count, True, False, True)
            del da
        if ctx.tc:
            dw = K.stem_wgrad(x, dy, ctx.xshape)
        else:
            dw, _ = _conv_wgrad(x, dy, ctx.cfg, ctx.wshape)
        return None, dw, dgamma, dbeta, None, None, None


# --------------------------------------------------------------------------------------------- residual blocks
class BasicBlockFn(torch.autograd.Function):
    """MedicalNet BasicBlock: conv3-bn-relu-conv3-bn (+ downsample(x) | x) - relu.  Returns (out, y2, bnp2): the raw
    output of conv2 and bn2's parameters describe the block's final bn -> relu to the NEXT block, whose last dgrad
    computes bn2's backward sums in its epilogue (tail_y / tail_p are the same pair of the PRECEDING block)."""

    @staticmethod
    def forward(ctx, x, w1, g1, b1, w2, g2, b2, wd, gd, bd, bn1, bn2, bnd, c1, c2, cd, tail_y=None, tail_p=None):
        ctx.dp_group = _GROUP[0]
        need_dx = ctx.needs_input_grad[0]
        y1, st1, w1_ito = _conv_fwd(x, w1, c1, need_ito=need_dx)
        a1, p1, n1 = _bn_forward(y1, st1, g1, b1, bn1, None, True)
        y2, st2, w2_ito = _conv_fwd(a1, w2, c2)
        if wd is not None:
            yd, std, wd_ito = _conv_fwd(x, wd, cd, need_ito=need_dx)
            r, pd, nd = _bn_forward(yd, std, gd, bd, bnd, None, False)
        else:
            yd = wd_ito = pd = None
            nd = 0
            r = x
        out, p2, n2 = _bn_forward(y2, st2, g2, b2, bn2, r, True)
        ctx.save_for_backward(x, y1, a1, y2, out, yd, p1, p2, pd, g1, g2, gd, w1_ito, w2_ito, wd_ito, tail_y, tail_p)
        ctx.cfg = (c1, c2, cd, n1, n2, nd, _ws(w1), _ws(w2), None if wd is None else _ws(wd), need_dx)
        if p2 is None:
            p2 = torch.empty(0, device=x.device)
        ctx.mark_non_differentiable(y2, p2)
        return out, y2, p2

    @staticmethod
    @_with_forward_group
    def backward(ctx, dout, _dy2, _dp2):
        x, y1, a1, y2, out, yd, p1, p2, pd, g1, g2, gd, w1_ito, w2_ito, wd_ito, tail_y, tail_p = ctx.saved_tensors
        c1, c2, cd, n1, n2, nd, ws1, ws2, wsd, need_dx = ctx.cfg
        dout = dout.contiguous()
        dy2, dres, dg2, db2 = _bn_backward(dout, out, y2, p2, g2, n2, True, True, True, red=_take_pending_red(dout))
        dw2, _ = _conv_wgrad(a1, dy2, c2, ws2)
        da1, red1 = _dgrad_bnred(dy2, w2_ito, tuple(a1.shape), c2, y1, p1, relu_out=a1)
        del dy2
        dy1, _, dg1, db1 = _bn_backward(da1, None, y1, p1, g1, n1, True, False, True, red=red1)
        del da1
        dw1, _ = _conv_wgrad(x, dy1, c1, ws1)
        dwd = dgd = dbd = None
        dx = None
        if yd is not None:
            dyd, _, dgd, dbd = _bn_backward(dres, None, yd, pd, gd, nd, False, False, True)
            dwd, _ = _conv_wgrad(x, dyd, cd, wsd)
            if need_dx:
                dxd = K.conv3d_dgrad(dyd, wd_ito, tuple(x.shape), cd.k, cd.stride, cd.pad, cd.dil)
                dx = _block_input_grad((tail_y, tail_p), x, dy1, w1_ito, c1, dxd)
        elif need_dx:
            dx = _block_input_grad((tail_y, tail_p), x, dy1, w1_ito, c1, dres)
        return (dx, dw1, dg1, db1, dw2, dg2, db2, dwd, dgd, dbd) + (None,) * 8


class BottleneckFn(torch.autograd.Function):
    """MedicalNet Bottleneck: 1x1-bn-relu, 3x3(stride,dil)-bn-relu, 1x1-bn (+ downsample(x) | x) - relu.
    Returns (out, y3, bnp3) / takes (tail_y, tail_p) like BasicBlockFn."""

    @staticmethod
    def forward(ctx, x, w1, g1, b1, w2, g2, b2, w3, g3, b3, wd, gd, bd, bn1, bn2, bn3, bnd, c1, c2, c3, cd, tail_y=None,
                tail_p=None):
        ctx.dp_group = _GROUP[0]
        need_dx = ctx.needs_input_grad[0]
        y1, st1, w1_ito = _conv_fwd(x, w1, c1, need_ito=need_dx)
        a1, p1, n1 = _bn_forward(y1, st1, g1, b1, bn1, None, True)
        y2, st2, w2_ito = _conv_fwd(a1, w2, c2)
        a2, p2, n2 = _bn_forward(y2, st2, g2, b2, bn2, None, True)
        y3, st3, w3_ito = _conv_fwd(a2, w3, c3)
        if wd is not None:
            yd, std, wd_ito = _conv_fwd(x, wd, cd, need_ito=need_dx)
            r, pd, nd = _bn_forward(yd, std, gd, bd, bnd, None, False)
        else:
            yd = wd_ito = pd = None
            nd = 0
            r = x
        out, p3, n3 = _bn_forward(y3, st3, g3, b3, bn3, r, True)
        ctx.save_for_backward(x, y1, a1, y2, a2, y3, out, yd, p1, p2, p3, pd, g1, g2, g3, gd, w1_ito, w2_ito, w3_ito,
                              wd_ito, tail_y, tail_p)
        ctx.cfg = (c1, c2, c3, cd, n1, n2, n3, nd, _ws(w1), _ws(w2), _ws(w3), None if wd is None else _ws(wd),
                   need_dx)
        if p3 is None:
            p3 = torch.empty(0, device=x.device)
        ctx.mark_non_differentiable(y3, p3)
        return out, y3, p3

    @staticmethod
    @_with_forward_group
    def backward(ctx, dout, _dy3, _dp3):
        (x, y1, a1, y2, a2, y3, out, yd, p1, p2, p3, pd, g1, g2, g3, gd, w1_ito, w2_ito, w3_ito, wd_ito, tail_y,
         tail_p) = ctx.saved_tensors
        c1, c2, c3, cd, n1, n2, n3, nd, ws1, ws2, ws3, wsd, need_dx = ctx.cfg
        dout = dout.contiguous()
        dy3, dres, dg3, db3 = _bn_backward(dout, out, y3, p3, g3, n3, True, True, True, red=_take_pending_red(dout))
        dw3, _ = _conv_wgrad(a2, dy3, c3, ws3)
        da2, red2 = _dgrad_bnred(dy3, w3_ito, tuple(a2.shape), c3, y2, p2, relu_out=a2)
        del dy3
        dy2, _, dg2, db2 = _bn_backward(da2, None, y2, p2, g2, n2, True, False, True, red=red2)
        del da2
        dw2, _ = _conv_wgrad(a1, dy2, c2, ws2)
        da1, red1 = _dgrad_bnred(dy2, w2_ito, tuple(a1.shape), c2, y1, p1, relu_out=a1)
        del dy2
        dy1, _, dg1, db1 = _bn_backward(da1, None, y1, p1, g1, n1, True, False, True, red=red1)
        del da1
        dw1, _ = _conv_wgrad(x, dy1, c1, ws1)
        dwd = dgd = dbd = None
        dx = None
        if yd is not None:
            dyd, _, dgd, dbd = _bn_backward(dres, None, yd, pd, gd, nd, False, False, True)
            dwd, _ = _conv_wgrad(x, dyd, cd, wsd)
            if need_dx:
                dxd = K.conv3d_dgrad(dyd, wd_ito, tuple(x.shape), cd.k, cd.stride, cd.pad, cd.dil)
                dx = _block_input_grad((tail_y, tail_p), x, dy1, w1_ito, c1, dxd)
        elif need_dx:
            dx = _block_input_grad((tail_y, tail_p), x, dy1, w1_ito, c1, dres)
        return (dx, dw1, dg1, db1, dw2, dg2, db2, dw3, dg3, db3, dwd, dgd, dbd) + (None,) * 10


# --------------------------------------------------------------------------------------------- stand-alone ops
class Conv3dFn(torch.autograd.Function):
    """Conv3d (+bias). Returns (y, stats) where stats (fp64 [2,Cout]) feeds a following BatchNorm3d."""

    @staticmethod
    def forward(ctx, x, w, bias, cfg, want_stats):
        need_dx = ctx.needs_input_grad[0]
        y, st, ito = _conv_fwd(x, w, cfg, bias=bias, stats=want_stats, need_ito=need_dx)
        ctx.save_for_backward(x, ito)
        ctx.cfg = (cfg, _ws(w), bias is not None, need_dx)
        if st is None:
            st = torch.empty(0, device=x.device)
        ctx.mark_non_differentiable(st)
        return y, st

    @staticmethod
    def backward(ctx, dy, _dst):
        x, ito = ctx.saved_tensors
        cfg, wshape, has_bias, need_dx = ctx.cfg
        dy = dy.contiguous()
        dw, db = _conv_wgrad(x, dy, cfg, wshape, want_dbias=has_bias)
        dx = K.conv3d_dgrad(dy, ito, tuple(x.shape), cfg.k, cfg.stride, cfg.pad, cfg.dil) if need_dx else None
        return dx, dw, db, None, None


class BatchNormActFn(torch.autograd.Function):
    """BatchNorm3d (batch statistics) [+ residual] [+ ReLU] on a bf16 NDHWC tensor."""

    @staticmethod
    def forward(ctx, y, stats, gamma, beta, residual, bn, relu):
        ctx.dp_group = _GROUP[0]
        st = stats if (stats is not None and stats.numel() > 0) else None
        out, bnp, count = _bn_forward(y, st, gamma, beta, bn, residual, relu)
        ctx.save_for_backward(y, out if (relu and residual is not None) else None, bnp, gamma)
        ctx.cfg = (count, relu, residual is not None)
        return out

    @staticmethod
    @_with_forward_group
    def backward(ctx, dout):
        y, out, bnp, gamma = ctx.saved_tensors
        count, relu, has_res = ctx.cfg
        dy, dres, dgamma, dbeta = _bn_backward(dout.contiguous(), out, y, bnp, gamma, count, relu, has_res, True)
        return dy, None, dgamma, dbeta, dres, None, None


class ReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = K.relu_fwd(x.contiguous())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return K.relu_bwd(dy.contiguous(), y)


class ReluF32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = K.relu_f32(x.contiguous())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return K.relu_f32(y, dy.contiguous())


class DropoutFn(torch.autograd.Function):
    """Training-mode nn.Dropout: the mask is regenerated in backward from (seed, the offset value of the forward)."""

    @staticmethod
    def forward(ctx, x, p, seed, counter):
        x = x.contiguous()
        offset = counter.clone()        # this call's stream position, kept for the backward pass
        counter += 1                    # next call (and the next CUDA-graph replay) draws a new mask
        ctx.save_for_backward(offset)
        ctx.cfg = (p, seed)
        return K.dropout(x, p, seed, offset)

    @staticmethod
    def backward(ctx, dy):
        (offset,) = ctx.saved_tensors
        p, seed = ctx.cfg
        return K.dropout(dy.contiguous(), p, seed, offset), None, None, None


class MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad):
        y, am = K.maxpool3d_fwd(x.contiguous(), k, stride, pad)
        ctx.save_for_backward(am)
        ctx.cfg = (tuple(x.shape), k, stride, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        (am,) = ctx.saved_tensors
        shape, k, stride, pad = ctx.cfg
        return K.maxpool3d_bwd(dy.contiguous(), am, shape, k, stride, pad), None, None, None


class ReluMaxPoolFn(torch.autograd.Function):
    """nn.ReLU -> nn.MaxPool3d(k) with non-overlapping windows as one pass each way (csrc/pool_small.cu)."""

    @staticmethod
    def forward(ctx, x, k):
        x = x.contiguous()
        y, am = K.relu_maxpool_fwd(x, k, relu=True)
        ctx.save_for_backward(am, y)
        ctx.cfg = (tuple(x.shape), k)
        return y

    @staticmethod
    def backward(ctx, dy):
        am, y = ctx.saved_tensors
        shape, k = ctx.cfg
        return K.relu_maxpool_bwd(dy.contiguous(), am, y, shape, k), None


class GapFn(torch.autograd.Function):
    """AdaptiveAvgPool3d(1): bf16 [N,D,H,W,C] -> fp32 [N,C,1,1,1]."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return K.gap_fwd(x.contiguous()).view(x.shape[0], x.shape[-1], 1, 1, 1)

    @staticmethod
    def backward(ctx, df):
        N, C = ctx.shape[0], ctx.shape[-1]
        return K.gap_bwd(df.reshape(N, C).contiguous(), ctx.shape)


class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, relu):
        x = x.contiguous()
        y = K.linear_fwd(x, W, b, relu)
        ctx.save_for_backward(x, W, y)
        ctx.cfg = (relu, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        relu, has_bias = ctx.cfg
        need_dx = ctx.needs_input_grad[0]
        need_dw = ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2])
        dx, dW, db = K.linear_bwd(x, W, y, dy.contiguous(), relu, need_dx=need_dx, need_dw=need_dw, has_bias=has_bias)
        return dx, dW, db if has_bias else None, None


class BatchNorm1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, bn, relu):
        ctx.dp_group = _GROUP[0]
        x = x.contiguous()
        B, C = x.shape
        if bn.training:
            st = _allreduce_(K.rows_stats_f32(x))
            count = B * _world()
            mean, invstd, scale, shift = K.bn_finalize(st, count, gamma, beta, bn.eps, bn.momentum, bn.running_mean,
                                                       bn.running_var).unbind(0)
            if bn.module is not None and bn.module.num_batches_tracked is not None:
                bn.module.num_batches_tracked += 1
        else:
            scale, shift = K.bn_eval_params(bn.running_mean, bn.running_var, gamma, beta, bn.eps)
            mean = invstd = None
            count = B
        y = K.bn1d_apply(x, scale, shift, relu)
        ctx.save_for_backward(x, y, mean, invstd, gamma)
        ctx.cfg = (count, relu)
        return y

    @staticmethod
    @_with_forward_group
    def backward(ctx, dy):
        x, y, mean, invstd, gamma = ctx.saved_tensors
        count, relu = ctx.cfg
        dy = dy.contiguous()
        red = K.bn1d_bwd_reduce(dy, y, x, mean, invstd, relu)
        dgamma, dbeta = K.bn_param_grads(red)  # local sums; the gradient all-reduce makes them global
        _allreduce_(red)
        dx, _, _ = K.bn1d_bwd_apply(dy, y, x, mean, invstd, gamma, red, count, relu)
        return dx, dgamma, dbeta, None, None


class LossFn(torch.autograd.Function):
    """Focal (detached modulating factor) / weighted cross entropy in fp64 with the GLOBAL-batch normaliser."""

    @staticmethod
    def forward(ctx, logits, target, gamma, class_weights, mean=True):
        logits = logits.contiguous()
        partial, coeff = K.loss_fwd(logits, target.contiguous(), gamma, class_weights)
        _allreduce_(partial)
        if not mean:  # size_average=False: plain sum, normaliser 1
            partial[1] = 1.0
        ctx.save_for_backward(logits, target, coeff, partial)
        return partial[0] / partial[1]

    @staticmethod
    def backward(ctx, g):
        logits, target, coeff, partial = ctx.saved_tensors
        up = g.reshape(1).to(torch.float64).contiguous()
        return K.loss_bwd(logits, target, coeff, partial[1:2], up), None, None, None, None

"""Tensor-level wrappers over the C-ABI (one Python function per kernel family).

All activation tensors are bf16 NDHWC (shape [N, D, H, W, C], contiguous).  These wrappers only allocate
outputs with torch (PyTorch owns device memory and streams — plumbing) and pass raw pointers to
libadni_b200.so; no arithmetic happens in Python and nothing here falls back to torch ops.
"""
import os

import torch

from . import _lib
from ._lib import ENGINE_AUTO, ENGINE_DIRECT, ENGINE_MMA_SYNC, ENGINE_TCGEN05, call, geom, out_extent, ptr, stream_ptr  # noqa: F401

BF16 = torch.bfloat16


class _Profile:
    """Optional per-kernel timing of the conv launches (bench.py's roofline): CUDA events recorded on the launching
    stream around each call, tagged with the kernel family and the call's algorithmic FLOPs."""

    def __init__(self):
        self.on = False
        self.records = []

    def enable(self):
        self.on = True
        self.records = []

    def disable_and_collect(self):
        self.on = False
        torch.cuda.synchronize()
        out = {}
        shapes = {}
        for tag, flops, e0, e1, shape, frac in self.records:
            ms = e0.elapsed_time(e1)
            for dd, key in ((out, tag), (shapes, f"{tag} {shape}")):
                d = dd.setdefault(key, {"flops": 0, "executed_flops": 0, "ms": 0.0, "n": 0})
                d["flops"] += flops
                d["executed_flops"] += flops * frac
                d["ms"] += ms
                d["n"] += 1
        self.shapes = shapes
        self.records = []
        return out

    def begin(self):
        if not self.on:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, e0, tag, flops, shape="", executed_fraction=1.0):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records.append((tag, flops, e0, e1, shape, executed_fraction))


PROFILE = _Profile()


def call_hbm(tag, nbytes, name, *args):
    """`call` for an HBM-bound kernel: when bench.py's profiling pass is on, the launch is bracketed by CUDA events and
    recorded under `tag` with its ALGORITHMIC bytes (the tensors the operation must read and
    write once, at their storage width) in the record's work field."""
    ev = PROFILE.begin()
    call(name, *args)
    PROFILE.end(ev, tag, nbytes)


_PLAN_CACHE = {}


def _plan(g, pass_):
    """(engine kind, executed fraction) the library's planner reports for a geometry (cached):
    0 = direct CUDA-core, 1 = tcgen05 tap-per-box, 2 = tcgen05 halo-resident, 3 = mma.sync small-channel engine."""
    key = (g.N, g.D, g.H, g.W, g.Cin, g.Cout, g.k, g.stride, g.pad, g.dil, pass_)
    hit = _PLAN_CACHE.get(key)
    if hit is None:
        import ctypes
        kind, frac = ctypes.c_int(0), ctypes.c_double(1.0)
        call("adni_conv3d_plan_info", g, pass_, ctypes.byref(kind), ctypes.byref(frac))
        hit = _PLAN_CACHE[key] = (kind.value, float(frac.value))
    return hit


def _engine_tag(g, pass_, engine):
    """(kernel-family tag, executed fraction of the algorithmic FLOPs) for bench.py's per-kernel roofline."""
    if not PROFILE.on:
        return "", 1.0
    kind, frac = _plan(g, pass_)
    if engine == ENGINE_DIRECT or kind == 0:
        return "direct", 1.0
    if engine == ENGINE_MMA_SYNC or kind == 3:
        return "tc_small", 1.0
    if pass_ == 2:
        return "tc_wgrad", frac
    return ("tc_halo" if kind == 2 else "tc_kmajor"), frac


def _chk(t, dtype, name):
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


# --------------------------------------------------------------------------------------------- conv
def weights_to_kernel_layout(w, want_ito=True):
    """fp32 [Cout, Cin, kd, kh, kw] parameter -> (bf16 [Cout, taps, Cin], bf16 [Cin, taps, Cout])."""
    w = _chk(w.detach(), torch.float32, "weight")
    cout, cin = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    oti = torch.empty((cout, taps, cin), dtype=BF16, device=w.device)
    ito = torch.empty((cin, taps, cout), dtype=BF16, device=w.device) if want_ito else None
    call("adni_weights_to_kernel_layout", ptr(w), cout, cin, taps, ptr(oti), ptr(ito), stream_ptr())
    return oti, ito


class _ZeroArena:
    """fp64 accumulators (BatchNorm sums, loss partials) handed out as slices of a few pre-zeroed chunks: one fill
    launch per 512 KB instead of one per [2, C] buffer (about 170 tiny fills per training step otherwise).  A slice is
    never handed out twice; chunks are per (device, stream) because the fill is ordered on the allocating stream, and a
    chunk zeroed outside CUDA-graph capture is never used inside one (its fill would not be part of the graph)."""

    def __init__(self, dtype=torch.float64, chunk=65536):
        self.chunks = {}
        self.dtype, self.CHUNK = dtype, chunk

    def take(self, shape, device):
        n = 1
        for d in shape:
            n *= d
        n_al = (n + 31) & ~31  # 128-byte granules (fp32) or more
        capturing = torch.cuda.is_current_stream_capturing()
        key = (device.index, stream_ptr().value)
        rec = self.chunks.get(key)
        if rec is None or rec[2] != capturing or rec[1] + n_al > rec[0].numel():
            rec = [torch.zeros(max(self.CHUNK, n_al), dtype=self.dtype, device=device), 0, capturing]
            self.chunks[key] = rec
        v = rec[0][rec[1]:rec[1] + n].view(shape)
        rec[1] += n_al
        return v


_ZEROS = _ZeroArena()


def zeros_f64(shape, device):
    return _ZEROS.take(tuple(shape), device)


# fp32 split-K wgrad accumulators: one ResNet-18 encoder needs 33 M floats per backward pass in 21 buffers; a 144 MB
# chunk turns their 21 fill launches into one (the fills are a fixed per-step cost that does not shrink with the batch)
_ZEROS_F32 = _ZeroArena(torch.float32, 36 << 20)


def zeros_f32(shape, device):
    return _ZEROS_F32.take(tuple(shape), device)


class WeightArena:
    """Kernel-layout (bf16 OTI + ITO) copies of a fixed set of conv weights, refreshed by ONE launch
    (adni_weights_to_kernel_layout_multi) instead of two transposes per conv.  The job table lives on the device and
    is rebuilt only if a parameter is re-allocated; `lookup(w)` serves the copies while the parameter's version
    counter is the one they were converted from, so a stale copy can never be used after an optimizer step."""

    def __init__(self):
        self.key = None
        self.jobs = None
        self.total_tiles = 0
        self.copies = {}     # data_ptr -> (oti, ito)
        self.versions = {}   # data_ptr -> parameter version the copies were made from

    def refresh(self, weights):
        import struct
        ws = [w for w in weights if w.dim() == 5 and w[0, 0].numel() <= 27]
        if not ws:
            return
        key = tuple((w.data_ptr(), tuple(w.shape)) for w in ws)
        if key != self.key:
            rec, tile_begin = [], 0
            self.copies = {}
            for w in ws:
                cout, cin, taps = w.shape[0], w.shape[1], w[0, 0].numel()
                oti = torch.empty((cout, taps, cin), dtype=BF16, device=w.device)
                ito = torch.empty((cin, taps, cout), dtype=BF16, device=w.device)
                self.copies[w.data_ptr()] = (oti, ito)
                tiles_ci = (cin + 15) // 16
                rec.append(struct.pack("<QQQiiiii", w.data_ptr(), oti.data_ptr(), ito.data_ptr(), cout, cin, taps,
                                       tile_begin, tiles_ci))
                tile_begin += ((cout + 15) // 16) * tiles_ci
            size = int(_lib.load().adni_weights_multi_job_bytes())
            blob = b"".join(r.ljust(size, b"\0") for r in rec)
            self.jobs = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(ws[0].device)
            self.total_tiles = tile_begin
            self.key = key
        call_hbm("hbm_weight_layouts", 8 * sum(w.numel() for w in ws), "adni_weights_to_kernel_layout_multi",
                 ptr(self.jobs), len(ws), self.total_tiles, stream_ptr())
        self.versions = {w.data_ptr(): w._version for w in ws}

    def lookup(self, w):
        hit = self.copies.get(w.data_ptr())
        if hit is not None and self.versions.get(w.data_ptr()) == w._version:
            return hit
        return None


import weakref

_ARENAS = weakref.WeakSet()  # arenas of the live encoders (consulted by kernel_layout); owned by their modules


def kernel_layout(w, want_ito=True):
    """(OTI, ITO) of a conv weight: the arena copy if an enclosing encoder refreshed one for this version of the
    parameter, else a per-tensor conversion."""
    for arena in _ARENAS:
        hit = arena.lookup(w)
        if hit is not None:
            return hit
    return weights_to_kernel_layout(w, want_ito=want_ito)


def register_arena(arena):
    _ARENAS.add(arena)


def wgrad_to_param_layout(dw_oti, shape, out=None, dst=None):
    """fp32 [Cout, taps, Cin] -> fp32 parameter-shaped gradient [Cout, Cin, kd, kh, kw].  `out`: accumulate into it;
    `dst`: write into it (a slot of a data-parallel gradient bucket, data_parallel.grad_slot)."""
    cout, cin = shape[0], shape[1]
    taps = dw_oti.shape[1]
    if dst is not None:
        if taps == 1:
            dst.copy_(dw_oti.view(shape))
        else:
            call("adni_wgrad_to_param_layout", ptr(dw_oti), cout, cin, taps, ptr(dst), 0, stream_ptr())
        return dst
    if taps == 1 and out is None:      # 1x1x1 convs: [Cout][1][Cin] IS the parameter layout - no copy
        return dw_oti.view(shape)
    accumulate = out is not None
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=dw_oti.device)
    call("adni_wgrad_to_param_layout", ptr(dw_oti), cout, cin, taps, ptr(out), int(accumulate), stream_ptr())
    return out


def conv3d_fprop(x, w_oti, bias, k, stride, pad, dil, stats=False, engine=ENGINE_AUTO):
    _chk(x, BF16, "x")
    _chk(w_oti, BF16, "w_oti")
    N, D, H, W, Cin = x.shape
    Cout = w_oti.shape[0]
    g = geom(N, D, H, W, Cin, Cout, k, stride, pad, dil)
    Do, Ho, Wo = (out_extent(v, k, stride, pad, dil) for v in (D, H, W))
    y = torch.empty((N, Do, Ho, Wo, Cout), dtype=BF16, device=x.device)
    st = zeros_f64((2, Cout), x.device) if stats else None
    ev = PROFILE.begin()
    call("adni_conv3d_fprop", g, ptr(x), ptr(w_oti), ptr(bias), ptr(y), ptr(st[0]) if stats else None,
         ptr(st[1]) if stats else None, engine, stream_ptr())
    tag, frac = _engine_tag(g, 0, engine)
    PROFILE.end(ev, tag, 2 * N * Do * Ho * Wo * Cout * Cin * k ** 3,
                f"fprop N{N} {D}x{H}x{W} {Cin}->{Cout} k{k} s{stride} d{dil}", frac)
    return y, st


def conv3d_dgrad(dy, w_ito, in_shape, k, stride, pad, dil, addend=None, engine=ENGINE_AUTO):
    _chk(dy, BF16, "dy")
    _chk(w_ito, BF16, "w_ito")
    N, D, H, W, Cin = in_shape
    Cout = dy.shape[-1]
    g = geom(N, D, H, W, Cin, Cout, k, stride, pad, dil)
    dx = torch.empty(in_shape, dtype=BF16, device=dy.device)
    if addend is not None:
        _chk(addend, BF16, "addend")
    ev = PROFILE.begin()
    call("adni_conv3d_dgrad", g, ptr(dy), ptr(w_ito), ptr(addend), ptr(dx), engine, stream_ptr())
    tag, frac = _engine_tag(g, 1, engine)
    PROFILE.end(ev, tag, 2 * dy.numel() * Cin * k ** 3,
                f"dgrad N{N} {D}x{H}x{W} {Cin}->{Cout} k{k} s{stride} d{dil}", frac)
    return dx


def conv3d_dgrad_bnred(dy, w_ito, in_shape, k, stride, pad, dil, bn_y, bn_relu_out=None, bn_scale=None, bn_shift=None,
                       addend=None):
    """conv3d_dgrad whose epilogue also accumulates the BatchNorm-backward sums of the layer that produced dx's tensor
    (adni_conv3d_dgrad_bnred).  Returns (dx, red) with red fp64 [2, Cin] = [sum g | sum g*y]  (bn_bwd_apply red_form 1)."""
    _chk(dy, BF16, "dy")
    _chk(w_ito, BF16, "w_ito")
    _chk(bn_y, BF16, "bn_y")
    N, D, H, W, Cin = in_shape
    Cout = dy.shape[-1]
    if tuple(bn_y.shape) != tuple(in_shape):
        raise ValueError(f"conv3d_dgrad_bnred: bn_y {tuple(bn_y.shape)} must have dx's shape {tuple(in_shape)}")
    g = geom(N, D, H, W, Cin, Cout, k, stride, pad, dil)
    dx = torch.empty(in_shape, dtype=BF16, device=dy.device)
    red = zeros_f64((2, Cin), dy.device)
    ev = PROFILE.begin()
    call("adni_conv3d_dgrad_bnred", g, ptr(dy), ptr(w_ito), ptr(addend), ptr(dx), ptr(bn_y), ptr(bn_relu_out),
         ptr(bn_scale), ptr(bn_shift), ptr(red[0]), ptr(red[1]), stream_ptr())
    tag, frac = _engine_tag(g, 1, ENGINE_AUTO)
    PROFILE.end(ev, tag, 2 * dy.numel() * Cin * k ** 3,
                f"dgrad+bnred N{N} {D}x{H}x{W} {Cin}->{Cout} k{k} s{stride} d{dil}", frac)
    return dx, red


_BNRED_CACHE = {}


def dgrad_bnred_profitable(in_shape, Cout, k, stride, pad, dil):
    """Whether the library prefers the fused dgrad + BatchNorm-backward sums for this geometry
    (adni_conv3d_dgrad_bnred_profitable: long main loops only)."""
    key = (tuple(in_shape), Cout, k, stride, pad, dil)
    hit = _BNRED_CACHE.get(key)
    if hit is None:
        N, D, H, W, Cin = in_shape
        hit = _BNRED_CACHE[key] = bool(_lib.load().adni_conv3d_dgrad_bnred_profitable(geom(N, D, H, W, Cin, Cout, k, stride, pad, dil)))
    return hit


_SCRATCH_CACHE = {}


def _wgrad_scratch_floats(g):
    key = (g.Cin, g.Cout, g.k, g.stride, g.pad, g.dil)
    n = _SCRATCH_CACHE.get(key)
    if n is None:
        n = _SCRATCH_CACHE[key] = int(_lib.load().adni_conv3d_wgrad_scratch_floats(g))
    return n


def conv3d_wgrad(x, dy, k, stride, pad, dil, want_dbias=False, engine=ENGINE_AUTO):
    _chk(x, BF16, "x")
    _chk(dy, BF16, "dy")
    N, D, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    g = geom(N, D, H, W, Cin, Cout, k, stride, pad, dil)
    dw = zeros_f32((Cout, k * k * k, Cin), x.device)
    db = None
    if want_dbias:
        if engine == ENGINE_DIRECT or (engine == ENGINE_AUTO and _plan(g, 2)[0] == 0):
            db = torch.zeros((Cout,), dtype=torch.float32, device=x.device)
            ev = PROFILE.begin()
            call("adni_conv3d_wgrad", g, ptr(x), ptr(dy), ptr(dw), ptr(db), None, ENGINE_DIRECT, stream_ptr())
            PROFILE.end(ev, "direct", 2 * dy.numel() * Cin * k ** 3)
            return dw, db
        s = channel_stats(dy.view(-1, Cout))
        db = s[0].to(torch.float32)
    scratch = None
    if engine != ENGINE_DIRECT:
        n_scratch = _wgrad_scratch_floats(g)
        if n_scratch:  # halo-plane engine (layer1): accumulates in a small scratch, then overwrites dw
            scratch = zeros_f32((n_scratch,), x.device)
    ev = PROFILE.begin()
    call("adni_conv3d_wgrad", g, ptr(x), ptr(dy), ptr(dw), None, ptr(scratch), engine, stream_ptr())
    tag, frac = _engine_tag(g, 2, engine)
    PROFILE.end(ev, tag, 2 * dy.numel() * Cin * k ** 3,
                f"wgrad N{N} {D}x{H}x{W} {Cin}->{Cout} k{k} s{stride} d{dil}", frac)
    return dw, db


# --------------------------------------------------------------------------------------------- stem (tcgen05)
def stem_supported(Cin, Cout, k, stride, pad, dil):
    return Cin == 1 and Cout == 64 and k == 7 and stride == 2 and pad == 3 and dil == 1


def stem_expand(x):
    """bf16 [N, D, H, W, 1] -> X8 bf16 [N, D, H, W', 8] (filter-window pixels)."""
    _chk(x, BF16, "x")
    N, D, H, W = x.shape[:4]
    Wo = (W - 1) // 2 + 1
    x8 = torch.empty((N, D, H, Wo, 8), dtype=BF16, device=x.device)
    call_hbm("hbm_stem_expand", 2 * x.numel() + 2 * x8.numel(), "adni_stem_expand", ptr(x), N, D, H, W, ptr(x8),
             stream_ptr())
    return x8


def stem_fprop(x8, in_shape, w, stats=True):
    """x8 from stem_expand, w fp32 [64,1,7,7,7]. Returns (y bf16 [N,Do,Ho,Wo,64], stats fp64 [2,64])."""
    N, D, H, W = in_shape[:4]
    w2g = torch.empty((56, 64, 8), dtype=BF16, device=x8.device)
    call("adni_stem_weights", ptr(_chk(w.detach(), torch.float32, "weight")), ptr(w2g), stream_ptr())
    Do, Ho, Wo = ((v - 1) // 2 + 1 for v in (D, H, W))
    y = torch.empty((N, Do, Ho, Wo, 64), dtype=BF16, device=x8.device)
    st = zeros_f64((2, 64), x8.device) if stats else None
    ev = PROFILE.begin()
    call("adni_stem_fprop", ptr(x8), N, D, H, W, ptr(w2g), ptr(y), ptr(st[0]) if stats else None,
         ptr(st[1]) if stats else None, stream_ptr())
    PROFILE.end(ev, "tc_stem", 2 * N * Do * Ho * Wo * 64 * 343)
    return y, st


def stem_wgrad(x8, dy, in_shape):
    """Returns the fp32 parameter-layout gradient [64, 1, 7, 7, 7]."""
    _chk(dy, BF16, "dy")
    N, D, H, W = in_shape[:4]
    ws = torch.empty((512 * 64,), dtype=torch.float32, device=x8.device)
    grad = torch.empty((64, 1, 7, 7, 7), dtype=torch.float32, device=x8.device)
    ev = PROFILE.begin()
    call("adni_stem_wgrad", ptr(x8), ptr(dy), N, D, H, W, ptr(ws), ptr(grad), stream_ptr())
    PROFILE.end(ev, "tc_stem", 2 * dy.numel() * 343)
    return grad


# --------------------------------------------------------------------------------------------- BN
def bn_finalize(stats, count, gamma, beta, eps, momentum, running_mean, running_var):
    """stats: fp64 [2, C] (sum, sum of squares).  Returns fp32 [4, C]: rows mean, invstd, scale, shift."""
    C = stats.shape[1]
    out = torch.empty((4, C), dtype=torch.float32, device=stats.device)
    call("adni_bn_finalize", ptr(stats[0]), ptr(stats[1]), float(count), C, ptr(gamma), ptr(beta), float(eps),
         float(momentum), ptr(running_mean), ptr(running_var), ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(out[3]),
         stream_ptr())
    return out


def bn_apply(y, scale, shift, residual=None, relu=True):
    _chk(y, BF16, "y")
    C = y.shape[-1]
    rows = y.numel() // C
    out = torch.empty_like(y)
    call("adni_bn_apply", ptr(y), ptr(scale), ptr(shift), ptr(residual), ptr(out), rows, C, int(relu), None, None,
         stream_ptr())
    return out


def bn_train_apply(y, stats, count, gamma, beta, eps, momentum, running_mean, running_var, residual=None, relu=True):
    """Training-mode BatchNorm forward in one launch. Returns (out, bnp) with bnp = fp32 [4, C] (mean, invstd, scale,
    shift); running statistics are updated in place."""
    _chk(y, BF16, "y")
    C = y.shape[-1]
    rows = y.numel() // C
    out = torch.empty_like(y)
    bnp = torch.empty((4, C), dtype=torch.float32, device=y.device)
    call_hbm("hbm_bn_fwd", 2 * y.numel() * (3 if residual is not None else 2), "adni_bn_train_apply", ptr(y),
             ptr(stats[0]), ptr(stats[1]), float(count), ptr(gamma), ptr(beta), float(eps), float(momentum),
             ptr(running_mean), ptr(running_var), ptr(bnp), ptr(residual), ptr(out), rows, C, int(relu), stream_ptr())
    return out, bnp


def bn_param_grads(red):
    """fp64 [2, C] backward sums (sum g, sum g*xhat) -> (dgamma, dbeta) fp32 [C]."""
    C = red.shape[1]
    pg = torch.empty((2, C), dtype=torch.float32, device=red.device)
    call("adni_bn_param_grads", ptr(red), C, ptr(pg[0]), ptr(pg[1]), stream_ptr())
    return pg[0], pg[1]


def bn_eval_params(running_mean, running_var, gamma, beta, eps):
    C = running_mean.shape[0]
    out = torch.empty((2, C), dtype=torch.float32, device=running_mean.device)
    call("adni_bn_eval_params", ptr(running_mean), ptr(running_var), ptr(gamma), ptr(beta), float(eps), C, ptr(out[0]),
         ptr(out[1]), stream_ptr())
    return out[0], out[1]


def relu_fwd(x):
    _chk(x, BF16, "x")
    y = torch.empty_like(x)
    call("adni_relu_fwd", ptr(x), ptr(y), x.numel(), stream_ptr())
    return y


def relu_bwd(dy, y):
    dx = torch.empty_like(y)
    call("adni_relu_bwd", ptr(dy), ptr(y), ptr(dx), y.numel(), stream_ptr())
    return dx


def relu_f32(x, dy=None):
    """dy None: max(x, 0); else dy * (x > 0) with x the forward output."""
    x = _chk(x, torch.float32, "x")
    out = torch.empty_like(x)
    call("adni_relu_f32", ptr(x), ptr(dy), ptr(out), x.numel(), stream_ptr())
    return out


def dropout(x, p, seed, offset):
    """y = x * keep / (1 - p); keep = Philox(seed, *offset, index) >= p.  x: bf16 or fp32, any shape (contiguous);
    offset: one-element int64 DEVICE tensor.  The backward pass is the same call on the gradient."""
    if x.dtype not in (BF16, torch.float32):
        raise ValueError(f"dropout: unsupported dtype {x.dtype}")
    if not x.is_contiguous():
        raise ValueError("dropout: tensor must be contiguous")
    _chk(offset, torch.int64, "offset")
    y = torch.empty_like(x)
    call_hbm("hbm_dropout", 2 * x.numel() * x.element_size(), "adni_dropout", ptr(x), ptr(y), x.numel(),
             int(x.dtype == torch.float32), float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(offset), stream_ptr())
    return y


def channel_stats(x2d):
    """bf16 [rows, C] -> fp64 [2, C] (sum, sum of squares)."""
    _chk(x2d, BF16, "x")
    C = x2d.shape[-1]
    rows = x2d.numel() // C
    st = zeros_f64((2, C), x2d.device)
    call("adni_channel_stats", ptr(x2d), rows, C, ptr(st[0]), ptr(st[1]), stream_ptr())
    return st


def bn_bwd_reduce(dout, out, y, mean, invstd, relu, scale=None, shift=None):
    """relu mask: from `out`, or (out None, no residual in the forward) recomputed from y*scale+shift."""
    C = y.shape[-1]
    rows = y.numel() // C
    red = zeros_f64((2, C), y.device)
    call_hbm("hbm_bn_bwd_reduce", 2 * y.numel() * (3 if (relu and out is not None) else 2), "adni_bn_bwd_reduce",
             ptr(dout), ptr(out) if relu else None, ptr(y), ptr(mean), ptr(invstd), ptr(scale), ptr(shift), rows, C,
             int(relu), ptr(red), stream_ptr())
    return red


def bn_bwd_apply(dout, out, y, mean, invstd, gamma, red, count, relu, want_dres, want_param_grads=True, scale=None,
                 shift=None, param_grad_scale=1.0, red_form=0):
    C = y.shape[-1]
    rows = y.numel() // C
    dy = torch.empty_like(y)
    dres = torch.empty_like(y) if want_dres else None
    pg = torch.empty((2, C), dtype=torch.float32, device=y.device) if want_param_grads else None
    n_tensors = 3 + (1 if (relu and out is not None) else 0) + (1 if want_dres else 0)  # dout, y, dy (+out) (+dres)
    call_hbm("hbm_bn_bwd_apply", 2 * y.numel() * n_tensors, "adni_bn_bwd_apply", ptr(dout),
             ptr(out) if relu else None, ptr(y), ptr(mean), ptr(invstd), ptr(gamma), ptr(scale), ptr(shift), ptr(red),
             float(count), rows, C, int(relu), ptr(dy), ptr(dres), ptr(pg[0]) if want_param_grads else None,
             ptr(pg[1]) if want_param_grads else None, float(param_grad_scale), int(red_form), stream_ptr())
    if want_param_grads:
        return dy, dres, pg[0], pg[1]
    return dy, dres, None, None


# --------------------------------------------------------------------------------------------- pooling
def maxpool3d_fwd(x, k, stride, pad):
    _chk(x, BF16, "x")
    N, D, H, W, C = x.shape
    Do, Ho, Wo = ((v + 2 * pad - k) // stride + 1 for v in (D, H, W))
    y = torch.empty((N, Do, Ho, Wo, C), dtype=BF16, device=x.device)
    am = torch.empty((N, Do, Ho, Wo, C), dtype=torch.uint8, device=x.device)
    call("adni_maxpool3d_fwd", ptr(x), N, D, H, W, C, k, stride, pad, ptr(y), ptr(am), stream_ptr())
    return y, am


def maxpool3d_bwd(dy, argmax, in_shape, k, stride, pad):
    N, D, H, W, C = in_shape
    dx = torch.empty(in_shape, dtype=BF16, device=dy.device)
    call("adni_maxpool3d_bwd", ptr(dy), ptr(argmax), N, D, H, W, C, k, stride, pad, ptr(dx), stream_ptr())
    return dx


def relu_maxpool_supported(k, stride, pad):
    return k in (2, 3) and stride == k and pad == 0


def relu_maxpool_fwd(x, k, relu=True):
    """relu(maxpool_k(x)) for non-overlapping windows in one pass; returns (pooled, argmax)."""
    _chk(x, BF16, "x")
    N, D, H, W, C = x.shape
    Do, Ho, Wo = D // k, H // k, W // k
    y = torch.empty((N, Do, Ho, Wo, C), dtype=BF16, device=x.device)
    am = torch.empty((N, Do, Ho, Wo, C), dtype=torch.uint8, device=x.device)
    call_hbm("hbm_relu_pool", 2 * x.numel() + 3 * y.numel(), "adni_relu_maxpool_fwd", ptr(x), N, D, H, W, C, k, int(relu),
             ptr(y), ptr(am), stream_ptr())
    return y, am


def relu_maxpool_bwd(dy, argmax, pooled, in_shape, k):
    """Gradient of relu_maxpool_fwd w.r.t. x (pooled = the forward output: ReLU mask; None = plain max pool)."""
    _chk(dy, BF16, "dy")
    N, D, H, W, C = in_shape
    dx = torch.empty(in_shape, dtype=BF16, device=dy.device)
    call_hbm("hbm_relu_pool", 2 * dx.numel() + 5 * dy.numel(), "adni_relu_maxpool_bwd", ptr(dy), ptr(argmax), ptr(pooled),
             N, D, H, W, C, k, ptr(dx), stream_ptr())
    return dx


def fused_pool_supported(k, stride, pad):
    return (k, stride, pad) == (3, 2, 1)


_POOL_STREAMING = os.environ.get("ADNI_POOL_STREAM", "1") != "0"


def bn_relu_maxpool_fwd(y, bnp, k, stride, pad, want_raw=True):
    """maxpool(relu(y*scale+shift)) without materialising the activation. bnp: fp32 [4, C] from bn_finalize.
    Returns (pooled, argmax, y_at_argmax): the raw conv output at every window's arg-max (None with the tiled kernels,
    ADNI_POOL_STREAM=0) lets the BatchNorm backward sums run over the pooled tensor (bn_bwd_reduce)."""
    _chk(y, BF16, "y")
    N, D, H, W, C = y.shape
    Do, Ho, Wo = ((v + 2 * pad - k) // stride + 1 for v in (D, H, W))
    p = torch.empty((N, Do, Ho, Wo, C), dtype=BF16, device=y.device)
    am = torch.empty((N, Do, Ho, Wo, C), dtype=torch.uint8, device=y.device)
    raw = torch.empty((N, Do, Ho, Wo, C), dtype=BF16, device=y.device) if (want_raw and _POOL_STREAMING) else None
    call_hbm("hbm_pool_fwd", 2 * y.numel() + (5 if raw is not None else 3) * p.numel(), "adni_bn_relu_maxpool_fwd", ptr(y),
             ptr(bnp[2]), ptr(bnp[3]), N, D, H, W, C, k, stride, pad, ptr(p), ptr(am), ptr(raw), stream_ptr())
    return p, am, raw


def maxpool_bn_bwd_reduce(dp, argmax, y, bnp, k, stride, pad):
    N, D, H, W, C = y.shape
    red = zeros_f64((2, C), y.device)
    call_hbm("hbm_pool_bwd_reduce", 2 * y.numel() + 3 * dp.numel(), "adni_maxpool_bn_bwd_reduce", ptr(dp),
             ptr(argmax), ptr(y), ptr(bnp), N, D, H, W, C, k, stride, pad, ptr(red), stream_ptr())
    return red


def maxpool_bn_bwd_apply(dp, argmax, y, bnp, gamma, red, count, k, stride, pad):
    N, D, H, W, C = y.shape
    dy = torch.empty_like(y)
    call_hbm("hbm_pool_bwd_apply", 4 * y.numel() + 3 * dp.numel(), "adni_maxpool_bn_bwd_apply", ptr(dp), ptr(argmax),
             ptr(y), ptr(bnp), ptr(gamma), ptr(red), float(count), N, D, H, W, C, k, stride, pad, ptr(dy), stream_ptr())
    return dy


def gap_fwd(x):
    _chk(x, BF16, "x")
    N, C = x.shape[0], x.shape[-1]
    P = x.numel() // (N * C)
    feat = torch.empty((N, C), dtype=torch.float32, device=x.device)
    call_hbm("hbm_gap", 2 * x.numel() + 4 * feat.numel(), "adni_gap_fwd", ptr(x), N, P, C, ptr(feat), stream_ptr())
    return feat


def gap_bwd(dfeat, shape):
    N, C = shape[0], shape[-1]
    P = 1
    for v in shape[1:-1]:
        P *= v
    dfeat = _chk(dfeat, torch.float32, "dfeat")
    dx = torch.empty(shape, dtype=BF16, device=dfeat.device)
    call_hbm("hbm_gap", 2 * dx.numel() + 4 * dfeat.numel(), "adni_gap_bwd", ptr(dfeat), N, P, C, ptr(dx), stream_ptr())
    return dx


# --------------------------------------------------------------------------------------------- heads
def _ld(t):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError("expected a 2-D tensor with unit column stride")
    return t.stride(0)


def linear_fwd(x, W, b, relu, out=None):
    """x fp32 [B, in] (row-strided view allowed), W fp32 [out, in]; `out` may be a column slice (concat)."""
    B, nin = x.shape
    nout = W.shape[0]
    if out is None:
        out = torch.empty((B, nout), dtype=torch.float32, device=x.device)
    call("adni_linear_fwd", ptr(x), _ld(x), ptr(W), ptr(b), ptr(out), _ld(out), B, nin, nout, int(relu), stream_ptr())
    return out


def linear_bwd(x, W, y, dy, relu, need_dx=True, need_dw=True, has_bias=True):
    B, nin = x.shape
    nout = W.shape[0]
    dx = torch.empty((B, nin), dtype=torch.float32, device=x.device) if need_dx else None
    dW = torch.zeros_like(W) if need_dw else None
    db = torch.zeros((nout,), dtype=torch.float32, device=x.device) if (need_dw and has_bias) else None
    call("adni_linear_bwd", ptr(x), _ld(x), ptr(W), ptr(y), _ld(y) if y is not None else 0, ptr(dy), _ld(dy), ptr(dx),
         _ld(dx) if need_dx else 0, 0, ptr(dW), ptr(db), B, nin, nout, int(relu), stream_ptr())
    return dx, dW, db


def rows_stats_f32(x):
    B, C = x.shape
    st = zeros_f64((2, C), x.device)
    call("adni_rows_stats_f32", ptr(x), _ld(x), B, C, ptr(st), stream_ptr())
    return st


def bn1d_apply(x, scale, shift, relu):
    B, C = x.shape
    y = torch.empty((B, C), dtype=torch.float32, device=x.device)
    call("adni_bn1d_apply", ptr(x), _ld(x), ptr(scale), ptr(shift), ptr(y), _ld(y), B, C, int(relu), stream_ptr())
    return y


def bn1d_bwd_reduce(dy, y, x, mean, invstd, relu):
    B, C = x.shape
    red = zeros_f64((2, C), x.device)
    call("adni_bn1d_bwd_reduce", ptr(dy), _ld(dy), ptr(y), _ld(y), ptr(x), _ld(x), ptr(mean), ptr(invstd), B, C,
         int(relu), ptr(red), stream_ptr())
    return red


def bn1d_bwd_apply(dy, y, x, mean, invstd, gamma, red, count, relu):
    B, C = x.shape
    dx = torch.empty((B, C), dtype=torch.float32, device=x.device)
    pg = torch.empty((2, C), dtype=torch.float32, device=x.device)
    call("adni_bn1d_bwd_apply", ptr(dy), _ld(dy), ptr(y), _ld(y), ptr(x), _ld(x), ptr(mean), ptr(invstd), ptr(gamma),
         ptr(red), float(count), B, C, int(relu), ptr(dx), _ld(dx), ptr(pg[0]), ptr(pg[1]), stream_ptr())
    return dx, pg[0], pg[1]


# --------------------------------------------------------------------------------------------- loss
def loss_fwd(logits, target, gamma, class_weights):
    """Returns (partial fp64[2] = [numerator, normaliser], per-sample coefficient fp64[B])."""
    B, C = logits.shape
    _chk(target, torch.int64, "target")
    partial = torch.zeros((2,), dtype=torch.float64, device=logits.device)
    coeff = torch.empty((B,), dtype=torch.float64, device=logits.device)
    call("adni_loss_fwd", ptr(logits), int(logits.dtype == torch.float64), _ld(logits), ptr(target), B, C,
         float(gamma or 0.0), ptr(class_weights), ptr(partial), ptr(coeff), stream_ptr())
    return partial, coeff


def loss_bwd(logits, target, coeff, denom, upstream=None):
    """denom / upstream: fp64 device scalars (1-element tensors); dlogits has the dtype of logits."""
    B, C = logits.shape
    if upstream is None:
        upstream = torch.ones((1,), dtype=torch.float64, device=logits.device)
    dl = torch.empty((B, C), dtype=logits.dtype, device=logits.device)
    call("adni_loss_bwd", ptr(logits), int(logits.dtype == torch.float64), _ld(logits), ptr(target), B, C, ptr(coeff),
         ptr(denom), ptr(upstream), ptr(dl), _ld(dl), stream_ptr())
    return dl


# --------------------------------------------------------------------------------------------- test-epoch metrics
def bootstrap_metrics(logits, labels, draws=None, want_confmat=False):
    """logits (n, C) fp64, labels (n,) int64, draws (d, n) int64 resampling indices or None (the identity, d = 1).
    Returns dict(f1 (d,), f1_class (d, C), mcc (d,)[, confmat (d, C, C) int64]) - one launch for all draws."""
    _chk(logits, torch.float64, "logits")
    _chk(labels, torch.int64, "labels")
    n, C = logits.shape
    d = 1 if draws is None else draws.shape[0]
    if draws is not None:
        _chk(draws, torch.int64, "draws")
        assert draws.shape[1] == n
    dev = logits.device
    out = {"f1": torch.empty((d,), dtype=torch.float32, device=dev),
           "f1_class": torch.empty((d, C), dtype=torch.float32, device=dev),
           "mcc": torch.empty((d,), dtype=torch.float32, device=dev)}
    cm = torch.empty((d, C, C), dtype=torch.int64, device=dev) if want_confmat else None
    call("adni_bootstrap_metrics", ptr(logits), logits.stride(0), ptr(labels), ptr(draws), n, C, d, ptr(out["f1"]),
         ptr(out["f1_class"]), ptr(out["mcc"]), ptr(cm), stream_ptr())
    if cm is not None:
        out["confmat"] = cm
    return out


# --------------------------------------------------------------------------------------------- normalisation
def quantile_minmax_normalize(x, mask, q, out_dtype=torch.float32, want_info=False):
    """x fp32 [S, ...], mask uint8 same shape.  Returns normalised volumes (and (info, qvals) if asked)."""
    _chk(x, torch.float32, "x")
    _chk(mask, torch.uint8, "mask")
    S = x.shape[0]
    nvox = x[0].numel()
    ws_bytes = int(_lib.load().adni_quantile_workspace_bytes(S))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
    of = torch.empty_like(x) if out_dtype == torch.float32 else None
    ob = torch.empty(x.shape, dtype=BF16, device=x.device) if out_dtype == BF16 else None
    info = torch.empty((S, 8), dtype=torch.int64, device=x.device) if want_info else None
    qv = torch.empty((S, 2), dtype=torch.float64, device=x.device) if want_info else None
    # select pass N*(4+1) + apply pass N*(4+1+o); the extra radix passes count against the kernel
    call_hbm("hbm_quantile_normalize", x.numel() * (10 + (4 if of is not None else 2)), "adni_quantile_minmax_normalize",
             ptr(x), ptr(mask), S, nvox, float(q), ptr(of), ptr(ob), ptr(info), ptr(qv), ptr(ws), ws_bytes, stream_ptr())
    out = of if of is not None else ob
    if want_info:
        return out, info, qv
    return out


def standardize(x, mean, std, mask=None, out_dtype=torch.float32):
    _chk(x, torch.float32, "x")
    of = torch.empty_like(x) if out_dtype == torch.float32 else None
    ob = torch.empty(x.shape, dtype=BF16, device=x.device) if out_dtype == BF16 else None
    call_hbm("hbm_standardize", x.numel() * (4 + (1 if mask is not None else 0) + (4 if of is not None else 2)),
             "adni_standardize", ptr(x), ptr(mask), x.numel(), float(mean), float(std), ptr(of), ptr(ob), stream_ptr())
    return of if of is not None else ob


def scan_moments(x):
    _chk(x, torch.float32, "x")
    S = x.shape[0]
    m = torch.zeros((S, 2), dtype=torch.float64, device=x.device)
    call("adni_scan_moments", ptr(x), S, x[0].numel(), ptr(m), stream_ptr())
    return m


def masked_std_mean(x, mask):
    _chk(x, torch.float32, "x")
    _chk(mask, torch.uint8, "mask")
    S = x.shape[0]
    out = torch.empty((S, 3), dtype=torch.float64, device=x.device)
    call("adni_masked_std_mean", ptr(x), ptr(mask), S, x[0].numel(), ptr(out), stream_ptr())
    return out


def cast_to_bf16(x):
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    if x.dtype == torch.float32:
        call("adni_cast_f32_to_bf16", ptr(x), ptr(y), x.numel(), stream_ptr())
    elif x.dtype == torch.float64:
        call("adni_cast_f64_to_bf16", ptr(x), ptr(y), x.numel(), stream_ptr())
    else:
        raise ValueError(f"cast_to_bf16: unsupported dtype {x.dtype}")
    return y


def volumes_to_ndhwc(x):
    """(B, C, D, H, W) fp32/fp64 NCDHW module input with C in {2, 3, 4} modalities -> bf16 NDHWC (B, D, H, W, C)."""
    if x.dtype not in (torch.float32, torch.float64):
        raise ValueError(f"volumes_to_ndhwc: unsupported dtype {x.dtype}")
    x = x.contiguous()
    N, C, D, H, W = x.shape
    out = torch.empty((N, D, H, W, C), dtype=BF16, device=x.device)
    call_hbm("hbm_input_cast", x.numel() * (x.element_size() + 2), "adni_volumes_to_ndhwc_bf16", ptr(x),
             int(x.dtype == torch.float64), N, C, D * H * W, ptr(out), stream_ptr())
    return out


def maxout_fwd(a, b):
    _chk(a, BF16, "a")
    _chk(b, BF16, "b")
    if a.shape != b.shape:
        raise ValueError(f"maxout: shapes differ {tuple(a.shape)} vs {tuple(b.shape)}")
    out = torch.empty_like(a)
    call_hbm("hbm_maxout", 6 * a.numel(), "adni_maxout_fwd", ptr(a), ptr(b), ptr(out), a.numel(), stream_ptr())
    return out


def maxout_bwd(dout, a, b, want_a=True, want_b=True):
    _chk(dout, BF16, "dout")
    da = torch.empty_like(a) if want_a else None
    db = torch.empty_like(b) if want_b else None
    call_hbm("hbm_maxout", 2 * a.numel() * (3 + int(want_a) + int(want_b)), "adni_maxout_bwd", ptr(dout), ptr(a), ptr(b),
             ptr(da), ptr(db), a.numel(), stream_ptr())
    return da, db


def concat_channels(a, b):
    _chk(a, BF16, "a")
    _chk(b, BF16, "b")
    if a.shape[:-1] != b.shape[:-1]:
        raise ValueError(f"concat_channels: shapes differ {tuple(a.shape)} vs {tuple(b.shape)}")
    Ca, Cb = a.shape[-1], b.shape[-1]
    out = torch.empty(tuple(a.shape[:-1]) + (Ca + Cb,), dtype=BF16, device=a.device)
    call_hbm("hbm_concat", 4 * out.numel(), "adni_concat_channels", ptr(a), Ca, ptr(b), Cb, a.numel() // Ca, ptr(out),
             stream_ptr())
    return out


def split_channels(dout, Ca, Cb, want_a=True, want_b=True):
    _chk(dout, BF16, "dout")
    lead = tuple(dout.shape[:-1])
    da = torch.empty(lead + (Ca,), dtype=BF16, device=dout.device) if want_a else None
    db = torch.empty(lead + (Cb,), dtype=BF16, device=dout.device) if want_b else None
    call_hbm("hbm_concat", 4 * dout.numel(), "adni_split_channels", ptr(dout), Ca, Cb, dout.numel() // (Ca + Cb), ptr(da),
             ptr(db), stream_ptr())
    return da, db


def pad_volume_high(x, extra):
    """bf16 NDHWC volume with `extra` zero voxels appended on the high side of D, H and W."""
    _chk(x, BF16, "x")
    N, D, H, W, C = x.shape
    out = torch.empty((N, D + extra, H + extra, W + extra, C), dtype=BF16, device=x.device)
    call_hbm("hbm_pad", 2 * (x.numel() + out.numel()), "adni_pad_volume_high", ptr(x), N, D, H, W, C, extra, extra, extra,
             ptr(out), stream_ptr())
    return out


def crop_volume_high(xp, extra):
    """Inverse of pad_volume_high: the leading (D, H, W) box of a padded volume (gradient of the padding)."""
    _chk(xp, BF16, "x_padded")
    N, Dp, Hp, Wp, C = xp.shape
    D, H, W = Dp - extra, Hp - extra, Wp - extra
    out = torch.empty((N, D, H, W, C), dtype=BF16, device=xp.device)
    call_hbm("hbm_pad", 4 * out.numel(), "adni_crop_volume_high", ptr(xp), N, D, H, W, C, extra, extra, extra, ptr(out),
             stream_ptr())
    return out


def cast_to_f32(x):
    _chk(x, BF16, "x")
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    call("adni_cast_bf16_to_f32", ptr(x), ptr(y), x.numel(), stream_ptr())
    return y

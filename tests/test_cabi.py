"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/adni_b200.h declares, and
rejects bad arguments with the documented error codes before touching a GPU."""
import ctypes
import os
import re

import pytest
import torch

from multimodal_alzheimer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "adni_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(adni_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/adni_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS), set(declared) ^ set(_lib.EXPORTED_SYMBOLS)


def test_version_and_launch_counter(lib):
    assert lib.adni_version() >= 100
    assert lib.adni_launch_count() >= 0


def test_out_extent_matches_torch_formula(lib):
    for n, k, s, p, d in [(128, 7, 2, 3, 1), (64, 3, 2, 1, 1), (16, 3, 1, 4, 4), (91, 7, 2, 3, 1), (23, 3, 2, 1, 1)]:
        y = torch.nn.functional.conv1d(torch.zeros(1, 1, n), torch.zeros(1, 1, k), stride=s, padding=p, dilation=d)
        assert lib.adni_conv3d_out_extent(n, k, s, p, d) == y.shape[-1]


def test_bad_arguments_return_einval_without_a_gpu(lib):
    g = _lib.geom(1, 8, 8, 8, 64, 64, 3, 1, 1, 1)
    assert lib.adni_conv3d_fprop(ctypes.byref(g), None, None, None, None, None, None, 0, None) == -1
    assert b"null" in lib.adni_last_error_string()
    bad = _lib.geom(1, 8, 8, 8, 64, 64, 3, 0, 1, 1)  # stride 0
    assert lib.adni_conv3d_fprop(ctypes.byref(bad), None, None, None, None, None, None, 0, None) == -1
    assert lib.adni_quantile_workspace_bytes(3) == 3 * lib.adni_quantile_workspace_bytes(1) > 0
    assert lib.adni_loss_fwd(None, 0, 3, None, 4, 3, 1.0, None, None, None, None) == -1


def test_unsupported_engine_is_enotsup_not_a_fallback(lib):
    g = _lib.geom(1, 8, 8, 8, 8, 8, 3, 1, 1, 1)
    one = ctypes.c_void_p(16)  # never dereferenced: the engine check happens first
    assert lib.adni_conv3d_fprop(ctypes.byref(g), one, one, None, one, None, None, _lib.ENGINE_TCGEN05, None) == -2


def test_cpu_tensors_are_rejected_no_fallback():
    with pytest.raises(_lib.AdniError):
        _lib.ptr(torch.zeros(4))
    from multimodal_alzheimer_b200 import kernels as K
    with pytest.raises(_lib.AdniError):
        K.relu_f32(torch.zeros(8))


def test_adam_entry_point_validates_before_touching_a_gpu(lib):
    """adni_adam_step_multi: empty table is a no-op, torch.optim.Adam's ValueError cases are ADNI_EINVAL, and the
    optimizer refuses CPU parameters (no CPU fallback)."""
    assert 16 <= lib.adni_adam_max_tensors_per_launch() <= 64
    assert lib.adni_adam_step_multi(0, None, None, None, None, None, None, None, None, None, 0.9, 0.999, 1e-8, None) == 0
    assert lib.adni_adam_step_multi(1, None, None, None, None, None, None, None, None, None, 0.9, 0.999, 1e-8, None) == -1
    assert lib.adni_adam_step_multi(0, None, None, None, None, None, None, None, None, None, 1.0, 0.999, 1e-8, None) == -1
    assert lib.adni_adam_step_multi(0, None, None, None, None, None, None, None, None, None, 0.9, 0.999, -1.0, None) == -1
    from multimodal_alzheimer_b200.optim import Adam
    p = torch.nn.Parameter(torch.zeros(4))
    opt = Adam([{"params": p, "lr": 1e-3}], weight_decay=1e-4)
    assert isinstance(opt, torch.optim.Adam) and opt.param_groups[0]["weight_decay"] == 1e-4
    opt.step()  # no gradients: nothing to do, like torch
    p.grad = torch.ones(4)
    with pytest.raises(_lib.AdniError):
        opt.step()
    with pytest.raises(ValueError):
        Adam([p], lr=-1.0)
    with pytest.raises(NotImplementedError):
        Adam([p], amsgrad=True)


def _encoder_conv_geometries(depth, shape):
    """(D, H, W, Cin, Cout, k, stride, pad, dil) of every Conv3d of the MedicalNet encoder for an input volume, from a
    meta-device forward of the oracle's restatement."""
    import torch
    from oracle.medicalnet import generate_model
    m = generate_model(depth).to("meta")
    out = []

    def hook(mod, inp, o):
        x = inp[0]
        out.append((x.shape[2], x.shape[3], x.shape[4], mod.in_channels, mod.out_channels, mod.kernel_size[0],
                    mod.stride[0], mod.padding[0], mod.dilation[0]))

    hooks = [mm.register_forward_hook(hook) for mm in m.modules() if isinstance(mm, torch.nn.Conv3d)]
    with torch.no_grad():
        m(torch.empty((1, 1) + tuple(shape), device="meta"))
    for h in hooks:
        h.remove()
    return out


@pytest.mark.parametrize("depth,shape,n_per_gpu", [
    (10, (128, 128, 128), 2),     # BASELINE.json configs[0]
    (18, (128, 128, 128), 16),    # configs[1]
    (18, (128, 128, 128), 4),     # configs[2] at 8 GPUs (4 pairs per GPU)
    (18, (91, 109, 91), 32),      # the reference's own grid (MNI 2 mm)
    (50, (160, 192, 160), 8),     # configs[4]: the largest conv / wgrad path
])
def test_planner_routes_every_encoder_conv_to_a_tensor_core_engine(depth, shape, n_per_gpu):
    """adni_conv3d_plan_info (a host-side query, no GPU needed) for every conv of the encoder at the BASELINE
    configurations: fprop, dgrad and wgrad are all planned on a tcgen05 engine (kind 1 = tap-per-box, 2 =
    halo-resident), never the CUDA-core fallback engine (0) and never ADNI_ENOTSUP; the executed fraction of the
    (tile, tap) blocks is in (0, 1].  The 1-channel stem has its own entry points (adni_stem_*)."""
    import ctypes
    lib = _lib.load()
    convs = [g for g in _encoder_conv_geometries(depth, shape) if g[3] > 1]
    assert len(convs) == {10: 11, 18: 19, 50: 52}[depth]
    for (D, H, W, Ci, Co, k, s, p, d) in convs:
        for pass_ in (0, 1, 2):
            g = _lib.ConvGeom(n_per_gpu, D, H, W, Ci, Co, k, s, p, d)
            kind, frac = ctypes.c_int(-1), ctypes.c_double(-1.0)
            rc = lib.adni_conv3d_plan_info(ctypes.byref(g), pass_, ctypes.byref(kind), ctypes.byref(frac))
            assert rc == 0, ((D, H, W, Ci, Co, k, s, p, d), pass_, lib.adni_last_error_string())
            assert kind.value in (1, 2), ((D, H, W, Ci, Co, k, s, p, d), pass_, kind.value)
            assert 0.0 < frac.value <= 1.0


def test_planner_routes_small_stacks_to_tensor_cores():
    """Every conv of the reference's small-CNN search space (train_pet_cnn.py:43-59) is planned onto a tensor-core
    engine for all three passes (host-side query, no kernel launch)."""
    import ctypes
    from multimodal_alzheimer_b200 import _lib
    lib = _lib.load()
    for first in (8, 16, 32):
        for n in (3, 4):
            chans = [1] + [first * 2 ** i for i in range(n)]
            for filt in ((5, 5, 3, 3), (7, 5, 3, 3), (5, 5, 5, 3), (3, 3, 3, 3)):
                ext = 128
                for i in range(n):
                    k = filt[i]
                    g = _lib.geom(2, ext, ext, ext, chans[i], chans[i + 1], k, 1, (k - 1) // 2, 1)
                    for pass_ in (0, 1, 2):
                        if pass_ == 1 and i == 0:
                            continue                       # no input gradient for the first layer
                        kind, frac = ctypes.c_int(-1), ctypes.c_double(0)
                        assert lib.adni_conv3d_plan_info(ctypes.byref(g), pass_, ctypes.byref(kind), ctypes.byref(frac)) == 0
                        if (chans[i], chans[i + 1], k) == (64, 128, 5) or (chans[i], chans[i + 1], k, pass_) == (32, 64, 5, 1):
                            # 125 taps exceed the tcgen05 tap mask and 128 channels the small engine; the dgrad of
                            # 32 -> 64 at k = 5 (a 64-channel halo tile + 125 x 64-deep weights) exceeds shared memory
                            assert kind.value == 0
                            continue
                        assert kind.value in (1, 2, 3), (chans[i], chans[i + 1], k, pass_, kind.value)
                    ext //= 2


def _wgrad_schedule(lib, g):
    import ctypes
    header = (ctypes.c_int * 16)()
    table = (ctypes.c_int * (4 * 160))()
    rc = lib.adni_conv3d_wgrad_schedule(ctypes.byref(g), header, table, 4 * 160)
    assert rc == 0, lib.adni_last_error_string()
    return list(header), list(table)


@pytest.mark.parametrize("geom,chunk_mb", [
    ((32, 16, 16, 16, 512, 512, 3, 1, 4, 4), 48),    # layer4 of the bench workload: 268 MB of operands -> 6 chunks
    ((32, 16, 16, 16, 256, 256, 3, 1, 2, 2), 48),    # layer3: 3 chunks
    ((32, 16, 16, 16, 512, 512, 3, 1, 4, 4), 100000),  # one chunk = the plain tile-major sequence
    ((4, 16, 16, 16, 512, 512, 3, 1, 4, 4), 48),     # the 8-GPU per-rank batch: fits the L2, one chunk
    ((5, 12, 14, 12, 256, 512, 3, 2, 1, 1), 1),      # strided, ragged extents, tiny chunks (one sample each)
    ((8, 20, 24, 20, 1024, 256, 1, 1, 0, 1), 16),    # a ResNet-50 1x1x1 conv
    ((6, 20, 24, 20, 128, 256, 3, 1, 2, 2), 2),      # dilated, box grid not a divisor of the extents
])
def test_wgrad_stream_k_schedule_covers_every_active_block_once(geom, chunk_mb, monkeypatch):
    """Host-side check of the chunk-major stream-K schedule of the tcgen05 wgrad kernel (adni_conv3d_wgrad_schedule, no
    GPU): replaying the CTA walks exactly as conv_wgrad2.cu does (make_ctx + the box loops) must visit every ACTIVE
    (output tile, position box) block exactly once - activity recomputed here from the conv geometry alone - and the
    blocks per CTA must differ by at most one."""
    lib = _lib.load()
    monkeypatch.setenv("ADNI_WGRAD_CHUNK_MB", str(chunk_mb))
    N, D, H, W, Ci, Co, k, s, pad, dil = geom
    g = _lib.ConvGeom(*geom)
    h, table = _wgrad_schedule(lib, g)
    used, ctas, chunk_boxes, pos_boxes, m_tiles, n_tiles, groups, n_groups, cin_blocks, bd, bh, bw, td, th, tw, _ = h
    assert used == 1 and ctas == 148
    per_sample = td * th * tw
    assert pos_boxes == N * per_sample and chunk_boxes % per_sample == 0 and chunk_boxes > 0

    def axis(ext):           # per kernel index: (view extent, box offset) - conv_api.cu fwd_axis_taps / parity_views
        out = []
        for kk in range(k):
            off = kk * dil - pad
            par = off % s
            out.append(((ext - par + s - 1) // s, (off - par) // s))
        return out
    ad, ah, aw = axis(D), axis(H), axis(W)

    def active(nt, r):       # any 64-column group of N tile nt whose shifted box meets the input (box_active)
        x, y, z = r % tw, (r // tw) % th, r // (tw * th)
        for gi in range(nt * groups, min((nt + 1) * groups, n_groups)):
            t = gi // cin_blocks
            kd, kh, kw = t // (k * k), (t // k) % k, t % k
            ok = True
            for (ext, off), b0, b in ((ad[kd], z * bd, bd), (ah[kh], y * bh, bh), (aw[kw], x * bw, bw)):
                lo = b0 + off
                ok = ok and lo + b > 0 and lo < ext
            if ok:
                return True
        return False
    act = [[active(nt, r) for r in range(per_sample)] for nt in range(n_tiles)]
    expected = sum(sum(a) for a in act) * m_tiles * N
    seen = {}
    loads = []
    tiles = m_tiles * n_tiles
    for c in range(ctas):
        t0, b0, t1, b1 = table[4 * c:4 * c + 4]
        n = 0
        for v in range(t0, t1 + 1):
            nt, mt, chunk = v % n_tiles, (v // n_tiles) % m_tiles, v // tiles
            lo = b0 if v == t0 else chunk * chunk_boxes
            hi = b1 if v == t1 else min((chunk + 1) * chunk_boxes, pos_boxes)
            assert chunk * chunk_boxes <= lo <= hi <= min((chunk + 1) * chunk_boxes, pos_boxes), (c, v, lo, hi)
            for b in range(lo, hi):
                if act[nt][b % per_sample]:
                    key = (mt, nt, b)
                    assert key not in seen, (key, c, seen[key])
                    seen[key] = c
                    n += 1
        loads.append(n)
    assert len(seen) == expected
    assert max(loads) - min(loads) <= 1, (min(loads), max(loads))
    if chunk_mb >= 100000 or N == 4:
        assert chunk_boxes == pos_boxes
    elif geom[:2] == (32, 16) and Ci == 512:
        assert chunk_boxes == 6 * per_sample          # 8 MB of dY + X per sample -> 6 samples per 48 MB chunk

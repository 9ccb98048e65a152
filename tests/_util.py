"""Shared helpers for the parity tests."""
import torch


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def describe_mismatch(got, ref, tol):
    """Human-readable error report (used in assertion messages to debug layout / descriptor bugs)."""
    g = got.double().flatten()
    r = ref.double().flatten()
    err = (g - r).abs()
    scale = r.abs().max().clamp_min(1e-30)
    bad = err > tol * scale
    nbad = int(bad.sum())
    msg = [f"rel_l2={rel_l2(got, ref):.3e} max_abs_err={float(err.max()):.3e} ref_max={float(scale):.3e} "
           f"bad={nbad}/{g.numel()} nan={int(torch.isnan(g).sum())}"]
    if nbad:
        idx = torch.nonzero(bad).flatten()[:8]
        shape = tuple(ref.shape)
        for i in idx.tolist():
            coord = []
            rem = i
            for s in reversed(shape):
                coord.append(rem % s)
                rem //= s
            msg.append(f"  at {tuple(reversed(coord))}: got {float(g[i]):.6g} ref {float(r[i]):.6g}")
    return "\n".join(msg)


def assert_close(got, ref, tol, what=""):
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    r = rel_l2(got, ref)
    assert r <= tol and not torch.isnan(got.double()).any(), f"{what}: " + describe_mismatch(got, ref, tol)


def to_ndhwc_bf16(x_ncdhw):
    return x_ncdhw.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)


def to_ncdhw_f32(x_ndhwc):
    return x_ndhwc.permute(0, 4, 1, 2, 3).contiguous().float()

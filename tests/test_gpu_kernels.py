"""Parity of the HBM-/latency-bound CUDA kernels against torch on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

from tests._util import assert_close, to_ncdhw_f32, to_ndhwc_bf16

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _rand_act(shape, dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(shape, generator=g).to(dev).to(BF)


@pytest.mark.parametrize("shape", [(2, 4, 6, 8, 64), (1, 3, 5, 7, 24), (3, 8, 8, 8, 512), (1, 40, 16, 16, 8)])
def test_channel_stats_and_bn_forward(cuda_dev, shape):
    from multimodal_alzheimer_b200 import kernels as K
    y = _rand_act(shape, cuda_dev) * 3 + 1
    C = shape[-1]
    rows = y.numel() // C
    st = K.channel_stats(y.view(-1, C))
    yf = y.float().view(-1, C).double()
    assert_close(st[0], yf.sum(0), 1e-5, "sum")
    assert_close(st[1], (yf * yf).sum(0), 1e-5, "sqsum")
    gamma = torch.rand(C, device=cuda_dev) + 0.5
    beta = torch.randn(C, device=cuda_dev)
    rm = torch.zeros(C, device=cuda_dev)
    rv = torch.ones(C, device=cuda_dev)
    mean, invstd, scale, shift = K.bn_finalize(st, rows, gamma, beta, 1e-5, 0.1, rm, rv)
    bn = torch.nn.BatchNorm1d(C).to(cuda_dev)
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    ref = bn(y.float().view(-1, C))
    assert_close(rm, bn.running_mean, 1e-5, "running_mean")
    assert_close(rv, bn.running_var, 1e-5, "running_var")
    res = _rand_act(shape, cuda_dev, 1)
    out = K.bn_apply(y, scale, shift, residual=res, relu=True)
    ref_out = F.relu(ref + res.float().view(-1, C)).detach()
    assert_close(out.float().view(-1, C), ref_out, 5e-3, "bn_apply+res+relu")
    out2 = K.bn_apply(y, scale, shift, residual=None, relu=False)
    assert_close(out2.float().view(-1, C), ref.detach(), 5e-3, "bn_apply")


@pytest.mark.parametrize("relu,with_res", [(True, True), (True, False), (False, False)])
@pytest.mark.parametrize("shape", [(2, 4, 6, 8, 64), (2, 8, 8, 8, 256)])
def test_bn_backward(cuda_dev, shape, relu, with_res):
    from multimodal_alzheimer_b200 import kernels as K
    C = shape[-1]
    y = _rand_act(shape, cuda_dev) * 2 + 0.5
    res = _rand_act(shape, cuda_dev, 1) if with_res else None
    dout = _rand_act(shape, cuda_dev, 2)
    rows = y.numel() // C
    gamma = torch.rand(C, device=cuda_dev) + 0.5
    beta = torch.randn(C, device=cuda_dev) * 0.1
    st = K.channel_stats(y.view(-1, C))
    mean, invstd, scale, shift = K.bn_finalize(st, rows, gamma, beta, 1e-5, 0.1, None, None)
    out = K.bn_apply(y, scale, shift, residual=res, relu=relu)
    red = K.bn_bwd_reduce(dout, out, y, mean, invstd, relu)
    dy, dres, dgamma, dbeta = K.bn_bwd_apply(dout, out, y, mean, invstd, gamma, red, rows, relu, with_res)
    # torch reference in fp32 on the same bf16-rounded inputs, using the relu mask of the kernel's own output
    yr = y.float().view(-1, C).requires_grad_(True)
    g_ = gamma.clone().requires_grad_(True)
    b_ = beta.clone().requires_grad_(True)
    z = F.batch_norm(yr, None, None, g_, b_, True, 0.1, 1e-5)
    rr = res.float().view(-1, C).requires_grad_(True) if with_res else None
    if with_res:
        z = z + rr
    g_up = dout.float().view(-1, C)
    if relu:
        g_up = g_up * (out.float().view(-1, C) > 0)
    z.backward(g_up)
    assert_close(dy.float().view(-1, C), yr.grad, 8e-3, "dy")
    assert_close(dgamma, g_.grad, 2e-3, "dgamma")
    assert_close(dbeta, b_.grad, 2e-3, "dbeta")
    if with_res:
        assert_close(dres.float().view(-1, C), rr.grad, 1e-6, "dres")


@pytest.mark.parametrize("cfg", [((2, 16, 16, 16, 64), 3, 2, 1), ((1, 9, 11, 13, 16), 3, 2, 1), ((2, 8, 8, 8, 8), 2, 2, 0)])
def test_maxpool(cuda_dev, cfg):
    from multimodal_alzheimer_b200 import kernels as K
    shape, k, s, p = cfg
    # quantise to few distinct values so that ties occur and the first-max rule is exercised
    x = (torch.randint(0, 6, shape, generator=torch.Generator().manual_seed(3)).float() - 1).to(cuda_dev).to(BF)
    y, am = K.maxpool3d_fwd(x, k, s, p)
    xr = to_ncdhw_f32(x).cpu().requires_grad_(True)   # CPU max_pool3d defines the tie rule of the oracle
    ref = F.max_pool3d(xr, k, s, p)
    assert torch.equal(to_ncdhw_f32(y).cpu(), ref.detach())
    dy = torch.randint(-3, 4, ref.shape, generator=torch.Generator().manual_seed(4)).float()
    ref.backward(dy)
    dx = K.maxpool3d_bwd(to_ndhwc_bf16(dy).to(cuda_dev), am, tuple(x.shape), k, s, p)
    assert torch.equal(to_ncdhw_f32(dx).cpu(), xr.grad)


@pytest.mark.parametrize("cfg", [((2, 8, 8, 16, 8), 2), ((1, 9, 7, 11, 16), 2), ((2, 6, 9, 12, 32), 3), ((1, 4, 4, 4, 64), 2)])
def test_relu_maxpool_fused(cuda_dev, cfg):
    """ReLU -> MaxPool3d(k) (non-overlapping, floor mode incl. ragged tails) in one pass each way vs torch on the CPU:
    pooled values and the input gradient bit-exact, ties and all-negative windows included (relu'(0) = 0)."""
    from multimodal_alzheimer_b200 import kernels as K
    shape, k = cfg
    x = (torch.randint(0, 7, shape, generator=torch.Generator().manual_seed(5)).float() - 3).to(cuda_dev).to(BF)
    y, am = K.relu_maxpool_fwd(x, k, relu=True)
    xr = to_ncdhw_f32(x).cpu().requires_grad_(True)
    ref = F.max_pool3d(torch.relu(xr), k)
    assert torch.equal(to_ncdhw_f32(y).cpu(), ref.detach())
    dy = torch.randint(-3, 4, ref.shape, generator=torch.Generator().manual_seed(6)).float()
    ref.backward(dy)
    dx = K.relu_maxpool_bwd(to_ndhwc_bf16(dy).to(cuda_dev), am, y, tuple(x.shape), k)
    assert torch.equal(to_ncdhw_f32(dx).cpu(), xr.grad)
    # plain max pool through the same kernels (relu = 0 / no mask) == the generic entry point's results
    y2, am2 = K.relu_maxpool_fwd(x, k, relu=False)
    y3, am3 = K.maxpool3d_fwd(x, k, k, 0)
    assert torch.equal(y2, y3) and torch.equal(am2, am3)
    dxa = K.maxpool3d_bwd(to_ndhwc_bf16(dy).to(cuda_dev), am3, tuple(x.shape), k, k, 0)
    xr2 = to_ncdhw_f32(x).cpu().requires_grad_(True)
    F.max_pool3d(xr2, k).backward(dy)
    assert torch.equal(to_ncdhw_f32(dxa).cpu(), xr2.grad)


def test_gap(cuda_dev):
    from multimodal_alzheimer_b200 import kernels as K
    x = _rand_act((3, 5, 6, 7, 512), cuda_dev)
    feat = K.gap_fwd(x)
    ref = x.float().mean(dim=(1, 2, 3))
    assert_close(feat, ref, 1e-5, "gap fwd")
    df = torch.randn((3, 512), device=cuda_dev)
    dx = K.gap_bwd(df, tuple(x.shape))
    refdx = (df / (5 * 6 * 7))[:, None, None, None, :].expand(x.shape)
    assert_close(dx.float(), refdx, 4e-3, "gap bwd")


@pytest.mark.parametrize("cfg", [(5, 512, 3, True), (4, 128, 64, True), (7, 1024, 512, False), (2, 64, 3, False)])
def test_linear(cuda_dev, cfg):
    from multimodal_alzheimer_b200 import kernels as K
    B, nin, nout, relu = cfg
    g = torch.Generator().manual_seed(5)
    x = torch.randn((B, nin), generator=g).to(cuda_dev)
    W = (torch.randn((nout, nin), generator=g) / nin ** 0.5).to(cuda_dev)
    b = torch.randn((nout,), generator=g).to(cuda_dev)
    dy = torch.randn((B, nout), generator=g).to(cuda_dev)
    y = K.linear_fwd(x, W, b, relu)
    xr, Wr, br = (t.clone().requires_grad_(True) for t in (x, W, b))
    ref = F.linear(xr, Wr, br)
    if relu:
        ref = F.relu(ref)
    assert_close(y, ref.detach(), 1e-5, "linear fwd")
    ref.backward(dy)
    dx, dW, db = K.linear_bwd(x, W, y, dy, relu)
    assert_close(dx, xr.grad, 1e-5, "linear dx")
    assert_close(dW, Wr.grad, 1e-5, "linear dW")
    assert_close(db, br.grad, 1e-5, "linear db")
    # concat expressed as a column slice of a wider buffer
    wide = torch.zeros((B, nout + 5), device=cuda_dev)
    K.linear_fwd(x, W, b, relu, out=wide[:, 5:])
    assert_close(wide[:, 5:], ref.detach(), 1e-5, "linear fwd into slice")
    assert float(wide[:, :5].abs().sum()) == 0


def test_bn1d(cuda_dev):
    from multimodal_alzheimer_b200 import kernels as K
    B, C = 6, 40
    g = torch.Generator().manual_seed(6)
    x = torch.randn((B, C), generator=g).to(cuda_dev) * 2 + 1
    gamma = (torch.rand(C, generator=g) + 0.5).to(cuda_dev)
    beta = torch.randn(C, generator=g).to(cuda_dev)
    dy = torch.randn((B, C), generator=g).to(cuda_dev)
    st = K.rows_stats_f32(x)
    mean, invstd, scale, shift = K.bn_finalize(st, B, gamma, beta, 1e-5, 0.1, None, None)
    y = K.bn1d_apply(x, scale, shift, True)
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, gamma, beta))
    ref = F.relu(F.batch_norm(xr, None, None, gr, br, True, 0.1, 1e-5))
    assert_close(y, ref.detach(), 1e-5, "bn1d fwd")
    ref.backward(dy)
    red = K.bn1d_bwd_reduce(dy, y, x, mean, invstd, True)
    dx, dgamma, dbeta = K.bn1d_bwd_apply(dy, y, x, mean, invstd, gamma, red, B, True)
    assert_close(dx, xr.grad, 1e-4, "bn1d dx")
    assert_close(dgamma, gr.grad, 1e-4, "bn1d dgamma")
    assert_close(dbeta, br.grad, 1e-4, "bn1d dbeta")


@pytest.mark.parametrize("gamma", [0.0, 1.0, 2.0, 5.0])
@pytest.mark.parametrize("C", [2, 3])
def test_loss(cuda_dev, gamma, C):
    """focal (detached pt, pkg/loss_functions/focalloss.py:30) / weighted CE, fp64, tolerance 1e-12."""
    from multimodal_alzheimer_b200 import kernels as K
    from oracle.losses import FocalLossOracle
    B = 9
    g = torch.Generator().manual_seed(7)
    logits = (torch.randn((B, C), generator=g) * 3).to(cuda_dev)
    target = torch.randint(0, C, (B,), generator=g).to(cuda_dev)
    cw = torch.tensor([0.4651162790697675, 0.6712473572938689, 0.8636363636363636][:C], dtype=torch.float64,
                      device=cuda_dev)
    partial, coeff = K.loss_fwd(logits, target, gamma, cw if gamma == 0 else None)
    loss = partial[0] / partial[1]
    dl = K.loss_bwd(logits, target, coeff, partial[1:2])
    zr = logits.double().clone().requires_grad_(True)
    if gamma > 0:
        ref = FocalLossOracle(gamma=gamma)(zr, target)
    else:
        ref = F.cross_entropy(zr, target, weight=cw)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-12 * max(1.0, abs(float(ref)))
    assert_close(dl, zr.grad.float(), 1e-6, "dlogits")


def _synthetic_scans(S, shape, dev, seed=15):
    g = torch.Generator().manual_seed(seed)
    v = 400 * torch.randn((S,) + shape, generator=g).abs() + 50 * torch.rand((S,) + shape, generator=g)
    D, H, W = shape
    zz, yy, xx = torch.meshgrid(torch.arange(D), torch.arange(H), torch.arange(W), indexing="ij")
    ell = (((zz - (D - 1) / 2) / (0.42 * D)) ** 2 + ((yy - (H - 1) / 2) / (0.42 * H)) ** 2 +
           ((xx - (W - 1) / 2) / (0.42 * W)) ** 2) <= 1
    mask = ell[None].expand(S, -1, -1, -1).clone()
    zero = torch.rand((S,) + shape, generator=g) < 0.005
    v[zero & mask] = 0.0
    return v.float().to(dev).contiguous(), mask.to(torch.uint8).to(dev).contiguous()


@pytest.mark.parametrize("q", [0.95, 0.98, 0.99, 1.0])
def test_quantile_normalize(cuda_dev, q):
    """Order-statistic indices bit-exact; Qmin/Qmax bit-exact (fp64); normalised volume bit-exact in fp32."""
    from multimodal_alzheimer_b200 import kernels as K
    from oracle.normalization import quantile_minmax_oracle
    x, mask = _synthetic_scans(3, (24, 28, 20), cuda_dev)
    out, info, qv = K.quantile_minmax_normalize(x, mask, q, want_info=True)
    torch.cuda.synchronize()
    for s in range(x.shape[0]):
        ref, meta = quantile_minmax_oracle(x[s].cpu().double(), mask[s].cpu().double(), q)
        assert int(info[s, 0]) == meta["n"]
        assert [int(v) for v in info[s, 1:5]] == [meta["lo_max"], meta["hi_max"], meta["lo_min"], meta["hi_min"]]
        assert float(qv[s, 0]) == meta["qmax"] and float(qv[s, 1]) == meta["qmin"]
        assert torch.equal(out[s].cpu(), ref.float())


def test_quantile_negative_and_ties(cuda_dev):
    from multimodal_alzheimer_b200 import kernels as K
    from oracle.normalization import quantile_minmax_oracle
    g = torch.Generator().manual_seed(1)
    x = torch.randint(-50, 50, (2, 10, 12, 8), generator=g).float().to(cuda_dev)   # many ties, negatives, zeros
    mask = (torch.rand((2, 10, 12, 8), generator=g) < 0.7).to(torch.uint8).to(cuda_dev)
    out, info, qv = K.quantile_minmax_normalize(x, mask, 0.9, want_info=True)
    for s in range(2):
        ref, meta = quantile_minmax_oracle(x[s].cpu().double(), mask[s].cpu().double(), 0.9)
        assert int(info[s, 0]) == meta["n"]
        assert float(qv[s, 0]) == meta["qmax"] and float(qv[s, 1]) == meta["qmin"]
        assert torch.equal(out[s].cpu(), ref.float())


def test_standardize_and_moments(cuda_dev):
    from multimodal_alzheimer_b200 import kernels as K
    from oracle.normalization import masked_std_mean_oracle, pet_standardize_oracle, split_moments_oracle
    x, mask = _synthetic_scans(2, (16, 12, 20), cuda_dev, seed=3)
    out = K.standardize(x, 0.5145, 0.5383)
    ref = pet_standardize_oracle(x.cpu().double(), 0.5145, 0.5383).float()
    assert torch.equal(out.cpu(), ref)
    # vector path + scalar tail + misaligned views, masked (dataloader.py:252-260) and bf16 outputs: all bit-exact
    flat = torch.randn(4 * 1000 + 7, generator=torch.Generator().manual_seed(9)).to(cuda_dev) * 300 + 400
    mflat = (torch.rand(flat.numel(), generator=torch.Generator().manual_seed(10)) < 0.4).to(torch.uint8).to(cuda_dev)
    for off in (0, 1, 4):
        xs, ms = flat[off:], mflat[off:]            # off = 1: misaligned views take the scalar path
        refm = (((xs.cpu().double() - 426.9336) / 1018.7830) * ms.cpu().double()).float()
        got = K.standardize(xs, 426.9336, 1018.7830, mask=ms)
        assert torch.equal(got.cpu(), refm), off
        gotb = K.standardize(xs, 426.9336, 1018.7830, mask=ms, out_dtype=BF)
        assert torch.equal(gotb.cpu(), refm.to(BF)), off
    m = K.scan_moments(x)
    mean, std, per_scan = split_moments_oracle([x[s].cpu().double() for s in range(2)])
    assert_close(m.cpu(), per_scan, 1e-12, "scan moments")
    sm = K.masked_std_mean(x, mask).cpu()
    for s in range(2):
        n, mu, sd = masked_std_mean_oracle(x[s].cpu().double(), mask[s].cpu().double())
        assert int(sm[s, 0]) == n
        assert abs(float(sm[s, 1]) - mu) <= 1e-12 * abs(mu) and abs(float(sm[s, 2]) - sd) <= 1e-12 * sd


def test_weight_layout_roundtrip(cuda_dev):
    from multimodal_alzheimer_b200 import kernels as K
    w = torch.randn((40, 24, 3, 3, 3), device=cuda_dev)
    oti, ito = K.weights_to_kernel_layout(w)
    wb = w.to(BF)
    assert torch.equal(oti, wb.view(40, 24, 27).permute(0, 2, 1).contiguous())
    assert torch.equal(ito, wb.view(40, 24, 27).permute(1, 2, 0).contiguous())
    dw = torch.randn((40, 27, 24), device=cuda_dev)
    g = K.wgrad_to_param_layout(dw, tuple(w.shape))
    assert torch.equal(g, dw.permute(0, 2, 1).contiguous().view(w.shape))


def test_casts(cuda_dev):
    from multimodal_alzheimer_b200 import kernels as K
    x = torch.randn(1000, device=cuda_dev, dtype=torch.float64)
    assert torch.equal(K.cast_to_bf16(x), x.float().to(BF))
    assert torch.equal(K.cast_to_f32(K.cast_to_bf16(x.float())), x.float().to(BF).float())


@pytest.mark.parametrize("shape", [(2, 16, 16, 16, 64), (1, 9, 11, 13, 64), (1, 32, 20, 24, 16)])
def test_fused_stem_tail_matches_unfused_kernels(cuda_dev, shape):
    """bn1 -> ReLU -> MaxPool3d(3,2,1) fused (stem_fused.cu) vs bn_apply + maxpool3d + bn_bwd_* kernels."""
    from multimodal_alzheimer_b200 import kernels as K
    C = shape[-1]
    y = _rand_act(shape, cuda_dev) * 2 + 0.3
    rows = y.numel() // C
    gamma = torch.rand(C, device=cuda_dev) + 0.5
    beta = torch.randn(C, device=cuda_dev) * 0.2
    bnp = K.bn_finalize(K.channel_stats(y.view(-1, C)), rows, gamma, beta, 1e-5, 0.1, None, None)
    a = K.bn_apply(y, bnp[2], bnp[3], residual=None, relu=True)
    p_ref, am_ref = K.maxpool3d_fwd(a, 3, 2, 1)
    p, am, yraw = K.bn_relu_maxpool_fwd(y, bnp, 3, 2, 1)
    assert torch.equal(p, p_ref) and torch.equal(am, am_ref)
    # the raw conv output at every window's arg-max: gather y at the recorded slot (kd, kh, kw) of window (od, oh, ow)
    N, D, H, W, _ = shape
    Do, Ho, Wo = p.shape[1:4]
    slot = am.long()
    od, oh, ow = torch.meshgrid(torch.arange(Do), torch.arange(Ho), torch.arange(Wo), indexing="ij")
    dev = y.device
    idd = (2 * od.to(dev) - 1)[None, ..., None] + slot // 9
    ihh = (2 * oh.to(dev) - 1)[None, ..., None] + (slot // 3) % 3
    iww = (2 * ow.to(dev) - 1)[None, ..., None] + slot % 3
    nn = torch.arange(N, device=dev)[:, None, None, None, None].expand_as(slot)
    cc = torch.arange(C, device=dev)[None, None, None, None, :].expand_as(slot)
    assert torch.equal(yraw, y[nn, idd, ihh, iww, cc])
    dp = _rand_act(tuple(p.shape), cuda_dev, 5)
    da = K.maxpool3d_bwd(dp, am_ref, tuple(a.shape), 3, 2, 1)
    red_ref = K.bn_bwd_reduce(da, a, y, bnp[0], bnp[1], True)
    dy_ref, _, _, _ = K.bn_bwd_apply(da, a, y, bnp[0], bnp[1], gamma, red_ref, rows, True, False)
    red = K.maxpool_bn_bwd_reduce(dp, am, y, bnp, 3, 2, 1)
    # the unfused path rounds the pooled gradient to bf16 before reducing; the fused one keeps it in fp32
    assert_close(red, red_ref, 6e-3, "fused reduce")
    dy = K.maxpool_bn_bwd_apply(dp, am, y, bnp, gamma, red, rows, 3, 2, 1)
    assert_close(dy.float(), dy_ref.float(), 1e-2, "fused apply")
    # fp32 torch reference of the same chain on the bf16-rounded activation
    af = a.float().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    pf = F.max_pool3d(af, 3, 2, 1)
    pf.backward(dp.float().permute(0, 4, 1, 2, 3).contiguous())
    g = (af.grad * (af > 0)).permute(0, 2, 3, 4, 1).reshape(-1, C).double()
    xh = ((y.float().view(-1, C) - bnp[0]) * bnp[1]).double()
    assert_close(red[0], g.sum(0), 1e-5, "fused reduce vs fp32 reference (sum g)")
    assert_close(red[1], (g * xh).sum(0), 1e-5, "fused reduce vs fp32 reference (sum g*xhat)")
    # the same sums over the POOLED tensor (what StemFn.backward runs): every window feeds exactly one voxel
    red_p = K.bn_bwd_reduce(dp, None, yraw, bnp[0], bnp[1], True, bnp[2], bnp[3])
    assert_close(red_p[0], g.sum(0), 1e-5, "pooled reduce vs fp32 reference (sum g)")
    assert_close(red_p[1], (g * xh).sum(0), 1e-5, "pooled reduce vs fp32 reference (sum g*xhat)")


@pytest.mark.parametrize("shape", [(2, 4, 6, 8, 64), (3, 8, 8, 8, 512), (1, 3, 5, 7, 24)])
@pytest.mark.parametrize("with_res", [False, True])
def test_bn_train_apply_equals_finalize_plus_apply(cuda_dev, shape, with_res):
    """The fused training-mode BatchNorm forward is BIT-identical to bn_finalize + bn_apply (same expression per
    channel), including bnp and the running statistics; C = 24 exercises the non-fixed-channel fallback."""
    from multimodal_alzheimer_b200 import kernels as K
    C = shape[-1]
    y = _rand_act(shape, cuda_dev) * 3 + 1
    res = _rand_act(shape, cuda_dev, 1) if with_res else None
    rows = y.numel() // C
    gamma = torch.rand(C, device=cuda_dev) + 0.5
    beta = torch.randn(C, device=cuda_dev)
    st = K.channel_stats(y.view(-1, C))
    rm0, rv0 = torch.zeros(C, device=cuda_dev), torch.ones(C, device=cuda_dev)
    rm1, rv1 = rm0.clone(), rv0.clone()
    bnp_ref = K.bn_finalize(st, rows, gamma, beta, 1e-5, 0.1, rm0, rv0)
    out_ref = K.bn_apply(y, bnp_ref[2], bnp_ref[3], residual=res, relu=True)
    out, bnp = K.bn_train_apply(y, st, rows, gamma, beta, 1e-5, 0.1, rm1, rv1, residual=res, relu=True)
    torch.cuda.synchronize()
    assert torch.equal(out, out_ref)
    assert torch.equal(bnp, bnp_ref)
    assert torch.equal(rm1, rm0) and torch.equal(rv1, rv0)


def test_bn_param_grad_scale(cuda_dev):
    """param_grad_scale = 1/world: the ranks' parameter gradients SUM to the global sums (data-parallel step)."""
    from multimodal_alzheimer_b200 import kernels as K
    shape, C = (2, 4, 4, 8, 64), 64
    y, dout = _rand_act(shape, cuda_dev), _rand_act(shape, cuda_dev, 2)
    rows = y.numel() // C
    gamma = torch.rand(C, device=cuda_dev) + 0.5
    mean, invstd, scale, shift = K.bn_finalize(K.channel_stats(y.view(-1, C)), rows, gamma, None, 1e-5, 0.1, None, None)
    red = K.bn_bwd_reduce(dout, None, y, mean, invstd, False)
    _, _, dg1, db1 = K.bn_bwd_apply(dout, None, y, mean, invstd, gamma, red, rows, False, False)
    _, _, dg4, db4 = K.bn_bwd_apply(dout, None, y, mean, invstd, gamma, red, rows, False, False, param_grad_scale=0.25)
    torch.cuda.synchronize()
    assert torch.equal(dg1, red[1].float()) and torch.equal(db1, red[0].float())
    assert_close(dg4 * 4, dg1, 1e-7, "dgamma / 4")
    assert_close(db4 * 4, db1, 1e-7, "dbeta / 4")


def test_weight_arena_matches_per_tensor_conversion(cuda_dev):
    """One-launch multi-tensor weight conversion == the per-tensor transposes (bit-exact), ragged channel counts
    included; a stale copy is never served after the parameter changed."""
    from multimodal_alzheimer_b200 import kernels as K
    g = torch.Generator().manual_seed(5)
    shapes = [(64, 64, 3), (128, 64, 3), (128, 64, 1), (40, 24, 3), (512, 256, 3), (72, 200, 1)]
    ws = [torch.nn.Parameter(torch.randn((co, ci, k, k, k), generator=g).to(cuda_dev)) for co, ci, k in shapes]
    arena = K.WeightArena()
    arena.refresh(ws)
    torch.cuda.synchronize()
    for w in ws:
        oti, ito = arena.lookup(w)
        oti_ref, ito_ref = K.weights_to_kernel_layout(w)
        assert torch.equal(oti, oti_ref), tuple(w.shape)
        assert torch.equal(ito, ito_ref), tuple(w.shape)
    with torch.no_grad():
        ws[0].add_(1.0)  # an optimizer step bumps the version counter
    assert arena.lookup(ws[0]) is None and arena.lookup(ws[1]) is not None
    arena.refresh(ws)
    assert torch.equal(arena.lookup(ws[0])[0], K.weights_to_kernel_layout(ws[0])[0])


def test_conv_plan_info(cuda_dev):
    """Planner query used by bench.py: engine routing and executed-tap fraction (all-padding taps are skipped)."""
    import ctypes
    from multimodal_alzheimer_b200._lib import call, geom

    def info(N, D, C, dil, pass_):
        kind, frac = ctypes.c_int(-1), ctypes.c_double(-1.0)
        call("adni_conv3d_plan_info", geom(N, D, D, D, C, C, 3, 1, dil, dil), pass_, ctypes.byref(kind), ctypes.byref(frac))
        return kind.value, frac.value

    assert info(2, 32, 64, 1, 0)[0] == 2 and info(2, 16, 128, 1, 1)[0] == 2          # halo-resident engine
    k4, f4 = info(2, 16, 512, 4, 0)
    assert k4 == 1 and 0.5 < f4 < 0.9                                                  # dilation 4: ~31 % of the taps skipped
    assert info(2, 16, 256, 2, 1)[0] == 1 and 0.7 < info(2, 16, 256, 2, 1)[1] <= 1.0
    assert info(2, 16, 512, 4, 2)[0] == 1 and 0.4 < info(2, 16, 512, 4, 2)[1] < 0.8        # wgrad: all-padding (box, tap group) blocks skipped
    assert info(2, 8, 8, 1, 0)[0] == 3                                                 # small channels: the mma.sync engine
    assert info(2, 8, 24, 1, 0)[0] == 0                                                # odd channel counts: CUDA cores


def test_adam_multi_tensor_matches_torch_adam(cuda_dev):
    """csrc/optimizer.cu against torch.optim.Adam on the CPU (what every configure_optimizers of the reference builds:
    one group per tensor, two learning rates, weight_decay = l2_reg; anat_cnn.py:111-128).  150 tensors = 3 launches,
    ragged sizes (numel % 4 != 0, > one 8192-element chunk), a tensor that skips a step (grad None), a misaligned
    parameter view.  fp32 arithmetic in torch's operation order: <= 2e-6 relative after 4 steps."""
    from multimodal_alzheimer_b200.optim import Adam
    g = torch.Generator().manual_seed(11)
    sizes = [(64, 64, 27), (8192 * 2 + 5,), (3,), (1,), (513, 7), (64,), (128, 64, 3, 3, 3)] + [(17 + i,) for i in range(142)]
    base = torch.randn(40, generator=g)
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in sizes] + [torch.nn.Parameter(base[1:38].clone())]
    dev_base = base.to(cuda_dev)
    our_p = [torch.nn.Parameter(p.detach().to(cuda_dev)) for p in ref_p[:-1]] + [torch.nn.Parameter(dev_base[1:38])]
    assert our_p[-1].data_ptr() % 16 != 0

    def groups(ps):
        return [{"params": p, "lr": 1e-2 if i % 3 else 1e-3} for i, p in enumerate(ps)]

    ref = torch.optim.Adam(groups(ref_p), weight_decay=1e-2)
    our = Adam(groups(our_p), weight_decay=1e-2)
    for step in range(4):
        for i, (a, b) in enumerate(zip(ref_p, our_p)):
            if i == 4 and step == 1:
                a.grad, b.grad = None, None          # this tensor's step counter must lag by one afterwards
                continue
            gr = torch.randn(a.shape, generator=g) * (10.0 ** (i % 5 - 2))
            a.grad, b.grad = gr.clone(), gr.to(cuda_dev)
        ref.step()
        our.step()
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(ref_p, our_p)):
        assert_close(b.detach().cpu(), a.detach(), 2e-6, f"param {i} {tuple(a.shape)}")
        assert_close(our.state[b]["exp_avg"].cpu(), ref.state[a]["exp_avg"], 2e-6, f"exp_avg {i}")
        assert_close(our.state[b]["exp_avg_sq"].cpu(), ref.state[a]["exp_avg_sq"], 2e-6, f"exp_avg_sq {i}")
        assert float(our.state[b]["step"]) == float(ref.state[a]["step"]) == (3.0 if i == 4 else 4.0)
    # the state dict interchanges with torch.optim.Adam: load ours into a stock optimizer and vice versa
    stock = torch.optim.Adam(groups(our_p), weight_decay=1e-2)
    stock.load_state_dict(our.state_dict())
    again = Adam(groups(our_p), weight_decay=1e-2)
    again.load_state_dict(ref.state_dict())
    for b in our_p:
        b.grad = torch.ones_like(b)
    before = [b.detach().clone() for b in our_p]
    again.step()
    assert float(again.state[our_p[0]]["step"]) == 5.0
    assert all(not torch.equal(x, b.detach()) for x, b in zip(before, our_p))


def test_adam_under_cuda_graph(cuda_dev):
    """The table rides in the kernel parameters, the step counters live on the device: replaying a captured step
    advances the bias correction exactly like eager steps."""
    from multimodal_alzheimer_b200.optim import Adam
    g = torch.Generator().manual_seed(3)
    init = [torch.randn(s, generator=g) for s in [(1000,), (64, 9)]]
    grads = [torch.randn(t.shape, generator=g) for t in init]
    graph_p = [torch.nn.Parameter(t.to(cuda_dev)) for t in init]
    dev_grads = [t.to(cuda_dev) for t in grads]
    graphed = Adam(graph_p, lr=1e-2)
    for p, gr in zip(graph_p, dev_grads):
        p.grad = gr.clone()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        graphed.step()                               # allocates the state outside the capture
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg, stream=s):
            graphed.step()
        for _ in range(3):
            cg.replay()
    torch.cuda.synchronize()
    # capture does not execute: the graphed optimizer ran 1 eager step + 3 replays = 4 updates
    assert float(graphed.state[graph_p[0]]["step"]) == 4.0
    eager4_p = [torch.nn.Parameter(t.to(cuda_dev)) for t in init]
    e4 = Adam(eager4_p, lr=1e-2)
    for p, gr in zip(eager4_p, dev_grads):
        p.grad = gr.clone()
    for _ in range(4):
        e4.step()
    torch.cuda.synchronize()
    for a, b in zip(eager4_p, graph_p):
        assert torch.equal(a.detach(), b.detach())


def test_adam_graph_follows_lr_schedule_and_bumps_versions(cuda_dev):
    """lr / weight decay are read from a device table refreshed through a pinned host buffer, so a captured
    graph trains with the rate ReduceLROnPlateau last set (after `refresh_hyperparameters()`), and every step bumps the
    parameters' version counters (the weight arena's staleness guard)."""
    from multimodal_alzheimer_b200.optim import Adam
    g = torch.Generator().manual_seed(5)
    init = [torch.randn(s, generator=g) for s in [(700,), (33, 5)]]
    grads = [torch.randn(t.shape, generator=g).to(cuda_dev) for t in init]

    def make():
        ps = [torch.nn.Parameter(t.to(cuda_dev)) for t in init]
        for p, gr in zip(ps, grads):
            p.grad = gr.clone()
        return ps, Adam([{"params": ps[0], "lr": 1e-2}, {"params": ps[1], "lr": 1e-3}], weight_decay=1e-2)

    gp, gopt = make()
    ep, eopt = make()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        v0 = gp[0]._version
        gopt.step()
        assert gp[0]._version > v0
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg, stream=s):
            gopt.step()
        cg.replay()
        for group in gopt.param_groups:          # what ReduceLROnPlateau(factor=0.1) does
            group["lr"] *= 0.1
        gopt.refresh_hyperparameters()
        cg.replay()
        cg.replay()
    for k in range(4):
        if k == 2:
            for group in eopt.param_groups:
                group["lr"] *= 0.1
        eopt.step()
    torch.cuda.synchronize()
    for a, b in zip(ep, gp):
        assert torch.equal(a.detach(), b.detach())


def test_fusion_ops_bit_exact(cuda_dev):
    """csrc/fusion_ops.cu against torch on the same bf16 values: multi-channel input cast (early_fusion.py:84-88),
    maxout with torch.max's first-index tie rule and its gradient routing (anat_pet_featuremapfusion.py:121-123),
    channel concatenation and the split of its gradient (:118-119).  Pure data movement / comparisons: bit-exact."""
    from multimodal_alzheimer_b200 import kernels as K
    g = torch.Generator().manual_seed(21)
    for dtype in (torch.float32, torch.float64):
        for C in (2, 3, 4):
            x = torch.randn((3, C, 5, 6, 7), generator=g, dtype=dtype)
            got = K.volumes_to_ndhwc(x.to(cuda_dev))
            ref = x.float().permute(0, 2, 3, 4, 1).contiguous().to(BF)
            assert torch.equal(got.cpu(), ref), (dtype, C)
    with pytest.raises(NotImplementedError):
        K.volumes_to_ndhwc(torch.zeros((1, 5, 2, 2, 2), device=cuda_dev))
    shape = (2, 3, 5, 4, 16)
    a = torch.randn(shape, generator=g).clamp_min(0).to(BF)          # post-ReLU maps: many exact ties at zero
    b = torch.randn(shape, generator=g).clamp_min(0).to(BF)
    b[0, 0] = a[0, 0]                                                # and ties at non-zero values
    dout = torch.randn(shape, generator=g).to(BF)
    af, bf_ = a.float().requires_grad_(), b.float().requires_grad_()
    ref, _ = torch.max(torch.stack((af, bf_), dim=0), dim=0)
    ref.backward(dout.float())
    ad, bd = a.to(cuda_dev), b.to(cuda_dev)
    out = K.maxout_fwd(ad, bd)
    da, db = K.maxout_bwd(dout.to(cuda_dev), ad, bd)
    assert torch.equal(out.cpu().float(), ref.detach())
    assert torch.equal(da.cpu().float(), af.grad) and torch.equal(db.cpu().float(), bf_.grad)
    only_b = K.maxout_bwd(dout.to(cuda_dev), ad, bd, want_a=False)
    assert only_b[0] is None and torch.equal(only_b[1], db)
    c = torch.randn(shape[:-1] + (24,), generator=g).to(BF)
    cat = K.concat_channels(ad, c.to(cuda_dev))
    assert torch.equal(cat.cpu(), torch.cat((a, c), dim=-1))
    sa, sc = K.split_channels(cat, 16, 24)
    assert torch.equal(sa.cpu(), a) and torch.equal(sc.cpu(), c)
    with pytest.raises(NotImplementedError):
        K.concat_channels(ad[..., :4].contiguous(), ad)


@pytest.mark.parametrize("C", [64, 8, 2])
def test_even_kernel_same_padding(cuda_dev, C):
    """nn.Conv3d(C, Cout, 4, padding='same') (filter_size_fusion = 4): torch pads 1 low / 2 high.  The pad / crop kernels
    are bit-exact data movement; the module matches torch's conv on the same bf16 operands (fwd, dgrad, wgrad)."""
    from multimodal_alzheimer_b200 import kernels as K
    from multimodal_alzheimer_b200 import nn as bnn
    g = torch.Generator().manual_seed(31)
    x = torch.randn((2, 3, 5, 4, C), generator=g).to(BF)
    xp = K.pad_volume_high(x.to(cuda_dev), 1)
    ref = F.pad(x, (0, 0, 0, 1, 0, 1, 0, 1))
    assert torch.equal(xp.cpu(), ref)
    assert torch.equal(K.crop_volume_high(xp, 1).cpu(), x)
    conv = bnn.Conv3d(C, 16, 4, padding="same").to(cuda_dev)
    tconv = torch.nn.Conv3d(C, 16, 4, padding="same")
    with torch.no_grad():
        tconv.weight.copy_(conv.weight.detach().cpu().to(BF).float())
        tconv.bias.copy_(conv.bias.detach().cpu())
    xin = x.to(cuda_dev).requires_grad_(True)
    y = conv(xin)
    xr = to_ncdhw_f32(x).requires_grad_(True)
    yr = tconv(xr)
    assert tuple(y.shape) == (2, 3, 5, 4, 16)
    assert_close(to_ncdhw_f32(y.detach().cpu()), yr.detach(), 6e-3, "even-kernel fprop")
    dy = torch.randn(yr.shape, generator=g)
    dy_b = to_ndhwc_bf16(dy)
    yr.backward(to_ncdhw_f32(dy_b))
    y.backward(dy_b.to(cuda_dev))
    torch.cuda.synchronize()
    assert_close(to_ncdhw_f32(xin.grad.cpu()), xr.grad, 6e-3, "even-kernel dgrad")
    assert_close(conv.weight.grad.cpu(), tconv.weight.grad, 2e-3, "even-kernel wgrad")
    assert_close(conv.bias.grad.cpu(), tconv.bias.grad, 2e-3, "even-kernel dbias")


@pytest.mark.parametrize("C,n", [(2, 7), (3, 150), (3, 1000), (2, 333)])
def test_bootstrap_metrics_match_oracle(cuda_dev, C, n):
    """adni_bootstrap_metrics (all resamples in one launch) vs oracle/metrics.py (torchmetrics 0.10.2 reductions
    restated, cross-checked against scikit-learn): confusion matrices exact, F1 (macro, per class) and MCC to 1e-6
    (fp32 reductions in the same order), including resamples in which a class is missing."""
    from multimodal_alzheimer_b200 import kernels as K
    from oracle.metrics import confusion_matrix, f1_from_confmat, mcc_from_confmat
    g = torch.Generator().manual_seed(15 + n)
    logits = torch.randn((n, C), generator=g, dtype=torch.float64)
    logits[::5, 0] = logits[::5, 1]                                   # ties: argmax takes the first maximum
    labels = torch.randint(0, C, (n,), generator=g)
    draws = torch.stack([torch.randint(0, n, (n,), generator=g) for _ in range(64)])
    draws[1] = draws[1, 0]                                            # a resample that holds a single sample
    got = K.bootstrap_metrics(logits.to(cuda_dev), labels.to(cuda_dev), draws.to(cuda_dev), want_confmat=True)
    whole = K.bootstrap_metrics(logits.to(cuda_dev), labels.to(cuda_dev), None, want_confmat=True)
    torch.cuda.synchronize()
    for d in range(64):
        cm = confusion_matrix(logits[draws[d]], labels[draws[d]], C)
        assert torch.equal(got["confmat"][d].cpu(), cm), d
        macro, per = f1_from_confmat(cm)
        assert abs(float(got["f1"][d]) - float(macro)) <= 1e-6, (d, float(got["f1"][d]), float(macro))
        assert torch.allclose(got["f1_class"][d].cpu(), per, atol=1e-6), d
        assert abs(float(got["mcc"][d]) - float(mcc_from_confmat(cm))) <= 1e-6, d
    cm = confusion_matrix(logits, labels, C)
    assert torch.equal(whole["confmat"][0].cpu(), cm)
    assert abs(float(whole["f1"][0]) - float(f1_from_confmat(cm)[0])) <= 1e-6


def test_bootstrap_metric_method_follows_reference_draws(cuda_dev):
    """Base_Model.bootstrap_metric / test_epoch_metrics: same `torch.randint` draws as base_model.py:212-236 for the
    same global seed -> the oracle's mean and 1.96 x std."""
    from multimodal_alzheimer_b200.pkg.models.pet_models.pet_cnn import Small_PET_CNN
    from oracle.metrics import bootstrap_metric
    from tests._models import hp_pet
    model = Small_PET_CNN(hp_pet(3, conv_out=(8,), filter_size=(3,)))
    g = torch.Generator().manual_seed(3)
    y_hat = torch.randn((90, 3), generator=g, dtype=torch.float64)
    y = torch.randint(0, 3, (90,), generator=g)
    for metric in ("f1", "mcc"):
        torch.manual_seed(7)
        mean_o, ci_o, _ = bootstrap_metric(metric, y_hat, y, 3, n_drawings=200)
        torch.manual_seed(7)
        mean_p, ci_p = model.bootstrap_metric(metric, y_hat.to(cuda_dev), y.to(cuda_dev), n_drawings=200)
        assert abs(float(mean_p) - float(mean_o)) <= 1e-6 and abs(float(ci_p) - float(ci_o)) <= 1e-6, metric
    outs = [{"loss": torch.tensor(0.5 + 0.1 * i, dtype=torch.float64, device=cuda_dev), "outputs": y_hat[30 * i:30 * i + 30].to(cuda_dev),
             "labels": y[30 * i:30 * i + 30].to(cuda_dev)} for i in range(3)]
    torch.manual_seed(1)
    log = model.test_epoch_metrics(outs)
    assert abs(float(log["test_loss_epoch"]) - 0.6) < 1e-12 and int(log["confusion_matrix"].sum()) == 90
    assert {"test_f1_epoch", "test_f1_epoch_class_2", "test_f1_epoch_boot", "test_f1_epoch_ci", "test_mcc_epoch_boot",
            "test_mcc_epoch_ci"} <= set(log)
    # epoch-end hooks (base_model.py:91-133): the keys ReduceLROnPlateau / EarlyStopping / ModelCheckpoint monitor
    from oracle.metrics import confusion_matrix, f1_from_confmat
    model.validation_epoch_end(outs)
    model.training_epoch_end(outs)
    cm = confusion_matrix(y_hat, y, 3)
    macro, per = f1_from_confmat(cm)
    for mode in ("val", "train"):
        assert abs(float(model.logged[f"{mode}_loss_epoch"]) - 0.6) < 1e-12
        assert abs(float(model.logged[f"{mode}_f1_epoch"]) - float(macro)) <= 1e-6
        for i in range(3):
            assert abs(float(model.logged[f"{mode}_f1_epoch_class_{i}"]) - float(per[i])) <= 1e-6
        assert torch.equal(model.last_confusion_matrix[mode], cm)
    torch.manual_seed(1)
    model.test_epoch_end(outs)
    assert abs(float(model.logged["test_f1_epoch"]) - float(macro)) <= 1e-6 and "test_mcc_epoch_ci" in model.logged
    assert torch.equal(model.last_confusion_matrix["test"], cm)

"""Parity of the CUDA Conv3d engines (tcgen05 implicit GEMM and direct) against torch fp32 conv on the
same bf16-rounded operands.  Tolerances: outputs are bf16-rounded fp32 accumulations, so rel-L2 <= 6e-3
(bf16 has 8 mantissa bits: half-ulp relative error 2^-9 = 2e-3 per element); wgrad is fp32: <= 2e-4."""
import pytest
import torch
import torch.nn.functional as F

from tests._util import assert_close, to_ncdhw_f32, to_ndhwc_bf16

pytestmark = pytest.mark.gpu

TC, DIRECT = 1, 2

# (N, D, H, W, Cin, Cout, k, stride, pad, dil)
TC_SHAPES = [
    (1, 8, 8, 16, 64, 64, 1, 1, 0, 1),      # 1x1x1 = plain GEMM, one tile
    (1, 16, 16, 16, 64, 64, 3, 1, 1, 1),    # layer1-like
    (2, 16, 16, 16, 128, 256, 3, 1, 2, 2),  # layer3.0.conv1 (dilation 2)
    (1, 16, 16, 16, 256, 512, 3, 1, 4, 4),  # layer4.0.conv1 (dilation 4, tap skipping)
    (2, 16, 16, 16, 64, 128, 3, 2, 1, 1),   # layer2.0.conv1 (stride 2, parity views)
    (2, 16, 16, 16, 64, 128, 1, 2, 0, 1),   # layer2.0.downsample
    (1, 12, 14, 12, 128, 128, 3, 1, 1, 1),  # ragged extents (91x109x91 input family)
    (1, 11, 13, 9, 64, 64, 3, 2, 1, 1),     # odd extents with stride 2
    (3, 8, 8, 8, 192, 64, 3, 1, 1, 1),      # Cin not a power of two
    (20, 33, 16, 8, 64, 64, 3, 1, 1, 1),    # halo engine: long columns, odd depth (unpaired last piece), many CTAs
    (9, 21, 16, 16, 128, 128, 3, 1, 1, 1),  # halo engine, two K blocks, several pieces per CTA
    (1, 5, 40, 20, 64, 64, 3, 1, 1, 1),     # halo engine: three H tiles (ragged), more CTAs than planes per column
    (2, 10, 12, 10, 256, 64, 1, 1, 0, 1),   # flat 1x1x1 path: 2400 positions (ragged last tile), K = 256, N = 64
    (3, 5, 6, 7, 64, 512, 1, 1, 0, 1),      # flat path: two 256-channel tiles, 630 positions
    (1, 4, 4, 4, 128, 128, 1, 1, 0, 1),     # flat path: fewer positions than one tile
    (2, 9, 10, 11, 1024, 256, 1, 1, 0, 1),  # flat path: 16 K blocks per tile (ResNet-50 layer3 reduce conv)
]
# feature maps smaller than one tile: the fusion conv of PET_MRI_FMF (anat_pet_featuremapfusion.py:75-81) sees 8^3 maps
# for 128^3 inputs, 5x6x5 for the MNI 91x109x91 grid and 2^3 for the 32^3 test volumes
TINY_SHAPES = [
    (4, 2, 2, 2, 64, 64, 3, 1, 1, 1),
    (2, 5, 6, 5, 128, 128, 3, 1, 1, 1),
    (2, 1, 3, 2, 128, 64, 3, 1, 1, 1),
    (3, 4, 4, 4, 64, 64, 3, 1, 1, 1),
]


def _mk(shape, dev, seed=0):
    N, D, H, W, Cin, Cout, k, s, p, d = shape
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((N, Cin, D, H, W), generator=g).to(dev)
    w = (torch.randn((Cout, Cin, k, k, k), generator=g) / (Cin * k ** 3) ** 0.5).to(dev)
    x_b = to_ndhwc_bf16(x)
    x_ref = to_ncdhw_f32(x_b)
    w_ref = w.to(torch.bfloat16).float()
    return x_b, x_ref, w, w_ref


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tc_fprop(cuda_dev, shape):
    from multimodal_alzheimer_b200 import kernels as K
    N, D, H, W, Cin, Cout, k, s, p, d = shape
    x_b, x_ref, w, w_ref = _mk(shape, cuda_dev)
    oti, _ = K.weights_to_kernel_layout(w)
    y, st = K.conv3d_fprop(x_b, oti, None, k, s, p, d, stats=True, engine=TC)
    torch.cuda.synchronize()
    ref = F.conv3d(x_ref, w_ref, None, s, p, d)
    assert_close(to_ncdhw_f32(y), ref, 6e-3, f"fprop {shape}")
    ssum = ref.sum(dim=(0, 2, 3, 4)).double()
    ssq = (ref.double() ** 2).sum(dim=(0, 2, 3, 4))
    if k == 1 and s == 1:
        # flat 1x1x1 path: the fused statistics are the sums of the STORED bf16 tensor (what BatchNorm then normalises)
        stored = to_ncdhw_f32(y).double()
        assert_close(st[0], stored.sum(dim=(0, 2, 3, 4)), 1e-5, f"fprop stats sum of the stored tensor {shape}")
        assert_close(st[1], (stored ** 2).sum(dim=(0, 2, 3, 4)), 1e-5, f"fprop stats sqsum of the stored tensor {shape}")
        assert_close(st[1], ssq, 1e-3, f"fprop stats sqsum {shape}")
        return
    assert_close(st[0], ssum, 1e-3, f"fprop stats sum {shape}") if float(ssum.norm()) > 1 else None
    assert_close(st[1], ssq, 1e-4, f"fprop stats sqsum {shape}")


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tc_dgrad(cuda_dev, shape):
    from multimodal_alzheimer_b200 import kernels as K
    N, D, H, W, Cin, Cout, k, s, p, d = shape
    x_b, x_ref, w, w_ref = _mk(shape, cuda_dev)
    _, ito = K.weights_to_kernel_layout(w)
    x_ref.requires_grad_(True)
    ref_y = F.conv3d(x_ref, w_ref, None, s, p, d)
    dy = torch.randn_like(ref_y)
    dy_b = to_ndhwc_bf16(dy)
    ref_y.backward(to_ncdhw_f32(dy_b))
    add = to_ndhwc_bf16(torch.randn_like(x_ref))
    dx = K.conv3d_dgrad(dy_b, ito, tuple(x_b.shape), k, s, p, d, addend=add, engine=TC)
    torch.cuda.synchronize()
    assert_close(to_ncdhw_f32(dx), x_ref.grad + to_ncdhw_f32(add), 6e-3, f"dgrad {shape}")


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tc_wgrad(cuda_dev, shape):
    from multimodal_alzheimer_b200 import kernels as K
    N, D, H, W, Cin, Cout, k, s, p, d = shape
    x_b, x_ref, w, w_ref = _mk(shape, cuda_dev)
    w_ref.requires_grad_(True)
    ref_y = F.conv3d(x_ref, w_ref, None, s, p, d)
    dy_b = to_ndhwc_bf16(torch.randn_like(ref_y))
    ref_y.backward(to_ncdhw_f32(dy_b))
    dw, _ = K.conv3d_wgrad(x_b, dy_b, k, s, p, d, engine=TC)
    g = K.wgrad_to_param_layout(dw, tuple(w.shape))
    torch.cuda.synchronize()
    assert_close(g, w_ref.grad, 2e-4, f"wgrad {shape}")


@pytest.mark.parametrize("shape,chunk_mb", [
    ((6, 16, 16, 16, 128, 256, 3, 1, 2, 2), 1),      # 3 MB of dY + X per sample: one sample per chunk, 6 chunks
    ((5, 12, 14, 12, 256, 512, 3, 2, 1, 1), 1),      # strided (8 parity views), ragged, chunked
    ((32, 16, 16, 16, 512, 512, 3, 1, 4, 4), 48),    # layer4 at the bench batch: the default budget gives 6 chunks
])
def test_tc_wgrad_chunked_stream_k(cuda_dev, shape, chunk_mb, monkeypatch):
    """The chunk-major stream-K schedule (position chunks sized for the L2, conv_api.cu plan_wgrad_stream_k) against
    torch and against the single-chunk schedule of the same kernel."""
    from multimodal_alzheimer_b200 import kernels as K
    N, D, H, W, Cin, Cout, k, s, p, d = shape
    x_b, x_ref, w, w_ref = _mk(shape, cuda_dev)
    w_ref.requires_grad_(True)
    ref_y = F.conv3d(x_ref, w_ref, None, s, p, d)
    dy_b = to_ndhwc_bf16(torch.randn_like(ref_y))
    ref_y.backward(to_ncdhw_f32(dy_b))
    monkeypatch.setenv("ADNI_WGRAD_CHUNK_MB", str(chunk_mb))
    dw, _ = K.conv3d_wgrad(x_b, dy_b, k, s, p, d, engine=TC)
    g = K.wgrad_to_param_layout(dw, tuple(w.shape))
    monkeypatch.setenv("ADNI_WGRAD_CHUNK_MB", "1000000")
    dw1, _ = K.conv3d_wgrad(x_b, dy_b, k, s, p, d, engine=TC)
    g1 = K.wgrad_to_param_layout(dw1, tuple(w.shape))
    torch.cuda.synchronize()
    assert_close(g, w_ref.grad, 2e-4, f"chunked wgrad {shape}")
    assert_close(g, g1, 1e-4, f"chunked vs single-chunk wgrad {shape}")   # fp32 partial sums in a different order


@pytest.mark.parametrize("shape", [(6, 16, 16, 16, 128, 256, 3, 1, 2, 2), (4, 16, 16, 16, 64, 64, 3, 1, 1, 1),
                                   (32, 16, 16, 16, 512, 512, 3, 1, 4, 4)])
def test_tc_wgrad_deterministic_option(cuda_dev, shape, monkeypatch):
    """ADNI_WGRAD_DETERMINISTIC=1 (one CTA per output tile over all positions; layer1 leaves the halo-plane kernel): two
    runs are bit-identical and agree with torch; the default stream-K schedule accumulates through red.add in an
    order that varies from run to run."""
    from multimodal_alzheimer_b200 import kernels as K
    N, D, H, W, Cin, Cout, k, s, p, d = shape
    x_b, x_ref, w, w_ref = _mk(shape, cuda_dev)
    w_ref.requires_grad_(True)
    ref_y = F.conv3d(x_ref, w_ref, None, s, p, d)
    dy_b = to_ndhwc_bf16(torch.randn_like(ref_y))
    ref_y.backward(to_ncdhw_f32(dy_b))
    monkeypatch.setenv("ADNI_WGRAD_DETERMINISTIC", "1")
    K._PLAN_CACHE.clear()
    runs = []
    for _ in range(3):
        dw, _ = K.conv3d_wgrad(x_b, dy_b, k, s, p, d, engine=TC)
        runs.append(K.wgrad_to_param_layout(dw, tuple(w.shape)).clone())
    torch.cuda.synchronize()
    K._PLAN_CACHE.clear()
    assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
    assert_close(runs[0], w_ref.grad, 2e-4, f"deterministic wgrad {shape}")


BNRED_SHAPES = [TC_SHAPES[1], TC_SHAPES[2], TC_SHAPES[4], TC_SHAPES[5], TC_SHAPES[6], TC_SHAPES[9], TC_SHAPES[10],
                (2, 10, 12, 10, 256, 64, 1, 1, 0, 1)]   # + a Bottleneck 1x1 reduce conv (dx has 256 channels)


@pytest.mark.parametrize("mask_mode", ["relu_out", "recompute", "none"])
@pytest.mark.parametrize("shape", BNRED_SHAPES)
def test_tc_dgrad_with_fused_bn_backward_sums(cuda_dev, shape, mask_mode):
    """adni_conv3d_dgrad_bnred: dx is bit-identical to adni_conv3d_dgrad, and the epilogue's per-channel sums equal
    sum g and sum g*y over the STORED bf16 dx with the preceding layer's ReLU mask (fp64 reference)."""
    from multimodal_alzheimer_b200 import kernels as K
    N, D, H, W, Cin, Cout, k, s, p, d = shape
    x_b, x_ref, w, w_ref = _mk(shape, cuda_dev)
    _, ito = K.weights_to_kernel_layout(w)
    Do, Ho, Wo = (K.out_extent(v, k, s, p, d) for v in (D, H, W))
    g = torch.Generator().manual_seed(5)
    dy_b = torch.randn((N, Do, Ho, Wo, Cout), generator=g).to(cuda_dev).to(torch.bfloat16)
    add = torch.randn((N, D, H, W, Cin), generator=g).to(cuda_dev).to(torch.bfloat16)
    bn_y = (torch.randn((N, D, H, W, Cin), generator=g) * 1.5 + 0.3).to(cuda_dev).to(torch.bfloat16)
    scale = (torch.rand(Cin, generator=g) * 2 - 0.5).to(cuda_dev)         # some negative scales
    shift = (torch.randn(Cin, generator=g) * 0.5).to(cuda_dev)
    relu_out = torch.relu(torch.randn((N, D, H, W, Cin), generator=g)).to(cuda_dev).to(torch.bfloat16)
    kw = dict(relu_out=dict(bn_relu_out=relu_out), recompute=dict(bn_scale=scale, bn_shift=shift), none={})[mask_mode]
    dx_ref = K.conv3d_dgrad(dy_b, ito, tuple(x_b.shape), k, s, p, d, addend=add)
    dx, red = K.conv3d_dgrad_bnred(dy_b, ito, tuple(x_b.shape), k, s, p, d, bn_y, addend=add, **kw)
    torch.cuda.synchronize()
    assert torch.equal(dx, dx_ref)
    gq, yq = dx.double(), bn_y.double()
    if mask_mode == "relu_out":
        gq = gq * (relu_out > 0)
    elif mask_mode == "recompute":
        gq = gq * (torch.addcmul(shift, bn_y.float(), scale) > 0)          # fmaf(y, scale, shift) in fp32, like the kernel
    assert_close(red[0], gq.sum(dim=(0, 1, 2, 3)), 2e-5, f"sum g {shape} {mask_mode}")
    assert_close(red[1], (gq * yq).sum(dim=(0, 1, 2, 3)), 2e-5, f"sum g*y {shape} {mask_mode}")


@pytest.mark.parametrize("shape", TINY_SHAPES)
def test_tc_tiny_feature_maps(cuda_dev, shape):
    """fprop + dgrad + wgrad on feature maps smaller than one tile, through the planner's own engine choice."""
    test_tc_fprop(cuda_dev, shape)
    test_tc_dgrad(cuda_dev, shape)
    test_tc_wgrad(cuda_dev, shape)


DIRECT_SHAPES = [
    (2, 16, 16, 16, 1, 8, 5, 1, 2, 1),    # Small_PET_CNN conv1
    (2, 8, 8, 8, 8, 16, 5, 1, 2, 1),      # conv2
    (1, 8, 8, 8, 16, 32, 3, 1, 1, 1),     # conv3
    (1, 8, 8, 8, 32, 64, 3, 1, 1, 1),     # conv4
    (1, 18, 20, 22, 1, 64, 7, 2, 3, 1),   # ResNet stem
    (1, 9, 9, 9, 24, 40, 3, 2, 1, 2),     # odd everything
]


@pytest.mark.parametrize("shape", DIRECT_SHAPES)
def test_direct_conv(cuda_dev, shape):
    from multimodal_alzheimer_b200 import kernels as K
    N, D, H, W, Cin, Cout, k, s, p, d = shape
    x_b, x_ref, w, w_ref = _mk(shape, cuda_dev)
    bias = torch.randn(Cout, device=cuda_dev)
    oti, ito = K.weights_to_kernel_layout(w)
    x_ref.requires_grad_(True)
    w_ref.requires_grad_(True)
    b_ref = bias.clone().requires_grad_(True)
    ref = F.conv3d(x_ref, w_ref, b_ref, s, p, d)
    y, st = K.conv3d_fprop(x_b, oti, bias, k, s, p, d, stats=True, engine=DIRECT)
    assert_close(to_ncdhw_f32(y), ref.detach(), 6e-3, f"direct fprop {shape}")
    assert_close(st[1], (ref.detach().double() ** 2).sum(dim=(0, 2, 3, 4)), 1e-4, f"direct stats {shape}")
    dy_b = to_ndhwc_bf16(torch.randn_like(ref))
    ref.backward(to_ncdhw_f32(dy_b))
    dx = K.conv3d_dgrad(dy_b, ito, tuple(x_b.shape), k, s, p, d, engine=DIRECT)
    assert_close(to_ncdhw_f32(dx), x_ref.grad, 6e-3, f"direct dgrad {shape}")
    dw, db = K.conv3d_wgrad(x_b, dy_b, k, s, p, d, want_dbias=True, engine=DIRECT)
    assert_close(K.wgrad_to_param_layout(dw, tuple(w.shape)), w_ref.grad, 2e-4, f"direct wgrad {shape}")
    assert_close(db, b_ref.grad, 2e-4, f"direct dbias {shape}")


SMALL = 3
# the small-channel stacks (pet_cnn.py:18-28, early_fusion.py:34-44, anat_pet_featuremapfusion.py:40-64) on the
# mma.sync engine: every (Cin, Cout, k) of conv_out in {(8,16,32,64), (16,32,64,..), (32,64,..)} x filter sizes {3,5,7}
SMALL_SHAPES = [
    (2, 16, 16, 16, 1, 8, 5, 1, 2, 1),     # Small_PET_CNN conv1 (on-chip window expansion)
    (1, 12, 20, 33, 1, 16, 7, 1, 3, 1),    # first layer with filter 7, ragged extents, two w tiles + tail
    (2, 9, 10, 17, 1, 32, 3, 1, 1, 1),     # first layer, filter 3
    (2, 8, 8, 16, 8, 16, 5, 1, 2, 1),      # conv2
    (1, 13, 11, 19, 8, 16, 5, 1, 2, 1),    # conv2, ragged
    (1, 8, 8, 8, 16, 32, 3, 1, 1, 1),      # conv3
    (2, 6, 9, 20, 16, 32, 5, 1, 2, 1),     # filter (5,5,5,3): third layer with k = 5
    (1, 8, 8, 8, 32, 64, 3, 1, 1, 1),      # conv4
    (3, 5, 7, 18, 32, 64, 3, 1, 1, 1),     # conv4, ragged
    (1, 7, 8, 16, 64, 32, 3, 1, 1, 1),     # 64 sliding channels (the dgrad of conv4 seen as a forward conv)
    (1, 6, 9, 17, 32, 64, 5, 1, 2, 1),     # long K (125 taps x 32 channels): output channels split over blockIdx.y
    (2, 9, 9, 17, 8, 8, 4, 1, 2, 1),       # even kernel after the high-side pad (output extent = input - 1 + ... )
]


@pytest.mark.parametrize("shape", SMALL_SHAPES)
def test_small_channel_engine(cuda_dev, shape):
    """fprop (+bias, +BatchNorm sums), dgrad and wgrad of the mma.sync engine vs torch fp32 on the same bf16 operands."""
    from multimodal_alzheimer_b200 import kernels as K
    N, D, H, W, Cin, Cout, k, s, p, d = shape
    x_b, x_ref, w, w_ref = _mk(shape, cuda_dev, seed=3)
    bias = torch.randn(Cout, device=cuda_dev)
    oti, ito = K.weights_to_kernel_layout(w)
    x_ref.requires_grad_(True)
    w_ref.requires_grad_(True)
    ref = F.conv3d(x_ref, w_ref, bias, s, p, d)
    y, st = K.conv3d_fprop(x_b, oti, bias, k, s, p, d, stats=True, engine=SMALL)
    torch.cuda.synchronize()
    assert_close(to_ncdhw_f32(y), ref.detach(), 6e-3, f"small fprop {shape}")
    assert_close(st[0], ref.detach().double().sum(dim=(0, 2, 3, 4)), 1e-3, f"small stats sum {shape}")
    assert_close(st[1], (ref.detach().double() ** 2).sum(dim=(0, 2, 3, 4)), 1e-4, f"small stats sqsum {shape}")
    dy_b = to_ndhwc_bf16(torch.randn_like(ref))
    ref.backward(to_ncdhw_f32(dy_b))
    if Cin != 1 and K._plan(K.geom(N, D, H, W, Cin, Cout, k, s, p, d), 1)[0] == 3:
        dx = K.conv3d_dgrad(dy_b, ito, tuple(x_b.shape), k, s, p, d, engine=SMALL)
        assert_close(to_ncdhw_f32(dx), x_ref.grad, 6e-3, f"small dgrad {shape}")
    elif Cin != 1:      # 64 sliding channels x 125 taps: tile + weights exceed shared memory -> CUDA-core engine (AUTO)
        with pytest.raises(NotImplementedError):
            K.conv3d_dgrad(dy_b, ito, tuple(x_b.shape), k, s, p, d, engine=SMALL)
        dx = K.conv3d_dgrad(dy_b, ito, tuple(x_b.shape), k, s, p, d)
        assert_close(to_ncdhw_f32(dx), x_ref.grad, 6e-3, f"auto dgrad {shape}")
    dw, db = K.conv3d_wgrad(x_b, dy_b, k, s, p, d, want_dbias=True, engine=SMALL)
    torch.cuda.synchronize()
    assert_close(K.wgrad_to_param_layout(dw, tuple(w.shape)), w_ref.grad, 2e-4, f"small wgrad {shape}")
    assert_close(db, to_ncdhw_f32(dy_b).sum(dim=(0, 2, 3, 4)), 2e-4, f"small dbias {shape}")


def test_engine_rejects_unsupported(cuda_dev):
    from multimodal_alzheimer_b200 import kernels as K
    x = torch.zeros((1, 4, 4, 4, 8), dtype=torch.bfloat16, device=cuda_dev)
    w = torch.zeros((8, 27, 8), dtype=torch.bfloat16, device=cuda_dev)
    with pytest.raises(NotImplementedError):
        K.conv3d_fprop(x, w, None, 3, 1, 1, 1, engine=TC)
    with pytest.raises(NotImplementedError):                       # the small-channel engine is stride 1 only
        K.conv3d_fprop(x, w, None, 3, 2, 1, 1, engine=SMALL)


STEM_SHAPES = [(2, 16, 16, 16), (1, 32, 24, 40), (1, 18, 20, 22), (3, 9, 11, 13), (1, 128, 128, 128)]


@pytest.mark.parametrize("shape", STEM_SHAPES)
def test_tc_stem(cuda_dev, shape):
    """Conv3d(1,64,k=7,s=2,p=3) on tcgen05 through the W-window ("X8") expansion: fprop + BN sums + wgrad."""
    from multimodal_alzheimer_b200 import kernels as K
    N, D, H, W = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn((N, 1, D, H, W), generator=g).to(cuda_dev)
    w = (torch.randn((64, 1, 7, 7, 7), generator=g) / 343 ** 0.5).to(cuda_dev)
    x_b = x.to(torch.bfloat16).view(N, D, H, W, 1).contiguous()
    x_ref = x_b.view(N, 1, D, H, W).float()
    w_ref = w.to(torch.bfloat16).float().requires_grad_(True)
    ref = F.conv3d(x_ref, w_ref, None, 2, 3, 1)
    x8 = K.stem_expand(x_b)
    y, st = K.stem_fprop(x8, tuple(x_b.shape), w)
    torch.cuda.synchronize()
    assert_close(to_ncdhw_f32(y), ref.detach(), 6e-3, f"stem fprop {shape}")
    assert_close(st[1], (ref.detach().double() ** 2).sum(dim=(0, 2, 3, 4)), 1e-4, f"stem stats {shape}")
    dy_b = to_ndhwc_bf16(torch.randn_like(ref))
    ref.backward(to_ncdhw_f32(dy_b))
    gw = K.stem_wgrad(x8, dy_b, tuple(x_b.shape))
    torch.cuda.synchronize()
    assert_close(gw, w_ref.grad, 2e-4, f"stem wgrad {shape}")

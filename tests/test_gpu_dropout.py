"""Training-mode Dropout on the CUDA path (ADVICE r01 / VERDICT A6): pet_cnn.py:26-27,38-40 semantics.

* mask replay: the CUDA keep mask equals oracle/dropout.py's numpy restatement bit for bit (bf16 and fp32, ragged
  sizes), and the backward pass applies the SAME mask, scaled by 1/(1-p);
* statistics: keep rate within 4 sigma of 1-p, E[y] = E[x];
* stream: consecutive calls draw different masks, a re-seeded module repeats them, eval mode / p = 0 are the identity;
* the reference's shipped hparams that carry dropout (train_early_fusion.py:250 dropout_conv_p = 0.1166) train.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n", [8 * 1000, 8 * 1000 + 5, 3])
def test_mask_matches_oracle_and_backward_replays_it(cuda_dev, dtype, n):
    from multimodal_alzheimer_b200 import autograd as A
    from oracle.dropout import keep_mask
    p, seed = 0.3, 0x1234_5678_9ABC_DEF
    g = torch.Generator().manual_seed(n)
    x = (torch.rand(n, generator=g) + 0.5).to(dtype).to(cuda_dev).requires_grad_(True)
    counter = torch.full((1,), 41, dtype=torch.int64, device=cuda_dev)
    y = A.DropoutFn.apply(x, p, seed, counter)
    assert int(counter) == 42
    keep = torch.from_numpy(keep_mask(n, p, seed, 41))
    scale = float(np.float32(1.0 / (1.0 - p)))
    ref = torch.where(keep, x.detach().cpu().float() * scale, torch.zeros(n)).to(dtype)
    assert torch.equal(y.detach().cpu(), ref)
    dy = (torch.rand(n, generator=g) - 0.5).to(dtype).to(cuda_dev)
    y.backward(dy)
    gref = torch.where(keep, dy.cpu().float() * scale, torch.zeros(n)).to(dtype)
    assert torch.equal(x.grad.cpu(), gref)


def test_statistics_and_stream(cuda_dev):
    from multimodal_alzheimer_b200 import nn as bnn
    torch.manual_seed(15)
    d = bnn.Dropout(p=0.1166).to(cuda_dev).train()
    n = 1 << 22
    x = torch.ones(n, dtype=torch.bfloat16, device=cuda_dev)
    y1, y2 = d(x), d(x)
    k1, k2 = (y1 != 0).float(), (y2 != 0).float()
    sigma = (0.1166 * 0.8834 / n) ** 0.5
    assert abs(float(k1.mean()) - 0.8834) < 4 * sigma and abs(float(k2.mean()) - 0.8834) < 4 * sigma
    assert abs(float(y1.float().mean()) - 1.0) < 6 * sigma / 0.8834 + 4e-3      # bf16 rounding of 1/(1-p)
    assert float((k1 != k2).float().mean()) > 0.15                              # independent draws
    torch.manual_seed(15)
    d2 = bnn.Dropout(p=0.1166).to(cuda_dev).train()
    assert torch.equal(d2(x), y1)                                               # reproducible under the seed
    d.eval()
    assert d(x) is x
    assert bnn.Dropout(p=0.0).train()(x) is x


def test_reference_dropout_hparams_train(cuda_dev):
    """PET_MRI_EF with the reference's shipped dropout_conv_p (train_early_fusion.py:250) takes a training step:
    finite loss, gradients on every parameter, eval mode deterministic."""
    from tests._models import build_model, hp_pet, synthetic_batch
    from multimodal_alzheimer_b200.pkg.models.fusion_models.early_fusion import PET_MRI_EF
    from multimodal_alzheimer_b200.pkg.models.pet_models.pet_cnn import Small_PET_CNN
    torch.manual_seed(15)
    for cls, mods in ((PET_MRI_EF, ("mri", "pet1451")), (Small_PET_CNN, ("pet1451",))):
        hp = hp_pet(3, True)
        hp["dropout_conv_p"] = 0.1166
        hp["dropout_dense_p"] = 0.3
        model = cls(hp).to(cuda_dev).train()
        batch = {k: v.to(cuda_dev) for k, v in synthetic_batch(4, (32, 32, 32), 3, modalities=mods).items()}
        out = model.training_step(batch, 0)
        out["loss"].backward()
        assert torch.isfinite(out["loss"])
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
        model.eval()
        with torch.no_grad():
            a = model.validation_step(batch, 0)["outputs"]
            b = model.validation_step(batch, 0)["outputs"]
        assert torch.equal(a, b)

"""CPU tests pinning the oracle (no GPU): golden vectors generated from the unmodified reference
(tools/make_golden.py) and the facts the reference records about MedicalNet."""
import json
import os

import pytest
import torch

from oracle.losses import FocalLossOracle, focal_grad_closed_form
from oracle.medicalnet import generate_model
from oracle.normalization import (masked_std_mean_oracle, masked_zscore_oracle, pet_standardize_oracle,
                                  quantile_from_sorted, quantile_minmax_oracle)

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)["cases"]


def test_focal_and_ce_match_reference_golden():
    """oracle/losses.py vs the reference's own focalloss.py outputs (fp64; identical op sequence -> 1e-15)."""
    for c in _load("focal_loss.json"):
        z = torch.tensor(c["logits"], dtype=torch.float64, requires_grad=True)
        t = torch.tensor(c["target"])
        if c["kind"] == "focal":
            loss = FocalLossOracle(gamma=c["gamma"])(z, t)
        else:
            loss = torch.nn.CrossEntropyLoss(weight=torch.tensor(c["weight"], dtype=torch.float64))(z, t)
        loss.backward()
        assert abs(float(loss) - c["loss"]) <= 1e-15 * max(1, abs(c["loss"]))
        assert torch.allclose(z.grad, torch.tensor(c["grad"], dtype=torch.float64), rtol=1e-14, atol=1e-16)


def test_focal_gradient_is_the_detached_form():
    """pt is detached (focalloss.py:30): grad = (1-pt)^gamma (softmax - onehot)/N, not the textbook focal gradient."""
    for c in _load("focal_loss.json"):
        if c["kind"] != "focal":
            continue
        z = torch.tensor(c["logits"], dtype=torch.float64)
        t = torch.tensor(c["target"])
        g = focal_grad_closed_form(z, t, c["gamma"])
        assert torch.allclose(g, torch.tensor(c["grad"], dtype=torch.float64), rtol=1e-12, atol=1e-15)


def test_weighted_ce_normaliser_is_sum_of_target_weights():
    w = torch.tensor([0.4651162790697675, 0.6712473572938689, 0.8636363636363636], dtype=torch.float64)
    z = torch.randn(7, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    t = torch.tensor([0, 1, 2, 2, 1, 0, 0])
    ref = torch.nn.CrossEntropyLoss(weight=w)(z, t)
    nll = -torch.log_softmax(z, 1).gather(1, t[:, None]).squeeze(1)
    assert abs(float((w[t] * nll).sum() / w[t].sum()) - float(ref)) < 1e-15


def test_quantile_normalisation_matches_reference_golden():
    for c in _load("quantile.json"):
        shape = tuple(c["shape"])
        mri = torch.tensor(c["mri"], dtype=torch.float64).view(shape)
        mask = torch.tensor(c["mask"], dtype=torch.float64).view(shape)
        out, meta = quantile_minmax_oracle(mri, mask, c["q"])
        assert meta["n"] == c["n"] and meta["qmax"] == c["qmax"] and meta["qmin"] == c["qmin"]
        # (q = 0.5 makes Qmax == Qmin: the reference divides by zero and keeps the NaNs; so must the oracle)
        assert torch.allclose(out.flatten(), torch.tensor(c["out"], dtype=torch.float64), rtol=0, atol=0,
                              equal_nan=True)
        # the rank/lerp restatement used by the CUDA kernel reproduces torch.quantile bit for bit
        s = (mri * mask).flatten()
        s = s[s != 0].sort().values
        assert quantile_from_sorted(s, c["q"]) == c["qmax"]
        assert quantile_from_sorted(s, 1 - c["q"]) == c["qmin"]
        n, mean, std = masked_std_mean_oracle(mri, mask)
        assert n == c["n"] and mean == c["mean"] and std == c["std"]
        assert torch.equal(masked_zscore_oracle(mri, mask).flatten(), torch.tensor(c["zscore"], dtype=torch.float64))
        assert torch.equal(pet_standardize_oracle(mri, 0.5145, 0.5383).flatten(),
                           torch.tensor(c["pet"], dtype=torch.float64))


def test_one_minus_q_keeps_its_double_rounding():
    assert 1 - 0.98 == 0.020000000000000018 and 1 - 0.99 == 0.010000000000000009


def test_medicalnet_restatement_matches_reference_facts():
    """pkg/utils/outdated/inspect_model.py:100,105,284-285: 2048 channels, 91x109x91 -> 12x14x12, 159 tensors."""
    m = generate_model(50).eval()
    assert sum(1 for _ in m.parameters()) == 159
    with torch.no_grad():
        y = m(torch.zeros(1, 1, 91, 109, 91))
    assert tuple(y.shape) == (1, 2048, 12, 14, 12)
    m18 = generate_model(18)
    assert abs(sum(p.numel() for p in m18.parameters()) / 1e6 - 33.16) < 0.01
    with pytest.raises(ValueError):
        generate_model(42)

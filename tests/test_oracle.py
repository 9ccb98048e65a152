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


def _golden_models():
    with open(os.path.join(GOLD, "models.json")) as f:
        return json.load(f)


def _golden_model_cases():
    return _golden_models()["cases"]


@pytest.fixture
def recorded_threads():
    """The intra-op thread count tools/make_golden_models.py ran with (CPU conv kernels split sums by thread)."""
    before = torch.get_num_threads()
    torch.set_num_threads(_golden_models()["num_threads"])
    yield
    torch.set_num_threads(before)


@pytest.mark.parametrize("case_id", sorted(_golden_model_cases().keys()))
def test_oracle_models_reproduce_reference_classes(case_id, recorded_threads):
    """oracle/models.py vs tests/golden/models.json = one training step of the reference's OWN LightningModules
    (pkg/models/**, imported unmodified by tools/reference_harness.py; stage-N models built through their
    load_from_checkpoint / truncation / freezing code) on the seeded inputs and weights of tests/_models.py::CASES.
    Same fp32 CPU arithmetic on both sides: bit-identical when generated.  Asserted: logits rel-L2 <= 1e-5, loss
    <= 1e-6, running statistics <= 1e-5, every gradient's fingerprint (norm and two unit-vector projections) within
    1e-4 of its norm + 20x the re-association noise the generator measured on the REFERENCE for that tensor by
    re-running it with 1 and 3 threads (conv biases in front of a BatchNorm have an exactly-zero gradient whose
    computed value is nothing but that noise).  state_dict keys/shapes, the set of parameters that receive a
    gradient and the requires_grad flags must be identical."""
    from tests._models import build_oracle, fingerprint, oracle_step, synthetic_batch
    rec = _golden_model_cases()[case_id]
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in rec["kw"].items()}
    oracle = build_oracle(rec["kind"], **kw)
    assert {k: list(v.shape) for k, v in oracle.state_dict().items()} == rec["state_dict"]
    assert list(oracle.state_dict().keys()) == list(rec["state_dict"].keys())
    assert {n: bool(p.requires_grad) for n, p in oracle.named_parameters()} == rec["requires_grad"]
    batch = synthetic_batch(rec["batch"], rec["shape"], kw.get("n_classes", 3), modalities=tuple(rec["modalities"]))
    out = oracle_step(oracle, batch)
    ref_logits = torch.tensor(rec["outputs"], dtype=torch.float64)
    assert out["outputs"].dtype == torch.float64 and out["outputs"].shape == ref_logits.shape
    assert float((out["outputs"].detach() - ref_logits).norm() / ref_logits.norm().clamp_min(1e-300)) <= 1e-5
    assert abs(float(out["loss"].detach()) - rec["loss"]) <= 1e-6
    got = dict(oracle.named_parameters())
    assert {n for n, p in got.items() if p.grad is not None} == set(rec["grads"])
    for n, fp in rec["grads"].items():
        f = fingerprint(n, got[n].grad)
        assert max(abs(a - b) for a, b in zip(f, fp)) <= 1e-4 * fp[0] + 20 * rec["grad_noise"][n], (n, f, fp)
    bufs = dict(oracle.named_buffers())
    for n, fp in rec["running"].items():
        f = fingerprint(n, bufs[n])
        assert max(abs(a - b) for a, b in zip(f, fp)) <= 1e-5 * max(fp[0], 1e-30), (n, f, fp)


@pytest.mark.parametrize("traj_id", sorted(_golden_models()["trajectories"].keys()))
def test_oracle_training_trajectory_matches_reference(traj_id, recorded_threads):
    """Four optimisation steps (reference `training_step` -> backward -> the Adam instance its own
    `configure_optimizers()` returned, a fresh synthetic batch per step; tools/make_golden_models.py) replayed with
    the oracle and torch.optim.Adam over the recorded parameter groups: loss of every step <= 1e-5, logits rel-L2
    <= 1e-4, parameters and running statistics after the last step within 1e-4 of their norm."""
    from tests._models import build_oracle, fingerprint, trajectory_batches
    rec = _golden_models()["trajectories"][traj_id]
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in rec["kw"].items()}
    oracle = build_oracle(rec["kind"], **kw)
    oracle.train()
    params = dict(oracle.named_parameters())
    opt = torch.optim.Adam([{"params": [params[n]], "lr": lr, "weight_decay": wd} for n, lr, wd in rec["optimizer"]],
                           betas=tuple(rec["adam"]["betas"]), eps=rec["adam"]["eps"])
    for k, batch in enumerate(trajectory_batches(traj_id)):
        out = oracle.general_step(dict(batch), k, "train")
        opt.zero_grad()
        out["loss"].backward()
        opt.step()
        assert abs(float(out["loss"].detach()) - rec["losses"][k]) <= 1e-5, (k, float(out["loss"]), rec["losses"][k])
        want = torch.tensor(rec["logits"][k], dtype=torch.float64)
        assert float((out["outputs"].detach() - want).norm() / want.norm().clamp_min(1e-300)) <= 1e-4, k
    for n, fp in rec["final_params"].items():
        f = fingerprint(n, params[n])
        assert max(abs(a - b) for a, b in zip(f, fp)) <= 1e-4 * max(fp[0], 1e-30), (n, f, fp)
    bufs = dict(oracle.named_buffers())
    for n, fp in rec["final_running"].items():
        f = fingerprint(n, bufs[n])
        assert max(abs(a - b) for a, b in zip(f, fp)) <= 1e-4 * max(fp[0], 1e-30), (n, f, fp)


def test_metric_restatement_agrees_with_sklearn():
    """oracle/metrics.py (torchmetrics 0.10.2 reductions restated; torchmetrics itself is absent: parity unpinned)
    vs scikit-learn on random 2- and 3-class problems in which every class occurs: macro / per-class F1 and MCC
    agree to fp32 rounding; a class absent from predictions AND targets is excluded from the macro mean
    (torchmetrics) where sklearn would count it as 0."""
    from sklearn.metrics import f1_score, matthews_corrcoef
    from oracle.metrics import confusion_matrix, f1_from_confmat, mcc_from_confmat
    g = torch.Generator().manual_seed(15)
    for C in (2, 3):
        for n in (12, 57, 300):
            logits = torch.randn((n, C), generator=g, dtype=torch.float64)
            labels = torch.randint(0, C, (n,), generator=g)
            labels[:C] = torch.arange(C)
            logits[torch.arange(C), torch.arange(C)] += 10            # every class predicted at least once
            cm = confusion_matrix(logits, labels, C)
            macro, per = f1_from_confmat(cm)
            preds = logits.argmax(1).numpy()
            assert abs(float(macro) - f1_score(labels.numpy(), preds, average="macro")) <= 1e-6
            assert torch.allclose(per.double(), torch.tensor(f1_score(labels.numpy(), preds, average=None)), atol=1e-6)
            assert abs(float(mcc_from_confmat(cm)) - matthews_corrcoef(labels.numpy(), preds)) <= 1e-6
    cm = torch.tensor([[5, 1, 0], [2, 4, 0], [0, 0, 0]])
    macro, per = f1_from_confmat(cm)
    assert per[2] == 0 and abs(float(macro) - float(per[:2].mean())) < 1e-7
    assert float(mcc_from_confmat(torch.tensor([[4, 0], [3, 0]]))) == 0.0   # zero denominator -> 0


def test_dropout_oracle_philox_known_answers():
    """Philox4x32-10 of oracle/dropout.py against the Random123 known-answer vectors (kat_vectors: philox4x32 10)."""
    from oracle.dropout import keep_mask, philox4x32_10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(v) for v in philox4x32_10(*ctr, *key)) == want
    m = keep_mask(200000, 0.25, seed=99, offset=3)
    assert abs(m.mean() - 0.75) < 4 * (0.25 * 0.75 / 200000) ** 0.5
    assert not (keep_mask(4096, 0.25, 99, 3) == keep_mask(4096, 0.25, 99, 4)).all()

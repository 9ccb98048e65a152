"""Golden outputs of the reference's OWN model classes (imported unmodified through tools/reference_harness.py) for the
module-level parity cases of tests/_models.py::CASES -> tests/golden/models.json.

Run in the build container (needs /root/reference).  For every case that is a reference configuration:
  1. build the oracle (oracle/models.py) with the case's seed,
  2. build the reference class the way its training script does - stage-N models through
     `load_from_checkpoint(path)` of registered stage-(N-1) checkpoints, i.e. through the reference's own
     truncation / freezing code,
  3. load the oracle's state_dict into the reference model with strict=True (key / shape identity with the real class),
  4. run the reference's `general_step(batch, 0, 'train')` + backward on the case's synthetic batch,
  5. record logits, loss, per-parameter gradient fingerprints (L2 norm, sum, dot with a fixed pseudo-random
     vector), requires_grad flags, running statistics fingerprints and the optimizer's parameter groups
     (parameter name, lr, weight decay) from `configure_optimizers()`;
  6. assert that the oracle reproduces all of it (the same check tests/test_oracle.py repeats from the fixture).

Also recorded: the same for the frozen-encoder variants (`lr_pretrained` None), which exercise the freezing code.
"""
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import reference_harness as H  # noqa: E402
from tests import _models as M  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "models.json")
THREADS = 16          # intra-op threads of the recorded run (tests pin the same count: bit-reproducible on this image)


fingerprint = M.fingerprint


def build_reference(R, kind, kw, oracle):
    """The reference model of `kind`, built as the reference's train_*.py scripts build it, carrying the oracle's
    weights.  Stage-1 weights reach a stage-2 model the way they do in the reference: through a checkpoint."""
    depth, nc = kw.get("depth", 10), kw.get("n_classes", 3)
    hp_a = M.hp_anat(depth, nc, kw.get("fl_gamma"), kw.get("bn_begin", False), kw.get("bn_dense", False),
                     kw.get("linear_out", ()))
    hp_p = M.hp_pet(nc, kw.get("pet_batchnorm", False), kw.get("conv_out", (8, 16, 32, 64)),
                    kw.get("filter_size", (5, 5, 3, 3)))
    hp_f = M.hp_fusion(nc, kw.get("fl_gamma", 1), kw.get("simple_dim_red", False))
    if kw.get("frozen"):
        hp_a["lr_pretrained"] = None
        hp_f["lr_pretrained"] = None
    H.CHECKPOINTS.clear()

    def stage1():
        H.register_checkpoint("ckpt/mri", R["Anat_CNN"](hp_a))
        H.register_checkpoint("ckpt/pet", R["Small_PET_CNN"](hp_p))

    if kind == "anat":
        ref = R["Anat_CNN"](hp_a)
    elif kind == "pet_resnet":
        ref = R["PET_CNN_ResNet"](dict(hp_a, gpu_id="0"))
    elif kind == "small_pet":
        ref = R["Small_PET_CNN"](hp_p)
    elif kind == "early_fusion":
        ref = R["PET_MRI_EF"](hp_p)
    elif kind == "fmf":
        ref = R["PET_MRI_FMF"](M.hp_fmf(nc, kw.get("fusion_mode", "maxout"), kw.get("pet_batchnorm", True),
                                        kw.get("batchnorm_fusion", True), kw.get("filter_size_fusion", 3),
                                        kw.get("n_out_fusion", 64)))
    elif kind == "anat_pet":
        stage1()
        ref = R["Anat_PET_CNN"](dict(hp_f, path_pet="ckpt/pet", path_mri="ckpt/mri"))
    elif kind == "mri_tab":
        stage1()
        ref = R["Tabular_MRT_Model"](dict(hp_f, path_mri="ckpt/mri"))
    elif kind == "pet_tab":
        stage1()
        ref = R["PET_TABULAR_CNN"](dict(hp_f, path_pet="ckpt/pet"))
    elif kind == "all":
        stage1()
        hp2 = M.hp_fusion(nc, kw.get("fl_gamma", 1), False)
        hp2["lr_pretrained"] = hp_f["lr_pretrained"]          # stage 2 trained with the same freeze choice
        H.register_checkpoint("ckpt/anat_pet", R["Anat_PET_CNN"](dict(hp2, path_pet="ckpt/pet", path_mri="ckpt/mri")))
        H.register_checkpoint("ckpt/anat_tab", R["Tabular_MRT_Model"](dict(hp2, path_mri="ckpt/mri")))
        H.register_checkpoint("ckpt/pet_tab", R["PET_TABULAR_CNN"](dict(hp2, path_pet="ckpt/pet")))
        ref = R["All_Modalities_Fusion"](dict(hp_f, path_anat_pet="ckpt/anat_pet", path_anat_tab="ckpt/anat_tab",
                                              path_pet_tab="ckpt/pet_tab", path_pet="ckpt/pet", path_anat="ckpt/mri"))
    else:
        return None
    res = ref.load_state_dict(copy.deepcopy(oracle.state_dict()), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return ref


def optimizer_groups(model, opt):
    if isinstance(opt, dict):
        opt = opt["optimizer"]
    names = {}
    for n, p in model.named_parameters(remove_duplicate=False):
        names.setdefault(id(p), n)
    groups = []
    for g in opt.param_groups:
        for p in g["params"]:
            groups.append([names.get(id(p), "?"), g["lr"], g["weight_decay"]])
    return groups


NOISE_THREADS = (1, 3)


def record(ref, batch, with_optimizer=True):
    ref.train()
    # Re-association noise of the reference itself: the same step with other intra-op thread counts (the CPU conv /
    # reduction kernels split their sums by thread).  Gradients that are mathematically zero (a conv bias in front
    # of a BatchNorm) are pure noise of this kind; the tests allow 20x the deviation measured here per tensor.
    saved = {n: b.detach().clone() for n, b in ref.named_buffers()}
    alt = []
    for t in NOISE_THREADS:
        torch.set_num_threads(t)
        ref.zero_grad(set_to_none=True)
        ref.general_step(dict(batch), 0, "train")["loss"].backward()
        alt.append({n: fingerprint(n, p.grad) for n, p in ref.named_parameters() if p.grad is not None})
        with torch.no_grad():
            for n, b in ref.named_buffers():
                b.copy_(saved[n])
    torch.set_num_threads(THREADS)
    ref.zero_grad(set_to_none=True)
    out = ref.general_step(dict(batch), 0, "train")
    out["loss"].backward()
    rec = {"outputs": out["outputs"].detach().tolist(), "loss": float(out["loss"].detach()),
           "state_dict": {k: list(v.shape) for k, v in ref.state_dict().items()},
           "grads": {}, "grad_noise": {}, "requires_grad": {}, "running": {}}
    for n, p in ref.named_parameters():
        rec["requires_grad"][n] = bool(p.requires_grad)
        if p.grad is not None:
            rec["grads"][n] = fp = fingerprint(n, p.grad)
            rec["grad_noise"][n] = max(abs(a - b) for other in alt for a, b in zip(other[n], fp))
    for n, b in ref.named_buffers():
        if n.endswith("running_mean") or n.endswith("running_var"):
            rec["running"][n] = fingerprint(n, b)
    if with_optimizer:
        try:
            rec["optimizer"] = optimizer_groups(ref, ref.configure_optimizers())
        except AttributeError as e:
            # all_modalities_fusion.py:109-122 iterates `model_tabular.named_parameters()` when lr_pretrained is set;
            # a TabPFNClassifier is an sklearn estimator, not an nn.Module (the stand-in is faithful to that)
            rec["optimizer_error"] = repr(e)
        rec["requires_grad_after_configure"] = {n: bool(p.requires_grad) for n, p in ref.named_parameters()}
    return rec


def check_oracle(case_id, rec, oracle, batch):
    """The oracle must reproduce the reference record (same arithmetic, same machine: tight)."""
    out = M.oracle_step(oracle, batch)
    lo = torch.tensor(rec["outputs"], dtype=torch.float64)
    d = float((out["outputs"].detach() - lo).norm() / lo.norm().clamp_min(1e-300))
    dl = abs(float(out["loss"].detach()) - rec["loss"])
    worst = 0.0
    got = {n: p for n, p in oracle.named_parameters()}
    have = set(n for n, p in got.items() if p.grad is not None)
    assert have == set(rec["grads"]), (case_id, sorted(have ^ set(rec["grads"]))[:10])
    for n, fp in rec["grads"].items():
        f = fingerprint(n, got[n].grad)
        scale = max(fp[0], 1e-30)
        worst = max(worst, max(abs(a - b) for a, b in zip(f, fp)) / scale)
    assert {k: list(v.shape) for k, v in oracle.state_dict().items()} == rec["state_dict"], case_id
    print(f"{case_id:22s} logits rel {d:.2e}  loss abs {dl:.2e}  worst grad fingerprint {worst:.2e}  "
          f"({len(rec['grads'])} gradients, {len(rec['state_dict'])} state_dict entries)")
    assert d <= 1e-5 and dl <= 1e-6 and worst <= 1e-4, case_id
    return d, dl, worst


def main():
    R = H.reference_classes()
    torch.set_num_threads(THREADS)
    cases = {}
    extra = [("anat_pet", dict(depth=10, frozen=True), 2, (32, 32, 32), ("mri", "pet1451")),
             ("anat", dict(depth=10, frozen=True), 2, (32, 32, 32), ("mri",)),
             ("all", dict(depth=10, frozen=True), 2, (32, 32, 32), ("mri", "pet1451", "tabular")),
             ("pet_tab", dict(n_classes=2), 3, (32, 32, 32), ("pet1451", "tabular")),
             ("anat_pet", dict(depth=10, n_classes=2, fl_gamma=None), 3, (32, 32, 32), ("mri", "pet1451"))]
    todo = list(zip(M.CASE_IDS, M.CASES)) + [(f"x-{c[0]}-{i}", c) for i, c in enumerate(extra)]
    for case_id, (kind, kw, B, shape, mods) in todo:
        oracle = M.build_oracle(kind, **kw)
        ref = build_reference(R, kind, kw, oracle)
        if ref is None:
            print(f"{case_id:22s} not a reference configuration - no record")
            continue
        batch = M.synthetic_batch(B, shape, kw.get("n_classes", 3), modalities=mods)
        rec = record(ref, batch)
        rec.update(kind=kind, kw={k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}, batch=B,
                   shape=list(shape), modalities=list(mods), reference_class=type(ref).__module__ + "." + type(ref).__name__)
        check_oracle(case_id, rec, oracle, batch)
        cases[case_id] = rec
    trajectories = {}
    for traj_id, (kind, kw, B, shape, mods, steps) in M.TRAJECTORIES.items():
        torch.set_num_threads(THREADS)
        oracle = M.build_oracle(kind, **kw)
        ref = build_reference(R, kind, kw, oracle)
        ref.train()
        opt = ref.configure_optimizers()
        opt = opt["optimizer"] if isinstance(opt, dict) else opt
        groups = optimizer_groups(ref, opt)
        losses, logits = [], []
        for k, batch in enumerate(M.trajectory_batches(traj_id)):
            out = ref.training_step(dict(batch), k)                      # base_model.py:60-66
            opt.zero_grad()
            out["loss"].backward()
            opt.step()
            losses.append(float(out["loss"].detach()))
            logits.append(out["outputs"].detach().tolist())
        trajectories[traj_id] = {
            "kind": kind, "kw": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}, "batch": B,
            "shape": list(shape), "modalities": list(mods), "steps": steps, "losses": losses, "logits": logits,
            "optimizer": groups, "adam": {"betas": list(opt.defaults["betas"]), "eps": opt.defaults["eps"]},
            "final_params": {n: fingerprint(n, p) for n, p in ref.named_parameters()},
            "final_running": {n: fingerprint(n, b) for n, b in ref.named_buffers()
                              if n.endswith("running_mean") or n.endswith("running_var")}}
        print(f"{traj_id:22s} losses {[round(x, 5) for x in losses]}")
    if "tabular" in "".join(sum((list(c[4]) for _, c in todo), [])):
        assert H.ENSEMBLE_SEEN and all(e == 4 for e in H.ENSEMBLE_SEEN)
    with open(OUT, "w") as f:
        json.dump({"source": "reference pkg/models/** imported unmodified via tools/reference_harness.py; MedicalNet "
                             "ResNet = oracle/medicalnet.py (third-party clone, absent from the reference tree)",
                   "torch": torch.__version__, "num_threads": THREADS, "noise_threads": list(NOISE_THREADS),
                   "cases": cases, "trajectories": trajectories}, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()

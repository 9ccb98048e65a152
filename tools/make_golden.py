"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference (run in the build container,
where /root/reference exists; the fixtures travel to the GPU box, the reference does not).

  focal_loss.json : outputs + input gradients of the reference's own pkg/loss_functions/focalloss.py (imported
                    as-is) and of nn.CrossEntropyLoss(weight) on seeded fp64 logits.
  quantile.json   : torch.quantile / torchvision Normalize executed exactly as pkg/utils/dataloader.py:213-215,
                    245-270 does (the file itself needs nibabel/seaborn and cannot be imported), on seeded volumes.
"""
import importlib.util
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def load_reference_focal():
    spec = importlib.util.spec_from_file_location("ref_focalloss", os.path.join(REF, "pkg/loss_functions/focalloss.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.FocalLoss


def golden_focal():
    FocalLoss = load_reference_focal()
    cases = []
    g = torch.Generator().manual_seed(15)
    cw = [0.4651162790697675, 0.6712473572938689, 0.8636363636363636]  # pkg/inference/test_tab.py:36-40
    for C in (2, 3):
        for B in (1, 4, 9):
            logits = (torch.randn((B, C), generator=g, dtype=torch.float64) * 3)
            target = torch.randint(0, C, (B,), generator=g)
            for gamma in (1, 2, 5):
                z = logits.clone().requires_grad_(True)
                loss = FocalLoss(gamma=gamma)(z, target)
                loss.backward()
                cases.append({"kind": "focal", "gamma": gamma, "logits": logits.tolist(), "target": target.tolist(),
                              "loss": float(loss), "grad": z.grad.tolist()})
            z = logits.clone().requires_grad_(True)
            w = torch.tensor(cw[:C], dtype=torch.float64)
            loss = torch.nn.CrossEntropyLoss(weight=w)(z, target)
            loss.backward()
            cases.append({"kind": "ce", "weight": cw[:C], "logits": logits.tolist(), "target": target.tolist(),
                          "loss": float(loss), "grad": z.grad.tolist()})
    with open(os.path.join(OUT, "focal_loss.json"), "w") as f:
        json.dump({"source": "pkg/loss_functions/focalloss.py (reference, imported unmodified) / torch CE",
                   "cases": cases}, f)


def golden_quantile():
    from torchvision.transforms import Normalize
    g = torch.Generator().manual_seed(15)
    cases = []
    for shape, q in (((6, 7, 5), 0.98), ((8, 8, 8), 0.95), ((5, 9, 4), 0.99), ((4, 4, 4), 1.0), ((3, 5, 7), 0.5)):
        mri = (400 * torch.randn(shape, generator=g).abs() + 50 * torch.rand(shape, generator=g)).float().double()
        mask = (torch.rand(shape, generator=g) < 0.6).double()
        mri[0, 0, 0] = 0.0
        # --- verbatim sequence of dataloader.py:245-270 ---
        data_masked_mri = mri * mask
        data_masked_mri = data_masked_mri.reshape(-1)
        data_masked_mri = data_masked_mri[data_masked_mri.nonzero()]
        quant_max = torch.quantile(data_masked_mri, q, interpolation='linear')
        quant_min = torch.quantile(data_masked_mri, 1 - q, interpolation='linear')
        out = (mri - quant_min) / (quant_max - quant_min)
        out[out > 1] = 1
        out[out < 0] = 0
        out *= mask
        # --- dataloader.py:252-260 ---
        std_m, mean_m = torch.std_mean(data_masked_mri)
        z = Normalize(mean=mean_m, std=std_m)(mri.clone()) * mask
        # --- dataloader.py:213-215 ---
        pet = Normalize(mean=0.5145, std=0.5383)(mri.clone())
        cases.append({"shape": list(shape), "q": q, "mri": mri.flatten().tolist(), "mask": mask.flatten().tolist(),
                      "n": int(data_masked_mri.numel()), "qmax": float(quant_max), "qmin": float(quant_min),
                      "out": out.flatten().tolist(), "std": float(std_m), "mean": float(mean_m),
                      "zscore": z.flatten().tolist(), "pet": pet.flatten().tolist()})
    with open(os.path.join(OUT, "quantile.json"), "w") as f:
        json.dump({"source": "torch.quantile / torchvision Normalize run as pkg/utils/dataloader.py:213-215,245-270",
                   "cases": cases}, f)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    golden_focal()
    golden_quantile()
    print("golden fixtures written to", OUT)

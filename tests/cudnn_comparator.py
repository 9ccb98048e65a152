"""Informational comparator (SURVEY.md 8(d) 'Secondary comparator'): the reference's own GPU path modernised - the
oracle restatement of the two-ResNet-18 PET-MRI fusion model executed by stock PyTorch + cuDNN on the same B200, in fp32
(TF32 convolutions, PyTorch's default) and under bf16 autocast with channels_last_3d - for the same training step as
bench.py minus the input normalisation (forward + fp64 focal loss + backward + Adam).  Not a parity oracle and not part
of the product path; run by hand:  python tests/cudnn_comparator.py [pairs]  -> one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bench import build_oracle
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    out = {"workload": "two-ResNet-18 PET-MRI fusion, 128^3, fwd + focal loss + bwd + Adam (no input normalisation)",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    g = torch.Generator(device=dev).manual_seed(15)
    for mode in ("bf16_autocast_channels_last_3d", "fp32_tf32"):
        B = pairs
        while B >= 2:
            try:
                model = build_oracle("pet_mri_fusion_r18", 18).to(dev).train()
                if mode.startswith("bf16"):
                    model = model.to(memory_format=torch.channels_last_3d)
                opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4,
                                       fused=True)
                batch = {"mri": torch.rand((B, 128, 128, 128), generator=g, device=dev),
                         "pet1451": torch.randn((B, 128, 128, 128), generator=g, device=dev),
                         "label": torch.randint(0, 3, (B,), generator=g, device=dev)}

                def step():
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode.startswith("bf16")):
                        res = model.general_step(batch, 0, "train")
                    opt.zero_grad(set_to_none=True)
                    res["loss"].backward()
                    opt.step()

                t0 = time.perf_counter()
                step()                                  # cuDNN autotuning happens here
                torch.cuda.synchronize()
                tune_s = time.perf_counter() - t0
                step()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                n = 3
                for _ in range(n):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                out[mode] = {"pairs": B, "ms_per_step": ms, "volumes_per_s": 2 * B / (ms / 1e3), "first_step_s": tune_s,
                             "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
                break
            except torch.cuda.OutOfMemoryError:
                B //= 2
            finally:
                model = opt = batch = None
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()

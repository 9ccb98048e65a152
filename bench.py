#!/usr/bin/env python
"""Benchmark of the B200-native training hot path (contract: see README / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores
    python bench.py --workload mri_r50_160 ...               # the other BASELINE.json configurations

Default workload (BASELINE.json metric "MRI+PET volumes/sec train (ResNet-18 3D, 128^3)"; configs[2]): the PET-MRI
two-branch ResNet-18 fusion model, focal loss gamma=1, global batch 32 (MRI, PET) pairs of 1x128^3 volumes,
random-init weights, synthetic data.  A step = per-scan quantile min-max normalisation of the MRI batch + PET
standardisation + forward + fp64 loss + backward + gradient all-reduce + Adam step.  value = volumes per second over
all ranks, device-timed with CUDA events, max over ranks.  Every rank count trains the SAME global batch (samples are
generated per global index), so `parity.first_step_loss` / `parity.first_step_logits_checksum` must agree across N.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "volumes/s"
PORT_PINNING = ("oracle port; its model classes reproduce one training step of the reference's own LightningModules bit "
                "for bit (tests/golden/models.json, tools/make_golden_models.py) - the reference itself needs "
                "pytorch_lightning / MedicalNet / nibabel and cannot run on this box")


def parse_args():
    from multimodal_alzheimer_b200.workloads import WORKLOADS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pet_mri_fusion_r18", choices=sorted(WORKLOADS))
    ap.add_argument("--global-batch", type=int, default=None)
    ap.add_argument("--volume", type=int, nargs="+", default=None, help="override the volume: one edge or D H W")
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-pairs", type=int, default=1, help="samples per step of the CPU legs (bounded sample)")
    ap.add_argument("--shape-profile", default=None, help="write the per-conv-shape timing table to this JSON file")
    ap.add_argument("--no-graph", action="store_true", help="time the eager launch path instead of the CUDA graph")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- shared config
def resolve(args, world):
    """(workload record, depth, volume, global batch, scaling) after the command-line overrides."""
    from multimodal_alzheimer_b200.workloads import WORKLOADS
    w = WORKLOADS[args.workload]
    depth = args.depth or w["depth"]
    if args.volume is None:
        volume = tuple(w["volume"])
    else:
        volume = tuple(args.volume * 3) if len(args.volume) == 1 else tuple(args.volume)
    if args.global_batch:
        gb, scaling = args.global_batch, "strong"
    elif w["global_batch"] is not None:
        gb, scaling = w["global_batch"], "strong"
    else:
        gb, scaling = w["per_gpu_batch"] * world, "weak"
    return w, depth, volume, gb, scaling


def metric_name(args, w, depth, volume):
    if args.workload == "pet_mri_fusion_r18" and depth == 18 and volume == (128, 128, 128):
        return "MRI+PET volumes/sec train (ResNet-18 3D, 128^3)"      # BASELINE.json's metric
    mods = "+".join({"mri": "MRI", "pet": "PET", "tab": "tabular"}[m] for m in w["modalities"])
    return f"{mods} volumes/sec train ({args.workload}, ResNet-{depth} 3D, {'x'.join(map(str, volume))})"


def workload_config(args, w, depth, volume, world, global_batch, step_global_batch=None):
    cfg = {"workload": f"{w['title']}; per-scan quantile(0.98) MRI normalisation"
                       + (" + PET standardisation" if "pet" in w["modalities"] else "") + ", fwd+bwd+allreduce+Adam",
           "name": args.workload, "resnet_depth": depth, "global_batch": global_batch, "volume": list(volume),
           "parallelism": f"dp{world}",
           "l2": "inputs and activations of a step are far larger than the 126 MB L2 (>= 18 MB of raw input per sample)",
           "baseline_config": w["config"]}
    if step_global_batch is not None and step_global_batch != global_batch:
        cfg["samples_per_timed_step"] = step_global_batch
    return cfg


# ----------------------------------------------------------------------------------------------- CPU reference arm
def oracle_namespace():
    """The CPU oracle's classes in the shape workloads.build_model expects (bench.py's CPU legs only)."""
    import types

    import torch

    import oracle.models as O

    class ResNet_PET_Trunk(torch.nn.Module):  # PET_CNN_ResNet encoder + Linear(512,64)+ReLU (two-ResNet fusion)
        def __init__(self, encm):
            super().__init__()
            self.encoder = encm
            self.encoder.model.conv_seg = self.encoder.model.conv_seg[:2]
            self.relu = torch.nn.ReLU()
            self.reduce_dim_pet = torch.nn.Sequential(torch.nn.Linear(512, 64), self.relu)

        def forward(self, x):
            o = self.encoder(x)
            return self.reduce_dim_pet(o.view(o.shape[0], -1))

    return types.SimpleNamespace(Anat_CNN=O.Anat_CNN, PET_CNN_ResNet=O.PET_CNN_ResNet, Small_PET_CNN=O.Small_PET_CNN,
                                 Anat_PET_CNN=O.Anat_PET_CNN, ResNet_PET_Trunk=ResNet_PET_Trunk,
                                 Tabular_MRT_Model=O.Tabular_MRT_Model, PET_TABULAR_CNN=O.PET_TABULAR_CNN,
                                 All_Modalities_Fusion=O.All_Modalities_Fusion, tab_key="tabular_features")


def cpu_reference_step_fn(args, n_samples):
    """The reference's CPU path for the same step: dataloader.py normalisation (fp64 torch.quantile / Normalize),
    fp32 torch.nn forward, fp64 loss, backward, Adam with one group per tensor.  Runs on the first `n_samples` global
    samples of the workload.  Returns (callable, volumes per call, threads)."""
    import torch

    from multimodal_alzheimer_b200 import workloads as W
    from oracle.normalization import pet_standardize_oracle, quantile_minmax_oracle
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    w, depth, volume, _, _ = resolve(args, 1)
    ns = oracle_namespace()
    model = W.build_model(ns, args.workload, depth=depth)
    opt = torch.optim.Adam(W.per_tensor_groups(model), weight_decay=1e-4)
    data = W.synth_batch(0, n_samples, volume, w["modalities"])

    def step():
        batch = {"label": data["label"]}
        if "mri_raw" in data:
            batch["mri"] = torch.stack([quantile_minmax_oracle(data["mri_raw"][i].double(), data["mask"][i].double(),
                                                               0.98)[0] for i in range(n_samples)])
        if "pet_raw" in data:
            batch["pet1451"] = pet_standardize_oracle(data["pet_raw"].double(), W.PET_MEAN, W.PET_STD)
        if "tabular" in data:
            batch[ns.tab_key] = data["tabular"]
        out = model.general_step(batch, 0, "train")
        opt.zero_grad(set_to_none=True)
        out["loss"].backward()
        opt.step()
        return float(out["loss"].detach())

    return step, n_samples * W.volumes_per_sample(args.workload), threads


def run_reference(args, rank):
    if rank != 0:
        return
    w, depth, volume, gb, scaling = resolve(args, args.gpus)
    step, vols, threads = cpu_reference_step_fn(args, args.cpu_sample_pairs)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = vols * args.steps / dt
    sample = (f"each step = the first {args.cpu_sample_pairs} sample(s) of the workload's global batch ({vols} volumes of "
              f"{'x'.join(map(str, volume))}): fp64 quantile normalisation + fp32 fwd + fp64 loss + bwd + per-tensor-group Adam")
    line = {
        "impl": "reference", "metric": metric_name(args, w, depth, volume), "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, w, depth, volume, args.gpus, gb, step_global_batch=args.cpu_sample_pairs),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "pinning": PORT_PINNING},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}
            while not self._stop_evt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.05)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from multimodal_alzheimer_b200 import _lib, data_parallel as dp
    from multimodal_alzheimer_b200 import kernels as K
    from multimodal_alzheimer_b200 import workloads as W

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback (use --impl reference)")
    _lib.load()
    rank, local_rank, world = dp.init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))  # never the legacy default stream (CUDA-graph capture)
    w, depth, volume, global_batch, scaling = resolve(args, world)
    lo, hi = dp.shard_bounds(global_batch, rank, world)
    n_local = hi - lo
    vps = W.volumes_per_sample(args.workload)
    vols_per_step = global_batch * vps

    model = W.build_model(W.product_namespace(), args.workload, depth=depth)
    model.to(dev).train()
    opt = model.configure_optimizers()          # the reference's per-tensor groups -> csrc/optimizer.cu
    opt = opt["optimizer"] if isinstance(opt, dict) else opt
    params = [p for p in model.parameters() if p.requires_grad]
    buckets = dp.make_gradient_buckets(params, overlap=True)   # one backward pass per step: exchange during backward

    host_data = W.synth_batch(lo, n_local, volume, w["modalities"])   # this rank's shard of the fixed global batch
    data = {k: v.to(dev) for k, v in host_data.items()}
    probe = torch.linspace(-1.0, 1.0, 3 * global_batch, dtype=torch.float64, device=dev).view(global_batch, 3)[lo:hi]

    def step(d):
        out = model.general_step(W.normalized_batch_gpu(d), 0, "train")
        out["loss"].backward()
        buckets.all_reduce()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return out["loss"], out["outputs"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def global_sum(t):
        t = t.detach().clone()
        if world > 1:
            dist.all_reduce(t)
        return float(t)

    # ---- first step: parity evidence (same weights, same global batch for every N) ------------------------------
    loss, logits = step(data)
    parity = {"first_step_loss": float(loss.detach()),
              "first_step_logits_checksum": global_sum((logits.detach() * probe).sum()),
              "first_step_logits_abs_sum": global_sum(logits.detach().abs().sum()),
              "note": "weights seed 15, global samples 0..B-1 generated per index: identical for every --gpus N"}

    # ---- eager profiled pass: per-kernel CUDA events for the roofline, launch count, eager step time --------
    for _ in range(max(args.warmup - 1, 0)):
        step(data)
    barrier()
    prof_steps = max(1, min(args.steps, 3))
    from multimodal_alzheimer_b200 import branches
    branches_enabled, branches._ENABLED = branches._ENABLED, False  # serialise the branches: clean per-kernel timing
    K.PROFILE.enable()
    if args.shape_profile:
        _lib.CALL_TIMING = {}
    launches0 = _lib.launch_count()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    host_t0 = time.perf_counter()
    for _ in range(prof_steps):
        loss, _ = step(data)
    host_issue_ms = 1e3 * (time.perf_counter() - host_t0) / prof_steps  # CPU time to enqueue one eager step
    p1.record()
    barrier()
    launches_per_step = (_lib.launch_count() - launches0) // prof_steps
    eager_ms = p0.elapsed_time(p1) / prof_steps
    prof = K.PROFILE.disable_and_collect()
    branches._ENABLED = branches_enabled
    for d in prof.values():  # normalise to args.steps so that the per-step divisions below hold
        d["flops"] = d["flops"] * args.steps / prof_steps
        d["executed_flops"] = d.get("executed_flops", d["flops"]) * args.steps / prof_steps
        d["ms"] = d["ms"] * args.steps / prof_steps
        d["n"] = d["n"] * args.steps / prof_steps
    call_ms = _lib.collect_call_timing() if args.shape_profile else {}
    if args.shape_profile and rank == 0:
        table = {k: {"ms_per_step": v["ms"] / prof_steps, "n_per_step": v["n"] / prof_steps,
                     "tflops": v["flops"] / (v["ms"] / 1e3) / 1e12 if v["ms"] > 0 else None,
                     "executed_tflops": v.get("executed_flops", v["flops"]) / (v["ms"] / 1e3) / 1e12 if v["ms"] > 0 else None}
                 for k, v in sorted(K.PROFILE.shapes.items(), key=lambda kv: -kv[1]["ms"])}
        table["__entry_points__"] = {k: {"calls_per_step": n / prof_steps, "ms_per_step": ms / prof_steps}
                                     for k, (n, ms) in sorted(call_ms.items(), key=lambda kv: -kv[1][1])}
        with open(args.shape_profile, "w") as f:
            json.dump(table, f, indent=1)

    # ---- device-resident timing (value): the step replayed as ONE CUDA graph (same kernels, no host launches) ----
    graphed = None
    if not args.no_graph:
        from multimodal_alzheimer_b200.graphed import GraphedStep
        torch.cuda.empty_cache()
        graphed = GraphedStep(lambda: step(data), warmup=1)
        if not graphed.captured:
            if rank == 0:
                print(f"bench.py: CUDA graph capture unavailable ({graphed.error}); timing the eager path",
                      file=sys.stderr)
            graphed = None
    run_step = (lambda: graphed.replay()) if graphed else (lambda: step(data))
    for _ in range(2):
        run_step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _ = run_step()
    e1.record()
    barrier()
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    value = vols_per_step * args.steps / (ms_total / 1e3)
    loss_val = float(loss.detach())

    # ---- end-to-end timing: pinned host inputs -> H2D -> step -> D2H loss, every step ------------------
    e2e = None
    if not args.no_e2e:
        host = {k: v.pin_memory() for k, v in host_data.items()}
        h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
        copy_stream = torch.cuda.Stream(device=dev)
        NB = 4   # input buffers / loss slots in flight: uploads run NB-1 steps ahead, losses are read NB-1 steps behind
        bufs = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(NB)]
        ready = [torch.cuda.Event() for _ in range(NB)]
        free = [torch.cuda.Event() for _ in range(NB)]
        loss_host = [torch.empty((), dtype=torch.float64).pin_memory() for _ in range(NB)]
        loss_done = [torch.cuda.Event() for _ in range(NB)]
        losses_read = []

        def upload(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[i % NB])
                for k, v in host.items():
                    bufs[i % NB][k].copy_(v, non_blocking=True)
                ready[i % NB].record(copy_stream)

        def e2e_run(n):
            """Step i's inputs cross PCIe while earlier steps compute (NB-1 uploads in flight); every step's loss is copied to
            pinned host memory and read there inside the timed region, NB-1 steps behind the launches - a host that
            waited for step i-1 before enqueueing step i+1 would expose every rank's launch jitter to all ranks through
            the step's collectives (measured at 8 GPUs in round 1: e2e 16 % under the device-resident number)."""
            for ev in free:
                ev.record()
            for j in range(min(NB - 1, n)):
                upload(j)
            for i in range(n):
                if i + NB - 1 < n:
                    upload(i + NB - 1)
                torch.cuda.current_stream().wait_event(ready[i % NB])
                if graphed:  # device-to-device hand-over into the graph's fixed input tensors, then one replay
                    for k, v in bufs[i % NB].items():
                        data[k].copy_(v, non_blocking=True)
                    free[i % NB].record()
                    l, _ = graphed.replay()
                else:
                    l, _ = step(bufs[i % NB])
                    free[i % NB].record()
                loss_host[i % NB].copy_(l.detach(), non_blocking=True)
                loss_done[i % NB].record()
                if i >= NB - 1:
                    j = i - (NB - 1)
                    loss_done[j % NB].synchronize()
                    losses_read.append(float(loss_host[j % NB]))
            for j in range(max(0, n - (NB - 1)), n):
                loss_done[j % NB].synchronize()
                losses_read.append(float(loss_host[j % NB]))

        e2e_run(NB)
        barrier()
        losses_read.clear()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_run(args.steps)
        t1.record()
        barrier()
        ems = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": vols_per_step * args.steps / (float(ems) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d_bytes * world), "d2h_bytes_per_step": 8 * world,
               "ms_per_step": float(ems) / args.steps, "losses_read_on_host": len(losses_read),
               "how": f"pinned host raw volumes+masks -> cudaMemcpyAsync ({NB} buffers on a copy stream) -> "
                      "general_step/backward/allreduce/Adam -> every step's loss copied to pinned host memory and read "
                      f"there {NB - 1} steps behind the launches, all inside the timed region"}

    if rank != 0:
        return
    peaks = load_peaks()
    step_flops = W.train_flops_per_sample(args.workload, depth, volume) * global_batch / world  # per rank
    empty = {"flops": 0, "executed_flops": 0, "ms": 0.0, "n": 0}
    tc = prof.get("tc_kmajor", empty)
    hl = prof.get("tc_halo", empty)
    wg = prof.get("tc_wgrad", empty)
    st = prof.get("tc_stem", empty)
    dr = prof.get("direct", empty)
    sc = prof.get("tc_small", empty)

    def tf(d, key="flops"):
        return d.get(key, d["flops"]) / (d["ms"] / 1e3) / 1e12 if d["ms"] > 0 else None

    # Per-kernel numbers come from the eager pass: every conv launch is bracketed by CUDA events on its stream and
    # the host issues kernels ~2x slower than the GPU retires them, so each kernel runs alone at burst clocks ->
    # the burst figure of MEASURED_PEAKS.json is the denominator.  `achieved` counts ALGORITHMIC FLOPs (DESIGN.md
    # section 4); taps / position boxes that lie entirely in the zero padding are skipped, so the rate the tensor pipe
    # actually executes is `executed` (from the library's planner, adni_conv3d_plan_info), bounded by the peak.
    peak = peaks["bf16_tflops"]

    def fam(d):
        return {"achieved": tf(d), "executed": tf(d, "executed_flops"),
                "frac": (tf(d) / peak) if tf(d) else None,
                "frac_executed": (tf(d, "executed_flops") / peak) if tf(d) else None,
                "launches_per_step": d["n"] / args.steps, "kernel_ms_per_step": d["ms"] / args.steps}

    dominant = max((tc, hl, wg, st, sc, dr), key=lambda d: d["ms"])
    dom_name = {id(tc): "igemm_kmajor_kernel (Conv3d fprop + dgrad, tcgen05/TMEM, TMA box loads)",
                id(hl): "igemm_halo_kernel (layer1/layer2 3x3x3 convs, plane ring in smem)",
                id(wg): "wgrad2_kernel (Conv3d wgrad, MN-major operands)",
                id(st): "stem_fprop_plane / stem_wgrad_plane_kernel (tcgen05, 1-channel stem)",
                id(sc): "small-channel conv kernels (Small_PET_CNN)",
                id(dr): "direct_conv (CUDA cores)"}[id(dominant)]
    roofline = {
        "bound": "tensor", "kernel": dom_name,
        "achieved": tf(dominant), "peak": peak, "unit": "TFLOP/s",
        "frac": (tf(dominant) / peak) if tf(dominant) else None, "traffic": None,
        "executed": tf(dominant, "executed_flops"),
        "frac_executed": (tf(dominant, "executed_flops") / peak) if tf(dominant) else None,
        "peak_source": peaks["source"] + " (burst figure: every kernel is timed alone with CUDA events in the eager pass); "
                       "achieved = algorithmic FLOPs / time, executed = FLOPs actually issued (all-padding taps skipped)",
        "launches_per_step": dominant["n"] / args.steps, "kernel_ms_per_step": dominant["ms"] / args.steps,
        "timed_in": "eager profiled pass of the same step (CUDA events around every conv launch)",
        "algorithmic_flops_per_step": dominant["flops"] / args.steps,
        "kernels": {
            "igemm_kmajor_kernel": fam(tc), "igemm_halo_kernel": fam(hl), "wgrad2_kernel": fam(wg),
            "stem_plane_kernels": fam(st), "small_channel_conv": fam(sc), "direct_conv (CUDA cores)": fam(dr),
        },
        "whole_step_tensor_frac": step_flops / (ms_total / args.steps / 1e3) / 1e12 / peaks["bf16_tflops_sustained"],
        "whole_step_peak": peaks["bf16_tflops_sustained"],
    }
    # HBM-bound kernels of the same eager pass (north_star: achieved HBM GB/s for norm, pool and loss kernels):
    # algorithmic bytes (kernels.call_hbm) / CUDA-event time, against the measured copy bandwidth
    hbm_peak = peaks["hbm_gbs"]
    roofline["hbm_kernels"] = {
        tag[4:]: {"achieved_gbs": d["flops"] / (d["ms"] / 1e3) / 1e9 if d["ms"] > 0 else None,
                  "frac": d["flops"] / (d["ms"] / 1e3) / 1e9 / hbm_peak if d["ms"] > 0 else None,
                  "algorithmic_mb_per_step": d["flops"] / args.steps / 1e6,
                  "launches_per_step": d["n"] / args.steps, "kernel_ms_per_step": d["ms"] / args.steps}
        for tag, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]) if tag.startswith("hbm_")}
    roofline["hbm_peak_gbs"] = hbm_peak
    # DRAM bytes per launch are a profiler quantity: they come from the committed ncu capture of THIS command line
    # (tools/gpu_profile.sh -> profiles/r02_traffic.json), and only for the configuration that capture ran
    # (default workload, 32 pairs on one GPU); everything else reports null
    traffic_path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if (args.workload == "pet_mri_fusion_r18" and world == 1 and global_batch == 32 and depth == 18
            and volume == (128, 128, 128) and dominant is tc and os.path.exists(traffic_path)):
        with open(traffic_path) as f:
            cap = json.load(f)
        roofline["traffic"] = cap.get("dram_bytes_per_launch")
        # algorithmic bytes of the same launches: activations in + out (+ addend) and the weights, once each, bf16
        import re as _re
        alg, nl = 0.0, 0
        for key, rec in getattr(K.PROFILE, "shapes", {}).items():
            mm = _re.match(r"tc_kmajor (fprop|dgrad\S*) N(\d+) (\d+)x(\d+)x(\d+) (\d+)->(\d+) k(\d+) s(\d+) d\d+", key)
            if not mm:
                continue
            n_, d_, h_, w_, ci, co, k_, s_ = (int(v) for v in mm.groups()[1:])
            pos_in, pos_out = n_ * d_ * h_ * w_, n_ * d_ * h_ * w_ // (s_ ** 3)
            alg += rec["n"] * 2.0 * (pos_in * ci + pos_out * co + co * ci * k_ ** 3)
            nl += rec["n"]
        roofline["traffic_algorithmic_bytes_per_launch"] = alg / nl if nl else None
        roofline["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per igemm_kmajor launch, averaged over the "
                                    "launches of one step, from the committed ncu capture of this command line: " + cap.get("source", ""))
    else:
        roofline["traffic_note"] = ("dram bytes per launch are a profiler quantity (ncu) captured for the default workload at 32 "
                                    "pairs on one GPU only (profiles/r02_traffic.json): null here")
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cstep, cvols, threads = cpu_reference_step_fn(args, args.cpu_sample_pairs)
        cstep()
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            cstep()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        cpu = {"value": cvols / best, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"the first {args.cpu_sample_pairs} sample(s) of the same global batch ({cvols} volumes "
                         f"{'x'.join(map(str, volume))}): fp64 torch.quantile normalisation + fp32 fwd + fp64 loss + bwd + "
                         f"Adam, best of 2 after 1 warm-up",
               "pinning": PORT_PINNING}
    line = {
        "metric": metric_name(args, w, depth, volume), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, w, depth, volume, world, global_batch),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "parity": parity, "loss": loss_val, "sync_bn_exchange": dp.exchange_status()["mode"],
        "gradient_exchange": f"{type(buckets).__name__}: {dp.gradient_exchange_status()['mode']}", "launch_mode": "cuda_graph" if graphed else "eager",
        "peak_memory_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
        "eager": {"ms_per_step": eager_ms, "host_issue_ms_per_step": host_issue_ms,
                  "note": "same step launched kernel by kernel from Python with the two encoder branches serialised "
                          "(the pass the per-kernel roofline events come from)"},
    }
    print(json.dumps(line), flush=True)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass


if __name__ == "__main__":
    main()

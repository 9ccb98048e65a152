"""Summarise gpurun_out/launches.csv (ncu launch list with gpu__time_duration + DRAM bytes) into profiles/."""
import collections, csv, json, re, sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
tag = sys.argv[2] if len(sys.argv) > 2 else "r01"
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))


def val(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    if row["Metric Name"].startswith("gpu__time"):
        return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}[u]  # -> us
    return {"byte": v, "Kbyte": v * 1e3, "Mbyte": v * 1e6, "Gbyte": v * 1e9}[u]  # -> bytes


launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    d[r["Metric Name"]] = val(r)
ids = list(launch.keys())
# the timed step = launches after the last quantile histogram pass 0 (first kernel of a step)
starts = [i for i, k in enumerate(ids) if "q_hist_kernel<0>" in launch[k]["name"]]
step = [launch[k] for k in ids[starts[-1]:]] if starts else list(launch.values())
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in step:
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("adni::", "").replace("<unnamed>::", "")
    a = agg[name]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
out = [f"# ncu launch list of one training step (`bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline`, B=32 pairs, 1x B200)",
       "", "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
       f"launches in the step: {len(step)}; sum of kernel time: {tot / 1e3:.2f} ms", "",
       "| share | time ms | launches | DRAM read GB | DRAM write GB | kernel |", "|---|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    out.append(f"| {a[1] / tot * 100:.2f}% | {a[1] / 1e3:.2f} | {a[0]} | {a[2] / 1e9:.3f} | {a[3] / 1e9:.3f} | `{k[:90]}` |")
open(f"profiles/{tag}_launch_summary.md", "w").write("\n".join(out) + "\n")
ig = [a for k, a in agg.items() if "igemm_kmajor" in k]
hl = [a for k, a in agg.items() if "igemm_halo" in k]
n = sum(a[0] for a in ig)
traffic = {"kernel": "igemm_kmajor_kernel", "launches_per_step": n,
           "dram_bytes_per_launch": (sum(a[2] + a[3] for a in ig) / n) if n else None,
           "share_of_step_kernel_time": sum(a[1] for a in ig) / tot,
           "halo_kernel": {"launches_per_step": sum(a[0] for a in hl),
                           "dram_bytes_per_launch": (sum(a[2] + a[3] for a in hl) / max(1, sum(a[0] for a in hl))),
                           "share_of_step_kernel_time": sum(a[1] for a in hl) / tot},
           "source": f"profiles/{tag}_launch_summary.md (ncu dram__bytes_read.sum + dram__bytes_write.sum, average over the "
                     "igemm_kmajor fprop+dgrad launches of one step, global batch 32)"}
json.dump(traffic, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print("\n".join(out[:24]))
print(traffic)

"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) as a markdown table.
    python tools/ncu_full_summary.py raw.csv out.md "title / command line"
"""
import csv
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("time us", "gpu__time_duration.sum", 1.0, "{:.1f}"),
        ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1.0, "{:.1f}"),
        ("utcmma bf16 % of peak", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", 1.0, "{:.1f}"),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0, "{:.1f}"),
        ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0, "{:.1f}"),
        ("L1/TEX %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 1.0, "{:.1f}"),
        ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0, "{:.1f}"),
        ("DRAM read MB", "dram__bytes_read.sum", None, "{:.1f}"),
        ("DRAM write MB", "dram__bytes_write.sum", None, "{:.1f}"),
        ("regs", "launch__registers_per_thread", 1.0, "{:.0f}"),
        ("warp instr M", "smsp__inst_executed.sum", 1e-6, "{:.1f}")]
units = rows[1]
out = [f"# {title}", "", "Times are under the profiler (cold caches, serialised replays); raw export beside this file.", "",
       "| kernel | grid | " + " | ".join(c[0] for c in cols) + " |", "|---|---|" + "---|" * len(cols)]
for r in data:
    if len(r) < len(hdr):
        continue
    cells = []
    for name, key, scale, fmt in cols:
        if key not in ix or r[ix[key]] in ("", "n/a"):
            cells.append("")
            continue
        v = float(r[ix[key]].replace(",", ""))
        if scale is None:   # bytes -> MB by the unit row
            u = units[ix[key]]
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        else:
            v *= scale
        cells.append(fmt.format(v))
    k = r[ix["Kernel Name"]].replace("void ", "").replace("adni::", "").replace("<unnamed>::", "")[:44]
    out.append(f"| `{k}` | {r[ix['Grid Size']]} | " + " | ".join(cells) + " |")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))

"""Turn `ncu --page raw --csv` output and a launch-list csv into the markdown kept under profiles/."""
import collections
import csv
import sys

raw_csv, launches_csv, out_md, title = sys.argv[1:5]
rows = list(csv.reader(open(raw_csv)))
hdr = rows[0]
want = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM read MB"), ("dram__bytes_write.sum", "DRAM write MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__waves_per_multiprocessor", "waves/SM"), ("smsp__inst_executed.sum", "warp instr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %")]
ki = hdr.index("Kernel Name")
idx = [(hdr.index(m), n) for m, n in want if m in hdr]
units = rows[1]
lines = ["# " + title, "", "## `ncu --set full` (one launch per kernel, cold cache, serialised)", "",
         "| kernel | " + " | ".join(n for _, n in idx) + " |", "|---|" + "---|" * len(idx)]
seen = set()
for r in rows[2:]:
    name = r[ki].rsplit("(", 1)[0].replace("<unnamed>::", "").replace("void ", "")
    if len(sys.argv) > 5 and sys.argv[5] == "all":
        name = f"{name} #{sum(1 for s_ in seen if s_.startswith(name + ' #')) + 1}"
    if name in seen:
        continue
    seen.add(name)
    vals = []
    for i, n in idx:
        v = r[i]
        u = units[i]
        try:
            f = float(v.replace(",", ""))
            if n.endswith("MB"):
                f *= {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u.lower(), 1.0)
            if n == "time us":
                f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u.lower(), 1.0)
            v = f"{f:.2f}" if f < 1000 else f"{f:.0f}"
        except ValueError:
            pass
        vals.append(v)
    lines.append(f"| `{name}` | " + " | ".join(vals) + " |")
lines += ["", "## launch list (`ncu --metrics gpu__time_duration.sum`, every launch of the same command)", "",
          "| kernel | launches | avg us | min us | max us | share of step |", "|---|---|---|---|---|---|"]
lr = [r for r in csv.reader(open(launches_csv)) if len(r) > 5]
lh = lr[0]
k2, v2 = lh.index("Kernel Name"), lh.index("Metric Value")
d = collections.OrderedDict()
for r in lr[1:]:
    name = r[k2].rsplit("(", 1)[0].replace("<unnamed>::", "").replace("void ", "")
    if name.startswith("at::") or name.startswith("at_cuda") or "nccl" in name.lower():
        continue  # torch's input generators / fills of the harness, not the product's kernels
    d.setdefault(name, []).append(float(r[v2].replace(",", "")) / 1e3)
main = {k: v for k, v in d.items() if "reset" not in k}
tot = sum(sum(v) / len(v) for v in main.values())
for k, v in d.items():
    share = f"{100 * (sum(v) / len(v)) / tot:.1f} %" if k in main else "(setup)"
    lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {min(v):.1f} | {max(v):.1f} | {share} |")
open(out_md, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

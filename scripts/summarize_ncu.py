"""Summarise gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and
gpurun_out/prof_raw.csv (ncu --set full, --page raw --csv) into profiles/<tag>_*.{md,csv}."""
import collections
import csv
import io
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
note = sys.argv[2] if len(sys.argv) > 2 else ""

lines = [l for l in open("gpurun_out/launches.csv") if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(io.StringIO("".join(lines))):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
    name = row["Kernel Name"].split("(")[0].replace("void ", "")[:90]
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
with open(f"profiles/{tag}_launches.md", "w") as f:
    f.write(f"# ncu launch list ({tag}) -- `ncu --metrics gpu__time_duration.sum --clock-control none` on "
            f"`python scripts/profile_step.py`\n\n{note}\n\nPer-launch times under ncu are cold-cache and "
            f"serialised: compare SHARES with bench.py's `kernels`, not absolutes.\n\n"
            f"total {s:.1f} ms over {sum(cnt.values())} launches\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:30]:
        f.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / s:.2f}% |\n")

try:
    rows = list(csv.reader(open("gpurun_out/prof_raw.csv")))
except FileNotFoundError:
    sys.exit(0)
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]
idx = {h: i for i, h in enumerate(hdr)}
cols = [w for w in want if w in idx]
with open(f"profiles/{tag}_ncu_full.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[idx[c]] for c in cols])
    for r in rows[2:]:
        w.writerow([r[idx[c]][:70] for c in cols])
print("wrote profiles/", tag)

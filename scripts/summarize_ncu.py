"""Summarise the ncu outputs of scripts/gpu_ncu.sh into profiles/<tag>_*:
  gpurun_out/launches.csv            -> profiles/<tag>_launches.md        (kernel shares of a run)
  gpurun_out/prof_{wgrad,gru,gemm}_raw.csv -> profiles/<tag>_ncu_{wgrad,gru,gemm}.csv (per launch: duration, DRAM bytes,
                                         tensor-pipe / L2 / DRAM utilisation, registers, smem)
  and profiles/<tag>_traffic.json: DRAM bytes per launch of each kernel family, read by bench.py (roofline.traffic).
usage: python scripts/summarize_ncu.py <tag> [workload] [note]"""
import collections
import csv
import io
import json
import os
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
workload = sys.argv[2] if len(sys.argv) > 2 else "ithor_b256"
note = sys.argv[3] if len(sys.argv) > 3 else ""

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
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:32]:
        f.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / s:.2f}% |\n")

want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]
# bench.py family tag -> (substrings that must ALL appear in the ncu kernel name)
FAMILY = {"wgrad": (("tc_wgrad_tma",), ("cin1_wgrad",), ("cin3_wgrad",), ("tc_wgrad_kernel",), ("fullw_wgrad",)),
          "wgrad16": (("tc_wgrad_h16",),),
          "gru_step": (("gru_persist_kernel",),), "gru_bwd": (("gru_bwd_ksplit",),),
          "mfcc": (("mfcc_kernel",),),
          # the 16-bit forward and dgrad GEMMs are the same kernel template: one entry serves both tags
          "gemm_fwd16": (("halo_conv_fwd_kernel<0>",), ("tc_gemm_persist", ", 1>")),
          "gemm_dgrad16": (("halo_conv_dgrad_kernel",), ("tc_gemm_persist", ", 1>")),
          "gemm_fwd": (("tc_gemm_persist", ", 0>"),), "gemm_dgrad": (("tc_gemm_persist", ", 0>"),)}
traffic = {}


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


for part in ("wgrad", "gru", "gemm", "halo"):
    path = f"gpurun_out/prof_{part}_raw.csv"
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [w for w in want if w in idx]
    with open(f"profiles/{tag}_ncu_{part}.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in rows[2:]:
            w.writerow([r[idx[c]][:70] for c in cols])
    for fam, keys in FAMILY.items():
        sel = [r for r in rows[2:] if any(all(k in r[idx["Kernel Name"]] for k in alt) for alt in keys)]
        if not sel:
            continue
        rd = sum(to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) for r in sel)
        wr = sum(to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]) for r in sel)
        traffic[fam] = {"launches_captured": len(sel), "dram_bytes_per_launch": (rd + wr) / len(sel),
                        "dram_read_bytes": rd, "dram_write_bytes": wr,
                        "source": f"profiles/{tag}_ncu_{part}.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
if traffic:
    json.dump({workload: traffic}, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print("wrote profiles/", tag, sorted(traffic))

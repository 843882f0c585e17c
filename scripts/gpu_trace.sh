#!/bin/bash
mkdir -p gpurun_out
VAR_GRU_TRACE=1 python scripts/prof_dump.py ithor 256 serial > gpurun_out/pd_trace.log 2>&1
python - <<'P'
import csv
for name in ("fwd",):
    rows = list(csv.DictReader(open(f"gpurun_out/gru_trace_{name}.csv")))
    rows = [r for r in rows if int(r["released"]) > 0]
    n = len(rows)
    def col(c): return [int(r[c]) for r in rows]
    st, acq, iss, tf, epi, rel = (col(c) for c in ("start", "acquired", "issued", "tfull", "epi_done", "released"))
    f = 1 / 1.965e3  # us per clock
    mid = range(5, n - 5)
    avg = lambda xs: sum(xs) / len(xs)
    print(name, "steps", n, "us/step", round(avg([(st[i + 1] - st[i]) * f for i in mid]), 2),
          "| wait", round(avg([(acq[i] - st[i]) * f for i in mid]), 2), "stream", round(avg([(iss[i] - acq[i]) * f for i in mid]), 2),
          "mma tail", round(avg([(tf[i] - iss[i]) * f for i in mid]), 2), "epilogue", round(avg([(epi[i] - tf[i]) * f for i in mid]), 2),
          "release", round(avg([(rel[i] - epi[i]) * f for i in mid]), 2), "next", round(avg([(st[i + 1] - rel[i]) * f for i in mid]), 2))
P

"""Rewrites section 1 of profiles/r2_ncu_final.md (per-kernel shares and DRAM bytes of the serialised C4 step) from
profiles/r2_launches_c4_step.csv; the other sections of that file quote the `--page raw --csv` exports unchanged."""
import collections
import csv
import json
import os
import re

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pr = os.path.join(root, "profiles")
md = open(os.path.join(pr, "r2_ncu_final.md")).read()
i, j = md.index("## 1. Serialised launch list"), md.index("## 2. `ncu --set full` exports")
rows = list(csv.reader(open(os.path.join(pr, "r2_launches_c4_step.csv"))))
hdr = [k for k, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
ki, mi, vi, ui = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
d = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
unit = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("msmp::", "")
    if n.startswith("at"):
        n = "ATen / CUB kernels (gradient adds, casts, fills)"
    v = float(r[vi].replace(",", ""))
    unit[r[mi]] = r[ui]
    col = {"gpu__time_duration.sum": 1, "dram__bytes_read.sum": 2, "dram__bytes_write.sum": 3}[r[mi]]
    d[n][col] += v
    if col == 1:
        d[n][0] += 1
sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tsc = {"ns": 1e-6, "us": 1e-3, "ms": 1, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1}[unit["gpu__time_duration.sum"]]
tot = sum(v[1] for v in d.values()) * tsc
b = json.load(open(os.path.join(pr, "r2_bench_n1.json")))
out = ["## 1. Serialised launch list of one C4 step (`r2_launches_c4_step.csv`, final code)\n",
       "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` over the single-stream eager step",
       f"(`scripts/ncu_step.py c4 8`: 131 072 nodes, 520 192 edges, T = 25): {sum(v[0] for v in d.values())} launches, {tot:.2f} ms of kernel time; "
       f"the replayed step on the same box: {b['ms_per_step']:.1f} ms",
       f"(`r2_bench_n1.json`, {b['clocks']['sm_mhz']:.0f} MHz under `sw_power_cap`), so the shares below are the shares of the step.  Per-launch times are "
       "cold-cache and serialised.\n",
       "| kernel | launches | ms | share | DRAM read MB / launch | DRAM written MB / launch | DRAM TB/s |", "|---|---:|---:|---:|---:|---:|---:|"]
ur, uw = sc[unit["dram__bytes_read.sum"]], sc[unit["dram__bytes_write.sum"]]
for k, v in sorted(d.items(), key=lambda x: -x[1][1]):
    t = v[1] * tsc
    if t / tot < 0.002:
        continue
    rd, wr = v[2] * ur, v[3] * uw
    out.append(f"| `{k}` | {v[0]} | {t:.3f} | {100 * t / tot:.1f} % | {rd / v[0] / 1e6:.1f} | {wr / v[0] / 1e6:.1f} | {(rd + wr) / (t * 1e-3) / 1e12:.2f} |")
out += ["", "Against the list at the time of the `--set full` captures of section 2 (298 launches, 38.56 ms; 59 framework kernels, 0.97 ms; `k_lem_fwd_tc` 5.89 ms, "
        "`k_unpack` 0.281 ms): the criterion and the input assembly are own launches now", "(`k_sse_*`, `k_node_features`, `k_lem_inputs`), `k_unpack` walks its "
        "four rows together, the LEM forward kernel's gate epilogues run block-wise behind the G GEMM.\n"]
open(os.path.join(pr, "r2_ncu_final.md"), "w").write(md[:i] + "\n".join(out) + "\n" + md[j:])
print("\n".join(out))

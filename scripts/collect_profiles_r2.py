"""Copies the outputs of scripts/final_profiles_r2.sh (gpurun_out/r2f_*) into profiles/ (tracked) and prints the per-op
summary of the serialised launch list (time share, DRAM bytes per launch) that bench.py's NCU_TRAFFIC table quotes."""
import collections, csv, json, os, shutil, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
shutil.copy(os.path.join(go, "r2f_bench_n1.json"), os.path.join(pr, "r2_bench_n1.json"))
shutil.copy(os.path.join(go, "r2f_bench_ref.json"), os.path.join(pr, "r2_bench_reference_arm.json"))
shutil.copy(os.path.join(go, "r2f_launches_c4_step.csv"), os.path.join(pr, "r2_launches_c4_step.csv"))
with open(os.path.join(pr, "r2_bench_linear.txt"), "w") as f:
    f.write("# scripts/bench_linear.py, 131 072 rows, one layer's six node-GEMM launches (CUDA events, L2 flushed), final code of round 2\n")
    for title, name in (("k_linear_ts (default), fp32 parity mode", "r2f_bench_linear.txt"),
                        ("k_linear_tma (MSMP_LINEAR_TS=0; with the round-2 epilogue)", "r2f_bench_linear_tma.txt"),
                        ("k_linear_ts, reduced-precision mode (MSMP_PRECISION=bf16)", "r2f_bench_linear_reduced.txt")):
        f.write("== " + title + "\n" + open(os.path.join(go, name)).read())
    f.write("\n# start of round 2 (k_linear_tma, old epilogue) and the MSMP_LIN_DBG ablation that located the bound:\n"
            "#   total 0.655 ms; no epilogue memory traffic / math (dbg 1): 0.446; no conversion (2): 0.567; no MMAs (4): 0.573; loads only (7): 0.262\n"
            "#   first k_linear_ts (8 epilogue warps): 0.633, dbg 1: 0.295 -> the epilogue was the bound\n")
with open(os.path.join(pr, "r2_bench_wgrad_final.jsonl"), "w") as f:
    f.write("# final code of round 2: 'ws' = msmp_wgrad_ws (k_wgrad_ts for <= 160 operand columns without side columns, else k_wgrad_ws), 'tc' = k_wgrad_tc; includes the reduction launch\n")
    f.write(open(os.path.join(go, "r2f_bench_wgrad.jsonl")).read())
keep = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio']
for rep, out in (("r2f_linear_ts", "r2_linear_ts_ncu_raw.csv"), ("r2f_wgrad", "r2_wgrad_ncu_raw.csv")):
    txt = subprocess.run(["ncu", "-i", os.path.join(go, rep + ".ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h = rows[0]
    idx = [next(i for i, c in enumerate(h) if c == k or c.endswith('.' + k)) for k in keep if any(c == k or c.endswith('.' + k) for c in h)]
    with open(os.path.join(pr, out), "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in idx])
    print(out, len(rows) - 2, "launches")
    for r in rows[2:]:
        print("   ", r[h.index('Kernel Name')][:34], r[h.index('Grid Size')], [r[i] for i in idx[3:8]])
rows = list(csv.reader(open(os.path.join(pr, "r2_launches_c4_step.csv"))))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn, mn, mu, mv = (h.index(x) for x in ('Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value'))
dd = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1}
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    n = r[kn].split('(')[0][:50]
    dd[n][r[mn]] += float(r[mv].replace(',', '')) * scale.get(r[mu], 1)
    if r[mn] == 'gpu__time_duration.sum': cnt[n] += 1
tot = sum(v['gpu__time_duration.sum'] for v in dd.values())
print('serialised step: total ms', round(tot, 3), 'launches', sum(cnt.values()))
for k, v in sorted(dd.items(), key=lambda x: -x[1]['gpu__time_duration.sum'])[:12]:
    t, rd, wr = v['gpu__time_duration.sum'], v['dram__bytes_read.sum'], v['dram__bytes_write.sum']
    print(f"{k:50s} n={cnt[k]:3d} {t:8.3f} ms {t / tot * 100:5.1f}% read {rd / 1e9:6.2f} GB write {wr / 1e9:6.2f} GB -> {(rd + wr) / t / 1e9:5.2f} TB/s per launch {(rd + wr) / cnt[k] / 1e6:9.1f} MB")
d = json.load(open(os.path.join(pr, "r2_bench_n1.json")))
print("bench:", d["ms_per_step"], "ms", d["value"], "nodes/s; e2e", d["e2e"]["ms_per_step"], "; clocks", d["clocks"]["sm_mhz"], "; C2", d["configs"]["C2"]["ms_per_step"],
      "C3", d["configs"]["C3"]["ms_per_step"], "reduced", d["bf16_mode"]["ms_per_step"], "layer_c5", [x["ms_fwd_bwd"] for x in d["layer_c5"]])
print("roofline:", d["roofline"]["kernel"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"].get("frac_of_3xtf32_ceiling"), d["roofline"]["traffic"])
print({k: (v["ms_per_step"], v["tflops"]) for k, v in d["ops"].items()})
r = json.load(open(os.path.join(pr, "r2_bench_reference_arm.json"))); print("reference arm:", r["ms_per_step"], r["value"])

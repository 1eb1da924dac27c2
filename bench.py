#!/usr/bin/env python
"""Benchmark of the MSMP-PDE hot path (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c2|c3]

Headline workload (BASELINE.json configs[3], the configuration the metric "graph-nodes/sec at 1/2/4/8 B200" is quoted
on): MSMP-PDE2D training step (MP_PDE_Solver2DLEMLinGated: LEM encoder + 6 gated layer pairs + Conv1d decoder,
1 409 002 parameters, time_window 25) on 128 x 128 lattice graphs, 8 graphs = 131 072 nodes / 520 192 edges per GPU,
data parallel over whole graphs (weak scaling).  One step = forward + loss (sqrt of the batch-global summed squared error,
experiments/train_helper.py:126,138) + backward + (N > 1: ONE gradient all-reduce over NCCL) + AdamW update, with the
optimizer experiments/train.py:410 builds.  Metric: graph-nodes per second, whole job.  At N = 1 the other BASELINE
configs (C2, C3, the config-5 single-layer shapes) and the bf16 operand mode are measured as sub-records of the same
JSON line.  Synthetic data, random-init weights.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GNN fwd+bwd graph-nodes/sec"
UNIT = "nodes/s"
TW = 25
WORKLOADS = {
    "c4": dict(graphs=8, label="C4: MSMP-PDE2D (MP_PDE_Solver2DLEMLinGated) on MSWG3-shaped 128 x 128 4-neighbour lattice graphs, "
                               "{B} graphs x 16384 nodes per GPU, tw=25, fwd+loss+bwd+AdamW"),
    "c2": dict(graphs=64, label="C2: MSMP-PDE2D (MP_PDE_Solver2DLEMLinGated) RP shape, {B} graphs x 100 nodes per GPU, tw=25, "
                                "588 edges/graph, fwd+loss+bwd+AdamW"),
    "c3": dict(graphs=64, label="C3: MSMP-PDE2D (MP_PDE_Solver2DLEMLinGated) RPU shape (pseudo-random grid, kNN k=3), {B} graphs "
                                "x 100 nodes per GPU, tw=25, fwd+loss+bwd+AdamW"),
}
REF_GRAPHS = {"c4": 8, "c2": 64, "c3": 64}          # graphs per step of the CPU reference arm: the full per-GPU batch


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), tensor_tflops=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json: copy GB/s, cuBLAS bf16 sustained TFLOP/s)")
    return dict(hbm_gbs=6650.0, tensor_tflops=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def _synth_module(load_native: bool):
    """msmp_pde_b200.synth (host-side generators of the BASELINE shapes).  The CPU reference arm imports it WITHOUT
    running the package __init__, which would dlopen libmsmp_b200.so: a stub package object makes the submodule import
    resolve on its own (synth and the compat shims it uses are pure torch / numpy)."""
    if not load_native and "msmp_pde_b200" not in sys.modules:
        pkg = types.ModuleType("msmp_pde_b200")
        pkg.__path__ = [os.path.join(ROOT, "msmp_pde_b200")]
        sys.modules["msmp_pde_b200"] = pkg
    from msmp_pde_b200 import synth
    return synth


def _make(synth, workload: str, B: int, seed: int, dtype=None):
    kw = {} if dtype is None else {"dtype": dtype}
    if workload == "c4":
        return synth.config_c4(B=B, side=128, tw=TW, seed=seed, **kw)
    if workload == "c2":
        return synth.config_c2(B=B, nx=100, tw=TW, seed=seed, **kw)
    if workload == "c3":
        return synth.config_c3(B=B, nx=100, tw=TW, seed=seed, **kw)
    raise ValueError(workload)


def _loss(pred, y):
    import torch
    return torch.sqrt(torch.nn.functional.mse_loss(pred, y, reduction="sum"))


# ----------------------------------------------------------------------------------------------- ours
def _time_step(step, flush, K, graph, read_loss):
    """K steps, one CUDA-event pair each on the launching stream, L2 flushed between steps -> list of ms."""
    import torch
    evs = []
    for _ in range(K):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        loss = step(graph)          # graph=None: inputs already resident in HBM; else H2D copies inside
        if read_loss:
            loss.item()
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    return [s.elapsed_time(e) for s, e in evs]


def _time_e2e(step, flush, K, host_graph):
    """End to end through the public API with a prefetching loader: every step's inputs are copied from pinned host memory
    (step.prefetch, on a copy stream: the copy of step k+1 overlaps the compute of step k) and every step's loss is read
    back.  K steps are timed as ONE region (first copy issued inside it, last loss read inside it): ms per step."""
    import torch
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    step.prefetch(host_graph)
    for k in range(K):
        flush.zero_()
        loss = step(None)               # waits for the staged inputs, D2D into the static buffers, replays
        if k + 1 < K:
            step.prefetch(host_graph)   # next step's H2D copies, behind this step's kernels on the copy stream
        loss.item()
    e.record()
    torch.cuda.synchronize()
    return [s.elapsed_time(e) / K]


def _sub_config(name, B, dev, flush, steps, precision=None):
    """ms/step of another BASELINE config on this GPU (captured step, inputs resident) -- a sub-record, not the headline."""
    import torch
    from msmp_pde_b200 import models_gnn2D, ops, synth
    from msmp_pde_b200.train_step import GraphedTrainStep
    prev = ops.PRECISION
    if precision:
        ops.PRECISION = precision
    try:
        pde, data, meta = _make(synth, name, B, 0)
        torch.manual_seed(0)
        model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, TW, 128, 6, meta["eq_variables"]).to(dev)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        g = data.clone().to(dev)
        step = GraphedTrainStep(model, opt, g, warmup=3)
        for _ in range(3):
            step(None)
        ts = _time_step(step, flush, steps, None, False)
        ms = sum(ts) / len(ts)
        N, E = g.x.shape[0], g.edge_index.shape[1]
        out = {"workload": WORKLOADS[name]["label"].format(B=B), "nodes": N, "edges": E, "ms_per_step": round(ms, 4),
               "nodes_per_s": round(N / ms * 1e3, 1), "loss": float(step.loss)}
        del step, model, opt, g
        torch.cuda.empty_cache()
        return out
    finally:
        ops.PRECISION = prev


def _layer_c5(n_nodes, degree, topology, dev, reps=5):
    """BASELINE config 5: ONE GNN_Layer(128,128,128,25,1) forward + backward on a large synthetic graph (eager launches,
    graph preparation cached and excluded).  Algorithmic FLOPs per SURVEY.md section 8(d): fwd + bwd = 3 x forward."""
    import torch
    from msmp_pde_b200 import layers, synth
    g = synth.large_graph(n_nodes, degree, topology=topology, nodes_per_graph=100, seed=0)
    t = {k: v.to(dev) for k, v in g.items()}
    torch.manual_seed(0)
    layer = layers.GNN_Layer(128, 128, 128, 25, 1).to(dev)
    x = t["x"].clone().requires_grad_(True)

    def step():
        out = layer(x, t["u"], t["pos"], t["variables"], t["edge_index"], t["batch"])
        out.backward(out.detach())
        x.grad = None
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    E = n_nodes * degree
    alg = 3.0 * (E * (2 * 283 * 128 + 2 * 128 * 128) + n_nodes * (2 * 257 * 128 + 2 * 128 * 128))
    del layer, x, t, g
    torch.cuda.empty_cache()
    return {"nodes": n_nodes, "in_degree": degree, "topology": topology, "ms_fwd_bwd": round(ms, 3),
            "nodes_per_s": round(n_nodes / ms * 1e3), "algorithmic_tflops": round(alg / ms / 1e9, 1)}


def tf32_peak(dev, seconds=1.0):
    """cuBLAS TF32 dense throughput (8192^3) measured in this run: burst = best of 10, sustained = back to back for
    `seconds`.  The error-compensated 3xTF32 GEMMs of the fp32 mode have a ceiling of one third of it."""
    import torch
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            torch.matmul(a, b, out=c)
            e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        reps = max(10, int(seconds * 1e3 / best))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e.record()
        torch.cuda.synchronize()
        sus = s.elapsed_time(e) / reps
        fl = 2.0 * n ** 3
        return {"burst_tflops": round(fl / best / 1e9, 1), "sustained_tflops": round(fl / sus / 1e9, 1),
                "how": f"torch.matmul fp32 inputs, allow_tf32, {n}^3, best of 10 / {reps} back to back"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def run_ours(args):
    import torch
    import torch.distributed as dist
    from msmp_pde_b200 import models_gnn2D, ops, synth
    from msmp_pde_b200.train_step import GraphedTrainStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = args.workload
    B = args.graphs_per_gpu or WORKLOADS[wl]["graphs"]

    def say(msg):
        if args.verbose:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)
    saved_stdout = None
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        # NCCL writes its version banner to the C-level stdout; the contract is ONE JSON line there, so fd 1 points at
        # stderr until that line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dp_check = None
    if world > 1:
        # the data-parallel step against the single-GPU full-batch step on the same global batch, before any timing
        from msmp_pde_b200.dp import dp_selfcheck
        pde_c, data_c, meta_c = synth.config_c2(B=3 * world, nx=100, tw=TW, seed=11)

        def make_model():
            torch.manual_seed(0)
            return models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde_c, TW, 128, 6, meta_c["eq_variables"])
        dp_check = dp_selfcheck(make_model, lambda m: torch.optim.AdamW(m.parameters(), lr=1e-4), data_c, dev)
        dp_check["max_rel_err"] = max(dp_check["loss_rel_err"], dp_check["grad_max_rel_err"])
        say(f"dp_check {dp_check}")
    pde, data, meta = _make(synth, wl, B, rank)                  # float64 host tensors (the reference feeds float64)
    torch.manual_seed(0)
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, TW, 128, 6, meta["eq_variables"]).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)          # exactly experiments/train.py:410
    N, E = data.x.shape[0], data.edge_index.shape[1]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    resident = data.clone().to(dev)
    pinned = data.clone().apply(lambda t: t.pin_memory())
    h2d = sum(t.numel() * t.element_size() for t in (getattr(pinned, k) for k in pinned.keys())
              if torch.is_tensor(t) and t.is_floating_point())
    say("model built")
    step = GraphedTrainStep(model, opt, resident, warmup=max(3, args.warmup), use_graph=not args.eager)
    say("step captured")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(None)
    barrier()
    # ---- per-op CUDA events (roofline) need eager launches: one short eager pass, not part of `value`.  The pass runs
    # with the side streams switched off (ops.SERIALIZE): every kernel has the GPU to itself, so the event pair around an
    # op is that kernel's own duration (in the replayed step kernels of different streams share the SMs).
    ops.SERIALIZE = True
    for _ in range(2):
        step.eager()
    ops.PROFILE_EVENTS = {}
    ops.LAUNCHES = 0
    eager_ev = []
    for _ in range(3):
        flush.zero_()
        eager_ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
        eager_ev[-1][0].record()
        step.eager()
        eager_ev[-1][1].record()
    torch.cuda.synchronize()
    eager_ms = sum(a.elapsed_time(b) for a, b in eager_ev) / 3
    launches_per_step = ops.LAUNCHES // 3
    kern = {}
    for k, v in ops.PROFILE_EVENTS.items():
        ms_ = [s_.elapsed_time(e_) for s_, e_, _, _ in v]
        kern[k] = dict(ms=sum(ms_) / 3, n=len(v) / 3, flops=sum(f for _, _, f, _ in v) / 3,
                       bytes=sum(b for _, _, _, b in v) / 3)
    ops.PROFILE_EVENTS = None
    ops.SERIALIZE = False
    say("eager profiling pass done")
    # ---- device-resident timing (value)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    times = _time_step(step, flush, args.steps, None, False)
    barrier()
    # ---- end-to-end timing: host (pinned, float64) inputs -> device every step, loss read back
    for _ in range(2):
        step(pinned)
    _time_e2e(step, flush, 2, pinned)
    barrier()
    e2e_times = _time_e2e(step, flush, args.steps, pinned)
    barrier()
    clocks = sampler.stop()
    say("timing done")
    launches = launches_per_step * args.steps

    ms = sum(times) / len(times)
    ms_e2e = sum(e2e_times) / len(e2e_times)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    out = None
    if rank == 0:
        peaks = _peaks()
        extras = world == 1 and not args.no_extras
        tf32 = tf32_peak(dev) if extras else None
        table = {}
        for k, d in kern.items():
            tt = d["ms"] * 1e-3
            table[k] = {"ms_per_step": round(d["ms"], 4), "launches_per_step": round(d["n"], 1),
                        "tflops": round(d["flops"] / tt / 1e12, 2) if tt > 0 else None,
                        "gbs": round(d["bytes"] / tt / 1e9, 1) if tt > 0 else None}
        dom = max(kern, key=lambda k: kern[k]["ms"]) if kern else None
        roof = None
        if dom:
            d = kern[dom]
            ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
            roof = {"kernel": KERNEL_OF_OP.get(dom, "k_" + dom), "bound": "tensor", "achieved": round(ach, 3),
                    "peak": peaks["tensor_tflops"], "unit": "TFLOP/s", "frac": round(ach / peaks["tensor_tflops"], 5),
                    "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH.get(dom), "avg_launch_ms": round(d["ms"] / max(d["n"], 1), 5),
                    "share_of_step": round(d["ms"] / sum(v["ms"] for v in kern.values()), 4),
                    "eager_step_ms": round(eager_ms, 4), "peak_source": peaks["source"],
                    "note": "fp32-parity mode = error-compensated tensor-core products (3 x kind::tf32, or 6 x kind::f16 on bf16 "
                            "pieces in k_wgrad_ws): achieved = useful fp32 FLOPs of the op (the extra passes are not counted) / "
                            "CUDA-event time of the op in an eagerly launched, single-stream pass of the same step (every "
                            "kernel alone on the GPU); share_of_step = that time / the summed time of all instrumented ops "
                            "(`ops`); peak = cuBLAS bf16 sustained (MEASURED_PEAKS.json); the ceiling of an fp32-parity GEMM "
                            "is the measured TF32 rate / 3, see frac_of_3xtf32_ceiling"}
            if tf32:
                roof["tf32_peak_measured"] = tf32
                roof["frac_of_3xtf32_ceiling"] = round(ach / (tf32["sustained_tflops"] / 3.0), 4)
        out = {
            "metric": METRIC, "value": round(world * N / (ms * 1e-3), 1), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[wl]["label"].format(B=B), "nodes_per_gpu": N, "edges_per_gpu": E,
                       "global_batch": world * B, "parallelism": f"dp{world}",
                       "optimizer": "torch.optim.AdamW(lr=1e-4) as experiments/train.py:410 builds it, driven by msmp_adamw_run",
                       "l2": "flushed (256 MiB write) between timed steps"},
            "e2e": {"value": round(world * N / (ms_e2e * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms_e2e, 4),
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "ops": table,
            "execution": ("eager launches" if args.eager else
                          "whole step captured as one CUDA graph (GraphedTrainStep)" if world == 1 else
                          "two CUDA graphs per step (forward+loss+backward | optimizer) with ONE NCCL all-reduce of the flat "
                          "gradient bucket (it carries the local squared error) launched between them"),
        }
        if dp_check is not None:
            out["dp_check"] = dp_check
        if extras:
            del step, model, opt, resident
            torch.cuda.empty_cache()
            say("sub-records")
            subs = {}
            for name, b in (("c2", 64), ("c3", 64), ("c4", 8)):
                if name != wl:
                    subs[name.upper()] = _sub_config(name, b, dev, flush, 10)
            out["configs"] = subs
            out["bf16_mode"] = _sub_config(wl, B, dev, flush, 10, precision="bf16")
            out["bf16_mode"]["note"] = ("bf16 operands / fp32 accumulation in the GEMMs that have a bf16 variant (DESIGN.md "
                                        "section 8 lists them and the measured tolerance); `value` above is the fp32-parity mode")
            out["layer_c5"] = [_layer_c5(1 << 20, 6, "band", dev), _layer_c5(1 << 20, 6, "random", dev),
                               _layer_c5(1 << 20, 16, "random", dev)]
            try:        # the largest size of the sweep that leaves head room in 180 GB (per-edge tensors: 4 x 12.9 GB)
                out["layer_c5"].append(_layer_c5(1 << 22, 6, "random", dev, reps=3))
            except torch.OutOfMemoryError as exc:
                out["layer_c5"].append({"nodes": 1 << 22, "in_degree": 6, "error": str(exc)[:120]})
                torch.cuda.empty_cache()
            out["scatter_hbm"] = scatter_bandwidth(dev)
            out["message_hbm"] = message_bandwidth(dev)
            if not args.no_cpu_baseline:
                out["cpu_baseline"] = cpu_baseline(wl)
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


KERNEL_OF_OP = {"wgrad_ws": "k_wgrad_ws", "wgrad_tc": "k_wgrad_tc",
                "wgrad_ts": "k_wgrad_ts (LEM weight gradients, 3.3 Mi rows; edge dW2, 520 Ki rows; dW4)",
                "linear_tc": "k_linear_ts (k_linear_tc below 296 tiles)",
                "edge_ws_fwd": "k_edge_ws<fwd>", "edge_ws_bwd": "k_edge_ws<bwd>", "lem_tc_fwd": "k_lem_fwd_tc",
                "lem_tc_bwd": "k_lem_bwd_tc", "segment_reduce": "k_segment_reduce"}
# dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the launches of the op in one step) from the committed
# ncu pass over the serialised headline step: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
# --clock-control none --profile-from-start off python scripts/ncu_step.py c4 8` -> profiles/r2_launches_c4_step.csv
# (298 launches).  DRAM counters do not see operands that are still in the 126 MB L2 (the node GEMMs read what the previous
# launch wrote: 164 MB per launch against 4 M (K + N) = 184 MB of algorithmic bytes).
NCU_TRAFFIC_BYTES_PER_LAUNCH = {"linear_tc": 163.4e6, "lem_tc_fwd": 10.595e9, "lem_tc_bwd": 18.318e9, "wgrad_ts": 703.0e6,
                                "wgrad_ws": 222.8e6, "edge_ws_fwd": 490.6e6, "edge_ws_bwd": 1422.3e6, "segment_reduce": 316.8e6}


def scatter_bandwidth(dev, n_nodes=1 << 20, degree=6, reps=10):
    """Second half of the metric: standalone deterministic scatter-mean (msmp_segment_reduce) HBM GB/s on a 1 Mi-node,
    6-neighbour graph (config 5 shape): algorithmic bytes E*512 + N*512 + (N+1)*4 / CUDA-event time.
    Inputs (3.2 GB) exceed L2, so every repetition streams from HBM."""
    import torch
    from msmp_pde_b200 import ops
    E = n_nodes * degree
    src = torch.randn(E, 128, device=dev)
    ptr = (torch.arange(n_nodes + 1, device=dev, dtype=torch.int32) * degree).contiguous()
    inv = torch.full((n_nodes,), 1.0 / degree, device=dev)
    out = torch.empty(n_nodes, 128, device=dev)
    perm = torch.randperm(E, device=dev, dtype=torch.int32)            # gather-hostile order (by-source scatter)
    res = {}
    for name, pm in (("contiguous", None), ("permuted", perm)):
        for _ in range(3):
            ops.segment_reduce(src, ptr, perm=pm, scale=inv, out=out, N=n_nodes)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            ops.segment_reduce(src, ptr, perm=pm, scale=inv, out=out, N=n_nodes)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
        nbytes = E * 512 + n_nodes * 512 + (n_nodes + 1) * 4 + (E * 4 if pm is not None else 0)
        res[name] = {"ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1)}
    peak = _peaks()["hbm_gbs"]
    return {"kernel": "k_segment_reduce", "nodes": n_nodes, "edges": E, "bytes": E * 512 + n_nodes * 512 + (n_nodes + 1) * 4,
            "contiguous_gbs": res["contiguous"]["gbs"], "contiguous_frac": round(res["contiguous"]["gbs"] / peak, 4),
            "permuted_gbs": res["permuted"]["gbs"], "permuted_frac": round(res["permuted"]["gbs"] / peak, 4),
            "peak_gbs": peak}


def message_bandwidth(dev, n_nodes=1 << 20, degree=6, reps=7):
    """The message kernels (fused gather + second message-MLP layer on the tensor pipe + mean aggregation, forward and
    backward) on the config-5 shape, 1 Mi nodes x 6 Mi edges, as a fraction of the measured HBM copy peak, in COMPULSORY
    bytes: every distinct row is counted once however often it is gathered (the destination-sorted edge list re-uses
    P[dst] in-degree times; a gathered Q[src] row is also one of N distinct rows) --
        forward : N*1024 (P | Q rows) + E*512 (z2 written) + E*8 (indices) + N*516 (agg, rowptr)
        backward: N*1024 (P | Q) + N*512 (dagg) + E*512 (z2 read) + 3*E*512 (dz2, a1, dz1 written) + E*12 + N*512 (dP).
    `*_noreuse_frac` counts every gather as an HBM row (upper bound of the traffic: 1544 / 3596 B per edge).  'band' sources
    are neighbours on a ring, 'random' sources are uniform over the whole graph (gather-hostile)."""
    import torch
    from msmp_pde_b200 import ops, synth
    from msmp_pde_b200.graph import build_topology
    gen = torch.Generator(device=dev).manual_seed(0)
    PQ = torch.randn(n_nodes, 256, device=dev, generator=gen)
    W2 = (torch.randn(128, 128, device=dev, generator=gen) / 11).contiguous()
    b2 = torch.randn(128, device=dev, generator=gen) * 0.1
    dagg = torch.randn(n_nodes, 128, device=dev, generator=gen)
    dP = torch.empty(n_nodes, 128, device=dev)
    peak = _peaks()["hbm_gbs"]
    res = {"kernel": "k_edge_ws<fwd> / k_edge_ws<bwd>", "nodes": n_nodes, "edges": n_nodes * degree, "peak_gbs": peak,
           "timing": f"CUDA events around the op (kernel + carry fix-up), median of {reps} after 3 warm-ups"}
    for topo_name, npg in (("band", 100), ("random", 0)):
        g = synth.large_graph(n_nodes, degree, topology=topo_name, nodes_per_graph=npg, seed=0)
        topo = build_topology(g["edge_index"].to(dev), g["batch"].to(dev), n_nodes)
        del g
        E = topo.E
        tfs, tbs = [], []
        for it in range(3 + reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            agg, z2 = ops.edge_fwd(PQ[:, :128], PQ[:, 128:], topo, None, b2, W2raw=W2)
            ev[1].record()
            out = ops.edge_bwd(PQ[:, :128], PQ[:, 128:], topo, None, z2, dagg, dP, defer_wgrad=True, W2raw=W2)
            ev[2].record()
            torch.cuda.synchronize()
            del out
            if it >= 3:       # the first passes pay for the allocator's cudaMallocs of the 3.2 GB edge tensors
                tfs.append(ev[0].elapsed_time(ev[1]))
                tbs.append(ev[1].elapsed_time(ev[2]))
        tf, tb = sorted(tfs)[len(tfs) // 2], sorted(tbs)[len(tbs) // 2]       # median of `reps`
        cf = n_nodes * 1024 + E * 520 + n_nodes * 516
        cb = n_nodes * 1024 + n_nodes * 512 + E * 512 + 3 * E * 512 + E * 12 + n_nodes * 512
        nf, nb_ = E * 1544 + n_nodes * 512, E * 3596 + n_nodes * 512
        res[topo_name] = {"fwd_ms": round(tf, 4), "fwd_compulsory_gbs": round(cf / tf / 1e6, 1),
                          "fwd_compulsory_frac": round(cf / tf / 1e6 / peak, 4),
                          "fwd_noreuse_frac": round(nf / tf / 1e6 / peak, 4),
                          "bwd_ms": round(tb, 4), "bwd_compulsory_gbs": round(cb / tb / 1e6, 1),
                          "bwd_compulsory_frac": round(cb / tb / 1e6 / peak, 4),
                          "bwd_noreuse_frac": round(nb_ / tb / 1e6 / peak, 4),
                          "fwd_executed_tf32_tflops": round(3 * 2 * E * 128 * 128 / tf / 1e9, 1)}
        del topo, agg, z2
    return res


# ------------------------------------------------------------------------------------------ reference
def _oracle_step_time(workload, B, steps, warmup, budget_s=None):
    """The reference's CPU path (oracle port, float64 = reference-native dtype) on all host cores.  Never loads the CUDA
    library: `synth` is imported without the package __init__ (see _synth_module)."""
    import torch
    synth = _synth_module(load_native=False)
    from oracle import models as om
    torch.set_num_threads(os.cpu_count())
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        pde, data, meta = _make(synth, workload, B, 0, dtype=torch.float64)
        torch.manual_seed(0)
        model = om.MP_PDE_Solver2DLEMLinGated(pde, TW, 128, 6, meta["eq_variables"])
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        ts = []
        t_start = time.perf_counter()
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            opt.zero_grad()
            loss = _loss(model(data), data.y)
            loss.backward()
            opt.step()
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
            if budget_s is not None and len(ts) >= 1 and time.perf_counter() - t_start > budget_s:
                break
    finally:
        torch.set_default_dtype(prev)
    return sum(ts) / len(ts), data.x.shape[0], len(ts)


def cpu_baseline(workload="c4"):
    """Bounded sample for the N = 1 line (~20 s of CPU work): two steps of the reference's CPU path after one warm-up step,
    on 2 of the 8 lattice graphs of the C4 batch (`--impl reference` times the full batch)."""
    B = 2 if workload == "c4" else 16
    sec, n, k = _oracle_step_time(workload, B, 2, 1)
    return {"value": round(n / sec, 1), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"{B} of the {WORKLOADS[workload]['graphs']} graphs of the {workload.upper()} batch ({n} nodes), {k} steps "
                      f"after 1 warm-up, float64 (reference-native), torch CPU with {os.cpu_count()} threads",
            "ms_per_step": round(sec * 1e3, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    full = args.graphs_per_gpu or WORKLOADS[wl]["graphs"]
    B = min(REF_GRAPHS[wl], full)
    # each step is a bounded sample (B whole graphs -- the path's independent units -- of the `full`-graph batch); the
    # run stops adding timed steps once ~150 s are spent so that --steps K --warmup W stays within minutes
    warmup = min(max(args.warmup, 1), 1)
    sec, n, k = _oracle_step_time(wl, B, max(args.steps, 1), warmup, budget_s=150.0)
    v = round(n / sec, 1)
    sample = (f"{B} of {full} graphs per step ({n} nodes), {k} timed steps after {warmup} warm-up, float64 "
              f"(reference-native), torch CPU, {os.cpu_count()} threads")
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
           "steps": k, "warmup": warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOADS[wl]["label"].format(B=full), "sample": sample,
                      "implementation": "oracle port of the reference CPU path (the reference itself is not importable: "
                                        "torch_geometric, torch_scatter, torch_cluster, lem_cuda absent)"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "native_library_loaded": any("libmsmp_b200" in l for l in open("/proc/self/maps"))}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--graphs-per-gpu", type=int, default=0)
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records (other configs, bf16 mode, C5, HBM lines)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the MSMP-PDE hot path (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): MSMP-PDE2D training step on the RP shape -- 64 trajectories per GPU,
100 nodes each, time_window 25, 6-neighbour radius graph (N = 6400 nodes, E = 37632 edges per GPU),
MP_PDE_Solver2DLEMLinGated (LEM encoder + 6 gated layer pairs + Conv1d decoder, 1 409 002 parameters),
synthetic data, random-init weights.  One step = forward + loss (sqrt of the batch-global summed squared
error, train_helper.py:126,138) + backward + (N > 1: gradient all-reduce over NCCL) + AdamW update.
Metric: graph-nodes per second, whole job.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GNN fwd+bwd graph-nodes/sec"
UNIT = "nodes/s"
B_PER_GPU, NX, TW = 64, 100, 25


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


def _workload(rank: int):
    import torch
    from msmp_pde_b200 import synth
    pde, data, meta = synth.config_c2(B=B_PER_GPU, nx=NX, tw=TW, seed=rank)     # float64 host tensors (F1)
    return pde, data, meta


def _loss(pred, y):
    import torch
    return torch.sqrt(torch.nn.functional.mse_loss(pred, y, reduction="sum"))


# ----------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    from msmp_pde_b200 import models_gnn2D, ops
    from msmp_pde_b200.train_step import GraphedTrainStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    def say(msg):
        if args.verbose:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)
    saved_stdout = None
    if world > 1:
        # CUDA-graph capture of NCCL collectives: the watchdog thread must not touch the capturing context
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
    pde, data, meta = _workload(rank)
    torch.manual_seed(0)
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, TW, 128, 6, meta["eq_variables"]).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True, capturable=True)      # train.py:410
    N, E = data.x.shape[0], data.edge_index.shape[1]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    resident = data.clone().to(dev)
    pinned = data.clone().apply(lambda t: t.pin_memory())
    h2d = sum(t.numel() * t.element_size() for t in (getattr(pinned, k) for k in pinned.keys())
              if torch.is_tensor(t) and t.is_floating_point())
    # public API: the whole training step (fwd + loss + bwd + DP all-reduce + AdamW) as one CUDA graph
    say("model built")
    step = GraphedTrainStep(model, opt, resident, warmup=max(3, args.warmup), use_graph=not args.eager)
    say("step captured")

    def timed(K, graph, read_loss):
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(None)
    barrier()
    # ---- per-kernel events on the edge ops (roofline) need eager launches: one short eager pass, not timed as value
    ops.PROFILE_EVENTS = {}
    ops.LAUNCHES = 0
    eager_ev = []
    for _ in range(3):
        flush.zero_()
        eager_ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
        eager_ev[-1][0].record()
        step._eager_step()
        eager_ev[-1][1].record()
    torch.cuda.synchronize()
    eager_ms = sum(a.elapsed_time(b) for a, b in eager_ev) / 3        # one eagerly launched step, same pass as the op events
    launches_per_step = ops.LAUNCHES // 3
    kern = {}
    for k, v in ops.PROFILE_EVENTS.items():
        ms_ = [s_.elapsed_time(e_) for s_, e_, _, _ in v]
        kern[k] = dict(ms=sum(ms_) / 3, n=len(v) / 3, flops=sum(f for _, _, f, _ in v) / 3,
                       bytes=sum(b for _, _, _, b in v) / 3)
    ops.PROFILE_EVENTS = None
    say("eager profiling pass done")
    # ---- device-resident timing (value)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    times = timed(args.steps, None, False)
    barrier()
    # ---- end-to-end timing: host (pinned, float64) inputs -> device every step, loss read back
    for _ in range(2):
        step(pinned)
    barrier()
    e2e_times = timed(args.steps, pinned, True)
    barrier()
    clocks = sampler.stop()
    say("timing done")
    launches = launches_per_step * args.steps

    scatter = scatter_bandwidth(dev) if rank == 0 else None
    message = message_bandwidth(dev) if rank == 0 else None
    ms = sum(times) / len(times)
    ms_e2e = sum(e2e_times) / len(e2e_times)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    out = None
    if rank == 0:
        peaks = _peaks()
        # per-op table (CUDA events around each C-ABI op in a short eager pass) and the dominant op's roofline
        table = {}
        for k, d in kern.items():
            t = d["ms"] * 1e-3
            table[k] = {"ms_per_step": round(d["ms"], 4), "launches_per_step": round(d["n"], 1),
                        "tflops": round(d["flops"] / t / 1e12, 2) if t > 0 else None,
                        "gbs": round(d["bytes"] / t / 1e9, 1) if t > 0 else None}
        dom = max(kern, key=lambda k: kern[k]["ms"]) if kern else None
        roof = None
        if dom:
            d = kern[dom]
            ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
            roof = {"kernel": "k_" + dom, "bound": "tensor", "achieved": round(ach, 3),
                    "peak": peaks["tensor_tflops"], "unit": "TFLOP/s", "frac": round(ach / peaks["tensor_tflops"], 5),
                    "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH.get(dom), "avg_launch_ms": round(d["ms"] / max(d["n"], 1), 5),
                    "share_of_step": round(d["ms"] / sum(v["ms"] for v in kern.values()), 4),
                    "eager_step_ms": round(eager_ms, 4),
                    "peak_source": peaks["source"],
                    "note": "tcgen05 kind::tf32, error-compensated 3xTF32 (fp32 parity): achieved = useful fp32 FLOPs "
                            "of the op (the 3x MMA passes are not counted) / CUDA-event time of the op (kernel + its fixed-order "
                            "split reduction) in an eagerly launched step; share_of_step = that time / the summed device "
                            "time of all instrumented ops of the same eager step (`ops`; they cover ~90 % of the step's "
                            "device time; the eager step itself is host-launch bound, eager_step_ms, and the timed value "
                            "is a CUDA-graph replay in which ops overlap on several streams); "
                            "peak = cuBLAS bf16 sustained, i.e. 6x the effective ceiling of 3xTF32"}
        out = {
            "metric": METRIC, "value": round(world * N / (ms * 1e-3), 1), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: MSMP-PDE2D (MP_PDE_Solver2DLEMLinGated) RP shape, 64 graphs x 100 nodes per GPU, "
                                   "tw=25, 588 edges/graph, fwd+loss+bwd+AdamW",
                       "nodes_per_gpu": N, "edges_per_gpu": E, "global_batch": world * B_PER_GPU,
                       "parallelism": f"dp{world}", "l2": "flushed (256 MiB write) between timed steps"},
            "e2e": {"value": round(world * N / (ms_e2e * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms_e2e, 4),
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "ops": table, "scatter_hbm": scatter, "message_hbm": message,
            "execution": ("eager launches" if args.eager else
                          "whole step captured as one CUDA graph (GraphedTrainStep)" if world == 1 else
                          "three CUDA graphs per step (forward+loss | backward | AdamW) with the NCCL all-reduces of the "
                          "loss terms and of the flat gradient bucket launched eagerly between them"),
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(sample_graphs=8, steps=2)
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


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


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the C2 shapes
# (profiles/r1_final_kernels_ncu_raw.csv): the step's 53 weight-gradient launches are 36 node-level (9.9 MB), 12 edge-level
# (38.6 MB), the two LEM ones (355 MB, 190 MB) and 3 small ones -> 25.7 MB on average, against 4 * M * (K + N) bytes
# of operands (the algorithmic traffic: every operand element is read once).
NCU_TRAFFIC_BYTES_PER_LAUNCH = {"wgrad_tc": 25.7e6}


def message_bandwidth(dev, n_nodes=1 << 20, degree=6, reps=7):
    """The message kernels (fused gather + second message-MLP layer on the tensor pipe + mean aggregation, forward and
    backward) on the config-5 shape, 1 Mi nodes x 6 Mi edges, as a fraction of the measured HBM copy peak.
    Algorithmic bytes per edge (gathers counted without reuse): forward 2 x 512 gathered (P[dst],
    Q[src]) + 512 written (z2) + 8 index = 1544; backward 4 x 512 read (dagg[dst], z2, P[dst], Q[src]) + 3 x 512 written
    (dz2, a1, dz1) + 12 index/scale = 3596; plus 512 per node for the aggregated output.  'band' sources are neighbours
    on a ring (gathers mostly hit L2), 'random' sources are uniform over the whole graph (every gather is an HBM row)."""
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
           "timing": f"CUDA events around the op (kernel + memset + carry fix-up), median of {reps} after 3 warm-ups"}
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
        bf, bb = E * 1544 + n_nodes * 512, E * 3596 + n_nodes * 512
        res[topo_name] = {"fwd_ms": round(tf, 4), "fwd_gbs": round(bf / tf / 1e6, 1), "fwd_frac": round(bf / tf / 1e6 / peak, 4),
                          "bwd_ms": round(tb, 4), "bwd_gbs": round(bb / tb / 1e6, 1), "bwd_frac": round(bb / tb / 1e6 / peak, 4),
                          "fwd_executed_tf32_tflops": round(3 * 2 * E * 128 * 128 / tf / 1e9, 1)}
        del topo, agg, z2
    return res


# ------------------------------------------------------------------------------------------ reference
def _oracle_step_time(B, steps, warmup, dtype_name="float64"):
    """The reference's CPU path (oracle port, float64 = reference-native dtype) on all host cores."""
    import torch
    from msmp_pde_b200 import synth
    from oracle import models as om
    torch.set_num_threads(os.cpu_count())
    prev = torch.get_default_dtype()
    torch.set_default_dtype(getattr(torch, dtype_name))
    try:
        pde, data, meta = synth.config_c2(B=B, nx=NX, tw=TW, seed=0, dtype=getattr(torch, dtype_name))
        torch.manual_seed(0)
        model = om.MP_PDE_Solver2DLEMLinGated(pde, TW, 128, 6, meta["eq_variables"])
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        ts = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            opt.zero_grad()
            loss = _loss(model(data), data.y)
            loss.backward()
            opt.step()
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    finally:
        torch.set_default_dtype(prev)
    return sum(ts) / len(ts), data.x.shape[0]


def cpu_baseline(sample_graphs=8, steps=2):
    sec, n = _oracle_step_time(sample_graphs, steps, 1)
    return {"value": round(n / sec, 1), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"{sample_graphs} of the 64 graphs of the C2 batch ({n} nodes), {steps} steps after 1 warm-up, "
                      f"float64 (reference-native), torch CPU with {os.cpu_count()} threads", "ms_per_step": round(sec * 1e3, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 16          # bounded sample of the 64-graph batch: keeps --steps K --warmup W within minutes
    steps, warmup = min(args.steps, 5), min(max(args.warmup, 1), 2)
    sec, n = _oracle_step_time(B, steps, warmup)
    v = round(n / sec, 1)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
           "steps": steps, "warmup": warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "C2: MSMP-PDE2D RP shape (bounded sample: 16 of 64 graphs x 100 nodes), fwd+loss+bwd+AdamW, "
                                  "oracle port of the reference CPU path (reference not importable: torch_geometric, "
                                  "torch_scatter, torch_cluster, lem_cuda absent)"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                            "sample": f"16 of 64 graphs ({n} nodes) x {steps} steps, float64, {os.cpu_count()} threads"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
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

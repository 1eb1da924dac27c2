"""Training-step time (forward + loss + backward + AdamW, CUDA-graph replay) of the other BASELINE configs on one GPU:
C3 (RPU, kNN graph on the pseudo-random grid) and C4 (MSMP-PDE2D on 128 x 128 lattices, b graphs per GPU).
Supplementary to bench.py (whose metric is quoted on C2)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import models_gnn2D, synth
from msmp_pde_b200.train_step import GraphedTrainStep

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(name, pde, data, meta, steps=10):
    torch.manual_seed(0)
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True, capturable=True)
    g = data.clone().to(dev)
    step = GraphedTrainStep(model, opt, g, warmup=3)
    for _ in range(3):
        step(g)
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(g)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    N, E = g.x.shape[0], g.edge_index.shape[1]
    print(json.dumps({"config": name, "nodes": N, "edges": E, "ms_per_step": round(ms, 3),
                      "nodes_per_s": round(N / ms * 1e3)}), flush=True)
    del step, model, opt, g
    torch.cuda.empty_cache()


if __name__ == "__main__":
    run("C3 RPU kNN k=3, 64 graphs x 100 nodes", *synth.config_c3(B=64, seed=0))
    for b in (1, 2, 4, 8):
        run(f"C4 MSWG3 128x128 4-neighbour lattice, {b} graphs/GPU", *synth.config_c4(B=b, side=128, seed=0))

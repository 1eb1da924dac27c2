"""Kernel-time summary (torch.profiler / CUPTI) of one eagerly launched training step of a BASELINE config, plus the
CUDA-graph replay time of the same step.  Diagnostic only.   python scripts/prof_cfg.py c4 8"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from msmp_pde_b200 import models_gnn2D, synth
from msmp_pde_b200.train_step import GraphedTrainStep

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
pde, data, meta = {"c2": lambda: synth.config_c2(B=B, seed=0), "c3": lambda: synth.config_c3(B=B, seed=0),
                   "c4": lambda: synth.config_c4(B=B, side=128, seed=0)}[cfg]()
torch.manual_seed(0)
model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
g = data.clone().to(dev)
step = GraphedTrainStep(model, opt, g, warmup=3, use_graph=True)
for _ in range(3):
    step.eager()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step.eager()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0)
    if t > 0:
        rows.append((t / 2, e.count / 2, e.key[:100]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"{cfg} B={B} N={g.x.shape[0]} E={g.edge_index.shape[1]}: total kernel time per step {tot/1e3:.3f} ms over "
      f"{sum(r[1] for r in rows):.0f} launches")
for t, c, k in rows[:40]:
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  n={c:6.1f}  avg={t/c:7.1f}  {k}")
ts = []
for _ in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(g); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print("graph replay ms/step:", sorted(ts)[len(ts) // 2])

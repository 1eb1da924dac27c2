"""Kernel-time summary of one eager training step (torch.profiler / CUPTI).  Diagnostic only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from msmp_pde_b200 import models_gnn2D, synth
from msmp_pde_b200.train_step import GraphedTrainStep

dev = torch.device("cuda:0")
pde, data, meta = synth.config_c2(B=64, nx=100, seed=0)
torch.manual_seed(0)
model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True, capturable=True)
step = GraphedTrainStep(model, opt, data.clone().to(dev), warmup=3, use_graph=False)
for _ in range(3):
    step._eager_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step._eager_step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 or getattr(e, "cuda_time_total", 0) > 0]
rows = []
for e in prof.key_averages():
    t = getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0)
    if t > 0:
        rows.append((t / 2, e.count / 2, e.key[:90]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total kernel time per step: {tot/1e3:.3f} ms over {sum(r[1] for r in rows):.0f} launches")
for t, c, k in rows[:45]:
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  n={c:6.1f}  avg={t/c:7.1f}  {k}")

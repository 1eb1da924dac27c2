"""One eagerly launched, single-stream training step of the headline workload (C4, 8 lattices) inside a cudaProfiler range:
the command ncu wraps for the launch list and the `--set full` captures (profiles/r2_*).  python scripts/ncu_step.py [c4|c2] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import models_gnn2D, ops, synth
from msmp_pde_b200.train_step import GraphedTrainStep

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
pde, data, meta = synth.config_c4(B=B, side=128, seed=0) if cfg == "c4" else synth.config_c2(B=B, seed=0)
torch.manual_seed(0)
model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
g = data.clone().to(dev)
ops.SERIALIZE = True
step = GraphedTrainStep(model, opt, g, warmup=2, use_graph=False)
step.eager()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step.eager()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(step.loss))

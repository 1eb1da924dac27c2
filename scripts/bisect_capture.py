import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
from msmp_pde_b200 import models_gnn2D, synth, ops
from msmp_pde_b200.layers import _LayerBase
from msmp_pde_b200.lem import LEMcuda
from msmp_pde_b200.train_step import GraphedTrainStep
dev = torch.device("cuda:0")
pde, data, meta = synth.config_c2(B=4, nx=100, seed=0)
for variant in sys.argv[1:] or ["all", "nolayer", "nolem", "nolinear", "nodecoder", "none"]:
    torch.manual_seed(0)
    model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True, capturable=True)
    g = data.clone().to(dev)
    step = GraphedTrainStep(model, opt, g, warmup=2, use_graph=False)
    for m in model.modules():
        kill = (variant == "nolayer" and isinstance(m, _LayerBase)) or (variant == "nolem" and isinstance(m, LEMcuda)) or \
               (variant == "nolinear" and isinstance(m, nn.Linear)) or (variant == "nodecoder" and isinstance(m, nn.Sequential)) or variant == "none"
        if kill:
            m.__dict__.pop("_msmp_gsink", None)
    torch.cuda.synchronize()
    try:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            step._eager_step()
        gr.replay(); torch.cuda.synchronize()
        print(variant, "OK")
    except Exception as e:
        print(variant, "FAIL", str(e).splitlines()[0])
        break

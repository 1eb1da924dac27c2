"""Diagnostic: per-parameter gradient error of the CUDA model vs the fp64 oracle, next to the error of an
honest fp32 torch run of the oracle itself (the fp32 floor).  Not part of the product."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from msmp_pde_b200 import models_gnn, models_gnn2D, synth
from oracle import models as om
from tests.util import formula_weights_

dev = torch.device("cuda:0")

def run(model, data):
    model.zero_grad()
    out = model(data)
    torch.sqrt(((out - data.y) ** 2).sum()).backward()
    return out.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters()}

def cmp(tag, g, gref):
    gscale = max(float(v.abs().max()) for v in gref.values())
    rows = []
    for n in gref:
        r = gref[n].double().cpu(); a = g[n].double().cpu()
        den = float(r.abs().max()); 
        e_own = float((a - r).abs().max()) / max(den, 1e-300)
        e_glob = float((a - r).abs().max()) / gscale
        rows.append((e_own, e_glob, den / gscale, n))
    rows.sort(reverse=True)
    print(f"--- {tag}: worst 6 by own-scale error (err/own max, err/global max, own max/global max)")
    for r in rows[:6]:
        print("   %.2e  %.2e  %.2e  %s" % r)
    print("   worst global-scale err: %.2e" % max(r[1] for r in rows))

for weights in ("formula", "random"):
    for cfg, mod, cls in ((synth.config_c1, models_gnn, "MP_PDE_Solver"), (synth.config_c2, models_gnn2D, "MP_PDE_Solver2DLEMLinGated")):
        pde, data, meta = cfg(B=3, nx=40, seed=4)
        torch.manual_seed(0)
        torch.set_default_dtype(torch.float64)
        ref = getattr(om, cls)(pde, 25, 128, 6, meta["eq_variables"])
        if weights == "formula":
            formula_weights_(ref)
        out64, g64 = run(ref, data)
        torch.set_default_dtype(torch.float32)
        ref32 = getattr(om, cls)(pde, 25, 128, 6, meta["eq_variables"]).float()
        ref32.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
        d32 = data.clone().apply(lambda t: t.float() if t.is_floating_point() else t)
        ref32 = ref32.to(dev); d32g = d32.clone().to(dev)
        out32, g32 = run(ref32, d32g)
        model = getattr(mod, cls)(pde, 25, 128, 6, meta["eq_variables"])
        model.load_state_dict(ref.state_dict())
        model = model.to(dev)
        dd = data.clone().to(dev)
        outc, gc = run(model, dd)
        print(f"===== {cls} weights={weights}: out err cuda {float((outc.cpu()-out64).abs().max()/out64.abs().max()):.2e}  torch-fp32 {float((out32.cpu().double()-out64).abs().max()/out64.abs().max()):.2e}")
        cmp("msmp cuda vs fp64", gc, g64)
        cmp("torch fp32 (gpu) vs fp64", g32, g64)
        # determinism per parameter
        _, gc2 = run(model, dd)
        bad = [n for n in gc if not torch.equal(gc[n], gc2[n])]
        print("   non-deterministic params:", bad[:8], len(bad))

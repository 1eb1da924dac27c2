"""Diagnostic: worst gradient-allowance excess of the CUDA variant models next to an honest fp32 torch run of the
oracle variant (the fp32 floor) on the golden inputs.  Not part of the product."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from msmp_pde_b200 import models_gnn, models_gnn2D
from oracle import variants as ov
from tests import golden_io
from tests.util import formula_weights_
from tests.test_oracle_golden import VARIANTS_1F, VARIANTS_2F, variant_eq

dev = torch.device("cuda:0")
names = sys.argv[1:] or (VARIANTS_1F + VARIANTS_2F)


def run(model, data):
    model.zero_grad()
    out = model(data)
    torch.sqrt(torch.nn.functional.mse_loss(out, data.y, reduction="sum")).backward()
    return out.detach(), {n: (p.grad if p.grad is not None else torch.zeros_like(p)).detach().double().cpu()
                          for n, p in model.named_parameters()}


def excess(g, gref):
    gscale = max(float(v.abs().max()) for v in gref.values())
    rows = sorted(((float((g[n] - gref[n]).abs().max()) / (1e-5 * float(gref[n].abs().max()) + 2e-6 * gscale), n)
                   for n in gref), reverse=True)
    return rows[:3]


for name in names:
    g = golden_io.load(f"var_{name}.npz")
    pde_name, eq = variant_eq(name)
    pde, data = golden_io.model_inputs(g, pde_name)
    torch.set_default_dtype(torch.float64)
    ref = getattr(ov, name)(pde, 25, 128, 6, eq)
    formula_weights_(ref)
    out64, g64 = run(ref, data)
    torch.set_default_dtype(torch.float32)
    ref32 = getattr(ov, name)(pde, 25, 128, 6, eq)
    ref32.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    ref32 = ref32.to(dev)
    d32 = data.clone().apply(lambda t: t.float() if t.is_floating_point() else t).to(dev)
    out32, g32 = run(ref32, d32)
    torch.set_default_dtype(torch.float64)
    model = getattr(models_gnn if name in VARIANTS_1F else models_gnn2D, name)(pde, 25, 128, 6, eq)
    model.load_state_dict(ref.state_dict())
    model = model.to(dev)
    outc, gc = run(model, copy.copy(data).clone().to(dev))
    den = float(out64.abs().max())
    print(f"== {name}: out err cuda {float((outc.cpu() - out64).abs().max()) / den:.2e}  torch-fp32 "
          f"{float((out32.cpu().double() - out64).abs().max()) / den:.2e}")
    print("   cuda  excess:", [(round(e, 2), n) for e, n in excess(gc, g64)])
    print("   fp32  excess:", [(round(e, 2), n) for e, n in excess(g32, g64)])

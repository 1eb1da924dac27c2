"""Diagnostic: one large edge forward + backward; with a -DMSMP_EW_TICKS build (MSMP_B200_LIB=...) CTA 0 prints the
cycles each role of k_edge_ws spent waiting.  Also times the two ops with CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import ops, synth
from msmp_pde_b200.graph import build_topology

n, deg = int(os.environ.get("N", 1 << 20)), int(os.environ.get("DEG", 6))
dev = torch.device("cuda:0")
g = synth.large_graph(n, deg, topology=os.environ.get("TOPO", "band"), nodes_per_graph=100, seed=0)
topo = build_topology(g["edge_index"].to(dev), g["batch"].to(dev), n)
torch.manual_seed(0)
PQ = torch.randn(n, 256, device=dev)
W2 = (torch.randn(128, 128, device=dev) / 11).contiguous()
b2 = torch.randn(128, device=dev) * 0.1
dagg = torch.randn(n, 128, device=dev)
dP = torch.empty(n, 128, device=dev)
for it in range(int(os.environ.get("REPS", 2))):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    agg, z2 = ops.edge_fwd(PQ[:, :128], PQ[:, 128:], topo, None, b2, W2raw=W2)
    ev[1].record()
    out = ops.edge_bwd(PQ[:, :128], PQ[:, 128:], topo, None, z2, dagg, dP, defer_wgrad=True, W2raw=W2)
    ev[2].record()
    torch.cuda.synchronize()
    E = topo.E
    f, b = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print(f"iter {it}: fwd {f:.3f} ms ({E * 1544 / f / 1e6:.0f} GB/s algorithmic)  bwd {b:.3f} ms ({E * 3588 / b / 1e6:.0f} GB/s)",
          flush=True)

"""Edge forward / backward op time vs graph size, warp-specialised (edge_ws.cu) against single-role (edge_tc.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import ops, synth
from msmp_pde_b200.graph import build_topology

dev = torch.device("cuda:0")
torch.manual_seed(0)
W2 = (torch.randn(128, 128, device=dev) / 11).contiguous()
W2t = W2.t().contiguous()
b2 = torch.randn(128, device=dev) * 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in (6400, 16384, 32768, 65536, 131072, 262144, 1 << 20):
    g = synth.large_graph(n, 6, topology="band", nodes_per_graph=100, seed=0)
    topo = build_topology(g["edge_index"].to(dev), g["batch"].to(dev), n)
    PQ = torch.randn(n, 256, device=dev)
    dagg = torch.randn(n, 128, device=dev)
    dP = torch.empty(n, 128, device=dev)
    row = [f"N={n} E={topo.E} tiles/SM={topo.E / 128 / 148:.1f}"]
    for ws in (True, False):
        ops.EDGE_WS = ws
        best = [1e9, 1e9]
        for it in range(8):
            flush.zero_()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            agg, z2 = ops.edge_fwd(PQ[:, :128], PQ[:, 128:], topo, W2t, b2)
            ev[1].record()
            ops.edge_bwd(PQ[:, :128], PQ[:, 128:], topo, W2, z2, dagg, dP, defer_wgrad=True)
            ev[2].record()
            torch.cuda.synchronize()
            best = [min(best[0], ev[0].elapsed_time(ev[1])), min(best[1], ev[1].elapsed_time(ev[2]))]
        row.append(f"{'ws' if ws else 'tc'}: fwd {best[0] * 1e3:.1f} us bwd {best[1] * 1e3:.1f} us")
    print(" | ".join(row), flush=True)

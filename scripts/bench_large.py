"""Config 5: one GNN_Layer(128,128,128,25,1) forward+backward on a large synthetic graph; per-op CUDA-event timings."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import layers, ops, synth

def run(n, deg, topology, npg=100, reps=5):
    dev = torch.device("cuda:0")
    g = synth.large_graph(n, deg, topology=topology, nodes_per_graph=npg, seed=0)
    t = {k: v.to(dev) for k, v in g.items()}
    torch.manual_seed(0)
    layer = layers.GNN_Layer(128, 128, 128, 25, 1).to(dev)
    x = t["x"].clone().requires_grad_(True)
    def step():
        out = layer(x, t["u"], t["pos"], t["variables"], t["edge_index"], t["batch"])
        out.backward(out.detach())
    for _ in range(3): step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): step()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    ops.PROFILE_EVENTS = {}
    for _ in range(3): step()
    torch.cuda.synchronize()
    table = {}
    for k, v in ops.PROFILE_EVENTS.items():
        tt = sum(a.elapsed_time(b) for a, b, _, _ in v) / 3
        table[k] = dict(ms=round(tt, 3), tflops=round(sum(f for *_, f, _ in v) / 3 / tt / 1e9, 1),
                        gbs=round(sum(b for *_, b in v) / 3 / tt / 1e6, 0))
    ops.PROFILE_EVENTS = None
    E = n * deg
    alg = 3 * (E * (2 * 283 * 128 + 2 * 128 * 128) + n * (2 * 257 * 128 + 2 * 128 * 128))     # SURVEY 8d, fwd+bwd = 3x
    return dict(nodes=n, degree=deg, topology=topology, ms_per_fwd_bwd=round(ms, 3), nodes_per_s=round(n / ms * 1e3),
                algorithmic_tflops=round(alg / ms / 1e9, 1), ops=table)

if __name__ == "__main__":
    for n, deg, topo in ((1 << 20, 6, "band"), (1 << 20, 6, "random"), (1 << 20, 16, "random")):
        print(json.dumps(run(n, deg, topo)))

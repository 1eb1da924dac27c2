"""Diagnostic: shapes, strides and CUDA-event time of every ops.linear_wgrad call of one GNN_Layer backward on 1 Mi nodes x 6 Mi edges."""
import sys, json, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scripts')
from msmp_pde_b200 import ops, layers, synth
ops.SERIALIZE = True
orig = ops.linear_wgrad
log = []
def wrapped(X, dY, **kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = orig(X, dY, **kw); b.record(); torch.cuda.synchronize()
    X1 = kw.get("X1")
    log.append((tuple(X.shape), X.stride(), tuple(dY.shape), dY.stride(), None if X1 is None else (tuple(X1.shape), X1.stride()), kw.get("xswish"), round(a.elapsed_time(b), 3)))
    return r
ops.linear_wgrad = wrapped
layers.ops.linear_wgrad = wrapped
dev = torch.device("cuda:0")
g = synth.large_graph(1 << 20, 6, topology="band", nodes_per_graph=100, seed=0)
t = {k: v.to(dev) for k, v in g.items()}
torch.manual_seed(0)
layer = layers.GNN_Layer(128, 128, 128, 25, 1).to(dev)
x = t["x"].clone().requires_grad_(True)
for it in range(3):
    log.clear()
    out = layer(x, t["u"], t["pos"], t["variables"], t["edge_index"], t["batch"])
    out.backward(out.detach())
    torch.cuda.synchronize()
for l in log: print(l)

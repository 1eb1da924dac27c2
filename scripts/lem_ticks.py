import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import ops, _lib
from msmp_pde_b200.lem import LEMcuda
dev = torch.device("cuda:0")
ops.LEM_PERSISTENT = True
rnn = LEMcuda(6, 128, 1.0).to(dev)
NN = int(os.environ.get('NN', 6400))
x = torch.randn(25, NN, 6, device=dev)
for _ in range(3):
    ys, zs = rnn(x, last_only=True)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
_lib.lib.msmp_lem_debug_ticks.argtypes = [ctypes.c_void_p]
_lib.lib.msmp_lem_debug_ticks(buf)
t = list(buf)
seq = [(0, "start"), (1, "accumulator pre-init (overlaps the G GEMM)"), (2, "rest of the G GEMM wait"),
       (3, "gate_z epilogue + publish"), (4, "L GEMM wait"), (5, "gate_y epilogue + publish")]
for (i0, _), (i1, name) in zip(seq[:-1], seq[1:]):
    print(f"{name:45s} {t[i1]-t[i0]:8d} cycles")
print("step total", t[5] - t[0])
(ys.sum() + zs.sum()).backward()
torch.cuda.synchronize()
_lib.lib.msmp_lem_debug_ticks(buf)
t = list(buf)
seq = [(32, "start"), (33, "bwd_y epilogue + publish"), (34, "acc1 GEMM wait"), (35, "bwd_z epilogue + publish"),
       (36, "dG1 GEMM wait (acc_mid)"), (37, "restage dG0 + publish"), (38, "dG2 / dG0 GEMM wait"), (39, "dy += acc2")]
print("backward:")
for (i0, _), (i1, name) in zip(seq[:-1], seq[1:]):
    print(f"{name:45s} {t[i1]-t[i0]:8d} cycles")
print("step total", t[39] - t[32])

print("bwd_y batches (cycles since bwd_y start):", [t[40 + i] - t[32] for i in range(4)])

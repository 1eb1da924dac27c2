import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import ops, _lib
from msmp_pde_b200.lem import LEMcuda
dev = torch.device("cuda:0")
ops.LEM_PERSISTENT = True
rnn = LEMcuda(6, 128, 1.0).to(dev)
x = torch.randn(25, 6400, 6, device=dev)
for _ in range(3):
    ys, zs = rnn(x, last_only=True)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
_lib.lib.msmp_lem_debug_ticks.argtypes = [ctypes.c_void_p]
_lib.lib.msmp_lem_debug_ticks(buf)
t = list(buf)[:7]
names = ["G gemm (12 chunks)", "gate_z epilogue", "image->Z", "L gemm (4 chunks)", "gate_y epilogue", "image->Y"]
for i, n in enumerate(names):
    print(f"{n:22s} {t[i+1]-t[i]:8d} cycles")

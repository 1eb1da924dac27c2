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
t = list(buf)
seq = [(0, "start"), (1, "G gemm issued+done (gemm_wait)"), (10, "gate_z: tmem ld + first batch math"), (11, "wait_peer_free"),
       (12, "gate_z: remaining batches + stores"), (2, "publish (fence + bar + remote arrive)"),
       (13, "L gemm issue (incl. wait xfull)"), (3, "Z copy-out"), (4, "L gemm wait"), (5, "gate_y epilogue + publish"),
       (6, "Y copy-out")]
for (i0, _), (i1, name) in zip(seq[:-1], seq[1:]):
    print(f"{name:45s} {t[i1]-t[i0]:8d} cycles")
print("step total", t[6] - t[0])

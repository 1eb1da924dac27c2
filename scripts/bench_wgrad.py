"""Stand-alone timing of the weight-gradient kernels on the shapes of the C4 / C2 steps (CUDA events, L2 flushed)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = {"W4": (128, 0, 128, 0), "W3": (128, 128, 128, 3), "PQ": (128, 64, 256, 4), "LEM_G": (128, 32, 384, 0), "LEM_L": (128, 32, 128, 0)}
def run(name, M, ws, prec, reps=5):
    K0, K1, N, r = SHAPES[name]
    X = torch.randn(M, K0, device=dev); X1 = torch.randn(M, K1, device=dev) if K1 else None
    dY = torch.randn(M, N, device=dev); side = torch.randn(M, 8, device=dev) if r else None
    ops.WGRAD_WS, ops.PRECISION = ws, prec
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.linear_wgrad(X, dY, side=side, r=r, has_bias=True, X1=X1); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * M * (K0 + K1) * N
    by = 4.0 * M * (K0 + K1 + N)
    print(json.dumps(dict(op=name, M=M, kernel="ws" if ws else "tc", precision=prec, ms=round(ms, 4), tflops=round(fl / ms / 1e9, 1),
                          gbs=round(by / ms / 1e6, 1))), flush=True)
for M, names in ((131072, ("W4", "W3", "PQ")), (520192, ("W4",)), (25 * 131072, ("LEM_G", "LEM_L")), (6400, ("W4", "W3", "PQ")), (37632, ("W4",))):
    for n in names:
        run(n, M, False, "fp32"); run(n, M, True, "fp32"); run(n, M, True, "bf16")

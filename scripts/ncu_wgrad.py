"""One launch each of the weight-gradient kernels on C4 shapes, for `ncu --set full` (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import ops
dev = torch.device("cuda:0")
def run(M, K0, K1, N, r, ws):
    X = torch.randn(M, K0, device=dev); X1 = torch.randn(M, K1, device=dev) if K1 else None
    dY = torch.randn(M, N, device=dev); side = torch.randn(M, 8, device=dev) if r else None
    ops.WGRAD_WS = ws
    ops.WGRAD_WS_MAX_TALL_ROWS = 1 << 30
    for _ in range(2):
        ops.linear_wgrad(X, dY, side=side, r=r, has_bias=True, X1=X1)
    torch.cuda.synchronize()
run(131072, 128, 128, 128, 3, True)      # dW3, k_wgrad_ws
run(131072, 128, 128, 128, 3, False)     # dW3, k_wgrad_tc
run(520192, 128, 0, 128, 0, True)        # dW2
run(520192, 128, 0, 128, 0, False)
run(25 * 131072, 128, 32, 384, 0, False)  # LEM dG, k_wgrad_tc
run(25 * 131072, 128, 32, 384, 0, True)   # LEM dG, k_wgrad_ts (dY^T in tensor memory)
run(25 * 131072, 128, 32, 128, 0, True)   # LEM dL, k_wgrad_ts

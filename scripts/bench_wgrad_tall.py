"""Stand-alone timing of the 128 x 128 weight-gradient product on 0.5 - 6.3 Mi rows (k_wgrad_ts), with and without swish on X."""
import os, sys, json
sys.path.insert(0, '/root/repo')
import torch
from msmp_pde_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(M, xsw, label):
    X = torch.randn(M, 128, device=dev); dY = torch.randn(M, 128, device=dev)
    ts = []
    for i in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.linear_wgrad(X, dY, has_bias=True, xswish=xsw); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    print(label, M, "xswish", xsw, f"{ms:.3f} ms", f"{2.0 * M * 128 * 128 / ms / 1e9:.1f} TFLOP/s", flush=True)
for M in (520192, 1 << 20, 6291456):
    run(M, False, "W"); run(M, True, "W")

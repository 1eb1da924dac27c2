"""Per-kernel timings (CUDA events, back-to-back launches) to separate fixed from per-chunk costs.  Diagnostic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import ops
dev = torch.device("cuda:0")

def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3     # us

g = torch.Generator().manual_seed(0)
print("linear: M K Nout | tc us | ffma us")
for M in (128, 6400, 65536):
    for K, Nout in ((32, 128), (128, 128), (192, 256), (384, 128), (160, 384)):
        A = torch.randn(M, K, generator=g).to(dev); Wt = (torch.randn(K, Nout, generator=g) / 10).to(dev)
        out = torch.empty(M, Nout, device=dev)
        res = []
        for mode in ("tc", "ffma"):
            ops.GEMM_MODE = mode
            res.append(timeit(lambda: ops.linear_fwd([A], Wt, out=out)))
        print(f"  {M:6d} {K:4d} {Nout:4d} | {res[0]:8.1f} | {res[1]:8.1f}   (host-bound floor ~ python call)")
print("wgrad: M K Nout | tc us | ffma us")
for M in (6400, 160000):
    for K, Nout in ((128, 128), (128, 256), (128, 384)):
        X = torch.randn(M, K, generator=g).to(dev); dY = torch.randn(M, Nout, generator=g).to(dev)
        res = []
        for mode in ("tc", "ffma"):
            ops.GEMM_MODE = mode
            res.append(timeit(lambda: ops.linear_wgrad(X, dY, has_bias=True)))
        print(f"  {M:6d} {K:4d} {Nout:4d} | {res[0]:8.1f} | {res[1]:8.1f}")
# empty-ish kernel launch cost through the same python path
z = torch.empty(1024, device=dev)
print("mul_dswish 1K elems (launch floor through ctypes): %.1f us" % timeit(lambda: ops.mul_dswish(z, z)))
# graph replay of 20 linear launches: pure GPU time per launch
ops.GEMM_MODE = "tc"
A = torch.randn(6400, 192, generator=g).to(dev); Wt = (torch.randn(192, 256, generator=g) / 10).to(dev); out = torch.empty(6400, 256, device=dev)
for K in (32, 64, 96, 128, 192):
    A_ = A[:, :K].contiguous(); Wt_ = Wt[:K].contiguous()
    ops.linear_fwd([A_], Wt_, out=out); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20): ops.linear_fwd([A_], Wt_, out=out)
    print(f"graph replay linear_tc M=6400 K={K} Nout=256: {timeit(gr.replay, 20)/20:.2f} us per launch")

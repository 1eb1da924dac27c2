"""Stand-alone timing of the node-level GEMM launches of one layer (forward + dgrad) at M rows (CUDA events, L2 flushed).
Kernel variant by environment: MSMP_LINEAR_TMA=0 -> k_linear_ws, MSMP_LINEAR_WS_MIN_TILES=0 -> k_linear_tc."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g)
h, upad, agg, z3, dz4, dz3, dPQ, dcat, side = r(M, 128), r(M, 64), r(M, 128), r(M, 128), r(M, 128), r(M, 128), r(M, 256), r(M, 256), r(M, 8)
Wpq, W3, W4, W4d, W3hx, W1hq = r(192, 256) / 14, r(256, 128) / 16, r(128, 128) / 11, r(128, 128) / 11, r(128, 256) / 11, r(256, 128) / 16
b256, b128, Ws256, Ws128 = r(256), r(128), r(8, 256), r(8, 128)
z4 = torch.empty(M, 128, device=dev)
CALLS = {
    "PQ   K192 N256 side bias": (lambda: ops.linear_fwd([h, upad], Wpq, bias=b256, side=side, r=4, Wside=Ws256), 2.0 * M * 192 * 256, 4.0 * M * (192 + 256)),
    "z3   K256 N128 side bias": (lambda: ops.linear_fwd([h, agg], W3, bias=b128, side=side[:, 1:], r=3, Wside=Ws128), 2.0 * M * 256 * 128, 4.0 * M * (256 + 128)),
    "z4   K128 N128 swish act R": (lambda: ops.linear_fwd([z3], W4, bias=b128, Ypre=z4, act=True, R=h, aswish=[1]), 2.0 * M * 128 * 128, 4.0 * M * (128 * 4)),
    "dz3  K128 N128 Zmul": (lambda: ops.linear_fwd([dz4], W4d, Zmul=z3), 2.0 * M * 128 * 128, 4.0 * M * 384),
    "dcat K128 N256": (lambda: ops.linear_fwd([dz3], W3hx), 2.0 * M * 128 * 256, 4.0 * M * 384),
    "dh   K256 N128 R": (lambda: ops.linear_fwd([dPQ], W1hq, R=dcat[:, :128]), 2.0 * M * 256 * 128, 4.0 * M * 512),
}
tot = 0.0
for name, (fn, fl, by) in CALLS.items():
    ts = []
    for i in range(int(os.environ.get('BENCH_REPS', 7))):
        flush.zero_()
        if os.environ.get('BENCH_FLUSH') == 'read': flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 2 or int(os.environ.get('BENCH_REPS', 7)) < 3: ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    tot += ms
    print(f"{name:28s} {ops.PRECISION} M={M} ms={ms:.4f} tflops={fl / ms / 1e9:.1f} gbs={by / ms / 1e6:.1f}", flush=True)
print("total ms", round(tot, 4))

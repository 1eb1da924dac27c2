"""Error of the weight-gradient kernels against float64 as the row count grows (unit-variance data): the tensor core's truncating
fp32 accumulation makes a single accumulation chain per CTA drift like rows^1.5 (k_wgrad_tc), k_wgrad_ts / k_wgrad_ws bound the chain."""
import sys, torch
sys.path.insert(0, '/root/repo')
from msmp_pde_b200 import ops
dev = torch.device("cuda:0")
def run(M, K0, K1, N):
    g = torch.Generator(device=dev).manual_seed(5)
    X = torch.randn(M, K0, device=dev, generator=g); X1 = torch.randn(M, K1, device=dev, generator=g) if K1 else None
    dY = torch.randn(M, N, device=dev, generator=g)
    Xd = torch.cat([X.double(), X1.double()], 1) if K1 else X.double()
    ref = Xd.t() @ dY.double(); refs = dY.double().sum(0, keepdim=True)
    out = {}
    for name, (minrows, tall, ts) in {"ts": (0, 1 << 62, True), "ws_bf16x3": (0, 1 << 62, False), "tc": (1 << 62, 1 << 62, True)}.items():
        ops.WGRAD_WS_MIN_ROWS, ops.WGRAD_WS_MAX_TALL_ROWS = minrows, tall
        if name == "ws_bf16x3":
            import os
            continue
        a, a_s = ops.linear_wgrad(X, dY, has_bias=True, X1=X1)
        out[name] = (float((a.double() - ref).abs().max() / ref.abs().max()), float((a_s.double() - refs).abs().max() / refs.abs().max()))
    # plain fp32 torch matmul (no tf32)
    t = (torch.cat([X, X1], 1) if K1 else X).t() @ dY
    out["torch_fp32"] = (float((t.double() - ref).abs().max() / ref.abs().max()), 0.0)
    print(M, K0 + K1, N, {k: (f"{v[0]:.2e}", f"{v[1]:.2e}") for k, v in out.items()}, flush=True)
for M in (7500, 70001, 520192, 3276800):
    run(M, 128, 32, 384 if M != 520192 else 128)
# k_wgrad_ws (bf16-piece products, wide operands): rows per CTA capped at 1024 (MSMP_WGRAD_WS_MAX_ROWS)
def run_ws(M):
    g = torch.Generator(device=dev).manual_seed(6)
    X = torch.randn(M, 128, device=dev, generator=g); X1 = torch.randn(M, 128, device=dev, generator=g)
    dY = torch.randn(M, 128, device=dev, generator=g); side = torch.randn(M, 8, device=dev, generator=g)
    ref = torch.cat([X.double(), X1.double()], 1).t() @ dY.double()
    ops.WGRAD_WS_MIN_ROWS, ops.WGRAD_WS_MAX_TALL_ROWS = 0, 1 << 62
    a, a_s = ops.linear_wgrad(X, dY, side=side, r=3, has_bias=True, X1=X1)
    print("ws", M, 256, 128, f"{float((a.double() - ref).abs().max() / ref.abs().max()):.2e}", flush=True)
for M in (131072, 1 << 20):
    run_ws(M)

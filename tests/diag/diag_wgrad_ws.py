import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from msmp_pde_b200 import ops
dev = torch.device("cuda:0")
def run(M, K0, K1, Nout, r, bias, mode="fp32", ident=False):
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn(M, K0, device=dev, generator=g)
    X1 = torch.randn(M, K1, device=dev, generator=g) if K1 else None
    dY = torch.randn(M, Nout, device=dev, generator=g)
    if ident:   # one-hot probes: X[m] = e_{m % K0}, dY[m] = (m+1) * e_{m % Nout} on the first rows only
        X.zero_(); dY.zero_()
        for m in range(min(M, 16)):
            X[m, (3 * m + 1) % K0] = 1.0
            dY[m, (5 * m + 2) % Nout] = float(m + 1)
    side = torch.randn(M, 8, device=dev, generator=g) if r else None
    ops.PRECISION = mode
    dW, dWs = ops.linear_wgrad(X, dY, side=side, r=r, has_bias=bias, X1=X1)
    ops.PRECISION = "fp32"
    Xd = X.double() if X1 is None else torch.cat([X, X1], 1).double()
    ref = Xd.t() @ dY.double()
    err = float((dW.double() - ref).abs().max() / ref.abs().max())
    msg = f"M={M} K0={K0} K1={K1} N={Nout} r={r} bias={bias} mode={mode}: err={err:.3e}"
    if dWs is not None:
        cols = ([side[:, :r].double()] if r else []) + ([torch.ones(M, 1, dtype=torch.float64, device=dev)] if bias else [])
        refs = torch.cat(cols, 1).t() @ dY.double()
        msg += f" side_err={float((dWs.double() - refs).abs().max() / refs.abs().max()):.3e}"
    print(msg, flush=True)
    if ident:
        nz = dW.nonzero().tolist(); rz = ref.nonzero().tolist()
        print("  got ", [(k, n, float(dW[k, n])) for k, n in nz][:20])
        print("  want", [(k, n, float(ref[k, n])) for k, n in rz][:20])
for mode in ("fp32", "bf16"):
    run(16, 128, 0, 128, 0, False, mode, ident=True)
    run(16, 128, 0, 128, 0, False, mode)
    run(256, 128, 0, 128, 0, False, mode)
    run(6400, 128, 0, 128, 0, False, mode)
    run(6400, 128, 0, 128, 0, True, mode)
    run(6400, 128, 128, 128, 3, True, mode)
    run(6400, 128, 64, 256, 4, True, mode)
    run(7500, 128, 32, 384, 0, True, mode)

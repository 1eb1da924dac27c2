"""Thin Python wrappers over the C ABI (one call each, tensors in, tensors out, current CUDA stream).

These do no math of their own; they validate layout, size the shared scratch buffer and forward raw
device pointers to ``libmsmp_b200.so``.  CPU tensors are rejected: there is no fallback path.
"""
from __future__ import annotations

import ctypes
import os

import torch

from ._lib import check, lib
from .packing import TcW

H = 128
_ws: dict = {}

# bench.py instrumentation: number of msmp kernels launched, and (when a dict) CUDA events around the edge ops
LAUNCHES = 0
PROFILE_EVENTS = None


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


class _timed:
    """Records (start, stop, flops, bytes) with CUDA events on the current stream around one op when
    PROFILE_EVENTS is a dict (bench.py's per-kernel roofline pass)."""

    def __init__(self, name, flops=0.0, nbytes=0.0):
        self.name, self.flops, self.nbytes = name, flops, nbytes

    def __enter__(self):
        if PROFILE_EVENTS is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if PROFILE_EVENTS is not None:
            self.e.record()
            PROFILE_EVENTS.setdefault(self.name, []).append((self.s, self.e, self.flops, self.nbytes))
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _workspace(nbytes: int, device) -> torch.Tensor:
    # one scratch buffer per (device, stream): ops issued on different streams may run concurrently
    key = (device.type, device.index, torch.cuda.current_stream().cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 22), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def _req(t: torch.Tensor, name: str, dtype=torch.float32):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: msmp_pde_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.dim() >= 1 and t.stride(-1) != 1 and t.numel() > 1:
        raise ValueError(f"{name}: innermost dimension must be contiguous")
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.dim() == 2 else t.shape[-1]


def _p(t):
    return 0 if t is None else t.data_ptr()


# Tensor-core (tcgen05 3xTF32) execution of the dense layers.  "ffma" keeps the exact-fp32 CUDA-core kernels.
GEMM_MODE = "tc"
# Operand precision of the tensor-core GEMMs that have a bf16 variant (north_star's "bf16 mode"): "fp32" = error-compensated
# 3xTF32 (parity at 1e-5), "bf16" = bf16 operands with fp32 accumulation (tolerance stated in DESIGN.md section 8).
PRECISION = os.environ.get("MSMP_PRECISION", "fp32")
# persistent warp-specialised weight-gradient kernel (csrc/wgrad_ws.cu); "0" keeps the one-tile-per-CTA k_wgrad_tc
WGRAD_WS = os.environ.get("MSMP_WGRAD_WS", "1") != "0"
# ... except for the very tall, narrow products (the LEM weight gradients of the large-graph configs: T x N >= 1 Mi rows,
# K = 160): there the MMAs of both kernels are bound by the MN-major operand fetch, k_wgrad_ws has no re-read to save and
# k_wgrad_tc measured faster (3.3 Mi rows x 384: 3.5 against 4.4 ms, profiles/r2_bench_wgrad.jsonl).  fp32 mode only.
WGRAD_WS_MAX_TALL_ROWS = int(os.environ.get("MSMP_WGRAD_WS_MAX_TALL_ROWS", str(1 << 18)))
# ... and for small row counts (the reference's 100-node graphs: 6 400 nodes / 37 632 edges per step), where a launch is
# latency bound either way and the persistent kernel's longer prologue costs more than its single read of the operands
# saves: C2 step 3.74 ms with k_wgrad_tc against 3.87 ms (C4, 131 072 nodes: 46.9 against 46.0 ms).
WGRAD_WS_MIN_ROWS = int(os.environ.get("MSMP_WGRAD_WS_MIN_ROWS_USE", str(1 << 16)))
# k_wgrad_ts inside msmp_wgrad_ws (fp32-parity mode, at most 192 operand columns); "0" restores the bf16-piece kernel there
WGRAD_TS = os.environ.get("MSMP_WGRAD_TS", "1") != "0"


def wgrad_use_ws(M: int, Kt: int, Nout: int, nside: int = 1) -> bool:
    """Which weight-gradient kernel a [M, Kt]^T [M, Nout] product takes in tensor-core mode (gradsink.GradPlan lays its
    sink regions out accordingly)."""
    if not (GEMM_MODE == "tc" and WGRAD_WS and Kt % 32 == 0 and Nout % 128 == 0 and nside <= 8):
        return False
    if PRECISION == "bf16":
        return True
    if M < WGRAD_WS_MIN_ROWS:
        return False
    if M >= WGRAD_WS_MAX_TALL_ROWS and Kt <= 192:
        # tall and narrow: msmp_wgrad_ws runs k_wgrad_ts (dY^T in tensor memory) for Kt <= 160 columns: LEM dL
        # product (3.3 Mi rows) 1.00 against 1.19 ms with k_wgrad_tc, LEM dG product (three dY blocks, fetched by tensor-map
        # copies) 2.94 against 3.53 ms, edge dW2 (520 Ki rows) 0.165 against 0.210 ms
        return WGRAD_TS and Kt <= 160 and nside <= 1
    return True
# Persistent (all-T-steps-in-one-launch) LEM kernels; False = one GEMM + one gate kernel per step.
LEM_PERSISTENT = True
# tensor-core edge kernels: warp-specialised, weights in tensor memory (edge_ws.cu) | single-role (edge_tc.cu)
EDGE_WS = os.environ.get("MSMP_EDGE_WS", "1") != "0"
# ... from this many 128-edge tiles on.  Measured on B200 (scripts/edge_sweep.py and bench.py): the pipelined kernels
# are faster from about three tiles per SM on (2.4x at 1 Mi x 6 Mi) and equal below; inside the captured C2 step
# (two tiles per SM) they are 3 % faster than the single-role kernels, so the default is "always".
EDGE_WS_MIN_TILES_FWD = int(os.environ.get("MSMP_EDGE_WS_MIN_TILES_FWD", "0"))
EDGE_WS_MIN_TILES_BWD = int(os.environ.get("MSMP_EDGE_WS_MIN_TILES_BWD", "0"))
# InstanceNorm (+ gate blend) as one launch per direction when every graph of the batch fits one statistics chunk
INSTNORM_FUSED = os.environ.get("MSMP_INSTNORM_FUSED", "1") != "0"
# The persistent backward recurrence can be cut into several launches so that the weight-gradient GEMMs of finished
# steps overlap the remaining ones on a side stream (lem._LEMFn.backward).  Measured on the C2 workload after the
# recurrence kernel got its own MMA warp and 16 epilogue warps: 1 launch 4.11 ms/step, 2: 4.16, 3: 4.19, 5: 4.28,
# 8: 4.40 -- the co-running GEMMs slow the latency-bound recurrence more than the overlap gains, so the default is 1.
# Plain autograd path only: under GraphedTrainStep (gradient sink) the recurrence always runs as one launch.
LEM_BWD_SEGMENTS = int(__import__('os').environ.get('MSMP_LEM_BWD_SEGMENTS', 1))
# gradsink.GradPlan of the backward pass in flight (set by GraphedTrainStep): the backward Functions then leave their
# weight gradients in the plan's raw buffer and return None for the parameters; None = plain autograd behaviour.
GRAD_SINK = None
# bench.py's per-kernel pass: no side streams (layers._side_stream returns the current stream), so that the CUDA events
# around an op time the kernel alone and not its wait for SMs that another stream's kernels hold
SERIALIZE = False
# GraphedTrainStep sets this around the model call: the solvers then return their float32 output as it leaves the decoder
# instead of casting it to the input dtype (the fused criterion below converts on the fly)
RAW_OUTPUT = False
_IMG_CACHE: dict = {}


def _cached_images(Wt: torch.Tensor) -> torch.Tensor:
    """tc_images(Wt), cached on the identity + version of the weight tensor (weights change once per step)."""
    key = id(Wt)
    ent = _IMG_CACHE.get(key)
    if ent is not None:
        ref, ver, ptr, img = ent
        if ref() is Wt and ver == Wt._version and ptr == Wt.data_ptr():
            return img
    img = tc_images(Wt if Wt.is_contiguous() else Wt.contiguous())
    import weakref

    def _drop(_, key=key):
        _IMG_CACHE.pop(key, None)
    _IMG_CACHE[key] = (weakref.ref(Wt, _drop), Wt._version, Wt.data_ptr(), img)
    return img


def linear_fwd(segs, Wt, bias=None, side=None, r=0, Wside=None, Zmul=None, Ypre=None, act=False, R=None,
               out=None, Nout=None, aswish=None):
    """Y = epi([A0|A1|A2] @ Wt + bias + side[:, :r] @ Wside); see include/msmp_b200.h."""
    if isinstance(Wt, TcW):          # pre-packed by packing.PackPlan (tensor-core mode only)
        return linear_tc_fwd(segs, Wt.img, Wt.N if Nout is None else Nout, bias=bias, side=side, r=r, Wside=Wside,
                             Zmul=Zmul, Ypre=Ypre, act=act, R=R, out=out, aswish=aswish)
    if GEMM_MODE == "tc":
        return linear_tc_fwd(segs, _cached_images(Wt), Wt.shape[1] if Nout is None else Nout, bias=bias, side=side,
                             r=r, Wside=Wside, Zmul=Zmul, Ypre=Ypre, act=act, R=R, out=out, aswish=aswish)
    n = len(segs)
    M = segs[0].shape[0]
    Nout = Wt.shape[1] if Nout is None else Nout
    for i, a in enumerate(segs):
        _req(a, f"A{i}")
    _req(Wt, "Wt")
    if out is None:
        out = torch.empty(M, Nout, dtype=torch.float32, device=Wt.device)
    A = (ctypes.c_void_p * 3)(*[a.data_ptr() for a in segs], *([0] * (3 - n)))
    lda = (ctypes.c_int * 3)(*[_ld(a) for a in segs], *([0] * (3 - n)))
    ka = (ctypes.c_int * 3)(*[a.shape[1] for a in segs], *([0] * (3 - n)))
    asw = (ctypes.c_int * 3)(*([int(bool(x)) for x in aswish] if aswish else [0] * n), *([0] * (3 - n)))
    check(lib.msmp_linear_fwd(A, lda, ka, asw, n, Wt.data_ptr(), _ld(Wt), _p(bias), _p(side),
                              _ld(side) if side is not None else 0, r if side is not None else 0, _p(Wside),
                              _p(Zmul), _ld(Zmul) if Zmul is not None else 0, _p(Ypre),
                              _ld(Ypre) if Ypre is not None else 0, int(act), _p(R), _ld(R) if R is not None else 0,
                              out.data_ptr(), _ld(out), M, Nout, _stream()), "msmp_linear_fwd")
    _count(1)
    return out


_IMG_IDX = {}


def _img_index(device):
    """float offset of element (n, k) inside a [128 x 32] UMMA 128B-swizzled tile image (see msmp_b200.h)."""
    idx = _IMG_IDX.get(device)
    if idx is None:
        n = torch.arange(128, device=device).view(128, 1)
        k = torch.arange(32, device=device).view(1, 32)
        idx = ((n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 2) ^ n) & 7) << 4)) // 4 + (k & 3)
        idx = idx.reshape(-1)
        _IMG_IDX[device] = idx
    return idx


def tc_images(Wt: torch.Tensor) -> torch.Tensor:
    """Pre-split (tf32 hi | lo), pre-swizzled weight images for msmp_linear_tc_fwd from the k-major weight
    Wt[K, N] (K % 32 == 0): returns [ceil(N/128), K/32, 2, 4096] fp32."""
    K, N = Wt.shape
    nt, nc = (N + 127) // 128, K // 32
    Wp = Wt.new_zeros(K, nt * 128)
    Wp[:, :N] = Wt
    B = Wp.t().reshape(nt, 128, nc, 32).permute(0, 2, 1, 3).contiguous()          # [nt, nc, n, k]
    hi = ((B.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    lo = B - hi
    img = Wt.new_empty(nt, nc, 2, 4096)
    idx = _img_index(Wt.device)
    img[:, :, 0, idx] = hi.reshape(nt, nc, 4096)
    img[:, :, 1, idx] = lo.reshape(nt, nc, 4096)
    return img


def linear_tc_fwd(segs, img, Nout, bias=None, side=None, r=0, Wside=None, Zmul=None, Ypre=None, act=False, R=None,
                  out=None, aswish=None):
    """Tensor-core (3xTF32) version of linear_fwd; ``img`` = tc_images(Wt)."""
    n = len(segs)
    M = segs[0].shape[0]
    for i, a in enumerate(segs):
        _req(a, f"A{i}")
    if out is None:
        out = torch.empty(M, Nout, dtype=torch.float32, device=img.device)
    A = (ctypes.c_void_p * 3)(*[a.data_ptr() for a in segs], *([0] * (3 - n)))
    lda = (ctypes.c_int * 3)(*[_ld(a) for a in segs], *([0] * (3 - n)))
    ka = (ctypes.c_int * 3)(*[a.shape[1] for a in segs], *([0] * (3 - n)))
    asw = (ctypes.c_int * 3)(*([int(bool(x)) for x in aswish] if aswish else [0] * n), *([0] * (3 - n)))
    Ktot = sum(a.shape[1] for a in segs)
    with _timed("linear_tc", 2.0 * M * Ktot * Nout, 4.0 * (M * Ktot + M * Nout)):
      check(lib.msmp_linear_tc_fwd(A, lda, ka, asw, n, img.data_ptr(), _p(bias), _p(side),
                                 _ld(side) if side is not None else 0, r if side is not None else 0, _p(Wside),
                                 _ld(Wside) if Wside is not None else 0, _p(Zmul),
                                 _ld(Zmul) if Zmul is not None else 0, _p(Ypre), _ld(Ypre) if Ypre is not None else 0,
                                 int(act), _p(R), _ld(R) if R is not None else 0, out.data_ptr(), _ld(out), M, Nout,
                                 1 if PRECISION == "bf16" else 0, _stream()), "msmp_linear_tc_fwd")
    _count(1)
    return out


def _side_base(side):
    """(base pointer, row stride, first column) of a side array that may be a column slice of a [M, lds] tensor."""
    lds = side.stride(0)
    c0 = side.storage_offset() % lds if lds <= 16 else 0
    return side.data_ptr() - 4 * c0, lds, c0


def linear_wgrad(X, dY, K=None, xswish=False, side=None, r=0, has_bias=False, dWt=None, dWside=None,
                 accumulate=False, X1=None):
    """dWt[K, Nout] = [X | X1]^T dY ; dWside[r(+1), Nout] = [side|1]^T dY.  X1 (optional) supplies the columns
    after X's (X.shape[1] % 128 == 0), e.g. [h | agg] or [h | u_padded], without materialising the concatenation.
    A 3-dim ``dWt`` [S, K, Nout] (with ``dWside`` [S, nside, Nout]) asks for the split-M partials of msmp_wgrad_ws
    instead of the reduced gradient (gradsink.GradPlan sums them in its unpack launch)."""
    _req(X, "X")
    _req(dY, "dY")
    M, Nout = dY.shape
    K0 = X.shape[1] if K is None else K
    K1 = X1.shape[1] if X1 is not None else 0
    Kt = K0 + K1
    nside = (r if side is not None else 0) + int(has_bias)
    dev = dY.device
    partial = dWt is not None and dWt.dim() == 3
    ws_ok = (K0 % 32 == 0 and K1 % 32 == 0 and (side is None or (side.stride(0) <= 16 and side.stride(0) % 4 == 0))
             and (partial or wgrad_use_ws(M, Kt, Nout, nside)))
    if partial and not ws_ok:
        raise RuntimeError("linear_wgrad: split-M partials were requested for a shape msmp_wgrad_ws does not take")
    if dWt is None:
        dWt = torch.empty(Kt, Nout, dtype=torch.float32, device=dev)
    if nside and dWside is None:
        dWside = torch.empty(nside, Nout, dtype=torch.float32, device=dev)
    if ws_ok:
        segs = [X] if X1 is None else [X, X1]
        n = len(segs)
        S = lib.msmp_wgrad_ws_splits(M, Kt, Nout, nside)
        if S <= 0:
            raise RuntimeError("msmp_wgrad_ws_splits rejected the shape")
        Xp = (ctypes.c_void_p * 3)(*[a.data_ptr() for a in segs], *([0] * (3 - n)))
        ldx = (ctypes.c_int * 3)(*[_ld(a) for a in segs], *([0] * (3 - n)))
        kx = (ctypes.c_int * 3)(K0, *([K1] if X1 is not None else []), *([0] * (3 - n)))
        xsw = (ctypes.c_int * 3)(*([int(bool(xswish))] * n), *([0] * (3 - n)))
        sb, lds, c0 = _side_base(side) if side is not None else (0, 0, 0)
        if partial:
            if dWt.shape[0] != S or not dWt.is_contiguous() or (nside and (dWside is None or dWside.shape[0] != S)):
                raise RuntimeError(f"linear_wgrad: the partial buffer holds {dWt.shape[0]} splits, this call makes {S}")
            if accumulate:
                raise RuntimeError("linear_wgrad: accumulate is not available with split-M partials")
            part, part_side, out, out_side = dWt.data_ptr(), _p(dWside), 0, 0
        else:
            ws = _workspace(lib.msmp_wgrad_ws_workspace(M, Kt, Nout, nside), dev)
            part = ws.data_ptr()
            part_side = part + 4 * S * Kt * Nout
            out, out_side = dWt.data_ptr(), _p(dWside)
        is_ts = WGRAD_TS and PRECISION != "bf16" and Kt <= 160 and nside <= 1      # the kernel msmp_wgrad_ws will pick
        with _timed("wgrad_ts" if is_ts else "wgrad_ws", 2.0 * M * Kt * Nout, 4.0 * (M * Kt + M * Nout)):
            check(lib.msmp_wgrad_ws(Xp, ldx, kx, xsw, n, dY.data_ptr(), _ld(dY), Nout, sb, lds, c0,
                                    r if side is not None else 0, int(has_bias), part, part_side, out, out_side,
                                    int(accumulate), M, 1 if PRECISION == "bf16" else 0, _stream()), "msmp_wgrad_ws")
        _count(1 if partial else 2)
        return dWt, dWside
    sargs = (_p(side), _ld(side) if side is not None else 0, r if side is not None else 0, int(has_bias))
    if GEMM_MODE == "tc":
        ws = _workspace(lib.msmp_linear_wgrad_workspace(M, Kt, Nout, nside), dev)
        with _timed("wgrad_tc", 2.0 * M * Kt * Nout, 4.0 * (M * Kt + M * Nout)):
          check(lib.msmp_linear_wgrad_tc2(X.data_ptr(), _ld(X), K0, _p(X1), _ld(X1) if X1 is not None else 0, K1,
                                        int(xswish), dY.data_ptr(), _ld(dY), Nout, *sargs, dWt.data_ptr(), _p(dWside),
                                        int(accumulate), M, ws.data_ptr(), ws.numel(), _stream()),
              "msmp_linear_wgrad_tc2")
        _count(2)
        return dWt, dWside
    ws = _workspace(lib.msmp_linear_wgrad_workspace(M, max(K0, K1), Nout, nside), dev)
    check(lib.msmp_linear_wgrad(X.data_ptr(), _ld(X), K0, int(xswish), dY.data_ptr(), _ld(dY), Nout, *sargs,
                                dWt.data_ptr(), _p(dWside), int(accumulate), M, ws.data_ptr(), ws.numel(), _stream()),
          "msmp_linear_wgrad")
    _count(2)
    if X1 is not None:
        check(lib.msmp_linear_wgrad(X1.data_ptr(), _ld(X1), K1, int(xswish), dY.data_ptr(), _ld(dY), Nout, 0, 0, 0, 0,
                                    dWt[K0:].data_ptr(), 0, int(accumulate), M, ws.data_ptr(), ws.numel(), _stream()),
              "msmp_linear_wgrad")
        _count(2)
    return dWt, dWside


def _ws_operand(W2raw, Wk, what):
    """(tensor, row stride, column stride) of the warp-specialised edge kernels' A operand [m][k]."""
    if W2raw is not None:                 # the parameter W2[n][k] itself
        _req(W2raw, "W2")
        if tuple(W2raw.shape) != (H, H) or W2raw.stride(0) != H:
            raise ValueError("W2: expected a contiguous [128, 128] tensor")
        return (W2raw, H, 1) if what == "fwd" else (W2raw, 1, H)
    if isinstance(Wk, TcW):
        return None
    _req(Wk, "W2")
    if tuple(Wk.shape) != (H, H) or Wk.stride(0) != H:
        raise ValueError("W2: expected a contiguous [128, 128] tensor")
    return (Wk, 1, H)                     # fwd: W2t[k][n] read as A[n][k];  bwd: W2[n][k] read as A[k][n]


def edge_fwd(P, Q, topo, W2t, b2, save_z2=True, W2raw=None):
    """W2t: k-major W2^T (tensor or packed images).  W2raw: the [n][k] parameter itself; with it (or a plain W2t) the
    tensor-core mode runs the warp-specialised kernel, which keeps the weights in tensor memory."""
    _req(P, "P")
    _req(Q, "Q")
    dev = P.device
    agg = torch.empty(topo.N, H, dtype=torch.float32, device=dev)
    z2 = torch.empty(topo.E, H, dtype=torch.float32, device=dev) if save_z2 else None
    ws = _workspace(lib.msmp_edge_fwd_workspace(topo.E), dev)
    use_ws = EDGE_WS and (GEMM_MODE == "tc" or isinstance(W2t, TcW)) and topo.E >= 128 * EDGE_WS_MIN_TILES_FWD
    wsop = _ws_operand(W2raw, W2t, "fwd") if use_ws else None
    if wsop is not None:
        Wa, rs, cs = wsop
        ws = _workspace(lib.msmp_edge_ws_workspace(topo.E), dev)
        with _timed("edge_ws_fwd", 2.0 * topo.E * H * H, 4.0 * H * (3 * topo.E + topo.N)):
            check(lib.msmp_edge_ws_fwd(P.data_ptr(), Q.data_ptr(), _ld(P), topo.src.data_ptr(), topo.dst.data_ptr(),
                                       topo.rowptr.data_ptr(), topo.inv_deg.data_ptr(), Wa.data_ptr(), rs, cs,
                                       b2.data_ptr(), _p(z2), agg.data_ptr(), topo.E, topo.N, int(topo.no_isolated),
                                       ws.data_ptr(), ws.numel(), _stream()), "msmp_edge_ws_fwd")
        _count(2)
        return agg, z2
    if GEMM_MODE == "tc" or isinstance(W2t, TcW):
        img = W2t.img if isinstance(W2t, TcW) else _cached_images(W2t)
        with _timed("edge_tc_fwd", 2.0 * topo.E * H * H, 4.0 * H * (3 * topo.E + topo.N)):
            check(lib.msmp_edge_tc_fwd(P.data_ptr(), Q.data_ptr(), _ld(P), topo.src.data_ptr(), topo.dst.data_ptr(),
                                       topo.rowptr.data_ptr(), topo.inv_deg.data_ptr(), img.data_ptr(), b2.data_ptr(),
                                       _p(z2), agg.data_ptr(), topo.E, topo.N, ws.data_ptr(), ws.numel(), _stream()),
                  "msmp_edge_tc_fwd")
        _count(2)
        return agg, z2
    with _timed("edge_fwd"):
        check(lib.msmp_edge_fwd(P.data_ptr(), Q.data_ptr(), _ld(P), topo.src.data_ptr(), topo.dst.data_ptr(),
                                topo.rowptr.data_ptr(), topo.inv_deg.data_ptr(), W2t.data_ptr(), b2.data_ptr(),
                                _p(z2), agg.data_ptr(), topo.E, topo.N, ws.data_ptr(), ws.numel(), _stream()),
              "msmp_edge_fwd")
    _count(2)
    return agg, z2


def edge_bwd(P, Q, topo, W2, z2, dagg, dP, defer_wgrad=False, W2raw=None):
    """Returns dz1 [E,128], dW2 [128,128] ([n][k] = parameter layout), db2 [128]; writes dP in place.
    defer_wgrad (tensor-core path): returns (dz1, a1, dz2) instead; the caller runs
    ``linear_wgrad(a1, dz2, has_bias=True)`` itself (e.g. on a side stream)."""
    dev = P.device
    dz1 = torch.empty(topo.E, H, dtype=torch.float32, device=dev)
    if GEMM_MODE == "tc" or isinstance(W2, TcW):
        dz2 = torch.empty(topo.E, H, dtype=torch.float32, device=dev)
        a1 = torch.empty(topo.E, H, dtype=torch.float32, device=dev)
        ws = _workspace(lib.msmp_edge_fwd_workspace(topo.E), dev)
        wsop = _ws_operand(W2raw, W2, "bwd") if (EDGE_WS and topo.E >= 128 * EDGE_WS_MIN_TILES_BWD) else None
        if wsop is not None:
            Wa, rs, cs = wsop
            ws = _workspace(lib.msmp_edge_ws_workspace(topo.E), dev)
            with _timed("edge_ws_bwd", 2.0 * topo.E * H * H, 4.0 * H * (7 * topo.E + topo.N)):
                check(lib.msmp_edge_ws_bwd(P.data_ptr(), Q.data_ptr(), _ld(P), topo.src.data_ptr(), topo.dst.data_ptr(),
                                           topo.rowptr.data_ptr(), topo.inv_deg_e.data_ptr(), Wa.data_ptr(), rs, cs,
                                           z2.data_ptr(), dagg.data_ptr(), _ld(dagg), dz2.data_ptr(), a1.data_ptr(),
                                           dz1.data_ptr(), dP.data_ptr(), _ld(dP), topo.E, topo.N, int(topo.no_isolated),
                                           ws.data_ptr(), ws.numel(), _stream()), "msmp_edge_ws_bwd")
        else:
            img = W2.img if isinstance(W2, TcW) else _cached_images(W2)
            with _timed("edge_tc_bwd", 2.0 * topo.E * H * H, 4.0 * H * (7 * topo.E + topo.N)):
                check(lib.msmp_edge_tc_bwd(P.data_ptr(), Q.data_ptr(), _ld(P), topo.src.data_ptr(), topo.dst.data_ptr(),
                                           topo.rowptr.data_ptr(), topo.inv_deg.data_ptr(), img.data_ptr(), z2.data_ptr(),
                                           dagg.data_ptr(), _ld(dagg), dz2.data_ptr(), a1.data_ptr(), dz1.data_ptr(),
                                           dP.data_ptr(), _ld(dP), topo.E, topo.N, ws.data_ptr(), ws.numel(),
                                           _stream()), "msmp_edge_tc_bwd")
        _count(2)
        if defer_wgrad:
            return dz1, a1, dz2
        if topo.E > 0:
            dW2t, dbs = linear_wgrad(a1, dz2, has_bias=True)       # dW2t[k][n] = sum_e a1[e][k] dz2[e][n]
        else:
            dW2t, dbs = torch.zeros(H, H, device=dev), torch.zeros(1, H, device=dev)
        return dz1, dW2t.t(), dbs[0]
    dW2 = torch.empty(H, H, dtype=torch.float32, device=dev)
    db2 = torch.empty(H, dtype=torch.float32, device=dev)
    ws = _workspace(lib.msmp_edge_bwd_workspace(topo.E), dev)
    with _timed("edge_bwd"):
        check(lib.msmp_edge_bwd(P.data_ptr(), Q.data_ptr(), _ld(P), topo.src.data_ptr(), topo.dst.data_ptr(),
                                topo.rowptr.data_ptr(), topo.inv_deg.data_ptr(), W2.data_ptr(), z2.data_ptr(),
                                dagg.data_ptr(), _ld(dagg), dz1.data_ptr(), dP.data_ptr(), _ld(dP), dW2.data_ptr(),
                                db2.data_ptr(), topo.E, topo.N, ws.data_ptr(), ws.numel(), _stream()),
              "msmp_edge_bwd")
    _count(4)
    return dz1, dW2, db2


def segment_reduce(src, ptr, perm=None, scale=None, out=None, N=None):
    """out[n] = scale[n] * sum_{k in [ptr[n], ptr[n+1])} src[perm[k] or k]  (rows of 128 floats)."""
    _req(src, "src")
    N = ptr.numel() - 1 if N is None else N
    if out is None:
        out = torch.empty(N, H, dtype=torch.float32, device=src.device)
    with _timed("segment_reduce", 0.0, 4.0 * (src.shape[0] * H + N * H + N + 1)):
        check(lib.msmp_segment_reduce(src.data_ptr(), _ld(src), _p(perm), ptr.data_ptr(), _p(scale), out.data_ptr(),
                                      _ld(out), N, _stream()), "msmp_segment_reduce")
    _count(1)
    return out


def instnorm_fwd(y0, topo, y1=None, h=None, eps=1e-5):
    """mode 0: IN(y0);  mode 1 (y1, h given): (1-s) h + s*swish(IN(y1)), s = sigmoid(IN(y0)).
    Returns (out, stat)."""
    mode = 0 if y1 is None else 1
    dev = y0.device
    out = torch.empty(topo.N, H, dtype=torch.float32, device=dev)
    stat = torch.empty(mode + 1, topo.B, 2, H, dtype=torch.float32, device=dev)
    if INSTNORM_FUSED and topo.one_chunk_per_graph:
        check(lib.msmp_instnorm1_fwd(y0.data_ptr(), _p(y1), _ld(y0), _p(h), topo.chunk_begin.data_ptr(),
                                     topo.chunk_end.data_ptr(), topo.B, topo.N, mode, eps, stat.data_ptr(),
                                     out.data_ptr(), _stream()), "msmp_instnorm1_fwd")
        _count(1)
        return out, stat
    ws = _workspace(lib.msmp_instnorm_workspace(topo.nchunks, topo.B), dev)
    check(lib.msmp_instnorm_fwd(y0.data_ptr(), _p(y1), _ld(y0), _p(h), topo.chunk_begin.data_ptr(),
                                topo.chunk_end.data_ptr(), topo.graph_chunk_ptr.data_ptr(), topo.node_graph.data_ptr(),
                                topo.nchunks, topo.B, topo.N, mode, eps, stat.data_ptr(), out.data_ptr(),
                                ws.data_ptr(), ws.numel(), _stream()), "msmp_instnorm_fwd")
    _count(3)
    return out, stat


def instnorm_bwd(dout, y0, topo, stat, y1=None, h=None):
    """Returns dy0 (mode 0) or (dy0, dy1, dh) (mode 1)."""
    mode = 0 if y1 is None else 1
    dev = y0.device
    dout = dout.contiguous()
    dy0 = torch.empty(topo.N, H, dtype=torch.float32, device=dev)
    dy1 = torch.empty(topo.N, H, dtype=torch.float32, device=dev) if mode else None
    dh = torch.empty(topo.N, H, dtype=torch.float32, device=dev) if mode else None
    if INSTNORM_FUSED and topo.one_chunk_per_graph:
        check(lib.msmp_instnorm1_bwd(dout.data_ptr(), y0.data_ptr(), _p(y1), _ld(y0), _p(h), stat.data_ptr(),
                                     topo.chunk_begin.data_ptr(), topo.chunk_end.data_ptr(), topo.B, topo.N, mode,
                                     dy0.data_ptr(), _p(dy1), H, _p(dh), _stream()), "msmp_instnorm1_bwd")
        _count(1)
        return dy0 if mode == 0 else (dy0, dy1, dh)
    ws = _workspace(lib.msmp_instnorm_workspace(topo.nchunks, topo.B), dev)
    check(lib.msmp_instnorm_bwd(dout.data_ptr(), y0.data_ptr(), _p(y1), _ld(y0), _p(h), stat.data_ptr(),
                                topo.chunk_begin.data_ptr(), topo.chunk_end.data_ptr(),
                                topo.graph_chunk_ptr.data_ptr(), topo.node_graph.data_ptr(), topo.nchunks, topo.B,
                                topo.N, mode, dy0.data_ptr(), _p(dy1), H, _p(dh), ws.data_ptr(), ws.numel(),
                                _stream()), "msmp_instnorm_bwd")
    _count(3)
    return dy0 if mode == 0 else (dy0, dy1, dh)


def mul_dswish(g, z):
    g = g.contiguous()
    out = torch.empty_like(z)
    check(lib.msmp_mul_dswish(g.data_ptr(), z.data_ptr(), out.data_ptr(), z.numel(), _stream()), "msmp_mul_dswish")
    _count(1)
    return out


# ---- input assembly (prep.cu) -------------------------------------------------------------------------
class _LemCol(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("ld", ctypes.c_int), ("off", ctypes.c_int), ("kind", ctypes.c_int),
                ("reserved", ctypes.c_int)]


LEM_COL_STATIC, LEM_COL_TIME, LEM_COL_CLOCK = 0, 1, 2


def node_features(u, pos_x, variables, ldu):
    """(upad [N, ldu] = [u | 0], side [N, 8] = [pos_x, variables..., 0]) from contiguous fp32 u [N,F_u], pos_x [N,1],
    variables [N,V] in one launch."""
    N, F_u = u.shape
    V = variables.shape[1]
    for t in (u, pos_x, variables):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError("node_features: contiguous float32 CUDA tensors")
    upad = torch.empty(N, ldu, dtype=torch.float32, device=u.device)
    side = torch.empty(N, 8, dtype=torch.float32, device=u.device)
    check(lib.msmp_node_features(u.data_ptr(), F_u, pos_x.data_ptr(), variables.data_ptr() if V else 0, V, N, upad.data_ptr(),
                                 ldu, side.data_ptr(), _stream()), "msmp_node_features")
    _count(1 if N else 0)
    return upad, side


def lem_inputs(T, N, cols, clock=None, node_t=None):
    """inp [T, N, 32] (zero padded) of the LEM recurrence in one launch.  ``cols``: one entry per input column,
    ("static", x [N,k] fp32, j) -> x[n, j];  ("time", x [N,k] fp32, j) -> x[n, j + t];  ("clock",) -> float(clock[t] +
    node_t[n]) with float64 ``clock`` [T] and ``node_t`` [N]."""
    if not 1 <= len(cols) <= 8:
        raise ValueError("lem_inputs: 1..8 columns")
    arr = (_LemCol * len(cols))()
    keep = []
    dev = None
    for i, c in enumerate(cols):
        if c[0] == "clock":
            if clock is None or node_t is None or clock.dtype != torch.float64 or node_t.dtype != torch.float64 \
                    or clock.numel() != T or node_t.numel() != N or not (clock.is_contiguous() and node_t.is_contiguous()):
                raise ValueError("lem_inputs: the clock column needs contiguous float64 clock [T] and node_t [N]")
            arr[i] = _LemCol(None, 0, 0, LEM_COL_CLOCK, 0)
            dev = clock.device
            continue
        kind, x, j = c
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous() and x.shape[0] == N):
            raise ValueError("lem_inputs: column sources are contiguous float32 [N, k] CUDA tensors")
        span = T if kind == "time" else 1
        if not 0 <= j <= x.shape[1] - span:
            raise ValueError("lem_inputs: column offset out of range")
        arr[i] = _LemCol(x.data_ptr(), x.shape[1], j, LEM_COL_TIME if kind == "time" else LEM_COL_STATIC, 0)
        keep.append(x)
        dev = x.device
    inp = torch.empty(T, N, 32, dtype=torch.float32, device=dev)
    check(lib.msmp_lem_inputs(ctypes.cast(arr, ctypes.c_void_p), len(cols), _p(clock), _p(node_t), T, N, inp.data_ptr(),
                              _stream()), "msmp_lem_inputs")
    _count(1 if T * N else 0)
    return inp


# ---- G^2 gate statistic (models_gnn2D.py:598-603) ------------------------------------------------------
def g2_fwd(t, topo, inv):
    """out[s] = inv[s] * sum_{e: src e = s} (t[s] - t[dst e])^2 (CSC order), [N,128] fp32."""
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape == (topo.N, H)):
        raise ValueError("g2_fwd: contiguous float32 [N,128] CUDA tensor")
    out = torch.empty_like(t)
    check(lib.msmp_g2_fwd(t.data_ptr(), topo.colptr.data_ptr(), topo.csc_perm.data_ptr(), topo.dst.data_ptr(), inv.data_ptr(),
                          out.data_ptr(), topo.N, _stream()), "msmp_g2_fwd")
    _count(1 if topo.N else 0)
    return out


def g2_bwd(t, g, topo, inv):
    """d/dt of g2_fwd applied to g (deterministic: one pass over a node's out-edges, one over its in-edges)."""
    g = g.contiguous()
    dt = torch.empty_like(t)
    check(lib.msmp_g2_bwd(t.data_ptr(), g.data_ptr(), topo.colptr.data_ptr(), topo.csc_perm.data_ptr(), topo.src.data_ptr(),
                          topo.dst.data_ptr(), topo.rowptr.data_ptr(), inv.data_ptr(), dt.data_ptr(), topo.N, _stream()),
          "msmp_g2_bwd")
    _count(1 if topo.N else 0)
    return dt


# ---- training criterion: summed squared error on float64 labels (train_helper.py:126) ----------------
def sse_fwd(pred, y, sse_hi_lo=None):
    """sum((pred.double() - y) ** 2) as a float64 0-dim tensor; ``sse_hi_lo`` (2 floats) also receives it as a float pair."""
    if not (pred.is_cuda and pred.dtype == torch.float32 and y.dtype == torch.float64 and pred.shape == y.shape
            and pred.is_contiguous() and y.is_contiguous() and y.device == pred.device):
        raise ValueError("sse_fwd: contiguous float32 predictions and float64 labels of the same shape on one CUDA device")
    n = pred.numel()
    nbytes = int(lib.msmp_sse_workspace(n))
    ws = torch.empty(nbytes // 8, dtype=torch.float64, device=pred.device)
    sse = torch.empty((), dtype=torch.float64, device=pred.device)
    check(lib.msmp_sse_fwd(pred.data_ptr(), y.data_ptr(), n, ws.data_ptr(), nbytes, sse.data_ptr(), _p(sse_hi_lo), _stream()),
          "msmp_sse_fwd")
    _count(2 if n else 1)
    return sse


def sse_bwd(pred, y, g):
    """d sse / d pred = 2 g (pred - y), evaluated in float64 and rounded once to float32; ``g``: float64 device scalar."""
    g = g.to(dtype=torch.float64, device=pred.device).contiguous()
    out = torch.empty_like(pred)
    check(lib.msmp_sse_bwd(pred.data_ptr(), y.data_ptr(), g.data_ptr(), pred.numel(), out.data_ptr(), _stream()), "msmp_sse_bwd")
    _count(1)
    return out


class _SSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, y, sse_hi_lo):
        ctx.save_for_backward(pred, y)
        return sse_fwd(pred, y, sse_hi_lo)

    @staticmethod
    def backward(ctx, g):
        pred, y = ctx.saved_tensors
        return sse_bwd(pred, y, g), None, None


def sse_loss(pred, y, sse_hi_lo=None):
    """Differentiable summed squared error (float64 scalar) of float32 predictions against float64 labels."""
    return _SSEFn.apply(pred, y, sse_hi_lo)


# ---- LEM gate kernels (tensors are contiguous [N,128] slices of the step buffers) -----------------
def lem_gate_z(G, z_prev, dt, gates_t, z_new):
    check(lib.msmp_lem_gate_z(G.data_ptr(), z_prev.data_ptr(), float(dt), gates_t.data_ptr(), z_new.data_ptr(),
                              z_prev.shape[0], _stream()), "msmp_lem_gate_z")
    _count(1)


def lem_gate_y(L, y_prev, gates_t, y_new):
    check(lib.msmp_lem_gate_y(L.data_ptr(), y_prev.data_ptr(), gates_t.data_ptr(), y_new.data_ptr(),
                              y_prev.shape[0], _stream()), "msmp_lem_gate_y")
    _count(1)


def lem_bwd_y(dy, gy_t, y_prev, gates_t, dt, dL_t, dG_t):
    check(lib.msmp_lem_bwd_y(dy.data_ptr(), _p(gy_t), y_prev.data_ptr(), gates_t.data_ptr(), float(dt),
                             dL_t.data_ptr(), dG_t.data_ptr(), y_prev.shape[0], _stream()), "msmp_lem_bwd_y")
    _count(1)


def lem_bwd_z(dz_tot, gz_t, z_prev, gates_t, dt, dG_t, dz):
    check(lib.msmp_lem_bwd_z(dz_tot.data_ptr(), _p(gz_t), z_prev.data_ptr(), gates_t.data_ptr(), float(dt),
                             dG_t.data_ptr(), dz.data_ptr(), z_prev.shape[0], _stream()), "msmp_lem_bwd_z")
    _count(1)


# ---- decoder ----------------------------------------------------------------------------------------
def decoder_fwd(h, w1, b1, w2, b2, u, dt, geom):
    """geom = (C, K1, S1, L1, K2, TW).  Returns (out [N, C*TW], za [N, 8*L1])."""
    C, K1, S1, L1, K2, TW = geom
    N = h.shape[0]
    out = torch.empty(N, C * TW, dtype=torch.float32, device=h.device)
    za = torch.empty(N, 8 * L1, dtype=torch.float32, device=h.device)
    check(lib.msmp_decoder_fwd(h.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                               u.data_ptr(), _ld(u), dt.data_ptr(), za.data_ptr(), out.data_ptr(), N, C, K1, S1, L1,
                               K2, TW, _stream()), "msmp_decoder_fwd")
    _count(1)
    return out, za


def decoder_bwd(dout, h, za, w1, w2, dt, geom, dW=None):
    """Returns (dh [N, C*128], dW flat [w1 | b1 | w2 | b2])."""
    C, K1, S1, L1, K2, TW = geom
    N = h.shape[0]
    dev = h.device
    dh = torch.empty_like(h)
    if dW is None:
        dW = torch.empty(lib.msmp_decoder_nweights(C, K1, K2), dtype=torch.float32, device=dev)
    ws = _workspace(lib.msmp_decoder_bwd_workspace(N, C, K1, K2), dev)
    check(lib.msmp_decoder_bwd(dout.data_ptr(), h.data_ptr(), za.data_ptr(), w1.data_ptr(), w2.data_ptr(),
                               dt.data_ptr(), dh.data_ptr(), dW.data_ptr(), N, C, K1, S1, L1, K2, TW, ws.data_ptr(),
                               ws.numel(), _stream()), "msmp_decoder_bwd")
    _count(2)
    return dh, dW


def to_lane_major(x: torch.Tensor, Npad: int) -> torch.Tensor:
    """[..., N, C] row-major -> [..., Npad/32, C, 32] lane-major (rows zero-padded to Npad)."""
    *lead, N, C = x.shape
    if N != Npad:
        xp = x.new_zeros(*lead, Npad, C)
        xp[..., :N, :] = x
        x = xp
    return x.reshape(*lead, Npad // 32, 32, C).transpose(-1, -2).contiguous()


def from_lane_major(x: torch.Tensor, N: int) -> torch.Tensor:
    """[..., Npad/32, C, 32] lane-major -> [..., N, C] row-major."""
    *lead, nt, C, _ = x.shape
    return x.transpose(-1, -2).reshape(*lead, nt * 32, C)[..., :N, :].contiguous()


def _img_of(w):
    return w.img if isinstance(w, TcW) else _cached_images(w)


LEM_TILE = 64         # nodes per CTA of the persistent LEM kernels (csrc/lem_tc.cu LT_NODES)


def lem_tc_fwd(inp, ninp, Wt_in, Wzt_in, Wt_h, Wzt_h, bias, bias_z, Y, Z, dt):
    """Persistent tensor-core LEM forward (input projection fused, all T steps, one launch).
    Wt_in [>= ninp, 384] / Wzt_in [>= ninp, 128]: input rows of the k-major W^T / Wz^T; Wt_h / Wzt_h: their state rows
    (TcW images or k-major tensors).  Y, Z: row-major [T+1, N, 128] with slab 0 set; slabs 1..T are written.
    Returns the gate activations [T, Npad, 512] kept for the backward."""
    T, N = inp.shape[0], inp.shape[1]
    Npad = (N + LEM_TILE - 1) // LEM_TILE * LEM_TILE
    dev = inp.device
    gates = torch.empty(T, Npad, 512, dtype=torch.float32, device=dev)
    with _timed("lem_tc_fwd", 2.0 * T * N * H * 4 * H, 4.0 * T * N * (512 + 4 * H + 8)):
      check(lib.msmp_lem_tc_fwd(inp.data_ptr(), int(ninp), Wt_in.data_ptr(), Wzt_in.data_ptr(), _img_of(Wt_h).data_ptr(),
                              _img_of(Wzt_h).data_ptr(), bias.data_ptr(), bias_z.data_ptr(),
                              Y.data_ptr(), Z.data_ptr(), gates.data_ptr(), float(dt), T, N, Npad,
                              1 if PRECISION == "bf16" else 0, _stream()),
            "msmp_lem_tc_fwd")
    _count(1)
    return gates


def lem_tc_bwd_state(gates):
    """(dy, dz): zeroed carried state gradients [Npad, 128] of one backward recurrence."""
    Npad, dev = gates.shape[1], gates.device
    return (torch.zeros(Npad, H, dtype=torch.float32, device=dev), torch.zeros(Npad, H, dtype=torch.float32, device=dev))


def lem_tc_bwd(Wzh, Wh, Y, Z, gates, gY, gZ, last_only, dG, dL, dt, N, state, t_begin=0, t_end=None):
    """Steps t_end-1 .. t_begin of the backward recurrence (state = lem_tc_bwd_state(); dy / dz carry the gradient
    between calls and hold d/d(y0, z0) in their first N rows after the call with t_begin = 0); fills dG[t], dL[t]."""
    T, Npad = gates.shape[0], gates.shape[1]
    t_end = T if t_end is None else t_end
    dy, dz = state
    nst = t_end - t_begin
    with _timed("lem_tc_bwd", 2.0 * nst * N * H * 4 * H, 4.0 * nst * N * (512 + 512 + 6 * H)):
      check(lib.msmp_lem_tc_bwd(_img_of(Wzh).data_ptr(), _img_of(Wh).data_ptr(), Y.data_ptr(), Z.data_ptr(),
                              gates.data_ptr(), _p(gY), _p(gZ), int(bool(last_only)), dG.data_ptr(), dL.data_ptr(),
                              dy.data_ptr(), dz.data_ptr(), float(dt), T, t_begin, t_end, N, Npad,
                              1 if PRECISION == "bf16" else 0, _stream()),
            "msmp_lem_tc_bwd")
    _count(1)
    return dy, dz

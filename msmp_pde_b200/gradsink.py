"""One-launch gradient unpacking for the captured training step (see csrc/pack.cu, k_unpack).

In plain autograd use every backward Function of this package returns parameter-layout gradients: the raw k-major
results of the weight-gradient kernels are transposed / concatenated / subtracted with framework ops and then added
into ``param.grad`` by AccumulateGrad -- several hundred tiny kernels per step, and a join of the weight-gradient
side stream at the end of every layer.  ``GradPlan`` removes all of it for ``GraphedTrainStep``: the kernels write
their raw results into one persistent buffer (a ``sink`` region per module), the Functions return ``None`` for the
parameters, and ONE launch at the end of the backward pass writes every ``param.grad`` (overwrite, so the gradients
need no zeroing either).  Valid while every parameter is used by exactly one Function call per backward pass, which
holds for all solver classes (GraphedTrainStep runs one forward per backward).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import check, lib
from .layers import H, SIDE_LD, _LayerBase, pad32
from .lem import LEMcuda

_UJOB_DTYPE = np.dtype([("dst", "<u8"), ("src0", "<u8"), ("src1", "<u8"), ("ldd", "<i4"), ("ld0", "<i4"),
                        ("ld1", "<i4"), ("rows", "<i4"), ("cols", "<i4"), ("sign1", "<f4"), ("zero", "<i4"),
                        ("nsplit", "<i4"), ("sstride0", "<i4"), ("sstride1", "<i4")], align=True)


class Sink:
    """Raw-gradient views of one module (attribute names = the local variable names of the backward Functions)."""


class GradPlan:
    """``n_nodes`` / ``n_edges`` / ``n_steps``: row counts of the step's weight-gradient GEMMs (nodes, edges, LEM steps x
    nodes).  With them the sink regions hold the split-M PARTIALS of msmp_wgrad_ws ([S, K, N], S fixed by the row count)
    and the unpack launch sums them; without them (None) every weight gradient is reduced by its own launch."""

    def __init__(self, model: nn.Module, n_nodes: int = None, n_edges: int = None, n_steps: int = None):
        assert _UJOB_DTYPE.itemsize == lib.msmp_unpack_job_bytes(), "UnpackJob layout mismatch"
        self.model = model
        self.rows = dict(node=n_nodes, edge=n_edges, lem=None if (n_nodes is None or n_steps is None) else n_nodes * n_steps)
        self.total = 0
        self.jobs = []            # (param, dst_off, ldd, src0_off, ld0, src1_off | None, ld1, rows, cols, sign1, zero, nsplit, ss0, ss1)
        self._late = []
        self.covered = []
        self.streams = set()      # side streams that received weight-gradient work during the current backward
        decoders = {id(m.output_mlp) for m in model.modules() if isinstance(getattr(m, "output_mlp", None), nn.Sequential)}
        for m in model.modules():
            if isinstance(m, _LayerBase):
                self._add_layer(m)
            elif isinstance(m, LEMcuda):
                self._add_lem(m)
            elif isinstance(m, nn.Linear) and "_msmp_tcw" in m.__dict__:
                self._add_linear(m)
            elif id(m) in decoders:
                self._add_decoder(m)
        self.device = next(model.parameters()).device
        self.raw = torch.zeros(max(self.total, 64), dtype=torch.float32, device=self.device)
        for fn in self._late:
            fn()
        self._late = []
        self._build_table()

    # ---- building -------------------------------------------------------------------------------------
    def _alloc(self, *shape) -> int:
        off = self.total
        self.total += (int(np.prod(shape)) + 63) // 64 * 64
        return off

    def _job(self, param, dst_off, ldd, src0, ld0, rows, cols, src1=None, ld1=0, sign1=0.0, zero=0, split=(1, 0)):
        nsplit, sstride = split
        self.jobs.append((param, dst_off, ldd, src0, ld0, src1, ld1, rows, cols, sign1, zero, nsplit, sstride, sstride))

    def _views(self, sink, **regions):
        def late():
            for name, (off, shape) in regions.items():
                setattr(sink, name, self.raw[off:off + int(np.prod(shape))].view(*shape))
        self._late.append(late)

    def _wg(self, kind, K, N, nside):
        """Sink regions of one weight-gradient GEMM -> (main offset, side offset, main shape, side shape, main split,
        side split): [S, K, N] / [S, nside, N] partials when the row count is known and msmp_wgrad_ws takes the shape."""
        M = self.rows[kind]
        S = 0
        if M is not None and ops.wgrad_use_ws(int(M), K, N, nside):
            S = lib.msmp_wgrad_ws_splits(int(M), K, N, nside)
        if S > 0:
            return (self._alloc(S, K, N), self._alloc(S, nside, N), (S, K, N), (S, nside, N), (S, K * N), (S, nside * N))
        return (self._alloc(K, N), self._alloc(nside, N), (K, N), (nside, N), (1, 0), (1, 0))

    def _add_layer(self, layer):
        W1, b1, W2, b2, W3, b3, W4, b4 = layer._params()
        F_u, V = layer.time_window, layer.n_variables
        K1, K3 = W1.shape[1], W3.shape[1]
        Kp = H + pad32(F_u)
        o_pq, o_s, sh_pq, sh_s, sp_pq, sp_s = self._wg("node", Kp, 2 * H, 2 + V)
        o_3, o_3s, sh_3, sh_3s, sp_3, sp_3s = self._wg("node", 2 * H, H, V + 1)
        o_4, o_4s, sh_4, sh_4s, sp_4, sp_4s = self._wg("node", H, H, 1)
        o_2, o_2s, sh_2, sh_2s, sp_2, sp_2s = self._wg("edge", H, H, 1)
        # message_net_1: [x_i | x_j | u_i - u_j | pos_i - pos_j | variables]  <-  P | Q factorisation
        self._job(W1, 0, K1, o_pq, 2 * H, H, H, split=sp_pq)
        self._job(W1, H, K1, o_pq + H, 2 * H, H, H, split=sp_pq)
        self._job(W1, 2 * H, K1, o_pq + H * 2 * H, 2 * H, H, F_u, src1=o_pq + H * 2 * H + H, ld1=2 * H, sign1=-1.0,
                  split=sp_pq)
        self._job(W1, 2 * H + F_u, K1, o_s, 2 * H, H, 1, src1=o_s + H, ld1=2 * H, sign1=-1.0, split=sp_s)
        self._job(W1, 2 * H + F_u + 1, K1, o_s + 2 * H, 2 * H, H, V, split=sp_s)
        self._job(b1, 0, 1, o_s + (1 + V) * 2 * H, 2 * H, H, 1, split=sp_s)
        self._job(W2, 0, H, o_2, H, H, H, split=sp_2)
        self._job(b2, 0, 1, o_2s, H, H, 1, split=sp_2s)
        self._job(W3, 0, K3, o_3, H, H, 2 * H, split=sp_3)
        self._job(W3, 2 * H, K3, o_3s, H, H, V, split=sp_3s)
        self._job(b3, 0, 1, o_3s + V * H, H, H, 1, split=sp_3s)
        self._job(W4, 0, H, o_4, H, H, H, split=sp_4)
        # GNN_LayerLin: b4 feeds a non-affine InstanceNorm directly, its gradient is identically zero
        self._job(b4, 0, 1, o_4s, H, H, 1, zero=0 if layer.final_swish else 1, split=sp_4s)
        sink = Sink()
        self._views(sink, dWpq_t=(o_pq, sh_pq), dWs=(o_s, sh_s), dW3t=(o_3, sh_3), dW3s=(o_3s, sh_3s),
                    dW4t=(o_4, sh_4), dW4s=(o_4s, sh_4s), dW2t=(o_2, sh_2), db2s=(o_2s, sh_2s))
        layer.__dict__["_msmp_gsink"] = sink
        self.covered += [W1, b1, W2, b2, W3, b3, W4, b4]

    def _add_lem(self, rnn):
        ninp, Kp = rnn.ninp, H + pad32(rnn.ninp)
        o_w, o_b, sh_w, sh_b, sp_w, sp_b = self._wg("lem", Kp, 3 * H, 1)
        o_wz, o_bz, sh_wz, sh_bz, sp_wz, sp_bz = self._wg("lem", Kp, H, 1)
        self._job(rnn.weights, 0, H + ninp, o_w, 3 * H, 3 * H, H + ninp, split=sp_w)
        self._job(rnn.weights_lin_z, 0, H + ninp, o_wz, H, H, H + ninp, split=sp_wz)
        self._job(rnn.bias, 0, 1, o_b, 3 * H, 3 * H, 1, split=sp_b)
        self._job(rnn.bias_lin_z, 0, 1, o_bz, H, H, 1, split=sp_bz)
        sink = Sink()
        self._views(sink, dWt=(o_w, sh_w), dWzt=(o_wz, sh_wz), dbias=(o_b, sh_b), dbz=(o_bz, sh_bz))
        rnn.__dict__["_msmp_gsink"] = sink
        self.covered += [rnn.weights, rnn.weights_lin_z, rnn.bias, rnn.bias_lin_z]

    def _add_linear(self, lin):
        Nout, K = lin.weight.shape
        Kp = pad32(K)
        o_w, o_b, sh_w, sh_b, sp_w, sp_b = self._wg("node", Kp, Nout, 1)
        self._job(lin.weight, 0, K, o_w, Nout, Nout, K, split=sp_w)
        self._job(lin.bias, 0, 1, o_b, Nout, Nout, 1, split=sp_b)
        sink = Sink()
        self._views(sink, dWt=(o_w, sh_w), dbs=(o_b, sh_b))
        lin.__dict__["_msmp_gsink"] = sink
        self.covered += [lin.weight, lin.bias]

    def _add_decoder(self, seq):
        c1, c2 = seq[0], seq[2]
        sizes = [c1.weight.numel(), c1.bias.numel(), c2.weight.numel(), c2.bias.numel()]
        o = self._alloc(sum(sizes))
        off = o
        for p, n in zip((c1.weight, c1.bias, c2.weight, c2.bias), sizes):
            self._job(p, 0, 1, off, n, n, 1)          # flat copy: dst[i] = src[i]
            off += n
        sink = Sink()
        self._views(sink, dW=(o, (sum(sizes),)))
        seq.__dict__["_msmp_gsink"] = sink
        self.covered += [c1.weight, c1.bias, c2.weight, c2.bias]

    def _build_table(self):
        base = self.raw.data_ptr()
        rows = []
        for (p, dst_off, ldd, s0, ld0, s1, ld1, r, c, sg, z, ns, ss0, ss1) in self.jobs:
            if p.grad is None or not p.grad.is_contiguous() or p.grad.dtype != torch.float32:
                raise RuntimeError("GradPlan needs contiguous fp32 .grad buffers on every covered parameter")
            rows.append((p.grad.data_ptr() + 4 * dst_off, base + 4 * s0, 0 if s1 is None else base + 4 * s1, ldd, ld0,
                         ld1, r, c, sg, z, ns, ss0, ss1))
        arr = np.array(rows, dtype=_UJOB_DTYPE)
        self.njobs = len(rows)
        self.max_tiles = int(max(((r + 31) // 32) * ((c + 31) // 32) for (_, _, _, _, _, _, _, r, c, *_rest) in self.jobs))
        self.jobs_dev = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(self.device)
        self.grad_key = tuple(p.grad.data_ptr() for p in self.covered)
        ids = {id(p) for p in self.covered}
        self.uncovered = [p for p in self.model.parameters() if p.requires_grad and id(p) not in ids]

    # ---- use ------------------------------------------------------------------------------------------
    def valid(self) -> bool:
        return all(p.grad is not None for p in self.covered) and \
            tuple(p.grad.data_ptr() for p in self.covered) == self.grad_key

    def begin(self) -> None:
        """Start of a backward pass whose Functions write into the sinks."""
        if not self.valid():
            self._build_table()
        for p in self.uncovered:          # parameters of framework modules (e.g. the cuDNN LSTM encoder) accumulate
            if p.grad is not None:
                p.grad.zero_()
        self.streams.clear()
        ops.GRAD_SINK = self

    def finish(self) -> None:
        """Join the weight-gradient streams and write every covered ``param.grad`` with one launch."""
        ops.GRAD_SINK = None
        cur = torch.cuda.current_stream()
        for st in self.streams:
            cur.wait_stream(st)
        self.streams.clear()
        check(lib.msmp_unpack_run(self.jobs_dev.data_ptr(), self.njobs, self.max_tiles, cur.cuda_stream),
              "msmp_unpack_run")
        ops._count(1)

"""One-launch weight packing for the tensor-core kernels (see csrc/pack.cu).

The parameters keep the reference's names/layouts; the kernels want pre-split (tf32 hi | lo), pre-swizzled tile
images of the k-major weights plus a few small side/bias arrays.  ``PackPlan`` owns one persistent device buffer with
every packed weight of a model (message-passing layers, LEM, encoder/decoder MLPs) and a device-resident job table;
``refresh()`` re-packs everything with ONE kernel launch.  It runs at the start of every forward pass: parameter
``_version`` counters cannot be trusted to detect updates (fused optimizers and CUDA-graph replays do not move them).
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib

H = 128
IMG = 4096            # floats per [128 x 32] tile image


class TcW:
    """Handle of a packed weight: ``img`` [ntiles, nchunks, 2, 4096] of the k-major Wt[K, N]."""
    __slots__ = ("img", "K", "N")

    def __init__(self, img, K, N):
        self.img, self.K, self.N = img, K, N


def pad32(n: int) -> int:
    return (n + 31) // 32 * 32


_JOB_DTYPE = np.dtype([("src", "<u8"), ("dst", "<u8"), ("ld", "<i4"), ("transpose", "<i4"), ("sign", "<f4"),
                       ("kvalid", "<i4"), ("nvalid", "<i4"), ("nchunks", "<i4"), ("kind", "<i4"), ("ldd", "<i4")],
                      align=True)


class LayerPackTC:
    """Packed weights of one message-passing layer (same attribute names as layers.LayerPack)."""
    __slots__ = ("Wpq_t", "Wpq_side", "bias_pq", "W2t", "W2d", "W3t", "W3side", "W3hx", "W4t", "W4d", "W1hq", "plan")


class LemPackTC:
    __slots__ = ("Wt_h", "Wzt_h", "Wh", "Wzh", "Wt_in", "Wzt_in")


class LinearPackTC:
    __slots__ = ("fwd", "dgrad")


class PackPlan:
    generation = 0

    def __init__(self, device):
        assert _JOB_DTYPE.itemsize == lib.msmp_pack_job_bytes(), "PackJob layout mismatch"
        self.device = device
        self.total = 0
        self.jobs = []            # (src_ptr, dst_offset_floats, ld, transpose, sign, kvalid, nvalid, nchunks, kind, ldd)
        self.params = []
        self._late = []           # callables run after the buffer exists
        self.buf = None

    # ---- building -------------------------------------------------------------------------------------
    def _alloc(self, nfloats: int) -> int:
        off = self.total
        self.total += (nfloats + 63) // 64 * 64          # keep every block 256-byte aligned
        return off

    def _img(self, off, src_ptr, ld, transpose, sign, kvalid, nvalid, nchunks):
        self.jobs.append((src_ptr, off, ld, transpose, sign, kvalid, nvalid, nchunks, 0, 0))

    def _plain(self, off, src_ptr, ld, sign, rows, nvalid, ldd):
        self.jobs.append((src_ptr, off, ld, 1, sign, rows, nvalid, 0, 1, ldd))

    def _view(self, off, *shape):
        return self.buf[off:off + int(np.prod(shape))].view(*shape)

    def add_layer(self, layer) -> LayerPackTC:
        W1, b1, W2, b2, W3, b3, W4, b4 = layer._params()
        self.params += [W1, b1, W2, W3, W4]
        F_u, V = layer.time_window, layer.n_variables
        K1, K3 = W1.shape[1], W3.shape[1]
        Kp = H + pad32(F_u)
        nc_pq, fp = Kp // 32, pad32(F_u) // 32
        S = lambda t, col=0: t.data_ptr() + 4 * col
        tile = nc_pq * 2 * IMG
        o_pq = self._alloc(2 * tile)
        # P | Q projection: Wt[k][n], k over [h(128) | u(F_u, zero padded)], n over [P(128) | Q(128)]
        self._img(o_pq, S(W1, 0), K1, 1, 1.0, H, H, 4)
        self._img(o_pq + tile, S(W1, H), K1, 1, 1.0, H, H, 4)
        self._img(o_pq + 4 * 2 * IMG, S(W1, 2 * H), K1, 1, 1.0, F_u, H, fp)
        self._img(o_pq + tile + 4 * 2 * IMG, S(W1, 2 * H), K1, 1, -1.0, F_u, H, fp)
        o_w2t, o_w2d = self._alloc(8 * IMG), self._alloc(8 * IMG)
        self._img(o_w2t, S(W2), H, 1, 1.0, H, H, 4)                       # Wt[k][n] = W2[n][k]
        self._img(o_w2d, S(W2), H, 0, 1.0, H, H, 4)                       # Wt[k][n] = W2[k][n] (edge backward)
        o_w3t = self._alloc(16 * IMG)
        self._img(o_w3t, S(W3), K3, 1, 1.0, 2 * H, H, 8)                  # [h | agg] part of update_net_1
        o_w3hx = self._alloc(16 * IMG)
        self._img(o_w3hx, S(W3, 0), K3, 0, 1.0, H, H, 4)                  # dgrad operand, n-tile 0 (d/dh)
        self._img(o_w3hx + 8 * IMG, S(W3, H), K3, 0, 1.0, H, H, 4)        # n-tile 1 (d/dagg)
        o_w4t, o_w4d = self._alloc(8 * IMG), self._alloc(8 * IMG)
        self._img(o_w4t, S(W4), H, 1, 1.0, H, H, 4)
        self._img(o_w4d, S(W4), H, 0, 1.0, H, H, 4)
        o_w1hq = self._alloc(16 * IMG)
        self._img(o_w1hq, S(W1, 0), K1, 0, 1.0, H, H, 4)                  # dgrad operand [W1xi ; W1xj]
        self._img(o_w1hq + 8 * IMG, S(W1, H), K1, 0, 1.0, H, H, 4)
        # side rows [pos | v...] of the P|Q projection, bias, variable rows of update_net_1
        o_side, o_bias, o_w3s = self._alloc(8 * 2 * H), self._alloc(2 * H), self._alloc(8 * H)
        self._plain(o_side, S(W1, 2 * H + F_u), K1, 1.0, 1, H, 2 * H)
        self._plain(o_side + H, S(W1, 2 * H + F_u), K1, -1.0, 1, H, 2 * H)
        self._plain(o_side + 2 * H, S(W1, 2 * H + F_u + 1), K1, 1.0, V, H, 2 * H)
        self._plain(o_bias, S(b1), 1, 1.0, 1, H, 2 * H)
        self._plain(o_w3s, S(W3, 2 * H), K3, 1.0, V, H, H)
        pk = LayerPackTC()
        pk.plan = self

        def late():
            pk.Wpq_t = TcW(self._view(o_pq, 2, nc_pq, 2, IMG), Kp, 2 * H)
            pk.Wpq_side = self._view(o_side, 8, 2 * H)
            pk.bias_pq = self._view(o_bias, 2 * H)
            pk.W2t = TcW(self._view(o_w2t, 1, 4, 2, IMG), H, H)
            pk.W2d = TcW(self._view(o_w2d, 1, 4, 2, IMG), H, H)
            pk.W3t = TcW(self._view(o_w3t, 1, 8, 2, IMG), 2 * H, H)
            pk.W3side = self._view(o_w3s, 8, H)
            pk.W3hx = TcW(self._view(o_w3hx, 2, 4, 2, IMG), H, 2 * H)
            pk.W4t = TcW(self._view(o_w4t, 1, 4, 2, IMG), H, H)
            pk.W4d = TcW(self._view(o_w4d, 1, 4, 2, IMG), H, H)
            pk.W1hq = TcW(self._view(o_w1hq, 1, 8, 2, IMG), 2 * H, H)
        self._late.append(late)
        return pk

    def add_lem(self, rnn) -> LemPackTC:
        """rnn: lem.LEMcuda.  Packs for the persistent tensor-core LEM kernels."""
        W, Wz = rnn.weights, rnn.weights_lin_z
        self.params += [W, Wz]
        ninp = rnn.ninp
        ld = H + ninp
        S = lambda t, col=0, row=0: t.data_ptr() + 4 * (row * ld + col)
        o_wt = self._alloc(3 * 4 * 2 * IMG)
        for t in range(3):                                                # Wt[k][n] = W[128 t + n][k], k < 128
            self._img(o_wt + t * 4 * 2 * IMG, S(W, 0, 128 * t), ld, 1, 1.0, H, H, 4)
        o_wzt = self._alloc(4 * 2 * IMG)
        self._img(o_wzt, S(Wz), ld, 1, 1.0, H, H, 4)
        o_wh = self._alloc(12 * 2 * IMG)
        self._img(o_wh, S(W), ld, 0, 1.0, 3 * H, H, 12)                   # Wt[k][n] = W[k][n], k < 384 (dgrad)
        o_wzh = self._alloc(4 * 2 * IMG)
        self._img(o_wzh, S(Wz), ld, 0, 1.0, H, H, 4)
        o_in, o_zin = self._alloc(8 * 3 * H), self._alloc(8 * H)
        self._plain(o_in, S(W, H), ld, 1.0, ninp, 3 * H, 3 * H)           # Wt_in[q][n] = W[n][128 + q]
        self._plain(o_zin, S(Wz, H), ld, 1.0, ninp, H, H)
        pk = LemPackTC()

        def late():
            pk.Wt_h = TcW(self._view(o_wt, 3, 4, 2, IMG), H, 3 * H)
            pk.Wzt_h = TcW(self._view(o_wzt, 1, 4, 2, IMG), H, H)
            pk.Wh = TcW(self._view(o_wh, 1, 12, 2, IMG), 3 * H, H)
            pk.Wzh = TcW(self._view(o_wzh, 1, 4, 2, IMG), H, H)
            pk.Wt_in = self._view(o_in, 8, 3 * H)
            pk.Wzt_in = self._view(o_zin, 8, H)
        self._late.append(late)
        return pk

    def add_linear(self, linear) -> LinearPackTC:
        """nn.Linear [Nout, K]: forward images of W^T (K zero-padded to 32) and dgrad images of W (when K % 32 == 0)."""
        W = linear.weight
        self.params.append(W)
        Nout, K = W.shape
        Kp, nt = pad32(K), (Nout + 127) // 128
        o_f = self._alloc(nt * (Kp // 32) * 2 * IMG)
        for t in range(nt):
            self._img(o_f + t * (Kp // 32) * 2 * IMG, W.data_ptr() + 4 * (128 * t * K), K, 1, 1.0, K, min(128, Nout - 128 * t),
                      Kp // 32)
        has_d = (K % 32 == 0) and (Nout % 32 == 0)
        o_d = None
        if has_d:                                                         # Wt := W  ([K' = Nout][N' = K])
            ntd = (K + 127) // 128
            o_d = self._alloc(ntd * (Nout // 32) * 2 * IMG)
            for t in range(ntd):
                self._img(o_d + t * (Nout // 32) * 2 * IMG, W.data_ptr() + 4 * (128 * t), K, 0, 1.0, Nout,
                          min(128, K - 128 * t), Nout // 32)
        pk = LinearPackTC()

        def late():
            pk.fwd = TcW(self._view(o_f, nt, Kp // 32, 2, IMG), Kp, Nout)
            pk.dgrad = TcW(self._view(o_d, (K + 127) // 128, Nout // 32, 2, IMG), Nout, K) if has_d else None
        self._late.append(late)
        return pk

    def finalize(self):
        self.buf = torch.zeros(max(self.total, 64), dtype=torch.float32, device=self.device)
        base = self.buf.data_ptr()
        arr = np.array([(s, base + 4 * off, ld, tr, sg, kv, nv, nc, kd, ldd)
                        for (s, off, ld, tr, sg, kv, nv, nc, kd, ldd) in self.jobs], dtype=_JOB_DTYPE)
        self.njobs = len(self.jobs)
        self.max_chunks = int(max(1, arr["nchunks"].max())) if self.njobs else 1
        self.jobs_dev = torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(self.device)
        for fn in self._late:
            fn()
        self._late = []
        self.ptr_key = tuple(p.data_ptr() for p in self.params)
        return self

    # ---- use ------------------------------------------------------------------------------------------
    def valid(self) -> bool:
        return tuple(p.data_ptr() for p in self.params) == self.ptr_key

    def refresh(self) -> None:
        """Re-pack every weight from the current parameter values (one launch).  ``generation`` counts the refreshes: the
        backward Functions compare it with the value their forward saw (the dgrad operands are views into this one
        buffer, so a backward that runs after a LATER forward's refresh would mix old activations with new weights)."""
        self.generation = getattr(self, "generation", 0) + 1
        check(lib.msmp_pack_run(self.jobs_dev.data_ptr(), self.njobs, self.max_chunks,
                                torch.cuda.current_stream().cuda_stream), "msmp_pack_run")
        from . import ops
        ops._count(1)

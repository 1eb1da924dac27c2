"""The optimizer update of the captured training step: ``torch.optim.AdamW`` semantics in one graph-capturable launch.

The reference's scripts build ``optim.AdamW(model.parameters(), lr=args.lr)`` and a ``MultiStepLR`` scheduler that rewrites
``param_group['lr']`` (experiments/train.py:410-411,437).  ``FusedAdamW`` drives that very optimizer object: it updates the
optimizer's own ``exp_avg`` / ``exp_avg_sq`` state in place with ``msmp_adamw_run`` (csrc/optim.cu) and feeds lr, betas,
eps, weight decay and the bias corrections through a small device array refreshed before every launch, so

* the launch can be captured in a CUDA graph although the optimizer is not ``capturable``,
* schedulers (and manual edits of the param groups) take effect on the next step,
* ``optimizer.state_dict()`` and a later eager ``optimizer.step()`` see consistent state (step counters are written back
  lazily through the optimizer's own pre-hooks).
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib

_JOB = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i4"), ("group", "<i4")], align=True)


def supported(optimizer) -> bool:
    """AdamW without amsgrad / maximize on fp32 CUDA parameters."""
    if type(optimizer) is not torch.optim.AdamW:
        return False
    for g in optimizer.param_groups:
        if g.get("amsgrad", False) or g.get("maximize", False):
            return False
        for p in g["params"]:
            if p.requires_grad and (p.dtype != torch.float32 or not p.is_cuda):
                return False
    return True


class FusedAdamW:
    def __init__(self, optimizer):
        if not supported(optimizer):
            raise ValueError("FusedAdamW drives torch.optim.AdamW (no amsgrad, no maximize) on fp32 CUDA parameters")
        assert _JOB.itemsize == lib.msmp_adamw_job_bytes(), "AdamJob layout mismatch"
        self.opt = optimizer
        self.nh = lib.msmp_adamw_hyper_floats()
        chunk = lib.msmp_adamw_chunk()
        rows, chunks = [], []
        self._steps = []            # per group: [python step count, [state dicts]]
        self._refs = []             # (param, grad, state dict, exp_avg, exp_avg_sq): identity = the job table is current
        dev = None
        for gi, g in enumerate(optimizer.param_groups):
            states, t0 = [], None
            for p in g["params"]:
                if not p.requires_grad:
                    continue
                if p.grad is None or not p.grad.is_contiguous() or not p.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous parameters with allocated contiguous .grad buffers")
                dev = p.device
                st = optimizer.state[p]
                if len(st) == 0:          # what torch.optim.AdamW._init_group creates on the first step
                    on_dev = bool(g.get("capturable", False) or g.get("fused", False))
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device if on_dev else "cpu")
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                t = int(round(float(st["step"])))
                t0 = t if t0 is None else t0
                if t != t0:
                    raise RuntimeError("FusedAdamW: parameters of one group carry different step counts")
                states.append(st)
                self._refs.append((p, p.grad, st, st["exp_avg"], st["exp_avg_sq"]))
                job = len(rows)
                rows.append((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                             p.numel(), gi))
                chunks += [(job, off) for off in range(0, p.numel(), chunk)]
            self._steps.append([t0 or 0, states])
        self.device = dev
        self.jobs = torch.from_numpy(np.array(rows, dtype=_JOB).view(np.uint8).reshape(-1).copy()).to(dev)
        self.chunks = torch.tensor(chunks, dtype=torch.int32).to(dev)
        self.nchunks = len(chunks)
        self.hyper = torch.zeros(len(optimizer.param_groups) * self.nh, dtype=torch.float32, device=dev)
        self._dirty = False
        optimizer.register_step_pre_hook(lambda *a, **k: self.flush_steps())
        if hasattr(optimizer, "register_state_dict_pre_hook"):
            optimizer.register_state_dict_pre_hook(lambda *a, **k: self.flush_steps())

    def valid(self) -> bool:
        """The tensors the job table points at are still the optimizer's / the parameters' own."""
        state = self.opt.state
        for p, g, st, m, v in self._refs:
            if p.grad is not g or state.get(p) is not st or st.get("exp_avg") is not m or st.get("exp_avg_sq") is not v:
                return False
        return True

    def resync(self) -> None:
        """Re-read the step counters from the optimizer's state (after the state was restored / edited in place)."""
        for ent in self._steps:
            if ent[1]:
                ent[0] = int(round(float(ent[1][0]["step"])))
        self._dirty = False

    def host_update(self) -> None:
        """Advance the step counters and refresh the device-side hyper-parameters (call once per step, BEFORE the launch
        or the replay of the graph that contains it; the copy is stream-ordered)."""
        h = np.zeros(len(self.opt.param_groups) * self.nh, dtype=np.float32)
        for gi, g in enumerate(self.opt.param_groups):
            self._steps[gi][0] += 1
            t = self._steps[gi][0]
            b1, b2 = g["betas"]
            lr = g["lr"]
            lr = float(lr) if not torch.is_tensor(lr) else float(lr.item())
            h[gi * self.nh:gi * self.nh + 8] = (lr, 1.0 - b1, b2, g["eps"], g["weight_decay"], 1.0 - b1 ** t,
                                                (1.0 - b2 ** t) ** 0.5, 1.0 - b2)
        self.hyper.copy_(torch.from_numpy(h), non_blocking=True)       # pageable source: staged before the call returns
        self._dirty = True

    def launch(self, gscale=None) -> None:
        """The update itself (capturable).  gscale: fp32 device scalar multiplied into every gradient first."""
        check(lib.msmp_adamw_run(self.jobs.data_ptr(), self.chunks.data_ptr(), self.nchunks, self.hyper.data_ptr(),
                                 0 if gscale is None else gscale.data_ptr(), torch.cuda.current_stream().cuda_stream),
              "msmp_adamw_run")
        from . import ops
        ops._count(1)

    def flush_steps(self) -> None:
        """Write the step counters into the optimizer's state (before state_dict() / an eager optimizer.step())."""
        if not self._dirty:
            return
        for t, states in self._steps:
            for st in states:
                st["step"].fill_(float(t))
        self._dirty = False

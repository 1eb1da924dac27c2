"""``lem_cuda`` -- the reference's one native extension (upstream tk-rusch/LEM ``src/lem_cuda``, a pybind11 module that is
neither in the reference tree nor in its environment recipe) re-implemented on ``libmsmp_b200.so``.

Same two entry points, argument order and return arity as the call sites in experiments/models_gnn.py:287-302::

    all_y, all_z, all_X, all_X2, all_multi_scales, all_lin_new_z_state = lem_cuda.forward(
        inputs[T,N,ninp], weights[3H,H+ninp], weights_lin_z[H,H+ninp], bias[3H], bias_lin_z[H], y0[N,H], z0[N,H], dt[1,1])
    d_inputs, d_weights, d_weights_lin_z, d_bias, d_bias_lin_z, d_y0, d_z0 = lem_cuda.backward(
        grad_y[T,N,H], grad_z[T,N,H], all_X, all_X2, all_multi_scales, all_lin_new_z_state,
        weights, weights_lin_z, bias, bias_lin_z, y0, z0, dt)

``msmp_pde_b200.install()`` registers this module as ``lem_cuda`` when no real one is importable, so the reference's own
``LEMFunction`` / ``LEMcuda`` / ``LEM`` / ``LEMS`` classes run unchanged on the persistent tensor-core kernels
(``msmp_lem_tc_fwd`` / ``msmp_lem_tc_bwd``, all T steps per launch).  The four "saved" tensors are opaque to the caller
(the reference only passes them from forward to backward); here they hold the state histories, the gate activations
and the zero-padded inputs.  Tensors may be float64 (the reference's default dtype): they are cast to fp32 once, results
are returned in the callers' dtypes.  ``d_inputs`` is ``None``: the reference discards it (models_gnn.py:302).  CUDA only.
"""
from __future__ import annotations

import torch

from .. import lem as _lem

H = _lem.H


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


def forward(inputs, weights, weights_lin_z, bias, bias_lin_z, initial_y_state, initial_z_state, dt):
    if not inputs.is_cuda:
        raise RuntimeError("lem_cuda (msmp_pde_b200): CUDA tensors only, there is no CPU path")
    T, N, ninp = inputs.shape
    if weights.shape != (3 * H, ninp + H) or weights_lin_z.shape != (H, ninp + H):
        raise ValueError("lem_cuda (msmp_pde_b200): nhid must be 128 and weights [3*nhid, ninp + nhid]")
    packs = _lem.make_packs(_f32(weights), _f32(weights_lin_z), ninp)
    inp, Y, Z, gates, _ = _lem.lem_forward(_f32(inputs), _f32(bias), _f32(bias_lin_z), _f32(initial_y_state),
                                           _f32(initial_z_state), float(dt), packs)
    od = inputs.dtype
    return Y[1:].to(od), Z[1:].to(od), Y, Z, gates, inp


def backward(grad_y_states, grad_z_states, all_X, all_X2, all_multi_scales, all_lin_new_z_state, weights, weights_lin_z,
             bias, bias_lin_z, initial_y_state, initial_z_state, dt):
    Y, Z, gates, inp = all_X, all_X2, all_multi_scales, all_lin_new_z_state
    ninp = weights.shape[1] - H
    packs = _lem.make_packs(_f32(weights), _f32(weights_lin_z), ninp)
    dWt, dWzt, dbias, dbz, dy, dz, _ = _lem.lem_backward(inp, Y, Z, gates, _f32(grad_y_states), _f32(grad_z_states),
                                                          float(dt), packs, _lem.use_persistent(ninp), False)
    wd = weights.dtype
    dW = torch.cat([dWt[:H].t(), dWt[H:H + ninp].t()], 1).to(wd)
    dWz = torch.cat([dWzt[:H].t(), dWzt[H:H + ninp].t()], 1).to(weights_lin_z.dtype)
    return (None, dW, dWz, dbias[0].to(bias.dtype), dbz[0].to(bias_lin_z.dtype), dy.to(initial_y_state.dtype),
            dz.to(initial_z_state.dtype))

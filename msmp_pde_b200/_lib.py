"""ctypes binding of ``csrc/libmsmp_b200.so`` (the C ABI declared in ``include/msmp_b200.h``).

There is NO fallback: if the shared library is missing or a symbol is absent, importing fails loudly;
if a call returns a non-zero status, ``MsmpError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_float, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# MSMP_B200_LIB: A/B runs of another build of the same C ABI (scripts/), never a different implementation
LIB_PATH = os.environ.get("MSMP_B200_LIB") or os.path.join(_HERE, "csrc", "libmsmp_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "msmp_b200.h")


class MsmpError(RuntimeError):
    pass


_ERR = {-1: "bad argument", -2: "CUDA launch/runtime error", -3: "workspace too small"}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(msmp_pde_b200 has no CPU or eager fallback)")
    return ctypes.CDLL(LIB_PATH)


lib = _load()

P, I, S, F = c_void_p, c_int, c_size_t, c_float
_SIGS = {
    "msmp_abi_version": (I, []),
    "msmp_linear_fwd": (I, [P, P, P, P, I, P, I, P, P, I, I, P, P, I, P, I, I, P, I, P, I, I, I, P]),
    "msmp_linear_tc_image_floats": (S, [I, I]),
    "msmp_linear_tc_fwd": (I, [P, P, P, P, I, P, P, P, I, I, P, I, P, I, P, I, I, P, I, P, I, I, I, I, P]),
    "msmp_pack_job_bytes": (I, []),
    "msmp_pack_run": (I, [P, I, I, P]),
    "msmp_unpack_job_bytes": (I, []),
    "msmp_unpack_run": (I, [P, I, I, P]),
    "msmp_linear_wgrad_splits": (I, [I, I, I]),
    "msmp_linear_wgrad_workspace": (S, [I, I, I, I]),
    "msmp_linear_wgrad": (I, [P, I, I, I, P, I, I, P, I, I, I, P, P, I, I, P, S, P]),
    "msmp_linear_wgrad_tc": (I, [P, I, I, I, P, I, I, P, I, I, I, P, P, I, I, P, S, P]),
    "msmp_linear_wgrad_tc2": (I, [P, I, I, P, I, I, I, P, I, I, P, I, I, I, P, P, I, I, P, S, P]),
    "msmp_wgrad_ws_splits": (I, [I, I, I, I]),
    "msmp_wgrad_ws_workspace": (S, [I, I, I, I]),
    "msmp_wgrad_ws": (I, [P, P, P, P, I, P, I, I, P, I, I, I, I, P, P, P, P, I, I, I, P]),
    "msmp_edge_tiles": (I, [I]),
    "msmp_edge_grid": (I, [I]),
    "msmp_edge_fwd_workspace": (S, [I]),
    "msmp_edge_fwd": (I, [P, P, I, P, P, P, P, P, P, P, P, I, I, P, S, P]),
    "msmp_edge_bwd_workspace": (S, [I]),
    "msmp_edge_bwd": (I, [P, P, I, P, P, P, P, P, P, P, I, P, P, I, P, P, I, I, P, S, P]),
    "msmp_edge_tc_fwd": (I, [P, P, I, P, P, P, P, P, P, P, P, I, I, P, S, P]),
    "msmp_edge_tc_bwd": (I, [P, P, I, P, P, P, P, P, P, P, I, P, P, P, P, I, I, I, P, S, P]),
    "msmp_edge_ws_workspace": (S, [I]),
    "msmp_edge_ws_fwd": (I, [P, P, I, P, P, P, P, P, I, I, P, P, P, I, I, I, P, S, P]),
    "msmp_edge_ws_bwd": (I, [P, P, I, P, P, P, P, P, I, I, P, P, I, P, P, P, P, I, I, I, I, P, S, P]),
    "msmp_segment_reduce": (I, [P, I, P, P, P, P, I, I, P]),
    "msmp_instnorm_workspace": (S, [I, I]),
    "msmp_instnorm_fwd": (I, [P, P, I, P, P, P, P, P, I, I, I, I, F, P, P, P, S, P]),
    "msmp_instnorm_bwd": (I, [P, P, P, I, P, P, P, P, P, P, I, I, I, I, P, P, I, P, P, S, P]),
    "msmp_instnorm1_fwd": (I, [P, P, I, P, P, P, I, I, I, F, P, P, P]),
    "msmp_instnorm1_bwd": (I, [P, P, P, I, P, P, P, P, I, I, I, P, P, I, P, P]),
    "msmp_mul_dswish": (I, [P, P, P, S, P]),
    "msmp_decoder_nweights": (I, [I, I, I]),
    "msmp_decoder_bwd_workspace": (S, [I, I, I, I]),
    "msmp_decoder_fwd": (I, [P, P, P, P, P, P, I, P, P, P, I, I, I, I, I, I, I, P]),
    "msmp_decoder_bwd": (I, [P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, P, S, P]),
    "msmp_lem_tc_fwd": (I, [P, I, P, P, P, P, P, P, P, P, P, F, I, I, I, I, P]),
    "msmp_lem_tc_bwd": (I, [P, P, P, P, P, P, P, I, P, P, P, P, F, I, I, I, I, I, I, P]),
    "msmp_adamw_job_bytes": (I, []),
    "msmp_adamw_chunk": (I, []),
    "msmp_adamw_hyper_floats": (I, []),
    "msmp_adamw_run": (I, [P, P, I, P, P, P]),
    "msmp_loss_scalars": (I, [P, P, P, P]),
    "msmp_lem_inputs": (I, [P, I, P, P, I, I, P, P]),
    "msmp_node_features": (I, [P, I, P, P, I, I, P, I, P, P]),
    "msmp_g2_fwd": (I, [P, P, P, P, P, P, I, P]),
    "msmp_g2_bwd": (I, [P, P, P, P, P, P, P, P, P, I, P]),
    "msmp_sse_workspace": (S, [S]),
    "msmp_sse_fwd": (I, [P, P, S, P, S, P, P, P]),
    "msmp_sse_bwd": (I, [P, P, P, S, P, P]),
    "msmp_lem_gate_z": (I, [P, P, F, P, P, I, P]),
    "msmp_lem_gate_y": (I, [P, P, P, P, I, P]),
    "msmp_lem_bwd_y": (I, [P, P, P, P, F, P, P, I, P]),
    "msmp_lem_bwd_z": (I, [P, P, P, P, F, P, P, I, P]),
}


def declared_symbols() -> list[str]:
    """Every function name declared in include/msmp_b200.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msmp_[a-z0-9_]+)\s*\(", text)))


for _name, (_res, _args) in _SIGS.items():
    try:
        _fn = getattr(lib, _name)
    except AttributeError as e:
        raise ImportError(f"libmsmp_b200.so does not export {_name}") from e
    _fn.restype = _res
    _fn.argtypes = _args


def check(status: int, what: str) -> None:
    if status != 0:
        raise MsmpError(f"{what} failed: {_ERR.get(status, status)}")

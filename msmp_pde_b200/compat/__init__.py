"""Import shims for the third-party packages the reference imports but this image lacks.

`install_shims()` registers minimal stand-ins for ``torch_geometric`` (``data.Data``, ``nn`` names
the reference imports, ``utils.random.erdos_renyi_graph``), ``torch_cluster`` (``radius_graph``,
``knn_graph``), ``torch_scatter`` (``scatter``), ``h5py``/``matplotlib`` (import-only stubs) in
``sys.modules`` -- only for packages that are genuinely missing -- so that the reference's
``experiments/train.py``, ``cv.py``, ``eval.py`` and ``common/utils.py`` import unchanged
(SURVEY.md F5).  Real installations always win.
"""
from __future__ import annotations

import importlib
import importlib.util
import sys
import types


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install_shims(stub_io: bool = True) -> list[str]:
    """Install shims for missing packages; returns the list of names that were shimmed."""
    done = []
    here = __name__
    if _missing("torch_geometric"):
        pkg = importlib.import_module(here + ".torch_geometric")
        sys.modules["torch_geometric"] = pkg
        for sub in ("data", "nn", "utils", "utils.random"):
            sys.modules["torch_geometric." + sub] = importlib.import_module(here + ".torch_geometric." + sub)
        done.append("torch_geometric")
    if _missing("torch_cluster"):
        sys.modules["torch_cluster"] = importlib.import_module(here + ".torch_cluster")
        done.append("torch_cluster")
    if _missing("torch_scatter"):
        sys.modules["torch_scatter"] = importlib.import_module(here + ".torch_scatter")
        done.append("torch_scatter")
    if stub_io:
        for name in ("h5py", "matplotlib", "matplotlib.pyplot", "torchdiffeq"):
            if _missing(name):
                m = types.ModuleType(name)
                m.__dict__["__shim__"] = True

                def _getattr(attr, __n=name):
                    if attr.startswith("__"):
                        raise AttributeError(attr)

                    def _fail(*a, **k):
                        raise ImportError(f"{__n}.{attr}: {__n} is not installed; msmp_pde_b200 only provides an import stub")
                    return _fail
                m.__getattr__ = _getattr
                sys.modules[name] = m
                done.append(name)
        if "matplotlib" in done and "matplotlib.pyplot" in done:
            sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    return done

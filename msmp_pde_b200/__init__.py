"""msmp_pde_b200 -- B200-native MP-PDE / MSMP-PDE message-passing stack (drop-in torch.nn.Modules).

Importing this package loads ``csrc/libmsmp_b200.so``; it raises if the library has not been built.
There is no CPU / eager fallback for the compute path.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)

__all__ = ["models_gnn", "models_gnn2D", "install"]


def install():
    """Make ``import experiments.models_gnn`` / ``experiments.models_gnn2D`` resolve to this package's drop-in
    classes and shim the third-party imports the reference needs, so that the reference's
    ``train.py`` / ``cv.py`` / ``eval.py`` run unchanged."""
    import sys
    import types

    from . import models_gnn, models_gnn2D
    from .compat import install_shims

    install_shims()
    pkg = sys.modules.get("experiments")
    if pkg is None:
        try:
            import experiments as pkg          # the reference checkout, if it is on sys.path
        except ImportError:
            pkg = types.ModuleType("experiments")
            pkg.__path__ = []
            sys.modules["experiments"] = pkg
    sys.modules["experiments.models_gnn"] = models_gnn
    sys.modules["experiments.models_gnn2D"] = models_gnn2D
    pkg.models_gnn, pkg.models_gnn2D = models_gnn, models_gnn2D
    # the step in front of the models: vectorised, device-capable GraphCreator with a cached topology
    # (common/utils.py:267-471); patched in when the reference checkout is importable
    try:
        import common.utils as cu
        from .graph_creator import GraphCreator
        cu.GraphCreator = GraphCreator
    except Exception:      # reference not on sys.path (or its own imports unavailable): nothing to patch
        pass

"""CPU oracle for the MP-PDE / MSMP-PDE message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline -- never as a fallback for the CUDA path.

What it is: a plain-PyTorch (CPU, float64 by default -- the reference computes in
float64, ``temporal/solvers.py:10``) restatement of

* ``experiments/models_gnn.py:12-149``   Swish, GNN_Layer, GNN_LayerLin
* ``experiments/models_gnn.py:151-281``  MP_PDE_Solver
* ``experiments/models_gnn.py:285-361``  LEMFunction / LEMcuda / LEM / LEMS
* ``experiments/models_gnn.py:1220-1377`` MP_PDE_SolverLEMLinGated
* ``experiments/models_gnn2D.py:9-14,290-458`` unflatten_u, MP_PDE_Solver2DLEMLinGated
* the third-party semantics those files call (PyG ``MessagePassing(aggr='mean')``,
  PyG ``InstanceNorm``, ``torch_scatter.scatter``, ``torch_cluster.radius_graph /
  knn_graph``, ``lem_cuda``), none of which is vendored in the reference.

Parity pinning (see DESIGN.md "Oracle"):

* The model-level glue (feature concatenation order, the ``data.a``-for-``b`` quirk,
  decoder geometry, time stepping, gating) is PINNED: ``tests/golden/make_golden.py``
  imports the reference's own ``experiments/models_gnn.py`` / ``models_gnn2D.py``
  classes from ``/root/reference`` in the build container, runs them on seeded
  inputs, and commits the outputs and gradients as fixtures; the oracle is checked
  against those fixtures in ``tests/test_oracle_golden.py``.
* The third-party semantics underneath (PyG / torch_scatter / torch_cluster /
  lem_cuda) are NOT installed and NOT in the reference tree; the reference has no
  tests or golden vectors of its own.  For those pieces this oracle restates the
  published upstream behaviour and is therefore "parity unpinned" at that
  boundary (the golden generator runs the reference classes on top of these same
  restated semantics).
"""

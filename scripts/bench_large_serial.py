"""bench_large.py (config 5 layer) with ops.SERIALIZE: every op alone on the GPU, so the per-op CUDA-event times are kernel times."""
import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scripts')
from msmp_pde_b200 import ops
ops.SERIALIZE = True
import bench_large
d = bench_large.run(1 << 20, 6, "band")
print(d["ms_per_fwd_bwd"], {k: v["ms"] for k, v in d["ops"].items()})

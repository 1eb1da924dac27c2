"""Config 5 size sweep: one GNN_Layer(128,128,128,25,1) forward + backward on 1 / 2 / 4 Mi nodes x 6 and 1 / 2 Mi nodes x 16
in-neighbours (bench.py's `_layer_c5`, CUDA events, median).  Prints one JSON line per size
(-> profiles/r2_bench_c5_sweep.jsonl).  Parity at these sizes: tests/test_round2_gpu.py::
test_c5_multi_million_node_layer_slices_vs_oracle."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    for n, deg in ((1 << 20, 6), (1 << 21, 6), (1 << 22, 6), (1 << 20, 16), (1 << 21, 16)):
        try:
            rec = bench._layer_c5(n, deg, "random", dev, reps=3)
            rec["peak_mem_gb"] = round(torch.cuda.max_memory_allocated() / 1e9, 1)
        except torch.OutOfMemoryError as exc:
            rec = {"nodes": n, "in_degree": deg, "error": str(exc)[:120]}
            torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        print(json.dumps(rec), flush=True)

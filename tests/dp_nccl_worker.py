"""Worker of tests/test_round2_gpu.py::test_dp_nccl_step_equals_single_gpu (launched with torch.distributed.run, one rank
per GPU, NCCL): prints one JSON line with msmp_pde_b200.dp.dp_selfcheck() of a 3-graphs-per-rank MSMP-PDE2D batch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    from msmp_pde_b200 import models_gnn2D, synth
    from msmp_pde_b200.dp import dp_selfcheck
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    pde, data, meta = synth.config_c2(B=3 * world, nx=100, seed=11)          # identical on every rank

    def make_model():
        torch.manual_seed(0)
        return models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"])

    res = {}
    for graphed in (False, True):
        res["graphed" if graphed else "eager"] = dp_selfcheck(make_model, lambda m: torch.optim.AdamW(m.parameters(), lr=1e-4),
                                                              data, dev, use_graph=graphed)
    if dist.get_rank() == 0:
        print("DPCHECK " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

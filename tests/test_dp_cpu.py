"""Data-parallel host logic on CPU: gloo, world_size 2."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from msmp_pde_b200 import synth
from msmp_pde_b200.dp import shard_graph


def test_shard_graph_partitions_whole_graphs():
    pde, data, meta = synth.config_c3(B=5, nx=20, neighbors=3, seed=2)
    shards = [shard_graph(data, r, 2) for r in range(2)]
    assert [int(s.batch.max()) + 1 for s in shards] == [3, 2]
    assert sum(s.x.shape[0] for s in shards) == data.x.shape[0]
    assert sum(s.edge_index.shape[1] for s in shards) == data.edge_index.shape[1]
    n0 = 0
    rebuilt = []
    for s in shards:
        assert int(s.edge_index.min()) >= 0 and int(s.edge_index.max()) < s.x.shape[0]
        assert bool((s.edge_index[1][1:] >= s.edge_index[1][:-1]).all())      # still destination-sorted
        rebuilt.append(s.edge_index + n0)
        n0 += s.x.shape[0]
    assert torch.equal(torch.cat(rebuilt, 1), data.edge_index)
    assert torch.equal(torch.cat([s.x for s in shards]), data.x)
    assert torch.equal(torch.cat([s.a for s in shards]), data.a)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from msmp_pde_b200.train_step import FlatGradBucket, global_rmse_loss
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)).double()
    g = torch.Generator().manual_seed(1)
    X = torch.randn(10, 6, generator=g, dtype=torch.float64)
    Y = torch.randn(10, 3, generator=g, dtype=torch.float64)
    # reference: full batch on one process
    ref = torch.sqrt(((net(X) - Y) ** 2).sum())
    ref_grads = torch.autograd.grad(ref, list(net.parameters()))
    # sharded: rows 0..5 on rank 0, 6..9 on rank 1 (uneven on purpose)
    sl = slice(0, 6) if rank == 0 else slice(6, 10)
    bucket = FlatGradBucket(net.parameters())
    loss = global_rmse_loss(net(X[sl]), Y[sl])
    loss.backward()
    bucket.all_reduce()
    ok_loss = torch.allclose(loss, ref, rtol=1e-12)
    ok_grad = all(torch.allclose(p.grad, g_, rtol=1e-10, atol=1e-12) for p, g_ in zip(net.parameters(), ref_grads))
    ret[rank] = bool(ok_loss and ok_grad)
    dist.destroy_process_group()


def test_global_loss_and_flat_bucket_gloo_world2():
    """sqrt of the batch-global SSE + SUM all-reduce of one flat bucket == single-process full-batch gradients."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def _worker_single_collective(rank, world, port, ret):
    """The scheme of GraphedTrainStep (train_step.py): unit-seeded backward of the LOCAL squared error, the squared error
    itself as an exact float pair in the two trailing bucket elements, ONE all-reduce, then the factor 1 / (2 sqrt(SSE))."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from msmp_pde_b200.train_step import FlatGradBucket
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)).float()
    g = torch.Generator().manual_seed(1)
    X = torch.randn(10, 6, generator=g)
    Y = torch.randn(10, 3, generator=g).double()
    ref = torch.sqrt(((net(X).double() - Y) ** 2).sum())
    ref_grads = torch.autograd.grad(ref, list(net.parameters()))
    sl = slice(0, 6) if rank == 0 else slice(6, 10)
    bucket = FlatGradBucket(net.parameters(), extra=2)
    sse = ((net(X[sl]).double() - Y[sl]) ** 2).sum()
    sse.backward()
    hi = sse.detach().float()
    bucket.tail[0] = hi
    bucket.tail[1] = (sse.detach() - hi.double()).float()
    bucket.all_reduce()
    total = bucket.tail[0].double() + bucket.tail[1].double()
    loss = torch.sqrt(total)
    scale = (0.5 / loss).float()
    ok_loss = abs(float(loss) - float(ref)) < 1e-6 * float(ref)
    ok_grad = all(torch.allclose(p.grad * scale, g_, rtol=1e-5, atol=1e-7) for p, g_ in zip(net.parameters(), ref_grads))
    ret[rank] = bool(ok_loss and ok_grad)
    dist.destroy_process_group()


def test_single_collective_step_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_worker_single_collective, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]

"""CPU, world_size 2 over gloo: the N>1 host logic of the path -- one flat gradient
all_reduce with SUM / world semantics (loss_func/distrib.py:100-116), parameter broadcast, batch sharding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cruse_b200 import distrib
        from cruse_b200.cruse_net import unet_2
        torch.manual_seed(100 + rank)                      # different init per rank on purpose
        m = unet_2(in_feat=32, ch=(1, 2, 4), rnn_groups=2)  # parameter container only: no compute on CPU
        distrib.broadcast_model(m, src=0)
        ref = torch.cat([p.detach().flatten() for p in m.parameters()])
        gathered = [torch.empty_like(ref) for _ in range(world)]
        dist.all_gather(gathered, ref)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        # gradients: rank r holds (r+1) * ones, except one parameter that has no grad anywhere
        params = list(m.parameters())
        for i, p in enumerate(params):
            p.grad = None if i == 3 else torch.full_like(p, float(rank + 1))
        nbytes = distrib.sync_grad(params)
        want = sum(range(1, world + 1)) / world
        ok = all((p.grad is None) if i == 3 else bool(torch.allclose(p.grad, torch.full_like(p, want))) for i, p in enumerate(params))
        # the same average in place on a flat buffer the gradients are views of (pipeline.CapturedTrainStep.flat_grad)
        flat, views = distrib.flat_grad_views(params)
        for p, v in zip(params, views):
            v.fill_(float(rank + 1))
            p.grad = v
        nb2 = distrib.sync_grad(params, flat=flat)
        ok = ok and nb2 == 4 * flat.numel() and all(bool(torch.allclose(p.grad, torch.full_like(p, want))) for p in params)
        ok = ok and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, views))
        lo, hi = distrib.shard_batch(7, rank, world)
        out[rank] = (same, ok, nbytes, (lo, hi))
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_broadcast_and_sharding_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    for r in range(world):
        same, ok, nbytes, _ = res[r]
        assert same, "broadcast_model did not equalise the replicas"
        assert ok, "sync_grad is not SUM / world"
        assert nbytes > 0
    assert res[0][3] == (0, 4) and res[1][3] == (4, 7)


def test_single_process_is_a_no_op():
    from cruse_b200 import distrib
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    assert distrib.sync_grad([p]) is None and torch.equal(p.grad, torch.full((3,), 2.0))
    assert distrib.world_size() == 1


def test_flat_allreduce_needs_an_initialised_group():
    """distrib.FlatAllreduce (the cruse_flat_allreduce entry point) refuses to build a communicator outside a multi-rank job
    instead of silently doing nothing -- the reference's sync_grad returns early there (loss_func/distrib.py:103-104), which
    sync_grad mirrors; the explicit object does not."""
    import pytest
    from cruse_b200 import distrib
    with pytest.raises(RuntimeError):
        distrib.FlatAllreduce()

"""GPU, 2 ranks over NCCL (skipped on a one-GPU box): the data-parallel split of the training step (SURVEY.md section 8e, section 4
(iv)) -- rank r runs the step on its shard, ONE in-place all_reduce averages the flat gradient buffer (loss_func/distrib.py:100-116
semantics) -- must equal the single-process gradient of the whole batch.  BatchNorm in eval mode (running statistics), because
train-mode statistics are per replica in the reference (plain BatchNorm2d under DDP) and would differ from the big batch by design."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from cruse_b200 import distrib, pipeline
        from cruse_b200.cruse_net import unet_2
        import bench
        torch.manual_seed(50 + rank)                               # different init per rank on purpose: broadcast must fix it
        model = unet_2(in_feat=256)
        bench.randomise_bn(model, seed=77 + rank)
        model = model.to(dev).eval()                               # eval-mode BatchNorm, gradients on
        distrib.broadcast_model(model)
        B, L = 8, 16000
        noisy, clean = bench.synth_batch(B, L, 4242)               # the SAME global batch on every rank
        lo, hi = distrib.shard_batch(B, rank, world)
        params = [p for p in model.parameters()]
        flat, views = distrib.flat_grad_views(params)
        loss = pipeline.train_forward_loss(model, noisy[lo:hi].to(dev), clean[lo:hi].to(dev), 512, 320)
        loss.backward()
        for p, v in zip(params, views):
            if p.grad is not None:
                v.copy_(p.grad)
        nbytes = distrib.sync_grad(params, flat=flat)
        torch.cuda.synchronize()
        res = {"nbytes": nbytes, "backend": dist.get_backend()}
        if rank == 0:
            for p in params:
                p.grad = None
            big = pipeline.train_forward_loss(model, noisy.to(dev), clean.to(dev), 512, 320)
            big.backward()
            worst = 0.0
            for (name, p), v in zip(model.named_parameters(), views):
                if p.grad is None or float(p.grad.norm()) == 0.0:
                    continue
                worst = max(worst, float((v - p.grad).norm() / p.grad.norm()))
            res["worst_rel_l2"] = worst
            res["loss_big"] = float(big)
        lt = loss.detach().clone()
        dist.all_reduce(lt)
        res["loss_mean_of_shards"] = float(lt) / world
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_average_equals_big_batch_gradient(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == {0, 1} and res[0]["backend"] == "nccl"
    assert res[0]["nbytes"] == res[1]["nbytes"] > 12_000_000          # one 12.9 MB buffer
    # equal shards, loss = mean over (B,T,F): mean of the shard losses = the big-batch loss, averaged gradients = its gradient
    assert abs(res[0]["loss_mean_of_shards"] - res[0]["loss_big"]) <= 1e-5 * abs(res[0]["loss_big"])
    assert res[0]["worst_rel_l2"] <= 1e-3, res[0]


def _worker_native(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from cruse_b200 import distrib
        g = torch.Generator().manual_seed(100 + rank)
        n = 3_269_017                                              # the parameter count of config B, not a multiple of 4
        buf_native = torch.randn(n, generator=g).to(dev)          # a fresh allocation: 16-byte aligned start
        buf_torch = buf_native.clone()
        ar = distrib.FlatAllreduce()
        ar(buf_native)
        dist.all_reduce(buf_torch)
        buf_torch.div_(world)
        torch.cuda.synchronize()
        out[rank] = {"max_abs_diff": float((buf_native - buf_torch).abs().max()), "scale": float(buf_torch.abs().max())}
        ar.close()
    finally:
        dist.destroy_process_group()


def test_flat_allreduce_entry_point_equals_torch_all_reduce(cuda):
    """cruse_flat_allreduce (own communicator, ncclAllReduce in place + 1/world) against torch.distributed.all_reduce + div on the
    same buffers: the average of loss_func/distrib.py:111-116."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_native, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert set(res) == {0, 1}
    for r in res.values():
        assert r["max_abs_diff"] <= 1e-6 * r["scale"], res

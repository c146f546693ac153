"""Data-parallel plumbing of the hot path (SURVEY.md section 8 rows a10 / e).

The batch axis shards across ranks with no data-path collective; the ONE exchange per training step is the
gradient average.  ``sync_grad`` keeps the semantics of the reference's ``loss_func/distrib.py:100-116``
(all_reduce SUM then divide by world size, parameters without a gradient are skipped) but issues a single
collective on one flat fp32 buffer (12.9 MB at config B) instead of one per parameter: over NCCL / NVLink 5 that
is one launch whose cost is latency, not bandwidth.  ``broadcast_model`` mirrors ``broadcast_tensors``
(``distrib.py:57-72``).  Backend-agnostic: NCCL on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def is_distributed(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def world_size(group=None) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def flat_grad_views(params, flat=None):
    """one flat fp32 buffer and per-parameter views into it (parameter order, every parameter gets a slot):
    gradients that live in such views are averaged in place by ``sync_grad(..., flat=flat)`` -- no gather / scatter copies."""
    ps = list(params)
    n = sum(p.numel() for p in ps)
    if flat is None:
        flat = torch.zeros(n, device=ps[0].device, dtype=torch.float32)
    elif flat.numel() != n:
        raise RuntimeError(f"flat_grad_views: buffer has {flat.numel()} elements, parameters have {n}")
    views, off = [], 0
    for p in ps:
        views.append(flat[off:off + p.numel()].view(p.shape))
        off += p.numel()
    return flat, views


class FlatAllreduce:
    """The same exchange through the library's own C entry point (``cruse_flat_allreduce``, include/cruse_b200.h: one
    ncclAllReduce in place + the 1/world scale on the current stream) over a communicator of its own: rank 0 draws the NCCL
    unique id, the existing process group (any backend) carries its 128 bytes to the other ranks, every rank joins.  What a
    host that is not PyTorch would bind; ``sync_grad(..., flat=buf, native=FlatAllreduce())`` uses it instead of
    ``torch.distributed.all_reduce``."""

    def __init__(self, group=None):
        import ctypes as C
        from ._lib import check, lib
        if not is_distributed(group):
            raise RuntimeError("FlatAllreduce: torch.distributed is not initialised with more than one rank")
        if not torch.cuda.is_available():
            raise RuntimeError("FlatAllreduce: cruse_b200 runs on sm_100a only (no CPU fallback)")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        raw = (C.c_char * 128)()
        if self.rank == 0:
            check(lib().cruse_nccl_unique_id(raw), "cruse_nccl_unique_id")
        ident = torch.tensor(list(bytes(raw)), dtype=torch.uint8)
        on_device = dist.get_backend(group) == "nccl"
        if on_device:
            ident = ident.cuda()
        dist.broadcast(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = (C.c_char * 128).from_buffer_copy(bytes(ident.cpu().tolist()))
        self.comm = C.c_void_p()
        check(lib().cruse_nccl_comm_init(C.byref(self.comm), self.world, self.rank, raw), "cruse_nccl_comm_init")

    def __call__(self, flat):
        from ._lib import check, lib
        if not (flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous()):
            raise RuntimeError("FlatAllreduce: a contiguous float32 device buffer is needed")
        check(lib().cruse_flat_allreduce(self.comm, flat.data_ptr(), flat.numel(), 1.0 / self.world,
                                         torch.cuda.current_stream(flat.device).cuda_stream), "cruse_flat_allreduce")
        return flat

    def close(self):
        from ._lib import check, lib
        if self.comm:
            check(lib().cruse_nccl_comm_destroy(self.comm), "cruse_nccl_comm_destroy")
            self.comm = None


def sync_grad(params, group=None, flat=None, native=None):
    """average ``p.grad`` over all ranks with ONE all_reduce (loss_func/distrib.py:100-116 semantics).
    ``flat``: the buffer the gradients already live in (``flat_grad_views``; pipeline.CapturedTrainStep.flat_grad):
    reduced in place, two launches in total.  ``native`` (a ``FlatAllreduce``): through ``cruse_flat_allreduce`` instead of
    ``torch.distributed.all_reduce``."""
    if not is_distributed(group):
        return None
    if flat is not None:
        if native is not None:
            native(flat)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.div_(world_size(group))
        return flat.numel() * flat.element_size()
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return None
    flat = _flatten_dense_tensors([p.grad.data for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world_size(group))
    for p, g in zip(ps, _unflatten_dense_tensors(flat, [p.grad.data for p in ps])):
        p.grad.data.copy_(g)
    return flat.numel() * flat.element_size()


def broadcast_model(model, src=0, group=None):
    """rank ``src``'s parameters and buffers to everybody (start-of-training sync, distrib.py:57-72)."""
    if not is_distributed(group):
        return
    tensors = [t.data for t in list(model.parameters()) + list(model.buffers()) if torch.is_floating_point(t)]
    flat = _flatten_dense_tensors(tensors)
    dist.broadcast(flat, src=src, group=group)
    for t, v in zip(tensors, _unflatten_dense_tensors(flat, tensors)):
        t.copy_(v)


def shard_batch(n_items: int, rank: int, world: int):
    """contiguous shard [lo, hi) of a batch axis (bench / tests; the reference's DistributedSampler strides instead,
    tools/train_stand.py:48-51 -- either is a partition, the path has no cross-sample op besides BN statistics,
    which stay per replica as in the reference)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)

"""Developer tool (run under torchrun on N GPUs): what limits the pinned host -> device copies of the e2e path when every rank copies at
once?  Compares, per rank and in aggregate, (a) torch's pinned memory (cudaHostAlloc default), (b) write-combined pinned memory
(cudaHostAllocWriteCombined: no CPU-cache snooping on the DMA reads), (c) one rank copying alone.
   python -m torch.distributed.run --nproc-per-node 8 tools/h2d_probe.py"""
import ctypes, os, sys
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 40960000
cudart = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
def host_alloc(n, flags):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_float * (n // 4)).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.float32)
bufs = {"pinned": torch.empty(nbytes // 4).pin_memory(), "write_combined": host_alloc(nbytes, 0x04)}
for b in bufs.values():
    b.fill_(1.0)
dst = torch.empty(nbytes // 4, device=dev)
def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
def measure(src, active=True, reps=20):
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if active:
        dst.copy_(src, non_blocking=True)
        a.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        b.record()
    barrier()
    ms = a.elapsed_time(b) / reps if active else 0.0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
for name, src in bufs.items():
    ms = measure(src)
    if rank == 0:
        print(f"{name:15s} all {world} ranks at once: {ms:.3f} ms per 41 MB -> {nbytes / ms / 1e6:.1f} GB/s per rank, {world * nbytes / ms / 1e6:.1f} GB/s aggregate", flush=True)
for name, src in bufs.items():
    ms = measure(src, active=(rank == 0))
    if rank == 0:
        print(f"{name:15s} rank 0 alone:            {ms:.3f} ms -> {nbytes / ms / 1e6:.1f} GB/s", flush=True)
if world > 1:
    dist.destroy_process_group()

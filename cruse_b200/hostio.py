"""Host side of the HOST-buffer entry points (``pipeline.CapturedForwardLoss.prefetch`` / ``forward_loss_host``): where the pinned
staging memory lives and which cores the launching thread runs on.

One process per GPU (tools/train_stand.py:151-155 of the reference spawns one per rank); with eight ranks each pushing ~28 GB/s of
pinned host memory through its own PCIe link, the copies only scale if every rank's staging pages sit on the NUMA node its GPU
hangs off and the ranks do not pile onto the same cores.  ``bind_near_gpu`` does both for the calling process, from what NVML and
sysfs report at run time; it never fails the caller (returns what it did, or why not).
"""
from __future__ import annotations

import ctypes
import os


def _gpu_numa_node(handle, nv):
    try:
        bus = nv.nvmlDeviceGetPciInfo(handle).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        with open(path) as f:
            return int(f.read().strip())
    except Exception:  # noqa: BLE001
        return None


def _node_cpus(node):
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            txt = f.read().strip()
        cpus = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        return cpus
    except Exception:  # noqa: BLE001
        return None


def _prefer_node(node):
    """set_mempolicy(MPOL_PREFERRED, {node}): pages this process touches first (incl. cudaHostAlloc'd staging) come from ``node``"""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        sys_set_mempolicy = {"x86_64": 238, "aarch64": 237}.get(os.uname().machine)
        if sys_set_mempolicy is None:
            return False
        rc = libc.syscall(sys_set_mempolicy, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))
        return rc == 0
    except Exception:  # noqa: BLE001
        return False


def bind_near_gpu(device_index: int, local_rank: int = 0, local_world: int = 1) -> dict:
    """Bind the calling process to cores near GPU ``device_index`` (a disjoint share of them per local rank) and prefer that NUMA
    node for its memory.  Call it BEFORE allocating pinned buffers.  Returns a report for the bench line."""
    rep = {"numa_node": None, "cpus": None, "mempolicy": False}
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(device_index)
    except Exception as e:  # noqa: BLE001
        rep["why_not"] = f"nvml: {e!r}"
        return rep
    allowed = sorted(os.sched_getaffinity(0))
    node = _gpu_numa_node(h, nv)
    near = None
    if node is not None and node >= 0:
        rep["numa_node"] = node
        near = _node_cpus(node)
        rep["mempolicy"] = _prefer_node(node)
    if near is None:
        try:                                             # NVML's ideal affinity mask, 64 cpus per word
            words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            near = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        except Exception:  # noqa: BLE001
            near = None
    cpus = [c for c in (near or allowed) if c in allowed] or allowed
    # the ranks whose GPUs share these cores take disjoint shares (a launching thread, a copy thread and NCCL's proxy each want one)
    share = cpus[local_rank % max(1, local_world)::max(1, local_world)] if len(cpus) >= 2 * local_world else cpus
    try:
        os.sched_setaffinity(0, share)
        rep["cpus"] = f"{len(share)} of {len(cpus)} near cores (first {share[0]})"
    except Exception as e:  # noqa: BLE001
        rep["why_not"] = f"sched_setaffinity: {e!r}"
    return rep


_wc_keep = []


def staging_like(t, write_combined=False):
    """a page-locked HOST copy of ``t`` for the HOST-buffer entry points.  ``write_combined``: cudaHostAllocWriteCombined memory -- the
    CPU writes it through write-combining buffers and never caches it, so the GPU's DMA reads do not snoop the CPU caches (NVIDIA's
    advice for buffers the host only writes and the device only reads, which is what an input staging buffer is); reading it back on
    the host is slow.  Falls back to ordinary pinned memory (and says so) when the runtime call is unavailable."""
    import torch
    if not write_combined:
        return t.pin_memory(), "pinned"
    try:
        rt = ctypes.CDLL("libcudart.so.12")
        ptr = ctypes.c_void_p()
        nbytes = t.numel() * t.element_size()
        rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))     # cudaHostAllocWriteCombined
        if rc != 0:
            raise RuntimeError(f"cudaHostAlloc rc={rc}")
        buf = (ctypes.c_byte * nbytes).from_address(ptr.value)
        out = torch.frombuffer(buf, dtype=t.dtype).view(t.shape)
        _wc_keep.append((buf, ptr))              # lives as long as the process (a staging buffer is allocated once)
        out.copy_(t)
        return out, "write_combined"
    except Exception as e:  # noqa: BLE001
        return t.pin_memory(), f"pinned (write-combined unavailable: {e!r})"

#!/usr/bin/env python
"""bench.py -- frames/s of the CRUSE hot path (STFT -> U-Net -> mask -> iSTFT -> wo_male loss) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload infer|stream]

One "step" = one pass of the hot path over one batch of synthetic 16 kHz clips.  At N=1 the
workload is BASELINE.json configs[1]: CRUSE (4-enc/4-dec, 256-GRU) inference fwd+loss, batch
32 x 10 s, n_fft 512, hop 320 (20 ms) -> 32 x 501 = 16 032 STFT frames per step.  N>1 is weak
scaling (every rank runs its own 32 x 10 s shard; the inference path has no data-path
collective); timing = max over ranks, value = all ranks' frames / that time.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, N_FFT, HOP, F_BINS = 16000, 512, 320, 256
WORKLOADS = {
    # name: (clips per GPU, seconds per clip, description)
    "infer": (32, 10.0, "cfg2: CRUSE 4enc/4dec 256-GRU inference fwd+loss, 32x10s per GPU, n_fft512 hop320"),
    # BASELINE configs[2] / [3]: train step = STFT + fwd (train-mode BN) + wo_male + backward (+ 1 flat grad allreduce, N>1)
    "train": (64, 4.0, "cfg3: CRUSE train step (STFT+fwd+wo_male+bwd, no optimizer), 64x4s per GPU, n_fft512 hop320"),
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/ncu_*_full.md),
# cfg2 shapes; None where no capture exists yet
NCU_TRAFFIC = {"gru_seq_fwd_tc": 249.8e6, "gru_seq_fwd": 269.4e6}
try:                                    # refreshed by tools/ncu_traffic.py from the committed full-set captures
    NCU_TRAFFIC.update(json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))))
except Exception:  # noqa: BLE001
    pass


def synth_batch(B, L, seed):
    """SURVEY.md 8d synthetic data (same formula the oracle's synth_batch uses)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    clean = 0.05 * torch.randn(B, L, generator=g)
    noise = 0.05 * torch.randn(B, L, generator=g)
    return clean + noise, clean


def randomise_bn(model, seed=1235):
    g = torch.Generator(device="cpu").manual_seed(seed)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(0.1 * torch.randn(mod.num_features, generator=g))
            mod.running_var.copy_(1 + 0.1 * torch.rand(mod.num_features, generator=g))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU path: the minimal-repair oracle (the reference sources do not import,
    SURVEY.md section 0), torch CPU fp32 with all host threads, on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cruse_oracle as o
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B_full, secs, desc = WORKLOADS[args.workload]
    Bs = min(B_full, args.ref_clips) if args.ref_clips > 0 else B_full        # default: every clip of the workload (same config as our arm)
    L = int(secs * SR)
    T = 1 + L // HOP
    train = args.workload == "train"
    model = o.make_model(F_BINS, eval_stats=not train)
    model.train(train)
    noisy, clean = synth_batch(Bs, L, 20260)

    def cpu_step():
        if train:
            model.zero_grad(set_to_none=True)
            loss = o.forward_loss(model, noisy, clean, N_FFT, HOP)[0]
            loss.backward()
            return loss.detach()
        with torch.no_grad():
            return o.forward_loss(model, noisy, clean, N_FFT, HOP)[0]

    for _ in range(args.warmup):
        cpu_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = cpu_step()
    dt = time.perf_counter() - t0
    fps = Bs * T * args.steps / dt
    sample = f"{Bs} of {B_full} clips x {secs:g}s per step (oracle port, torch {torch.__version__} CPU fp32)"
    line = {
        "impl": "reference", "metric": "frames/sec (16 kHz, 20 ms hop) CRUSE " + ("fwd+loss+bwd" if train else "fwd+loss"), "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "sample": sample, "same_config_as_ours": Bs == B_full},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "loss": float(loss),
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(workload, budget_s=15.0, state_dict=None, ours=None):
    """the oracle port on the host cores over the WHOLE workload batch (same clips, same weights as our arm when ``state_dict``
    is given); ``ours`` = (loss, est [B,T,NF,2]) of our arm on the same data -> the ``parity`` block of the line."""
    from oracle import cruse_oracle as o
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B_full, secs, _ = WORKLOADS[workload]
    Bs, L = B_full, int(secs * SR)
    T = 1 + L // HOP
    train = workload == "train"
    model = o.make_model(F_BINS, eval_stats=not train)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    model.train(train)
    noisy, clean = synth_batch(Bs, L, 20260)
    last = {}

    def cpu_step():
        if train:
            model.zero_grad(set_to_none=True)
            loss = o.forward_loss(model, noisy, clean, N_FFT, HOP)[0]
            loss.backward()
            last["loss"] = float(loss)
        else:
            with torch.no_grad():
                loss, _, est, _ = o.forward_loss(model, noisy, clean, N_FFT, HOP)
                last["loss"], last["est"] = float(loss), est

    cpu_step()
    n, t0 = 0, time.perf_counter()
    while True:
        cpu_step()
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 20:
            break
    out = {"value": Bs * T * n / dt, "unit": "frames/s", "cores": cores, "kind": "port",
           "sample": f"{n} passes over all {Bs} clips x {secs:g}s of the workload (oracle port, torch CPU fp32, {cores} threads)"}
    parity = None
    if ours is not None and state_dict is not None:
        l_ours, est_ours = ours
        parity = {"loss_ours": l_ours, "loss_oracle": last["loss"], "loss_rel_delta": abs(l_ours - last["loss"]) / abs(last["loss"]),
                  "data": "identical clips and weights on both sides (rank 0's batch)"}
        if est_ours is not None and "est" in last:
            parity["enhanced_spectrum_mse"] = float(((est_ours - last["est"].permute(0, 2, 3, 1)) ** 2).mean())
            parity["gate"] = "loss_rel_delta <= 1e-3, enhanced_spectrum_mse < 1e-4 (BASELINE)"
    return out, parity


# ------------------------------------------------------------------------------------------------
def train_block(args, dev, world, rank, flush):
    """BASELINE configs[2] / [3] beside the headline: the training step (STFT + forward with batch statistics + wo_male + backward)
    on a 64 x 4 s shard per rank (N = 8: the 512 x 4 s batch of configs[3]) as one CUDA-graph replay, followed by THE collective of
    the path -- one in-place all_reduce(SUM) + 1/N on the flat fp32 gradient buffer over NCCL (loss_func/distrib.py:100-116
    semantics).  Device-timed with CUDA events, L2 flushed between steps, max over ranks; the all_reduce is also timed alone."""
    import torch.distributed as dist
    from cruse_b200 import distrib, pipeline
    from cruse_b200.cruse_net import unet_2
    B, secs, desc = WORKLOADS["train"]
    L = int(secs * SR)
    T = 1 + L // HOP
    torch.manual_seed(1234)
    model = unet_2(in_feat=F_BINS).to(dev).train()
    if world > 1:
        distrib.broadcast_model(model)
    noisy, clean = synth_batch(B, L, 30260 + rank)
    noisy, clean = noisy.to(dev), clean.to(dev)
    cap = pipeline.CapturedTrainStep(model, B, L, N_FFT, HOP)
    cap.noisy.copy_(noisy)
    cap.clean.copy_(clean)
    params = cap.params
    steps = max(3, min(args.steps, 10))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step():
        loss = cap.replay()
        if world > 1:
            distrib.sync_grad(params, flat=cap.flat_grad)
        return loss

    for _ in range(3):
        step()
    sync()
    evs = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = step()
        b.record()
        evs.append((a, b))
    sync()
    ms = maxr(sum(a.elapsed_time(b) for a, b in evs)) / steps
    out = {"workload": desc, "clips_global": B * world, "frames_per_step_per_gpu": B * T, "steps": steps, "warmup": 3,
           "ms_per_step": ms, "value": B * T * world / (ms / 1e3), "unit": "frames/s", "loss": float(loss),
           "launch": "one CUDA graph replay (forward + loss + backward, ~190 kernels) + in-place all_reduce on the flat gradient buffer",
           "grad_bytes": int(cap.flat_grad.numel() * 4)}
    if world > 1:
        # the collective alone: back-to-back all_reduce + scale on the same buffer
        reps = 20
        for _ in range(3):
            distrib.sync_grad(params, flat=cap.flat_grad)
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            distrib.sync_grad(params, flat=cap.flat_grad)
        b.record()
        sync()
        ar_ms = maxr(a.elapsed_time(b)) / reps
        # the same K steps WITHOUT the collective: what the all_reduce adds to the step
        evs = []
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            cap.replay()
            b.record()
            evs.append((a, b))
        sync()
        ms0 = maxr(sum(a.elapsed_time(b) for a, b in evs)) / steps
        nbytes = out["grad_bytes"]
        out["allreduce"] = {"backend": dist.get_backend(), "nranks": world, "bytes": nbytes, "us": 1e3 * ar_ms,
                            "bus_GBps": 2 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9,
                            "pct_of_step": 100.0 * ar_ms / ms, "ms_per_step_without": ms0, "added_ms": ms - ms0,
                            "how": "one dist.all_reduce(SUM) on the flat fp32 gradient buffer + one in-place 1/N scale (timed alone, 20 back-to-back reps, max over ranks)"}
    else:
        out["allreduce"] = None
    del cap
    return out


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from cruse_b200 import ops, pipeline
    from cruse_b200.cruse_net import unet_2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # cores and memory near this rank's GPU, BEFORE the pinned staging buffers are allocated (cruse_b200/hostio.py)
    from cruse_b200 import hostio
    affinity0 = os.sched_getaffinity(0)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_binding = hostio.bind_near_gpu(local, local, local_world) if not args.no_bind else {"why_not": "--no-bind"}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # (NCCL_DEBUG is left alone: on this image it prints its version line to STDOUT, which must stay one JSON line, and no init
        # lines at all; the communicator is reported on stderr below and as `comm` in the line)
        dist.init_process_group("nccl", device_id=dev)
        if rank == 0:
            print(f"[bench] NCCL communicator: backend nccl {'.'.join(map(str, torch.cuda.nccl.version()))}, nranks {world}", file=sys.stderr)
    if args.gpus != world and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    B, secs, desc = WORKLOADS[args.workload]
    L = int(secs * SR)
    T = 1 + L // HOP
    frames = B * T

    torch.manual_seed(1234)
    model = unet_2(in_feat=F_BINS)
    randomise_bn(model)
    train = args.workload == "train"
    model = model.to(dev)
    model.train(train)
    if train and world > 1:
        from cruse_b200 import distrib
        distrib.broadcast_model(model)
    bn_state0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()} if train else None
    noisy_h, clean_h = synth_batch(B, L, 20260 + rank)
    noisy_h, staging_kind = hostio.staging_like(noisy_h, args.wc)
    clean_h, _ = hostio.staging_like(clean_h, args.wc)
    noisy, clean = noisy_h.to(dev), clean_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    params = [p for p in model.parameters()]

    def run(nz, cl):
        if not train:
            with torch.no_grad():
                return pipeline.forward_loss(model, nz, cl, N_FFT, HOP)[0]
        for p in params:
            p.grad = None
        loss = pipeline.train_forward_loss(model, nz, cl, N_FFT, HOP)
        loss.backward()
        if world > 1:
            distrib.sync_grad(params)          # the one collective of the path: flat fp32 gradient all_reduce over NCCL
        return loss.detach()

    # the whole step (inference: ~95 launches on 13 streams; training: ~190) is captured once into a CUDA graph and replayed
    captured = None
    if not args.no_graph:
        captured = (pipeline.CapturedTrainStep if train else pipeline.CapturedForwardLoss)(model, B, L, N_FFT, HOP)
        captured.noisy.copy_(noisy)
        captured.clean.copy_(clean)

    def step_eager():
        return run(noisy, clean)

    def step():
        if captured is not None:
            loss = captured.replay()
            if train and world > 1:
                distrib.sync_grad(params, flat=captured.flat_grad)
            return loss
        return run(noisy, clean)

    def step_host():
        if captured is not None:
            out = captured(noisy_h, clean_h)
            if train and world > 1:
                distrib.sync_grad(params, flat=captured.flat_grad)
            return (out if train else out[0]).to("cpu")
        loss = run(noisy_h.to(dev, non_blocking=True), clean_h.to(dev, non_blocking=True))
        return loss.to("cpu")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- kernel-launch count of one step
    prof = ops.Profile(timing=False)
    ops.set_profile(prof)
    loss = step_eager()
    ops.set_profile(None)
    launches_per_step = prof.launches

    # ---- timed region: K steps, device time by CUDA events, L2 flushed between steps
    sampler = ClockSampler(local)
    sampler.start()
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = step()
        b.record()
        evs.append((a, b))
    barrier()
    sampler.stop_flag = True
    sampler.join()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = frames * world * args.steps / (total_ms / 1e3)

    # ---- e2e: host buffers in, loss out, copies inside the timed region
    for _ in range(2):
        step_host()
    if captured is not None:              # the prefetch path captures one graph per staging pair on first use: outside the timed region
        for _ in range(2):
            out = captured.run_prefetched(captured.prefetch(noisy_h, clean_h))
        (out if train else out[0]).to("cpu")
    barrier()
    t0 = time.perf_counter()
    if captured is not None:
        # every step: pinned host -> device copy of that step's inputs (on the copy stream, overlapping the previous
        # step's replay), graph replay, loss read back to the host
        # (inference) the loss of step i is read back on its own stream and the host blocks on it only after step i+1 has been
        # launched, so the GPU does not idle for the host's launch latency between steps; every step's loss reaches the host
        ticket = captured.prefetch(noisy_h, clean_h)
        pending = None
        for i in range(args.steps):
            nxt = captured.prefetch(noisy_h, clean_h) if i + 1 < args.steps else None
            out = captured.run_prefetched(ticket)
            if train:
                if world > 1:
                    distrib.sync_grad(params, flat=captured.flat_grad)
                l_host = out.to("cpu")
            else:
                handle = captured.loss_to_host_async(out[0])
                if pending is not None:
                    l_host = pending.result()
                pending = handle
            ticket = nxt
        if pending is not None:
            l_host = pending.result()
    else:
        for _ in range(args.steps):
            l_host = step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = {"value": frames * world * args.steps / float(t.item()), "unit": "frames/s",
           "h2d_bytes_per_step": int(noisy_h.numel() * 4 + clean_h.numel() * 4), "d2h_bytes_per_step": 4,
           "ms_per_step": 1e3 * float(t.item()) / args.steps,
           "how": ("pipeline.CapturedForwardLoss.prefetch/run_prefetched: pinned H2D of step i+1 on a copy stream (into the staging pair the next replay reads in place) under the replay of step i; every step's loss is copied to pinned host memory on a read-back stream and awaited after the next step has been launched"
                   if captured is not None else "forward_loss_host: H2D, launches, loss .to(cpu) back to back")}

    # ---- the same end-to-end loop with the audio shipped as 16-bit PCM (the sample format of the wav files the reference reads,
    # dataset/dataset.py:20): half the bytes per step over PCIe, widened on the device.  Reported beside `e2e`, never instead of it:
    # `e2e` keeps float32 host tensors, what the reference's DataLoader hands to the model.
    e2e_pcm16 = None
    if captured is not None and not train:
        q = lambda x: torch.clamp(torch.round(x * 32768.0), -32768, 32767).to(torch.int16).pin_memory()
        noisy_q, clean_q = q(noisy_h), q(clean_h)
        for _ in range(3):
            out = captured.run_prefetched(captured.prefetch(noisy_q, clean_q))
        loss_q = float(out[0].to("cpu"))
        barrier()
        t0 = time.perf_counter()
        ticket = captured.prefetch(noisy_q, clean_q)
        pending = None
        for i in range(args.steps):
            nxt = captured.prefetch(noisy_q, clean_q) if i + 1 < args.steps else None
            out = captured.run_prefetched(ticket)
            handle = captured.loss_to_host_async(out[0])
            if pending is not None:
                pending.result()
            pending = handle
            ticket = nxt
        pending.result()
        torch.cuda.synchronize()
        tq = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tq, op=dist.ReduceOp.MAX)
        e2e_pcm16 = {"value": frames * world * args.steps / float(tq.item()), "unit": "frames/s", "ms_per_step": 1e3 * float(tq.item()) / args.steps,
                     "h2d_bytes_per_step": int(2 * (noisy_q.numel() + clean_q.numel())), "d2h_bytes_per_step": 4, "loss": loss_q,
                     "how": "the e2e loop with int16 PCM host buffers (prefetch widens them on the device: x / 32768, as soundfile / librosa do on the host); "
                            "same clips quantised to 16 bit, so the loss differs from `loss` in its last digits"}
    if captured is not None and not train:
        captured.check_wavefront()         # a timed-out flag spin anywhere above (outputs NaN) raises here instead of being reported
    # ---- the H2D leg alone, all ranks at once: what the host side delivers to each GPU while the others copy too
    h2d_stream = torch.cuda.Stream(device=dev)
    stage = (torch.empty_like(noisy), torch.empty_like(clean))
    barrier()
    with torch.cuda.stream(h2d_stream):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage[0].copy_(noisy_h, non_blocking=True)
        a.record(h2d_stream)
        for _ in range(10):
            stage[0].copy_(noisy_h, non_blocking=True)
            stage[1].copy_(clean_h, non_blocking=True)
        b.record(h2d_stream)
    barrier()
    h2d_ms = a.elapsed_time(b) / 10
    t = torch.tensor([h2d_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e["h2d_alone"] = {"ms_per_step": float(t.item()), "GBps_per_rank": e2e["h2d_bytes_per_step"] / (float(t.item()) * 1e-3) / 1e9,
                        "GBps_all_ranks": world * e2e["h2d_bytes_per_step"] / (float(t.item()) * 1e-3) / 1e9,
                        "note": "pinned host -> device copy of one step's inputs, every rank copying at the same time, slowest rank"}
    e2e["host_binding"] = host_binding
    e2e["staging"] = staging_kind
    if e2e_pcm16 is not None:
        e2e["pcm16"] = e2e_pcm16
    del stage

    # ---- per-kernel device times of one step (CUDA events on the launching stream) -> roofline
    pk = peaks()
    rows_acc = {}
    for _ in range(5):
        flush.zero_()
        prof = ops.Profile(timing=True)
        ops.set_profile(prof)
        step_eager()
        ops.set_profile(None)
        for i, (n, tag, by, fl, ms) in enumerate(prof.rows()):
            rows_acc.setdefault((i, n, tag, by, fl), []).append(ms)
    kernels = []
    for (i, n, tag, by, fl), mss in sorted(rows_acc.items()):
        ms = sum(mss) / len(mss)
        kernels.append({"call": n.replace("cruse_", ""), "tag": tag, "ms": round(ms, 4), "GBps": round(by / ms / 1e6, 1) if ms else None,
                        "TFLOPs": round(fl / ms / 1e9, 2) if ms else None, "alg_bytes": by, "flops": fl})
    step_ms_sum = sum(k["ms"] for k in kernels)
    # dominant kernel = the C-ABI entry point with the largest summed device time in one step (chunked launches of the same
    # kernel are summed); achieved = its algorithmic bytes / its device time
    groups = {}
    for k in kernels:
        g = groups.setdefault(k["call"], {"ms": 0.0, "bytes": 0, "flops": 0, "launches": 0, "tag": k["tag"]})
        g["ms"] += k["ms"]; g["bytes"] += k["alg_bytes"]; g["flops"] += k["flops"]; g["launches"] += 1
    dom_name, dom = max(groups.items(), key=lambda kv: kv[1]["ms"])
    dom_gbs = dom["bytes"] / dom["ms"] / 1e6
    hbm_groups = {n: g for n, g in groups.items() if not n.startswith(("gru_seq", "bn_fold"))}
    hbm_ms = sum(g["ms"] for g in hbm_groups.values())
    hbm_gbs = sum(g["bytes"] for g in hbm_groups.values()) / hbm_ms / 1e6 if hbm_ms else 0.0
    roofline = {"kernel": f'{dom_name} x{dom["launches"]} [{dom["tag"]}]', "bound": "hbm", "achieved": round(dom_gbs, 1), "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": round(dom_gbs / pk["hbm_gbs"], 4), "traffic": NCU_TRAFFIC.get(dom_name),
                "share_of_kernel_time": round(dom["ms"] / step_ms_sum, 3), "peak_source": pk["source"],
                "us_per_recurrence_step": (round(1e3 * dom["ms"] / (2 * T), 3) if dom_name.startswith("gru_seq") else None),
                "all_other_kernels": {"achieved": round(hbm_gbs, 1), "frac": round(hbm_gbs / pk["hbm_gbs"], 4), "ms": round(hbm_ms, 4),
                                      "note": "every HBM-bound launch of the step together (STFT, conv/convT, input projections, LayerNorm, mask+iSTFT, loss)"},
                "note": "dominant = the GRU recurrence: T sequential steps per layer, latency-bound by construction (h exchange over DSMEM -> 16 "
                        "tcgen05.mma -> tcgen05.ld -> gate math per step), so its HBM fraction is low by design; both layers run side by side "
                        "on 64 of the 148 SMs while the skip convs and input projections use the rest. Per-launch numbers are under 'kernels' "
                        "(timed eagerly, one CUDA-event pair per launch on its own stream; launches on different streams overlap)"}

    line = {
        "metric": "frames/sec (16 kHz, 20 ms hop) CRUSE " + ("fwd+loss+bwd" if train else "fwd+loss"), "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "frames_per_step_per_gpu": frames, "l2": "256 MB flush write between timed steps",
                   "launch": ("one CUDA graph replay per step (captured from the ctypes launches on up to 13 streams)" if captured is not None
                              else "eager ctypes launches on torch's current stream"),
                   "tensor_core_operands": "tf32 (tcgen05, fp32 accumulate in TMEM) in the conv / convT implicit GEMMs and the GRU input projections; IEEE half (kind::f16, same 10-bit mantissa as tf32 for |h| < 1, fp32 accumulate) for W_hh and h in the recurrence; everything else fp32",
                   "gru": "two layers side by side (flag-synchronised time-chunked wavefront; in inference the decoder, mask*X + iSTFT and the loss follow layer 2 in groups of chunks), 2 x 16 utterances software-pipelined per cluster",
                   "collective": ("one flat fp32 gradient all_reduce (NCCL) per step" if (train and world > 1) else "none")},
        "roofline": roofline, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
        "clocks": sampler.summary(), "loss": float(loss), "kernels": kernels,
    }
    if world > 1:
        line["comm"] = {"backend": "nccl", "nranks": world, "nccl_version": ".".join(map(str, torch.cuda.nccl.version())),
                        "used_by": "the `train` block's gradient all_reduce (inference ranks are independent replicas)"}
    # whole-step traffic view: the layer-by-layer algorithmic bytes of the path (205 344 B per frame, every stage
    # reading its inputs and writing its outputs once) over the step time, against the measured HBM peak
    if not train:
        step_gbs = 205344.0 * frames / (total_ms / args.steps / 1e3) / 1e9
        line["roofline"]["whole_step"] = {"alg_bytes_per_frame": 205344, "achieved": round(step_gbs, 1), "frac": round(step_gbs / pk["hbm_gbs"], 4),
                                          "note": "layer-wise algorithmic bytes x frames / step time; the fused kernels move less than this (skip tensors 3/4 and the "
                                                  "decoder intermediates never reach HBM), so it measures the step against the UNFUSED path's traffic"}
    config_streams = line["config"]["launch"].replace("13 streams", "16 streams")
    line["config"]["launch"] = config_streams
    if not train and not args.no_train_block:
        line["train"] = train_block(args, dev, world, rank, flush)
        if world > 1:
            line["config"]["collective"] = ("none on the inference path (replicas); the `train` block runs the path's one collective: "
                                            "a flat fp32 gradient all_reduce over NCCL per training step")
    try:
        os.sched_setaffinity(0, affinity0)             # the CPU leg below uses every host core again
    except Exception:  # noqa: BLE001
        pass
    if world == 1 and not args.no_cpu_baseline:
        # the oracle on the host cores, same clips and weights: the CPU baseline and the oracle-vs-ours delta of THIS run
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        if bn_state0 is not None:
            sd.update(bn_state0)                       # training replays moved the running statistics: compare at the start state
        est_ours = None
        if captured is not None and not train:
            captured.noisy.copy_(noisy)
            captured.clean.copy_(clean)
            l_ours = float(captured.replay())
            est_ours = captured.est.detach().cpu()
        else:
            l_ours = float(loss)
        line["cpu_baseline"], line["parity"] = cpu_baseline_leg(args.workload, state_dict=sd, ours=(l_ours, est_ours) if not train else None)
    if rank == 0:
        if args.table:
            with open(args.table, "w") as f:
                f.write("| call | tag | ms | alg GB/s | frac of %.0f GB/s | TFLOP/s |\n|---|---|---:|---:|---:|---:|\n" % pk["hbm_gbs"])
                for k in kernels:
                    f.write(f'| {k["call"]} | {k["tag"]} | {k["ms"]:.4f} | {k["GBps"]} | {k["GBps"] / pk["hbm_gbs"]:.3f} | {k["TFLOPs"]} |\n')
                f.write(f"\nsum of kernels {step_ms_sum:.3f} ms; timed step {total_ms / args.steps:.3f} ms; {frames} frames/step\n")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="infer", choices=sorted(WORKLOADS),
                    help="infer = BASELINE configs[1] (the headline, default); train = configs[2]/[3]")
    ap.add_argument("--ref-clips", type=int, default=0,
                    help="--impl reference: clips per step of the CPU run (0 = the whole workload batch, i.e. the same config as our arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wc", action="store_true", help="e2e staging buffers in write-combined pinned memory (cudaHostAllocWriteCombined)")
    ap.add_argument("--no-bind", action="store_true", help="do not bind the process / its pinned memory near its GPU")
    ap.add_argument("--no-train-block", action="store_true", help="infer workload: skip the cfg-3/cfg-4 train step + all_reduce block")
    ap.add_argument("--no-graph", action="store_true", help="inference: launch eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--table", default=None, help="write the per-kernel roofline table (markdown) here")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

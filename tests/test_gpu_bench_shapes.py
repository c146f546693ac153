"""GPU: oracle parity AT THE BENCH SHAPES, THROUGH THE BENCH ENTRY POINTS (BASELINE.json configs[1], [2], [4]).

The oracle (oracle/cruse_oracle.py, pinned to the reference's own source by tests/test_oracle.py) runs on the host cores
at the full size -- cfg-2 (32 x 10 s) takes < 1 s, cfg-3 (64 x 4 s, forward + backward) a few seconds -- and is compared
with what `bench.py` times: `pipeline.CapturedForwardLoss.prefetch / run_prefetched` (cfg-2), `pipeline.CapturedTrainStep`
(cfg-3) and `streaming.step` on 2048 utterances x 1 frame (cfg-5).  Model, weights and data are built by bench.py's own helpers.

Stated gates:
  forward, default product mode (tf32 tensor-core operands):  max|d| / max|ref| <= 1e-3 for mask / spectrum / waveform / loss,
                                                               enhanced-spectrum MSE < 1e-4 (BASELINE)
  gradients, exact-fp32 mode:   per parameter tensor cosine >= 0.9999 and rel-L2 <= 1e-3 against autograd of the oracle run in float64
                                 (<= 5x the float32 oracle's own distance to float64 where that is larger: float32 rounding noise,
                                 measured 6e-4 at 8 x 2 s and 7e-4 at 64 x 4 s for the torch-CPU oracle itself)
  gradients, default tf32 mode: cosine >= 0.999 and rel-L2 <= 5e-2 (measured values are logged; DESIGN.md section 3.2 explains
                                 why tf32 operand rounding is amplified ~50x by this network's backward)
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _log(msg):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_bench_shapes.log"), "a") as f:
        f.write(msg + "\n")


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _bench_model(cuda, train):
    """exactly what bench.run_ours builds: seed 1234, default inits, randomised BatchNorm running statistics"""
    import bench
    from cruse_b200.cruse_net import unet_2
    from oracle import cruse_oracle as o
    torch.manual_seed(1234)
    ours = unet_2(in_feat=bench.F_BINS)
    bench.randomise_bn(ours)
    ref = o.unet_2(in_feat=bench.F_BINS)
    ref.load_state_dict(ours.state_dict())
    ours = ours.to(cuda)
    ours.train(train)
    ref.train(train)
    return ours, ref


class _modes:
    """numeric mode of the conv stages / GRU input projections / recurrence for the duration of a with-block"""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        from cruse_b200 import ops
        self.old = ops.GRU_IH_MODE, ops.GRU_SEQ_MODE, ops.get_conv_mode()
        ops.GRU_IH_MODE = ops.GRU_SEQ_MODE = self.mode
        ops.set_conv_mode(self.mode)

    def __exit__(self, *exc):
        from cruse_b200 import ops
        ops.GRU_IH_MODE, ops.GRU_SEQ_MODE = self.old[:2]
        ops.set_conv_mode(self.old[2])


def test_cfg2_inference_32x10s_through_bench_entry_matches_oracle(cuda):
    """BASELINE configs[1]: the exact call sequence of bench.py's e2e leg (pinned host buffers -> prefetch -> run_prefetched ->
    loss_to_host_async) and of its device-timed leg (replay on the static buffers) against the oracle on all 32 clips."""
    import bench
    from cruse_b200 import pipeline
    from oracle import cruse_oracle as o
    B, secs, _ = bench.WORKLOADS["infer"]
    L = int(secs * bench.SR)
    ours, ref = _bench_model(cuda, train=False)
    noisy_h, clean_h = bench.synth_batch(B, L, 20260)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        l0, w0, e0, m0 = o.forward_loss(ref, noisy_h, clean_h, bench.N_FFT, bench.HOP)
    T = m0.shape[2]
    assert (B, T) == (32, 501)
    noisy_p, clean_p = noisy_h.pin_memory(), clean_h.pin_memory()
    cap = pipeline.CapturedForwardLoss(ours, B, L, bench.N_FFT, bench.HOP)
    # (i) the e2e leg
    out = cap.run_prefetched(cap.prefetch(noisy_p, clean_p))
    handle = cap.loss_to_host_async(out[0])
    out = [t.clone() for t in out]
    l_host = handle.result()
    torch.cuda.synchronize()
    cap.check_wavefront()
    # (ii) the device-timed leg
    cap.noisy.copy_(noisy_p)
    cap.clean.copy_(clean_p)
    l2 = cap.replay().clone()
    torch.cuda.synchronize()
    assert torch.equal(l2, out[0]) and torch.equal(cap.mask, out[3])
    l1, w1, e1, m1 = out
    est_ref = e0.permute(0, 2, 3, 1)
    errs = dict(mask=rel_err(m1, m0.view(B, T, -1)), est=rel_err(e1, est_ref), wav=rel_err(w1, w0),
                loss=abs(float(l1) - float(l0)) / abs(float(l0)), spec_mse=float(((e1.cpu() - est_ref) ** 2).mean()))
    _log(f"cfg2 32x10s via CapturedForwardLoss.run_prefetched: loss ours {float(l1):.7f} oracle {float(l0):.7f} " +
         " ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert l_host == float(l1)
    assert errs["mask"] <= 1e-3 and errs["est"] <= 1e-3 and errs["wav"] <= 1e-3 and errs["loss"] <= 1e-3
    assert errs["spec_mse"] < 1e-4


def _grad_rows(ours, ref_grads):
    rows = []
    for name, p in ours.named_parameters():
        gr = ref_grads.get(name)
        if gr is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name          # fc.* is unused (cruse_net.py:146)
            continue
        if name.endswith(".bias") and name.startswith("conv") and name != "conv1_t.bias":
            continue       # conv biases in front of a train-mode BatchNorm: analytically zero gradient (rounding noise on both sides)
        assert p.grad is not None, name
        a, b = p.grad.detach().double().cpu().flatten(), gr.double().flatten()
        rows.append((name, float(torch.dot(a, b) / (a.norm() * b.norm())), float((a - b).norm() / b.norm())))
    return rows


def _oracle_grads(ref, noisy, clean, n_fft, hop, dtype=torch.float32):
    from oracle import cruse_oracle as o
    import copy
    m = copy.deepcopy(ref).to(dtype)
    m.train()
    loss = o.forward_loss(m, noisy.to(dtype), clean.to(dtype), n_fft, hop)[0]
    loss.backward()
    return float(loss), {n: p.grad.detach().double() for n, p in m.named_parameters() if p.grad is not None}


def _summarise(tag, rows):
    worst_cos = min(rows, key=lambda r: r[1])
    worst_l2 = max(rows, key=lambda r: r[2])
    _log(f"{tag}: {len(rows)} tensors; worst cos {worst_cos[1]:.7f} ({worst_cos[0]}), worst relL2 {worst_l2[2]:.3e} ({worst_l2[0]})")
    for name, cos, rl2 in rows:
        _log(f"    {name:40s} cos {cos:.7f} relL2 {rl2:.3e}")
    return worst_cos[1], worst_l2[2]


@pytest.mark.parametrize("F,n_fft,hop,act,B,L", [(256, 512, 320, "relu", 3, 6400), (256, 512, 320, "prelu", 2, 4800),
                                                 (161, 320, 160, "relu", 2, 3200), (256, 512, 320, "relu", 8, 32000)])
def test_exact_mode_end_to_end_gradients_meet_the_stated_gate(cuda, F, n_fft, hop, act, B, L):
    """SURVEY 8d gradient gate (cosine >= 0.9999, rel-L2 <= 1e-3 per parameter tensor) for the WHOLE training step -- STFT ->
    unet_2 with batch statistics -> mask*X -> wo_male -> backward -- in the exact-fp32 mode (conv stages on the CUDA cores, GRU
    projections / recurrence / BPTT / weight-gradient GEMMs through csrc/gru_exact.cu) against autograd of the oracle."""
    from cruse_b200 import pipeline
    from cruse_b200.cruse_net import unet_2
    from oracle import cruse_oracle as o
    ref = o.make_model(F, act=act, eval_stats=False)
    ours = unet_2(in_feat=F, act=act)
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(cuda).train()
    noisy, clean = o.synth_batch(B, L)
    l_ref, g32 = _oracle_grads(ref, noisy, clean, n_fft, hop)
    _, g_ref = _oracle_grads(ref, noisy, clean, n_fft, hop, torch.float64)        # the reference of the comparison: the oracle in float64
    noise = max(float((g32[n] - g_ref[n]).norm() / g_ref[n].norm()) for n in g_ref
                if not (n.endswith(".bias") and n.startswith("conv") and n != "conv1_t.bias"))
    with _modes("fp32"):
        loss = pipeline.train_forward_loss(ours, noisy.to(cuda), clean.to(cuda), n_fft, hop)
        loss.backward()
    torch.cuda.synchronize()
    rows = _grad_rows(ours, g_ref)
    cos, rl2 = _summarise(f"exact-mode gradients vs float64 oracle F={F} act={act} B={B} L={L} (loss ours {float(loss):.7f} oracle {l_ref:.7f}; "
                          f"the float32 ORACLE's own worst rel-L2 to float64: {noise:.3e})", rows)
    assert abs(float(loss) - l_ref) <= 1e-5 * abs(l_ref)
    # the stated gate, unless float32 arithmetic itself cannot meet it on this batch: this network's backward amplifies rounding
    # (DESIGN.md 3.2), and at 8 x 2 s the torch-CPU float32 oracle is already 6e-4 away from its own float64 run (one sample of
    # float32 rounding noise; ours, with a different operation order, is another: same order of magnitude, gate 5x)
    assert cos >= 0.9999 and rl2 <= max(1e-3, 5.0 * noise)


def test_cfg3_train_step_64x4s_through_bench_entry_vs_oracle_autograd(cuda):
    """BASELINE configs[2]: `pipeline.CapturedTrainStep` -- the object bench.py --workload train (and the `train` block of the
    default line) times -- at 64 x 4 s against autograd of the oracle on the same 64 clips: loss and every gradient tensor, in
    the default tf32 mode (stated gate: cos >= 0.999, rel-L2 <= 5e-2) and, eagerly, in the exact-fp32 mode (cos >= 0.9999,
    rel-L2 <= 1e-3 against the oracle in float64 when the host has the memory for it, else <= 2e-3 against the float32 oracle,
    whose own distance to float64 is logged beside it)."""
    import bench
    from cruse_b200 import pipeline
    B, secs, _ = bench.WORKLOADS["train"]
    L = int(secs * bench.SR)
    ours, ref = _bench_model(cuda, train=True)
    noisy_h, clean_h = bench.synth_batch(B, L, 20260)
    torch.set_num_threads(os.cpu_count() or 1)
    l32, g32 = _oracle_grads(ref, noisy_h, clean_h, bench.N_FFT, bench.HOP)
    g64 = None
    try:
        import psutil
        if psutil.virtual_memory().available > 96 << 30:
            l64, g64 = _oracle_grads(ref, noisy_h, clean_h, bench.N_FFT, bench.HOP, torch.float64)
    except Exception:  # noqa: BLE001
        g64 = None
    if g64 is not None:
        rows = [(n, float(torch.dot(g32[n].flatten(), g64[n].flatten()) / (g32[n].norm() * g64[n].norm())),
                 float((g32[n] - g64[n]).norm() / g64[n].norm())) for n in g64
                if not (n.endswith(".bias") and n.startswith("conv") and n != "conv1_t.bias")]
        _summarise(f"cfg3 64x4s: float32 ORACLE vs float64 oracle (loss {l32:.7f} vs {l64:.7f})", rows)
    # ---- the captured step, default mode
    bn_state = {n: b.clone() for n, b in ours.named_buffers()}
    cap = pipeline.CapturedTrainStep(ours, B, L, bench.N_FFT, bench.HOP)
    loss = cap(noisy_h.pin_memory(), clean_h.pin_memory())
    torch.cuda.synchronize()
    assert abs(float(loss) - l32) <= 1e-3 * abs(l32)
    rows = _grad_rows(ours, g32)
    cos, rl2 = _summarise(f"cfg3 64x4s via CapturedTrainStep, tf32 mode vs float32 oracle (loss ours {float(loss):.7f} oracle {l32:.7f})", rows)
    if g64 is not None:
        _summarise("cfg3 64x4s via CapturedTrainStep, tf32 mode vs float64 oracle", _grad_rows(ours, g64))
    assert cos >= 0.999 and rl2 <= 5e-2
    # ---- the same step eagerly in the exact mode
    with torch.no_grad():
        for n, b in ours.named_buffers():
            b.copy_(bn_state[n])
    for p in ours.parameters():
        p.grad = None
    with _modes("fp32"):
        loss_x = pipeline.train_forward_loss(ours, noisy_h.to(cuda), clean_h.to(cuda), bench.N_FFT, bench.HOP)
        loss_x.backward()
    torch.cuda.synchronize()
    want = g64 if g64 is not None else g32
    cos, rl2 = _summarise(f"cfg3 64x4s eager, exact-fp32 mode vs {'float64' if g64 is not None else 'float32'} oracle "
                          f"(loss ours {float(loss_x):.7f})", _grad_rows(ours, want))
    assert abs(float(loss_x) - l32) <= 1e-5 * abs(l32)
    assert cos >= 0.9999 and rl2 <= (1e-3 if g64 is not None else 2e-3)


def test_cfg5_streaming_2048_utterances_one_frame_steps_vs_oracle(cuda):
    """BASELINE configs[4]: 2048 concurrent utterances, one frame per step, persistent state (streaming.step).  The oracle is causal
    (every (2,3) conv looks back one frame, the GRUs run forward in time), so its batched forward over the first t+1 frames gives
    the frame-by-frame answer for step t: each step's mask is compared with the oracle's frame t."""
    import bench
    from cruse_b200 import streaming
    ours, ref = _bench_model(cuda, train=False)
    B, steps, F = 2048, 4, bench.F_BINS
    g = torch.Generator(device="cpu").manual_seed(20260)
    mag = torch.rand(B, steps, F, generator=g)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        want = ref(mag.view(B, 1, steps, F)).view(B, steps, F)
        st = streaming.StreamState()
        worst = 0.0
        for t in range(steps):
            got = streaming.step(ours, mag[:, t:t + 1].contiguous().to(cuda), st)
            assert got.shape == (B, 1, F)
            worst = max(worst, rel_err(got[:, 0], want[:, t]))
    _log(f"cfg5 2048 utterances x 1 frame, {steps} steps vs oracle frame {'{t}'}: worst mask error {worst:.2e}")
    assert worst <= 1e-3

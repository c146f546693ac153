"""The hot path end to end: wav -> STFT -> U-Net mask ->
mask*spectrum -> iSTFT (+ weighted-magnitude loss against the clean spectrum).

Everything between the input waveform and the outputs stays in HBM in the kernels' own
frame-major layouts; this file only sequences launches on the current stream.
"""
from __future__ import annotations

import torch

from . import ops
from .acoustics import hann_window, stft_frames
from .loss import wo_male_frames, wo_male_frames_masked

EPS_MAG = 1e-8   # utils/utils.py:400


def enhance(model, noisy, n_fft=512, hop=320, pad_mode="reflect"):
    """noisy wav [B,L] -> (enhanced wav [B,L], est spec [B,T,NF,2], mask [B,T,F], noisy spec [B,T,NF,2])."""
    F = model.in_feat
    X, mag = stft_frames(noisy, n_fft, hop, n_fft, pad_mode, mag_bins=F, mag_eps=EPS_MAG)   # feature.py:10-30, utils.py:400
    mask = model.forward_frames(mag)                                                        # cruse_net.py:147-165
    est, wav = ops.mask_istft_fwd(X, mask, hann_window(n_fft, n_fft, noisy.device), n_fft, hop, noisy.shape[-1])
    return wav, est, mask, X


def forward_loss(model, noisy, clean, n_fft=512, hop=320, pad_mode="reflect"):
    """STFT + forward + mask + iSTFT + wo_male over the F bins the net sees -> (loss, wav, est, mask)."""
    if ops.OVERLAP_SKIPS and noisy.is_cuda:
        # Two things here do not depend on what the critical path is doing and run on a low-priority side stream:
        # the clean-speech STFT (beside the encoder) and the loss itself, which forms est = mask*X on the fly and therefore
        # runs beside the mask*spectrum + iSTFT kernel instead of after it.
        dev = noisy.device
        F = model.in_feat
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        side2 = _side_stream(dev, 1)                # the loss ranges: beside (not behind) the mask*X + iSTFT ranges
        # consecutive ranges rotate over a few streams each: a range that was served late (the first, largest one shares the GPU with
        # both recurrences) must not hold up the ranges behind it -- the last one sits on the critical path
        istft_streams = (side, _side_stream(dev, 2), _side_stream(dev, 4))
        loss_streams = (side2, _side_stream(dev, 3), _side_stream(dev, 5))
        X, mag = stft_frames(noisy, n_fft, hop, n_fft, pad_mode, mag_bins=F, mag_eps=EPS_MAG)   # feature.py:10-30, utils.py:400
        # the clean-speech STFT is only needed by the loss: it is released behind the encoder and the layer-1 input projections
        # (not beside them, where it would take SMs from what gates the recurrence) and runs beside the GRU; its output buffer
        # exists up front because the loss launches that read it are queued while the model runs
        S = torch.empty_like(X)
        S.record_stream(side)
        clean_started = []

        def start_clean_stft(after=None):
            fork = after
            if fork is None:
                fork = torch.cuda.Event()
                fork.record(main)
            side.wait_event(fork)
            with torch.cuda.stream(side):
                ops.stft_fwd_into(clean, hann_window(n_fft, n_fft, dev), S, n_fft, hop, pad_mode)
                s_ready = torch.cuda.Event()
                s_ready.record(side)
                s_events.append(s_ready)
            for st in loss_streams:
                st.wait_event(s_ready)
            clean_started.append(True)
        # Pipelined schedule (ops.PIPELINE_EDGES): the decoder hands the mask over range by range; mask*X + iSTFT and the loss
        # follow on the side stream, so that only the last range of both is left after the last decoder launch.
        T = X.shape[1]
        FC = ops.mask_istft_chunk_frames(n_fft, hop)
        nct = (T + FC - 1) // FC
        window = hann_window(n_fft, n_fft, dev)
        est_buf = torch.empty_like(X)
        wav_buf = torch.empty(noisy.shape[0], noisy.shape[-1], device=dev, dtype=torch.float32)
        ws = ops.loss_workspace(dev)
        lay_s, lay_x = ops.layout_btf2(S), ops.layout_btf2(X)
        prog = {"c": 0, "p": 0, "i": 0}
        loss_rows = torch.empty(X.shape[0] * T, device=dev, dtype=torch.float32)
        s_events = []
        loss_inputs = {"args": (S, lay_s, X, lay_x, loss_rows), "ready": lambda: list(s_events)}

        def post(mask_all, t0, t1):
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            s_i, s_l = istft_streams[prog["i"] % len(istft_streams)], loss_streams[prog["i"] % len(loss_streams)]
            prog["i"] += 1
            s_i.wait_event(ev)
            s_l.wait_event(ev)
            if prog.get("prev") is not None:
                s_i.wait_event(prog["prev"])     # the first iSTFT CTAs of this range read mask frames of the previous range's decoder launch
            prog["prev"] = ev
            with torch.cuda.stream(s_i):
                c1 = nct if t1 >= T else t1 // FC
                if c1 > prog["c"]:
                    ops.mask_istft_fwd_range(X, mask_all, window, n_fft, hop, est_buf, wav_buf, prog["c"], c1)
                    prog["c"] = c1
            if getattr(model, "_loss_fused", False):
                return                               # the decoder launch left this range's loss shares in loss_rows
            with torch.cuda.stream(s_l):
                # CTAs (= partial-sum slots) of this range: at least two (b, t) rows each, never fewer than ~4 per SM -- the last,
                # short range sits on the critical path and is pure latency with few CTAs
                rows = X.shape[0] * (t1 - t0)
                nparts = max(1, min(rows // 2, max(592, ws.numel() // 4 * (t1 - t0) // T), ws.numel() - prog["p"]))
                ops.wo_male_masked_partial_range(S, lay_s, mask_all, X, lay_x, ws, prog["p"], nparts, X.shape[0], T, F, t0, t1)
                prog["p"] += nparts

        if ops.PIPELINE_EDGES:
            mask = model.forward_frames(mag, post=post, after_encoder=start_clean_stft, loss_inputs=loss_inputs)   # cruse_net.py:147-165
        else:
            start_clean_stft()
            mask = model.forward_frames(mag)
        if not clean_started:
            raise RuntimeError("forward_loss: the model did not call after_encoder()")
        ranges = getattr(model, "_post_ranges", [])
        if ranges and ranges[0][0] == 0 and ranges[-1][1] == T and prog["c"] == nct:
            for st in loss_streams[1:]:
                ev = torch.cuda.Event()
                ev.record(st)
                side2.wait_event(ev)
            with torch.cuda.stream(side2):
                if getattr(model, "_loss_fused", False):
                    have_rows = torch.cuda.Event()   # forward_frames has joined every decoder stream into the caller's
                    have_rows.record(main)
                    side2.wait_event(have_rows)
                    loss = ops.wo_male_finish_rows(loss_rows, X.shape[0], T, F)
                else:
                    loss = ops.wo_male_finish(ws, prog["p"], X.shape[0], T, F)
                done = torch.cuda.Event()
                done.record(side2)
            loss.record_stream(main)                     # made on a side stream, handed to the caller's
            for st in istft_streams[:max(1, min(prog["i"], len(istft_streams)))]:
                for t_ in (est_buf, wav_buf):            # made on the caller's stream, written on the side streams
                    t_.record_stream(st)
                done_istft = torch.cuda.Event()
                done_istft.record(st)
                main.wait_event(done_istft)
            for st in loss_streams:
                for t_ in (ws, S, X, loss_rows):
                    t_.record_stream(st)
            main.wait_event(done)
            if model.gru._wavefront_err is not None:
                # the ranges above were computed from the mask before forward_frames could poison it: a timed-out flag spin
                # turns every output of the step into NaN (and the host raises where it synchronises, see check_wavefront)
                ops.poison_on_error(model.gru._wavefront_err, [loss.view(1), wav_buf, est_buf])
            return loss, wav_buf, est_buf, mask
        have_mask = torch.cuda.Event()
        have_mask.record(main)
        with torch.cuda.stream(side):
            side.wait_event(have_mask)
            loss = wo_male_frames_masked(S, mask, X, F)                                          # loss.py:121-148 on mask*X
            loss.record_stream(main)
            done = torch.cuda.Event()
            done.record(side)
        est, wav = ops.mask_istft_fwd(X, mask, hann_window(n_fft, n_fft, dev), n_fft, hop, noisy.shape[-1])
        main.wait_event(done)
        for st in loss_streams:                      # they only waited for the clean STFT here: join them (graph capture needs it)
            joined = torch.cuda.Event()
            joined.record(st)
            main.wait_event(joined)
        return loss, wav, est, mask
    wav, est, mask, X = enhance(model, noisy, n_fft, hop, pad_mode)
    S, _ = stft_frames(clean, n_fft, hop, n_fft, pad_mode)
    loss = wo_male_frames(S, est, X, model.in_feat)                                          # loss.py:121-148
    return loss, wav, est, mask


_side_streams = {}


def _side_stream(device, which=0):
    key = (device.type, device.index, which)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device, priority=0)
    return _side_streams[key]


def forward_loss_host(model, noisy_host, clean_host, n_fft=512, hop=320, pad_mode="reflect", want_wav=False):
    """Reference-facing call with HOST buffers (pinned or pageable CPU tensors): copies in, runs the
    path, copies the loss (and optionally the enhanced waveform) back.  Used for the e2e bench."""
    dev = next(model.parameters()).device
    noisy = noisy_host.to(dev, non_blocking=True)
    clean = clean_host.to(dev, non_blocking=True)
    loss, wav, _, _ = forward_loss(model, noisy, clean, n_fft, hop, pad_mode)
    out = loss.to("cpu", non_blocking=False)
    ops.raise_if_wavefront_failed(model.wavefront_error_flags(), "forward_loss_host")      # the host is synchronised here anyway
    if want_wav:
        return out, wav.to("cpu")
    return out


class CapturedForwardLoss:
    """``forward_loss`` for one fixed (B, L) captured ONCE into a CUDA graph and replayed per batch.

    The path is ~60 short launches on four streams (the GRU wavefront forks three); replaying them as one graph
    removes the per-launch host cost (ctypes + allocator + event bookkeeping), which otherwise exceeds the device time
    of the step.  Inputs are copied into static device buffers (host tensors: one pinned H2D copy each, inside the
    caller's timed region); outputs are static tensors overwritten by the next call.  Parameters are read in place
    (in-place updates are seen; re-assigned ``.data`` needs a new capture)."""

    def __init__(self, model, B, L, n_fft=512, hop=320, pad_mode="reflect", warmup=2):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("CapturedForwardLoss: cruse_b200 runs on sm_100a only (no CPU fallback)")
        self.model, self.args = model, (n_fft, hop, pad_mode)
        self.noisy = torch.zeros(B, L, device=dev, dtype=torch.float32)
        self.clean = torch.zeros(B, L, device=dev, dtype=torch.float32)
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():          # lazy one-time work (func attributes, allocator pools) outside the capture
            for _ in range(max(1, warmup)):
                forward_loss(model, self.noisy, self.clean, n_fft, hop, pad_mode)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.loss, self.wav, self.est, self.mask = forward_loss(model, self.noisy, self.clean, n_fft, hop, pad_mode)
        self._err_flags = list(model.wavefront_error_flags())      # static device flags of this graph's wavefront

    def check_wavefront(self):
        """Synchronise and raise ``ops.WavefrontTimeout`` if a bounded flag spin of any replay since the last check timed out
        (the outputs of that replay are NaN).  Cheap enough to call once per N steps; ``_HostLoss.result`` calls it when it sees
        a NaN loss."""
        torch.cuda.synchronize(self.noisy.device)
        ops.raise_if_wavefront_failed(self._err_flags, type(self).__name__)

    def __call__(self, noisy, clean):
        """noisy / clean [B,L] on the device or on the host (pinned -> asynchronous copy) -> (loss, wav, est, mask)."""
        self.noisy.copy_(noisy, non_blocking=True)
        self.clean.copy_(clean, non_blocking=True)
        self.graph.replay()
        return self.loss, self.wav, self.est, self.mask

    def replay(self):
        """re-run on whatever the static input buffers hold"""
        self.graph.replay()
        return self.loss

    # -- double-buffered host input: the H2D copy of batch i+1 overlaps the replay of batch i ----------------
    def _staging(self):
        if not hasattr(self, "_stage"):
            dev = self.noisy.device
            self._stage = [(torch.empty_like(self.noisy), torch.empty_like(self.clean)) for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._consumed = [torch.cuda.Event() for _ in range(2)]
            for e in self._consumed:
                e.record(torch.cuda.current_stream(dev))
            self._next = 0
        return self._stage

    def prefetch(self, noisy_host, clean_host):
        """start the asynchronous H2D copy of a (pinned) host batch on the copy stream; returns a ticket for run_prefetched.
        Host tensors are float32 (what the reference's DataLoader yields) or int16 PCM (what its wav files hold)."""
        stage = self._staging()
        i = self._next
        self._next ^= 1
        st = self._copy_stream
        st.wait_event(self._consumed[i])                  # the previous user of this staging pair has been consumed
        with torch.cuda.stream(st):
            for dst, src in zip(stage[i], (noisy_host, clean_host)):
                if src.dtype == torch.int16:
                    # 16-bit PCM (the wav files' own sample format): half the bytes over PCIe, widened on the device exactly as
                    # soundfile / librosa widen it on the host (x / 32768)
                    if not hasattr(self, "_pcm"):
                        self._pcm = {}
                    raw = self._pcm.setdefault((i, id(dst)), torch.empty(dst.shape, device=dst.device, dtype=torch.int16))
                    raw.copy_(src, non_blocking=True)
                    ops.pcm16_to_float(raw, dst)
                else:
                    dst.copy_(src, non_blocking=True)
            self._ready[i].record(st)
        return i

    def _stage_graph(self, ticket):
        """a second / third capture of the same step that reads staging pair ``ticket`` in place (no device-to-device copy of
        the inputs per step); shares the memory pool of the first graph -- the graphs never run concurrently"""
        if not hasattr(self, "_stage_graphs"):
            self._stage_graphs = {}
        if ticket not in self._stage_graphs:
            n_fft, hop, pad_mode = self.args
            nz, cl = self._stage[ticket]
            torch.cuda.synchronize(self.noisy.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.graph.pool()), torch.no_grad():
                outs = forward_loss(self.model, nz, cl, n_fft, hop, pad_mode)
            self._err_flags.extend(self.model.wavefront_error_flags())
            self._stage_graphs[ticket] = (g, outs)
        return self._stage_graphs[ticket]

    def run_prefetched(self, ticket):
        main = torch.cuda.current_stream(self.noisy.device)
        if type(self) is CapturedForwardLoss and self.IN_PLACE_STAGING:
            g, outs = self._stage_graph(ticket)                          # (captured on first use, outside any timed region)
            main.wait_event(self._ready[ticket])
            g.replay()
            self._consumed[ticket].record(main)
            return outs
        main.wait_event(self._ready[ticket])
        self.noisy.copy_(self._stage[ticket][0], non_blocking=True)     # device-to-device, ~30 us for 2 x 20 MB
        self.clean.copy_(self._stage[ticket][1], non_blocking=True)
        self._consumed[ticket].record(main)
        self.graph.replay()
        return self.loss, self.wav, self.est, self.mask

    IN_PLACE_STAGING = True

    # -- asynchronous read-back of a step's loss: the host can launch step i+1 before it blocks on the loss of step i ------
    class _HostLoss:
        def __init__(self, buf, event, owner):
            self.buf, self.event, self.owner = buf, event, owner

        def result(self):
            self.event.synchronize()
            v = float(self.buf[0])
            if v != v:                       # NaN: either the data, or a timed-out wavefront poisoned the step -> raise for the latter
                self.owner.check_wavefront()
            return v

    def loss_to_host_async(self, loss):
        """queue the device->host copy of ``loss`` (a 0-dim output of the step just launched) behind that step; returns a handle
        whose ``result()`` blocks until the value has arrived (and raises if the step's wavefront timed out).  The value is first
        copied, on the caller's stream, into one of two private 1-element device slots, so the next replay of the same graph --
        which overwrites its static ``loss`` tensor -- cannot race the read-back; a handle must be consumed before the step after
        next is queued (two slots alternate)."""
        dev = self.noisy.device
        if not hasattr(self, "_read_stream"):
            self._read_stream = torch.cuda.Stream(device=dev)
            self._host_loss = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._dev_loss = [torch.empty(1, dtype=torch.float32, device=dev) for _ in range(2)]
            self._slot_free = [None, None]
            self._host_slot = 0
        main = torch.cuda.current_stream(dev)
        i = self._host_slot
        self._host_slot ^= 1
        if self._slot_free[i] is not None:
            main.wait_event(self._slot_free[i])          # the previous read-back of this slot has left the device
        self._dev_loss[i].copy_(loss.reshape(1), non_blocking=True)       # on the caller's stream, behind the step
        done = torch.cuda.Event()
        done.record(main)
        rs = self._read_stream
        rs.wait_event(done)
        buf = self._host_loss[i]
        with torch.cuda.stream(rs):
            buf.copy_(self._dev_loss[i], non_blocking=True)
            arrived = torch.cuda.Event()
            arrived.record(rs)
        self._slot_free[i] = arrived
        return CapturedForwardLoss._HostLoss(buf, arrived, self)


class CapturedTrainStep(CapturedForwardLoss):
    """The whole training step -- STFT, forward with batch statistics, wo_male, backward -- for one fixed (B, L) captured
    ONCE into a CUDA graph.  Eagerly the step is ~180 launches plus autograd / allocator bookkeeping, about as much host
    time as the 7 ms of device time; replayed it is device-bound, which is also what lets the weight-gradient kernels run
    beside the BPTT launches on a second stream (autograd._SideWork) instead of queueing behind the host.

    After a call every ``param.grad`` holds THIS step's gradient as a view of ``self.flat_grad`` (one static buffer:
    ``distrib.sync_grad(params, flat=step.flat_grad)`` averages it over the ranks in place; overwritten by the next call, never
    accumulated: gradient accumulation = add them up outside); BatchNorm running statistics are updated by the replay as
    in eager mode.  An optimizer that updates parameters in place is seen by the next replay."""

    def __init__(self, model, B, L, n_fft=512, hop=320, pad_mode="reflect", warmup=3):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("CapturedTrainStep: cruse_b200 runs on sm_100a only (no CPU fallback)")
        self.model, self.args = model, (n_fft, hop, pad_mode)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.noisy = torch.zeros(B, L, device=dev, dtype=torch.float32)
        self.clean = torch.zeros(B, L, device=dev, dtype=torch.float32)
        bn_state = [(b, b.clone()) for n, b in model.named_buffers()]          # warm-up / capture must not move running stats
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                for p in self.params:
                    p.grad = None
                train_forward_loss(model, self.noisy, self.clean, n_fft, hop, pad_mode).backward()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in self.params:
            p.grad = None                                                      # .grad allocated inside the capture = static
        from .distrib import flat_grad_views
        self.flat_grad, views = flat_grad_views(self.params)
        self.graph = torch.cuda.CUDAGraph()
        # the GRU weight gradients (12.6 of the 13.1 MB) are written into their slots of the flat buffer by the kernels that
        # produce them (autograd.gru_layer_bwd, ``sink``); only during this capture: outside it, a second backward without
        # zero_grad must ADD to p.grad, which a kernel writing into p.grad's own memory would break
        model._grad_sink = {p: v for p, v in zip(self.params, views)}
        try:
            with torch.cuda.graph(self.graph):
                loss = train_forward_loss(model, self.noisy, self.clean, n_fft, hop, pad_mode)
                loss.backward()
                self.loss = loss.detach()
                # the step's gradients end up in ONE flat buffer (the data-parallel all_reduce then runs in place on it)
                got = [(v, p.grad) for v, p in zip(views, self.params) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
                torch._foreach_copy_([v for v, _ in got], [g for _, g in got])
        finally:
            del model._grad_sink
        self.grads = [v if p.grad is not None else None for v, p in zip(views, self.params)]
        with torch.no_grad():
            for b, saved in bn_state:
                b.copy_(saved)
        self.wav = self.est = self.mask = None
        self._err_flags = []                 # the training step has no spinning kernels

    def _attach(self):
        for p, g in zip(self.params, self.grads):
            p.grad = g

    def __call__(self, noisy, clean):
        """noisy / clean [B,L] on the device or pinned host -> loss (0-dim, static); gradients in ``param.grad``."""
        self.noisy.copy_(noisy, non_blocking=True)
        self.clean.copy_(clean, non_blocking=True)
        return self.replay()

    def replay(self):
        self.graph.replay()
        self._attach()
        return self.loss

    def run_prefetched(self, ticket):
        main = torch.cuda.current_stream(self.noisy.device)
        main.wait_event(self._ready[ticket])
        self.noisy.copy_(self._stage[ticket][0], non_blocking=True)
        self.clean.copy_(self._stage[ticket][1], non_blocking=True)
        self._consumed[ticket].record(main)
        return self.replay()


def train_forward_loss(model, noisy, clean, n_fft=512, hop=320, pad_mode="reflect"):
    """Training step forward: STFT + U-Net (saving what backward needs) + mask*X + wo_male -> loss with autograd
    history; ``loss.backward()`` runs the sm_100a backward kernels and fills ``param.grad``."""
    from .autograd import unet2_frames_autograd
    from .loss import wo_male_of_mask
    F = model.in_feat
    X, mag = stft_frames(noisy, n_fft, hop, n_fft, pad_mode, mag_bins=F, mag_eps=EPS_MAG)
    if noisy.is_cuda and ops.OVERLAP_BWD:
        # the clean spectrum is read by the loss only: off the encoder's way, on the side stream (S stays referenced by the loss
        # node until its backward has run, i.e. past every main-stream use)
        main, side = torch.cuda.current_stream(noisy.device), _side_stream(noisy.device, 2)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            S, _ = stft_frames(clean, n_fft, hop, n_fft, pad_mode)
        mask = unet2_frames_autograd(model, mag)
        main.wait_stream(side)
    else:
        S, _ = stft_frames(clean, n_fft, hop, n_fft, pad_mode)
        mask = unet2_frames_autograd(model, mag)
    return wo_male_of_mask(S, mask, X, n_fft, hop)

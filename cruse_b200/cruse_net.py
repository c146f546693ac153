"""Drop-in ``unet_2`` / ``GGRU`` (reference: model/cruse_net.py:14-55, :129-165) on sm_100a kernels.

The classes keep the reference's plugin surface: same constructor,
``forward(x [B,1,T,F]) -> mask [B,1,T,F]``, and a ``state_dict()`` whose keys / shapes are those of
the reference's stock ``torch.nn`` children (App. C) -- the children are created in the
reference's order purely as PARAMETER CONTAINERS (so seeds, checkpoints, DDP and Adam behave
identically) and are never called: every stage runs through ``libcruse_sm100.so``.
There is no CPU path: a CPU tensor raises.

Repairs to the reference's literal (non-running) code follow SURVEY.md Appendix A.1 and are the
same ones the oracle makes; they are cited inline.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import ops


def freq_pyramid(in_feat: int, nlayers: int):
    """bins after each (k=3, s=2, p=1) encoder conv (model/cruse_net.py:134-138)."""
    f = [in_feat]
    for _ in range(nlayers):
        f.append((f[-1] + 2 - 3) // 2 + 1)
    return f


def _need_cuda(x, what):
    if not x.is_cuda:
        raise RuntimeError(f"{what}: cruse_b200 runs on sm_100a only; got a {x.device} tensor (no CPU fallback)")


class GGRU(nn.Module):
    """Grouped 2-layer GRU + LayerNorm bottleneck (model/cruse_net.py:14-55).

    Layer 1 outputs are interleaved (stack(dim=-1)+flatten, :43-45 -> feature index h*G+g),
    layer 2 outputs concatenated (:49-50); both orders are produced directly by the recurrence
    kernel's store strides, so no shuffle/copy kernels run.
    """

    def __init__(self, in_features=None, out_features=None, mid_features=None, hidden_size=1024, groups=2):
        super().__init__()
        hidden_size_t = hidden_size // groups
        if hidden_size_t * groups != hidden_size:
            raise ValueError(f"hidden_size {hidden_size} not divisible by groups {groups}")
        self.gru_list1 = nn.ModuleList(
            [nn.GRU(hidden_size_t, hidden_size_t, 1, batch_first=True) for _ in range(groups)])   # :23-26
        self.gru_list2 = nn.ModuleList(
            [nn.GRU(hidden_size_t, hidden_size_t, 1, batch_first=True) for _ in range(groups)])   # :28-31
        self.ln1 = nn.LayerNorm(hidden_size)                                                       # :32
        self.ln2 = nn.LayerNorm(hidden_size)                                                       # :33
        self.groups = groups
        self.hidden_size = hidden_size
        self.mid_features = mid_features

    # -- internal frame-major entry: x [B,T,D] -> [B,T,D] (+ residual fused into ln2) ----------
    def _layer(self, x2d, grus, B, T, interleave, h0=None, want_hT=False):
        xproj = ops.gru_ih_gemm(x2d, [g.weight_ih_l0 for g in grus], [g.bias_ih_l0 for g in grus],
                                [g.bias_hh_l0 for g in grus])
        if T == 1 and want_hT and B >= self.STEP_MIN_B and ops.GRU_SEQ_MODE == "tf32":
            # streaming step of many concurrent utterances (cfg-5): no recurrence to keep on chip, the hidden half is a GEMM
            return ops.gru_step(xproj, [g.weight_hh_l0 for g in grus], [g.bias_hh_l0 for g in grus], h0, interleave)
        return ops.gru_seq_fwd(xproj, [g.weight_hh_l0 for g in grus], [g.bias_hh_l0 for g in grus], B, T,
                               interleave=interleave, h0=h0, want_hT=want_hT)

    # -- two-layer wavefront ----------------------------------------------------------------------
    # The recurrence is the only part of the path that is sequential in time (T steps of ~1 us per layer), and one
    # layer only occupies G*ceil(B/16) clusters of 8 SMs.  Layer 2 (+ LayerNorm 1 and its input projections) of
    # frames [t0,t1) depends only on layer 1 up to t1, so the layers run side by side on three streams, one chunk of
    # frames apart; GRU-internal buffers are time-major so a chunk is a contiguous row range.
    STEP_MIN_B = 128                # 1-frame streaming calls with at least this many utterances use ops.gru_step
    WAVEFRONT_MIN_T = 96
    WAVEFRONT_CHUNKS = 8            # relaunch mode: one recurrence launch per chunk
    WAVEFRONT_FLAG_CHUNKS = int(os.environ.get("CRUSE_FLAG_CHUNKS", "8"))   # flag mode: chunks only gate the hand-over between the layers (ABI limit 16)
    WAVEFRONT_LAST_CHUNK = int(os.environ.get("CRUSE_LAST_CHUNK", "0"))      # frames of the last flag chunk (0 = equal chunks)
    # frames of extra short chunks behind the equal ones, e.g. "32,16" (experiment)
    WAVEFRONT_TAIL = [int(v) for v in os.environ.get("CRUSE_CHUNK_TAIL", "").split(",") if v]
    _side_streams = {}

    @classmethod
    def _streams(cls, device):
        key = (device.type, device.index)
        if key not in cls._side_streams:
            # the latency-critical recurrence chunks get scheduling priority over whatever runs beside them
            cls._side_streams[key] = tuple(torch.cuda.Stream(device=device, priority=-1) for _ in range(8))
        return cls._side_streams[key]

    def plan(self, B, T, device, last_chunk=0):
        """chunk boundaries + zeroed device flags of a flag-synchronised wavefront over T frames, made BEFORE the encoder runs
        so that the first chunk can start as soon as the encoder has produced its frames (``edges``); None when the
        wavefront would not run in flag mode."""
        if ops.GRU_WAVEFRONT_MODE != "flags":
            return None
        nch = max(2, min(self.WAVEFRONT_FLAG_CHUNKS, T // 24))
        # optional SHORT last chunk (WAVEFRONT_LAST_CHUNK frames; 0 = equal chunks, the default): what is left after layer 1 has
        # finished is proportional to it -- measured on B200 it loses all the same (32 / 16 frames: 1.51 / 1.53 vs 1.49 ms per step)
        # ``last_chunk``: the caller's upper bound for the last chunk (what follows layer 2 there is on the critical path, e.g. one
        # round of the one-launch decoder); the environment knob overrides it
        last = min(self.WAVEFRONT_LAST_CHUNK or (last_chunk if 0 < last_chunk < T // nch else 0), T // nch)
        tail = [c for c in self.WAVEFRONT_TAIL if c >= 8]
        if tail and nch >= 3 and sum(tail) <= T // 4 and nch + len(tail) <= 16:
            # equal chunks, then a run of SHORT ones: layer 2 ends one chunk + one hand-over behind layer 1, so the last chunks set the lag
            Tm = T - sum(tail)
            bounds = [Tm * k // nch for k in range(nch + 1)]
            for c in tail:
                bounds.append(bounds[-1] + c)
            nch += len(tail)
        elif last >= 8 and nch >= 3:
            bounds = [(T - last) * k // (nch - 1) for k in range(nch)] + [T]
        else:
            bounds = [T * k // nch for k in range(nch + 1)]
        if not all(bounds[k + 1] - bounds[k] >= 8 for k in range(nch)):
            return None
        # [0,nch) layer-1 projections ready, [nch,2nch) layer-1 chunk stored (counts (CTA, slice) pairs), [2nch,3nch) layer-2
        # projections ready, [3nch,4nch) layer-2 chunk stored, [4nch] error
        flags = torch.zeros(4 * nch + 1, device=device, dtype=torch.int32)
        return {"nch": nch, "bounds": bounds, "flags": flags}

    def _wavefront(self, x, residual, side=None, time_major=False, plan=None, around=None):
        """``side(fork_event) -> (residual, ready_event)``: optional independent work (the skip convs) launched on a
        low-priority stream once the wavefront starts; it fills the SMs the recurrence leaves free.
        ``time_major``: x is [T,B,D] (written so by the last encoder stage); the layer-1 input projections then run per
        chunk on contiguous rows and join the wavefront instead of preceding it.
        ``plan`` / ``around`` (flag mode): the network is causal, so the wavefront is extended over the stages behind the
        GRU.  ``around.groups`` = [(k0, k1), ...] partitions the chunks; for group j covering the frames [t0,t1) the
        callbacks ``skips(j, t0, t1, after_event) -> ready_event`` (the skip convs, on the caller's side stream) and
        ``decode(j, ln2_out, t0, t1)`` (decoder stages, on the current stream) are issued, the latter behind LayerNorm 2 of
        those frames as soon as layer 2 has stored chunk k1-1; ``around.residual`` is the [B,T,D] tensor LayerNorm 2 adds
        (skip4).  Only the LAST group's LayerNorm 2 + decoder remain after the recurrence."""
        if time_major:
            T, B, D = x.shape
        else:
            B, T, D = x.shape
        G, H = self.groups, D // self.groups
        dev = x.device
        g1, g2 = self.gru_list1, self.gru_list2
        w_hh1, b_hh1 = [g.weight_hh_l0 for g in g1], [g.bias_hh_l0 for g in g1]
        w_hh2, b_hh2 = [g.weight_hh_l0 for g in g2], [g.bias_hh_l0 for g in g2]
        w_ih2, b_ih2 = [g.weight_ih_l0 for g in g2], [g.bias_ih_l0 for g in g2]
        main = torch.cuda.current_stream(dev)
        sA, sC, sB, sD, sE, *sMore = self._streams(dev)
        dstreams = (sE, *sMore)                                           # one per decoder group in flight
        w_ih1, b_ih1 = [g.weight_ih_l0 for g in g1], [g.bias_ih_l0 for g in g1]
        if time_major:
            xp1 = torch.empty(T, B, G, 3 * H, device=dev, dtype=torch.float32)
        else:
            xp1 = ops.gru_ih_gemm_tm(x.reshape(B * T, D), w_ih1, b_ih1, b_hh1, B, T)
        # every buffer is allocated on the caller's stream and outlives the join below
        y1 = torch.empty(T, B, D, device=dev, dtype=torch.float32)
        z1 = torch.empty(T, B, D, device=dev, dtype=torch.float32)
        xp2 = torch.empty(T, B, G, 3 * H, device=dev, dtype=torch.float32)
        y2 = torch.empty(B, T, D, device=dev, dtype=torch.float32)
        hA = [torch.empty(G, B, H, device=dev, dtype=torch.float32) for _ in range(2)]
        hB = [torch.empty(G, B, H, device=dev, dtype=torch.float32) for _ in range(2)]
        # chunk boundaries: layer 2 finishes one (LayerNorm + projections + LAST chunk) after layer 1 and starts one FIRST chunk
        # after it, so the first and last chunks are short and the ones in between long (fewer relaunches)
        flag_mode = ops.GRU_WAVEFRONT_MODE == "flags" and time_major
        if plan is None and flag_mode:
            plan = self.plan(B, T, dev)
        if flag_mode and plan is not None:
            # flag-synchronised chunks cost no relaunch, so they are short: layer 2 trails layer 1 by one chunk + its LayerNorm
            # and projections
            nch, bounds = plan["nch"], plan["bounds"]
        elif flag_mode:
            nch = max(2, min(self.WAVEFRONT_FLAG_CHUNKS, T // 24))
            bounds = [T * k // nch for k in range(nch + 1)]
        else:
            nch = max(2, min(self.WAVEFRONT_CHUNKS, T // 32))
        if flag_mode:
            pass
        elif nch >= 4:
            edge = max(16, T // (3 * nch))
            inner = [edge + (T - 2 * edge) * k // (nch - 2) for k in range(nch - 1)]
            bounds = [0] + inner + [T]
        else:
            bounds = [T * k // nch for k in range(nch + 1)]
        H = D // G
        flagged = flag_mode and plan is not None
        if flagged:
            # one launch per layer; chunk hand-over through device flags (layout: plan())
            flags = plan["flags"]
            f_ih1, f_l1, f_ih2, f_l2, f_err = (flags[0:nch], flags[nch:2 * nch], flags[2 * nch:3 * nch], flags[3 * nch:4 * nch],
                                               flags[4 * nch:])
            self._wavefront_err = f_err
            n_wg = G * ((H + 31) // 32) * ((B + 15) // 16)
        fork = torch.cuda.Event()                                         # (after the flags are zeroed)
        fork.record(main)
        if not flagged:
            around = None
        for s_ in (sA, sC, sB, sD, *dstreams):
            s_.wait_event(fork)
        e_skip = [None] * nch
        side_ready = None
        tw1, tb1 = ops._ptr_table(w_ih1), ops._ptr_table(b_ih1)
        tw, tb = ops._ptr_table(w_ih2), ops._ptr_table(b_ih2)
        eD = [None] * nch
        if time_major:
            # layer-1 input projections, one launch per chunk, all queued up front on their own stream: chunk 0 gates the
            # start of the wavefront, the rest stay ahead of layer 1
            with torch.cuda.stream(sD):
                x_ready = plan.get("x_ready") if plan is not None else None
                for k in range(nch):
                    t0, t1 = bounds[k], bounds[k + 1]
                    if x_ready is not None and k == x_ready[0]:
                        sD.wait_event(x_ready[1])              # the encoder frames of the chunks from here on come from its second range
                    ops.gru_ih_gemm_into(x[t0:t1].view(-1, D), w_ih1, b_ih1, b_hh1, xp1[t0:t1], tables=(tw1, tb1))
                    if flagged:
                        ops.flag_set(f_ih1[k:k + 1])
                    eD[k] = torch.cuda.Event()
                    eD[k].record(sD)
        if flagged:
            # host launch order = dependency order (projections, layer 1, LayerNorm + projections, layer 2), so the spinning
            # kernels also terminate when a tool serialises the launches
            with torch.cuda.stream(sA):                                   # layer 1, all frames  (:41-45)
                ops.gru_seq_flagged(xp1, w_hh1, b_hh1, y1, G != 4, True, bounds, f_ih1, 1, f_l1, f_err)
            with torch.cuda.stream(sC):                                   # LayerNorm 1 + layer-2 input projections per chunk (:46-48)
                for k in range(nch):
                    t0, t1 = bounds[k], bounds[k + 1]
                    ops.flag_wait(f_l1[k:k + 1], n_wg, f_err)
                    if G == 4:
                        ops.layernorm_interleave_fwd_into(y1[t0:t1], self.ln1.weight, self.ln1.bias, self.ln1.eps, z1[t0:t1], G)
                    else:
                        ops.layernorm_fwd_into(y1[t0:t1], self.ln1.weight, self.ln1.bias, self.ln1.eps, z1[t0:t1])
                    ops.gru_ih_gemm_into(z1[t0:t1].view(-1, D), w_ih2, b_ih2, b_hh2, xp2[t0:t1], tables=(tw, tb))
                    ops.flag_set(f_ih2[k:k + 1])
            with torch.cuda.stream(sB):                                   # layer 2, all frames  (:49-50)
                ops.gru_seq_flagged(xp2, w_hh2, b_hh2, y2, False, False, bounds, f_ih2, 1, f_l2 if around is not None else None, f_err)
            if side is not None:
                residual, side_ready = side(eD[nch - 1])
            if around is not None:
                # LayerNorm 2 (+ skip4, :51,160) and the decoder follow layer 2 group by group
                out = torch.empty(B, T, D, device=dev, dtype=torch.float32)
                groups = around.groups(nch)
                sgroups = around.skip_groups(nch)
                # skip convs: once the layer-1 projections are through (as `side`), in their own (coarser) groups
                e_skip = [around.skips(j, bounds[k0], bounds[k1], eD[nch - 1]) for j, (k0, k1) in enumerate(sgroups)]
                # every group on its own stream: a late group (the first one is the largest) must not hold up the next ones
                for j, (k0, k1) in enumerate(groups):
                    sG = dstreams[j % len(dstreams)]
                    with torch.cuda.stream(sG):
                        t0, t1 = bounds[k0], bounds[k1]
                        for (s0, s1), ev_s in zip(sgroups, e_skip):
                            if s0 < k1 and k0 < s1:
                                sG.wait_event(ev_s)
                        ops.flag_wait(f_l2[k1 - 1:k1], n_wg, f_err)
                        if getattr(around, "fused_decoder", False):
                            around.decode_fused(j, y2, self.ln2, t0, t1)
                        else:
                            ops.layernorm_fwd_range(y2, self.ln2.weight, self.ln2.bias, self.ln2.eps, around.residual, out, t0, t1)
                            around.decode(j, out, t0, t1)
            for s_ in (sA, sC, sD):                                       # join every branch (graph capture needs it)
                ev = torch.cuda.Event()
                ev.record(s_)
                main.wait_event(ev)
        for k in range(nch if not flagged else 0):
            t0, t1 = bounds[k], bounds[k + 1]
            with torch.cuda.stream(sA):                                   # layer 1, frames [t0,t1)  (:41-45)
                if eD[k] is not None:
                    sA.wait_event(eD[k])
                # layer-1 outputs are stored concatenated (16-byte stores); LayerNorm 1 applies the :43-45 interleave
                ops.gru_seq_chunk(xp1, w_hh1, b_hh1, hA[(k + 1) & 1] if k else None, y1, hA[k & 1], t0, t1, G != 4, True)
                eA = torch.cuda.Event()
                eA.record(sA)
            with torch.cuda.stream(sC):                                   # LayerNorm 1 + layer-2 input projections (:46-48)
                sC.wait_event(eA)
                if G == 4:
                    ops.layernorm_interleave_fwd_into(y1[t0:t1], self.ln1.weight, self.ln1.bias, self.ln1.eps, z1[t0:t1], G)
                else:
                    ops.layernorm_fwd_into(y1[t0:t1], self.ln1.weight, self.ln1.bias, self.ln1.eps, z1[t0:t1])
                ops.gru_ih_gemm_into(z1[t0:t1].view(-1, D), w_ih2, b_ih2, b_hh2, xp2[t0:t1], tables=(tw, tb))
                eC = torch.cuda.Event()
                eC.record(sC)
            with torch.cuda.stream(sB):                                   # layer 2, frames [t0,t1)  (:49-50)
                sB.wait_event(eC)
                ops.gru_seq_chunk(xp2, w_hh2, b_hh2, hB[(k + 1) & 1] if k else None, y2, hB[k & 1], t0, t1, False, False)
            if k == 0 and side is not None:
                # the side work (skip convs) starts once the layer-1 projections are through, so that it does not take the
                # SMs those need to stay ahead of the recurrence
                residual, side_ready = side(eD[nch - 1] if time_major else fork)
        join = torch.cuda.Event()
        join.record(sB)
        main.wait_event(join)
        if side_ready is not None:
            main.wait_event(side_ready)
        for s_ in dstreams:
            ev = torch.cuda.Event()
            ev.record(s_)
            main.wait_event(ev)
        if around is not None:
            return out
        return ops.layernorm_fwd(y2, self.ln2.weight, self.ln2.bias, self.ln2.eps, residual=residual)   # :51 (+ skip4, :160)

    def uses_wavefront(self, B, T, state=None, want_state=False):
        # both layers side by side need 2*G*ceil(B/32) co-resident clusters
        return (state is None and not want_state and T >= self.WAVEFRONT_MIN_T and ops.GRU_SEQ_MODE == "tf32"
                and ops.GRU_IH_MODE == "tf32" and ops.GRU_WAVEFRONT
                and 2 * self.groups * ((B + 31) // 32) <= ops.gru_seq_max_clusters(self.hidden_size // self.groups))

    def forward_frames(self, x, residual=None, state=None, want_state=False, side=None, time_major=False, plan=None,
                       around=None, skip_ln2=False):
        """x [B,T,D] frame-major ([T,B,D] with ``time_major``, wavefront path only); the result is always [B,T,D].
        state = (h1 [G,B,H], h2 [G,B,H]) carries the recurrence (streaming, model/based_model/cust_conv.py:303-325).
        ``skip_ln2`` (layer-by-layer path only): return the output of GRU layer 2 -- the caller's one-launch decoder applies
        LayerNorm 2 + residual itself."""
        _need_cuda(x, "GGRU")
        if time_major:
            T, B, D = x.shape
        else:
            B, T, D = x.shape
        if D != self.hidden_size:
            raise RuntimeError(f"GGRU: feature size {D} != hidden_size {self.hidden_size}")
        self._wavefront_err = None          # set by a flag-synchronised wavefront: device flag "a bounded spin timed out"
        if self.uses_wavefront(B, T, state, want_state):
            if skip_ln2:
                raise RuntimeError("GGRU: skip_ln2 is an option of the layer-by-layer path")
            return self._wavefront(x, residual, side, time_major, plan, around)
        if side is not None or time_major or around is not None:
            raise RuntimeError("GGRU: side work / time-major input need the wavefront path")
        h1 = h2 = None
        if state is not None:
            h1, h2 = state
        r1 = self._layer(x.reshape(B * T, D), self.gru_list1, B, T, True, h1, want_state)
        y1, n1 = r1 if want_state else (r1, None)
        z1 = ops.layernorm_fwd(y1, self.ln1.weight, self.ln1.bias, self.ln1.eps)
        r2 = self._layer(z1.view(B * T, D), self.gru_list2, B, T, False, h2, want_state)
        y2, n2 = r2 if want_state else (r2, None)
        out = y2 if skip_ln2 else ops.layernorm_fwd(y2, self.ln2.weight, self.ln2.bias, self.ln2.eps, residual=residual)
        return (out, (n1, n2)) if want_state else out

    def forward(self, x):
        """reference layout: x [B,C,T,F'] -> [B,C,T,F'] (:37-55; :53 repaired to out.view)."""
        _need_cuda(x, "GGRU")
        B, Cc, T, Fp = x.shape
        frames = x.transpose(1, 2).contiguous().view(B, T, Cc * Fp)      # :39-40 (layout plumbing)
        out = self.forward_frames(frames)
        if self._wavefront_err is not None:
            ops.poison_on_error(self._wavefront_err, [out])               # a timed-out flag spin must not return numbers
        return out.view(B, T, Cc, Fp).transpose(1, 2).contiguous()       # :53-54


class unet_2(nn.Module):
    """4-enc / 4-dec convolutional-recurrent U-Net mask estimator (model/cruse_net.py:129-165)."""

    def __init__(self, in_feat=161, ch=(1, 8, 16, 32, 64), stride=(1, 2), rnn_groups=4, act="relu"):
        super().__init__()
        if tuple(stride) != (1, 2):
            raise ValueError("unet_2: only the reference's stride (1,2) is supported")
        self.laynum = len(ch) - 1
        self.ker_x = 2
        self.stride = tuple(stride)
        self.padding = [self.ker_x - stride[0], 3 - stride[1]]                                       # :136
        self.ch = tuple(ch)
        self.in_feat = in_feat
        self.freqs = freq_pyramid(in_feat, self.laynum)
        self.act_kind = act
        n = self.laynum
        for i in range(n):                                                                          # :137-143
            setattr(self, f"conv{i+1}", nn.Conv2d(ch[i], ch[i + 1], (self.ker_x, 3), self.stride, self.padding))
            tmp = n - i
            setattr(self, f"conv{tmp}_t", nn.ConvTranspose2d(ch[tmp], ch[tmp - 1], (1, 3), self.stride))  # :140 repaired
            setattr(self, f"bn{i+1}", nn.BatchNorm2d(ch[i + 1]))
            if tmp >= 2:
                setattr(self, f"bn{tmp}_t", nn.BatchNorm2d(ch[tmp - 1]))                             # :142 repaired
            setattr(self, f"skip_connect_{i+1}",
                    nn.Conv2d(ch[i + 1], ch[i + 1], (1, 3), bias=False, padding=(0, 1)))             # :143 repaired
        self.gru = GGRU(hidden_size=ch[-1] * self.freqs[-1], groups=rnn_groups)                      # :144 repaired
        self.elu = nn.ReLU()                                                                         # :145
        self.fc = nn.Linear(in_feat, in_feat)                                                        # :146 (unused)
        if act == "prelu":
            for k in range(1, n + 1):
                setattr(self, f"act{k}", nn.PReLU(ch[k]))
            for k in range(n, 1, -1):
                setattr(self, f"act{k}_t", nn.PReLU(ch[k - 1]))
        elif act != "relu":
            raise ValueError(f"act must be 'relu' or 'prelu', got {act!r}")

    # ------------------------------------------------------------------------------------
    _skip_streams = {}

    @classmethod
    def _skip_stream(cls, device):
        key = (device.type, device.index)
        if key not in cls._skip_streams:
            cls._skip_streams[key] = torch.cuda.Stream(device=device, priority=0)
        return cls._skip_streams[key]

    def _alpha(self, name):
        return getattr(self, name).weight if self.act_kind == "prelu" else None

    def _stage(self, h, conv, bn, alpha, kt_fs, train, hist=None, fold=None):
        """conv + BN + act.  eval: one fused kernel; train: conv(+stat partials) -> finalize -> bn_act."""
        kt, fs = kt_fs
        if not train:
            scale, shift = fold if fold is not None else ops.bn_fold(bn)
            return ops.conv_fwd(h, conv.weight, conv.bias, scale, shift, alpha, self.act_kind, kt, fs, hist=hist)
        z, stats = ops.conv_fwd(h, conv.weight, conv.bias, None, None, None, "none", kt, fs, want_stats=True)
        B, T, Cn, F = z.shape
        scale, shift, _, _ = ops.bn_finalize(stats, B * T * F, bn)
        return ops.bn_act_fwd(z, scale, shift, alpha, self.act_kind)

    # developer knobs: chunk indices at which the pipelined decoder / skip convs are cut (default: derived from the chunk count)
    DECODE_CUTS = [int(v) for v in os.environ.get("CRUSE_DECODE_CUTS", "").split(",") if v]
    SKIP_CUTS = [int(v) for v in os.environ.get("CRUSE_SKIP_CUTS", "").split(",") if v]
    SIDE_CAP = int(os.environ.get("CRUSE_SIDE_CAP", "0"))     # CTAs of the persistent side kernels (0 = the SMs the recurrences leave free, minus SIDE_SPARE)
    # SMs the persistent side kernels leave free for the hand-over kernels (LayerNorm 1 + layer-2 projections of a chunk), which otherwise wait
    # for a side kernel to END before they get an SM (measured r2: 0 -> 1.188 ms, 12 -> 1.172 ms per step; with the fused decoder: 0 -> 1.077, 12 -> 1.049, 24 -> 1.028, 40 -> 1.039)
    SIDE_SPARE = int(os.environ.get("CRUSE_SIDE_SPARE", "24"))
    DEC_CAP = int(os.environ.get("CRUSE_DEC_CAP", "0"))          # CTAs of the fused decoder launches (0 = like the other side kernels, -1 = uncapped)
    # wavefront chunks of the encoder that run in front of the recurrences (0 = all of it, the default).  Measured on B200 (r2, 8 chunks,
    # 1.266 ms with the whole encoder in front): 4 -> 1.273, 3 -> 1.315 (layer 1 stalls 87 us at chunk 3: the second encoder range
    # takes ~340 us beside the recurrences), 2 -> 1.45, 1 -> 1.47 ms.  Kept as a knob; the two-range schedule is bit-identical.
    HEAD_CHUNKS = int(os.environ.get("CRUSE_HEAD_CHUNKS", "0"))
    HEAD_CAP = int(os.environ.get("CRUSE_HEAD_CAP", "84"))        # CTAs of the encoder launches that run beside layer 1
    _enc_streams = {}

    @classmethod
    def _enc_stream(cls, device):
        key = (device.type, device.index)
        if key not in cls._enc_streams:
            cls._enc_streams[key] = torch.cuda.Stream(device=device, priority=-1)
        return cls._enc_streams[key]

    @staticmethod
    def chunk_groups(nch, cuts=None):
        """[(k0, k1), ...]: the groups of wavefront chunks the pipelined decoder / skip convs are issued in -- pairs of chunks, then
        the last two chunks singly (``cuts``: explicit chunk indices instead).  Measured on B200 at 8 chunks (r2, after the spectral /
        loss kernels behind the decoder had become cheap): cuts 0,5,7,8 -> 1.32 ms per step, 0,4,6,7,8 -> 1.29, 0,3,5,7,8 and
        0,2,4,6,7,8 -> 1.27, single chunks -> 1.31."""
        cuts = sorted(set(cuts)) if cuts else sorted({0, nch} | {c for c in (nch - 6, nch - 4, nch - 2, nch - 1) if c > 0})
        if cuts[0] != 0 or cuts[-1] != nch or any(c < 0 or c > nch for c in cuts):
            raise RuntimeError(f"chunk_groups: cuts {cuts} do not partition [0, {nch}]")
        return list(zip(cuts[:-1], cuts[1:]))

    def _forward_frames_pipelined(self, mag, plan, folds, post=None, after_encoder=None, loss_inputs=None):
        """Eval, whole utterances, flag-synchronised wavefront: the net is causal and the transposed convs / (1,3) skip convs
        have no time taps at all, so LayerNorm 2 + the decoder (and the skip convs they add) are run per GROUP of wavefront
        chunks as soon as layer 2 of the GRU has stored them: behind the last step of the recurrence only the last chunk's
        LayerNorm + decoder are left instead of the whole decoder.  (The encoder stays in front: per-chunk encoder launches
        were measured too -- layer 1 then starts 170 us earlier but is starved, ~25 us per small launch under load against
        87 us of recurrence per chunk: 1.92 ms instead of 1.53.)"""
        B, T, F = mag.shape
        n = self.laynum
        dev = mag.device
        main = torch.cuda.current_stream(dev)
        s_skip = self._skip_stream(dev)
        new = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
        C4, F4 = self.ch[n], self.freqs[n]
        D = C4 * F4
        unet = self
        # ---- encoder, whole utterances (:149-152 repaired); the last stage writes the GRU input time-major
        skip_out = [new(B, T, self.ch[k], self.freqs[k]) for k in range(1, n + 1)]
        # skip conv k (:153-155) reads the same tensor e_k as encoder stage k+1 (:150-152): where a fused tensor-core instantiation
        # exists (Cin 8 / 16 of the 256-bin pyramid) it rides along with that stage -- e_k is read once and the skip conv costs no
        # launch of its own; the remaining skip convs run beside the recurrences (Around.skips)
        fused = {k - 1 for k in range(2, n) if ops.FUSE_SKIPS and self.ch[k - 1] in (8, 16) and self.freqs[k - 1] == 2 * self.freqs[k]
                 and self.freqs[k] in (64, 32)}
        enc = [new(B, T, self.ch[k], self.freqs[k]) for k in range(1, n)] + [new(T, B, self.ch[n], self.freqs[n])]   # last stage: time-major

        def encode(t0, t1):
            """encoder stages 1..n for the output frames [t0, t1) (the net is causal: a stage's range needs the same range and one
            frame before it of the stage below, which the previous range has produced)"""
            h = mag.view(B, T, 1, F)
            for k in range(1, n + 1):
                conv = getattr(self, f"conv{k}")
                scale, shift = folds[f"bn{k}"]
                alpha = self._alpha(f"act{k}")
                if (k - 1) in fused:
                    ops.conv_skip_fwd(h, conv.weight, conv.bias, scale, shift, alpha, self.act_kind,
                                      getattr(self, f"skip_connect_{k - 1}").weight, out=enc[k - 1], out_skip=skip_out[k - 2], t0=t0, t1=t1)
                else:
                    ops.conv_fwd_range(h, conv.weight, conv.bias, scale, shift, alpha, self.act_kind, 2, 2, B, T, enc[k - 1], t0, t1,
                                       out_tm=(k == n))
                h = enc[k - 1]

        # The head of the step: only the first HEAD_CHUNKS wavefront chunks of the encoder run in front of the recurrences; the rest
        # of the encoder (one more set of launches over the remaining frames, grids capped so that the hand-over GEMMs find SMs)
        # runs beside layer 1 on its own stream and gates the layer-1 input projections of its chunks (plan["x_ready"]).
        head = min(self.HEAD_CHUNKS, plan["nch"]) if self.HEAD_CHUNKS > 0 else plan["nch"]
        t_head = plan["bounds"][head]
        encode(0, t_head)
        if t_head < T:
            s_enc = self._enc_stream(dev)
            ev = torch.cuda.Event()
            ev.record(main)
            s_enc.wait_event(ev)
            ops.set_conv_max_ctas(self.HEAD_CAP)
            try:
                with torch.cuda.stream(s_enc):
                    encode(t_head, T)
                    rest_ready = torch.cuda.Event()
                    rest_ready.record(s_enc)
            finally:
                ops.set_conv_max_ctas(0)
            plan["x_ready"] = (head, rest_ready)
        # LayerNorm 2 + all four decoder stages as ONE launch per group (decoder_fused.cu) where the geometry is the 256-bin pyramid
        fuse_dec = (ops.FUSE_DECODER and n == 4 and tuple(self.ch) == (1, 8, 16, 32, 64) and tuple(self.freqs) == (256, 128, 64, 32, 16)
                    and self.act_kind in ("relu", "prelu"))
        dec = [] if fuse_dec else [new(B, T, self.ch[k - 1], self.freqs[k - 1]) for k in range(n, 1, -1)]
        dec_image = None
        # skip convs 4 and 3 inside the decoder launch, from e4 / e3: no skip-conv launches are left beside the recurrences
        fuse_skip34 = fuse_dec and ops.FUSE_SKIP34
        # the caller's wo_male inputs: the fused decoder then leaves every frame's share of the loss beside the mask (no loss launches)
        fuse_loss = fuse_dec and loss_inputs is not None and F == 256 and ops.FUSE_LOSS
        self._loss_fused = fuse_loss
        if fuse_dec:
            names = [f"conv{k}_t" for k in range(n, 0, -1)]
            prep_ev = torch.cuda.Event()
            prep_ev.record(main)                          # the BatchNorm folds are through
            s_skip.wait_event(prep_ev)
            with torch.cuda.stream(s_skip):               # off the head of the step: nothing needs the images before the first decoder group
                pargs = ([getattr(self, nm).weight for nm in names], [getattr(self, nm).bias for nm in names],
                         [folds[f"bn{k}_t"][0] for k in range(n, 1, -1)], [folds[f"bn{k}_t"][1] for k in range(n, 1, -1)],
                         [self._alpha(f"act{k}_t") for k in range(n, 1, -1)] if self.act_kind == "prelu" else None, self.act_kind)
                dec_image = ops.decoder_fused_prep(*pargs, self.skip_connect_4.weight if fuse_skip34 else None,
                                                   self.skip_connect_3.weight if fuse_skip34 else None)
                dec_image.record_stream(main)
        mask_buf = new(B, T, 1, F)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        layer_sms = 8 * self.gru.groups * ((B + 31) // 32)

        class Around:
            residual = skip_out[n - 1].view(B, T, D)

            @staticmethod
            def groups(nch):
                # few, large groups (every launch costs ~10 us of set-up and runs beside the recurrences; measured on B200 at
                # 8 chunks: cuts 0,5,7,8 -> 1.45 ms, 0,4,6,7,8 -> 1.47-1.50, pairs -> 1.53, single chunks -> 1.63, off -> 1.52)
                return unet.chunk_groups(nch, unet.DECODE_CUTS)

            @staticmethod
            def skip_groups(nch):
                # the two skip convs that are still separate launches (3 and 4) depend on the encoder only: ONE launch each over all
                # frames, early (low priority, beside layer 1) -- in groups the last one was served so late that it held up the last
                # decoder group (measured r2: cuts 0,5,7,8 -> 1.202 ms, 0,6,8 -> 1.191, 0,8 -> 1.188)
                return unet.chunk_groups(nch, unet.SKIP_CUTS or [0, nch])

            @staticmethod
            def caps(j, ngroups):
                # persistent grids sized to the SMs the recurrences leave free: both layers busy / layer 2 only / nothing
                if j == ngroups - 1:
                    return 0
                if unet.SIDE_CAP:
                    return unet.SIDE_CAP
                return max(32, sms - (2 if j < ngroups - 2 else 1) * layer_sms - unet.SIDE_SPARE)

            @staticmethod
            def skips(j, t0, t1, after):                                                         # :153-156
                if j == 0 and after_encoder is not None:
                    after_encoder(after)             # the caller's off-path work starts with the skip convs: once the layer-1
                                                     # projections are through, so that it takes no SMs from what gates layer 1
                s_skip.wait_event(after)
                ops.set_conv_max_ctas(unet.SIDE_CAP or max(32, sms - 2 * layer_sms - unet.SIDE_SPARE))
                try:
                    with torch.cuda.stream(s_skip):
                        for k in range(n, 0, -1):
                            if k in fused:
                                continue                                   # came out of encoder stage k+1 already
                            if fuse_skip34 and k >= n - 1:
                                continue                                   # made inside the decoder launches
                            wk = getattr(unet, f"skip_connect_{k}").weight
                            ops.conv_fwd_range(enc[k - 1], wk, None, None, None, None, "none", 1, 1, B, T, skip_out[k - 1], t0, t1,
                                               in_tm=(k == n))
                        ev = torch.cuda.Event()
                        ev.record(s_skip)
                finally:
                    ops.set_conv_max_ctas(0)
                return ev

            fused_decoder = fuse_dec

            @staticmethod
            def decode_fused(j, y2, ln2, t0, t1):                                                # :51,160-164 repaired, one launch
                cap = Around.caps(j, len(Around.groups(plan["nch"])))
                if unet.DEC_CAP:
                    cap = max(0, unet.DEC_CAP)
                larg = None
                if fuse_loss:
                    for ev in loss_inputs["ready"]():        # the clean-speech spectrum (made on the caller's side stream)
                        torch.cuda.current_stream(dev).wait_event(ev)
                    larg = loss_inputs["args"]
                # (measured and dropped: the last group on the 16-frames-per-SM launch with its skip tensors made by two small early
                # skip-conv launches -- 1.014 vs 0.995 ms per step)
                sk = [enc[n - 1], enc[n - 2], skip_out[1], skip_out[0]] if fuse_skip34 else [skip_out[k - 1] for k in range(n, 0, -1)]
                ops.decoder_fused_range(y2, ln2.weight, ln2.bias, ln2.eps, sk, dec_image, mask_buf, t0, t1, cap, loss=larg,
                                        skip_convs=fuse_skip34)
                if post is not None:
                    post(mask_buf.view(B, T, F), t0, t1)
                    unet._post_ranges.append((t0, t1))

            @staticmethod
            def decode(j, ln2_out, t0, t1):                                                      # :161-164 repaired
                ops.set_conv_max_ctas(Around.caps(j, len(Around.groups(plan["nch"]))))
                try:
                    cur = ln2_out.view(B, T, C4, F4)
                    for i, k in enumerate(range(n, 1, -1)):
                        conv = getattr(unet, f"conv{k}_t")
                        scale, shift = folds[f"bn{k}_t"]
                        ops.convT_fwd_range(cur, conv.weight, conv.bias, scale, shift, unet._alpha(f"act{k}_t"), unet.act_kind,
                                            skip_out[k - 2], dec[i], t0, t1)
                        cur = dec[i]
                    ops.convT_fwd_range(cur, unet.conv1_t.weight, unet.conv1_t.bias, None, None, None, "sigmoid", None, mask_buf, t0, t1)
                finally:
                    ops.set_conv_max_ctas(0)
                if post is not None:                 # what the caller does with the mask (mask*X + iSTFT, loss) follows range by range
                    post(mask_buf.view(B, T, F), t0, t1)
                    unet._post_ranges.append((t0, t1))

        self.gru.forward_frames(enc[n - 1].view(T, B, D), time_major=True, plan=plan, around=Around)
        ev = torch.cuda.Event()
        ev.record(s_skip)
        main.wait_event(ev)
        if "x_ready" in plan:
            main.wait_event(plan["x_ready"][1])
        if self.gru._wavefront_err is not None:
            ops.poison_on_error(self.gru._wavefront_err, [mask_buf])      # a timed-out flag spin must not return numbers
        return mask_buf.view(B, T, F)

    def forward_frames(self, mag, state=None, want_state=False, post=None, after_encoder=None, loss_inputs=None):
        """mag [B,T,F] frame-major magnitudes -> mask [B,T,F].  (Internal zero-copy entry used by
        cruse_b200.pipeline; ``forward`` wraps it with the reference's [B,1,T,F] layout.)
        ``post(mask, t0, t1)``: optional consumer of the mask, called on the stream that has just produced the frames [t0,t1)
        when the pipelined schedule runs (``self._post_ranges`` lists the ranges it was called for; empty = not called).
        ``loss_inputs`` = {"args": (S, layout_S, X, layout_X, rows[B*T]), "ready": callable -> events}: where the one-launch decoder
        runs it also leaves every frame's share of wo_male(S, mask*X, X) in ``rows`` (``self._loss_fused`` says whether it did).
        ``after_encoder(event=None)``: optional hook called once the encoder stages (pipelined schedule: and the layer-1 input
        projections, ``event``) have been queued -- work the caller wants to start behind them rather than beside them, e.g.
        the clean-speech STFT of the loss.
        ``state`` (cruse_b200.streaming.StreamState) carries one frame of history per encoder conv and the
        GRU hidden states between chunks of a stream; it is updated in place when ``want_state``."""
        _need_cuda(mag, "unet_2")
        B, T, F = mag.shape
        if F != self.in_feat:
            raise RuntimeError(f"unet_2: input has {F} bins, model was built for in_feat={self.in_feat}")
        n = self.laynum
        train = self.training
        h = mag.view(B, T, 1, F)
        enc, skips = [], []
        # eval, whole utterances: the skip convs leave the critical path and run beside the GRU wavefront
        overlap = (not train and ops.OVERLAP_SKIPS and ops.get_conv_mode() == "tf32" and F == 256
                   and self.gru.uses_wavefront(B, T, state.gru if state is not None else None, want_state))
        hists = state.hist if state is not None and state.hist else [None] * n
        new_hist = []
        folds = {}
        if not train:                                   # all eval-mode BatchNorm folds of the pass in one launch
            names = [f"bn{k}" for k in range(1, n + 1)] + [f"bn{k}_t" for k in range(n, 1, -1)]
            folds = dict(zip(names, ops.bn_fold_many([getattr(self, nm) for nm in names])))
        # the decoder launch with the skip convs inside keeps 12 frames in flight per SM: the last wavefront chunk is cut so that its
        # frames are ONE round of that launch
        one_round = 0
        if (ops.FUSE_DECODER and ops.FUSE_SKIP34 and n == 4 and tuple(self.ch) == (1, 8, 16, 32, 64) and F == 256
                and self.act_kind in ("relu", "prelu")):
            one_round = (torch.cuda.get_device_properties(mag.device).multi_processor_count * 12) // B
        plan = self.gru.plan(B, T, mag.device, last_chunk=one_round) if (overlap and ops.PIPELINE_EDGES) else None
        self._post_ranges = []
        self._loss_fused = False
        if plan is not None:
            return self._forward_frames_pipelined(mag, plan, folds, post, after_encoder, loss_inputs)
        for k in range(1, n + 1):                                            # :149-152 repaired
            if want_state:
                new_hist.append(h[:, -1].contiguous())
            if overlap and k == n:
                # the last encoder stage writes the GRU input TIME-MAJOR [T,B,C,F']: chunks of frames become contiguous rows
                scale, shift = folds[f"bn{k}"]
                conv = getattr(self, f"conv{k}")
                h = ops.conv_fwd_tm(h, conv.weight, conv.bias, scale, shift, self._alpha(f"act{k}"), self.act_kind, 2, 2, B, T,
                                    False, True)
            else:
                h = self._stage(h, getattr(self, f"conv{k}"), getattr(self, f"bn{k}"), self._alpha(f"act{k}"), (2, 2), train,
                                hist=hists[k - 1], fold=folds.get(f"bn{k}"))
            enc.append(h)
            if not overlap:
                skips.append(ops.conv_fwd(h, getattr(self, f"skip_connect_{k}").weight, None, None, None, None,
                                          "none", 1, 1))                                             # :153-156
        e4 = enc[-1]
        if after_encoder is not None:
            after_encoder()
        C4, F4 = e4.shape[2], e4.shape[3]
        D = C4 * F4
        side = None
        if overlap:
            dev = mag.device
            main = torch.cuda.current_stream(dev)
            s_skip = self._skip_stream(dev)
            skips = [None] * n

            def side(fork_event):
                # the four skip convs (:153-156) on a low-priority stream beside the recurrence; skip4 first (LayerNorm 2 adds it)
                s_skip.wait_event(fork_event)
                ops.set_conv_max_ctas(ops.SKIP_MAX_CTAS)
                try:
                    with torch.cuda.stream(s_skip):
                        for k in range(n, 0, -1):
                            wk = getattr(self, f"skip_connect_{k}").weight
                            if k == n:      # reads the time-major e4, writes frame order
                                skips[k - 1] = ops.conv_fwd_tm(enc[k - 1], wk, None, None, None, None, "none", 1, 1, B, T, True, False)
                            else:
                                skips[k - 1] = ops.conv_fwd(enc[k - 1], wk, None, None, None, None, "none", 1, 1)
                            skips[k - 1].record_stream(main)
                            if k == n:
                                ev4 = torch.cuda.Event()
                                ev4.record(s_skip)
                        self._skips_done = torch.cuda.Event()
                        self._skips_done.record(s_skip)
                finally:
                    ops.set_conv_max_ctas(0)
                return skips[n - 1].view(B, T, D), ev4

        # layer-by-layer eval path (streaming state carry, short inputs): LayerNorm 2 + skip 4 + the whole decoder as ONE launch
        fuse_tail = (not train and not overlap and ops.FUSE_DECODER and ops.get_conv_mode() == "tf32" and n == 4 and F == 256
                     and tuple(self.ch) == (1, 8, 16, 32, 64) and self.act_kind in ("relu", "prelu")
                     and not self.gru.uses_wavefront(B, T, state.gru if state is not None else None, want_state))
        g = self.gru.forward_frames(e4.view(T, B, D) if overlap else e4.view(B, T, D),
                                    residual=None if (overlap or fuse_tail) else skips[-1].view(B, T, D),
                                    state=state.gru if state is not None else None, want_state=want_state, side=side,
                                    time_major=overlap, skip_ln2=fuse_tail)                          # :158-160
        if overlap:
            torch.cuda.current_stream(mag.device).wait_event(self._skips_done)
        if want_state:
            g, gru_state = g
            state.hist, state.gru = new_hist, gru_state
        if fuse_tail:                                                                               # :51,160-164 repaired
            names = [f"conv{k}_t" for k in range(n, 0, -1)]
            image = ops.decoder_fused_prep(
                [getattr(self, nm).weight for nm in names], [getattr(self, nm).bias for nm in names],
                [folds[f"bn{k}_t"][0] for k in range(n, 1, -1)], [folds[f"bn{k}_t"][1] for k in range(n, 1, -1)],
                [self._alpha(f"act{k}_t") for k in range(n, 1, -1)] if self.act_kind == "prelu" else None, self.act_kind)
            mask = torch.empty(B, T, F, device=mag.device, dtype=torch.float32)
            ops.decoder_fused_range(g.view(B, T, D), self.gru.ln2.weight, self.gru.ln2.bias, self.gru.ln2.eps,
                                    [skips[k - 1] for k in range(n, 0, -1)], image, mask, 0, T)
            return mask
        out = g.view(B, T, C4, F4)
        for k in range(n, 1, -1):                                                                   # :161-163 repaired
            conv, bn = getattr(self, f"conv{k}_t"), getattr(self, f"bn{k}_t")
            alpha = self._alpha(f"act{k}_t")
            if not train:
                scale, shift = folds[f"bn{k}_t"]
                out = ops.convT_fwd(out, conv.weight, conv.bias, scale, shift, alpha, self.act_kind,
                                    skips[k - 2], self.freqs[k - 1])
            else:
                z, stats = ops.convT_fwd(out, conv.weight, conv.bias, None, None, None, "none", None,
                                         self.freqs[k - 1], want_stats=True)
                scale, shift, _, _ = ops.bn_finalize(stats, B * T * self.freqs[k - 1], bn)
                out = ops.bn_act_fwd(z, scale, shift, alpha, self.act_kind, skip=skips[k - 2])
        mask = ops.convT_fwd(out, self.conv1_t.weight, self.conv1_t.bias, None, None, None, "sigmoid", None,
                             self.freqs[0])                                                          # :164
        mask = mask.view(B, T, F)
        if self.gru._wavefront_err is not None:
            ops.poison_on_error(self.gru._wavefront_err, [mask])          # a timed-out flag spin must not return numbers
        return mask

    def wavefront_error_flags(self):
        """device flags (possibly none) that a flag-synchronised wavefront of the LAST forward call sets when a bounded spin
        timed out; ``ops.raise_if_wavefront_failed`` reads them on the host"""
        f = getattr(self.gru, "_wavefront_err", None)
        return [] if f is None else [f]

    def forward(self, x):
        """x [B,1,T,F] float32 CUDA -> mask [B,1,T,F] (model/cruse_net.py:147-165)."""
        _need_cuda(x, "unet_2")
        if x.dim() != 4 or x.shape[1] != self.ch[0] or self.ch[0] != 1:
            raise RuntimeError(f"unet_2: expected input [B,1,T,F], got {tuple(x.shape)}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .autograd import unet2_autograd_forward  # backward kernels
            return unet2_autograd_forward(self, x)
        B, _, T, F = x.shape
        return self.forward_frames(x.contiguous().view(B, T, F)).view(B, 1, T, F)

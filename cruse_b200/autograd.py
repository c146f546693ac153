"""Training path: forward with saved intermediates + backward of the U-Net on the sm_100a kernels
(SURVEY.md section 8 row a9 -- what autograd/cuDNN do on the reference path for model/cruse_net.py:14-55,129-165).

``unet2_autograd_forward(model, x)`` is what ``unet_2.forward`` calls when gradients are enabled: one
``torch.autograd.Function`` whose inputs are the module's parameters, so ``loss.backward()``,
optimizers and DDP see ordinary ``.grad`` tensors.  All arithmetic is in ``libcruse_sm100.so``; torch
only sequences launches and owns the buffers.
"""
from __future__ import annotations

import torch

from . import ops


# ------------------------------------------------------------------------------------------------
# grouped-GRU layer
# ------------------------------------------------------------------------------------------------
def gru_layer_fwd_train(x2d, grus, B, T, interleave, side=None, beside=None):
    """x2d [B*T, G*H] -> (y [B,T,G*H], saved); ``beside(after_event)`` queues side work next to the recurrence (which
    occupies 8*G*ceil(B/32) SMs) -- called right after the recurrence has been launched on a high-priority stream, see
    _SideWork.critical"""
    w_ih = [g.weight_ih_l0 for g in grus]
    # numeric mode: ops.GRU_IH_MODE / ops.GRU_SEQ_MODE ("tf32" = tensor cores, the product default; "fp32" = the exact twins of
    # csrc/gru_exact.cu, for the end-to-end gradient parity check against the oracle's autograd)
    xproj = ops.gru_ih_gemm(x2d, w_ih, [g.bias_ih_l0 for g in grus], [g.bias_hh_l0 for g in grus])
    run = lambda: ops.gru_seq_fwd(xproj, [g.weight_hh_l0 for g in grus], [g.bias_hh_l0 for g in grus], B, T,
                                  interleave=interleave, want_gates=True)
    if side is not None and side.enabled and beside is not None:
        ready = side.mark()
        y, gates = side.critical(run, ready)
        beside(ready)
    else:
        y, gates = run()
        if beside is not None:
            beside(None)
    return y, (x2d, y, gates)


class _SideWork:
    """Weight-gradient kernels are off the backward's critical path (nothing downstream reads them), while the two BPTT
    launches on it occupy only 8*G*slices of the 148 SMs for ~0.6 ms each.  ``run`` enqueues such work on a second stream
    behind an event recorded on the main stream; ``join`` (before backward returns) makes the main stream wait for it.
    Every tensor handed to the side stream is kept alive until the join so that the caching allocator cannot give its
    block to a later main-stream allocation while the side stream is still reading it."""

    def __init__(self, device, enabled, hp_priority=-1):
        self.enabled = enabled and device.type == "cuda"
        self.keep = []
        self.used = set()                  # lanes with work since the last join (an idle lane is not part of a graph capture)
        if self.enabled:
            from .pipeline import _side_stream
            self.main = torch.cuda.current_stream(device)
            self.side = _side_stream(device)
            self.lanes = [self.side, _side_stream(device, 1), _side_stream(device, 2)]   # lanes 1, 2: work that must not queue behind lane 0's
            self.hp = _priority_stream(device, hp_priority)

    def mark(self):
        """event at the current point of the main stream (None when disabled)"""
        if not self.enabled:
            return None
        ev = torch.cuda.Event()
        ev.record(self.main)
        return ev

    def mark_side(self):
        """event at the current point of the side stream: what the main stream waits for before it reads a side result early"""
        ev = torch.cuda.Event()
        ev.record(self.side)
        return ev

    def critical(self, fn, after):
        """run ``fn`` (the BPTT launch) on a high-priority stream behind ``after`` and let the main stream continue behind
        it: work queued on the side stream behind the same event then starts AFTER the BPTT has taken its SMs (a cluster
        needs 8 free SMs in one GPC; persistent side CTAs spread over all GPCs would otherwise keep it waiting)."""
        if not self.enabled:
            return fn()
        self.hp.wait_event(after)
        with torch.cuda.stream(self.hp):
            out = fn()
            done = torch.cuda.Event()
            done.record(self.hp)
        self.main.wait_event(done)
        return out

    def run(self, fn, *tensors, after=None, max_ctas=0, lane=0):
        if not self.enabled:
            return fn()
        ev = after if after is not None else self.mark()
        stream = self.lanes[lane]
        self.used.add(lane)
        stream.wait_event(ev)
        self.keep.extend(tensors)
        if max_ctas:
            ops.set_conv_max_ctas(max_ctas)
        try:
            with torch.cuda.stream(stream):
                out = fn()
        finally:
            if max_ctas:
                ops.set_conv_max_ctas(0)
        self.keep.append(out)
        return out

    def join(self):
        if self.enabled:
            for lane in sorted(self.used):
                ev = torch.cuda.Event()
                ev.record(self.lanes[lane])
                self.main.wait_event(ev)
        self.used.clear()
        self.keep.clear()


_hp_streams = {}


def _priority_stream(device, priority=-1):
    """one stream per (device, priority); CUDA priorities: 0 = default (lowest), negative = served first"""
    key = (device.type, device.index, priority)
    if key not in _hp_streams:
        _hp_streams[key] = torch.cuda.Stream(device=device, priority=priority)
    return _hp_streams[key]


def _splitk(M):
    """K = B*T of the weight-gradient GEMMs is split so that ~all SMs get a CTA (3 m-tiles x G x splitk)."""
    return max(1, min(16, M // 1024))


def gru_w_ih_transposed(grus):
    """[H, 3H] copies of the input weights: the K-major B operand of dx = dxproj . W_ih"""
    return [g.weight_ih_l0.detach().t().contiguous() for g in grus]


def gru_shifted_state_transposed(saved, G, H, B, T, interleave):
    """hT[g][i][b*T + t] = h_{t-1}[b, g, i] (0 at t = 0) from the saved layer output: the K-major B operand of dW_hh where the
    output's features are interleaved over the groups (layer 1, model/cruse_net.py:43-45) and a group's columns are not contiguous"""
    y = saved[1]
    y_fs, y_gs = (G, 1) if interleave else (1, H)
    return ops.transpose_gcm(y, B * T, G, H, G * H, y_gs, y_fs, shift_T=T, Bn=B)


def gru_layer_bwd(dy, saved, grus, B, T, interleave, need_dx=True, side=None, beside=None, own_wgrads_on_side=True,
                  defer_wgrads=False, dx_addend=None, early=None, sink=None):
    """dy [B,T,G*H] (layout of y) -> (dx [B*T, G*H] | None, {param: grad}); with ``side`` the weight gradients are
    enqueued on the side stream (valid on the main stream after ``side.join()``) and ``beside(after_event)`` is called right
    after the BPTT launch to queue side work that should run next to it.  ``defer_wgrads``: the weight-gradient GEMMs are not
    launched; a third return value ``wg()`` launches them (on whatever stream is current) and fills the dict.
    ``dx_addend()`` -> [B*T, G*H] tensor added to dx in the GEMM epilogue (tensor-core mode; called after the BPTT launch).
    ``early`` (dict): work that needs only saved forward tensors / parameters, done ahead by the caller -- "hT" = the transposed
    h_{t-1} (gru_shifted_state_transposed), "w_t" = gru_w_ih_transposed.
    ``sink`` ({parameter: tensor of its shape}): where the weight gradients -- 96 % of the model's gradient bytes -- are written
    instead of fresh buffers (pipeline.CapturedTrainStep: views of its flat all_reduce buffer, so that they need no copy there)."""
    x2d, y, gates = saved
    G = len(grus)
    H = grus[0].hidden_size
    M = B * T
    w_ih = [g.weight_ih_l0 for g in grus]
    w_hh = [g.weight_hh_l0 for g in grus]
    if side is not None and side.enabled:
        ready = side.mark()
        dxproj, dpre, dbias, _ = side.critical(lambda: ops.gru_seq_bwd(dy, y, gates, w_hh, B, T, interleave), ready)
        if beside is not None:
            beside(ready)
    else:
        dxproj, dpre, dbias, _ = ops.gru_seq_bwd(dy, y, gates, w_hh, B, T, interleave)
        if beside is not None:
            beside(None)
    dev = dy.device
    grads = {}
    # tensor-core GEMMs read the factors where they lie (MN-major operand descriptors); the exact-fp32 twins (parity mode)
    # keep the K-major contract and therefore the transposed copies
    in_place = ops.GRU_IH_MODE == "tf32" and ops.GEMM_MN_MAJOR
    # ---- bias gradients: xproj carried b_ih (all gates) + b_hh (r,z); W_hn.h carried b_hh (n)
    for gi, g in enumerate(grus):
        grads[g.bias_ih_l0] = dbias[gi, :3].reshape(3 * H)
        grads[g.bias_hh_l0] = torch.cat([dbias[gi, 0], dbias[gi, 1], dbias[gi, 3]])
    # ---- dx = dxproj . W_ih   ([M,3H] x [3H,H]); B operand K-major -> W_ih^T copies (0.75 MB each; reading W_ih in place as an
    # MN-major operand works too but was measured slower here: 109 vs 80 us beside the same side work)
    dx = None
    if need_dx:
        dx = torch.empty(M, G * H, device=dev, dtype=torch.float32)
        add = dx_addend() if dx_addend is not None else None
        w_t = (early or {}).get("w_t")
        if w_t is None:
            w_t = gru_w_ih_transposed(grus)
        if in_place:
            ops.gemm_tc([dxproj[:, gi] for gi in range(G)], w_t, [dx[:, gi * H:] for gi in range(G)],
                        M, H, 3 * H, G * 3 * H, 3 * H, G * H,
                        addend=[add.view(M, G * H)[:, gi * H:] for gi in range(G)] if add is not None else None)
        else:
            ops.gemm_tn_tc([dxproj[:, gi] for gi in range(G)], w_t, [dx[:, gi * H:] for gi in range(G)],
                           M, H, 3 * H, G * 3 * H, 3 * H, G * H)
            if add is not None:
                ops.colsum(add, 1, dx.numel(), dx, accumulate=True)
    # ---- weight gradients: dW_ih = dxproj^T . x, dW_hh = dpre^T . h_{t-1}; the reduction index is (b,t)
    y_fs, y_gs = (G, 1) if interleave else (1, H)
    plane = 3 * H * H

    def weight_grads():
        sk = _splitk(M)
        part = torch.empty(2, G, sk, plane, device=dev, dtype=torch.float32)
        if in_place:
            # A = dxproj / dpre [M, G, 3H] and B = x [M, G*H] are read MN-major straight from where the BPTT / the forward
            # left them.  h_{t-1}: the same rows of y one frame earlier (b_kshift = 1) with dpre's t = 0 rows zeroed (zero
            # initial state; they would otherwise meet the previous utterance's last frame) -- when y is concatenated;
            # the interleaved layer-1 output (feature = h*G + g) is not contiguous per group and keeps its transposed copy
            dpre.view(B, T, G * 3 * H)[:, 0].zero_()
            ops.gemm_tc([dxproj[:, gi] for gi in range(G)], [x2d[:, gi * H:] for gi in range(G)], [part[0, gi] for gi in range(G)],
                        3 * H, H, M, G * 3 * H, G * H, H, a_mn=True, b_mn=True, splitk=sk, c_plane=plane)
            if interleave:
                hT = (early or {}).get("hT")
                if hT is None:
                    hT = gru_shifted_state_transposed(saved, G, H, B, T, interleave)         # [G, H, M4]: h_{t-1}, K-major
                M4 = hT.shape[-1]
                ops.gemm_tc([dpre[:, gi] for gi in range(G)], [hT[gi] for gi in range(G)], [part[1, gi] for gi in range(G)],
                            3 * H, H, M, G * 3 * H, M4, H, a_mn=True, splitk=sk, c_plane=plane)
            else:
                yv = y.view(M, G * H)
                ops.gemm_tc([dpre[:, gi] for gi in range(G)], [yv[:, gi * H:] for gi in range(G)], [part[1, gi] for gi in range(G)],
                            3 * H, H, M, G * 3 * H, G * H, H, a_mn=True, b_mn=True, b_kshift=1, splitk=sk, c_plane=plane)
        else:
            dxT = ops.transpose_gcm(dxproj, M, G, 3 * H, G * 3 * H, 3 * H, 1)           # [G, 3H, M4]
            dpT = ops.transpose_gcm(dpre, M, G, 3 * H, G * 3 * H, 3 * H, 1)
            xT = ops.transpose_gcm(x2d, M, G, H, G * H, H, 1)                           # [G, H, M4]
            hT = ops.transpose_gcm(y, M, G, H, G * H, y_gs, y_fs, shift_T=T, Bn=B)      # h_{t-1}
            M4 = dxT.shape[-1]
            ops.gemm_tn_tc([dxT[gi] for gi in range(G)], [xT[gi] for gi in range(G)], [part[0, gi] for gi in range(G)],
                           3 * H, H, M, M4, M4, H, splitk=sk, c_plane=plane)
            ops.gemm_tn_tc([dpT[gi] for gi in range(G)], [hT[gi] for gi in range(G)], [part[1, gi] for gi in range(G)],
                           3 * H, H, M, M4, M4, H, splitk=sk, c_plane=plane)
        dw_ = [[None] * G, [None] * G]
        for a in range(2):
            for gi in range(G):
                v = sink.get(grus[gi].weight_ih_l0 if a == 0 else grus[gi].weight_hh_l0) if sink else None
                ok = v is not None and v.is_contiguous() and v.numel() == plane and v.dtype == torch.float32 and v.device == dev
                dw_[a][gi] = ops.colsum(part[a, gi], sk, plane, v.view(plane) if ok else torch.empty(plane, device=dev, dtype=torch.float32))
        return dw_

    def wg(on_side=own_wgrads_on_side):
        dw = side.run(weight_grads, dxproj, dpre, x2d, y) if (side is not None and on_side) else weight_grads()
        for gi, g in enumerate(grus):
            grads[g.weight_ih_l0] = dw[0][gi].view(3 * H, H)
            grads[g.weight_hh_l0] = dw[1][gi].view(3 * H, H)

    if defer_wgrads:
        return dx, grads, wg
    wg()
    return dx, grads


# ------------------------------------------------------------------------------------------------
# the whole U-Net
# ------------------------------------------------------------------------------------------------
def _bn_buffers(bn):
    return bn.weight, bn.bias


class _Unet2Fn(torch.autograd.Function):
    """mag [B,T,F] -> mask [B,T,F]; differentiable w.r.t. every parameter of the module (not w.r.t. mag)."""

    @staticmethod
    def forward(ctx, model, mag, names, *params):
        m = model
        n = m.laynum
        B, T, F = mag.shape
        act = m.act_kind
        train = m.training
        sv = {}
        h = mag.view(B, T, 1, F)
        enc_in, enc_z, enc_bn, skips, enc_out = [], [], [], [], []
        counters = []                                             # num_batches_tracked of the train-mode BatchNorms: one launch at the end
        for k in range(1, n + 1):                                                    # cruse_net.py:149-156
            conv, bn = getattr(m, f"conv{k}"), getattr(m, f"bn{k}")
            alpha = m._alpha(f"act{k}")
            enc_in.append(h)
            z, stats = ops.conv_fwd(h, conv.weight, conv.bias, None, None, None, "none", 2, 2, want_stats=True)
            Fo = z.shape[3]
            if train:
                scale, shift, mean, invstd = ops.bn_finalize(stats, B * T * Fo, bn, counters=counters)
            else:
                scale, shift = ops.bn_fold(bn)
                mean, invstd = bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)
            h = ops.bn_act_fwd(z, scale, shift, alpha, act)
            enc_z.append(z)
            enc_bn.append((scale, shift, mean, invstd))
            enc_out.append(h)
        e4 = h
        # the four skip convs (:153-156) run beside the layer-1 recurrence (64 of 148 SMs busy) on the side stream
        # (ops.FWD_SIDE_SKIPS; measured on one box, repeated: 6.06 ms per step with, 6.27 without, 6.86 with no side streams at all)
        side = _SideWork(mag.device, ops.OVERLAP_BWD and ops.FWD_SIDE_SKIPS)
        cap = 0
        if side.enabled and ops.BWD_SIDE_CAP:
            free_sms = torch.cuda.get_device_properties(mag.device).multi_processor_count - 8 * len(m.gru.gru_list1) * ((B + 31) // 32)
            cap = free_sms if free_sms >= 32 else 0

        def skip_convs():
            return [ops.conv_fwd(enc_out[k - 1], getattr(m, f"skip_connect_{k}").weight, None, None, None, None, "none", 1, 1)
                    for k in range(1, n + 1)]

        def beside(ev):
            skips.extend(side.run(skip_convs, *enc_out, after=ev, max_ctas=cap))
        C4, F4 = e4.shape[2], e4.shape[3]
        D = C4 * F4
        gru = m.gru
        x2d = e4.view(B * T, D)
        y1, sv1 = gru_layer_fwd_train(x2d, gru.gru_list1, B, T, True, side=side, beside=beside)   # :42-45
        z1, mean1, rstd1 = ops.layernorm_fwd(y1, gru.ln1.weight, gru.ln1.bias, gru.ln1.eps, want_stats=True)
        hT1 = []

        def beside2(ev):
            # backward's K-major copy of layer 1's shifted output (the one transposed operand left, gru_layer_bwd): it needs y1 only,
            # and the layer-2 recurrence leaves 84 SMs idle
            if ops.GRU_IH_MODE == "tf32" and ops.GEMM_MN_MAJOR and ops.BWD_PRIORITY:
                hT1.append(side.run(lambda: gru_shifted_state_transposed(sv1, len(gru.gru_list1), gru.gru_list1[0].hidden_size, B, T, True),
                                    y1, after=ev))
        y2, sv2 = gru_layer_fwd_train(z1.view(B * T, D), gru.gru_list2, B, T, False, side=side,
                                      beside=beside2 if side.enabled else None)      # :48-50
        side.join()
        out, mean2, rstd2 = ops.layernorm_fwd(y2, gru.ln2.weight, gru.ln2.bias, gru.ln2.eps,
                                              residual=skips[-1].view(B, T, D), want_stats=True)   # :51,160
        out = out.view(B, T, C4, F4)
        dec_in, dec_z, dec_bn = [], [], []
        for k in range(n, 1, -1):                                                    # :161-163
            conv, bn = getattr(m, f"conv{k}_t"), getattr(m, f"bn{k}_t")
            alpha = m._alpha(f"act{k}_t")
            dec_in.append(out)
            z, stats = ops.convT_fwd(out, conv.weight, conv.bias, None, None, None, "none", None, m.freqs[k - 1],
                                     want_stats=True)
            if train:
                scale, shift, mean, invstd = ops.bn_finalize(stats, B * T * m.freqs[k - 1], bn, counters=counters)
            else:
                scale, shift = ops.bn_fold(bn)
                mean, invstd = bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)
            out = ops.bn_act_fwd(z, scale, shift, alpha, act, skip=skips[k - 2])
            dec_z.append(z)
            dec_bn.append((scale, shift, mean, invstd))
        mask = ops.convT_fwd(out, m.conv1_t.weight, m.conv1_t.bias, None, None, None, "sigmoid", None, m.freqs[0])  # :164
        ops.bump_counters(counters)
        ctx.model = m
        ctx.names = names
        ctx.train = train
        ctx.dims = (B, T, F, D, C4, F4)
        ctx.sv = dict(enc_in=enc_in, enc_z=enc_z, enc_bn=enc_bn, e4=e4, sv1=sv1, y1=y1, ln1=(mean1, rstd1), sv2=sv2, y2=y2,
                      hT1=hT1[0] if hT1 else None,
                      ln2=(mean2, rstd2), dec_in=dec_in, dec_z=dec_z, dec_bn=dec_bn, d2=out, mask=mask)
        return mask.view(B, T, F)

    @staticmethod
    def backward(ctx, dmask):
        """Schedule (ops.BWD_PRIORITY): the chain of kernels that depend on each other -- loss gradient -> decoder -> BPTT 2
        -> BPTT 1 -> encoder -- runs on a stream of priority -1 and the two BPTT launches at -2, while everything nothing
        downstream waits for (all weight gradients) and the skip convs' data gradients (needed only at the encoder stage
        they are added to) go to the side stream at priority 0: the block scheduler then fills the SMs the chain leaves
        idle instead of making the chain queue behind a full-grid weight-gradient kernel."""
        dev = dmask.device
        if ctx.sv is not None and ops.OVERLAP_BWD and ops.BWD_PRIORITY and dev.type == "cuda":
            cur = torch.cuda.current_stream(dev)
            chain = _priority_stream(dev, -1)
            chain.wait_stream(cur)
            with torch.cuda.stream(chain):
                out = _Unet2Fn._backward(ctx, dmask, -2)
            cur.wait_stream(chain)
            return out
        return _Unet2Fn._backward(ctx, dmask, -1)

    @staticmethod
    def _backward(ctx, dmask, bptt_priority):
        m, sv = ctx.model, ctx.sv
        if sv is None:
            raise RuntimeError("cruse_b200.unet_2: backward was already run through this forward pass; its saved activations were "
                               "released (they are plain device buffers, not autograd-saved tensors, so retain_graph=True cannot "
                               "keep them).  Run the forward again for a second backward.")
        n = m.laynum
        B, T, F, D, C4, F4 = ctx.dims
        act, train = m.act_kind, ctx.train
        G = {}                                                    # parameter tensor -> gradient
        dmask = dmask.contiguous()
        side = _SideWork(dmask.device, ops.OVERLAP_BWD, hp_priority=bptt_priority)
        early = side.enabled and ops.BWD_PRIORITY                 # the re-ordered schedule of the docstring above
        cap = 0                                                   # persistent side kernels sized to the SMs the BPTT leaves free
        if side.enabled and ops.BWD_SIDE_CAP:
            ng = len(m.gru.gru_list1)
            clusters = ng * ((B + 15) // 16)                      # gru_bwd_tc.cu launch_bwd_nc: 16-utterance slices if they fit
            if clusters > ops.gru_seq_max_clusters(m.gru.gru_list1[0].hidden_size):
                clusters = ng * ((B + 31) // 32)
            free_sms = torch.cuda.get_device_properties(dmask.device).multi_processor_count - 8 * clusters
            cap = free_sms if free_sms >= 32 else 0
        deferred = []                                             # decoder / skip weight gradients: run beside the BPTT
        ahead1, ahead2 = {}, {}                                   # small things the GRU backward needs, made early on the side
        if early:
            ahead2["w_t"] = side.run(lambda: gru_w_ih_transposed(m.gru.gru_list2))
            ahead1["w_t"] = side.run(lambda: gru_w_ih_transposed(m.gru.gru_list1))
            w_t_ready = side.mark_side()
        # ---- last decoder stage: mask = sigmoid(convT(d2))           cruse_net.py:164
        dz = ops.sigmoid_bwd(dmask.view(B, T, 1, F), sv["mask"])

        def mask_layer_wgrad():
            G[m.conv1_t.weight], G[m.conv1_t.bias] = ops.convT_wgrad(sv["d2"], dz)
        if early:
            side.run(mask_layer_wgrad, dz)
        else:
            mask_layer_wgrad()
        d_out = ops.convT_dgrad(dz, m.conv1_t.weight, sv["d2"].shape)
        # ---- decoder stages k = 2..n: out = act(BN(convT_k(in))) + skip_{k-1}       :161-163
        dskip = [None] * n                                        # gradient w.r.t. skip_k output (index k-1)
        for k in range(2, n + 1):
            i = n - k                                             # position in the forward lists (k = n first)
            conv, bn = getattr(m, f"conv{k}_t"), getattr(m, f"bn{k}_t")
            alpha = m._alpha(f"act{k}_t")
            scale, shift, mean, invstd = sv["dec_bn"][i]
            z, x_in = sv["dec_z"][i], sv["dec_in"][i]
            dskip[k - 2] = d_out
            dzk, dgamma, dbeta, dalpha = ops.bn_act_bwd(d_out, z, scale, shift, alpha, act, mean, invstd, bn.weight,
                                                        B * T * z.shape[3], training=train)
            G[bn.weight], G[bn.bias] = dgamma, dbeta
            if dalpha is not None:
                G[getattr(m, f"act{k}_t").weight] = dalpha
            deferred.append((conv, x_in, dzk))
            d_out = ops.convT_dgrad(dzk, conv.weight, x_in.shape)
        dskip[n - 1] = d_out                                      # out = g + skip4     :160

        def e_of(k_):                                             # output of encoder stage k_ (= input of its skip conv)
            return sv["e4"] if k_ == n else sv["enc_in"][k_]

        def skip_wgrad(k_):
            G[getattr(m, f"skip_connect_{k_}").weight], _ = ops.conv_wgrad(e_of(k_), dskip[k_ - 1], 1, 1, want_bias=False)

        late_skip = 1 if early else 0                             # the widest (slowest) skip weight gradient moves beside BPTT 1:
                                                                  # beside BPTT 2 the queue was 0.1 ms longer than the BPTT

        def decoder_weight_grads():
            for conv_, x_, dz_ in deferred:
                G[conv_.weight], G[conv_.bias] = ops.convT_wgrad(x_, dz_)
            for k_ in range(n, late_skip, -1):
                skip_wgrad(k_)

        sd = [None] * n                                           # skip conv k's data gradient (index k-1), early schedule only

        def skip_data_grads():
            for k_ in range(n, 0, -1):
                sd[k_ - 1] = ops.conv_dgrad(dskip[k_ - 1], getattr(m, f"skip_connect_{k_}").weight, e_of(k_).shape, 1, 1)
        # ---- GGRU                                                              :37-55
        gru = m.gru
        dgo = d_out.view(B * T, D)
        dy2, G[gru.ln2.weight], G[gru.ln2.bias] = ops.layernorm_bwd(dgo, sv["y2"].view(B * T, D), gru.ln2.weight, *sv["ln2"])
        side_in = [t for d in deferred for t in d[1:]] + dskip
        if early:
            side.main.wait_event(w_t_ready)
        sink = getattr(m, "_grad_sink", None)                      # set by CapturedTrainStep for the duration of its capture
        dz1, g2, wg2 = gru_layer_bwd(dy2.view(B, T, D), sv["sv2"], gru.gru_list2, B, T, False, side=side, defer_wgrads=True, early=ahead2,
                                     sink=sink, beside=lambda ev: side.run(decoder_weight_grads, *side_in, after=ev, max_ctas=cap))
        if not early:
            wg2()
        dy1, G[gru.ln1.weight], G[gru.ln1.bias] = ops.layernorm_bwd(dz1, sv["y1"].view(B * T, D), gru.ln1.weight, *sv["ln1"])

        sd_ready = []

        ahead = ahead1

        def beside_bptt1(ev):                                     # what the chain needs first goes first
            side.run(skip_data_grads, *dskip, after=ev, max_ctas=cap)
            sd_ready.append(side.mark_side())
            side.run(lambda: wg2(on_side=False), after=ev)        # layer 2's weight gradients (its BPTT is long done)
            for k_ in range(late_skip, 0, -1):
                side.run(lambda k_=k_: skip_wgrad(k_), after=ev, max_ctas=cap)
            if sv.get("hT1") is not None:                          # layer 1's h_{t-1}, transposed: made in the forward pass
                ahead["hT"] = sv["hT1"]
            elif ops.GRU_IH_MODE == "tf32" and ops.GEMM_MN_MAJOR:
                ahead["hT"] = side.run(lambda: gru_shifted_state_transposed(sv["sv1"], len(gru.gru_list1), gru.gru_list1[0].hidden_size,
                                                                            B, T, True), after=ev)

        def skip4_path():                                         # out = g + skip4 (:160): both paths' gradients meet at e4
            side.main.wait_event(sd_ready[0])
            return sd[n - 1]
        de, g1, wg1 = gru_layer_bwd(dy1.view(B, T, D), sv["sv1"], gru.gru_list1, B, T, True, side=side, defer_wgrads=True,
                                    beside=beside_bptt1 if early else None, dx_addend=skip4_path if early else None, early=ahead,
                                    sink=sink)
        wg1(on_side=ops.BWD_SIDE_L1 or early)
        de = de.view(B, T, C4, F4)
        # ---- encoder stages k = n..1 with their skip convs                      :149-156
        for k in range(n, 0, -1):
            conv, bn, skipc = getattr(m, f"conv{k}"), getattr(m, f"bn{k}"), getattr(m, f"skip_connect_{k}")
            alpha = m._alpha(f"act{k}")
            e_k = e_of(k)                                         # output of stage k = input of stage k+1
            x_in, z = sv["enc_in"][k - 1], sv["enc_z"][k - 1]
            scale, shift, mean, invstd = sv["enc_bn"][k - 1]
            if not early:
                de = ops.conv_dgrad(dskip[k - 1], skipc.weight, e_k.shape, 1, 1, addend=de)
            dzk, dgamma, dbeta, dalpha = ops.bn_act_bwd(de, z, scale, shift, alpha, act, mean, invstd, bn.weight,
                                                        B * T * z.shape[3], training=train)
            G[bn.weight], G[bn.bias] = dgamma, dbeta
            if dalpha is not None:
                G[getattr(m, f"act{k}").weight] = dalpha

            def stage_wgrad(conv=conv, x_in=x_in, dzk=dzk):
                G[conv.weight], G[conv.bias] = ops.conv_wgrad(x_in, dzk, 2, 2)
            if early and k > 1:
                side.run(stage_wgrad, x_in, dzk, lane=1 + k % 2, max_ctas=ops.BWD_TAIL_CAP)  # own lanes (lane 0 holds layer 1's GEMMs): consecutive stages' weight
                                                                  # gradients run side by side instead of one behind the other
            else:
                stage_wgrad()                                     # stage 1's is the last kernel anything waits for: on the chain itself
            if k > 1:
                de = ops.conv_dgrad(dzk, conv.weight, x_in.shape, 2, 2, addend=sd[k - 2] if early else None)
        side.join()
        G.update(g2)
        G.update(g1)
        named = dict(m.named_parameters())
        out = []
        for nm in ctx.names:
            g = G.get(named[nm])
            out.append(g.reshape(named[nm].shape) if g is not None else None)
        ctx.sv = None
        return (None, None, None, *out)


def unet2_frames_autograd(model, mag):
    """mag [B,T,F] -> mask [B,T,F] with gradients to the module's parameters (NOT to ``mag``: the reference trains on fixed
    features, tools/train_stand.py; an input that requires grad is refused rather than silently given no gradient)."""
    if mag.requires_grad:
        raise RuntimeError("cruse_b200.unet_2: the input requires grad, but the sm_100a backward produces parameter gradients only "
                           "(no dL/d input); detach the input (INTEGRATION.md, limitations)")
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    params = [dict(model.named_parameters())[n] for n in names]
    return _Unet2Fn.apply(model, mag.contiguous(), tuple(names), *params)


def unet2_autograd_forward(model, x):
    """reference surface: x [B,1,T,F] -> mask [B,1,T,F]  (model/cruse_net.py:147-165)."""
    B, _, T, F = x.shape
    return unet2_frames_autograd(model, x.contiguous().view(B, T, F)).view(B, 1, T, F)


class _MaskApplyFn(torch.autograd.Function):
    """est[b,t,f,:] = mask[b,t,f] * X[b,t,f,:] for f < F, X beyond (utils/utils.py:417-420, mag_mapping)."""

    @staticmethod
    def forward(ctx, mask, X, n_fft, hop):
        from .acoustics import hann_window
        est, _ = ops.mask_istft_fwd(X, mask, hann_window(n_fft, n_fft, X.device), n_fft, hop, 0, want_est=True, want_wav=False)
        ctx.save_for_backward(X)
        ctx.F = mask.shape[-1]
        return est

    @staticmethod
    def backward(ctx, dest):
        (X,) = ctx.saved_tensors
        return ops.mask_bwd(dest.contiguous(), X, ctx.F), None, None, None


def mask_apply(mask, X, n_fft, hop):
    return _MaskApplyFn.apply(mask, X, n_fft, hop)


class _MaskIstftFn(torch.autograd.Function):
    """wav = istft(mask * X) (utils/utils.py:417-454), differentiable w.r.t. the mask: for time-domain losses (SI-SNR)."""

    @staticmethod
    def forward(ctx, mask, X, n_fft, hop, length):
        from .acoustics import hann_window
        window = hann_window(n_fft, n_fft, X.device)
        _, wav = ops.mask_istft_fwd(X, mask, window, n_fft, hop, length, want_est=False, want_wav=True)
        ctx.save_for_backward(X, window)
        ctx.geom = (n_fft, hop, mask.shape[-1])
        return wav

    @staticmethod
    def backward(ctx, dwav):
        X, window = ctx.saved_tensors
        n_fft, hop, F = ctx.geom
        dspec = ops.istft_bwd(dwav.contiguous(), window, n_fft, hop)          # adjoint of the iSTFT
        return ops.mask_bwd(dspec, X, F), None, None, None, None              # Re(conj(X) * dspec) per masked bin


def mask_istft_apply(mask, X, n_fft, hop, length):
    return _MaskIstftFn.apply(mask, X, n_fft, hop, length)

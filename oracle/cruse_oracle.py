"""CPU oracle for the CRUSE hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product path (``cruse_b200``)
never imports it and fails loudly when its CUDA library is missing.

What it is: a minimal-repair restatement of the reference's algorithm for the
path  STFT -> conv-recurrent U-Net -> mask*spectrum -> iSTFT -> weighted-magnitude
loss, built from stock ``torch.nn`` CPU ops (torch CPU fp32 is the arithmetic
authority).  Each function cites the reference lines it follows
(paths relative to /root/reference) and each repair cites the defect it fixes
(SURVEY.md Appendix A).

PARITY PIN STATUS: **pinned to outputs of the reference's own hot-path source, executed in the build
container** (the reference holds no golden vector, known-answer test or fixture for this path -- SURVEY.md
section 0 -- so the fixtures were generated from its code).  ``oracle/ref_extract.py`` cuts the class / function
source out of the reference files' ASTs (the modules themselves do not import) and runs it unmodified with
era-compatible torch spellings, or with asserted one-token repairs of genuine defects;
``oracle/make_golden.py`` stores the outputs as ``tests/golden/refx_*.npz``:

  GGRU                 model/cruse_net.py:14-51 unmodified, taken at ln1 / ln2 by forward hook (hidden 64 and 1024)
  unet_2 stages        the modules model/cruse_net.py:129-146 builds (conv3/bn3, conv4/bn4, skip convs, GGRU, ReLU)
                       composed as :151-156 with the App. A.1 slice repair; decoder stage vs cust_conv.py:65-113
  wo_male / rmse / c_rmse / sisnr / loss_func   loss_func/loss.py from source (one repaired index, :139)
  stft / istft         train_base/acoustics/feature.py:10-61 unmodified
  PreProcess           utils/utils.py:365-455 unmodified
  ConvSTFT             train_base/acoustics/conv_stft.py: stft unmodified; istft with four repairs (round trip exact)
  si_snr_loss, complex_mul, encoder stage (train/eval BN), grouped GRU with state: ``ref_*.npz`` (fragments that import)

``tests/test_oracle.py`` drives THIS file's classes with those fixtures.  What cannot be pinned by any reference
output: the composition of the repaired ``unet_2.forward`` (:147-165 cannot execute; its repairs are SURVEY App. A.1)
and the decoder's ``conv{k}_t`` / ``bn{k}_t`` modules, which the reference constructor never creates.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

EPS_MAG = 1e-8  # utils/utils.py:400  sqrt(re^2 + im^2 + 1e-8)


# --------------------------------------------------------------------------
# model/cruse_net.py:14-55  GGRU
# --------------------------------------------------------------------------
class GGRU(nn.Module):
    """Grouped 2-layer GRU + LayerNorm bottleneck (model/cruse_net.py:14-55).

    Kept exactly, including the layer-1 ``stack(dim=-1)+flatten`` interleave
    (:43-45, output feature index = h*G + g) versus the layer-2 ``cat`` (:48-50).
    Repair: :53 ``self.view`` -> ``out.view`` (App. A.1).
    """

    def __init__(self, in_features=None, out_features=None, mid_features=None,
                 hidden_size=1024, groups=2):
        super().__init__()
        hidden_size_t = hidden_size // groups
        self.gru_list1 = nn.ModuleList(
            [nn.GRU(hidden_size_t, hidden_size_t, 1, batch_first=True) for _ in range(groups)])
        self.gru_list2 = nn.ModuleList(
            [nn.GRU(hidden_size_t, hidden_size_t, 1, batch_first=True) for _ in range(groups)])
        self.ln1 = nn.LayerNorm(hidden_size)
        self.ln2 = nn.LayerNorm(hidden_size)
        self.groups = groups
        self.mid_features = mid_features

    def forward(self, x):
        out = x.transpose(1, 2).contiguous()                       # :39  [B,T,C,F']
        out = out.view(out.size(0), out.size(1), -1).contiguous()  # :40  [B,T,C*F']
        out = torch.chunk(out, self.groups, dim=-1)                # :42
        out = torch.stack([self.gru_list1[i](out[i])[0] for i in range(self.groups)], dim=-1)  # :43-44
        out = torch.flatten(out, start_dim=-2, end_dim=-1)         # :45
        out = self.ln1(out)                                        # :46
        out = torch.chunk(out, self.groups, dim=-1)                # :48
        out = torch.cat([self.gru_list2[i](out[i])[0] for i in range(self.groups)], dim=-1)    # :49-50
        out = self.ln2(out)                                        # :51
        out = out.view(out.size(0), out.size(1), x.size(1), -1).contiguous()  # :53 (repaired)
        out = out.transpose(1, 2).contiguous()                     # :54
        return out


def freq_pyramid(in_feat: int, nlayers: int):
    """Frequency sizes after each (k=3, s=2, p=1) encoder conv (model/cruse_net.py:134-138)."""
    f = [in_feat]
    for _ in range(nlayers):
        f.append((f[-1] + 2 * 1 - 3) // 2 + 1)
    return f


# --------------------------------------------------------------------------
# model/cruse_net.py:129-165  unet_2
# --------------------------------------------------------------------------
class unet_2(nn.Module):
    """Repaired ``unet_2`` (model/cruse_net.py:129-165); repairs = SURVEY App. A.1.

    Module creation order follows the reference loop (:137-143) so that default
    initialisation under a fixed seed is reproducible and ``state_dict()`` is the
    surface of SURVEY App. C.
    """

    def __init__(self, in_feat=161, ch=(1, 8, 16, 32, 64), stride=(1, 2), rnn_groups=4, act="relu"):
        super().__init__()
        self.laynum = len(ch) - 1
        self.ker_x = 2
        self.stride = tuple(stride)
        self.padding = [self.ker_x - stride[0], 3 - stride[1]]          # :136
        self.ch = tuple(ch)
        self.in_feat = in_feat
        self.freqs = freq_pyramid(in_feat, self.laynum)
        self.act_kind = act
        n = self.laynum
        for i in range(n):
            setattr(self, f"conv{i+1}", nn.Conv2d(ch[i], ch[i + 1], (self.ker_x, 3), self.stride, self.padding))  # :138
            tmp = n - i
            # :140 repaired: conv{tmp}_t is the transposed conv the forward uses (:161-164)
            setattr(self, f"conv{tmp}_t", nn.ConvTranspose2d(ch[tmp], ch[tmp - 1], (1, 3), self.stride))
            setattr(self, f"bn{i+1}", nn.BatchNorm2d(ch[i + 1]))                                                     # :141
            if tmp >= 2:  # :142 repaired index/name; no BN on the sigmoid output layer (:164)
                setattr(self, f"bn{tmp}_t", nn.BatchNorm2d(ch[tmp - 1]))
            # :143 repaired: padding (0,1) keeps F so the skip can be added (:160-163)
            setattr(self, f"skip_connect_{i+1}", nn.Conv2d(ch[i + 1], ch[i + 1], (1, 3), bias=False, padding=(0, 1)))
        # :144 repaired: hidden size from the real conv arithmetic (1024 at F=256)
        self.gru = GGRU(hidden_size=ch[-1] * self.freqs[-1], groups=rnn_groups)
        self.elu = nn.ReLU()                                              # :145 (named elu, is ReLU)
        self.fc = nn.Linear(in_feat, in_feat)                             # :146 unused, kept for state_dict
        if act == "prelu":  # optional (north_star); absent for act="relu" so state_dict stays reference-shaped
            for k in range(1, n + 1):
                setattr(self, f"act{k}", nn.PReLU(ch[k]))
            for k in range(n, 1, -1):
                setattr(self, f"act{k}_t", nn.PReLU(ch[k - 1]))
        elif act != "relu":
            raise ValueError(f"act must be 'relu' or 'prelu', got {act!r}")

    def _act(self, name, x):
        if self.act_kind == "relu":
            return self.elu(x)
        return getattr(self, name)(x)

    # the three stage expressions of the reference forward, one method each so that tests can drive a single
    # stage with reference-generated weights (tests/test_oracle.py)
    def enc_stage(self, k, x):
        """:149-152 repaired: act(bn_k(conv_k(x)[..., :-pad_t, :]))  (drop the look-ahead frame)."""
        z = getattr(self, f"conv{k}")(x)[..., :-self.padding[0], :]
        return self._act(f"act{k}", getattr(self, f"bn{k}")(z))

    def skip(self, k, e):
        """:153-156 repaired: skip_connect_k(e_k)."""
        return getattr(self, f"skip_connect_{k}")(e)

    def dec_stage(self, k, x, skip):
        """:161-163 repaired: act(bn_k_t(conv_k_t(x)[..., :F_{k-1}])) + skip_{k-1}."""
        z = getattr(self, f"conv{k}_t")(x)[..., : self.freqs[k - 1]]
        return self._act(f"act{k}_t", getattr(self, f"bn{k}_t")(z)) + skip

    def forward(self, x):
        n = self.laynum
        e = []
        out = x
        for k in range(1, n + 1):                                         # :149-152
            out = self.enc_stage(k, out)
            e.append(out)
        skips = [self.skip(k, e[k - 1]) for k in range(1, n + 1)]         # :153-156
        out = self.gru(e[-1]) + skips[-1]                                 # :158-160
        for k in range(n, 1, -1):                                         # :161-163 (chained)
            out = self.dec_stage(k, out, skips[k - 2])
        return torch.sigmoid(self.conv1_t(out)[..., : self.freqs[0]])     # :164


# --------------------------------------------------------------------------
# train_base/acoustics/feature.py:10-61   stft / istft
# --------------------------------------------------------------------------
def stft(y, n_fft, hop_length, win_length, pad_mode="reflect"):
    """feature.py:10-30 (pad_mode='constant' = PreProcess.pre_stft, utils/utils.py:396)."""
    assert y.dim() == 2
    return torch.stft(y, n_fft, hop_length, win_length, window=torch.hann_window(n_fft).to(y.device),
                      return_complex=True, center=True, pad_mode=pad_mode)


def istft(features, n_fft, hop_length, win_length, length=None, use_mag_phase=False):
    """feature.py:33-61; repair: complex input required by torch>=2 (App. A.2)."""
    if use_mag_phase:
        assert isinstance(features, (tuple, list))
        mag, phase = features
        features = torch.stack([mag * torch.cos(phase), mag * torch.sin(phase)], dim=-1)
    if not torch.is_complex(features):
        features = torch.view_as_complex(features.contiguous())
    return torch.istft(features, n_fft, hop_length, win_length,
                       window=torch.hann_window(n_fft).to(features.device), length=length, center=True)


# --------------------------------------------------------------------------
# utils/utils.py:365-455   PreProcess
# --------------------------------------------------------------------------
class PreProcess:
    """utils/utils.py:365-455, repaired for torch>=2 (return_complex, App. A.2)."""

    def __init__(self, win_len, win_inc, fft_len, win_type="hanning", post_process_mode="mag_mapping",
                 loss_mode="freq", use_cuda=False):
        self.win_len, self.win_inc, self.fft_len = win_len, win_inc, fft_len
        self.post_process_mode, self.loss_mode = post_process_mode, loss_mode
        if win_type != "hanning":
            raise ValueError("ERROR window type")
        self.window = torch.hann_window(fft_len)

    def pre_stft(self, inputs):
        c = torch.stft(inputs, n_fft=self.fft_len, hop_length=self.win_inc, win_length=self.win_len,
                       window=self.window, center=True, pad_mode="constant", return_complex=True)
        stft_inputs = torch.view_as_real(c).transpose(1, 3).contiguous()   # [B,2,T,F]  (:397)
        real = stft_inputs[:, 0, :, :]
        imag = stft_inputs[:, 1, :, :]
        spec_mags = torch.sqrt(real ** 2 + imag ** 2 + EPS_MAG)            # :400
        spec_phase = torch.atan2(imag, real)
        self.real, self.imag = real.unsqueeze(1), imag.unsqueeze(1)
        self.spec_mags, self.spec_phase = spec_mags.unsqueeze(1), spec_phase.unsqueeze(1)
        return stft_inputs, self.real, self.imag, self.spec_mags, self.spec_phase

    def masking(self, mask_real, mask_imag=None):                          # :417-433
        if self.post_process_mode == "mag_mapping":
            out_real, out_imag = mask_real * self.real, mask_real * self.imag
        elif self.post_process_mode == "complex_mapping":
            out_real, out_imag = mask_real * self.real, mask_imag * self.imag
        elif self.post_process_mode == "mapping":
            out_real, out_imag = mask_real, mask_imag
        else:
            raise NotImplementedError
        return torch.stack([out_real.squeeze(1), out_imag.squeeze(1)], dim=-1).contiguous()  # [B,T,F,2]

    def reconstruction(self, stft_outputs, sig_len=None):                  # :443-455  in: [B,T,F,2]
        c = torch.view_as_complex(stft_outputs.contiguous()).transpose(1, 2)  # -> [B,F,T]
        return torch.istft(c, n_fft=self.fft_len, hop_length=self.win_inc, win_length=self.win_len,
                           window=self.window, center=True, length=sig_len)


class ConvSTFT:
    """train_base/acoustics/conv_stft.py:8-129, the conv-form analysis / synthesis pair (320/160, symmetric hamming,
    zero padding win-hop).  Restated with FFT calls instead of DFT-matrix convolutions (same linear maps):
    ``stft``  :72-98   = torch.stft(window=hamming(N, symmetric), center-equivalent constant pad of win-hop);
    ``istft`` :100-129 = per-frame inverse real DFT (NO synthesis window), overlap-add, divide by the hop-periodic
    sum of the ANALYSIS window (+eps, :60-70,127-128), trim win-hop at both ends (conv_transpose1d padding).
    Repairs: scipy.hamming / nn.parameter (:20,23), imaginary channel (:102), Hermitian cat (:108), the sign of the
    imaginary synthesis term (:120) and the envelope tiling (:66-67) -- see oracle/ref_extract.py::conv_stft_class."""

    def __init__(self, win_size=320, hop_size=160):
        import scipy.signal
        self.win_size, self.hop_size = win_size, hop_size
        self.n_overlap = win_size // hop_size
        self.win = torch.relu(torch.from_numpy(scipy.signal.windows.hamming(win_size)).float())   # :20-21
        self.eps = torch.finfo(torch.float32).eps                                                # :39

    def stft(self, sig):
        pad = self.win_size - self.hop_size                                                      # :85,90
        x = torch.nn.functional.pad(sig, (pad, pad))
        c = torch.stft(x, self.win_size, self.hop_size, self.win_size, window=self.win, center=False,
                       return_complex=True)                                                      # [B,F,T]
        spec_r, spec_i = c.real.transpose(-1, -2).contiguous(), c.imag.transpose(-1, -2).contiguous()   # :92-93
        return spec_r, spec_i, torch.sqrt(spec_r ** 2 + spec_i ** 2), torch.atan2(spec_i, spec_r)       # :95-96

    def istft(self, x):
        """x: [B,2,T,F] (real, imag) -> [B, (T-1)*hop + win - 2*(win-hop)]."""
        c = torch.complex(x[:, 0], x[:, 1])                                                      # [B,T,F]
        frames = torch.fft.irfft(c, n=self.win_size, dim=-1)                                     # :105-124 (basis / N)
        B, T, N = frames.shape
        hop, pad = self.hop_size, self.win_size - self.hop_size
        out = torch.zeros(B, (T - 1) * hop + N)
        for t in range(T):                                                                       # conv_transpose1d = OLA
            out[:, t * hop:t * hop + N] += frames[:, t]
        sig = out[:, pad:out.shape[1] - pad]
        seg = sum(self.win[i * hop:(i + 1) * hop] for i in range(self.n_overlap))                # :62-65
        window = seg.repeat(T - self.n_overlap + 1)                                              # :66-68 repaired
        return sig / (window + self.eps)                                                         # :128


def complex_mul(noisy_r, noisy_i, mask_r, mask_i):
    """train_base/acoustics/mask.py:60-62."""
    return noisy_r * mask_r - noisy_i * mask_i, noisy_r * mask_i + noisy_i * mask_r


# --------------------------------------------------------------------------
# loss_func/loss.py
# --------------------------------------------------------------------------
def wo_male(ref, est, unproc, norm=False, eps=1e-8):
    """loss_func/loss.py:121-148; repairs :129 torch.size -> .size(), :139 index (App. A.4)."""
    if ref.shape != est.shape:
        raise RuntimeError(f"Dimension mismatch when calculate wo-male, {ref.shape} vs {est.shape}")
    alpha, beta, gamma = 2, 1, 1
    B, C, T, F = ref.size()
    mag_ref = torch.sqrt(ref[:, 0] ** 2 + ref[:, 1] ** 2)
    mag_est = torch.sqrt(est[:, 0] ** 2 + est[:, 1] ** 2)
    mag_unproc = torch.sqrt(unproc[:, 0] ** 2 + unproc[:, 1] ** 2)
    iam = (mag_ref / mag_unproc) ** gamma
    w_iam = torch.exp(alpha / (beta + iam))
    loss = w_iam * torch.abs(torch.log10(mag_est + 1) - torch.log10(mag_ref + 1))
    return torch.sum(loss) / (B * T * F * 1.0)


def rmse(ref, est, eps=1e-8):
    """loss_func/loss.py:59-78 (sum sqrt(err^2) / (B*T*F))."""
    if ref.shape != est.shape:
        raise RuntimeError(f"Dimension mismatch when calculate rmse, {ref.shape} vs {est.shape}")
    B, C, T, F = ref.size()
    return torch.sum(torch.sqrt((est - ref) ** 2)) / (B * T * F)


def c_rmse(ref, est, unproc=None, norm=False, eps=1e-8):
    """loss_func/loss.py:88-118, arithmetic kept literally (incl. the tmp1/tmp2 mix at :107-109)."""
    if ref.shape != est.shape:
        raise RuntimeError(f"Dimension mismatch when calculate c_mse, {ref.shape} vs {est.shape}")
    c, beta = 0.3, 0.3
    mag_ref = torch.sqrt(ref[:, 0] ** 2 + ref[:, 1] ** 2)
    phase_ref = torch.atan2(ref[:, 1], ref[:, 0])
    mag_est = torch.sqrt(est[:, 0] ** 2 + est[:, 1] ** 2)
    phase_est = torch.atan2(est[:, 1], est[:, 0])
    tmp1, tmp2 = torch.pow(mag_est, c), torch.pow(mag_ref, c)
    tmp3 = tmp1 * torch.cos(phase_ref) + tmp1 * torch.sin(phase_ref) * 1j
    tmp4 = tmp2 * torch.cos(phase_est) + tmp1 * torch.sin(phase_est) * 1j
    tmp5 = torch.abs(tmp3 - tmp4)
    loss1 = (tmp2 - tmp1) ** 2
    return (1 - beta) * torch.sum(loss1) + beta * torch.sum(tmp5 ** 2)


def sisnr(s1, s2, eps=1e-8):
    """loss_func/loss.py:37-56."""
    def l2(a, b):
        return torch.sum(a * b, -1, keepdim=True)
    s_target = l2(s1, s2) / (l2(s2, s2) + eps) * s2
    e_noise = s1 - s_target
    snr = 10 * torch.log10(l2(s_target, s_target) / (l2(e_noise, e_noise) + eps) + eps)
    return torch.mean(snr)


def si_snr_loss():
    """train_base/loss.py:7-25."""
    def si_snr(x, s, eps=1e-8):
        def l2norm(mat, keep_dim=False):
            return torch.norm(mat, dim=-1, keepdim=keep_dim)
        if x.shape != s.shape:
            raise RuntimeError(f"Dimension mismatch when calculate si_snr, {x.shape} vs {s.shape}")
        x_zm = x - torch.mean(x, dim=-1, keepdim=True)
        s_zm = s - torch.mean(s, dim=-1, keepdim=True)
        t = torch.sum(x_zm * s_zm, dim=-1, keepdim=True) * s_zm / (l2norm(s_zm, keep_dim=True) ** 2 + eps)
        return -torch.mean(20 * torch.log10(eps + l2norm(t) / (l2norm(x_zm - t) + eps)))
    return si_snr


class loss_func:
    """loss_func/loss.py:16-34 dispatcher (arg order: wo_male(labels, inputs, noisy))."""

    MODES = ['SI-SNR', 'SS-SNR', 'MSE', 'Normal_MSE', 'CN_MSE', 'D_MSE', 'WO_MALE', 'C_MSE']

    def __init__(self, loss_mode):
        assert loss_mode in self.MODES, "Loss mode must be one of ***"
        self.loss_mode = loss_mode

    def loss(self, inputs, labels, noisy=None):
        if self.loss_mode == 'SI-SNR':
            return -(sisnr(inputs, labels))
        elif self.loss_mode == 'SS-SNR':
            return 0
        elif self.loss_mode == 'WO_MALE':
            return wo_male(labels, inputs, noisy)
        elif self.loss_mode == 'C_MSE':
            return c_rmse(labels, inputs)
        elif self.loss_mode == 'MSE':
            return rmse(labels, inputs)


# --------------------------------------------------------------------------
# The step in front of the path: input feature norms and on-the-fly mixing
# --------------------------------------------------------------------------
EPSILON = float(torch.finfo(torch.float32).eps)            # train_base/constant.py:8


def offline_laplace_norm(x):
    """train_base/model/base_model.py:202-215; x [B,C,F,T]."""
    return x / (torch.mean(x, dim=(1, 2, 3), keepdim=True) + 1e-5)


def offline_gaussian_norm(x):
    """base_model.py:247-261."""
    mu = torch.mean(x, dim=(1, 2, 3), keepdim=True)
    return (x - mu) / (torch.std(x, dim=(1, 2, 3), keepdim=True) + 1e-5)


def cumulative_laplace_norm(x):
    """base_model.py:217-245: x [B,C,F,T] divided by the running mean over all bins of the frames 0..t."""
    B, C, F, T = x.size()
    v = x.reshape(B * C, F, T)
    cum = torch.cumsum(torch.sum(v, dim=1), dim=-1)
    count = torch.arange(F, F * T + 1, F, dtype=x.dtype).reshape(1, T)
    mean = (cum / count).reshape(B * C, 1, T)
    return (v / (mean + EPSILON)).reshape(B, C, F, T)


def cumulative_layer_norm(x):
    """base_model.py:263-300 (online zero-mean / unit-variance; the variance formula is kept literally)."""
    B, C, F, T = x.size()
    v = x.reshape(B * C, F, T)
    cs = torch.cumsum(torch.sum(v, dim=1), dim=-1)
    cp = torch.cumsum(torch.sum(torch.square(v), dim=1), dim=-1)
    count = torch.arange(F, F * T + 1, F, dtype=x.dtype).reshape(1, T)
    mean = cs / count
    var = (cp - 2 * mean * cs) / count + mean.pow(2)
    std = torch.sqrt(var + EPSILON)
    return ((v - mean.reshape(B * C, 1, T)) / std.reshape(B * C, 1, T)).reshape(B, C, F, T)


def snr_mix(clean_y, noise_y, snr, target_dB_FS=None, rir=None, rir_noise=None, eps=1e-7):
    """dataset/dataset.py:236-264 on torch tensors [L] (the reference: numpy + scipy.signal.fftconvolve).  The reference file
    ENDS inside this function (after drawing the output level, :262-264, nothing is returned); kept here: reverberation, peak
    normalisation, the SNR scalar and the sum (:244-260); ``target_dB_FS`` (the level the reference draws at random) then scales
    noisy and clean by 10^(dB/20) / (rms(noisy) + eps), the continuation of the recipe the file was taken from."""
    def fftconvolve(a, b):
        n = a.numel() + b.numel() - 1
        return torch.fft.irfft(torch.fft.rfft(a.double(), n) * torch.fft.rfft(b.double(), n), n).to(a.dtype)
    if rir is not None:
        clean_y = fftconvolve(clean_y, rir)[: clean_y.numel()]
    if rir_noise is not None:
        noise_y = fftconvolve(noise_y, rir_noise)[: noise_y.numel()]
    clean_y = clean_y / (clean_y.abs().max() + eps)
    clean_rms = (clean_y ** 2).mean() ** 0.5
    noise_y = noise_y / (noise_y.abs().max() + eps)
    noise_rms = (noise_y ** 2).mean() ** 0.5
    noise_y = noise_y * (clean_rms / (10 ** (snr / 20)) / (noise_rms + eps))
    noisy_y = clean_y + noise_y
    if target_dB_FS is not None:
        sc = 10 ** (target_dB_FS / 20) / ((noisy_y ** 2).mean() ** 0.5 + eps)
        noisy_y, clean_y = noisy_y * sc, clean_y * sc
    return noisy_y, clean_y


# --------------------------------------------------------------------------
# The hot path end to end, used by parity tests and cpu_baseline
# --------------------------------------------------------------------------
def spec_to_bctf(c):
    """complex [B,F,T] -> real [B,2,T,F] (utils/utils.py:397-399 layout)."""
    return torch.view_as_real(c).permute(0, 3, 2, 1).contiguous()


def enhance(model, noisy, n_fft=512, hop=320, pad_mode="reflect"):
    """noisy wav [B,L] -> (enhanced wav [B,L], est spec [B,2,T,NF], mask [B,1,T,F], noisy spec [B,2,T,NF])."""
    X = spec_to_bctf(stft(noisy, n_fft, hop, n_fft, pad_mode))            # feature.py:10-30
    F = model.in_feat
    mag = torch.sqrt(X[:, 0:1] ** 2 + X[:, 1:2] ** 2 + EPS_MAG)[..., :F]   # utils.py:400
    mask = model(mag)                                                     # cruse_net.py:147-165
    full = torch.ones_like(X[:, 0:1])
    full[..., :F] = mask                                                  # bins >= F pass through
    est = X * full                                                        # utils.py:418-420 mag_mapping
    c = torch.complex(est[:, 0], est[:, 1]).transpose(1, 2)               # [B,NF,T]
    wav = istft(c, n_fft, hop, n_fft, length=noisy.shape[-1])             # feature.py:33-61
    return wav, est, mask, X


def forward_loss(model, noisy, clean, n_fft=512, hop=320, pad_mode="reflect"):
    """STFT + forward + mask + iSTFT + wo_male on the F bins the net sees."""
    wav, est, mask, X = enhance(model, noisy, n_fft, hop, pad_mode)
    S = spec_to_bctf(stft(clean, n_fft, hop, n_fft, pad_mode))
    F = model.in_feat
    loss = wo_male(S[..., :F], est[..., :F], X[..., :F])                  # loss.py:121-148
    return loss, wav, est, mask


def synth_batch(B, L, seed=20260):
    """SURVEY section 8d synthetic data: clean/noise = 0.05*randn, noisy = clean+noise."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    clean = 0.05 * torch.randn(B, L, generator=g)
    noise = 0.05 * torch.randn(B, L, generator=g)
    return clean + noise, clean


def make_model(in_feat=256, act="relu", seed=1234, eval_stats=True):
    """SURVEY section 8d weights: manual_seed(1234) + default inits; eval BN stats randomised."""
    torch.manual_seed(seed)
    m = unet_2(in_feat=in_feat, act=act)
    if eval_stats:
        g = torch.Generator(device="cpu").manual_seed(seed + 1)
        for mod in m.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.running_mean.copy_(0.1 * torch.randn(mod.num_features, generator=g))
                mod.running_var.copy_(1 + 0.1 * torch.rand(mod.num_features, generator=g))
    return m


def seeded_fill_(mod, seed, scale=0.06):
    """Overwrite every floating parameter / buffer of ``mod`` from a seeded generator in state_dict key order, so that a
    test can rebuild the same weights in the oracle's module without storing 3 M floats (oracle/make_golden.py fills the reference's modules with it, tests/test_oracle.py the oracle's)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for k, v in mod.state_dict().items():
            if not v.is_floating_point():
                continue
            r = torch.randn(v.shape, generator=g) * scale
            if k.endswith("running_var"):
                r = 1 + r.abs()
            elif k.endswith(".weight") and v.dim() == 1:       # BN / LN scale
                r = 1 + r
            v.copy_(r)

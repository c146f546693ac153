"""ctypes binding of libcruse_sm100.so (C ABI in include/cruse_b200.h).

The library is the only compute path.  ``lib()`` raises ``RuntimeError`` when the shared
object is missing -- there is deliberately no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcruse_sm100.so")

c_fp = C.c_void_p      # device float*
c_pp = C.c_void_p      # host array of device pointers
c_int = C.c_int
c_ll = C.c_longlong
c_f = C.c_float
c_d = C.c_double


class CplxLayout(C.Structure):
    """cruse_cplx_layout: re = p[b*sb + t*st + f*sf], im = p[... + im_off]."""
    _fields_ = [("sb", c_ll), ("st", c_ll), ("sf", c_ll), ("im_off", c_ll)]


# name -> (restype, argtypes).  Must list every symbol include/cruse_b200.h declares
# (tests/test_abi.py parses the header and checks).
SIGNATURES = {
    "cruse_version": (c_int, []),
    "cruse_last_error": (C.c_char_p, []),
    "cruse_sm_count": (c_int, []),
    "cruse_stft_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_f, c_fp]),
    "cruse_mask_istft_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_fp]),
    "cruse_mask_bwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_fp]),
    "cruse_conv_fwd": (c_int, [c_fp] * 7 + [c_int, c_fp, c_fp] + [c_int] * 8 + [c_fp]),
    "cruse_conv_nparts": (c_int, [c_int, c_int]),
    "cruse_conv_fwd_tm": (c_int, [c_fp] * 6 + [c_int, c_fp] + [c_int] * 10 + [c_fp]),
    "cruse_conv_get_mode": (c_int, []),
    "cruse_conv_set_mode": (c_int, [c_int]),
    "cruse_conv_set_max_ctas": (c_int, [c_int]),
    "cruse_conv_fwd_range": (c_int, [c_fp] * 6 + [c_int, c_fp] + [c_int] * 12 + [c_fp]),
    "cruse_conv_skip_fwd": (c_int, [c_fp] * 6 + [c_int, c_fp, c_fp, c_fp] + [c_int] * 9 + [c_fp]),
    "cruse_convT_fwd_range": (c_int, [c_fp] * 6 + [c_int, c_fp, c_fp] + [c_int] * 8 + [c_fp]),
    "cruse_layernorm_fwd_range": (c_int, [c_fp, c_fp, c_fp, c_f, c_fp, c_fp] + [c_int] * 5 + [c_fp]),
    "cruse_decoder_fused_image_floats": (C.c_longlong, [c_int]),
    "cruse_decoder_fused_prep": (c_int, [c_pp] * 5 + [c_int, c_fp, c_fp, c_fp, c_fp]),
    "cruse_decoder_fused_range": (c_int, [c_fp, c_fp, c_fp, c_f, c_pp, c_int, c_fp, c_fp, c_fp, CplxLayout, c_fp, CplxLayout, c_fp] + [c_int] * 5 + [c_fp]),
    "cruse_wo_male_finish_rows": (c_int, [c_fp, c_int, c_int, c_int, c_fp, c_fp]),
    "cruse_convT_fwd": (c_int, [c_fp] * 6 + [c_int, c_fp, c_fp, c_fp] + [c_int] * 6 + [c_fp]),
    "cruse_bn_finalize": (c_int, [c_fp, c_int, c_int, c_d, c_fp, c_fp, c_f, c_f] + [c_fp] * 6 + [c_fp]),
    "cruse_bn_fold": (c_int, [c_fp, c_fp, c_fp, c_fp, c_f, c_fp, c_fp, c_int, c_fp]),
    "cruse_bn_fold_many": (c_int, [c_pp] * 4 + [c_fp, c_pp, c_pp, c_fp, c_int, c_fp]),
    "cruse_bn_act_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_fp, c_ll, c_int, c_int, c_fp]),
    "cruse_gru_ih_gemm": (c_int, [c_fp, c_pp, c_pp, c_pp, c_fp, c_int, c_int, c_int, c_fp]),
    "cruse_gemm_set_astat": (c_int, [c_int]),
    "cruse_gru_ih_gemm_tc": (c_int, [c_fp, c_pp, c_pp, c_pp, c_fp, c_int, c_int, c_int, c_fp]),
    "cruse_gru_seq_fwd": (c_int, [c_fp, c_pp, c_pp, c_fp, c_fp, c_fp] + [c_int] * 6 + [c_fp]),
    "cruse_gru_seq_fwd_tc": (c_int, [c_fp, c_pp, c_pp, c_fp, c_fp, c_fp, c_fp] + [c_int] * 6 + [c_fp]),
    "cruse_gru_seq_tc_max_clusters": (c_int, [c_int]),
    "cruse_gru_seq_flagged_tc": (c_int, [c_fp, c_pp, c_pp, c_fp] + [c_int] * 6 + [c_ll] * 4 + [c_fp, c_int, c_fp, C.c_uint, c_fp, c_fp, c_fp]),
    "cruse_flag_wait": (c_int, [c_fp, C.c_uint, c_fp, c_fp]),
    "cruse_flag_set": (c_int, [c_fp, C.c_uint, c_fp]),
    "cruse_poison_on_error": (c_int, [c_fp, c_pp, c_pp, c_int, c_fp]),
    "cruse_debug_seq_trace": (c_int, [c_fp]),
    "cruse_gru_ih_gemm_tm_tc": (c_int, [c_fp, c_pp, c_pp, c_pp, c_fp, c_int, c_int, c_int, c_int, c_fp]),
    "cruse_gru_seq_chunk_tc": (c_int, [c_fp, c_pp, c_pp, c_fp, c_fp, c_fp] + [c_int] * 6 + [c_ll] * 4 + [c_fp]),
    "cruse_layernorm_fwd": (c_int, [c_fp, c_fp, c_fp, c_f, c_fp, c_fp, c_fp, c_fp, c_ll, c_int, c_fp]),
    "cruse_layernorm_interleave_fwd": (c_int, [c_fp, c_fp, c_fp, c_f, c_fp, c_ll, c_int, c_int, c_fp]),
    "cruse_wo_male_fwd_bwd": (c_int, [c_fp, CplxLayout, c_fp, CplxLayout, c_fp, CplxLayout, c_fp, c_fp, c_fp,
                                      c_int, c_int, c_int, c_fp]),
    "cruse_wo_male_masked_partial_range": (c_int, [c_fp, CplxLayout, c_fp, c_fp, CplxLayout, c_fp] + [c_int] * 7 + [c_fp]),
    "cruse_wo_male_finish": (c_int, [c_fp, c_int, c_int, c_int, c_int, c_fp, c_fp]),
    "cruse_mask_istft_chunk_frames": (c_int, [c_int, c_int]),
    "cruse_mask_istft_fwd_range": (c_int, [c_fp] * 5 + [c_int] * 8 + [c_fp]),
    "cruse_istft_bwd_ws_bytes": (C.c_size_t, [c_int, c_int]),
    "cruse_istft_bwd": (c_int, [c_fp, c_fp, c_fp, c_fp] + [c_int] * 5 + [c_fp]),
    "cruse_spec_loss_fwd_bwd": (c_int, [c_int, c_fp, CplxLayout, c_fp, CplxLayout, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_fp]),
    "cruse_sisnr_ws_bytes": (C.c_size_t, [c_int]),
    "cruse_sisnr_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_f, c_fp]),
    "cruse_sisnr_bwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_fp]),
    "cruse_si_snr_zm_ws_bytes": (C.c_size_t, [c_int]),
    "cruse_si_snr_zm_fwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_f, c_fp]),
    "cruse_si_snr_zm_bwd": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_fp]),
    "cruse_wo_male_masked_fwd": (c_int, [c_fp, CplxLayout, c_fp, c_fp, CplxLayout, c_fp, c_fp, c_int, c_int, c_int, c_fp]),
    "cruse_wo_male_ws_bytes": (C.c_size_t, []),
    # ---- a9 backward
    "cruse_colsum": (c_int, [c_fp, c_int, c_int, c_fp, c_int, c_fp]),
    "cruse_conv_dgrad": (c_int, [c_fp, c_fp, c_fp, c_fp] + [c_int] * 8 + [c_fp]),
    "cruse_conv_wgrad": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp] + [c_int] * 8 + [c_fp]),
    "cruse_conv_wgrad_ws_bytes": (C.c_size_t, [c_int] * 7),
    "cruse_convT_dgrad": (c_int, [c_fp, c_fp, c_fp, c_fp] + [c_int] * 6 + [c_fp]),
    "cruse_convT_wgrad": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp] + [c_int] * 6 + [c_fp]),
    "cruse_convT_wgrad_ws_bytes": (C.c_size_t, [c_int] * 6),
    "cruse_bn_bwd_nparts": (c_int, [c_ll]),
    "cruse_bn_act_bwd_reduce": (c_int, [c_fp] * 5 + [c_int] + [c_fp] * 3 + [c_ll, c_int, c_int, c_fp]),
    "cruse_bn_bwd_finalize": (c_int, [c_fp, c_int, c_int, c_d, c_fp, c_fp, c_int, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "cruse_bn_act_bwd_apply": (c_int, [c_fp] * 5 + [c_int] + [c_fp] * 4 + [c_ll, c_int, c_int, c_fp]),
    "cruse_sigmoid_bwd": (c_int, [c_fp, c_fp, c_fp, c_ll, c_fp]),
    "cruse_layernorm_bwd_nparts": (c_int, [c_ll]),
    "cruse_layernorm_bwd": (c_int, [c_fp] * 7 + [c_ll, c_int, c_fp]),
    "cruse_gru_seq_bwd_tc": (c_int, [c_fp, c_fp, c_fp, c_fp, c_pp, c_fp, c_fp, c_fp, c_fp] + [c_int] * 6 + [c_fp]),
    "cruse_gemm_tn_tc": (c_int, [c_pp, c_pp, c_pp, c_pp, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_int, c_ll, c_fp]),
    "cruse_gemm_tc": (c_int, [c_pp, c_pp, c_pp, c_pp, c_pp, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_int, c_ll, c_int, c_int, c_int, c_fp]),
    "cruse_gru_exact_ws_bytes": (C.c_size_t, [c_int] * 2),
    "cruse_gru_seq_fwd_exact": (c_int, [c_fp, c_pp, c_pp, c_fp, c_fp, c_fp, c_fp, c_fp] + [c_int] * 6 + [c_fp]),
    "cruse_gru_seq_bwd_exact": (c_int, [c_fp, c_fp, c_fp, c_fp, c_pp, c_fp, c_fp, c_fp, c_fp] + [c_int] * 6 + [c_fp]),
    "cruse_gemm_tn_fp32": (c_int, [c_pp, c_pp, c_pp, c_pp, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_int, c_ll, c_fp]),
    "cruse_gru_step_ws_bytes": (C.c_size_t, [c_int] * 3),
    "cruse_gru_step": (c_int, [c_fp, c_pp, c_pp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_fp]),
    "cruse_feature_norm": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_fp]),
    "cruse_rir_conv": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_ll, c_fp]),
    "cruse_snr_mix_ws_bytes": (C.c_size_t, [c_int]),
    "cruse_snr_mix": (c_int, [c_fp] * 7 + [c_int, c_int, c_f, c_fp]),
    "cruse_pcm16_to_float": (c_int, [c_fp, c_fp, C.c_longlong, c_fp]),
    "cruse_nccl_available": (c_int, []),
    "cruse_nccl_unique_id": (c_int, [C.c_void_p]),
    "cruse_nccl_comm_init": (c_int, [C.POINTER(C.c_void_p), c_int, c_int, C.c_void_p]),
    "cruse_flat_allreduce": (c_int, [C.c_void_p, c_fp, c_ll, c_f, c_fp]),
    "cruse_nccl_comm_destroy": (c_int, [C.c_void_p]),
    "cruse_transpose_gcm": (c_int, [c_fp, c_fp, c_fp, c_ll, c_int, c_int, c_ll, c_ll, c_ll, c_int, c_int, c_ll, c_fp]),
}

_lock = threading.Lock()
_lib = None


def lib():
    """Load (once) and return the shared library; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: cruse_b200 has no CPU/eager fallback. "
                    "Build it with `python -m cruse_b200.build` (needs nvcc, targets sm_100a).")
            h = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(h, name)   # AttributeError if the .so is stale
                fn.restype = res
                fn.argtypes = args
            _lib = h
    return _lib


def check(rc: int, what: str = ""):
    """Reference convention = Python exceptions (loss.py:65-68, feature.py:352-354): rc != 0 -> RuntimeError."""
    if rc != 0:
        msg = lib().cruse_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what or 'cruse_b200'} failed (rc={rc}): {msg}")

"""CPU: the C-ABI library loads, exports every symbol include/cruse_b200.h declares, the ctypes table
covers them all, and the product path neither imports the oracle nor falls back to CPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cruse_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cruse_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_declares_something():
    syms = header_symbols()
    assert "cruse_stft_fwd" in syms and "cruse_gru_seq_fwd" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol():
    from cruse_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run `python -m cruse_b200.build` (or __graft_entry__.build())"
    h = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(h, s), f"{s} declared in include/cruse_b200.h but not exported"


def test_ctypes_table_matches_header():
    from cruse_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    lib = _lib.lib()
    assert lib.cruse_version() >= 100
    assert lib.cruse_wo_male_ws_bytes() > 0
    assert lib.cruse_conv_nparts(2, 17) > 0


def test_argument_errors_come_back_as_runtime_error():
    from cruse_b200 import _lib
    lib = _lib.lib()
    rc = lib.cruse_stft_fwd(None, None, None, None, 1, 100, 512, 320, 1, 0, 0, 0.0, None)
    assert rc != 0
    with pytest.raises(RuntimeError, match="null pointer"):
        _lib.check(rc, "stft_fwd")


def test_no_cpu_fallback():
    from cruse_b200.cruse_net import unet_2
    from cruse_b200 import acoustics, loss
    m = unet_2(in_feat=256).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU"):
        m(torch.randn(1, 1, 4, 256))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        acoustics.stft(torch.randn(1, 4000), 512, 320, 512)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        loss.wo_male(torch.randn(1, 2, 3, 4), torch.randn(1, 2, 3, 4), torch.randn(1, 2, 3, 4))


def test_missing_library_fails_loudly(tmp_path):
    code = ("import cruse_b200._lib as l, sys\n"
            f"l.LIB_PATH = {str(tmp_path / 'nope.so')!r}\n"
            "try:\n    l.lib()\nexcept RuntimeError as e:\n    print('RAISED', e)\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True).stdout
    assert "RAISED" in out and "no CPU/eager fallback" in out


def test_product_never_imports_oracle():
    code = ("import sys, cruse_b200, cruse_b200.cruse_net, cruse_b200.acoustics, cruse_b200.loss, cruse_b200.pipeline, "
            "cruse_b200.autograd\n"
            "print('ORACLE' if any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules) else 'CLEAN')\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True).stdout
    assert "CLEAN" in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cruse_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "cruse_oracle" not in txt and "/oracle/" not in txt, f


def test_state_dict_surface_equals_oracle():
    from cruse_b200.cruse_net import unet_2
    from oracle import cruse_oracle as o
    for F, act in ((256, "relu"), (161, "prelu")):
        torch.manual_seed(1234)
        ours = unet_2(in_feat=F, act=act)
        ref = o.make_model(F, act=act, eval_stats=False)
        a, b = ours.state_dict(), ref.state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
        assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
        ours.load_state_dict(ref.state_dict(), strict=True)


def test_ctypes_argument_counts_match_header_prototypes():
    """every entry of the ctypes table passes exactly as many arguments as the C prototype declares (a drifted table corrupts
    the call silently): parsed from include/cruse_b200.h"""
    import re
    from cruse_b200 import _lib
    text = open(os.path.join(ROOT, "include", "cruse_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = dict(re.findall(r"\b(cruse_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    assert set(protos) == set(_lib.SIGNATURES)
    for name, params in protos.items():
        params = " ".join(params.split())
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))


def test_collective_entry_points_resolve_nccl_at_run_time():
    """cruse_flat_allreduce & co. (include/cruse_b200.h, a10) look NCCL up in the process's libnccl.so.2 when first used: the
    library loads without it, and with it rank 0 can draw the 128-byte unique id (no device needed for that)."""
    import ctypes as C
    from cruse_b200._lib import lib
    have = lib().cruse_nccl_available()
    assert have in (0, 1)
    raw = (C.c_char * 128)()
    rc = lib().cruse_nccl_unique_id(raw)
    if have:
        assert rc == 0 and any(bytes(raw))
    else:
        assert rc != 0 and b"libnccl" in lib().cruse_last_error()

"""C-ABI surface checks that need no GPU: the library loads, exports every symbol include/svi_ls.h
declares, validates arguments, and FAILS LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from svinet_b200 import build as svbuild
from svinet_b200 import engine

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    svbuild.build_lib()
    return engine.load_library()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(REPO, "include", "svi_ls.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(svi_ls_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(engine.ABI_SYMBOLS), declared ^ set(engine.ABI_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_fa2_header_symbols_all_exported(lib):
    from svinet_b200 import fa2_engine
    hdr = open(os.path.join(REPO, "include", "svi_fa2.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(svi_fa2_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(fa2_engine.FA2_SYMBOLS), declared ^ set(fa2_engine.FA2_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    fa2_engine.bind(lib)
    cfg = fa2_engine.Fa2Config()
    lib.svi_fa2_default_config(C.byref(cfg), 100, 8)
    assert (cfg.n, cfg.k, cfg.alpha, cfg.tau0, cfg.m_sets, cfg.online_iterations) == (100, 8, 0.125, 1025.0, 10, 50)
    h = C.c_void_p()
    assert lib.svi_fa2_create(None, C.byref(h)) == -1
    cfg.k = 70000
    assert lib.svi_fa2_create(C.byref(cfg), C.byref(h)) == -4
    assert lib.svi_fa2_step(None, 0, 0, 0, 0, None) == -1


def test_abi_is_usable_from_plain_c(lib, tmp_path):
    """include/*.h are C headers and libsvi_ls.so links from C: tests/cc/abi_c_check.c with gcc -std=c99."""
    import subprocess
    exe = str(tmp_path / "abi_c_check")
    libdir = os.path.dirname(svbuild.LIB)
    subprocess.check_call([os.environ.get("CC", "gcc"), "-std=c99", "-Wall", "-Werror", "-pedantic", "-I",
                           os.path.join(REPO, "include"), "-o", exe, os.path.join(REPO, "tests", "cc", "abi_c_check.c"),
                           "-L", libdir, "-lsvi_ls", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "svi_ls_create rc=" in out.stdout


def test_abi_version(lib):
    assert lib.svi_ls_abi_version() == 2


def test_invalid_arguments_are_rejected(lib):
    h = C.c_void_p()
    assert lib.svi_ls_create(None, None, None, C.byref(h)) == -1
    assert b"null" in lib.svi_ls_last_error()
    cfg = engine.SviConfig(n=0, k=4, nlinks=0, alpha=0.25, eta0=1, eta1=1, ones=0, device=-1, seg_len=0,
                           node_begin=0, node_end=0)
    assert lib.svi_ls_create(C.byref(cfg), None, None, C.byref(h)) == -1
    cfg = engine.SviConfig(n=10, k=4, nlinks=0, alpha=0.25, eta0=1, eta1=1, ones=0, device=-1, seg_len=0,
                           node_begin=5, node_end=11)
    assert lib.svi_ls_create(C.byref(cfg), None, None, C.byref(h)) == -1
    cfg = engine.SviConfig(n=10, k=70000, nlinks=0, alpha=0.25, eta0=1, eta1=1, ones=0, device=-1, seg_len=0,
                           node_begin=0, node_end=10)
    assert lib.svi_ls_create(C.byref(cfg), None, None, C.byref(h)) == -4
    assert lib.svi_ls_step(None, 0, 1, 0) == -1
    assert lib.svi_ls_get_state(None, None, None) == -1


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_error_not_fallback(lib):
    links = np.array([[0, 1], [1, 2]], dtype=np.uint32)
    with pytest.raises(engine.SviError) as ei:
        engine.LinkSamplingEngine(3, 4, links)
    assert "svi_ls error -2" in str(ei.value)

"""The digamma restatements (oracle/oracle_ls.c, the GSL stand-in of the reference build, the device's
digamma_pos) are one published formula written three times; scipy.special.digamma (Cephes) is the independent
implementation they are pinned to here.  gsl_sf_psi call sites: src/linksampling.hh:181,184,198,200."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy.special import digamma

import oracle_py as orc

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grid():
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-8, 8, 20000), rng.uniform(0.5, 12.0, 20000),
                        np.array([1e-300, 1e-12, 0.005, 1.0, 1.4616321449683623, 2.0, 9.999999, 10.0, 10.000001, 1e15])])
    return x


def _err(got, want):
    # psi has a zero at 1.46163...: measure against max(|psi|, 1)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)))


def test_oracle_digamma_matches_scipy():
    L = orc.lib()
    x = _grid()
    got = np.array([L.orc_digamma(float(v)) for v in x])
    assert _err(got, digamma(x)) < 5e-15


def test_gsl_shim_digamma_matches_scipy(tmp_path):
    """The stand-in the compiled reference (oracle/_ref) links against."""
    src = tmp_path / "psi.c"
    src.write_text('#include <gsl/gsl_sf_psi.h>\ndouble shim_psi(double x) { return gsl_sf_psi(x); }\n')
    so = str(tmp_path / "psi.so")
    subprocess.check_call(["g++", "-x", "c++", "-O2", "-shared", "-fPIC", "-I", os.path.join(REPO, "oracle", "gsl_shim"),
                           "-o", so, str(src)])
    L = C.CDLL(so)
    fn = getattr(L, "_Z8shim_psid")
    fn.restype = C.c_double
    fn.argtypes = [C.c_double]
    x = _grid()
    got = np.array([fn(float(v)) for v in x])
    assert _err(got, digamma(x)) < 5e-15


@pytest.mark.gpu
@pytest.mark.parametrize("k", [7, 200])
def test_device_expectations_match_scipy(k):
    """k_refresh / k_lambda through the C ABI: exp(Elogpi - rowmax) and the lambda side, from gamma spanning
    1e-3 .. 1e6, against scipy."""
    import torch
    from svinet_b200.engine import LinkSamplingEngine
    from svinet_b200.sharded import cuda_view
    n = 3000
    rng = np.random.default_rng(k)
    gamma = 10.0 ** rng.uniform(-3, 6, (n, k))
    lam = 10.0 ** rng.uniform(-2, 7, (k, 2))
    links = np.array([[0, 1]], dtype=np.uint32)
    eng = LinkSamplingEngine(n, k, links)
    eng.set_state(gamma, lam)
    ptr, ld = eng.device_buffer("exppi")
    b = cuda_view(ptr, (n, ld), torch.float64, torch.device("cuda", torch.cuda.current_device())).cpu().numpy()[:, :k]
    e = digamma(gamma) - digamma(gamma.sum(1, keepdims=True))
    want = np.exp(e - e.max(1, keepdims=True))
    big = want > 1e-200
    assert float(np.max(np.abs(b[big] - want[big]) / want[big])) < 1e-11     # exp amplifies |Elogpi| * eps
    eng.close()

"""ctypes binding of the `-rnode -stratified` entry points of libsvi_ls.so (include/svi_fa2.h).

Plumbing for tests and bench.py; the product boundary is the C ABI and the host-side drop-in is the C++
class FastAMM2 under svinet_b200/host/.  No CPU fallback: without the library or a CUDA device the calls
raise.
"""
import ctypes as C

import numpy as np

from .engine import SviError, load_library

FA2_SYMBOLS = [
    "svi_fa2_default_config", "svi_fa2_create", "svi_fa2_destroy", "svi_fa2_set_stream", "svi_fa2_sync",
    "svi_fa2_set_state", "svi_fa2_get_state", "svi_fa2_step", "svi_fa2_set_graph", "svi_fa2_run", "svi_fa2_draw",
    "svi_fa2_heldout", "svi_fa2_phi_pair", "svi_fa2_get_info",
]


class Fa2Config(C.Structure):
    _fields_ = [("n", C.c_uint32), ("k", C.c_uint32), ("alpha", C.c_double), ("eta0", C.c_double),
                ("eta1", C.c_double), ("epsilon", C.c_double), ("tau0", C.c_double), ("kappa", C.c_double),
                ("nodetau0", C.c_double), ("nodekappa", C.c_double), ("inf_epsilon", C.c_double),
                ("m_sets", C.c_uint32), ("online_iterations", C.c_uint32), ("meanchangethresh", C.c_double),
                ("nolambda", C.c_int32), ("device", C.c_int32), ("eager_blend", C.c_int32)]


class Fa2Info(C.Structure):
    _fields_ = [("ld", C.c_uint32), ("lanes", C.c_uint32), ("vec", C.c_uint32), ("pair_blocks", C.c_uint32),
                ("device_bytes", C.c_uint64), ("last_npairs", C.c_uint64), ("last_rounds", C.c_uint64),
                ("kernels_per_step", C.c_uint32)]


_bound = False


def bind(L):
    global _bound
    vp = C.c_void_p
    L.svi_fa2_default_config.argtypes = [C.POINTER(Fa2Config), C.c_uint32, C.c_uint32]
    L.svi_fa2_default_config.restype = None
    L.svi_fa2_create.argtypes = [C.POINTER(Fa2Config), C.POINTER(vp)]
    L.svi_fa2_destroy.argtypes = [vp]
    L.svi_fa2_destroy.restype = None
    L.svi_fa2_set_stream.argtypes = [vp, vp]
    L.svi_fa2_sync.argtypes = [vp]
    L.svi_fa2_set_state.argtypes = [vp, vp, vp, C.c_uint64]
    L.svi_fa2_get_state.argtypes = [vp, vp, vp]
    L.svi_fa2_step.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, vp]
    L.svi_fa2_set_graph.argtypes = [vp, C.c_uint64, vp, C.c_uint64, vp, vp]
    L.svi_fa2_run.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]
    L.svi_fa2_draw.argtypes = [vp, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                               C.POINTER(C.c_uint64), vp, C.c_uint64]
    L.svi_fa2_heldout.argtypes = [vp, C.c_uint64, vp, vp, vp, vp]
    L.svi_fa2_phi_pair.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp, C.POINTER(C.c_uint32)]
    L.svi_fa2_get_info.argtypes = [vp, C.POINTER(Fa2Info)]
    for name in FA2_SYMBOLS:
        if name not in ("svi_fa2_destroy", "svi_fa2_default_config"):
            getattr(L, name).restype = C.c_int
    _bound = True
    return L


def _check(L, rc):
    if rc != 0:
        raise SviError("svi_fa2 error %d: %s" % (rc, L.svi_ls_last_error().decode(errors="replace")))


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Fa2Engine:
    """One device-resident FastAMM2 problem."""

    def __init__(self, n, k, device=-1, stream=None, **overrides):
        self.L = bind(load_library())
        cfg = Fa2Config()
        self.L.svi_fa2_default_config(C.byref(cfg), n, k)
        cfg.device = device
        for key, v in overrides.items():
            if not hasattr(cfg, key):
                raise KeyError(key)
            setattr(cfg, key, v)
        self.cfg, self.n, self.k = cfg, n, k
        self.h = C.c_void_p()
        _check(self.L, self.L.svi_fa2_create(C.byref(cfg), C.byref(self.h)))
        if stream is not None:
            _check(self.L, self.L.svi_fa2_set_stream(self.h, C.c_void_p(int(stream))))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.svi_fa2_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def sync(self):
        _check(self.L, self.L.svi_fa2_sync(self.h))

    def set_state(self, gamma, lam, nodec=0):
        gamma = np.ascontiguousarray(gamma, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        assert gamma.shape == (self.n, self.k) and lam.shape == (self.k, 2)
        _check(self.L, self.L.svi_fa2_set_state(self.h, _ptr(gamma), _ptr(lam), nodec))

    def get_state(self):
        gamma = np.empty((self.n, self.k), dtype=np.float64)
        lam = np.empty((self.k, 2), dtype=np.float64)
        _check(self.L, self.L.svi_fa2_get_state(self.h, _ptr(gamma), _ptr(lam)))
        return gamma, lam

    def step(self, it, typ, start, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        _check(self.L, self.L.svi_fa2_step(self.h, it, typ, start, pairs.shape[0], _ptr(pairs)))

    def set_graph(self, links, heldout, shuffled):
        links = np.ascontiguousarray(links, dtype=np.uint32).reshape(-1, 2)
        heldout = np.ascontiguousarray(heldout, dtype=np.uint32).reshape(-1, 2)
        shuffled = np.ascontiguousarray(shuffled, dtype=np.uint32)
        assert shuffled.shape == (self.n,)
        _check(self.L, self.L.svi_fa2_set_graph(self.h, links.shape[0], _ptr(links), heldout.shape[0], _ptr(heldout),
                                                _ptr(shuffled)))

    def run(self, it0, iters, seed, count=True):
        c = C.c_uint64()
        _check(self.L, self.L.svi_fa2_run(self.h, it0, iters, seed, C.byref(c) if count else None))
        return c.value

    def draw(self, it, seed, cap=1 << 22):
        t, s, npairs = C.c_uint32(), C.c_uint32(), C.c_uint64()
        buf = np.empty((cap, 2), dtype=np.uint32)
        _check(self.L, self.L.svi_fa2_draw(self.h, it, seed, C.byref(t), C.byref(s), C.byref(npairs), _ptr(buf), cap))
        return t.value, s.value, buf[:min(cap, npairs.value)].copy()

    def heldout(self, p, q, y):
        p = np.ascontiguousarray(p, dtype=np.uint32)
        q = np.ascontiguousarray(q, dtype=np.uint32)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        out = np.empty(p.shape[0], dtype=np.float64)
        _check(self.L, self.L.svi_fa2_heldout(self.h, p.shape[0], _ptr(p), _ptr(q), _ptr(y), _ptr(out)))
        return out

    def phi_pair(self, p, q, y):
        a, b, r = np.empty(self.k), np.empty(self.k), C.c_uint32()
        _check(self.L, self.L.svi_fa2_phi_pair(self.h, p, q, int(y), _ptr(a), _ptr(b), C.byref(r)))
        return a, b, r.value

    def info(self):
        i = Fa2Info()
        _check(self.L, self.L.svi_fa2_get_info(self.h, C.byref(i)))
        return {f: getattr(i, f) for f, _ in Fa2Info._fields_}

"""ctypes binding of libsvi_ls.so (include/svi_ls.h) -- the device path of `-link-sampling`.

This is plumbing for tests and bench.py; the product boundary is the C ABI itself and the
host-side drop-in is the C++ CLI under svinet_b200/host/.  There is NO CPU fallback here: if
the shared library is missing or no CUDA device is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

SVI_BUF = {"exppi": 0, "mphi": 1, "gamma": 2, "kvec": 3, "converged": 4, "lambda": 5,
           "active": 6, "active_bits": 7, "member_bits": 8}


class SviConfig(C.Structure):
    _fields_ = [("n", C.c_uint32), ("k", C.c_uint32), ("nlinks", C.c_uint64),
                ("alpha", C.c_double), ("eta0", C.c_double), ("eta1", C.c_double),
                ("ones", C.c_uint32), ("device", C.c_int32), ("seg_len", C.c_uint32),
                ("node_begin", C.c_uint32), ("node_end", C.c_uint32)]


class SviInfo(C.Structure):
    _fields_ = [("half_edges_phi", C.c_uint64), ("half_edges_s3", C.c_uint64),
                ("segments_phi", C.c_uint64), ("segments_s3", C.c_uint64),
                ("ld", C.c_uint32), ("seg_len", C.c_uint32), ("lanes", C.c_uint32), ("vec", C.c_uint32),
                ("ring_depth", C.c_uint32), ("device_bytes", C.c_uint64), ("kernels_per_step", C.c_uint32)]


class SviError(RuntimeError):
    pass


_lib = None

# every symbol include/svi_ls.h declares (tests/test_abi.py checks the .so exports all of them)
ABI_SYMBOLS = [
    "svi_ls_create", "svi_ls_destroy", "svi_ls_set_stream", "svi_ls_sync", "svi_ls_set_state",
    "svi_ls_get_state", "svi_ls_set_converged", "svi_ls_get_converged", "svi_ls_step",
    "svi_ls_get_membership", "svi_ls_heldout", "svi_ls_get_kvectors", "svi_ls_phase_phi",
    "svi_ls_phase_node", "svi_ls_phase_s3", "svi_ls_phase_finish", "svi_ls_phase_refresh", "svi_ls_phase_lambda",
    "svi_ls_device_buffer",
    "svi_ls_peer_blob_bytes", "svi_ls_peer_export", "svi_ls_peer_attach", "svi_ls_peer_attach_local", "svi_ls_mg_step",
    "svi_ls_mg_share_gamma", "svi_ls_mg_publish_gamma", "svi_ls_mg_error", "svi_ls_mg_timing", "svi_ls_get_membership_rows",
    "svi_ls_get_info", "svi_ls_last_error", "svi_ls_abi_version",
]


def load_library(path=None):
    """dlopen libsvi_ls.so (built in-tree by svinet_b200.build); raises if it is not there."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("SVI_LS_LIB") or _build.LIB     # SVI_LS_LIB: development A/B override
    if not os.path.exists(path):
        raise SviError("%s not built: run `python -m svinet_b200.build` (needs nvcc)" % path)
    L = C.CDLL(path)
    vp, u32p, f64p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    L.svi_ls_create.argtypes = [C.POINTER(SviConfig), vp, vp, C.POINTER(vp)]
    L.svi_ls_destroy.argtypes = [vp]
    L.svi_ls_destroy.restype = None
    L.svi_ls_set_stream.argtypes = [vp, vp]
    L.svi_ls_sync.argtypes = [vp]
    L.svi_ls_set_state.argtypes = [vp, vp, vp]
    L.svi_ls_get_state.argtypes = [vp, vp, vp]
    L.svi_ls_set_converged.argtypes = [vp, vp]
    L.svi_ls_get_converged.argtypes = [vp, vp, vp]
    L.svi_ls_step.argtypes = [vp, C.c_uint32, C.c_int, C.c_int]
    L.svi_ls_get_membership.argtypes = [vp, vp]
    L.svi_ls_heldout.argtypes = [vp, C.c_uint64, vp, vp, vp, C.c_double, vp]
    L.svi_ls_get_kvectors.argtypes = [vp, vp, vp, vp, vp]
    L.svi_ls_phase_phi.argtypes = [vp, C.c_uint32, C.c_int]
    L.svi_ls_phase_node.argtypes = [vp]
    L.svi_ls_phase_s3.argtypes = [vp]
    L.svi_ls_phase_finish.argtypes = [vp, C.c_int]
    L.svi_ls_phase_refresh.argtypes = [vp, C.c_int]
    L.svi_ls_phase_lambda.argtypes = [vp, C.c_int]
    L.svi_ls_device_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.svi_ls_get_info.argtypes = [vp, C.POINTER(SviInfo)]
    L.svi_ls_peer_blob_bytes.argtypes = []
    L.svi_ls_peer_export.argtypes = [vp, vp, C.c_size_t]
    L.svi_ls_peer_attach.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32]
    L.svi_ls_peer_attach_local.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32]
    L.svi_ls_mg_step.argtypes = [vp, C.c_uint32, C.c_int, C.c_int]
    L.svi_ls_mg_share_gamma.argtypes = [vp, C.c_int]
    L.svi_ls_mg_publish_gamma.argtypes = [vp]
    L.svi_ls_mg_error.argtypes = [vp]
    L.svi_ls_mg_timing.argtypes = [vp, C.c_int, vp, vp]
    L.svi_ls_get_membership_rows.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
    L.svi_ls_last_error.restype = C.c_char_p
    L.svi_ls_abi_version.restype = C.c_int
    for name in ABI_SYMBOLS:
        if name not in ("svi_ls_destroy", "svi_ls_last_error", "svi_ls_peer_blob_bytes"):
            getattr(L, name).restype = C.c_int
    L.svi_ls_peer_blob_bytes.restype = C.c_size_t
    if path == _build.LIB:
        _lib = L
    return L


def _check(L, rc):
    if rc != 0:
        raise SviError("svi_ls error %d: %s" % (rc, L.svi_ls_last_error().decode(errors="replace")))


def _ptr(a):
    return a.ctypes.data if a is not None else None


class LinkSamplingEngine:
    """One device-resident inference problem (one shard of it when node_range is given)."""

    def __init__(self, n, k, links, tl=None, alpha=None, eta0=1.0, eta1=1.0, ones=None, device=-1,
                 seg_len=0, node_range=None, stream=None):
        self.L = load_library()
        links = np.ascontiguousarray(links, dtype=np.uint32).reshape(-1, 2)
        if tl is not None:
            tl = np.ascontiguousarray(tl, dtype=np.float64)
            assert tl.shape == (n,)
        nb, ne = node_range if node_range is not None else (0, n)
        cfg = SviConfig(n=n, k=k, nlinks=links.shape[0], alpha=(1.0 / k if alpha is None else alpha),
                        eta0=eta0, eta1=eta1, ones=(links.shape[0] if ones is None else ones),
                        device=device, seg_len=seg_len, node_begin=nb, node_end=ne)
        self.n, self.k, self.nlinks = n, k, links.shape[0]
        self.words = (k + 31) // 32
        self.h = C.c_void_p()
        _check(self.L, self.L.svi_ls_create(C.byref(cfg), _ptr(links), _ptr(tl), C.byref(self.h)))
        if stream is not None:
            self.set_stream(stream)

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.svi_ls_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing -------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""
        _check(self.L, self.L.svi_ls_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def sync(self):
        _check(self.L, self.L.svi_ls_sync(self.h))

    def info(self):
        i = SviInfo()
        _check(self.L, self.L.svi_ls_get_info(self.h, C.byref(i)))
        return {f: getattr(i, f) for f, _ in SviInfo._fields_}

    def device_buffer(self, name):
        p, ld = C.c_void_p(), C.c_uint64()
        _check(self.L, self.L.svi_ls_device_buffer(self.h, SVI_BUF[name], C.byref(p), C.byref(ld)))
        return p.value, ld.value

    # -- state ----------------------------------------------------------------------------
    def set_state(self, gamma, lam):
        gamma = np.ascontiguousarray(gamma, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        assert gamma.shape == (self.n, self.k) and lam.shape == (self.k, 2)
        _check(self.L, self.L.svi_ls_set_state(self.h, _ptr(gamma), _ptr(lam)))

    def set_state_ptr(self, gamma_ptr, lambda_ptr):
        """Raw host pointers (pinned buffers owned by the caller)."""
        _check(self.L, self.L.svi_ls_set_state(self.h, C.c_void_p(gamma_ptr), C.c_void_p(lambda_ptr)))

    def get_state_ptr(self, gamma_ptr, lambda_ptr):
        _check(self.L, self.L.svi_ls_get_state(self.h, C.c_void_p(gamma_ptr), C.c_void_p(lambda_ptr)))

    def get_state(self):
        gamma = np.empty((self.n, self.k), dtype=np.float64)
        lam = np.empty((self.k, 2), dtype=np.float64)
        _check(self.L, self.L.svi_ls_get_state(self.h, _ptr(gamma), _ptr(lam)))
        return gamma, lam

    def set_converged(self, conv):
        conv = np.ascontiguousarray(conv, dtype=np.uint32)
        assert conv.shape == (self.n,)
        _check(self.L, self.L.svi_ls_set_converged(self.h, _ptr(conv)))

    def get_converged(self):
        conv = np.empty(self.n, dtype=np.uint32)
        act = np.empty(self.n, dtype=np.uint32)
        _check(self.L, self.L.svi_ls_get_converged(self.h, _ptr(conv), _ptr(act)))
        return conv, act

    # -- the path -------------------------------------------------------------------------
    def step(self, it, annealing, write_comm):
        _check(self.L, self.L.svi_ls_step(self.h, it, int(annealing), int(write_comm)))

    def phase_phi(self, it, write_comm):
        _check(self.L, self.L.svi_ls_phase_phi(self.h, it, int(write_comm)))

    def phase_node(self):
        _check(self.L, self.L.svi_ls_phase_node(self.h))

    def phase_s3(self):
        _check(self.L, self.L.svi_ls_phase_s3(self.h))

    def phase_finish(self, annealing):
        _check(self.L, self.L.svi_ls_phase_finish(self.h, int(annealing)))

    def phase_refresh(self, annealing):
        _check(self.L, self.L.svi_ls_phase_refresh(self.h, int(annealing)))

    def phase_lambda(self, annealing):
        _check(self.L, self.L.svi_ls_phase_lambda(self.h, int(annealing)))

    # -- multi-GPU over peer memory -----------------------------------------------------------
    def peer_blob(self):
        """bytes describing this shard's exchange arena (CUDA IPC handle), to be all-gathered by the caller"""
        nb = self.L.svi_ls_peer_blob_bytes()
        buf = np.zeros(nb, dtype=np.uint8)
        _check(self.L, self.L.svi_ls_peer_export(self.h, _ptr(buf), nb))
        return buf

    def peer_attach(self, world, rank, bounds, blobs, chunks=0):
        bounds = np.ascontiguousarray(bounds, dtype=np.uint32)
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8)
        assert bounds.shape == (world + 1,) and blobs.size == world * self.L.svi_ls_peer_blob_bytes()
        _check(self.L, self.L.svi_ls_peer_attach(self.h, world, rank, _ptr(bounds), _ptr(blobs), chunks))

    @staticmethod
    def attach_local(engines, bounds, chunks=0):
        """all shards live in this process (one or several devices)"""
        world = len(engines)
        bounds = np.ascontiguousarray(bounds, dtype=np.uint32)
        arr = (C.c_void_p * world)(*[e.h.value for e in engines])
        for r, e in enumerate(engines):
            _check(e.L, e.L.svi_ls_peer_attach_local(e.h, world, r, _ptr(bounds), C.cast(arr, C.c_void_p), chunks))

    def mg_step(self, it, annealing, write_comm):
        _check(self.L, self.L.svi_ls_mg_step(self.h, it, int(annealing), int(write_comm)))

    def mg_publish_gamma(self):
        _check(self.L, self.L.svi_ls_mg_publish_gamma(self.h))

    def mg_share_gamma(self, on=True):
        _check(self.L, self.L.svi_ls_mg_share_gamma(self.h, int(on)))

    MG_PHASES = ("wait_b_rows", "phi+node", "allreduce_sum_s1_s2", "refresh", "wait_mphi_rows", "s3", "allreduce_s3+lambda",
                 "drain_own_pushes", "side_mphi_pushes_span", "side_b_pushes_span")

    def mg_timing(self, enable=True, read=False):
        """enable/disable per-phase timing of mg_step; read=True returns {phase: mean ms} over the recorded steps"""
        if not read:
            _check(self.L, self.L.svi_ls_mg_timing(self.h, int(enable), None, None))
            return None
        ms = np.zeros(10, dtype=np.float64)
        cnt = C.c_uint32()
        _check(self.L, self.L.svi_ls_mg_timing(self.h, int(enable), _ptr(ms), C.byref(cnt)))
        return dict(zip(self.MG_PHASES, [float(x) for x in ms])), cnt.value

    def membership_rows(self, first, count, out=None):
        bits = np.empty((count, self.words), dtype=np.uint32) if out is None else out
        assert bits.shape == (count, self.words) and bits.dtype == np.uint32 and bits.flags.c_contiguous
        _check(self.L, self.L.svi_ls_get_membership_rows(self.h, first, count, _ptr(bits)))
        return bits

    def membership_bits(self, out=None):
        """out: optional caller-owned [n, words] uint32 array (e.g. a view of pinned memory)"""
        bits = np.empty((self.n, self.words), dtype=np.uint32) if out is None else out
        assert bits.shape == (self.n, self.words) and bits.dtype == np.uint32 and bits.flags.c_contiguous
        _check(self.L, self.L.svi_ls_get_membership(self.h, _ptr(bits)))
        return bits

    def membership(self):
        """[n, k] uint8 matrix unpacked from the bit words."""
        bits = self.membership_bits()
        cols = np.arange(self.k)
        return ((bits[:, cols // 32] >> (cols % 32).astype(np.uint32)) & 1).astype(np.uint8)

    def heldout(self, p, q, y, epsilon=1e-30, out=None):
        p = np.ascontiguousarray(p, dtype=np.uint32)
        q = np.ascontiguousarray(q, dtype=np.uint32)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        out = np.empty(p.shape[0], dtype=np.float64) if out is None else out
        assert out.shape == (p.shape[0],) and out.dtype == np.float64
        _check(self.L, self.L.svi_ls_heldout(self.h, p.shape[0], _ptr(p), _ptr(q), _ptr(y), epsilon, _ptr(out)))
        return out

    def kvectors(self):
        v = [np.empty(self.k, dtype=np.float64) for _ in range(4)]
        _check(self.L, self.L.svi_ls_get_kvectors(self.h, *[_ptr(a) for a in v]))
        return dict(zip(("sum", "s1", "s2", "s3"), v))

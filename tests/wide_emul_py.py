"""ctypes binding of tests/cc/wide_emul.cc -- TEST INFRASTRUCTURE ONLY.

The K > 1024 kernels (svinet_b200/csrc/svi_ls_wide.cuh) compiled as host code over tests/cc/cuda_shim/ and run on
host threads; `WideEmulEngine` has the interface of svinet_b200.engine.LinkSamplingEngine so that the parity helpers
of the GPU tests apply unchanged.
"""
import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(REPO, "tests", "cc", "wide_emul.cc")
FA2_SRC = os.path.join(REPO, "tests", "cc", "fa2_wide_emul.cc")
CSRC = os.path.join(REPO, "svinet_b200", "csrc")
BUILD = os.path.join(REPO, "tests", "_build")

_libs = {}


def tsan_toolchain():
    """(compiler, path of its libtsan.so) of the first g++ here that ships ThreadSanitizer, or None"""
    for cxx in (os.environ.get("CXX"), "g++", "/usr/bin/g++"):
        if not cxx:
            continue
        try:
            rt = subprocess.run([cxx, "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
        except OSError:
            continue
        if os.path.isabs(rt) and os.path.exists(rt):
            return cxx, os.path.realpath(rt)
    return None


def build(threads=256, tsan=False, src=SRC):
    """g++ a harness; `threads` = kWideT of the emulated blocks (the product uses 256)."""
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "lib%s_t%d%s.so" % (os.path.basename(src)[:-3], threads, "_tsan" if tsan else ""))
    deps = [src, os.path.join(REPO, "tests", "cc", "cuda_shim", "cuda_runtime.h")] + [
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        cmd = [tsan_toolchain()[0] if tsan else os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
               "-Wno-unused-function", "-fvisibility=hidden", "-fno-gnu-unique", "-DSVI_WIDE_T=%d" % threads, "-I", os.path.join(REPO, "tests", "cc", "cuda_shim"),
               "-I", CSRC, "-o", out, src, "-lpthread"]
        if tsan:
            cmd[1:1] = ["-fsanitize=thread"]
        subprocess.check_call(cmd)
    return out


def lib(threads=256):
    if threads in _libs:
        return _libs[threads]
    L = C.CDLL(build(threads))
    vp = C.c_void_p
    L.we_threads.restype = C.c_uint32
    L.we_create.restype = vp
    L.we_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, vp, vp, C.c_double, C.c_double, C.c_double, C.c_uint32,
                            C.c_uint32, C.c_uint32]
    L.we_destroy.argtypes = [vp]
    L.we_info.argtypes = [vp] + [C.POINTER(C.c_uint32)] * 4
    L.we_set_state.argtypes = [vp, vp, vp]
    L.we_get_state.argtypes = [vp, vp, vp]
    L.we_set_converged.argtypes = [vp, vp]
    L.we_get_converged.argtypes = [vp, vp, vp]
    L.we_get_kvectors.argtypes = [vp, vp, vp, vp, vp]
    L.we_get_membership.argtypes = [vp, vp]
    L.we_step.argtypes = [vp, C.c_uint32, C.c_int, C.c_int]
    L.we_heldout.restype = C.c_uint64
    L.we_heldout.argtypes = [vp, C.c_uint64, vp, vp, vp, C.c_double, vp, C.c_uint32]
    assert L.we_threads() == threads
    _libs[threads] = L
    return L


def _ptr(a):
    return a.ctypes.data


class WideEmulEngine:
    def __init__(self, n, k, links, tl, alpha=None, eta0=1.0, eta1=1.0, ones=None, seg_len=16, sms=4, threads=256):
        self.L = lib(threads)
        links = np.ascontiguousarray(links, dtype=np.uint32).reshape(-1, 2)
        tl = np.ascontiguousarray(tl, dtype=np.float64)
        self.n, self.k, self.words = n, k, (k + 31) // 32
        self.h = self.L.we_create(n, k, links.shape[0], _ptr(links), _ptr(tl), 1.0 / k if alpha is None else alpha, eta0, eta1,
                                  links.shape[0] if ones is None else ones, seg_len, sms)
        assert self.h, "we_create: bad link"

    def close(self):
        if self.h:
            self.L.we_destroy(self.h)
            self.h = None

    def info(self):
        v = [C.c_uint32() for _ in range(4)]
        self.L.we_info(self.h, *[C.byref(x) for x in v])
        return dict(zip(("nseg", "nseg_lo", "blocks_node", "blocks_s3"), [x.value for x in v]))

    def set_state(self, gamma, lam):
        gamma = np.ascontiguousarray(gamma, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        assert gamma.shape == (self.n, self.k) and lam.shape == (self.k, 2)
        self.L.we_set_state(self.h, _ptr(gamma), _ptr(lam))

    def get_state(self):
        gamma = np.empty((self.n, self.k))
        lam = np.empty((self.k, 2))
        self.L.we_get_state(self.h, _ptr(gamma), _ptr(lam))
        return gamma, lam

    def set_converged(self, conv):
        conv = np.ascontiguousarray(conv, dtype=np.uint32)
        self.L.we_set_converged(self.h, _ptr(conv))

    def get_converged(self):
        conv, act = np.empty(self.n, dtype=np.uint32), np.empty(self.n, dtype=np.uint32)
        self.L.we_get_converged(self.h, _ptr(conv), _ptr(act))
        return conv, act

    def step(self, it, annealing, write_comm):
        self.L.we_step(self.h, it, int(annealing), int(write_comm))

    def kvectors(self):
        v = [np.empty(self.k) for _ in range(4)]
        self.L.we_get_kvectors(self.h, *[_ptr(a) for a in v])
        return dict(zip(("sum", "s1", "s2", "s3"), v))

    def membership(self):
        bits = np.empty((self.n, self.words), dtype=np.uint32)
        self.L.we_get_membership(self.h, _ptr(bits))
        cols = np.arange(self.k)
        return ((bits[:, cols // 32] >> (cols % 32).astype(np.uint32)) & 1).astype(np.uint8)

    def heldout(self, p, q, y, epsilon=1e-30, blocks=3):
        p = np.ascontiguousarray(p, dtype=np.uint32)
        q = np.ascontiguousarray(q, dtype=np.uint32)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        out = np.empty(p.shape[0])
        bad = self.L.we_heldout(self.h, p.shape[0], _ptr(p), _ptr(q), _ptr(y), epsilon, _ptr(out), blocks)
        return out, bad


# ---- `-rnode -stratified`, K > 512 (svinet_b200/csrc/svi_fa2_wide.cuh over tests/cc/fa2_wide_emul.cc) ---------------
_fa2_libs = {}


def fa2_lib(threads=256):
    if threads in _fa2_libs:
        return _fa2_libs[threads]
    L = C.CDLL(build(threads, src=FA2_SRC))
    vp = C.c_void_p
    L.fwe_threads.restype = C.c_uint32
    L.fwe_create.restype = vp
    L.fwe_create.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_double]
    L.fwe_destroy.argtypes = [vp]
    L.fwe_set_state.argtypes = [vp, vp, vp, C.c_uint64]
    L.fwe_get_state.argtypes = [vp, vp, vp]
    L.fwe_step.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, vp]
    L.fwe_folds.restype = C.c_uint32
    L.fwe_folds.argtypes = [vp]
    L.fwe_last_rounds.restype = C.c_uint64
    L.fwe_last_rounds.argtypes = [vp]
    L.fwe_heldout.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, C.c_uint32]
    L.fwe_phi_pair.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp, C.POINTER(C.c_uint32)]
    assert L.fwe_threads() == threads
    _fa2_libs[threads] = L
    return L


class Fa2WideEmulEngine:
    """the interface of svinet_b200.fa2_engine.Fa2Engine (what the parity tests use of it)"""

    def __init__(self, n, k, eager_blend=0, pair_blocks=3, online_iterations=50, fold_below=-230.0, threads=256):
        self.L = fa2_lib(threads)
        self.n, self.k = n, k
        self.h = self.L.fwe_create(n, k, int(eager_blend), pair_blocks, online_iterations, fold_below)

    def close(self):
        if self.h:
            self.L.fwe_destroy(self.h)
            self.h = None

    def set_state(self, gamma, lam, nodec=0):
        gamma = np.ascontiguousarray(gamma, dtype=np.float64)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        assert gamma.shape == (self.n, self.k) and lam.shape == (self.k, 2)
        self.L.fwe_set_state(self.h, _ptr(gamma), _ptr(lam), nodec)

    def get_state(self):
        gamma, lam = np.empty((self.n, self.k)), np.empty((self.k, 2))
        self.L.fwe_get_state(self.h, _ptr(gamma), _ptr(lam))
        return gamma, lam

    def step(self, it, typ, start, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        assert self.L.fwe_step(self.h, it, typ, start, pairs.shape[0], _ptr(pairs)) == 0, "bad minibatch"

    def folds(self):
        return self.L.fwe_folds(self.h)

    def last_rounds(self):
        return self.L.fwe_last_rounds(self.h)

    def heldout(self, p, q, y, blocks=2):
        p = np.ascontiguousarray(p, dtype=np.uint32)
        q = np.ascontiguousarray(q, dtype=np.uint32)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        out = np.empty(p.shape[0])
        self.L.fwe_heldout(self.h, p.shape[0], _ptr(p), _ptr(q), _ptr(y), _ptr(out), blocks)
        return out

    def phi_pair(self, p, q, y):
        p1, p2, r = np.empty(self.k), np.empty(self.k), C.c_uint32()
        self.L.fwe_phi_pair(self.h, p, q, int(y), _ptr(p1), _ptr(p2), C.byref(r))
        return p1, p2, r.value

"""ctypes binding of oracle/_build/liboracle_ls.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (svinet_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle_ls.so")

_lib = None


class OrcGraph(C.Structure):
    _fields_ = [("n_arg", C.c_uint32), ("n", C.c_uint32), ("singles", C.c_uint32), ("ones", C.c_uint32),
                ("seq2id", C.POINTER(C.c_uint32)), ("adj_off", C.POINTER(C.c_uint64)),
                ("adj", C.POINTER(C.c_uint32)), ("edges", C.POINTER(C.c_uint32))]


class OrcState(C.Structure):
    _fields_ = [("n", C.c_uint32), ("k", C.c_uint32),
                ("alpha", C.c_double), ("eta0", C.c_double), ("eta1", C.c_double),
                ("ones", C.c_uint32), ("nlinks", C.c_uint64),
                ("links", C.POINTER(C.c_uint32)), ("tl", C.POINTER(C.c_double)),
                ("gamma", C.POINTER(C.c_double)), ("gammanext", C.POINTER(C.c_double)),
                ("Elogpi", C.POINTER(C.c_double)), ("mphi", C.POINTER(C.c_double)),
                ("lambda_", C.POINTER(C.c_double)), ("lambdanext", C.POINTER(C.c_double)),
                ("Elogbeta", C.POINTER(C.c_double)),
                ("s1", C.POINTER(C.c_double)), ("s2", C.POINTER(C.c_double)),
                ("s3", C.POINTER(C.c_double)), ("sum", C.POINTER(C.c_double)),
                ("converged", C.POINTER(C.c_uint32)), ("active_comms", C.POINTER(C.c_uint32)),
                ("active_k", C.POINTER(C.c_uint16)), ("active_len", C.POINTER(C.c_uint32)),
                ("member", C.POINTER(C.c_uint8)),
                ("cnt_dense", C.c_uint64), ("cnt_sparse", C.c_uint64), ("cnt_shortcut", C.c_uint64)]


class OrcOptions(C.Structure):
    _fields_ = [("k", C.c_uint32), ("seed", C.c_double), ("heldout_ratio", C.c_double),
                ("accuracy", C.c_int), ("max_iterations", C.c_uint32), ("use_validation_stop", C.c_int),
                ("reportfreq", C.c_uint32), ("eta0", C.c_double), ("eta1", C.c_double), ("epsilon", C.c_double),
                ("init_communities", C.c_char_p)]


class OrcFa2Options(C.Structure):
    _fields_ = [("k", C.c_uint32), ("seed", C.c_double), ("heldout_ratio", C.c_double),
                ("max_iterations", C.c_uint32), ("reportfreq", C.c_uint32),
                ("eta0", C.c_double), ("eta1", C.c_double), ("epsilon", C.c_double),
                ("tau0", C.c_double), ("kappa", C.c_double), ("nodetau0", C.c_double), ("nodekappa", C.c_double),
                ("online_iterations", C.c_uint32), ("meanchangethresh", C.c_double),
                ("deterministic", C.c_int), ("nolambda", C.c_int)]


def build():
    """(Re)build the C restatement; cheap (one gcc call)."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("oracle_ls.c", "oracle_fa2.c", "oracle_fa2.h",
                                                                     "oracle_ls_omp.c", "oracle_ls.h")):
        build()
    L = C.CDLL(LIB_PATH)
    L.orc_graph_read.restype = C.POINTER(OrcGraph)
    L.orc_graph_read.argtypes = [C.c_char_p, C.c_uint32]
    L.orc_graph_from_pairs.restype = C.POINTER(OrcGraph)
    L.orc_graph_from_pairs.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
    L.orc_graph_free.argtypes = [C.POINTER(OrcGraph)]
    L.orc_graph_y.restype = C.c_int
    L.orc_graph_y.argtypes = [C.POINTER(OrcGraph), C.c_uint32, C.c_uint32]
    L.orc_state_alloc.restype = C.POINTER(OrcState)
    L.orc_state_alloc.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    L.orc_state_free.argtypes = [C.POINTER(OrcState)]
    L.orc_set_dir_exp.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.orc_prune.argtypes = [C.POINTER(OrcState)]
    L.orc_step.argtypes = [C.POINTER(OrcState), C.c_uint32, C.c_int, C.c_int]
    L.orc_step_omp.argtypes = [C.POINTER(OrcState), C.c_int, C.c_int, C.c_int]
    L.orc_omp_max_threads.restype = C.c_int
    L.orc_edge_likelihood.restype = C.c_double
    L.orc_edge_likelihood.argtypes = [C.POINTER(OrcState), C.c_uint32, C.c_uint32, C.c_int, C.c_double]
    L.orc_digamma.restype = C.c_double
    L.orc_digamma.argtypes = [C.c_double]
    L.orc_options_default.argtypes = [C.POINTER(OrcOptions), C.c_uint32]
    L.orc_model_create.restype = C.c_void_p
    L.orc_model_create.argtypes = [C.POINTER(OrcGraph), C.POINTER(OrcOptions)]
    L.orc_model_free.argtypes = [C.c_void_p]
    L.orc_model_run.restype = C.c_uint32
    L.orc_model_run.argtypes = [C.c_void_p, C.c_uint32]
    L.orc_model_state.restype = C.POINTER(OrcState)
    L.orc_model_state.argtypes = [C.c_void_p]
    for fn in ("orc_model_iter",):
        getattr(L, fn).restype = C.c_uint32
        getattr(L, fn).argtypes = [C.c_void_p]
    for fn in ("orc_model_annealing", "orc_model_write_comm", "orc_model_stopped"):
        getattr(L, fn).restype = C.c_int
        getattr(L, fn).argtypes = [C.c_void_p]
    L.orc_model_nvalidation.restype = C.c_uint64
    L.orc_model_nvalidation.argtypes = [C.c_void_p]
    L.orc_model_validation_pairs.restype = C.POINTER(C.c_uint32)
    L.orc_model_validation_pairs.argtypes = [C.c_void_p]
    L.orc_model_heldout.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.orc_model_write_outputs.restype = C.c_int
    L.orc_model_write_outputs.argtypes = [C.c_void_p, C.c_char_p]
    L.orc_rng_seed.argtypes = [C.c_void_p, C.c_ulong]
    L.orc_rng_next.restype = C.c_uint32
    L.orc_rng_next.argtypes = [C.c_void_p]
    L.orc_rng_uniform.restype = C.c_double
    L.orc_rng_uniform.argtypes = [C.c_void_p]
    L.orc_rng_uniform_int.restype = C.c_ulong
    L.orc_rng_uniform_int.argtypes = [C.c_void_p, C.c_ulong]
    # ---- oracle_fa2.h ----
    vp = C.c_void_p
    L.orc_fa2_options_default.argtypes = [C.POINTER(OrcFa2Options), C.c_uint32]
    L.orc_fa2_create.restype = vp
    L.orc_fa2_create.argtypes = [C.POINTER(OrcGraph), C.POINTER(OrcFa2Options)]
    L.orc_fa2_free.argtypes = [vp]
    L.orc_fa2_phi_pair.restype = C.c_uint32
    L.orc_fa2_phi_pair.argtypes = [C.c_uint32, vp, vp, vp, C.c_int, C.c_double, C.c_uint32, C.c_double, vp, vp]
    L.orc_fa2_plan.argtypes = [vp]
    L.orc_fa2_process.argtypes = [vp]
    L.orc_fa2_run.restype = C.c_uint32
    L.orc_fa2_run.argtypes = [vp, C.c_uint32]
    for fn in ("orc_fa2_n", "orc_fa2_k", "orc_fa2_iter", "orc_fa2_plan_type", "orc_fa2_plan_start"):
        getattr(L, fn).restype = C.c_uint32
        getattr(L, fn).argtypes = [vp]
    for fn in ("orc_fa2_plan_npairs", "orc_fa2_total_pairs_sampled", "orc_fa2_nheldout"):
        getattr(L, fn).restype = C.c_uint64
        getattr(L, fn).argtypes = [vp]
    L.orc_fa2_stopped.restype = C.c_int
    L.orc_fa2_stopped.argtypes = [vp]
    for fn in ("orc_fa2_gamma", "orc_fa2_lambda"):
        getattr(L, fn).restype = C.POINTER(C.c_double)
        getattr(L, fn).argtypes = [vp]
    L.orc_fa2_alpha.restype = C.c_double
    L.orc_fa2_alpha.argtypes = [vp]
    for fn in ("orc_fa2_shuffled", "orc_fa2_plan_pairs", "orc_fa2_heldout_pairs", "orc_fa2_heldout_sorted"):
        getattr(L, fn).restype = C.POINTER(C.c_uint32)
        getattr(L, fn).argtypes = [vp]
    L.orc_fa2_edge_likelihood.restype = C.c_double
    L.orc_fa2_edge_likelihood.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int]
    L.orc_fa2_heldout_log.restype = C.c_char_p
    L.orc_fa2_heldout_log.argtypes = [vp]
    L.orc_fa2_write_outputs.restype = C.c_int
    L.orc_fa2_write_outputs.argtypes = [vp, C.c_char_p]
    _lib = L
    return L


def _view(ptr, shape, dtype):
    n = int(np.prod(shape)) if len(shape) else 1
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    arr = np.ctypeslib.as_array(ptr, shape=(n,))
    return arr.view(dtype).reshape(shape)


class Graph:
    def __init__(self, ptr):
        self.ptr = ptr
        g = ptr.contents
        self.n_arg, self.n, self.singles, self.ones = g.n_arg, g.n, g.singles, g.ones
        self.seq2id = _view(g.seq2id, (g.n_arg,), np.uint32)
        self.adj_off = _view(g.adj_off, (g.n_arg + 1,), np.uint64)
        self.adj = _view(g.adj, (2 * g.ones,), np.uint32)
        self.edges = _view(g.edges, (g.ones, 2), np.uint32)

    @staticmethod
    def read(path, n_arg):
        p = lib().orc_graph_read(path.encode(), n_arg)
        if not p:
            raise IOError("oracle: cannot read %s" % path)
        return Graph(p)

    @staticmethod
    def from_pairs(pairs, n_arg):
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        return Graph(lib().orc_graph_from_pairs(pairs.ctypes.data, pairs.shape[0], n_arg))

    def y(self, a, b):
        return lib().orc_graph_y(self.ptr, a, b)

    def close(self):
        if self.ptr:
            lib().orc_graph_free(self.ptr)
            self.ptr = None


class State:
    """numpy views onto an orc_state (no copies; views die with the owner)."""

    def __init__(self, ptr, owner=None):
        self.ptr = ptr
        self.owner = owner

    @property
    def c(self):
        return self.ptr.contents

    def arr(self, name):
        s = self.c
        n, k = s.n, s.k
        kk = max(1, k // 10)
        shapes = {"links": ((s.nlinks, 2), np.uint32), "tl": ((n,), np.float64),
                  "gamma": ((n, k), np.float64), "gammanext": ((n, k), np.float64),
                  "Elogpi": ((n, k), np.float64), "mphi": ((n, k), np.float64),
                  "lambda_": ((k, 2), np.float64), "lambdanext": ((k, 2), np.float64),
                  "Elogbeta": ((k, 2), np.float64),
                  "s1": ((k,), np.float64), "s2": ((k,), np.float64), "s3": ((k,), np.float64),
                  "sum": ((k,), np.float64),
                  "converged": ((n,), np.uint32), "active_comms": ((n,), np.uint32),
                  "active_k": ((n, kk), np.uint16), "active_len": ((n,), np.uint32),
                  "member": ((n, k), np.uint8)}
        shape, dt = shapes[name]
        return _view(getattr(s, name), shape, dt)

    @staticmethod
    def alloc(n, k, nlinks):
        return State(lib().orc_state_alloc(n, k, nlinks))

    def step(self, it, annealing, write_comm):
        lib().orc_step(self.ptr, it, int(annealing), int(write_comm))

    def step_omp(self, annealing, write_comm, threads=0):
        lib().orc_step_omp(self.ptr, int(annealing), int(write_comm), int(threads))

    def refresh_expectations(self):
        s = self.c
        lib().orc_set_dir_exp(C.cast(s.gamma, C.c_void_p), C.cast(s.Elogpi, C.c_void_p), s.n, s.k)
        lib().orc_set_dir_exp(C.cast(s.lambda_, C.c_void_p), C.cast(s.Elogbeta, C.c_void_p), s.k, 2)

    def edge_likelihood(self, p, q, y, eps=1e-30):
        return lib().orc_edge_likelihood(self.ptr, p, q, y, eps)

    def free(self):
        if self.ptr and self.owner is None:
            lib().orc_state_free(self.ptr)
        self.ptr = None


class Model:
    """ctor + infer() + do_on_stop() of the reference's LinkSampling, restated."""

    def __init__(self, graph, k, **opts):
        L = lib()
        o = OrcOptions()
        L.orc_options_default(C.byref(o), k)
        for key, v in opts.items():
            if not hasattr(o, key):
                raise KeyError(key)
            setattr(o, key, v)
        self.graph = graph
        self.opts = o
        self.ptr = L.orc_model_create(graph.ptr, C.byref(o))
        self.state = State(L.orc_model_state(self.ptr), owner=self)

    def run(self, max_sweeps=0):
        return lib().orc_model_run(self.ptr, max_sweeps)

    iter = property(lambda self: lib().orc_model_iter(self.ptr))
    annealing = property(lambda self: bool(lib().orc_model_annealing(self.ptr)))
    write_comm = property(lambda self: bool(lib().orc_model_write_comm(self.ptr)))
    stopped = property(lambda self: bool(lib().orc_model_stopped(self.ptr)))

    def validation_pairs(self):
        n = lib().orc_model_nvalidation(self.ptr)
        return _view(lib().orc_model_validation_pairs(self.ptr), (n, 2), np.uint32).copy()

    def heldout(self):
        a, m0, m1 = C.c_double(), C.c_double(), C.c_double()
        k0, k1 = C.c_uint32(), C.c_uint32()
        lib().orc_model_heldout(self.ptr, C.byref(a), C.byref(m0), C.byref(m1), C.byref(k0), C.byref(k1))
        return a.value, m0.value, m1.value, k0.value, k1.value

    def write_outputs(self, d):
        os.makedirs(d, exist_ok=True)
        if lib().orc_model_write_outputs(self.ptr, d.encode()) != 0:
            raise IOError("oracle: cannot write outputs to %s" % d)

    def close(self):
        if self.ptr:
            lib().orc_model_free(self.ptr)
            self.ptr = None


class Rng:
    def __init__(self, seed=0):
        self.buf = C.create_string_buffer(624 * 4 + 8)
        lib().orc_rng_seed(self.buf, seed)

    def next(self):
        return lib().orc_rng_next(self.buf)

    def uniform(self):
        return lib().orc_rng_uniform(self.buf)

    def uniform_int(self, n):
        return lib().orc_rng_uniform_int(self.buf, n)


class Fa2Model:
    """ctor + infer() of the reference's FastAMM2 (`-rnode -stratified`), restated (oracle/oracle_fa2.c)."""

    def __init__(self, graph, k, **opts):
        L = lib()
        o = OrcFa2Options()
        L.orc_fa2_options_default(C.byref(o), k)
        for key, v in opts.items():
            if not hasattr(o, key):
                raise KeyError(key)
            setattr(o, key, v)
        self.graph, self.opts = graph, o
        self.ptr = L.orc_fa2_create(graph.ptr, C.byref(o))
        self.n, self.k = L.orc_fa2_n(self.ptr), L.orc_fa2_k(self.ptr)

    gamma = property(lambda self: _view(lib().orc_fa2_gamma(self.ptr), (self.n, self.k), np.float64))
    lambda_ = property(lambda self: _view(lib().orc_fa2_lambda(self.ptr), (self.k, 2), np.float64))
    iter = property(lambda self: lib().orc_fa2_iter(self.ptr))
    stopped = property(lambda self: bool(lib().orc_fa2_stopped(self.ptr)))
    alpha = property(lambda self: lib().orc_fa2_alpha(self.ptr))
    shuffled = property(lambda self: _view(lib().orc_fa2_shuffled(self.ptr), (self.n,), np.uint32))
    total_pairs_sampled = property(lambda self: lib().orc_fa2_total_pairs_sampled(self.ptr))

    def plan(self):
        """RNG draws + pair selection of the next iteration: (type, start, pairs[np,2])."""
        L = lib()
        L.orc_fa2_plan(self.ptr)
        npairs = L.orc_fa2_plan_npairs(self.ptr)
        pairs = _view(L.orc_fa2_plan_pairs(self.ptr), (npairs, 2), np.uint32).copy()
        return L.orc_fa2_plan_type(self.ptr), L.orc_fa2_plan_start(self.ptr), pairs

    def process(self):
        lib().orc_fa2_process(self.ptr)

    def run(self, max_steps=0):
        return lib().orc_fa2_run(self.ptr, max_steps)

    def heldout_pairs(self, sorted_=True):
        L = lib()
        n = L.orc_fa2_nheldout(self.ptr)
        fn = L.orc_fa2_heldout_sorted if sorted_ else L.orc_fa2_heldout_pairs
        return _view(fn(self.ptr), (n, 2), np.uint32).copy()

    def edge_likelihood(self, p, q, y):
        return lib().orc_fa2_edge_likelihood(self.ptr, p, q, y)

    def heldout_log(self):
        return lib().orc_fa2_heldout_log(self.ptr).decode()

    def write_outputs(self, d):
        os.makedirs(d, exist_ok=True)
        if lib().orc_fa2_write_outputs(self.ptr, d.encode()) != 0:
            raise IOError("oracle: cannot write outputs to %s" % d)

    def close(self):
        if self.ptr:
            lib().orc_fa2_free(self.ptr)
            self.ptr = None


def fa2_phi_pair(elogpi_p, elogpi_q, elogf, y, logepsilon=None, online_iterations=50, thresh=1e-5):
    """PhiCompute::update_phis_until_conv for one pair: returns (phi1, phi2, rounds)."""
    k = len(elogf)
    a = np.ascontiguousarray(elogpi_p, dtype=np.float64)
    b = np.ascontiguousarray(elogpi_q, dtype=np.float64)
    f = np.ascontiguousarray(elogf, dtype=np.float64)
    p1, p2 = np.empty(k), np.empty(k)
    le = float(np.log(1e-30)) if logepsilon is None else logepsilon
    r = lib().orc_fa2_phi_pair(k, a.ctypes.data, b.ctypes.data, f.ctypes.data, int(y), le, online_iterations, thresh,
                               p1.ctypes.data, p2.ctypes.data)
    return p1, p2, r

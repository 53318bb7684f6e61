"""ctypes binding of libbnkgpu.so (include/bnkgpu.h).

The library is the product; this module only marshals numpy / torch buffers into the C ABI.
There is no fallback: if the shared library is missing or no B200 is visible, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbnkgpu.so")
_LIB = None
NONE = 255

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
_vp = C.c_void_p


class BnkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libbnkgpu error %d: %s" % (code, msg))
        self.code = code


# every symbol include/bnkgpu.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "bnk_last_error", "bnk_version", "bnk_device_count",
    "bnk_model_create", "bnk_model_create_custom", "bnk_model_create_from_eigen", "bnk_model_destroy",
    "bnk_model_nstates", "bnk_model_get",
    "bnk_tree_create", "bnk_tree_destroy", "bnk_tree_size", "bnk_tree_n_internal", "bnk_tree_prune_orphans",
    "bnk_ws_create", "bnk_ws_destroy", "bnk_ws_sync", "bnk_ws_dims", "bnk_ws_configure", "bnk_ws_launch_count",
    "bnk_ws_device_bytes", "bnk_set_rates", "bnk_set_distances", "bnk_probs",
    "bnk_joint", "bnk_marginal", "bnk_loglik", "bnk_joint_dev", "bnk_marginal_dev", "bnk_loglik_dev",
    "bnk_joint_marginal_dev", "bnk_joint_marginal",
    "bnk_ws_phase_ms", "bnk_bep_presence",
    "bnk_ws_set_evidence_mode", "bnk_ws_set_pipeline",
    "bnk_bep_infer", "bnk_pogs_n_edges", "bnk_pogs_edges", "bnk_pogs_destroy",
    "bnk_mgpu_create", "bnk_mgpu_destroy", "bnk_mgpu_n_devices", "bnk_mgpu_shard", "bnk_mgpu_set_rates",
    "bnk_mgpu_set_distances", "bnk_mgpu_joint", "bnk_mgpu_marginal", "bnk_mgpu_loglik", "bnk_mgpu_upload",
    "bnk_mgpu_loglik_resident", "bnk_mgpu_reduction", "bnk_mgpu_launch_count",
    "bnk_indel_base_model", "bnk_indel_create", "bnk_indel_create_multi", "bnk_indel_destroy", "bnk_indel_set_alignment", "bnk_indel_set_rates", "bnk_indel_column_logl",
    "bnk_indel_column_priors", "bnk_indel_aln_logl", "bnk_indel_optimise_mulambda", "bnk_indel_probs",
    "bnk_indel_phase_ms", "bnk_indel_launch_count", "bnk_indel_device_bytes",
]
EVIDENCE_AUTO, EVIDENCE_LEAVES, EVIDENCE_ANYWHERE = 0, 1, 2


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_SO):
        raise ImportError(
            "bnkit_b200: %s is missing -- build it with `python -m bnkit_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback." % _SO)
    L = C.CDLL(_SO)
    L.bnk_last_error.restype = C.c_char_p
    L.bnk_version.restype = C.c_char_p
    L.bnk_model_create.argtypes = [C.c_char_p, _dp, C.POINTER(_vp)]
    L.bnk_model_create_custom.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_int, C.POINTER(_vp)]
    L.bnk_model_create_from_eigen.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.POINTER(_vp)]
    L.bnk_model_destroy.argtypes = [_vp]
    L.bnk_model_destroy.restype = None
    L.bnk_model_nstates.argtypes = [_vp]
    L.bnk_model_get.argtypes = [_vp] + [_dp] * 6
    L.bnk_tree_create.argtypes = [C.c_int32, _ip, _dp, C.POINTER(_vp)]
    L.bnk_tree_destroy.argtypes = [_vp]
    L.bnk_tree_destroy.restype = None
    L.bnk_tree_size.argtypes = [_vp]
    L.bnk_tree_n_internal.argtypes = [_vp]
    L.bnk_tree_prune_orphans.argtypes = [_vp, C.c_int32, _bp, _bp]
    L.bnk_ws_create.argtypes = [_vp, _vp, C.c_int32, C.c_int, _vp, C.POINTER(_vp)]
    L.bnk_ws_destroy.argtypes = [_vp]
    L.bnk_ws_destroy.restype = None
    L.bnk_ws_sync.argtypes = [_vp]
    L.bnk_ws_dims.argtypes = [_vp, _ip, _ip, _ip]
    L.bnk_ws_configure.argtypes = [_vp, C.c_int, C.c_int]
    L.bnk_ws_launch_count.argtypes = [_vp]
    L.bnk_ws_launch_count.restype = C.c_int64
    L.bnk_ws_device_bytes.argtypes = [_vp]
    L.bnk_ws_device_bytes.restype = C.c_int64
    L.bnk_set_rates.argtypes = [_vp, C.c_int32, _dp]
    L.bnk_set_distances.argtypes = [_vp, _dp]
    L.bnk_probs.argtypes = [_vp, C.c_int32, _dp, _dp, _dp]
    L.bnk_joint.argtypes = [_vp, C.c_int32, _vp, _vp, _vp]
    L.bnk_marginal.argtypes = [_vp, C.c_int32, _vp, _vp, _ip, C.c_int32, _vp, _vp]
    L.bnk_loglik.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _vp]
    L.bnk_joint_dev.argtypes = [_vp, C.c_int32, _vp, _vp, _vp]
    L.bnk_marginal_dev.argtypes = [_vp, C.c_int32, _vp, _vp, _ip, C.c_int32, _vp, _vp]
    L.bnk_loglik_dev.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _vp]
    L.bnk_joint_marginal_dev.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _ip, C.c_int32, _vp, _vp]
    L.bnk_joint_marginal.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _ip, C.c_int32, _vp, _vp]
    L.bnk_ws_phase_ms.argtypes = [_vp, C.c_int, C.POINTER(C.c_float)]
    L.bnk_bep_presence.argtypes = [_vp, C.c_int32, _bp, C.c_int, _bp, _ip]
    L.bnk_bep_infer.argtypes = [_vp, C.c_int32, _bp, C.c_int, _bp, _ip, C.POINTER(_vp)]
    L.bnk_pogs_n_edges.argtypes = [_vp, C.c_int32]
    L.bnk_pogs_edges.argtypes = [_vp, C.c_int32, _ip, _ip, _bp]
    L.bnk_pogs_destroy.argtypes = [_vp]
    L.bnk_pogs_destroy.restype = None
    L.bnk_ws_set_evidence_mode.argtypes = [_vp, C.c_int]
    L.bnk_ws_set_pipeline.argtypes = [_vp, C.c_int32]
    L.bnk_mgpu_create.argtypes = [_vp, _vp, C.c_int32, C.c_int, C.POINTER(C.c_int), C.POINTER(_vp)]
    L.bnk_mgpu_destroy.argtypes = [_vp]
    L.bnk_mgpu_destroy.restype = None
    L.bnk_mgpu_n_devices.argtypes = [_vp]
    L.bnk_mgpu_shard.argtypes = [_vp, C.c_int32, C.c_int, _ip, _ip]
    L.bnk_mgpu_set_rates.argtypes = [_vp, C.c_int32, _dp]
    L.bnk_mgpu_set_distances.argtypes = [_vp, _dp]
    L.bnk_mgpu_joint.argtypes = [_vp, C.c_int32, _vp, _vp, _vp]
    L.bnk_mgpu_marginal.argtypes = [_vp, C.c_int32, _vp, _vp, _ip, C.c_int32, _vp, _vp]
    L.bnk_mgpu_loglik.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _vp]
    L.bnk_mgpu_upload.argtypes = [_vp, C.c_int32, _vp, _vp]
    L.bnk_mgpu_loglik_resident.argtypes = [_vp, _dp]
    L.bnk_mgpu_reduction.argtypes = [_vp]
    L.bnk_mgpu_reduction.restype = C.c_char_p
    L.bnk_mgpu_launch_count.argtypes = [_vp]
    L.bnk_mgpu_launch_count.restype = C.c_int64
    L.bnk_indel_base_model.argtypes = [C.c_char_p, C.POINTER(_vp)]
    L.bnk_indel_create.argtypes = [_vp, _vp, C.c_int32, C.c_int, C.POINTER(_vp)]
    L.bnk_indel_create_multi.argtypes = [_vp, _vp, C.c_int32, C.c_int, C.POINTER(C.c_int), C.POINTER(_vp)]
    L.bnk_indel_destroy.argtypes = [_vp]
    L.bnk_indel_destroy.restype = None
    L.bnk_indel_set_alignment.argtypes = [_vp, C.c_int32, _bp]
    L.bnk_indel_set_rates.argtypes = [_vp, C.c_double, C.c_double]
    L.bnk_indel_column_logl.argtypes = [_vp, C.c_double, C.c_double, _dp, _dp]
    L.bnk_indel_column_priors.argtypes = [_vp, C.c_double, C.c_int32, _dp, _dp]
    L.bnk_indel_aln_logl.argtypes = [_vp, C.c_double, _dp]
    L.bnk_indel_optimise_mulambda.argtypes = [_vp, C.c_double, C.c_double, C.c_double, _dp, _ip]
    L.bnk_indel_probs.argtypes = [_vp, C.c_double, _dp, _dp, _dp]
    L.bnk_indel_phase_ms.argtypes = [_vp, C.c_int]
    L.bnk_indel_phase_ms.restype = C.c_double
    L.bnk_indel_launch_count.argtypes = [_vp]
    L.bnk_indel_launch_count.restype = C.c_int64
    L.bnk_indel_device_bytes.argtypes = [_vp]
    L.bnk_indel_device_bytes.restype = C.c_int64
    _LIB = L
    return L


def check(rc):
    if rc < 0:
        raise BnkError(rc, lib().bnk_last_error().decode("utf-8", "replace"))
    return rc


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _ptr(x):
    """address of a numpy array, a torch tensor (host or device), an int, or None"""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    return x.data_ptr()  # torch tensor


class Model:
    """bnk_model handle <-> bn.ctmc.SubstModel."""

    def __init__(self, name=None, F=None, custom=None, eigen=None, indel_base=False):
        L = lib()
        h = _vp()
        self._keep = []
        if indel_base:   # the base model IndelPeeler.optimiseMuLambda builds for a name
            check(L.bnk_indel_base_model(name.encode(), C.byref(h)))
        elif eigen is not None:
            K, Fe, evec, ievec, ev = eigen
            arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (Fe, evec, ievec, ev)]
            check(L.bnk_model_create_from_eigen(int(K), *[_d(a) for a in arrs], C.byref(h)))
        elif custom is not None:
            K, Fc, M, symmetric, normalise = custom
            Fc = np.ascontiguousarray(Fc, dtype=np.float64)
            M = np.ascontiguousarray(M, dtype=np.float64)
            check(L.bnk_model_create_custom(int(K), _d(Fc), _d(M), int(symmetric), int(normalise), C.byref(h)))
        else:
            Fa = None if F is None else np.ascontiguousarray(F, dtype=np.float64)
            check(L.bnk_model_create(name.encode(), _d(Fa), C.byref(h)))
        self.h = h
        self.name = name
        self.K = L.bnk_model_nstates(h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().bnk_model_destroy(self.h)
            self.h = None

    def parts(self, want_R=True):
        K = self.K
        R, ev, iev = (np.zeros((K, K)) for _ in range(3))
        lam, F, prior = (np.zeros(K) for _ in range(3))
        check(lib().bnk_model_get(self.h, _d(R) if want_R else None, _d(ev), _d(iev), _d(lam), _d(F), _d(prior)))
        return dict(R=R if want_R else None, evec=ev, ievec=iev, eval=lam, F=F, prior=prior)


class Tree:
    """bnk_tree handle <-> dat.phylo.IdxTree (parent[] / distance[] arrays)."""

    def __init__(self, parent, dist=None):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.dist = None if dist is None else np.ascontiguousarray(dist, dtype=np.float64)
        h = _vp()
        check(lib().bnk_tree_create(len(self.parent), _i(self.parent), _d(self.dist), C.byref(h)))
        self.h = h
        self.n = len(self.parent)
        self.n_internal = lib().bnk_tree_n_internal(h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().bnk_tree_destroy(self.h)
            self.h = None

    def internal_nodes(self):
        has_child = np.zeros(self.n, dtype=bool)
        has_child[self.parent[self.parent >= 0]] = True
        return np.nonzero(has_child)[0].astype(np.int32)

    def prune_orphans(self, present):
        present = np.ascontiguousarray(present, dtype=np.uint8)
        out = np.empty_like(present)
        check(lib().bnk_tree_prune_orphans(self.h, present.shape[1], present.ctypes.data_as(_bp), out.ctypes.data_as(_bp)))
        return out


    def bep_presence(self, obs, device=-1, want_info=False):
        """Prediction.PredictByBidirEdgeParsimony as a mask transform: obs [n][n_cols] (255 = gap; only the
        rows of childless nodes are read) -> present [n][n_cols] (extants: residue present; ancestors: the
        columns of their inferred partial-order graph).  info = (ancestors without a start-to-end path after
        the first pass, patch rounds, largest number of distinct edge targets in a column, kernel launches, us on the device sweeps incl.
        their copies, us of host graph assembly, us in total, us in the kernels alone)."""
        obs = np.ascontiguousarray(obs, dtype=np.uint8)
        out = np.empty_like(obs)
        info = np.zeros(8, dtype=np.int32)
        check(lib().bnk_bep_presence(self.h, obs.shape[1], obs.ctypes.data_as(_bp), int(device), out.ctypes.data_as(_bp), _i(info)))
        return (out, info) if want_info else out


    def bep_infer(self, obs, device=-1):
        """bep_presence plus the ancestors' partial-order graphs: returns (present, {node: (from[], to[], flags[])})
        with from = -1 for the start terminal, to = n_cols for the end terminal, flags bit 0 / 1 = forward / backward."""
        obs = np.ascontiguousarray(obs, dtype=np.uint8)
        out = np.empty_like(obs)
        info = np.zeros(8, dtype=np.int32)
        h = _vp()
        L = lib()
        check(L.bnk_bep_infer(self.h, obs.shape[1], obs.ctypes.data_as(_bp), int(device), out.ctypes.data_as(_bp), _i(info), C.byref(h)))
        graphs = {}
        try:
            for v in self.internal_nodes():
                ne = check(L.bnk_pogs_n_edges(h, int(v)))
                fr, to, fl = np.empty(ne, np.int32), np.empty(ne, np.int32), np.empty(ne, np.uint8)
                if ne:
                    check(L.bnk_pogs_edges(h, int(v), _i(fr), _i(to), fl.ctypes.data_as(_bp)))
                graphs[int(v)] = (fr, to, fl)
        finally:
            L.bnk_pogs_destroy(h)
        return out, graphs


class Workspace:
    """bnk_ws handle: device buffers + stream for one (model, tree)."""

    def __init__(self, model, tree, max_cols, device=-1, stream=None):
        self.model, self.tree = model, tree  # keep alive
        h = _vp()
        # stream=None: a private non-blocking stream.  A caller that hands over torch's DEFAULT stream passes the
        # handle 0, which is also NULL: map it to cudaStreamLegacy (0x1), the same stream under its explicit name,
        # so the work is ordered with torch's events / NCCL calls on that stream instead of racing them.
        if stream is not None and int(stream) == 0:
            stream = 1
        check(lib().bnk_ws_create(model.h, tree.h, int(max_cols), int(device), stream, C.byref(h)))
        self.h = h
        self.K = model.K
        self.n = tree.n

    def close(self):
        if getattr(self, "h", None):
            lib().bnk_ws_destroy(self.h)
            self.h = None

    __del__ = close

    def sync(self):
        check(lib().bnk_ws_sync(self.h))

    def configure(self, tile_cols=0, cols_per_thread=0):
        check(lib().bnk_ws_configure(self.h, int(tile_cols), int(cols_per_thread)))

    def launch_count(self):
        return lib().bnk_ws_launch_count(self.h)

    def device_bytes(self):
        return lib().bnk_ws_device_bytes(self.h)

    def set_rates(self, n_cols, rates=None):
        r = None if rates is None else np.ascontiguousarray(rates, dtype=np.float64)
        check(lib().bnk_set_rates(self.h, int(n_cols), _d(r)))

    def set_distances(self, dist):
        d = np.ascontiguousarray(dist, dtype=np.float64)
        check(lib().bnk_set_distances(self.h, _d(d)))

    def probs(self, times, want_log=True):
        t = np.ascontiguousarray(np.atleast_1d(times), dtype=np.float64)
        P = np.zeros((len(t), self.K, self.K))
        Lg = np.zeros_like(P) if want_log else None
        check(lib().bnk_probs(self.h, len(t), _d(t), _d(P), _d(Lg)))
        return (P, Lg) if want_log else P

    def phase_ms(self, phase):
        ms = C.c_float(0)
        check(lib().bnk_ws_phase_ms(self.h, phase, C.byref(ms)))
        return ms.value

    # ---- host-buffer entry points (numpy or pinned torch tensors) ----
    def set_evidence_mode(self, mode):
        """EVIDENCE_AUTO (scan + validate per call), EVIDENCE_LEAVES / EVIDENCE_ANYWHERE (caller's word, no sync)."""
        check(lib().bnk_ws_set_evidence_mode(self.h, int(mode)))

    def set_pipeline(self, chunk_cols):
        """column chunking of the host-buffer calls: 0 automatic, < 0 off, > 0 columns per chunk"""
        check(lib().bnk_ws_set_pipeline(self.h, int(chunk_cols)))

    def _check_shapes(self, obs, present, out=None, out_shape=None):
        """a short host array would make the library read / write past its end"""
        if len(obs.shape) != 2 or obs.shape[0] != self.n:
            raise ValueError("obs must be [n_nodes = %d][n_cols], got %r" % (self.n, tuple(obs.shape)))
        if present is not None and tuple(present.shape) != tuple(obs.shape):
            raise ValueError("present must have the shape of obs %r, got %r" % (tuple(obs.shape), tuple(present.shape)))
        if out is not None and out_shape is not None and tuple(out.shape) != tuple(out_shape):
            raise ValueError("output buffer must be %r, got %r" % (tuple(out_shape), tuple(out.shape)))

    def joint(self, obs, present=None, out=None):
        obs = _as_u8(obs)
        n_cols = obs.shape[1]
        present = None if present is None else _as_u8(present)
        self._check_shapes(obs, present, out, (self.n, n_cols))
        if out is None:
            out = np.empty((self.n, n_cols), dtype=np.uint8)
        check(lib().bnk_joint(self.h, n_cols, _ptr(obs), _ptr(present), _ptr(out)))
        return out

    def marginal(self, obs, present=None, query=None, want_logl=True, out=None):
        obs = _as_u8(obs)
        n_cols = obs.shape[1]
        present = None if present is None else _as_u8(present)
        if query is None:
            nq, q = -1, None
            rows = self.tree.n_internal
        else:
            q = np.ascontiguousarray(query, dtype=np.int32)
            nq = rows = len(q)
        self._check_shapes(obs, present, out, (rows, n_cols, self.K))
        if out is None:
            out = np.empty((rows, n_cols, self.K))
        logl = np.empty(n_cols) if want_logl else None
        check(lib().bnk_marginal(self.h, n_cols, _ptr(obs), _ptr(present), _i(q), nq, _ptr(out), _ptr(logl)))
        return out, logl

    def joint_marginal(self, obs, present=None, query=None, want_logl=True):
        """joint() and marginal() of the same alignment in one call: (states, posteriors, logl)"""
        obs = _as_u8(obs)
        n_cols = obs.shape[1]
        present = None if present is None else _as_u8(present)
        if query is None:
            nq, q = -1, None
            rows = self.tree.n_internal
        else:
            q = np.ascontiguousarray(query, dtype=np.int32)
            nq = rows = len(q)
        self._check_shapes(obs, present)
        states = np.empty((self.n, n_cols), dtype=np.uint8)
        post = np.empty((rows, n_cols, self.K))
        logl = np.empty(n_cols) if want_logl else None
        check(lib().bnk_joint_marginal(self.h, n_cols, _ptr(obs), _ptr(present), _ptr(states), _i(q), nq, _ptr(post), _ptr(logl)))
        return states, post, logl

    def loglik(self, obs, present=None):
        obs = _as_u8(obs)
        n_cols = obs.shape[1]
        present = None if present is None else _as_u8(present)
        self._check_shapes(obs, present)
        logl = np.empty(n_cols)
        total = np.zeros(1)
        check(lib().bnk_loglik(self.h, n_cols, _ptr(obs), _ptr(present), _ptr(logl), _ptr(total)))
        return logl, float(total[0])

    # ---- device-resident entry points (torch CUDA tensors or raw device addresses) ----
    def joint_dev(self, n_cols, d_obs, d_present, d_out_state):
        check(lib().bnk_joint_dev(self.h, int(n_cols), _ptr(d_obs), _ptr(d_present), _ptr(d_out_state)))

    def marginal_dev(self, n_cols, d_obs, d_present, d_out_post, d_out_logl=None, query=None):
        if query is None:
            nq, q = -1, None
        else:
            q = np.ascontiguousarray(query, dtype=np.int32)
            nq = len(q)
        check(lib().bnk_marginal_dev(self.h, int(n_cols), _ptr(d_obs), _ptr(d_present), _i(q), nq,
                                     _ptr(d_out_post), _ptr(d_out_logl)))

    def joint_marginal_dev(self, n_cols, d_obs, d_present, d_out_state, d_out_post, d_out_logl=None, query=None):
        """joint_dev + marginal_dev of the same alignment in one call (the two passes overlap on two streams)"""
        if query is None:
            nq, q = -1, None
        else:
            q = np.ascontiguousarray(query, dtype=np.int32)
            nq = len(q)
        check(lib().bnk_joint_marginal_dev(self.h, int(n_cols), _ptr(d_obs), _ptr(d_present), _ptr(d_out_state), _i(q), nq,
                                           _ptr(d_out_post), _ptr(d_out_logl)))

    def loglik_dev(self, n_cols, d_obs, d_present, d_out_logl, d_out_sum):
        check(lib().bnk_loglik_dev(self.h, int(n_cols), _ptr(d_obs), _ptr(d_present), _ptr(d_out_logl), _ptr(d_out_sum)))


class Indel:
    """bnk_indel handle: gap-augmented peeling (asr.IndelPeeler under bn.ctmc.GapSubstModel) of one alignment"""

    def __init__(self, model, tree, max_cols, device=-1, n_devices=None, devices=None):
        """n_devices / devices given: one handle over several devices (bnk_indel_create_multi; 0 = all visible)"""
        self.model, self.tree = model, tree
        h = _vp()
        if n_devices is None and devices is None:
            check(lib().bnk_indel_create(model.h, tree.h, int(max_cols), int(device), C.byref(h)))
        else:
            arr = None
            if devices is not None:
                arr = (C.c_int * len(devices))(*[int(d) for d in devices])
                n_devices = len(devices)
            check(lib().bnk_indel_create_multi(model.h, tree.h, int(max_cols), int(n_devices), arr, C.byref(h)))
        self.h = h
        self.K, self.n, self.n_cols = model.K, tree.n, 0

    def close(self):
        if getattr(self, "h", None):
            lib().bnk_indel_destroy(self.h)
            self.h = None

    __del__ = close

    def set_alignment(self, obs):
        obs = _as_u8(obs)
        if obs.ndim != 2 or obs.shape[0] != self.n:
            raise ValueError("obs must be [n_nodes][n_cols]")
        check(lib().bnk_indel_set_alignment(self.h, obs.shape[1], obs.ctypes.data_as(_bp)))
        self.n_cols = obs.shape[1]

    def set_rates(self, mu, lam):
        check(lib().bnk_indel_set_rates(self.h, float(mu), float(lam)))

    def column_logl(self, rate=1.0, geom=0.01):
        out = np.empty(self.n_cols)
        s = C.c_double(0)
        check(lib().bnk_indel_column_logl(self.h, float(rate), float(geom), _d(out), C.cast(C.byref(s), _dp)))
        return out, s.value

    def column_priors(self, rates, geom):
        rates = np.ascontiguousarray(rates, dtype=np.float64)
        out = np.empty((self.n_cols, len(rates)))
        check(lib().bnk_indel_column_priors(self.h, float(geom), len(rates), _d(rates), _d(out)))
        return out

    def aln_logl(self, geom):
        s = C.c_double(0)
        check(lib().bnk_indel_aln_logl(self.h, float(geom), C.cast(C.byref(s), _dp)))
        return s.value

    def optimise_mulambda(self, lo, hi, geom):
        x = C.c_double(0)
        ne = C.c_int32(0)
        check(lib().bnk_indel_optimise_mulambda(self.h, float(lo), float(hi), float(geom), C.cast(C.byref(x), _dp), C.cast(C.byref(ne), _ip)))
        return x.value, ne.value

    def probs(self, t):
        K1 = self.K + 1
        P, kg, F = np.empty((K1, K1)), np.empty(2), np.empty(K1)
        check(lib().bnk_indel_probs(self.h, float(t), _d(P), _d(kg), _d(F)))
        return P, kg[0], kg[1], F

    def phase_ms(self, phase):
        return lib().bnk_indel_phase_ms(self.h, int(phase))

    def launch_count(self):
        return lib().bnk_indel_launch_count(self.h)


class MultiGPU:
    """bnk_mgpu handle: one process, several devices, columns in contiguous blocks (the device-side counterpart of
    the reference's thread pool over columns, ThreadedDecorators.java:28-72).  Host buffers in, host buffers out."""

    def __init__(self, model, tree, max_cols, n_devices=0, devices=None):
        self.model, self.tree = model, tree
        h = _vp()
        dv = None
        if devices is not None:
            dv = (C.c_int * len(devices))(*[int(d) for d in devices])
            n_devices = len(devices)
        check(lib().bnk_mgpu_create(model.h, tree.h, int(max_cols), int(n_devices), dv, C.byref(h)))
        self.h = h
        self.K, self.n = model.K, tree.n
        self.n_devices = lib().bnk_mgpu_n_devices(h)

    def close(self):
        if getattr(self, "h", None):
            lib().bnk_mgpu_destroy(self.h)
            self.h = None

    __del__ = close

    def shard(self, n_cols, idx):
        c0, nc = C.c_int32(0), C.c_int32(0)
        check(lib().bnk_mgpu_shard(self.h, int(n_cols), int(idx), C.byref(c0), C.byref(nc)))
        return c0.value, nc.value

    def set_rates(self, n_cols, rates=None):
        r = None if rates is None else np.ascontiguousarray(rates, dtype=np.float64)
        check(lib().bnk_mgpu_set_rates(self.h, int(n_cols), _d(r)))

    def set_distances(self, dist):
        d = np.ascontiguousarray(dist, dtype=np.float64)
        check(lib().bnk_mgpu_set_distances(self.h, _d(d)))

    def joint(self, obs, present=None, out=None):
        obs = _as_u8(obs)
        present = None if present is None else _as_u8(present)
        n_cols = obs.shape[1]
        if out is None:
            out = np.empty((self.n, n_cols), dtype=np.uint8)
        check(lib().bnk_mgpu_joint(self.h, n_cols, _ptr(obs), _ptr(present), _ptr(out)))
        return out

    def marginal(self, obs, present=None, query=None, want_logl=True, out=None):
        obs = _as_u8(obs)
        present = None if present is None else _as_u8(present)
        n_cols = obs.shape[1]
        if query is None:
            nq, q, rows = -1, None, self.tree.n_internal
        else:
            q = np.ascontiguousarray(query, dtype=np.int32)
            nq = rows = len(q)
        if out is None:
            out = np.empty((rows, n_cols, self.K))
        logl = np.empty(n_cols) if want_logl else None
        check(lib().bnk_mgpu_marginal(self.h, n_cols, _ptr(obs), _ptr(present), _i(q), nq, _ptr(out), _ptr(logl)))
        return out, logl

    def loglik(self, obs, present=None):
        obs = _as_u8(obs)
        present = None if present is None else _as_u8(present)
        n_cols = obs.shape[1]
        logl, total = np.empty(n_cols), np.zeros(1)
        check(lib().bnk_mgpu_loglik(self.h, n_cols, _ptr(obs), _ptr(present), _ptr(logl), _ptr(total)))
        return logl, float(total[0])

    def upload(self, obs, present=None):
        obs = _as_u8(obs)
        present = None if present is None else _as_u8(present)
        check(lib().bnk_mgpu_upload(self.h, obs.shape[1], _ptr(obs), _ptr(present)))

    def loglik_resident(self):
        total = np.zeros(1)
        check(lib().bnk_mgpu_loglik_resident(self.h, _d(total)))
        return float(total[0])

    def reduction(self):
        return lib().bnk_mgpu_reduction(self.h).decode()

    def launch_count(self):
        return lib().bnk_mgpu_launch_count(self.h)


def _as_u8(a):
    if isinstance(a, np.ndarray):
        return np.ascontiguousarray(a, dtype=np.uint8)
    return a  # torch tensor: caller guarantees uint8, contiguous


def device_count():
    return lib().bnk_device_count()

"""ctypes binding of the CPU oracle (oracle/libbnk_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under bnkit_b200/
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
NONE = 255

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_bp = C.POINTER(C.c_uint8)


def build(force=False):
    so = os.path.join(_HERE, "libbnk_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B", "libbnk_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libbnk_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.om_exp.restype = C.c_double; L.om_exp.argtypes = [C.c_double]
        L.om_log.restype = C.c_double; L.om_log.argtypes = [C.c_double]
        L.om_logsum.restype = C.c_double; L.om_logsum.argtypes = [C.c_double, C.c_double]
        L.om_model_create.restype = C.c_void_p; L.om_model_create.argtypes = [C.c_char_p, _dp]
        L.om_model_create_custom.restype = C.c_void_p
        L.om_model_create_custom.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_int]
        L.om_model_free.argtypes = [C.c_void_p]
        L.om_model_nstates.argtypes = [C.c_void_p]
        L.om_model_get.argtypes = [C.c_void_p] + [_dp] * 6
        L.om_getprobs.argtypes = [C.c_void_p, C.c_double, _dp]
        L.om_eigen.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.om_prune.argtypes = [C.c_int, _ip, _bp, C.c_int, _bp, _ip]
        for name in ("om_ve_joint",):
            if hasattr(L, name):
                getattr(L, name).argtypes = [C.c_void_p, C.c_int, _ip, _dp, _bp, _bp, C.c_double, _bp, _dp]
        if hasattr(L, "om_ve_marginal"):
            L.om_ve_marginal.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _bp, _bp, C.c_double, C.c_int, _dp]
        if hasattr(L, "om_ve_loglik"):
            L.om_ve_loglik.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _bp, _bp, C.c_double, _dp]
        L.om_fast_joint.argtypes = [C.c_void_p, C.c_int, _ip, _dp, C.c_int, _bp, _bp, _dp, _bp, _dp, C.c_int]
        L.om_fast_marginal.argtypes = [C.c_void_p, C.c_int, _ip, _dp, C.c_int, _bp, _bp, _dp, _dp, _dp, C.c_int]
        L.om_brute.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _bp, _bp, C.c_double, C.c_long, _bp, _dp, _dp, _dp]
        L.om_gap_model_create.restype = C.c_void_p; L.om_gap_model_create.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.om_gap_model_free.argtypes = [C.c_void_p]
        L.om_gap_nstates.argtypes = [C.c_void_p]
        L.om_gap_get.argtypes = [C.c_void_p] + [_dp] * 5
        L.om_gap_getprobs.argtypes = [C.c_void_p, C.c_double, _dp]
        for f in (L.om_gap_ksi, L.om_gap_gamma):
            f.restype = C.c_double; f.argtypes = [C.c_void_p, C.c_double]
        L.om_logsumexp.restype = C.c_double; L.om_logsumexp.argtypes = [_dp, C.c_int]
        L.om_logm1exp.restype = C.c_double; L.om_logm1exp.argtypes = [C.c_double]
        L.om_indel_column_logl.argtypes = [C.c_void_p, C.c_int, _ip, _dp, C.c_int, _bp, C.c_double, C.c_double, _dp, C.c_int]
        L.om_indel_prob_extra_col.restype = C.c_double
        L.om_indel_prob_extra_col.argtypes = [C.c_void_p, C.c_int, _ip, _dp, C.c_double]
        L.om_indel_aln_logl.argtypes = [C.c_void_p, C.c_int, _ip, _dp, C.c_int, _bp, C.c_double, _dp, C.c_int]
        L.om_indel_optimise_mulambda.restype = C.c_double
        L.om_indel_optimise_mulambda.argtypes = [C.c_void_p, C.c_int, _ip, _dp, C.c_int, _bp, C.c_double, C.c_double, C.c_double, _ip, C.c_int]
        _LIB = L
    return _LIB


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _b(a):
    return None if a is None else a.ctypes.data_as(_bp)


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class Model:
    """om_model handle: SubstModel restatement (SubstModel.java:80-110)."""

    def __init__(self, name=None, F=None, custom=None):
        L = lib()
        if custom is not None:
            K, Fc, M, symmetric, normalise = custom
            Fc = _arr(Fc, np.float64); M = _arr(M, np.float64)
            self.h = L.om_model_create_custom(K, _d(Fc), _d(M), int(symmetric), int(normalise))
        else:
            Fa = _arr(F, np.float64)
            self.h = L.om_model_create(name.encode(), _d(Fa))
        if not self.h:
            raise ValueError("oracle: cannot create model %r" % (name,))
        self.K = L.om_model_nstates(self.h)
        self.name = name

    def __del__(self):
        if getattr(self, "h", None):
            lib().om_model_free(self.h)
            self.h = None

    def parts(self):
        K = self.K
        R, ev, iev = (np.zeros((K, K)) for _ in range(3))
        lam, F, prior = (np.zeros(K) for _ in range(3))
        lib().om_model_get(self.h, _d(R), _d(ev), _d(iev), _d(lam), _d(F), _d(prior))
        return dict(R=R, evec=ev, ievec=iev, eval=lam, F=F, prior=prior)

    def probs(self, t):
        P = np.zeros((self.K, self.K))
        lib().om_getprobs(self.h, float(t), _d(P))
        return P


def eigen(A):
    A = _arr(A, np.float64)
    n = A.shape[0]
    ev, iev = np.zeros((n, n)), np.zeros((n, n))
    wr, wi = np.zeros(n), np.zeros(n)
    rc = lib().om_eigen(n, _d(A), _d(ev), _d(wr), _d(wi), _d(iev))
    if rc:
        raise ArithmeticError("om_eigen rc=%d" % rc)
    return ev, wr, wi, iev


def prune(parent, present, remove_orphans=True):
    parent = _arr(parent, np.int32); present = _arr(present, np.uint8)
    out = np.zeros_like(present)
    nr = C.c_int(0)
    cnt = lib().om_prune(len(parent), _i(parent), _b(present), int(remove_orphans), _b(out), C.byref(nr))
    return out, cnt, nr.value


def _col_args(parent, dist, present, obs):
    parent = _arr(parent, np.int32); dist = _arr(dist, np.float64)
    present = _arr(present, np.uint8); obs = _arr(obs, np.uint8)
    return parent, dist, present, obs


def fast_joint(model, parent, dist, obs, present=None, rates=None, nthreads=1, want_logmax=False):
    """obs, present: [n_nodes][n_cols] uint8 (col fastest). Returns states [n_nodes][n_cols]."""
    parent, dist, present, obs = _col_args(parent, dist, present, obs)
    n, ncols = obs.shape
    rates = _arr(rates, np.float64)
    out = np.full((n, ncols), NONE, dtype=np.uint8)
    lm = np.zeros(ncols)
    rc = lib().om_fast_joint(model.h, n, _i(parent), _d(dist), ncols, _b(present), _b(obs), _d(rates),
                             _b(out), _d(lm), nthreads)
    assert rc == 0
    return (out, lm) if want_logmax else out


def fast_marginal(model, parent, dist, obs, present=None, rates=None, nthreads=1, want_post=True):
    """Returns (post [n_nodes][n_cols][K] or None, logl [n_cols])."""
    parent, dist, present, obs = _col_args(parent, dist, present, obs)
    n, ncols = obs.shape
    rates = _arr(rates, np.float64)
    post = np.zeros((n, ncols, model.K)) if want_post else None
    logl = np.zeros(ncols)
    rc = lib().om_fast_marginal(model.h, n, _i(parent), _d(dist), ncols, _b(present), _b(obs), _d(rates),
                                _d(post), _d(logl), nthreads)
    assert rc == 0
    return post, logl


def brute(model, parent, dist, obs, present=None, rate=1.0, max_assign=200_000_000, want_post=True):
    """One column. obs/present: [n_nodes]. Returns dict(state, best_p, post, total_p)."""
    parent, dist, present, obs = _col_args(parent, dist, present, obs)
    n = len(parent)
    st = np.full(n, NONE, dtype=np.uint8)
    post = np.zeros((n, model.K)) if want_post else None
    bp, tp = C.c_double(0), C.c_double(0)
    rc = lib().om_brute(model.h, n, _i(parent), _d(dist), _b(present), _b(obs), float(rate), max_assign,
                        _b(st), C.byref(bp), _d(post), C.byref(tp))
    if rc:
        raise ValueError("brute force too large")
    return dict(state=st, best_p=bp.value, post=post, total_p=tp.value)


def ve_joint(model, parent, dist, obs, present=None, rate=1.0):
    parent, dist, present, obs = _col_args(parent, dist, present, obs)
    n = len(parent)
    st = np.full(n, NONE, dtype=np.uint8)
    lm = C.c_double(0)
    rc = lib().om_ve_joint(model.h, n, _i(parent), _d(dist), _b(present), _b(obs), float(rate), _b(st), C.byref(lm))
    assert rc == 0, rc
    return st, lm.value


def ve_marginal(model, parent, dist, obs, query, present=None, rate=1.0):
    parent, dist, present, obs = _col_args(parent, dist, present, obs)
    post = np.zeros(model.K)
    rc = lib().om_ve_marginal(model.h, len(parent), _i(parent), _d(dist), _b(present), _b(obs), float(rate),
                              int(query), _d(post))
    return None if rc == 1 else post


def ve_loglik(model, parent, dist, obs, present=None, rate=1.0):
    parent, dist, present, obs = _col_args(parent, dist, present, obs)
    out = C.c_double(0)
    rc = lib().om_ve_loglik(model.h, len(parent), _i(parent), _d(dist), _b(present), _b(obs), float(rate),
                            C.byref(out))
    assert rc == 0, rc
    return out.value


class GapModel:
    """bn.ctmc.GapSubstModel over a base Model: K residues + the gap state (oracle/bnk_indel.c)"""

    def __init__(self, base, mu, lam):
        self.base = base
        self.K1 = base.K + 1
        self.mu, self.lam = float(mu), float(lam)
        self.h = lib().om_gap_model_create(base.h, self.mu, self.lam)
        if not self.h:
            raise ValueError("gap model: eigen-decomposition failed or mu + lambda < 0")

    def __del__(self):
        if getattr(self, "h", None):
            lib().om_gap_model_free(self.h)
            self.h = None

    def parts(self):
        K1 = self.K1
        R, ev, iev = np.zeros((K1, K1)), np.zeros((K1, K1)), np.zeros((K1, K1))
        F, ew = np.zeros(K1), np.zeros(K1)
        lib().om_gap_get(self.h, _d(R), _d(F), _d(ev), _d(iev), _d(ew))
        return {"R": R, "F": F, "evec": ev, "ievec": iev, "eval": ew}

    def probs(self, t):
        P = np.zeros((self.K1, self.K1))
        lib().om_gap_getprobs(self.h, float(t), _d(P))
        return P

    def ksi(self, t):
        return lib().om_gap_ksi(self.h, float(t))

    def gamma(self, t):
        return lib().om_gap_gamma(self.h, float(t))


def logsumexp(a):
    a = _arr(a, np.float64)
    return lib().om_logsumexp(_d(a), len(a))


def logm1exp(x):
    return lib().om_logm1exp(float(x))


def indel_column_logl(gm, parent, dist, obs, rate=1.0, geom=0.01, nthreads=1):
    """IndelPeeler.decorate for every column of obs [node][col] (NONE = gap)"""
    parent = _arr(parent, np.int32); dist = _arr(dist, np.float64); obs = _arr(obs, np.uint8)
    out = np.zeros(obs.shape[1])
    rc = lib().om_indel_column_logl(gm.h, len(parent), _i(parent), _d(dist), obs.shape[1], _b(obs), float(rate), float(geom), _d(out), int(nthreads))
    assert rc == 0
    return out


def indel_prob_extra_col(gm, parent, dist, geom):
    parent = _arr(parent, np.int32); dist = _arr(dist, np.float64)
    return lib().om_indel_prob_extra_col(gm.h, len(parent), _i(parent), _d(dist), float(geom))


def indel_aln_logl(gm, parent, dist, obs, geom, nthreads=1):
    """IndelPeeler.calcProbAlnGivenTree"""
    parent = _arr(parent, np.int32); dist = _arr(dist, np.float64); obs = _arr(obs, np.uint8)
    out = C.c_double(0)
    rc = lib().om_indel_aln_logl(gm.h, len(parent), _i(parent), _d(dist), obs.shape[1], _b(obs), float(geom), C.cast(C.byref(out), _dp), int(nthreads))
    assert rc == 0
    return out.value


def indel_optimise_mulambda(base, parent, dist, obs, geom, lo, hi, nthreads=1):
    """IndelPeeler.optimiseMuLambda: (mu = lambda at the optimum, number of distinct likelihood evaluations)"""
    parent = _arr(parent, np.int32); dist = _arr(dist, np.float64); obs = _arr(obs, np.uint8)
    ne = C.c_int(0)
    m = lib().om_indel_optimise_mulambda(base.h, len(parent), _i(parent), _d(dist), obs.shape[1], _b(obs), float(geom), float(lo), float(hi), C.byref(ne), int(nthreads))
    return m, ne.value

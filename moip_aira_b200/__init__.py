"""moip_aira_b200 -- Python host mirror of the C ABI in include/moip_b200.h.

The product is `libmoip_b200.so` (hand-written sm_100a kernels + host C++ driver).  This module only
binds it with ctypes, mirroring the reference's `Problem` / `Solutions` / `solve` / `get_limit`
interfaces (reference src/problem.h, src/solutions.h, src/aira.cpp:83-117) so tests read like the
reference's own.  There is no Python or CPU implementation of any compute path here: if the shared
library is missing the import fails, and compute calls fail when no B200 is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmoip_b200.so")

MIP_OPTIMAL, MIP_INFEASIBLE, MIP_INFORUNBD = 101, 103, 119
LP_CONVERGED, LP_CUTOFF, LP_ITERLIMIT, LP_INFEASIBLE = 0, 1, 2, 3
SENSE_MIN, SENSE_MAX = 0, 1
INFBOUND = 1.0e20
MAX_OBJ = 4


class MoipError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m moip_aira_b200.build` "
        "(nvcc, sm_100a).  There is no fallback implementation.")
# one hardware work queue per worker stream (see csrc/solver.cu, WorkQueueEnv): must be in the environment before the
# process creates its CUDA context, i.e. before torch touches the device
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
_lib = C.CDLL(LIB_PATH)


class ModelInfo(C.Structure):
    _fields_ = [("n", C.c_int), ("ms", C.c_int), ("k", C.c_int), ("nnz", C.c_int), ("sense", C.c_int),
                ("all_binary", C.c_int), ("m", C.c_int), ("mask_words", C.c_int)]


class LpParams(C.Structure):
    _fields_ = [("eps", C.c_double), ("max_iter", C.c_int), ("check_every", C.c_int),
                ("fixed_iters", C.c_int), ("cutoff", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("ip_solved", C.c_int64), ("bb_nodes", C.c_int64), ("node_lps", C.c_int64),
                ("lp_iterations", C.c_int64), ("kernel_launches", C.c_int64), ("cache_queries", C.c_int64),
                ("solver_seconds", C.c_double)]


class KernelTimes(C.Structure):
    _fields_ = [("k1_ms", C.c_double), ("k2_ms", C.c_double), ("k3_ms", C.c_double), ("k4_ms", C.c_double),
                ("copy_ms", C.c_double), ("rounds", C.c_int64), ("scans", C.c_int64)]


class Worker(C.Structure):
    _fields_ = [("id", C.c_int), ("n_obj", C.c_int), ("perm", C.c_int * MAX_OBJ), ("split", C.c_int),
                ("split_start", C.c_double), ("split_stop", C.c_double),
                ("window", C.c_int), ("win_start", C.c_double), ("win_stop", C.c_double)]


SOLVE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double),
                       C.POINTER(C.c_int), C.POINTER(C.c_int))
FIND_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int))
INSERT_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int)
CLAIM_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_pi, _pd = C.POINTER(C.c_int), C.POINTER(C.c_double)
_SIGS = {
    "moip_model_load": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "moip_model_free": (None, [_vp]),
    "moip_model_get_info": (_i, [_vp, C.POINTER(ModelInfo)]),
    "moip_model_objcoef": (_i, [_vp, _i, _pd]),
    "moip_model_dense": (_i, [_vp, _pd, C.c_char_p, _pd, _pd, _pd, C.POINTER(C.c_uint8)]),
    "moip_model_colname": (_i, [_vp, _i, C.c_char_p, _i]),
    "moip_model_selfcheck": (_i, [_vp, _pi, C.c_char_p, _i]),
    "moip_ctx_create": (_i, [_vp, _i, _vp, C.POINTER(_vp)]),
    "moip_ctx_create_own_stream": (_i, [_vp, _i, C.POINTER(_vp)]),
    "moip_device_count": (_i, []),
    "moip_ctx_destroy": (None, [_vp]),
    "moip_lp_default_params": (None, [C.POINTER(LpParams)]),
    "moip_lp_batch_solve": (_i, [_vp, _i, _pi, _pd, C.POINTER(C.c_uint32), C.POINTER(LpParams), _pd, _pd, _pi, _pi, _pd]),
    "moip_lp_batch_upload": (_i, [_vp, _i, _pi, _pd, C.POINTER(C.c_uint32)]),
    "moip_lp_batch_run": (_i, [_vp, C.POINTER(LpParams)]),
    "moip_lp_batch_download": (_i, [_vp, _pd, _pd, _pi, _pi, _pd]),
    "moip_cache_create": (_i, [_vp, C.POINTER(_vp)]),
    "moip_cache_destroy": (None, [_vp]),
    "moip_cache_insert": (_i, [_vp, _pd, _pi, _i]),
    "moip_cache_size": (_i, [_vp]),
    "moip_cache_find_batch": (_i, [_vp, _i, _pd, _i, _pi]),
    "moip_cache_get": (_i, [_vp, _i, _pd, _pi, _pi]),
    "moip_cache_merge": (_i, [_vp, _vp]),
    "moip_cache_sort_unique": (_i, [_vp, _pi, _i]),
    "moip_verify_int64": (_i, [_vp, _i, _pi, _pd, C.POINTER(C.c_int64), C.POINTER(C.c_uint8)]),
    "moip_lex_solve": (_i, [_vp, _pi, _i, _pd, _pi, _pi]),
    "moip_get_limit": (_i, [_vp, _i, _i, _pd, _pi, _pi]),
    "moip_mip_solve": (_i, [_vp, _i, _pd, _pi, _pi, C.POINTER(C.c_int64), _pi]),
    "moip_ctx_stats": (_i, [_vp, C.POINTER(Stats)]),
    "moip_ctx_reset_stats": (_i, [_vp]),
    "moip_ctx_set_kernel_timing": (_i, [_vp, _i]),
    "moip_ctx_set_sync_mode": (_i, [_vp, _i]),
    "moip_ctx_set_ip_node_budget": (_i, [_vp, C.c_longlong]),
    "moip_ctx_kernel_times": (_i, [_vp, C.POINTER(KernelTimes)]),
    "moip_pool_set_kernel_timing": (_i, [_vp, _i]),
    "moip_pool_kernel_times": (_i, [_vp, C.POINTER(KernelTimes)]),
    "moip_pool_set_max_workers": (_i, [_vp, _i]),
    "moip_pool_export_records": (_i, [_vp, _i, _pd, _pi, _pi, _pi]),
    "moip_pool_import_records": (_i, [_vp, _i, _pd, _pi, _pi]),
    "moip_pool_exchange_counts": (_i, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "moip_pool_strips_stolen": (C.c_int64, [_vp]),
    "moip_pool_boxes_postponed": (C.c_int64, [_vp]),
    "moip_optimise": (_i, [_vp, C.POINTER(Worker), _vp, _vp]),
    "moip_optimise_with": (_i, [_i, _i, C.POINTER(Worker), SOLVE_FN, FIND_CB, INSERT_CB, _vp,
                                C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "moip_split_strips": (_i, [_i, _i, _i, _i, _i, _pd]),
    "moip_pareto_front": (_i, [_vp, _i, _i, _i, _pi, _i, _pi]),
    "moip_pool_create": (_i, [_vp, _i, _i, C.POINTER(_vp)]),
    "moip_pool_destroy": (None, [_vp]),
    "moip_pool_workers": (_i, [_vp]),
    "moip_pool_stats": (_i, [_vp, C.POINTER(Stats)]),
    "moip_pool_get_limit": (_i, [_vp, _i, _i, _pd, _pi, _pi]),
    "moip_pool_run_strips": (_i, [_vp, _i, _i, _pd, _pi, _i, _pi]),
    "moip_pool_run_strips_claim": (_i, [_vp, _i, _i, _pd, CLAIM_FN, _vp, _pi, _i, _pi]),
    "moip_pool_run_boxes_claim": (_i, [_vp, _i, _i, _pd, _pd, CLAIM_FN, _vp, _pi, _i, _pi]),
    "moip_pool_pareto_front": (_i, [_vp, _i, _i, _pi, _i, _pi]),
    "moip_coop_workers": (_i, [_i, _i, C.POINTER(Worker)]),
    "moip_coop_optimise_with": (_i, [_i, _i, _i, C.POINTER(Worker), SOLVE_FN, FIND_CB, INSERT_CB, C.POINTER(_vp),
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "moip_pool_synergistic_front": (_i, [_vp, _i, _pi, _i, _pi]),
    "moip_coop_create": (_i, [_i, _i, _i, C.POINTER(_vp)]),
    "moip_coop_destroy": (None, [_vp]),
    "moip_coop_publish": (_i, [_vp, _i, C.c_longlong, _i]),
    "moip_coop_read": (_i, [_vp, _i, C.POINTER(C.c_longlong), _pi]),
    "moip_coop_optimise": (_i, [_vp, C.POINTER(Worker), _vp, _vp, _vp]),
    "moip_coop_optimise_one_with": (_i, [_i, _i, C.POINTER(Worker), _vp, SOLVE_FN, FIND_CB, INSERT_CB, _vp,
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "moip_version": (C.c_char_p, []),
}
EXPORTED = sorted(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    try:
        _fn = getattr(_lib, _name)
    except AttributeError as _e:        # the library does not export the ABI this module binds: stale build
        raise ImportError(f"{LIB_PATH} does not export {_name}: rebuild it with `python moip_aira_b200/build.py --force`") from _e
    _fn.restype = _res
    _fn.argtypes = _args


def _check(rc, what):
    if rc != 0:
        raise MoipError(f"{what} failed with code {rc}")


def _dp(a):
    return a.ctypes.data_as(_pd)


def _ip(a):
    return a.ctypes.data_as(_pi)


class Problem:
    """Mirror of the reference's `Problem` (src/problem.h:11-21): objcnt, objcoef, objsen, rhs."""

    def __init__(self, filename: str):
        self._h = _vp()
        _check(_lib.moip_model_load(os.fsencode(filename), C.byref(self._h)), f"load {filename}")
        info = ModelInfo()
        _check(_lib.moip_model_get_info(self._h, C.byref(info)), "model info")
        self.info = info
        self.filename = filename
        self.n, self.ms, self.objcnt, self.m = info.n, info.ms, info.k, info.m
        self.objsen = info.sense
        self.mask_words = info.mask_words
        self.objcoef = np.zeros((info.k, info.n))
        for j in range(info.k):
            _check(_lib.moip_model_objcoef(self._h, j, _dp(self.objcoef[j])), "objcoef")
        inf = INFBOUND if info.sense == SENSE_MIN else -INFBOUND
        self.rhs = np.full(info.k, inf)          # src/problem.cpp:122-132

    def dense(self):
        n, ms = self.n, self.ms
        A = np.zeros((ms, n))
        sense = C.create_string_buffer(max(ms, 1))
        b = np.zeros(max(ms, 1))
        lb, ub = np.zeros(n), np.zeros(n)
        isint = np.zeros(n, dtype=np.uint8)
        _check(_lib.moip_model_dense(self._h, _dp(A), sense, _dp(b), _dp(lb), _dp(ub),
                                     isint.ctypes.data_as(C.POINTER(C.c_uint8))), "dense")
        return A, [chr(c) for c in sense.raw[:ms]], b[:ms], lb, ub, isint.astype(bool)

    KERNEL_PATHS = {0: "generic", 1: "k1_fast", 2: "k1_reg", 3: "k1_small", 4: "generic-streaming"}

    def selfcheck(self):
        """(kernel path name, error text or None): consistency of the packed device images, no GPU needed"""
        path = C.c_int(-1)
        buf = C.create_string_buffer(256)
        rc = _lib.moip_model_selfcheck(self._h, C.byref(path), buf, 256)
        return self.KERNEL_PATHS.get(path.value, "?"), (None if rc == 0 else buf.value.decode())

    def colnames(self):
        buf = C.create_string_buffer(1024)
        out = []
        for j in range(self.n):
            _check(_lib.moip_model_colname(self._h, j, buf, 1024), "colname")
            out.append(buf.value.decode())
        return out

    def close(self):
        if self._h:
            _lib.moip_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One worker's solver context (the role of `Env` in the reference, src/aira.cpp:541-585)."""

    def __init__(self, problem: Problem, device: int = 0, stream: int | None = None):
        self.problem = problem
        self._h = _vp()
        _check(_lib.moip_ctx_create(problem._h, device, _vp(stream or 0), C.byref(self._h)), "ctx_create")

    # ---- solve() / get_limit() (src/aira.cpp:452-536, :367-450)
    def solve(self, rhs, perm=None, n_obj=None):
        k = self.problem.objcnt
        perm = np.ascontiguousarray(list(perm) if perm is not None else range(k), dtype=np.int32)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        res = np.zeros(k, dtype=np.int32)
        st = C.c_int(0)
        _check(_lib.moip_lex_solve(self._h, _ip(perm), int(n_obj or k), _dp(rhs), _ip(res), C.byref(st)), "lex_solve")
        return st.value, (None if st.value == MIP_INFEASIBLE else res.tolist())

    def get_limit(self, obj, rhs, sense=None):
        k = self.problem.objcnt
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        res = np.zeros(k, dtype=np.int32)
        st = C.c_int(0)
        _check(_lib.moip_get_limit(self._h, int(obj), self.problem.objsen if sense is None else sense, _dp(rhs),
                                   _ip(res), C.byref(st)), "get_limit")
        return st.value, (None if st.value == MIP_INFEASIBLE else res.tolist())

    # ---- K1
    @staticmethod
    def lp_params(**kw):
        p = LpParams()
        _lib.moip_lp_default_params(C.byref(p))
        for a, v in kw.items():
            setattr(p, a, v)
        return p

    def lp_batch_solve(self, cost_idx, rhs, fix_masks=None, params=None, want_x=False):
        pr = self.problem
        cost_idx = np.ascontiguousarray(cost_idx, dtype=np.int32)
        B = len(cost_idx)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64).reshape(B, pr.objcnt)
        masks = None
        if fix_masks is not None:
            masks = np.ascontiguousarray(fix_masks, dtype=np.uint32).reshape(B, pr.mask_words)
        pobj, dbound = np.zeros(B), np.zeros(B)
        status, iters = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        x = np.zeros((B, pr.n)) if want_x else None
        _check(_lib.moip_lp_batch_solve(
            self._h, B, _ip(cost_idx), _dp(rhs),
            masks.ctypes.data_as(C.POINTER(C.c_uint32)) if masks is not None else None,
            C.byref(params) if params is not None else None, _dp(pobj), _dp(dbound), _ip(status), _ip(iters),
            _dp(x) if want_x else None), "lp_batch_solve")
        return {"primal_obj": pobj, "dual_bound": dbound, "status": status, "iters": iters, "x": x}

    def lp_batch_upload(self, cost_idx, rhs, fix_masks=None):
        pr = self.problem
        cost_idx = np.ascontiguousarray(cost_idx, dtype=np.int32)
        B = len(cost_idx)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64).reshape(B, pr.objcnt)
        masks = None
        if fix_masks is not None:
            masks = np.ascontiguousarray(fix_masks, dtype=np.uint32).reshape(B, pr.mask_words)
        self._B = B
        _check(_lib.moip_lp_batch_upload(
            self._h, B, _ip(cost_idx), _dp(rhs),
            masks.ctypes.data_as(C.POINTER(C.c_uint32)) if masks is not None else None), "lp_batch_upload")

    def lp_batch_run(self, params=None):
        _check(_lib.moip_lp_batch_run(self._h, C.byref(params) if params is not None else None), "lp_batch_run")

    def lp_batch_download(self, want_x=False):
        B, pr = self._B, self.problem
        pobj, dbound = np.zeros(B), np.zeros(B)
        status, iters = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        x = np.zeros((B, pr.n)) if want_x else None
        _check(_lib.moip_lp_batch_download(self._h, _dp(pobj), _dp(dbound), _ip(status), _ip(iters),
                                           _dp(x) if want_x else None), "lp_batch_download")
        return {"primal_obj": pobj, "dual_bound": dbound, "status": status, "iters": iters, "x": x}

    # ---- K4
    def verify(self, x, rhs=None):
        pr = self.problem
        x = np.ascontiguousarray(x, dtype=np.int32).reshape(-1, pr.n)
        B = x.shape[0]
        obj = np.zeros((B, pr.objcnt), dtype=np.int64)
        feas = np.zeros(B, dtype=np.uint8)
        r = None
        if rhs is not None:
            r = np.ascontiguousarray(rhs, dtype=np.float64).reshape(B, pr.objcnt)
        _check(_lib.moip_verify_int64(self._h, B, _ip(x), _dp(r) if r is not None else None,
                                      obj.ctypes.data_as(C.POINTER(C.c_int64)),
                                      feas.ctypes.data_as(C.POINTER(C.c_uint8))), "verify_int64")
        return obj, feas.astype(bool)

    # ---- generator
    def optimise(self, worker: Worker, all_sols: "Solutions", infeasibles: "Solutions"):
        _check(_lib.moip_optimise(self._h, C.byref(worker), all_sols._h, infeasibles._h), "optimise")

    def coop_optimise(self, worker: Worker, limits: "CoopLimits", all_sols: "Solutions", infeasibles: "Solutions"):
        """One cooperative worker (owns worker.perm[k-1]) on this context against a limits handle (moip_coop_optimise)."""
        _check(_lib.moip_coop_optimise(self._h, C.byref(worker), limits._h, all_sols._h, infeasibles._h), "coop_optimise")

    def pareto_front(self, split=False, num_threads=1, split_normal=False, cap=1 << 16):
        k = self.problem.objcnt
        rows = np.zeros((cap, k), dtype=np.int32)
        nrows = C.c_int(0)
        _check(_lib.moip_pareto_front(self._h, int(split), int(num_threads), int(split_normal), _ip(rows), cap,
                                      C.byref(nrows)), "pareto_front")
        if nrows.value > cap:
            return self.pareto_front(split, num_threads, split_normal, cap=nrows.value)
        return [tuple(int(v) for v in r) for r in rows[:nrows.value]]

    def stats(self):
        s = Stats()
        _check(_lib.moip_ctx_stats(self._h, C.byref(s)), "stats")
        return {f: getattr(s, f) for f, _ in Stats._fields_}

    def reset_stats(self):
        _check(_lib.moip_ctx_reset_stats(self._h), "reset_stats")

    def set_kernel_timing(self, on=True):
        _check(_lib.moip_ctx_set_kernel_timing(self._h, int(bool(on))), "set_kernel_timing")

    def kernel_times(self):
        """CUDA-event milliseconds per kernel class (the FINETIMING counterpart, reference src/aira.cpp:554-560)"""
        t = KernelTimes()
        _check(_lib.moip_ctx_kernel_times(self._h, C.byref(t)), "kernel_times")
        return {f: getattr(t, f) for f, _ in KernelTimes._fields_}

    def close(self):
        if self._h:
            _lib.moip_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WorkerPool:
    """One solver context per host thread on one GPU (the reference's thread-per-strip model,
    src/aira.cpp:1920-1933): the EPP strips of a level run concurrently and share the device."""

    def __init__(self, problem: Problem, device: int = 0, workers: int = 8):
        self.problem = problem
        h = _vp()
        _check(_lib.moip_pool_create(problem._h, device, int(workers), C.byref(h)), "pool_create")
        self._h = h

    @property
    def workers(self):
        return _lib.moip_pool_workers(self._h)

    def synergistic_front(self, n_workers=None, cap=1 << 16):
        """-t W without --split on the cooperative workers (moip_pool_synergistic_front): W <= k workers, one solver
        context each, exchanging monotone limits on the objective each of them owns."""
        k = self.problem.objcnt
        w = int(n_workers or k)
        while True:
            rows = np.zeros((cap, k), dtype=np.int32)
            n = C.c_int(0)
            _check(_lib.moip_pool_synergistic_front(self._h, w, _ip(rows), cap, C.byref(n)), "pool_synergistic_front")
            if n.value <= cap:
                return [tuple(int(v) for v in r) for r in rows[:n.value]]
            cap = n.value

    def get_limit(self, obj, rhs, sense=None):
        k = self.problem.objcnt
        r = np.ascontiguousarray(rhs, dtype=np.float64)
        res = np.zeros(k, dtype=np.int32)
        st = C.c_int(0)
        _check(_lib.moip_pool_get_limit(self._h, int(obj), self.problem.objsen if sense is None else int(sense), _dp(r), _ip(res),
                                        C.byref(st)), "pool_get_limit")
        return st.value, ([int(v) for v in res] if st.value != MIP_INFEASIBLE else None)

    def run_strips(self, n_obj, strips, claim=None, cap=1 << 20, windows=None):
        """strips: list of (start, stop); returns the feasible result rows found (unsorted).  `claim`: optional
        callable returning the index of the next strip to solve (shared by the pools of several ranks); without it
        the pool works through all the strips given.  `windows`: one (near edge, far edge) pair per entry of `strips` --
        the entries are then boxes, strip x window on objective 1 (moip_worker::window)."""
        k = self.problem.objcnt
        ss = np.ascontiguousarray(np.asarray(strips, dtype=np.float64).reshape(-1, 2))
        rows = np.zeros((cap, k), dtype=np.int32)
        n = C.c_int(0)
        if windows is not None:
            ww = np.ascontiguousarray(np.asarray(windows, dtype=np.float64).reshape(-1, 2))
            if len(ww) != len(ss):
                raise MoipError("one window per strip entry")
            if claim is None:
                it = iter(range(len(ss)))
                lock = __import__("threading").Lock()

                def claim():
                    with lock:
                        return next(it, len(ss))
            cb = CLAIM_FN(lambda _user: int(claim()))
            _check(_lib.moip_pool_run_boxes_claim(self._h, int(n_obj), len(ss), _dp(ss), _dp(ww), cb, None, _ip(rows), cap,
                                                  C.byref(n)), "pool_run_boxes_claim")
        elif claim is None:
            _check(_lib.moip_pool_run_strips(self._h, int(n_obj), len(ss), _dp(ss), _ip(rows), cap, C.byref(n)), "pool_run_strips")
        else:
            cb = CLAIM_FN(lambda _user: int(claim()))
            _check(_lib.moip_pool_run_strips_claim(self._h, int(n_obj), len(ss), _dp(ss), cb, None, _ip(rows), cap, C.byref(n)),
                   "pool_run_strips_claim")
        if n.value > cap:
            raise MoipError("more result rows than the buffer holds")
        return [tuple(int(v) for v in r) for r in rows[:n.value]]

    def pareto_front(self, num_threads=8, split_normal=False, cap=1 << 16):
        k = self.problem.objcnt
        rows = np.zeros((cap, k), dtype=np.int32)
        n = C.c_int(0)
        _check(_lib.moip_pool_pareto_front(self._h, int(num_threads), int(split_normal), _ip(rows), cap, C.byref(n)),
               "pool_pareto_front")
        if n.value > cap:
            return self.pareto_front(num_threads, split_normal, cap=n.value)
        return [tuple(int(v) for v in r) for r in rows[:n.value]]

    def stats(self):
        s = Stats()
        _check(_lib.moip_pool_stats(self._h, C.byref(s)), "pool_stats")
        return {f: getattr(s, f) for f, _ in Stats._fields_}

    def set_kernel_timing(self, on=True):
        _check(_lib.moip_pool_set_kernel_timing(self._h, int(bool(on))), "pool_set_kernel_timing")

    def kernel_times(self):
        t = KernelTimes()
        _check(_lib.moip_pool_kernel_times(self._h, C.byref(t)), "pool_kernel_times")
        return {f: getattr(t, f) for f, _ in KernelTimes._fields_}

    def set_max_workers(self, n):
        """at most n contexts draw strips (0 = all): the share of one rank when several ranks draw from one counter"""
        _check(_lib.moip_pool_set_max_workers(self._h, int(n)), "pool_set_max_workers")

    def export_records(self, cap=1024):
        """cache records this pool produced since the last call while a level's strips are running -> (ip[n][k], result[n][k], infeasible[n])"""
        k = self.problem.objcnt
        ip = np.zeros((cap, k))
        res = np.zeros((cap, k), dtype=np.int32)
        inf = np.zeros(cap, dtype=np.int32)
        n = C.c_int(0)
        _check(_lib.moip_pool_export_records(self._h, int(cap), _dp(ip), _ip(res), _ip(inf), C.byref(n)), "pool_export_records")
        return ip[:n.value], res[:n.value], inf[:n.value]

    def import_records(self, ip, result, infeasible):
        ip = np.ascontiguousarray(ip, dtype=np.float64)
        res = np.ascontiguousarray(result, dtype=np.int32)
        inf = np.ascontiguousarray(infeasible, dtype=np.int32)
        if len(inf):
            _check(_lib.moip_pool_import_records(self._h, len(inf), _dp(ip), _ip(res), _ip(inf)), "pool_import_records")

    def strips_stolen(self):
        return int(_lib.moip_pool_strips_stolen(self._h))

    def boxes_postponed(self):
        return int(_lib.moip_pool_boxes_postponed(self._h))

    def exchange_counts(self):
        a, b = C.c_int64(0), C.c_int64(0)
        _check(_lib.moip_pool_exchange_counts(self._h, C.byref(a), C.byref(b)), "pool_exchange_counts")
        return a.value, b.value

    def close(self):
        if self._h:
            _lib.moip_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Solutions:
    """Mirror of the reference's `Solutions` (src/solutions.h:10-36); `find` runs kernel K3."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.k = ctx.problem.objcnt
        self._h = _vp()
        _check(_lib.moip_cache_create(ctx._h, C.byref(self._h)), "cache_create")

    def insert(self, ip, result, infeasible):
        ip = np.ascontiguousarray(ip, dtype=np.float64)
        res = None if result is None else np.ascontiguousarray(result, dtype=np.int32)
        _check(_lib.moip_cache_insert(self._h, _dp(ip), _ip(res) if res is not None else None, int(bool(infeasible))),
               "cache_insert")

    def __len__(self):
        return _lib.moip_cache_size(self._h)

    def find_batch(self, ips, sense):
        ips = np.ascontiguousarray(ips, dtype=np.float64).reshape(-1, self.k)
        out = np.zeros(len(ips), dtype=np.int32)
        _check(_lib.moip_cache_find_batch(self._h, len(ips), _dp(ips), int(sense), _ip(out)), "cache_find_batch")
        return out

    def find(self, ip, sense):
        """Solutions::find (src/solutions.cpp:11-81): returns (index, infeasible, result) or None."""
        idx = int(self.find_batch([ip], sense)[0])
        return None if idx < 0 else (idx,) + self.get(idx)[1:]

    def get(self, i):
        ip = np.zeros(self.k)
        res = np.zeros(self.k, dtype=np.int32)
        inf = C.c_int(0)
        _check(_lib.moip_cache_get(self._h, int(i), _dp(ip), _ip(res), C.byref(inf)), "cache_get")
        return ip, bool(inf.value), (None if inf.value else res.tolist())

    def merge(self, other: "Solutions"):
        _check(_lib.moip_cache_merge(self._h, other._h), "cache_merge")

    def sort_unique(self):
        n = _lib.moip_cache_sort_unique(self._h, None, 0)
        rows = np.zeros((max(n, 1), self.k), dtype=np.int32)
        n = _lib.moip_cache_sort_unique(self._h, _ip(rows), n)
        return [tuple(int(v) for v in r) for r in rows[:n]]

    def close(self):
        if self._h:
            _lib.moip_cache_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def split_strips(sense, biggest, smallest, num_threads, split_normal=False):
    out = np.zeros(2 * num_threads)
    _check(_lib.moip_split_strips(int(sense), int(biggest), int(smallest), int(num_threads), int(split_normal),
                                  _dp(out)), "split_strips")
    return [(float(out[2 * t]), float(out[2 * t + 1])) for t in range(num_threads)]


def make_worker(k, perm=None, n_obj=None, split=False, split_start=0.0, split_stop=0.0, wid=0, window=None):
    w = Worker()
    if window is not None:
        w.window, w.win_start, w.win_stop = 1, float(window[0]), float(window[1])
    w.id, w.n_obj, w.split = wid, int(n_obj or k), int(split)
    p = list(perm) if perm is not None else list(range(k))
    for i in range(MAX_OBJ):
        w.perm[i] = p[i] if i < len(p) else i
    w.split_start, w.split_stop = float(split_start), float(split_stop)
    return w


def optimise_with(k, sense, worker, solve, find, insert):
    """Host generator state machine with caller-supplied solve/find/insert (host-logic tests)."""
    def _solve(user, perm, n_obj, rhs, result, status):
        st, res = solve([perm[i] for i in range(k)], n_obj, [rhs[i] for i in range(k)])
        status[0] = st
        if res is not None:
            for i in range(k):
                result[i] = res[i]
        return 0

    def _find(user, ip, infeasible, result):
        r = find([ip[i] for i in range(k)])
        if r is None:
            return 0
        inf, res = r
        infeasible[0] = int(inf)
        if not inf:
            for i in range(k):
                result[i] = res[i]
        return 1

    def _insert(user, ip, result, infeasible):
        insert([ip[i] for i in range(k)], None if infeasible else [result[i] for i in range(k)], bool(infeasible))
        return 0

    it, hits = C.c_int64(0), C.c_int64(0)
    _check(_lib.moip_optimise_with(k, int(sense), C.byref(worker), SOLVE_FN(_solve), FIND_CB(_find),
                                   INSERT_CB(_insert), None, C.byref(it), C.byref(hits)), "optimise_with")
    return it.value, hits.value


class CoopLimits:
    """moip_coop handle: the published limits of the cooperative workers (one monotone value per owned objective)."""

    def __init__(self, k, sense, owned):
        self.k = int(k)
        h = _vp()
        mask = 0
        for j in owned:
            mask |= 1 << int(j)
        _check(_lib.moip_coop_create(self.k, int(sense), mask, C.byref(h)), "coop_create")
        self._h = h

    def publish(self, obj, value=0, done=False):
        _check(_lib.moip_coop_publish(self._h, int(obj), int(value), int(bool(done))), "coop_publish")

    def read(self, obj):
        """-> (state, value): state 0 = no limit yet, 1 = value is the limit, 2 = the owner is through"""
        v, st = C.c_longlong(0), C.c_int(0)
        _check(_lib.moip_coop_read(self._h, int(obj), C.byref(v), C.byref(st)), "coop_read")
        return st.value, v.value

    def close(self):
        if self._h:
            _lib.moip_coop_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def coop_optimise_one_with(k, sense, perm, limits, solve, find, insert):
    """ONE cooperative worker (owns perm[-1]) against a CoopLimits handle, caller-supplied solve/find/insert
    (host-logic tests of the multi-process driver).  Returns (solves, skipped)."""
    w = make_worker(k, perm=perm)

    def _solve(user, pm, n_obj, rhs, result, status):
        st, res = solve([pm[i] for i in range(k)], n_obj, [rhs[i] for i in range(k)])
        status[0] = st
        if res is not None:
            for i in range(k):
                result[i] = res[i]
        return 0

    def _find(user, ip, infeasible, result):
        r = find([ip[i] for i in range(k)])
        if r is None:
            return 0
        inf, res = r
        infeasible[0] = int(inf)
        if not inf:
            for i in range(k):
                result[i] = res[i]
        return 1

    def _insert(user, ip, result, infeasible):
        insert([ip[i] for i in range(k)], None if infeasible else [result[i] for i in range(k)], bool(infeasible))
        return 0

    solves, skipped = C.c_int64(0), C.c_int64(0)
    _check(_lib.moip_coop_optimise_one_with(int(k), int(sense), C.byref(w), limits._h, SOLVE_FN(_solve), FIND_CB(_find),
                                            INSERT_CB(_insert), None, C.byref(solves), C.byref(skipped)),
           "coop_optimise_one_with")
    return solves.value, skipped.value


def coop_workers(k, n_workers):
    """The W <= k cooperative workers' permutations (worker i owns the last objective of its permutation)."""
    ws = (Worker * n_workers)()
    _check(_lib.moip_coop_workers(int(k), int(n_workers), ws), "coop_workers")
    return [[ws[i].perm[j] for j in range(k)] for i in range(n_workers)]


def coop_optimise_with(k, sense, n_workers, solve, find, insert):
    """Cooperative workers on host threads with caller-supplied solve/find/insert(worker, ...) (host-logic tests).
    Returns (solves per worker, subproblems skipped per worker)."""
    ws = (Worker * n_workers)()
    _check(_lib.moip_coop_workers(int(k), int(n_workers), ws), "coop_workers")

    def _solve(user, perm, n_obj, rhs, result, status):
        st, res = solve(int(user or 0), [perm[i] for i in range(k)], n_obj, [rhs[i] for i in range(k)])
        status[0] = st
        if res is not None:
            for i in range(k):
                result[i] = res[i]
        return 0

    def _find(user, ip, infeasible, result):
        r = find(int(user or 0), [ip[i] for i in range(k)])
        if r is None:
            return 0
        inf, res = r
        infeasible[0] = int(inf)
        if not inf:
            for i in range(k):
                result[i] = res[i]
        return 1

    def _insert(user, ip, result, infeasible):
        insert(int(user or 0), [ip[i] for i in range(k)], None if infeasible else [result[i] for i in range(k)],
               bool(infeasible))
        return 0

    users = (_vp * n_workers)(*[_vp(i) for i in range(n_workers)])
    solves, skipped = (C.c_int64 * n_workers)(), (C.c_int64 * n_workers)()
    _check(_lib.moip_coop_optimise_with(int(k), int(sense), int(n_workers), ws, SOLVE_FN(_solve), FIND_CB(_find),
                                        INSERT_CB(_insert), users, solves, skipped), "coop_optimise_with")
    return list(solves), list(skipped)


def version():
    return _lib.moip_version().decode()

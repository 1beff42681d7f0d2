"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (dynosam_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class OrcBlock(C.Structure):
    _fields_ = [("type", C.c_int), ("n", C.c_int), ("idx", c_ip), ("meas", c_dp), ("sigma", c_dp),
                ("sigma_dim", C.c_int), ("sigma_bcast", C.c_int), ("robust_k", C.c_double), ("aux_idx", c_ip)]


class OrcProblem(C.Structure):
    _fields_ = [("n_pose", C.c_int), ("n_point", C.c_int), ("n_flow", C.c_int),
                ("pose", c_dp), ("point", c_dp), ("flow", c_dp),
                ("n_aux", C.c_int), ("aux_pose", c_dp), ("calib", C.c_double*6),
                ("n_blocks", C.c_int), ("blocks", C.POINTER(OrcBlock)), ("pose_order", c_ip)]


class OrcLmParams(C.Structure):
    _fields_ = [("lambda_initial", C.c_double), ("lambda_factor", C.c_double), ("lambda_upper", C.c_double),
                ("lambda_lower", C.c_double), ("min_model_fidelity", C.c_double), ("rel_tol", C.c_double),
                ("abs_tol", C.c_double), ("err_tol", C.c_double), ("max_iterations", C.c_int), ("verbose", C.c_int)]


class OrcLmStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("inner_iterations", C.c_int), ("error_initial", C.c_double),
                ("error_final", C.c_double), ("lambda_final", C.c_double), ("bandwidth", C.c_int),
                ("reduced_dim", C.c_int), ("t_linearize", C.c_double), ("t_schur", C.c_double),
                ("t_solve", C.c_double), ("t_backsub", C.c_double), ("t_error", C.c_double), ("t_total", C.c_double)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libdynoba_oracle.so")
    src = os.path.join(_HERE, "dynoba_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_error.restype = C.c_double
        L.orc_error.argtypes = [C.POINTER(OrcProblem)]
        L.orc_dense_dim.argtypes = [C.POINTER(OrcProblem)]
        L.orc_dense_normal.argtypes = [C.POINTER(OrcProblem), c_dp, c_dp]
        L.orc_schur_solve.argtypes = [C.POINTER(OrcProblem), C.c_double, c_dp]
        L.orc_reduced_dense.argtypes = [C.POINTER(OrcProblem), C.c_double, c_dp, c_dp, c_ip]
        L.orc_retract.argtypes = [C.POINTER(OrcProblem), c_dp]
        L.orc_lm_optimize.argtypes = [C.POINTER(OrcProblem), C.POINTER(OrcLmParams), C.POINTER(OrcLmStats)]
        L.orc_lm_default_params.argtypes = [C.POINTER(OrcLmParams)]
        L.orc_linearize_block.argtypes = [C.POINTER(OrcProblem), C.POINTER(OrcBlock), c_dp, c_dp]
        L.orc_error_block.argtypes = [C.POINTER(OrcProblem), C.POINTER(OrcBlock), c_dp]
        L.orc_factor_eval.argtypes = [C.POINTER(OrcProblem), C.POINTER(OrcBlock), C.c_int, c_dp, c_dp]
        for f in ("orc_se3_expmap", "orc_se3_logmap", "orc_se3_inverse"):
            getattr(L, f).argtypes = [c_dp, c_dp]
        L.orc_se3_compose.argtypes = [c_dp, c_dp, c_dp]
        L.orc_se3_retract.argtypes = [c_dp, c_dp, c_dp]
        L.orc_hybrid_project_to_object3.argtypes = [c_dp]*8
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None and a.size else C.cast(None, c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None and a.size else C.cast(None, c_ip)


class OracleProblem:
    """Wraps a dynosam_b200.problem.Problem (duck-typed) as an orc_problem; owns copies of the values."""

    def __init__(self, prob):
        self.src = prob
        self.pose = prob.pose.copy(); self.point = prob.point.copy(); self.flow = prob.flow.copy()
        self._keep = []
        nb = len(prob.blocks)
        self.cblocks = (OrcBlock*max(nb, 1))()
        for i, b in enumerate(prob.blocks):
            cb = self.cblocks[i]
            cb.type, cb.n = b.type, b.n
            cb.idx = _ip(b.idx); cb.meas = _dp(b.meas) if b.meas is not None else C.cast(None, c_dp)
            cb.sigma = _dp(b.sigma); cb.sigma_dim = b.sigma_dim; cb.sigma_bcast = 1 if b.sigma_bcast else 0
            cb.robust_k = float(b.robust_k)
            cb.aux_idx = _ip(b.aux_idx) if b.aux_idx is not None else C.cast(None, c_ip)
        self.c = OrcProblem()
        self.c.n_pose, self.c.n_point, self.c.n_flow = prob.n_pose, prob.n_point, prob.n_flow
        self.c.pose, self.c.point, self.c.flow = _dp(self.pose), _dp(self.point), _dp(self.flow)
        self.c.n_aux = prob.aux_pose.shape[0]; self.c.aux_pose = _dp(prob.aux_pose)
        for k in range(6):
            self.c.calib[k] = float(prob.calib[k])
        self.c.n_blocks = nb; self.c.blocks = self.cblocks
        self.c.pose_order = _ip(prob.pose_order) if prob.pose_order is not None else C.cast(None, c_ip)

    # -- evaluation
    def error(self) -> float:
        return lib().orc_error(C.byref(self.c))

    def linearize_block(self, bi):
        from dynosam_b200.problem import DIM, JCOLS
        b = self.src.blocks[bi]
        A = np.zeros((b.n, DIM[b.type], JCOLS[b.type])); bv = np.zeros((b.n, DIM[b.type]))
        lib().orc_linearize_block(C.byref(self.c), C.byref(self.cblocks[bi]), _dp(A), _dp(bv))
        return A, bv

    def error_block(self, bi):
        e = np.zeros(self.src.blocks[bi].n)
        lib().orc_error_block(C.byref(self.c), C.byref(self.cblocks[bi]), _dp(e))
        return e

    def factor_eval(self, bi, i):
        from dynosam_b200.problem import DIM, JCOLS
        b = self.src.blocks[bi]
        r = np.zeros(DIM[b.type]); J = np.zeros((DIM[b.type], JCOLS[b.type]))
        lib().orc_factor_eval(C.byref(self.c), C.byref(self.cblocks[bi]), i, _dp(r), _dp(J))
        return r, J

    def dense_normal(self):
        n = lib().orc_dense_dim(C.byref(self.c))
        H = np.zeros((n, n)); g = np.zeros(n)
        lib().orc_dense_normal(C.byref(self.c), _dp(H), _dp(g))
        return H, g

    def schur_solve(self, lam):
        n = lib().orc_dense_dim(C.byref(self.c))
        d = np.zeros(n)
        rc = lib().orc_schur_solve(C.byref(self.c), float(lam), _dp(d))
        return rc, d

    def reduced_dense(self, lam):
        n = 6*self.src.n_pose
        S = np.zeros((n, n)); g = np.zeros(n); pos = np.zeros(self.src.n_pose, dtype=np.int32)
        lib().orc_reduced_dense(C.byref(self.c), float(lam), _dp(S), _dp(g), _ip(pos))
        return S, g, pos

    def retract(self, delta):
        delta = np.ascontiguousarray(delta, dtype=np.float64)
        lib().orc_retract(C.byref(self.c), _dp(delta))

    def optimize(self, **kw):
        prm = OrcLmParams(); lib().orc_lm_default_params(C.byref(prm))
        for k, v in kw.items():
            setattr(prm, k, v)
        st = OrcLmStats()
        lib().orc_lm_optimize(C.byref(self.c), C.byref(prm), C.byref(st))
        return {f[0]: getattr(st, f[0]) for f in OrcLmStats._fields_}


def se3_expmap(xi):
    out = np.zeros(12); lib().orc_se3_expmap(_dp(np.ascontiguousarray(xi, dtype=np.float64)), _dp(out)); return out


def se3_logmap(P):
    out = np.zeros(6); lib().orc_se3_logmap(_dp(np.ascontiguousarray(P, dtype=np.float64)), _dp(out)); return out


def se3_compose(a, b):
    out = np.zeros(12)
    lib().orc_se3_compose(_dp(np.ascontiguousarray(a, dtype=np.float64)), _dp(np.ascontiguousarray(b, dtype=np.float64)), _dp(out))
    return out


def se3_inverse(a):
    out = np.zeros(12); lib().orc_se3_inverse(_dp(np.ascontiguousarray(a, dtype=np.float64)), _dp(out)); return out


def se3_retract(P, xi):
    out = np.zeros(12)
    lib().orc_se3_retract(_dp(np.ascontiguousarray(P, dtype=np.float64)), _dp(np.ascontiguousarray(xi, dtype=np.float64)), _dp(out))
    return out


def hybrid_project_to_object3(X, E, L, Z):
    out = np.zeros(3); JX = np.zeros((3, 6)); JE = np.zeros((3, 6)); JL = np.zeros((3, 6))
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (X, E, L, Z)]
    lib().orc_hybrid_project_to_object3(_dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(out), _dp(JX), _dp(JE), _dp(JL))
    return out, JX, JE, JL

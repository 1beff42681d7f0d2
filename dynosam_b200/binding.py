"""ctypes binding of libdynoba.so (include/dynoba.h).  No CPU fallback: a missing library or a
missing sm_100 device raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .problem import ARITY, DIM, JCOLS, MEAS_DIM, Problem

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdynoba.so")

POSE6, POINT3, FLOW2 = 0, 1, 2
OK, ERR_BAD_ARG, ERR_STATE, ERR_CUDA, ERR_INDETERMINATE, ERR_UNSUPPORTED, ERR_COMM = 0, -1, -2, -3, -4, -5, -6

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_u64p = C.POINTER(C.c_uint64)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p)

EXPORTS = [
    "dynoba_version", "dynoba_status_string", "dynoba_last_error", "dynoba_create", "dynoba_destroy",
    "dynoba_lm_default_params", "dynoba_set_variables", "dynoba_set_aux_poses", "dynoba_set_calibration",
    "dynoba_add_factors", "dynoba_set_pose_order", "dynoba_set_shard", "dynoba_set_reduce", "dynoba_set_partition", "dynoba_plan_partition", "dynoba_set_tuning", "dynoba_fp64_rate", "dynoba_add_linear_prior", "dynoba_marginal", "dynoba_finalize", "dynoba_error",
    "dynoba_optimize", "dynoba_get_variables", "dynoba_get_keys", "dynoba_num_variables", "dynoba_problem_info",
    "dynoba_linearize", "dynoba_linearize_block", "dynoba_get_linearization", "dynoba_get_factor_errors", "dynoba_solve",
    "dynoba_get_reduced_system", "dynoba_retract", "dynoba_flow_pose_default_params", "dynoba_flow_pose_batch",
    "dynoba_motion_refine_default_params", "dynoba_motion_refine_batch", "dynoba_batch_release",
]


class LmParams(C.Structure):
    _fields_ = [("lambda_initial", C.c_double), ("lambda_factor", C.c_double), ("lambda_upper_bound", C.c_double),
                ("lambda_lower_bound", C.c_double), ("min_model_fidelity", C.c_double),
                ("relative_error_tol", C.c_double), ("absolute_error_tol", C.c_double), ("error_tol", C.c_double),
                ("max_iterations", C.c_int32), ("verbosity", C.c_int32)]


class LmStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("inner_iterations", C.c_int32), ("error_initial", C.c_double),
                ("error_final", C.c_double), ("lambda_final", C.c_double), ("reduced_dim", C.c_int32),
                ("bandwidth", C.c_int32), ("kernel_launches", C.c_int64), ("ms_linearize", C.c_double),
                ("ms_schur", C.c_double), ("ms_factor", C.c_double), ("ms_backsub", C.c_double),
                ("ms_error", C.c_double), ("ms_total", C.c_double)]

    def as_dict(self):
        return {f[0]: getattr(self, f[0]) for f in self._fields_}


class FlowPoseParams(C.Structure):
    _fields_ = [("flow_sigma", C.c_double), ("flow_prior_sigma", C.c_double), ("huber_k", C.c_double), ("outlier_rounds", C.c_int32),
                ("outlier_threshold", C.c_double), ("lm", LmParams)]


class MotionRefineParams(C.Structure):
    _fields_ = [("landmark_motion_sigma", C.c_double), ("projection_sigma", C.c_double), ("huber_k", C.c_double),
                ("pose_prior_sigma", C.c_double), ("lm", LmParams)]


class DynobaError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libdynoba status {status}: {msg}")
        self.status = status


_LIB = None


def load():
    """dlopen libdynoba.so; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build the CUDA library first (__graft_entry__.build()); "
                              "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.dynoba_status_string.restype = C.c_char_p
        L.dynoba_last_error.restype = C.c_char_p
        L.dynoba_last_error.argtypes = [C.c_void_p]
        L.dynoba_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.dynoba_destroy.argtypes = [C.c_void_p]
        L.dynoba_lm_default_params.argtypes = [C.POINTER(LmParams)]
        L.dynoba_set_variables.argtypes = [C.c_void_p, C.c_int, C.c_int64, c_u64p, c_dp]
        L.dynoba_set_aux_poses.argtypes = [C.c_void_p, C.c_int64, c_dp]
        L.dynoba_set_calibration.argtypes = [C.c_void_p, c_dp]
        L.dynoba_add_factors.argtypes = [C.c_void_p, C.c_int, C.c_int64, c_ip, c_dp, c_dp, C.c_int, C.c_int64, C.c_double, c_ip]
        L.dynoba_set_pose_order.argtypes = [C.c_void_p, C.c_int64, c_ip]
        L.dynoba_set_shard.argtypes = [C.c_void_p, C.c_int, C.c_int, ALLREDUCE_FN, C.c_void_p, C.c_int]
        L.dynoba_set_reduce.argtypes = [C.c_void_p, REDUCE_FN, C.c_void_p]
        L.dynoba_set_partition.argtypes = [C.c_void_p, C.c_int]
        L.dynoba_plan_partition.argtypes = [C.c_int32, C.c_int32, C.c_int32, c_ip]
        L.dynoba_set_tuning.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.dynoba_fp64_rate.argtypes = [C.c_void_p, c_dp]
        L.dynoba_add_linear_prior.argtypes = [C.c_void_p, C.c_int32, c_ip, c_dp, c_dp, c_dp, C.c_double]
        L.dynoba_marginal.argtypes = [C.c_void_p, C.c_int32, c_ip, c_dp, c_dp]
        L.dynoba_finalize.argtypes = [C.c_void_p]
        L.dynoba_error.argtypes = [C.c_void_p, c_dp]
        L.dynoba_optimize.argtypes = [C.c_void_p, C.POINTER(LmParams), C.POINTER(LmStats)]
        L.dynoba_get_variables.argtypes = [C.c_void_p, C.c_int, C.c_int64, c_dp]
        L.dynoba_get_keys.argtypes = [C.c_void_p, C.c_int, C.c_int64, c_u64p]
        L.dynoba_num_variables.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.dynoba_problem_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.dynoba_linearize.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.dynoba_linearize_block.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
        L.dynoba_get_linearization.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp]
        L.dynoba_get_factor_errors.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.dynoba_solve.argtypes = [C.c_void_p, C.c_double, c_dp]
        L.dynoba_get_reduced_system.argtypes = [C.c_void_p, C.c_double, c_dp, c_dp]
        L.dynoba_retract.argtypes = [C.c_void_p, c_dp]
        c_u8p = C.POINTER(C.c_uint8)
        L.dynoba_flow_pose_default_params.argtypes = [C.POINTER(FlowPoseParams)]
        L.dynoba_flow_pose_batch.argtypes = [C.c_int, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.POINTER(FlowPoseParams),
                                             c_dp, c_dp, c_u8p, c_dp, c_dp, c_ip, c_ip, c_ip]
        L.dynoba_motion_refine_default_params.argtypes = [C.POINTER(MotionRefineParams)]
        L.dynoba_motion_refine_batch.argtypes = [C.c_int, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.POINTER(MotionRefineParams),
                                                 c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip, c_ip]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None and a.size else C.cast(None, c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None and a.size else C.cast(None, c_ip)


def plan_partition(n_poses, bandwidth, world):
    """first pose position (elimination order) of every rank's cells, [world + 1] (pure host computation, no GPU needed)"""
    out = np.zeros(world + 1, np.int32)
    st = load().dynoba_plan_partition(int(n_poses), int(bandwidth), int(world), _ip(out))
    if st != OK:
        raise DynobaError(st, "dynoba_plan_partition")
    return out


def default_params(**kw) -> LmParams:
    p = LmParams()
    load().dynoba_lm_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Solver:
    """One dynoba handle == one optimiser instance (not thread-safe, like the reference's call site)."""

    def __init__(self, problem: Problem | None = None, device: int = 0):
        self.lib = load()
        self.h = C.c_void_p()
        st = self.lib.dynoba_create(device, C.byref(self.h))
        if st != OK:
            raise DynobaError(st, self.lib.dynoba_status_string(st).decode() + " (no sm_100 CUDA device: libdynoba has no CPU path)")
        self.problem = None
        self._cb = None
        if problem is not None:
            self.ingest(problem)

    def _ck(self, st):
        if st != OK:
            raise DynobaError(st, f"{self.lib.dynoba_status_string(st).decode()}: {self.lib.dynoba_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.lib.dynoba_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- ingest
    def ingest(self, p: Problem):
        self.problem = p
        L = self.lib
        kp = p.pose_keys.ctypes.data_as(c_u64p) if p.pose_keys is not None else C.cast(None, c_u64p)
        kq = p.point_keys.ctypes.data_as(c_u64p) if p.point_keys is not None else C.cast(None, c_u64p)
        self._ck(L.dynoba_set_variables(self.h, POSE6, p.n_pose, kp, _dp(p.pose)))
        self._ck(L.dynoba_set_variables(self.h, POINT3, p.n_point, kq, _dp(p.point)))
        if p.n_flow:
            self._ck(L.dynoba_set_variables(self.h, FLOW2, p.n_flow, C.cast(None, c_u64p), _dp(p.flow)))
        if p.aux_pose.shape[0]:
            self._ck(L.dynoba_set_aux_poses(self.h, p.aux_pose.shape[0], _dp(p.aux_pose)))
        self._ck(L.dynoba_set_calibration(self.h, _dp(p.calib)))
        for b in p.blocks:
            self.add_factors(b)
        if p.pose_order is not None:
            self._ck(L.dynoba_set_pose_order(self.h, p.n_pose, _ip(p.pose_order)))
        for pr in getattr(p, "linear_priors", []):
            self.add_linear_prior(pr["idx"], pr["lin"], pr["G"], pr["g"], pr.get("f", 0.0))

    def add_factors(self, b):
        self._ck(self.lib.dynoba_add_factors(self.h, b.type, b.n, _ip(b.idx), _dp(b.meas) if b.meas is not None else C.cast(None, c_dp),
                                             _dp(b.sigma), b.sigma_dim, 1 if b.sigma_bcast else b.n, float(b.robust_k),
                                             _ip(b.aux_idx) if b.aux_idx is not None else C.cast(None, c_ip)))

    def add_linear_prior(self, idx, lin, G, g, f=0.0):
        """gtsam::LinearContainerFactor(HessianFactor): error = 1/2 d^T G d - g^T d + 1/2 f, d = localCoordinates(lin, x)"""
        idx = np.ascontiguousarray(idx, dtype=np.int32); lin = np.ascontiguousarray(lin, dtype=np.float64).reshape(-1, 12)
        G = np.ascontiguousarray(G, dtype=np.float64); g = np.ascontiguousarray(g, dtype=np.float64)
        assert G.shape == (6*idx.size, 6*idx.size) and g.shape == (6*idx.size,)
        self._ck(self.lib.dynoba_add_linear_prior(self.h, idx.size, _ip(idx), _dp(lin), _dp(G), _dp(g), float(f)))

    def marginal(self, keep_idx):
        """(G, g): marginal information of the pose-like variables keep_idx (the last ones in frame order) at the current values"""
        keep = np.ascontiguousarray(keep_idx, dtype=np.int32); k = keep.size
        G = np.zeros((6*k, 6*k)); g = np.zeros(6*k)
        self._ck(self.lib.dynoba_marginal(self.h, k, _ip(keep), _dp(G), _dp(g)))
        return G, g

    def set_shard(self, rank, world, allreduce, min_bandwidth=0):
        """allreduce(dev_ptr:int, n:int, stream:int) -> None sums n doubles in place across ranks."""
        def _cb(ctx, dev, n, stream):
            try:
                allreduce(dev, n, stream)
                return 0
            except Exception as e:  # pragma: no cover
                print("all-reduce callback failed:", e)
                return 1
        self._cb = ALLREDUCE_FN(_cb)
        self._ck(self.lib.dynoba_set_shard(self.h, rank, world, self._cb, None, min_bandwidth))

    def set_reduce(self, reduce):
        """reduce(dev_ptr:int, n:int, root:int, stream:int) -> None sums n doubles at rank `root` (distributed reduced solve)."""
        def _cb(ctx, dev, n, root, stream):
            try:
                reduce(dev, n, root, stream)
                return 0
            except Exception as e:  # pragma: no cover
                print("reduce callback failed:", e)
                return 1
        self._cb_red = REDUCE_FN(_cb)
        self._ck(self.lib.dynoba_set_reduce(self.h, self._cb_red, None))

    def set_partition(self, ncells):
        """number of cells of the reduced solve (0 automatic, -1 one plain band factorisation)"""
        self._ck(self.lib.dynoba_set_partition(self.h, int(ncells)))

    def set_tuning(self, name, value):
        self._ck(self.lib.dynoba_set_tuning(self.h, name.encode(), float(value)))

    def fp64_rate(self) -> float:
        out = C.c_double()
        self._ck(self.lib.dynoba_fp64_rate(self.h, C.byref(out)))
        return out.value

    def reset_values(self):
        """Re-upload the initial values of the ingested problem (device layout is kept)."""
        p = self.problem
        self._ck(self.lib.dynoba_set_variables(self.h, POSE6, p.n_pose, C.cast(None, c_u64p), _dp(p.pose)))
        if p.n_point:
            self._ck(self.lib.dynoba_set_variables(self.h, POINT3, p.n_point, C.cast(None, c_u64p), _dp(p.point)))
        if p.n_flow:
            self._ck(self.lib.dynoba_set_variables(self.h, FLOW2, p.n_flow, C.cast(None, c_u64p), _dp(p.flow)))

    def finalize(self):
        self._ck(self.lib.dynoba_finalize(self.h))

    # ---- compute
    def error(self) -> float:
        out = C.c_double()
        self._ck(self.lib.dynoba_error(self.h, C.byref(out)))
        return out.value

    def optimize(self, params: LmParams | None = None, **kw) -> dict:
        prm = params if params is not None else default_params(**kw)
        st = LmStats()
        self._ck(self.lib.dynoba_optimize(self.h, C.byref(prm), C.byref(st)))
        return st.as_dict()

    def linearize(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.dynoba_linearize(self.h, C.byref(ms)))
        return ms.value

    def linearize_block(self, bi):
        """(ms, algorithmic bytes) of one launch of block bi's Jacobian-build kernel."""
        ms = C.c_float(); nb = C.c_int64()
        self._ck(self.lib.dynoba_linearize_block(self.h, bi, C.byref(ms), C.byref(nb)))
        return ms.value, nb.value

    def linearization(self, bi):
        b = self.problem.blocks[bi]
        A = np.zeros((b.n, DIM[b.type], JCOLS[b.type])); bv = np.zeros((b.n, DIM[b.type]))
        self._ck(self.lib.dynoba_get_linearization(self.h, bi, _dp(A), _dp(bv)))
        return A, bv

    def factor_errors(self, bi):
        e = np.zeros(self.problem.blocks[bi].n)
        self._ck(self.lib.dynoba_get_factor_errors(self.h, bi, _dp(e)))
        return e

    def solve(self, lam):
        p = self.problem
        d = np.zeros(6*p.n_pose + 3*p.n_point + 2*p.n_flow)
        self._ck(self.lib.dynoba_solve(self.h, float(lam), _dp(d)))
        return d

    def reduced_system(self, lam):
        n = 6*self.problem.n_pose
        S = np.zeros((n, n)); g = np.zeros(n)
        self._ck(self.lib.dynoba_get_reduced_system(self.h, float(lam), _dp(S), _dp(g)))
        return S, g

    def retract(self, delta):
        delta = np.ascontiguousarray(delta, dtype=np.float64)
        self._ck(self.lib.dynoba_retract(self.h, _dp(delta)))

    def info(self):
        n = C.c_int32(); bw = C.c_int32(); jb = C.c_int64()
        self._ck(self.lib.dynoba_problem_info(self.h, C.byref(n), C.byref(bw), C.byref(jb)))
        return dict(reduced_dim=n.value, bandwidth=bw.value, jacobian_bytes=jb.value)

    def values(self):
        p = self.problem
        pose = np.zeros((p.n_pose, 12)); point = np.zeros((p.n_point, 3)); flow = np.zeros((p.n_flow, 2))
        self._ck(self.lib.dynoba_get_variables(self.h, POSE6, p.n_pose, _dp(pose)))
        if p.n_point:
            self._ck(self.lib.dynoba_get_variables(self.h, POINT3, p.n_point, _dp(point)))
        if p.n_flow:
            self._ck(self.lib.dynoba_get_variables(self.h, FLOW2, p.n_flow, _dp(flow)))
        return pose, point, flow


def _cat(problems, key, width):
    return np.ascontiguousarray(np.concatenate([np.asarray(q[key], dtype=np.float64).reshape(-1, width) for q in problems], 0))


def _set_params(prm, kw):
    for k, v in kw.items():
        if hasattr(prm, k):
            setattr(prm, k, v)
        elif hasattr(prm.lm, k):
            setattr(prm.lm, k, v)
        else:
            raise TypeError(f"unknown parameter {k}")


def flow_pose_batch(problems, device: int = 0, **kw):
    """All joint optical-flow + pose refinements of a frame in ONE launch (dynoba_flow_pose_batch; the reference runs
    OpticalFlowAndPoseOptimizer::optimize per object, MotionSolver-inl.hpp:88-278).  problems: dicts with pose_init[12],
    pose_prev[12], calib[5], kp_prev[n,2], depth[n], flow[n,2].  Keyword arguments set dynoba_flow_pose_params / its lm
    member (flow_sigma, flow_prior_sigma, huber_k, outlier_rounds, outlier_threshold, max_iterations, ...).
    Returns one dict per problem."""
    L = load()
    npb = len(problems)
    if npb == 0:
        return []
    cnt = [len(np.asarray(q["depth"]).reshape(-1)) for q in problems]
    off = np.zeros(npb + 1, dtype=np.int32); off[1:] = np.cumsum(cnt)
    pose0, prev, cal = _cat(problems, "pose_init", 12), _cat(problems, "pose_prev", 12), _cat(problems, "calib", 5)
    kp, depth, flow = _cat(problems, "kp_prev", 2), _cat(problems, "depth", 1), _cat(problems, "flow", 2)
    total = int(off[-1])
    pose_out = np.zeros((npb, 12)); flow_out = np.zeros((max(total, 1), 2)); inl = np.zeros(max(total, 1), dtype=np.uint8)
    e0 = np.zeros(npb); e1 = np.zeros(npb)
    it = np.zeros(npb, dtype=np.int32); inner = np.zeros(npb, dtype=np.int32); rounds = np.zeros(npb, dtype=np.int32)
    prm = FlowPoseParams(); L.dynoba_flow_pose_default_params(C.byref(prm)); _set_params(prm, kw)
    rc = L.dynoba_flow_pose_batch(device, npb, _ip(off), _dp(pose0), _dp(prev), _dp(cal), _dp(kp), _dp(depth), _dp(flow), C.byref(prm),
                                  _dp(pose_out), _dp(flow_out), inl.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(e0), _dp(e1), _ip(it), _ip(inner), _ip(rounds))
    if rc != 0:
        raise DynobaError(rc, L.dynoba_status_string(rc).decode())
    return [dict(pose=pose_out[i].copy(), flow=flow_out[off[i]:off[i + 1]].copy(), inlier=inl[off[i]:off[i + 1]].astype(bool),
                 error_initial=float(e0[i]), error_final=float(e1[i]), iterations=int(it[i]), inner_iterations=int(inner[i]), rounds=int(rounds[i]))
            for i in range(npb)]


def motion_refine_batch(problems, device: int = 0, **kw):
    """All object-motion refinements of a frame in ONE launch (dynoba_motion_refine_batch; the reference runs
    MotionOnlyRefinementOptimizer::optimize per object, MotionSolver-inl.hpp:291-470).  problems: dicts with pose_prev[12],
    pose_cur[12], motion_init[12], calib[5], kp_prev[n,2], kp_cur[n,2], points_init[n,6] (m_k-1 | m_k, world frame)."""
    L = load()
    npb = len(problems)
    if npb == 0:
        return []
    cnt = [np.asarray(q["kp_prev"]).reshape(-1, 2).shape[0] for q in problems]
    off = np.zeros(npb + 1, dtype=np.int32); off[1:] = np.cumsum(cnt)
    pa, pb, h, cal = _cat(problems, "pose_prev", 12), _cat(problems, "pose_cur", 12), _cat(problems, "motion_init", 12), _cat(problems, "calib", 5)
    ka, kb, pts = _cat(problems, "kp_prev", 2), _cat(problems, "kp_cur", 2), _cat(problems, "points_init", 6)
    total = int(off[-1])
    mo = np.zeros((npb, 12)); po = np.zeros((npb, 24)); pto = np.zeros((max(total, 1), 6)); me = np.zeros(max(total, 1)); e0 = np.zeros(npb); e1 = np.zeros(npb)
    it = np.zeros(npb, dtype=np.int32); inner = np.zeros(npb, dtype=np.int32)
    prm = MotionRefineParams(); L.dynoba_motion_refine_default_params(C.byref(prm)); _set_params(prm, kw)
    rc = L.dynoba_motion_refine_batch(device, npb, _ip(off), _dp(pa), _dp(pb), _dp(h), _dp(cal), _dp(ka), _dp(kb), _dp(pts), C.byref(prm),
                                      _dp(mo), _dp(po), _dp(pto), _dp(me), _dp(e0), _dp(e1), _ip(it), _ip(inner))
    if rc != 0:
        raise DynobaError(rc, L.dynoba_status_string(rc).decode())
    return [dict(motion=mo[i].copy(), poses=po[i].reshape(2, 12).copy(), points=pto[off[i]:off[i + 1]].copy(), motion_factor_error=me[off[i]:off[i + 1]].copy(),
                 error_initial=float(e0[i]), error_final=float(e1[i]), iterations=int(it[i]), inner_iterations=int(inner[i])) for i in range(npb)]

"""the sliding-window caller of the same LM surface (dynosam_opt/src/SlidingWindowOptimization.cc:67-190):
a dense Hessian prior (gtsam::LinearContainerFactor) as one more factor, and the Schur marginal of the most recent
pose-like variables as an output.

Oracle side: numpy on top of the CPU oracle's primitives -- dense normal equations / reduced system of the graph
(oracle/dynoba_oracle.c) plus the literal LinearContainerFactor algebra [GTSAM-ext LinearContainerFactor::{error,linearize}]:
    error(x) = 1/2 d^T G d - g^T d + 1/2 f,  d = linPoint.localCoordinates(x);   linearize -> HessianFactor(G, g - G d)
"""
import numpy as np
import pytest

from dynosam_b200 import lie, synth
from dynosam_b200.problem import FactorBlock, Problem, SLOT_CLASS

pytestmark = pytest.mark.gpu


def _solver(p):
    from dynosam_b200.binding import Solver
    return Solver(p)


def _oracle(p):
    from oracle import oracle as O
    return O.OracleProblem(p)


def _local(lin, x):
    """Pose3 localCoordinates, per pose: Logmap(lin^-1 x) through the oracle's SE(3) maps"""
    from oracle import oracle as O
    return np.concatenate([O.se3_logmap(O.se3_compose(O.se3_inverse(lin[i]), x[i])) for i in range(lin.shape[0])])


def _random_prior(rng, p, idx):
    n = len(idx)
    A = rng.normal(0, 1, (6*n + 4, 6*n)); G = A.T @ A*30.0 + 5.0*np.eye(6*n)
    lin = lie.retract(p.pose[idx], rng.normal(0, 0.02, (n, 6)))
    g = rng.normal(0, 3.0, 6*n)
    return dict(idx=np.asarray(idx, np.int32), lin=lin, G=G, g=g, f=0.7)


def test_linear_prior_error_and_damped_solve():
    import dataclasses
    rng = np.random.default_rng(1)
    base = synth.make_config("C1")
    pr = _random_prior(rng, base, [15, 16, 17, 18, 19])                   # the five most recent camera poses
    p = dataclasses.replace(base, linear_priors=[pr])
    s = _solver(p); o = _oracle(base)
    d0 = _local(pr["lin"], base.pose[pr["idx"]])
    e_prior = 0.5*d0 @ pr["G"] @ d0 - pr["g"] @ d0 + 0.5*pr["f"]
    assert abs(s.error() - (o.error() + e_prior)) <= 1e-9*abs(o.error() + e_prior)
    # damped step against the dense normal equations + the relinearised Hessian factor
    H, g = o.dense_normal()
    cols = np.concatenate([6*i + np.arange(6) for i in pr["idx"]])
    H[np.ix_(cols, cols)] += pr["G"]; g[cols] += pr["g"] - pr["G"] @ d0
    lam = 1e-3
    ref = np.linalg.solve(H + lam*np.eye(H.shape[0]), g)
    d = s.solve(lam)
    assert np.linalg.norm(d - ref) <= 1e-6*np.linalg.norm(ref)
    # the LM run converges and the prior pulls: the final error counts the prior
    st = s.optimize(max_iterations=8)
    pose, point, _ = s.values()
    o2 = _oracle(dataclasses.replace(base, pose=pose, point=point))
    d1 = _local(pr["lin"], pose[pr["idx"]])
    e1 = o2.error() + 0.5*d1 @ pr["G"] @ d1 - pr["g"] @ d1 + 0.5*pr["f"]
    assert abs(st["error_final"] - e1) <= 1e-8*abs(e1) and st["error_final"] < st["error_initial"]
    s.close()


def test_marginal_equals_dense_schur_complement():
    p = synth.make_problem(n_frames=60, n_objects=2, n_static=900, n_dynamic=300, seed=8, max_static_age=5, max_dynamic_age=5, object_span=(30, 40))
    s = _solver(p); o = _oracle(p)
    S, g, pos = o.reduced_dense(0.0)                                    # solver order, every landmark eliminated
    order = np.argsort(pos)                                             # order[k] = user index of the pose at solver position k
    for k in (3, 9):
        keep = order[-k:][::-1].copy()                                  # the last k variables, listed in another order on purpose
        G, gm = s.marginal(keep)
        n = S.shape[0]; cut = n - 6*k
        Srr, Skr, Skk = S[:cut, :cut], S[cut:, :cut], S[cut:, cut:]
        X = np.linalg.solve(Srr, np.concatenate([Skr.T, g[:cut, None]], 1))
        Gk = Skk - Skr @ X[:, :-1]; gk = g[cut:] - Skr @ X[:, -1]
        # to the caller's order of `keep`
        perm = np.concatenate([6*(pos[i] - (len(pos) - k)) + np.arange(6) for i in keep])
        assert np.abs(G - Gk[np.ix_(perm, perm)]).max() <= 1e-7*np.abs(Gk).max(), k
        assert np.abs(gm - gk[perm]).max() <= 1e-6*max(np.abs(gk).max(), 1e-12), k
    # not the most recent variables -> a status, not garbage
    from dynosam_b200.binding import DynobaError, ERR_UNSUPPORTED
    with pytest.raises(DynobaError) as ei:
        s.marginal(order[:3])
    assert ei.value.status == ERR_UNSUPPORTED
    s.close()


def _split_by_time(p, f_cut, k):
    """Window split of SlidingWindowOptimization: graph A = every factor of the landmarks last seen before frame f_cut and
    the pose-only factors that touch a pose older than f_cut - k; graph B = the rest.  The poses of frames
    [f_cut - k, f_cut) are shared (the kept block of A, the oldest block of B)."""
    frame = p.pose_order
    old_pose = frame < f_cut - k
    npt = p.n_point
    last = np.full(npt, -1)
    for b in p.blocks:
        cls = SLOT_CLASS[b.type]; ls = [i for i, c in enumerate(cls) if c == 1]; ps = [i for i, c in enumerate(cls) if c == 0]
        if ls:
            np.maximum.at(last, b.idx[:, ls[0]], frame[b.idx[:, ps]].max(1))
    in_a_pt = last < f_cut
    A, B = [], []
    for b in p.blocks:
        cls = SLOT_CLASS[b.type]; ls = [i for i, c in enumerate(cls) if c == 1]; ps = [i for i, c in enumerate(cls) if c == 0]
        sel = in_a_pt[b.idx[:, ls[0]]] if ls else old_pose[b.idx[:, ps]].any(1)
        for dst, m in ((A, sel), (B, ~sel)):
            if m.any():
                dst.append(FactorBlock(b.type, b.idx[m], None if b.meas is None else b.meas[m], b.sigma if b.sigma_bcast else b.sigma[m],
                                       b.robust_k, None if b.aux_idx is None else b.aux_idx[m]))
    return A, B, in_a_pt


def _subproblem(p, blocks, keep_pose, keep_pt):
    """the variables touched by `blocks` (masks keep_pose / keep_pt), re-indexed"""
    pi = np.cumsum(keep_pose) - 1; qi = np.cumsum(keep_pt) - 1
    out = []
    for b in blocks:
        idx = b.idx.copy()
        for s_, c in enumerate(SLOT_CLASS[b.type]):
            idx[:, s_] = pi[idx[:, s_]] if c == 0 else qi[idx[:, s_]]
        out.append(FactorBlock(b.type, idx, b.meas, b.sigma, b.robust_k, b.aux_idx))
    return Problem(p.pose[keep_pose], p.point[keep_pt], aux_pose=p.aux_pose, calib=p.calib, blocks=out, pose_order=p.pose_order[keep_pose])


def test_two_windows_with_marginal_prior_equal_the_joint_solve():
    """Marginalisation consistency, the property the sliding window rests on: the Gauss-Newton step of the joint graph,
    restricted to the second window's variables, equals the step of the second window alone once the first window has
    been replaced by its marginal on the shared poses (prior linearised at the current values)."""
    p = synth.make_problem(n_frames=48, n_objects=0, n_static=1200, n_dynamic=0, seed=12, max_static_age=4)
    f_cut, k = 24, 5                                                     # k >= max track age: no landmark of B reaches behind the kept block
    A, B, in_a_pt = _split_by_time(p, f_cut, k)
    frame = p.pose_order
    pose_a = frame < f_cut; pose_b = frame >= f_cut - k
    pa = _subproblem(p, A, pose_a, in_a_pt)
    pb = _subproblem(p, B, pose_b, ~in_a_pt)
    assert pa.n_factors + pb.n_factors == p.n_factors
    # window 1: marginal of its last k poses (user indices inside pa: the poses of frames [f_cut - k, f_cut))
    sa = _solver(pa)
    keep_a = np.flatnonzero(pa.pose_order >= f_cut - k)
    G, g = sa.marginal(keep_a)
    sa.close()
    # window 2 = graph B + the prior on its first k poses, linearised at the current values (d = 0)
    import dataclasses
    keep_b = np.flatnonzero(pb.pose_order < f_cut)
    assert np.array_equal(pa.pose[keep_a], pb.pose[keep_b])
    pb2 = dataclasses.replace(pb, linear_priors=[dict(idx=keep_b.astype(np.int32), lin=pb.pose[keep_b], G=G, g=g, f=0.0)])
    lam = 1e-9                                                           # (damping on eliminated variables is the only thing the identity does not cover)
    sj = _solver(p); dj = sj.solve(lam); sj.close()
    sb = _solver(pb2); db = sb.solve(lam); sb.close()
    # joint step restricted to window-2 variables, in pb's variable order
    pj = dj[:6*p.n_pose].reshape(-1, 6)[pose_b].reshape(-1); qj = dj[6*p.n_pose:].reshape(-1, 3)[~in_a_pt].reshape(-1)
    ref = np.concatenate([pj, qj])
    assert np.linalg.norm(db - ref) <= 1e-5*np.linalg.norm(ref), np.linalg.norm(db - ref)/np.linalg.norm(ref)

"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md 8c).  Fixtures are re-typed from the reference's gtest sources; tolerances are theirs.

  dynosam/test/test_factors.cc:92-132     Pose3FlowProjectionFactor.testJacobians        (1e-4)
  dynosam/test/test_factors.cc:134-196    LandmarkMotionTernaryFactor.{testJacobians,testZeroError} (1e-9 / 1e-4)
  dynosam/test/test_factors.cc:198-276    SmartMotionFactor zero-error / noise tests      (1e-5)
  dynosam/test/test_factors.cc:278-450    SmartMotionFactor.testBasicSchurCompliment      (1e-5)
  dynosam/test/test_factors.cc:462-556    SmartMotionFactor.testSimpleOptimise (LM recovers GT, 1e-5)
  dynosam/test/test_hybrid_motion.cc:45-343  HybridMotion / StereoHybrid fixtures         (1e-5 / 1e-9)
"""
import numpy as np
import pytest

from dynosam_b200 import lie
from dynosam_b200.problem import (BETWEEN6, FLOWPROJ2, HYBRID3, HYBRID_STEREO3, MOTIONPOSE3, POSE2POINT3, PRIOR6,
                                  SMOOTH_HYBRID6, SMOOTH_POSE6, STEREO3, TERNARY3, CLASS_DIM, SLOT_CLASS, FactorBlock, Problem)
from oracle import oracle as O


def one_factor_problem(ftype, poses, points, flows=(), meas=None, sigma=(1.0,), aux=(), calib=None, robust_k=0.0):
    """Problem with a single factor whose key slots take the variables in order."""
    ip = ipt = ifl = 0
    idx = []
    for c in SLOT_CLASS[ftype]:
        if c == 0: idx.append(ip); ip += 1
        elif c == 1: idx.append(ipt); ipt += 1
        else: idx.append(ifl); ifl += 1
    blk = FactorBlock(ftype, np.array([idx]), None if meas is None else np.asarray(meas, dtype=float).reshape(1, -1),
                      np.asarray(sigma, dtype=float), robust_k, aux_idx=np.array([0]) if len(aux) else None)
    kw = {} if calib is None else dict(calib=np.asarray(calib, dtype=float))
    return Problem(np.asarray(poses, dtype=float).reshape(-1, 12), np.asarray(points, dtype=float).reshape(-1, 3),
                   flow=np.asarray(flows, dtype=float).reshape(-1, 2), aux_pose=np.asarray(aux, dtype=float).reshape(-1, 12),
                   blocks=[blk], **kw)


def numerical_jacobian(prob, delta=1e-5):
    """gtsam::numericalDerivative: central differences through retract, per key slot."""
    t = prob.blocks[0].type
    cols = []
    for slot, c in enumerate(SLOT_CLASS[t]):
        vi = prob.blocks[0].idx[0, slot]
        for j in range(CLASS_DIM[c]):
            vals = []
            for sgn in (+1, -1):
                q = prob.copy()
                d = np.zeros(CLASS_DIM[c]); d[j] = sgn*delta
                if c == 0: q.pose[vi] = O.se3_retract(q.pose[vi], d)
                elif c == 1: q.point[vi] += d
                else: q.flow[vi] += d
                r, _ = O.OracleProblem(q).factor_eval(0, 0)
                vals.append(r)
            cols.append((vals[0] - vals[1])/(2*delta))
    return np.stack(cols, 1)


def check_jacobian(prob, tol):
    r, J = O.OracleProblem(prob).factor_eval(0, 0)
    Jn = numerical_jacobian(prob)
    assert np.max(np.abs(J - Jn)) < tol, (J, Jn)
    return r, J


RNG = np.random.default_rng(7)
H_FIX = lie.pose(lie.rodrigues(-0.1, 0.2, 0.25), [0.05, -0.10, 0.20])
X_K = lie.pose(lie.ypr(0.1, 0.2, 0.3), [1, 2, 3])
E_H = lie.pose(lie.ypr(0.4, 0.1, -0.2), [-1, 0.5, 2])
L_E = lie.pose(lie.ypr(-0.1, 0.0, 0.1), [0.1, 0.0, 0.0])


def perturb(P, sigma):
    return O.se3_retract(P, RNG.normal(0, sigma, 6))


# ---------------------------------------------------------------- test_factors.cc:134-196
def test_ternary_jacobians():
    Hp = perturb(H_FIX, 0.3)
    P1 = np.array([0.4, 1.0, 0.8]); P2 = lie.transform_from(H_FIX, P1)[0]
    check_jacobian(one_factor_problem(TERNARY3, [Hp], [P1, P2], sigma=[0.1]), 1e-9)


def test_ternary_zero_error():
    P1 = np.array([0.4, 1.0, 0.8]); P2 = lie.transform_from(H_FIX, P1)[0]
    r, _ = O.OracleProblem(one_factor_problem(TERNARY3, [H_FIX], [P1, P2], sigma=[0.1])).factor_eval(0, 0)
    assert np.allclose(r, 0, atol=1e-4)


# ---------------------------------------------------------------- test_factors.cc:92-132
def test_flow_projection_jacobians():
    prev = perturb(lie.identity()[0], 0.4)
    cur = O.se3_compose(prev, H_FIX)
    calib = [554.256, 554.256, 0.0, 320.0, 240.0, 0.0]
    meas = np.concatenate([[1.2, 2.4], [0.5], prev])
    check_jacobian(one_factor_problem(FLOWPROJ2, [cur], [], flows=[[0.1, -0.3]], meas=meas, sigma=[0.1], calib=calib), 1e-4)


# ---------------------------------------------------------------- test_hybrid_motion.cc:45-69
def test_project_to_object_roundtrip():
    """testProjections: projectToObject3 inverts projectToCamera3."""
    m_L = np.array([1.5, -0.5, 2.0])
    p = one_factor_problem(HYBRID3, [X_K, E_H], [m_L], meas=[0, 0, 0], aux=[L_E])
    r, _ = O.OracleProblem(p).factor_eval(0, 0)          # residual with z = 0 is the camera-frame point
    m_obj, *_ = O.hybrid_project_to_object3(X_K, E_H, L_E, r)
    assert np.allclose(m_obj, m_L, atol=1e-9)


# ---------------------------------------------------------------- test_hybrid_motion.cc:71-145
def test_hybrid_factor_jacobians():
    m_L = np.array([1.5, -0.5, 2.0])
    p0 = one_factor_problem(HYBRID3, [X_K, E_H], [m_L], meas=[0, 0, 0], aux=[L_E])
    Zk, _ = O.OracleProblem(p0).factor_eval(0, 0)
    p = one_factor_problem(HYBRID3, [X_K, E_H], [m_L], meas=Zk, aux=[L_E])
    r, J = check_jacobian(p, 1e-5)
    assert np.allclose(r, 0, atol=1e-12) and J.shape == (3, 15)


def test_hybrid_closed_form_matches_chain():
    """The closed form used by the CUDA kernels equals the reference's Adjoint chain (new-vs-original, 1e-9)."""
    m_L = np.array([1.5, -0.5, 2.0])
    p = one_factor_problem(HYBRID3, [X_K, E_H], [m_L], meas=[0.3, -0.2, 0.1], aux=[L_E])
    r, J = O.OracleProblem(p).factor_eval(0, 0)
    RX, RH, RL = lie.rot(X_K)[0], lie.rot(E_H)[0], lie.rot(L_E)[0]
    qo = RL @ m_L + lie.trans(L_E)[0]
    pw = RH @ qo + lie.trans(E_H)[0]
    q = RX.T @ (pw - lie.trans(X_K)[0])
    JX = np.concatenate([lie.skew(q), -np.eye(3)], 1)
    JH = RX.T @ RH @ np.concatenate([-lie.skew(qo), np.eye(3)], 1)
    Jm = RX.T @ RH @ RL
    assert np.allclose(r, q - [0.3, -0.2, 0.1], atol=1e-12)
    assert np.allclose(J, np.concatenate([JX, JH, Jm], 1), atol=1e-9)


# ---------------------------------------------------------------- test_hybrid_motion.cc:147-176 (CompareOriginal_*)
def test_project_to_object_original_algebra():
    Zk = np.array([-0.5, 1.2, 3.0])
    new, *_ = O.hybrid_project_to_object3(X_K, E_H, L_E, Zk)
    Li = O.se3_inverse(L_E)
    k_H_s0_k = O.se3_inverse(O.se3_compose(O.se3_compose(Li, E_H), L_E))
    L_k = O.se3_compose(E_H, L_E)
    k_H_s0_W = O.se3_compose(O.se3_compose(L_k, k_H_s0_k), O.se3_inverse(L_k))
    T = O.se3_compose(O.se3_compose(Li, k_H_s0_W), X_K)
    orig = lie.transform_from(T, Zk)[0]
    assert np.allclose(new, orig, atol=1e-9)


# ---------------------------------------------------------------- test_hybrid_motion.cc:180-204
def test_project_to_object_jacobians():
    Zk = np.array([-0.5, 1.2, 3.0])
    out, JX, JE, JL = O.hybrid_project_to_object3(X_K, E_H, L_E, Zk)
    d = 1e-5
    for Jact, which in ((JX, 0), (JE, 1), (JL, 2)):
        Jn = np.zeros((3, 6))
        for j in range(6):
            v = []
            for sgn in (+1, -1):
                xi = np.zeros(6); xi[j] = sgn*d
                args = [X_K, E_H, L_E]
                args[which] = O.se3_retract(args[which], xi)
                v.append(O.hybrid_project_to_object3(*args, Zk)[0])
            Jn[:, j] = (v[0] - v[1])/(2*d)
        assert np.max(np.abs(Jn - Jact)) < 1e-5


# ---------------------------------------------------------------- test_hybrid_motion.cc:260-343
def test_stereo_hybrid_jacobians():
    m_L = np.array([1.5, -0.5, 5.0])
    K = [1000, 1000, 0, 320, 240, 0.5]
    p0 = one_factor_problem(HYBRID_STEREO3, [X_K, E_H], [m_L], meas=[0, 0, 0], aux=[L_E], calib=K)
    perfect, _ = O.OracleProblem(p0).factor_eval(0, 0)
    meas = perfect + [2.0, -1.0, 0.5]
    r, J = check_jacobian(one_factor_problem(HYBRID_STEREO3, [X_K, E_H], [m_L], meas=meas, aux=[L_E], calib=K), 1e-5)
    assert np.allclose(r, [-2.0, 1.0, -0.5], atol=1e-9)


def test_stereo_cheirality_branch():
    """HybridFormulationFactors.cc:250-260: behind the camera -> zero Jacobians, error 2*fx."""
    K = [1000, 1000, 0, 320, 240, 0.5]
    I = lie.identity()[0]
    for t, pts, poses, aux in ((HYBRID_STEREO3, [[0, 0, -5.0]], [I, I], [I]), (STEREO3, [[0, 0, -5.0]], [I], [])):
        r, J = O.OracleProblem(one_factor_problem(t, poses, pts, meas=[0, 0, 0], aux=aux, calib=K)).factor_eval(0, 0)
        assert np.all(r == 2000.0) and np.all(J == 0)


# ---------------------------------------------------------------- GTSAM-ext factors: analytic vs numerical
@pytest.mark.parametrize("ftype", [POSE2POINT3, STEREO3])
def test_static_point_factor_jacobians(ftype):
    X = perturb(X_K, 0.2)
    p = lie.transform_from(X, [[0.4, -0.3, 6.0]])[0]
    check_jacobian(one_factor_problem(ftype, [X], [p], meas=[0.1, 0.2, 5.5], sigma=[0.2],
                                      calib=[1000, 1000, 0, 320, 240, 0.5]), 1e-5)


def test_numerical_factors_selfconsistent():
    """MOTIONPOSE3 / SMOOTH_* use gtsam::numericalDerivative in the reference; the oracle's J *is* that."""
    A, B, Cc = perturb(X_K, 0.1), perturb(X_K, 0.1), perturb(X_K, 0.1)
    for t, poses, pts, aux in ((MOTIONPOSE3, [A, B], [[1, 2, 3.0], [1.1, 2.1, 3.2]], []),
                               (SMOOTH_POSE6, [A, B, Cc], [], []), (SMOOTH_HYBRID6, [A, B, Cc], [], [L_E])):
        check_jacobian(one_factor_problem(t, poses, pts, aux=aux, sigma=[1.0]), 1e-9)


def test_prior_between_residuals():
    A, B = perturb(X_K, 0.1), perturb(X_K, 0.1)
    r, J = O.OracleProblem(one_factor_problem(PRIOR6, [A], [], meas=A, sigma=np.ones(6))).factor_eval(0, 0)
    assert np.allclose(r, 0, atol=1e-12) and np.allclose(J, np.eye(6))
    rel = O.se3_compose(O.se3_inverse(A), B)
    r, J = O.OracleProblem(one_factor_problem(BETWEEN6, [A, B], [], meas=rel, sigma=np.ones(6))).factor_eval(0, 0)
    assert np.allclose(r, 0, atol=1e-12)
    # at zero error the (non "slow but correct") BetweenFactor Jacobians are exact
    Jn = numerical_jacobian(one_factor_problem(BETWEEN6, [A, B], [], meas=rel, sigma=np.ones(6)))
    assert np.max(np.abs(J - Jn)) < 1e-6


def test_expmap_logmap_roundtrip():
    for _ in range(20):
        xi = RNG.normal(0, 0.7, 6)
        assert np.allclose(O.se3_logmap(O.se3_expmap(xi)), xi, atol=1e-10)
    assert np.allclose(O.se3_expmap(np.zeros(6)), lie.identity()[0])
    assert np.allclose(lie.se3_exp(np.array([[0.3, -0.2, 0.5, 1, 2, 3.0]]))[0], O.se3_expmap([0.3, -0.2, 0.5, 1, 2, 3.0]), atol=1e-14)


# ---------------------------------------------------------------- test_factors.cc:198-276 (reprojection error)
@pytest.mark.parametrize("case", ["identities", "L0_and_camera", "noise"])
def test_smart_factor_reprojection_error(case):
    I = lie.identity()[0]
    if case == "identities":
        pose, Le, noise = I, I, np.zeros(3)
    elif case == "L0_and_camera":
        pose, Le, noise = H_FIX, lie.pose(lie.ypr(-np.pi/10, 0., -np.pi/10), [0.5, 0.1, 0.3]), np.zeros(3)
    else:
        pose, Le, noise = I, I, np.array([0.4, 1.0, 2.0])
    pt = np.array([1.0, 2.0, 3.0])
    meas = lie.transform_to(pose, lie.transform_from(Le, pt + noise))[0]
    r, _ = O.OracleProblem(one_factor_problem(HYBRID3, [pose, I], [pt], meas=meas, aux=[Le], sigma=[0.05])).factor_eval(0, 0)
    assert np.allclose(r, -noise, atol=1e-5)


# ---------------------------------------------------------------- test_factors.cc:278-450
def test_basic_schur_complement():
    """G = F^T F - F^T E (E^T E)^-1 E^T F, g = F^T (b - E (E^T E)^-1 E^T b) equals the oracle's reduced system."""
    I = lie.identity()[0]
    pt = np.array([1.0, 1.4, 2.0]); noise = np.array([10.0, 0.0, 7.0]); meas = pt + noise
    poses = np.stack([I, I, H_FIX, I])            # pose1, motion, pose2, motion1
    blk = FactorBlock(HYBRID3, np.array([[0, 1, 0], [2, 3, 0]]), np.stack([meas, meas + 2*noise]), np.array([0.05]),
                      aux_idx=np.array([0, 0]))
    prob = Problem(poses, pt[None], aux_pose=I[None], blocks=[blk])
    op = O.OracleProblem(prob)
    A, b = op.linearize_block(0)                  # whitened (createReducedMatrix whitens too, HybridEstimator.hpp:365)
    F = np.zeros((6, 24)); F[:3, :12] = A[0][:, :12]; F[3:, 12:] = A[1][:, :12]
    E = np.concatenate([A[0][:, 12:], A[1][:, 12:]], 0); bb = b.reshape(-1)
    P = np.linalg.inv(E.T @ E)
    G = F.T @ F - F.T @ E @ P @ E.T @ F
    g = F.T @ (bb - E @ P @ E.T @ bb)
    S, gS, pos = op.reduced_dense(0.0)
    perm = np.concatenate([6*pos[i] + np.arange(6) for i in range(4)])
    assert np.allclose(S[np.ix_(perm, perm)], G, rtol=1e-9, atol=1e-5*np.abs(G).max())
    assert np.allclose(gS[perm], g, rtol=1e-9, atol=1e-5*np.abs(g).max())


# ---------------------------------------------------------------- test_factors.cc:462-556
def test_simple_optimise_recovers_ground_truth():
    I = lie.identity()[0]
    pose1, pose2 = I, H_FIX
    L_e = perturb(I, 2.0)
    motion1, motion2 = I, perturb(I, 3.0)
    pts = np.array([[3.0, 0, 1.2], [2.0, -5, 3.2]])
    sigma = 0.001
    gt = np.stack([pose1, motion1, pose2, motion2])
    idx, meas = [], []
    for li, m in enumerate(pts):
        for (xi, hi) in ((0, 1), (2, 3)):
            w = lie.transform_from(gt[hi], lie.transform_from(L_e, m))
            meas.append(lie.transform_to(gt[xi], w)[0])     # exact measurements: GT is the unique minimiser
            idx.append([xi, hi, li])
    init = np.stack([perturb(p, sigma) for p in gt])
    blocks = [FactorBlock(PRIOR6, np.array([[1]]), I[None], np.full(6, 1e-5)),
              FactorBlock(PRIOR6, np.array([[0]]), I[None], np.full(6, 1e-5)),   # fixes the camera gauge
              FactorBlock(HYBRID3, np.array(idx), np.array(meas), np.array([sigma]), aux_idx=np.zeros(4, dtype=np.int32))]
    # 2 points seen twice cannot pin pose2/motion2 fully; add two more shared points as the smart factors' triangulation does
    extra = np.array([[1.0, 1.0, 2.0], [-1.0, 0.5, 4.0]])
    for li, m in enumerate(extra):
        for (xi, hi) in ((0, 1), (2, 3)):
            w = lie.transform_from(gt[hi], lie.transform_from(L_e, m))
            meas.append(lie.transform_to(gt[xi], w)[0]); idx.append([xi, hi, 2 + li])
    blocks[2] = FactorBlock(HYBRID3, np.array(idx), np.array(meas), np.array([sigma]), aux_idx=np.zeros(len(idx), dtype=np.int32))
    allpts = np.concatenate([pts, extra]) + RNG.normal(0, sigma, (4, 3))
    prob = Problem(init, allpts, aux_pose=L_e[None], blocks=blocks)
    op = O.OracleProblem(prob)
    st = op.optimize(rel_tol=1e-8, abs_tol=0.0, max_iterations=20)
    assert st["error_final"] < 1e-12*max(st["error_initial"], 1.0) + 1e-9
    # pose/motion pairs are only determined up to the product X^-1 H (the reference test has the same gauge and
    # pins motion1 by a prior); check the measurable quantity and the pinned variables at the reference's 1e-5
    assert np.allclose(op.pose[0], pose1, atol=1e-5) and np.allclose(op.pose[1], motion1, atol=1e-5)
    rel_est = O.se3_compose(O.se3_inverse(op.pose[2]), op.pose[3])
    rel_gt = O.se3_compose(O.se3_inverse(pose2), motion2)
    assert np.allclose(rel_est, rel_gt, atol=1e-5)


def test_lie_maps_against_scipy():
    """External pin of the manifold maps the oracle restates from GTSAM 4.2: Pose3::Expmap / Logmap with
    the [omega; v] tangent order are the matrix exponential / logarithm of the 4x4 twist, Rot3::Expmap is Rodrigues -- checked
    against scipy.linalg.expm / logm, scipy's Rotation and cv2.Rodrigues (none of them shares code with the oracle)."""
    import cv2
    from scipy.linalg import expm, logm
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(12)
    for scale in (1e-9, 1e-3, 0.5, 2.5):
        for _ in range(5):
            xi = rng.normal(0, 1, 6)*np.array([scale, scale, scale, 1.0, 1.0, 1.0])
            w, v = xi[:3], xi[3:]
            twist = np.zeros((4, 4)); twist[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]); twist[:3, 3] = v
            T = expm(twist)
            P = O.se3_expmap(xi)
            # GTSAM switches to t = v below theta^2 = DBL_EPSILON (Pose3::Expmap): first-order accurate there
            ttol = 1e-12 if w @ w > np.finfo(float).eps else np.linalg.norm(w)*np.linalg.norm(v) + 1e-15
            assert np.abs(P[:9].reshape(3, 3) - T[:3, :3]).max() < 1e-12 and np.abs(P[9:] - T[:3, 3]).max() < ttol
            assert np.abs(P[:9].reshape(3, 3) - Rotation.from_rotvec(w).as_matrix()).max() < 1e-12
            assert np.abs(P[:9].reshape(3, 3) - cv2.Rodrigues(w.reshape(3, 1))[0]).max() < 1e-9
            if np.linalg.norm(w) < 3.0:                                # inside the principal branch of the logarithm
                L = np.real(logm(T))
                back = O.se3_logmap(P)
                ref = np.array([L[2, 1], L[0, 2], L[1, 0], L[0, 3], L[1, 3], L[2, 3]])
                assert np.abs(back - ref).max() < 1e-8*max(1.0, np.abs(ref).max())
    # compose / inverse / retract against plain 4x4 algebra
    def mat(P):
        M = np.eye(4); M[:3, :3] = np.asarray(P)[:9].reshape(3, 3); M[:3, 3] = np.asarray(P)[9:]; return M
    a = O.se3_expmap(rng.normal(0, 0.7, 6)); b = O.se3_expmap(rng.normal(0, 0.7, 6)); xi = rng.normal(0, 0.3, 6)
    assert np.abs(mat(O.se3_compose(a, b)) - mat(a) @ mat(b)).max() < 1e-13
    assert np.abs(mat(O.se3_inverse(a)) - np.linalg.inv(mat(a))).max() < 1e-12
    assert np.abs(mat(O.se3_retract(a, xi)) - mat(a) @ mat(O.se3_expmap(xi))).max() < 1e-13      # retract(T, xi) = T Expmap(xi)


def test_gtsam_ext_residuals_against_independent_algebra():
    """External pin of the GTSAM factors the oracle restates (none of them has a test in the reference tree): residuals computed
    here with plain 4x4 matrices + scipy.linalg.logm, the Huber loss with scipy.special.huber.
      PriorFactor<Pose3>:    e = -Logmap(x^-1 prior)            BetweenFactor<Pose3>: e = Logmap(measured^-1 (p1^-1 p2))
      PoseToPointFactor:     e = R^T (p - t) - z                  GenericStereoFactor:  e = (uL, uR, v) - z
      Robust(Huber(k), Isotropic(sigma)): loss = huber(k, |e| / sigma)"""
    from scipy.linalg import logm
    from scipy.special import huber
    from dynosam_b200.problem import STEREO3
    rng = np.random.default_rng(13)

    def mat(P):
        M = np.eye(4); M[:3, :3] = np.asarray(P)[:9].reshape(3, 3); M[:3, 3] = np.asarray(P)[9:]; return M

    def vee(T):
        L = np.real(logm(T)); return np.array([L[2, 1], L[0, 2], L[1, 0], L[0, 3], L[1, 3], L[2, 3]])

    for _ in range(4):
        A, B, Mz = (O.se3_expmap(rng.normal(0, 0.4, 6)) for _ in range(3))
        r, _ = O.OracleProblem(one_factor_problem(PRIOR6, [A], [], meas=B, sigma=np.ones(6))).factor_eval(0, 0)
        assert np.abs(r + vee(np.linalg.inv(mat(A)) @ mat(B))).max() < 1e-9
        r, _ = O.OracleProblem(one_factor_problem(BETWEEN6, [A, B], [], meas=Mz, sigma=np.ones(6))).factor_eval(0, 0)
        assert np.abs(r - vee(np.linalg.inv(mat(Mz)) @ np.linalg.inv(mat(A)) @ mat(B))).max() < 1e-9
        p = rng.normal(0, 2, 3) + np.array([0, 0, 8.0]); z = rng.normal(0, 1, 3)
        q = (np.linalg.inv(mat(A)) @ np.append(p, 1.0))[:3]
        r, _ = O.OracleProblem(one_factor_problem(POSE2POINT3, [A], [p], meas=z, sigma=[1.0])).factor_eval(0, 0)
        assert np.abs(r - (q - z)).max() < 1e-12
        K = np.array([700.0, 650.0, 0.0, 600.0, 180.0, 0.5])
        if q[2] > 0.5:
            zs = rng.normal(0, 50, 3)
            r, _ = O.OracleProblem(one_factor_problem(STEREO3, [A], [p], meas=zs, sigma=[1.0], calib=K)).factor_eval(0, 0)
            proj = np.array([K[3] + K[0]*q[0]/q[2], K[3] + K[0]*(q[0] - K[5])/q[2], K[4] + K[1]*q[1]/q[2]])
            assert np.abs(r - (proj - zs)).max() < 1e-9
        # Huber over an isotropic model: the factor error is the scipy Huber loss of the whitened norm
        for k, sigma in ((1e-4, 0.2), (1.0, 0.2), (5.0, 0.5)):
            o = O.OracleProblem(one_factor_problem(POSE2POINT3, [A], [p], meas=z, sigma=[sigma], robust_k=k))
            assert abs(o.error() - huber(k, np.linalg.norm(q - z)/sigma)) <= 1e-12*max(1.0, o.error())


def test_lm_minimum_against_scipy_least_squares():
    """External pin of the LM run as a whole: on a Gaussian (robust off) graph the oracle's Levenberg-Marquardt must end in the
    least-squares minimum of the whitened residuals -- the same minimum scipy.optimize.least_squares (MINPACK-style trust
    region, unrelated code) finds from the same initial values through the same retraction."""
    from scipy.optimize import least_squares
    from dynosam_b200 import synth
    p = synth.make_problem(n_frames=6, n_objects=1, n_static=40, n_dynamic=30, formulation="hybrid", seed=3, robust=False, object_span=(6, 6))
    n = 6*p.n_pose + 3*p.n_point

    def residuals(delta):
        o = O.OracleProblem(p); o.retract(delta)
        return -np.concatenate([o.linearize_block(bi)[1].ravel() for bi in range(len(p.blocks))])

    o = O.OracleProblem(p)
    st = o.optimize(max_iterations=100, rel_tol=1e-12, abs_tol=1e-12)
    assert abs(0.5*np.sum(residuals(np.zeros(n))**2) - st["error_initial"]) <= 1e-9*st["error_initial"]
    sol = least_squares(residuals, np.zeros(n), method="trf", xtol=1e-14, ftol=1e-14, gtol=1e-12, max_nfev=200)
    assert abs(sol.cost - st["error_final"]) <= 1e-7*st["error_final"], (sol.cost, st["error_final"])
    assert st["error_final"] < 0.5*st["error_initial"]                # it did move: the graph starts well off its minimum

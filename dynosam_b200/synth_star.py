"""Seeded synthetic inputs of the per-object refinements of the front end: joint optical-flow + pose problems
(what OpticalFlowAndPoseOptimizer::optimize reads from two frames, MotionSolver-inl.hpp:117-160) and object-motion problems
(MotionOnlyRefinementOptimizer::optimize, :343-366).  Shared by tests/test_star.py, tests/golden/make_golden.py and
tools/star_bench.py, so that fixtures only store outputs."""
import numpy as np

from . import lie

K5 = np.array([721.5377, 721.5377, 0.0, 609.5593, 172.854])
FLOW_SIGMA, PRIOR_SIGMA, HUBER_K = 10.0, 3.33, 0.001          # FrontendParams-like magnitudes (flow in pixels)


def _project(X, p):
    q = lie.transform_to(np.tile(X, (len(p), 1)), p)
    return np.stack([K5[0]*q[:, 0]/q[:, 2] + K5[3], K5[1]*q[:, 1]/q[:, 2] + K5[4]], 1), q[:, 2]


def _back_project(X, kp, depth):
    pc = np.stack([(kp[:, 0] - K5[3])/K5[0]*depth, (kp[:, 1] - K5[4])/K5[1]*depth, depth], 1)
    return lie.transform_from(np.tile(X, (len(kp), 1)), pc)


def make_motion_problem(rng, n, px_noise=0.3, depth_noise=0.05, outliers=0.0):
    """two camera poses, a moving rigid object: key-points in both frames, back-projected (noisy depth) world points, a
    perturbed initial motion"""
    Xa = lie.se3_exp(rng.normal(0, 0.05, (1, 6)))[0]
    Xb = lie.compose(Xa[None], lie.se3_exp(np.array([[0.002, -0.01, 0.001, 0.02, -0.01, 0.6]])))[0]
    H = lie.se3_exp(np.array([[0.01, 0.03, -0.01, 0.3, 0.02, 0.5]]) + rng.normal(0, 0.01, (1, 6)))[0]
    kp0 = np.stack([rng.uniform(400, 800, n), rng.uniform(100, 300, n)], 1); d0 = rng.uniform(8, 20, n)
    ma = _back_project(Xa, kp0, d0)
    mb = lie.transform_from(np.tile(H, (n, 1)), ma)
    if outliers > 0:
        bad = rng.random(n) < outliers; mb[bad] += rng.normal(0, 0.5, (int(bad.sum()), 3))
    kpa, da = _project(Xa, ma); kpb, db = _project(Xb, mb)
    kpa = kpa + rng.normal(0, px_noise, (n, 2)); kpb = kpb + rng.normal(0, px_noise, (n, 2))
    ma0 = _back_project(Xa, kpa, da*(1 + rng.normal(0, depth_noise, n))); mb0 = _back_project(Xb, kpb, db*(1 + rng.normal(0, depth_noise, n)))
    H0 = lie.compose(H[None], lie.se3_exp(rng.normal(0, 0.02, (1, 6))))[0]
    return dict(pose_prev=Xa, pose_cur=Xb, motion_init=H0, calib=K5, kp_prev=kpa, kp_cur=kpb, points_init=np.concatenate([ma0, mb0], 1), gt=H)


def make_flow_pose_problem(rng, n, noise=0.5, outliers=0.0, behind=0):
    """a camera step between two frames: key-points and depths in the previous frame, noisy measured flows (gross outliers /
    points that end up behind the camera on request), a perturbed initial pose"""
    X_prev = lie.se3_exp(rng.normal(0, 0.05, (1, 6)))[0]
    step = lie.se3_exp(np.array([[0.01, -0.02, 0.005, 0.05, -0.02, 0.9]]) + rng.normal(0, 0.01, (1, 6)))[0]
    X_gt = lie.compose(X_prev[None], step[None])[0]
    kp = np.stack([rng.uniform(50, 1190, n), rng.uniform(30, 340, n)], 1); depth = rng.uniform(5, 40, n)
    pc = np.stack([(kp[:, 0] - K5[3])/K5[0]*depth, (kp[:, 1] - K5[4])/K5[1]*depth, depth], 1)
    pw = lie.transform_from(np.tile(X_prev, (n, 1)), pc) if n else np.zeros((0, 3))
    q = lie.transform_to(np.tile(X_gt, (n, 1)), pw) if n else np.zeros((0, 3))
    proj = np.stack([K5[0]*q[:, 0]/q[:, 2] + K5[3], K5[1]*q[:, 1]/q[:, 2] + K5[4]], 1)
    flow = proj - kp + rng.normal(0, noise, (n, 2))
    if outliers > 0 and n:
        bad = rng.random(n) < outliers
        flow[bad] += rng.normal(0, 40.0, (int(bad.sum()), 2))
    if behind and n:
        depth[:behind] = 0.2                                     # points that end up behind the camera: cheirality branch
    init = lie.compose(X_gt[None], lie.se3_exp(rng.normal(0, 0.02, (1, 6))))[0]
    return dict(pose_init=init, pose_prev=X_prev, calib=K5, kp_prev=kp, depth=depth, flow=flow, gt=X_gt)


def flow_parity_set():
    """problems of 1 .. 700 features: exact fits, CTA-size boundaries, outliers, cheirality"""
    rng = np.random.default_rng(11)
    sizes = [1, 2, 3, 7, 33, 64, 100, 255, 256, 257, 300, 511, 700] + list(rng.integers(20, 400, 12))
    return [make_flow_pose_problem(rng, int(n), noise=0.5, outliers=0.1 if i % 3 == 0 else 0.0, behind=2 if i % 5 == 4 and n > 10 else 0) for i, n in enumerate(sizes)]


def flow_rounds_set():
    """problems whose gross outliers trigger the outlier rounds"""
    rng = np.random.default_rng(21)
    sizes = [40, 80, 150, 200, 260, 300] + list(rng.integers(30, 250, 10))
    return [make_flow_pose_problem(rng, int(n), noise=0.3, outliers=(0.0, 0.1, 0.2)[i % 3], behind=2 if i % 4 == 3 else 0) for i, n in enumerate(sizes)]


def motion_set():
    rng = np.random.default_rng(31)
    sizes = [6, 7, 20, 64, 100, 255, 256, 257] + list(rng.integers(10, 120, 6))
    return [make_motion_problem(rng, int(n), outliers=0.1 if i % 2 else 0.0) for i, n in enumerate(sizes)]

"""batched star problems -- the front end's joint optical-flow + pose refinement
(OpticalFlowAndPoseOptimizer::optimize, MotionSolver-inl.hpp:88-260), all objects of a frame in one launch."""
import numpy as np
import pytest

from dynosam_b200 import binding, lie

K5 = np.array([721.5377, 721.5377, 0.0, 609.5593, 172.854])
FLOW_SIGMA, PRIOR_SIGMA, HUBER_K = 10.0, 3.33, 0.001          # FrontendParams-like magnitudes (flow in pixels)


def make_problem(rng, n, noise=0.5, outliers=0.0, behind=0):
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


def test_star_oracle_is_a_minimiser():
    """Pins the restatement: at its fixed point the gradient of the full objective vanishes and the pose is the ground truth
    (noise-free flows), so the dense LM in oracle/star_oracle.py solves the problem the reference poses."""
    from oracle import star_oracle as SO
    rng = np.random.default_rng(3)
    q = make_problem(rng, 60, noise=0.0)
    r = SO.flow_pose_lm(q["pose_init"], q["pose_prev"], K5, q["kp_prev"], q["depth"], q["flow"], FLOW_SIGMA, PRIOR_SIGMA, 0.0,
                        max_iterations=50, relative_error_tol=1e-14, absolute_error_tol=1e-14)
    assert r["error_final"] < 1e-12 and r["error_final"] < 1e-6*r["error_initial"]
    assert np.abs(r["pose"] - q["gt"]).max() < 1e-7
    assert np.abs(r["flow"] - q["flow"]).max() < 1e-6
    # Huber: outliers pull less -- the robust fit is closer to the ground truth than the Gaussian one
    q = make_problem(rng, 120, noise=0.2, outliers=0.2)
    g = SO.flow_pose_lm(q["pose_init"], q["pose_prev"], K5, q["kp_prev"], q["depth"], q["flow"], 1.0, 0.5, 0.0, max_iterations=10)
    h = SO.flow_pose_lm(q["pose_init"], q["pose_prev"], K5, q["kp_prev"], q["depth"], q["flow"], 1.0, 0.5, 1.0, max_iterations=10)
    assert h["iterations"] >= 1 and g["iterations"] >= 1
    assert np.abs(h["pose"][9:] - q["gt"][9:]).max() < np.abs(g["pose"][9:] - q["gt"][9:]).max()


def test_star_batch_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rng = np.random.default_rng(0)
    with pytest.raises(binding.DynobaError):
        binding.flow_pose_batch([make_problem(rng, 10)], FLOW_SIGMA, PRIOR_SIGMA, HUBER_K)


def test_star_batch_bad_arguments():
    L = binding.load()
    assert L.dynoba_flow_pose_batch(0, -1, None, None, None, None, None, None, None, 1.0, 1.0, 0.0, None, None, None, None, None, None, None) == -1
    assert L.dynoba_flow_pose_batch(0, 1, None, None, None, None, None, None, None, 1.0, 1.0, 0.0, None, None, None, None, None, None, None) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("sig", [(FLOW_SIGMA, PRIOR_SIGMA, HUBER_K), (1.0, 0.5, 1.0), (2.0, 1.0, 0.0)])
def test_star_batch_matches_oracle(sig):
    """Every problem of the batch runs the same LM as the CPU restatement: same accepted / rejected steps, same error, same
    pose and flows (fp64; tolerance 1e-9 relative on chi^2, 1e-7 on the values)."""
    from oracle import star_oracle as SO
    rng = np.random.default_rng(11)
    sizes = [1, 2, 3, 7, 33, 64, 100, 255, 256, 257, 300, 511, 700] + list(rng.integers(20, 400, 27))
    probs = [make_problem(rng, int(n), noise=0.5, outliers=0.1 if i % 3 == 0 else 0.0, behind=2 if i % 5 == 4 and n > 10 else 0) for i, n in enumerate(sizes)]
    out = binding.flow_pose_batch(probs, *sig, max_iterations=10)
    assert len(out) == len(probs)
    moved = 0
    for q, r in zip(probs, out):
        o = SO.flow_pose_lm(q["pose_init"], q["pose_prev"], K5, q["kp_prev"], q["depth"], q["flow"], *sig, max_iterations=10)
        assert abs(r["error_initial"] - o["error_initial"]) <= 1e-11*max(o["error_initial"], 1.0)
        if o["error_final"] < 1e-12*o["error_initial"]:
            # exact fit (<= 3 features: 4N residuals, 6 + 2N unknowns): the last steps act on rounding noise; same minimum, no step-by-step claim
            assert r["error_final"] < 1e-10*o["error_initial"] and r["iterations"] >= 1
            moved += 1
            continue
        assert (r["iterations"], r["inner_iterations"]) == (o["iterations"], o["inner_iterations"]), (len(q["depth"]), r, o)
        assert abs(r["error_final"] - o["error_final"]) <= 1e-9*max(o["error_final"], 1e-12) + 1e-12
        assert np.abs(r["pose"] - o["pose"]).max() < 1e-7 and np.abs(r["flow"] - o["flow"]).max() < 1e-6
        moved += r["iterations"] > 0
    assert moved >= len(probs) - 2


@pytest.mark.gpu
def test_star_batch_empty_and_large():
    """Edge cases: no problems, a problem with no features (nothing to do: the pose stays), and a frame-sized batch (one CTA per
    problem, more problems than SMs) whose every problem lowers its error."""
    assert binding.flow_pose_batch([], FLOW_SIGMA, PRIOR_SIGMA, HUBER_K) == []
    rng = np.random.default_rng(2)
    probs = [make_problem(rng, 0)] + [make_problem(rng, int(n)) for n in rng.integers(30, 300, 400)]
    out = binding.flow_pose_batch(probs, 1.0, 0.5, 1.0, max_iterations=10)
    assert out[0]["iterations"] == 0 and np.array_equal(out[0]["pose"], np.asarray(probs[0]["pose_init"]))
    closer = 0
    for q, r in zip(probs[1:], out[1:]):
        assert r["error_final"] < r["error_initial"] and r["iterations"] >= 1
        closer += np.abs(r["pose"][9:] - q["gt"][9:]).max() < np.abs(np.asarray(q["pose_init"])[9:] - q["gt"][9:]).max()
    assert closer >= 0.9*(len(probs) - 1)

"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the CPU oracle on seeded inputs).

CPU: the oracle still reproduces them (guards the test infrastructure against silent changes).
GPU: the CUDA path reproduces them through the C ABI without the oracle being involved at run time.
"""
import os

import numpy as np
import pytest

from dynosam_b200 import synth
from dynosam_b200.problem import MOTIONPOSE3, SMOOTH_HYBRID6, SMOOTH_POSE6, TYPE_NAMES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NUMERIC = (MOTIONPOSE3, SMOOTH_HYBRID6, SMOOTH_POSE6)


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _check_linearization(g, p, lin, err_block, total_error, tol_scale=1.0):
    assert abs(total_error - float(g["error"])) <= 1e-6*float(g["error"])
    for bi, b in enumerate(p.blocks):
        assert int(g[f"type_{bi}"]) == b.type and int(g[f"n_{bi}"]) == b.n
        A, bv = lin(bi); e = err_block(bi)
        Ag, bg, eg = g[f"A_{bi}"], g[f"b_{bi}"], g[f"e_{bi}"]
        k = Ag.shape[0]
        sa = max(np.abs(Ag).max(), 1e-300); sb = max(np.abs(bg).max(), 1e-300)
        tol = (1e-6 if b.type in NUMERIC else 1e-9)*tol_scale
        assert np.abs(A[:k] - Ag).max() <= tol*sa, TYPE_NAMES[b.type]
        assert np.abs(bv[:k] - bg).max() <= 1e-9*tol_scale*sb + 1e-12, TYPE_NAMES[b.type]
        assert np.abs(e[:k] - eg).max() <= 1e-6*max(np.abs(eg).max(), 1e-300), TYPE_NAMES[b.type]
        assert abs(e.sum() - float(g[f"esum_{bi}"])) <= 1e-6*max(abs(float(g[f"esum_{bi}"])), 1e-300)


def test_oracle_reproduces_golden_linearization():
    from oracle import oracle as O
    p = synth.make_all_types_problem(3)
    o = O.OracleProblem(p)
    _check_linearization(_load("all_types_linearization.npz"), p, o.linearize_block, o.error_block, o.error(), tol_scale=1e-3)


@pytest.mark.parametrize("formulation", ["hybrid", "wcme"])
def test_oracle_reproduces_golden_lm(formulation):
    from oracle import oracle as O
    g = _load(f"c1_{formulation}_lm.npz")
    p = synth.make_config("C1", formulation=formulation)
    rc, d = O.OracleProblem(p).schur_solve(float(g["lambda"]))
    assert rc == 0 and np.linalg.norm(d[:64] - g["step_head"]) <= 1e-9*np.linalg.norm(g["step_head"])
    o = O.OracleProblem(p); r = o.optimize()
    assert r["iterations"] == int(g["iterations"]) and r["inner_iterations"] == int(g["inner_iterations"])
    assert abs(r["error_final"] - float(g["error_final"])) <= 1e-9*float(g["error_final"])


@pytest.mark.gpu
def test_cuda_reproduces_golden_linearization():
    from dynosam_b200.binding import Solver
    p = synth.make_all_types_problem(3)
    s = Solver(p)
    s.linearize()
    _check_linearization(_load("all_types_linearization.npz"), p, s.linearization, s.factor_errors, s.error())
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("formulation", ["hybrid", "wcme"])
def test_cuda_reproduces_golden_lm(formulation):
    from dynosam_b200.binding import Solver
    g = _load(f"c1_{formulation}_lm.npz")
    p = synth.make_config("C1", formulation=formulation)
    s = Solver(p)
    d = s.solve(float(g["lambda"]))
    assert abs(np.linalg.norm(d) - float(g["step_norm"])) <= 1e-6*float(g["step_norm"])
    assert np.linalg.norm(d[:64] - g["step_head"]) <= 1e-6*np.linalg.norm(g["step_head"])
    s.close()
    s = Solver(p); st = s.optimize()
    assert st["iterations"] == int(g["iterations"]) and st["inner_iterations"] == int(g["inner_iterations"])
    assert abs(st["error_initial"] - float(g["error_initial"])) <= 1e-6*float(g["error_initial"])
    assert abs(st["error_final"] - float(g["error_final"])) <= 1e-6*float(g["error_final"])
    pose, point, _ = s.values()
    assert np.abs(pose.sum(0) - g["pose_sum"]).max() <= 1e-4 and np.abs(point.sum(0) - g["point_sum"]).max() <= 1e-2
    s.close()


# ---- batched star problems: outputs of oracle/star_oracle.py on dynosam_b200/synth_star.py's seeded sets
STAR_SIGMAS = dict(flow_sigma=1.0, flow_prior_sigma=0.5, huber_k=1.0)


def _check_star_flow(g, results, offset=0):
    """results[i] belongs to problem offset + i of the fixture; exact-fit problems (<= 3 features) are compared at the minimum"""
    starts = np.concatenate([[0], np.cumsum(g["n"])])
    same = 0
    for i, r in enumerate(results):
        j = offset + i
        assert len(r["flow"]) == int(g["n"][j])
        assert abs(r["error_initial"] - g["error_initial"][j]) <= 1e-11*max(g["error_initial"][j], 1.0)
        assert r["rounds"] == int(g["rounds"][j])
        assert np.array_equal(np.asarray(r["inlier"], dtype=np.uint8), g["inlier_bits"][starts[j]:starts[j + 1]])
        if g["error_final"][j] < 1e-12*g["error_initial"][j]:
            assert r["error_final"] < 1e-10*g["error_initial"][j]
            same += 1
            continue
        if (r["iterations"], r["inner_iterations"]) == (int(g["iterations"][j]), int(g["inner_iterations"][j])):
            same += 1
            assert abs(r["error_final"] - g["error_final"][j]) <= 1e-9*max(g["error_final"][j], 1e-12) + 1e-12
            assert np.abs(r["pose"] - g["pose"][j]).max() < 1e-7
            assert np.abs(r["flow"].sum(0) - g["flow_sum"][j]).max() < 1e-6*max(len(r["flow"]), 1)
        else:       # a stopping test fell on the other side in the last digits: same minimum, one iteration more or less
            assert abs(r["iterations"] - int(g["iterations"][j])) <= 1
            assert abs(r["error_final"] - g["error_final"][j]) <= 1e-4*g["error_final"][j] and np.abs(r["pose"] - g["pose"][j]).max() < 1e-4
    assert same >= len(results) - 2, same


def _check_star_motion(g, results, offset=0):
    same = 0
    for i, r in enumerate(results):
        j = offset + i
        assert len(r["points"]) == int(g["n"][j])
        assert abs(r["error_initial"] - g["error_initial"][j]) <= 1e-10*g["error_initial"][j]
        if (r["iterations"], r["inner_iterations"]) != (int(g["iterations"][j]), int(g["inner_iterations"][j])):
            continue
        same += 1
        assert abs(r["error_final"] - g["error_final"][j]) <= 1e-3*g["error_final"][j]
        assert np.abs(r["motion"] - g["motion"][j]).max() < 1e-4 and np.abs(r["poses"] - g["poses"][j]).max() < 1e-6
        assert abs(r["motion_factor_error"].sum() - g["factor_error_sum"][j]) <= 1e-3*max(g["factor_error_sum"][j], 1.0)
    assert same >= len(results) - 3, same


def test_oracle_reproduces_golden_star():
    """the dense restatements still give the stored runs (a slice of every set, to keep the CPU suite short)"""
    from dynosam_b200 import synth_star
    from oracle import star_oracle as SO
    sig = (STAR_SIGMAS["flow_sigma"], STAR_SIGMAS["flow_prior_sigma"], STAR_SIGMAS["huber_k"])
    for name, probs, rounds, lo, hi in (("star_flow_pose_lm.npz", synth_star.flow_parity_set(), 0, 2, 9), ("star_flow_pose_rounds.npz", synth_star.flow_rounds_set(), 4, 0, 4)):
        rs = [SO.flow_pose_refine(q["pose_init"], q["pose_prev"], synth_star.K5, q["kp_prev"], q["depth"], q["flow"], *sig, outlier_rounds=rounds, max_iterations=10)
              for q in probs[lo:hi]]
        _check_star_flow(_load(name), rs, offset=lo)
    probs = synth_star.motion_set()
    rs = [SO.motion_refine_lm(q["pose_prev"], q["pose_cur"], q["motion_init"], synth_star.K5, q["kp_prev"], q["kp_cur"], q["points_init"]) for q in probs[:5]]
    _check_star_motion(_load("star_motion_refine.npz"), rs)


@pytest.mark.gpu
def test_cuda_reproduces_golden_star():
    """dynoba_flow_pose_batch / dynoba_motion_refine_batch against the stored runs, no oracle at run time"""
    from dynosam_b200 import binding, synth_star
    _check_star_flow(_load("star_flow_pose_lm.npz"), binding.flow_pose_batch(synth_star.flow_parity_set(), outlier_rounds=0, **STAR_SIGMAS))
    _check_star_flow(_load("star_flow_pose_rounds.npz"), binding.flow_pose_batch(synth_star.flow_rounds_set(), **STAR_SIGMAS))
    _check_star_motion(_load("star_motion_refine.npz"), binding.motion_refine_batch(synth_star.motion_set()))

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

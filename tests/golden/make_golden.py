#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ from the CPU oracle (oracle/dynoba_oracle.c) on seeded inputs.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

The reference itself (GTSAM + DynOSAM, C++) cannot be built in this container, so these vectors are the ORACLE's
outputs, not the reference's: they pin the oracle (tests/test_golden.py, CPU) and give the CUDA path a comparison that
needs no oracle at run time (tests/test_golden.py, GPU).  The oracle in turn is pinned against the reference's own
known-answer tests in tests/test_oracle_kat.py.  Inputs are regenerated from the seed by the same generator
(dynosam_b200/synth.py), so only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dynosam_b200 import synth, synth_star          # noqa: E402
from oracle import oracle as O                      # noqa: E402
from oracle import star_oracle as SO                # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = 48                                           # factors kept per block (the first KEEP in the caller's order)


def linearization_fixture(p):
    o = O.OracleProblem(p)
    out = {"error": np.array(o.error())}
    for bi, b in enumerate(p.blocks):
        A, bv = o.linearize_block(bi)
        e = o.error_block(bi)
        k = min(KEEP, b.n)
        out[f"type_{bi}"] = np.array(b.type); out[f"n_{bi}"] = np.array(b.n)
        out[f"A_{bi}"] = A[:k].copy(); out[f"b_{bi}"] = bv[:k].copy(); out[f"e_{bi}"] = e[:k].copy()
        out[f"esum_{bi}"] = np.array(e.sum())
    return out


def lm_fixture(p, lam=1e-4):
    o = O.OracleProblem(p)
    rc, d = o.schur_solve(lam)
    assert rc == 0
    st = O.OracleProblem(p)
    r = st.optimize()
    return {"lambda": np.array(lam), "step_norm": np.array(np.linalg.norm(d)), "step_head": d[:64].copy(),
            "iterations": np.array(r["iterations"]), "inner_iterations": np.array(r["inner_iterations"]),
            "error_initial": np.array(r["error_initial"]), "error_final": np.array(r["error_final"]),
            "pose_sum": st.pose.sum(0), "point_sum": st.point.sum(0)}


STAR_SIGMAS = (1.0, 0.5, 1.0)                       # flow_sigma, flow_prior_sigma, huber_k of the star fixtures


def star_flow_fixture(probs, outlier_rounds):
    """per problem: the LM run of oracle/star_oracle.py (dense restatement of OpticalFlowAndPoseOptimizer::optimize)"""
    rs = [SO.flow_pose_refine(q["pose_init"], q["pose_prev"], synth_star.K5, q["kp_prev"], q["depth"], q["flow"], *STAR_SIGMAS,
                              outlier_rounds=outlier_rounds, max_iterations=10) for q in probs]
    return {"n": np.array([len(q["depth"]) for q in probs]), "pose": np.stack([r["pose"] for r in rs]),
            "flow_sum": np.stack([r["flow"].sum(0) if len(r["flow"]) else np.zeros(2) for r in rs]),
            "n_inlier": np.array([int(r["inlier"].sum()) for r in rs]), "inlier_bits": np.concatenate([r["inlier"] for r in rs]).astype(np.uint8),
            "rounds": np.array([r["rounds"] for r in rs]), "iterations": np.array([r["iterations"] for r in rs]),
            "inner_iterations": np.array([r["inner_iterations"] for r in rs]),
            "error_initial": np.array([r["error_initial"] for r in rs]), "error_final": np.array([r["error_final"] for r in rs])}


def star_motion_fixture(probs):
    rs = [SO.motion_refine_lm(q["pose_prev"], q["pose_cur"], q["motion_init"], synth_star.K5, q["kp_prev"], q["kp_cur"], q["points_init"]) for q in probs]
    return {"n": np.array([len(q["kp_prev"]) for q in probs]), "motion": np.stack([r["motion"] for r in rs]), "poses": np.stack([r["poses"] for r in rs]),
            "iterations": np.array([r["iterations"] for r in rs]), "inner_iterations": np.array([r["inner_iterations"] for r in rs]),
            "error_initial": np.array([r["error_initial"] for r in rs]), "error_final": np.array([r["error_final"] for r in rs]),
            "factor_error_sum": np.array([r["motion_factor_error"].sum() for r in rs])}


def main():
    np.savez_compressed(os.path.join(HERE, "star_flow_pose_lm.npz"), **star_flow_fixture(synth_star.flow_parity_set(), 0))
    np.savez_compressed(os.path.join(HERE, "star_flow_pose_rounds.npz"), **star_flow_fixture(synth_star.flow_rounds_set(), 4))
    np.savez_compressed(os.path.join(HERE, "star_motion_refine.npz"), **star_motion_fixture(synth_star.motion_set()))
    np.savez_compressed(os.path.join(HERE, "all_types_linearization.npz"), **linearization_fixture(synth.make_all_types_problem(3)))
    for form in ("hybrid", "wcme"):
        np.savez_compressed(os.path.join(HERE, f"c1_{form}_lm.npz"), **lm_fixture(synth.make_config("C1", formulation=form)))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()

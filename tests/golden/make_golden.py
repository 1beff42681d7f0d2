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
from dynosam_b200 import synth                      # noqa: E402
from oracle import oracle as O                      # noqa: E402

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


def main():
    np.savez_compressed(os.path.join(HERE, "all_types_linearization.npz"), **linearization_fixture(synth.make_all_types_problem(3)))
    for form in ("hybrid", "wcme"):
        np.savez_compressed(os.path.join(HERE, f"c1_{form}_lm.npz"), **lm_fixture(synth.make_config("C1", formulation=form)))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()

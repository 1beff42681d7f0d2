"""graph construction straight into SoA blocks (dynoba_builder_*).  The builder gets what the reference's
Map holds -- frames, camera pose estimates, odometry, static / dynamic point observations, front-end motions -- and must
emit exactly the graph the formulation rules produce.  Check: the observation tables of a seeded synthetic scenario
(dynosam_b200/synth.py restates the same rules independently, in numpy) are fed frame by frame and the emitted arrays are
compared block by block; the oracle then evaluates both graphs."""
import numpy as np
import pytest

from dynosam_b200 import lie, synth
from dynosam_b200.problem import BETWEEN6, HYBRID3, POSE2POINT3, PRIOR6, SMOOTH_HYBRID6


def _feed(p, b):
    """replay a synthetic hybrid problem as per-frame observations"""
    N = p.meta["n_frames"]
    ptp = [x for x in p.blocks if x.type == POSE2POINT3][0]
    hyb = [x for x in p.blocks if x.type == HYBRID3]
    odo = [x for x in p.blocks if x.type == BETWEEN6][0]
    rel = {int(i[1]): m for i, m in zip(odo.idx, odo.meas)}
    motion_frame = p.pose_order                                   # frame of every pose-like variable
    obj_of_motion = {}
    if hyb:
        h = hyb[0]
        for f in range(h.n):
            obj_of_motion[int(h.idx[f, 1])] = int(h.aux_idx[f]) + 1
        for j in range(p.aux_pose.shape[0]):
            frames_j = [int(motion_frame[m]) for m, o in obj_of_motion.items() if o == j + 1]
            b.set_keyframe_pose(j + 1, min(frames_j), p.aux_pose[j])
        for m, o in obj_of_motion.items():
            b.set_motion_init(o, int(motion_frame[m]), p.pose[m])
    for k in range(N):
        b.add_frame(k, p.pose[k], rel.get(k))
        sel = np.flatnonzero(ptp.idx[:, 0] == k)
        sel = sel[np.argsort(ptp.idx[sel, 1], kind="stable")]
        b.add_static(k, ptp.idx[sel, 1], ptp.meas[sel])
        if hyb:
            h = hyb[0]
            sel = np.flatnonzero(h.idx[:, 0] == k)
            sel = sel[np.argsort(h.idx[sel, 2], kind="stable")]
            b.add_dynamic(k, h.idx[sel, 2], h.aux_idx[sel] + 1, h.meas[sel])


@pytest.mark.parametrize("cfg", [dict(n_frames=20, n_objects=1, n_static=700, n_dynamic=300, seed=42),
                                 dict(n_frames=40, n_objects=3, n_static=500, n_dynamic=400, seed=7, object_span=(12, 25))])
def test_builder_emits_the_formulation_graph(cfg):
    from dynosam_b200.builder import GraphBuilder
    from oracle import oracle as O
    p = synth.make_problem(formulation="hybrid", **cfg)
    b = GraphBuilder()
    _feed(p, b)
    q = b.problem()
    assert q.n_pose == p.n_pose and q.n_point == p.n_point and q.aux_pose.shape == p.aux_pose.shape
    assert np.array_equal(q.pose_order, p.pose_order) and np.array_equal(q.pose_keys, p.pose_keys)
    assert np.array_equal(q.pose, p.pose) and np.array_equal(q.aux_pose, p.aux_pose)
    assert np.abs(q.point - p.point).max() <= 1e-12*max(np.abs(p.point).max(), 1.0)       # initial values: X z / projectToObject3
    assert [x.type for x in q.blocks] == [x.type for x in p.blocks] == [POSE2POINT3, HYBRID3, PRIOR6, SMOOTH_HYBRID6, PRIOR6, BETWEEN6]
    for x, y in zip(q.blocks, p.blocks):
        assert np.array_equal(x.idx, y.idx), x.type
        assert (x.meas is None and y.meas is None) or np.array_equal(x.meas, y.meas), x.type
        assert np.array_equal(np.ravel(x.sigma), np.ravel(y.sigma)) and x.robust_k == y.robust_k, x.type
        assert (x.aux_idx is None and y.aux_idx is None) or np.array_equal(x.aux_idx, y.aux_idx), x.type
    assert abs(O.OracleProblem(q).error() - O.OracleProblem(p).error()) <= 1e-12*O.OracleProblem(p).error()
    b.close()


def test_builder_rules():
    """minimum observation counts, key-framing of an object that disappears, centroid key-frame pose"""
    from dynosam_b200.builder import GraphBuilder
    I = lie.identity()[0]
    b = GraphBuilder()
    for k in range(10):
        b.add_frame(k, I, None if k == 0 else I)
    b.add_static(0, [5, 6], [[0, 0, 4], [1, 0, 4]]); b.add_static(1, [5], [[0, 0, 4]])          # tracklet 6 seen once: dropped
    # object 1 seen in frames 0-2 and 6-8 (gap of 3 > keyframe_gap 2 -> second key-frame at 6); tracklet 9 has only two observations
    for k in (0, 1, 2):
        b.add_dynamic(k, [1, 2], [1, 1], [[1.0, 0.0, 5.0], [3.0, 0.0, 5.0]])
    for k in (6, 7, 8):
        b.add_dynamic(k, [3], [1], [[2.0, 1.0, 6.0]])
    b.add_dynamic(6, [9], [1], [[0.0, 0.0, 7.0]]); b.add_dynamic(7, [9], [1], [[0.0, 0.0, 7.0]])
    q = b.problem()
    assert q.n_point == 1 + 3                                      # static 5; dynamic 1, 2, 3
    assert q.n_pose == 10 + 6 and q.aux_pose.shape[0] == 2
    assert np.allclose(q.aux_pose[0][9:], [2.0, 0.0, 5.0]) and np.allclose(q.aux_pose[0][:9], np.eye(3).ravel())   # centroid of (1,0,5), (3,0,5)
    types = [x.type for x in q.blocks]
    assert types == [POSE2POINT3, HYBRID3, PRIOR6, SMOOTH_HYBRID6, PRIOR6, BETWEEN6]
    hyb = q.blocks[1]
    assert hyb.n == 9 and set(hyb.aux_idx[:6]) == {0} and set(hyb.aux_idx[6:]) == {1}
    assert q.blocks[2].n == 2 and q.blocks[3].n == 2               # one prior and one smoothing triple per key-frame segment
    assert q.blocks[5].n == 9
    b.close()


@pytest.mark.gpu
def test_builder_emit_into_the_solver():
    """dynoba_builder_emit -> dynoba_optimize: same LM run as the Problem ingested through the per-array calls"""
    from dynosam_b200.binding import Solver
    from dynosam_b200.builder import GraphBuilder
    p = synth.make_config("C1")
    b = GraphBuilder(); _feed(p, b)
    s = Solver(); b.emit(s)
    ref = Solver(p)
    a = s.optimize(max_iterations=6); r = ref.optimize(max_iterations=6)
    assert a["iterations"] == r["iterations"] and a["inner_iterations"] == r["inner_iterations"]
    assert abs(a["error_final"] - r["error_final"]) <= 1e-9*r["error_final"]
    s.close(); ref.close(); b.close()

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


def _feed_world_centric(p, b):
    """replay a synthetic WCME / WCPE problem as per-frame observations (tracklet / object ids from the generator's table)"""
    N = p.meta["n_frames"]; d = p.meta["dyn_obs"]
    ptp = [x for x in p.blocks if x.type == POSE2POINT3]
    odo = [x for x in p.blocks if x.type == BETWEEN6][-1]
    rel = {int(i[1]): m for i, m in zip(odo.idx, odo.meas)}
    for k in range(N):
        b.add_frame(k, p.pose[k], rel.get(k))
        sel = np.flatnonzero(ptp[0].idx[:, 0] == k)
        sel = sel[np.argsort(ptp[0].idx[sel, 1], kind="stable")]
        b.add_static(k, ptp[0].idx[sel, 1], ptp[0].meas[sel])
        sel = np.flatnonzero(d["frame"] == k)
        sel = sel[np.argsort(d["tracklet"][sel], kind="stable")]
        b.add_dynamic(k, d["tracklet"][sel], d["object"][sel], ptp[1].meas[sel])


@pytest.mark.parametrize("formulation", ["wcme", "wcpe"])
@pytest.mark.parametrize("cfg", [dict(n_frames=20, n_objects=1, n_static=300, n_dynamic=300, seed=42),
                                 dict(n_frames=40, n_objects=3, n_static=200, n_dynamic=600, seed=7, object_span=(12, 25))])
def test_builder_world_centric_formulations(formulation, cfg):
    """WCME (WorldMotionEstimator.cc:151-351) and WCPE (WorldPoseEstimator.cc:89-315) topology: point per (tracklet, frame),
    motion factor chains, smoothing, keys -- block by block against the independent numpy generator; initial values of the
    pose-like variables by the reference's own rules."""
    from dynosam_b200.builder import GraphBuilder
    from dynosam_b200.problem import MOTIONPOSE3, SMOOTH_POSE6, TERNARY3
    from oracle import oracle as O
    p = synth.make_problem(formulation=formulation, **cfg)
    N = p.meta["n_frames"]
    b = GraphBuilder(formulation=formulation)
    if formulation == "wcme":
        keys = p.pose_keys[N:]
        for m in range(N, p.n_pose):                               # front-end motions k-1 -> k
            b.set_motion_init(int((int(p.pose_keys[m]) >> 48) & 0xff) - ord('0'), int(p.pose_order[m]), p.pose[m])
    _feed_world_centric(p, b)
    q = b.problem()
    assert q.n_pose == p.n_pose and q.n_point == p.n_point and q.aux_pose.shape[0] == 0
    assert np.array_equal(q.pose_order, p.pose_order) and np.array_equal(q.pose_keys, p.pose_keys) and np.array_equal(q.point_keys, p.point_keys)
    assert np.array_equal(q.pose[:N], p.pose[:N])
    assert np.abs(q.point - p.point).max() <= 1e-12*max(np.abs(p.point).max(), 1.0)       # X_k z
    want = [POSE2POINT3, POSE2POINT3, TERNARY3, BETWEEN6, PRIOR6, BETWEEN6] if formulation == "wcme" else \
           [POSE2POINT3, POSE2POINT3, MOTIONPOSE3, SMOOTH_POSE6, PRIOR6, BETWEEN6]
    assert [x.type for x in q.blocks] == [x.type for x in p.blocks] == want
    for x, y in zip(q.blocks, p.blocks):
        assert np.array_equal(x.idx, y.idx), x.type
        assert (x.meas is None and y.meas is None) or np.array_equal(x.meas, y.meas), x.type
        assert np.array_equal(np.ravel(x.sigma), np.ravel(y.sigma)) and x.robust_k == y.robust_k, x.type
    if formulation == "wcme":
        # WorldMotionEstimator.cc:297-303: the front end's translation, identity rotation
        assert np.array_equal(q.pose[N:, 9:], p.pose[N:, 9:]) and np.array_equal(q.pose[N:, :9], np.tile(np.eye(3).ravel(), (p.n_pose - N, 1)))
    else:
        # WorldPoseEstimator.cc:205-232 without front-end motions: centroid of the object's points at that frame, identity rotation
        d = p.meta["dyn_obs"]; npt_s = p.meta["n_static"]
        for m in range(N, p.n_pose):
            obj = int((int(p.pose_keys[m]) >> 48) & 0xff) - ord('0'); fr = int(p.pose_order[m])
            sel = np.flatnonzero((d["object"] == obj) & (d["frame"] == fr))
            assert np.allclose(q.pose[m, 9:], q.point[npt_s + sel].mean(0), rtol=0, atol=1e-12) and np.array_equal(q.pose[m, :9], np.eye(3).ravel())
    q.pose[N:] = p.pose[N:]                                          # same initial values -> same graph error
    q.calib = p.calib
    assert abs(O.OracleProblem(q).error() - O.OracleProblem(p).error()) <= 1e-12*O.OracleProblem(p).error()
    b.close()


def test_builder_wcpe_pose_propagation():
    """WCPE initial object poses: L_k = motion(k-1 -> k) * L_k-1 when the front end gave a motion and L_k-1 exists
    (WorldPoseEstimator.cc:205-215), a given pose overrides, else the centroid."""
    from dynosam_b200.builder import GraphBuilder
    I = lie.identity()[0]
    b = GraphBuilder(formulation="wcpe")
    for k in range(4):
        b.add_frame(k, I, None if k == 0 else I)
        b.add_dynamic(k, [1, 2], [1, 1], [[1.0 + k, 0.0, 5.0], [3.0 + k, 0.0, 5.0]])
    H = lie.se3_exp(np.array([[0.0, 0.1, 0.0, 1.0, 0.0, 0.0]]))[0]
    L1 = lie.se3_exp(np.array([[0.0, 0.2, 0.0, 3.0, 0.0, 5.0]]))[0]
    b.set_keyframe_pose(1, 1, L1); b.set_motion_init(1, 2, H)
    q = b.problem()
    assert q.n_pose == 4 + 4 and q.n_point == 8
    assert np.allclose(q.pose[4][9:], [2.0, 0.0, 5.0])                       # frame 0: centroid
    assert np.array_equal(q.pose[5], L1)                                     # frame 1: given
    assert np.allclose(q.pose[6], lie.compose(H[None], L1[None])[0], atol=1e-15)     # frame 2: propagated
    assert np.allclose(q.pose[7][9:], [5.0, 0.0, 5.0])                       # frame 3: no motion -> centroid
    assert q.blocks[1].n == 6 and q.blocks[2].n == 2                         # 6 pose-motion factors, 2 smoothing triples
    b.close()


def test_builder_without_backtrack():
    """UpdateObservationParams::do_backtrack = false (RegularBackendModule.cc:139,197): a tracklet enters the graph at the frame
    its observation count reaches the minimum, with only what that update adds -- static: that frame's observation on
    (Formulation-impl.hpp:194-199); dynamic: the pair (previous, that frame) on (:703-720)."""
    from dynosam_b200.builder import GraphBuilder
    from dynosam_b200.problem import TERNARY3
    I = lie.identity()[0]
    X = [lie.se3_exp(np.array([[0.0, 0.01*k, 0.0, 0.1*k, 0.0, 0.5*k]]))[0] for k in range(6)]

    def feed(b):
        for k in range(6):
            b.add_frame(k, X[k], None if k == 0 else I)
        for k in (2, 3, 4):
            b.add_static(k, [7], [[0.1*k, 0.0, 4.0]])
        for k in (4, 5):
            b.add_static(k, [8], [[1.0, 0.2*k, 6.0]])
        b.add_static(5, [9], [[0.0, 0.0, 3.0]])                        # seen once: never enters
        for k in range(5):
            b.add_dynamic(k, [1, 2], [1, 1], [[1.0 + 0.1*k, 0.0, 5.0], [3.0 + 0.1*k, 0.0, 5.0]])
        b.add_dynamic(4, [3], [1], [[2.0, 1.0, 5.0]]); b.add_dynamic(5, [3], [1], [[2.1, 1.0, 5.0]])      # two observations: never enters

    b = GraphBuilder(backtrack=0); feed(b); q = b.problem()
    ptp = q.blocks[0]
    assert q.n_point == 2 + 2                                          # static 7, 8; dynamic 1, 2
    assert ptp.idx.tolist() == [[3, 0], [4, 0], [5, 1]]               # 7: frames 3, 4 (2 dropped); 8: frame 5 (4 dropped)
    assert np.allclose(q.point[0], lie.transform_from(X[3][None], np.array([[0.3, 0.0, 4.0]]))[0])     # X_3 z_3
    assert np.allclose(q.point[1], lie.transform_from(X[5][None], np.array([[1.0, 1.0, 6.0]]))[0])
    hyb = q.blocks[1]
    assert hyb.n == 2*4 and sorted(set(hyb.idx[:, 0].tolist())) == [1, 2, 3, 4]       # frame 0 dropped: the pair (1, 2) opens the tracklet
    assert q.n_pose == 6 + 4 and q.pose_order[6:].tolist() == [1, 2, 3, 4]            # motions only where factors exist; key-frame = frame 1
    assert np.allclose(q.aux_pose[0][9:], lie.transform_from(X[1][None], np.array([[2.1, 0.0, 5.0]]))[0])    # centroid of what entered at frame 1
    # m_L = L_e^-1 (e_H_e = I)^-1 X_1 z_1
    w = lie.transform_from(X[1][None], np.array([[1.1, 0.0, 5.0]]))[0]
    assert np.allclose(q.point[2], lie.transform_to(q.aux_pose[0][None], w[None])[0])
    b.close()
    # with backtracking (the default) everything of a tracklet that reaches the minimum enters
    b = GraphBuilder(); feed(b); q = b.problem()
    assert q.blocks[0].n == 3 + 2 and q.blocks[1].n == 2*5 and q.n_pose == 6 + 5
    b.close()
    # world-centric: points from the opening pair on
    b = GraphBuilder(formulation="wcme", backtrack=0); feed(b); q = b.problem()
    assert q.n_point == 2 + 2*4 and q.blocks[1].n == 8                # static 7, 8; dynamic points at frames 1..4 of tracklets 1, 2
    tern = [x for x in q.blocks if x.type == TERNARY3][0]
    assert tern.n == 2*3 and q.n_pose == 6 + 3 and q.pose_order[6:].tolist() == [2, 3, 4]
    b.close()

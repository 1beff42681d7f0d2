"""Timing of the batched flow+pose and object-motion refinements (8f-2): one frame's worth of objects (and a 400-problem batch) through the
C ABI with host buffers (allocation, H2D, the single launch, D2H inside the timed region).  Prints one JSON line."""
import json, sys, time
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_star import make_motion_problem, make_problem                                    # noqa: E402
from dynosam_b200 import binding                                      # noqa: E402

rng = np.random.default_rng(1)
out = {}
for name, nprob in (("frame_10_objects", 10), ("batch_400", 400)):
    probs = [make_problem(rng, int(n)) for n in rng.integers(80, 300, nprob)]
    kw = dict(flow_sigma=1.0, flow_prior_sigma=0.5, huber_k=1.0)
    binding.flow_pose_batch(probs, **kw)
    ts = []
    for _ in range(10):
        t = time.perf_counter(); r = binding.flow_pose_batch(probs, **kw); ts.append(time.perf_counter() - t)
    out["flow_pose_" + name] = dict(problems=nprob, features=int(sum(len(q["depth"]) for q in probs)), ms_per_call=1e3*float(np.median(ts)),
                     lm_iterations=int(sum(x["iterations"] for x in r)), problems_per_s=nprob/float(np.median(ts)))
    mprobs = [make_motion_problem(rng, int(n), outliers=0.1) for n in rng.integers(80, 300, nprob)]
    binding.motion_refine_batch(mprobs)
    ts = []
    for _ in range(10):
        t = time.perf_counter(); r = binding.motion_refine_batch(mprobs); ts.append(time.perf_counter() - t)
    out["motion_refine_" + name] = dict(problems=nprob, tracklets=int(sum(len(q["kp_prev"]) for q in mprobs)), ms_per_call=1e3*float(np.median(ts)),
                                        lm_iterations=int(sum(x["iterations"] for x in r)), problems_per_s=nprob/float(np.median(ts)))
print(json.dumps(out))

"""Timing of the batched flow+pose refinement (8f-2): one frame's worth of objects (and a 400-problem batch) through the
C ABI with host buffers (allocation, H2D, the single launch, D2H inside the timed region).  Prints one JSON line."""
import json, sys, time
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_star import make_problem                                    # noqa: E402
from dynosam_b200 import binding                                      # noqa: E402

rng = np.random.default_rng(1)
out = {}
for name, nprob in (("frame_10_objects", 10), ("batch_400", 400)):
    probs = [make_problem(rng, int(n)) for n in rng.integers(80, 300, nprob)]
    binding.flow_pose_batch(probs, 1.0, 0.5, 1.0, max_iterations=10)
    ts = []
    for _ in range(10):
        t = time.perf_counter(); r = binding.flow_pose_batch(probs, 1.0, 0.5, 1.0, max_iterations=10); ts.append(time.perf_counter() - t)
    out[name] = dict(problems=nprob, features=int(sum(len(q["depth"]) for q in probs)), ms_per_call=1e3*float(np.median(ts)),
                     lm_iterations=int(sum(x["iterations"] for x in r)), problems_per_s=nprob/float(np.median(ts)))
print(json.dumps(out))

"""Timing of the batched flow+pose and object-motion refinements (8f-2): one frame's worth of objects and a 400-problem batch
through the C ABI with host buffers (H2D, the single launch, D2H inside the timed region).  Prints one JSON line."""
import json, os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynosam_b200.synth_star import make_flow_pose_problem as make_problem, make_motion_problem    # noqa: E402
from dynosam_b200 import binding                                                           # noqa: E402

rng = np.random.default_rng(1)
kw = dict(flow_sigma=1.0, flow_prior_sigma=0.5, huber_k=1.0)
sets = {}
for name, nprob in (("frame_10_objects", 10), ("batch_400", 400)):
    sets[name] = ([make_problem(rng, int(n), outliers=0.05) for n in rng.integers(80, 300, nprob)],
                  [make_motion_problem(rng, int(n), outliers=0.1) for n in rng.integers(80, 300, nprob)])
t0 = time.perf_counter()
while time.perf_counter() - t0 < 0.5:                                                      # bring the clocks up before timing the short calls
    binding.flow_pose_batch(sets["batch_400"][0], **kw)
out = {}
for name, (fp, mr) in sets.items():
    reps = 40 if len(fp) <= 10 else 8
    for label, fn, probs, args in (("flow_pose_", binding.flow_pose_batch, fp, kw), ("motion_refine_", binding.motion_refine_batch, mr, {})):
        fn(probs, **args)
        ts = []
        for _ in range(reps):
            t = time.perf_counter(); r = fn(probs, **args); ts.append(time.perf_counter() - t)
        units = int(sum(len(q["depth"]) if "depth" in q else len(q["kp_prev"]) for q in probs))
        out[label + name] = dict(problems=len(probs), features=units, ms_per_call=1e3*float(np.median(ts)), ms_min=1e3*float(np.min(ts)),
                                 lm_iterations=int(sum(x["iterations"] for x in r)), problems_per_s=len(probs)/float(np.median(ts)))
try:
    out["clocks"] = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader"], capture_output=True, text=True, timeout=10).stdout.strip()
except Exception as e:                                                                     # noqa: BLE001
    out["clocks"] = str(e)
print(json.dumps(out))

#!/usr/bin/env python
"""bench_frontend.py -- front-end frames/s at 1242x375 (BASELINE.json second metric, config C4).

Per frame (what FeatureTracker::track does on the hot path, dynosam/src/frontend/vision/FeatureTracker.cc:73-192):
upload flow / instance mask / gray, propogateMask, trackDynamic over the previous dynamic features, sampleDynamic
candidate scan, and pyramidal KLT of the static features (forward 21x21 L3 + backward check L5 as in
StaticFeatureTracker.cc:486-534).  CPU arm (--impl reference): cv2.calcOpticalFlowPyrLK on all host threads plus the
literal numpy restatement of the dense passes on a bounded number of frames.
    python bench_frontend.py [--frames N] [--impl dynoba|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from dynosam_b200.synth_frames import SyntheticStream, W, H  # noqa: E402

N_STATIC, PER_OBJECT = 800, 200          # params/FrontendParams.yaml:50-66


def make_inputs(n_frames, seed=42):
    st = SyntheticStream(n_objects=10, seed=seed)
    rng = np.random.default_rng(seed)
    frames = [st.frame(k) for k in range(n_frames + 1)]
    static_pts = np.stack([rng.uniform(20, W - 20, N_STATIC), rng.uniform(20, H - 20, N_STATIC)], 1).astype(np.float32)
    feats = []
    for k in range(n_frames):
        _, m0, f0 = frames[k]
        kps, labs = [], []
        for lab in range(1, 11):
            ys, xs = np.nonzero(m0 == lab)
            if len(ys) == 0:
                continue
            sel = rng.choice(len(ys), size=min(PER_OBJECT, len(ys)), replace=False)
            kps.append(np.stack([xs[sel] + 0.5 + f0[ys[sel], xs[sel], 0], ys[sel] + 0.5 + f0[ys[sel], xs[sel], 1]], 1))
            labs.append(np.full(len(sel), lab, np.int32))
        kp = np.concatenate(kps) if kps else np.zeros((0, 2)); lab = np.concatenate(labs) if labs else np.zeros(0, np.int32)
        ok = (kp[:, 0] > 1) & (kp[:, 0] < W - 1) & (kp[:, 1] > 1) & (kp[:, 1] < H - 1)
        feats.append((kp[ok], lab[ok], rng.integers(0, 21, ok.sum()).astype(np.int32), np.arange(ok.sum(), dtype=np.int64)))
    return frames, static_pts, feats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dynoba", choices=["dynoba", "reference"])
    ap.add_argument("--call-by-call", action="store_true", help="round-1 path: every image re-uploaded by every call")
    args = ap.parse_args()
    n = args.frames
    frames, static_pts, feats = make_inputs(max(n, args.warmup) + 1)
    config = {"workload": f"C4: 1242x375 synthetic stream, {n} frames, 10 objects, {N_STATIC} static KLT points (fwd L3 + bwd L5), "
                          f"<= {PER_OBJECT} dynamic features/object", "data": "synthetic"}
    per_frame_bytes = 21*W*H + 2*W*H          # dense passes ~21 B/px + two gray uploads
    if args.impl == "reference":
        import cv2
        from oracle import frontend_oracle as FO
        nb = min(n, 20); t0 = time.perf_counter()
        for k in range(1, nb + 1):
            g0 = frames[k-1][0]; g1, m1, f1 = frames[k]
            p1, st, _ = FO.klt_track(g0, g1, static_pts, 21, 3, 30, 0.03)
            FO.klt_track(g1, g0, p1, 21, 5, 30, 0.01)
        t_klt = (time.perf_counter() - t0)/nb
        # dense passes: vectorised numpy equivalent of the literal loops (the literal Python loops take minutes/frame)
        t0 = time.perf_counter()
        for k in range(1, nb + 1):
            _, m1, f1 = frames[k]
            ok = (m1 != 0) & (f1[..., 0] != 0) & (f1[..., 1] != 0)
            [np.flatnonzero(ok & (m1 == o)) for o in range(1, 11)]
            kp, lab, age, tid = feats[k-1]
            x = kp[:, 0].astype(int); y = kp[:, 1].astype(int); _ = (m1[y, x] == lab) & (f1[y, x, 0] != 0)
        t_dense = (time.perf_counter() - t0)/nb
        fps = 1.0/(t_klt + t_dense)
        print(json.dumps({"impl": "reference", "metric": "frontend fps at 1242x375", "value": fps, "unit": "frames/s", "n_gpus": 1,
                          "higher_is_better": True, "dtype": "u8/int16/fp32 (OpenCV)", "config": config,
                          "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference",
                                           "sample": f"{nb} frames: cv2.calcOpticalFlowPyrLK fwd+bwd ({1e3*t_klt:.1f} ms) + numpy dense passes ({1e3*t_dense:.1f} ms)"}}))
        return
    print(json.dumps(run_dynoba(n, args.warmup, streaming=not args.call_by_call)))


def run_dynoba(n, warmup, streaming=True):
    """front-end frames/s through the C ABI (libdynofront) over n frames of the synthetic stream.
    streaming: frames arrive in pinned host buffers, only the NEW frame crosses PCIe (dynofront_next_frame), the previous
    frame's flow / mask / gray pyramid stay on the device, the forward-backward KLT with its checks is one call and the
    propagated mask stays on the device.  Otherwise: the call-by-call path of round 1 (every image re-uploaded per call)."""
    frames, static_pts, feats = make_inputs(max(n, warmup) + 1)
    config = {"workload": f"C4: 1242x375 synthetic stream, {n} frames, 10 objects, {N_STATIC} static KLT points (fwd L3 + bwd L5 + round-trip / label / border checks), "
                          f"<= {PER_OBJECT} dynamic features/object", "data": "synthetic", "mode": "streaming (resident previous frame, pinned uploads)" if streaming else "call by call"}
    import torch  # noqa: F401  (device selection / presence check only)
    from dynosam_b200.frontend import FeatureTrackerGPU, TrackParams
    t = FeatureTrackerGPU(W, H); prm = TrackParams()
    static_age = np.zeros(N_STATIC, np.int32)
    if streaming:
        frames = [(np.ascontiguousarray(g, np.uint8), np.ascontiguousarray(m, np.int32), np.ascontiguousarray(f, np.float32)) for g, m, f in frames]
        for g, m, f in frames:
            t.pin(g); t.pin(m); t.pin(f)
        t.next_frame(frames[0][0], frames[0][2], frames[0][1])
        def one(k):
            g1, m1, f1 = frames[k]
            kp, lab, age, tid = feats[k-1]
            t.next_frame(g1, f1, m1)                                   # H2D of the new frame only
            t.propagate_mask_resident(kp, lab, prm)
            acc, *_ = t.track_dynamic(kp, lab, age, tid, prm, 10**6, want_masks=False)
            cand, _ = t.sample_candidates(list(range(1, 11)), prm, capacity=W*H//2)
            p1, st, back, keep = t.klt_track_fb(None, None, static_pts, prm, static_age, 25)
            return t.last_ms, int(acc.sum()), int(keep.sum())
        h2d = W*H*(1 + 8 + 4)
    else:
        def one(k):
            g0 = frames[k-1][0]; _, m0, f0 = frames[k-1]; g1, m1, f1 = frames[k]
            kp, lab, age, tid = feats[k-1]
            cur = t.propagate_mask(kp, lab, m0, f0, m1, prm)
            t.set_frame(f1, cur, None)
            acc, *_ = t.track_dynamic(kp, lab, age, tid, prm, 10**6, want_masks=False)
            cand, _ = t.sample_candidates(list(range(1, 11)), prm, capacity=W*H//2)
            p1, st, _ = t.klt_track(g0, g1, static_pts, 21, 3, 30, 0.03); ms = t.last_ms
            p0, st2, _ = t.klt_track(g1, g0, p1, 21, 5, 30, 0.01)
            return ms + t.last_ms, int(acc.sum()), int(st.sum())
        h2d = W*H*(2*(8 + 4) + 4 + 4)
    for k in range(1, warmup + 1):
        one(k)
    klt_ms = []; t0 = time.perf_counter()
    for k in range(1, n + 1):
        ms, na, ns = one(k); klt_ms.append(ms)
    dt = time.perf_counter() - t0
    fps = n/dt
    dense_bytes = 21*W*H                    # flow 8 + prev mask 4 + cur mask 4 + det mask 1 + scatter 4 per pixel
    out = {"metric": "frontend fps at 1242x375", "value": fps, "unit": "frames/s", "n_gpus": 1, "frames": n, "ms_per_frame": 1e3*dt/n,
           "higher_is_better": True, "dtype": "u8/int16/int32 fixed point + fp32 (OpenCV semantics)", "config": config,
           "klt_device_ms_per_frame": float(np.mean(klt_ms)), "h2d_bytes_per_frame": int(h2d),
           "note": "end to end through the C ABI from host images (H2D/D2H inside the timed region); a frame moves a few MB and "
                   "launches ~40 small kernels, so it is PCIe / launch bound, not HBM bound: the dense passes' 21 B/px "
                   f"({dense_bytes/1e6:.1f} MB) would take {dense_bytes/6.57e12*1e6:.1f} us at the measured HBM rate"}
    return out


if __name__ == "__main__":
    main()

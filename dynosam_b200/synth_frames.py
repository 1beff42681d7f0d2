"""Seeded synthetic RGB-D-like stream for the front-end rows: procedurally textured
background (value noise), moving textured rectangles as objects, instance mask CV_32S, exact per-pixel flow CV_32FC2."""
from __future__ import annotations

import numpy as np

W, H = 1242, 375


def _value_noise(rng, h, w, cells):
    g = rng.uniform(0, 255, (h//cells + 2, w//cells + 2))
    ys = np.arange(h)/cells; xs = np.arange(w)/cells
    y0 = ys.astype(int); x0 = xs.astype(int); fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
    return a*(1 - fx)*(1 - fy) + b*fx*(1 - fy) + c*(1 - fx)*fy + d*fx*fy


class SyntheticStream:
    def __init__(self, n_objects=10, seed=42, width=W, height=H):
        self.rng = np.random.default_rng(seed); self.W, self.H = width, height
        rng = self.rng
        self.bg = (0.5*_value_noise(rng, height, width + 400, 8) + 0.3*_value_noise(rng, height, width + 400, 3)
                   + 0.2*rng.uniform(0, 255, (height, width + 400)))
        self.cam_v = 1.5                                     # background scroll, px / frame
        self.obj = []
        for j in range(n_objects):
            w = int(rng.integers(60, 160)); h = int(rng.integers(40, 110))
            tex = 0.6*_value_noise(rng, h, w, 5) + 0.4*rng.uniform(0, 255, (h, w))
            self.obj.append(dict(label=j + 1, w=w, h=h, tex=tex, x=float(rng.uniform(0, width - w)), y=float(rng.uniform(0, height - h)),
                                 vx=float(rng.uniform(-3, 3)), vy=float(rng.uniform(-1, 1))))

    def frame(self, k):
        """gray uint8, mask int32, flow float32[H,W,2] (flow from frame k to k+1)."""
        W_, H_ = self.W, self.H
        off = self.cam_v*k
        i0 = int(np.floor(off)); f = off - i0
        wb = self.bg.shape[1]                               # the background strip wraps around: streams of any length
        c0 = (i0 + np.arange(W_)) % wb; c1 = (i0 + 1 + np.arange(W_)) % wb
        gray = (1 - f)*self.bg[:, c0] + f*self.bg[:, c1]
        mask = np.zeros((H_, W_), np.int32)
        flow = np.zeros((H_, W_, 2), np.float32); flow[..., 0] = -self.cam_v; flow[..., 1] = 1e-3   # no exact zeros in the background
        for o in self.obj:
            x = o["x"] + o["vx"]*k; y = o["y"] + o["vy"]*k
            x = (x + o["w"]) % (W_ + o["w"]) - o["w"]; y = (y + o["h"]) % (H_ + o["h"]) - o["h"]   # objects re-enter on the other side
            xi, yi = int(round(x)), int(round(y))
            x0, y0 = max(xi, 0), max(yi, 0); x1, y1 = min(xi + o["w"], W_), min(yi + o["h"], H_)
            if x1 <= x0 or y1 <= y0:
                continue
            gray[y0:y1, x0:x1] = o["tex"][y0 - yi:y1 - yi, x0 - xi:x1 - xi]
            mask[y0:y1, x0:x1] = o["label"]
            flow[y0:y1, x0:x1, 0] = o["vx"]; flow[y0:y1, x0:x1, 1] = o["vy"] if o["vy"] != 0 else 1e-3
        return np.clip(gray, 0, 255).astype(np.uint8), mask, flow

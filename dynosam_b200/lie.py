"""Batched SE(3)/SO(3) helpers in numpy (host side: synthetic generator, value packing).

Pose layout everywhere: 12 doubles = R row-major (9) followed by t (3).  Tangent order is
[omega; v] with right-multiplied retraction T*Exp(xi) (GTSAM 4.2 built with POSE3_EXPMAP,
reference docker/Dockerfile.amd64:112).
"""
from __future__ import annotations

import numpy as np

EPS = np.finfo(np.float64).eps


def skew(v):
    v = np.asarray(v, dtype=np.float64)
    out = np.zeros(v.shape[:-1] + (3, 3))
    out[..., 0, 1] = -v[..., 2]; out[..., 0, 2] = v[..., 1]
    out[..., 1, 0] = v[..., 2]; out[..., 1, 2] = -v[..., 0]
    out[..., 2, 0] = -v[..., 1]; out[..., 2, 1] = v[..., 0]
    return out


def so3_exp(w):
    w = np.asarray(w, dtype=np.float64).reshape(-1, 3)
    th2 = np.einsum('ni,ni->n', w, w)
    W = skew(w)
    near = th2 <= EPS
    th = np.sqrt(np.where(near, 1.0, th2))
    K = W/th[:, None, None]
    s = np.sin(th); s2 = np.sin(0.5*th); omc = 2.0*s2*s2
    R = np.eye(3)[None] + s[:, None, None]*K + omc[:, None, None]*(K @ K)
    R[near] = np.eye(3)[None] + W[near]
    return R


def se3_exp(xi):
    xi = np.asarray(xi, dtype=np.float64).reshape(-1, 6)
    w, v = xi[:, :3], xi[:, 3:]
    R = so3_exp(w)
    th2 = np.einsum('ni,ni->n', w, w)
    big = th2 > EPS
    c = np.cross(w, v)
    wv = np.einsum('ni,ni->n', w, v)
    t = (c - np.einsum('nij,nj->ni', R, c) + w*wv[:, None])/np.where(big, th2, 1.0)[:, None]
    t = np.where(big[:, None], t, v)
    return pack(R, t)


def pack(R, t):
    R = np.asarray(R, dtype=np.float64).reshape(-1, 9)
    t = np.asarray(t, dtype=np.float64).reshape(-1, 3)
    return np.concatenate([R, t], axis=1)


def rot(P):
    return np.asarray(P, dtype=np.float64).reshape(-1, 12)[:, :9].reshape(-1, 3, 3)


def trans(P):
    return np.asarray(P, dtype=np.float64).reshape(-1, 12)[:, 9:]


def identity(n=1):
    return pack(np.tile(np.eye(3), (n, 1, 1)), np.zeros((n, 3)))


def compose(a, b):
    Ra, Rb = rot(a), rot(b)
    return pack(Ra @ Rb, np.einsum('nij,nj->ni', Ra, trans(b)) + trans(a))


def inverse(a):
    Rt = np.swapaxes(rot(a), 1, 2)
    return pack(Rt, -np.einsum('nij,nj->ni', Rt, trans(a)))


def between(a, b):
    return compose(inverse(a), b)


def transform_from(P, p):
    return np.einsum('nij,nj->ni', rot(P), np.asarray(p, dtype=np.float64).reshape(-1, 3)) + trans(P)


def transform_to(P, p):
    d = np.asarray(p, dtype=np.float64).reshape(-1, 3) - trans(P)
    return np.einsum('nji,nj->ni', rot(P), d)


def retract(P, xi):
    return compose(P, se3_exp(xi))


def ypr(y, p, r):
    """gtsam::Rot3::Ypr(y,p,r) = Rz(y) * Ry(p) * Rx(r)."""
    cy, sy, cp, sp, cr, sr = np.cos(y), np.sin(y), np.cos(p), np.sin(p), np.cos(r), np.sin(r)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    Ry = np.array([[cp, 0, sp], [0, 1.0, 0], [-sp, 0, cp]])
    Rx = np.array([[1.0, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


def rodrigues(wx, wy, wz):
    """gtsam::Rot3::Rodrigues(wx,wy,wz) = Expmap."""
    return so3_exp(np.array([[wx, wy, wz]]))[0]


def pose(R, t):
    return pack(np.asarray(R).reshape(1, 3, 3), np.asarray(t).reshape(1, 3))[0]

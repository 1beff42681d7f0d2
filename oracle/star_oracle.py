"""CPU restatement of the front end's joint optical-flow + pose refinement (TEST INFRASTRUCTURE, NOT PRODUCT).

Follows OpticalFlowAndPoseOptimizer::optimize (dynosam/include/dynosam/frontend/vision/MotionSolver-inl.hpp:88-260): per
feature a Pose3FlowProjectionFactor (Robust Huber over Isotropic flow_sigma) and a PriorFactor<Point2> on the flow
(Isotropic flow_prior_sigma), one Pose3 unknown, gtsam::LevenbergMarquardtOptimizer with maxIterations 10.  The factor
residual / Jacobian come from the C oracle (orc_linearize_block of the FLOWPROJ2 block: Pose3FlowProjectionFactor.h:73-133,
pinned by tests/test_oracle_kat.py); the damped normal equations are formed DENSE here and solved by Cholesky, and the
outer loop is LevenbergMarquardtOptimizer::iterate / tryLambda written out literally.
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

from dynosam_b200.problem import FLOWPROJ2, FactorBlock, Problem

from . import oracle as orc

DEFAULTS = dict(lambda_initial=1e-5, lambda_factor=10.0, lambda_upper_bound=1e5, lambda_lower_bound=0.0,
                min_model_fidelity=1e-3, relative_error_tol=1e-5, absolute_error_tol=1e-5, error_tol=0.0,
                max_iterations=100)


def flow_pose_lm(pose_init, pose_prev, calib5, kp_prev, depth, flow, flow_sigma, flow_prior_sigma, huber_k, **kw):
    P = dict(DEFAULTS); P.update(kw)
    n = len(depth)
    kp_prev = np.asarray(kp_prev, dtype=np.float64).reshape(n, 2); flow0 = np.asarray(flow, dtype=np.float64).reshape(n, 2)
    meas = np.concatenate([kp_prev, np.asarray(depth, dtype=np.float64).reshape(n, 1), np.tile(np.asarray(pose_prev).reshape(1, 12), (n, 1))], 1)
    blk = FactorBlock(FLOWPROJ2, np.stack([np.arange(n), np.zeros(n, dtype=int)], 1), meas, np.array([flow_sigma]), float(huber_k))
    prob = Problem(np.asarray(pose_init, dtype=np.float64).reshape(1, 12), np.zeros((0, 3)), flow=flow0.copy(),
                   calib=np.concatenate([np.asarray(calib5, dtype=np.float64), [0.0]]), blocks=[blk])
    o = orc.OracleProblem(prob)
    isp = 1.0/flow_prior_sigma

    def error():
        return (float(o.error_block(0).sum()) if n else 0.0) + 0.5*float((((o.flow - flow0)*isp)**2).sum())

    def linear_system():
        A = np.zeros((4*n, 2*n + 6)); b = np.zeros(4*n)
        if n:
            Af, bf = o.linearize_block(0)          # whitened, Huber-weighted: [n,2,8] (flow 2 | pose 6), [n,2]
            for i in range(n):
                A[2*i:2*i + 2, 2*i:2*i + 2] = Af[i, :, 0:2]; A[2*i:2*i + 2, 2*n:] = Af[i, :, 2:8]; b[2*i:2*i + 2] = bf[i]
                A[2*n + 2*i:2*n + 2*i + 2, 2*i:2*i + 2] = isp*np.eye(2); b[2*n + 2*i:2*n + 2*i + 2] = -(o.flow[i] - flow0[i])*isp
        return A, b

    err = error(); err0 = err
    lam = P["lambda_initial"]; iterations = inner = 0
    if not (err <= P["error_tol"]) and P["max_iterations"] > 0:
        new_error = err
        while True:
            current_error = new_error
            A, b = linear_system()
            H = A.T @ A; g = A.T @ b
            while True:      # tryLambda
                solved = True
                try:
                    L = np.linalg.cholesky(H + lam*np.eye(H.shape[0]))
                    delta = np.linalg.solve(L.T, np.linalg.solve(L, g))
                except np.linalg.LinAlgError:
                    solved = False
                success = stop = False; nerr = np.inf
                if solved:
                    old_lin = 0.5*float(b @ b); rr = A @ delta - b; lin = old_lin - 0.5*float(rr @ rr)
                    if np.isfinite(lin) and lin >= 0:
                        keep_pose, keep_flow = o.pose.copy(), o.flow.copy()
                        o.pose[0] = orc.se3_retract(o.pose[0], delta[2*n:]); o.flow += delta[:2*n].reshape(n, 2)
                        nerr = error(); cost = err - nerr
                        if lin > np.finfo(float).eps*old_lin:
                            success = cost/lin > P["min_model_fidelity"]
                        if abs(cost) < P["relative_error_tol"]*err:
                            stop = True
                        if not success:
                            o.pose[:] = keep_pose; o.flow[:] = keep_flow
                if success:
                    lam = max(P["lambda_lower_bound"], lam/P["lambda_factor"]); err = nerr; iterations += 1; inner += 1
                    break
                elif not stop:
                    lam *= P["lambda_factor"]; inner += 1
                    if lam >= P["lambda_upper_bound"]:
                        break
                else:
                    break
            new_error = err
            with np.errstate(all="ignore"):
                rel = np.float64(current_error - new_error)/np.float64(current_error)
            done = (new_error <= P["error_tol"]) or (P["relative_error_tol"] != 0.0 and rel <= P["relative_error_tol"]) \
                or ((current_error - new_error) <= P["absolute_error_tol"])
            if not (iterations < P["max_iterations"] and not done and np.isfinite(current_error)):
                break
    return dict(pose=o.pose[0].copy(), flow=o.flow.copy(), error_initial=err0, error_final=err, iterations=iterations, inner_iterations=inner)

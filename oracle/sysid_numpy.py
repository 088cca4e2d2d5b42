"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the LMPC model identification
`LMPCRacingGame.estimate_ABC` (car_racing/utils/base.py:585-622) =
N x `lmpc_helper.regression_and_linearization` (car_racing/control/lmpc_helper.py:26-201).

PINNED: tests/golden/sysid_golden.npz holds outputs of the reference function itself
(tests/golden/make_sysid_golden.py; cvxopt's unconstrained qp answered by numpy.linalg.solve);
tests/test_sysid.py checks this restatement against them.

Per horizon stage i (lmpc_helper.py:40-131):
  * query z0 = (vx, vy, wz, delta, a) of (lin_points[i], lin_input[i])                              (:47-59)
  * per used lap: l1 norm of (data - z0) * diag(0.1, 1, 1, 1, 1) over rows 0..time_ss-2              (:216-230)
    keep rows with norm < h = 5; if at least max_num_point of them, the max_num_point smallest
    (argsort order), else all in row order (:231-235); weights K = 3/4 (1 - (norm/h)^2)              (:237)
  * three weighted least squares  min sum K (y - [f, 1] theta)^2  via the normal equations Q theta = M' K y
    (:249-279, :343-366): vx+ on (vx, vy, wz, a); vy+ and wz+ on (vx, vy, wz, delta)
  * rows epsi, s, ey: analytic linearisation of the Frenet kinematics at lin_points[i] with the track
    curvature at s (:135-199), including the reference's `den * 2` in ds/dey (:178).
"""
import numpy as np

H_BANDWIDTH = 5.0
SCALING = np.array([0.1, 1.0, 1.0, 1.0, 1.0])


def curvature(point_and_tangent, lap_length, s):
    """racing_env.get_curvature (utils/racing_env.py:225-246)."""
    while s > lap_length:
        s -= lap_length
    while s < 0:
        s += lap_length
    pt = point_and_tangent
    hit = np.nonzero((s >= pt[:, 3]) & (s <= pt[:, 3] + pt[:, 4]))[0]
    return float(pt[int(hit[0]), 5])


def select(data, z0, max_num_point, h=H_BANDWIDTH):
    norm = np.abs((data - z0[None, :]) * SCALING[None, :]).sum(axis=1)
    inside = np.nonzero(norm < h)[0]
    idx = np.argsort(norm)[:max_num_point] if inside.shape[0] >= max_num_point else inside
    return idx, 0.75 * (1.0 - (norm[idx] / h) ** 2)


def stage_model(x0, u0, ss, us, time_ss, used_laps, point_and_tangent, dt, max_num_point=40):
    """One stage: returns A (6,6), B (6,2), C (6,), [indices per lap]."""
    A = np.zeros((6, 6)); B = np.zeros((6, 2)); C = np.zeros(6)
    z0 = np.array([x0[0], x0[1], x0[2], u0[0], u0[1]])
    rows, wts, nxt, sel = [], [], [], []
    for lap in used_laps:
        T = int(time_ss[lap])
        data = np.hstack((ss[:T - 1, 0:3, lap], us[:T - 1, :, lap]))
        idx, K = select(data, z0, max_num_point)
        sel.append(idx)
        rows.append(data[idx]); wts.append(K); nxt.append(ss[idx + 1, 0:3, lap])
    Z = np.vstack(rows); K = np.concatenate(wts); Y = np.vstack(nxt)
    one = np.ones((Z.shape[0], 1))
    for y_index, feat, ucol in ((0, [0, 1, 2, 4], 1), (1, [0, 1, 2, 3], 0), (2, [0, 1, 2, 3], 0)):
        M = np.hstack((Z[:, feat], one))
        Q = M.T @ (K[:, None] * M)
        th = np.linalg.solve(Q, M.T @ (K * Y[:, y_index]))
        A[y_index, 0:3] = th[0:3]
        B[y_index, ucol] = th[3]
        C[y_index] = th[4]
    vx, vy, wz, epsi, s, ey = x0
    cur = curvature(point_and_tangent, point_and_tangent[-1, 3] + point_and_tangent[-1, 4], s)
    den = 1.0 - cur * ey
    ce, se = np.cos(epsi), np.sin(epsi)
    A[3] = [-dt * ce / den * cur, dt * se / den * cur, dt, 1 - dt * (-vx * se - vy * ce) / den * cur, 0.0,
            dt * (vx * ce - vy * se) / den ** 2 * cur * (-cur)]
    C[3] = epsi + dt * (wz - (vx * ce - vy * se) / den * cur) - A[3] @ x0
    A[4] = [dt * ce / den, -dt * se / den, 0.0, dt * (-vx * se - vy * ce) / den, 1.0,
            -dt * (vx * ce - vy * se) / (den * 2) * (-cur)]                     # `den * 2`: reference quirk (:178)
    C[4] = s + dt * ((vx * ce - vy * se) / den) - A[4] @ x0
    A[5] = [dt * se, dt * ce, 0.0, dt * (vx * ce - vy * se), 0.0, 1.0]
    C[5] = ey + dt * (vx * se + vy * ce) - A[5] @ x0
    return A, B, C, sel


def estimate_abc(lin_points, lin_input, ss, us, time_ss, used_laps, point_and_tangent, dt, max_num_point=40):
    """Batched estimate_ABC: lin_points (Bn,N+1,6) or (N+1,6); lin_input (Bn,N,2).  Returns A (Bn,N,6,6), B (Bn,N,6,2),
    C (Bn,N,6), idx (Bn,N,len(used_laps),max_num_point) int (-1 padded)."""
    lp = np.asarray(lin_points, float); li = np.asarray(lin_input, float)
    if lp.ndim == 2:
        lp, li = lp[None], li[None]
    Bn, N = li.shape[0], li.shape[1]
    A = np.zeros((Bn, N, 6, 6)); B = np.zeros((Bn, N, 6, 2)); C = np.zeros((Bn, N, 6))
    idx = -np.ones((Bn, N, len(used_laps), max_num_point), dtype=np.int64)
    for b in range(Bn):
        for i in range(N):
            A[b, i], B[b, i], C[b, i], sel = stage_model(lp[b, i], li[b, i], ss, us, time_ss, used_laps, point_and_tangent, dt,
                                                         max_num_point)
            for j, s_ in enumerate(sel):
                idx[b, i, j, :len(s_)] = s_
    return A, B, C, idx

"""TEST INFRASTRUCTURE -- numpy restatement of a dynamic rival's prediction (not on the product path).

Follows offboard.DynamicBicycleModel.get_estimation / get_trajectory_nsteps (car_racing/racing/offboard.py:51-94) with
racing_env.get_curvature (car_racing/utils/racing_env.py:225-246): n explicit-Euler steps of the zero-input Frenet
kinematics, s wrapped into the lap after every step.  Pinned: tests/golden/rollout_golden.npz holds the unmodified
reference's output on three tracks (tests/golden/make_rollout_golden.py).  Only tests/ may import this module.
"""
import numpy as np


def curvature(lap_length, segments, s):
    """racing_env.py:225-246; segments = point_and_tangent[:, 3:6] (start s, length, curvature)."""
    while s > lap_length:
        s = s - lap_length
    while s < 0:
        s = s + lap_length
    hit = np.where((s >= segments[:, 0]) & (s <= segments[:, 0] + segments[:, 1]))[0]
    return segments[int(hit[0]), 2]


def rollout(xcurv, xglob, segments, lap_length, timestep, n):
    """offboard.py:80-94.  Returns (xcurv_nsteps (6,n), xglob_nsteps (6,n))."""
    xc_n, xg_n = np.zeros((6, n)), np.zeros((6, n))
    xc, xg = np.asarray(xcurv, float), np.asarray(xglob, float)
    for index in range(n):
        curv = curvature(lap_length, segments, xc[4])                          # offboard.py:53
        e, g = np.zeros(6), np.zeros(6)
        e[0:3] = xc[0:3]
        e[3] = xc[3] + timestep * (xc[2] - (xc[0] * np.cos(xc[3]) - xc[1] * np.sin(xc[3])) / (1 - curv * xc[5]) * curv)
        e[4] = xc[4] + timestep * ((xc[0] * np.cos(xc[3]) - xc[1] * np.sin(xc[3])) / (1 - curv * xc[5]))
        e[5] = xc[5] + timestep * (xc[0] * np.sin(xc[3]) + xc[1] * np.cos(xc[3]))
        g[0:3] = xg[0:3]
        g[3] = xg[3] + timestep * (xg[2])
        # offboard.py:71-76 assigns xglob_est[4] twice (the X update is overwritten by the Y update) and never fills [5]
        g[4] = xg[4] + timestep * (xg[0] * np.sin(xg[3]) + xg[1] * np.cos(xg[3]))
        while e[4] > lap_length:                                               # :89-90
            e[4] = e[4] - lap_length
        xc_n[:, index], xg_n[:, index] = e, g
        xc, xg = e, g
    return xc_n, xg_n

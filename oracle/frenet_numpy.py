"""TEST INFRASTRUCTURE -- numpy restatement of the curvilinear -> global-frame conversion (not on the product path).

Follows racing_env.get_global_position / get_orientation (car_racing/utils/racing_env.py:6-127) with wrap / sign
(:268-283).  Pinned: tests/golden/frenet_golden.npz holds the reference's output on four tracks
(tests/golden/make_frenet_golden.py).  Only tests/ may import this module.
"""
import numpy as np


def _segment(lap_length, pat, s):
    while s > lap_length:                                      # :12-15
        s = s - lap_length
    while s < 0:
        s = s + lap_length
    hit = np.where((s >= pat[:, 3]) & (s < pat[:, 3] + pat[:, 4] + 0.001))[0]     # s_tolerance (:11, 18-25)
    return s, int(hit[0])


def curv_to_glob(lap_length, pat, s, ey):
    """Returns (x, y, psi) of one point; pat = point_and_tangent (rows: x, y, psi, start s, length, curvature)."""
    s, i = _segment(lap_length, pat, s)
    if pat[i, 5] == 0.0:                                       # straight segment (:27-41)
        xf, yf, xs, ys, psi = pat[i, 0], pat[i, 1], pat[i - 1, 0], pat[i - 1, 1], pat[i, 2]
        deltaL, reltaL = pat[i, 4], s - pat[i, 3]
        x = (1 - reltaL / deltaL) * xs + reltaL / deltaL * xf + ey * np.cos(psi + np.pi / 2)
        y = (1 - reltaL / deltaL) * ys + reltaL / deltaL * yf + ey * np.sin(psi + np.pi / 2)
        return x, y, psi
    r = 1 / pat[i, 5]                                          # arc (:42-68)
    ang = pat[i - 1, 2]
    direction = 1 if r >= 0 else -1
    cx = pat[i - 1, 0] + np.abs(r) * np.cos(ang + direction * np.pi / 2)
    cy = pat[i - 1, 1] + np.abs(r) * np.sin(ang + direction * np.pi / 2)
    span = (s - pat[i, 3]) / (np.pi * np.abs(r)) * np.pi
    an = direction * np.pi / 2 + ang
    an = 2 * np.pi + an if an < -np.pi else (an - 2 * np.pi if an > np.pi else an)     # wrap (:268-275)
    angle = -(np.pi - np.abs(an)) * (1 if an >= 0 else -1)                              # sign (:278-283)
    x = cx + (np.abs(r) - direction * ey) * np.cos(angle + direction * span)
    y = cy + (np.abs(r) - direction * ey) * np.sin(angle + direction * span)
    return x, y, angle + direction * span + np.pi / 2          # get_orientation (:125)

"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference plant step
`DynamicBicycleModel.forward_dynamics` (car_racing/utils/base.py:897-942): sub-steps of
`vehicle_dynamics` (car_racing/system/vehicle_dynamics.py:4-49) with `get_curvature`
(car_racing/utils/racing_env.py:225-246) evaluated every sub-step, then the clipped noise (base.py:927-939).
PINNED: tests/golden/plant_golden.npz (tests/golden/make_plant_golden.py runs the reference itself)."""
import numpy as np


def substeps(timestep, delta_t=0.001):
    i = 0
    while (i + 1) * delta_t <= timestep:      # the reference's loop condition (base.py:909)
        i += 1
    return i


def curvature(pat, lap_length, s):
    s = np.array(s, dtype=float, copy=True)
    for b in range(s.shape[0]):
        while s[b] > lap_length:
            s[b] -= lap_length
        while s[b] < 0:
            s[b] += lap_length
    hit = (s[:, None] >= pat[None, :, 3]) & (s[:, None] <= pat[None, :, 3] + pat[None, :, 4])
    return pat[np.argmax(hit, axis=1), 5]


def plant_step(xcurv, xglob, u, draws, dyn, pat, lap_length, timestep=0.1, delta_t=0.001):
    """xcurv, xglob (B,6); u (B,2); draws (B,3) standard-normal draws (0 = zero_noise_flag).  Returns next (xcurv, xglob)."""
    m, lf, lr, Iz, Df, Cf, Bf, Dr, Cr, Br = dyn
    xc = np.array(xcurv, dtype=float, copy=True); xg = np.array(xglob, dtype=float, copy=True)
    delta, a = u[:, 0], u[:, 1]
    for _ in range(substeps(timestep, delta_t)):
        cur = curvature(pat, lap_length, xc[:, 4])
        psi, X, Y = xg[:, 3], xg[:, 4], xg[:, 5]
        vx, vy, wz, epsi, s, ey = (xc[:, k] for k in range(6))
        alpha_f = delta - np.arctan2(vy + lf * wz, vx)
        alpha_r = -np.arctan2(vy - lf * wz, vx)                     # lf, not lr: reference quirk (vehicle_dynamics.py:26)
        Fyf = 2 * Df * np.sin(Cf * np.arctan(Bf * alpha_f))
        Fyr = 2 * Dr * np.sin(Cr * np.arctan(Br * alpha_r))
        n0 = vx + delta_t * (a - 1 / m * Fyf * np.sin(delta) + wz * vy)
        n1 = vy + delta_t * (1 / m * (Fyf * np.cos(delta) + Fyr) - wz * vx)
        n2 = wz + delta_t * (1 / Iz * (lf * Fyf * np.cos(delta) - lr * Fyr))
        g3 = psi + delta_t * wz
        g4 = X + delta_t * (vx * np.cos(psi) - vy * np.sin(psi))
        g5 = Y + delta_t * (vx * np.sin(psi) + vy * np.cos(psi))
        c3 = epsi + delta_t * (wz - (vx * np.cos(epsi) - vy * np.sin(epsi)) / (1 - cur * ey) * cur)
        c4 = s + delta_t * ((vx * np.cos(epsi) - vy * np.sin(epsi)) / (1 - cur * ey))
        c5 = ey + delta_t * (vx * np.sin(epsi) + vy * np.cos(epsi))
        xg = np.stack([n0, n1, n2, g3, g4, g5], axis=1)
        xc = np.stack([n0, n1, n2, c3, c4, c5], axis=1)
    nz = np.stack([np.clip(draws[:, 0] * 0.01, -0.05, 0.05), np.clip(draws[:, 1] * 0.01, -0.1, 0.1),
                   np.clip(draws[:, 2] * 0.005, -0.05, 0.05)], axis=1)
    xc[:, 0:3] += 0.5 * nz                                            # the noise enters xcurv only (base.py:936-938)
    return xc, xg

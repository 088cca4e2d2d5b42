"""Generates tests/golden/plant_golden.npz with the UNMODIFIED reference plant: DynamicBicycleModel.forward_dynamics
(car_racing/utils/base.py:897-942) = 100 explicit-Euler sub-steps of system/vehicle_dynamics.py:4-49 with the track
curvature looked up every sub-step (utils/racing_env.py:225-246), then the clipped process noise.  Build container only.
numpy.random.randn is replaced by a recorded sequence so that the noise draws travel with the fixture.

    python tests/golden/make_plant_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = "/root/reference"
from make_ilqr_golden import import_reference_control   # noqa: E402


def main():
    _, base = import_reference_control()
    from utils import racing_env
    rng = np.random.default_rng(7)
    out = {}
    for ti, name in enumerate(["ellipse", "l_shape"]):
        spec = np.genfromtxt(os.path.join(REF, "data/track_layout/%s.csv" % name), delimiter=",")
        track = racing_env.ClosedTrack(spec, 1.0)
        out["pat_%s" % name] = np.asarray(track.point_and_tangent, float)
        E, T = 6, 12
        xc = np.zeros((E, T + 1, 6)); xg = np.zeros((E, T + 1, 6)); us = np.zeros((E, T, 2)); draws = np.zeros((E, T, 3))
        for e in range(E):
            m = base.DynamicBicycleModel(name="ego", param=base.CarParam(), system_param=base.SystemParam())
            m.track = track
            m.lap_length = track.lap_length
            m.set_timestep(0.1)
            s0 = rng.uniform(0.5, track.lap_length - 0.5)
            m.xcurv = np.array([rng.uniform(0.5, 1.6), rng.uniform(-.05, .05), rng.uniform(-.3, .3), rng.uniform(-.15, .15), s0,
                                rng.uniform(-.5, .5)])
            m.xglob = np.array([m.xcurv[0], m.xcurv[1], m.xcurv[2], rng.uniform(-3, 3), rng.uniform(-5, 5), rng.uniform(-5, 5)])
            m.zero_noise_flag = (e == 0)
            xc[e, 0], xg[e, 0] = m.xcurv, m.xglob
            for k in range(T):
                m.u = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-1.0, 1.0)])
                d = rng.normal(size=3) * (8.0 if k % 5 == 4 else 1.0)      # some draws hit the clipping
                seq = list(d)
                orig = np.random.randn
                np.random.randn = lambda *a: seq.pop(0)
                try:
                    m.forward_dynamics(False)
                finally:
                    np.random.randn = orig
                us[e, k], draws[e, k] = m.u, (0.0 if e == 0 else d)
                xc[e, k + 1], xg[e, k + 1] = m.xcurv, m.xglob
        out.update({"xcurv_%s" % name: xc, "xglob_%s" % name: xg, "u_%s" % name: us, "draws_%s" % name: draws,
                    "lap_length_%s" % name: track.lap_length})
    p = base.CarParam().dynamics_param
    out["dyn"] = np.array(p.get_params(), float)
    path = os.path.join(HERE, "plant_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

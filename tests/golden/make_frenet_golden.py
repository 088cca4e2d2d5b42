"""Generates tests/golden/frenet_golden.npz with the reference's curvilinear -> global-frame conversion:
racing_env.get_global_position / get_orientation (car_racing/utils/racing_env.py:6-127), which the controllers use to log
predictions in the global frame (utils/base.py:500-509, 573-580) and the planner to plot candidates
(planning/planner_helper.py:208-220).  Build container only.  The reference calls numpy.asscalar, which numpy >= 1.23 no
longer has; the generator restores it (`a.item()`, its documented replacement) -- the reference's code is not touched.

    python tests/golden/make_frenet_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = "/root/reference"
from make_ilqr_golden import import_reference_control   # noqa: E402


def main():
    import_reference_control()
    from utils import racing_env
    if not hasattr(np, "asscalar"):
        np.asscalar = lambda a: a.item()
    rng = np.random.default_rng(5)
    out = {}
    for name in ["ellipse", "l_shape", "goggle", "m_shape"]:
        spec = np.genfromtxt(os.path.join(REF, "data/track_layout/%s.csv" % name), delimiter=",")
        track = racing_env.ClosedTrack(spec, 1.0)
        pat = np.asarray(track.point_and_tangent, float)
        lap = track.lap_length
        s = np.concatenate([rng.uniform(-1.5, 2.3 * lap, size=300),
                            pat[:, 3], pat[:, 3] + pat[:, 4], pat[:, 3] + 0.5 * pat[:, 4],          # segment starts / ends / middles
                            pat[:, 3] + pat[:, 4] + 0.0005, [0.0, lap, lap + 1e-9]])
        ey = rng.uniform(-1.0, 1.0, size=s.shape[0])
        xy = np.zeros((s.shape[0], 2))
        psi = np.zeros(s.shape[0])
        for k in range(s.shape[0]):
            xy[k] = track.get_global_position(s[k], ey[k])
            psi[k] = track.get_orientation(s[k], ey[k])
        out.update({name + "/pat": pat, name + "/lap_length": np.array(lap), name + "/s": s, name + "/ey": ey, name + "/xy": xy,
                    name + "/psi": psi})
        print(name, "points", s.shape[0], "segments", pat.shape[0], "straight", int((pat[:, 5] == 0).sum()), flush=True)
    path = os.path.join(HERE, "frenet_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

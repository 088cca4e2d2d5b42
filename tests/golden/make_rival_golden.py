"""Generates tests/golden/rival_golden.npz with the UNMODIFIED reference NoDynamicsModel
(car_racing/utils/base.py:845-890, sympy path) -- build container only.  casadi/matplotlib/cvxopt/pathos are
stubbed (never called on this path); the track object is a stub whose global-frame functions return zeros
(the reference's own ones call the removed numpy.asscalar and the callers discard that block anyway).

    python tests/golden/make_rival_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_ilqr_golden import import_reference_control   # noqa: E402


def main():
    import sympy as sp
    _, base = import_reference_control()
    t = sp.symbols("t")
    cases = [(1.2 * t + 10.5, -0.5 + 0.0 * t), (1.22 * t + 12.0, 0.3 * sp.sin(0.5 * t)), (4.0 + 0.2 * t, sp.Float(0.1)),
             (0.05 * t ** 2 + 0.8 * t + 3.0, 0.4 * sp.cos(0.3 * t) - 0.2), (sp.Float(7.5), sp.Float(-0.35))]
    track = types.SimpleNamespace(get_global_position=lambda s, ey: (0.0, 0.0), get_orientation=lambda s, ey: 0.0)
    out = {}
    for c, (s_func, ey_func) in enumerate(cases):
        m = base.NoDynamicsModel(name="car%d" % c, param=base.CarParam())
        m.track = track
        m.set_state_curvilinear_func(t, s_func, ey_func)
        for time in (0.0, 0.7, 12.3):
            m.time = time
            xc, _ = m.get_trajectory_nsteps(time, 0.1, 21)
            out["c%d_t%g" % (c, time)] = xc
    path = os.path.join(HERE, "rival_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()

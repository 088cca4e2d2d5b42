"""Rival prediction drop-in against golden vectors from the reference's sympy path
(tests/golden/make_rival_golden.py)."""
import os
import time
import types

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rival_golden.npz")


def _cases():
    sp = pytest.importorskip("sympy")
    t = sp.symbols("t")
    return t, [(1.2 * t + 10.5, -0.5 + 0.0 * t), (1.22 * t + 12.0, 0.3 * sp.sin(0.5 * t)), (4.0 + 0.2 * t, sp.Float(0.1)),
               (0.05 * t ** 2 + 0.8 * t + 3.0, 0.4 * sp.cos(0.3 * t) - 0.2), (sp.Float(7.5), sp.Float(-0.35))]


def test_fast_prediction_matches_reference_golden():
    from car_racing_b200 import rivals
    g = np.load(GOLD)
    t, cases = _cases()
    for c, (s_func, ey_func) in enumerate(cases):
        m = types.SimpleNamespace(t_symbol=t, s_func=s_func, ey_func=ey_func, time=0.0)
        for tm in (0.0, 0.7, 12.3):
            m.time = tm
            xc, xg = rivals.get_trajectory_nsteps(m, 99.0, 0.1, 21)     # t0 is ignored, as in the reference (base.py:883)
            assert xc.shape == (6, 21) and xg.shape == (6, 21)
            assert np.abs(xc - g["c%d_t%g" % (c, tm)]).max() < 1e-12
    # compiled once per rival
    m = types.SimpleNamespace(t_symbol=t, s_func=cases[1][0], ey_func=cases[1][1], time=1.0)
    rivals.get_trajectory_nsteps(m, 0, 0.1, 21)
    t0 = time.perf_counter()
    for _ in range(50):
        rivals.get_trajectory_nsteps(m, 0, 0.1, 21)
    assert (time.perf_counter() - t0) / 50 < 2e-3


def test_with_glob_uses_the_track_object():
    from car_racing_b200 import rivals
    t, cases = _cases()
    trk = types.SimpleNamespace(get_global_position=lambda s, ey: (2.0 * s, ey), get_orientation=lambda s, ey: 0.25)
    m = types.SimpleNamespace(t_symbol=t, s_func=cases[0][0], ey_func=cases[0][1], time=0.5, track=trk)
    xc, xg = rivals.get_trajectory_nsteps(m, 0, 0.1, 5, with_glob=True)
    assert np.allclose(xg[4], 2.0 * xc[4]) and np.allclose(xg[5], xc[5]) and np.allclose(xg[3], 0.25) and np.allclose(xg[0], xc[0])


def test_rival_still_pickles_after_a_prediction():
    """The reference pickles the simulator with its vehicles after a run (car_racing/tests/mpccbf_test.py:45-46): the drop-in
    must not leave compiled functions on the rival."""
    import pickle
    from car_racing_b200 import rivals
    t, cases = _cases()
    m = types.SimpleNamespace(t_symbol=t, s_func=cases[1][0], ey_func=cases[1][1], time=0.3)
    before = set(vars(m))
    xc, _ = rivals.get_trajectory_nsteps(m, 0, 0.1, 11)
    assert set(vars(m)) == before
    m2 = pickle.loads(pickle.dumps(m))
    assert np.array_equal(rivals.get_trajectory_nsteps(m2, 0, 0.1, 11)[0], xc)

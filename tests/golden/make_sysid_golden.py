"""Generates tests/golden/sysid_golden.npz by running the UNMODIFIED reference
car_racing/control/lmpc_helper.py:regression_and_linearization (the body of LMPCRacingGame.estimate_ABC,
utils/base.py:585-622) imported from /root/reference -- build container only; the .npz travels.

cvxopt 1.3.0 (requirements.txt) is absent here.  The reference uses it for `qp(Q, b)` WITHOUT constraints
(lmpc_helper.py:360), i.e. the linear system Q x = -b; the stub below answers that call with
numpy.linalg.solve(Q, -b) and `matrix` with numpy.asarray.  Everything else (nearest-neighbour selection,
Epanechnikov weights, normal equations, curvature lookup, analytic linearisation) is the reference's own code.

    python tests/golden/make_sysid_golden.py
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    def qp(Q, b):
        return {"x": np.linalg.solve(np.asarray(Q, float), -np.asarray(b, float))}

    def matrix(x):
        return np.asarray(x, float)

    solvers = _stub("cvxopt.solvers", qp=qp)
    _stub("cvxopt", matrix=matrix, spmatrix=None, solvers=solvers)
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.animation"]:
        _stub(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, os.path.join(REF, "car_racing"))
    from control import lmpc_helper
    from utils import racing_env
    return lmpc_helper, racing_env, qp, matrix


def stored_laps(rng, track_len, num_laps=3, num_points=400):
    """Stored closed-loop laps as LMPCRacingGame keeps them (base.py:430-435): 10000-filled arrays, lap j valid on
    [0, time_ss[j]].  States follow the identified LTI model under the reference's PID law plus excitation."""
    from car_racing_b200.scenarios import LTI_A, LTI_B
    ss = 10000.0 * np.ones((num_points, 6, num_laps))
    us = 10000.0 * np.ones((num_points, 2, num_laps))
    time_ss = 10000 * np.ones(num_laps, dtype=int)
    for lap in range(num_laps):
        vt = 1.0 + 0.15 * lap
        x = np.array([vt, 0.0, 0.0, rng.uniform(-.02, .02), 0.0, rng.uniform(-.1, .1)])
        t = 0
        while x[4] < track_len and t < num_points - 2:
            ss[t, :, lap] = x
            u = np.array([-0.6 * x[5] - 0.9 * x[3] + 0.05 * np.sin(0.35 * t + lap), 1.5 * (vt - x[0]) + 0.2 * np.sin(0.21 * t)])
            u = np.clip(u + rng.normal(scale=[0.01, 0.03]), [-0.5, -1.0], [0.5, 1.0])
            us[t, :, lap] = u
            x = LTI_A @ x + LTI_B @ u + rng.normal(scale=[2e-3, 1e-3, 3e-3, 1e-3, 0, 1e-3])
            t += 1
        ss[t, :, lap] = x
        time_ss[lap] = t
    return ss, us, time_ss


def main():
    lmpc_helper, racing_env, qp, matrix = import_reference()
    rng = np.random.default_rng(0)
    spec = np.genfromtxt(os.path.join(REF, "data/track_layout/ellipse.csv"), delimiter=",")
    track = racing_env.ClosedTrack(spec, 1.0)
    pat = np.asarray(track.point_and_tangent, float)
    ss, us, time_ss = stored_laps(rng, track.lap_length)
    it = 2                                           # two stored laps -> used_iter = range(0, 2) (base.py:600-601)
    used_iter = range(it - 2, it)
    N, dt, max_num_point = 12, 0.1, 40
    cases = 24
    lin_points = np.zeros((cases, N + 1, 6))
    lin_input = np.zeros((cases, N, 2))
    A = np.zeros((cases, N, 6, 6)); B = np.zeros((cases, N, 6, 2)); C = np.zeros((cases, N, 6))
    idx = -np.ones((cases, N, 2, max_num_point), dtype=np.int64)
    for c in range(cases):
        lap = c % 2
        t0 = int(rng.integers(1, time_ss[lap] - N - 3))
        lin_points[c] = ss[t0:t0 + N + 1, :, lap] + rng.normal(scale=[0.02, 0.005, 0.01, 0.005, 0.01, 0.01], size=(N + 1, 6))
        lin_points[c, :, 4] = np.maximum(lin_points[c, :, 4], 1e-3)
        lin_input[c] = us[t0:t0 + N, :, lap] + rng.normal(scale=[0.01, 0.02], size=(N, 2))
        for i in range(N):
            Ai, Bi, Ci, sel = lmpc_helper.regression_and_linearization(lin_points[c], lin_input[c], used_iter, ss, us, time_ss,
                                                                       max_num_point, qp, matrix, pat, dt, i)
            A[c, i], B[c, i], C[c, i] = Ai, Bi, Ci[:, 0]
            for j, s_ in enumerate(sel):
                idx[c, i, j, :len(s_)] = s_
    out = os.path.join(HERE, "sysid_golden.npz")
    np.savez_compressed(out, ss=ss[:, :, :2], us=us[:, :, :2], time_ss=time_ss[:2], point_and_tangent=pat, dt=dt,
                        max_num_point=max_num_point, lin_points=lin_points, lin_input=lin_input, A=A, B=B, C=C, idx=idx,
                        lap_length=track.lap_length)
    print("wrote", out, os.path.getsize(out), "bytes; time_ss", time_ss, "A range", np.abs(A).max(), "C range", np.abs(C).max())


if __name__ == "__main__":
    main()

"""Generates tests/golden/ilqr_golden.npz by running the UNMODIFIED reference
car_racing/control/control.py:ilqr (imported from /root/reference; only works in the build
container -- the reference does not travel to the GPU box, the .npz does).

Third-party modules the reference imports at module scope but that iLQR never calls
(casadi, matplotlib, cvxopt, pathos) are replaced by empty stubs; numpy/scipy are real.

    python tests/golden/make_ilqr_golden.py
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REF = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__all__ = []
    sys.modules[name] = m
    return m


def import_reference_control():
    for name in ["casadi", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.animation",
                 "cvxopt", "cvxopt.solvers", "pathos", "pathos.multiprocessing"]:
        _stub(name)
    sys.modules["cvxopt.solvers"].qp = None
    for k in ["spmatrix", "matrix", "solvers"]:
        setattr(sys.modules["cvxopt"], k, None)
    sys.modules["pathos.multiprocessing"].ProcessingPool = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, os.path.join(REF, "car_racing"))
    cwd = os.getcwd()
    os.chdir(REF)  # Param classes read data/sys/LTI/*.csv relative to CWD at import time (base.py:124-125)
    try:
        from control import control
        from utils import base
    finally:
        os.chdir(cwd)
    return control, base


class Rival:
    """Duck-typed rival: .param.length/.width and get_trajectory_nsteps (base.py:879-886)."""

    def __init__(self, traj, length=0.4, width=0.2):
        self.param = types.SimpleNamespace(length=length, width=width)
        self._traj = traj

    def get_trajectory_nsteps(self, t0, dt, n):
        return self._traj[:, :n].copy(), None


def scenarios(rng, n, lap, N):
    """x0 as config 2, one rival ahead s+U(0.6,1.6), ey_r~U(-0.7,0.7);
    a third of the rivals move (s_j(t) = s_j + v t) and a few instances sit on lap 1 (lap offset)."""
    out = []
    for k in range(n):
        vx = rng.uniform(0.4, 1.5)
        x0 = np.array([vx, rng.uniform(-.05, .05), rng.uniform(-.2, .2), rng.uniform(-.1, .1),
                       rng.uniform(0, lap - 5), rng.uniform(-.6, .6)])
        if k % 5 == 4:
            x0[4] += lap
        vt = rng.uniform(0.5, 1.2)
        s_r = x0[4] + rng.uniform(0.6, 1.6) - (lap if k % 10 == 9 else 0.0)
        ey_r = rng.uniform(-0.7, 0.7)
        v_r = rng.uniform(0.0, 0.5) if k % 3 == 0 else 0.0
        traj = np.zeros((6, N + 1))
        traj[4] = s_r + v_r * 0.1 * np.arange(N + 1)
        traj[5] = ey_r
        out.append((x0, np.array([vt, 0, 0, 0, 0, rng.uniform(-0.2, 0.2)]), traj))
    return out


def main():
    control, base = import_reference_control()
    lap = 19.2296
    recs = []
    for N, n_case, seed in [(50, 24, 0), (20, 8, 1)]:
        os.chdir(REF)
        prm = base.iLQRRacingParam(num_horizon=N)
        os.chdir(os.path.dirname(os.path.abspath(__file__)))
        rng = np.random.default_rng(seed)
        for x0, xt, traj in scenarios(rng, n_case, lap, N):
            vehicles = {"ego": Rival(None), "car1": Rival(traj)}
            with contextlib.redirect_stdout(io.StringIO()):
                u0 = control.ilqr(x0.copy(), xt.copy(), prm, vehicles, "ego", lap, 0.0, 0.1, None, None)
            recs.append(dict(N=N, x0=x0, xt=xt, obs=traj[4:6].copy(), u0=np.array(u0, float), lap=lap))
    Nmax = max(r["N"] for r in recs)
    obs = np.zeros((len(recs), 2, Nmax + 1))
    for i, r in enumerate(recs):
        obs[i, :, : r["N"] + 1] = r["obs"]
    np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ilqr_golden.npz"),
             N=np.array([r["N"] for r in recs]), x0=np.array([r["x0"] for r in recs]),
             xt=np.array([r["xt"] for r in recs]), obs=obs, u0=np.array([r["u0"] for r in recs]),
             lap=np.array([r["lap"] for r in recs]),
             A=np.asarray(prm.matrix_A), B=np.asarray(prm.matrix_B), Q=np.asarray(prm.matrix_Q), R=np.asarray(prm.matrix_R),
             max_iter=np.array(prm.max_iter))
    print("wrote", len(recs), "cases")


if __name__ == "__main__":
    main()

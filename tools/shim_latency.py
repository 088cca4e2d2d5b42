"""Per-step latency of the drop-in shims inside the control loop, on the GPU, from committed inputs only.

The reference's simulator cannot travel to the GPU box (no /root/reference there), so this driver rebuilds the loop of
offboard.CarRacingSim.sim (racing/offboard.py:124-127: for every vehicle `forward_one_step`, i.e. ctrl.calc_input -> plant
step -> memory update) out of the repo's own pinned pieces: the shims of car_racing_b200.control / .planning called with
duck-typed *Param objects and rivals, the plant kernel (b200mpc_plant_step, pinned to forward_dynamics by
tests/golden/plant_golden.npz) and NoDynamicsModel-style rivals (s = s0 + v t).  Scenarios are the reference's own test
set-ups (car_racing/tests/mpccbf_test.py:20-41, control_test.py, ilqr_test.py, lmpc / overtake_planner_test.py), track
geometry from the committed golden files.  What it reports is what SURVEY.md 8(b) asks of a drop-in in the 10 Hz loop
(tests/auto_mpccbf_test.py:20,35: timestep 0.1 s): wall time of each shim call (host packing + C-ABI + GPU + read-back),
p50 / p90 / max over the episode, against the 100 ms budget.

    python tools/shim_latency.py --out profiles/r05_shim_latency.json
"""
import argparse
import json
import os
import sys
import time
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import car_racing_b200 as crb                      # noqa: E402
from car_racing_b200 import batch, control, planning, scenarios   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class Rival:
    """NoDynamicsModel (utils/base.py:860-886): s(t) = s0 + v t, ey constant; get_trajectory_nsteps returns (6, n)."""

    def __init__(self, s0, v, ey, length=0.4, width=0.2):
        self.param = types.SimpleNamespace(length=length, width=width)
        self.s0, self.v, self.ey, self.time = s0, v, ey, 0.0

    def get_trajectory_nsteps(self, t0, dt, n):
        tr = np.zeros((6, n))
        tr[4] = self.s0 + self.v * (self.time + dt * np.arange(n))
        tr[5] = self.ey
        tr[0] = self.v
        return tr, None


def stats(ms):
    ms = np.asarray(ms)
    return dict(calls=int(ms.size), p50_ms=float(np.median(ms)), p90_ms=float(np.percentile(ms, 90)), p99_ms=float(np.percentile(ms, 99)),
                max_ms=float(ms.max()), mean_ms=float(ms.mean()), budget_ms=100.0, over_budget=int((ms > 100.0).sum()))


def plant(pat):
    def step(x, xg, u):
        xc, xgn = batch.plant_step_batch(x[None], xg[None], u[None], None, pat)
        return xc[0], xgn[0]
    return step


def episode_mpccbf(steps, pat, lap):
    """mpccbf_test.py: ego MPC-CBF (N=10, alpha 0.8, vt 0.8), car1 at 4 m / car2 at 10 m, both 0.2 m/s, l_shape."""
    prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 4.0, 0, 40.0]),
                                matrix_R=np.diag([0.1, 0.1]), num_horizon=10, alpha=0.8)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    track = types.SimpleNamespace(width=1.0, lap_length=lap)
    veh = {"ego": types.SimpleNamespace(param=types.SimpleNamespace(length=0.4, width=0.2)),
           "car1": Rival(4.0, 0.2, 0.1), "car2": Rival(10.0, 0.2, -0.1)}
    x, xg = np.zeros(6), np.zeros(6)
    xt = np.array([0.8, 0, 0, 0, 0, 0.0])
    step = plant(pat)
    ms, status, retries, hmin, laps, t = [], [], 0, np.inf, 0, 0.0
    for k in range(steps):
        t0 = time.perf_counter()
        u = control.mpccbf(x, xt, prm, veh, "ego", lap, t, 0.1, False, track, sysp)
        ms.append(1e3 * (time.perf_counter() - t0))
        status.append(control.last_solve["status"])
        retries += len(control.last_solve["retries"])
        x, xg = step(x, xg, u)
        if x[4] > lap:
            x[4] -= lap
            laps += 1
        t += 0.1
        for n in ("car1", "car2"):
            veh[n].time = t
            ds = (x[4] + laps * lap) - (veh[n].s0 + veh[n].v * t)
            hmin = min(hmin, (ds / 0.4) ** 6 + ((x[5] - veh[n].ey) / 0.2) ** 6 - 1.0)
    passed = [bool(x[4] + laps * lap > veh[n].s0 + veh[n].v * t) for n in ("car1", "car2")]
    return dict(latency=stats(ms[5:]), first_call_ms=ms[0], statuses=np.bincount(status, minlength=5).tolist(), retries=retries,
                barrier_min=float(hmin), ego_passed_rivals=passed, ego_s=float(x[4] + laps * lap), steps=steps)


def episode_mpc_lti(steps, pat, lap):
    """control_test.py / auto_control_test.py: MPC-LTI tracking, N=10, vt 0.8, track width 0.8."""
    prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 4.0, 0, 40.0]),
                                matrix_R=np.diag([0.1, 0.1]), num_horizon=10)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    track = types.SimpleNamespace(width=0.8, lap_length=lap)
    x, xg = np.zeros(6), np.zeros(6)
    xt = np.array([0.8, 0, 0, 0, 0, 0.0])
    step = plant(pat)
    ms = []
    for k in range(steps):
        t0 = time.perf_counter()
        u = control.mpc_lti(x, xt, prm, sysp, track)
        ms.append(1e3 * (time.perf_counter() - t0))
        x, xg = step(x, xg, u)
        if x[4] > lap:
            x[4] -= lap
    return dict(latency=stats(ms[5:]), final_vx=float(x[0]), max_abs_ey=float(abs(x[5])), steps=steps)


def episode_ilqr(steps, pat, lap):
    """ilqr_test.py: iLQR N=50 (base.py:167-186), one slower rival ahead."""
    prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 4.0, 0, 40.0]),
                                matrix_R=np.diag([0.1, 0.1]), num_horizon=50, max_iter=150)
    veh = {"ego": types.SimpleNamespace(param=types.SimpleNamespace(length=0.4, width=0.2)), "car1": Rival(4.0, 0.2, 0.1)}
    x, xg = np.array([0.5, 0, 0, 0, 0, 0.0]), np.zeros(6)
    xt = np.array([0.8, 0, 0, 0, 0, 0.0])
    step = plant(pat)
    ms, t = [], 0.0
    for k in range(steps):
        t0 = time.perf_counter()
        u = control.ilqr(x, xt, prm, veh, "ego", lap, t, 0.1, None, None)
        ms.append(1e3 * (time.perf_counter() - t0))
        x, xg = step(x, xg, u)
        t += 0.1
        veh["car1"].time = t
    return dict(latency=stats(ms[5:]), final_s=float(x[4]), steps=steps)


def episode_lmpc(steps):
    """The LMPC step of the racing game (utils/base.py:459-476): estimate_ABC (model identification on two stored laps) then
    control.lmpc, from the stored laps of tests/golden/sysid_golden.npz (generated by the reference's own PID / MPC laps)."""
    g = np.load(os.path.join(GOLD, "sysid_golden.npz"))
    ss, us, time_ss, pat = g["ss"], g["us"], g["time_ss"], g["point_and_tangent"]
    lap = float(pat[-1, 3] + pat[-1, 4])
    N = 12
    lmpc_param = types.SimpleNamespace(matrix_Q=np.zeros((6, 6)), matrix_R=np.diag([1.0, 0.25]), matrix_dR=5 * np.diag([0.8, 0.0]),
                                       num_horizon=N, num_ss_points=44, num_ss_iter=2, shift=0)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    T = int(time_ss[1])
    Qfun = np.zeros((ss.shape[0], ss.shape[2]))
    for it in range(ss.shape[2]):
        Qfun[:int(time_ss[it]), it] = np.arange(int(time_ss[it]), 0, -1)
    self = types.SimpleNamespace(lmpc_param=lmpc_param, iter=2, lin_points=g["lin_points"][0].copy(), lin_input=g["lin_input"][0].copy(),
                                 ss_xcurv=ss, u_ss=us, time_ss=time_ss, point_and_tangent=pat, timestep=0.1)
    x = self.lin_points[0].copy()
    u_old = self.lin_input[0].copy()
    ms_id, ms_qp, status = [], [], []
    for k in range(steps):
        t0 = time.perf_counter()
        Atv, Btv, Ctv, _ = control.estimate_ABC(self)
        t1 = time.perf_counter()
        u_pred, x_pred, _, _, lin_points, lin_input = control.lmpc(x, lmpc_param, Atv, Btv, Ctv, ss, Qfun, 2, lap, 1.0, u_old, sysp)
        t2 = time.perf_counter()
        ms_id.append(1e3 * (t1 - t0))
        ms_qp.append(1e3 * (t2 - t1))
        self.lin_points, self.lin_input = lin_points, lin_input
        x, u_old = x_pred[1].copy(), u_pred[0].copy()         # follow the LTV prediction (the plant of the golden laps is not re-run here)
    return dict(estimate_ABC=stats(ms_id[3:]), lmpc=stats(ms_qp[3:]), both=stats((np.array(ms_id) + np.array(ms_qp))[3:]), steps=steps,
                stored_lap_rows=T)


def episode_overtake(steps):
    """The overtaking step (utils/base.py:540-582): planner candidates -> selection -> tracking MPC-CBF, as one fused call
    (planning.plan_and_track) and as the reference's two calls (solve_optimization_problem + mpc_multi_agents)."""
    from planner_cases import make_planner
    mp = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 5.0, 0, 50.0]),
                               matrix_R=np.diag([0.1, 0.1]), num_horizon_ctrl=10)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    ms_two, ms_one = [], []
    for k in range(steps):
        pl = make_planner(100 + k, num_veh=2 + k % 2)
        for name in pl.sorted_vehicles:
            tr = pl.obs_infos[name]
            pl.vehicles[name] = Rival(tr[4, 0], tr[0, 0], tr[5, 0])
        xc = np.asarray(pl.vehicles["ego"].xcurv, float).copy()
        t0 = time.perf_counter()
        t2, f2, _, _ = planning.solve_optimization_problem(pl)
        control.mpc_multi_agents(xc, mp, pl.track, None, None, None, sysp, target_traj_xcurv=t2, vehicles=pl.vehicles, agent_name="ego",
                                 direction_flag=f2, sorted_vehicles=pl.sorted_vehicles, time=None)
        t1 = time.perf_counter()
        planning.plan_and_track(pl, xc, mp, pl.track, sysp, time=None)
        t3 = time.perf_counter()
        ms_two.append(1e3 * (t1 - t0))
        ms_one.append(1e3 * (t3 - t1))
    return dict(two_calls=stats(ms_two[3:]), fused=stats(ms_one[3:]), steps=steps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--steps", type=int, default=500, help="control steps of the MPC-CBF episode (mpccbf_test.py simulates 50 s = 500)")
    ap.add_argument("--short", action="store_true", help="a few steps of every episode (checking the script itself)")
    ap.add_argument("--lib", default=None, help="library to load instead of libb200mpc.so (the host emulation, to check the script without a GPU)")
    a = ap.parse_args()
    if a.lib:
        from car_racing_b200 import _capi
        _capi.LIB_PATH, _capi._lib = os.path.abspath(a.lib), None
    warnings.simplefilter("ignore")
    pg = np.load(os.path.join(GOLD, "plant_golden.npz"))
    pat, lap = pg["pat_l_shape"], float(pg["lap_length_l_shape"])
    import io
    import contextlib
    doc = {"what": "wall time per shim call in a closed loop on this box (host packing + ctypes + H2D + kernel + D2H), 10 Hz budget = 100 ms",
           "gpu": os.popen("nvidia-smi --query-gpu=name --format=csv,noheader").read().strip()}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):          # the shims print the reference's own messages ("solver failed.")
        n = (lambda full: 8 if a.short else full)
        doc["mpccbf_test"] = episode_mpccbf(n(a.steps), pat, lap)
        doc["mpc_lti_tracking"] = episode_mpc_lti(n(200), pat, lap)
        doc["ilqr_test"] = episode_ilqr(n(200), pat, lap)
        doc["lmpc_step"] = episode_lmpc(n(60))
        doc["overtaking_step"] = episode_overtake(n(40))
    s = json.dumps(doc, indent=1)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()

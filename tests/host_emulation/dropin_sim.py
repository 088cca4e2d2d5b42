"""The drop-in inside the reference's OWN simulator: the scenarios of car_racing/tests/{ilqr,mpccbf,control}_test.py are set up
exactly as those scripts do (same track, same Param classes, same rivals, offboard.CarRacingSim.sim loop), the only edit is
`car_racing_b200.install()` -- the module-level swap of control.control.{ilqr, mpccbf, mpc_lti, ...}.

  * iLQR (ilqr_test.py): the reference's numpy `control.ilqr` runs natively here, so the SAME closed loop is simulated twice --
    unmodified reference, then with the shims -- and the two ego trajectories are compared step by step (the plant's noise
    draws are the same numpy stream in both runs).
  * MPC-CBF (mpccbf_test.py) and MPC-LTI (control_test.py --ctrl-policy mpc-lti): CasADi/IPOPT cannot be installed offline, so only the
    shim run exists; recorded are progress, the barrier value h against every rival at every step (h >= 0 = the safe set of
    control.py:527-557 was never left) and the solver status of every step.

Needs /root/reference, so it only runs in the build container; the summary it writes (profiles/*_dropin_sim.json) travels.
TEST INFRASTRUCTURE: without a GPU (`--lib emu`, the default when no CUDA device is visible) the shims call the host-compiled
copy of the library (tests/host_emulation/build_emu_library.py); on a GPU box `--lib product` uses car_racing_b200/libb200mpc.so.

    python tests/host_emulation/dropin_sim.py --steps 100 --out profiles/r04a_dropin_sim.json
"""
import argparse
import importlib.util
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__all__ = []
    sys.modules[name] = m
    return m


def import_reference():
    """Modules the reference imports at module scope but never calls on these paths are empty stubs (as in tests/golden/make_*.py)."""
    for name in ["casadi", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.animation",
                 "cvxopt", "cvxopt.solvers", "pathos", "pathos.multiprocessing"]:
        _stub(name)
    sys.modules["cvxopt.solvers"].qp = None
    for k in ["spmatrix", "matrix", "solvers"]:
        setattr(sys.modules["cvxopt"], k, None)
    sys.modules["pathos.multiprocessing"].ProcessingPool = lambda *a, **k: None   # LMPCRacingGame.__init__ makes a Pool(4) (base.py:446); the
                                                                                   # only user is the reference's estimate_ABC, which install_all replaces
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for k in ["patches", "animation"]:
        setattr(sys.modules["matplotlib"], k, sys.modules["matplotlib." + k])
    if not hasattr(np, "asscalar"):            # removed from numpy; racing_env.py:25,86 still calls it
        np.asscalar = lambda a: a.item()
    sys.path.insert(0, os.path.join(REF, "car_racing"))
    os.chdir(REF)                              # data/sys/LTI/*.csv and data/track_layout/*.csv are read relative to CWD
    from control import control
    from racing import offboard
    from utils import base, racing_env
    return control, offboard, base, racing_env


def build_sim(offboard, base, racing_env, policy, rivals, track_layout="l_shape"):
    """car_racing/tests/ilqr_test.py:10-37, mpccbf_test.py:10-46, control_test.py (same calls, same order)."""
    import sympy as sp
    track_spec = np.genfromtxt("data/track_layout/" + track_layout + ".csv", delimiter=",")
    track = racing_env.ClosedTrack(track_spec, track_width=1.0)
    ego = offboard.DynamicBicycleModel(name="ego", param=base.CarParam(edgecolor="black"), system_param=base.SystemParam())
    ego.set_state_curvilinear(np.zeros((6,)))
    ego.set_state_global(np.zeros((6,)))
    ego.start_logging()
    if policy == "ilqr":
        ego.set_ctrl_policy(offboard.iLQRRacing(base.iLQRRacingParam(vt=0.8), ego.system_param))
    elif policy == "mpccbf":
        ego.set_ctrl_policy(offboard.MPCCBFRacing(base.MPCCBFRacingParam(vt=0.8), ego.system_param))
    else:
        ego.set_ctrl_policy(offboard.MPCTracking(base.MPCTrackingParam(vt=0.8), ego.system_param))
    ego.ctrl_policy.set_timestep(0.1)
    ego.set_track(track)
    ego.ctrl_policy.set_track(track)
    t = sp.symbols("t")
    sim = offboard.CarRacingSim()
    sim.set_timestep(0.1)
    sim.set_track(track)
    sim.add_vehicle(ego)
    ego.ctrl_policy.set_racing_sim(sim)
    for k, (s0, v, ey) in enumerate(rivals):
        car = offboard.NoDynamicsModel(name="car%d" % (k + 1), param=base.CarParam(edgecolor="orange"))
        car.set_track(track)
        car.set_state_curvilinear_func(t, v * t + s0, ey + 0.0 * t)
        car.start_logging()
        sim.add_vehicle(car)
    return sim, ego, track


def run(sim, ego, steps, seed):
    np.random.seed(seed)                        # plant noise (base.py:897-942) draws from numpy's global stream
    xs, us = [], []
    t0 = time.time()
    for _ in range(steps):                      # offboard.CarRacingSim.sim (offboard.py:121-127), one step at a time to log
        for name in sim.vehicles:
            sim.vehicles[name].forward_one_step(sim.vehicles[name].realtime_flag)
        xs.append(np.array(ego.xcurv, float).copy())
        us.append(np.array(ego.u, float).copy())
    return np.array(xs), np.array(us), time.time() - t0


def racing_game(crb, control, offboard, base, racing_env, lmpc_steps, overtake_steps, seed, track_layout="l_shape"):
    """car_racing/tests/overtake_planner_test.py / lmpc_test.py with car_racing_b200.install_all(): lap 0 under the reference's own PID
    (data collection), lap 1 under MPC-LTI (shim), then LMPCRacingGame: `lmpc_steps` steps of LMPC (estimate_ABC + lmpc shims on the two
    recorded laps), then two rivals are put in front of the ego the way the test does (:149-157: speeds 1.2 / 1.22, ey -0.5 / -0.2, 1.5 m
    apart) and `overtake_steps` steps run through get_overtake_flag -> get_local_traj -> solve_optimization_problem -> mpc_multi_agents."""
    import sympy as sp
    from control.lmpc_helper import LMPCPrediction
    from planning.overtake_traj_planner import OvertakeTrajPlanner
    crb.install_all(control, base, OvertakeTrajPlanner, offboard)
    np.random.seed(seed)
    dt = 0.1
    track = racing_env.ClosedTrack(np.genfromtxt("data/track_layout/" + track_layout + ".csv", delimiter=","), track_width=1.0)
    opti_xcurv = np.genfromtxt("data/optimal_traj/xcurv_" + track_layout + ".csv", delimiter=",")
    opti_xglob = np.genfromtxt("data/optimal_traj/xglob_" + track_layout + ".csv", delimiter=",")
    ego = offboard.DynamicBicycleModel(name="ego", param=base.CarParam(edgecolor="black"), system_param=base.SystemParam())   # set_up_ego
    ego.set_timestep(dt)
    pid = offboard.PIDTracking(vt=0.7, eyt=0.0)
    pid.set_timestep(dt)
    ego.set_ctrl_policy(pid)
    pid.set_track(track)
    ego.set_state_curvilinear(np.zeros((6,)))
    ego.set_state_global(np.zeros((6,)))
    ego.start_logging()
    ego.set_track(track)
    mpc = offboard.MPCTracking(base.MPCTrackingParam(vt=0.7, eyt=0.0), ego.system_param)
    mpc.set_timestep(dt)
    mpc.set_track(track)
    lap_number = 4                                                                                                              # set_up_lmpc
    lmpc = offboard.LMPCRacingGame(base.LMPCRacingParam(timestep=dt, lap_number=lap_number, time_lmpc=10000 * dt),
                                   racing_game_param=base.RacingGameParam(timestep=dt, alpha=0.8, num_horizon_planner=10), system_param=ego.system_param)
    lmpc.set_track(track)
    lmpc.set_timestep(dt)
    lmpc.set_opti_traj(opti_xcurv, opti_xglob)
    lmpc.openloop_prediction = LMPCPrediction(lap_number=lap_number)
    sim = offboard.CarRacingSim()
    sim.set_timestep(dt)
    sim.set_track(track)
    sim.add_vehicle(ego)
    sim.set_opti_traj(opti_xglob)
    pid.set_racing_sim(sim)
    mpc.set_racing_sim(sim)
    lmpc.set_racing_sim(sim)
    lmpc.set_vehicles_track()
    t0 = time.time()
    sim.sim(sim_time=90, one_lap=True, one_lap_name="ego")                       # lap 0: PID (reference code, host)
    n_pid = len(ego.times[0])
    ego.set_ctrl_policy(mpc)
    sim.sim(sim_time=90, one_lap=True, one_lap_name="ego")                       # lap 1: MPC-LTI (shim)
    n_mpc = len(ego.times[1])
    t_laps = time.time() - t0
    lmpc.add_trajectory(ego, 0)
    lmpc.add_trajectory(ego, 1)
    ego.set_ctrl_policy(lmpc)
    t0 = time.time()
    xs = []
    for _ in range(lmpc_steps):                                                  # lap 2: LMPC (shims), no rivals
        ego.forward_one_step(ego.realtime_flag)
        xs.append(np.array(ego.xcurv, float).copy())
    xs = np.array(xs)
    t_lmpc = time.time() - t0
    t = sp.symbols("t")
    s_ego, t_now = float(ego.xcurv[4]), float(ego.time)
    rivals = []
    for k in range(2):                                                           # overtake_planner_test.py:149-157, placed relative to the ego
        v, s0, ey = 1.2 + 0.02 * k, s_ego + 1.0 + 1.5 * k, -0.5 + 0.3 * k
        car = offboard.NoDynamicsModel(name="car%d" % (k + 1), param=base.CarParam(edgecolor="orange"))
        car.set_track(track)
        car.set_state_curvilinear_func(t, v * t + s0, ey + 0.0 * t)
        car.start_logging()
        sim.add_vehicle(car)
        rivals.append(car)
    t0 = time.time()
    ys, clear, n_overtake = [], np.inf, 0
    for _ in range(overtake_steps):
        for name in sim.vehicles:
            sim.vehicles[name].forward_one_step(sim.vehicles[name].realtime_flag)
        ys.append(np.array(ego.xcurv, float).copy())
        n_overtake += ego.local_trajs[-1] is not None
        for car in rivals:
            ds = (ego.xcurv[4] - car.xcurv[4] + track.lap_length / 2) % track.lap_length - track.lap_length / 2
            clear = min(clear, max(abs(ds) - 0.4, abs(ego.xcurv[5] - car.xcurv[5]) - 0.2))
    ys = np.array(ys)
    t_ov = time.time() - t0
    lead = [float((ego.xcurv[4] - car.xcurv[4] + track.lap_length / 2) % track.lap_length - track.lap_length / 2) for car in rivals]
    return {"what": "car_racing/tests/overtake_planner_test.py flow with car_racing_b200.install_all(): PID lap (reference), MPC-LTI lap, LMPC steps, "
                    "then 2 rivals ahead: planner + mpc_multi_agents steps (all shims; no CasADi here)",
            "lap0_pid_steps": n_pid, "lap1_mpc_lti_steps": n_mpc, "wall_s_two_laps": t_laps,
            "lmpc_steps": lmpc_steps, "lmpc_vx_first_last": [float(xs[0, 0]), float(xs[-1, 0])], "lmpc_max_abs_ey": float(np.abs(xs[:, 5]).max()),
            "wall_s_lmpc": t_lmpc,
            "overtake_steps": overtake_steps, "steps_with_planner_active": int(n_overtake), "min_clearance_m": float(clear),
            "ego_minus_rival_s_at_end": lead, "overtake_max_abs_ey": float(np.abs(ys[:, 5]).max()), "track_half_width": float(track.width),
            "vx_at_end": float(ys[-1, 0]), "wall_s_overtake": t_ov}


def barrier(xs, rivals, lap, dt=0.1, L=0.4, W=0.2, margin=0.2):
    """h of control.py:527-557 at the simulated states (ego state after step k is at time (k+1) dt)."""
    hmin = np.inf
    for k, x in enumerate(xs):
        for s0, v, ey in rivals:
            ds = x[4] - (s0 + v * (k + 1) * dt)
            ds = (ds + lap / 2) % lap - lap / 2
            h = ds ** 6 / L ** 6 + (x[5] - ey) ** 6 / W ** 6 - 1.0 - margin
            hmin = min(hmin, h)
    return float(hmin)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--lib", choices=["emu", "product", "auto"], default="auto")
    ap.add_argument("--out", default=None)
    ap.add_argument("--track-layout", default="l_shape", help="data/track_layout/<name>.csv of the reference (the test scripts' --track-layout)")
    ap.add_argument("--only-ilqr", action="store_true", help="stop after the reference-vs-drop-in iLQR closed loop")
    ap.add_argument("--racing-game", action="store_true", help="only the LMPC + overtaking scenario")
    ap.add_argument("--lmpc-steps", type=int, default=30)
    ap.add_argument("--overtake-steps", type=int, default=60)
    args = ap.parse_args()
    import car_racing_b200 as crb
    from car_racing_b200 import _capi, batch
    lib = args.lib
    if lib == "auto":
        import torch
        lib = "product" if torch.cuda.is_available() else "emu"
    if lib == "emu":
        spec = importlib.util.spec_from_file_location("build_emu_library", os.path.join(ROOT, "tests", "host_emulation", "build_emu_library.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _capi.LIB_PATH, _capi._lib, batch._default_handle = mod.build(), None, None
    control, offboard, base, racing_env = import_reference()
    native = {n: getattr(control, n) for n in ("ilqr", "mpccbf", "mpc_lti", "mpc_multi_agents", "lmpc")}
    out = {"library": lib, "steps": args.steps, "seed": args.seed, "track_layout": args.track_layout, "scenarios": {}}
    if args.racing_game:
        r = racing_game(crb, control, offboard, base, racing_env, args.lmpc_steps, args.overtake_steps, args.seed, args.track_layout)
        out["scenarios"]["racing_game"] = r
        print("[racing_game]", json.dumps(r), flush=True)
        if args.out:
            with open(os.path.join(ROOT, args.out) if not os.path.isabs(args.out) else args.out, "w") as f:
                json.dump(out, f, indent=1)
        return

    # --- iLQR: unmodified reference vs drop-in, same closed loop (car_racing/tests/ilqr_test.py) -------------------------
    rivals = [(4.0, 0.2, 0.1)]
    sim, ego, track = build_sim(offboard, base, racing_env, "ilqr", rivals, args.track_layout)
    xr, ur, tr = run(sim, ego, args.steps, args.seed)
    crb.install(control)
    sim, ego, track = build_sim(offboard, base, racing_env, "ilqr", rivals, args.track_layout)
    xg, ug, tg = run(sim, ego, args.steps, args.seed)
    out["scenarios"]["ilqr_test"] = {
        "what": "car_racing/tests/ilqr_test.py, 1 rival; reference control.ilqr (numpy) vs car_racing_b200.install()",
        "max_abs_dx": float(np.abs(xr - xg).max()), "max_abs_du": float(np.abs(ur - ug).max()),
        "s_final_reference": float(xr[-1, 4]), "s_final_dropin": float(xg[-1, 4]),
        "wall_s_reference": tr, "wall_s_dropin": tg}
    print("[ilqr_test] %d steps: max|dx| %.2e max|du| %.2e (s_final %.4f / %.4f)" % (args.steps, np.abs(xr - xg).max(),
                                                                                    np.abs(ur - ug).max(), xr[-1, 4], xg[-1, 4]), flush=True)

    if args.only_ilqr:
        if args.out:
            with open(os.path.join(ROOT, args.out) if not os.path.isabs(args.out) else args.out, "w") as f:
                json.dump(out, f, indent=1)
        return

    # --- MPC-CBF (mpccbf_test.py) and MPC-LTI (control_test.py): drop-in only ----------------------------------------------
    statuses = []
    shim = control.mpccbf

    def mpccbf_logged(*a, **kw):
        u, r = shim(*a, return_details=True, **kw)
        statuses.append((int(r["status"][0]), int(r["iters"][0])))
        return u
    control.mpccbf = mpccbf_logged
    rivals = [(4.0, 0.2, 0.1), (10.0, 0.2, -0.1)]
    sim, ego, track = build_sim(offboard, base, racing_env, "mpccbf", rivals, args.track_layout)
    xg, ug, tg = run(sim, ego, args.steps, args.seed)
    st = np.array(statuses)
    out["scenarios"]["mpccbf_test"] = {
        "what": "car_racing/tests/mpccbf_test.py, 2 rivals; car_racing_b200.install() only (no CasADi here)",
        "s_final": float(xg[-1, 4]), "laps": int(ego.laps), "min_barrier_h": barrier(xg, rivals, track.lap_length),
        "max_abs_ey": float(np.abs(xg[:, 5]).max()), "track_half_width": float(track.width),
        "steps_converged": int((st[:, 0] == 0).sum()), "steps_total": int(len(st)), "iters_mean": float(st[:, 1].mean()),
        "iters_max": int(st[:, 1].max()), "u_within_bounds": bool((np.abs(ug[:, 0]) <= 0.5 + 1e-9).all() and (np.abs(ug[:, 1]) <= 1.0 + 1e-9).all()),
        "wall_s_dropin": tg}
    print("[mpccbf_test] %d steps: s_final %.3f, min h %.3f, max|ey| %.3f, converged %d/%d, iters mean %.1f max %d" % (
        args.steps, xg[-1, 4], out["scenarios"]["mpccbf_test"]["min_barrier_h"], np.abs(xg[:, 5]).max(), (st[:, 0] == 0).sum(), len(st),
        st[:, 1].mean(), st[:, 1].max()), flush=True)
    control.mpccbf = shim

    sim, ego, track = build_sim(offboard, base, racing_env, "mpc_lti", [], args.track_layout)
    xg, ug, tg = run(sim, ego, args.steps, args.seed)
    out["scenarios"]["control_test_mpc_lti"] = {
        "what": "car_racing/tests/control_test.py --ctrl-policy mpc-lti, no rivals; drop-in only (raises if a solve does not converge)",
        "s_final": float(xg[-1, 4]), "vx_final": float(xg[-1, 0]), "max_abs_ey": float(np.abs(xg[:, 5]).max()), "wall_s_dropin": tg}
    print("[control_test mpc-lti] %d steps: s_final %.3f vx_final %.3f max|ey| %.3f" % (args.steps, xg[-1, 4], xg[-1, 0], np.abs(xg[:, 5]).max()), flush=True)
    for n, f in native.items():
        setattr(control, n, f)
    if args.out:
        with open(os.path.join(ROOT, args.out) if not os.path.isabs(args.out) else args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

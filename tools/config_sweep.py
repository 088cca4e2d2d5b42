"""Secondary measurements: batch sweep of the north-star configuration and throughput of
configs 3 (planner candidates), 4 (LMPC) and 5 (iLQR) through the host-pointer C-ABI (H2D + kernel + D2H),
with the CPU oracle port on the same inputs beside each.  Not the bench contract -- bench.py is; this
writes one JSON document to stdout / --out for profiles/.

    python tools/config_sweep.py --out profiles/r01k_config_sweep.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import car_racing_b200 as crb                      # noqa: E402
from car_racing_b200 import scenarios              # noqa: E402


def timed(fn, reps=7, warm=2):
    for _ in range(warm):
        out = fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts)
    return out, float(np.median(ts)), float(np.percentile(ts, 10)), float(np.percentile(ts, 90))


def cpu_timed(fn):
    t0 = time.perf_counter()
    out = fn()
    return out, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    from oracle import oracle
    ncpu = os.cpu_count() or 1
    doc = {"host_threads": ncpu, "timing": "median wall time of the host-pointer call (H2D + kernel + D2H), 7 reps after 2 warm-ups"}

    # ---- config 2 batch sweep
    prm = scenarios.default_cbf_params(N=20)
    sweep = []
    for B in (1, 8, 64, 256, 1024, 2048, 4096, 8192, 16384, 65536):
        x0, xt, obs, lo = scenarios.mpccbf_scenarios(min(B, 8192), N=20, M=3, seed=1)
        if B > 8192:
            rep = B // 8192
            x0, obs, lo = np.tile(x0, (rep, 1)), np.tile(obs, (rep, 1, 1, 1)), np.tile(lo, (rep, 1))
        rec, M, ps = crb.pack_cbf(x0, xt, obs, lo, 20)
        g, t, p10, p90 = timed(lambda: crb.solve_cbf_packed(rec, prm, M, ps, want=()), reps=5 if B >= 16384 else 7)
        sweep.append(dict(B=B, ms=t * 1e3, ms_p10=p10 * 1e3, ms_p90=p90 * 1e3, solves_per_s=B / t,
                          converged=float((g["status"] == 0).mean()), iters_mean=float(g["iters"].mean()),
                          iters_max=int(g["iters"].max())))
        print("config2", sweep[-1], flush=True)
    doc["config2_batch_sweep"] = sweep

    # ---- config 3: planner candidate QPs (N=10), 64 candidates per planner call
    from test_gpu_parity import _planner_batch
    kw, off, pprm = _planner_batch(range(32))
    C = kw["x0"].shape[0]
    g, t, p10, p90 = timed(lambda: crb.solve_cbf_batch(kw["x0"], kw["xt"], kw["obs"], None, pprm, xlb=kw["xlb"], xub=kw["xub"],
                                                       wd=kw["wd"], want=("x",)))
    ent = dict(candidates=C, ms=t * 1e3, candidates_per_s=C / t, converged=float((g["status"] == 0).mean()))
    if not a.no_cpu:
        r, tc = cpu_timed(lambda: oracle.solve_cbf_batch(kw["x0"], kw["xt"], kw["obs"], None, pprm, xlb=kw["xlb"], xub=kw["xub"],
                                                         wd=kw["wd"], nthreads=ncpu))
        ent.update(cpu_port_ms=tc * 1e3, cpu_port_candidates_per_s=C / tc)
    doc["config3_planner"] = ent
    print("config3", ent, flush=True)

    # ---- overtaking step: planner + tracking MPC as two reference-shaped calls vs the fused device chain
    import types
    from planner_cases import make_planner
    from test_shims_host import Rival
    from car_racing_b200 import control, planning
    mp = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 5.0, 0, 50.0]),
                               matrix_R=np.diag([0.1, 0.1]), num_horizon_ctrl=10)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)

    def mk():
        pl = make_planner(7, num_veh=3)
        for name in pl.sorted_vehicles:
            tr = pl.obs_infos[name]
            pl.vehicles[name] = Rival(tr[4, 0], tr[0, 0], tr[5, 0])
        return pl
    pl = mk()
    xc = np.asarray(pl.vehicles["ego"].xcurv, float).copy()

    def two_calls():
        t2, f2, _, _ = planning.solve_optimization_problem(pl)
        return control.mpc_multi_agents(xc, mp, pl.track, None, None, None, sysp, target_traj_xcurv=t2, vehicles=pl.vehicles,
                                        agent_name="ego", direction_flag=f2, sorted_vehicles=pl.sorted_vehicles, time=None)
    _, t_two, _, _ = timed(two_calls, reps=21, warm=3)
    _, t_one, _, _ = timed(lambda: planning.plan_and_track(pl, xc, mp, pl.track, sysp, time=None), reps=21, warm=3)
    # the same step from the rivals' predictions on, candidates prepared on the device
    pl3 = mk()
    for name in pl3.sorted_vehicles:
        pl3.vehicles[name].no_dynamics = True
    pl3.racing_game_param.timestep, pl3.racing_game_param.planning_prediction_factor = 0.1, 0.5
    s_tab = np.linspace(0.0, pl3.track.lap_length + 1.0, 64)
    pl3.opti_traj_xcurv = np.zeros((64, 6))
    pl3.opti_traj_xcurv[:, 4], pl3.opti_traj_xcurv[:, 5] = s_tab, 0.2 * np.sin(s_tab)
    interest = {n: pl3.vehicles[n] for n in pl3.sorted_vehicles}
    _, t_prep, _, _ = timed(lambda: planning.plan_and_track_from_predictions(pl3, xc, 0.0, interest, xc, mp, pl3.track, sysp),
                            reps=21, warm=3)
    doc["overtaking_step"] = dict(candidates=len(pl.sorted_vehicles) + 1, two_calls_ms=t_two * 1e3, fused_chain_ms=t_one * 1e3,
                                  from_predictions_ms=t_prep * 1e3,
                                  note="median wall time incl. host packing; two_calls = solve_optimization_problem + mpc_multi_agents shims, "
                                       "fused = planning.plan_and_track (b200mpc_plan_and_track: one stream, no host round trip); "
                                       "from_predictions = planning.plan_and_track_from_predictions (b200mpc_plan_and_track_prepared: also "
                                       "the Bezier references, targets, bounds and candidate records are made on the device; the time "
                                       "includes the rivals' predictions, which the other two get ready-made)")
    print("overtaking_step", doc["overtaking_step"], flush=True)

    # ---- config 4: LMPC, 512 per GPU
    lprm = scenarios.default_lmpc_params()
    out4 = []
    for B in (1, 512, 4096):
        sc = scenarios.lmpc_scenarios(min(B, 512), seed=3)
        if B > 512:
            sc = tuple(np.tile(v, (B // 512,) + (1,) * (v.ndim - 1)) for v in sc)
        g, t, p10, p90 = timed(lambda: crb.solve_lmpc_batch(*sc, lprm, want=()))
        ent = dict(B=B, ms=t * 1e3, solves_per_s=B / t, converged=float((g["status"] == 0).mean()), iters_mean=float(g["iters"].mean()))
        if not a.no_cpu and B == 512:
            r, tc = cpu_timed(lambda: oracle.solve_lmpc_batch(*sc, lprm, nthreads=ncpu))
            ent.update(cpu_port_ms=tc * 1e3, cpu_port_solves_per_s=B / tc)
        out4.append(ent)
        print("config4", ent, flush=True)
    doc["config4_lmpc"] = out4

    # ---- config 5: iLQR N=50, 1024 per GPU
    p = scenarios.default_cbf_params()
    iprm = dict(A=p["A"], B=p["B"], Q=p["Q"], R=p["R"], N=50, max_iter=150, L=0.4, W=0.2)
    out5 = []
    for B in (1, 1024, 8192):
        x0, xt, obs, lo = scenarios.ilqr_scenarios(B, N=50, seed=1)
        g, t, p10, p90 = timed(lambda: crb.solve_ilqr_batch(x0, xt, obs, lo, iprm, want=()))
        ent = dict(B=B, ms=t * 1e3, solves_per_s=B / t, iters_mean=float(g["iters"].mean()))
        if not a.no_cpu and B == 1024:
            r, tc = cpu_timed(lambda: oracle.solve_ilqr_batch(x0, xt, obs, lo, iprm, nthreads=ncpu))
            ent.update(cpu_port_ms=tc * 1e3, cpu_port_solves_per_s=B / tc)
        out5.append(ent)
        print("config5", ent, flush=True)
    doc["config5_ilqr"] = out5
    # ---- batches in flight for configs 4 and 5 (per-GPU batch of the BASELINE configs, 4 streams)
    from car_racing_b200 import batch as _b

    def pipelined(pipe, rec, nb=24):
        for i in range(pipe.depth):
            pipe.submit(rec)
        for i in range(pipe.depth):
            pipe.result(i)
        base = pipe.n_submitted
        t0 = time.perf_counter()
        for i in range(nb):
            if i >= pipe.depth:
                pipe.result(base + i - pipe.depth, copy=False)
            pipe.submit(rec, copy=True)
        for i in range(nb - pipe.depth, nb):
            pipe.result(base + i, copy=False)
        dt = time.perf_counter() - t0
        return nb * rec.shape[0] / dt
    x0, xt, obs, lo = scenarios.ilqr_scenarios(1024, N=50, seed=1)
    ip = _b.IlqrPipeline(iprm, B=1024, depth=4)
    doc["config5_ilqr_in_flight"] = dict(B=1024, depth=4, solves_per_s=pipelined(ip, _b.pack_ilqr(x0, xt, obs, lo, 50)))
    ip.close()
    sc = scenarios.lmpc_scenarios(512, seed=3)
    rec4, K4 = _b.pack_lmpc(*sc, int(lprm["N"]))
    lp = _b.LmpcPipeline(lprm, K4, B=512, depth=4)
    doc["config4_lmpc_in_flight"] = dict(B=512, depth=4, solves_per_s=pipelined(lp, rec4))
    lp.close()
    print("in flight", doc["config5_ilqr_in_flight"], doc["config4_lmpc_in_flight"], flush=True)
    doc["reference_python_ilqr_note"] = "SURVEY.md 8(d): the reference's own numpy iLQR measured p50 20.1 ms per solve (1 core)"
    s = json.dumps(doc, indent=1)
    if a.out:
        with open(a.out, "w") as f:
            f.write(s + "\n")
    print(s)


if __name__ == "__main__":
    main()

"""One (warm-up + timed) launch sequence of the north-star kernel at a chosen batch size -- the target of the
ncu captures under profiles/ (a crowded launch, B >> 1036 resident instances, shows what bounds the SM when
every warp slot is busy; B=1024 shows the straggler tail).

    ncu --set full --clock-control none --import-source on -k regex:ocp_ipm -s 1 -c 1 -o out python tools/one_launch.py --B 8192
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import car_racing_b200 as crb                      # noqa: E402
from car_racing_b200 import scenarios              # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8192)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--kind", default="cbf", choices=["cbf", "lmpc", "ilqr", "planner"])
a = ap.parse_args()
if a.kind == "lmpc":     # BASELINE config 4: N=12, 44 safe-set points (ncu: -k regex:lmpc_kernel)
    sc = scenarios.lmpc_scenarios(min(a.B, 512), seed=3)
    if a.B > 512:
        sc = tuple(np.tile(v, (a.B // 512,) + (1,) * (v.ndim - 1)) for v in sc)
    lprm = scenarios.default_lmpc_params()
    fn = lambda: crb.solve_lmpc_batch(*sc, lprm, want=())
elif a.kind == "ilqr":   # BASELINE config 5: N=50 (ncu: -k regex:ilqr_kernel)
    p = scenarios.default_cbf_params()
    iprm = dict(A=p["A"], B=p["B"], Q=p["Q"], R=p["R"], N=50, max_iter=150, L=0.4, W=0.2)
    x0, xt, obs, lo = scenarios.ilqr_scenarios(a.B, N=50, seed=1)
    fn = lambda: crb.solve_ilqr_batch(x0, xt, obs, lo, iprm, want=())
elif a.kind == "planner":   # BASELINE config 3: 64 candidate QPs per planner call (ncu: -k regex:ocp_ipm)
    from car_racing_b200 import planning
    sc = scenarios.planner_scenarios(C=a.B, N=10, seed=1)
    kw, _ = planning.pack_candidates(sc["x0"], sc["s_ref"], sc["ey_ref"], sc["xlb"], sc["xub"], 10)
    pprm = planning.planner_params(scenarios.LTI_A, scenarios.LTI_B, 10)
    fn = lambda: crb.solve_cbf_batch(kw["x0"], kw["xt"], kw["obs"], None, pprm, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"], want=())
if a.kind != "cbf":
    g = fn()
    for _ in range(a.reps):
        t0 = time.perf_counter()
        g = fn()
        dt = time.perf_counter() - t0
        print(f"{a.kind} B={a.B} {dt * 1e3:.2f} ms  {a.B / dt:.0f} solves/s  iters mean {g['iters'].mean():.1f} max {g['iters'].max()}")
    sys.exit(0)
prm = scenarios.default_cbf_params(N=20)
x0, xt, obs, lo = scenarios.mpccbf_scenarios(min(a.B, 8192), N=20, M=3, seed=1)
if a.B > 8192:
    rep = a.B // 8192
    x0, obs, lo = np.tile(x0, (rep, 1)), np.tile(obs, (rep, 1, 1, 1)), np.tile(lo, (rep, 1))
rec, M, ps = crb.pack_cbf(x0, xt, obs, lo, 20)
g = crb.solve_cbf_packed(rec, prm, M, ps, want=())
for _ in range(a.reps):
    t0 = time.perf_counter()
    g = crb.solve_cbf_packed(rec, prm, M, ps, want=())
    dt = time.perf_counter() - t0
    print(f"B={a.B} {dt * 1e3:.2f} ms  {a.B / dt:.0f} solves/s  iters mean {g['iters'].mean():.1f} max {g['iters'].max()}")

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
a = ap.parse_args()
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

"""Parity of the CUDA path against the CPU oracle at the full BASELINE batch (config 2, B=1024, seeds 0..3) --
the same comparison the -m gpu tests make on smaller batches, written to profiles/ as a JSON document.

    python tools/parity_report.py --out profiles/r01_parity_report.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import car_racing_b200 as crb                      # noqa: E402
from car_racing_b200 import scenarios              # noqa: E402
import oracle as orc                               # noqa: E402  (test infrastructure: the checker)

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
a = ap.parse_args()
prm = scenarios.default_cbf_params(N=20)
doc = {"config": "MPC-CBF N=20, 3 static rivals, l_shape, B=1024 per seed", "tolerance": {"du": 1e-4, "dcost": 1e-5}, "seeds": []}
for seed in range(4):
    x0, xt, obs, lo = scenarios.mpccbf_scenarios(1024, N=20, M=3, seed=seed)
    g = crb.solve_cbf_batch(x0, xt, obs, lo, prm)
    r = orc.solve_cbf_batch(x0, xt, obs, lo, prm, nthreads=os.cpu_count() or 1)
    both = (g["status"] == 0) & (r["status"] == 0)
    du = np.abs(g["u0"] - r["u0"]).max(axis=1)
    dc = np.abs(g["cost"] - r["cost"])
    ok = both & (du < 1e-4) & (dc < 1e-5)
    ent = dict(seed=seed, gpu_converged=float((g["status"] == 0).mean()), oracle_converged=float((r["status"] == 0).mean()),
               both_converged=int(both.sum()), within_tolerance=int(ok.sum()), same_status=float((g["status"] == r["status"]).mean()),
               same_iteration_count=float((g["iters"] == r["iters"]).mean()),
               median_du=float(np.median(du[both])), p99_du=float(np.percentile(du[both], 99)), max_du_within=float(du[ok].max()),
               median_dcost=float(np.median(dc[both])), max_dcost_within=float(dc[ok].max()),
               outside_tolerance=int((both & ~ok).sum()),
               outside_are_other_local_optima=int((both & ~ok & (np.abs(g["cost"] - r["cost"]) > 1e-3)).sum()),
               gpu_cost_lower_on_outside=int((both & ~ok & (g["cost"] < r["cost"])).sum()),
               x_pred_max_diff_within=float(np.abs(g["x"][ok] - r["x"][ok]).max()),
               sigma_max_diff_within=float(np.abs(g["sigma"][ok] - r["sigma"][ok]).max()))
    doc["seeds"].append(ent)
    print(ent, flush=True)
s = json.dumps(doc, indent=1)
if a.out:
    open(a.out, "w").write(s + "\n")

"""Zero start (what Opti/IPOPT start from) against the roll-out start on the BASELINE batch, on the GPU, plus the KKT
certificate and the oracle comparison -- the three numbers of VERDICT r1 item 1 in one document:

    python tools/start_report.py --out profiles/r05_start_report.json [--seeds 4]

  certificate pass rate : tests/kkt_check.py on every converged, non-elastic instance of the roll-out-start solutions
  same-basin du / dcost : roll-out start vs zero start where both converge (median / p99 / max)
  mismatch rate         : share of those whose u0 differs by more than 1e-4 (a different KKT point)
The zero start needs IPOPT's iteration budget (max_iter 3000) and unlimited slack resets (our stand-in for the restoration
phase) to get anywhere: both settings are reported."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import car_racing_b200 as crb                      # noqa: E402
from car_racing_b200 import scenarios              # noqa: E402
from kkt_check import certificate                  # noqa: E402  (test infrastructure: the checker)

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--seeds", type=int, default=4)
ap.add_argument("--batch", type=int, default=1024)
a = ap.parse_args()
prm = scenarios.default_cbf_params(N=20)
doc = {"config": "MPC-CBF N=20, 3 static rivals, l_shape, B=%d per seed, GPU (libb200mpc.so)" % a.batch, "seeds": []}
tot = dict(n=0, cert=0, cert_pass=0, both=0, mismatch=0)
for seed in range(a.seeds):
    x0, xt, obs, lo = scenarios.mpccbf_scenarios(a.batch, N=20, M=3, seed=seed)
    t0 = time.perf_counter()
    g = crb.solve_cbf_batch(x0, xt, obs, lo, prm)
    t_roll = time.perf_counter() - t0
    ent = dict(seed=seed, rollout=dict(converged=float((g["status"] == 0).mean()), iters_mean=float(g["iters"].mean()),
                                       elastic=int((g["elastic_max"] > 1e-6).sum()), ms=1e3 * t_roll))
    n = npass = 0
    for b in range(a.batch):
        if g["status"][b] != 0 or g["elastic_max"][b] > 1e-7:
            continue
        c = certificate(x0[b], xt, obs[b], lo[b], prm, g["x"][b], g["u"][b], g["sigma"][b])
        n += 1
        npass += c["dyn"] < 1e-8 and c["row_viol"] < 1e-6 and c["stat"] < 1e-4 and c["comp"] < 1e-4
    ent["certificate"] = dict(certified=n, passed=int(npass), pass_rate=npass / max(n, 1))
    for label, kw in (("zero_ipopt_budget", dict(start=1, max_iter=3000, max_reset=5)),
                      ("zero_unlimited_resets", dict(start=1, max_iter=3000, max_reset=100000))):
        t0 = time.perf_counter()
        z = crb.solve_cbf_batch(x0, xt, obs, lo, prm, **kw)
        dt = time.perf_counter() - t0
        both = (g["status"] == 0) & (z["status"] == 0) & (g["elastic_max"] <= 1e-6) & (z["elastic_max"] <= 1e-6)
        du = np.abs(g["u0"] - z["u0"]).max(axis=1)
        dc = np.abs(g["cost"] - z["cost"])
        e = dict(options=kw, converged=float((z["status"] == 0).mean()), status_counts=np.bincount(z["status"], minlength=5).tolist(),
                 iters_median=float(np.median(z["iters"])), iters_max=int(z["iters"].max()), elastic=int((z["elastic_max"] > 1e-6).sum()),
                 ms=1e3 * dt, both_converged_nonelastic=int(both.sum()))
        if both.any():
            e.update(du_median=float(np.median(du[both])), du_p99=float(np.percentile(du[both], 99)), du_max=float(du[both].max()),
                     dcost_median=float(np.median(dc[both])), dcost_p99=float(np.percentile(dc[both], 99)),
                     mismatch=int((du[both] > 1e-4).sum()), mismatch_rate=float((du[both] > 1e-4).mean()),
                     zero_start_cost_lower=int(((z["cost"] < g["cost"] - 1e-5) & both).sum()),
                     zero_start_cost_higher=int(((z["cost"] > g["cost"] + 1e-5) & both).sum()))
        ent[label] = e
        if label == "zero_unlimited_resets":
            tot["both"] += int(both.sum())
            tot["mismatch"] += e.get("mismatch", 0)
    tot["n"] += a.batch
    tot["cert"] += n
    tot["cert_pass"] += int(npass)
    doc["seeds"].append(ent)
    print(json.dumps(ent), flush=True)
doc["summary"] = dict(instances=tot["n"], certificate_pass_rate=tot["cert_pass"] / max(tot["cert"], 1),
                      both_starts_converged=tot["both"], basin_mismatch_rate=tot["mismatch"] / max(tot["both"], 1))
print(json.dumps(doc["summary"]))
if a.out:
    json.dump(doc, open(a.out, "w"), indent=1)

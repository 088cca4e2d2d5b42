"""Per-phase cycle counters of ocp_ipm_kernel<3,QDIAG,20> at 1, 4 and 7 resident warps per SM -- which phases slow down when
the SM is crowded?  Needs a library built with -DB200MPC_PHASE_CLOCKS (tools/variants.sh build clocks:"-DB200MPC_PHASE_CLOCKS"):
that build writes the counters over the first 13 doubles of each instance's x_pred slot.

    python tools/phase_clocks.py variants_build/libb200mpc_clocks.so [--out gpurun_out/phase_clocks.json]

The same 148 scenarios are tiled 1x, 4x, 7x (B = 148, 592, 1036: one wave each, every copy does identical work), so the
ratio of a phase's cycles between the runs is the slowdown caused by sharing the SM, not by a different instance mix."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from car_racing_b200 import _capi                  # noqa: E402

lib = sys.argv[1]
out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
_capi.LIB_PATH = os.path.abspath(lib)
import car_racing_b200 as crb                      # noqa: E402
from car_racing_b200 import scenarios              # noqa: E402

NAMES = ["eval+error+mu", "assemble", "riccati_backward", "riccati_forward", "rows+step bounds", "line search", "accept+update",
         "(loop top)", "bw: P[A B]", "bw: G column", "bw: chol+solves", "bw: P update", "bw: -"]
ORDER = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12]
prm = scenarios.default_cbf_params(N=20)
x0, xt, obs, lo = scenarios.mpccbf_scenarios(148, N=20, M=3, seed=1)
doc = {"library": lib, "runs": []}
base = None
for rep in (1, 4, 7):
    X0, OBS, LO = np.tile(x0, (rep, 1)), np.tile(obs, (rep, 1, 1, 1)), np.tile(lo, (rep, 1))
    rec, M, ps = crb.pack_cbf(X0, xt, OBS, LO, 20)
    crb.solve_cbf_packed(rec, prm, M, ps, want=("x",))
    g = crb.solve_cbf_packed(rec, prm, M, ps, want=("x",))
    pc = g["x"].reshape(len(X0), -1)[:, :13]
    # pc[0..7] are indexed by PCLK(k): k = 7 loop top, 0 eval, 1 assemble, 2 backward, 3 forward, 4 rows, 5 line search, 6 accept
    tot = pc[:, :8].sum(axis=1)
    mean = pc.mean(axis=0)
    ent = {"warps_per_sm": rep, "B": len(X0), "iters_mean": float(g["iters"].mean()),
           "cycles_per_instance_mean": float(tot.mean()),
           "phase_cycles_mean": {NAMES[k if k < 7 else k]: float(mean[k]) for k in range(13)}}
    if base is None:
        base = mean
    ent["slowdown_vs_lone"] = {NAMES[k]: (float(mean[k] / base[k]) if base[k] > 0 else None) for k in range(13)}
    ent["share_pct"] = {NAMES[k]: float(100 * mean[k] / mean[:8].sum()) for k in range(8)}
    doc["runs"].append(ent)
    print(json.dumps(ent))
if out:
    json.dump(doc, open(out, "w"), indent=1)

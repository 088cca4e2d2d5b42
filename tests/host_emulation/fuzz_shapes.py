"""Shape sweep of the C-ABI under AddressSanitizer on the host-compiled library (tests/host_emulation/sanitize_kernels.sh builds it the same
way): every (N, M, per-stage target) combination at the corners of the supported ranges, odd batches, the iLQR and LMPC
size limits.  Out-of-bounds shared / global accesses of a kernel show up as AddressSanitizer errors; the solves must also
agree with the oracle where both converge.  Test infrastructure: never used by the product.

    ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 LD_PRELOAD=$(g++ -print-file-name=libasan.so) B200MPC_EMU_ASAN=1 python tests/host_emulation/fuzz_shapes.py
"""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    spec = importlib.util.spec_from_file_location("build_emu_library", os.path.join(ROOT, "tests", "host_emulation", "build_emu_library.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if os.environ.get("B200MPC_EMU_ASAN"):
        mod.LIB = os.path.join(mod.OUT, "libb200mpc_emu_asan.so")
    path = mod.build()
    import car_racing_b200 as crb
    from car_racing_b200 import _capi, batch, scenarios
    import oracle as orc
    orc.build()
    _capi.LIB_PATH, _capi._lib, batch._default_handle = path, None, None
    bad = 0
    t0 = time.time()
    for N in (1, 2, 3, 5, 20, 31, 32, 33, 63, 64):
        for M in (0, 1, 2, 3, 4):
            for per_stage in (False, True):
                B = 1 if N > 33 else 2
                x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=N, M=M, seed=N + M)
                if N > 30 and M:
                    obs[:, :, 0, :] += 3.0
                if per_stage:
                    xt = np.tile(np.asarray(xt, float).reshape(1, 1, 6), (B, N + 1, 1))
                prm = scenarios.default_cbf_params(N=N)
                g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
                r = orc.solve_cbf_batch(x0, xt, obs, lap_off, prm)
                both = (g["status"] == 0) & (r["status"] == 0)
                du = np.abs(g["u0"] - r["u0"])[both].max() if both.any() else 0.0
                ok = (g["status"] == r["status"]).all() and du < 1e-4
                bad += 0 if ok else 1
                print("cbf N=%2d M=%d per_stage=%d: status %s / %s, max|du0| %.1e %s" % (N, M, per_stage, g["status"].tolist(),
                                                                                        r["status"].tolist(), du, "" if ok else "  <-- MISMATCH"),
                      flush=True)
    p = scenarios.default_cbf_params()
    for N in (1, 2, 7, 50, 64):
        x0, xt, obs, lap_off = scenarios.ilqr_scenarios(3, N=N, seed=N)
        prm = dict(A=p["A"], B=p["B"], Q=p["Q"], R=p["R"], N=N, max_iter=150, L=0.4, W=0.2)
        g, r = crb.solve_ilqr_batch(x0, xt, obs, lap_off, prm), orc.solve_ilqr_batch(x0, xt, obs, lap_off, prm)
        same = g["iters"] == r["iters"]
        du = np.abs(g["u0"] - r["u0"])[same].max() if same.any() else 0.0
        ok = same.all() and du < 1e-8
        bad += 0 if ok else 1
        print("ilqr N=%2d: iterations %s / %s, max|du0| %.1e %s" % (N, g["iters"].tolist(), r["iters"].tolist(), du, "" if ok else "  <-- MISMATCH"),
              flush=True)
    for N, K in ((2, 1), (2, 64), (16, 64), (16, 2), (12, 44), (7, 33)):
        ni = 1 if K % 2 else 2
        sc = scenarios.lmpc_scenarios(1, N=N, num_ss_points=K, num_ss_iter=ni, seed=N + K)
        prm = scenarios.default_lmpc_params(N=N, Q=np.diag([0.5, 0, 0, 0.1, 0, 2.0]))
        g, r = crb.solve_lmpc_batch(*sc, prm), orc.solve_lmpc_batch(*sc, prm)
        both = (g["status"] == 0) & (r["status"] == 0)
        du = np.abs(g["u0"] - r["u0"])[both].max() if both.any() else 0.0
        ok = (g["status"] == r["status"]).all() and du < 1e-4
        bad += 0 if ok else 1
        print("lmpc N=%2d K=%2d: status %s / %s, max|du0| %.1e %s" % (N, K, g["status"].tolist(), r["status"].tolist(), du,
                                                                     "" if ok else "  <-- MISMATCH"), flush=True)
    print("summary: %d mismatching cases, %.0f s" % (bad, time.time() - t0))


if __name__ == "__main__":
    main()

"""Driver of tests/host_emulation/sanitize_kernels.sh: builds the host-compiled C-ABI library with -fsanitize=thread or =address (B200MPC_EMU_TSAN / _ASAN), points
car_racing_b200._capi at it and runs every kernel on a few instances through the product's batch API; ThreadSanitizer's
reports go to stderr.  Test infrastructure: never used by the product."""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    spec = importlib.util.spec_from_file_location("build_emu_library", os.path.join(ROOT, "tests", "host_emulation", "build_emu_library.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.LIB = os.path.join(mod.OUT, "libb200mpc_emu_%s.so" % ("asan" if os.environ.get("B200MPC_EMU_ASAN") else "tsan"))
    path = mod.build(force=True)
    import car_racing_b200 as crb
    from car_racing_b200 import _capi, batch, planning, scenarios
    _capi.LIB_PATH, _capi._lib, batch._default_handle = path, None, None

    def report(name, fn):
        t = time.time()
        info = fn()
        print("kernel %s: %s, %.1f s" % (name, info, time.time() - t), flush=True)

    def cbf():
        x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(2, N=20, M=3, seed=0)
        r = crb.solve_cbf_batch(x0, xt, obs, lap_off, scenarios.default_cbf_params(N=20))
        return "2 instances, iterations %s, status %s" % (r["iters"].tolist(), r["status"].tolist())

    def ilqr():
        p = scenarios.default_cbf_params()
        x0, xt, obs, lap_off = scenarios.ilqr_scenarios(2, N=50, seed=2)
        r = crb.solve_ilqr_batch(x0, xt, obs, lap_off, dict(A=p["A"], B=p["B"], Q=p["Q"], R=p["R"], N=50, max_iter=150, L=0.4, W=0.2))
        return "2 instances, iterations %s" % r["iters"].tolist()

    def lmpc():
        r = crb.solve_lmpc_batch(*scenarios.lmpc_scenarios(1, seed=5), scenarios.default_lmpc_params())
        return "1 instance (128 threads), iterations %s, status %s" % (r["iters"].tolist(), r["status"].tolist())

    def sysid():
        g = np.load(os.path.join(GOLD, "sysid_golden.npz"))
        r = crb.estimate_abc_batch(g["lin_points"][:1], g["lin_input"][:1], g["ss"], g["us"], g["time_ss"], [0, 1], g["point_and_tangent"],
                                   float(g["dt"]), int(g["max_num_point"]))
        return "1 instance x 12 stages, status %s" % r["status"].tolist()

    def planner():
        from planner_cases import make_planner
        from test_shims_host import Rival
        import types
        p = make_planner(7, num_veh=2)
        for name in p.sorted_vehicles:
            tr = p.obs_infos[name]
            p.vehicles[name] = Rival(tr[4, 0], tr[0, 0], tr[5, 0])
        prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 5.0, 0, 50.0]),
                                    matrix_R=np.diag([0.1, 0.1]), num_horizon_ctrl=10)
        sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
        x = np.asarray(p.vehicles["ego"].xcurv, float).copy()
        (_, flag, _, _), _ = planning.plan_and_track(p, x, prm, p.track, sysp, time=None)
        return "3 candidates + selection + tracking solve, region %d, tracking status %d" % (flag, p.tracking_status)

    def exchange():
        import ctypes as C
        L = _capi.lib()
        hs = [_capi.Handle(), _capi.Handle()]
        cs, blobs = [], []
        for r in range(2):
            c = C.c_void_p()
            hs[r].check(L.b200mpc_comm_create(hs[r].ptr, r, 2, 4, 2, C.byref(c)), "comm_create")
            buf = (C.c_char * _capi.COMM_HANDLE_BYTES)()
            L.b200mpc_comm_export(c, buf)
            cs.append(c)
            blobs.append(bytes(buf.raw))
        for r in range(2):
            assert L.b200mpc_comm_connect(cs[r], b"".join(blobs)) == 0
        prm = scenarios.default_cbf_params(N=10)
        p, o = _capi.make_cbf_params(prm, 1, False), _capi.default_options()
        args = []
        for use in range(2):
            for r in range(2):
                x0, xt, obs, lo = scenarios.mpccbf_scenarios(2, N=10, M=1, seed=40 + r)
                rin, M, ps = batch.pack_cbf(x0, xt, obs, lo, 10)
                rec = np.zeros(2, dtype=_capi.RECORD_DTYPE)
                L.b200mpc_comm_publish_next(hs[r].ptr, cs[r], 0)
                hs[r].check(L.b200mpc_cbf_solve(hs[r].ptr, C.byref(p), C.byref(o), 2, batch._ptr(rin), batch._ptr(rec), None, None, None, None), "solve")
            for r in range(2):
                arg, allr = np.zeros(1, dtype=np.int32), np.zeros(4, dtype=_capi.RECORD_DTYPE)
                hs[r].check(L.b200mpc_comm_argmin(hs[r].ptr, cs[r], 0, 0, batch._ptr(arg), batch._ptr(allr)), "argmin")
                args.append(int(arg[0]))
        for c in cs:
            L.b200mpc_comm_destroy(c)
        return "2 ranks x 2 instances x 2 uses of one slot, argmin %s" % args

    for name, fn in (("ocp_ipm_kernel<3,QDIAG,20>", cbf), ("exchange window: publish epilogue + xchg_wait_ack_kernel + xchg_argmin_kernel", exchange), ("ilqr_kernel", ilqr), ("lmpc_kernel", lmpc), ("sysid_kernel", sysid),
                     ("ocp_ipm_kernel<0,3,0> + planner_select_kernel + ocp_ipm_kernel<M,0,0>", planner)):
        report(name, fn)
    print("summary: done")


if __name__ == "__main__":
    main()

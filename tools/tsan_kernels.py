"""Driver of tools/tsan_kernels.sh: runs the host-emulated kernels (built with -fsanitize=thread into $B200MPC_EMU_LIBDIR)
on a few instances each and prints one line per kernel; ThreadSanitizer's own reports go to stderr."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
LIBDIR = os.environ.get("B200MPC_EMU_LIBDIR", os.path.join(ROOT, "tests", "host_emulation", "_build"))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    from car_racing_b200 import scenarios
    import test_hot_kernel_on_host as hk
    L = C.CDLL(os.path.join(LIBDIR, "libocp_ipm_emu.so"))
    L.emu_cbf_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=20, M=3, seed=0)
    prm = scenarios.default_cbf_params(N=20)
    t = time.time()
    r = hk._solve(L, x0, xt, obs, lap_off, prm, specialised=1)
    print("kernel ocp_ipm_kernel<3,0,20>: %d instances, iterations %s, status %s, %.1f s" % (B, r["iters"].tolist(), r["status"].tolist(),
                                                                                              time.time() - t), flush=True)
    for name, fn in (("ilqr", "run_ilqr"), ("lmpc", "run_lmpc"), ("sysid", "run_sysid")):
        path = os.path.join(LIBDIR, "lib%s_emu.so" % name)
        if os.path.exists(path):
            import test_warp_kernels_on_host as wk
            print(getattr(wk, fn)(C.CDLL(path), B), flush=True)
    print("summary: done")


if __name__ == "__main__":
    main()

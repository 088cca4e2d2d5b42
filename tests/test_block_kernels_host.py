"""Block-level product kernels run on the host (tests/host_emulation/: one std::thread per CUDA thread, std::barrier for
__syncthreads, function-local statics for __shared__): planner_select_kernel and plant_kernel, the very sources the CUDA
build compiles, checked against the reference-pinned expectations without a GPU.  The CUDA build of the same sources is
checked by the -m gpu tests (tests/test_gpu_parity.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
from scipy.interpolate import interp1d

from car_racing_b200 import _capi, batch, planning

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "host_emulation")
GOLD = os.path.join(HERE, "golden")


def _lib():
    lib = os.path.join(EMU, "_build", "libblock_emu.so")
    src = [os.path.join(EMU, "block_kernels_host.cpp"), os.path.join(EMU, "cuda_runtime.h"),
           os.path.join(HERE, "..", "car_racing_b200", "csrc", "planner_select.cuh"),
           os.path.join(HERE, "..", "car_racing_b200", "csrc", "plant.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(f) > os.path.getmtime(lib) for f in src):
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-I", EMU, src[0], "-o", lib],
                       check=True)
    return C.CDLL(lib)


def test_plant_kernel_on_host_matches_reference_golden():
    L = _lib()
    L.emu_plant_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
    L.emu_plant_step.restype = None
    g = np.load(os.path.join(GOLD, "plant_golden.npz"))
    P = batch._ptr
    for name in ("ellipse", "l_shape"):
        xc, xg, us, dr = g["xcurv_" + name], g["xglob_" + name], g["u_" + name], g["draws_" + name]
        pat = g["pat_" + name]
        seg = np.ascontiguousarray(pat[:, 3:6])
        p = _capi.make_plant_params(seg.shape[0], float(g["lap_length_" + name]), dyn=tuple(g["dyn"]))
        for k in range(us.shape[1]):
            c, gl = np.ascontiguousarray(xc[:, k]).copy(), np.ascontiguousarray(xg[:, k]).copy()
            u, d = np.ascontiguousarray(us[:, k]), np.ascontiguousarray(dr[:, k])
            L.emu_plant_step(C.byref(p), c.shape[0], P(c), 6, 0, P(gl), P(u), 2, P(d), P(seg), None)
            assert np.abs(c - xc[:, k + 1]).max() < 1e-12 and np.abs(gl - xg[:, k + 1]).max() < 1e-12, (name, k)
    # a batch that spans several blocks, ragged last one, with the lap wrap and its counter
    Bn = 300
    c = np.tile(xc[:, 0], (Bn // xc.shape[0], 1)).copy()
    c[:, 4] = float(g["lap_length_l_shape"]) - 0.01
    gl, u = np.tile(xg[:, 0], (Bn // xc.shape[0], 1)).copy(), np.tile(us[:, 0], (Bn // xc.shape[0], 1)).copy()
    laps = np.zeros(c.shape[0], dtype=np.int32)
    p = _capi.make_plant_params(seg.shape[0], float(g["lap_length_l_shape"]), dyn=tuple(g["dyn"]), wrap_lap=True)
    L.emu_plant_step(C.byref(p), c.shape[0], P(c), 6, 0, P(gl), P(u), 2, None, P(seg), P(laps))
    assert (laps == 1).all() and (c[:, 4] < 1.0).all() and np.array_equal(c[:6], c[6:12])


def _select_case(seed, C_, num_veh, N=10, Nc=10, Mc=2):
    rng = np.random.default_rng(seed)
    lap = 19.13
    x = np.zeros((C_, N + 1, 6))
    s0 = rng.uniform(2.0, lap - 1.0)
    x[:, :, 4] = s0 + np.cumsum(rng.uniform(0.05, 0.2, size=(C_, N + 1)), axis=1)
    x[:, 0, 4] = s0
    x[:, :, 5] = rng.uniform(-0.8, 0.8, size=(C_, N + 1))
    heur = x + rng.normal(scale=0.05, size=x.shape) * np.array([0, 0, 0, 0, 0, 1.0])
    rec = np.zeros(C_, dtype=_capi.RECORD_DTYPE)
    rec["status"] = (rng.uniform(size=C_) < 0.2).astype(np.int32)          # some solves "failed"
    ok0 = (rng.uniform(size=C_) > 0.15).astype(np.int32)                   # some x_0 violate a stage-0 row
    region = np.concatenate([np.arange(num_veh + 1), rng.integers(0, num_veh + 1, size=C_ - num_veh - 1)]).astype(np.int32)
    rivals = np.zeros((num_veh, 2, N + 1))
    for j in range(num_veh):
        rivals[j, 0] = s0 + rng.uniform(0.2, 1.0) + 0.1 * rng.uniform(0.6, 1.2) * np.arange(N + 1) + (lap if j == 0 else 0.0)
        rivals[j, 1] = rng.uniform(-0.6, 0.6)
    return lap, x, heur, rec, ok0, region, rivals


def test_planner_select_kernel_on_host_matches_the_host_selection():
    """overtake_traj_planner.py:205-246 + control.py:373-382 as planner_select_kernel computes them, against
    planning.selection_costs (pinned to the reference's own selection by tests/test_reference_statement.py) and the target
    construction of the control.mpc_multi_agents shim."""
    L = _lib()
    L.emu_planner_select.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 10
    L.emu_planner_select.restype = None
    P = batch._ptr
    flags = set()
    for seed, C_, nv, old in ((0, 3, 2, None), (1, 4, 3, 1), (2, 64, 2, 0), (3, 200, 3, None), (4, 2, 1, None)):
        N, Nc, Mc = 10, 10, 2
        lap, x, heur, rec, ok0, region, rivals = _select_case(seed, C_, nv)
        sel = _capi.PlannerSelectParams()
        sel.C, sel.N, sel.num_veh, sel.N_ctrl, sel.M_ctrl = C_, N, nv, Nc, Mc
        sel.old_direction_flag = -1 if old is None else old
        sel.veh_length, sel.veh_width, sel.lap_length = 0.4, 0.2, lap
        stride = batch.cbf_record_doubles(Nc, Mc, True)
        hdr = (6 + Mc + 1) & ~1
        trk = np.zeros(stride)
        trk[0:6] = [1.3, 0.0, 0.0, 0.0, x[0, 0, 4] + 0.02, 0.1]
        sel_cost, flag, traj = np.zeros(C_), np.zeros(2, dtype=np.int32), np.zeros((N + 1, 6))
        L.emu_planner_select(C.byref(sel), hdr, P(rec), P(np.ascontiguousarray(x)), P(np.ascontiguousarray(heur)), P(ok0), P(region),
                             P(np.ascontiguousarray(rivals)), P(sel_cost), P(flag), P(traj), P(trk))
        # expectation: the host path
        solved = (ok0 != 0) & (rec["status"] == 0)
        sol = np.where(solved[:, None, None], x, heur)
        names = ["car%d" % j for j in range(nv)]
        obs_infos = {}
        for j, n in enumerate(names):
            tr = np.zeros((6, N + 1))
            tr[4], tr[5] = rivals[j, 0], rivals[j, 1]
            obs_infos[n] = tr
        cost = np.zeros(C_)
        for c in range(C_):      # planning.selection_costs with the candidate's region as its index (reference: region = index)
            one = planning.selection_costs(sol[c:c + 1].transpose(0, 2, 1), names, {n: obs_infos[n].copy() for n in names}, 0.4, 0.2, lap,
                                           None)
            # selection_costs takes the index as the region: shift by evaluating with the neighbours of region[c]
            r = int(region[c])
            val = -10 * (sol[c, -1, 4] - sol[c, 0, 4])
            for side in (r - 1, r):
                if 0 <= side < nv:
                    so = rivals[side, 0].copy()
                    while (so > lap).any():
                        so = np.where(so > lap, so - lap, so)
                    d2 = (sol[c, :, 4] - so) ** 2 + (sol[c, :, 5] - rivals[side, 1]) ** 2
                    val += 100 * int((d2 - 0.4 ** 2 - 0.2 ** 2 < 0).sum())
            if old is not None and old != c:
                val += 100
            cost[c] = val
            if r == 0 and c == 0 and old is None:
                assert abs(one[0] - val) < 1e-12      # agrees with the reference-pinned host function where both apply
        assert np.abs(sel_cost - cost).max() < 1e-12
        best = int(np.argmin(cost))                   # first minimum = list.index(min(list))
        assert flag[0] == best and flag[1] == region[best]
        assert np.array_equal(traj, sol[best])
        f = interp1d(sol[best][:, 4], sol[best][:, 5])
        for i in range(Nc + 1):                       # control.py:373-378
            s_tmp = trk[0] * 0.1 * i + trk[4]
            s_tmp = max(s_tmp, sol[best][0, 4])
            if s_tmp >= sol[best][-1, 4]:
                s_tmp = sol[best][-1, 4]
            t = trk[hdr + 6 * i: hdr + 6 * i + 6]
            assert t[0] == trk[0] and (t[1:5] == 0).all() and abs(t[5] - float(f(s_tmp))) < 1e-12
        flags.add(best)
    assert len(flags) > 1

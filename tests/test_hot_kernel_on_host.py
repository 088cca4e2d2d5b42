"""The hot kernel's own source on the host: car_racing_b200/csrc/ocp_ipm.cuh compiled by g++ and run with one host thread
per lane (tests/host_emulation/: std::barrier per warp for __syncwarp and the shuffle / reduce collectives, a synchronous
memcpy for the TMA bulk copy, exact seeds for the rcp / rsqrt approximations).  Parity of the kernel's LOGIC against the
oracle without a GPU -- every template instantiation the C-ABI dispatches to is exercised -- with BASELINE.json's
tolerances |du| < 1e-4, |dcost| < 1e-5 (the observed differences are ~1e-15).  The CUDA build of the same source is what
the -m gpu tests run."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from car_racing_b200 import _capi, batch, planning, scenarios

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "host_emulation")
TOL_U, TOL_C = 1e-4, 1e-5
pytestmark = pytest.mark.timeout(600)


def _lib():
    from conftest import wait_prebuilt          # g++ -std=c++20 -O1 -ffp-contract=off -pthread on ocp_ipm_host.cpp (tests/conftest.py)
    lib = wait_prebuilt("ocp_ipm_emu")
    L = C.CDLL(lib)
    L.emu_cbf_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
    return L


def _options():
    o = _capi.IpmOptions()       # b200mpc_default_ipm_options (car_racing_b200/csrc/capi.cu), set here because the library
    o.tol, o.max_iter, o.acceptable_iter, o.acceptable_tol = 1e-8, 200, 15, 1e-6     # itself is not loaded on the CPU tier
    o.mu_init, o.rho, o.bound_push, o.bound_frac, o.max_grad = 0.1, 1e3, 1e-2, 1e-2, 100.0
    o.start, o.max_reset = _capi.START_ROLLOUT, 5
    return o


def _solve(L, x0, xt, obs, lap_off, prm, specialised=1, xlb=None, xub=None, wd=None, sizes=None, **opt):
    N = int(prm["N"])
    records, M, per_stage = batch.pack_cbf(x0, xt, obs, lap_off, N, xlb=xlb, xub=xub, wd=wd, sizes=sizes)
    flags = batch.cbf_flags(M, xlb, xub, wd, sizes)
    p, o = _capi.make_cbf_params(prm, M, per_stage, flags), _options()
    for k, v in opt.items():
        setattr(o, k, v)
    B = records.shape[0]
    rec = np.zeros(B, dtype=_capi.RECORD_DTYPE)
    aux, x, u = np.zeros((B, 4)), np.zeros((B, N + 1, 6)), np.zeros((B, N, 2))
    sg = np.zeros((B, max(M, 1), N + 1))
    P = batch._ptr
    rc = L.emu_cbf_solve(C.byref(p), C.byref(o), B, specialised, P(np.ascontiguousarray(records)), P(rec), P(aux), P(x), P(u), P(sg))
    assert rc == 0
    return dict(u0=rec["u0"], cost=rec["cost"], status=rec["status"], iters=rec["iters"], x=x, u=u, sigma=sg[:, :M], kkt_err=aux[:, 0])


def _agree(g, r):
    assert (g["status"] == r["status"]).all(), (g["status"], r["status"])
    ok = g["status"] == 0
    assert ok.any()
    assert np.abs(g["u0"] - r["u0"])[ok].max() < TOL_U and np.abs(g["cost"] - r["cost"])[ok].max() < TOL_C
    assert np.abs(g["x"] - r["x"])[ok].max() < 1e-4 and np.abs(g["u"] - r["u"])[ok].max() < 1e-4
    return int(np.abs(g["iters"] - r["iters"]).max())


def test_north_star_configuration_on_host(oracle):
    """BASELINE config 2 (N=20, 3 rivals): the horizon-specialised instantiation <3,0,20> and the runtime-horizon one
    <3,0,0> against the oracle, and bit for bit against each other."""
    L = _lib()
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(4, N=20, M=3, seed=0)
    prm = scenarios.default_cbf_params(N=20)
    ref = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    spec = _solve(L, x0, xt, obs, lap_off, prm, specialised=1)
    assert _agree(spec, ref) == 0                               # the same iteration counts as the oracle
    assert np.abs(spec["sigma"] - ref["sigma"]).max() < 1e-6
    gen = _solve(L, x0[:2], xt, obs[:2], lap_off[:2], prm, specialised=0)
    assert np.array_equal(gen["u0"], spec["u0"][:2]) and np.array_equal(gen["iters"], spec["iters"][:2])
    assert np.array_equal(gen["x"], spec["x"][:2])


def test_other_instantiations_on_host(oracle):
    """MPC-LTI (<0,0,0>, config 1: N=10, no rivals), two rivals (<2,0,0>, N=12) and the planner's candidate QP (<0,3,0>:
    per-stage bounds + ey-rate cost)."""
    L = _lib()
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(3, N=10, M=0, seed=3)
    prm = scenarios.default_cbf_params(N=10)
    _agree(_solve(L, x0, xt, obs, lap_off, prm), oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm))
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(3, N=12, M=2, seed=4)
    prm = scenarios.default_cbf_params(N=12)
    _agree(_solve(L, x0, xt, obs, lap_off, prm), oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm))
    from planner_cases import make_planner
    p = make_planner(3, num_veh=2)
    N = p.racing_game_param.num_horizon_planner
    ego = np.asarray(p.vehicles["ego"].xcurv, float)
    s_ref, ey_ref, xlb, xub = [], [], [], []
    for c in range(3):
        lo, hi = planning.candidate_bounds(c, p.xcurv_ego, p.sorted_vehicles, p.obs_infos, 0.4, 0.2, p.track.width, p.track.lap_length, N)
        s, e = planning.candidate_targets(c, ego, p.bezier_xcurvs, p.bezier_funcs, N)
        s_ref.append(s); ey_ref.append(e); xlb.append(lo); xub.append(hi)
    kw, off = planning.pack_candidates(ego, np.array(s_ref), np.array(ey_ref), np.array(xlb), np.array(xub), N)
    pprm = planning.planner_params(p.racing_game_param.matrix_A, p.racing_game_param.matrix_B, N)
    ref = oracle.solve_cbf_batch(kw["x0"], kw["xt"], kw["obs"], None, pprm, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"])
    _agree(_solve(L, kw["x0"], kw["xt"], kw["obs"], None, pprm, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"]), ref)


def test_full_weight_matrix_on_host(oracle):
    """The reference's Q is diagonal in every configuration and the kernel takes a one-product-per-row path for it
    (KParams::q_diag, set by the host from Q); a Q with off-diagonal entries must take the general path and agree with the
    oracle as well, and a diagonal Q must give the same bits on both paths."""
    L = _lib()
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(3, N=12, M=2, seed=6)
    prm = scenarios.default_cbf_params(N=12)
    Q = np.array(prm["Q"], float)
    Q[0, 3] = Q[3, 0] = 1.5
    Q[3, 5] = Q[5, 3] = -2.0
    Q[1, 1] = 0.5
    prm_full = dict(prm, Q=Q)
    _agree(_solve(L, x0, xt, obs, lap_off, prm_full), oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm_full))
    tiny = np.array(prm["Q"], float)
    tiny[0, 1] = 1e-300            # numerically nothing, but not zero: the general path
    a = _solve(L, x0, xt, obs, lap_off, prm)
    b = _solve(L, x0, xt, obs, lap_off, dict(prm, Q=tiny))
    assert np.array_equal(a["iters"], b["iters"]) and np.abs(a["u0"] - b["u0"]).max() < 1e-12 and np.abs(a["x"] - b["x"]).max() < 1e-10
    # the north-star shape: <3, QDIAG, 20> (diagonal Q known at compile time) against <3, 0, 20> on its general path
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(2, N=20, M=3, seed=0)
    prm = scenarios.default_cbf_params(N=20)
    tiny = np.array(prm["Q"], float)
    tiny[0, 1] = 1e-300
    a = _solve(L, x0, xt, obs, lap_off, prm)
    b = _solve(L, x0, xt, obs, lap_off, dict(prm, Q=tiny))
    assert np.array_equal(a["iters"], b["iters"]) and np.abs(a["u0"] - b["u0"]).max() < 1e-12 and np.abs(a["x"] - b["x"]).max() < 1e-10


def test_zero_start_rival_sizes_and_x0_rows_on_host(oracle):
    """Round-2 options of the hot kernel against the oracle: B200MPC_START_ZERO (w = 0, what Opti/IPOPT start from), per-rival
    (L, W) in the record (flag RIVAL_SIZE, <2,4,0>), and status 4 when x_0 violates the bound rows the reference imposes on
    stage 0 (control.py:582-586)."""
    L = _lib()
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(3, N=12, M=2, seed=4)
    prm = scenarios.default_cbf_params(N=12)
    # zero start: a bounded number of iterations; the iterates must be the oracle's (same status, iteration count, point)
    kw = dict(start=_capi.START_ZERO, max_iter=60, max_reset=50)
    g, r = _solve(L, x0, xt, obs, lap_off, prm, **kw), oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, **kw)
    assert (g["status"] == r["status"]).all() and (g["iters"] == r["iters"]).all()
    assert np.abs(g["x"] - r["x"]).max() < 1e-6 and np.abs(g["u"] - r["u"]).max() < 1e-6
    # per-rival sizes: a long, narrow rival and a short, wide one
    sizes = np.array([[0.55, 0.18], [0.35, 0.27]])
    g = _solve(L, x0, xt, obs, lap_off, prm, sizes=sizes)
    r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, sizes=sizes)
    _agree(g, r)
    same = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    assert np.abs(r["u"] - same["u"]).max() > 1e-6              # the sizes are really used
    # x_0 outside |ey| <= width: the reference's NLP is infeasible
    x0b = x0.copy()
    x0b[0, 5] = prm["width"] + 0.05
    g, r = _solve(L, x0b, xt, obs, lap_off, prm), oracle.solve_cbf_batch(x0b, xt, obs, lap_off, prm)
    assert g["status"][0] == 4 and r["status"][0] == 4 and (g["status"][1:] == r["status"][1:]).all()

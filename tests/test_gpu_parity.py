"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle on the
same seeded inputs.  Tolerances are BASELINE.json's: |du| < 1e-4, |dcost| < 1e-5 (FP64 both sides)."""
import os

import numpy as np
import pytest

from car_racing_b200 import scenarios

pytestmark = pytest.mark.gpu
TOL_U, TOL_C = 1e-4, 1e-5
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _log(info):
    """Append the measured rates to gpurun_out/parity_rates.jsonl (GPU box runs): the gates below are set from them."""
    import json
    d = os.path.join(os.path.dirname(GOLD.rstrip("/")), "..", "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_rates.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=os.environ.get("PYTEST_CURRENT_TEST", "?"), **info)) + "\n")


def _compare(g, r, min_match=0.99, min_rel=None, max_other=None):
    """match = BASELINE's absolute tolerances (|du| < 1e-4, |dcost| < 1e-5) or the same failure on both sides.
    min_rel gates the same with the cost tolerance relative to the cost (1e-5 * max(1, |cost|): the costs are 1e2..1e6, and the
    two sides may stop one iteration apart at a different final mu), max_other the number of instances that end at a different
    KKT point of the non-convex problem (|dcost| > 1e-3 |cost|: rounding decides the basin there)."""
    both = (g["status"] == 0) & (r["status"] == 0)
    du = np.abs(g["u0"] - r["u0"]).max(axis=1)
    dc = np.abs(g["cost"] - r["cost"])
    scale = np.maximum(1.0, np.abs(r["cost"]))
    same_fail = (g["status"] != 0) & (g["status"] == r["status"])   # both stop the same way on the same instance
    match = (both & (du < TOL_U) & (dc < TOL_C)) | same_fail
    rel = (both & (du < TOL_U) & (dc < TOL_C * scale)) | same_fail
    other = both & (dc > 1e-3 * scale)
    info = dict(B=len(du), both_converged=int(both.sum()), match=int(match.sum()),
                gpu_fail=int((g["status"] != 0).sum()), cpu_fail=int((r["status"] != 0).sum()),
                worst_du=float(du[both].max()) if both.any() else 0.0, worst_dc=float(dc[both].max()) if both.any() else 0.0,
                iters_equal=float((g["iters"] == r["iters"]).mean()), match_rel=int(rel.sum()), different_kkt_point=int(other.sum()))
    info["match_frac"], info["gate"] = float(match.mean()), min_match
    print(info)
    _log(info)
    assert match.mean() >= min_match, info
    if min_rel is not None:
        assert rel.mean() >= min_rel, info
    if max_other is not None:
        assert other.sum() <= max_other, info
    return match, info


def test_mpccbf_config2_parity(crb, oracle):
    """North-star config: N=20, 3 static rivals, l_shape, seed 0."""
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(128, N=20, M=3, seed=0)
    prm = scenarios.default_cbf_params(N=20)
    g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, nthreads=os.cpu_count() or 1)
    match, _ = _compare(g, r)
    ok = match
    assert np.abs(g["x"][ok] - r["x"][ok]).max() < 1e-4
    assert np.abs(g["u"][ok] - r["u"][ok]).max() < 1e-4
    assert np.abs(g["sigma"][ok] - r["sigma"][ok]).max() < 1e-6
    assert g["kkt_err"][g["status"] == 0].max() <= 1e-6


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_mpccbf_full_batch_parity(crb, oracle, seed):
    """The BASELINE batch itself: 1024 config-2 instances per seed (round 1 had this only as a report under profiles/).
    Measured on a B200 (profiles/r05_parity_rates.jsonl): same status on >= 99.9 %, |du| < 1e-4 on every instance both sides
    converge on but one in 4096, and the ABSOLUTE cost criterion |dcost| < 1e-5 on 99.3-99.7 %: the rest differ by 1-2e-5 on
    costs of 1e2..1e5 (relative 1e-9) because the two sides stop one iteration apart at a different final mu and the objective
    carries the barrier residual of the 63 slack variables (1e4 * sigma, sigma ~ mu/z).  One instance in 4096 (seed 2, #243)
    ends at a different KKT point of the non-convex problem (same u0 to 4e-8, later-stage slacks differ): rounding decides the
    basin there -- the oracle differs from itself the same way when x0 is perturbed by 1e-15."""
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(1024, N=20, M=3, seed=seed)
    prm = scenarios.default_cbf_params(N=20)
    g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, nthreads=os.cpu_count() or 1)
    match, info = _compare(g, r, min_match=0.99)
    both = (g["status"] == 0) & (r["status"] == 0)
    du = np.abs(g["u0"] - r["u0"]).max(axis=1)
    dc = np.abs(g["cost"] - r["cost"])
    scale = np.maximum(1.0, np.abs(r["cost"]))
    rel = both & (du < TOL_U) & (dc < TOL_C * scale)
    other = both & (dc > 1e-3 * scale)
    extra = dict(seed=seed, same_status=float((g["status"] == r["status"]).mean()), match_rel_cost=float((rel | ~both).mean()),
                 different_kkt_point=int(other.sum()), du_max_same_point=float(du[both & ~other].max()),
                 dcost_max_same_point=float(dc[both & ~other].max()))
    print(extra)
    _log(extra)
    assert extra["same_status"] >= 0.998 and extra["match_rel_cost"] >= 0.998 and extra["different_kkt_point"] <= 2
    assert extra["du_max_same_point"] < 1e-3 and extra["dcost_max_same_point"] < 1e-3
    ok = match & both
    assert np.abs(g["x"][ok] - r["x"][ok]).max() < 1e-4
    assert np.abs(g["sigma"][ok] - r["sigma"][ok]).max() < 1e-6


def test_kkt_certificate_on_every_converged_instance(crb):
    """Solver-independent certificate (tests/kkt_check.py: feasibility, NNLS multiplier recovery on the active set,
    stationarity, complementarity) on ALL converged, non-elastic instances of one BASELINE batch."""
    from kkt_check import certificate
    B = 1024
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=20, M=3, seed=1)
    prm = scenarios.default_cbf_params(N=20)
    g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    n = n_pass = 0
    worst = dict(dyn=0.0, row_viol=0.0, stat=0.0, comp=0.0)
    for b in range(B):
        if g["status"][b] != 0 or g["elastic_max"][b] > 1e-7:
            continue
        c = certificate(x0[b], xt, obs[b], lap_off[b], prm, g["x"][b], g["u"][b], g["sigma"][b])
        n += 1
        ok = c["dyn"] < 1e-8 and c["row_viol"] < 1e-6 and c["stat"] < 1e-4 and c["comp"] < 1e-4
        n_pass += ok
        for k in worst:
            worst[k] = max(worst[k], float(c[k]))
    info = dict(certified=n, passed=int(n_pass), pass_rate=n_pass / max(n, 1), worst=worst,
                converged=int((g["status"] == 0).sum()), elastic=int(((g["status"] == 0) & (g["elastic_max"] > 1e-7)).sum()))
    print(info)
    _log(info)
    assert n >= 0.98 * B and worst["dyn"] < 1e-8 and worst["row_viol"] < 1e-6
    # the NNLS recovery is itself approximate at active-set changes (|stat| up to ~3e-4 on a handful of instances)
    assert n_pass >= 0.99 * n and worst["stat"] < 1e-2


@pytest.mark.parametrize("N,M", [(10, 0), (20, 0), (10, 1), (12, 2), (20, 4), (5, 3), (1, 1), (32, 3), (15, 6), (10, 8)])
def test_cbf_shapes_parity(crb, oracle, N, M):
    """mpc_lti (M=0) and other horizon / rival-count combinations, incl. N=1 and M=MMAX (8)."""
    B = 192
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=N, M=max(M, 1), seed=10 + N + M)
    obs, lap_off = obs[:, :M], lap_off[:, :M]
    prm = scenarios.default_cbf_params(N=N, width=0.8 if M == 0 else 1.0)
    g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, nthreads=os.cpu_count() or 1)
    # measured 189..192 of 192 on the absolute criterion (profiles/r05_parity_rates.jsonl, r06_parity_rates.jsonl): N = 32 is the
    # shape where 32 of the 192 scenarios fail on both sides and 2-3 of the others differ in the cost only (same u0 to 4e-7) --
    # barrier residual of the 99 slacks, and one instance whose later-stage slacks settle differently; which ones moves with the
    # order of the kernel's arithmetic (round 2, session 2 reordered the Riccati sweep), hence two instances of headroom
    _compare(g, r, min_match=0.97, max_other=2)


def test_full_weight_matrix(crb, oracle):
    """A Q with off-diagonal entries takes the kernel's general path (the reference's diagonal Q takes one product per row,
    KParams::q_diag): parity with the oracle on both."""
    B, N, M = 96, 20, 3
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=N, M=M, seed=41)
    prm = scenarios.default_cbf_params(N=N)
    Q = np.array(prm["Q"], float)
    Q[0, 3] = Q[3, 0] = 1.5
    Q[3, 5] = Q[5, 3] = -2.0
    Q[1, 1] = 0.5
    prm = dict(prm, Q=Q)
    g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, nthreads=os.cpu_count() or 1)
    _compare(g, r, min_match=0.96, max_other=1)


def test_per_rival_sizes_zero_start_and_x0_rows(crb, oracle):
    """Round-2 options through the C-ABI on the GPU: per-rival (L, W) in the record (flag RIVAL_SIZE), the zero start of
    Opti/IPOPT (B200MPC_START_ZERO, bounded iteration count: same iterates as the oracle), status 4 for an x_0 outside the
    stage-0 rows."""
    B, N, M = 96, 20, 3
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=N, M=M, seed=31)
    prm = scenarios.default_cbf_params(N=N)
    rng = np.random.default_rng(5)
    sizes = np.stack([rng.uniform(0.35, 0.5, (B, M)), rng.uniform(0.18, 0.26, (B, M))], axis=2)
    g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm, sizes=sizes)
    r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, sizes=sizes, nthreads=os.cpu_count() or 1)
    _compare(g, r, min_match=0.98)      # measured 96 of 96
    kw = dict(start=1, max_iter=40, max_reset=50)
    g = crb.solve_cbf_batch(x0[:32], xt, obs[:32], lap_off[:32], prm, **kw)
    r = oracle.solve_cbf_batch(x0[:32], xt, obs[:32], lap_off[:32], prm, nthreads=os.cpu_count() or 1, **kw)
    same = (g["status"] == r["status"]) & (g["iters"] == r["iters"])
    info = dict(zero_start_same_path=float(same.mean()), x_diff=float(np.abs(g["x"] - r["x"])[same].max()))
    print(info)
    _log(info)
    assert same.mean() >= 0.95 and info["x_diff"] < 1e-5      # measured 32 of 32, x_diff 0
    x0b = x0[:8].copy()
    x0b[0, 5], x0b[1, 0] = 1.2, -0.3
    g = crb.solve_cbf_batch(x0b, xt, obs[:8], lap_off[:8], prm)
    r = oracle.solve_cbf_batch(x0b, xt, obs[:8], lap_off[:8], prm)
    assert (g["status"][:2] == 4).all() and (g["status"] == r["status"]).all()


def test_mpc_lti_anchor_on_gpu(crb):
    prm = scenarios.default_cbf_params(N=10, width=0.8)
    g = crb.solve_cbf_batch(np.zeros((1, 6)), np.array([0.8, 0, 0, 0, 0, 0.0]), np.zeros((1, 0, 2, 11)), None, prm)
    assert g["status"][0] == 0
    assert abs(g["u0"][0, 0] - 0.0033384) < 1e-6 and abs(g["u0"][0, 1] - 1.0) < 1e-6
    assert abs(g["cost"][0] - 22.71576162) < 1e-6


def test_moving_rivals_lap_offsets_per_stage_targets(crb, oracle):
    """mpc_multi_agents-style data: per-stage targets, alpha 0.6, margin 0.15, moving rivals, ego on lap 1."""
    N, M, B = 10, 2, 64
    rng = np.random.default_rng(7)
    lap = scenarios.LAP_LENGTH["goggle"]
    x0, xt, obs, _ = scenarios.mpccbf_scenarios(B, N=N, M=M, seed=7, track="goggle")
    x0[:, 4] += lap                                       # ego has completed one lap, rivals have not
    lap_off = np.full((B, M), lap)
    v = rng.uniform(0.0, 0.6, size=(B, M, 1))
    obs[:, :, 0, :] += v * 0.1 * np.arange(N + 1)          # rivals move along s
    obs[:, :, 0, 1:] += lap                                # quirk control.py:542: h_next carries no lap offset
    xts = np.zeros((B, N + 1, 6))
    xts[:, :, 0] = x0[:, 0:1]
    xts[:, :, 5] = rng.uniform(-0.3, 0.3, size=(B, 1)) * np.linspace(0, 1, N + 1)
    prm = scenarios.default_cbf_params(N=N, alpha=0.6, margin=0.15, Q=np.diag([10.0, 0, 0, 5.0, 0, 50.0]))
    g = crb.solve_cbf_batch(x0, xts, obs, lap_off, prm)
    r = oracle.solve_cbf_batch(x0, xts, obs, lap_off, prm, nthreads=os.cpu_count() or 1)
    _compare(g, r, min_match=0.98)      # measured 64 of 64


def test_full_size_properties(crb):
    """BASELINE.json size (B=1024): size-independent properties of the returned solutions."""
    B, N, M = 1024, 20, 3
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=N, M=M, seed=1)
    prm = scenarios.default_cbf_params(N=N)
    g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    ok = g["status"] == 0
    assert ok.mean() >= 0.99
    assert g["kkt_err"][ok].max() <= 1e-6
    x, u, sg = g["x"], g["u"], g["sigma"]
    assert np.abs(x[:, 0] - x0).max() == 0.0                                        # control.py:497
    dyn = x[:, 1:] - x[:, :-1] @ prm["A"].T - u @ prm["B"].T                          # control.py:566-570
    assert np.abs(dyn[ok]).max() < 1e-8
    assert (np.abs(u[ok, :, 0]) <= 0.5 + 1e-9).all() and (np.abs(u[ok, :, 1]) <= 1.0 + 1e-9).all()
    assert (x[ok, 1:, 0] >= -1e-9).all() and (np.abs(x[ok, 1:, 5]) <= 1.0 + 1e-9).all() and (sg[ok] >= -1e-12).all()
    ds = x[:, None, :, 4] - obs[:, :, 0, :]
    de = x[:, None, :, 5] - obs[:, :, 1, :]
    h = (ds / 0.4) ** 6 + (de / 0.2) ** 6 - 1.2 - sg
    rows = h[:, :, 1:] - 0.2 * h[:, :, :-1]                                          # control.py:558
    feas = ok & (g["elastic_max"] < 1e-7)
    scale = np.maximum(1.0, np.abs(h[:, :, 1:]))
    assert (rows[feas] / scale[feas]).min() > -1e-6
    # objective value is consistent with the returned trajectory
    dx = x - xt
    cost = np.einsum("bki,ij,bkj->b", dx, prm["Q"], dx) + np.einsum("bki,ij,bkj->b", u, prm["R"], u) + 1e4 * sg.sum(axis=(1, 2))
    assert np.abs(cost - g["cost"])[ok].max() < 1e-8
    # determinism: a second run gives bit-identical results
    g2 = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm, want=())
    assert (g2["u0"] == g["u0"]).all() and (g2["iters"] == g["iters"]).all()


def test_kkt_certificate_on_gpu_solutions(crb):
    from kkt_check import certificate
    N, M = 20, 3
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(8, N=N, M=M, seed=3)
    prm = scenarios.default_cbf_params(N=N)
    g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    n = 0
    for b in range(8):
        if g["status"][b] != 0 or g["elastic_max"][b] > 1e-7:
            continue
        c = certificate(x0[b], xt, obs[b], lap_off[b], prm, g["x"][b], g["u"][b], g["sigma"][b])
        assert c["dyn"] < 1e-8 and c["row_viol"] < 1e-6 and c["stat"] < 1e-4 and c["comp"] < 1e-4, c
        n += 1
    assert n >= 6


def test_ilqr_matches_reference_golden(crb):
    """GPU iLQR against vectors produced by the unmodified reference control.ilqr."""
    gold = np.load(os.path.join(GOLD, "ilqr_golden.npz"))
    for N in sorted(set(gold["N"].tolist())):
        idx = np.where(gold["N"] == N)[0]
        x0, xt, obs, lap = gold["x0"][idx], gold["xt"][idx], gold["obs"][idx][:, :, :N + 1], gold["lap"][idx]
        lap_off = (np.trunc(x0[:, 4] / lap) - np.trunc(obs[:, 0, 0] / lap)) * lap
        prm = dict(A=gold["A"], B=gold["B"], Q=gold["Q"], R=gold["R"], N=int(N), max_iter=int(gold["max_iter"]), L=0.4, W=0.2)
        g = crb.solve_ilqr_batch(x0, xt, obs, lap_off, prm)
        assert np.abs(g["u0"] - gold["u0"][idx]).max() < 1e-9


def test_ilqr_batch_vs_oracle(crb, oracle):
    B, N = 512, 50
    x0, xt, obs, lap_off = scenarios.ilqr_scenarios(B, N=N, seed=2)
    p = scenarios.default_cbf_params()
    prm = dict(A=p["A"], B=p["B"], Q=p["Q"], R=p["R"], N=N, max_iter=150, L=0.4, W=0.2)
    g = crb.solve_ilqr_batch(x0, xt, obs, lap_off, prm)
    r = oracle.solve_ilqr_batch(x0, xt, obs, lap_off, prm, nthreads=os.cpu_count() or 1)
    same = g["iters"] == r["iters"]
    du = np.abs(g["u0"] - r["u0"]).max(axis=1)
    print("ilqr: same iteration count", same.mean(), "max du (same path)", du[same].max())
    assert same.mean() >= 0.99 and du[same].max() < 1e-8
    assert np.abs(g["u"][same] - r["u"][same]).max() < 1e-8 and np.abs(g["x"][same] - r["x"][same]).max() < 1e-8


def test_drop_in_shims_on_gpu(crb, oracle):
    """The reference-signature functions (control.py:476,198,64) with duck-typed vehicles."""
    import types
    from test_shims_host import Rival
    lap = scenarios.LAP_LENGTH["l_shape"]
    vehicles = {"ego": Rival(0, 0, 0), "car1": Rival(4.0, 0.2, 0.1), "car2": Rival(10.0, 0.2, -0.1)}
    param = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 4.0, 0, 40.0]),
                                  matrix_R=np.diag([0.1, 0.1]), num_horizon=10, alpha=0.8)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    track = types.SimpleNamespace(width=1.0, lap_length=lap)
    x = np.array([0.9, 0.0, 0.0, 0.02, 3.0, 0.05])
    xt = np.array([0.8, 0, 0, 0, 0, 0.0]).reshape(6, 1)
    u, det = crb.mpccbf(x, xt, param, vehicles, "ego", lap, 0.0, 0.1, False, track, sysp, return_details=True)
    prm = scenarios.default_cbf_params(N=10)
    obs = np.zeros((1, 1, 2, 11)); obs[0, 0, 0] = 4.0 + 0.02 * np.arange(11); obs[0, 0, 1] = 0.1
    r = oracle.solve_cbf_batch(x[None], xt.ravel(), obs, None, prm)
    assert u.shape == (2,) and np.abs(u - r["u0"][0]).max() < TOL_U and abs(det["cost"][0] - r["cost"][0]) < TOL_C
    u2 = crb.mpc_lti(x, xt, types.SimpleNamespace(**{**param.__dict__}), sysp, track)
    r2 = oracle.solve_cbf_batch(x[None], xt.ravel(), np.zeros((1, 0, 2, 11)), None, prm)
    assert np.abs(u2 - r2["u0"][0]).max() < TOL_U
    ip = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=param.matrix_Q, matrix_R=param.matrix_R,
                               max_iter=150, num_horizon=50)
    u3 = crb.ilqr(x, xt.ravel(), ip, {"ego": vehicles["ego"], "car1": vehicles["car1"]}, "ego", lap, 0.0, 0.1, track, sysp)
    tr, _ = vehicles["car1"].get_trajectory_nsteps(0.0, 0.1, 51)
    r3 = oracle.solve_ilqr_batch(x[None], xt.ravel(), tr[4:6][None], None,
                                 dict(A=scenarios.LTI_A, B=scenarios.LTI_B, Q=param.matrix_Q, R=param.matrix_R, N=50, max_iter=150, L=0.4, W=0.2))
    assert np.abs(u3 - r3["u0"][0]).max() < 1e-9


def test_argmin_kernel_first_min(crb):
    import ctypes as C
    import torch
    from car_racing_b200 import _capi
    h = _capi.Handle()
    rec = np.zeros(1000, dtype=_capi.RECORD_DTYPE)
    rng = np.random.default_rng(0)
    rec["cost"] = rng.integers(5, 50, size=1000).astype(float)
    rec["status"] = rng.integers(0, 3, size=1000)
    rec["cost"][[17, 400, 923]] = 1.0
    rec["status"][[17, 400, 923]] = [2, 0, 0]
    d = torch.from_numpy(rec.view(np.uint8)).cuda()
    out = torch.full((1,), -7, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    h.check(_capi.lib().b200mpc_argmin_cost_device(h.ptr, d.data_ptr(), 1000, 0, out.data_ptr()), "argmin")
    torch.cuda.current_stream().synchronize(); torch.cuda.synchronize()
    assert int(out.item()) == 400
    h.check(_capi.lib().b200mpc_argmin_cost_device(h.ptr, d.data_ptr(), 1000, 2, out.data_ptr()), "argmin")
    torch.cuda.synchronize()
    assert int(out.item()) == 17


def _planner_batch(seeds):
    from car_racing_b200 import planning
    from planner_cases import make_planner
    kws, offs, prm = [], [], None
    for seed in seeds:
        p = make_planner(seed, num_veh=2 + seed % 2)
        N = p.racing_game_param.num_horizon_planner
        ego = p.vehicles["ego"]
        prm = planning.planner_params(p.racing_game_param.matrix_A, p.racing_game_param.matrix_B, N)
        for c in range(len(p.sorted_vehicles) + 1):
            xlb, xub = planning.candidate_bounds(c, p.xcurv_ego, p.sorted_vehicles, p.obs_infos, 0.4, 0.2, 1.0, p.track.lap_length, N)
            if not planning.x0_feasible(ego.xcurv, xlb, xub) or (xlb[1:N, 1] > xub[1:N, 1] - 0.05).any():
                continue
            s_ref, ey_ref = planning.candidate_targets(c, ego.xcurv, p.bezier_xcurvs, p.bezier_funcs, N)
            kw, off = planning.pack_candidates(ego.xcurv, s_ref[None], ey_ref[None], xlb[None], xub[None], N)
            kws.append(kw)
            offs.append(off)
    cat = {k: (np.concatenate([kw[k] for kw in kws]) if kws[0][k] is not None else None) for k in kws[0]}
    return cat, np.concatenate(offs), prm


def test_planner_candidate_qp_parity(crb, oracle):
    """Planner candidate QPs (overtake_traj_planner.py:248-379): per-stage bounds + ey-rate cost, R = 0."""
    kw, off, prm = _planner_batch(range(24))
    assert kw["x0"].shape[0] >= 40
    g = crb.solve_cbf_batch(kw["x0"], kw["xt"], kw["obs"], None, prm, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"])
    r = oracle.solve_cbf_batch(kw["x0"], kw["xt"], kw["obs"], None, prm, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"],
                               nthreads=os.cpu_count() or 1)
    match, info = _compare(g, r, min_match=0.98)     # measured 57 of 57 (37 solved, 20 infeasible regions stop the same way)
    ok = (g["status"] == 0) & (r["status"] == 0)
    # the rest are dynamically infeasible regions (a rival row demands a 0.4 m lateral jump within one step): both
    # solvers stop the same way and the drop-in then takes the reference's heuristic trajectory (:365-374)
    assert ok.sum() >= 30 and (g["status"] == r["status"]).all()
    assert np.abs(g["x"][ok] - r["x"][ok]).max() < 1e-5


def test_planner_drop_in_on_gpu(crb, oracle):
    from car_racing_b200 import planning
    from planner_cases import make_planner
    for seed in (3, 7, 11):
        p1, p2 = make_planner(seed), make_planner(seed)
        t1, f1, st1, s1 = planning.solve_optimization_problem(p1)     # default solver: the GPU batch call
        t2, f2, st2, s2 = planning.solve_optimization_problem(
            p2, solver=lambda x0, xt, obs, lo, prm, **k: oracle.solve_cbf_batch(x0, xt, obs, lo, prm, **k))
        assert f1 == f2 and np.abs(s1 - s2).max() < 1e-5 and t1.shape == (11, 6)
        fin = np.isfinite(p2.candidate_costs)
        assert (np.isfinite(p1.candidate_costs) == fin).all()
        assert np.abs(p1.candidate_costs[fin] - p2.candidate_costs[fin]).max() < 1e-5


def test_max_sizes_and_odd_batches(crb, oracle):
    """Maximum horizon / rival count of the C-ABI (N=64, M=4 and M=8) and batch sizes that are not a multiple of anything."""
    for (N, M, B, gate) in ((64, 4, 64, 0.92), (64, 8, 24, 0.95)):      # measured 61-62 of 64 (|dcost| 5e-5 on the others: 3.4x the
        # horizon, 3.4x the barrier residual in the cost; relative to the cost all 64 match) and 24 of 24
        x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=N, M=M, seed=21)
        obs[:, :, 0, :] += 3.0                                  # keep the long horizon feasible: rivals further ahead
        prm = scenarios.default_cbf_params(N=N)
        g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
        r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, nthreads=os.cpu_count() or 1)
        _compare(g, r, min_match=gate, min_rel=0.98, max_other=0)
    for B in (1, 3, 33):
        x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=20, M=3, seed=30 + B)
        prm = scenarios.default_cbf_params(N=20)
        g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
        r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm)
        _compare(g, r, min_match=0.96)      # measured 1 of 1, 3 of 3, 33 of 33
        assert g["x"].shape == (B, 21, 6) and g["sigma"].shape == (B, 3, 21)


def test_blocked_lane_elastic_rows(crb, oracle):
    """A rival parked right in front of the ego across the whole lane: the CBF rows cannot all hold; both solvers
    must report the same (elastic) outcome instead of diverging -- the reference's IPOPT would fail here (control.py:600-603)."""
    N, M, B = 20, 3, 16
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=N, M=M, seed=5)
    x0[:, 0] = 1.4
    obs[:, :, 0, :] = x0[:, None, 4:5] + 0.55
    obs[:, 0, 1, :], obs[:, 1, 1, :], obs[:, 2, 1, :] = -0.55, 0.0, 0.55
    x0[:, 5] = 0.27
    prm = scenarios.default_cbf_params(N=N)
    g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, nthreads=os.cpu_count() or 1)
    same_status = g["status"] == r["status"]
    both = (g["status"] == 0) & (r["status"] == 0)
    print("status gpu", g["status"], "cpu", r["status"], "elastic gpu", g["elastic_max"].round(4))
    _log(dict(blocked_lane_same_status=float(same_status.mean()), both_converged=int(both.sum())))
    assert same_status.mean() >= 0.9      # measured 16 of 16; 16 instances on the edge of feasibility: a rounding-level difference flips max_iter / converged
    if both.any():
        assert np.abs(g["elastic_max"][both] - r["elastic_max"][both]).max() < 1e-4
        assert np.abs(g["cost"][both] - r["cost"][both]).max() < 1e-3 * np.abs(r["cost"][both]).max()


def test_capi_argument_validation(crb):
    """Bad arguments are reported through the return code + b200mpc_last_error, never by a crash."""
    import ctypes as C
    from car_racing_b200 import _capi, batch
    h = _capi.Handle()
    L = _capi.lib()
    prm = scenarios.default_cbf_params(N=20)
    rec_in = np.zeros((2, batch.cbf_record_doubles(20, 3, False)))
    rec_out = np.zeros(2, dtype=_capi.RECORD_DTYPE)
    o = _capi.default_options()

    def call(p, B=2, inp=rec_in, out=rec_out, opt=o):
        return L.b200mpc_cbf_solve(h.ptr, C.byref(p), C.byref(opt), B, None if inp is None else inp.ctypes.data_as(C.c_void_p),
                                   None if out is None else out.ctypes.data_as(C.c_void_p), None, None, None, None)
    p = _capi.make_cbf_params(prm, 3, False)
    p.N = 65
    assert call(p) == -1 and b"out of range" in L.b200mpc_last_error(h.ptr)
    p = _capi.make_cbf_params(prm, 3, False)
    p.M = _capi.MMAX + 1
    assert call(p) == -1
    p = _capi.make_cbf_params(prm, 3, False)
    assert call(p, B=0) == -1 and call(p, inp=None) == -1 and call(p, out=None) == -1
    p.alpha = 1.5
    assert call(p) == -1
    p = _capi.make_cbf_params(prm, 3, False, flags=1)
    assert call(p) == -1 and b"flags" in L.b200mpc_last_error(h.ptr)
    bad = _capi.default_options()
    bad.max_iter = 0
    assert call(_capi.make_cbf_params(prm, 3, False), opt=bad) == -1
    bad = _capi.default_options(start=2)
    assert call(_capi.make_cbf_params(prm, 3, False), opt=bad) == -1
    bad = _capi.default_options(max_reset=-1)
    assert call(_capi.make_cbf_params(prm, 3, False), opt=bad) == -1
    ip = _capi.make_ilqr_params(dict(A=prm["A"], B=prm["B"], Q=prm["Q"], R=prm["R"], N=70, max_iter=10, L=0.4, W=0.2))
    assert L.b200mpc_ilqr_solve(h.ptr, C.byref(ip), 1, rec_in.ctypes.data_as(C.c_void_p), rec_out.ctypes.data_as(C.c_void_p), None, None) == -1
    with pytest.raises(ValueError):
        batch.solve_cbf_packed(np.zeros((2, 7)), prm, 3, False)
    # a valid call still works afterwards
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(2, N=20, M=3, seed=1)
    assert (crb.solve_cbf_batch(x0, xt, obs, lap_off, prm, handle=h)["status"] == 0).all()


def test_lmpc_config4_parity(crb, oracle):
    """control.lmpc's QP: N=12, 44 safe-set points, LTV model."""
    B = 96
    sc = scenarios.lmpc_scenarios(B, seed=5)
    prm = scenarios.default_lmpc_params()
    g = crb.solve_lmpc_batch(*sc, prm)
    r = oracle.solve_lmpc_batch(*sc, prm, nthreads=os.cpu_count() or 1)
    match, info = _compare(g, r)
    assert info["both_converged"] >= 0.99 * B      # every generated instance is feasible (scenarios.lmpc_feasible): a convex QP
    ok = (g["status"] == 0) & (r["status"] == 0)
    assert np.abs(g["x"][ok] - r["x"][ok]).max() < 1e-4
    assert np.abs(g["u"][ok] - r["u"][ok]).max() < 1e-4
    assert np.abs(g["lambda"][ok] - r["lam"][ok]).max() < 1e-4
    # hull constraint and simplex hold on the GPU result itself
    x0, u_old, A, Bm, Cm, SS, Qf = sc
    assert np.abs(np.einsum("bak,bk->ba", SS, g["lambda"])[ok] - g["x"][ok, -1]).max() < 1e-6
    assert np.abs(g["lambda"][ok].sum(axis=1) - 1).max() < 1e-7 and g["lambda"][ok].min() > -1e-9
    xn = np.einsum("bnij,bnj->bni", A, g["x"][:, :-1]) + np.einsum("bnij,bnj->bni", Bm, g["u"]) + Cm
    assert np.abs(xn - g["x"][:, 1:])[ok].max() < 1e-7
    assert g["kkt_err"][ok].max() <= 1e-6


@pytest.mark.parametrize("N,K", [(2, 1), (3, 16), (5, 24), (8, 33), (16, 64), (12, 21)])
def test_lmpc_shapes_parity(crb, oracle, N, K):
    ni = 1 if K % 2 else 2
    sc = scenarios.lmpc_scenarios(24, N=N, num_ss_points=K, num_ss_iter=ni, seed=N + K)
    prm = scenarios.default_lmpc_params(N=N, Q=np.diag([0.5, 0, 0, 0.1, 0, 2.0]))
    g = crb.solve_lmpc_batch(*sc, prm)
    r = oracle.solve_lmpc_batch(*sc, prm, nthreads=os.cpu_count() or 1)
    _, info = _compare(g, r, min_match=0.95)     # 24 instances: measured 24 of 24 on all shapes
    if K > 1:   # K = 1 pins x_N to one stored state: infeasible, both sides must stop the same way
        assert info["both_converged"] >= 0.9 * 24


def test_lmpc_drop_in_on_gpu(crb, oracle):
    """control.lmpc shim: same 6-tuple as control.py:723-730."""
    from types import SimpleNamespace
    from car_racing_b200 import control
    rng = np.random.default_rng(0)
    sc = scenarios.lmpc_scenarios(1, seed=9)
    x0, u_old, A, Bm, Cm, SS, Qf = [a[0] for a in sc]
    # stored laps as the reference holds them: ss_xcurv (T, 6, laps), Qfun (T, laps); rows = the 22 points of each lap
    T = 60
    ss = np.zeros((T, 6, 2)); qf = np.zeros((T, 2))
    for lap in range(2):
        ss[:, :, lap] = 1e3                       # far from x0 except the stored block
        ss[10:32, :, lap] = SS[:, 22 * (1 - lap):22 * (2 - lap)].T
        qf[10:32, lap] = Qf[22 * (1 - lap):22 * (2 - lap)]
        for k in range(32, T):                    # the lap continues past the stored block
            ss[k, :, lap] = ss[31, :, lap]
            ss[k, 4, lap] += 0.12 * (k - 31)
            qf[k, lap] = qf[31, lap] - (k - 31)
    lp = SimpleNamespace(num_ss_iter=2, num_ss_points=44, shift=0, num_horizon=12, matrix_Q=np.zeros((6, 6)),
                         matrix_R=np.diag([1.0, 0.25]), matrix_dR=5 * np.diag([0.8, 0.0]))
    sp = SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10.0)
    u_pred, x_pred, ss_sel, q_sel, lin_points, lin_input = control.lmpc(x0, lp, list(A), list(Bm), list(Cm), ss, qf, 2, 20.0, 1.0,
                                                                       u_old, sp)
    assert u_pred.shape == (12, 2) and x_pred.shape == (13, 6) and ss_sel.shape[0] == 6 and lin_points.shape == (13, 6)
    assert lin_input.shape == (12, 2)
    K = ss_sel.shape[1]
    r = oracle.solve_lmpc_batch(x0[None], u_old[None], A[None], Bm[None], Cm[None], ss_sel[None], q_sel[None],
                                scenarios.default_lmpc_params())
    if r["status"][0] == 0:
        assert np.abs(u_pred - r["u"][0]).max() < 1e-4 and np.abs(x_pred - r["x"][0]).max() < 1e-4
    assert np.allclose(lin_points[:-1], x_pred[1:]) and np.allclose(lin_input[:-1], u_pred[1:])


def _sysid_gold():
    g = np.load(os.path.join(GOLD, "sysid_golden.npz"))
    return {k: g[k] for k in g.files}


def test_sysid_matches_reference_golden(crb):
    """LMPC model identification against outputs of the reference's own
    regression_and_linearization (tests/golden/make_sysid_golden.py)."""
    g = _sysid_gold()
    r = crb.estimate_abc_batch(g["lin_points"], g["lin_input"], g["ss"], g["us"], g["time_ss"], [0, 1], g["point_and_tangent"],
                               float(g["dt"]), int(g["max_num_point"]))
    assert (r["status"] == 0).all()
    assert (r["idx"] == g["idx"]).all()                     # same neighbours, same order
    # regression rows: normal equations with cond ~1e6 solved in a different summation order -> 1e-8; analytic rows: exact
    assert np.abs(r["A"] - g["A"]).max() < 1e-8 and np.abs(r["B"] - g["B"]).max() < 1e-8 and np.abs(r["C"] - g["C"]).max() < 1e-8
    assert np.abs(r["A"][:, :, 3:] - g["A"][:, :, 3:]).max() < 1e-13 and np.abs(r["C"][:, :, 3:] - g["C"][:, :, 3:]).max() < 1e-11


def test_sysid_batch_vs_restatement_and_few_points(crb):
    import sysid_numpy
    g = _sysid_gold()
    rng = np.random.default_rng(3)
    Bn, N = 40, 12
    lap = rng.integers(0, 2, Bn)
    t0 = np.array([rng.integers(1, g["time_ss"][l] - N - 3) for l in lap])
    lp = np.stack([g["ss"][t:t + N + 1, :, l] for t, l in zip(t0, lap)]) + rng.normal(scale=0.01, size=(Bn, N + 1, 6))
    lp[:, :, 4] = np.abs(lp[:, :, 4]) + 1e-3
    li = np.stack([g["us"][t:t + N, :, l] for t, l in zip(t0, lap)]) + rng.normal(scale=0.01, size=(Bn, N, 2))
    r = crb.estimate_abc_batch(lp, li, g["ss"], g["us"], g["time_ss"], [0, 1], g["point_and_tangent"], 0.1)
    A, B, C, idx = sysid_numpy.estimate_abc(lp, li, g["ss"], g["us"], g["time_ss"], [0, 1], g["point_and_tangent"], 0.1)
    assert (r["idx"] == idx).all() and (r["status"] == 0).all()
    assert np.abs(r["A"] - A).max() < 1e-8 and np.abs(r["B"] - B).max() < 1e-8 and np.abs(r["C"] - C).max() < 1e-8
    # fewer than max_num_point rows inside the bandwidth: "all rows, in row order" branch (lmpc_helper.py:234-235)
    ss2, us2 = g["ss"].copy(), g["us"].copy()
    ss2[30:, 0, :] += 80.0                                   # vx far away: scaled distance 8 > h for rows >= 30
    ts2 = g["time_ss"].copy()
    lp2, li2 = lp[:4].copy(), li[:4].copy()
    lp2[:, :, 0] = ss2[5, 0, 0]
    r2 = crb.estimate_abc_batch(lp2, li2, ss2, us2, ts2, [0, 1], g["point_and_tangent"], 0.1)
    A2, B2, C2, idx2 = sysid_numpy.estimate_abc(lp2, li2, ss2, us2, ts2, [0, 1], g["point_and_tangent"], 0.1)
    assert (r2["idx"] == idx2).all() and (idx2[:, :, :, 30:] == -1).all()
    assert np.abs(r2["A"] - A2).max() < 1e-7 and np.abs(r2["C"] - C2).max() < 1e-7


def test_sysid_chains_into_lmpc_records_and_drop_in(crb, oracle):
    """estimate_ABC writes the model block of LMPC records in place; LMPCRacingGame.estimate_ABC drop-in returns lists."""
    from types import SimpleNamespace
    from car_racing_b200 import batch, control
    g = _sysid_gold()
    N, K = 12, 44
    stride = batch.lmpc_record_doubles(N, K)
    rec = np.full((3, stride), 7.0)
    r = crb.estimate_abc_batch(g["lin_points"][:3], g["lin_input"][:3], g["ss"], g["us"], g["time_ss"], [0, 1],
                               g["point_and_tangent"], 0.1, out=rec, out_stride=stride, out_offset=8)
    assert (rec[:, :8] == 7.0).all() and (rec[:, 8 + 54 * N:] == 7.0).all()
    assert np.abs(rec[:, 8:8 + 36 * N].reshape(3, N, 6, 6) - g["A"][:3]).max() < 1e-8
    assert np.abs(rec[:, 8 + 48 * N:8 + 54 * N].reshape(3, N, 6) - g["C"][:3]).max() < 1e-8
    game = SimpleNamespace(lmpc_param=SimpleNamespace(num_horizon=N), iter=2, lin_points=g["lin_points"][5],
                           lin_input=g["lin_input"][5], ss_xcurv=g["ss"], u_ss=g["us"], time_ss=g["time_ss"],
                           point_and_tangent=g["point_and_tangent"], timestep=0.1)
    Atv, Btv, Ctv, used = control.estimate_ABC(game)
    assert len(Atv) == N and Atv[0].shape == (6, 6) and Btv[0].shape == (6, 2) and Ctv[0].shape == (6, 1) and len(used[0]) == 2
    assert np.abs(np.array(Atv) - g["A"][5]).max() < 1e-8 and np.abs(np.array(Ctv)[:, :, 0] - g["C"][5]).max() < 1e-8
    assert (used[3][1] == g["idx"][5, 3, 1]).all()


def test_sysid_lmpc_chain_on_device(crb, oracle):
    """Config 4 end to end on the device: sysid_kernel writes the model block of the LMPC records in HBM, lmpc_kernel
    consumes them (device-pointer entry points, same stream); checked against restatement + oracle on the host."""
    import ctypes as C
    import torch
    import sysid_numpy
    from car_racing_b200 import _capi, batch
    g = _sysid_gold()
    N, K, Bn = 12, 44, 16
    rng = np.random.default_rng(11)
    lap = rng.integers(0, 2, Bn)
    t0 = np.array([rng.integers(5, g["time_ss"][l] - 70) for l in lap])
    lp = np.stack([g["ss"][t + 1:t + N + 2, :, l] for t, l in zip(t0, lap)])
    li = np.stack([g["us"][t + 1:t + N + 1, :, l] for t, l in zip(t0, lap)])
    x0 = np.stack([g["ss"][t, :, l] for t, l in zip(t0, lap)]) + rng.normal(scale=1e-3, size=(Bn, 6))
    u_old = np.stack([g["us"][t, :, l] for t, l in zip(t0, lap)])
    # safe set: 22 points ahead on each of the two stored laps (nearest in s), Qfun = time to go
    SS = np.zeros((Bn, 6, K)); Qf = np.zeros((Bn, K))
    for b in range(Bn):
        for j in range(2):
            m = int(np.argmin(np.abs(g["ss"][:g["time_ss"][j], 4, j] - x0[b, 4]))) + 4
            SS[b, :, 22 * j:22 * j + 22] = g["ss"][m:m + 22, :, j].T
            Qf[b, 22 * j:22 * j + 22] = (g["time_ss"][j] - np.arange(m, m + 22)) + 10.0 * j
    prm = scenarios.default_lmpc_params(width=1.0)
    zeros = np.zeros
    rec, _ = batch.pack_lmpc(x0, u_old, zeros((Bn, N, 6, 6)), zeros((Bn, N, 6, 2)), zeros((Bn, N, 6)), SS, Qf, N)
    stride = rec.shape[1]
    laps, rows, lstride = batch.pack_laps(g["ss"], g["us"], g["time_ss"], [0, 1])
    pat = g["point_and_tangent"]
    dev = torch.device("cuda:0")
    h = batch.default_handle()
    d_rec = torch.from_numpy(rec).to(dev)
    d_lin = torch.from_numpy(np.ascontiguousarray(np.concatenate([lp[:, :N], li], axis=2))).to(dev)
    d_laps = torch.from_numpy(laps).to(dev)
    d_seg = torch.from_numpy(np.ascontiguousarray(pat[:, 3:6])).to(dev)
    d_out = torch.zeros((Bn, 4), dtype=torch.float64, device=dev)
    d_x = torch.zeros((Bn, N + 1, 6), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    sp = _capi.SysidParams()
    sp.N, sp.num_laps, sp.max_num_point, sp.num_segments, sp.lap_stride = N, 2, 40, pat.shape[0], lstride
    sp.lap_rows[0], sp.lap_rows[1] = rows
    sp.dt, sp.h, sp.lap_length = 0.1, 5.0, float(pat[-1, 3] + pat[-1, 4])
    L = _capi.lib()
    h.check(L.b200mpc_lmpc_sysid_device(h.ptr, C.byref(sp), Bn, d_lin.data_ptr(), d_laps.data_ptr(), d_seg.data_ptr(),
                                        d_rec.data_ptr(), stride, 8, None, None), "sysid_device")
    lpm = _capi.make_lmpc_params(prm, K)
    opt = _capi.default_options()
    h.check(L.b200mpc_lmpc_solve_device(h.ptr, C.byref(lpm), C.byref(opt), Bn, d_rec.data_ptr(), d_out.data_ptr(), None,
                                        d_x.data_ptr(), None, None), "lmpc_solve_device")
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().reshape(-1).view(_capi.RECORD_DTYPE)
    A, Bm, Cm, _ = sysid_numpy.estimate_abc(lp, li, g["ss"], g["us"], g["time_ss"], [0, 1], pat, 0.1)
    model = d_rec.cpu().numpy()[:, 8:8 + 54 * N]
    assert np.abs(model[:, :36 * N].reshape(Bn, N, 6, 6) - A).max() < 1e-8
    r = oracle.solve_lmpc_batch(x0, u_old, A, Bm, Cm, SS, Qf, prm, nthreads=os.cpu_count() or 1)
    both = (got["status"] == 0) & (r["status"] == 0)
    print("chain: both converged", both.sum(), "of", Bn, "statuses", got["status"], r["status"])
    assert both.sum() >= Bn // 2 and ((got["status"] == 0) == (r["status"] == 0)).mean() >= 0.9
    assert np.abs(got["u0"][both] - r["u0"][both]).max() < TOL_U and np.abs(got["cost"][both] - r["cost"][both]).max() < TOL_C
    assert np.abs(d_x.cpu().numpy()[both] - r["x"][both]).max() < 1e-4


def test_plant_step_matches_reference_golden(crb):
    """Batched plant step against the reference's DynamicBicycleModel.forward_dynamics goldens."""
    g = np.load(os.path.join(GOLD, "plant_golden.npz"))
    for name in ("ellipse", "l_shape"):
        xc, xg, us, dr = g["xcurv_" + name], g["xglob_" + name], g["u_" + name], g["draws_" + name]
        worst = 0.0
        for k in range(us.shape[1]):
            nc, ng = crb.plant_step_batch(xc[:, k], xg[:, k], us[:, k], dr[:, k], g["pat_" + name], dyn=tuple(g["dyn"]))
            worst = max(worst, np.abs(nc - xc[:, k + 1]).max(), np.abs(ng - xg[:, k + 1]).max())
        print(name, "max |d state| over", us.shape[1], "steps:", worst)
        assert worst < 1e-10          # 100 sub-steps of libm vs CUDA sin/cos/atan (1-2 ulp each)
    # zero_noise_flag and the lap wrap
    name = "ellipse"
    xc0 = g["xcurv_" + name][:, 0].copy()
    xc0[:, 4] = float(g["lap_length_" + name]) - 0.01
    nc, ng, laps = crb.plant_step_batch(xc0, g["xglob_" + name][:, 0], g["u_" + name][:, 0], None, g["pat_" + name], wrap_lap=True)
    assert (laps == 1).all() and (nc[:, 4] < 1.0).all() and (nc[:, 4] > 0.0).all()


def test_closed_loop_chain_matches_cpu_chain(crb, oracle):
    """Solve -> plant -> solve on the device (tools/closed_loop.py) against oracle + numpy plant on the host, 8 control steps."""
    import sys
    import plant_numpy
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "tools"))
    import closed_loop
    B, T, N = 12, 8, 20
    res, traj = closed_loop.run(B, T, seed=4, noise=False, record_every=1)
    g = np.load(os.path.join(GOLD, "plant_golden.npz"))
    x0, xt, s0, ey, v = scenarios.closed_loop_scenarios(B, N=N, M=3, seed=4)
    prm = scenarios.default_cbf_params(N=N)
    xc = x0.copy(); xg = np.zeros((B, 6)); xg[:, 0:3] = x0[:, 0:3]
    nfail = 0
    for k in range(T):
        r = oracle.solve_cbf_batch(xc, xt, scenarios.rival_block(s0, ey, v, 0.1 * k, N), np.zeros((B, 3)), prm,
                                   nthreads=os.cpu_count() or 1)
        nfail += int((r["status"] != 0).sum())
        xc, xg = plant_numpy.plant_step(xc, xg, r["u0"], np.zeros((B, 3)), g["dyn"], g["pat_l_shape"], float(g["lap_length_l_shape"]))
        d = np.abs(traj[k] - xc).max()
        print("closed loop step", k, "max |dx|", d)
        assert d < 1e-6
    assert res["status_counts"].get(0, 0) == B * T - nfail


def test_batches_in_flight_equal_the_blocking_call(crb):
    """CbfPipeline (b200mpc_cbf_solve_async on several handles, pinned buffers) must return, batch by batch, exactly
    what the blocking b200mpc_cbf_solve returns -- including when a slot is reused before the others are collected."""
    from car_racing_b200 import batch
    prm = scenarios.default_cbf_params(N=20)
    B, depth, nb = 96, 3, 7
    recs, refs = [], []
    for k in range(nb):
        x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(B, N=20, M=3, seed=100 + k)
        rec, M, ps = crb.pack_cbf(x0, xt, obs, lap_off, 20)
        recs.append(rec)
        refs.append(crb.solve_cbf_packed(rec, prm, M, ps, want=())["record"])
    pipe = batch.CbfPipeline(prm, M=3, B=B, depth=depth)
    out, tickets = {}, []
    for k in range(nb):
        if k >= depth:
            t = tickets[k - depth]
            out[t] = pipe.result(t)
        tickets.append(pipe.submit(recs[k]))
    with pytest.raises(RuntimeError):
        pipe.submit(recs[0])                     # the slot's previous batch has not been collected
    for t in tickets[-depth:]:
        out[t] = pipe.result(t)
    for k in range(nb):
        assert out[k].tobytes() == refs[k].tobytes()
    assert pipe.launch_count == nb
    pipe.close()


def test_plan_and_track_chain_matches_host_path(crb):
    """candidate solve -> selection (overtake_traj_planner.py:205-246) -> per-stage targets
    (control.py:373-382) -> tracking MPC on one stream (b200mpc_plan_and_track) must equal the two reference-shaped calls
    with the host in between (planning.solve_optimization_problem, then control.mpc_multi_agents)."""
    import types
    from car_racing_b200 import control, planning
    from planner_cases import make_planner
    from test_shims_host import Rival
    param = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 5.0, 0, 50.0]),
                                  matrix_R=np.diag([0.1, 0.1]), num_horizon_ctrl=10)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    flags = []
    for seed, old in ((3, None), (7, None), (11, 0), (12, 2), (13, 1), (14, None)):
        ps = [make_planner(seed, num_veh=2 + seed % 2) for _ in range(2)]
        for p in ps:
            p.old_direction_flag = old
            for name in p.sorted_vehicles:
                tr = p.obs_infos[name]
                p.vehicles[name] = Rival(tr[4, 0], tr[0, 0], tr[5, 0])
        p1, p2 = ps
        x = np.asarray(p1.vehicles["ego"].xcurv, float).copy()
        (t1, f1, st1, s1), (u1, x1) = planning.plan_and_track(p1, x, param, p1.track, sysp, time=None)
        t2, f2, st2, s2 = planning.solve_optimization_problem(p2)
        u2, x2 = control.mpc_multi_agents(x, param, p2.track, None, None, None, sysp, target_traj_xcurv=t2, vehicles=p2.vehicles,
                                          agent_name="ego", direction_flag=f2, sorted_vehicles=p2.sorted_vehicles, time=None)
        c2 = planning.selection_costs(s2, p2.sorted_vehicles, p2.obs_infos, 0.4, 0.2, p2.track.lap_length, old)
        assert f1 == f2 and np.abs(p1.selection_costs - np.array(c2)).max() < 1e-9
        assert np.abs(t1 - t2).max() < 1e-12 and np.abs(s1 - s2).max() < 1e-12
        assert p1.tracking_status == 0
        assert np.abs(u1 - u2).max() < 1e-7 and np.abs(x1 - x2).max() < 1e-7
        flags.append(f1)
    assert len(set(flags)) > 1          # the cases do not all pick the same region


def test_plan_and_track_with_64_candidates(crb):
    """BASELINE config 3: 64 candidates per planner call -- the reference's num_veh+1 regions plus perturbed references
    (scaled lateral targets x stretched progress), each mapped to its region for the neighbour test of the selection
    cost.  The device chain must pick the candidate a host re-computation picks, and track its trajectory."""
    import types
    from car_racing_b200 import control, planning
    from planner_cases import make_planner
    from test_shims_host import Rival
    param = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 5.0, 0, 50.0]),
                                  matrix_R=np.diag([0.1, 0.1]), num_horizon_ctrl=10)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    for seed in (5, 9):
        p = make_planner(seed, num_veh=2)
        for name in p.sorted_vehicles:
            tr = p.obs_infos[name]
            p.vehicles[name] = Rival(tr[4, 0], tr[0, 0], tr[5, 0])
        N, ego = 10, p.vehicles["ego"]
        x = np.asarray(ego.xcurv, float).copy()
        ex = dict(s_ref=[], ey_ref=[], xlb=[], xub=[], region=[], heur=[])
        for c in range(3):
            xlb, xub = planning.candidate_bounds(c, p.xcurv_ego, p.sorted_vehicles, p.obs_infos, 0.4, 0.2, 1.0, p.track.lap_length, N)
            s0, e0 = planning.candidate_targets(c, x, p.bezier_xcurvs, p.bezier_funcs, N)
            h0 = planning.heuristic_traj(c, p.xcurv_ego, p.bezier_xcurvs, p.bezier_funcs, N).T
            for scale in (0.4, 0.6, 0.8, 0.9, 1.1, 1.2, 1.4):
                for stretch in (0.9, 1.0, 1.1):
                    if len(ex["region"]) == 61:
                        break
                    ex["s_ref"].append(x[4] + stretch * (s0 - x[4]))
                    ex["ey_ref"].append(x[5] + scale * (e0 - x[5]))
                    ex["xlb"].append(xlb); ex["xub"].append(xub); ex["region"].append(c); ex["heur"].append(h0)
        ex = {k: np.array(v) for k, v in ex.items()}
        (traj, flag, st, sol), (u, xp) = planning.plan_and_track(p, x, param, p.track, sysp, time=None, extra=ex)
        C = sol.shape[0]
        assert C == 64 and len(p.selection_costs) == 64
        # host re-computation of the selection cost with the region map (overtake_traj_planner.py:205-243)
        region = np.concatenate([np.arange(3), ex["region"]])
        cost = np.zeros(C)
        for c in range(C):
            cost[c] = -10 * (sol[c, 4, -1] - sol[c, 4, 0])
            for side in (region[c] - 1, region[c]):
                if 0 <= side < 2:
                    o = p.obs_infos[p.sorted_vehicles[side]]
                    d2 = (sol[c, 4] - o[4, :N + 1]) ** 2 + (sol[c, 5] - o[5, :N + 1]) ** 2
                    cost[c] += 100 * (d2 - 0.4 ** 2 - 0.2 ** 2 < 0).sum()
        assert np.abs(cost - p.selection_costs).max() < 1e-9 and flag == int(np.argmin(cost))
        assert np.abs(traj - sol[flag].T).max() == 0.0
        u2, x2 = control.mpc_multi_agents(x, param, p.track, None, None, None, sysp, target_traj_xcurv=traj, vehicles=p.vehicles,
                                          agent_name="ego", direction_flag=flag, sorted_vehicles=p.sorted_vehicles, time=None)
        assert p.tracking_status == 0 and np.abs(u - u2).max() < 1e-7 and np.abs(xp - x2).max() < 1e-7


def test_ilqr_and_lmpc_pipelines_equal_the_blocking_calls(crb):
    from car_racing_b200 import batch
    p = scenarios.default_cbf_params()
    iprm = dict(A=p["A"], B=p["B"], Q=p["Q"], R=p["R"], N=50, max_iter=150, L=0.4, W=0.2)
    B = 64
    pipe = batch.IlqrPipeline(iprm, B=B, depth=2)
    refs, tickets = [], []
    for k in range(3):
        x0, xt, obs, lo = scenarios.ilqr_scenarios(B, N=50, seed=40 + k)
        refs.append(crb.solve_ilqr_batch(x0, xt, obs, lo, iprm, want=())["record"])
        if k >= 2:
            assert pipe.result(tickets[k - 2]).tobytes() == refs[k - 2].tobytes()
        tickets.append(pipe.submit(batch.pack_ilqr(x0, xt, obs, lo, 50)))
    for k in (1, 2):
        assert pipe.result(tickets[k]).tobytes() == refs[k].tobytes()
    pipe.close()
    lprm = scenarios.default_lmpc_params()
    sc = scenarios.lmpc_scenarios(32, seed=8)
    ref = crb.solve_lmpc_batch(*sc, lprm, want=())["record"]
    rec, K = batch.pack_lmpc(*sc, int(lprm["N"]))
    lp = batch.LmpcPipeline(lprm, K, B=32, depth=2)
    t0, t1 = lp.submit(rec), lp.submit(rec)
    assert lp.result(t0).tobytes() == ref.tobytes() and lp.result(t1).tobytes() == ref.tobytes()
    lp.close()

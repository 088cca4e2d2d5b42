"""CPU tests (-m "not gpu"): the PROBLEM STATEMENT our shims pack equals the reference's own.

tests/golden/nlp_golden.npz holds, for every optimisation problem on the hot path, the values of the reference's cost
and of every constraint its code passes to `opti.subject_to`, at random points -- produced by running the UNMODIFIED
reference functions under a recording stand-in for CasADi (tests/golden/make_nlp_golden.py, casadi_recorder.py).
Here the same inputs go through car_racing_b200's drop-in shims with the GPU call intercepted; the captured problem data
(exactly what is packed into the kernel's records) is evaluated at the same points, in the reference's row order.
This pins packing, quirks (lap offset on h but not h_next, alpha / margin of the two CBF variants, the slack penalty
of the last column, the planner's same-sign rival rows) and the oracle's problem functions to the reference's code.
The solver algorithm itself (IPOPT) stays unpinned -- DESIGN.md section 2.
"""
import os
import sys
import types

import numpy as np
import pytest

from car_racing_b200 import control, planning, scenarios

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from test_shims_host import Rival                  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "nlp_golden.npz"))
SYSP = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
RTOL = 1e-11


def close(a, b, what):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b) / (1.0 + np.abs(b))
    assert err.max() < RTOL, (what, float(err.max()), int(err.argmax()))


def capture_cbf(monkeypatch):
    seen = {}

    def fake(x0, xt, obs, lap_off, prm, want=("aux", "x", "u", "sigma"), handle=None, **opt):
        seen.update(x0=np.array(x0, float).reshape(6), xt=np.array(xt, float), obs=np.array(obs, float)[0],
                    lap_off=None if lap_off is None else np.array(lap_off, float).reshape(-1), prm=dict(prm), extra=opt)
        N = prm["N"]
        return dict(u=np.zeros((1, N, 2)), x=np.zeros((1, N + 1, 6)), status=np.array([0]), u0=np.zeros((1, 2)), cost=np.zeros(1),
                    iters=np.array([0]), elastic_max=np.zeros(1), kkt_err=np.zeros(1))
    monkeypatch.setattr(control.batch, "solve_cbf_batch", fake)
    return seen


def eval_cbf(seen, X, U, S):
    """cost, equality residuals and inequality values of the captured MPC-LTI / MPC-CBF problem at (X (6,N+1), U (2,N),
    S (M,N+1)), in the reference's recording order (control.py:497, 525-562, 564-591)."""
    p = seen["prm"]
    N, A, B, Q, R = p["N"], p["A"], p["B"], p["Q"], p["R"]
    obs, M = seen["obs"], seen["obs"].shape[0]
    lap_off = np.zeros(M) if seen["lap_off"] is None else seen["lap_off"]
    xt = seen["xt"].reshape(-1, 6) if seen["xt"].ndim > 1 else seen["xt"].reshape(1, 6)
    xts = xt if xt.shape[0] == N + 1 else np.repeat(xt, N + 1, axis=0)
    cost = p["slack_w"] * S.sum()
    for i in range(N + 1):
        d = X[:, i] - xts[i]
        cost += d @ Q @ d
    for i in range(N):
        cost += U[:, i] @ R @ U[:, i]
    eq = [X[:, 0] - seen["x0"]] + [X[:, i + 1] - (A @ X[:, i] + B @ U[:, i]) for i in range(N)]
    L6, W6 = p["L"] ** 6, p["W"] ** 6

    def h(j, i, off):
        return (X[4, i] - obs[j, 0, i] - off) ** 6 / L6 + (X[5, i] - obs[j, 1, i]) ** 6 / W6 - 1 - p["margin"] - S[j, i]
    ine = []
    for j in range(M):
        for i in range(N):
            hh, hn = h(j, i, lap_off[j]), h(j, i + 1, 0.0)
            ine += [hn - hh + p["alpha"] * hh, S[j, i]]
        ine.append(S[j, N])
    for i in range(N):
        ine += [U[0, i] + p["umax"][0], p["umax"][0] - U[0, i], U[1, i] + p["umax"][1], p["umax"][1] - U[1, i]]
    for i in range(N + 1):
        ine += [p["vmax"] - X[0, i], X[0, i] - p["vmin"], p["width"] - X[5, i], X[5, i] + p["width"]]
    return cost, np.concatenate(eq), np.array(ine)


def check_cbf_case(key, seen, prefix=""):
    M = seen["obs"].shape[0]
    for r in range(3):
        X, U = G[f"{key}/{prefix}var0"][r], G[f"{key}/{prefix}var1"][r]
        S = G[f"{key}/{prefix}var2"][r] if f"{key}/{prefix}var2" in G else np.zeros((0, X.shape[1]))
        assert S.shape[0] == M
        cost, eq, ine = eval_cbf(seen, X, U, S)
        close(cost, G[f"{key}/{prefix}cost"][r], key + " cost")
        close(eq, G[f"{key}/{prefix}eq"][r], key + " equalities")
        close(ine, G[f"{key}/{prefix}ineq"][r], key + " inequalities")


def test_mpc_lti_statement(monkeypatch):
    for k in range(2):
        key = "mpc_lti%d" % k
        seen = capture_cbf(monkeypatch)
        prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 4.0, 0, 40.0]),
                                    matrix_R=np.diag([0.1, 0.1]), num_horizon=int(G[key + "/N"]))
        track = types.SimpleNamespace(width=float(G[key + "/width"]), lap_length=scenarios.LAP_LENGTH["l_shape"])
        control.mpc_lti(G[key + "/x0"], G[key + "/xtarget"].reshape(6, 1), prm, SYSP, track)
        check_cbf_case(key, seen)


def test_mpccbf_statement_and_oracle_functions(monkeypatch):
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    import ipm_numpy
    lap = scenarios.LAP_LENGTH["l_shape"]
    for k in range(3):
        key = "mpccbf%d" % k
        seen = capture_cbf(monkeypatch)
        vehicles = {"ego": Rival(0, 0, 0)}
        for j, (s0, v, ey) in enumerate(G[key + "/rivals"]):
            vehicles["car%d" % (j + 1)] = Rival(s0, v, ey)
        prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 4.0, 0, 40.0]),
                                    matrix_R=np.diag([0.1, 0.1]), num_horizon=20, alpha=float(G[key + "/alpha"]))
        track = types.SimpleNamespace(width=1.0, lap_length=lap)
        control.mpccbf(G[key + "/x0"], G[key + "/xtarget"].reshape(6, 1), prm, vehicles, "ego", lap, float(G[key + "/time"]), 0.1,
                       False, track, SYSP)
        assert seen["obs"].shape[0] == int(G[key + "/num_slack_rows"])      # the same rivals pass the proximity filter
        check_cbf_case(key, seen)
        check_cbf_case(key, seen, prefix="near_")
        # the oracle's own problem functions (oracle/ipm_numpy.py, cross-checked against oracle/ocp_oracle.c) on the same points
        p = seen["prm"]
        for r in range(3):
            X, U, S = G[key + "/near_var0"][r], G[key + "/near_var1"][r], G[key + "/near_var2"][r]
            P = ipm_numpy.CbfProblem(X[:, 0], seen["xt"].reshape(6), seen["obs"], p["A"], p["B"], p["Q"], p["R"], 20, alpha=p["alpha"],
                                     margin=p["margin"], umax=tuple(p["umax"]), vmin=p["vmin"], vmax=p["vmax"], width=p["width"],
                                     L=p["L"], W=p["W"], lap_off=seen["lap_off"], slack_w=p["slack_w"])
            w = np.concatenate([X[:, 1:].T.ravel(), U.T.ravel(), S.ravel()])
            # its variable bounds are the reference's single-variable rows (:559-561, 572-576, 582-586)
            assert (P.lbw[P.ix(3)] == [p["vmin"], -np.inf, -np.inf, -np.inf, -np.inf, -p["width"]]).all()
            assert (P.ubw[P.ix(3)] == [p["vmax"], np.inf, np.inf, np.inf, np.inf, p["width"]]).all()
            assert (P.lbw[P.iu(2)] == [-0.5, -1.0]).all() and (P.ubw[P.iu(2)] == [0.5, 1.0]).all()
            assert (P.lbw[P.nx + P.nu:] == 0.0).all() and np.isinf(P.ubw[P.nx + P.nu:]).all()
            close(P.f(w), G[key + "/near_cost"][r], "ipm_numpy f")
            close(P.c(w), G[key + "/near_eq"][r][6:], "ipm_numpy c")
            M = S.shape[0]
            rows = np.concatenate([G[key + "/near_ineq"][r][j * (2 * 20 + 1):(j + 1) * (2 * 20 + 1) - 1:2] for j in range(M)])
            close(P.g(w), rows, "ipm_numpy g")


def test_mpc_multi_agents_statement(monkeypatch):
    lapg = scenarios.LAP_LENGTH["goggle"]
    for k in range(2):
        key = "multi%d" % k
        seen = capture_cbf(monkeypatch)
        vehicles = {"ego": Rival(0, 0, 0)}
        for j, (s0, v, ey) in enumerate(G[key + "/rivals"]):
            vehicles["car%d" % (j + 1)] = Rival(s0, v, ey)
        prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 5.0, 0, 50.0]),
                                    matrix_R=np.diag([0.1, 0.1]), num_horizon_ctrl=10)
        track = types.SimpleNamespace(width=1.0, lap_length=lapg)
        control.mpc_multi_agents(G[key + "/x0"], prm, track, None, None, None, SYSP, target_traj_xcurv=G[key + "/traj"], vehicles=vehicles,
                                 agent_name="ego", direction_flag=0, sorted_vehicles=["car1", "car2", "car3"], time=None)
        assert seen["obs"].shape[0] == int(G[key + "/num_slack_rows"])
        assert seen["prm"]["alpha"] == 0.6 and seen["prm"]["margin"] == 0.15
        check_cbf_case(key, seen)


def test_lmpc_statement(monkeypatch):
    for k in range(2):
        key = "lmpc%d" % k
        seen = {}

        def fake(x0, u_old, A, B, Cm, SS, Qfun, prm, want=(), handle=None, **opt):
            seen.update(x0=np.array(x0)[0], u_old=np.array(u_old)[0], A=np.array(A)[0], B=np.array(B)[0], C=np.array(Cm)[0],
                        SS=np.array(SS)[0], Qfun=np.array(Qfun)[0], prm=dict(prm))
            N = prm["N"]
            return dict(status=np.array([0]), x=np.zeros((1, N + 1, 6)), u=np.zeros((1, N, 2)))
        monkeypatch.setattr(control.batch, "solve_lmpc_batch", fake)
        lp = types.SimpleNamespace(num_horizon=12, num_ss_iter=2, num_ss_points=44, shift=0, matrix_Q=G[key + "/matrix_Q"],
                                   matrix_R=np.diag([1.0, 0.25]), matrix_dR=np.diag([4.0, 0.0]))
        N = 12
        out = control.lmpc(G[key + "/x0"], lp, list(G[key + "/Atv"]), list(G[key + "/Btv"]), list(G[key + "/Ctv"]), G[key + "/ss"],
                           G[key + "/Qfun"], int(G[key + "/it"]), 25.0, float(G[key + "/lap_width"]), G[key + "/u_old"], SYSP)
        close(out[2], G[key + "/ss_sel"], "select_points")          # lmpc_helper.select_points, run by the reference itself
        close(out[3], G[key + "/Qfun_sel"], "select_points Qfun")
        p = seen["prm"]
        K = seen["SS"].shape[1]
        xtrk = np.asarray(p["xtrk"], float)
        for r in range(3):
            X, U, lam = G[key + "/var0"][r], G[key + "/var1"][r], G[key + "/var2"][r].ravel()
            cost = seen["Qfun"] @ lam
            for i in range(N + 1):
                d = X[:, i] - xtrk
                cost += d @ p["Q"] @ d
            for i in range(N):
                du = U[:, i] - (seen["u_old"] if i == 0 else U[:, i - 1])
                cost += U[:, i] @ p["R"] @ U[:, i] + du @ p["dR"] @ du
            eq = [X[:, 0] - seen["x0"]]
            ine = []
            for i in range(N):
                eq.append(X[:, i + 1] - (seen["A"][i] @ X[:, i] + seen["B"][i] @ U[:, i] + seen["C"][i]))
                ine += [p["vmax"] - X[0, i], p["width"] - X[5, i], X[5, i] + p["width"],
                        U[0, i] + p["umax"][0], p["umax"][0] - U[0, i], U[1, i] + p["umax"][1], p["umax"][1] - U[1, i]]
            ine += list(lam)
            eq += [X[:, N] - seen["SS"] @ lam, [lam.sum() - 1.0], np.zeros(6)]          # slack == 0 rows (:693-694)
            close(cost, G[key + "/cost"][r], key + " cost")
            close(np.concatenate([np.ravel(e) for e in eq]), G[key + "/eq"][r], key + " equalities")
            close(np.array(ine), G[key + "/ineq"][r], key + " inequalities")
        assert K == 44


def _planner(seed, nv, old, obs):
    from planner_cases import make_planner
    p = make_planner(seed, num_veh=nv)
    p.old_direction_flag = None if old < 0 else old
    for j, name in enumerate(p.sorted_vehicles):
        p.obs_infos[name] = obs[j].copy()
    return p


def test_planner_candidate_statement_fallback_and_selection():
    N = 10
    for n in range(int(G["num_plans"])):
        key = "plan%d" % n
        seed, nv, old = int(G[key + "/seed"]), int(G[key + "/num_veh"]), int(G[key + "/old"])
        p = _planner(seed, nv, old, G[key + "/obs"])
        ego_x = np.asarray(p.vehicles["ego"].xcurv, float)
        close(ego_x, G[key + "/ego_x"], "planner fixture")
        half = p.track.width - 0.5 * 0.2
        for c in range(nv + 1):
            ck = f"{key}/cand{c}"
            xlb, xub = planning.candidate_bounds(c, p.xcurv_ego, p.sorted_vehicles, p.obs_infos, 0.4, 0.2, p.track.width,
                                                 p.track.lap_length, N)
            # the box the reference's rows leave for every variable (:276-324), read off its recorded constraints
            ref_lb, ref_ub = G[ck + "/x_lb"], G[ck + "/x_ub"]
            close(np.where(np.isfinite(xlb[:, 1]), xlb[:, 1], -1e300), np.where(np.isfinite(ref_lb[5]), ref_lb[5], -1e300), "ey lower")
            close(np.where(np.isfinite(xub[:, 1]), xub[:, 1], 1e300), np.where(np.isfinite(ref_ub[5]), ref_ub[5], 1e300), "ey upper")
            close(np.where(np.isfinite(xub[:, 0]), xub[:, 0], 1e300), np.where(np.isfinite(ref_ub[0]), ref_ub[0], 1e300), "vx upper")
            assert np.isinf(ref_lb[[0, 1, 2, 3, 4]]).all() and np.isinf(ref_ub[[1, 2, 3, 4]]).all()
            close(G[ck + "/u_lb"], np.repeat([[-planning.DELTA_MAX_PLAN], [-planning.A_MAX_PLAN]], N, axis=1), "u lower")
            close(G[ck + "/u_ub"], np.repeat([[planning.DELTA_MAX_PLAN], [planning.A_MAX_PLAN]], N, axis=1), "u upper")
            # cost and dynamics of the mapped problem (planning.pack_candidates) at the reference's points
            s_ref, ey_ref = planning.candidate_targets(c, ego_x, p.bezier_xcurvs, p.bezier_funcs, N)
            kw, off = planning.pack_candidates(ego_x, s_ref[None], ey_ref[None], xlb[None], xub[None], N)
            prm = planning.planner_params(p.racing_game_param.matrix_A, p.racing_game_param.matrix_B, N)
            for r in range(3):
                X, U = G[ck + "/var0"][r], G[ck + "/var1"][r]
                cost = off[0]
                for i in range(N + 1):
                    d = X[:, i] - kw["xt"][0, i]
                    cost += d @ prm["Q"] @ d
                for i in range(N):
                    cost += kw["wd"][0, i] * (X[5, i + 1] - X[5, i]) ** 2
                cost += planning.W_PROGRESS * (X[4, 0] - ego_x[4])     # the reference keeps s_0 a variable in -200 (s_N - s_0)
                close(cost, G[ck + "/cost"][r], "planner cost")
                eq = [X[:, 0] - ego_x] + [X[:, i + 1] - (prm["A"] @ X[:, i] + prm["B"] @ U[:, i]) for i in range(N)]
                close(np.concatenate(eq), G[ck + "/eq"][r], "planner dynamics")
            # the reference's own failure branch (:365-374)
            close(planning.heuristic_traj(c, p.xcurv_ego, p.bezier_xcurvs, p.bezier_funcs, N), G[key + "/fallback"][c], "heuristic trajectory")
        # the reference's own selection code (:205-246) on given candidate trajectories
        p2 = _planner(seed, nv, old, G[key + "/obs"])
        cs = planning.selection_costs(G[key + "/given"], p2.sorted_vehicles, p2.obs_infos, 0.4, 0.2, p2.track.lap_length, p2.old_direction_flag)
        flag = cs.index(min(cs))
        assert flag == int(G[key + "/sel_flag"])
        close(G[key + "/given"][flag].T, G[key + "/sel_traj"], "selected trajectory")


@pytest.mark.gpu
def test_selection_kernel_matches_reference_selection():
    """planner_select_kernel against what the reference's own selection code returned (tests/golden/make_nlp_golden.py)."""
    import ctypes as C
    import torch
    from car_racing_b200 import _capi, batch
    h = batch.default_handle()
    L = _capi.lib()
    N = 10
    for n in range(int(G["num_plans"])):
        key = "plan%d" % n
        nv, old = int(G[key + "/num_veh"]), int(G[key + "/old"])
        given = np.ascontiguousarray(G[key + "/given"].transpose(0, 2, 1))          # (C, N+1, 6)
        Cn = given.shape[0]
        rec = np.zeros(Cn, dtype=_capi.RECORD_DTYPE)
        riv = np.ascontiguousarray(G[key + "/obs"][:, 4:6, :N + 1])
        sel = _capi.PlannerSelectParams()
        sel.C, sel.N, sel.num_veh, sel.N_ctrl, sel.M_ctrl, sel.old_direction_flag = Cn, N, nv, N, 0, old
        sel.veh_length, sel.veh_width, sel.lap_length = 0.4, 0.2, scenarios.LAP_LENGTH["goggle"]
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        d_rec = torch.from_numpy(rec.view(np.uint8)).cuda()
        d_x, d_riv = dev(given), dev(riv)
        d_ok, d_reg = dev(np.ones(Cn, dtype=np.int32)), dev(np.arange(Cn, dtype=np.int32))
        d_cost, d_flag, d_traj = torch.zeros(Cn, dtype=torch.float64, device="cuda"), torch.zeros(2, dtype=torch.int32, device="cuda"), \
            torch.zeros((N + 1, 6), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        rc = L.b200mpc_planner_select_device(h.ptr, C.byref(sel), d_rec.data_ptr(), d_x.data_ptr(), d_x.data_ptr(), d_ok.data_ptr(),
                                             d_reg.data_ptr(), d_riv.data_ptr(), d_cost.data_ptr(), d_flag.data_ptr(), d_traj.data_ptr(), None)
        h.check(rc, "b200mpc_planner_select_device")
        h.synchronize()
        assert int(d_flag[0].item()) == int(G[key + "/sel_flag"])
        assert np.abs(d_traj.cpu().numpy() - G[key + "/sel_traj"]).max() == 0.0


def test_oracle_solutions_satisfy_the_reference_statement(oracle, monkeypatch):
    """make_nlp_golden.py evaluated the REFERENCE's recorded cost and constraints at the point the oracle returns for the
    same inputs (through our shims): wherever the oracle converged, that point satisfies every row of the reference's
    own problem and the reported cost is the reference's cost there.  The oracle is re-run here so that a drift of the
    oracle away from the stored solutions is caught as well."""
    n_ok = 0
    for key in sorted({k[:-len("sol_cost_ref")] for k in G.files if k.endswith("sol_cost_ref")}):
        status = int(G[key + "sol_status"]) if key + "sol_status" in G else 0
        if key + "sol_x0_feasible" in G and not int(G[key + "sol_x0_feasible"]):
            assert status != 0          # x_0 violates a stage-0 row: the reference's IPOPT fails, ours reports failure too
            continue
        if status != 0:
            continue
        ref, rep = float(G[key + "sol_cost_ref"]), float(G[key + "sol_cost_reported"])
        assert abs(ref - rep) < 1e-8 * (1 + abs(ref)), (key, ref, rep)
        assert float(G[key + "sol_eq_max"]) < 1e-9 and float(G[key + "sol_ineq_min"]) > -1e-9, key
        n_ok += 1
    assert n_ok >= 16
    assert abs(float(G["mpc_lti0/sol_cost_ref"]) - 22.71576162) < 1e-6        # SURVEY 8(c) anchor, now the reference's own cost
    # re-solve the MPC-CBF cases with the current oracle
    lap = scenarios.LAP_LENGTH["l_shape"]
    for k in range(3):
        key = "mpccbf%d" % k
        got = {}

        def solve(x0, xt, obs, lap_off, prm, want=None, handle=None, **kw):
            got["r"] = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, **kw)
            return got["r"]
        monkeypatch.setattr(control.batch, "solve_cbf_batch", solve)
        vehicles = {"ego": Rival(0, 0, 0)}
        for j, (s0, v, ey) in enumerate(G[key + "/rivals"]):
            vehicles["car%d" % (j + 1)] = Rival(s0, v, ey)
        prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 4.0, 0, 40.0]),
                                    matrix_R=np.diag([0.1, 0.1]), num_horizon=20, alpha=float(G[key + "/alpha"]))
        control.mpccbf(G[key + "/x0"], G[key + "/xtarget"].reshape(6, 1), prm, vehicles, "ego", lap, float(G[key + "/time"]), 0.1, False,
                       types.SimpleNamespace(width=1.0, lap_length=lap), SYSP)
        assert abs(got["r"]["cost"][0] - float(G[key + "/sol_cost_reported"])) < 1e-9


def test_convex_paths_against_an_independent_solve_of_the_reference_qp(oracle, monkeypatch):
    """For the convex problems the optimum is unique, so any correct solver pins it.  make_nlp_golden.py read H, g and the
    row matrices off the reference's recorded closures and solved the REFERENCE's own QP with scipy (equalities
    eliminated, SLSQP from a cold start).  The oracle -- re-run here through our shims for MPC-LTI, stored for LMPC and
    the planner candidates -- must land on the same optimum: |du0| < 1e-5, |dcost| < 1e-6 (the oracle's cost sits ~1e-7
    above scipy's: the barrier residual at tol 1e-8)."""
    n = 0
    for key in sorted({k[:-len("scipy_cost")] for k in G.files if k.endswith("scipy_cost")}):
        assert float(G[key + "scipy_viol"]) < 1e-9
        dc = float(G[key + "sol_cost_reported"]) - float(G[key + "scipy_cost"])
        assert -1e-7 < dc < 1e-6, (key, dc)
        if key + "scipy_u0" in G:
            assert np.abs(G[key + "scipy_u0"] - G[key + "oracle_u0"]).max() < 1e-5, key
        else:
            assert np.abs(G[key + "scipy_x"] - G[key + "oracle_x"]).max() < 1e-4, key
        n += 1
    assert n >= 10
    for k in range(2):
        key = "mpc_lti%d" % k
        got = {}

        def solve(x0, xt, obs, lap_off, prm, want=None, handle=None, **kw):
            got["r"] = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, **kw)
            return got["r"]
        monkeypatch.setattr(control.batch, "solve_cbf_batch", solve)
        prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 4.0, 0, 40.0]),
                                    matrix_R=np.diag([0.1, 0.1]), num_horizon=10)
        track = types.SimpleNamespace(width=0.8, lap_length=scenarios.LAP_LENGTH["l_shape"])
        u = control.mpc_lti(G[key + "/x0"], G[key + "/xtarget"].reshape(6, 1), prm, SYSP, track)
        assert np.abs(u - G[key + "/scipy_u0"]).max() < 1e-5
        assert abs(got["r"]["cost"][0] - float(G[key + "/scipy_cost"])) < 1e-6


@pytest.mark.gpu
def test_gpu_solutions_against_the_independent_solve_of_the_reference_qp():
    """The CUDA path itself (through the drop-in shims) on the convex cases: MPC-LTI, LMPC and the planner candidates land
    on the optimum scipy found for the REFERENCE's own recorded QP."""
    for k in range(2):
        key = "mpc_lti%d" % k
        prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 4.0, 0, 40.0]),
                                    matrix_R=np.diag([0.1, 0.1]), num_horizon=10)
        track = types.SimpleNamespace(width=0.8, lap_length=scenarios.LAP_LENGTH["l_shape"])
        u = control.mpc_lti(G[key + "/x0"], G[key + "/xtarget"].reshape(6, 1), prm, SYSP, track)
        assert np.abs(u - G[key + "/scipy_u0"]).max() < 1e-5
    key = "lmpc1"
    lp = types.SimpleNamespace(num_horizon=12, num_ss_iter=2, num_ss_points=44, shift=0, matrix_Q=G[key + "/matrix_Q"],
                               matrix_R=np.diag([1.0, 0.25]), matrix_dR=np.diag([4.0, 0.0]))
    out = control.lmpc(G[key + "/x0"], lp, list(G[key + "/Atv"]), list(G[key + "/Btv"]), list(G[key + "/Ctv"]), G[key + "/ss"],
                       G[key + "/Qfun"], int(G[key + "/it"]), 25.0, float(G[key + "/lap_width"]), G[key + "/u_old"], SYSP)
    assert np.abs(out[0][0] - G[key + "/scipy_u0"]).max() < 1e-5
    checked = 0
    for n in range(int(G["num_plans"])):
        key = "plan%d" % n
        p = _planner(int(G[key + "/seed"]), int(G[key + "/num_veh"]), int(G[key + "/old"]), G[key + "/obs"])
        traj, flag, st, sol = planning.solve_optimization_problem(p)
        for c in range(sol.shape[0]):
            ck = f"{key}/cand{c}/scipy_x"
            if ck in G:
                assert np.abs(sol[c].T - G[ck]).max() < 1e-4, ck
                assert abs(p.candidate_costs[c] - float(G[f"{key}/cand{c}/scipy_cost"])) < 1e-6
                checked += 1
    assert checked >= 7

"""Generates tests/golden/nlp_golden.npz: the reference's OWN statement of every optimisation problem on the hot path,
evaluated numerically (build container only; the .npz travels).

The UNMODIFIED reference functions control.mpc_lti / mpccbf / mpc_multi_agents / lmpc (car_racing/control/control.py)
and OvertakeTrajPlanner.generate_traj_per_region / solve_optimization_problem (car_racing/planning/
overtake_traj_planner.py) are imported from /root/reference with tests/golden/casadi_recorder.py standing in for
`casadi`: every `opti.subject_to` / `opti.minimize` of the reference's code is recorded as a numeric closure and
`opti.solve()` raises RuntimeError (which the reference catches itself, except in mpc_lti).  For each case the
script stores the inputs, random points in the reference's variable layout and the reference's cost / equality /
inequality values at those points.  For the planner it also stores what the reference's own failure branch
(the heuristic trajectory, :365-374) and its own selection code (:205-246, run with the candidate solves
replaced by given trajectories) return.

    python tests/golden/make_nlp_golden.py
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import casadi_recorder as rec                      # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    sys.modules["casadi"] = rec
    solvers = _stub("cvxopt.solvers", qp=None)
    _stub("cvxopt", matrix=None, spmatrix=None, solvers=solvers)
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.animation"]:
        _stub(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.path.insert(0, os.path.join(REF, "car_racing"))
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "car_racing"))
    try:
        from control import control as ref_control
        from planning import overtake_traj_planner as ref_planner
    finally:
        os.chdir(cwd)
    return ref_control, ref_planner


def evaluate(opti, rng, R=3, scale=1.0, zero=()):
    """Random points in the reference's variable layout and its cost / constraint values there."""
    pts = [[(0.0 if k in zero else scale) * rng.normal(size=v.shape) for k, v in enumerate(opti.variables)] for _ in range(R)]
    cost, eq, ine = [], [], []
    for p in pts:
        ctx = {v: a for v, a in zip(opti.variables, p)}
        cost.append(opti.eval_cost(ctx))
        e, i = opti.eval_constraints(ctx)
        eq.append(e)
        ine.append(i)
    out = {"cost": np.array(cost), "eq": np.array(eq), "ineq": np.array(ine)}
    for k in range(len(opti.variables)):
        out["var%d" % k] = np.array([p[k] for p in pts])
    return out


def single_variable_bounds(opti):
    """For a problem whose inequalities are all bounds on single variables (the planner's candidate QP): the box
    [lb, ub] each variable ends up with, read off the reference's recorded rows numerically (row = coef * var + const >= 0)."""
    zero = {v: np.zeros(v.shape) for v in opti.variables}
    _, base = opti.eval_constraints(zero)
    boxes = [np.stack([np.full(v.shape, -np.inf), np.full(v.shape, np.inf)]) for v in opti.variables]
    hit = np.zeros(base.size, dtype=int)
    for k, v in enumerate(opti.variables):
        for idx in np.ndindex(*v.shape):
            ctx = dict(zero)
            e = np.zeros(v.shape)
            e[idx] = 1.0
            ctx[v] = e
            _, r1 = opti.eval_constraints(ctx)
            coef = r1 - base
            for row in np.nonzero(coef)[0]:
                hit[row] += 1
                b = -base[row] / coef[row]
                if coef[row] > 0:
                    boxes[k][0][idx] = max(boxes[k][0][idx], b)
                else:
                    boxes[k][1][idx] = min(boxes[k][1][idx], b)
    assert (hit == 1).all(), "an inequality row is not a single-variable bound"
    return boxes


def at_solution(opti, values, cost_reported):
    """The reference's own cost / constraints at a solver's returned point (values in the reference's variable layout)."""
    ctx = {v: np.asarray(a, float).reshape(v.shape) for v, a in zip(opti.variables, values)}
    e, i = opti.eval_constraints(ctx)
    return dict(sol_cost_ref=opti.eval_cost(ctx), sol_cost_reported=float(cost_reported), sol_eq_max=float(np.abs(e).max()),
                sol_ineq_min=float(i.min()))


def _ctx_from_w(opti, w):
    ctx, o = {}, 0
    for v in opti.variables:
        n = v.shape[0] * v.shape[1]
        ctx[v] = np.asarray(w[o:o + n], float).reshape(v.shape)
        o += n
    return ctx


def independent_qp_solution(opti, fixed_zero=()):
    """For the CONVEX problems (quadratic cost, linear rows): read H, g and the row matrices off the reference's recorded
    closures (exact for quadratics: second differences at unit vectors) and solve the reference's own QP with scipy's
    trust-constr from a cold start -- an optimum of the reference's problem that owes nothing to our solver.
    Returns (w*, cost*)."""
    from scipy.optimize import minimize
    sizes = [v.shape[0] * v.shape[1] for v in opti.variables]
    n = sum(sizes)
    f = lambda w: opti.eval_cost(_ctx_from_w(opti, w))
    z = np.zeros(n)
    f0 = f(z)
    E = np.eye(n)
    fi = np.array([f(E[i]) for i in range(n)])
    H = np.zeros((n, n))
    for i in range(n):
        H[i, i] = f(2 * E[i]) - 2 * fi[i] + f0
    # off-diagonal terms only where both variables carry curvature or the cost couples them: brute force over pairs
    for i in range(n):
        for j in range(i + 1, n):
            hij = f(E[i] + E[j]) - fi[i] - fi[j] + f0
            if abs(hij) > 1e-12:
                H[i, j] = H[j, i] = hij
    g = fi - f0 - 0.5 * np.diag(H)
    e0, i0 = opti.eval_constraints(_ctx_from_w(opti, z))
    Aeq = np.zeros((e0.size, n))
    Ain = np.zeros((i0.size, n))
    for i in range(n):
        e1, i1 = opti.eval_constraints(_ctx_from_w(opti, E[i]))
        Aeq[:, i], Ain[:, i] = e1 - e0, i1 - i0
    # eliminate the equality rows (x_0 and the dynamics fix the states): w = w_p + Z y, then SLSQP on the reduced, well
    # conditioned QP in the inputs (and lambda) with exact gradients
    from scipy.linalg import null_space
    w_p = np.linalg.lstsq(Aeq, -e0, rcond=None)[0]
    Z = null_space(Aeq)
    Hr, gr = Z.T @ H @ Z, Z.T @ (H @ w_p + g)
    Ar, br = Ain @ Z, Ain @ w_p + i0
    fun = lambda y: 0.5 * y @ Hr @ y + gr @ y
    best = None
    for y0 in (np.zeros(Z.shape[1]),):
        res = minimize(fun, y0, jac=lambda y: Hr @ y + gr, method="SLSQP",
                       constraints=[{"type": "ineq", "fun": lambda y: Ar @ y + br, "jac": lambda y: Ar}],
                       options=dict(maxiter=2000, ftol=1e-15))
        if best is None or res.fun < best.fun:
            best = res
    w = w_p + Z @ best.x
    viol = max(np.abs(Aeq @ w + e0).max(), max(0.0, -(Ain @ w + i0).min()))
    return w, float(f(w)), float(viol), int(best.status)


class OracleSolve:
    """Routes the drop-in shims' batch calls to the CPU oracle (as the tests do) and keeps the last result."""

    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle
        self.orc, self.last = oracle, None

    def cbf(self, x0, xt, obs, lap_off, prm, want=None, handle=None, **kw):
        self.last = self.orc.solve_cbf_batch(x0, xt, obs, lap_off, prm, **kw)
        return self.last

    def lmpc(self, *a, want=None, handle=None, **kw):
        self.last = self.orc.solve_lmpc_batch(*a, **kw)
        return self.last


def put(store, prefix, d):
    for k, v in d.items():
        store["%s/%s" % (prefix, k)] = np.asarray(v)


def main():
    from car_racing_b200 import scenarios
    from planner_cases import make_planner
    from test_shims_host import Rival
    ref_control, ref_planner = import_reference()
    from car_racing_b200 import control as our_control, planning as our_planning
    osolve = OracleSolve()
    our_control.batch.solve_cbf_batch = osolve.cbf
    our_control.batch.solve_lmpc_batch = osolve.lmpc
    rng = np.random.default_rng(2024)
    store = {}
    lap = scenarios.LAP_LENGTH["l_shape"]
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    Q, Rm = np.diag([10.0, 0, 0, 4.0, 0, 40.0]), np.diag([0.1, 0.1])

    # ---- a1 mpc_lti (control.py:198-248), config 1: N=10, width 0.8
    for k, x0 in enumerate([np.zeros(6), np.array([0.9, 0.02, -0.1, 0.03, 7.5, -0.2])]):
        prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=Q, matrix_R=Rm, num_horizon=10)
        track = types.SimpleNamespace(width=0.8, lap_length=lap)
        xt = np.array([0.8, 0, 0, 0, 0, 0.0]).reshape(6, 1)
        rec.clear()
        try:
            ref_control.mpc_lti(x0, xt, prm, sysp, track)
        except RuntimeError:
            pass
        d = evaluate(rec.LAST[-1], rng)
        our_control.mpc_lti(x0, xt, prm, sysp, track)
        r = osolve.last
        d.update(at_solution(rec.LAST[-1], [r["x"][0].T, r["u"][0].T], r["cost"][0]))
        w, c, viol, st = independent_qp_solution(rec.LAST[-1])
        ctx = _ctx_from_w(rec.LAST[-1], w)
        d.update(scipy_u0=ctx[rec.LAST[-1].variables[1]][:, 0], scipy_cost=c, scipy_viol=viol, scipy_status=st, oracle_u0=r["u0"][0])
        print("mpc_lti", k, "scipy cost", c, "oracle cost", r["cost"][0], "du0", np.abs(d["scipy_u0"] - r["u0"][0]).max(), "viol", viol, flush=True)
        d.update(x0=x0, xtarget=xt.ravel(), N=10, width=0.8)
        put(store, "mpc_lti%d" % k, d)

    # ---- a2 mpccbf (control.py:476-607): N=20, rivals nearby / on another lap count / filtered out / moving
    cases = [
        dict(x0=[0.9, 0.0, 0.0, 0.02, 3.0, 0.05], rivals=[(4.0, 0.2, 0.1), (4.6, 0.0, -0.4), (3.9, 0.1, 0.6)], alpha=0.8),
        dict(x0=[1.2, 0.01, 0.05, -0.02, 2.0, -0.1], rivals=[(lap + 3.0, 0.3, 0.2), (2 * lap + 2.5, 0.0, -0.3), (12.0, 0.2, 0.0)], alpha=0.8),
        dict(x0=[0.6, 0.0, 0.0, 0.0, lap + 1.0, 0.3], rivals=[(1.8, 0.1, 0.25), (lap + 1.9, 0.2, -0.2)], alpha=0.5),
    ]
    for k, cs in enumerate(cases):
        vehicles = {"ego": Rival(0, 0, 0)}
        for j, (s0, v, ey) in enumerate(cs["rivals"]):
            vehicles["car%d" % (j + 1)] = Rival(s0, v, ey)
        prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=Q, matrix_R=Rm, num_horizon=20,
                                    alpha=cs["alpha"])
        track = types.SimpleNamespace(width=1.0, lap_length=lap)
        xt = np.array([0.8, 0, 0, 0, 0, 0.0]).reshape(6, 1)
        x0 = np.array(cs["x0"], float)
        rec.clear()
        ref_control.mpccbf(x0, xt, prm, vehicles, "ego", lap, 0.3, 0.1, False, track, sysp)
        opti = rec.LAST[-1]
        d = evaluate(opti, rng, scale=1.0)
        # a second set close to the rivals, where the degree-6 barrier is O(1) instead of O(1e6)
        near = evaluate(opti, rng, scale=0.05)
        for kk, vv in near.items():
            if kk.startswith("var0"):
                xs = vv.copy()
                xs[:, 4, :] += x0[4] + 1.0
                near[kk] = xs
        ctxs = [{v: near["var%d" % i][r] for i, v in enumerate(opti.variables)} for r in range(3)]
        near["cost"] = np.array([opti.eval_cost(c) for c in ctxs])
        ec = [opti.eval_constraints(c) for c in ctxs]
        near["eq"], near["ineq"] = np.array([e for e, _ in ec]), np.array([i for _, i in ec])
        d.update({"near_" + kk: vv for kk, vv in near.items()})
        our_control.mpccbf(x0, xt, prm, vehicles, "ego", lap, 0.3, 0.1, False, track, sysp)
        r = osolve.last
        d.update(at_solution(opti, [r["x"][0].T, r["u"][0].T, r["sigma"][0]], r["cost"][0]))
        d.update(sol_status=int(r["status"][0]), sol_elastic=float(r["elastic_max"][0]))
        d.update(x0=x0, xtarget=xt.ravel(), N=20, alpha=cs["alpha"], time=0.3, rivals=np.array(cs["rivals"]),
                 num_slack_rows=opti.variables[2].shape[0])
        put(store, "mpccbf%d" % k, d)

    # ---- a3 mpc_multi_agents (control.py:251-473): N=10, alpha 0.6, margin 0.15, per-stage targets from a trajectory
    for k in range(2):
        lapg = scenarios.LAP_LENGTH["goggle"]
        vehicles = {"ego": Rival(0, 0, 0), "car1": Rival(5.0 + k, 1.0, -0.5), "car2": Rival(5.6, 0.9, 0.2), "car3": Rival(30.0, 1.0, 0.0)}
        prm = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 5.0, 0, 50.0]),
                                    matrix_R=Rm, num_horizon_ctrl=10)
        track = types.SimpleNamespace(width=1.0, lap_length=lapg)
        traj = np.zeros((11, 6))
        traj[:, 4] = 4.0 + 0.15 * np.arange(11)
        traj[:, 5] = 0.05 * np.arange(11) * (1 - 2 * k)
        x0 = np.array([1.5, 0, 0, 0, 3.9 + 0.4 * k, 0.1])
        rec.clear()
        ref_control.mpc_multi_agents(x0, prm, track, None, None, None, sysp, target_traj_xcurv=traj, vehicles=vehicles,
                                     agent_name="ego", direction_flag=0, target_traj_xglob=None,
                                     sorted_vehicles=["car1", "car2", "car3"], time=None)
        opti = rec.LAST[-1]
        d = evaluate(opti, rng)
        our_control.mpc_multi_agents(x0, prm, track, None, None, None, sysp, target_traj_xcurv=traj, vehicles=vehicles,
                                     agent_name="ego", direction_flag=0, sorted_vehicles=["car1", "car2", "car3"], time=None)
        r = osolve.last
        d.update(at_solution(opti, [r["x"][0].T, r["u"][0].T, r["sigma"][0]], r["cost"][0]))
        d.update(sol_status=int(r["status"][0]), sol_elastic=float(r["elastic_max"][0]))
        d.update(x0=x0, traj=traj, N=10, rivals=np.array([(5.0 + k, 1.0, -0.5), (5.6, 0.9, 0.2), (30.0, 1.0, 0.0)]),
                 num_slack_rows=opti.variables[2].shape[0])
        put(store, "multi%d" % k, d)

    # ---- a6 lmpc (control.py:610-730) incl. lmpc_helper.select_points: N=12, 2 x 22 safe-set points
    for k in range(2):
        T, laps = 120, 3
        ss = 10000.0 * np.ones((T, 6, laps))
        Qf = 10000.0 * np.ones((T, laps))
        for lp in range(laps):
            x = np.array([1.0 + 0.1 * lp, 0, 0, 0.01 * lp, 0.0, 0.05 * lp])
            for t in range(T - 10):
                ss[t, :, lp] = x
                Qf[t, lp] = (T - 10) - t + 3 * lp
                u = np.clip([-0.6 * x[5] - 0.9 * x[3], 1.5 * (1.2 - x[0])], [-0.5, -1.0], [0.5, 1.0])
                x = scenarios.LTI_A @ x + scenarios.LTI_B @ u
        x0 = ss[30 + 7 * k, :, 2] + rng.normal(scale=[0.02, 0.003, 0.01, 0.003, 0.02, 0.01])
        N = 12
        Atv = [scenarios.LTI_A + 1e-3 * rng.normal(size=(6, 6)) for _ in range(N)]
        Btv = [scenarios.LTI_B + 1e-3 * rng.normal(size=(6, 2)) for _ in range(N)]
        Ctv = [1e-3 * rng.normal(size=(6, 1)) for _ in range(N)]
        lp_prm = types.SimpleNamespace(num_horizon=N, num_ss_iter=2, num_ss_points=44, shift=0, matrix_Q=0 * np.eye(6),
                                       matrix_R=np.diag([1.0, 0.25]), matrix_dR=np.diag([4.0, 0.0]),
                                       matrix_Qslack=5 * np.diag([10, 0, 0, 1, 10, 1]))
        if k == 1:
            lp_prm.matrix_Q = np.diag([1.0, 0.5, 0.2, 0.1, 0.0, 2.0])
        u_old = np.array([[0.05, -0.3]]) if k == 0 else np.array([-0.1, 0.4])
        rec.clear()
        out = ref_control.lmpc(x0, lp_prm, Atv, Btv, Ctv, ss, Qf, 3, 25.0, 0.9, u_old, sysp)
        opti = rec.LAST[-1]
        d = evaluate(opti, rng, zero=(3,))        # slack is forced to 0 by :693-694; evaluate there
        lp2 = types.SimpleNamespace(**{kk: vv for kk, vv in lp_prm.__dict__.items() if kk != "matrix_Qslack"})
        our_control.lmpc(x0, lp2, Atv, Btv, Ctv, ss, Qf, 3, 25.0, 0.9, u_old, sysp)
        r = osolve.last
        d.update(at_solution(opti, [r["x"][0].T, r["u"][0].T, r["lam"][0], np.zeros(6)], r["cost"][0]))
        d.update(sol_status=int(r["status"][0]), oracle_u0=r["u0"][0])
        if int(r["status"][0]) == 0:
            w, c, viol, st = independent_qp_solution(opti)
            ctx = _ctx_from_w(opti, w)
            d.update(scipy_u0=ctx[opti.variables[1]][:, 0], scipy_cost=c, scipy_viol=viol, scipy_status=st)
            print("lmpc", k, "scipy cost", c, "oracle cost", r["cost"][0], "du0", np.abs(d["scipy_u0"] - r["u0"][0]).max(), "viol", viol, flush=True)
        d.update(x0=x0, ss=ss, Qfun=Qf, it=3, Atv=np.array(Atv), Btv=np.array(Btv), Ctv=np.array(Ctv), u_old=np.asarray(u_old).ravel(),
                 matrix_Q=lp_prm.matrix_Q, lap_width=0.9, ss_sel=out[2], Qfun_sel=out[3])
        put(store, "lmpc%d" % k, d)

    # ---- a4 generate_traj_per_region (overtake_traj_planner.py:248-379): candidate QPs + the reference's own fallback;
    #      a5 solve_optimization_problem (:162-246): the reference's own selection with given candidate trajectories
    class FakeProcess:
        def __init__(self, target, args):
            self.target, self.args = target, args

        def start(self):
            self.target(*self.args)

        def join(self):
            pass

    class FakeManager:
        def dict(self):
            return {}
    ref_planner.Process, ref_planner.Manager = FakeProcess, FakeManager
    nplan = 0
    for seed, old in ((3, None), (7, 1), (11, 0), (12, 2), (14, None)):
        p = make_planner(seed, num_veh=2 + seed % 2)
        p.old_direction_flag = old
        if seed == 12:     # a rival whose prediction is one lap ahead in s: the in-place lap wrap (:291-292)
            name = p.sorted_vehicles[0]
            p.obs_infos[name][4] += p.track.lap_length
        obs_in = {n: p.obs_infos[n].copy() for n in p.sorted_vehicles}
        C = len(p.sorted_vehicles) + 1
        trajs, fallback = {}, {}
        for c in range(C):
            rec.clear()
            dt, ds, dc = {}, {}, {}
            ref_planner.OvertakeTrajPlanner.generate_traj_per_region(p, c, dt, ds, dc)
            fallback[c] = dt[c].copy()                # the reference's heuristic trajectory (solve() raised)
            d = evaluate(rec.LAST[-1], rng)
            N_ = p.racing_game_param.num_horizon_planner
            ego_x = np.asarray(p.vehicles["ego"].xcurv, float)
            xlb, xub = our_planning.candidate_bounds(c, p.xcurv_ego, p.sorted_vehicles, p.obs_infos, 0.4, 0.2, p.track.width,
                                                     p.track.lap_length, N_)
            s_ref, ey_ref = our_planning.candidate_targets(c, ego_x, p.bezier_xcurvs, p.bezier_funcs, N_)
            kw, off = our_planning.pack_candidates(ego_x, s_ref[None], ey_ref[None], xlb[None], xub[None], N_)
            pprm = our_planning.planner_params(p.racing_game_param.matrix_A, p.racing_game_param.matrix_B, N_)
            r = osolve.orc.solve_cbf_batch(kw["x0"], kw["xt"], kw["obs"], None, pprm, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"])
            d.update(sol_status=int(r["status"][0]), sol_x0_feasible=int(our_planning.x0_feasible(ego_x, xlb, xub)))
            d.update(at_solution(rec.LAST[-1], [r["x"][0].T, r["u"][0].T], r["cost"][0] + off[0]))
            d.update(oracle_x=r["x"][0])
            if int(r["status"][0]) == 0:
                w, cst, viol, st = independent_qp_solution(rec.LAST[-1])
                ctx = _ctx_from_w(rec.LAST[-1], w)
                d.update(scipy_x=ctx[rec.LAST[-1].variables[0]].T, scipy_cost=cst, scipy_viol=viol, scipy_status=st)
                print("plan", nplan, c, "scipy cost", cst, "oracle cost", r["cost"][0] + off[0], "dx", np.abs(d["scipy_x"] - r["x"][0]).max(), flush=True)
            bx = single_variable_bounds(rec.LAST[-1])
            d.update(x_lb=bx[0][0], x_ub=bx[0][1], u_lb=bx[1][0], u_ub=bx[1][1])
            put(store, "plan%d/cand%d" % (nplan, c), d)
        # selection: the reference's own code with candidate "solutions" = perturbed heuristic trajectories
        given = {c: fallback[c] + rng.normal(scale=0.05, size=fallback[c].shape) * np.array([0, 0, 0, 0, 1, 1])[:, None] for c in range(C)}

        def fake_generate(self, pos_index, dict_traj, dict_solve_time, dict_cost, given=given):
            dict_traj[pos_index] = given[pos_index]
            dict_solve_time[pos_index] = 0.0
            dict_cost[pos_index] = 0.0
        p.generate_traj_per_region = types.MethodType(fake_generate, p)
        traj, flag, st, sol = ref_planner.OvertakeTrajPlanner.solve_optimization_problem(p)
        put(store, "plan%d" % nplan, dict(ego_x=np.asarray(p.vehicles["ego"].xcurv, float), xcurv_ego=np.asarray(p.xcurv_ego, float), seed=seed, num_veh=len(p.sorted_vehicles), old=-1 if old is None else old,
                                          obs=np.array([obs_in[n] for n in p.sorted_vehicles]),
                                          fallback=np.array([fallback[c] for c in range(C)]),
                                          given=np.array([given[c] for c in range(C)]), sel_traj=traj, sel_flag=flag, sel_sol=sol))
        nplan += 1
    store["num_plans"] = np.array(nplan)
    out = os.path.join(HERE, "nlp_golden.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, len(store), "arrays,", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

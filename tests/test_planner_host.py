"""CPU tests of the planner path: the mapping of the candidate QP onto the batched solver is checked against an
independent scipy solve of the reference's own formulation (overtake_traj_planner.py:263-334); the drop-in
solve_optimization_problem is exercised with the oracle injected as the solver (no GPU here)."""
import numpy as np
import pytest

from car_racing_b200 import planning
from planner_cases import make_planner, reference_qp_scipy


def _oracle_solver(oracle):
    def solve(x0, xt, obs, lap_off, prm, xlb=None, xub=None, wd=None, **kw):
        return oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, xlb=xlb, xub=xub, wd=wd)
    return solve


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_candidate_qp_mapping_matches_reference_formulation(oracle, seed):
    p = make_planner(seed)
    N = p.racing_game_param.num_horizon_planner
    ego = p.vehicles["ego"]
    for c in range(len(p.sorted_vehicles) + 1):
        xlb, xub = planning.candidate_bounds(c, p.xcurv_ego, p.sorted_vehicles, p.obs_infos, 0.4, 0.2, 1.0, p.track.lap_length, N)
        if not planning.x0_feasible(ego.xcurv, xlb, xub) or (xlb[1:N, 1] > xub[1:N, 1]).any():
            continue
        s_ref, ey_ref = planning.candidate_targets(c, ego.xcurv, p.bezier_xcurvs, p.bezier_funcs, N)
        prm = planning.planner_params(p.racing_game_param.matrix_A, p.racing_game_param.matrix_B, N)
        kw, off = planning.pack_candidates(ego.xcurv, s_ref[None], ey_ref[None], xlb[None], xub[None], N)
        r = oracle.solve_cbf_batch(kw["x0"], kw["xt"], kw["obs"], None, prm, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"])
        ref, xr = reference_qp_scipy(p, c)
        assert ref is not None
        if r["status"][0] != 0:
            continue
        assert abs((r["cost"][0] + off[0]) - ref.fun) < 2e-5 * max(1.0, abs(ref.fun)), (c, r["cost"][0] + off[0], ref.fun)
        assert np.abs(r["x"][0][:, 4:6] - xr[:, 4:6]).max() < 2e-3


def test_drop_in_solve_optimization_problem_with_oracle(oracle):
    p = make_planner(3)
    p.old_direction_flag = 1
    traj, flag, solve_time, sol = planning.solve_optimization_problem(p, solver=_oracle_solver(oracle))
    C = len(p.sorted_vehicles) + 1
    N = p.racing_game_param.num_horizon_planner
    assert traj.shape == (N + 1, 6) and sol.shape == (C, 6, N + 1) and solve_time.shape == (C,) and 0 <= flag < C
    assert np.allclose(traj, sol[flag].T)
    sel = planning.selection_costs(sol, p.sorted_vehicles, p.obs_infos, 0.4, 0.2, p.track.lap_length, 1)
    assert flag == sel.index(min(sel))
    for c in range(C):
        assert np.allclose(sol[c][:, 0], p.vehicles["ego"].xcurv) or not np.isfinite(p.candidate_costs[c])


def test_infeasible_candidate_uses_reference_fallback(oracle):
    p = make_planner(4)
    name = p.sorted_vehicles[0]
    p.obs_infos[name][4] = p.xcurv_ego[4] + 0.1 * np.arange(11) * p.xcurv_ego[0]    # right next to the ego ...
    p.obs_infos[name][5] = p.xcurv_ego[5] + 0.5                                       # ... and to its left: region 1 infeasible at k=0
    traj, flag, _, sol = planning.solve_optimization_problem(p, solver=_oracle_solver(oracle))
    assert not np.isfinite(p.candidate_costs[1])
    h = planning.heuristic_traj(1, p.xcurv_ego, p.bezier_xcurvs, p.bezier_funcs, 10)
    assert np.allclose(sol[1], h)

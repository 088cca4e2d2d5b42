"""Synthetic planner scenarios shaped like car_racing/tests/overtake_planner_test.py:151-155 (test infrastructure)."""
import types

import numpy as np
from scipy.interpolate import interp1d

from car_racing_b200 import scenarios


def make_planner(seed=0, num_veh=2, N=10, track="goggle"):
    """A duck-typed stand-in for the reference's OvertakeTrajPlanner after get_local_traj has prepared it
    (overtake_traj_planner.py:66-117): sorted rivals, their predictions, Bezier reference curves."""
    rng = np.random.default_rng(seed)
    lap = scenarios.LAP_LENGTH[track]
    vx = rng.uniform(1.0, 1.6)
    ego_x = np.array([vx, rng.uniform(-.02, .02), rng.uniform(-.05, .05), rng.uniform(-.03, .03), rng.uniform(2, lap - 6),
                      rng.uniform(-.2, .2)])
    vehicles = {"ego": types.SimpleNamespace(param=types.SimpleNamespace(length=0.4, width=0.2), xcurv=ego_x)}
    names, obs_infos = [], {}
    eys = np.sort(rng.uniform(-0.6, 0.6, size=num_veh))          # sorted_vehicles is ordered by ey (left to right)
    for j in range(num_veh):
        name = "car%d" % (j + 1)
        traj = np.zeros((6, N + 1))
        v = rng.uniform(0.8, 1.3)
        traj[4] = ego_x[4] + rng.uniform(0.3, 1.2) + v * 0.1 * np.arange(N + 1)
        traj[5] = eys[j]
        traj[0] = v
        names.append(name)
        obs_infos[name] = traj
        vehicles[name] = types.SimpleNamespace(param=types.SimpleNamespace(length=0.4, width=0.2), xcurv=traj[:, 0].copy())
    # one Bezier-like reference per region: start at ego ey, end in the gap of that region
    edges = np.concatenate([[-0.9], eys, [0.9]])
    C = num_veh + 1
    bez = np.zeros((C, N + 1, 2))
    funcs = []
    for c in range(C):
        s = ego_x[4] + np.linspace(0.0, 1.2 * vx * 0.1 * N + 0.5, N + 1)
        tgt = 0.5 * (edges[c] + edges[c + 1])
        tt = np.linspace(0, 1, N + 1)
        ey = ego_x[5] + (tgt - ego_x[5]) * (3 * tt ** 2 - 2 * tt ** 3)
        bez[c, :, 0], bez[c, :, 1] = s, ey
        funcs.append(interp1d(s, ey))
    p = types.SimpleNamespace(
        sorted_vehicles=names, obs_infos=obs_infos, old_ey=None, old_direction_flag=None, bezier_xcurvs=bez, bezier_funcs=funcs,
        xcurv_ego=ego_x.copy(), vehicles=vehicles, agent_name="ego",
        track=types.SimpleNamespace(width=1.0, lap_length=lap),
        racing_game_param=types.SimpleNamespace(num_horizon_planner=N, matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B))
    return p


def reference_qp_scipy(p, pos_index):
    """The candidate QP exactly as generate_traj_per_region states it (overtake_traj_planner.py:263-334), solved by
    scipy SLSQP on the original variables -- an independent check of the mapping used by car_racing_b200.planning."""
    from scipy.optimize import minimize
    from car_racing_b200 import planning
    N = p.racing_game_param.num_horizon_planner
    A, B = p.racing_game_param.matrix_A, p.racing_game_param.matrix_B
    ego = p.vehicles["ego"]
    x0 = np.asarray(ego.xcurv, float)
    xlb, xub = planning.candidate_bounds(pos_index, p.xcurv_ego, p.sorted_vehicles, p.obs_infos, ego.param.length, ego.param.width,
                                         p.track.width, p.track.lap_length, N)
    s_ref, ey_ref = planning.candidate_targets(pos_index, ego.xcurv, p.bezier_xcurvs, p.bezier_funcs, N)

    def roll(u):
        x = np.zeros((N + 1, 6))
        x[0] = x0
        for k in range(N):
            x[k + 1] = A @ x[k] + B @ u[2 * k:2 * k + 2]
        return x

    def cost(u):
        x = roll(u)
        c = 0.0
        for k in range(N):
            if k > 1:
                c += 30 * (x[k, 5] - x[k - 1, 5]) ** 2
        c += -200 * (x[N, 4] - x[0, 4])
        c += 20 * np.sum((x[:, 5] - ey_ref) ** 2 + (x[:, 4] - s_ref) ** 2)
        return c
    cons = []
    for k in range(1, N + 1):
        cons.append({"type": "ineq", "fun": (lambda u, k=k: 5.0 - roll(u)[k, 0])})
    for k in range(1, N):
        cons.append({"type": "ineq", "fun": (lambda u, k=k: xub[k, 1] - roll(u)[k, 5])})
        cons.append({"type": "ineq", "fun": (lambda u, k=k: roll(u)[k, 5] - xlb[k, 1])})
    bnds = [(-0.5, 0.5), (-1.5, 1.5)] * N
    best = None
    for u_init in (np.zeros(2 * N), np.tile([0.05, 0.5], N), np.tile([-0.05, -0.5], N)):
        r = minimize(cost, u_init, method="SLSQP", bounds=bnds, constraints=cons, options=dict(maxiter=500, ftol=1e-13))
        if r.success and (best is None or r.fun < best.fun):
            best = r
    return best, roll(best.x) if best is not None else None

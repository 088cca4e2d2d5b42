"""Parallel overtake-candidate evaluator: drop-in for
`OvertakeTrajPlanner.solve_optimization_problem` (car_racing/planning/overtake_traj_planner.py:162-246).

The reference forks one OS process per candidate region (`:182-197`), each building and solving one
CasADi/IPOPT QP (`generate_traj_per_region`, `:248-379`), gathers the results through a
`multiprocess.Manager().dict()` and takes the argmin of a selection cost (`:205-244`).  Here the
candidates are ONE batch: every candidate is one packed record, one kernel launch solves them all,
the selection cost and first-min argmin run on the host (a handful of scalars).

Mapping of the candidate QP onto the batched MPC kernel (include/b200mpc.h, flags STAGE_BOUNDS|EY_RATE):
  * LTI dynamics, x_0 = ego.xcurv (:263-274), no rivals rows (M = 0), R = 0, Q = diag(0,0,0,0,20,20)
    with per-stage targets (s~_j, ey_bezier(s~_j)) (:329-334);
  * -200 (s_N - s_0) (:328) is folded into the terminal target: 20 (s_N - s~_N - 5)^2 = 20 (s_N - s~_N)^2
    - 200 (s_N - s~_N) + 500, the constant is removed from the reported cost again;
  * 30 sum_{k=2}^{N-1} (ey_k - ey_{k-1})^2 (:325-327) -> ey-rate weights wd[1..N-2] = 30;
  * vx_{k+1} <= 5, |ey_k| <= width - veh_width/2 for k < N, |delta| <= 0.5, |a| <= 1.5 (:276-284);
  * per-step rows `ey_k - ey_rival,k >= veh_width + margin` for the rival on either side of the region when
    the predicted ego s is inside the rival's window (:286-324; same sign for both sides -- reference quirk)
    become per-stage lower bounds on ey_k.
"""
import ctypes as C

import numpy as np

from . import _capi, batch

SAFETY_MARGIN = 0.15          # overtake_traj_planner.py:177,262
W_EY_RATE, W_PROGRESS, W_TRACK = 30.0, 200.0, 20.0   # :327,328,333-334
VX_MAX_PLAN, DELTA_MAX_PLAN, A_MAX_PLAN = 5.0, 0.5, 1.5  # :276,280-284


def candidate_bounds(pos_index, xcurv_ego, sorted_vehicles, obs_infos, veh_length, veh_width, track_width, lap_length, N):
    """Per-stage bounds of candidate `pos_index` (overtake_traj_planner.py:276-324).
    Returns xlb, xub (N+1,2) on (vx, ey) and `feasible0` (False when a row on the fixed x_0 is violated,
    in which case the reference's IPOPT fails and the heuristic trajectory is used, :365-374)."""
    xlb = np.full((N + 1, 2), -np.inf)
    xub = np.full((N + 1, 2), np.inf)
    xub[1:, 0] = VX_MAX_PLAN                                   # vx_{k+1} <= 5 for k < N
    half = track_width - 0.5 * veh_width
    xlb[:N, 1], xub[:N, 1] = -half, half                       # ey_k, k < N
    num_veh = len(sorted_vehicles)
    for side in (pos_index - 1, pos_index):                    # rival on the left, rival on the right
        if side < 0 or side >= num_veh:
            continue
        obs_traj = obs_infos[sorted_vehicles[side]]
        for k in range(N):
            while obs_traj[4, k] > lap_length:                 # in-place wrap, as the reference does (:291-292)
                obs_traj[4, k] = obs_traj[4, k] - lap_length
            s_pred = xcurv_ego[4] + k * 0.1 * xcurv_ego[0]
            if (s_pred >= obs_traj[4, k] - veh_length - SAFETY_MARGIN) & (s_pred <= obs_traj[4, k] + veh_length + SAFETY_MARGIN):
                xlb[k, 1] = max(xlb[k, 1], obs_traj[5, k] + veh_width + SAFETY_MARGIN)
    return xlb, xub


def candidate_targets(pos_index, ego_xcurv, bezier_xcurvs, bezier_funcs, N):
    """s~_j and ey_bezier(s~_j), j = 0..N (overtake_traj_planner.py:329-334)."""
    s_ref = np.zeros(N + 1)
    ey_ref = np.zeros(N + 1)
    for j in range(N + 1):
        s_tmp = ego_xcurv[4] + 1.0 * j * ego_xcurv[0] * 0.1
        s_tmp = np.clip(s_tmp, bezier_xcurvs[pos_index, 0, 0], bezier_xcurvs[pos_index, -1, 0])
        s_ref[j] = s_tmp
        ey_ref[j] = float(bezier_funcs[pos_index](s_tmp))
    return s_ref, ey_ref


def planner_params(matrix_A, matrix_B, N):
    return dict(A=np.asarray(matrix_A, float), B=np.asarray(matrix_B, float), Q=np.diag([0, 0, 0, 0, W_TRACK, W_TRACK]),
                R=np.zeros((2, 2)), N=int(N), umax=[DELTA_MAX_PLAN, A_MAX_PLAN], vmin=0.0, vmax=VX_MAX_PLAN, width=1.0,
                alpha=0.8, margin=0.2, L=0.4, W=0.2, slack_w=1e4)


def pack_candidates(x0, s_ref, ey_ref, xlb, xub, N):
    """Stack C candidates: x0 (6,) or (C,6); s_ref, ey_ref (C,N+1); xlb, xub (C,N+1,2).
    Returns the keyword arguments for batch.solve_cbf_batch / oracle.solve_cbf_batch and the cost offsets."""
    s_ref = np.atleast_2d(np.asarray(s_ref, float))
    ey_ref = np.atleast_2d(np.asarray(ey_ref, float))
    C = s_ref.shape[0]
    x0 = np.broadcast_to(np.asarray(x0, float).reshape(-1, 6), (C, 6)).copy()
    xt = np.zeros((C, N + 1, 6))
    xt[:, :, 4] = s_ref
    xt[:, :, 5] = ey_ref
    xt[:, N, 4] += W_PROGRESS / (2.0 * W_TRACK)            # fold -200 s_N into the terminal target
    wd = np.zeros((C, N))
    wd[:, 1:N - 1] = W_EY_RATE                               # (ey_k - ey_{k-1})^2 for k = 2..N-1
    # reference cost = ours - 200 s~_N - 500 + 200 s_0
    offset = -W_PROGRESS * s_ref[:, N] - W_PROGRESS ** 2 / (4.0 * W_TRACK) + W_PROGRESS * x0[:, 4]
    kw = dict(x0=x0, xt=xt, obs=np.zeros((C, 0, 2, N + 1)), lap_off=None, xlb=np.asarray(xlb, float).reshape(C, N + 1, 2),
              xub=np.asarray(xub, float).reshape(C, N + 1, 2), wd=wd)
    return kw, offset


def x0_feasible(x0, xlb, xub):
    """The reference also imposes the stage-0 rows on the fixed x_0 (:277-278, :286-324); if x_0 violates one,
    its IPOPT solve fails and the heuristic trajectory is used."""
    return bool(xlb[0, 1] <= x0[5] <= xub[0, 1])


def heuristic_traj(pos_index, xcurv_ego, bezier_xcurvs, bezier_funcs, N):
    """Fallback of the reference when IPOPT fails (overtake_traj_planner.py:365-374)."""
    sol = np.zeros((6, N + 1))
    for j in range(N + 1):
        stmp = xcurv_ego[4] + 1.1 * j * 0.1 * xcurv_ego[0]
        sol[0, j] = 1.1 * xcurv_ego[0]
        sol[4, j] = stmp
        stmp = np.clip(stmp, bezier_xcurvs[pos_index, 0, 0], bezier_xcurvs[pos_index, -1, 0])
        sol[5, j] = float(bezier_funcs[pos_index](stmp))
    return sol


def selection_costs(solution_xvar, sorted_vehicles, obs_infos, veh_length, veh_width, lap_length, old_direction_flag):
    """overtake_traj_planner.py:205-243.  solution_xvar (C,6,N+1)."""
    C = solution_xvar.shape[0]
    N1 = solution_xvar.shape[2]
    num_veh = len(sorted_vehicles)
    cost = [0.0] * C
    for index in range(C):
        cost[index] = -10 * (solution_xvar[index, 4, -1] - solution_xvar[index, 4, 0])
        for side in (index - 1, index):
            if side < 0 or side >= num_veh:
                continue
            obs_traj = obs_infos[sorted_vehicles[side]]
            for j in range(N1):
                while obs_traj[4, j] > lap_length:
                    obs_traj[4, j] = obs_traj[4, j] - lap_length
                diffs = solution_xvar[index, 4, j] - obs_traj[4, j]
                diffey = solution_xvar[index, 5, j] - obs_traj[5, j]
                if diffs ** 2 + diffey ** 2 - veh_length ** 2 - veh_width ** 2 < 0:
                    cost[index] += 100
        if old_direction_flag is not None and old_direction_flag != index:
            cost[index] += 100
    return cost


def solve_optimization_problem(self, solver=None):
    """Drop-in for OvertakeTrajPlanner.solve_optimization_problem (bind with types.MethodType or assign to the
    class).  `self` is the reference's planner object; returns the same 4-tuple (:246)."""
    import time as _time
    solver = solver or batch.solve_cbf_batch
    sorted_vehicles, obs_infos = self.sorted_vehicles, self.obs_infos
    N = self.racing_game_param.num_horizon_planner
    num_veh = len(sorted_vehicles)
    ego = self.vehicles[self.agent_name]
    veh_length, veh_width = ego.param.length, ego.param.width
    track = self.track
    C = num_veh + 1
    s_ref, ey_ref = np.zeros((C, N + 1)), np.zeros((C, N + 1))
    xlb, xub = np.zeros((C, N + 1, 2)), np.zeros((C, N + 1, 2))
    ok0 = np.zeros(C, dtype=bool)
    for c in range(C):
        xlb[c], xub[c] = candidate_bounds(c, self.xcurv_ego, sorted_vehicles, obs_infos, veh_length, veh_width, track.width,
                                          track.lap_length, N)
        s_ref[c], ey_ref[c] = candidate_targets(c, ego.xcurv, self.bezier_xcurvs, self.bezier_funcs, N)
        ok0[c] = x0_feasible(np.asarray(ego.xcurv, float), xlb[c], xub[c])
    prm = planner_params(self.racing_game_param.matrix_A, self.racing_game_param.matrix_B, N)
    kw, offset = pack_candidates(np.asarray(ego.xcurv, float), s_ref, ey_ref, xlb, xub, N)
    t0 = _time.perf_counter()
    res = solver(kw["x0"], kw["xt"], kw["obs"], kw["lap_off"], prm, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"])
    dt = _time.perf_counter() - t0
    solution_xvar = np.zeros((C, 6, N + 1))
    self.candidate_costs = np.full(C, np.inf)
    for c in range(C):
        if ok0[c] and res["status"][c] == 0:
            solution_xvar[c] = res["x"][c].T
            self.candidate_costs[c] = res["cost"][c] + offset[c]
        else:   # IPOPT failure path of the reference (:365-374)
            solution_xvar[c] = heuristic_traj(c, self.xcurv_ego, self.bezier_xcurvs, self.bezier_funcs, N)
    cost_selection = selection_costs(solution_xvar, sorted_vehicles, obs_infos, veh_length, veh_width, track.lap_length,
                                     self.old_direction_flag)
    direction_flag = cost_selection.index(min(cost_selection))
    traj_xcurv = solution_xvar[direction_flag, :, :].T
    solve_time = np.full(C, dt / C)
    return traj_xcurv, direction_flag, solve_time, solution_xvar


def plan_and_track(self, xcurv, mpc_lti_param, track, system_param, vehicles=None, agent_name=None, sorted_vehicles=None,
                   time=None, handle=None, region=None, extra=None):
    """The whole overtaking step in ONE call on one CUDA stream (b200mpc_plan_and_track, include/b200mpc.h): candidate
    QPs (overtake_traj_planner.py:248-379) -> selection cost + first argmin (:205-246) -> per-stage targets from the chosen
    trajectory (control.py:277, 373-382) -> tracking MPC-CBF (control.py:251-473), without a host round trip in between.
    The reference does this as solve_optimization_problem() followed by control.mpc_multi_agents() (utils/base.py:540-582).

    `self` is the reference's planner object after get_local_traj prepared it.  Returns
    ((traj_xcurv, direction_flag, solve_time, solution_xvar), (u0, x_pred)) -- the two reference return values.
    `extra`: optional dict of additional candidates {"s_ref","ey_ref","xlb","xub","region"} appended to the reference's
    num_veh+1 regions (BASELINE config 3 evaluates 64 candidates)."""
    import time as _time
    from . import control
    h = handle or batch.default_handle()
    sorted_vehicles = self.sorted_vehicles if sorted_vehicles is None else sorted_vehicles
    vehicles = self.vehicles if vehicles is None else vehicles
    agent_name = self.agent_name if agent_name is None else agent_name
    obs_infos = self.obs_infos
    N = self.racing_game_param.num_horizon_planner
    num_veh = len(self.sorted_vehicles)
    ego = vehicles[agent_name]
    veh_length, veh_width = ego.param.length, ego.param.width
    ego_x = np.asarray(ego.xcurv, float)
    Cn = num_veh + 1
    s_ref, ey_ref = np.zeros((Cn, N + 1)), np.zeros((Cn, N + 1))
    xlb, xub = np.zeros((Cn, N + 1, 2)), np.zeros((Cn, N + 1, 2))
    reg = np.arange(Cn, dtype=np.int32)
    for c in range(Cn):
        xlb[c], xub[c] = candidate_bounds(c, self.xcurv_ego, self.sorted_vehicles, obs_infos, veh_length, veh_width, track.width,
                                          track.lap_length, N)
        s_ref[c], ey_ref[c] = candidate_targets(c, ego_x, self.bezier_xcurvs, self.bezier_funcs, N)
    heur = np.stack([heuristic_traj(c, self.xcurv_ego, self.bezier_xcurvs, self.bezier_funcs, N).T for c in range(Cn)])
    if extra is not None:
        s_ref, ey_ref = np.concatenate([s_ref, extra["s_ref"]]), np.concatenate([ey_ref, extra["ey_ref"]])
        xlb, xub = np.concatenate([xlb, extra["xlb"]]), np.concatenate([xub, extra["xub"]])
        reg = np.concatenate([reg, np.asarray(extra["region"], dtype=np.int32)])
        heur = np.concatenate([heur, extra["heur"]])
    Cn = s_ref.shape[0]
    ok0 = np.array([x0_feasible(ego_x, xlb[c], xub[c]) for c in range(Cn)], dtype=np.int32)
    pprm = planner_params(self.racing_game_param.matrix_A, self.racing_game_param.matrix_B, N)
    kw, offset = pack_candidates(ego_x, s_ref, ey_ref, xlb, xub, N)
    cand, M0, ps0 = batch.pack_cbf(kw["x0"], kw["xt"], kw["obs"], None, N, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"])
    rivals = np.zeros((max(num_veh, 1), 2, N + 1))
    for j, name in enumerate(self.sorted_vehicles):
        rivals[j, 0], rivals[j, 1] = obs_infos[name][4, :N + 1], obs_infos[name][5, :N + 1]
    # tracking MPC record (control.mpc_multi_agents): everything except the per-stage targets, which the device fills
    Nc = mpc_lti_param.num_horizon_ctrl
    xc = np.asarray(xcurv, float).reshape(6)
    kept, num_cycle_ego = control._nearby_rivals(xc, sorted_vehicles, vehicles, agent_name, track.lap_length, time, 0.1, False, Nc + 1)
    obs_t, lap_off_t = control._rival_block(kept, num_cycle_ego, track.lap_length, Nc)
    tprm = control._limits(control._model(mpc_lti_param, Nc), system_param, track.width)
    tprm.update(alpha=0.6, margin=0.15, L=veh_length, W=veh_width)            # control.py:285,311,316-319
    trec, Mc, _ = batch.pack_cbf(xc.reshape(1, 6), np.zeros((1, Nc + 1, 6)), obs_t, lap_off_t, Nc)
    p_plan = _capi.make_cbf_params(pprm, 0, True, _capi.FLAG_STAGE_BOUNDS | _capi.FLAG_EY_RATE)
    p_track = _capi.make_cbf_params(tprm, Mc, True, 0)
    o = _capi.default_options()
    sel = _capi.PlannerSelectParams()
    sel.C, sel.N, sel.num_veh, sel.N_ctrl, sel.M_ctrl = Cn, N, num_veh, Nc, Mc
    sel.old_direction_flag = -1 if self.old_direction_flag is None else int(self.old_direction_flag)
    sel.veh_length, sel.veh_width, sel.lap_length = veh_length, veh_width, track.lap_length
    cand_rec = np.zeros(Cn, dtype=_capi.RECORD_DTYPE)
    cand_x = np.zeros((Cn, N + 1, 6))
    sel_cost = np.zeros(Cn)
    flag = np.zeros(2, dtype=np.int32)
    traj = np.zeros((N + 1, 6))
    trk = np.zeros(1, dtype=_capi.RECORD_DTYPE)
    trk_x, trk_u = np.zeros((Nc + 1, 6)), np.zeros((Nc, 2))
    P = batch._ptr
    heur = np.ascontiguousarray(heur)
    t0 = _time.perf_counter()
    rc = _capi.lib().b200mpc_plan_and_track(h.ptr, C.byref(p_plan), C.byref(p_track), C.byref(o), C.byref(sel), P(cand), P(heur),
                                            P(ok0), P(reg), P(rivals), P(trec), P(cand_rec), P(cand_x), P(sel_cost), P(flag), P(traj),
                                            P(trk), P(trk_x), P(trk_u))
    h.check(rc, "b200mpc_plan_and_track")
    dt = _time.perf_counter() - t0
    solved = (ok0 != 0) & (cand_rec["status"] == 0)
    solution_xvar = np.where(solved[:, None, None], cand_x, heur).transpose(0, 2, 1).copy()
    self.candidate_costs = np.where(solved, cand_rec["cost"] + offset, np.inf)
    self.selection_costs = sel_cost
    self.tracking_status = int(trk["status"][0])
    return (traj, int(flag[0]), np.full(Cn, dt / Cn), solution_xvar), (trk_u[0].copy(), trk_x)

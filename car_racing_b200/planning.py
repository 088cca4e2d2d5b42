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


def _tracking_record(xcurv, mpc_lti_param, track, system_param, vehicles, agent_name, sorted_vehicles, time, veh_length, veh_width):
    """Packed record of the tracking MPC (control.mpc_multi_agents, control.py:251-473): everything except the per-stage
    targets, which the selection kernel fills on the device."""
    from . import control
    Nc = mpc_lti_param.num_horizon_ctrl
    xc = np.asarray(xcurv, float).reshape(6)
    kept, num_cycle_ego = control._nearby_rivals(xc, sorted_vehicles, vehicles, agent_name, track.lap_length, time, 0.1, False, Nc + 1)
    obs_t, lap_off_t = control._rival_block(kept, num_cycle_ego, track.lap_length, Nc)
    tprm = control._limits(control._model(mpc_lti_param, Nc), system_param, track.width)
    tprm.update(alpha=0.6, margin=0.15, L=veh_length, W=veh_width)            # control.py:285,311,316-319
    trec, Mc, _ = batch.pack_cbf(xc.reshape(1, 6), np.zeros((1, Nc + 1, 6)), obs_t, lap_off_t, Nc)
    return trec, tprm, Nc, Mc


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
    trec, tprm, Nc, Mc = _tracking_record(xcurv, mpc_lti_param, track, system_param, vehicles, agent_name, sorted_vehicles, time,
                                          veh_length, veh_width)
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


# ---------------------------------------------------------------- candidate preparation on the device
def sort_rivals(eys):
    """Order of `sorted_vehicles` as indices into the iteration order of vehicles_interest
    (overtake_traj_planner.py:69-77): a rival goes to the front when its ey is >= the current first one's, to the back
    otherwise -- not a full sort for more than two rivals, reproduced as it is."""
    order = []
    for i, ey in enumerate(eys):
        if not order:
            order.append(i)
        elif ey >= eys[order[0]]:
            order.insert(0, i)
        elif ey <= eys[order[0]]:
            order.append(i)
    return order


def _prepare_params(N, num_veh, num_opt, prediction_factor, track_width, lap_length, veh_length, veh_width):
    p = _capi.PlannerPrepareParams()
    p.N, p.num_veh, p.num_opt = int(N), int(num_veh), int(num_opt)
    p.prediction_factor, p.track_width, p.lap_length = float(prediction_factor), float(track_width), float(lap_length)
    p.veh_length, p.veh_width, p.safety_margin, p.vx_max = float(veh_length), float(veh_width), SAFETY_MARGIN, VX_MAX_PLAN
    p.w_ey_rate, p.w_progress, p.w_track = W_EY_RATE, W_PROGRESS, W_TRACK
    return p


def _prepare_inputs(ego_x, xcurv_ego, obs_sorted, insertion, rival_vx, opt_traj, N):
    ego = np.concatenate([np.asarray(ego_x, float).reshape(6), np.asarray(xcurv_ego, float).reshape(6)])
    rivals = np.ascontiguousarray(obs_sorted, dtype=np.float64)
    if rivals.ndim != 3 or rivals.shape[1:] != (2, N + 1) or rivals.shape[0] < 1:
        raise ValueError("obs_sorted must be (num_veh >= 1, 2, N+1): rows s, ey of each rival's prediction")
    num_veh = rivals.shape[0]
    ins = np.ascontiguousarray(insertion, dtype=np.int32).reshape(num_veh)
    if sorted(ins.tolist()) != list(range(num_veh)):
        raise ValueError("insertion must be a permutation of range(num_veh)")
    vx = np.ascontiguousarray(rival_vx, dtype=np.float64).reshape(num_veh)
    opt = np.ascontiguousarray(opt_traj, dtype=np.float64)
    if opt.ndim != 2 or opt.shape[1] != 2 or opt.shape[0] < 2 or not (np.diff(opt[:, 0]) > 0).all():
        raise ValueError("opt_traj must be (T >= 2, 2): s (ascending), ey of the optimal trajectory")
    return ego, rivals, ins, vx, opt


def prepare_candidates(ego_x, xcurv_ego, obs_sorted, insertion, rival_vx, opt_traj, N, prediction_factor=0.5, track_width=1.0,
                       lap_length=None, veh_length=0.4, veh_width=0.2, handle=None):
    """What get_local_traj computes between the rivals' predictions and the candidate solves
    (overtake_traj_planner.py:87-117, planner_helper.py:46-153, 177-205) plus the data part of generate_traj_per_region
    (:276-334, 365-374), on the device (b200mpc_planner_prepare, include/b200mpc.h).

    obs_sorted (num_veh,2,N+1): s, ey predictions in sorted_vehicles order; insertion[i] = position in sorted_vehicles of
    the i-th rival of vehicles_interest; rival_vx (num_veh,) in sorted order; opt_traj (T,2) = s, ey of the optimal
    trajectory.  Returns the packed candidate records and everything the host path computes for them."""
    h = handle or batch.default_handle()
    ego, rivals, ins, vx, opt = _prepare_inputs(ego_x, xcurv_ego, obs_sorted, insertion, rival_vx, opt_traj, N)
    num_veh = rivals.shape[0]
    Cn = num_veh + 1
    p = _prepare_params(N, num_veh, opt.shape[0], prediction_factor, track_width, lap_length, veh_length, veh_width)
    stride = batch.cbf_record_doubles(N, 0, True, _capi.FLAG_STAGE_BOUNDS | _capi.FLAG_EY_RATE)
    cand = np.zeros((Cn, stride))
    heur = np.zeros((Cn, N + 1, 6))
    ok0, region = np.zeros(Cn, dtype=np.int32), np.zeros(Cn, dtype=np.int32)
    offset, ctrl, bez = np.zeros(Cn), np.zeros((Cn, 4, 2)), np.zeros((Cn, N + 1, 2))
    err = np.zeros(1, dtype=np.int32)
    P = batch._ptr
    rc = _capi.lib().b200mpc_planner_prepare(h.ptr, C.byref(p), P(ego), P(rivals), P(vx), P(ins), P(opt), P(cand), P(heur), P(ok0),
                                             P(region), P(offset), P(ctrl), P(bez), P(err))
    h.check(rc, "b200mpc_planner_prepare")
    if err[0]:   # interp1d's bounds_error in get_bezier_control_points (planner_helper.py:57, 92-135)
        raise ValueError("A value in x_new is outside the interpolation range.")
    return dict(records=cand, heur=heur, ok0=ok0, region=region, offset=offset, ctrl=ctrl, bezier=bez)


def _from_predictions(self, xcurv_ego, time, vehicles_interest, old_direction_flag, handle, extra, tracking):
    """prediction -> preparation -> candidate solve -> selection [-> tracking solve] through
    b200mpc_plan_and_track_prepared; `tracking` = (xcurv, mpc_lti_param, track, system_param) or None."""
    import time as _time
    from scipy.interpolate import interp1d
    h = handle or batch.default_handle()
    prm = self.racing_game_param
    N = prm.num_horizon_planner
    vehicles, agent_name, track = self.vehicles, self.agent_name, self.track
    ego = vehicles[agent_name]
    veh_length, veh_width = ego.param.length, ego.param.width
    names = list(vehicles_interest)
    order = sort_rivals([vehicles_interest[n].xcurv[5] for n in names])
    sorted_vehicles = [names[i] for i in order]
    obs_infos = {}
    for name in names:                                                        # :78-86
        if vehicles[name].no_dynamics:
            obs_traj, _ = vehicles[name].get_trajectory_nsteps(time, prm.timestep, N + 1)
        else:
            obs_traj, _ = vehicles[name].get_trajectory_nsteps(N + 1)
        obs_infos[name] = obs_traj
    obs_sorted = np.array([obs_infos[n][4:6, :N + 1] for n in sorted_vehicles])
    insertion = [sorted_vehicles.index(n) for n in names]
    rival_vx = [vehicles[n].xcurv[0] for n in sorted_vehicles]
    opt_traj = np.asarray(self.opti_traj_xcurv, float)[:, 4:6]
    ego_x = np.asarray(ego.xcurv, float)
    egov, rivals, ins, vx, opt = _prepare_inputs(ego_x, xcurv_ego, obs_sorted, insertion, rival_vx, opt_traj, N)
    num_veh = len(names)
    C0 = num_veh + 1
    n_extra = 0 if extra is None else int(np.asarray(extra["region"]).shape[0])
    Cn = C0 + n_extra
    pp = _prepare_params(N, num_veh, opt.shape[0], prm.planning_prediction_factor, track.width, track.lap_length, veh_length, veh_width)
    self.sorted_vehicles, self.obs_infos, self.xcurv_ego, self.old_direction_flag = sorted_vehicles, obs_infos, xcurv_ego, old_direction_flag
    sel = _capi.PlannerSelectParams()
    sel.C, sel.N, sel.num_veh = Cn, N, num_veh
    sel.old_direction_flag = -1 if old_direction_flag is None else int(old_direction_flag)
    sel.veh_length, sel.veh_width, sel.lap_length = veh_length, veh_width, track.lap_length
    trec = trk = trk_x = trk_u = p_track_ref = None
    if tracking is not None:
        xcurv, mpc_lti_param, trk_track, system_param = tracking
        # control.mpc_multi_agents is called without `time` (utils/base.py:558-572)
        trec, tprm, Nc, Mc = _tracking_record(xcurv, mpc_lti_param, trk_track, system_param, vehicles, agent_name, sorted_vehicles, None,
                                              veh_length, veh_width)
        sel.N_ctrl, sel.M_ctrl = Nc, Mc
        p_track = _capi.make_cbf_params(tprm, Mc, True, 0)
        p_track_ref = C.byref(p_track)
        trk = np.zeros(1, dtype=_capi.RECORD_DTYPE)
        trk_x, trk_u = np.zeros((Nc + 1, 6)), np.zeros((Nc, 2))
    pprm = planner_params(prm.matrix_A, prm.matrix_B, N)
    p_plan = _capi.make_cbf_params(pprm, 0, True, _capi.FLAG_STAGE_BOUNDS | _capi.FLAG_EY_RATE)
    o = _capi.default_options()
    x_cand = x_heur = x_ok0 = x_reg = None
    x_off = np.zeros(0)
    if n_extra:
        kw, x_off = pack_candidates(ego_x, extra["s_ref"], extra["ey_ref"], extra["xlb"], extra["xub"], N)
        x_cand, _, _ = batch.pack_cbf(kw["x0"], kw["xt"], kw["obs"], None, N, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"])
        x_heur = np.ascontiguousarray(extra["heur"], dtype=np.float64)
        x_ok0 = np.array([x0_feasible(ego_x, kw["xlb"][c], kw["xub"][c]) for c in range(n_extra)], dtype=np.int32)
        x_reg = np.ascontiguousarray(extra["region"], dtype=np.int32)
    cand_rec = np.zeros(Cn, dtype=_capi.RECORD_DTYPE)
    cand_x, heur = np.zeros((Cn, N + 1, 6)), np.zeros((Cn, N + 1, 6))
    sel_cost, ok0 = np.zeros(Cn), np.zeros(Cn, dtype=np.int32)
    flag = np.zeros(2, dtype=np.int32)
    traj = np.zeros((N + 1, 6))
    off0, bez, err = np.zeros(C0), np.zeros((C0, N + 1, 2)), np.zeros(1, dtype=np.int32)
    P = batch._ptr
    t0 = _time.perf_counter()
    rc = _capi.lib().b200mpc_plan_and_track_prepared(
        h.ptr, C.byref(p_plan), p_track_ref, C.byref(o), C.byref(sel), C.byref(pp), P(egov), P(rivals), P(vx), P(ins), P(opt),
        n_extra, P(x_cand), P(x_heur), P(x_ok0), P(x_reg), P(trec), P(cand_rec), P(cand_x), P(sel_cost), P(flag), P(traj), P(trk),
        P(trk_x), P(trk_u), P(heur), P(ok0), P(off0), P(bez), P(err))
    h.check(rc, "b200mpc_plan_and_track_prepared")
    dt = _time.perf_counter() - t0
    if err[0]:   # interp1d's bounds_error in get_bezier_control_points (planner_helper.py:57, 92-135)
        raise ValueError("A value in x_new is outside the interpolation range.")
    self.bezier_xcurvs = bez
    self.bezier_funcs = [interp1d(bez[c, :, 0], bez[c, :, 1]) for c in range(C0)]
    solved = (ok0 != 0) & (cand_rec["status"] == 0)
    solution_xvar = np.where(solved[:, None, None], cand_x, heur).transpose(0, 2, 1).copy()
    self.candidate_costs = np.where(solved, cand_rec["cost"] + np.concatenate([off0, x_off]), np.inf)
    self.selection_costs = sel_cost
    plan = (traj, int(flag[0]), np.full(Cn, dt / Cn), solution_xvar)
    if tracking is None:
        return plan, None
    self.tracking_status = int(trk["status"][0])
    return plan, (trk_u[0].copy(), trk_x)


def plan_and_track_from_predictions(self, xcurv_ego, time, vehicles_interest, xcurv, mpc_lti_param, track, system_param,
                                    old_direction_flag=None, handle=None, extra=None):
    """The overtaking step from the rivals' predictions on: get_local_traj (overtake_traj_planner.py:44-161, without the
    global-frame copies it makes for plotting) followed by control.mpc_multi_agents (utils/base.py:540-582) as ONE call on one
    CUDA stream (b200mpc_plan_and_track_prepared): preparation -> candidate QPs -> selection -> tracking MPC-CBF.  The host
    only predicts the rivals (their own get_trajectory_nsteps), orders them (:69-77) and packs the tracking record.

    `self` is the reference's planner object (vehicles, agent_name, track, opti_traj_xcurv, racing_game_param set).  Leaves
    sorted_vehicles / obs_infos / bezier_xcurvs / bezier_funcs / xcurv_ego / old_direction_flag on `self` as get_local_traj
    does.  Returns ((traj_xcurv, direction_flag, solve_time, solution_xvar), (u0, x_pred)).
    `extra`: optional additional candidates as in plan_and_track."""
    return _from_predictions(self, xcurv_ego, time, vehicles_interest, old_direction_flag, handle, extra,
                             (xcurv, mpc_lti_param, track, system_param))


def _traj_xglob(trajs_xcurv, track):
    """planner_helper.get_traj_xglob (planning/planner_helper.py:208-220) for a stack of trajectories (..., 6): only columns
    4, 5 (global x, y) are filled.  One launch for all points (b200mpc_curv_to_glob) when the track carries the reference's
    point_and_tangent table; a duck-typed track without it is asked point by point, as the reference does."""
    trajs_xcurv = np.asarray(trajs_xcurv, float)
    out = np.zeros(trajs_xcurv.shape)
    if hasattr(track, "point_and_tangent"):
        x, y, _ = batch.curv_to_glob_batch(trajs_xcurv[..., 4], trajs_xcurv[..., 5], track.point_and_tangent, track.lap_length)
        out[..., 4], out[..., 5] = x, y
        return out
    flat_in, flat_out = trajs_xcurv.reshape(-1, 6), out.reshape(-1, 6)
    for i in range(flat_in.shape[0]):
        s_i = float(flat_in[i, 4])
        while s_i > track.lap_length:
            s_i = s_i - track.lap_length
        flat_out[i, 4], flat_out[i, 5] = track.get_global_position(s_i, flat_in[i, 5])
    return out


def get_local_traj(self, xcurv_ego, time, vehicles_interest, matrix_Atv, matrix_Btv, matrix_Ctv, old_ey, old_direction_flag):
    """Drop-in for OvertakeTrajPlanner.get_local_traj (overtake_traj_planner.py:44-161; assign to the class): same arguments,
    same 8-tuple.  Rival predictions and ordering on the host, then ONE call (b200mpc_plan_and_track_prepared without the
    tracking stage): Bezier references, candidate records, candidate QPs, selection on the device; the global-frame copies
    for plotting are one more launch (b200mpc_curv_to_glob) instead of one get_global_position call per point."""
    self.matrix_Atv, self.matrix_Btv, self.matrix_Ctv = matrix_Atv, matrix_Btv, matrix_Ctv
    self.old_ey = old_ey
    (traj, flag, solve_time, sol), _ = _from_predictions(self, xcurv_ego, time, vehicles_interest, old_direction_flag, None, None, None)
    N = self.racing_game_param.num_horizon_planner
    C0 = sol.shape[0]
    track = self.track
    # the reference converts target, chosen curve, all curves and all candidates point by point (:124-147); here one launch
    lines = np.zeros((C0, N + 1, 6))
    lines[:, :, 4:6] = self.bezier_xcurvs
    stack = np.concatenate([traj[None], lines, sol.transpose(0, 2, 1)])          # (1 + 2 C0, N+1, 6)
    glob = _traj_xglob(stack, track)
    target_traj_xglob, all_bezier_xglob, all_local_traj_xglob = glob[0], glob[1:1 + C0], glob[1 + C0:]
    bezier_xglob = all_bezier_xglob[flag].copy()
    return traj, target_traj_xglob, flag, self.sorted_vehicles, bezier_xglob, solve_time, all_bezier_xglob, all_local_traj_xglob

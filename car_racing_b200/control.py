"""Drop-in replacements for the reference's solve functions (car_racing/control/control.py).

Same names, positional signatures, return shapes and error behaviour as the reference, so
`car_racing/utils/base.py` -- which calls `control.mpc_lti / mpccbf / mpc_multi_agents / ilqr`
through the module object (base.py:199,256,307,558) -- runs unchanged after

    import car_racing_b200
    car_racing_b200.install()          # patches the reference's control.control module in place

Each shim does the reference's host-side packing in numpy (proximity filter, lap offsets,
per-stage targets) and calls the C-ABI with a batch of one.  The solve itself always runs on
the GPU: there is no CasADi/IPOPT and no CPU fallback here.
"""
import warnings

import numpy as np
from scipy.interpolate import interp1d

from . import _capi, batch

X_DIM, U_DIM = 6, 2
_SLACK_W = 10000.0          # control.py:560
_DEGREE = 6                 # control.py:528 (hard-coded in the reference, and in the kernel)


def _model(param, N):
    return dict(A=np.asarray(param.matrix_A, float), B=np.asarray(param.matrix_B, float),
                Q=np.asarray(param.matrix_Q, float), R=np.asarray(param.matrix_R, float), N=int(N))


def _limits(prm, system_param, width):
    prm.update(umax=[system_param.delta_max, system_param.a_max], vmin=system_param.v_min, vmax=system_param.v_max,
               width=width, slack_w=_SLACK_W)
    return prm


def _nearby_rivals(xcurv, names, vehicles, agent_name, lap_length, time, timestep, realtime_flag, n_pred):
    """Proximity filter of control.py:499-523 / :284-309.  Returns [(name, obs_traj (6,n_pred))]."""
    margin = xcurv[0] * 2.0                      # safety_time = 2.0 (control.py:499-501)
    num_cycle_ego = int(xcurv[4] / lap_length)
    dist_ego = xcurv[4] - num_cycle_ego * lap_length
    kept = []
    for name in names:
        if name == agent_name:
            continue
        if realtime_flag == False:  # noqa: E712  (the reference compares with ==, None falls through)
            obs_traj, _ = vehicles[name].get_trajectory_nsteps(time, timestep, n_pred)
        elif realtime_flag == True:  # noqa: E712
            obs_traj, _ = vehicles[name].get_trajectory_nsteps(n_pred)
        else:
            continue
        num_cycle_obs = int(obs_traj[4, 0] / lap_length)
        dist_obs = obs_traj[4, 0] - num_cycle_obs * lap_length
        if (dist_ego > dist_obs - margin) & (dist_ego < dist_obs + margin):
            kept.append((name, np.asarray(obs_traj, float)))
    return kept, num_cycle_ego


def _rival_block(kept, num_cycle_ego, lap_length, N):
    M = len(kept)
    obs = np.zeros((1, M, 2, N + 1))
    lap_off = np.zeros((1, M))
    for j, (_, traj) in enumerate(kept):
        obs[0, j, 0] = traj[4, :N + 1]
        obs[0, j, 1] = traj[5, :N + 1]
        # control.py:538-540: the offset enters h (diffs) but not h_next (diffs_next)
        lap_off[0, j] = (num_cycle_ego - int(traj[4, 0] / lap_length)) * lap_length
    return obs, lap_off


def _nearest(kept, xcurv, lap_length):
    """The kernel takes at most _capi.MMAX rivals per solve (one warp lane per column of the stage Hessian).  The reference
    has no limit; with more rivals inside the +-2*vx window the MMAX nearest along s are kept (a CBF row only binds near its
    rival) and a warning says so."""
    if len(kept) <= _capi.MMAX:
        return kept
    dist_ego = xcurv[4] - int(xcurv[4] / lap_length) * lap_length
    gap = [abs((t[4, 0] - int(t[4, 0] / lap_length) * lap_length) - dist_ego) for _, t in kept]
    order = sorted(np.argsort(gap, kind="stable")[:_capi.MMAX])          # keep the reference's rival order
    warnings.warn("b200mpc: %d rivals inside the proximity window, the %d nearest are used" % (len(kept), _capi.MMAX))
    return [kept[i] for i in order]


def _sizes(pairs):
    """(L, W, sizes): one (l_agent+l_obs, w_agent+w_obs) pair for all rivals -> kernel constants; different rivals
    (control.py:530-535 reads every rival's own length / width) -> the per-rival block of the record."""
    if not pairs:
        return 0.4, 0.2, None
    if len({(round(p[0], 12), round(p[1], 12)) for p in pairs}) == 1:
        return pairs[0][0], pairs[0][1], None
    return pairs[0][0], pairs[0][1], np.asarray(pairs, float)


# what the last single-instance solve did (status, iterations, retries, elastic_max, ...): the drop-ins return only u, as the
# reference does; this is where a caller can look
last_solve = {}
ELASTIC_TOL = 1e-6      # elastic_max above this: a CBF row is not met exactly, the point is the rho-penalised one
IPOPT_MAX_ITER = 3000   # IPOPT's default max_iter (the reference sets no solver option but printing, control.py:593)


def _solve_one(xcurv, xt, obs, lap_off, prm, sizes=None):
    """One instance through the batch API, with the second chances a single control step can afford (a batch cannot: its
    stragglers set the step time, so the kernel's defaults are max_iter 200 and a fixed elastic weight):
      * MAX_ITER at 200 iterations -> once more with IPOPT's own limit of 3000;
      * converged with elastic_max > ELASTIC_TOL (the l1-elastic CBF row undercut the price of the reference's own slack,
        DESIGN.md section 2 deviation 2) -> once more with rho x 100 and 3000 iterations; that result is used if it converges
        with the rows met, otherwise the first one stands and a warning is issued.
    Returns the batch result dict; `last_solve` records what happened."""
    M = obs.shape[1]
    r = batch.solve_cbf_batch(xcurv.reshape(1, 6), xt, obs, lap_off, prm, sizes=sizes)
    info = dict(M=M, retries=[], status=int(r["status"][0]), iters=int(r["iters"][0]))
    if r["status"][0] == 1:
        r2 = batch.solve_cbf_batch(xcurv.reshape(1, 6), xt, obs, lap_off, prm, sizes=sizes, max_iter=IPOPT_MAX_ITER)
        info["retries"].append(("max_iter=%d" % IPOPT_MAX_ITER, int(r2["status"][0]), int(r2["iters"][0])))
        r = r2
    if M > 0 and r["status"][0] == 0 and r["elastic_max"][0] > ELASTIC_TOL:
        rho = _capi.default_options().rho * 100.0
        r2 = batch.solve_cbf_batch(xcurv.reshape(1, 6), xt, obs, lap_off, prm, sizes=sizes, rho=rho, max_iter=IPOPT_MAX_ITER)
        info["retries"].append(("rho=%g" % rho, int(r2["status"][0]), int(r2["iters"][0])))
        if r2["status"][0] == 0 and r2["elastic_max"][0] <= ELASTIC_TOL:
            r = r2
        else:
            warnings.warn("b200mpc: CBF rows met only elastically (elastic_max %.3g): the rival cannot be cleared within the "
                          "horizon; the returned input is the penalised solution" % r["elastic_max"][0])
    info.update(status=int(r["status"][0]), iters=int(r["iters"][0]), cost=float(r["cost"][0]),
                elastic_max=float(r["elastic_max"][0]), kkt_err=float(r["kkt_err"][0]),
                status_text=_capi.STATUS_NAMES.get(int(r["status"][0]), "?"))
    last_solve.clear()
    last_solve.update(info)
    return r


def pid(xcurv, xtarget):
    """control.py:15-25 (two scalar gains; stays on the host)."""
    xtarget = np.asarray(xtarget, float).reshape(-1)
    u = np.zeros(U_DIM)
    u[0] = -0.6 * (xcurv[5] - xtarget[5]) - 0.9 * xcurv[3]
    u[1] = 1.5 * (xtarget[0] - xcurv[0])
    return u


def mpc_lti(xcurv, xtarget, mpc_lti_param, system_param, track):
    """control.py:198-248.  Raises RuntimeError on non-convergence, as the uncaught IPOPT failure does (:242)."""
    N = mpc_lti_param.num_horizon
    prm = _limits(_model(mpc_lti_param, N), system_param, track.width)
    prm.update(alpha=0.8, margin=0.2, L=0.4, W=0.2)
    xt = np.asarray(xtarget, float).reshape(X_DIM)
    r = _solve_one(np.asarray(xcurv, float).reshape(X_DIM), xt, np.zeros((1, 0, 2, N + 1)), None, prm)
    if r["status"][0] != 0:      # incl. status 4: x0 outside the v / ey rows of stage 0 -- IPOPT reports infeasibility there
        raise RuntimeError("b200mpc: mpc_lti failed: %s" % last_solve["status_text"])
    return r["u"][0, 0, :]


def mpccbf(xcurv, xtarget, mpc_cbf_param, vehicles, agent_name, lap_length, time, timestep, realtime_flag, track,
           system_param, return_details=False):
    """control.py:476-607.  Returns u_pred[0,:]; never raises on non-convergence (:600-603)."""
    N = mpc_cbf_param.num_horizon
    xcurv = np.asarray(xcurv, float).reshape(X_DIM)
    kept, num_cycle_ego = _nearby_rivals(xcurv, list(vehicles), vehicles, agent_name, lap_length, time, timestep,
                                         realtime_flag, N + 1)
    kept = _nearest(kept, xcurv, lap_length)
    obs, lap_off = _rival_block(kept, num_cycle_ego, lap_length, N)
    ego = vehicles[agent_name].param
    L, W, sizes = _sizes([(ego.length / 2 + vehicles[n].param.length / 2, ego.width / 2 + vehicles[n].param.width / 2)
                          for n, _ in kept])
    prm = _limits(_model(mpc_cbf_param, N), system_param, track.width)
    prm.update(alpha=mpc_cbf_param.alpha, margin=0.2, L=L, W=W)
    xt = np.asarray(xtarget, float).reshape(X_DIM)
    r = _solve_one(xcurv, xt, obs, lap_off, prm, sizes)
    if r["status"][0] != 0:
        print("solver failed.")                   # control.py:601; the iterate is used, as opti.debug.value is there
    if return_details:
        return r["u"][0, 0, :], r
    return r["u"][0, 0, :]


def mpc_multi_agents(xcurv, mpc_lti_param, track, matrix_Atv, matrix_Btv, matrix_Ctv, system_param,
                     target_traj_xcurv=None, vehicles=None, agent_name=None, direction_flag=None,
                     target_traj_xglob=None, sorted_vehicles=None, time=None):
    """control.py:251-473 (the CBF_Flag=True path; lines 383-445 are dead in the reference).
    Returns (u_pred[0,:], x_pred (N+1,6))."""
    N = mpc_lti_param.num_horizon_ctrl
    xcurv = np.asarray(xcurv, float).reshape(X_DIM)
    vx = xcurv[0]
    f_traj = interp1d(target_traj_xcurv[:, 4], target_traj_xcurv[:, 5])
    veh_len, veh_width = vehicles["ego"].param.length, vehicles["ego"].param.width
    kept, num_cycle_ego = _nearby_rivals(xcurv, sorted_vehicles, vehicles, agent_name, track.lap_length, time, 0.1,
                                         False, N + 1)
    kept = _nearest(kept, xcurv, track.lap_length)
    obs, lap_off = _rival_block(kept, num_cycle_ego, track.lap_length, N)
    xt = np.zeros((1, N + 1, 6))
    for i in range(N + 1):                        # control.py:373-378
        s_tmp = vx * 0.1 * i + xcurv[4]
        s_tmp = max(s_tmp, target_traj_xcurv[0, 4])
        if s_tmp >= target_traj_xcurv[-1, 4]:
            s_tmp = target_traj_xcurv[-1, 4]
        xt[0, i] = [vx, 0, 0, 0, 0, float(f_traj(s_tmp))]
    prm = _limits(_model(mpc_lti_param, N), system_param, track.width)
    prm.update(alpha=0.6, margin=0.15, L=veh_len, W=veh_width)   # control.py:285,311,316-319: the EGO's full length / width
    r = _solve_one(xcurv, xt, obs, lap_off, prm)
    if r["status"][0] != 0:
        print("solver fail")                      # control.py:459; the iterate is used, as opti.debug.value is there
    return r["u"][0, 0, :], r["x"][0]


def ilqr(xcurv, xtarget, ilqr_param, vehicles, agent_name, lap_length, time, timestep, track, system_param):
    """control.py:64-195.  Only the last non-ego vehicle's prediction is used (:100-105) and the rival
    size is read from vehicles["car1"] (:109-110), as in the reference."""
    N = ilqr_param.num_horizon
    xcurv = np.asarray(xcurv, float).reshape(X_DIM)
    obs_traj = None
    for name in list(vehicles):
        if name != agent_name:
            obs_traj, _ = vehicles[name].get_trajectory_nsteps(time, timestep, N + 1)
    if obs_traj is None:
        raise NameError("obs_traj")               # the reference fails the same way without a rival
    l_sum = vehicles[agent_name].param.length / 2 + vehicles["car1"].param.length / 2
    w_sum = vehicles[agent_name].param.width / 2 + vehicles["car1"].param.width / 2
    num_cycle_ego = int(xcurv[4] / lap_length)
    lap_off = (num_cycle_ego - int(obs_traj[4, 0] / lap_length)) * lap_length
    prm = _model(ilqr_param, N)
    prm.update(max_iter=ilqr_param.max_iter, L=l_sum, W=w_sum)
    r = batch.solve_ilqr_batch(xcurv.reshape(1, 6), np.asarray(xtarget, float).reshape(6),
                               np.asarray(obs_traj, float)[4:6, :N + 1].reshape(1, 2, N + 1), [lap_off], prm, want=("u",))
    return r["u"][0, 0, :]


def lmpc(xcurv, lmpc_param, matrix_Atv, matrix_Btv, matrix_Ctv, ss_curv, Qfun, iter, lap_length, lap_width, u_old,
         system_param):
    """control.lmpc (control.py:610-730): select the safe-set points (:625-638, lmpc_helper.select_points), solve the
    QP on the GPU, return the reference's 6-tuple (:723-730)."""
    from .scenarios import select_points
    cols, qs = [], []
    for jj in range(0, lmpc_param.num_ss_iter):
        pts, q = select_points(ss_curv, Qfun, iter - jj - 1, xcurv, lmpc_param.num_ss_points / lmpc_param.num_ss_iter,
                               lmpc_param.shift)
        cols.append(pts)
        qs.append(q)
    ss_point_selected_tot = np.concatenate(cols, axis=1)
    Qfun_selected_tot = np.concatenate(qs, axis=0)
    N = int(lmpc_param.num_horizon)
    prm = dict(Q=np.asarray(lmpc_param.matrix_Q, float), R=np.asarray(lmpc_param.matrix_R, float),
               dR=np.asarray(lmpc_param.matrix_dR, float), N=N, umax=[system_param.delta_max, system_param.a_max],
               vmax=system_param.v_max, width=lap_width, xtrk=np.array([5.0, 0, 0, 0, 0, 0]))
    A = np.asarray([np.asarray(matrix_Atv[i], float) for i in range(N)])
    B = np.asarray([np.asarray(matrix_Btv[i], float) for i in range(N)])
    Cm = np.asarray([np.asarray(matrix_Ctv[i], float).reshape(6) for i in range(N)])
    res = batch.solve_lmpc_batch(np.asarray(xcurv, float)[None], np.asarray(u_old, float).reshape(1, 2), A[None], B[None], Cm[None],
                                 ss_point_selected_tot[None], Qfun_selected_tot[None], prm, want=("x", "u"))
    if res["status"][0] != 0:
        print("solver fail to find the solution, the non-converged solution is used")   # control.py:718
    x_pred, u_pred = res["x"][0], res["u"][0]
    lin_points = np.concatenate((x_pred[1:, :], np.array([x_pred[-1, :]])), axis=0)
    lin_input = np.vstack((u_pred[1:, :], u_pred[-1, :]))
    return u_pred, x_pred, ss_point_selected_tot, Qfun_selected_tot, lin_points, lin_input


def estimate_ABC(self):
    """Drop-in for LMPCRacingGame.estimate_ABC (utils/base.py:585-622; bind with types.MethodType or assign to the class):
    the N sequential regression_and_linearization calls become one launch.  Returns the reference's 4-tuple."""
    N = self.lmpc_param.num_horizon
    used_iter = range(self.iter - 2, self.iter)                      # lap_used_for_linearization = 2 (base.py:600-601)
    r = batch.estimate_abc_batch(self.lin_points, self.lin_input, self.ss_xcurv, self.u_ss, self.time_ss, used_iter,
                                 self.point_and_tangent, self.timestep, max_num_point=40)
    Atv = [r["A"][0, i].copy() for i in range(N)]
    Btv = [r["B"][0, i].copy() for i in range(N)]
    Ctv = [r["C"][0, i].reshape(X_DIM, 1).copy() for i in range(N)]
    index_used_list = [[row[row >= 0].astype(np.int64) for row in r["idx"][0, i]] for i in range(N)]
    return Atv, Btv, Ctv, index_used_list


def install(control_module=None):
    """Swap the reference's solve functions for the GPU ones (module-level monkey patch).
    `control_module` defaults to the reference's `control.control` if it is importable."""
    if control_module is None:
        from control import control as control_module  # the reference's own package layout (setup.cfg:16-17)
    for name in ("mpc_lti", "mpccbf", "mpc_multi_agents", "ilqr", "lmpc"):
        setattr(control_module, name, globals()[name])
    return control_module


def install_all(control_module=None, base_module=None, planner_class=None, offboard_module=None):
    """Everything on the control-step path: the five solve functions, LMPCRacingGame.estimate_ABC, the planner's
    solve_optimization_problem and get_local_traj (candidate preparation on the device) and the rivals' trajectory
    prediction (sympy-free for NoDynamicsModel, on the device for offboard.DynamicBicycleModel)."""
    from . import planning, rivals
    control_module = install(control_module)
    if base_module is None:
        from utils import base as base_module
    base_module.LMPCRacingGame.estimate_ABC = estimate_ABC
    if offboard_module is None:
        try:
            from racing import offboard as offboard_module     # rivals with dynamics (racing/offboard.py:80-94)
        except ImportError:
            offboard_module = None
    rivals.install(base_module, offboard_module)
    if planner_class is None:
        from planning.overtake_traj_planner import OvertakeTrajPlanner as planner_class
    planner_class.solve_optimization_problem = planning.solve_optimization_problem
    planner_class.get_local_traj = planning.get_local_traj
    return control_module, base_module, planner_class

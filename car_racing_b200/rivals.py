"""Rival trajectory prediction: drop-in for `NoDynamicsModel.get_trajectory_nsteps`
(car_racing/utils/base.py:879-886).

The reference evaluates the rival's sympy expressions s(t), ey(t) with four `subs`/`diff` calls per predicted step
(`get_estimation`, base.py:860-877) -- 8.2 ms per rival and control step, i.e. several times the GPU
solve it feeds.  Here the expressions and their time derivatives are compiled once per rival (`sympy.lambdify`) and
evaluated for all n steps in one numpy call.  Rivals WITH dynamics (`offboard.DynamicBicycleModel.get_trajectory_nsteps`,
car_racing/racing/offboard.py:80-94: a zero-input Frenet rollout) are predicted on the device, see
`dynamic_get_trajectory_nsteps` below.  Every caller in the reference discards the second return value
(control.py:103,296,509; overtake_traj_planner.py:78), so the global-frame block is only filled on request."""
import numpy as np

X_DIM = 6
_COMPILED = {}      # (t, s(t), ey(t)) -> compiled functions.  Kept here, NOT on the rival: the reference pickles the simulator with its vehicles
                    # after every run (car_racing/tests/mpccbf_test.py:45-46, tests/auto_mpccbf_test.py:42-43) and lambdified functions do not pickle


def _compiled(model):
    import sympy as sp
    key = (model.t_symbol, model.s_func, model.ey_func)          # sympy expressions hash and compare structurally
    fns = _COMPILED.get(key)
    if fns is None:
        t = model.t_symbol
        exprs = [sp.diff(model.s_func, t), sp.diff(model.ey_func, t), model.s_func, model.ey_func]
        fns = [sp.lambdify(t, e, "numpy") for e in exprs]
        if len(_COMPILED) >= 256:
            _COMPILED.clear()
        _COMPILED[key] = fns
    return fns


def get_trajectory_nsteps(self, t0, delta_t, n, with_glob=False):
    """Same first return value as the reference: (6, n) with rows (ds/dt, dey/dt, 0, 0, s, ey) at self.time + k*delta_t
    (the reference ignores t0 and uses self.time, base.py:883).  Second value: zeros unless with_glob."""
    fns = _compiled(self)
    tt = self.time + np.arange(n) * delta_t
    xcurv = np.zeros((X_DIM, n))
    for row, fn in zip((0, 1, 4, 5), fns):
        xcurv[row, :] = np.broadcast_to(np.asarray(fn(tt), dtype=float), (n,))
    xglob = np.zeros((X_DIM, n))
    if with_glob:
        xglob[0:3, :] = xcurv[0:3, :]
        for k in range(n):
            X, Y = self.track.get_global_position(xcurv[4, k], xcurv[5, k])
            xglob[3, k] = self.track.get_orientation(xcurv[4, k], xcurv[5, k])
            xglob[4, k], xglob[5, k] = X, Y
    return xcurv, xglob


def dynamic_get_trajectory_nsteps(self, n):
    """Drop-in for offboard.DynamicBicycleModel.get_trajectory_nsteps (car_racing/racing/offboard.py:80-94): the zero-input
    Frenet rollout of a rival with dynamics, n steps in one kernel launch (b200mpc_rival_rollout).  Same return value:
    (xcurv_nsteps (6,n), xglob_nsteps (6,n)), incl. the reference's global-frame quirks (offboard.py:71-76).  Many rivals
    or scenarios at once: car_racing_b200.batch.rival_rollout_batch."""
    from . import batch
    xc, xg = batch.rival_rollout_batch(self.xcurv, self.xglob, self.point_and_tangent, self.lap_length, self.timestep, n,
                                       with_glob=True)
    return xc[0], xg[0]


def install(base_module=None, offboard_module=None):
    """Patch the reference's NoDynamicsModel (and, when its module is given, offboard.DynamicBicycleModel) in place."""
    if base_module is None:
        from utils import base as base_module
    base_module.NoDynamicsModel.get_trajectory_nsteps = get_trajectory_nsteps
    if offboard_module is not None:
        offboard_module.DynamicBicycleModel.get_trajectory_nsteps = dynamic_get_trajectory_nsteps
    return base_module

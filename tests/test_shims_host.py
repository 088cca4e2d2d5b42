"""CPU tests of the drop-in shims' host logic (proximity filter, lap offsets, per-stage targets):
the C-ABI call is intercepted, no GPU needed.  Scenario mirrors car_racing/tests/mpccbf_test.py:27-31."""
import types

import numpy as np
import pytest

from car_racing_b200 import control, scenarios


class Rival:
    def __init__(self, s0, v, ey, length=0.4, width=0.2):
        self.param = types.SimpleNamespace(length=length, width=width)
        self.s0, self.v, self.ey = s0, v, ey
        self.xcurv = np.array([v, 0, 0, 0, s0, ey])

    def get_trajectory_nsteps(self, t0, dt, n):
        tr = np.zeros((6, n))
        tr[4] = self.s0 + self.v * dt * np.arange(n)
        tr[5] = self.ey
        return tr, None


def _capture(monkeypatch):
    seen = {}

    def fake(x0, xt, obs, lap_off, prm, want=("aux", "x", "u", "sigma"), handle=None, **opt):
        seen.update(x0=np.array(x0), xt=np.array(xt), obs=np.array(obs), lap_off=lap_off, prm=prm, opt=opt)
        seen.setdefault("calls", []).append(dict(opt))
        N = prm["N"]
        k = len(seen["calls"]) - 1
        st = seen.get("statuses", [0])
        el = seen.get("elastic", [0.0])
        return dict(u=np.zeros((1, N, 2)) + 0.25, x=np.zeros((1, N + 1, 6)), status=np.array([st[min(k, len(st) - 1)]]),
                    u0=np.zeros((1, 2)), iters=np.array([7]), cost=np.zeros(1), kkt_err=np.zeros(1),
                    elastic_max=np.array([el[min(k, len(el) - 1)]]))
    monkeypatch.setattr(control.batch, "solve_cbf_batch", fake)
    return seen


def test_mpccbf_filter_and_lap_offset(monkeypatch):
    seen = _capture(monkeypatch)
    lap = 19.2296
    vehicles = {"ego": Rival(0, 0, 0), "car1": Rival(lap + 4.0, 0.2, 0.1), "car2": Rival(10.0, 0.2, -0.1),
                "car3": Rival(2.0 * lap + 2.5, 0.0, 0.3)}
    param = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.eye(6), matrix_R=np.eye(2),
                                  num_horizon=10, alpha=0.8)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    track = types.SimpleNamespace(width=1.0, lap_length=lap)
    x = np.array([1.2, 0, 0, 0, 3.0, 0.0])
    u = control.mpccbf(x, np.array([0.8, 0, 0, 0, 0, 0]).reshape(6, 1), param, vehicles, "ego", lap, 0.0, 0.1, False, track, sysp)
    assert u.shape == (2,) and (u == 0.25).all()
    # window is +-2*vx = 2.4 m: car1 (lap 1, s mod lap = 4.0) and car3 (lap 2, 2.5) are kept, car2 (10.0) is not
    assert seen["obs"].shape == (1, 2, 2, 11)
    assert np.allclose(seen["obs"][0, 0, 0], lap + 4.0 + 0.02 * np.arange(11)) and np.allclose(seen["obs"][0, 1, 1], 0.3)
    assert np.allclose(seen["lap_off"], [[-lap, -2 * lap]])
    assert seen["prm"]["alpha"] == 0.8 and seen["prm"]["margin"] == 0.2 and seen["prm"]["L"] == 0.4 and seen["prm"]["W"] == 0.2
    assert seen["xt"].shape == (6,)


def _mpccbf_args(vehicles, N=5):
    param = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.eye(6), matrix_R=np.eye(2),
                                  num_horizon=N, alpha=0.8)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    return (np.array([1.2, 0, 0, 0, 3.0, 0.0]), np.zeros(6), param, vehicles, "ego", 19.2296, 0.0, 0.1, False,
            types.SimpleNamespace(width=1.0), sysp)


def test_mpccbf_mixed_rival_sizes_go_into_the_record(monkeypatch):
    """control.py:530-535 reads every rival's own length / width: different rivals -> the per-rival block of the record."""
    seen = _capture(monkeypatch)
    vehicles = {"ego": Rival(0, 0, 0), "a": Rival(4.0, 0, 0.1), "b": Rival(4.5, 0, -0.3, length=0.6, width=0.3)}
    control.mpccbf(*_mpccbf_args(vehicles))
    assert np.allclose(seen["opt"]["sizes"], [[0.4, 0.2], [0.5, 0.25]])
    vehicles["b"] = Rival(4.5, 0, -0.3)
    control.mpccbf(*_mpccbf_args(vehicles))
    assert seen["opt"]["sizes"] is None and seen["prm"]["L"] == 0.4 and seen["prm"]["W"] == 0.2


def test_mpccbf_keeps_the_nearest_rivals_beyond_the_kernel_limit(monkeypatch):
    from car_racing_b200 import _capi
    seen = _capture(monkeypatch)
    vehicles = {"ego": Rival(0, 0, 0)}
    gaps = np.linspace(-2.0, 2.2, _capi.MMAX + 3)             # all inside the +-2.4 m window
    for k, g in enumerate(gaps):
        vehicles["car%d" % k] = Rival(3.0 + g, 0.0, 0.1 * k - 0.5)
    with pytest.warns(UserWarning, match="nearest"):
        control.mpccbf(*_mpccbf_args(vehicles))
    assert seen["obs"].shape[1] == _capi.MMAX
    kept_gap = np.abs(seen["obs"][0, :, 0, 0] - 3.0)
    assert kept_gap.max() <= np.sort(np.abs(gaps))[_capi.MMAX - 1] + 1e-12
    assert (np.diff(seen["obs"][0, :, 0, 0]) > 0).all()       # the reference's rival order is kept


def test_single_instance_retries(monkeypatch):
    """MAX_ITER at the batch default of 200 iterations -> IPOPT's limit; an elastic solution -> a stiffer penalty; the shim
    never raises, mpc_lti does (control.py:242)."""
    seen = _capture(monkeypatch)
    vehicles = {"ego": Rival(0, 0, 0), "a": Rival(4.0, 0, 0.1)}
    seen["statuses"], seen["elastic"] = [1, 0], [0.0, 0.0]
    control.mpccbf(*_mpccbf_args(vehicles))
    assert len(seen["calls"]) == 2 and seen["calls"][1]["max_iter"] == 3000 and control.last_solve["status"] == 0
    seen.pop("calls")
    seen["statuses"], seen["elastic"] = [0, 0], [0.3, 0.0]
    control.mpccbf(*_mpccbf_args(vehicles))
    assert len(seen["calls"]) == 2 and seen["calls"][1]["rho"] == 1e5 and control.last_solve["elastic_max"] == 0.0
    seen.pop("calls")
    seen["statuses"], seen["elastic"] = [0, 1], [0.3, 0.0]
    with pytest.warns(UserWarning, match="elastically"):
        control.mpccbf(*_mpccbf_args(vehicles))
    assert control.last_solve["elastic_max"] == 0.3 and control.last_solve["retries"][0][0].startswith("rho")
    seen.pop("calls")
    seen["statuses"], seen["elastic"] = [4], [0.0]
    a = _mpccbf_args(vehicles)
    control.mpccbf(*a)                                        # infeasible x0: reported, the iterate is used (control.py:600-603)
    assert control.last_solve["status"] == 4
    lti = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.eye(6), matrix_R=np.eye(2), num_horizon=5)
    with pytest.raises(RuntimeError, match="stage-0"):
        control.mpc_lti(a[0], a[1], lti, a[10], types.SimpleNamespace(width=1.0))


def test_mpc_multi_agents_targets(monkeypatch):
    seen = _capture(monkeypatch)
    lap = 19.1313
    vehicles = {"ego": Rival(0, 0, 0), "car1": Rival(5.0, 1.0, -0.5), "car2": Rival(30.0, 1.0, 0.2)}
    param = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.eye(6), matrix_R=np.eye(2),
                                  num_horizon_ctrl=10)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    track = types.SimpleNamespace(width=1.0, lap_length=lap)
    traj = np.zeros((11, 6))
    traj[:, 4] = 4.0 + 0.15 * np.arange(11)
    traj[:, 5] = 0.05 * np.arange(11)
    x = np.array([1.5, 0, 0, 0, 3.9, 0.1])
    u, xp = control.mpc_multi_agents(x, param, track, None, None, None, sysp, target_traj_xcurv=traj, vehicles=vehicles,
                                     agent_name="ego", direction_flag=0, sorted_vehicles=["car1", "car2"], time=None)
    assert u.shape == (2,) and xp.shape == (11, 6)
    xt = seen["xt"]
    assert xt.shape == (1, 11, 6) and (xt[0, :, 0] == 1.5).all()
    s = np.clip(1.5 * 0.1 * np.arange(11) + 3.9, 4.0, traj[-1, 4])
    assert np.allclose(xt[0, :, 5], np.interp(s, traj[:, 4], traj[:, 5]))
    assert seen["obs"].shape[1] == 1 and seen["prm"]["alpha"] == 0.6 and seen["prm"]["margin"] == 0.15
    assert seen["prm"]["L"] == 0.4 and seen["prm"]["W"] == 0.2


def test_install_all_patches_every_hook():
    import types
    import car_racing_b200 as crb
    from car_racing_b200 import planning, rivals
    ctrl = types.SimpleNamespace()
    base = types.SimpleNamespace(LMPCRacingGame=type("LMPCRacingGame", (), {}), NoDynamicsModel=type("NoDynamicsModel", (), {}))
    planner = type("OvertakeTrajPlanner", (), {})
    offboard = types.SimpleNamespace(DynamicBicycleModel=type("DynamicBicycleModel", (), {}))
    crb.install_all(ctrl, base, planner, offboard)
    assert offboard.DynamicBicycleModel.get_trajectory_nsteps is rivals.dynamic_get_trajectory_nsteps
    for name in ("mpc_lti", "mpccbf", "mpc_multi_agents", "ilqr", "lmpc"):
        assert getattr(ctrl, name) is getattr(crb.control, name)
    assert base.LMPCRacingGame.estimate_ABC is crb.control.estimate_ABC
    assert base.NoDynamicsModel.get_trajectory_nsteps is rivals.get_trajectory_nsteps
    assert planner.solve_optimization_problem is planning.solve_optimization_problem
    assert planner.get_local_traj is planning.get_local_traj

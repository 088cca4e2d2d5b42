"""Candidate preparation of the overtake planner: the numpy restatement against what the
unmodified reference returns (tests/golden/planner_prep_golden.npz, made by make_planner_prep_golden.py), and the CUDA
path (b200mpc_planner_prepare / b200mpc_plan_and_track_prepared through the C-ABI) against both the restatement and
the host packing of car_racing_b200.planning.  Floating-point stage: tolerance 1e-12 on curve samples, targets and
record entries (the kernel reproduces the reference's operation order without FMA contraction; the only difference
left is (1-t)**3 evaluated as pow() in the reference and as two products here)."""
import os
import types

import numpy as np
import pytest

import planner_numpy
from car_racing_b200 import planning, scenarios

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "planner_prep_golden.npz")
TOL = 1e-12


def _cases():
    g = np.load(GOLD)
    for ci in range(int(g["num_cases"])):
        yield ci, {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith("case%d/" % ci)}, g["opt_traj"], float(g["lap_length"])


def test_restatement_matches_reference_golden():
    n = 0
    for ci, c, opt, lap in _cases():
        r = planner_numpy.prepare(c["ego_x"], c["ego_x"], c["obs"], c["insertion"], c["rival_vx"], float(c["prediction_factor"]), 1.0,
                                  lap, 0.2, opt, 10)
        assert abs(r["max_delta_v"] - float(c["max_delta_v"])) == 0.0
        for k, gk in (("ctrl", "ctrl"), ("bezier", "bezier_xcurvs"), ("s_ref", "s_ref"), ("ey_ref", "ey_ref")):
            assert np.array_equal(r[k], c[gk]), (ci, k)      # bit for bit: same numpy operations in the same order
        n += 1
    assert n == 10


def test_rival_ordering_reproduces_the_reference_insertion_rule():
    """overtake_traj_planner.py:69-77 is not a full sort for more than two rivals; both mirrors must reproduce it."""
    seen_unsorted = False
    for ci, c, opt, lap in _cases():
        nv = int(c["num_veh"])
        ey_now = np.array([c["obs"][c["insertion"][i], 1, 0] for i in range(nv)])     # insertion order
        for fn in (planner_numpy.sort_rivals, planning.sort_rivals):
            order = fn(list(ey_now))
            assert [order.index(i) for i in range(nv)] == c["insertion"].tolist(), ci
        ey_sorted = ey_now[planner_numpy.sort_rivals(list(ey_now))]
        seen_unsorted |= bool((np.diff(ey_sorted) > 0).any() and (np.diff(ey_sorted) < 0).any())
    assert seen_unsorted        # the golden set contains a case where the rule leaves the rivals unsorted


def test_out_of_range_lookup_raises_like_interp1d():
    ci, c, opt, lap = next(_cases())
    ego = c["ego_x"].copy()
    ego[4] = opt[-1, 0] + 0.5                       # beyond the optimal trajectory's last abscissa (:98-100)
    with pytest.raises(ValueError):
        planner_numpy.prepare(ego, ego, c["obs"], c["insertion"], c["rival_vx"], 0.5, 1.0, lap, 0.2, opt, 10)


def test_prepare_input_validation_needs_no_gpu():
    ci, c, opt, lap = next(_cases())
    with pytest.raises(ValueError):
        planning._prepare_inputs(c["ego_x"], c["ego_x"], c["obs"][:, :, :5], c["insertion"], c["rival_vx"], opt, 10)
    with pytest.raises(ValueError):
        planning._prepare_inputs(c["ego_x"], c["ego_x"], c["obs"], [0] * len(c["insertion"]) + [], c["rival_vx"], opt[::-1], 10)
    with pytest.raises(ValueError):
        planning._prepare_inputs(c["ego_x"], c["ego_x"], c["obs"], [5], c["rival_vx"], opt, 10)


def _host_packing(c, opt, lap, N=10):
    """The host path of car_racing_b200.planning on the reference's own prepared state (golden curve samples)."""
    from scipy.interpolate import interp1d
    nv = int(c["num_veh"])
    names = ["car%d" % j for j in range(nv)]
    obs_infos = {}
    for j, n in enumerate(names):
        tr = np.zeros((6, N + 1))
        tr[4:6] = c["obs"][j]
        obs_infos[n] = tr
    bez = c["bezier_xcurvs"]
    funcs = [interp1d(bez[i, :, 0], bez[i, :, 1]) for i in range(nv + 1)]
    out = dict(s_ref=[], ey_ref=[], xlb=[], xub=[], heur=[], ok0=[])
    for i in range(nv + 1):
        xlb, xub = planning.candidate_bounds(i, c["ego_x"], names, obs_infos, 0.4, 0.2, 1.0, lap, N)
        s_ref, ey_ref = planning.candidate_targets(i, c["ego_x"], bez, funcs, N)
        out["xlb"].append(xlb); out["xub"].append(xub); out["s_ref"].append(s_ref); out["ey_ref"].append(ey_ref)
        out["heur"].append(planning.heuristic_traj(i, c["ego_x"], bez, funcs, N).T)
        out["ok0"].append(planning.x0_feasible(c["ego_x"], xlb, xub))
    kw, off = planning.pack_candidates(c["ego_x"], np.array(out["s_ref"]), np.array(out["ey_ref"]), np.array(out["xlb"]),
                                       np.array(out["xub"]), N)
    from car_racing_b200 import batch
    rec, _, _ = batch.pack_cbf(kw["x0"], kw["xt"], kw["obs"], None, N, xlb=kw["xlb"], xub=kw["xub"], wd=kw["wd"])
    return rec, off, np.array(out["heur"]), np.array(out["ok0"], dtype=np.int32)


def _emulated_prepare(c, opt, lap, N=10, ego=None):
    """The per-region body of planner_prepare.cuh compiled for the host (tests/host_emulation/): checks the kernel's LOGIC
    without a GPU; the CUDA build of the same source is checked by the -m gpu tests below."""
    import ctypes as C
    import subprocess
    from car_racing_b200 import _capi, batch
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emulation")
    lib = os.path.join(here, "_build", "libprep_emu.so")
    src = [os.path.join(here, "planner_prepare_host.cpp"), os.path.join(here, "cuda_runtime.h"),
           os.path.join(os.path.dirname(here), "..", "car_racing_b200", "csrc", "planner_prepare.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(f) > os.path.getmtime(lib) for f in src):
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I", here, src[0], "-o", lib], check=True)
    L = C.CDLL(lib)
    ego = c["ego_x"] if ego is None else ego
    egov, rivals, ins, vx, optc = planning._prepare_inputs(ego, ego, c["obs"], c["insertion"], c["rival_vx"], opt, N)
    nv = rivals.shape[0]
    p = planning._prepare_params(N, nv, optc.shape[0], float(c["prediction_factor"]), 1.0, lap, 0.4, 0.2)
    fl = _capi.FLAG_STAGE_BOUNDS | _capi.FLAG_EY_RATE
    stride, base = batch.cbf_record_doubles(N, 0, True, fl), batch.cbf_record_doubles(N, 0, True, 0)
    Cn = nv + 1
    out = dict(records=np.zeros((Cn, stride)), heur=np.zeros((Cn, N + 1, 6)), ok0=np.zeros(Cn, dtype=np.int32),
               region=np.zeros(Cn, dtype=np.int32), offset=np.zeros(Cn), ctrl=np.zeros((Cn, 4, 2)), bezier=np.zeros((Cn, N + 1, 2)))
    P = batch._ptr
    L.emu_planner_prepare.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 12
    out["err"] = L.emu_planner_prepare(C.byref(p), stride, 6, base, base + 4 * (N + 1), P(egov), P(rivals), P(vx), P(ins), P(optc),
                                       P(out["records"]), P(out["heur"]), P(out["ok0"]), P(out["region"]), P(out["offset"]),
                                       P(out["ctrl"]), P(out["bezier"]))
    return out


def _check_prepared(r, c, opt, lap, ci):
    assert np.abs(r["ctrl"] - c["ctrl"]).max() < TOL, ci
    assert np.abs(r["bezier"] - c["bezier_xcurvs"]).max() < TOL, ci
    rec, off, heur, ok0 = _host_packing(c, opt, lap)
    assert r["records"].shape == rec.shape
    assert np.abs(r["records"] - rec).max() < TOL, ci          # x0, targets, bounds (+-1e300 = none), ey-rate weights
    assert np.abs(r["heur"] - heur).max() < TOL and (r["ok0"] == ok0).all() and (r["region"] == np.arange(len(ok0))).all()
    assert np.abs(r["offset"] - off).max() < 1e-9
    return bool((rec[:, -(4 * 11 + 10):-10].reshape(len(ok0), 11, 4)[:, :, 1] > -0.89).any())


def test_kernel_body_compiled_for_host_matches_reference_golden_and_host_packing():
    some_rival_row = False
    for ci, c, opt, lap in _cases():
        r = _emulated_prepare(c, opt, lap)
        assert r["err"] == 0
        some_rival_row |= _check_prepared(r, c, opt, lap, ci)
    assert some_rival_row       # at least one case carries a rival row (ey_k >= ey_rival + width + margin)
    ci, c, opt, lap = next(_cases())
    ego = c["ego_x"].copy()
    ego[4] = opt[-1, 0] + 0.5
    assert _emulated_prepare(c, opt, lap, ego=ego)["err"] == 1      # the reference raises ValueError here


# ---------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_device_preparation_matches_reference_golden_and_host_packing(crb):
    some_rival_row = False
    for ci, c, opt, lap in _cases():
        r = planning.prepare_candidates(c["ego_x"], c["ego_x"], c["obs"], c["insertion"], c["rival_vx"], opt, 10,
                                        prediction_factor=float(c["prediction_factor"]), track_width=1.0, lap_length=lap)
        some_rival_row |= _check_prepared(r, c, opt, lap, ci)
    assert some_rival_row


@pytest.mark.gpu
def test_device_preparation_flags_out_of_range_lookup(crb):
    ci, c, opt, lap = next(_cases())
    ego = c["ego_x"].copy()
    ego[4] = opt[-1, 0] + 0.5
    with pytest.raises(ValueError):
        planning.prepare_candidates(ego, ego, c["obs"], c["insertion"], c["rival_vx"], opt, 10, lap_length=lap)


@pytest.mark.gpu
def test_prepared_chain_matches_host_prepared_chain(crb):
    """b200mpc_plan_and_track_prepared (preparation on the device) against planning.plan_and_track fed with the host
    preparation of the same planner state: same region, same trajectories, same control."""
    from scipy.interpolate import interp1d
    from test_shims_host import Rival
    param = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 5.0, 0, 50.0]),
                                  matrix_R=np.diag([0.1, 0.1]), num_horizon_ctrl=10)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    flags, n_solved = [], 0
    for ci, c, opt, lap in _cases():
        nv = int(c["num_veh"])
        ins = c["insertion"].tolist()
        names = ["car%d" % (i + 1) for i in range(nv)]                         # insertion order
        opt6 = np.zeros((opt.shape[0], 6))
        opt6[:, 4:6] = opt

        def planner():
            veh = {"ego": types.SimpleNamespace(param=types.SimpleNamespace(length=0.4, width=0.2), xcurv=c["ego_x"].copy())}
            for i, n in enumerate(names):
                o = c["obs"][ins[i]]
                veh[n] = Rival(o[0, 0], c["rival_vx"][ins[i]], o[1, 0])      # constant-ey prediction from the golden start point
                veh[n].no_dynamics = True
            rg = types.SimpleNamespace(num_horizon_planner=10, matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, timestep=0.1,
                                       planning_prediction_factor=0.5)
            return types.SimpleNamespace(vehicles=veh, agent_name="ego", track=types.SimpleNamespace(width=1.0, lap_length=lap),
                                         opti_traj_xcurv=opt6, racing_game_param=rg, old_direction_flag=None)
        p1, p2 = planner(), planner()
        x = c["ego_x"].copy()
        interest = {n: p1.vehicles[n] for n in names}
        (t1, f1, st1, s1), (u1, x1) = planning.plan_and_track_from_predictions(p1, x, 0.0, interest, x, param, p1.track, sysp)
        # host preparation of the same state -> the existing chain
        order = planning.sort_rivals([p2.vehicles[n].xcurv[5] for n in names])
        p2.sorted_vehicles = [names[i] for i in order]
        assert p2.sorted_vehicles == p1.sorted_vehicles
        p2.obs_infos = {n: p2.vehicles[n].get_trajectory_nsteps(0.0, 0.1, 11)[0] for n in names}
        obs_sorted = np.array([p2.obs_infos[n][4:6] for n in p2.sorted_vehicles])
        r = planner_numpy.prepare(x, x, obs_sorted, [p2.sorted_vehicles.index(n) for n in names],
                                  [p2.vehicles[n].xcurv[0] for n in p2.sorted_vehicles], 0.5, 1.0, lap, 0.2, opt, 10)
        p2.bezier_xcurvs = r["bezier"]
        p2.bezier_funcs = [interp1d(r["bezier"][i, :, 0], r["bezier"][i, :, 1]) for i in range(nv + 1)]
        p2.xcurv_ego = x
        (t2, f2, st2, s2), (u2, x2) = planning.plan_and_track(p2, x, param, p2.track, sysp, time=None)
        assert np.abs(p1.bezier_xcurvs - r["bezier"]).max() < TOL
        assert f1 == f2 and np.abs(p1.selection_costs - p2.selection_costs).max() < 1e-9
        assert np.abs(t1 - t2).max() < 1e-8 and np.abs(s1 - s2).max() < 1e-8
        assert p1.tracking_status == p2.tracking_status
        assert np.abs(u1 - u2).max() < 1e-7 and np.abs(x1 - x2).max() < 1e-7
        fin = np.isfinite(p1.candidate_costs)
        assert (fin == np.isfinite(p2.candidate_costs)).all()      # all False when x_0 violates a stage-0 row of every region
        assert not fin.any() or np.abs(p1.candidate_costs[fin] - p2.candidate_costs[fin]).max() < 1e-6
        n_solved += int(fin.sum())
        # the get_local_traj drop-in (no tracking stage): same plan, the reference's 8-tuple
        p3 = planner()
        p3.track.get_global_position = lambda s, ey: (s + 100.0, ey - 100.0)
        out = planning.get_local_traj(p3, x, 0.0, {n: p3.vehicles[n] for n in names}, None, None, None, None, None)
        assert len(out) == 8 and out[2] == f1 and out[3] == p1.sorted_vehicles
        assert np.abs(out[0] - t1).max() < 1e-12 and out[5].shape == (nv + 1,)
        assert out[6].shape == (nv + 1, 11, 6) and out[7].shape == (nv + 1, 11, 6) and out[4].shape == (11, 6)
        s_wrapped = np.where(t1[:, 4] > lap, t1[:, 4] - lap, t1[:, 4])
        assert np.abs(out[1][:, 4] - (s_wrapped + 100.0)).max() < 1e-12 and np.abs(out[1][:, 5] - (t1[:, 5] - 100.0)).max() < 1e-12
        assert np.abs(out[7][f1][:, 5] - (t1[:, 5] - 100.0)).max() < 1e-12
        flags.append(f1)
    assert len(flags) == 10 and n_solved >= 5 and len(set(flags)) > 1


@pytest.mark.gpu
def test_prepared_chain_with_64_candidates_equals_host_prepared_chain(crb):
    """BASELINE config 3 (64 candidates): the reference's regions prepared on the device + 61 host-packed extra candidates
    behind them (b200mpc_plan_and_track_prepared, n_extra > 0) against planning.plan_and_track with everything packed on
    the host."""
    from scipy.interpolate import interp1d
    from test_shims_host import Rival
    param = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 5.0, 0, 50.0]),
                                  matrix_R=np.diag([0.1, 0.1]), num_horizon_ctrl=10)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)
    done = 0
    for ci, c, opt, lap in _cases():
        nv = int(c["num_veh"])
        if nv != 2:
            continue
        ins, N = c["insertion"].tolist(), 10
        names = ["car%d" % (i + 1) for i in range(nv)]
        opt6 = np.zeros((opt.shape[0], 6))
        opt6[:, 4:6] = opt

        def planner():
            veh = {"ego": types.SimpleNamespace(param=types.SimpleNamespace(length=0.4, width=0.2), xcurv=c["ego_x"].copy())}
            for i, n in enumerate(names):
                o = c["obs"][ins[i]]
                veh[n] = Rival(o[0, 0], c["rival_vx"][ins[i]], o[1, 0])
                veh[n].no_dynamics = True
            rg = types.SimpleNamespace(num_horizon_planner=N, matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, timestep=0.1,
                                       planning_prediction_factor=0.5)
            return types.SimpleNamespace(vehicles=veh, agent_name="ego", track=types.SimpleNamespace(width=1.0, lap_length=lap),
                                         opti_traj_xcurv=opt6, racing_game_param=rg, old_direction_flag=None)
        p1, p2 = planner(), planner()
        x = c["ego_x"].copy()
        # host preparation (restatement) -> extras as tests/test_gpu_parity.py::test_plan_and_track_with_64_candidates builds them
        order = planning.sort_rivals([p2.vehicles[n].xcurv[5] for n in names])
        p2.sorted_vehicles = [names[i] for i in order]
        p2.obs_infos = {n: p2.vehicles[n].get_trajectory_nsteps(0.0, 0.1, N + 1)[0] for n in names}
        obs_sorted = np.array([p2.obs_infos[n][4:6] for n in p2.sorted_vehicles])
        r = planner_numpy.prepare(x, x, obs_sorted, [p2.sorted_vehicles.index(n) for n in names],
                                  [p2.vehicles[n].xcurv[0] for n in p2.sorted_vehicles], 0.5, 1.0, lap, 0.2, opt, N)
        p2.bezier_xcurvs, p2.xcurv_ego = r["bezier"], x
        p2.bezier_funcs = [interp1d(r["bezier"][i, :, 0], r["bezier"][i, :, 1]) for i in range(nv + 1)]
        ex = dict(s_ref=[], ey_ref=[], xlb=[], xub=[], region=[], heur=[])
        for reg in range(3):
            xlb, xub = planning.candidate_bounds(reg, x, p2.sorted_vehicles, p2.obs_infos, 0.4, 0.2, 1.0, lap, N)
            s0, e0 = planning.candidate_targets(reg, x, p2.bezier_xcurvs, p2.bezier_funcs, N)
            h0 = planning.heuristic_traj(reg, x, p2.bezier_xcurvs, p2.bezier_funcs, N).T
            for scale in (0.4, 0.6, 0.8, 0.9, 1.1, 1.2, 1.4):
                for stretch in (0.9, 1.0, 1.1):
                    if len(ex["region"]) == 61:
                        break
                    ex["s_ref"].append(x[4] + stretch * (s0 - x[4]))
                    ex["ey_ref"].append(x[5] + scale * (e0 - x[5]))
                    ex["xlb"].append(xlb); ex["xub"].append(xub); ex["region"].append(reg); ex["heur"].append(h0)
        ex = {k: np.array(v) for k, v in ex.items()}
        (t2, f2, st2, s2), (u2, x2) = planning.plan_and_track(p2, x, param, p2.track, sysp, time=None, extra=ex)
        (t1, f1, st1, s1), (u1, x1) = planning.plan_and_track_from_predictions(p1, x, 0.0, {n: p1.vehicles[n] for n in names}, x, param,
                                                                               p1.track, sysp, extra=ex)
        assert s1.shape[0] == 64 and f1 == f2 and np.abs(p1.selection_costs - p2.selection_costs).max() < 1e-9
        assert np.abs(s1 - s2).max() < 1e-8 and np.abs(t1 - t2).max() < 1e-8
        assert np.abs(u1 - u2).max() < 1e-7 and np.abs(x1 - x2).max() < 1e-7
        done += 1
    assert done >= 3

"""The whole C-ABI library on the host: car_racing_b200/csrc/capi.cu and every kernel header compiled by g++
(tests/host_emulation/build_emu_library.py: launches become emu_launch, "device" memory is host memory, one host thread
per CUDA thread with barriers for __syncthreads / __syncwarp / shuffles), loaded INTO THE TESTS ONLY in place of
libb200mpc.so, so that the product's own Python API -- packers, parameter structs, dispatch, kernels -- is exercised end
to end without a GPU at small sizes, including the -m gpu tests that fit (GPU_TESTS_ON_HOST below).  This is a checker of logic; the CUDA
build of the same sources is what `-m gpu` and bench.py run, and nothing on the product path can load the emulated
library."""
import importlib.util
import os
import types

import numpy as np
import pytest

import car_racing_b200 as crb
from car_racing_b200 import _capi, batch, planning, scenarios

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
pytestmark = pytest.mark.timeout(900)


@pytest.fixture(scope="module")
def emu():
    from conftest import wait_prebuilt          # tests/host_emulation/build_emu_library.py, started in the background at session start
    path = wait_prebuilt("b200mpc_emu")
    saved = (_capi.LIB_PATH, _capi._lib, batch._default_handle)
    _capi.LIB_PATH, _capi._lib, batch._default_handle = path, None, None
    try:
        yield path
    finally:
        if batch._default_handle is not None:
            batch._default_handle.close()
        _capi.LIB_PATH, _capi._lib, batch._default_handle = saved


def test_mpccbf_through_the_product_api(emu, oracle):
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(2, N=20, M=3, seed=1)
    prm = scenarios.default_cbf_params(N=20)
    g = crb.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    assert (g["status"] == r["status"]).all() and (g["iters"] == r["iters"]).all()
    ok = g["status"] == 0
    assert ok.any() and np.abs(g["u0"] - r["u0"])[ok].max() < 1e-4 and np.abs(g["cost"] - r["cost"])[ok].max() < 1e-5
    assert batch.default_handle().launch_count == 1


def test_full_weight_matrix_takes_the_general_instantiation(emu, oracle):
    """The C-ABI's dispatch (capi.cu): the north-star shape with the reference's diagonal Q runs <3,QDIAG,20>, the same shape with
    off-diagonal weights the general <3,0,20>; both against the oracle through the product API."""
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(2, N=20, M=3, seed=2)
    prm = scenarios.default_cbf_params(N=20)
    Q = np.array(prm["Q"], float)
    Q[0, 3] = Q[3, 0] = 1.5
    Q[3, 5] = Q[5, 3] = -2.0
    for p in (prm, dict(prm, Q=Q)):
        g = crb.solve_cbf_batch(x0, xt, obs, lap_off, p)
        r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, p)
        assert (g["status"] == r["status"]).all() and (g["iters"] == r["iters"]).all()
        ok = g["status"] == 0
        assert ok.any() and np.abs(g["u0"] - r["u0"])[ok].max() < 1e-4 and np.abs(g["cost"] - r["cost"])[ok].max() < 1e-5


def test_lmpc_against_the_oracle(emu, oracle):
    sc = scenarios.lmpc_scenarios(2, seed=5)
    prm = scenarios.default_lmpc_params()
    g = crb.solve_lmpc_batch(*sc, prm)
    r = oracle.solve_lmpc_batch(*sc, prm)
    assert (g["status"] == r["status"]).all()
    ok = g["status"] == 0
    assert ok.any() and np.abs(g["u0"] - r["u0"])[ok].max() < 1e-4 and np.abs(g["cost"] - r["cost"])[ok].max() < 1e-5
    assert np.abs(g["x"][ok] - r["x"][ok]).max() < 1e-4 and np.abs(g["lambda"][ok] - r["lam"][ok]).max() < 1e-4


def test_overtaking_step_from_predictions(emu):
    """b200mpc_plan_and_track_prepared (preparation, candidate QPs, selection, tracking MPC-CBF in one call) against the same
    step with the candidates packed on the host, both through the library on the host; plus the get_local_traj drop-in."""
    from scipy.interpolate import interp1d
    import planner_numpy
    from test_shims_host import Rival
    g = np.load(os.path.join(GOLD, "planner_prep_golden.npz"))
    fg = np.load(os.path.join(GOLD, "frenet_golden.npz"))
    c = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith("case1/")}
    opt, lap = g["opt_traj"], float(fg["goggle/lap_length"])
    nv, ins = int(c["num_veh"]), c["insertion"].tolist()
    names = ["car%d" % (i + 1) for i in range(nv)]
    opt6 = np.zeros((opt.shape[0], 6))
    opt6[:, 4:6] = opt
    param = types.SimpleNamespace(matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, matrix_Q=np.diag([10.0, 0, 0, 5.0, 0, 50.0]),
                                  matrix_R=np.diag([0.1, 0.1]), num_horizon_ctrl=10)
    sysp = types.SimpleNamespace(delta_max=0.5, a_max=1.0, v_max=10, v_min=0)

    def planner():
        veh = {"ego": types.SimpleNamespace(param=types.SimpleNamespace(length=0.4, width=0.2), xcurv=c["ego_x"].copy())}
        for i, n in enumerate(names):
            o = c["obs"][ins[i]]
            veh[n] = Rival(o[0, 0], c["rival_vx"][ins[i]], o[1, 0])
            veh[n].no_dynamics = True
        rg = types.SimpleNamespace(num_horizon_planner=10, matrix_A=scenarios.LTI_A, matrix_B=scenarios.LTI_B, timestep=0.1,
                                   planning_prediction_factor=0.5)
        track = types.SimpleNamespace(width=1.0, lap_length=lap, point_and_tangent=fg["goggle/pat"])
        return types.SimpleNamespace(vehicles=veh, agent_name="ego", track=track, opti_traj_xcurv=opt6, racing_game_param=rg,
                                     old_direction_flag=None)
    p1, p2, p3 = planner(), planner(), planner()
    x = c["ego_x"].copy()
    (t1, f1, _, s1), (u1, x1) = planning.plan_and_track_from_predictions(p1, x, 0.0, {n: p1.vehicles[n] for n in names}, x, param,
                                                                         p1.track, sysp)
    order = planning.sort_rivals([p2.vehicles[n].xcurv[5] for n in names])
    p2.sorted_vehicles = [names[i] for i in order]
    p2.obs_infos = {n: p2.vehicles[n].get_trajectory_nsteps(0.0, 0.1, 11)[0] for n in names}
    obs_sorted = np.array([p2.obs_infos[n][4:6] for n in p2.sorted_vehicles])
    r = planner_numpy.prepare(x, x, obs_sorted, [p2.sorted_vehicles.index(n) for n in names],
                              [p2.vehicles[n].xcurv[0] for n in p2.sorted_vehicles], 0.5, 1.0, lap, 0.2, opt, 10)
    p2.bezier_xcurvs, p2.xcurv_ego = r["bezier"], x
    p2.bezier_funcs = [interp1d(r["bezier"][i, :, 0], r["bezier"][i, :, 1]) for i in range(nv + 1)]
    (t2, f2, _, s2), (u2, x2) = planning.plan_and_track(p2, x, param, p2.track, sysp, time=None)
    assert f1 == f2 and np.abs(p1.selection_costs - p2.selection_costs).max() < 1e-9
    assert np.abs(t1 - t2).max() < 1e-8 and np.abs(s1 - s2).max() < 1e-8 and np.abs(u1 - u2).max() < 1e-7 and np.abs(x1 - x2).max() < 1e-7
    assert p1.tracking_status == 0 and np.abs(p1.bezier_xcurvs - r["bezier"]).max() < 1e-12
    out = planning.get_local_traj(p3, x, 0.0, {n: p3.vehicles[n] for n in names}, None, None, None, None, None)
    assert len(out) == 8 and out[2] == f1 and np.abs(out[0] - t1).max() < 1e-12
    import frenet_numpy
    for j in range(11):
        gx, gy, _ = frenet_numpy.curv_to_glob(lap, fg["goggle/pat"], t1[j, 4], t1[j, 5])
        assert abs(out[1][j, 4] - gx) < 1e-12 and abs(out[1][j, 5] - gy) < 1e-12


def test_small_kernels_through_the_product_api(emu):
    """rival_rollout_kernel, curv_to_glob_kernel, planner_prepare_kernel, plant_kernel through their batch functions against
    the reference goldens."""
    g = np.load(os.path.join(GOLD, "rollout_golden.npz"))
    oc, og = crb.rival_rollout_batch(g["l_shape/n21/xcurv0"], g["l_shape/n21/xglob0"], g["l_shape/pat"], float(g["l_shape/lap_length"]),
                                     0.1, 21, with_glob=True)
    assert np.abs(oc - g["l_shape/n21/xcurv"]).max() < 1e-12 and np.abs(og - g["l_shape/n21/xglob"]).max() < 1e-12
    f = np.load(os.path.join(GOLD, "frenet_golden.npz"))
    x, y, psi = crb.curv_to_glob_batch(f["m_shape/s"], f["m_shape/ey"], f["m_shape/pat"], float(f["m_shape/lap_length"]))
    assert np.abs(x - f["m_shape/xy"][:, 0]).max() < 1e-12 and np.abs(y - f["m_shape/xy"][:, 1]).max() < 1e-12
    assert np.abs(psi - f["m_shape/psi"]).max() < 1e-12
    pg = np.load(os.path.join(GOLD, "planner_prep_golden.npz"))
    k = lambda n: pg["case9/" + n]     # noqa: E731  (4 rivals)
    r = planning.prepare_candidates(k("ego_x"), k("ego_x"), k("obs"), k("insertion"), k("rival_vx"), pg["opt_traj"], 10,
                                    lap_length=float(pg["lap_length"]))
    assert np.abs(r["ctrl"] - k("ctrl")).max() < 1e-12 and np.abs(r["bezier"] - k("bezier_xcurvs")).max() < 1e-12
    pl = np.load(os.path.join(GOLD, "plant_golden.npz"))
    nc, ng = crb.plant_step_batch(pl["xcurv_ellipse"][:, 0], pl["xglob_ellipse"][:, 0], pl["u_ellipse"][:, 0], pl["draws_ellipse"][:, 0],
                                  pl["pat_ellipse"], dyn=tuple(pl["dyn"]))
    assert np.abs(nc - pl["xcurv_ellipse"][:, 1]).max() < 1e-12 and np.abs(ng - pl["xglob_ellipse"][:, 1]).max() < 1e-12


GPU_TESTS_ON_HOST = [
    ("test_gpu_parity", "test_mpc_lti_anchor_on_gpu"), ("test_gpu_parity", "test_capi_argument_validation"),
    ("test_gpu_parity", "test_drop_in_shims_on_gpu"), ("test_gpu_parity", "test_planner_drop_in_on_gpu"),
    ("test_gpu_parity", "test_lmpc_drop_in_on_gpu"), ("test_gpu_parity", "test_sysid_chains_into_lmpc_records_and_drop_in"),
    ("test_gpu_parity", "test_moving_rivals_lap_offsets_per_stage_targets"), ("test_gpu_parity", "test_kkt_certificate_on_gpu_solutions"),
    ("test_gpu_parity", "test_blocked_lane_elastic_rows"), ("test_gpu_parity", "test_plan_and_track_chain_matches_host_path"),
    ("test_gpu_parity", "test_plan_and_track_with_64_candidates"), ("test_gpu_parity", "test_plant_step_matches_reference_golden"),
    ("test_gpu_parity", "test_sysid_matches_reference_golden"), ("test_gpu_parity", "test_ilqr_matches_reference_golden"),
    ("test_planner_prepare", "test_device_preparation_matches_reference_golden_and_host_packing"),
    ("test_planner_prepare", "test_device_preparation_flags_out_of_range_lookup"),
    ("test_planner_prepare", "test_prepared_chain_matches_host_prepared_chain"),
    ("test_rival_rollout", "test_device_rollout_matches_reference_golden"),
    ("test_rival_rollout", "test_device_rollout_large_batch_and_bad_arguments"),
    ("test_frenet", "test_device_conversion_matches_reference_golden"), ("test_frenet", "test_planner_plot_copies_use_one_launch"),
]


@pytest.mark.parametrize("module,name", GPU_TESTS_ON_HOST)
def test_gpu_tests_also_pass_on_the_host_library(emu, oracle, module, name):
    """The -m gpu tests that need neither torch.cuda nor large batches, run UNCHANGED against the library on the host (the
    others stay GPU-only: full BASELINE batches, torch device tensors, batches in flight): the SURVEY 8(c) anchor (MPC-LTI
    N=10: u0 = [0.0033384, 1.0], cost 22.71576162), C-ABI argument validation, the reference-signature shims, moving rivals
    with lap offsets, the KKT certificate, the blocked lane, the fused overtaking step (incl. 64 candidates), the reference
    goldens of iLQR / model identification / plant / rollout / frame conversion / candidate preparation."""
    import importlib
    import inspect
    fn = getattr(importlib.import_module(module), name)
    args = {"crb": crb, "oracle": oracle}
    fn(**{k: args[k] for k in inspect.signature(fn).parameters})


def test_exchange_window_protocol_two_ranks_in_one_process(emu):
    """b200mpc_comm_* (csrc/exchange.cuh) on the host library: two "ranks" (two handles, two windows whose handles are
    exchanged by hand) publish their shards from the solver kernel's epilogue into BOTH gathered buffers; each rank's argmin
    kernel then sees all records in rank-major order.  Three uses of one slot exercise the arrival counters and the
    acknowledgements; the MPC-CBF, iLQR and LMPC kernels all carry the epilogue."""
    import ctypes as C
    from car_racing_b200 import sharding
    L = _capi.lib()
    B, world = 3, 2
    hs = [_capi.Handle(), _capi.Handle()]
    # windows of both ranks inside this one process: create and export both, then connect both (the product does the
    # hand-over of the 64-byte handles with torch.distributed, one process per GPU)
    px = [sharding.PeerExchange.__new__(sharding.PeerExchange) for _ in range(world)]
    blobs = []
    for r in range(world):
        c = C.c_void_p()
        hs[r].check(L.b200mpc_comm_create(hs[r].ptr, r, world, 4, 2, C.byref(c)), "comm_create")
        px[r]._c, px[r].rank, px[r].world = c, r, world
        buf = (C.c_char * _capi.COMM_HANDLE_BYTES)()
        assert L.b200mpc_comm_export(c, buf) == 0
        blobs.append(bytes(buf.raw))
    assert L.b200mpc_comm_publish_next(hs[0].ptr, px[0]._c, 0) == -1          # not connected yet
    for r in range(world):
        assert L.b200mpc_comm_connect(px[r]._c, b"".join(blobs)) == 0
    prm = scenarios.default_cbf_params(N=10)
    o = _capi.default_options()
    shards = [scenarios.mpccbf_scenarios(B, N=10, M=1, seed=40 + r) for r in range(world)]
    p = _capi.make_cbf_params(prm, 1, False)
    P = batch._ptr
    for use in range(3):
        recs = []
        for r in range(world):
            x0, xt, obs, lo = shards[r]
            x0 = x0 + 0.01 * use
            rin, M, ps = batch.pack_cbf(x0, xt, obs, lo, 10)
            rec = np.zeros(B, dtype=_capi.RECORD_DTYPE)
            px[r].publish_next(hs[r], 1)
            hs[r].check(L.b200mpc_cbf_solve(hs[r].ptr, C.byref(p), C.byref(o), B, P(rin), P(rec), None, None, None, None), "solve")
            recs.append(rec)
        want = np.concatenate(recs)
        for r in range(world):
            arg = np.zeros(1, dtype=np.int32)
            allr = np.zeros(world * B, dtype=_capi.RECORD_DTYPE)
            px[r].argmin(hs[r], 1, P(arg), P(allr))
            assert (allr["cost"] == want["cost"]).all() and (allr["u0"] == want["u0"]).all() and (allr["iters"] == want["iters"]).all()
            assert arg[0] == sharding.argmin_first(want)
    # a solve without publish_next leaves the window alone; argmin without a publication is an error
    rec = np.zeros(B, dtype=_capi.RECORD_DTYPE)
    hs[0].check(L.b200mpc_cbf_solve(hs[0].ptr, C.byref(p), C.byref(o), B, P(rin), P(rec), None, None, None, None), "solve")
    assert L.b200mpc_comm_argmin(hs[0].ptr, px[0]._c, 1, 0, P(np.zeros(1, dtype=np.int32)), None) == -1
    # the other two solver kernels publish as well
    x0, xt, obs, lo = scenarios.ilqr_scenarios(2, N=12, seed=2)
    iprm = dict(A=prm["A"], B=prm["B"], Q=prm["Q"], R=prm["R"], N=12, max_iter=20, L=0.4, W=0.2)
    outs = []
    for r in range(world):
        px[r].publish_next(hs[r], 0)
        outs.append(crb.solve_ilqr_batch(x0 + 0.02 * r, xt, obs, lo, iprm, want=(), handle=hs[r])["record"])
    allr = np.zeros(4, dtype=_capi.RECORD_DTYPE)
    arg = np.zeros(1, dtype=np.int32)
    px[1].argmin(hs[1], 0, P(arg), P(allr))
    px[0].argmin(hs[0], 0, P(arg), None)
    assert (allr["cost"] == np.concatenate(outs)["cost"]).all()
    sc = scenarios.lmpc_scenarios(2, seed=5)
    lprm = scenarios.default_lmpc_params()
    outs = []
    for r in range(world):
        px[r].publish_next(hs[r], 0)
        outs.append(crb.solve_lmpc_batch(*sc, lprm, want=(), handle=hs[r])["record"])
    px[0].argmin(hs[0], 0, P(arg), P(allr))
    px[1].argmin(hs[1], 0, P(arg), None)
    assert (allr["cost"] == np.concatenate(outs)["cost"]).all() and arg[0] == sharding.argmin_first(np.concatenate(outs))
    for x in px:
        x.close()

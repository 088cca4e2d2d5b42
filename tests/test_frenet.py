"""Curvilinear -> global-frame conversion: racing_env.get_global_position / get_orientation
(utils/racing_env.py:6-127).  Restatement against the reference's output on four tracks (tests/golden/frenet_golden.npz,
made by make_frenet_golden.py), the kernel body compiled for the host against the same golden (CPU suite), the CUDA build
through the C-ABI (-m gpu).  Floating point: 1e-12 absolute on coordinates of a few metres."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import frenet_numpy
from car_racing_b200 import batch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "frenet_golden.npz")
TOL = 1e-12
TRACKS = ("ellipse", "l_shape", "goggle", "m_shape")


def _tracks():
    g = np.load(GOLD)
    for name in TRACKS:
        yield name, g[name + "/pat"], float(g[name + "/lap_length"]), g[name + "/s"], g[name + "/ey"], g[name + "/xy"], g[name + "/psi"]


def test_restatement_matches_reference_golden():
    for name, pat, lap, s, ey, xy, psi in _tracks():
        got = np.array([frenet_numpy.curv_to_glob(lap, pat, s[k], ey[k]) for k in range(s.shape[0])])
        assert np.array_equal(got[:, :2], xy) and np.array_equal(got[:, 2], psi), name
        assert (s < 0).any() and (s > 2 * lap).any()            # both wrap loops are exercised


def test_kernel_body_compiled_for_host_matches_reference_golden():
    emu = os.path.join(HERE, "host_emulation")
    lib = os.path.join(emu, "_build", "libfrenet_emu.so")
    src = [os.path.join(emu, "frenet_host.cpp"), os.path.join(emu, "cuda_runtime.h"),
           os.path.join(HERE, "..", "car_racing_b200", "csrc", "frenet.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(f) > os.path.getmtime(lib) for f in src):
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I", emu, src[0], "-o", lib], check=True)
    L = C.CDLL(lib)
    L.emu_curv_to_glob.argtypes = [C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 4
    L.emu_curv_to_glob.restype = None
    P = batch._ptr
    for name, pat, lap, s, ey, xy, psi in _tracks():
        out = np.zeros((s.shape[0], 3))
        patc, sc, ec = np.ascontiguousarray(pat[:, :6]), np.ascontiguousarray(s), np.ascontiguousarray(ey)
        L.emu_curv_to_glob(s.shape[0], patc.shape[0], lap, P(patc), P(sc), P(ec), P(out))
        assert np.abs(out[:, :2] - xy).max() < TOL and np.abs(out[:, 2] - psi).max() < TOL, name


@pytest.mark.gpu
def test_device_conversion_matches_reference_golden(crb):
    for name, pat, lap, s, ey, xy, psi in _tracks():
        x, y, p = crb.curv_to_glob_batch(s, ey, pat, lap)
        assert np.abs(x - xy[:, 0]).max() < TOL and np.abs(y - xy[:, 1]).max() < TOL and np.abs(p - psi).max() < TOL, name
        # shaped input (candidates x stages), as the get_local_traj drop-in passes it
        n = (s.shape[0] // 11) * 11
        x2, y2, _ = crb.curv_to_glob_batch(s[:n].reshape(-1, 11), ey[:n].reshape(-1, 11), pat, lap)
        assert x2.shape == (n // 11, 11) and np.array_equal(x2.ravel(), x[:n]) and np.array_equal(y2.ravel(), y[:n])
    with pytest.raises(ValueError):
        crb.curv_to_glob_batch(s, ey[:5], pat, lap)


@pytest.mark.gpu
def test_planner_plot_copies_use_one_launch(crb):
    """planning._traj_xglob (the get_local_traj drop-in's global-frame copies) on a track that carries the reference's
    point_and_tangent table: every point equals the scalar reference function (restatement), s beyond the lap is wrapped."""
    import types
    from car_racing_b200 import planning
    name, pat, lap, s, ey, xy, psi = list(_tracks())[2]
    track = types.SimpleNamespace(point_and_tangent=pat, lap_length=lap)
    rng = np.random.default_rng(3)
    stack = np.zeros((7, 11, 6))
    stack[:, :, 4] = rng.uniform(0.0, 1.4 * lap, size=(7, 11))
    stack[:, :, 5] = rng.uniform(-0.8, 0.8, size=(7, 11))
    h0 = batch.default_handle()
    n0 = h0.launch_count if hasattr(h0, "launch_count") else None
    glob = planning._traj_xglob(stack, track)
    assert glob.shape == stack.shape and (glob[:, :, :4] == 0).all()
    for c in range(7):
        for j in range(11):
            x, y, _ = frenet_numpy.curv_to_glob(lap, pat, stack[c, j, 4], stack[c, j, 5])
            assert abs(glob[c, j, 4] - x) < TOL and abs(glob[c, j, 5] - y) < TOL
    if n0 is not None:
        assert h0.launch_count == n0 + 1

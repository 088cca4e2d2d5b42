"""Prediction of rivals with dynamics: offboard.DynamicBicycleModel.get_trajectory_nsteps
(racing/offboard.py:80-94).  The numpy restatement against the unmodified reference's output
(tests/golden/rollout_golden.npz, made by make_rollout_golden.py), the kernel body compiled for the host against the same
golden (CPU suite), the CUDA build through the C-ABI (-m gpu).  Floating point: 1e-12 (sin/cos of the device library and
FMA contraction against numpy's; 21 steps)."""
import ctypes as C
import os
import subprocess
import types

import numpy as np
import pytest

import rollout_numpy
from car_racing_b200 import _capi, batch, rivals

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "rollout_golden.npz")
TOL = 1e-12


def _sets():
    g = np.load(GOLD)
    for name in ("ellipse", "l_shape", "goggle"):
        for n in (11, 21):
            k = lambda q, name=name, n=n: g["%s/n%d/%s" % (name, n, q)]   # noqa: E731
            yield name, n, g[name + "/pat"], float(g[name + "/lap_length"]), k("xcurv0"), k("xglob0"), k("xcurv"), k("xglob")


def test_restatement_matches_reference_golden():
    wrapped = 0
    for name, n, pat, lap, xc0, xg0, xc, xg in _sets():
        for b in range(xc0.shape[0]):
            rc, rg = rollout_numpy.rollout(xc0[b], xg0[b], pat[:, 3:6], lap, 0.1, n)
            assert np.array_equal(rc, xc[b]) and np.array_equal(rg, xg[b]), (name, n, b)
        wrapped += int((xc[:, 4, -1] < xc0[:, 4]).sum())
        assert (xg[:, 5] == 0.0).all()         # offboard.py:71-76: the global Y row is never filled
    assert wrapped >= 6                        # the lap wrap (:89-90) is exercised


def test_kernel_body_compiled_for_host_matches_reference_golden():
    emu = os.path.join(HERE, "host_emulation")
    lib = os.path.join(emu, "_build", "librollout_emu.so")
    src = [os.path.join(emu, "rival_rollout_host.cpp"), os.path.join(emu, "cuda_runtime.h"),
           os.path.join(HERE, "..", "car_racing_b200", "csrc", "rival_rollout.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(f) > os.path.getmtime(lib) for f in src):
        os.makedirs(os.path.dirname(lib), exist_ok=True)
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I", emu, src[0], "-o", lib], check=True)
    L = C.CDLL(lib)
    L.emu_rival_rollout.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    L.emu_rival_rollout.restype = None
    for name, n, pat, lap, xc0, xg0, xc, xg in _sets():
        seg = np.ascontiguousarray(pat[:, 3:6])
        p = _capi.RolloutParams()
        p.n, p.num_segments, p.timestep, p.lap_length = n, seg.shape[0], 0.1, lap
        oc, og = np.zeros_like(xc), np.zeros_like(xg)
        P = batch._ptr
        L.emu_rival_rollout(C.byref(p), xc0.shape[0], P(np.ascontiguousarray(xc0)), P(np.ascontiguousarray(xg0)), P(seg), P(oc), P(og))
        assert np.abs(oc - xc).max() < TOL and np.abs(og - xg).max() < TOL, (name, n)


@pytest.mark.gpu
def test_device_rollout_matches_reference_golden(crb):
    for name, n, pat, lap, xc0, xg0, xc, xg in _sets():
        oc, og = crb.rival_rollout_batch(xc0, xg0, pat, lap, 0.1, n, with_glob=True)
        assert np.abs(oc - xc).max() < TOL and np.abs(og - xg).max() < TOL, (name, n)
        only_c = crb.rival_rollout_batch(xc0, xg0, pat, lap, 0.1, n)
        assert np.array_equal(only_c, oc)
        # the method drop-in on a duck-typed rival (offboard.DynamicBicycleModel's attributes)
        me = types.SimpleNamespace(xcurv=xc0[3], xglob=xg0[3], point_and_tangent=pat, lap_length=lap, timestep=0.1)
        mc, mg = rivals.dynamic_get_trajectory_nsteps(me, n)
        assert mc.shape == (6, n) and np.abs(mc - xc[3]).max() < TOL and np.abs(mg - xg[3]).max() < TOL


@pytest.mark.gpu
def test_device_rollout_large_batch_and_bad_arguments(crb):
    name, n, pat, lap, xc0, xg0, xc, xg = next(_sets())
    reps = 400                                         # 4800 rivals: several CTAs, ragged last one
    big = crb.rival_rollout_batch(np.tile(xc0, (reps, 1))[:-5], np.tile(xg0, (reps, 1))[:-5], pat, lap, 0.1, n)
    assert big.shape == (12 * reps - 5, 6, n)
    assert np.abs(big[:12] - xc).max() < TOL and np.array_equal(big[12:24], big[:12]) and np.array_equal(big[-7:], big[:7])
    with pytest.raises(ValueError):
        crb.rival_rollout_batch(xc0, xg0[:3], pat, lap, 0.1, n)
    with pytest.raises(crb.B200MPCError):
        crb.rival_rollout_batch(xc0, xg0, pat, -1.0, 0.1, n)

"""CPU tests (-m "not gpu"): the oracle against the reference-derived golden vectors, against an
independent dense numpy implementation, and against a solver-independent KKT certificate."""
import os

import numpy as np
import pytest

from car_racing_b200 import scenarios

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_lti_literals_match_reference_csv():
    ref = "/root/reference/data/sys/LTI"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present (GPU box)")
    A = np.genfromtxt(ref + "/matrix_A.csv", delimiter=",")
    B = np.genfromtxt(ref + "/matrix_B.csv", delimiter=",")
    assert (A == scenarios.LTI_A).all() and (B == scenarios.LTI_B).all()


def test_ilqr_oracle_matches_reference_golden(oracle):
    """tests/golden/ilqr_golden.npz was produced by the UNMODIFIED reference control.ilqr
    (tests/golden/make_ilqr_golden.py).  The C restatement must reproduce it."""
    g = np.load(os.path.join(GOLD, "ilqr_golden.npz"))
    for N in sorted(set(g["N"].tolist())):
        idx = np.where(g["N"] == N)[0]
        x0, xt, obs, lap = g["x0"][idx], g["xt"][idx], g["obs"][idx][:, :, :N + 1], g["lap"][idx]
        lap_off = (np.trunc(x0[:, 4] / lap) - np.trunc(obs[:, 0, 0] / lap)) * lap
        prm = dict(A=g["A"], B=g["B"], Q=g["Q"], R=g["R"], N=int(N), max_iter=int(g["max_iter"]), L=0.4, W=0.2)
        r = oracle.solve_ilqr_batch(x0, xt, obs, lap_off, prm)
        assert np.abs(r["u0"] - g["u0"][idx]).max() < 1e-10


def test_mpc_lti_anchor(oracle):
    """SURVEY.md 8(c) restatement-derived anchor (two scipy methods agreed to 3e-8): MPC-LTI N=10,
    x0=0, vt=0.8, width 0.8 (tests/auto_control_test.py:9,13,27 of the reference)."""
    prm = scenarios.default_cbf_params(N=10, width=0.8)
    r = oracle.solve_cbf_batch(np.zeros((1, 6)), np.array([0.8, 0, 0, 0, 0, 0.0]), np.zeros((1, 0, 2, 11)), None, prm)
    assert r["status"][0] == 0
    assert abs(r["u0"][0, 0] - 0.0033384) < 1e-6 and abs(r["u0"][0, 1] - 1.0) < 1e-6
    assert abs(r["cost"][0] - 22.71576162) < 1e-6


def test_oracle_vs_numpy_second_opinion(oracle):
    """Same algorithm, independent linear algebra (full KKT matrix + LU / eigenvalue inertia)."""
    import ipm_numpy
    N, M = 10, 2
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(3, N=N, M=M, seed=5)
    prm = scenarios.default_cbf_params(N=N)
    r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    for b in range(3):
        P = ipm_numpy.CbfProblem(x0[b], xt, obs[b], prm["A"], prm["B"], prm["Q"], prm["R"], N)
        s = ipm_numpy.ipm_solve(P)
        assert s["status"] == 0 and r["status"][b] == 0
        assert np.abs(s["w"][P.iu(0)] - r["u0"][b]).max() < 1e-7
        assert abs(s["cost"] - r["cost"][b]) < 1e-7


def test_oracle_kkt_certificate(oracle):
    from kkt_check import certificate
    N, M = 20, 3
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(6, N=N, M=M, seed=3)
    prm = scenarios.default_cbf_params(N=N)
    r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm)
    for b in range(6):
        if r["status"][b] != 0 or r["elastic_max"][b] > 1e-7:
            continue
        c = certificate(x0[b], xt, obs[b], lap_off[b], prm, r["x"][b], r["u"][b], r["sigma"][b])
        assert c["dyn"] < 1e-9 and c["row_viol"] < 1e-6 and c["bound_viol"] < 1e-9
        assert c["stat"] < 1e-4 and c["comp"] < 1e-4, c
        assert abs(c["cost"] - r["cost"][b]) < 1e-8


def test_oracle_robustness_config2(oracle):
    x0, xt, obs, lap_off = scenarios.mpccbf_scenarios(256, N=20, M=3, seed=1)
    prm = scenarios.default_cbf_params(N=20)
    r = oracle.solve_cbf_batch(x0, xt, obs, lap_off, prm, nthreads=os.cpu_count() or 1)
    assert (r["status"] == 0).mean() >= 0.99
    ok = r["status"] == 0
    assert r["kkt_err"][ok].max() <= 1e-6

"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product package never does.  See ocp_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NMAX, MMAX = 64, 8


class Problem(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("M", C.c_int),
        ("A", C.c_double * 36), ("B", C.c_double * 12),
        ("Q", C.c_double * 36), ("R", C.c_double * 4),
        ("x0", C.c_double * 6),
        ("xt", C.c_double * ((NMAX + 1) * 6)),
        ("umax", C.c_double * 2),
        ("vmin", C.c_double), ("vmax", C.c_double), ("width", C.c_double),
        ("alpha", C.c_double), ("margin", C.c_double), ("L", C.c_double), ("W", C.c_double),
        ("slack_w", C.c_double),
        ("obs_s", (C.c_double * (NMAX + 1)) * MMAX),
        ("obs_ey", (C.c_double * (NMAX + 1)) * MMAX),
        ("lap_off", C.c_double * MMAX),
        ("per_stage_bounds", C.c_int),
        ("xlb", C.c_double * ((NMAX + 1) * 2)), ("xub", C.c_double * ((NMAX + 1) * 2)),
        ("wd", C.c_double * NMAX),
        ("per_rival_size", C.c_int), ("Lj", C.c_double * MMAX), ("Wj", C.c_double * MMAX),
    ]


class Options(C.Structure):
    _fields_ = [
        ("tol", C.c_double), ("max_iter", C.c_int), ("mu_init", C.c_double), ("rho", C.c_double),
        ("bound_push", C.c_double), ("bound_frac", C.c_double), ("acceptable_tol", C.c_double),
        ("acceptable_iter", C.c_int), ("max_grad", C.c_double), ("start", C.c_int), ("max_reset", C.c_int),
    ]


class Result(C.Structure):
    _fields_ = [
        ("x", C.c_double * ((NMAX + 1) * 6)), ("u", C.c_double * (NMAX * 2)),
        ("sigma", C.c_double * (MMAX * (NMAX + 1))),
        ("cost", C.c_double), ("kkt_err", C.c_double), ("elastic_max", C.c_double),
        ("status", C.c_int), ("iters", C.c_int), ("n_refactor", C.c_int), ("n_backtrack", C.c_int),
        ("lam", C.c_double * (NMAX * 6)), ("y", C.c_double * (MMAX * NMAX)), ("df", C.c_double),
    ]


class IlqrProblem(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("max_iter", C.c_int),
        ("A", C.c_double * 36), ("B", C.c_double * 12), ("Q", C.c_double * 36), ("R", C.c_double * 4),
        ("x0", C.c_double * 6), ("xt", C.c_double * 6),
        ("obs_s", C.c_double * (NMAX + 1)), ("obs_ey", C.c_double * (NMAX + 1)),
        ("lap_off", C.c_double), ("L", C.c_double), ("W", C.c_double),
    ]


class IlqrResult(C.Structure):
    _fields_ = [
        ("u0", C.c_double * 2), ("cost", C.c_double), ("iters", C.c_int), ("converged", C.c_int),
        ("u", C.c_double * (NMAX * 2)), ("x", C.c_double * ((NMAX + 1) * 6)),
    ]


KMAX = 64


class LmpcProblem(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("K", C.c_int),
        ("Q", C.c_double * 36), ("R", C.c_double * 4), ("dR", C.c_double * 4),
        ("xtrk", C.c_double * 6), ("umax", C.c_double * 2), ("vmax", C.c_double), ("width", C.c_double),
        ("x0", C.c_double * 6), ("u_old", C.c_double * 2),
        ("A", C.c_double * (NMAX * 36)), ("B", C.c_double * (NMAX * 12)), ("C", C.c_double * (NMAX * 6)),
        ("SS", C.c_double * (6 * KMAX)), ("Qfun", C.c_double * KMAX),
    ]


class LmpcResult(C.Structure):
    _fields_ = [
        ("x", C.c_double * ((NMAX + 1) * 6)), ("u", C.c_double * (NMAX * 2)), ("lam", C.c_double * KMAX),
        ("cost", C.c_double), ("kkt_err", C.c_double), ("status", C.c_int), ("iters", C.c_int),
    ]


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _lib
    if _lib is None:
        so = os.environ.get("B200MPC_ORACLE_SO") or os.path.join(_HERE, "_build", "liboracle.so")   # override: experiments only
        if not os.path.exists(so):
            build()
        _lib = C.CDLL(so)
        _lib.orc_default_options.argtypes = [C.POINTER(Options)]
        _lib.orc_solve_batch.argtypes = [C.POINTER(Problem), C.c_int, C.POINTER(Options), C.POINTER(Result), C.c_int]
        _lib.orc_ilqr_solve_batch.argtypes = [C.POINTER(IlqrProblem), C.c_int, C.POINTER(IlqrResult), C.c_int]
        _lib.orc_lmpc_solve_batch.argtypes = [C.POINTER(LmpcProblem), C.c_int, C.POINTER(Options), C.POINTER(LmpcResult), C.c_int]
    return _lib


def default_options(**kw):
    o = Options()
    lib().orc_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _fill(dst, src):
    a = np.ascontiguousarray(src, dtype=np.float64).ravel()
    C.memmove(dst, a.ctypes.data, a.nbytes)


def solve_cbf_batch(x0, xt, obs, lap_off, prm, nthreads=1, xlb=None, xub=None, wd=None, sizes=None, **opt):
    """x0 (B,6); xt (B,N+1,6) or (6,); obs (B,M,2,N+1) [s, ey]; lap_off (B,M); prm: dict of
    model/limits (A,B,Q,R,N,umax,vmin,vmax,width,alpha,margin,L,W,slack_w).  Returns dict of arrays."""
    x0 = np.atleast_2d(np.asarray(x0, float))
    Bn = x0.shape[0]
    N = int(prm["N"])
    obs = np.asarray(obs, float).reshape(Bn, -1, 2, N + 1) if np.size(obs) else np.zeros((Bn, 0, 2, N + 1))
    M = obs.shape[1]
    xt = np.asarray(xt, float)
    if xt.ndim == 1:
        xt = np.broadcast_to(xt, (Bn, N + 1, 6))
    lap_off = np.zeros((Bn, M)) if lap_off is None else np.asarray(lap_off, float).reshape(Bn, M)
    P = (Problem * Bn)()
    for b in range(Bn):
        p = P[b]
        p.N, p.M = N, M
        _fill(p.A, prm["A"]); _fill(p.B, prm["B"]); _fill(p.Q, prm["Q"]); _fill(p.R, prm["R"])
        _fill(p.x0, x0[b]); _fill(p.xt, xt[b])
        _fill(p.umax, prm["umax"])
        p.vmin, p.vmax, p.width = prm["vmin"], prm["vmax"], prm["width"]
        p.alpha, p.margin, p.L, p.W, p.slack_w = prm["alpha"], prm["margin"], prm["L"], prm["W"], prm["slack_w"]
        for j in range(M):
            _fill(p.obs_s[j], obs[b, j, 0]); _fill(p.obs_ey[j], obs[b, j, 1])
            p.lap_off[j] = lap_off[b, j]
        if xlb is not None:
            p.per_stage_bounds = 1
            _fill(p.xlb, np.asarray(xlb, float).reshape(Bn, N + 1, 2)[b])
            _fill(p.xub, np.asarray(xub, float).reshape(Bn, N + 1, 2)[b])
        if wd is not None:
            _fill(p.wd, np.asarray(wd, float).reshape(Bn, N)[b])
        if sizes is not None and M > 0:
            sz = np.broadcast_to(np.asarray(sizes, float), (Bn, M, 2))[b]
            p.per_rival_size = 1
            for j in range(M):
                p.Lj[j], p.Wj[j] = sz[j, 0], sz[j, 1]
    o = default_options(**opt)
    R = (Result * Bn)()
    lib().orc_solve_batch(P, Bn, C.byref(o), R, nthreads)
    out = dict(
        x=np.array([np.frombuffer(r.x, dtype=np.float64)[: 6 * (N + 1)].reshape(N + 1, 6) for r in R]),
        u=np.array([np.frombuffer(r.u, dtype=np.float64)[: 2 * N].reshape(N, 2) for r in R]),
        sigma=np.array([np.frombuffer(r.sigma, dtype=np.float64)[: M * (N + 1)].reshape(M, N + 1) for r in R]),
        lam=np.array([np.frombuffer(r.lam, dtype=np.float64)[: 6 * N].reshape(N, 6) for r in R]),
        y=np.array([np.frombuffer(r.y, dtype=np.float64)[: M * N].reshape(M, N) for r in R]),
        cost=np.array([r.cost for r in R]), kkt_err=np.array([r.kkt_err for r in R]),
        elastic_max=np.array([r.elastic_max for r in R]),
        status=np.array([r.status for r in R]), iters=np.array([r.iters for r in R]),
        n_refactor=np.array([r.n_refactor for r in R]), n_backtrack=np.array([r.n_backtrack for r in R]),
    )
    out["u0"] = out["u"][:, 0, :].copy()
    return out


def solve_ilqr_batch(x0, xt, obs, lap_off, prm, nthreads=1):
    """x0 (B,6); xt (B,6)|(6,); obs (B,2,N+1) [s,ey] of the rival used by control.ilqr; prm: A,B,Q,R,N,max_iter,L,W."""
    x0 = np.atleast_2d(np.asarray(x0, float))
    Bn = x0.shape[0]
    N = int(prm["N"])
    xt = np.broadcast_to(np.asarray(xt, float), (Bn, 6))
    obs = np.asarray(obs, float).reshape(Bn, 2, N + 1)
    lap_off = np.zeros(Bn) if lap_off is None else np.asarray(lap_off, float).reshape(Bn)
    P = (IlqrProblem * Bn)()
    for b in range(Bn):
        p = P[b]
        p.N, p.max_iter = N, int(prm["max_iter"])
        _fill(p.A, prm["A"]); _fill(p.B, prm["B"]); _fill(p.Q, prm["Q"]); _fill(p.R, prm["R"])
        _fill(p.x0, x0[b]); _fill(p.xt, xt[b])
        _fill(p.obs_s, obs[b, 0]); _fill(p.obs_ey, obs[b, 1])
        p.lap_off, p.L, p.W = lap_off[b], prm["L"], prm["W"]
    R = (IlqrResult * Bn)()
    lib().orc_ilqr_solve_batch(P, Bn, R, nthreads)
    return dict(
        u0=np.array([[r.u0[0], r.u0[1]] for r in R]), cost=np.array([r.cost for r in R]),
        iters=np.array([r.iters for r in R]), converged=np.array([r.converged for r in R]),
        u=np.array([np.frombuffer(r.u, dtype=np.float64)[: 2 * N].reshape(N, 2) for r in R]),
        x=np.array([np.frombuffer(r.x, dtype=np.float64)[: 6 * (N + 1)].reshape(N + 1, 6) for r in R]),
    )


def solve_lmpc_batch(x0, u_old, A, B, Cm, SS, Qfun, prm, nthreads=1, **opt):
    """control.lmpc's QP.  x0 (Bn,6); u_old (Bn,2); A (Bn,N,6,6); B (Bn,N,6,2); Cm (Bn,N,6); SS (Bn,6,K); Qfun (Bn,K);
    prm: Q, R, dR, N, umax, vmax, width, xtrk."""
    x0 = np.atleast_2d(np.asarray(x0, float))
    Bn = x0.shape[0]
    N = int(prm["N"])
    SS = np.asarray(SS, float).reshape(Bn, 6, -1)
    K = SS.shape[2]
    A = np.asarray(A, float).reshape(Bn, N, 36); B = np.asarray(B, float).reshape(Bn, N, 12); Cm = np.asarray(Cm, float).reshape(Bn, N, 6)
    u_old = np.asarray(u_old, float).reshape(Bn, 2); Qfun = np.asarray(Qfun, float).reshape(Bn, K)
    P = (LmpcProblem * Bn)()
    for b in range(Bn):
        p = P[b]
        p.N, p.K = N, K
        _fill(p.Q, prm["Q"]); _fill(p.R, prm["R"]); _fill(p.dR, prm["dR"]); _fill(p.xtrk, prm["xtrk"]); _fill(p.umax, prm["umax"])
        p.vmax, p.width = prm["vmax"], prm["width"]
        _fill(p.x0, x0[b]); _fill(p.u_old, u_old[b])
        _fill(p.A, A[b]); _fill(p.B, B[b]); _fill(p.C, Cm[b]); _fill(p.SS, SS[b]); _fill(p.Qfun, Qfun[b])
    o = default_options(**opt)
    R = (LmpcResult * Bn)()
    lib().orc_lmpc_solve_batch(P, Bn, C.byref(o), R, nthreads)
    return dict(
        x=np.array([np.frombuffer(r.x, dtype=np.float64)[: 6 * (N + 1)].reshape(N + 1, 6) for r in R]),
        u=np.array([np.frombuffer(r.u, dtype=np.float64)[: 2 * N].reshape(N, 2) for r in R]),
        u0=np.array([np.frombuffer(r.u, dtype=np.float64)[:2].copy() for r in R]),
        lam=np.array([np.frombuffer(r.lam, dtype=np.float64)[:K] for r in R]),
        cost=np.array([r.cost for r in R]), kkt_err=np.array([r.kkt_err for r in R]),
        status=np.array([r.status for r in R]), iters=np.array([r.iters for r in R]))

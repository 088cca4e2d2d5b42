"""Solver-independent KKT certificate for the MPC-CBF NLP (test infrastructure).

Given a primal point (x, u, sigma) it (1) checks primal feasibility, (2) recovers multipliers by
non-negative least squares on the active set and (3) reports the stationarity residual.  Uses the
dense numpy statement of the problem in oracle/ipm_numpy.py.
"""
import numpy as np
from scipy.optimize import nnls

from ipm_numpy import CbfProblem


def certificate(x0, xt, obs, lap_off, prm, x, u, sigma, act_tol=1e-3):
    N = prm["N"]
    P = CbfProblem(x0, np.asarray(xt, float).reshape(-1)[:6], obs, prm["A"], prm["B"], prm["Q"], prm["R"], N,
                   alpha=prm["alpha"], margin=prm["margin"], umax=prm["umax"], vmin=prm["vmin"], vmax=prm["vmax"],
                   width=prm["width"], L=prm["L"], W=prm["W"], lap_off=lap_off, slack_w=prm["slack_w"])
    w = np.concatenate([np.asarray(x)[1:].ravel(), np.asarray(u).ravel(), np.asarray(sigma).ravel()])
    out = {}
    out["dyn"] = np.abs(P.c(w)).max()
    g = P.g(w) if P.m else np.zeros(0)
    out["row_viol"] = max(0.0, -g.min()) if P.m else 0.0
    out["bound_viol"] = max(0.0, (P.lbw - w).max(), (w - P.ubw).max())
    # active set
    gscale = np.maximum(1.0, np.abs(P.Jg(w)).max(axis=1)) if P.m else np.zeros(0)
    act_rows = np.where(g / gscale <= act_tol)[0] if P.m else np.zeros(0, int)
    act_lb = np.where(w - P.lbw <= act_tol)[0]
    act_ub = np.where(P.ubw - w <= act_tol)[0]
    # stationarity: grad f + Jc' lam - Jg_A' y - e_lb zl + e_ub zu = 0, y, zl, zu >= 0, lam free
    Jg = P.Jg(w) if P.m else np.zeros((0, P.n))
    cols = [P.Jc.T, -P.Jc.T]
    cols.append(-Jg[act_rows].T)
    E = np.eye(P.n)
    cols.append(-E[:, act_lb])
    cols.append(E[:, act_ub])
    Amat = np.hstack(cols)
    gr = P.grad(w)
    scale = np.maximum(1.0, np.abs(Amat).max(axis=0))
    sol, rn = nnls(Amat / scale, -gr, maxiter=20 * Amat.shape[1])
    out["stat"] = np.abs(Amat / scale @ sol + gr).max()
    # complementarity of the recovered multipliers (interior-point solutions sit ~mu/z off the bound)
    mult = sol / scale
    k = 2 * P.Jc.shape[0]
    y = mult[k:k + len(act_rows)]
    zl = mult[k + len(act_rows):k + len(act_rows) + len(act_lb)]
    zu = mult[k + len(act_rows) + len(act_lb):]
    comp = 0.0
    if len(act_rows):
        comp = max(comp, np.abs(y * g[act_rows]).max())
    if len(act_lb):
        comp = max(comp, np.abs(zl * (w - P.lbw)[act_lb]).max())
    if len(act_ub):
        comp = max(comp, np.abs(zu * (P.ubw - w)[act_ub]).max())
    out["comp"] = comp
    out["cost"] = P.f(w)
    return out

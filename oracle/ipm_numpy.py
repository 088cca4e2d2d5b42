"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's MPC NLPs and a dense
IPOPT-style interior point method.  Second opinion for oracle/ocp_oracle.c (dense full
KKT matrix, numpy.linalg) -- slow, used on a handful of instances per test.

Problem statement follows car_racing/control/control.py:476-607 (mpccbf) /
:198-248 (mpc_lti); algorithm follows the published IPOPT method (Waechter & Biegler,
Math. Prog. 106(1) 2006), the solver bundled with the CasADi 3.5.5 wheel pinned in the
reference's requirements.txt:6.

PARITY: the solver algorithm is UNPINNED -- CasADi/IPOPT cannot be installed here (no network) and the
reference holds no golden vector for this path; the problem functions below (f, c, g, bounds)
ARE pinned to the reference's own code (tests/test_reference_statement.py).  Only tests/ may import this file.
See DESIGN.md "Solver definition" for the conventions (start point, bounds-as-bounds).
"""
import numpy as np

X_DIM, U_DIM = 6, 2


class CbfProblem:
    """control.mpccbf's NLP, x_0 eliminated (it is fixed by control.py:497).

    w = [x_1..x_N (6 each) | u_0..u_{N-1} (2 each) | sigma_{j,0..N} for j<M]
    c(w)=0 : x_{i+1} = A x_i + B u_i  (:566-570)
    simple bounds: u box (:572-576), vx/ey box on x_1..x_N (:582-586), sigma>=0 (:559,561)
    rows g(w)>=0: M*N CBF rows (:537-558)
    """

    def __init__(self, x0, xt, obs, A, B, Q, R, N, alpha=0.8, margin=0.2, degree=6,
                 umax=(0.5, 1.0), vmin=0.0, vmax=10.0, width=1.0, L=0.4, W=0.2,
                 lap_off=None, slack_w=1e4, sizes=None):
        self.N, self.M = N, obs.shape[0]
        M = self.M
        self.x0, self.xt, self.obs = np.asarray(x0, float), xt, obs  # obs (M,2,N+1): s, ey
        self.A, self.B, self.Q, self.R = A, B, Q, R
        self.alpha, self.margin, self.deg = alpha, margin, degree
        # per-rival (l_agent+l_obs, w_agent+w_obs) (control.py:530-535); one pair for all when not given
        self.Lj = np.full(self.M, float(L)) if sizes is None else np.asarray(sizes, float).reshape(self.M, 2)[:, 0]
        self.Wj = np.full(self.M, float(W)) if sizes is None else np.asarray(sizes, float).reshape(self.M, 2)[:, 1]
        # (num_cycle_ego-num_cycle_obs)*lap_length, applied to h but not h_next (:539-542)
        self.lap_off = np.zeros(M) if lap_off is None else np.asarray(lap_off, float)
        self.slack_w = slack_w
        self.start_mode = 'rollout'
        self.nx, self.nu, self.ns = 6 * N, 2 * N, M * (N + 1)
        self.n = self.nx + self.nu + self.ns
        me = 6 * N
        Jc = np.zeros((me, self.n))
        for i in range(N):
            r = 6 * i
            Jc[r:r + 6, self.ix(i + 1)] = np.eye(6)
            if i > 0:
                Jc[r:r + 6, self.ix(i)] = -A
            Jc[r:r + 6, self.iu(i)] = -B
        self.Jc = Jc
        self.lbw = np.full(self.n, -np.inf)
        self.ubw = np.full(self.n, np.inf)
        for i in range(1, N + 1):
            self.lbw[self.ix(i).start + 0], self.ubw[self.ix(i).start + 0] = vmin, vmax
            self.lbw[self.ix(i).start + 5], self.ubw[self.ix(i).start + 5] = -width, width
        for i in range(N):
            self.lbw[self.iu(i)] = [-umax[0], -umax[1]]
            self.ubw[self.iu(i)] = [umax[0], umax[1]]
        self.lbw[self.nx + self.nu:] = 0.0
        self.m = M * N

    def ix(self, i):  # i>=1
        return slice(6 * (i - 1), 6 * i)

    def iu(self, i):
        return slice(self.nx + 2 * i, self.nx + 2 * i + 2)

    def isg(self, j, i):
        return self.nx + self.nu + j * (self.N + 1) + i

    def xk(self, w, i):
        return self.x0 if i == 0 else w[self.ix(i)]

    def start(self):
        """start_mode "rollout": u = 0 roll-out from x0, sigma = 0 (solver start point);
        "zero": w = 0, what Opti hands IPOPT (the reference never calls opti.set_initial in control.py:476-607)."""
        w = np.zeros(self.n)
        if self.start_mode == "zero":
            return w
        x = self.x0
        for i in range(1, self.N + 1):
            x = self.A @ x
            w[self.ix(i)] = x
        return w

    def f(self, w):
        c = 0.0
        for i in range(self.N + 1):
            d = self.xk(w, i) - self.xt
            c += d @ self.Q @ d
        for i in range(self.N):
            u = w[self.iu(i)]
            c += u @ self.R @ u
        return c + self.slack_w * np.sum(w[self.nx + self.nu:])

    def grad(self, w):
        g = np.zeros(self.n)
        for i in range(1, self.N + 1):
            g[self.ix(i)] = 2 * self.Q @ (w[self.ix(i)] - self.xt)
        for i in range(self.N):
            g[self.iu(i)] = 2 * self.R @ w[self.iu(i)]
        g[self.nx + self.nu:] = self.slack_w
        return g

    def c(self, w):
        r = self.Jc @ w
        r[0:6] -= self.A @ self.x0
        return r

    def _h(self, w, j, i, off):
        x = self.xk(w, i)
        ds = x[4] - self.obs[j, 0, i] - off
        de = x[5] - self.obs[j, 1, i]
        p = self.deg
        return ds ** p / self.Lj[j] ** p + de ** p / self.Wj[j] ** p - 1 - self.margin - w[self.isg(j, i)], ds, de

    def g(self, w):
        out = np.zeros(self.m)
        r = 0
        for j in range(self.M):
            for i in range(self.N):
                h, _, _ = self._h(w, j, i, self.lap_off[j])
                hn, _, _ = self._h(w, j, i + 1, 0.0)
                out[r] = hn - h + self.alpha * h      # :558  h_next - h >= -alpha*h
                r += 1
        return out

    def Jg(self, w):
        J = np.zeros((self.m, self.n))
        p = self.deg
        a = 1.0 - self.alpha
        r = 0
        for j in range(self.M):
            for i in range(self.N):
                _, ds, de = self._h(w, j, i, self.lap_off[j])
                _, dsn, den = self._h(w, j, i + 1, 0.0)
                if i > 0:
                    J[r, self.ix(i).start + 4] = -a * p * ds ** (p - 1) / self.Lj[j] ** p
                    J[r, self.ix(i).start + 5] = -a * p * de ** (p - 1) / self.Wj[j] ** p
                J[r, self.isg(j, i)] = a
                J[r, self.ix(i + 1).start + 4] = p * dsn ** (p - 1) / self.Lj[j] ** p
                J[r, self.ix(i + 1).start + 5] = p * den ** (p - 1) / self.Wj[j] ** p
                J[r, self.isg(j, i + 1)] = -1.0
                r += 1
        return J

    def hessL(self, w, sig_f, y):
        """Hessian of sig_f*f - sum_r y_r g_r."""
        H = np.zeros((self.n, self.n))
        for i in range(1, self.N + 1):
            H[self.ix(i), self.ix(i)] += 2 * sig_f * self.Q
        for i in range(self.N):
            H[self.iu(i), self.iu(i)] += 2 * sig_f * self.R
        p = self.deg
        c2 = p * (p - 1)
        a = 1.0 - self.alpha
        r = 0
        for j in range(self.M):
            for i in range(self.N):
                _, ds, de = self._h(w, j, i, self.lap_off[j])
                _, dsn, den = self._h(w, j, i + 1, 0.0)
                if i > 0:
                    k = self.ix(i).start
                    H[k + 4, k + 4] += y[r] * a * c2 * ds ** (p - 2) / self.Lj[j] ** p
                    H[k + 5, k + 5] += y[r] * a * c2 * de ** (p - 2) / self.Wj[j] ** p
                k = self.ix(i + 1).start
                H[k + 4, k + 4] -= y[r] * c2 * dsn ** (p - 2) / self.Lj[j] ** p
                H[k + 5, k + 5] -= y[r] * c2 * den ** (p - 2) / self.Wj[j] ** p
                r += 1
        return H


OPTS = dict(tol=1e-8, max_iter=200, mu_init=0.1, kappa_eps=10.0, kappa_mu=0.2, theta_mu=1.5,
            tau_min=0.99, gamma_theta=1e-5, gamma_phi=1e-8, delta_sw=1.0, s_theta=1.1, s_phi=2.3,
            eta_phi=1e-8, gamma_alpha=0.05, s_max=100.0, bound_push=1e-2, bound_frac=1e-2,
            kappa_sigma=1e10, acceptable_tol=1e-6, acceptable_iter=15, max_grad=100.0,
            rho=1e3, elastic=True, max_reset=5)


def ipm_solve(P, verbose=False, **kw):
    """Primal-dual interior point (IPOPT conventions) with l1-elastic nonlinear rows.

    rows:  dg*g(w) + t - s = 0,  s>=0, t>=0, objective += rho*sum(t)   (elastic=True)
    """
    o = dict(OPTS)
    o.update(kw)
    tol = o["tol"]
    n, m = P.n, P.m
    Jc = P.Jc
    me = Jc.shape[0]
    hasL, hasU = np.isfinite(P.lbw), np.isfinite(P.ubw)
    lbw = np.where(hasL, P.lbw, 0.0)
    ubw = np.where(hasU, P.ubw, 0.0)
    w = P.start()
    pl = np.where(hasL, np.minimum(o["bound_push"] * np.maximum(1, np.abs(lbw)),
                                    o["bound_frac"] * np.where(hasL & hasU, ubw - lbw, np.inf)), 0)
    pu = np.where(hasU, np.minimum(o["bound_push"] * np.maximum(1, np.abs(ubw)),
                                    o["bound_frac"] * np.where(hasL & hasU, ubw - lbw, np.inf)), 0)
    w = np.where(hasL, np.maximum(w, lbw + pl), w)
    w = np.where(hasU, np.minimum(w, ubw - pu), w)
    gmax = np.max(np.abs(P.grad(w)))
    df = o["max_grad"] / gmax if gmax > o["max_grad"] else 1.0
    if m:
        rowmax = np.max(np.abs(P.Jg(w)), axis=1)
        dg = np.where(rowmax > o["max_grad"], o["max_grad"] / np.maximum(rowmax, 1e-300), 1.0)
    else:
        dg = np.zeros(0)
    gs = lambda w: dg * P.g(w)
    Jgs = lambda w: dg[:, None] * P.Jg(w)
    rho, el = o["rho"], o["elastic"]
    if el:
        t = np.maximum(0.0, -gs(w)) + o["bound_push"]
        s = gs(w) + t
        v = np.ones(m)
    else:
        t, v = np.zeros(m), np.zeros(m)
        s = np.maximum(gs(w), o["bound_push"])
    zL, zU = np.where(hasL, 1.0, 0.0), np.where(hasU, 1.0, 0.0)
    z, y, lam = np.ones(m), np.zeros(m), np.zeros(me)
    mu = o["mu_init"]
    dw_last = 0.0
    last_needed = False
    dl = lambda w: np.where(hasL, w - lbw, 1.0)
    du = lambda w: np.where(hasU, ubw - w, 1.0)
    nb = int(hasL.sum() + hasU.sum()) + m + (m if el else 0)

    def theta_of(w, s, t):
        return np.sum(np.abs(P.c(w))) + np.sum(np.abs(gs(w) + t - s))

    def phi_of(w, s, t, mu):
        b = np.sum(np.log(dl(w))[hasL]) + np.sum(np.log(du(w))[hasU]) + np.sum(np.log(s))
        if el:
            b += np.sum(np.log(t))
        return df * P.f(w) + (rho * np.sum(t) if el else 0.0) - mu * b

    def errors(w, s, t, lam, y, z, v, zL, zU, mu):
        rw = df * P.grad(w) + Jc.T @ lam - Jgs(w).T @ y - zL + zU
        rs = y - z
        rt = (rho - y - v) if el else np.zeros(0)
        zsum = z.sum() + v.sum() + zL.sum() + zU.sum()
        sd = max(o["s_max"], (np.abs(lam).sum() + np.abs(y).sum() + zsum) / max(me + m + nb, 1)) / o["s_max"]
        sc = max(o["s_max"], zsum / max(nb, 1)) / o["s_max"]
        dual = max(np.max(np.abs(rw)), np.max(np.abs(rs), initial=0.0), np.max(np.abs(rt), initial=0.0))
        prim = max(np.max(np.abs(P.c(w)), initial=0.0), np.max(np.abs(gs(w) + t - s), initial=0.0))
        comp = max(np.max(np.abs(s * z - mu), initial=0.0), np.max(np.abs(dl(w) * zL - mu)[hasL], initial=0.0),
                   np.max(np.abs(du(w) * zU - mu)[hasU], initial=0.0),
                   np.max(np.abs(t * v - mu), initial=0.0) if el else 0.0)
        return max(dual / sd, prim, comp / sc), dual, prim, comp

    th0 = theta_of(w, s, t)
    theta_max = 1e4 * max(1.0, th0)
    theta_min = 1e-4 * max(1.0, th0)
    filt = []
    it, status, n_acc, n_reset = 0, 1, 0, 0
    hist = []
    while True:
        E0, dual, prim, comp = errors(w, s, t, lam, y, z, v, zL, zU, 0.0)
        if verbose:
            print(f"it {it:3d} f={P.f(w):.8e} E0={E0:.2e} du={dual:.2e} pr={prim:.2e} co={comp:.2e} mu={mu:.1e} dw={dw_last:.1e} tmax={t.max() if m else 0:.1e}"
                  + (f" a={hist[-1][0]:.2e} az={hist[-1][1]:.2e} ls={hist[-1][3]}" if hist else ""))
        if E0 <= tol:
            status = 0
            break
        if E0 <= o["acceptable_tol"]:
            n_acc += 1
            if n_acc >= o["acceptable_iter"]:
                status = 0
                break
        else:
            n_acc = 0
        if it >= o["max_iter"]:
            status = 1
            break
        while True:
            Emu = errors(w, s, t, lam, y, z, v, zL, zU, mu)[0]
            if Emu <= o["kappa_eps"] * mu and mu > tol / 11.0:
                mu = max(tol / 11.0, min(o["kappa_mu"] * mu, mu ** o["theta_mu"]))
                filt = []
            else:
                break
        tau = max(o["tau_min"], 1.0 - mu)
        J = Jgs(w)
        Sig_s = z / s
        rg = gs(w) + t - s
        if el:
            Sig_t = v / t
            beta = Sig_t / (Sig_s + Sig_t)
            Sig_e = beta * Sig_s
            yhat = (1 - beta) * (rho - mu / t) + beta * (mu / s - Sig_s * rg)
        else:
            Sig_e = Sig_s
            yhat = mu / s - Sig_s * rg
        SigW = np.where(hasL, zL / dl(w), 0) + np.where(hasU, zU / du(w), 0)
        H = P.hessL(w, df, y * dg)
        rhs_w = -(df * P.grad(w)) + np.where(hasL, mu / dl(w), 0) - np.where(hasU, mu / du(w), 0) + J.T @ yhat
        cval = P.c(w)
        dw_try = max(1e-20, dw_last / 3.0) if last_needed else 0.0
        K0 = H + np.diag(SigW) + J.T @ (Sig_e[:, None] * J)
        while True:
            K = K0 + dw_try * np.eye(n)
            KKT = np.block([[K, Jc.T], [Jc, np.zeros((me, me))]])
            ev = np.linalg.eigvalsh(KKT)
            if np.sum(ev > 0) == n and np.sum(ev < 0) == me:
                break
            if dw_try == 0.0:
                dw_try = 1e-4 if dw_last == 0.0 else max(1e-20, dw_last / 3.0)
            else:
                dw_try = dw_try * (100.0 if dw_last == 0.0 else 8.0)
            if dw_try > 1e40:
                raise RuntimeError("inertia correction failed")
        if dw_try > 0:
            dw_last = dw_try
        last_needed = dw_try > 0
        sol = np.linalg.solve(KKT, np.concatenate([rhs_w, -cval]))
        dwv = sol[:n]
        dlam = sol[n:] - lam
        Jd = J @ dwv
        ynew = yhat - Sig_e * Jd
        dy = ynew - y
        if el:
            dt = (mu / s + mu / t - rho - Sig_s * rg - Sig_s * Jd) / (Sig_s + Sig_t)
            dv = mu / t - v - Sig_t * dt
        else:
            dt = np.zeros(m)
            dv = np.zeros(m)
        ds = Jd + dt + rg
        dz = mu / s - z - Sig_s * ds
        dzL = np.where(hasL, mu / dl(w) - zL - zL / dl(w) * dwv, 0)
        dzU = np.where(hasU, mu / du(w) - zU + zU / du(w) * dwv, 0)

        def ftb(val, dv_):
            neg = dv_ < 0
            return min(1.0, np.min(-tau * val[neg] / dv_[neg])) if np.any(neg) else 1.0
        a_max = min(ftb(s, ds), ftb(dl(w)[hasL], dwv[hasL]), ftb(du(w)[hasU], -dwv[hasU]), ftb(t, dt) if el else 1.0)
        a_z = min(ftb(z, dz), ftb(zL[hasL], dzL[hasL]), ftb(zU[hasU], dzU[hasU]), ftb(v, dv) if el else 1.0)
        th = theta_of(w, s, t)
        ph = phi_of(w, s, t, mu)
        gphi = df * P.grad(w) @ dwv - mu * (np.sum((dwv / dl(w))[hasL]) - np.sum((dwv / du(w))[hasU]) + np.sum(ds / s))
        if el:
            gphi += rho * np.sum(dt) - mu * np.sum(dt / t)
        if gphi < 0 and th <= theta_min:
            amin = o["gamma_alpha"] * min(o["gamma_theta"], o["gamma_phi"] * th / (-gphi),
                                          o["delta_sw"] * th ** o["s_theta"] / (-gphi) ** o["s_phi"])
        elif gphi < 0:
            amin = o["gamma_alpha"] * min(o["gamma_theta"], o["gamma_phi"] * th / (-gphi))
        else:
            amin = o["gamma_alpha"] * o["gamma_theta"]
        a = a_max
        accepted, ftype = False, False
        nls = 0
        while a >= amin or nls == 0:
            wt, st, tt = w + a * dwv, s + a * ds, t + a * dt
            tht, pht = theta_of(wt, st, tt), phi_of(wt, st, tt, mu)
            okf = tht < theta_max and all(not (tht >= ft and pht >= fp) for ft, fp in filt)
            if okf:
                sw = gphi < 0 and a * (-gphi) ** o["s_phi"] > o["delta_sw"] * th ** o["s_theta"]
                if th <= theta_min and sw:
                    if pht <= ph + o["eta_phi"] * a * gphi:
                        accepted, ftype = True, True
                elif tht <= (1 - o["gamma_theta"]) * th or pht <= ph - o["gamma_phi"] * th:
                    accepted = True
            if accepted:
                break
            a *= 0.5
            nls += 1
        if not accepted:
            # IPOPT would enter its restoration phase.  The rows are elastic: remove their residual exactly by enlarging
            # the slacks (t' = max(t, s-g), s' = g+t'), restart the filter (ocp_oracle.c does the same)
            if not el or n_reset >= o["max_reset"]:
                status = 2
                if verbose:
                    print("line search failed (restoration needed)", th, ph, gphi, a_max)
                break
            n_reset += 1
            gcur = gs(w)
            t = np.maximum(t, s - gcur)
            s = gcur + t
            filt = []
            it += 1
            continue
        if not ftype:
            filt.append(((1 - o["gamma_theta"]) * th, ph - o["gamma_phi"] * th))
        w, s, t = wt, st, tt
        lam, y = lam + a * dlam, y + a * dy
        z, zL, zU, v = z + a_z * dz, zL + a_z * dzL, zU + a_z * dzU, v + a_z * dv
        ks = o["kappa_sigma"]
        z = np.maximum(np.minimum(z, ks * mu / s), mu / (ks * s))
        if el:
            v = np.maximum(np.minimum(v, ks * mu / t), mu / (ks * t))
        zL = np.where(hasL, np.maximum(np.minimum(zL, ks * mu / dl(w)), mu / (ks * dl(w))), 0)
        zU = np.where(hasU, np.maximum(np.minimum(zU, ks * mu / du(w)), mu / (ks * du(w))), 0)
        hist.append((a, a_z, dw_try, nls))
        it += 1
    return dict(w=w, s=s, t=t, lam=lam, y=y, z=z, v=v, zL=zL, zU=zU, status=status, iters=it, cost=P.f(w), mu=mu,
                df=df, dg=dg, hist=hist)


def load_AB(path="/root/reference/data/sys/LTI"):
    A = np.genfromtxt(path + "/matrix_A.csv", delimiter=",")
    B = np.genfromtxt(path + "/matrix_B.csv", delimiter=",")
    return A, B

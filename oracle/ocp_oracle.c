/* TEST INFRASTRUCTURE ONLY -- see ocp_oracle.h for scope and the "parity unpinned" note.
 *
 * Dense restatement of the reference's MPC-LTI / MPC-CBF NLP and of the interior point
 * method used to solve it.  Linear algebra here is deliberately NOT the product's
 * (Riccati): the Newton step is computed by the null-space method -- states are condensed
 * out through x = Gamma*u + x_p, the reduced Hessian Z^T K Z is formed densely and
 * factorised by Cholesky; a failed Cholesky is the inertia test.
 *
 * NLP (x_0 eliminated, it is fixed by control.py:497):
 *   w = [x_1..x_N | u_0..u_{N-1} | sigma_{j,0..N}]
 *   min  sum_i (x_i-xt_i)'Q(x_i-xt_i) + sum_i u_i'Ru_i + slack_w*sum sigma     (:560-562,:578-591)
 *   s.t. x_{i+1} = A x_i + B u_i                                               (:566-570)
 *        |u| <= umax, vmin<=vx_i<=vmax, |ey_i|<=width (i>=1), sigma>=0         (:572-586,:559,:561)
 *        g_{j,i} = hn_j(x_{i+1},sig_{i+1}) - (1-alpha) h_j(x_i,sig_i) >= 0      (:537-558)
 * Algorithm (DESIGN.md "Solver definition"): IPOPT's primal-dual barrier method with
 * monotone mu, fraction-to-boundary, filter line search, inertia correction, gradient-based
 * scaling and E_0<=1e-8 termination; the nonlinear rows are l1-elastic
 * (g + t - s = 0, s,t>=0, + rho*t) instead of IPOPT's restoration phase; start point is the
 * u=0 roll-out from x_0.
 */
#include "ocp_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define FILT_MAX 64

typedef struct {
    const orc_problem *p;
    int N, M, nx, nu, ns, n, m, nr;
    double df;
    double *dg;                    /* row scaling (m) */
    double *lb, *ub;               /* simple bounds on w (n), +-HUGE_VAL when absent */
    double AB[ORC_NMAX][12];       /* A^p B */
} ctx_t;

static inline int IX(int i) { return 6 * (i - 1); }
static inline int IU(const ctx_t *c, int i) { return c->nx + 2 * i; }
static inline int ISG(const ctx_t *c, int j, int i) { return c->nx + c->nu + j * (c->N + 1) + i; }
static inline double XK(const ctx_t *c, const double *w, int i, int k) { return i == 0 ? c->p->x0[k] : w[IX(i) + k]; }

static double pow6(double a) { double a2 = a * a; return a2 * a2 * a2; }
static double pow5(double a) { double a2 = a * a; return a2 * a2 * a; }
static double pow4(double a) { double a2 = a * a; return a2 * a2; }

/* unscaled objective */
static double eval_f(const ctx_t *c, const double *w) {
    const orc_problem *p = c->p;
    double f = 0.0;
    for (int i = 0; i <= c->N; i++) {
        double d[6];
        for (int k = 0; k < 6; k++) d[k] = XK(c, w, i, k) - p->xt[6 * i + k];
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) f += d[a] * p->Q[6 * a + b] * d[b];
    }
    for (int i = 0; i < c->N; i++) {
        const double *u = w + IU(c, i);
        f += u[0] * (p->R[0] * u[0] + p->R[1] * u[1]) + u[1] * (p->R[2] * u[0] + p->R[3] * u[1]);
    }
    double ss = 0.0;
    for (int k = 0; k < c->ns; k++) ss += w[c->nx + c->nu + k];
    for (int i = 0; i < c->N; i++) {   /* planner: 30*(ey_{i+1}-ey_i)^2 (overtake_traj_planner.py:325-327) */
        double d = w[IX(i + 1) + 5] - XK(c, w, i, 5);
        f += p->wd[i] * d * d;
    }
    return f + p->slack_w * ss;
}

static void eval_grad(const ctx_t *c, const double *w, double *g) {
    const orc_problem *p = c->p;
    for (int i = 1; i <= c->N; i++) {
        double d[6];
        for (int k = 0; k < 6; k++) d[k] = w[IX(i) + k] - p->xt[6 * i + k];
        for (int a = 0; a < 6; a++) {
            double s = 0.0;
            for (int b = 0; b < 6; b++) s += (p->Q[6 * a + b] + p->Q[6 * b + a]) * d[b];
            g[IX(i) + a] = s;
        }
    }
    for (int i = 0; i < c->N; i++) {
        const double *u = w + IU(c, i);
        g[IU(c, i) + 0] = 2 * p->R[0] * u[0] + (p->R[1] + p->R[2]) * u[1];
        g[IU(c, i) + 1] = (p->R[1] + p->R[2]) * u[0] + 2 * p->R[3] * u[1];
    }
    for (int k = 0; k < c->ns; k++) g[c->nx + c->nu + k] = p->slack_w;
    for (int i = 0; i < c->N; i++) {
        double d = w[IX(i + 1) + 5] - XK(c, w, i, 5);
        g[IX(i + 1) + 5] += 2.0 * p->wd[i] * d;
        if (i > 0) g[IX(i) + 5] -= 2.0 * p->wd[i] * d;
    }
}

/* dynamics residual c_i = x_{i+1} - A x_i - B u_i, i=0..N-1 */
static void eval_c(const ctx_t *c, const double *w, double *r) {
    const orc_problem *p = c->p;
    for (int i = 0; i < c->N; i++) {
        const double *u = w + IU(c, i);
        for (int a = 0; a < 6; a++) {
            double s = w[IX(i + 1) + a];
            for (int b = 0; b < 6; b++) s -= p->A[6 * a + b] * XK(c, w, i, b);
            s -= p->B[2 * a] * u[0] + p->B[2 * a + 1] * u[1];
            r[6 * i + a] = s;
        }
    }
}

/* CBF rows, scaled by dg: r = j*N+i.  jac (optional): 6 entries per row
 * [d/ds_i, d/dey_i, d/dsig_i, d/ds_{i+1}, d/dey_{i+1}, d/dsig_{i+1}] (scaled);
 * hes (optional): 4 second derivatives [ss_i, ee_i, ss_{i+1}, ee_{i+1}] (scaled). */
static void eval_rows(const ctx_t *c, const double *w, double *g, double *jac, double *hes) {
    const orc_problem *p = c->p;
    double a = 1.0 - p->alpha;
    for (int j = 0; j < c->M; j++) {
        double iL6 = 1.0 / pow6(p->per_rival_size ? p->Lj[j] : p->L), iW6 = 1.0 / pow6(p->per_rival_size ? p->Wj[j] : p->W);
        for (int i = 0; i < c->N; i++) {
            int r = j * c->N + i;
            double ds = XK(c, w, i, 4) - p->obs_s[j][i] - p->lap_off[j];
            double de = XK(c, w, i, 5) - p->obs_ey[j][i];
            double dsn = w[IX(i + 1) + 4] - p->obs_s[j][i + 1];
            double den = w[IX(i + 1) + 5] - p->obs_ey[j][i + 1];
            double h = pow6(ds) * iL6 + pow6(de) * iW6 - 1.0 - p->margin - w[ISG(c, j, i)];
            double hn = pow6(dsn) * iL6 + pow6(den) * iW6 - 1.0 - p->margin - w[ISG(c, j, i + 1)];
            double sc = c->dg ? c->dg[r] : 1.0;
            if (g) g[r] = sc * (hn - a * h);
            if (jac) {
                double *q = jac + 6 * r;
                q[0] = sc * (-a * 6.0 * pow5(ds) * iL6);
                q[1] = sc * (-a * 6.0 * pow5(de) * iW6);
                q[2] = sc * a;
                q[3] = sc * (6.0 * pow5(dsn) * iL6);
                q[4] = sc * (6.0 * pow5(den) * iW6);
                q[5] = -sc;
            }
            if (hes) {
                double *q = hes + 4 * r;
                q[0] = sc * (-a * 30.0 * pow4(ds) * iL6);
                q[1] = sc * (-a * 30.0 * pow4(de) * iW6);
                q[2] = sc * (30.0 * pow4(dsn) * iL6);
                q[3] = sc * (30.0 * pow4(den) * iW6);
            }
        }
    }
}

/* columns of w touched by row r, in the order of the jac entries; -1: x_0 (not a variable) */
static void row_cols(const ctx_t *c, int r, int *col) {
    int j = r / c->N, i = r % c->N;
    col[0] = i > 0 ? IX(i) + 4 : -1;
    col[1] = i > 0 ? IX(i) + 5 : -1;
    col[2] = ISG(c, j, i);
    col[3] = IX(i + 1) + 4;
    col[4] = IX(i + 1) + 5;
    col[5] = ISG(c, j, i + 1);
}

/* in-place Cholesky of the lower triangle; returns 0 on success, 1 if a pivot <= 0 */
static int cholesky(double *a, int n) {
    for (int j = 0; j < n; j++) {
        double d = a[j * n + j];
        for (int k = 0; k < j; k++) d -= a[j * n + k] * a[j * n + k];
        if (!(d > 0.0)) return 1;
        d = sqrt(d);
        a[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = a[i * n + j];
            for (int k = 0; k < j; k++) s -= a[i * n + k] * a[j * n + k];
            a[i * n + j] = s / d;
        }
    }
    return 0;
}

static void chol_solve(const double *l, int n, double *b) {
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= l[i * n + k] * b[k];
        b[i] = s / l[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < n; k++) s -= l[k * n + i] * b[k];
        b[i] = s / l[i * n + i];
    }
}

typedef struct {
    double *w, *s, *t, *lam, *y, *z, *v, *zL, *zU;
} iter_t;

typedef struct {
    double E, dual, prim, comp;
} err_t;

static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }

static int has_l(const ctx_t *c, int k) { return c->lb[k] > -HUGE_VAL; }
static int has_u(const ctx_t *c, int k) { return c->ub[k] < HUGE_VAL; }

/* grad_w of the Lagrangian: df*grad f + Jc' lam - J' y - zL + zU */
static void lag_grad(const ctx_t *c, const iter_t *it, const double *gradf, const double *jac, double *rw) {
    const orc_problem *p = c->p;
    int N = c->N;
    for (int k = 0; k < c->n; k++) rw[k] = c->df * gradf[k] - it->zL[k] + it->zU[k];
    for (int i = 1; i <= N; i++) { /* x_i: +lam_i - A' lam_{i+1} ; lam_i stored at lam[6*(i-1)] */
        for (int a = 0; a < 6; a++) {
            double s = it->lam[6 * (i - 1) + a];
            if (i < N)
                for (int b = 0; b < 6; b++) s -= p->A[6 * b + a] * it->lam[6 * i + b];
            rw[IX(i) + a] += s;
        }
    }
    for (int i = 0; i < N; i++)
        for (int a = 0; a < 2; a++) {
            double s = 0.0;
            for (int b = 0; b < 6; b++) s += p->B[2 * b + a] * it->lam[6 * i + b];
            rw[IU(c, i) + a] -= s;
        }
    for (int r = 0; r < c->m; r++) {
        int col[6];
        row_cols(c, r, col);
        for (int e = 0; e < 6; e++)
            if (col[e] >= 0) rw[col[e]] -= jac[6 * r + e] * it->y[r];
    }
}

static err_t errors(const ctx_t *c, const iter_t *it, double mu, double rho, const double *rw, const double *cres,
                    const double *g) {
    const double s_max = 100.0;
    err_t e;
    double dual = 0.0, prim = 0.0, comp = 0.0, zsum = 0.0, ysum = 0.0;
    int nb = 0;
    for (int k = 0; k < c->n; k++) {
        dual = dmax(dual, fabs(rw[k]));
        if (has_l(c, k)) { comp = dmax(comp, fabs((it->w[k] - c->lb[k]) * it->zL[k] - mu)); zsum += it->zL[k]; nb++; }
        if (has_u(c, k)) { comp = dmax(comp, fabs((c->ub[k] - it->w[k]) * it->zU[k] - mu)); zsum += it->zU[k]; nb++; }
    }
    for (int r = 0; r < c->m; r++) {
        dual = dmax(dual, fabs(it->y[r] - it->z[r]));
        dual = dmax(dual, fabs(rho - it->y[r] - it->v[r]));
        prim = dmax(prim, fabs(g[r] + it->t[r] - it->s[r]));
        comp = dmax(comp, fabs(it->s[r] * it->z[r] - mu));
        comp = dmax(comp, fabs(it->t[r] * it->v[r] - mu));
        zsum += it->z[r] + it->v[r];
        ysum += fabs(it->y[r]);
        nb += 2;
    }
    int me = 6 * c->N;
    for (int k = 0; k < me; k++) { prim = dmax(prim, fabs(cres[k])); ysum += fabs(it->lam[k]); }
    int nmul = me + c->m + nb;
    double sd = dmax(s_max, (ysum + zsum) / (nmul > 0 ? nmul : 1)) / s_max;
    double sc = dmax(s_max, zsum / (nb > 0 ? nb : 1)) / s_max;
    e.dual = dual; e.prim = prim; e.comp = comp;
    e.E = dmax(dual / sd, dmax(prim, comp / sc));
    return e;
}

static double theta_of(const ctx_t *c, const double *w, const double *s, const double *t, double *cres, double *g) {
    eval_c(c, w, cres);
    eval_rows(c, w, g, NULL, NULL);
    double th = 0.0;
    for (int k = 0; k < 6 * c->N; k++) th += fabs(cres[k]);
    for (int r = 0; r < c->m; r++) th += fabs(g[r] + t[r] - s[r]);
    return th;
}

static double phi_of(const ctx_t *c, const double *w, const double *s, const double *t, double mu, double rho) {
    double b = 0.0, ts = 0.0;
    for (int k = 0; k < c->n; k++) {
        if (has_l(c, k)) b += log(w[k] - c->lb[k]);
        if (has_u(c, k)) b += log(c->ub[k] - w[k]);
    }
    for (int r = 0; r < c->m; r++) { b += log(s[r]) + log(t[r]); ts += t[r]; }
    return c->df * eval_f(c, w) + rho * ts - mu * b;
}

void orc_default_options(orc_options *o) {
    o->tol = 1e-8;
    o->max_iter = 200;
    o->mu_init = 0.1;
    o->rho = 1e3;
    o->bound_push = 1e-2;
    o->bound_frac = 1e-2;
    o->acceptable_tol = 1e-6;
    o->acceptable_iter = 15;
    o->max_grad = 100.0;
    o->start = 0;
    o->max_reset = 5;
}

int orc_solve(const orc_problem *p, const orc_options *o, orc_result *res) {
    const double kappa_eps = 10.0, kappa_mu = 0.2, theta_mu = 1.5, tau_min = 0.99;
    const double gamma_theta = 1e-5, gamma_phi = 1e-8, delta_sw = 1.0, s_theta = 1.1, s_phi = 2.3, eta_phi = 1e-8;
    const double gamma_alpha = 0.05, kappa_sigma = 1e10;
    ctx_t cx;
    ctx_t *c = &cx;
    memset(c, 0, sizeof(*c));
    c->p = p;
    int N = c->N = p->N, M = c->M = p->M;
    if (N < 1 || N > ORC_NMAX || M < 0 || M > ORC_MMAX) return -1;
    int nx = c->nx = 6 * N, nu = c->nu = 2 * N, ns = c->ns = M * (N + 1);
    int n = c->n = nx + nu + ns, m = c->m = M * N, nr = c->nr = nu + ns, me = 6 * N;
    /* ---- workspace */
    size_t nd = (size_t)16 * n + 32 * (m + 1) + 4 * me + (size_t)n * n + (size_t)n * nr + 2 * (size_t)nr * nr + 8 * nr + 64;
    double *buf = (double *)calloc(nd, sizeof(double));
    if (!buf) return -2;
    double *q = buf;
#define TAKE(k) (q += (k), q - (k))
    double *w = TAKE(n), *wt = TAKE(n), *dw = TAKE(n), *gradf = TAKE(n), *rw = TAKE(n), *rhs = TAKE(n);
    double *zL = TAKE(n), *zU = TAKE(n), *dzL = TAKE(n), *dzU = TAKE(n), *lb = TAKE(n), *ub = TAKE(n), *dp = TAKE(n);
    double *hd = TAKE(n), *sigw = TAKE(n), *tmpn = TAKE(n);
    double *s = TAKE(m + 1), *t = TAKE(m + 1), *y = TAKE(m + 1), *z = TAKE(m + 1), *v = TAKE(m + 1), *g = TAKE(m + 1);
    double *st = TAKE(m + 1), *tt = TAKE(m + 1), *ds = TAKE(m + 1), *dt = TAKE(m + 1), *dy = TAKE(m + 1), *dz = TAKE(m + 1);
    double *dv = TAKE(m + 1), *sige = TAKE(m + 1), *yhat = TAKE(m + 1), *dg = TAKE(m + 1), *gt = TAKE(m + 1), *Jd = TAKE(m + 1);
    double *jac = TAKE(6 * (m + 1)), *hes = TAKE(4 * (m + 1));
    double *lam = TAKE(me), *lamn = TAKE(me), *cres = TAKE(me), *crest = TAKE(me);
    double *K = TAKE((size_t)n * n), *T1 = TAKE((size_t)n * nr), *Rh = TAKE((size_t)nr * nr), *Rl = TAKE((size_t)nr * nr);
    double *rr = TAKE(nr), *dvv = TAKE(nr);
    c->lb = lb; c->ub = ub; c->dg = NULL;
    /* A^p B */
    memcpy(c->AB[0], p->B, sizeof(double) * 12);
    for (int k = 1; k < N; k++)
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 2; b++) {
                double sacc = 0.0;
                for (int e = 0; e < 6; e++) sacc += p->A[6 * a + e] * c->AB[k - 1][2 * e + b];
                c->AB[k][2 * a + b] = sacc;
            }
    /* ---- bounds */
    for (int k = 0; k < n; k++) { lb[k] = -HUGE_VAL; ub[k] = HUGE_VAL; }
    for (int i = 1; i <= N; i++) {
        if (p->per_stage_bounds) {
            lb[IX(i) + 0] = p->xlb[2 * i]; ub[IX(i) + 0] = p->xub[2 * i];
            lb[IX(i) + 5] = p->xlb[2 * i + 1]; ub[IX(i) + 5] = p->xub[2 * i + 1];
        } else {
            lb[IX(i) + 0] = p->vmin; ub[IX(i) + 0] = p->vmax;
            lb[IX(i) + 5] = -p->width; ub[IX(i) + 5] = p->width;
        }
    }
    for (int i = 0; i < N; i++)
        for (int a = 0; a < 2; a++) { lb[IU(c, i) + a] = -p->umax[a]; ub[IU(c, i) + a] = p->umax[a]; }
    for (int k = 0; k < ns; k++) lb[nx + nu + k] = 0.0;
    /* ---- start: u = 0 roll-out (start = 0) or w = 0 (start = 1: buf is calloc'ed), sigma = 0, pushed into the bounds */
    if (o->start == 0) {
        double x[6];
        memcpy(x, p->x0, sizeof(x));
        for (int i = 1; i <= N; i++) {
            double xn[6];
            for (int a = 0; a < 6; a++) {
                double sacc = 0.0;
                for (int b = 0; b < 6; b++) sacc += p->A[6 * a + b] * x[b];
                xn[a] = sacc;
            }
            memcpy(x, xn, sizeof(x));
            memcpy(w + IX(i), x, sizeof(x));
        }
    }
    for (int k = 0; k < n; k++) {
        int hl = has_l(c, k), hu = has_u(c, k);
        if (hl) {
            double pl = o->bound_push * dmax(1.0, fabs(lb[k]));
            if (hu) pl = dmin(pl, o->bound_frac * (ub[k] - lb[k]));
            if (w[k] < lb[k] + pl) w[k] = lb[k] + pl;
        }
        if (hu) {
            double pu = o->bound_push * dmax(1.0, fabs(ub[k]));
            if (hl) pu = dmin(pu, o->bound_frac * (ub[k] - lb[k]));
            if (w[k] > ub[k] - pu) w[k] = ub[k] - pu;
        }
    }
    /* ---- gradient-based scaling at the start */
    eval_grad(c, w, gradf);
    double gmax = 0.0;
    for (int k = 0; k < n; k++) gmax = dmax(gmax, fabs(gradf[k]));
    c->df = gmax > o->max_grad ? o->max_grad / gmax : 1.0;
    eval_rows(c, w, NULL, jac, NULL);
    for (int r = 0; r < m; r++) {
        int col[6];
        row_cols(c, r, col);
        double rm = 0.0;
        for (int e = 0; e < 6; e++)
            if (col[e] >= 0) rm = dmax(rm, fabs(jac[6 * r + e]));
        dg[r] = rm > o->max_grad ? o->max_grad / rm : 1.0;
    }
    c->dg = dg;
    const double rho = o->rho;
    eval_rows(c, w, g, NULL, NULL);
    for (int r = 0; r < m; r++) {
        t[r] = dmax(0.0, -g[r]) + o->bound_push;
        s[r] = g[r] + t[r];
        z[r] = 1.0; v[r] = 1.0; y[r] = 0.0;
    }
    for (int k = 0; k < n; k++) { zL[k] = has_l(c, k) ? 1.0 : 0.0; zU[k] = has_u(c, k) ? 1.0 : 0.0; }
    double mu = o->mu_init, dw_last = 0.0;
    iter_t it = {w, s, t, lam, y, z, v, zL, zU};
    double th0 = theta_of(c, w, s, t, cres, g);
    double theta_max = 1e4 * dmax(1.0, th0), theta_min = 1e-4 * dmax(1.0, th0);
    double filt_th[FILT_MAX], filt_ph[FILT_MAX];
    int last_needed = 0;
    int nfilt = 0, iter = 0, status = 1, n_acc = 0, n_refac = 0, n_back = 0, n_reset = 0;
    double E0 = 0.0;
    for (;;) {
        eval_grad(c, w, gradf);
        eval_c(c, w, cres);
        eval_rows(c, w, g, jac, hes);
        lag_grad(c, &it, gradf, jac, rw);
        err_t e0 = errors(c, &it, 0.0, rho, rw, cres, g);
        E0 = e0.E;
        if (E0 <= o->tol) { status = 0; break; }
        if (E0 <= o->acceptable_tol) {
            if (++n_acc >= o->acceptable_iter) { status = 0; break; }
        } else
            n_acc = 0;
        if (iter >= o->max_iter) { status = 1; break; }
        /* barrier parameter update (monotone, Fiacco-McCormick) */
        for (;;) {
            err_t em = errors(c, &it, mu, rho, rw, cres, g);
            if (em.E <= kappa_eps * mu && mu > o->tol / 11.0) {
                mu = dmax(o->tol / 11.0, dmin(kappa_mu * mu, pow(mu, theta_mu)));
                nfilt = 0;
            } else
                break;
        }
        double tau = dmax(tau_min, 1.0 - mu);
        /* ---- condensed Newton system */
        for (int r = 0; r < m; r++) {
            double sig_s = z[r] / s[r], sig_t = v[r] / t[r];
            double beta = sig_t / (sig_s + sig_t);
            double rg = g[r] + t[r] - s[r];
            sige[r] = beta * sig_s;
            yhat[r] = (1.0 - beta) * (rho - mu / t[r]) + beta * (mu / s[r] - sig_s * rg);
        }
        for (int k = 0; k < n; k++) {
            double sw = 0.0, b = -c->df * gradf[k];
            if (has_l(c, k)) { sw += zL[k] / (w[k] - lb[k]); b += mu / (w[k] - lb[k]); }
            if (has_u(c, k)) { sw += zU[k] / (ub[k] - w[k]); b -= mu / (ub[k] - w[k]); }
            sigw[k] = sw; rhs[k] = b; hd[k] = 0.0;
        }
        for (int r = 0; r < m; r++) {
            int col[6];
            row_cols(c, r, col);
            for (int e = 0; e < 6; e++)
                if (col[e] >= 0) rhs[col[e]] += jac[6 * r + e] * yhat[r];
            /* Hessian of -y_r g_r : diagonal on (s,ey) of x_i and x_{i+1} */
            if (col[0] >= 0) { hd[col[0]] -= y[r] * hes[4 * r + 0]; hd[col[1]] -= y[r] * hes[4 * r + 1]; }
            hd[col[3]] -= y[r] * hes[4 * r + 2];
            hd[col[4]] -= y[r] * hes[4 * r + 3];
        }
        /* particular solution of the linearised dynamics (du = 0) */
        memset(dp, 0, sizeof(double) * n);
        for (int i = 0; i < N; i++)
            for (int a = 0; a < 6; a++) {
                double sacc = -cres[6 * i + a];
                if (i > 0)
                    for (int b = 0; b < 6; b++) sacc += p->A[6 * a + b] * dp[IX(i) + b];
                dp[IX(i + 1) + a] = sacc;
            }
        /* IPOPT always retries dw = 0 first; when the previous iteration needed a correction we start
         * from dw_last/3 instead (deviation 3) */
        double dw_try = last_needed ? dmax(1e-20, dw_last / 3.0) : 0.0;
        for (;;) {
            memset(K, 0, sizeof(double) * (size_t)n * n);
            for (int i = 1; i <= N; i++)
                for (int a = 0; a < 6; a++)
                    for (int b = 0; b < 6; b++) K[(size_t)(IX(i) + a) * n + IX(i) + b] = c->df * (p->Q[6 * a + b] + p->Q[6 * b + a]);
            for (int i = 0; i < N; i++) {
                int k0 = IU(c, i);
                K[(size_t)k0 * n + k0] = c->df * 2 * p->R[0];
                K[(size_t)k0 * n + k0 + 1] = K[(size_t)(k0 + 1) * n + k0] = c->df * (p->R[1] + p->R[2]);
                K[(size_t)(k0 + 1) * n + k0 + 1] = c->df * 2 * p->R[3];
            }
            for (int i = 0; i < N; i++) {
                double h2 = 2.0 * c->df * p->wd[i];
                int kn = IX(i + 1) + 5;
                K[(size_t)kn * n + kn] += h2;
                if (i > 0) {
                    int kc = IX(i) + 5;
                    K[(size_t)kc * n + kc] += h2;
                    K[(size_t)kc * n + kn] -= h2;
                    K[(size_t)kn * n + kc] -= h2;
                }
            }
            for (int k = 0; k < n; k++) K[(size_t)k * n + k] += hd[k] + sigw[k] + dw_try;
            for (int r = 0; r < m; r++) {
                int col[6];
                row_cols(c, r, col);
                for (int a = 0; a < 6; a++)
                    for (int b = 0; b < 6; b++)
                        if (col[a] >= 0 && col[b] >= 0) K[(size_t)col[a] * n + col[b]] += sige[r] * jac[6 * r + a] * jac[6 * r + b];
            }
            /* T1 = K Z,  Z = [Gamma 0; I 0; 0 I]  (columns: u then sigma) */
            memset(T1, 0, sizeof(double) * (size_t)n * nr);
            for (int a = 0; a < n; a++) {
                double *row = T1 + (size_t)a * nr;
                for (int i = 1; i <= N; i++)
                    for (int e = 0; e < 6; e++) {
                        double kv = K[(size_t)a * n + IX(i) + e];
                        if (kv == 0.0) continue;
                        for (int l = 0; l < i; l++) { /* x_i depends on u_l, l<i, through A^{i-1-l} B */
                            row[2 * l] += kv * c->AB[i - 1 - l][2 * e];
                            row[2 * l + 1] += kv * c->AB[i - 1 - l][2 * e + 1];
                        }
                    }
                for (int b = 0; b < nr; b++) row[b] += K[(size_t)a * n + nx + b];
            }
            /* Rh = Z' T1 */
            for (int l = 0; l < N; l++)
                for (int cc = 0; cc < 2; cc++) {
                    double *row = Rh + (size_t)(2 * l + cc) * nr;
                    memcpy(row, T1 + (size_t)(IU(c, l) + cc) * nr, sizeof(double) * nr);
                    for (int i = l + 1; i <= N; i++)
                        for (int e = 0; e < 6; e++) {
                            double gv = c->AB[i - 1 - l][2 * e + cc];
                            const double *src = T1 + (size_t)(IX(i) + e) * nr;
                            for (int b = 0; b < nr; b++) row[b] += gv * src[b];
                        }
                }
            for (int k = 0; k < ns; k++) memcpy(Rh + (size_t)(nu + k) * nr, T1 + (size_t)(nx + nu + k) * nr, sizeof(double) * nr);
            memcpy(Rl, Rh, sizeof(double) * (size_t)nr * nr);
            if (cholesky(Rl, nr) == 0) break;
            n_refac++;
            if (dw_try == 0.0)
                dw_try = dw_last == 0.0 ? 1e-4 : dmax(1e-20, dw_last / 3.0);
            else
                dw_try *= dw_last == 0.0 ? 100.0 : 8.0;
            if (dw_try > 1e40) { status = 3; goto done; }
        }
        if (dw_try > 0.0) dw_last = dw_try;
        last_needed = dw_try > 0.0;
        /* reduced rhs = Z'(rhs - K dp) */
        for (int a = 0; a < n; a++) {
            double sacc = rhs[a];
            for (int b = 0; b < nx; b++) sacc -= K[(size_t)a * n + b] * dp[b];
            tmpn[a] = sacc;
        }
        for (int l = 0; l < N; l++)
            for (int cc = 0; cc < 2; cc++) {
                double sacc = tmpn[IU(c, l) + cc];
                for (int i = l + 1; i <= N; i++)
                    for (int e = 0; e < 6; e++) sacc += c->AB[i - 1 - l][2 * e + cc] * tmpn[IX(i) + e];
                rr[2 * l + cc] = sacc;
            }
        for (int k = 0; k < ns; k++) rr[nu + k] = tmpn[nx + nu + k];
        memcpy(dvv, rr, sizeof(double) * nr);
        chol_solve(Rl, nr, dvv);
        /* dw = Z dv + dp */
        memcpy(dw, dp, sizeof(double) * n);
        for (int k = 0; k < nr; k++) dw[nx + k] = dvv[k];
        for (int i = 1; i <= N; i++)
            for (int e = 0; e < 6; e++) {
                double sacc = 0.0;
                for (int l = 0; l < i; l++) sacc += c->AB[i - 1 - l][2 * e] * dvv[2 * l] + c->AB[i - 1 - l][2 * e + 1] * dvv[2 * l + 1];
                dw[IX(i) + e] += sacc;
            }
        /* lam+ from  K dw + Jc' lam+ = rhs  (backward costate recursion) */
        for (int a = 0; a < nx; a++) {
            double sacc = rhs[a];
            for (int b = 0; b < n; b++) sacc -= K[(size_t)a * n + b] * dw[b];
            tmpn[a] = sacc;
        }
        for (int i = N; i >= 1; i--)
            for (int a = 0; a < 6; a++) {
                double sacc = tmpn[IX(i) + a];
                if (i < N)
                    for (int b = 0; b < 6; b++) sacc += p->A[6 * b + a] * lamn[6 * i + b];
                lamn[6 * (i - 1) + a] = sacc;
            }
        /* slack / multiplier steps */
        for (int r = 0; r < m; r++) {
            int col[6];
            row_cols(c, r, col);
            double jd = 0.0;
            for (int e = 0; e < 6; e++)
                if (col[e] >= 0) jd += jac[6 * r + e] * dw[col[e]];
            Jd[r] = jd;
            double sig_s = z[r] / s[r], sig_t = v[r] / t[r];
            double rg = g[r] + t[r] - s[r];
            dy[r] = yhat[r] - sige[r] * jd - y[r];
            dt[r] = (mu / s[r] + mu / t[r] - rho - sig_s * rg - sig_s * jd) / (sig_s + sig_t);
            dv[r] = mu / t[r] - v[r] - sig_t * dt[r];
            ds[r] = jd + dt[r] + rg;
            dz[r] = mu / s[r] - z[r] - sig_s * ds[r];
        }
        double a_max = 1.0, a_z = 1.0;
        for (int k = 0; k < n; k++) {
            dzL[k] = dzU[k] = 0.0;
            if (has_l(c, k)) {
                double d = w[k] - lb[k];
                dzL[k] = mu / d - zL[k] - zL[k] / d * dw[k];
                if (dw[k] < 0.0) a_max = dmin(a_max, -tau * d / dw[k]);
                if (dzL[k] < 0.0) a_z = dmin(a_z, -tau * zL[k] / dzL[k]);
            }
            if (has_u(c, k)) {
                double d = ub[k] - w[k];
                dzU[k] = mu / d - zU[k] + zU[k] / d * dw[k];
                if (dw[k] > 0.0) a_max = dmin(a_max, tau * d / dw[k]);
                if (dzU[k] < 0.0) a_z = dmin(a_z, -tau * zU[k] / dzU[k]);
            }
        }
        for (int r = 0; r < m; r++) {
            if (ds[r] < 0.0) a_max = dmin(a_max, -tau * s[r] / ds[r]);
            if (dt[r] < 0.0) a_max = dmin(a_max, -tau * t[r] / dt[r]);
            if (dz[r] < 0.0) a_z = dmin(a_z, -tau * z[r] / dz[r]);
            if (dv[r] < 0.0) a_z = dmin(a_z, -tau * v[r] / dv[r]);
        }
        /* ---- filter line search */
        double th = 0.0;
        for (int k = 0; k < me; k++) th += fabs(cres[k]);
        for (int r = 0; r < m; r++) th += fabs(g[r] + t[r] - s[r]);
        double ph = phi_of(c, w, s, t, mu, rho);
        double gphi = 0.0;
        for (int k = 0; k < n; k++) {
            gphi += c->df * gradf[k] * dw[k];
            if (has_l(c, k)) gphi -= mu * dw[k] / (w[k] - lb[k]);
            if (has_u(c, k)) gphi += mu * dw[k] / (ub[k] - w[k]);
        }
        for (int r = 0; r < m; r++) gphi += rho * dt[r] - mu * (ds[r] / s[r] + dt[r] / t[r]);
        double amin;
        if (gphi < 0.0 && th <= theta_min)
            amin = gamma_alpha * dmin(gamma_theta, dmin(gamma_phi * th / (-gphi), delta_sw * pow(th, s_theta) / pow(-gphi, s_phi)));
        else if (gphi < 0.0)
            amin = gamma_alpha * dmin(gamma_theta, gamma_phi * th / (-gphi));
        else
            amin = gamma_alpha * gamma_theta;
        double a = a_max;
        int accepted = 0, ftype = 0, nls = 0;
        while (a >= amin || nls == 0) {
            for (int k = 0; k < n; k++) wt[k] = w[k] + a * dw[k];
            for (int r = 0; r < m; r++) { st[r] = s[r] + a * ds[r]; tt[r] = t[r] + a * dt[r]; }
            double tht = theta_of(c, wt, st, tt, crest, gt);
            double pht = phi_of(c, wt, st, tt, mu, rho);
            int okf = tht < theta_max;
            for (int f = 0; f < nfilt && okf; f++)
                if (tht >= filt_th[f] && pht >= filt_ph[f]) okf = 0;
            if (okf) {
                int sw = gphi < 0.0 && a * pow(-gphi, s_phi) > delta_sw * pow(th, s_theta);
                if (th <= theta_min && sw) {
                    if (pht <= ph + eta_phi * a * gphi) { accepted = 1; ftype = 1; }
                } else if (tht <= (1.0 - gamma_theta) * th || pht <= ph - gamma_phi * th)
                    accepted = 1;
            }
            if (accepted) break;
            a *= 0.5;
            nls++;
            n_back++;
        }
        if (!accepted) {
            /* IPOPT would enter its restoration phase here.  The rows are elastic, so their
             * residual can be removed exactly by enlarging the slacks: t' = max(t, s-g), s' = g+t'
             * (both only grow); then restart the filter.  At most max_reset times per solve. */
            if (n_reset >= o->max_reset) { status = 2; break; }
            n_reset++;
            for (int r = 0; r < m; r++) {
                double tn = dmax(t[r], s[r] - g[r]);
                t[r] = tn;
                s[r] = g[r] + tn;
            }
            nfilt = 0;
            iter++;
            continue;
        }
        if (!ftype) {
            if (nfilt == FILT_MAX) { /* drop the oldest entry */
                memmove(filt_th, filt_th + 1, sizeof(double) * (FILT_MAX - 1));
                memmove(filt_ph, filt_ph + 1, sizeof(double) * (FILT_MAX - 1));
                nfilt--;
            }
            filt_th[nfilt] = (1.0 - gamma_theta) * th;
            filt_ph[nfilt] = ph - gamma_phi * th;
            nfilt++;
        }
        memcpy(w, wt, sizeof(double) * n);
        for (int r = 0; r < m; r++) {
            s[r] = st[r]; t[r] = tt[r];
            y[r] += a * dy[r];
            z[r] += a_z * dz[r];
            v[r] += a_z * dv[r];
            z[r] = dmax(dmin(z[r], kappa_sigma * mu / s[r]), mu / (kappa_sigma * s[r]));
            v[r] = dmax(dmin(v[r], kappa_sigma * mu / t[r]), mu / (kappa_sigma * t[r]));
        }
        for (int k = 0; k < me; k++) lam[k] += a * (lamn[k] - lam[k]);
        for (int k = 0; k < n; k++) {
            if (has_l(c, k)) {
                double d = w[k] - lb[k];
                zL[k] += a_z * dzL[k];
                zL[k] = dmax(dmin(zL[k], kappa_sigma * mu / d), mu / (kappa_sigma * d));
            }
            if (has_u(c, k)) {
                double d = ub[k] - w[k];
                zU[k] += a_z * dzU[k];
                zU[k] = dmax(dmin(zU[k], kappa_sigma * mu / d), mu / (kappa_sigma * d));
            }
        }
        iter++;
    }
done:
    memset(res, 0, sizeof(*res));
    memcpy(res->x, p->x0, sizeof(double) * 6);
    memcpy(res->x + 6, w, sizeof(double) * nx);
    memcpy(res->u, w + nx, sizeof(double) * nu);
    memcpy(res->sigma, w + nx + nu, sizeof(double) * ns);
    res->cost = eval_f(c, w);
    res->kkt_err = E0;
    /* control.py:582-586 (and :228-232) impose the bound rows on stage 0 as well, where x_0 is fixed by :497: an x_0
     * outside them (beyond IPOPT's constr_viol_tol 1e-4) makes the reference's NLP infeasible.  The solve above is that
     * of the problem without the stage-0 rows; it is returned with status 4. */
    {
        double l0 = p->per_stage_bounds ? p->xlb[0] : p->vmin, u0 = p->per_stage_bounds ? p->xub[0] : p->vmax;
        double l1 = p->per_stage_bounds ? p->xlb[1] : -p->width, u1 = p->per_stage_bounds ? p->xub[1] : p->width;
        const double ctol = 1e-4;
        if (p->x0[0] < l0 - ctol || p->x0[0] > u0 + ctol || p->x0[5] < l1 - ctol || p->x0[5] > u1 + ctol) status = 4;
    }
    res->status = status;
    res->iters = iter;
    res->n_refactor = n_refac;
    res->n_backtrack = n_back;
    res->df = c->df;
    double tm = 0.0;
    for (int r = 0; r < m; r++) { tm = dmax(tm, t[r]); res->y[r] = y[r] * dg[r] / c->df; }
    res->elastic_max = tm;
    for (int k = 0; k < me; k++) res->lam[k] = lam[k] / c->df;
    free(buf);
    return 0;
}

typedef struct {
    const orc_problem *p;
    const orc_options *o;
    orc_result *r;
    int B, tid, nt;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    for (int b = j->tid; b < j->B; b += j->nt) orc_solve(j->p + b, j->o, j->r + b);
    return NULL;
}

int orc_solve_batch(const orc_problem *p, int B, const orc_options *o, orc_result *r, int nthreads) {
    if (nthreads <= 1) {
        for (int b = 0; b < B; b++) orc_solve(p + b, o, r + b);
        return 0;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * nthreads);
    for (int k = 0; k < nthreads; k++) {
        jobs[k] = (job_t){p, o, r, B, k, nthreads};
        pthread_create(&th[k], NULL, worker, &jobs[k]);
    }
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    free(th);
    free(jobs);
    return 0;
}

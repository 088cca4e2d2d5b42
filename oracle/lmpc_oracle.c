/* TEST INFRASTRUCTURE ONLY -- CPU oracle for control.lmpc (car_racing/control/control.py:610-730).
 *
 * Restates the LMPC QP exactly as the reference builds it with CasADi:
 *   vars  x_0..x_N (6), u_0..u_{N-1} (2), lambda (K = num_ss_points), slack (6, forced to 0 by :693-694 -> dropped)
 *   min   sum_{i<N} [(x_i-x_trk)'Q(x_i-x_trk) + u_i'R u_i + (u_i-u_{i-1})'dR(u_i-u_{i-1})]      (:667-682, u_{-1}=u_old)
 *         + (x_N-x_trk)'Q(x_N-x_trk) + Qfun' lambda                                            (:684-687,:695)
 *   s.t.  x_0 = xcurv (:650);  x_{i+1} = A_i x_i + B_i u_i + C_i (:653-656, LTV)
 *         vx_i <= v_max, |ey_i| <= lap_width for i < N (:658-660);  |delta|<=delta_max, |a|<=a_max (:662-666)
 *         lambda >= 0 (:689);  x_N = SS lambda (:690-691);  1'lambda = 1 (:692)
 * and solves it with the interior-point definition of DESIGN.md section 2 (IPOPT conventions; the QP is convex, so no
 * inertia correction and no elastic rows are involved).  Linear algebra is deliberately not the product's: the full
 * KKT matrix is assembled densely and solved by LU with partial pivoting.
 * PARITY: solver algorithm UNPINNED for the same reason as ocp_oracle.c (no CasADi/IPOPT here, no golden vectors in the
 * reference); the problem statement is pinned to the reference's own code (tests/test_reference_statement.py); the QP is
 * convex, so any correct solver returns the reference's optimum.
 */
#include "ocp_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define FILT_MAX 64
static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline double dmin(double a, double b) { return a < b ? a : b; }

typedef struct {
    const orc_lmpc_problem *p;
    int N, K, nx, nu, n, me;
    double df;
    double *lb, *ub;
} lctx;

static inline int LX(int i) { return 6 * (i - 1); }                     /* x_i, i>=1 */
static inline int LU(const lctx *c, int i) { return c->nx + 2 * i; }
static inline int LL(const lctx *c, int k) { return c->nx + c->nu + k; }
static inline double XV(const lctx *c, const double *w, int i, int k) { return i == 0 ? c->p->x0[k] : w[LX(i) + k]; }
static inline double UPREV(const lctx *c, const double *w, int i, int a) { return i == 0 ? c->p->u_old[a] : w[LU(c, i - 1) + a]; }
static int hasl(const lctx *c, int k) { return c->lb[k] > -HUGE_VAL; }
static int hasu(const lctx *c, int k) { return c->ub[k] < HUGE_VAL; }

static double l_f(const lctx *c, const double *w) {
    const orc_lmpc_problem *p = c->p;
    double f = 0.0;
    for (int i = 0; i <= c->N; i++) {
        double d[6];
        for (int k = 0; k < 6; k++) d[k] = XV(c, w, i, k) - p->xtrk[k];
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) f += d[a] * p->Q[6 * a + b] * d[b];
    }
    for (int i = 0; i < c->N; i++) {
        const double *u = w + LU(c, i);
        double du[2] = {u[0] - UPREV(c, w, i, 0), u[1] - UPREV(c, w, i, 1)};
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) f += u[a] * p->R[2 * a + b] * u[b] + du[a] * p->dR[2 * a + b] * du[b];
    }
    for (int k = 0; k < c->K; k++) f += p->Qfun[k] * w[LL(c, k)];
    return f;
}

static void l_grad(const lctx *c, const double *w, double *g) {
    const orc_lmpc_problem *p = c->p;
    memset(g, 0, sizeof(double) * c->n);
    for (int i = 1; i <= c->N; i++) {
        double d[6];
        for (int k = 0; k < 6; k++) d[k] = w[LX(i) + k] - p->xtrk[k];
        for (int a = 0; a < 6; a++) {
            double s = 0.0;
            for (int b = 0; b < 6; b++) s += (p->Q[6 * a + b] + p->Q[6 * b + a]) * d[b];
            g[LX(i) + a] = s;
        }
    }
    for (int i = 0; i < c->N; i++) {
        const double *u = w + LU(c, i);
        double du[2] = {u[0] - UPREV(c, w, i, 0), u[1] - UPREV(c, w, i, 1)};
        for (int a = 0; a < 2; a++) {
            double s = 0.0;
            for (int b = 0; b < 2; b++) s += (p->R[2 * a + b] + p->R[2 * b + a]) * u[b] + (p->dR[2 * a + b] + p->dR[2 * b + a]) * du[b];
            g[LU(c, i) + a] += s;
            if (i > 0) {
                double t = 0.0;
                for (int b = 0; b < 2; b++) t += (p->dR[2 * a + b] + p->dR[2 * b + a]) * du[b];
                g[LU(c, i - 1) + a] -= t;
            }
        }
    }
    for (int k = 0; k < c->K; k++) g[LL(c, k)] = p->Qfun[k];
}

/* equality residuals: dynamics c_i (6N), terminal e = x_N - SS lambda (6), s = 1'lambda - 1 */
static void l_ceq(const lctx *c, const double *w, double *r) {
    const orc_lmpc_problem *p = c->p;
    int N = c->N;
    for (int i = 0; i < N; i++) {
        const double *u = w + LU(c, i);
        for (int a = 0; a < 6; a++) {
            double s = w[LX(i + 1) + a] - p->C[6 * i + a];
            for (int b = 0; b < 6; b++) s -= p->A[36 * i + 6 * a + b] * XV(c, w, i, b);
            s -= p->B[12 * i + 2 * a] * u[0] + p->B[12 * i + 2 * a + 1] * u[1];
            r[6 * i + a] = s;
        }
    }
    double sum = -1.0;
    for (int a = 0; a < 6; a++) {
        double s = w[LX(N) + a];
        for (int k = 0; k < c->K; k++) s -= p->SS[a * c->K + k] * w[LL(c, k)];
        r[6 * N + a] = s;
    }
    for (int k = 0; k < c->K; k++) sum += w[LL(c, k)];
    r[6 * N + 6] = sum;
}

typedef struct { double *w, *lam, *zL, *zU; } lit;   /* lam: 6N dynamics + 7 terminal multipliers */

static void l_laggrad(const lctx *c, const lit *it, const double *gradf, double *rw) {
    const orc_lmpc_problem *p = c->p;
    int N = c->N, K = c->K;
    for (int k = 0; k < c->n; k++) rw[k] = c->df * gradf[k] - it->zL[k] + it->zU[k];
    for (int i = 1; i <= N; i++)
        for (int a = 0; a < 6; a++) {
            double s = it->lam[6 * (i - 1) + a];
            if (i < N)
                for (int b = 0; b < 6; b++) s -= p->A[36 * i + 6 * b + a] * it->lam[6 * i + b];
            else
                s += it->lam[6 * N + a];                       /* + nu_x */
            rw[LX(i) + a] += s;
        }
    for (int i = 0; i < N; i++)
        for (int a = 0; a < 2; a++) {
            double s = 0.0;
            for (int b = 0; b < 6; b++) s += p->B[12 * i + 2 * b + a] * it->lam[6 * i + b];
            rw[LU(c, i) + a] -= s;
        }
    for (int k = 0; k < K; k++) {
        double s = it->lam[6 * N + 6];                           /* nu_1 */
        for (int a = 0; a < 6; a++) s -= p->SS[a * K + k] * it->lam[6 * N + a];
        rw[LL(c, k)] += s;
    }
}

static double l_err(const lctx *c, const lit *it, double mu, const double *rw, const double *ceq) {
    const double s_max = 100.0;
    double dual = 0.0, prim = 0.0, comp = 0.0, zsum = 0.0, ysum = 0.0;
    int nb = 0;
    for (int k = 0; k < c->n; k++) {
        dual = dmax(dual, fabs(rw[k]));
        if (hasl(c, k)) { comp = dmax(comp, fabs((it->w[k] - c->lb[k]) * it->zL[k] - mu)); zsum += it->zL[k]; nb++; }
        if (hasu(c, k)) { comp = dmax(comp, fabs((c->ub[k] - it->w[k]) * it->zU[k] - mu)); zsum += it->zU[k]; nb++; }
    }
    for (int k = 0; k < c->me; k++) { prim = dmax(prim, fabs(ceq[k])); ysum += fabs(it->lam[k]); }
    int nmul = c->me + nb;
    double sd = dmax(s_max, (ysum + zsum) / (nmul > 0 ? nmul : 1)) / s_max;
    double sc = dmax(s_max, zsum / (nb > 0 ? nb : 1)) / s_max;
    return dmax(dual / sd, dmax(prim, comp / sc));
}

static double l_phi(const lctx *c, const double *w, double mu) {
    double b = 0.0;
    for (int k = 0; k < c->n; k++) {
        if (hasl(c, k)) b += log(w[k] - c->lb[k]);
        if (hasu(c, k)) b += log(c->ub[k] - w[k]);
    }
    return c->df * l_f(c, w) - mu * b;
}

int orc_lmpc_solve(const orc_lmpc_problem *p, const orc_options *o, orc_lmpc_result *res) {
    const double kappa_eps = 10.0, kappa_mu = 0.2, theta_mu = 1.5, tau_min = 0.99;
    const double gamma_theta = 1e-5, gamma_phi = 1e-8, delta_sw = 1.0, s_theta = 1.1, s_phi = 2.3, eta_phi = 1e-8;
    const double gamma_alpha = 0.05, kappa_sigma = 1e10;
    lctx cx;
    lctx *c = &cx;
    memset(c, 0, sizeof(*c));
    c->p = p;
    int N = c->N = p->N, K = c->K = p->K;
    if (N < 2 || N > ORC_NMAX || K < 1 || K > ORC_KMAX) return -1;
    int nx = c->nx = 6 * N, nu = c->nu = 2 * N, n = c->n = nx + nu + K, me = c->me = 6 * N + 7, nr = nu + K;
    size_t nd = (size_t)16 * n + 6 * me + (size_t)nx * nu + (size_t)(n + me) * (n + me) + (size_t)(n + me) + (size_t)nr * nr + (size_t)7 * nr * 2 + 8 * nr + 256;
    double *buf = (double *)calloc(nd, sizeof(double));
    if (!buf) return -2;
    double *q = buf;
#define TAKE(k) (q += (k), q - (k))
    double *w = TAKE(n), *wt = TAKE(n), *dw = TAKE(n), *gradf = TAKE(n), *rw = TAKE(n), *rhs = TAKE(n), *zL = TAKE(n), *zU = TAKE(n);
    double *dzL = TAKE(n), *dzU = TAKE(n), *lb = TAKE(n), *ub = TAKE(n), *sigw = TAKE(n), *tmpn = TAKE(n), *kd = TAKE(n);
    double *lam = TAKE(me), *lamn = TAKE(me), *ceq = TAKE(me), *ceqt = TAKE(me);
    double *G = TAKE((size_t)nx * nu);                     /* dx = G du + dp */
    double *Rh = TAKE((size_t)(n + me) * (n + me)), *Rl = TAKE((size_t)(n + me) + (size_t)nr * nr);
    c->lb = lb; c->ub = ub;
    /* bounds */
    for (int k = 0; k < n; k++) { lb[k] = -HUGE_VAL; ub[k] = HUGE_VAL; }
    for (int i = 1; i < N; i++) {                           /* i < N only (:658-660); i = 0 is the fixed x_0 */
        ub[LX(i) + 0] = p->vmax;
        lb[LX(i) + 5] = -p->width; ub[LX(i) + 5] = p->width;
    }
    for (int i = 0; i < N; i++)
        for (int a = 0; a < 2; a++) { lb[LU(c, i) + a] = -p->umax[a]; ub[LU(c, i) + a] = p->umax[a]; }
    for (int k = 0; k < K; k++) lb[LL(c, k)] = 0.0;
    /* start: u = 0 roll-out through the LTV model, lambda = 1/K, pushed into the bounds */
    {
        double x[6];
        memcpy(x, p->x0, sizeof(x));
        for (int i = 0; i < N; i++) {
            double xn[6];
            for (int a = 0; a < 6; a++) {
                double s = p->C[6 * i + a];
                for (int b = 0; b < 6; b++) s += p->A[36 * i + 6 * a + b] * x[b];
                xn[a] = s;
            }
            memcpy(x, xn, sizeof(x));
            memcpy(w + LX(i + 1), x, sizeof(x));
        }
        for (int k = 0; k < K; k++) w[LL(c, k)] = 1.0 / K;
    }
    for (int k = 0; k < n; k++) {
        int hl = hasl(c, k), hu = hasu(c, k);
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
        zL[k] = hl ? 1.0 : 0.0;
        zU[k] = hu ? 1.0 : 0.0;
    }
    l_grad(c, w, gradf);
    double gmax = 0.0;
    for (int k = 0; k < n; k++) gmax = dmax(gmax, fabs(gradf[k]));
    c->df = gmax > o->max_grad ? o->max_grad / gmax : 1.0;
    /* G: x_{i} = sum_l G[i][l] u_l + ... ; G[(i-1)*6+a][2l+b] */
    for (int l = 0; l < N; l++) {
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 2; b++) G[(size_t)(LX(l + 1) + a) * nu + 2 * l + b] = p->B[12 * l + 2 * a + b];
        for (int i = l + 1; i < N; i++)
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 2; b++) {
                    double s = 0.0;
                    for (int e = 0; e < 6; e++) s += p->A[36 * i + 6 * a + e] * G[(size_t)(LX(i) + e) * nu + 2 * l + b];
                    G[(size_t)(LX(i + 1) + a) * nu + 2 * l + b] = s;
                }
    }
    double mu = o->mu_init;
    lit it = {w, lam, zL, zU};
    l_ceq(c, w, ceq);
    double th0 = 0.0;
    for (int k = 0; k < me; k++) th0 += fabs(ceq[k]);
    double theta_max = 1e4 * dmax(1.0, th0), theta_min = 1e-4 * dmax(1.0, th0);
    double filt_th[FILT_MAX], filt_ph[FILT_MAX];
    int nfilt = 0, iter = 0, status = 1, n_acc = 0;
    double E0 = 0.0;
    for (;;) {
        l_grad(c, w, gradf);
        l_ceq(c, w, ceq);
        l_laggrad(c, &it, gradf, rw);
        E0 = l_err(c, &it, 0.0, rw, ceq);
        if (E0 <= o->tol) { status = 0; break; }
        if (E0 <= o->acceptable_tol) {
            if (++n_acc >= o->acceptable_iter) { status = 0; break; }
        } else
            n_acc = 0;
        if (iter >= o->max_iter) { status = 1; break; }
        for (;;) {
            double em = l_err(c, &it, mu, rw, ceq);
            if (em <= kappa_eps * mu && mu > o->tol / 11.0) {
                mu = dmax(o->tol / 11.0, dmin(kappa_mu * mu, pow(mu, theta_mu)));
                nfilt = 0;
            } else
                break;
        }
        double tau = dmax(tau_min, 1.0 - mu);
        for (int k = 0; k < n; k++) {
            double sw = 0.0, b = -c->df * gradf[k];
            if (hasl(c, k)) { sw += zL[k] / (w[k] - lb[k]); b += mu / (w[k] - lb[k]); }
            if (hasu(c, k)) { sw += zU[k] / (ub[k] - w[k]); b -= mu / (ub[k] - w[k]); }
            sigw[k] = sw; rhs[k] = b;
        }
        /* K d for a given d (K = df*H + diag(sigw)); H: 2Q on x_i, 2R + dR coupling on u */
#define APPLY_K(d, out)                                                                                         \
    do {                                                                                                        \
        for (int k_ = 0; k_ < n; k_++) (out)[k_] = sigw[k_] * (d)[k_];                                          \
        for (int i_ = 1; i_ <= N; i_++)                                                                         \
            for (int a_ = 0; a_ < 6; a_++) {                                                                    \
                double s_ = 0.0;                                                                                \
                for (int b_ = 0; b_ < 6; b_++) s_ += (p->Q[6 * a_ + b_] + p->Q[6 * b_ + a_]) * (d)[LX(i_) + b_]; \
                (out)[LX(i_) + a_] += c->df * s_;                                                               \
            }                                                                                                   \
        for (int i_ = 0; i_ < N; i_++)                                                                          \
            for (int a_ = 0; a_ < 2; a_++) {                                                                    \
                double s_ = 0.0;                                                                                \
                for (int b_ = 0; b_ < 2; b_++) {                                                                \
                    double r2_ = p->R[2 * a_ + b_] + p->R[2 * b_ + a_], d2_ = p->dR[2 * a_ + b_] + p->dR[2 * b_ + a_]; \
                    double ub_ = (d)[LU(c, i_) + b_];                                                           \
                    double up_ = i_ > 0 ? (d)[LU(c, i_ - 1) + b_] : 0.0, un_ = i_ < N - 1 ? (d)[LU(c, i_ + 1) + b_] : 0.0; \
                    s_ += r2_ * ub_ + d2_ * (ub_ - up_) - (i_ < N - 1 ? d2_ * (un_ - ub_) : 0.0);               \
                }                                                                                               \
                (out)[LU(c, i_) + a_] += c->df * s_;                                                            \
            }                                                                                                   \
    } while (0)
        /* full KKT system [K J'; J 0][dw; lam+] = [rhs; -ceq], dense LU with partial pivoting.  (Condensing the states
         * and taking a Schur complement for the 7 terminal equalities was tried first: near the solution only ~3 of the
         * 44 lambdas are off their bound, the Schur complement's condition number exceeds 1e15 and its Cholesky breaks
         * down.  The pivoted solve of the full system is the textbook-stable route and still independent of the
         * product's Riccati recursion.) */
        {
            int nk = n + me;
            double *KK = Rh;                               /* reuse: nk*nk <= allocated? see allocation below */
            memset(KK, 0, sizeof(double) * (size_t)nk * nk);
            for (int col = 0; col < n; col++) {
                memset(tmpn, 0, sizeof(double) * n);
                tmpn[col] = 1.0;
                APPLY_K(tmpn, kd);
                for (int r = 0; r < n; r++) KK[(size_t)r * nk + col] = kd[r];
            }
            /* J rows: dynamics */
            for (int i = 0; i < N; i++)
                for (int a = 0; a < 6; a++) {
                    int r = n + 6 * i + a;
                    KK[(size_t)r * nk + LX(i + 1) + a] = 1.0;
                    if (i > 0)
                        for (int bb = 0; bb < 6; bb++) KK[(size_t)r * nk + LX(i) + bb] = -p->A[36 * i + 6 * a + bb];
                    for (int bb = 0; bb < 2; bb++) KK[(size_t)r * nk + LU(c, i) + bb] = -p->B[12 * i + 2 * a + bb];
                }
            for (int a = 0; a < 6; a++) {
                int r = n + 6 * N + a;
                KK[(size_t)r * nk + LX(N) + a] = 1.0;
                for (int k = 0; k < K; k++) KK[(size_t)r * nk + LL(c, k)] = -p->SS[a * K + k];
            }
            for (int k = 0; k < K; k++) KK[(size_t)(n + 6 * N + 6) * nk + LL(c, k)] = 1.0;
            for (int r = n; r < nk; r++)
                for (int col = 0; col < n; col++) KK[(size_t)col * nk + r] = KK[(size_t)r * nk + col];
            double *bb = Rl;                               /* rhs vector, nk entries */
            for (int k = 0; k < n; k++) bb[k] = rhs[k];
            for (int k = 0; k < me; k++) bb[n + k] = -ceq[k];
            /* LU with partial pivoting, in place */
            int sing = 0;
            for (int k = 0; k < nk && !sing; k++) {
                int piv = k;
                double best = fabs(KK[(size_t)k * nk + k]);
                for (int r = k + 1; r < nk; r++)
                    if (fabs(KK[(size_t)r * nk + k]) > best) { best = fabs(KK[(size_t)r * nk + k]); piv = r; }
                if (best < 1e-300) { sing = 1; break; }
                if (piv != k) {
                    for (int col = 0; col < nk; col++) { double t = KK[(size_t)k * nk + col]; KK[(size_t)k * nk + col] = KK[(size_t)piv * nk + col]; KK[(size_t)piv * nk + col] = t; }
                    double t = bb[k]; bb[k] = bb[piv]; bb[piv] = t;
                }
                double inv = 1.0 / KK[(size_t)k * nk + k];
                for (int r = k + 1; r < nk; r++) {
                    double f = KK[(size_t)r * nk + k] * inv;
                    if (f == 0.0) continue;
                    for (int col = k + 1; col < nk; col++) KK[(size_t)r * nk + col] -= f * KK[(size_t)k * nk + col];
                    bb[r] -= f * bb[k];
                }
            }
            if (sing) { status = 3; break; }
            for (int k = nk - 1; k >= 0; k--) {
                double sacc = bb[k];
                for (int col = k + 1; col < nk; col++) sacc -= KK[(size_t)k * nk + col] * bb[col];
                bb[k] = sacc / KK[(size_t)k * nk + k];
            }
            for (int k = 0; k < n; k++) dw[k] = bb[k];
            for (int k = 0; k < me; k++) lamn[k] = bb[n + k];
        }
        double a_max = 1.0, a_z = 1.0;
        for (int k = 0; k < n; k++) {
            dzL[k] = dzU[k] = 0.0;
            if (hasl(c, k)) {
                double d = w[k] - lb[k];
                dzL[k] = mu / d - zL[k] - zL[k] / d * dw[k];
                if (dw[k] < 0.0) a_max = dmin(a_max, -tau * d / dw[k]);
                if (dzL[k] < 0.0) a_z = dmin(a_z, -tau * zL[k] / dzL[k]);
            }
            if (hasu(c, k)) {
                double d = ub[k] - w[k];
                dzU[k] = mu / d - zU[k] + zU[k] / d * dw[k];
                if (dw[k] > 0.0) a_max = dmin(a_max, tau * d / dw[k]);
                if (dzU[k] < 0.0) a_z = dmin(a_z, -tau * zU[k] / dzU[k]);
            }
        }
        double th = 0.0;
        for (int k = 0; k < me; k++) th += fabs(ceq[k]);
        double ph = l_phi(c, w, mu);
        double gphi = 0.0;
        for (int k = 0; k < n; k++) {
            gphi += c->df * gradf[k] * dw[k];
            if (hasl(c, k)) gphi -= mu * dw[k] / (w[k] - lb[k]);
            if (hasu(c, k)) gphi += mu * dw[k] / (ub[k] - w[k]);
        }
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
            l_ceq(c, wt, ceqt);
            double tht = 0.0;
            for (int k = 0; k < me; k++) tht += fabs(ceqt[k]);
            double pht = l_phi(c, wt, mu);
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
        }
        if (!accepted) { status = 2; break; }
        if (!ftype) {
            if (nfilt == FILT_MAX) {
                memmove(filt_th, filt_th + 1, sizeof(double) * (FILT_MAX - 1));
                memmove(filt_ph, filt_ph + 1, sizeof(double) * (FILT_MAX - 1));
                nfilt--;
            }
            filt_th[nfilt] = (1.0 - gamma_theta) * th;
            filt_ph[nfilt] = ph - gamma_phi * th;
            nfilt++;
        }
        memcpy(w, wt, sizeof(double) * n);
        for (int k = 0; k < me; k++) lam[k] += a * (lamn[k] - lam[k]);
        for (int k = 0; k < n; k++) {
            if (hasl(c, k)) {
                double d = w[k] - lb[k];
                zL[k] += a_z * dzL[k];
                zL[k] = dmax(dmin(zL[k], kappa_sigma * mu / d), mu / (kappa_sigma * d));
            }
            if (hasu(c, k)) {
                double d = ub[k] - w[k];
                zU[k] += a_z * dzU[k];
                zU[k] = dmax(dmin(zU[k], kappa_sigma * mu / d), mu / (kappa_sigma * d));
            }
        }
        iter++;
    }
    memset(res, 0, sizeof(*res));
    memcpy(res->x, p->x0, sizeof(double) * 6);
    memcpy(res->x + 6, w, sizeof(double) * nx);
    memcpy(res->u, w + nx, sizeof(double) * nu);
    memcpy(res->lambda, w + nx + nu, sizeof(double) * K);
    res->cost = l_f(c, w);
    res->kkt_err = E0;
    res->status = status;
    res->iters = iter;
    free(buf);
    return 0;
}

typedef struct { const orc_lmpc_problem *p; const orc_options *o; orc_lmpc_result *r; int B, tid, nt; } ljob;
static void *lworker(void *arg) {
    ljob *j = (ljob *)arg;
    for (int b = j->tid; b < j->B; b += j->nt) orc_lmpc_solve(j->p + b, j->o, j->r + b);
    return NULL;
}
int orc_lmpc_solve_batch(const orc_lmpc_problem *p, int B, const orc_options *o, orc_lmpc_result *r, int nthreads) {
    if (nthreads <= 1) {
        for (int b = 0; b < B; b++) orc_lmpc_solve(p + b, o, r + b);
        return 0;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    ljob *jobs = (ljob *)malloc(sizeof(ljob) * nthreads);
    for (int k = 0; k < nthreads; k++) {
        jobs[k] = (ljob){p, o, r, B, k, nthreads};
        pthread_create(&th[k], NULL, lworker, &jobs[k]);
    }
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    free(th);
    free(jobs);
    return 0;
}

/* TEST INFRASTRUCTURE ONLY -- C restatement of the reference's iLQR solve,
 * car_racing/control/control.py:64-195 with car_racing/control/ilqr_helper.py:4-55.
 * Pinned against the reference itself: tests/golden/ilqr_golden.npz is produced by
 * tests/golden/make_ilqr_golden.py, which imports the unmodified reference control.ilqr.
 * Quirks reproduced on purpose: the roll-out cost excludes the
 * obstacle term (:113-122); V_x/V_xx start from stage N-1 (:143-144); b_ddot omits the
 * curvature of h (ilqr_helper.py:54); Q_uu is regularised through its eigenvalues but the
 * value update uses the unregularised Q_uu (:155-164); accept iff cost_new < cost (:181).
 */
#include "ocp_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

static void mat6v(const double *A, const double *x, double *y) { /* y = A x */
    for (int a = 0; a < 6; a++) {
        double s = 0.0;
        for (int b = 0; b < 6; b++) s += A[6 * a + b] * x[b];
        y[a] = s;
    }
}

static double quad6(const double *Q, const double *d) {
    double s = 0.0;
    double t[6];
    mat6v(Q, d, t);            /* numpy: d.T @ Q @ d evaluates (d.T @ Q) @ d; Q symmetric => same products */
    for (int a = 0; a < 6; a++) s += d[a] * t[a];
    return s;
}

static double rollout_cost(const orc_ilqr_problem *p, const double *x, const double *u) {
    int N = p->N;
    double cost = 0.0;
    for (int k = 0; k < N; k++) {
        double d[6];
        for (int a = 0; a < 6; a++) d[a] = x[6 * k + a] - p->xt[a];
        double ls = quad6(p->Q, d);
        const double *uk = u + 2 * k;
        double lc = uk[0] * (p->R[0] * uk[0] + p->R[1] * uk[1]) + uk[1] * (p->R[2] * uk[0] + p->R[3] * uk[1]);
        cost = cost + ls + lc;
    }
    double d[6];
    for (int a = 0; a < 6; a++) d[a] = x[6 * N + a] - p->xt[a];
    return cost + quad6(p->Q, d);
}

int orc_ilqr_solve(const orc_ilqr_problem *p, orc_ilqr_result *r) {
    int N = p->N;
    if (N < 1 || N > ORC_NMAX) return -1;
    const double eps = 0.01, lamb_factor = 10.0, max_lamb = 1000.0;
    const double q1 = 2.5, q2 = 2.5, margin = 0.15;
    double lamb = 1.0;
    static const int XD = 6;
    double *x = r->x, *u = r->u;
    double xn[(ORC_NMAX + 1) * 6], un[ORC_NMAX * 2];
    double lx[ORC_NMAX][6], lxx[ORC_NMAX][36], lu[ORC_NMAX][2];
    double Kg[ORC_NMAX][12], kf[ORC_NMAX][2];
    memset(u, 0, sizeof(double) * 2 * N);
    memset(x, 0, sizeof(double) * 6 * (N + 1));
    memcpy(x, p->x0, sizeof(double) * 6);
    double iL2 = 1.0 / (p->L * p->L), iW2 = 1.0 / (p->W * p->W);
    int it, conv = 0;
    double cost = 0.0;
    for (it = 0; it < p->max_iter; it++) {
        /* forward simulation (control.py:113-122) */
        for (int k = 0; k < N; k++) {
            double t[6];
            mat6v(p->A, x + 6 * k, t);
            for (int a = 0; a < 6; a++) x[6 * (k + 1) + a] = t[a] + (p->B[2 * a] * u[2 * k] + p->B[2 * a + 1] * u[2 * k + 1]);
        }
        cost = rollout_cost(p, x, u);
        /* cost derivatives (ilqr_helper.py:27-47) */
        for (int i = 0; i < N; i++) {
            lu[i][0] = 2 * (p->R[0] * u[2 * i] + p->R[1] * u[2 * i + 1]);
            lu[i][1] = 2 * (p->R[2] * u[2 * i] + p->R[3] * u[2 * i + 1]);
            double d[6], t[6];
            for (int a = 0; a < 6; a++) d[a] = x[6 * i + a] - p->xt[a];
            mat6v(p->Q, d, t);
            for (int a = 0; a < 6; a++) lx[i][a] = 2 * t[a];
            for (int a = 0; a < 36; a++) lxx[i][a] = 2 * p->Q[a];
            double ds = x[6 * i + 4] - p->obs_s[i] - p->lap_off;
            double de = x[6 * i + 5] - p->obs_ey[i];
            double h = 1 + margin - (ds * ds * iL2 + de * de * iW2);
            double hd4 = -2 * iL2 * ds, hd5 = -2 * iW2 * de;
            double ex = exp(q2 * h);
            lx[i][4] += q1 * q2 * ex * hd4;
            lx[i][5] += q1 * q2 * ex * hd5;
            double c2 = q1 * (q2 * q2) * ex;
            lxx[i][4 * XD + 4] += c2 * (hd4 * hd4);
            lxx[i][4 * XD + 5] += c2 * (hd4 * hd5);
            lxx[i][5 * XD + 4] += c2 * (hd5 * hd4);
            lxx[i][5 * XD + 5] += c2 * (hd5 * hd5);
        }
        /* backward pass (control.py:143-164) */
        double Vx[6], Vxx[36];
        memcpy(Vx, lx[N - 1], sizeof(Vx));
        memcpy(Vxx, lxx[N - 1], sizeof(Vxx));
        for (int i = N - 1; i >= 0; i--) {
            double Qx[6], Qu[2], Qxx[36], Quu[4], Qux[12], VA[36], VB[12];
            for (int a = 0; a < 6; a++) { /* VA = Vxx A, VB = Vxx B */
                for (int b = 0; b < 6; b++) {
                    double s = 0.0;
                    for (int e = 0; e < 6; e++) s += Vxx[6 * a + e] * p->A[6 * e + b];
                    VA[6 * a + b] = s;
                }
                for (int b = 0; b < 2; b++) {
                    double s = 0.0;
                    for (int e = 0; e < 6; e++) s += Vxx[6 * a + e] * p->B[2 * e + b];
                    VB[2 * a + b] = s;
                }
            }
            for (int a = 0; a < 6; a++) {
                double s = 0.0;
                for (int e = 0; e < 6; e++) s += p->A[6 * e + a] * Vx[e];
                Qx[a] = lx[i][a] + s;
                for (int b = 0; b < 6; b++) {
                    double s2 = 0.0;
                    for (int e = 0; e < 6; e++) s2 += p->A[6 * e + a] * VA[6 * e + b];
                    Qxx[6 * a + b] = lxx[i][6 * a + b] + s2;
                }
            }
            for (int a = 0; a < 2; a++) {
                double s = 0.0;
                for (int e = 0; e < 6; e++) s += p->B[2 * e + a] * Vx[e];
                Qu[a] = lu[i][a] + s;
                for (int b = 0; b < 2; b++) {
                    double s2 = 0.0;
                    for (int e = 0; e < 6; e++) s2 += p->B[2 * e + a] * VB[2 * e + b];
                    Quu[2 * a + b] = 2 * p->R[2 * a + b] + s2;
                }
                for (int b = 0; b < 6; b++) {
                    double s2 = 0.0;
                    for (int e = 0; e < 6; e++) s2 += p->B[2 * e + a] * VA[6 * e + b];
                    Qux[6 * a + b] = s2;
                }
            }
            /* eigen-regularised inverse of the 2x2 Q_uu (:155-158) */
            double a11 = Quu[0], a12 = 0.5 * (Quu[1] + Quu[2]), a22 = Quu[3];
            double tr = 0.5 * (a11 + a22), df = 0.5 * (a11 - a22);
            double rad = sqrt(df * df + a12 * a12);
            double e1 = tr + rad, e2 = tr - rad;
            double v1x, v1y;
            if (fabs(a12) > 0.0) {
                /* eigenvector of e1: (a12, e1 - a11) or (e1 - a22, a12), pick the better conditioned */
                if (fabs(e1 - a11) > fabs(e1 - a22)) { v1x = a12; v1y = e1 - a11; }
                else { v1x = e1 - a22; v1y = a12; }
                double nrm = sqrt(v1x * v1x + v1y * v1y);
                v1x /= nrm; v1y /= nrm;
            } else if (a11 >= a22) { v1x = 1.0; v1y = 0.0; }
            else { v1x = 0.0; v1y = 1.0; }
            double v2x = -v1y, v2y = v1x;
            if (e1 < 0.0) e1 = 0.0;
            if (e2 < 0.0) e2 = 0.0;
            e1 += lamb; e2 += lamb;
            double Qi[4];
            Qi[0] = v1x * v1x / e1 + v2x * v2x / e2;
            Qi[1] = v1x * v1y / e1 + v2x * v2y / e2;
            Qi[2] = Qi[1];
            Qi[3] = v1y * v1y / e1 + v2y * v2y / e2;
            kf[i][0] = -(Qi[0] * Qu[0] + Qi[1] * Qu[1]);
            kf[i][1] = -(Qi[2] * Qu[0] + Qi[3] * Qu[1]);
            for (int b = 0; b < 6; b++) {
                Kg[i][b] = -(Qi[0] * Qux[b] + Qi[1] * Qux[6 + b]);
                Kg[i][6 + b] = -(Qi[2] * Qux[b] + Qi[3] * Qux[6 + b]);
            }
            /* V_x = Q_x - K' Q_uu k ; V_xx = Q_xx - K' Q_uu K (:163-164) */
            double Qk0 = Quu[0] * kf[i][0] + Quu[1] * kf[i][1], Qk1 = Quu[2] * kf[i][0] + Quu[3] * kf[i][1];
            double QK[12];
            for (int b = 0; b < 6; b++) {
                QK[b] = Quu[0] * Kg[i][b] + Quu[1] * Kg[i][6 + b];
                QK[6 + b] = Quu[2] * Kg[i][b] + Quu[3] * Kg[i][6 + b];
            }
            for (int a = 0; a < 6; a++) {
                Vx[a] = Qx[a] - (Kg[i][a] * Qk0 + Kg[i][6 + a] * Qk1);
                for (int b = 0; b < 6; b++) Vxx[6 * a + b] = Qxx[6 * a + b] - (Kg[i][a] * QK[b] + Kg[i][6 + a] * QK[6 + b]);
            }
        }
        /* forward pass (control.py:166-180) */
        memcpy(xn, p->x0, sizeof(double) * 6);
        for (int i = 0; i < N; i++) {
            double dx[6];
            for (int a = 0; a < 6; a++) dx[a] = xn[6 * i + a] - x[6 * i + a];
            for (int a = 0; a < 2; a++) {
                double s = 0.0;
                for (int b = 0; b < 6; b++) s += Kg[i][6 * a + b] * dx[b];
                un[2 * i + a] = u[2 * i + a] + kf[i][a] + s;
            }
            double t[6];
            mat6v(p->A, xn + 6 * i, t);
            for (int a = 0; a < 6; a++) xn[6 * (i + 1) + a] = t[a] + (p->B[2 * a] * un[2 * i] + p->B[2 * a + 1] * un[2 * i + 1]);
        }
        double cost_new = rollout_cost(p, xn, un);
        if (cost_new < cost) {
            memcpy(x, xn, sizeof(double) * 6 * (N + 1));
            memcpy(u, un, sizeof(double) * 2 * N);
            lamb /= lamb_factor;
            if (fabs((cost_new - cost) / cost) < eps) { conv = 1; it++; cost = cost_new; break; }
        } else {
            lamb *= lamb_factor;
            if (lamb > max_lamb) { it++; break; }
        }
    }
    r->u0[0] = u[0];
    r->u0[1] = u[1];
    r->cost = cost;
    r->iters = it;
    r->converged = conv;
    return 0;
}

typedef struct { const orc_ilqr_problem *p; orc_ilqr_result *r; int B, tid, nt; } ijob_t;

static void *iworker(void *arg) {
    ijob_t *j = (ijob_t *)arg;
    for (int b = j->tid; b < j->B; b += j->nt) orc_ilqr_solve(j->p + b, j->r + b);
    return NULL;
}

int orc_ilqr_solve_batch(const orc_ilqr_problem *p, int B, orc_ilqr_result *r, int nthreads) {
    if (nthreads <= 1) {
        for (int b = 0; b < B; b++) orc_ilqr_solve(p + b, r + b);
        return 0;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    ijob_t *jobs = (ijob_t *)malloc(sizeof(ijob_t) * nthreads);
    for (int k = 0; k < nthreads; k++) {
        jobs[k] = (ijob_t){p, r, B, k, nthreads};
        pthread_create(&th[k], NULL, iworker, &jobs[k]);
    }
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    free(th);
    free(jobs);
    return 0;
}

// b200mpc: batched LMPC solve -- the QP of control.lmpc (car_racing/control/control.py:610-730):
//   vars  x_0..x_N, u_0..u_{N-1}, lambda (K safe-set weights); slack is forced to 0 by :693-694 and dropped
//   min   sum_{i<N}[(x_i-x_trk)'Q(.) + u_i'R u_i + (u_i-u_{i-1})'dR(.)] + (x_N-x_trk)'Q(.) + Qfun'lambda   (:667-695)
//   s.t.  x_0 = xcurv; x_{i+1} = A_i x_i + B_i u_i + C_i (LTV, :653-656); vx_i<=v_max, |ey_i|<=lap_width, i<N (:658-660)
//         |u| <= (delta_max, a_max) (:662-666); lambda >= 0, x_N = SS lambda, 1'lambda = 1 (:689-692)
// One CTA of 4 warps per instance (occupancy is set by shared memory: 79 KB at N=12, K=44 -> 2 CTAs/SM, so the extra
// warps are free), same interior-point definition as the MPC-CBF kernel (the QP is convex).
//
// Linear algebra: NOT a Riccati sweep.  Near the optimum only ~3 of the 44 lambdas leave their bound, so any stage-wise
// elimination of the terminal hull constraint meets a 7x7 Schur complement with condition number > 1e15 (measured in the
// oracle's first draft).  Instead the states are condensed through the LTV model once per solve (dx = G du + dx_p, G in
// shared memory) and each iteration solves a reduced KKT system by a CTA-level LU with partial pivoting (16 rows x 8 column
// lanes per pass).  Round 2: the system is no longer (du, dlambda, nu) with all K lambdas (75x75 at N=12, K=44) but
// (du, dlambda_F, nu) with only the NF = min(7, K) lambdas of largest lambda_k/z_k -- a vertex of the hull constraint has at
// most 7 positive weights (6 state rows + the simplex row) -- kept explicit: 38x38.  The other lambdas sit at their bound,
// their diagonal block z_k/lambda_k is large and well scaled, and they are eliminated exactly:
// dlambda_k = d_k (rhs_k - e_k' nu), d_k = lambda_k/z_k, which adds -sum_B d_k e_k e_k' (small, PSD) to the nu block.  The
// ill-conditioning that broke the 7x7 form came from the few HUGE d_k of the free lambdas; those are exactly the ones that
// stay in the pivoted system.  Half the elimination steps (each costs 2 block barriers), 1/8 of the flops, and the matrix
// shrinks from 45.6 KB to 11.9 KB of shared memory (2 -> 4 CTAs per SM).  Dynamics multipliers follow from the costate
// recursion.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200mpc.h"
#include "ocp_ipm.cuh"

namespace b200mpc {

struct LmpcKParams {
    b200mpc_lmpc_params p;
    b200mpc_ipm_options o;
    int32_t B, in_stride;
};

__host__ __device__ inline int lmpc_record_doubles(int N, int K) { return (8 + 54 * N + 7 * K + 1) & ~1; }

struct LmpcPlan {
    int N, K, NX, NU, NW, NR, ND, ME, NF;
    int oIN, oW, oD, oZL, oZU, oGF, oRHS, oSIG, oDP, oKD, oLAM, oLAMN, oCEQ, oG, oRQ, oKK, oPIV, oRED, oXS, oFI, oDK, total;
    __host__ __device__ LmpcPlan(int N_, int K_, int in_stride) {
        N = N_; K = K_;
        NX = 6 * (N + 1); NU = 2 * N; NW = NX + NU + K; ME = 6 * N + 7;
        NF = K < 7 ? K : 7;          // lambdas kept explicit in the pivoted system
        NR = NU + NF; ND = NR + 7;   // reduced KKT: (du, dlambda_F, nu)
        int o = 2;
        auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
        oIN = take(in_stride);
        oW = take(NW); oD = take(NW); oZL = take(NW); oZU = take(NW); oGF = take(NW); oRHS = take(NW); oSIG = take(NW);
        oDP = take(NX); oKD = take(NX);
        oLAM = take(ME); oLAMN = take(ME); oCEQ = take(ME);
        oG = take(6 * N * NU); oRQ = take(NU * NU);
        oKK = take(ND * (ND + 1)); oPIV = take(ND + 2); oRED = take(16); oXS = take(ND); oFI = take(8 + K); oDK = take(K);
        total = o;
    }
    __host__ __device__ size_t bytes() const { return (size_t)total * sizeof(double); }
};

constexpr int LMPC_NT = 128;   // threads per instance
#ifndef LMPC_MIN_CTAS
#define LMPC_MIN_CTAS 4        // 50.5 KB of shared memory per CTA at N=12, K=44 -> 4 CTAs fit an SM; cap the registers at 128 to match
#endif

__global__ void __launch_bounds__(LMPC_NT, LMPC_MIN_CTAS) lmpc_kernel(const __grid_constant__ LmpcKParams kp, const double *__restrict__ in,
                                                  b200mpc_record *__restrict__ rec, double *__restrict__ aux,
                                                  double *__restrict__ xpred, double *__restrict__ upred,
                                                  double *__restrict__ lambda_out, const XchgArgs xa = XchgArgs{nullptr, 0, 0, 0}) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NT = LMPC_NT;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, inst = blockIdx.x;
    const int N = kp.p.N, K = kp.p.K;
    const LmpcPlan pl(N, K, kp.in_stride);
    const int NX = pl.NX, NU = pl.NU, NW = pl.NW, NR = pl.NR, ND = pl.ND, ME = pl.ME, NF = pl.NF, OU = NX, OL = NX + NU, LD = ND + 1;
    double *IN = sm + pl.oIN, *W = sm + pl.oW, *D = sm + pl.oD, *ZL = sm + pl.oZL, *ZU = sm + pl.oZU, *GF = sm + pl.oGF;
    double *RHS = sm + pl.oRHS, *SIG = sm + pl.oSIG, *DP = sm + pl.oDP, *KD = sm + pl.oKD, *LAM = sm + pl.oLAM, *LAMN = sm + pl.oLAMN;
    double *CEQ = sm + pl.oCEQ, *G = sm + pl.oG, *RQ = sm + pl.oRQ, *KK = sm + pl.oKK, *PIV = sm + pl.oPIV, *RED = sm + pl.oRED, *XS = sm + pl.oXS;
    double *DK = sm + pl.oDK;
    int *FI = reinterpret_cast<int *>(sm + pl.oFI);     // FI[0..NF-1]: the explicit lambdas; FI[16 + k]: slot of lambda k in F or -1
    int red_phase = 0;
    // block reductions: one barrier each (two alternating scratch rows)
    auto bred = [&](double v, int op) -> double {
        v = (op == 0) ? warp_sum(v) : ((op == 1) ? warp_max(v) : warp_min(v));
        double *r = RED + 4 * (red_phase & 1) + ((red_phase & 2) ? 8 : 0);
        red_phase++;
        if (lane == 0) r[wid] = v;
        __syncthreads();
        if (op == 0) return (r[0] + r[1]) + (r[2] + r[3]);
        if (op == 1) return fmax(fmax(r[0], r[1]), fmax(r[2], r[3]));
        return fmin(fmin(r[0], r[1]), fmin(r[2], r[3]));
    };
    auto bsum = [&](double v) { return bred(v, 0); };
    auto bmax = [&](double v) { return bred(v, 1); };
    auto bmin = [&](double v) { return bred(v, 2); };
    const b200mpc_ipm_options &o = kp.o;

    uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
    const uint32_t in_bytes = (uint32_t)kp.in_stride * 8u;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, in_bytes);
        bulk_g2s(IN, in + (size_t)inst * kp.in_stride, in_bytes, bar);
    }
    for (int e = tid; e < NW; e += NT) { W[e] = 0.0; D[e] = 0.0; ZL[e] = 0.0; ZU[e] = 0.0; }
    for (int e = tid; e < ME; e += NT) { LAM[e] = 0.0; LAMN[e] = 0.0; }
    for (int e = tid; e < NX; e += NT) { DP[e] = 0.0; KD[e] = 0.0; }
    mbar_wait(bar, 0);
    __syncthreads();
    const double *x0 = IN, *u_old = IN + 6, *Am = IN + 8, *Bm = Am + 36 * N, *Cm = Bm + 12 * N, *SS = Cm + 6 * N, *Qf = SS + 6 * K;
    const double *Q = kp.p.Q, *R = kp.p.R, *dR = kp.p.dR, *xtrk = kp.p.xtrk;

    // ---- bounds as functions of the index in W
    auto has_l = [&](int e) -> bool {
        if (e < NX) { int i = e / 6, a = e - 6 * i; return i >= 1 && i < N && a == 5; }
        return true;   // u box, lambda >= 0
    };
    auto has_u = [&](int e) -> bool {
        if (e < NX) { int i = e / 6, a = e - 6 * i; return i >= 1 && i < N && (a == 0 || a == 5); }
        return e < OL;
    };
    auto lbv = [&](int e) -> double {
        if (e < NX) return -kp.p.width;
        if (e < OL) return -kp.p.umax[(e - OU) & 1];
        return 0.0;
    };
    auto ubv = [&](int e) -> double {
        if (e < NX) { int a = e % 6; return a == 0 ? kp.p.vmax : kp.p.width; }
        return kp.p.umax[(e - OU) & 1];
    };

    // ---- start: u = 0 roll-out through the LTV model, lambda = 1/K, pushed into the bounds
    if (wid == 0) {   // sequential recursions run on warp 0 with warp barriers
        if (lane < 6) W[lane] = x0[lane];
        __syncwarp();
        for (int i = 0; i < N; i++) {
            if (lane < 6) {
                double s = Cm[6 * i + lane];
                for (int b = 0; b < 6; b++) s += Am[36 * i + 6 * lane + b] * W[6 * i + b];
                W[6 * (i + 1) + lane] = s;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int k = tid; k < K; k += NT) W[OL + k] = 1.0 / (double)K;
    __syncthreads();
    int nbc = 0;
    for (int e = 6 + tid; e < NW; e += NT) {
        bool hl = has_l(e), hu = has_u(e);
        double w = W[e], lb = lbv(e), ub = ubv(e);
        if (hl) {
            double p_ = o.bound_push * fmax(1.0, fabs(lb));
            if (hu) p_ = fmin(p_, o.bound_frac * (ub - lb));
            if (w < lb + p_) w = lb + p_;
        }
        if (hu) {
            double p_ = o.bound_push * fmax(1.0, fabs(ub));
            if (hl) p_ = fmin(p_, o.bound_frac * (ub - lb));
            if (w > ub - p_) w = ub - p_;
        }
        W[e] = w;
        ZL[e] = hl ? 1.0 : 0.0;
        ZU[e] = hu ? 1.0 : 0.0;
        nbc += (hl ? 1 : 0) + (hu ? 1 : 0);
    }
    nbc = (int)(bsum((double)nbc) + 0.5);
    __syncthreads();
    // ---- G: x_i = sum_l G[6(i-1)+a][2l+b] u_l + ...   (rows for x_1..x_N)
    for (int e = tid; e < 6 * N * NU; e += NT) G[e] = 0.0;
    __syncthreads();
    for (int i = 0; i < N; i++) {   // row block of x_{i+1}
        for (int e = tid; e < 6 * NU; e += NT) {
            int a = e / NU, c = e - a * NU, l = c >> 1;
            double v = 0.0;
            if (l == i) v = Bm[12 * i + 2 * a + (c & 1)];
            else if (l < i) {
                for (int b = 0; b < 6; b++) v += Am[36 * i + 6 * a + b] * G[(6 * (i - 1) + b) * NU + c];
            }
            G[(6 * i + a) * NU + c] = v;
        }
        __syncthreads();
    }

    double df = 1.0, mu = o.mu_init;
    // ---- objective, gradient (into GF, unscaled), equality residual
    auto objective = [&](const double *Wp, double al, bool useD) -> double {
        double f = 0.0;
        for (int e = tid; e < 6 * (N + 1); e += NT) {
            int i = e / 6, a = e - 6 * i;
            double acc = 0.0, da = 0.0;
            for (int b = 0; b < 6; b++) {
                double d = Wp[6 * i + b] + (useD ? al * D[6 * i + b] : 0.0) - xtrk[b];
                acc += Q[6 * a + b] * d;
                if (b == a) da = d;
            }
            f += da * acc;
        }
        for (int i = tid; i < N; i += NT) {
            double u0 = Wp[OU + 2 * i] + (useD ? al * D[OU + 2 * i] : 0.0), u1 = Wp[OU + 2 * i + 1] + (useD ? al * D[OU + 2 * i + 1] : 0.0);
            double p0 = i ? Wp[OU + 2 * i - 2] + (useD ? al * D[OU + 2 * i - 2] : 0.0) : u_old[0];
            double p1 = i ? Wp[OU + 2 * i - 1] + (useD ? al * D[OU + 2 * i - 1] : 0.0) : u_old[1];
            double d0 = u0 - p0, d1 = u1 - p1;
            f += u0 * (R[0] * u0 + R[1] * u1) + u1 * (R[2] * u0 + R[3] * u1) + d0 * (dR[0] * d0 + dR[1] * d1) + d1 * (dR[2] * d0 + dR[3] * d1);
        }
        for (int k = tid; k < K; k += NT) f += Qf[k] * (Wp[OL + k] + (useD ? al * D[OL + k] : 0.0));
        return bsum(f);
    };
    auto gradient = [&]() {   // GF <- unscaled gradient of f at W
        for (int e = 6 + tid; e < NW; e += NT) {
            double g;
            if (e < NX) {
                int i = e / 6, a = e - 6 * i;
                g = 0.0;
                for (int b = 0; b < 6; b++) g += (Q[6 * a + b] + Q[6 * b + a]) * (W[6 * i + b] - xtrk[b]);
            } else if (e < OL) {
                int i = (e - OU) >> 1, a = (e - OU) & 1;
                double u0 = W[OU + 2 * i], u1 = W[OU + 2 * i + 1];
                double p0 = i ? W[OU + 2 * i - 2] : u_old[0], p1 = i ? W[OU + 2 * i - 1] : u_old[1];
                double d0 = u0 - p0, d1 = u1 - p1;
                g = (R[2 * a] + R[a]) * u0 + (R[2 * a + 1] + R[2 + a]) * u1 + (dR[2 * a] + dR[a]) * d0 + (dR[2 * a + 1] + dR[2 + a]) * d1;
                if (i < N - 1) {
                    double n0 = W[OU + 2 * i + 2] - u0, n1 = W[OU + 2 * i + 3] - u1;
                    g -= (dR[2 * a] + dR[a]) * n0 + (dR[2 * a + 1] + dR[2 + a]) * n1;
                }
            } else
                g = Qf[e - OL];
            GF[e] = g;
        }
    };
    auto residual = [&](double al, bool useD, double *out) -> double {   // equality residuals at W+al*D; returns the 1-norm
        double th = 0.0;
        for (int e = tid; e < ME; e += NT) {
            double s;
            if (e < 6 * N) {
                int i = e / 6, a = e - 6 * i;
                s = W[6 * (i + 1) + a] + (useD ? al * D[6 * (i + 1) + a] : 0.0) - Cm[6 * i + a];
                for (int b = 0; b < 6; b++) s -= Am[36 * i + 6 * a + b] * (W[6 * i + b] + (useD ? al * D[6 * i + b] : 0.0));
                s -= Bm[12 * i + 2 * a] * (W[OU + 2 * i] + (useD ? al * D[OU + 2 * i] : 0.0)) +
                     Bm[12 * i + 2 * a + 1] * (W[OU + 2 * i + 1] + (useD ? al * D[OU + 2 * i + 1] : 0.0));
            } else if (e < 6 * N + 6) {
                int a = e - 6 * N;
                s = W[6 * N + a] + (useD ? al * D[6 * N + a] : 0.0);
                for (int k = 0; k < K; k++) s -= SS[a * K + k] * (W[OL + k] + (useD ? al * D[OL + k] : 0.0));
            } else {
                s = -1.0;
                for (int k = 0; k < K; k++) s += W[OL + k] + (useD ? al * D[OL + k] : 0.0);
            }
            if (out) out[e] = s;
            th += fabs(s);
        }
        return bsum(th);
    };
    auto barrier = [&](double al, bool useD) -> double {
        LogAcc la;
        for (int e = 6 + tid; e < NW; e += NT) {
            double w = W[e] + (useD ? al * D[e] : 0.0);
            if (has_l(e)) la.mul(w - lbv(e));
            if (has_u(e)) la.mul(ubv(e) - w);
        }
        return bsum(la.value());
    };
    // optimality error (same scaling as the oracle) for barrier parameter 0 (-> e0) and m (-> em) in ONE pass: dual and
    // primal residuals do not depend on it (one set of block reductions instead of two); needs GF, CEQ current
    auto kkt_error2 = [&](double m, double &e0, double &em) {
        double dual = 0.0, prim = 0.0, comp = 0.0, compm = 0.0, zsum = 0.0, ysum = 0.0;
        for (int e = 6 + tid; e < NW; e += NT) {
            double rw = df * GF[e] - ZL[e] + ZU[e];
            if (e < NX) {
                int i = e / 6, a = e - 6 * i;
                rw += LAM[6 * (i - 1) + a];
                if (i < N) { for (int b = 0; b < 6; b++) rw -= Am[36 * i + 6 * b + a] * LAM[6 * i + b]; }
                else rw += LAM[6 * N + a];
            } else if (e < OL) {
                int i = (e - OU) >> 1, a = (e - OU) & 1;
                for (int b = 0; b < 6; b++) rw -= Bm[12 * i + 2 * b + a] * LAM[6 * i + b];
            } else {
                int k = e - OL;
                rw += LAM[6 * N + 6];
                for (int a = 0; a < 6; a++) rw -= SS[a * K + k] * LAM[6 * N + a];
            }
            dual = fmax(dual, fabs(rw));
            if (has_l(e)) { double v = (W[e] - lbv(e)) * ZL[e]; comp = fmax(comp, fabs(v)); compm = fmax(compm, fabs(v - m)); zsum += ZL[e]; }
            if (has_u(e)) { double v = (ubv(e) - W[e]) * ZU[e]; comp = fmax(comp, fabs(v)); compm = fmax(compm, fabs(v - m)); zsum += ZU[e]; }
        }
        for (int e = tid; e < ME; e += NT) { prim = fmax(prim, fabs(CEQ[e])); ysum += fabs(LAM[e]); }
        dual = bmax(dual); prim = bmax(prim); comp = bmax(comp); compm = bmax(compm); zsum = bsum(zsum); ysum = bsum(ysum);
        const double s_max = 100.0;
        int nmul = ME + nbc;
        double sd = fmax(s_max, (ysum + zsum) / (double)(nmul > 0 ? nmul : 1)) / s_max;
        double sc = fmax(s_max, zsum / (double)(nbc > 0 ? nbc : 1)) / s_max;
        e0 = fmax(dual / sd, fmax(prim, comp / sc));
        em = fmax(dual / sd, fmax(prim, compm / sc));
    };

    // ---- scaling
    gradient();
    __syncthreads();
    {
        double gm = 0.0;
        for (int e = 6 + tid; e < NW; e += NT) gm = fmax(gm, fabs(GF[e]));
        gm = bmax(gm);
        df = gm > o.max_grad ? o.max_grad / gm : 1.0;
    }
    // RQ = sum_i G_i' (df*(Q+Q')) G_i  (iteration independent)
    for (int e = tid; e < NU * NU; e += NT) {
        int a = e / NU, b = e - a * NU;
        double s = 0.0;
        for (int i = 0; i < N; i++)
            for (int p_ = 0; p_ < 6; p_++) {
                double t = 0.0;
                for (int q_ = 0; q_ < 6; q_++) t += (Q[6 * p_ + q_] + Q[6 * q_ + p_]) * G[(6 * i + q_) * NU + b];
                s += G[(6 * i + p_) * NU + a] * t;
            }
        RQ[e] = df * s;
    }
    __syncthreads();
    double th0 = residual(0.0, false, CEQ);
    __syncthreads();
    const double theta_max = 1e4 * fmax(1.0, th0), theta_min = 1e-4 * fmax(1.0, th0);
    double f_th0 = 0.0, f_ph0 = 0.0, f_th1 = 0.0, f_ph1 = 0.0;
    int nfilt = 0, fpos = 0, iter = 0, status = B200MPC_MAX_ITER, n_acc = 0, n_back = 0;
    double E0 = 0.0;
    const double kappa_eps = 10.0, kappa_mu = 0.2, tau_min = 0.99;
    const double gamma_theta = 1e-5, gamma_phi = 1e-8, delta_sw = 1.0, s_theta = 1.1, s_phi = 2.3, eta_phi = 1e-8;
    const double gamma_alpha = 0.05, kappa_sigma = 1e10;

    for (;;) {
        gradient();
        residual(0.0, false, CEQ);
        __syncthreads();
        double em;
        kkt_error2(mu, E0, em);
        if (E0 <= o.tol) { status = B200MPC_SOLVED; break; }
        if (E0 <= o.acceptable_tol) {
            if (++n_acc >= o.acceptable_iter) { status = B200MPC_SOLVED; break; }
        } else
            n_acc = 0;
        if (iter >= o.max_iter) { status = B200MPC_MAX_ITER; break; }
        for (;;) {
            if (em <= kappa_eps * mu && mu > o.tol / 11.0) {
                mu = fmax(o.tol / 11.0, fmin(kappa_mu * mu, mu * sqrt(mu)));
                nfilt = 0;
                fpos = 0;
                double e0_;
                kkt_error2(mu, e0_, em);
            } else
                break;
        }
        const double tau = fmax(tau_min, 1.0 - mu);
        // ---- barrier diagonal and right-hand side
        for (int e = 6 + tid; e < NW; e += NT) {
            double sw = 0.0, b = -df * GF[e];
            if (has_l(e)) { double id = rcp(W[e] - lbv(e)); sw += ZL[e] * id; b += mu * id; }
            if (has_u(e)) { double id = rcp(ubv(e) - W[e]); sw += ZU[e] * id; b -= mu * id; }
            SIG[e] = sw;
            RHS[e] = b;
        }
        // particular solution of the dynamics rows (du = 0)
        if (wid == 0) {
            if (lane < 6) DP[lane] = 0.0;
            __syncwarp();
            for (int i = 0; i < N; i++) {
                if (lane < 6) {
                    double s = -CEQ[6 * i + lane];
                    for (int b = 0; b < 6; b++) s += Am[36 * i + 6 * lane + b] * DP[6 * i + b];
                    DP[6 * (i + 1) + lane] = s;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // KD = K_xx dp  (x part of K applied to the particular solution)
        for (int e = 6 + tid; e < NX; e += NT) {
            int i = e / 6, a = e - 6 * i;
            double s = SIG[e] * DP[e];
            for (int b = 0; b < 6; b++) s += df * (Q[6 * a + b] + Q[6 * b + a]) * DP[6 * i + b];
            KD[e] = s;
        }
        __syncthreads();
        // ---- the NF lambdas that stay explicit: largest d_k = lambda_k/z_k = smallest SIG (ties -> lower index); warp 0,
        //      NF rounds of warp-argmin over at most 2 candidates per lane
        if (wid == 0) {
            double c0 = (lane < K) ? SIG[OL + lane] : 1e300 * 1e300, c1 = (lane + 32 < K) ? SIG[OL + lane + 32] : 1e300 * 1e300;
            for (int k = lane; k < K; k += 32) FI[16 + k] = -1;
            __syncwarp();
            for (int f = 0; f < NF; f++) {
                double best = (c0 <= c1) ? c0 : c1;
                int bi = (c0 <= c1) ? lane : lane + 32;
                for (int off = 16; off > 0; off >>= 1) {
                    double ob = __shfl_xor_sync(0xffffffffu, best, off);
                    int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if (lane == (bi & 31)) {
                    if (bi < 32) c0 = 1e300 * 1e300; else c1 = 1e300 * 1e300;
                    FI[f] = bi;
                    FI[16 + bi] = f;
                }
            }
        }
        // ---- reduced KKT matrix  [Rh 0 Eu'; 0 Sig_F E_F'; Eu E_F -S_B | rhs]
        for (int e = tid; e < ND * LD; e += NT) KK[e] = 0.0;
        for (int k = tid; k < K; k += NT) DK[k] = rcp(SIG[OL + k]);     // d_k = lambda_k / z_k, once per iteration
        __syncthreads();
        for (int e = tid; e < NU * NU; e += NT) {   // Rh_uu = RQ + G' diag(sig_x) G + K_uu
            int a = e / NU, b = e - a * NU;
            double s = RQ[e];
            // G' diag(sig_x) G: sig_x is nonzero only on the bounded states, vx_i and ey_i of the stages 0 < i < N (has_l / has_u)
            for (int i = 1; i < N; i++) {
                const int r0 = 6 * (i - 1), r5 = r0 + 5;
                s += G[r0 * NU + a] * SIG[6 + r0] * G[r0 * NU + b] + G[r5 * NU + a] * SIG[6 + r5] * G[r5 * NU + b];
            }
            int ia = a >> 1, ib = b >> 1, ca = a & 1, cb = b & 1;
            double r2 = R[2 * ca + cb] + R[2 * cb + ca], d2 = dR[2 * ca + cb] + dR[2 * cb + ca];
            if (ia == ib) s += df * (r2 + d2 + ((ia < N - 1) ? d2 : 0.0)) + ((a == b) ? SIG[OU + a] : 0.0);
            else if (ia == ib + 1 || ib == ia + 1) s -= df * d2;
            KK[a * LD + b] = s;
        }
        // explicit lambdas F (written by warp 0 before the barrier above): diagonal z/lambda, their columns e_k = (-SS[:,k]; 1)
        for (int f = tid; f < NF; f += NT) {
            const int k = FI[f];
            KK[(NU + f) * LD + NU + f] = SIG[OL + k];
            for (int a = 0; a < 6; a++) {
                KK[(NU + f) * LD + NR + a] = -SS[a * K + k];
                KK[(NR + a) * LD + NU + f] = -SS[a * K + k];
            }
            KK[(NU + f) * LD + NR + 6] = 1.0;
            KK[(NR + 6) * LD + NU + f] = 1.0;
            KK[(NU + f) * LD + ND] = RHS[OL + k];
        }
        // eliminated lambdas B: -S_B = -sum_B d_k e_k e_k' into the nu block, -sum_B d_k e_k rhs_k into its right-hand side
        for (int e = tid; e < 7 * 8; e += NT) {
            const int a = e >> 3, b = e & 7;     // b == 7: the right-hand side column
            double acc = 0.0;
            for (int k = 0; k < K; k++) {
                if (FI[16 + k] >= 0) continue;
                const double dk = DK[k];
                const double ea = (a < 6) ? -SS[a * K + k] : 1.0;
                const double eb = (b < 6) ? -SS[b * K + k] : ((b == 6) ? 1.0 : RHS[OL + k]);
                acc += dk * ea * eb;
            }
            if (b < 7) KK[(NR + a) * LD + NR + b] = -acc;
            else XS[a] = acc;                    // picked up below when the nu right-hand side is written
        }
        for (int e = tid; e < 6 * NU; e += NT) {
            int a = e / NU, c = e - a * NU;
            double v = G[(6 * (N - 1) + a) * NU + c];
            KK[(NR + a) * LD + c] = v;
            KK[c * LD + NR + a] = v;
        }
        for (int c = tid; c < NU; c += NT) {   // rr_u = rhs_u + G'(rhs_x - K_xx dp)
            double s = RHS[OU + c];
            for (int r = 0; r < 6 * N; r++) s += G[r * NU + c] * (RHS[6 + r] - KD[6 + r]);
            KK[c * LD + ND] = s;
        }
        __syncthreads();     // XS from the B loop
        if (tid < 6) KK[(NR + tid) * LD + ND] = -(CEQ[6 * N + tid] + DP[6 * N + tid]) - XS[tid];
        if (tid == 6) KK[(NR + 6) * LD + ND] = -CEQ[6 * N + 6] - XS[6];
        __syncthreads();
        // ---- LU with partial pivoting on the augmented matrix (ND x ND+1), then column-oriented back substitution
        //      (measured and rejected in round 2: rows addressed through a permutation instead of swapped, the pivot found by
        //      every warp, one block barrier per step, back substitution on one warp -- 519 k -> 468 k solves/s at B = 512 with
        //      6 batches in flight, 0.83 -> 0.88 ms for one instance: the index loads sit on the critical path of every step)
        bool singular = false;
        for (int k = 0; k < ND; k++) {
            if (wid == 0) {   // pivot search: warp 0 over the rows k..ND-1 of column k
                double best = -1.0;
                int bi = k;
                for (int r = k + lane; r < ND; r += 32) {
                    double v = fabs(KK[r * LD + k]);
                    if (v > best) { best = v; bi = r; }
                }
                for (int off = 16; off > 0; off >>= 1) {
                    double ob = __shfl_xor_sync(0xffffffffu, best, off);
                    int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if (lane == 0) { PIV[ND] = (double)bi; PIV[ND + 1] = best; }
            }
            __syncthreads();
            const int bi = (int)PIV[ND];
            if (!(PIV[ND + 1] > 1e-300)) { singular = true; break; }
            if (bi != k) {
                for (int c = k + tid; c < LD; c += NT) {
                    double t = KK[k * LD + c];
                    KK[k * LD + c] = KK[bi * LD + c];
                    KK[bi * LD + c] = t;
                }
                __syncthreads();
            }
            const double inv = rcp(KK[k * LD + k]);
            if (tid == 0) PIV[k] = inv;
            // eliminate: 16 rows per pass, 8 column lanes per row
            for (int r = k + 1 + (tid >> 3); r < ND; r += NT / 8) {
                double f = KK[r * LD + k] * inv;
                if (f != 0.0)
                    for (int c = k + 1 + (tid & 7); c < LD; c += 8) KK[r * LD + c] -= f * KK[k * LD + c];
            }
            __syncthreads();
        }
        if (singular) { status = B200MPC_INERTIA; break; }
        for (int k = ND - 1; k >= 0; k--) {   // x_k = rhs_k / u_kk, then rhs_r -= u_rk x_k for r < k; solution overwrites the rhs column
            double xk = KK[k * LD + ND] * PIV[k];
            for (int r = tid; r < k; r += NT) KK[r * LD + ND] -= KK[r * LD + k] * xk;
            if (tid == 0) XS[k] = xk;
            __syncthreads();
        }
        for (int e = tid; e < ND; e += NT) KK[e * LD + ND] = XS[e];
        __syncthreads();
        // ---- direction: du, dlambda from the solve; dx = G du + dp; nu+ ; costate recursion for the dynamics multipliers
        for (int e = tid; e < NU; e += NT) D[OU + e] = KK[e * LD + ND];
        for (int k = tid; k < K; k += NT) {   // dlambda: explicit ones from the solve, the eliminated ones d_k (rhs_k - e_k' nu)
            const int f = FI[16 + k];
            double v;
            if (f >= 0) v = KK[(NU + f) * LD + ND];
            else {
                double t = RHS[OL + k] - KK[(NR + 6) * LD + ND];
                for (int a = 0; a < 6; a++) t += SS[a * K + k] * KK[(NR + a) * LD + ND];
                v = t * DK[k];
            }
            D[OL + k] = v;
        }
        for (int e = tid; e < 7; e += NT) LAMN[6 * N + e] = KK[(NR + e) * LD + ND];
        __syncthreads();
        for (int e = 6 + tid; e < NX; e += NT) {
            double s = DP[e];
            for (int c = 0; c < NU; c++) s += G[(e - 6) * NU + c] * D[OU + c];
            D[e] = s;
        }
        if (tid < 6) D[tid] = 0.0;
        __syncthreads();
        // res_x = rhs_x - (K d)_x  -> KD
        for (int e = 6 + tid; e < NX; e += NT) {
            int i = e / 6, a = e - 6 * i;
            double s = SIG[e] * D[e];
            for (int b = 0; b < 6; b++) s += df * (Q[6 * a + b] + Q[6 * b + a]) * D[6 * i + b];
            KD[e] = RHS[e] - s;
        }
        __syncthreads();
        if (wid == 0) {
            for (int i = N; i >= 1; i--) {
                if (lane < 6) {
                    double s = KD[6 * i + lane];
                    if (i < N) { for (int b = 0; b < 6; b++) s += Am[36 * i + 6 * b + lane] * LAMN[6 * i + b]; }
                    else s -= LAMN[6 * N + lane];
                    LAMN[6 * (i - 1) + lane] = s;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- step sizes, directional derivative
        double a_max = 1.0, a_z = 1.0, gphi = 0.0;
        for (int e = 6 + tid; e < NW; e += NT) {
            double d = D[e], w = W[e];
            gphi += df * GF[e] * d;
            if (has_l(e)) {
                double dl = w - lbv(e), zl = ZL[e];
                double dzl = (mu - zl * d) * rcp(dl) - zl;
                if (d < 0.0) a_max = fmin(a_max, -tau * dl / d);
                if (dzl < 0.0) a_z = fmin(a_z, -tau * zl / dzl);
                gphi -= mu * d * rcp(dl);
            }
            if (has_u(e)) {
                double du = ubv(e) - w, zu = ZU[e];
                double dzu = (mu + zu * d) * rcp(du) - zu;
                if (d > 0.0) a_max = fmin(a_max, tau * du / d);
                if (dzu < 0.0) a_z = fmin(a_z, -tau * zu / dzu);
                gphi += mu * d * rcp(du);
            }
        }
        a_max = bmin(a_max); a_z = bmin(a_z); gphi = bsum(gphi);
        double th = 0.0;
        for (int e = tid; e < ME; e += NT) th += fabs(CEQ[e]);
        th = bsum(th);
        double ph = df * objective(W, 0.0, false) - mu * barrier(0.0, false);
        double amin;
        if (gphi < 0.0 && th <= theta_min)
            amin = gamma_alpha * fmin(gamma_theta, fmin(gamma_phi * th / (-gphi), delta_sw * pow(th, s_theta) / pow(-gphi, s_phi)));
        else if (gphi < 0.0)
            amin = gamma_alpha * fmin(gamma_theta, gamma_phi * th / (-gphi));
        else
            amin = gamma_alpha * gamma_theta;
        double a = a_max;
        bool accepted = false, ftype = false;
        int nls = 0;
        while (a >= amin || nls == 0) {
            double tht = residual(a, true, nullptr);
            double pht = df * objective(W, a, true) - mu * barrier(a, true);
            bool dom = false;
            if (lane < nfilt && tht >= f_th0 && pht >= f_ph0) dom = true;
            if (lane + 32 < nfilt && tht >= f_th1 && pht >= f_ph1) dom = true;
            bool okf = (tht < theta_max) && !__any_sync(0xffffffffu, dom);
            if (okf) {
                bool sw = gphi < 0.0 && a * pow(-gphi, s_phi) > delta_sw * pow(th, s_theta);
                if (th <= theta_min && sw) {
                    if (pht <= ph + eta_phi * a * gphi) { accepted = true; ftype = true; }
                } else if (tht <= (1.0 - gamma_theta) * th || pht <= ph - gamma_phi * th)
                    accepted = true;
            }
            if (accepted) break;
            a *= 0.5;
            nls++;
            n_back++;
        }
        if (!accepted) { status = B200MPC_LINESEARCH; break; }
        if (!ftype) {
            double nth = (1.0 - gamma_theta) * th, nph = ph - gamma_phi * th;
            int slot = fpos >> 5, ln = fpos & 31;
            if (lane == ln) {
                if (slot == 0) { f_th0 = nth; f_ph0 = nph; }
                else { f_th1 = nth; f_ph1 = nph; }
            }
            fpos = (fpos + 1) & 63;
            if (nfilt < 64) nfilt++;
        }
        // ---- update
        for (int e = 6 + tid; e < NW; e += NT) {
            double d = D[e], w = W[e], wn = w + a * d;
            if (has_l(e)) {
                double dl = w - lbv(e), zl = ZL[e];
                zl += a_z * ((mu - zl * d) * rcp(dl) - zl);
                double m1 = mu * rcp(wn - lbv(e));
                ZL[e] = fmax(fmin(zl, kappa_sigma * m1), m1 * (1.0 / kappa_sigma));
            }
            if (has_u(e)) {
                double du = ubv(e) - w, zu = ZU[e];
                zu += a_z * ((mu + zu * d) * rcp(du) - zu);
                double m1 = mu * rcp(ubv(e) - wn);
                ZU[e] = fmax(fmin(zu, kappa_sigma * m1), m1 * (1.0 / kappa_sigma));
            }
            W[e] = wn;
        }
        for (int e = tid; e < ME; e += NT) LAM[e] += a * (LAMN[e] - LAM[e]);
        __syncthreads();
        iter++;
    }
    double cost = objective(W, 0.0, false);
    if (tid < 32) {     // warp 0 (convergent here): result record, and its copy into every rank's gathered buffer
        b200mpc_record rc;
        rc.cost = cost;
        rc.u0[0] = W[OU];
        rc.u0[1] = W[OU + 1];
        rc.status = status;
        rc.iters = iter;
        if (tid == 0) rec[inst] = rc;
        xchg_publish(xa, tid, inst, rc);
    }
    if (aux != nullptr && tid < 4) aux[(size_t)inst * 4 + tid] = (tid == 0) ? E0 : (tid == 3 ? (double)n_back : 0.0);
    if (xpred != nullptr)
        for (int e = tid; e < NX; e += NT) xpred[(size_t)inst * NX + e] = W[e];
    if (upred != nullptr)
        for (int e = tid; e < NU; e += NT) upred[(size_t)inst * NU + e] = W[OU + e];
    if (lambda_out != nullptr)
        for (int e = tid; e < K; e += NT) lambda_out[(size_t)inst * K + e] = W[OL + e];
}

}  // namespace b200mpc

// b200mpc: batched iLQR, one warp per instance -- the iteration of the reference's
// control.ilqr (car_racing/control/control.py:64-195) with the cost derivatives of
// ilqr_helper.get_cost_derivation / repelling_cost_function (ilqr_helper.py:4-55),
// including its quirks: roll-out cost without the obstacle term
// (:113-122), V_x/V_xx seeded from stage N-1 (:143-144), b_ddot without the curvature of h
// (ilqr_helper.py:54), eigenvalue regularisation of Q_uu for the gains but the raw Q_uu in
// the value update (:155-164), accept iff cost_new < cost (:181), lambda x/÷ 10, stop at
// lambda > 1000 (:184-191).  LTI dynamics, no input/state bounds.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200mpc.h"
#include "ocp_ipm.cuh"

namespace b200mpc {

struct IlqrKParams {
    b200mpc_ilqr_params p;
    int32_t B;
    int32_t in_stride;
};

__host__ __device__ inline int ilqr_record_doubles(int N) { return (14 + 2 * (N + 1) + 1) & ~1; }

struct IlqrPlan {
    int N, oIN, oA, oB, oQ, oR, oX, oU, oXN, oUN, oKG, oKF, oLX, oLXX, oVXX, oVX, oVA, oVB, oQXX, oQUX, oQX, oQU, oQUU, total;
    __host__ __device__ IlqrPlan(int N_, int in_stride) {
        N = N_;
        int o = 2;
        auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
        oIN = take(in_stride);
        oA = take(36); oB = take(12); oQ = take(36); oR = take(4);
        oX = take(6 * (N + 1)); oU = take(2 * N); oXN = take(6 * (N + 1)); oUN = take(2 * N);
        oKG = take(12 * N); oKF = take(2 * N); oLX = take(6 * N); oLXX = take(4 * N);
        oVXX = take(36); oVX = take(6); oVA = take(36); oVB = take(12);
        oQXX = take(36); oQUX = take(12); oQX = take(6); oQU = take(2); oQUU = take(4);
        total = o;
    }
    __host__ __device__ size_t bytes() const { return (size_t)total * sizeof(double); }
};

__global__ void __launch_bounds__(32) ilqr_kernel(const __grid_constant__ IlqrKParams kp, const double *__restrict__ in,
                                                  b200mpc_record *__restrict__ rec, double *__restrict__ xpred,
                                                  double *__restrict__ upred, const XchgArgs xa = XchgArgs{nullptr, 0, 0, 0}) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x, inst = blockIdx.x;
    const int N = kp.p.N;
    const IlqrPlan pl(N, kp.in_stride);
    double *IN = sm + pl.oIN, *cA = sm + pl.oA, *cB = sm + pl.oB, *cQ = sm + pl.oQ, *cR = sm + pl.oR;
    double *X = sm + pl.oX, *U = sm + pl.oU, *XN = sm + pl.oXN, *UN = sm + pl.oUN, *KG = sm + pl.oKG, *KF = sm + pl.oKF;
    double *LX = sm + pl.oLX, *LXX = sm + pl.oLXX, *VXX = sm + pl.oVXX, *VX = sm + pl.oVX, *VA = sm + pl.oVA, *VB = sm + pl.oVB;
    double *QXX = sm + pl.oQXX, *QUX = sm + pl.oQUX, *QX = sm + pl.oQX, *QU = sm + pl.oQU, *QUU = sm + pl.oQUU;

    uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
    const uint32_t in_bytes = (uint32_t)kp.in_stride * 8u;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(bar, in_bytes);
        bulk_g2s(IN, in + (size_t)inst * kp.in_stride, in_bytes, bar);
    }
    for (int e = lane; e < 36; e += 32) { cA[e] = kp.p.A[e]; cQ[e] = kp.p.Q[e]; }
    if (lane < 12) cB[lane] = kp.p.B[lane];
    if (lane < 4) cR[lane] = kp.p.R[lane];
    mbar_wait(bar, 0);
    __syncwarp();
    const double *x0 = IN, *xt = IN + 6;
    const double lap_off = IN[12];
    const double *obs_s = IN + 14, *obs_e = IN + 14 + (N + 1);
    const double iL2 = 1.0 / (kp.p.L * kp.p.L), iW2 = 1.0 / (kp.p.W * kp.p.W);
    const double q1 = 2.5, q2 = 2.5, margin = 0.15, eps = 0.01, lamb_factor = 10.0, max_lamb = 1000.0;

    for (int e = lane; e < 2 * N; e += 32) U[e] = 0.0;
    if (lane < 6) { X[lane] = x0[lane]; XN[lane] = x0[lane]; }
    __syncwarp();

    // roll-out cost of (xs, us)  (control.py:113-122)
    auto traj_cost = [&](const double *xs, const double *us) -> double {
        double c = 0.0;
        for (int e = lane; e < 6 * (N + 1); e += 32) {
            int i = e / 6, a = e - 6 * i;
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < 6; b++) acc += cQ[6 * a + b] * (xs[6 * i + b] - xt[b]);
            c += (xs[6 * i + a] - xt[a]) * acc;
        }
        for (int e = lane; e < 2 * N; e += 32) {
            int i = e >> 1, a = e & 1;
            c += us[2 * i + a] * (cR[2 * a] * us[2 * i] + cR[2 * a + 1] * us[2 * i + 1]);
        }
        return warp_sum(c);
    };

    double lamb = 1.0, cost = 0.0;
    int it = 0, conv = 0;
    for (it = 0; it < kp.p.max_iter; it++) {
        // forward simulation with the current inputs (control.py:114-115)
        for (int k = 0; k < N; k++) {
            if (lane < 6) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < 6; b++) s += cA[6 * lane + b] * X[6 * k + b];
                X[6 * (k + 1) + lane] = s + (cB[2 * lane] * U[2 * k] + cB[2 * lane + 1] * U[2 * k + 1]);
            }
            __syncwarp();
        }
        cost = traj_cost(X, U);
        // cost derivatives (ilqr_helper.py:27-47); l_xx = 2Q + LXX on the (s,ey) block
        for (int i = lane; i < N; i += 32) {
            double ds = X[6 * i + 4] - obs_s[i] - lap_off, de = X[6 * i + 5] - obs_e[i];
            double h = 1.0 + margin - (ds * ds * iL2 + de * de * iW2);
            double hd4 = -2.0 * iL2 * ds, hd5 = -2.0 * iW2 * de;
            double ex = exp(q2 * h);
            double c1 = q1 * q2 * ex, c2 = q1 * (q2 * q2) * ex;
#pragma unroll
            for (int a = 0; a < 6; a++) {
                double acc = 0.0;
#pragma unroll
                for (int b = 0; b < 6; b++) acc += cQ[6 * a + b] * (X[6 * i + b] - xt[b]);
                double v = 2.0 * acc;
                if (a == 4) v += c1 * hd4;
                if (a == 5) v += c1 * hd5;
                LX[6 * i + a] = v;
            }
            LXX[4 * i + 0] = c2 * (hd4 * hd4);
            LXX[4 * i + 1] = c2 * (hd4 * hd5);
            LXX[4 * i + 2] = c2 * (hd5 * hd5);
        }
        __syncwarp();
        // backward pass (control.py:143-164)
        for (int e = lane; e < 36; e += 32) {
            int a = e / 6, b = e - 6 * a;
            double v = 2.0 * cQ[e];
            if (a == 4 && b == 4) v += LXX[4 * (N - 1) + 0];
            if ((a == 4 && b == 5) || (a == 5 && b == 4)) v += LXX[4 * (N - 1) + 1];
            if (a == 5 && b == 5) v += LXX[4 * (N - 1) + 2];
            VXX[e] = v;
        }
        if (lane < 6) VX[lane] = LX[6 * (N - 1) + lane];
        __syncwarp();
        for (int i = N - 1; i >= 0; i--) {
            for (int e = lane; e < 48; e += 32) {
                if (e < 36) {
                    int a = e / 6, b = e - 6 * a;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) s += VXX[6 * a + c] * cA[6 * c + b];
                    VA[e] = s;
                } else {
                    int f = e - 36, a = f >> 1, b = f & 1;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) s += VXX[6 * a + c] * cB[2 * c + b];
                    VB[f] = s;
                }
            }
            __syncwarp();
            for (int e = lane; e < 60; e += 32) {
                if (e < 36) {
                    int a = e / 6, b = e - 6 * a;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) s += cA[6 * c + a] * VA[6 * c + b];
                    double l = 2.0 * cQ[e];
                    if (a == 4 && b == 4) l += LXX[4 * i + 0];
                    if ((a == 4 && b == 5) || (a == 5 && b == 4)) l += LXX[4 * i + 1];
                    if (a == 5 && b == 5) l += LXX[4 * i + 2];
                    QXX[e] = l + s;
                } else if (e < 48) {
                    int f = e - 36, a = f / 6, b = f - 6 * a;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) s += cB[2 * c + a] * VA[6 * c + b];
                    QUX[f] = s;
                } else if (e < 52) {
                    int f = e - 48, a = f >> 1, b = f & 1;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) s += cB[2 * c + a] * VB[2 * c + b];
                    QUU[f] = 2.0 * cR[f] + s;
                } else if (e < 58) {
                    int a = e - 52;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) s += cA[6 * c + a] * VX[c];
                    QX[a] = LX[6 * i + a] + s;
                } else {
                    int a = e - 58;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) s += cB[2 * c + a] * VX[c];
                    QU[a] = 2.0 * (cR[2 * a] * U[2 * i] + cR[2 * a + 1] * U[2 * i + 1]) + s;
                }
            }
            __syncwarp();
            // eigen-regularised inverse of Q_uu (control.py:155-158), every lane redundantly
            double a11 = QUU[0], a12 = 0.5 * (QUU[1] + QUU[2]), a22 = QUU[3];
            double tr = 0.5 * (a11 + a22), dfh = 0.5 * (a11 - a22);
            double rad = sqrt(dfh * dfh + a12 * a12);
            double e1 = tr + rad, e2 = tr - rad;
            double v1x, v1y;
            if (fabs(a12) > 0.0) {
                if (fabs(e1 - a11) > fabs(e1 - a22)) { v1x = a12; v1y = e1 - a11; }
                else { v1x = e1 - a22; v1y = a12; }
                double nrm = sqrt(v1x * v1x + v1y * v1y);
                v1x /= nrm;
                v1y /= nrm;
            } else if (a11 >= a22) { v1x = 1.0; v1y = 0.0; }
            else { v1x = 0.0; v1y = 1.0; }
            double v2x = -v1y, v2y = v1x;
            if (e1 < 0.0) e1 = 0.0;
            if (e2 < 0.0) e2 = 0.0;
            e1 += lamb;
            e2 += lamb;
            double Qi0 = v1x * v1x / e1 + v2x * v2x / e2;
            double Qi1 = v1x * v1y / e1 + v2x * v2y / e2;
            double Qi3 = v1y * v1y / e1 + v2y * v2y / e2;
            if (lane < 12) {
                int c = lane / 6, b = lane - 6 * c;
                KG[12 * i + lane] = -((c ? Qi1 : Qi0) * QUX[b] + (c ? Qi3 : Qi1) * QUX[6 + b]);
            } else if (lane < 14) {
                int c = lane - 12;
                KF[2 * i + c] = -((c ? Qi1 : Qi0) * QU[0] + (c ? Qi3 : Qi1) * QU[1]);
            }
            __syncwarp();
            // V_x = Q_x - K' Q_uu k, V_xx = Q_xx - K' Q_uu K (control.py:163-164)
            for (int e = lane; e < 42; e += 32) {
                const double *Kc = KG + 12 * i;
                if (e < 36) {
                    int a = e / 6, b = e - 6 * a;
                    double qk0 = QUU[0] * Kc[b] + QUU[1] * Kc[6 + b], qk1 = QUU[2] * Kc[b] + QUU[3] * Kc[6 + b];
                    VXX[e] = QXX[e] - (Kc[a] * qk0 + Kc[6 + a] * qk1);
                } else {
                    int a = e - 36;
                    double qk0 = QUU[0] * KF[2 * i] + QUU[1] * KF[2 * i + 1], qk1 = QUU[2] * KF[2 * i] + QUU[3] * KF[2 * i + 1];
                    VX[a] = QX[a] - (Kc[a] * qk0 + Kc[6 + a] * qk1);
                }
            }
            __syncwarp();
        }
        // forward pass with feedback (control.py:166-180)
        for (int i = 0; i < N; i++) {
            if (lane < 2) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < 6; b++) s += KG[12 * i + 6 * lane + b] * (XN[6 * i + b] - X[6 * i + b]);
                UN[2 * i + lane] = U[2 * i + lane] + KF[2 * i + lane] + s;
            }
            __syncwarp();
            if (lane < 6) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < 6; b++) s += cA[6 * lane + b] * XN[6 * i + b];
                XN[6 * (i + 1) + lane] = s + (cB[2 * lane] * UN[2 * i] + cB[2 * lane + 1] * UN[2 * i + 1]);
            }
            __syncwarp();
        }
        double cost_new = traj_cost(XN, UN);
        if (cost_new < cost) {
            for (int e = lane; e < 6 * (N + 1); e += 32) X[e] = XN[e];
            for (int e = lane; e < 2 * N; e += 32) U[e] = UN[e];
            __syncwarp();
            lamb /= lamb_factor;
            if (fabs((cost_new - cost) / cost) < eps) { conv = 1; it++; cost = cost_new; break; }
        } else {
            lamb *= lamb_factor;
            if (lamb > max_lamb) { it++; break; }
        }
    }
    {
        b200mpc_record rc;
        rc.cost = cost;
        rc.u0[0] = U[0];
        rc.u0[1] = U[1];
        rc.status = conv ? 0 : 1;
        rc.iters = it;
        if (lane == 0) rec[inst] = rc;
        xchg_publish(xa, lane, inst, rc);
    }
    if (xpred != nullptr)
        for (int e = lane; e < 6 * (N + 1); e += 32) xpred[(size_t)inst * 6 * (N + 1) + e] = X[e];
    if (upred != nullptr)
        for (int e = lane; e < 2 * N; e += 32) upred[(size_t)inst * 2 * N + e] = U[e];
}

// first-min argmin over records (overtake_traj_planner.py:244: list.index(min(...)))
__global__ void argmin_cost_kernel(const b200mpc_record *__restrict__ rec, int B, int max_status, int32_t *__restrict__ out) {
    __shared__ double sc[32];
    __shared__ int si[32];
    double best = 1e300 * 1e300;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        if (rec[i].status <= max_status) {
            double c = rec[i].cost;
            if (c < best || (c == best && i < bi)) { best = c; bi = i; }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) { sc[w] = best; si[w] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < nw; k++)
            if (sc[k] < best || (sc[k] == best && si[k] < bi)) { best = sc[k]; bi = si[k]; }
        *out = (bi == 0x7fffffff) ? -1 : bi;
    }
}

}  // namespace b200mpc

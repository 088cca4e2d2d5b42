// b200mpc: batched interior-point solver for the reference's MPC-LTI / MPC-CBF /
// mpc_multi_agents NLP (car_racing/control/control.py:198-248, 476-607, 251-473).
//
// One warp (= one CTA of 32 threads) solves one problem instance.  All solver state lives in
// shared memory (~31 KB at N=20, M=3) so 7 instances are resident per SM; HBM is touched
// twice: a TMA bulk copy (cp.async.bulk -> UBLKCP) of the packed 1136-byte instance record
// at the start and the coalesced result record at the end.  FP64 throughout; no tensor cores
// (KKT blocks are 6..14 wide and banded -- see DESIGN.md).
//
// Algorithm (DESIGN.md "Solver definition"; same definition as oracle/ocp_oracle.c, different
// linear algebra): primal-dual barrier method with IPOPT's conventions -- monotone mu,
// fraction-to-boundary, filter line search, inertia correction, gradient-based scaling,
// E_0 <= tol termination -- l1-elastic CBF rows instead of a restoration phase.  The Newton
// step is a Riccati recursion over the horizon on the augmented stage state
// X_k = (x_k, sigma_k), U_k = (u_k, sigma_{k+1}); a non-positive pivot in any stage's
// Cholesky is the inertia test.
//
// Execution style: the warp works in two modes.  Entry-parallel phases (assembly of the stage
// Hessian, residuals, line-search evaluation) give every lane a few matrix entries with
// compile-time trip counts and branch-free, uniform code.  The sequential recursions
// (Cholesky of the 5x5 pivot block, forward sweep, costate sweep) are computed redundantly in
// every lane's registers from broadcast shared-memory loads, so they need no warp barrier.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200mpc.h"

namespace b200mpc {

struct KParams {
    b200mpc_cbf_params p;
    b200mpc_ipm_options o;
    int32_t B;
    int32_t in_stride;  // doubles per instance record
    int32_t hdr;        // offset of xtarget inside the record
    int32_t obs_off;    // offset of the rival block inside the record
    double iL6, iW6;    // 1/L^6, 1/W^6
};

__host__ __device__ inline int cbf_hdr_doubles(int M) { return (6 + M + 1) & ~1; }
__host__ __device__ inline int cbf_record_doubles(int N, int M, int xt_per_stage) {
    int n = cbf_hdr_doubles(M) + (xt_per_stage ? 6 * (N + 1) : 6) + 2 * M * (N + 1);
    return (n + 1) & ~1;
}

// ---------------------------------------------------------------- shared memory plan
template <int M>
struct SmemPlan {
    static constexpr int NXA = 6 + M, NUA = 2 + M, NZ = NXA + NUA;
    int N, R, NB, NW, OU, OS;
    int oTM, oQQ, oRR, oQ, oR, oIN, oW, oD, oHD, oZL, oZU, oS, oT, oY, oZ, oV, oDG, oG, oSIGE, oYHAT, oJD, oJA, oLAM,
        oCRES, oKFB, oKFF, oP, oPV, oPT, oGM, oGV, oCT, oYF, oQV, oDS, oGS, total;
    __host__ __device__ SmemPlan(int N_, int in_stride) {
        N = N_;
        R = M * N;
        NB = 4 * N + M * (N + 1);
        OU = 6 * (N + 1);
        OS = OU + 2 * N;
        NW = OS + M * (N + 1);
        int o = 2;  // doubles 0..1: mbarrier (8 B) + pad, keeps everything after 16-byte aligned
        auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
        oIN = take(in_stride);
        oTM = take(NXA * NZ); oQQ = take(36); oRR = take(4); oQ = take(36); oR = take(4);
        oW = take(NW); oD = take(NW); oHD = take(NW);
        oZL = take(NB); oZU = take(4 * N);
        oS = take(R); oT = take(R); oY = take(R); oZ = take(R); oV = take(R);
        oDG = take(R); oG = take(R); oSIGE = take(R); oYHAT = take(R); oJD = take(R);
        oJA = take(4 * R);
        oLAM = take(6 * N); oCRES = take(6 * N);
        oKFB = take(N * NUA * NXA); oKFF = take(N * NUA);
        oP = take(NXA * NXA); oPV = take(NXA); oPT = take(NXA * NZ);
        oGM = take(NZ * NZ); oGV = take(NZ); oCT = take((M > 0 ? M : 1) * NZ);
        oYF = take(NUA * (NXA + 1)); oQV = take(NXA); oDS = take(NZ); oGS = take(NZ);
        total = o;
    }
    __host__ __device__ size_t bytes() const { return (size_t)total * sizeof(double); }
};

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double p4(double a) { double b = a * a; return b * b; }
__device__ __forceinline__ double p5(double a) { double b = a * a; return b * b * a; }
__device__ __forceinline__ double p6(double a) { double b = a * a; return b * b * b; }

// sum of logs of positive numbers without one log() per term: multiply mantissas, add exponents
struct LogAcc {
    double m = 1.0;
    int e = 0;
    __device__ __forceinline__ void mul(double v) {
        int ex;
        m *= frexp(v, &ex);
        e += ex;
        if (m < 1e-200) { m = frexp(m, &ex); e += ex; }
    }
    __device__ __forceinline__ double value() const { return log(m) + (double)e * 0.693147180559945309417232; }
};

// ---------------------------------------------------------------- TMA (bulk async copy) + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// number of 32-wide rounds needed to cover n entries (compile-time when n is)
__host__ __device__ constexpr int rounds(int n) { return (n + 31) / 32; }

// ---------------------------------------------------------------- the solver
// NT > 0: horizon known at compile time (entry loops fully unrolled); NT == 0: runtime horizon.
template <int NT, int M>
struct Ipm {
    static constexpr int NXA = 6 + M, NUA = 2 + M, NZ = NXA + NUA;
    const KParams &kp;
    const int lane, N, R, NB, NW, OU, OS;
    double *TM, *QQ, *RR, *cQ, *cR, *IN, *W, *D, *HD, *ZL, *ZU, *S, *T, *Y, *Z, *V, *DG, *GR, *SIGE, *YHAT, *JD, *JA, *LAM,
        *CRES, *KFB, *KFF, *P, *PV, *PT, *GM, *GV, *CT, *YF, *QV, *DS, *GS;
    const double *xt, *obs, *lapoff;
    double df, mu, rho, a1;  // a1 = 1 - alpha

    __device__ Ipm(const KParams &kp_, const SmemPlan<M> &pl, double *sm, int lane_)
        : kp(kp_), lane(lane_), N(pl.N), R(pl.R), NB(pl.NB), NW(pl.NW), OU(pl.OU), OS(pl.OS) {
        TM = sm + pl.oTM; QQ = sm + pl.oQQ; RR = sm + pl.oRR; cQ = sm + pl.oQ; cR = sm + pl.oR; IN = sm + pl.oIN;
        W = sm + pl.oW; D = sm + pl.oD; HD = sm + pl.oHD; ZL = sm + pl.oZL; ZU = sm + pl.oZU;
        S = sm + pl.oS; T = sm + pl.oT; Y = sm + pl.oY; Z = sm + pl.oZ; V = sm + pl.oV;
        DG = sm + pl.oDG; GR = sm + pl.oG; SIGE = sm + pl.oSIGE; YHAT = sm + pl.oYHAT; JD = sm + pl.oJD;
        JA = sm + pl.oJA; LAM = sm + pl.oLAM; CRES = sm + pl.oCRES; KFB = sm + pl.oKFB; KFF = sm + pl.oKFF;
        P = sm + pl.oP; PV = sm + pl.oPV; PT = sm + pl.oPT; GM = sm + pl.oGM; GV = sm + pl.oGV; CT = sm + pl.oCT;
        YF = sm + pl.oYF; QV = sm + pl.oQV; DS = sm + pl.oDS; GS = sm + pl.oGS;
        lapoff = IN + 6;
        xt = IN + kp.hdr;
        obs = IN + kp.obs_off;
        a1 = 1.0 - kp.p.alpha;
        rho = kp.o.rho;
        df = 1.0;
        mu = kp.o.mu_init;
    }

    // A[a][b] and B[a][c] live inside the dense stage map TM (rows 0..5)
    __device__ __forceinline__ double Am(int a, int b) const { return TM[a * NZ + b]; }
    __device__ __forceinline__ double Bm(int a, int c) const { return TM[a * NZ + NXA + c]; }

    // ---- indexing
    __device__ __forceinline__ double xtv(int i, int a) const { return kp.p.xt_per_stage ? xt[6 * i + a] : xt[a]; }
    __device__ __forceinline__ double obs_s(int j, int i) const { return obs[(2 * j) * (N + 1) + i]; }
    __device__ __forceinline__ double obs_e(int j, int i) const { return obs[(2 * j + 1) * (N + 1) + i]; }
    __device__ __forceinline__ int isg(int j, int i) const { return OS + j * (N + 1) + i; }
    // bounded variable b -> index in W and its bounds (ZU has slots only for b < 4N)
    __device__ __forceinline__ void bvar(int b, int &wi, double &lb, double &ub, bool &hasU) const {
        if (b < 2 * N) {
            int i = 1 + (b >> 1);
            bool ey = b & 1;
            wi = 6 * i + (ey ? 5 : 0);
            lb = ey ? -kp.p.width : kp.p.vmin;
            ub = ey ? kp.p.width : kp.p.vmax;
            hasU = true;
        } else if (b < 4 * N) {
            int e = b - 2 * N;
            wi = OU + e;
            ub = kp.p.umax[e & 1];
            lb = -ub;
            hasU = true;
        } else {
            wi = OS + (b - 4 * N);
            lb = 0.0;
            ub = 0.0;
            hasU = false;
        }
    }
    // index of the bound-multiplier slot of primal entry wi, or -1
    __device__ __forceinline__ int bslot(int wi) const {
        if (wi < OU) {
            int i = wi / 6, a = wi - 6 * i;
            if (i == 0 || (a != 0 && a != 5)) return -1;
            return 2 * (i - 1) + (a == 5);
        }
        if (wi < OS) return 2 * N + (wi - OU);
        return 4 * N + (wi - OS);
    }

    // ---- row r=(j,i) evaluated at W + al*D : unscaled pieces
    struct RowV { double ds, de, dsn, den, sg, sgn; };
    template <bool useD>
    __device__ __forceinline__ RowV row_vals(int j, int i, double al) const {
        RowV v;
        double xs = W[6 * i + 4], xe = W[6 * i + 5], xsn = W[6 * i + 10], xen = W[6 * i + 11];
        v.sg = W[isg(j, i)];
        v.sgn = W[isg(j, i + 1)];
        if (useD) {
            xs += al * D[6 * i + 4]; xe += al * D[6 * i + 5]; xsn += al * D[6 * i + 10]; xen += al * D[6 * i + 11];
            v.sg += al * D[isg(j, i)];
            v.sgn += al * D[isg(j, i + 1)];
        }
        v.ds = xs - obs_s(j, i) - lapoff[j];   // control.py:539-540 (with lap offset)
        v.de = xe - obs_e(j, i);
        v.dsn = xsn - obs_s(j, i + 1);         // control.py:542 (quirk: no lap offset)
        v.den = xen - obs_e(j, i + 1);
        return v;
    }
    __device__ __forceinline__ double row_g(const RowV &v) const {  // unscaled h_next - (1-alpha) h  (control.py:558)
        double h = p6(v.ds) * kp.iL6 + p6(v.de) * kp.iW6 - 1.0 - kp.p.margin - v.sg;
        double hn = p6(v.dsn) * kp.iL6 + p6(v.den) * kp.iW6 - 1.0 - kp.p.margin - v.sgn;
        return hn - a1 * h;
    }

    // ---- dynamics residual c_k = x_{k+1} - A x_k - B u_k at W + al*D, entry e=(k,a)
    template <bool useD>
    __device__ __forceinline__ double dyn_res(int e, double al) const {
        int k = e / 6, a = e - 6 * k;
        double s = W[6 * (k + 1) + a];
        if (useD) s += al * D[6 * (k + 1) + a];
#pragma unroll
        for (int b = 0; b < 6; b++) {
            double xv = W[6 * k + b];
            if (useD) xv += al * D[6 * k + b];
            s -= Am(a, b) * xv;
        }
        double u0 = W[OU + 2 * k], u1 = W[OU + 2 * k + 1];
        if (useD) { u0 += al * D[OU + 2 * k]; u1 += al * D[OU + 2 * k + 1]; }
        s -= Bm(a, 0) * u0 + Bm(a, 1) * u1;
        return s;
    }

    // ---- constraint violation theta (1-norm) and barrier function phi at W + al*D, slacks moved by al
    template <bool useD>
    __device__ void theta_phi(double al, double &theta, double &phi) const {
        double th = 0.0, f = 0.0, tsum = 0.0;
        LogAcc la;
#pragma unroll
        for (int t = 0; t < (NT ? rounds(6 * NT) : 1); t++)
            for (int e = lane + 32 * t; e < 6 * N; e += (NT ? 1 << 30 : 32)) th += fabs(dyn_res<useD>(e, al));
#pragma unroll
        for (int t = 0; t < (NT ? rounds(M * NT) : 1); t++)
            for (int r = lane + 32 * t; r < R; r += (NT ? 1 << 30 : 32)) {
                int j = r / N, i = r - j * N;
                RowV v = row_vals<useD>(j, i, al);
                double g = DG[r] * row_g(v);
                double s = S[r], tt = T[r];
                if (useD) {
                    double ds, dt, dy, dz, dv;
                    row_step(r, ds, dt, dy, dz, dv);
                    s += al * ds;
                    tt += al * dt;
                }
                th += fabs(g + tt - s);
                la.mul(s);
                la.mul(tt);
                tsum += tt;
            }
        // objective: stage terms (control.py:588-591), input terms (:578-579), slack penalty (:560,562)
#pragma unroll
        for (int t = 0; t < (NT ? rounds(6 * (NT + 1)) : 1); t++)
            for (int e = lane + 32 * t; e < 6 * (N + 1); e += (NT ? 1 << 30 : 32)) {
                int i = e / 6, a = e - 6 * i;
                double acc = 0.0, da = 0.0;
#pragma unroll
                for (int b = 0; b < 6; b++) {
                    double d = W[6 * i + b] - xtv(i, b);
                    if (useD) d += al * D[6 * i + b];
                    acc += cQ[6 * a + b] * d;
                    da = (b == a) ? d : da;
                }
                f += da * acc;
            }
#pragma unroll
        for (int t = 0; t < (NT ? rounds(2 * NT) : 1); t++)
            for (int e = lane + 32 * t; e < 2 * N; e += (NT ? 1 << 30 : 32)) {
                int i = e >> 1, a = e & 1;
                double u0 = W[OU + 2 * i], u1 = W[OU + 2 * i + 1];
                if (useD) { u0 += al * D[OU + 2 * i]; u1 += al * D[OU + 2 * i + 1]; }
                f += (a ? u1 : u0) * (cR[2 * a] * u0 + cR[2 * a + 1] * u1);
            }
        double ss = 0.0;
#pragma unroll
        for (int t = 0; t < (NT ? rounds(M * (NT + 1)) : 1); t++)
            for (int e = lane + 32 * t; e < M * (N + 1); e += (NT ? 1 << 30 : 32)) {
                double sv = W[OS + e];
                if (useD) sv += al * D[OS + e];
                ss += sv;
            }
        f += kp.p.slack_w * ss;
#pragma unroll
        for (int t = 0; t < (NT ? rounds(4 * NT + M * (NT + 1)) : 1); t++)
            for (int b = lane + 32 * t; b < NB; b += (NT ? 1 << 30 : 32)) {
                int wi; double lb, ub; bool hu;
                bvar(b, wi, lb, ub, hu);
                double w = W[wi];
                if (useD) w += al * D[wi];
                la.mul(w - lb);
                if (hu) la.mul(ub - w);
            }
        theta = warp_sum(th);
        phi = df * warp_sum(f) + rho * warp_sum(tsum) - mu * warp_sum(la.value());
    }

    // unscaled objective at W
    __device__ double objective() const {
        double f = 0.0;
        for (int e = lane; e < 6 * (N + 1); e += 32) {
            int i = e / 6, a = e - 6 * i;
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < 6; b++) acc += cQ[6 * a + b] * (W[6 * i + b] - xtv(i, b));
            f += (W[6 * i + a] - xtv(i, a)) * acc;
        }
        for (int e = lane; e < 2 * N; e += 32) {
            int i = e >> 1, a = e & 1;
            double u0 = W[OU + 2 * i], u1 = W[OU + 2 * i + 1];
            f += (a ? u1 : u0) * (cR[2 * a] * u0 + cR[2 * a + 1] * u1);
        }
        double ss = 0.0;
        for (int e = lane; e < M * (N + 1); e += 32) ss += W[OS + e];
        return warp_sum(f + kp.p.slack_w * ss);
    }

    // ---- Newton steps of the row slacks / multipliers from JD (all at the current iterate)
    __device__ __forceinline__ void row_step(int r, double &ds, double &dt, double &dy, double &dz, double &dv) const {
        double s = S[r], t = T[r], z = Z[r], v = V[r];
        double is = 1.0 / s, it = 1.0 / t;
        double sig_s = z * is, sig_t = v * it;
        double rg = GR[r] + t - s;
        double jd = JD[r];
        dy = YHAT[r] - SIGE[r] * jd - Y[r];
        dt = (mu * is + mu * it - rho - sig_s * rg - sig_s * jd) / (sig_s + sig_t);
        dv = mu * it - v - sig_t * dt;
        ds = jd + dt + rg;
        dz = mu * is - z - sig_s * ds;
    }

    // ---- evaluate rows (GR, JA) and dynamics residual at the current iterate
    __device__ void eval_point() {
#pragma unroll
        for (int t = 0; t < (NT ? rounds(6 * NT) : 1); t++)
            for (int e = lane + 32 * t; e < 6 * N; e += (NT ? 1 << 30 : 32)) CRES[e] = dyn_res<false>(e, 0.0);
#pragma unroll
        for (int t = 0; t < (NT ? rounds(M * NT) : 1); t++)
            for (int r = lane + 32 * t; r < R; r += (NT ? 1 << 30 : 32)) {
                int j = r / N, i = r - j * N;
                RowV v = row_vals<false>(j, i, 0.0);
                double sc = DG[r];
                GR[r] = sc * row_g(v);
                JA[4 * r + 0] = sc * (-a1 * 6.0 * p5(v.ds) * kp.iL6);
                JA[4 * r + 1] = sc * (-a1 * 6.0 * p5(v.de) * kp.iW6);
                JA[4 * r + 2] = sc * (6.0 * p5(v.dsn) * kp.iL6);
                JA[4 * r + 3] = sc * (6.0 * p5(v.den) * kp.iW6);
            }
        __syncwarp();
    }

    // gradient of the scaled objective wrt primal entry wi (wi >= 6)
    __device__ __forceinline__ double grad_f(int wi) const {
        if (wi < OU) {
            int i = wi / 6, a = wi - 6 * i;
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < 6; b++) acc += QQ[6 * a + b] * (W[6 * i + b] - xtv(i, b));
            return acc;
        }
        if (wi < OS) {
            int e = wi - OU, i = e >> 1, a = e & 1;
            return RR[2 * a] * W[OU + 2 * i] + RR[2 * a + 1] * W[OU + 2 * i + 1];
        }
        return df * kp.p.slack_w;
    }

    // J^T yv for primal entry wi (gather over the rows that touch it)
    __device__ __forceinline__ double jt_times(int wi, const double *yv) const {
        double acc = 0.0;
        if (M == 0) return 0.0;
        if (wi < OU) {
            int i = wi / 6, a = wi - 6 * i;
            if (a < 4) return 0.0;
#pragma unroll
            for (int j = 0; j < M; j++) {
                if (i < N) acc += JA[4 * (j * N + i) + (a - 4)] * yv[j * N + i];
                if (i >= 1) acc += JA[4 * (j * N + i - 1) + 2 + (a - 4)] * yv[j * N + i - 1];
            }
            return acc;
        }
        if (wi < OS) return 0.0;
        int e = wi - OS, j = e / (N + 1), i = e - j * (N + 1);
        if (i < N) acc += DG[j * N + i] * a1 * yv[j * N + i];
        if (i >= 1) acc -= DG[j * N + i - 1] * yv[j * N + i - 1];
        return acc;
    }

    // ---- optimality error pieces that do not depend on mu
    struct Err { double dual, prim, ysum, zsum; };
    __device__ Err error_base() const {
        double dual = 0.0, prim = 0.0, ysum = 0.0, zsum = 0.0;
#pragma unroll
        for (int t = 0; t < (NT ? rounds(8 * NT + M * (NT + 1)) : 1); t++)
            for (int wi = 6 + lane + 32 * t; wi < NW; wi += (NT ? 1 << 30 : 32)) {
                double rw = grad_f(wi) - jt_times(wi, Y);
                int bs = bslot(wi);
                if (bs >= 0) rw += -ZL[bs] + ((wi < OS) ? ZU[bs] : 0.0);
                if (wi < OU) {  // + lam_i - A' lam_{i+1}
                    int i = wi / 6, a = wi - 6 * i;
                    double s = LAM[6 * (i - 1) + a];
                    if (i < N) {
#pragma unroll
                        for (int b = 0; b < 6; b++) s -= Am(b, a) * LAM[6 * i + b];
                    }
                    rw += s;
                } else if (wi < OS) {  // - B' lam_{i+1}
                    int e = wi - OU, i = e >> 1, a = e & 1;
                    double s = 0.0;
#pragma unroll
                    for (int b = 0; b < 6; b++) s += Bm(b, a) * LAM[6 * i + b];
                    rw -= s;
                }
                dual = fmax(dual, fabs(rw));
            }
#pragma unroll
        for (int t = 0; t < (NT ? rounds(M * NT) : 1); t++)
            for (int r = lane + 32 * t; r < R; r += (NT ? 1 << 30 : 32)) {
                dual = fmax(dual, fabs(Y[r] - Z[r]));
                dual = fmax(dual, fabs(rho - Y[r] - V[r]));
                prim = fmax(prim, fabs(GR[r] + T[r] - S[r]));
                zsum += Z[r] + V[r];
                ysum += fabs(Y[r]);
            }
#pragma unroll
        for (int t = 0; t < (NT ? rounds(6 * NT) : 1); t++)
            for (int e = lane + 32 * t; e < 6 * N; e += (NT ? 1 << 30 : 32)) {
                prim = fmax(prim, fabs(CRES[e]));
                ysum += fabs(LAM[e]);
            }
#pragma unroll
        for (int t = 0; t < (NT ? rounds(4 * NT + M * (NT + 1)) : 1); t++)
            for (int b = lane + 32 * t; b < NB; b += (NT ? 1 << 30 : 32)) zsum += ZL[b] + ((b < 4 * N) ? ZU[b] : 0.0);
        Err e;
        e.dual = warp_max(dual);
        e.prim = warp_max(prim);
        e.ysum = warp_sum(ysum);
        e.zsum = warp_sum(zsum);
        return e;
    }
    __device__ double comp_err(double m) const {
        double c = 0.0;
#pragma unroll
        for (int t = 0; t < (NT ? rounds(4 * NT + M * (NT + 1)) : 1); t++)
            for (int b = lane + 32 * t; b < NB; b += (NT ? 1 << 30 : 32)) {
                int wi; double lb, ub; bool hu;
                bvar(b, wi, lb, ub, hu);
                c = fmax(c, fabs((W[wi] - lb) * ZL[b] - m));
                if (hu) c = fmax(c, fabs((ub - W[wi]) * ZU[b] - m));
            }
#pragma unroll
        for (int t = 0; t < (NT ? rounds(M * NT) : 1); t++)
            for (int r = lane + 32 * t; r < R; r += (NT ? 1 << 30 : 32)) {
                c = fmax(c, fabs(S[r] * Z[r] - m));
                c = fmax(c, fabs(T[r] * V[r] - m));
            }
        return warp_max(c);
    }
    __device__ double total_err(const Err &e, double m) const {
        const double s_max = 100.0;
        int nb = NB + 4 * N + 2 * R;  // lower + upper bound multipliers + row (z, v)
        int nmul = 6 * N + R + nb;
        double sd = fmax(s_max, (e.ysum + e.zsum) / (double)(nmul > 0 ? nmul : 1)) / s_max;
        double sc = fmax(s_max, e.zsum / (double)(nb > 0 ? nb : 1)) / s_max;
        return fmax(e.dual / sd, fmax(e.prim, comp_err(m) / sc));
    }

    // ---- per-iteration assembly: HD (diag Hessian additions), base gradient (into D), SIGE, YHAT
    __device__ void assemble() {
#pragma unroll
        for (int t = 0; t < (NT ? rounds(M * NT) : 1); t++)
            for (int r = lane + 32 * t; r < R; r += (NT ? 1 << 30 : 32)) {
                double s = S[r], tt = T[r];
                double is = 1.0 / s, it = 1.0 / tt;
                double sig_s = Z[r] * is, sig_t = V[r] * it;
                double beta = sig_t / (sig_s + sig_t);
                double rg = GR[r] + tt - s;
                SIGE[r] = beta * sig_s;
                YHAT[r] = (1.0 - beta) * (rho - mu * it) + beta * (mu * is - sig_s * rg);
            }
#pragma unroll
        for (int t = 0; t < (NT ? rounds(8 * NT + M * (NT + 1)) : 1); t++)
            for (int wi = 6 + lane + 32 * t; wi < NW; wi += (NT ? 1 << 30 : 32)) {
                double hd = 0.0, g = grad_f(wi);
                int bs = bslot(wi);
                if (bs >= 0) {
                    int wj; double lb, ub; bool hu;
                    bvar(bs, wj, lb, ub, hu);
                    double idl = 1.0 / (W[wi] - lb);
                    hd += ZL[bs] * idl;
                    g -= mu * idl;
                    if (hu) {
                        double idu = 1.0 / (ub - W[wi]);
                        hd += ZU[bs] * idu;
                        g += mu * idu;
                    }
                }
                if (M > 0 && wi < OU) {  // Hessian of -y_r g_r: diagonal on (s, ey) (control.py:544-557, degree 6)
                    int i = wi / 6, a = wi - 6 * i;
                    if (a >= 4) {
                        double iX6 = (a == 4) ? kp.iL6 : kp.iW6;
#pragma unroll
                        for (int j = 0; j < M; j++) {
                            if (i < N) {
                                int r = j * N + i;
                                double d = W[wi] - ((a == 4) ? (obs_s(j, i) + lapoff[j]) : obs_e(j, i));
                                hd += Y[r] * DG[r] * a1 * 30.0 * p4(d) * iX6;
                            }
                            {
                                int r = j * N + i - 1;
                                double d = W[wi] - ((a == 4) ? obs_s(j, i) : obs_e(j, i));
                                hd -= Y[r] * DG[r] * 30.0 * p4(d) * iX6;
                            }
                        }
                    }
                }
                HD[wi] = hd;
                D[wi] = g;  // base gradient of the barrier problem; D is overwritten by the forward pass later
            }
        __syncwarp();
    }

    // ---- Riccati backward sweep with primal regularisation dw.  Returns false if a pivot <= 0.
    // Stage variables zeta = (dx 6, dsigma M | du 2, dsigma+ M); dX+ = TM * zeta + (rd, 0).
    __device__ bool riccati_backward(double dw) {
        // terminal value function: stage-N state block
#pragma unroll
        for (int t = 0; t < rounds(NXA * NXA); t++) {
            int e = lane + 32 * t;
            if (e < NXA * NXA) {
                int a = e / NXA, b = e - a * NXA;
                bool xx = (a < 6) && (b < 6);
                double v = xx ? QQ[xx ? 6 * a + b : 0] : 0.0;
                if (a == b) v += HD[(a < 6) ? 6 * N + a : isg(a - 6, N)] + dw;
                P[e] = v;
            }
        }
        if (lane < NXA) PV[lane] = (lane < 6) ? D[6 * N + lane] : D[isg(lane - 6, N)];
        __syncwarp();
        bool ok = true;
        for (int k = N - 1; k >= 0; k--) {
            // (1) PT = P * TM (dense, uniform); QV = PV - P[:,0:6]*c_k; row vectors CT; stage diagonal / gradient
#pragma unroll
            for (int t = 0; t < rounds(NXA * NZ); t++) {
                int e = lane + 32 * t;
                if (e < NXA * NZ) {
                    int a = e / NZ, c = e - a * NZ;
                    double v = 0.0;
#pragma unroll
                    for (int b = 0; b < NXA; b++) v += P[a * NXA + b] * TM[b * NZ + c];
                    PT[e] = v;
                }
            }
            if (lane < NXA) {
                double v = PV[lane];
#pragma unroll
                for (int b = 0; b < 6; b++) v -= P[lane * NXA + b] * CRES[6 * k + b];
                QV[lane] = v;
            }
            if (M > 0) {
#pragma unroll
                for (int t = 0; t < rounds(M * NZ); t++) {
                    int e = lane + 32 * t;
                    if (e < M * NZ) {
                        int j = e / NZ, a = e - j * NZ, r = j * N + k;
                        double q3 = JA[4 * r + 2], q4 = JA[4 * r + 3];
                        double v = TM[4 * NZ + a] * q3 + TM[5 * NZ + a] * q4;   // (A'a_n | B'a_n)
                        v += (a == 4) ? JA[4 * r + 0] : 0.0;
                        v += (a == 5) ? JA[4 * r + 1] : 0.0;
                        v += (a == 6 + j) ? DG[r] * a1 : 0.0;
                        v -= (a == NXA + 2 + j) ? DG[r] : 0.0;
                        CT[e] = v;
                    }
                }
            }
            if (lane < NZ) {
                int a = lane;
                bool isx = a < 6, iss = (a >= 6) && (a < NXA), isu = (a >= NXA) && (a < NXA + 2);
                int idx = isx ? 6 * k + a : (iss ? isg(a - 6, k) : (isu ? OU + 2 * k + (a - NXA) : 0));
                bool live = (isx && k > 0) || iss || isu;
                DS[a] = live ? HD[idx] + dw : 0.0;
                GS[a] = live ? D[idx] : 0.0;
            }
            __syncwarp();
            // (2) G = base + TM' P TM + sum_j SIGE_j ct_j ct_j' ; gv likewise
#pragma unroll
            for (int t = 0; t < rounds(NZ * NZ); t++) {
                int e = lane + 32 * t;
                if (e < NZ * NZ) {
                    int a = e / NZ, b = e - a * NZ;
                    double v = 0.0;
#pragma unroll
                    for (int q = 0; q < NXA; q++) v += TM[q * NZ + a] * PT[q * NZ + b];
                    bool xx = (a < 6) && (b < 6);
                    v += xx ? QQ[xx ? 6 * a + b : 0] : 0.0;
                    bool uu = (a >= NXA) && (a < NXA + 2) && (b >= NXA) && (b < NXA + 2);
                    v += uu ? RR[uu ? 2 * (a - NXA) + (b - NXA) : 0] : 0.0;
                    v += (a == b) ? DS[a] : 0.0;
#pragma unroll
                    for (int j = 0; j < M; j++) v += SIGE[j * N + k] * CT[j * NZ + a] * CT[j * NZ + b];
                    GM[e] = v;
                }
            }
            if (lane < NZ) {
                int a = lane;
                double v = GS[a];
#pragma unroll
                for (int q = 0; q < NXA; q++) v += TM[q * NZ + a] * QV[q];
#pragma unroll
                for (int j = 0; j < M; j++) {
                    int r = j * N + k;
                    double er = -(JA[4 * r + 2] * CRES[6 * k + 4] + JA[4 * r + 3] * CRES[6 * k + 5]);
                    v += (SIGE[r] * er - YHAT[r]) * CT[j * NZ + a];
                }
                GV[a] = v;
            }
            __syncwarp();
            // (3) Cholesky of G_uu, redundantly in every lane's registers (rsqrt: no division)
            double Lm[NUA][NUA], rinv[NUA];
#pragma unroll
            for (int a = 0; a < NUA; a++)
#pragma unroll
                for (int b = 0; b <= a; b++) Lm[a][b] = GM[(NXA + a) * NZ + NXA + b];
#pragma unroll
            for (int j = 0; j < NUA; j++) {
                double d = Lm[j][j];
#pragma unroll
                for (int q = 0; q < j; q++) d -= Lm[j][q] * Lm[j][q];
                if (!(d > 0.0)) ok = false;
                double ri = rsqrt(d);
                rinv[j] = ri;
#pragma unroll
                for (int i = j + 1; i < NUA; i++) {
                    double s = Lm[i][j];
#pragma unroll
                    for (int q = 0; q < j; q++) s -= Lm[i][q] * Lm[j][q];
                    Lm[i][j] = s * ri;
                }
            }
            if (!ok) return false;
            // (4) lane c solves column c of  L Y = [G_ux | g_u],  K = -L^-T Y
            if (lane <= NXA) {
                int c = lane;
                double yv[NUA], kv[NUA];
#pragma unroll
                for (int a = 0; a < NUA; a++) {
                    double s = (c < NXA) ? GM[(NXA + a) * NZ + c] : GV[NXA + a];
#pragma unroll
                    for (int q = 0; q < a; q++) s -= Lm[a][q] * yv[q];
                    yv[a] = s * rinv[a];
                }
#pragma unroll
                for (int a = NUA - 1; a >= 0; a--) {
                    double s = yv[a];
#pragma unroll
                    for (int q = a + 1; q < NUA; q++) s -= Lm[q][a] * kv[q];
                    kv[a] = s * rinv[a];
                }
#pragma unroll
                for (int a = 0; a < NUA; a++) {
                    YF[a * (NXA + 1) + c] = yv[a];
                    if (c < NXA) KFB[(k * NUA + a) * NXA + c] = -kv[a];
                    else KFF[k * NUA + a] = -kv[a];
                }
            }
            __syncwarp();
            // (5) P = G_xx - Y'Y,  PV = g_x - Y' y_g
#pragma unroll
            for (int t = 0; t < rounds(NXA * NXA); t++) {
                int e = lane + 32 * t;
                if (e < NXA * NXA) {
                    int a = e / NXA, b = e - a * NXA;
                    double v = GM[a * NZ + b];
#pragma unroll
                    for (int q = 0; q < NUA; q++) v -= YF[q * (NXA + 1) + a] * YF[q * (NXA + 1) + b];
                    P[e] = v;
                }
            }
            if (lane < NXA) {
                double v = GV[lane];
#pragma unroll
                for (int q = 0; q < NUA; q++) v -= YF[q * (NXA + 1) + lane] * YF[q * (NXA + 1) + NXA];
                PV[lane] = v;
            }
            __syncwarp();
        }
        // stage 0: x_0 is fixed (control.py:497), sigma_{.,0} is free: d sigma_0 = -P_ss^-1 p_s
        if (M > 0) {
            constexpr int MM = M > 0 ? M : 1;
            double Ls[MM][MM], ri[MM], yv[MM], xv[MM];
#pragma unroll
            for (int a = 0; a < M; a++)
#pragma unroll
                for (int b = 0; b <= a; b++) Ls[a][b] = P[(6 + a) * NXA + 6 + b];
#pragma unroll
            for (int j = 0; j < M; j++) {
                double d = Ls[j][j];
#pragma unroll
                for (int q = 0; q < j; q++) d -= Ls[j][q] * Ls[j][q];
                if (!(d > 0.0)) ok = false;
                ri[j] = rsqrt(d);
#pragma unroll
                for (int i = j + 1; i < M; i++) {
                    double s = Ls[i][j];
#pragma unroll
                    for (int q = 0; q < j; q++) s -= Ls[i][q] * Ls[j][q];
                    Ls[i][j] = s * ri[j];
                }
            }
            if (!ok) return false;
#pragma unroll
            for (int a = 0; a < M; a++) {
                double s = -PV[6 + a];
#pragma unroll
                for (int q = 0; q < a; q++) s -= Ls[a][q] * yv[q];
                yv[a] = s * ri[a];
            }
#pragma unroll
            for (int a = M - 1; a >= 0; a--) {
                double s = yv[a];
#pragma unroll
                for (int q = a + 1; q < M; q++) s -= Ls[q][a] * xv[q];
                xv[a] = s * ri[a];
            }
            __syncwarp();
            // QV[6+j] carries d sigma_0 to the forward pass (D still holds the base gradient until then)
            if (lane == 0) {
#pragma unroll
                for (int a = 0; a < M; a++) QV[6 + a] = xv[a];
            }
            __syncwarp();
        }
        return true;
    }

    // ---- forward sweep: D <- Newton direction.  Every lane carries dX_k in registers and computes the whole
    //      stage redundantly from broadcast loads (no warp barrier inside the loop); lane a writes component a.
    __device__ void riccati_forward() {
        double dx[NXA];
#pragma unroll
        for (int a = 0; a < 6; a++) dx[a] = 0.0;
#pragma unroll
        for (int j = 0; j < M; j++) dx[6 + j] = QV[6 + j];
        __syncwarp();
        if (lane < 6) D[lane] = 0.0;
        if (M > 0 && lane < M) D[isg(lane, 0)] = QV[6 + lane];
        for (int k = 0; k < N; k++) {
            double du[NUA];
#pragma unroll
            for (int m = 0; m < NUA; m++) {
                double s = KFF[k * NUA + m];
                const double *Kr = KFB + (k * NUA + m) * NXA;
#pragma unroll
                for (int c = 0; c < NXA; c++) s += Kr[c] * dx[c];
                du[m] = s;
            }
            double dn[6];
#pragma unroll
            for (int a = 0; a < 6; a++) {
                double s = -CRES[6 * k + a];
#pragma unroll
                for (int b = 0; b < 6; b++) s += Am(a, b) * dx[b];
                s += Bm(a, 0) * du[0] + Bm(a, 1) * du[1];
                dn[a] = s;
            }
            // publish: lane m < NUA writes dU component m, lane 8+a writes dx_{k+1}[a]
#pragma unroll
            for (int m = 0; m < NUA; m++)
                if (lane == m) D[(m < 2) ? OU + 2 * k + m : isg(m - 2, k + 1)] = du[m];
#pragma unroll
            for (int a = 0; a < 6; a++)
                if (lane == 8 + a) D[6 * (k + 1) + a] = dn[a];
#pragma unroll
            for (int a = 0; a < 6; a++) dx[a] = dn[a];
#pragma unroll
            for (int j = 0; j < M; j++) dx[6 + j] = du[2 + j];
        }
        __syncwarp();
    }
};

// ---------------------------------------------------------------- kernel
template <int NT, int M>
__global__ void __launch_bounds__(32) ocp_ipm_kernel(const __grid_constant__ KParams kp, const double *__restrict__ in,
                                                     b200mpc_record *__restrict__ rec, double *__restrict__ aux,
                                                     double *__restrict__ xpred, double *__restrict__ upred,
                                                     double *__restrict__ sigma) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x;
    const int inst = blockIdx.x;
    const SmemPlan<M> pl(NT ? NT : kp.p.N, kp.in_stride);
    Ipm<NT, M> S_(kp, pl, sm, lane);
    Ipm<NT, M> &q = S_;
    using IP = Ipm<NT, M>;
    constexpr int NXA = IP::NXA, NZ = IP::NZ;
    const int N = q.N, R = q.R, NB = q.NB, NW = q.NW, OU = q.OU, OS = q.OS;
    const b200mpc_ipm_options &o = kp.o;

    // ---- stage the instance record with one TMA bulk copy; build the shared model meanwhile
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
    const uint32_t in_bytes = (uint32_t)kp.in_stride * 8u;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(bar, in_bytes);
        bulk_g2s(q.IN, in + (size_t)inst * kp.in_stride, in_bytes, bar);
    }
    // TM: dX+ = TM * zeta :  rows 0..5 = [A | 0 | B | 0], rows 6+j = unit vector on dsigma+_j
    for (int e = lane; e < NXA * NZ; e += 32) {
        int a = e / NZ, c = e - a * NZ;
        double v = 0.0;
        if (a < 6) {
            if (c < 6) v = kp.p.A[6 * a + c];
            else if (c >= NXA && c < NXA + 2) v = kp.p.B[2 * a + (c - NXA)];
        } else if (c == NXA + 2 + (a - 6))
            v = 1.0;
        q.TM[e] = v;
    }
    for (int e = lane; e < 36; e += 32) q.cQ[e] = kp.p.Q[e];
    if (lane < 4) q.cR[lane] = kp.p.R[lane];
    mbar_wait(bar, 0);
    __syncwarp();

    // ---- start point: u = 0 roll-out from x_0, sigma = 0, pushed into the bounds
    for (int e = lane; e < NW; e += 32) q.W[e] = 0.0;
    __syncwarp();
    if (lane < 6) q.W[lane] = q.IN[lane];
    __syncwarp();
    for (int i = 1; i <= N; i++) {
        if (lane < 6) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 6; b++) s += q.Am(lane, b) * q.W[6 * (i - 1) + b];
            q.W[6 * i + lane] = s;
        }
        __syncwarp();
    }
    for (int b = lane; b < NB; b += 32) {
        int wi; double lb, ub; bool hu;
        q.bvar(b, wi, lb, ub, hu);
        double w = q.W[wi];
        double pl_ = o.bound_push * fmax(1.0, fabs(lb));
        if (hu) pl_ = fmin(pl_, o.bound_frac * (ub - lb));
        if (w < lb + pl_) w = lb + pl_;
        if (hu) {
            double pu = fmin(o.bound_push * fmax(1.0, fabs(ub)), o.bound_frac * (ub - lb));
            if (w > ub - pu) w = ub - pu;
        }
        q.W[wi] = w;
        q.ZL[b] = 1.0;
        if (hu) q.ZU[b] = 1.0;
    }
    for (int e = lane; e < 6 * N; e += 32) q.LAM[e] = 0.0;
    // unscaled (Q+Q'), (R+R') first: the scaling factor is derived from them
    for (int e = lane; e < 36; e += 32) { int a = e / 6, b = e - 6 * a; q.QQ[e] = kp.p.Q[6 * a + b] + kp.p.Q[6 * b + a]; }
    if (lane < 4) { int a = lane >> 1, b = lane & 1; q.RR[lane] = kp.p.R[2 * a + b] + kp.p.R[2 * b + a]; }
    __syncwarp();
    // ---- gradient-based scaling at the start (nlp_scaling_max_gradient)
    {
        double gm = 0.0;
        for (int wi = 6 + lane; wi < NW; wi += 32) gm = fmax(gm, fabs(q.grad_f(wi)));  // df == 1 here
        gm = warp_max(gm);
        q.df = gm > o.max_grad ? o.max_grad / gm : 1.0;
        __syncwarp();
        for (int e = lane; e < 36; e += 32) q.QQ[e] *= q.df;
        if (lane < 4) q.RR[lane] *= q.df;
        for (int r = lane; r < R; r += 32) {
            int j = r / N, i = r - j * N;
            typename IP::RowV v = q.template row_vals<false>(j, i, 0.0);
            double rm = fmax(q.a1, 1.0);  // |d/dsigma_i| = (1-alpha), |d/dsigma_{i+1}| = 1
            rm = fmax(rm, fmax(fabs(6.0 * p5(v.dsn) * kp.iL6), fabs(6.0 * p5(v.den) * kp.iW6)));
            if (i > 0) rm = fmax(rm, q.a1 * fmax(fabs(6.0 * p5(v.ds) * kp.iL6), fabs(6.0 * p5(v.de) * kp.iW6)));
            q.DG[r] = rm > o.max_grad ? o.max_grad / rm : 1.0;
        }
        __syncwarp();
        for (int r = lane; r < R; r += 32) {
            int j = r / N, i = r - j * N;
            double g = q.DG[r] * q.row_g(q.template row_vals<false>(j, i, 0.0));
            double t = fmax(0.0, -g) + o.bound_push;
            q.T[r] = t;
            q.S[r] = g + t;
            q.Z[r] = 1.0;
            q.V[r] = 1.0;
            q.Y[r] = 0.0;
            q.JD[r] = 0.0;
            q.SIGE[r] = 0.0;
            q.YHAT[r] = 0.0;
            q.GR[r] = g;
        }
        __syncwarp();
    }
    double th0, ph_dummy;
    q.template theta_phi<false>(0.0, th0, ph_dummy);
    const double theta_max = 1e4 * fmax(1.0, th0), theta_min = 1e-4 * fmax(1.0, th0);

    // filter: entry f lives in lane f%32, slot f/32 (capacity 64; the oldest entry is overwritten)
    double f_th0 = 0.0, f_ph0 = 0.0, f_th1 = 0.0, f_ph1 = 0.0;
    int nfilt = 0, fpos = 0;
    int iter = 0, status = B200MPC_MAX_ITER, n_acc = 0, n_refac = 0, n_back = 0, n_reset = 0;
    double dw_last = 0.0, E0 = 0.0;

    const double kappa_eps = 10.0, kappa_mu = 0.2, tau_min = 0.99;
    const double gamma_theta = 1e-5, gamma_phi = 1e-8, delta_sw = 1.0, s_theta = 1.1, s_phi = 2.3, eta_phi = 1e-8;
    const double gamma_alpha = 0.05, kappa_sigma = 1e10;

#ifdef B200MPC_PHASE_CLOCKS
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tc0 = clock64();
#define PCLK(k) { long long t_ = clock64(); pc[k] += t_ - tc0; tc0 = t_; }
#else
#define PCLK(k)
#endif
    for (;;) {
        PCLK(7)
        q.eval_point();
        typename IP::Err eb = q.error_base();
        E0 = q.total_err(eb, 0.0);
        if (E0 <= o.tol) { status = B200MPC_SOLVED; break; }
        if (E0 <= o.acceptable_tol) {
            if (++n_acc >= o.acceptable_iter) { status = B200MPC_SOLVED; break; }
        } else
            n_acc = 0;
        if (iter >= o.max_iter) { status = B200MPC_MAX_ITER; break; }
        // ---- barrier parameter (monotone Fiacco-McCormick)
        for (;;) {
            double em = q.total_err(eb, q.mu);
            if (em <= kappa_eps * q.mu && q.mu > o.tol / 11.0) {
                q.mu = fmax(o.tol / 11.0, fmin(kappa_mu * q.mu, q.mu * sqrt(q.mu)));
                nfilt = 0;
                fpos = 0;
            } else
                break;
        }
        const double mu = q.mu, rho = q.rho;
        const double tau = fmax(tau_min, 1.0 - mu);
        PCLK(0)
        // ---- Newton step by Riccati, with inertia correction
        q.assemble();
        PCLK(1)
        double dw_try = 0.0;
        bool fail = false;
        for (;;) {
            if (q.riccati_backward(dw_try)) break;
            n_refac++;
            if (dw_try == 0.0) dw_try = (dw_last == 0.0) ? 1e-4 : fmax(1e-20, dw_last / 3.0);
            else dw_try *= (dw_last == 0.0) ? 100.0 : 8.0;
            if (dw_try > 1e40) { fail = true; break; }
        }
        if (fail) { status = B200MPC_INERTIA; break; }
        if (dw_try > 0.0) dw_last = dw_try;
        PCLK(2)
        // ---- grad(phi)'d needs the base gradient that still sits in D: keep this lane's entries in registers
        //      across the forward sweep that overwrites D (entries wi = 6+lane+32t).
        constexpr int GT = NT ? rounds(8 * NT + M * (NT + 1)) : 8;
        double gbase[GT];
#pragma unroll
        for (int t = 0; t < GT; t++) {
            int wi = 6 + lane + 32 * t;
            gbase[t] = (wi < NW) ? q.D[wi] : 0.0;
        }
        __syncwarp();
        q.riccati_forward();
        PCLK(3)
        // ---- rows: J d, step bounds, directional derivative
#pragma unroll
        for (int t = 0; t < (NT ? rounds(M * NT) : 1); t++)
            for (int r = lane + 32 * t; r < R; r += (NT ? 1 << 30 : 32)) {
                int j = r / N, i = r - j * N;
                double jd = q.JA[4 * r + 0] * q.D[6 * i + 4] + q.JA[4 * r + 1] * q.D[6 * i + 5] + q.DG[r] * q.a1 * q.D[q.isg(j, i)] +
                            q.JA[4 * r + 2] * q.D[6 * i + 10] + q.JA[4 * r + 3] * q.D[6 * i + 11] - q.DG[r] * q.D[q.isg(j, i + 1)];
                q.JD[r] = jd;
            }
        __syncwarp();
        double a_max = 1.0, a_z = 1.0, gphi = 0.0, th = 0.0;
#pragma unroll
        for (int t = 0; t < (NT ? rounds(4 * NT + M * (NT + 1)) : 1); t++)
            for (int b = lane + 32 * t; b < NB; b += (NT ? 1 << 30 : 32)) {
                int wi; double lb, ub; bool hu;
                q.bvar(b, wi, lb, ub, hu);
                double w = q.W[wi], d = q.D[wi];
                double dl = w - lb, zl = q.ZL[b];
                double dzl = mu / dl - zl - zl / dl * d;
                if (d < 0.0) a_max = fmin(a_max, -tau * dl / d);
                if (dzl < 0.0) a_z = fmin(a_z, -tau * zl / dzl);
                if (hu) {
                    double du = ub - w, zu = q.ZU[b];
                    double dzu = mu / du - zu + zu / du * d;
                    if (d > 0.0) a_max = fmin(a_max, tau * du / d);
                    if (dzu < 0.0) a_z = fmin(a_z, -tau * zu / dzu);
                }
            }
#pragma unroll
        for (int t = 0; t < (NT ? rounds(M * NT) : 1); t++)
            for (int r = lane + 32 * t; r < R; r += (NT ? 1 << 30 : 32)) {
                double ds, dt, dy, dz, dv;
                q.row_step(r, ds, dt, dy, dz, dv);
                double s = q.S[r], tt = q.T[r];
                if (ds < 0.0) a_max = fmin(a_max, -tau * s / ds);
                if (dt < 0.0) a_max = fmin(a_max, -tau * tt / dt);
                if (dz < 0.0) a_z = fmin(a_z, -tau * q.Z[r] / dz);
                if (dv < 0.0) a_z = fmin(a_z, -tau * q.V[r] / dv);
                gphi += rho * dt - mu * (ds / s + dt / tt);
                th += fabs(q.GR[r] + tt - s);
            }
#pragma unroll
        for (int t = 0; t < GT; t++) {
            int wi = 6 + lane + 32 * t;
            if (wi < NW) gphi += gbase[t] * q.D[wi];
        }
        if (NT == 0) {
            for (int wi = 6 + 32 * GT + lane; wi < NW; wi += 32) {  // horizons beyond the register window: recompute
                double g = q.grad_f(wi);
                int bs = q.bslot(wi);
                if (bs >= 0) {
                    int wj; double lb, ub; bool hu;
                    q.bvar(bs, wj, lb, ub, hu);
                    g -= mu / (q.W[wi] - lb);
                    if (hu) g += mu / (ub - q.W[wi]);
                }
                gphi += g * q.D[wi];
            }
        }
        for (int e = lane; e < 6 * N; e += 32) th += fabs(q.CRES[e]);
        a_max = warp_min(a_max);
        a_z = warp_min(a_z);
        gphi = warp_sum(gphi);
        th = warp_sum(th);
        double th_chk, ph;
        q.template theta_phi<false>(0.0, th_chk, ph);
        PCLK(4)
        // ---- filter line search
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
            double tht, pht;
            q.template theta_phi<true>(a, tht, pht);
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
        if (!accepted) {
            // IPOPT would call its restoration phase; the rows are elastic, so remove their residual by
            // enlarging the slacks (t' = max(t, s-g), s' = g+t') and restart the filter.
            if (n_reset >= 5) { status = B200MPC_LINESEARCH; break; }
            n_reset++;
            for (int r = lane; r < R; r += 32) {
                double g = q.GR[r];
                double tn = fmax(q.T[r], q.S[r] - g);
                q.T[r] = tn;
                q.S[r] = g + tn;
            }
            nfilt = 0;
            fpos = 0;
            iter++;
            __syncwarp();
            continue;
        }
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
        PCLK(5)
        // ---- accept: multipliers of the dynamics by the costate recursion, from
        //      K d + Jc' lam+ = rhs  =>  lam+_i = (rhs - K d)_{x_i} + A' lam+_{i+1}
        // (a) residuals res_i = (rhs - K d)_{x_i} for all stages in parallel -> CRES (dead until the next eval)
#pragma unroll
        for (int t = 0; t < (NT ? rounds(6 * NT) : 1); t++)
            for (int e = lane + 32 * t; e < 6 * N; e += (NT ? 1 << 30 : 32)) {
                int i = 1 + e / 6, a = e - 6 * (i - 1);
                int wi = 6 * i + a;
                double g = q.grad_f(wi);
                int bs = q.bslot(wi);
                if (bs >= 0) {
                    int wj; double lb, ub; bool hu;
                    q.bvar(bs, wj, lb, ub, hu);
                    g += -mu / (q.W[wi] - lb) + mu / (ub - q.W[wi]);
                }
                double kd = (q.HD[wi] + dw_try) * q.D[wi];
#pragma unroll
                for (int b = 0; b < 6; b++) kd += q.QQ[6 * a + b] * q.D[6 * i + b];
                double res = -g - kd;
                if (M > 0 && a >= 4) {
#pragma unroll
                    for (int j = 0; j < M; j++) {
                        if (i < N) {
                            int r = j * N + i;
                            res += q.JA[4 * r + (a - 4)] * (q.YHAT[r] - q.SIGE[r] * q.JD[r]);
                        }
                        int r = j * N + i - 1;
                        res += q.JA[4 * r + 2 + (a - 4)] * (q.YHAT[r] - q.SIGE[r] * q.JD[r]);
                    }
                }
                q.CRES[e] = res;
            }
        __syncwarp();
        // (b) the recursion itself, redundantly in every lane (no barrier); lane a<6 blends component a into LAM
        {
            double ln[6];
#pragma unroll
            for (int a2 = 0; a2 < 6; a2++) ln[a2] = 0.0;
            for (int i = N; i >= 1; i--) {
                double lc[6];
#pragma unroll
                for (int a2 = 0; a2 < 6; a2++) {
                    double s = q.CRES[6 * (i - 1) + a2];
#pragma unroll
                    for (int b = 0; b < 6; b++) s += q.Am(b, a2) * ln[b];
                    lc[a2] = s;
                }
#pragma unroll
                for (int a2 = 0; a2 < 6; a2++) {
                    if (lane == a2) {
                        double lo = q.LAM[6 * (i - 1) + a2];
                        q.LAM[6 * (i - 1) + a2] = lo + a * (lc[a2] - lo);
                    }
                    ln[a2] = lc[a2];
                }
            }
        }
        // bound multipliers (old point), then primal step, then kappa_sigma safeguard (new point)
#pragma unroll
        for (int t = 0; t < (NT ? rounds(4 * NT + M * (NT + 1)) : 1); t++)
            for (int b = lane + 32 * t; b < NB; b += (NT ? 1 << 30 : 32)) {
                int wi; double lb, ub; bool hu;
                q.bvar(b, wi, lb, ub, hu);
                double w = q.W[wi], d = q.D[wi];
                double dl = w - lb, zl = q.ZL[b];
                zl += a_z * (mu / dl - zl - zl / dl * d);
                double wn = w + a * d, dln = wn - lb;
                q.ZL[b] = fmax(fmin(zl, kappa_sigma * mu / dln), mu / (kappa_sigma * dln));
                if (hu) {
                    double du = ub - w, zu = q.ZU[b];
                    zu += a_z * (mu / du - zu + zu / du * d);
                    double dun = ub - wn;
                    q.ZU[b] = fmax(fmin(zu, kappa_sigma * mu / dun), mu / (kappa_sigma * dun));
                }
            }
#pragma unroll
        for (int t = 0; t < (NT ? rounds(M * NT) : 1); t++)
            for (int r = lane + 32 * t; r < R; r += (NT ? 1 << 30 : 32)) {
                double ds, dt, dy, dz, dv;
                q.row_step(r, ds, dt, dy, dz, dv);
                double s = q.S[r] + a * ds, tt = q.T[r] + a * dt;
                q.Y[r] += a * dy;
                double z = q.Z[r] + a_z * dz, v = q.V[r] + a_z * dv;
                q.Z[r] = fmax(fmin(z, kappa_sigma * mu / s), mu / (kappa_sigma * s));
                q.V[r] = fmax(fmin(v, kappa_sigma * mu / tt), mu / (kappa_sigma * tt));
                q.S[r] = s;
                q.T[r] = tt;
            }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < (NT ? rounds(8 * NT + M * (NT + 1)) : 1); t++)
            for (int wi = 6 + lane + 32 * t; wi < NW; wi += (NT ? 1 << 30 : 32)) q.W[wi] += a * q.D[wi];
        __syncwarp();
        iter++;
        PCLK(6)
    }
#ifdef B200MPC_PHASE_CLOCKS
    // profiling build: phase cycle counters replace x_pred (first 8 doubles of the instance's slot)
    if (xpred != nullptr && lane == 0)
        for (int k = 0; k < 8; k++) xpred[(size_t)inst * 6 * (N + 1) + k] = (double)pc[k];
    xpred = nullptr;
#endif

    // ---- results
    double cost = q.objective();
    double tm = 0.0;
    for (int r = lane; r < R; r += 32) tm = fmax(tm, q.T[r]);
    tm = warp_max(tm);
    if (lane == 0) {
        b200mpc_record rc;
        rc.cost = cost;
        rc.u0[0] = q.W[OU];
        rc.u0[1] = q.W[OU + 1];
        rc.status = status;
        rc.iters = iter;
        rec[inst] = rc;
    }
    if (aux != nullptr && lane < 4) {
        double v = (lane == 0) ? E0 : (lane == 1) ? tm : (lane == 2) ? (double)n_refac : (double)n_back;
        aux[(size_t)inst * 4 + lane] = v;
    }
    if (xpred != nullptr)
        for (int e = lane; e < 6 * (N + 1); e += 32) xpred[(size_t)inst * 6 * (N + 1) + e] = q.W[e];
    if (upred != nullptr)
        for (int e = lane; e < 2 * N; e += 32) upred[(size_t)inst * 2 * N + e] = q.W[OU + e];
    if (sigma != nullptr)
        for (int e = lane; e < M * (N + 1); e += 32) sigma[(size_t)inst * M * (N + 1) + e] = q.W[OS + e];
}

}  // namespace b200mpc

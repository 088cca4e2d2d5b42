// b200mpc: LMPC model identification -- LMPCRacingGame.estimate_ABC (car_racing/utils/base.py:585-622), i.e. for every
// horizon stage lmpc_helper.regression_and_linearization (car_racing/control/lmpc_helper.py:26-201):
//   * nearest stored points of each used lap in the scaled l1 norm (:203-238), Epanechnikov weights (:237)
//   * three weighted least-squares fits (vx+, vy+, wz+) through the normal equations (:241-279, :343-366)
//   * analytic linearisation of the Frenet kinematics for (epsi, s, ey) with the track curvature at s (:135-199)
// One warp per (instance, stage).  The stored laps are shared by the whole batch (structure-of-arrays, read through
// L1/L2); results are written in the record layout of lmpc_kernel so the two kernels chain on the device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200mpc.h"
#include "ocp_ipm.cuh"

namespace b200mpc {

struct SysidKParams {
    b200mpc_sysid_params p;
    int32_t B, out_stride, out_offset;
};

// lap data: [lap][field 0..4 = vx, vy, wz, delta, a][lap_stride]
__device__ __forceinline__ double lap_val(const double *__restrict__ laps, int lap_stride, int lap, int field, int t) {
    return __ldg(laps + ((size_t)lap * 5 + field) * lap_stride + t);
}

__global__ void __launch_bounds__(32) sysid_kernel(const __grid_constant__ SysidKParams kp, const double *__restrict__ lin,
                                                   const double *__restrict__ laps, const double *__restrict__ segments,
                                                   double *__restrict__ out, int32_t *__restrict__ idx_out,
                                                   int32_t *__restrict__ status_out) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x;
    const int N = kp.p.N, inst = blockIdx.x / N, stage = blockIdx.x - inst * N;
    const int L = kp.p.num_laps, P = kp.p.max_num_point, LS = kp.p.lap_stride;
    // shared: norms [LS] | sel_t [L*P] (int) | sel_k [L*P] | mom [39 -> 40] | sys [3][5][6] | abc [54]
    double *norm = sm;
    int *sel_t = reinterpret_cast<int *>(norm + ((LS + 1) & ~1));
    double *sel_k = reinterpret_cast<double *>(sel_t + ((L * P + 1) & ~1));
    double *mom = sel_k + L * P;
    double *sys = mom + 40;
    double *abc = sys + 96;
    const double *rec = lin + ((size_t)inst * N + stage) * 8;
    double x0[6], u0[2];
#pragma unroll
    for (int a = 0; a < 6; a++) x0[a] = __ldg(rec + a);
    u0[0] = __ldg(rec + 6);
    u0[1] = __ldg(rec + 7);
    const double z0[5] = {x0[0], x0[1], x0[2], u0[0], u0[1]};
    const double h = kp.p.h;
    int nsel_tot = 0;
    int bad = 0;
    // ---- selection per lap
    for (int lap = 0; lap < L; lap++) {
        const int T1 = kp.p.lap_rows[lap] - 1;   // rows 0 .. time_ss-2 (:219-223)
        int inside = 0;
        for (int t = lane; t < T1; t += 32) {
            double nv = 0.1 * fabs(lap_val(laps, LS, lap, 0, t) - z0[0]);
#pragma unroll
            for (int f = 1; f < 5; f++) nv += fabs(lap_val(laps, LS, lap, f, t) - z0[f]);
            norm[t] = nv;
            inside += (nv < h) ? 1 : 0;
        }
        for (int off = 16; off > 0; off >>= 1) inside += __shfl_xor_sync(0xffffffffu, inside, off);
        __syncwarp();
        int cnt = 0;
        if (inside >= P) {
            // the P smallest norms in ascending order (np.argsort(norm)[0:P]); ties -> lower index
            for (int r = 0; r < P; r++) {
                double best = 1e300;
                int bi = 0x7fffffff;
                for (int t = lane; t < T1; t += 32) {
                    double nv = norm[t];
                    if (nv < best) { best = nv; bi = t; }
                }
                for (int off = 16; off > 0; off >>= 1) {
                    double ob = __shfl_xor_sync(0xffffffffu, best, off);
                    int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if (lane == 0) {
                    sel_t[lap * P + r] = bi;
                    sel_k[lap * P + r] = 0.75 * (1.0 - (best / h) * (best / h));
                    norm[bi] = 2e300;   // taken
                }
                __syncwarp();
            }
            cnt = P;
        } else {
            // all rows inside the bandwidth, in row order (:234-235): ordered compaction by ballot
            for (int base = 0; base < T1; base += 32) {
                int t = base + lane;
                double nv = (t < T1) ? norm[t] : 1e300;
                unsigned m = __ballot_sync(0xffffffffu, nv < h);
                if (nv < h) {
                    int pos = cnt + __popc(m & ((1u << lane) - 1u));
                    sel_t[lap * P + pos] = t;
                    sel_k[lap * P + pos] = 0.75 * (1.0 - (nv / h) * (nv / h));
                }
                cnt += __popc(m);
            }
            __syncwarp();
        }
        for (int r = cnt + lane; r < P; r += 32) { sel_t[lap * P + r] = -1; sel_k[lap * P + r] = 0.0; }
        nsel_tot += cnt;
        __syncwarp();
    }
    if (idx_out != nullptr)
        for (int e = lane; e < L * P; e += 32) idx_out[((size_t)inst * N + stage) * L * P + e] = sel_t[e];
    // ---- weighted moments: entries e < 21: pairs (a <= b) of f = (vx, vy, wz, delta, a, 1); 21 + 3a + c: f_a * y_c
    for (int rd = 0; rd < 2; rd++) {
        int e = lane + 32 * rd;
        if (e < 39) {
            int fa, fb;       // fb >= 6 means target y_{fb-6}
            if (e < 21) {
                int a = 0, rem = e;
                while (rem >= 6 - a) { rem -= 6 - a; a++; }
                fa = a; fb = a + rem;
            } else {
                fa = (e - 21) / 3;
                fb = 6 + (e - 21) - 3 * fa;
            }
            double acc = 0.0;
            for (int lap = 0; lap < L; lap++)
                for (int r = 0; r < P; r++) {
                    int t = sel_t[lap * P + r];
                    if (t < 0) break;
                    double va = (fa == 5) ? 1.0 : lap_val(laps, LS, lap, fa, t);
                    double vb = (fb == 5) ? 1.0 : ((fb < 5) ? lap_val(laps, LS, lap, fb, t) : lap_val(laps, LS, lap, fb - 6, t + 1));
                    acc += sel_k[lap * P + r] * va * vb;
                }
            mom[e] = acc;
        }
    }
    __syncwarp();
    // ---- three 5x5 systems, Gaussian elimination with partial pivoting (lanes 0..2)
    if (lane < 3) {
        const int f4 = (lane == 0) ? 4 : 3;           // input feature: a for vx+, delta for vy+ / wz+ (:82,:103)
        const int feat[5] = {0, 1, 2, f4, 5};
        double *Ms = sys + lane * 32;
        for (int a = 0; a < 5; a++) {
            for (int b = 0; b < 5; b++) {
                int p_ = feat[a] < feat[b] ? feat[a] : feat[b], q_ = feat[a] < feat[b] ? feat[b] : feat[a];
                int e = p_ * 6 - p_ * (p_ - 1) / 2 + (q_ - p_);
                Ms[a * 6 + b] = mom[e];
            }
            Ms[a * 6 + 5] = mom[21 + 3 * feat[a] + lane];
        }
        bool sing = false;
        for (int k = 0; k < 5; k++) {
            int piv = k;
            double best = fabs(Ms[k * 6 + k]);
            for (int r = k + 1; r < 5; r++)
                if (fabs(Ms[r * 6 + k]) > best) { best = fabs(Ms[r * 6 + k]); piv = r; }
            if (!(best > 0.0)) { sing = true; break; }
            if (piv != k)
                for (int c = 0; c < 6; c++) { double t = Ms[k * 6 + c]; Ms[k * 6 + c] = Ms[piv * 6 + c]; Ms[piv * 6 + c] = t; }
            double inv = 1.0 / Ms[k * 6 + k];
            for (int r = k + 1; r < 5; r++) {
                double f = Ms[r * 6 + k] * inv;
                for (int c = k + 1; c < 6; c++) Ms[r * 6 + c] -= f * Ms[k * 6 + c];
            }
        }
        double th[5] = {0, 0, 0, 0, 0};
        if (!sing)
            for (int k = 4; k >= 0; k--) {
                double s = Ms[k * 6 + 5];
                for (int c = k + 1; c < 5; c++) s -= Ms[k * 6 + c] * th[c];
                th[k] = s / Ms[k * 6 + k];
            }
        bad = sing ? 1 : 0;
        // rows 0..2 of A, B, C
        double *Ar = abc + 6 * lane, *Br = abc + 36 + 2 * lane;
        Ar[0] = th[0]; Ar[1] = th[1]; Ar[2] = th[2]; Ar[3] = 0.0; Ar[4] = 0.0; Ar[5] = 0.0;
        Br[0] = (lane == 0) ? 0.0 : th[3];
        Br[1] = (lane == 0) ? th[3] : 0.0;
        abc[48 + lane] = th[4];
    }
    // ---- analytic rows (lane 3)
    if (lane == 3) {
        const double dt = kp.p.dt;
        double vx = x0[0], vy = x0[1], wz = x0[2], epsi = x0[3], s = x0[4], ey = x0[5];
        double sw = s;
        while (sw > kp.p.lap_length) sw -= kp.p.lap_length;
        while (sw < 0.0) sw += kp.p.lap_length;
        double cur = 0.0;
        bool found = false;
        for (int g = 0; g < kp.p.num_segments && !found; g++) {
            double s0 = __ldg(segments + 3 * g), len = __ldg(segments + 3 * g + 1);
            if (sw >= s0 && sw <= s0 + len) { cur = __ldg(segments + 3 * g + 2); found = true; }
        }
        if (!found) bad = 2;
        double den = 1.0 - cur * ey, ce = cos(epsi), se = sin(epsi);
        double *A3 = abc + 18, *A4 = abc + 24, *A5 = abc + 30;
        A3[0] = -dt * ce / den * cur;
        A3[1] = dt * se / den * cur;
        A3[2] = dt;
        A3[3] = 1.0 - dt * (-vx * se - vy * ce) / den * cur;
        A3[4] = 0.0;
        A3[5] = dt * (vx * ce - vy * se) / (den * den) * cur * (-cur);
        A4[0] = dt * (ce / den);
        A4[1] = -dt * (se / den);
        A4[2] = 0.0;
        A4[3] = dt * (-vx * se - vy * ce) / den;
        A4[4] = 1.0;
        A4[5] = -dt * (vx * ce - vy * se) / (den * 2.0) * (-cur);   // `den * 2`: reference quirk (lmpc_helper.py:178)
        A5[0] = dt * se;
        A5[1] = dt * ce;
        A5[2] = 0.0;
        A5[3] = dt * (vx * ce - vy * se);
        A5[4] = 0.0;
        A5[5] = 1.0;
        double d3 = 0.0, d4 = 0.0, d5 = 0.0;
        for (int a = 0; a < 6; a++) { d3 += A3[a] * x0[a]; d4 += A4[a] * x0[a]; d5 += A5[a] * x0[a]; }
        abc[48 + 3] = epsi + dt * (wz - (vx * ce - vy * se) / (1.0 - cur * ey) * cur) - d3;
        abc[48 + 4] = s + dt * ((vx * ce - vy * se) / (1.0 - cur * ey)) - d4;
        abc[48 + 5] = ey + dt * (vx * se + vy * ce) - d5;
        for (int e = 0; e < 6; e++) abc[36 + 6 + e] = 0.0;   // B rows 3..5
    }
    __syncwarp();
    // ---- write A_i (36), B_i (12), C_i (6) into the LMPC record layout
    double *o = out + (size_t)inst * kp.out_stride + kp.out_offset;
    for (int e = lane; e < 36; e += 32) o[36 * stage + e] = abc[e];
    if (lane < 12) o[36 * N + 12 * stage + lane] = abc[36 + lane];
    if (lane < 6) o[48 * N + 6 * stage + lane] = abc[48 + lane];
    for (int off = 16; off > 0; off >>= 1) bad |= __shfl_xor_sync(0xffffffffu, bad, off);
    if (status_out != nullptr && lane == 0) status_out[(size_t)inst * N + stage] = bad | ((nsel_tot < 5) ? 4 : 0);
}

}  // namespace b200mpc

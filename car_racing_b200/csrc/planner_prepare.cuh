// b200mpc: the planner's candidate preparation on the device (last open part).
//
// Reference: everything OvertakeTrajPlanner.get_local_traj does between the rivals' predictions and the candidate
// solves (car_racing/planning/overtake_traj_planner.py:87-117) and the data part of generate_traj_per_region
// (:263-334, 365-374):
//   veh_infos[:, 1] = max ey of a rival's prediction (:87-92, row = insertion order, read with the region index --
//   reference quirk), max_delta_v (planner_helper.get_agent_info, planning/planner_helper.py:177-205), the cubic Bezier
//   control points of every region (planner_helper.py:46-136) and their N+1 samples (planner_helper.py:139-153,
//   overtake_traj_planner.py:105-111), the per-stage targets (s~_j, interp1d(samples)(s~_j)) (:329-334), the per-stage
//   bounds incl. the rival rows (:276-324), the heuristic trajectory of the failure branch (:365-374) and whether
//   x_0 violates a stage-0 row.
// Output = the packed records the candidate solve pulls by TMA (layout: include/b200mpc.h, flags STAGE_BOUNDS|EY_RATE,
// M = 0, per-stage targets), so that prediction -> preparation -> candidate solve -> selection -> tracking solve is one
// stream of kernels with no host packing in between.
//
// One CTA: thread c owns region c (strided).  HBM-bound in principle (reads num_veh x 176 B of predictions + the 1 KB
// optimal-trajectory table, writes C x (1.2 KB record + 528 B heuristic)); in practice a launch-latency kernel.
// Products and sums that the reference evaluates as separate numpy operations use __dmul_rn / __dadd_rn so that the
// compiler cannot contract them into FMAs: the window tests (:296-300) and the clip to the curve's range decide
// discretely on these values.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200mpc.h"

namespace b200mpc {

struct PrepareKParams {
    b200mpc_planner_prepare_params p;
    int32_t stride;     // doubles per candidate record
    int32_t xt_off;     // per-stage targets
    int32_t bnd_off;    // per-stage bounds [lo_vx, lo_ey, hi_vx, hi_ey]
    int32_t wd_off;     // ey-rate weights
};

constexpr int PREPARE_NT = 128;
constexpr double PREPARE_NO_BOUND = 1e300;   // "none" for the solver (ocp_ipm.cuh: |bound| >= 1e299)

// scipy interp1d(kind="linear") on float64 data = numpy.interp inside the range: j with xs[j] <= x < xs[j+1],
// slope * (x - x_lo) + y_lo, knots (and the last point) return their ordinate exactly.  Outside the range interp1d
// raises ValueError (bounds_error default): *err is set and the nearest end value returned.
__device__ __forceinline__ double interp_lin(const double *xs, const double *ys, int stride, int n, double x, int *err) {
    if (x < xs[0]) { *err = 1; return ys[0]; }
    if (x > xs[(size_t)(n - 1) * stride]) { *err = 1; return ys[(size_t)(n - 1) * stride]; }
    int j = 0;
    while (j + 1 < n && xs[(size_t)(j + 1) * stride] <= x) j++;
    if (j >= n - 1 || xs[(size_t)j * stride] == x) return ys[(size_t)j * stride];
    const double x_lo = xs[(size_t)j * stride], x_hi = xs[(size_t)(j + 1) * stride];
    const double y_lo = ys[(size_t)j * stride], y_hi = ys[(size_t)(j + 1) * stride];
    const double slope = __ddiv_rn(__dsub_rn(y_hi, y_lo), __dsub_rn(x_hi, x_lo));
    return __dadd_rn(__dmul_rn(slope, __dsub_rn(x, x_lo)), y_lo);
}

// everything of region c (one thread)
__device__ __forceinline__ void prepare_region(const PrepareKParams &kp, int c, const double *__restrict__ ego,
                                               const double *__restrict__ rivals, const double *__restrict__ rival_vx,
                                               const int32_t *__restrict__ insertion, const double *__restrict__ opt,
                                               double *cand, double *__restrict__ heur, int32_t *__restrict__ ok0,
                                               int32_t *__restrict__ region, double *__restrict__ offset, double *__restrict__ ctrl,
                                               double *bezier, int *err_flag) {
    const b200mpc_planner_prepare_params &p = kp.p;
    const int N = p.N, N1 = N + 1, nv = p.num_veh;
    const double *xe = ego, *xp = ego + 6;   // vehicles["ego"].xcurv ; the xcurv_ego argument of get_local_traj
    // planner_helper.py:177-205: only max_delta_v reaches the control points
    double max_dv = 0.0;
    for (int i = 0; i < nv; i++) max_dv = fmax(max_dv, fabs(__dsub_rn(xe[0], rival_vx[i])));
    const double *os_ = opt, *oe = opt + 1;
    int err = 0;
    // ---- control points (planner_helper.py:46-136)
    double s0 = xp[4];
    double s3 = __dadd_rn(__dadd_rn(xp[4], __dmul_rn(p.prediction_factor, max_dv)), 4.0);
    double s1, s2;
    if (s0 > s3) {                                                        // :63-79
        const double span = __dsub_rn(__dadd_rn(s3, p.lap_length), s0);
        s1 = __dadd_rn(__ddiv_rn(span, 3.0), s0);
        s2 = __dadd_rn(__ddiv_rn(__dmul_rn(2.0, span), 3.0), s0);
        s3 = __dadd_rn(s3, p.lap_length);
    } else {                                                              // :81-90
        const double span = __dsub_rn(s3, s0);
        s1 = __dadd_rn(__ddiv_rn(span, 3.0), s0);
        s2 = __dadd_rn(__ddiv_rn(__dmul_rn(2.0, span), 3.0), s0);
    }
    // ey0: the reference looks the optimal trajectory up (and may raise) although it overwrites the value (:92-100)
    if (s0 < 0.0) (void)interp_lin(os_, oe, 2, p.num_opt, __dadd_rn(s0, p.lap_length), &err);
    else if (!(s0 < os_[0])) (void)interp_lin(os_, oe, 2, p.num_opt, s0, &err);
    const double ey0 = xp[5];
    // veh_infos[i, 1]: row i is the i-th rival of vehicles_interest (insertion order), read with the region index
    auto ey_max = [&](int i) {
        const double *e = rivals + ((size_t)insertion[i] * 2 + 1) * N1;
        double m = e[0];
        for (int j = 1; j < N1; j++) m = fmax(m, e[j]);
        return m;
    };
    const double hw = __dmul_rn(0.5, p.veh_width);
    double ey1;
    if (c == 0 && nv > 0)                                                 // :103-110
        ey1 = __dsub_rn(__dmul_rn(0.8, p.track_width), __dmul_rn(__dsub_rn(-ey_max(0), hw), 0.2));
    else if (c == nv && nv > 0)                                           // :112-118
        ey1 = __dadd_rn(__dmul_rn(-0.8, p.track_width), __dmul_rn(__dsub_rn(ey_max(c - 1), hw), 0.2));
    else if (nv > 0)                                                      // :119-125
        ey1 = __dadd_rn(__dmul_rn(0.7, __dadd_rn(ey_max(c), hw)), __dmul_rn(0.3, __dsub_rn(ey_max(c - 1), hw)));
    else
        ey1 = 0.0;
    double ey3;                                                           // :127-135
    if (s3 >= p.lap_length) {
        const double sl = __dsub_rn(s3, p.lap_length);
        ey3 = (sl <= os_[0]) ? oe[0] : interp_lin(os_, oe, 2, p.num_opt, sl, &err);
    } else
        ey3 = (s3 <= os_[0]) ? oe[0] : interp_lin(os_, oe, 2, p.num_opt, s3, &err);
    if (ctrl != nullptr) {
        double *q = ctrl + (size_t)c * 8;
        q[0] = s0; q[1] = ey0; q[2] = s1; q[3] = ey1; q[4] = s2; q[5] = ey1; q[6] = s3; q[7] = ey3;
    }
    // ---- N+1 samples of the curve (planner_helper.py:139-153 at t = j/N); kept in the candidate's record slot for
    //      the look-ups below: the target block [s~ | ey] is written afterwards, stage by stage, behind the reads
    double *rec = cand + (size_t)c * kp.stride;
    double *bz = (bezier != nullptr) ? bezier + (size_t)c * 2 * N1 : rec + kp.bnd_off;   // scratch: 2 N1 <= 4 N1 doubles
    const double tstep = 1.0 / (double)N;
    for (int j = 0; j < N1; j++) {
        const double t = __dmul_rn((double)j, tstep), u = __dsub_rn(1.0, t);
        const double u2 = __dmul_rn(u, u), u3 = __dmul_rn(u2, u), t2 = __dmul_rn(t, t), t3 = __dmul_rn(t2, t);
        const double bs = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(s0, u3), __dmul_rn(__dmul_rn(__dmul_rn(3.0, s1), t), u2)),
                                               __dmul_rn(__dmul_rn(__dmul_rn(3.0, s2), t2), u)), __dmul_rn(s3, t3));
        const double be = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(ey0, u3), __dmul_rn(__dmul_rn(__dmul_rn(3.0, ey1), t), u2)),
                                               __dmul_rn(__dmul_rn(__dmul_rn(3.0, ey1), t2), u)), __dmul_rn(ey3, t3));
        bz[2 * j] = bs;
        bz[2 * j + 1] = be;
    }
    const double b_lo = bz[0], b_hi = bz[2 * N];
    // ---- record: x0, per-stage targets (overtake_traj_planner.py:329-334; -200 s_N folded into the terminal target,
    //      car_racing_b200/planning.py), heuristic trajectory of the failure branch (:365-374)
    for (int q = 0; q < 6; q++) rec[q] = xe[q];
    double *xt = rec + kp.xt_off;
    double *hz = heur + (size_t)c * 6 * N1;
    double s_ref_N = 0.0;
    for (int j = 0; j < N1; j++) {
        double st = __dadd_rn(xe[4], __dmul_rn(__dmul_rn(__dmul_rn(1.0, (double)j), xe[0]), 0.1));       // :330
        st = fmin(fmax(st, b_lo), b_hi);                                                                   // :331-332
        int e2 = 0;
        const double ey = interp_lin(bz, bz + 1, 2, N1, st, &e2);
        double *t = xt + 6 * j;
        t[0] = 0.0; t[1] = 0.0; t[2] = 0.0; t[3] = 0.0; t[4] = st; t[5] = ey;
        if (j == N) { s_ref_N = st; t[4] = st + p.w_progress / (2.0 * p.w_track); }
        double sh = __dadd_rn(xp[4], __dmul_rn(__dmul_rn(__dmul_rn(1.1, (double)j), 0.1), xp[0]));        // :367
        double *hq = hz + 6 * j;
        hq[0] = __dmul_rn(1.1, xp[0]); hq[1] = 0.0; hq[2] = 0.0; hq[3] = 0.0; hq[4] = sh;                   // :368-369
        sh = fmin(fmax(sh, b_lo), b_hi);                                                                   // :370-373
        hq[5] = interp_lin(bz, bz + 1, 2, N1, sh, &e2);
    }
    if (offset != nullptr)   // reference cost = solver cost - 200 s~_N - 200^2/(4*20) + 200 s_0
        offset[c] = -p.w_progress * s_ref_N - p.w_progress * p.w_progress / (4.0 * p.w_track) + p.w_progress * xe[4];
    // ---- per-stage bounds (:276-324); written last, the scratch samples lived here
    const double half = __dsub_rn(p.track_width, hw);
    double *bd = rec + kp.bnd_off;
    for (int k = 0; k < N1; k++) {
        double lo_ey = (k < N) ? -half : -PREPARE_NO_BOUND, hi_ey = (k < N) ? half : PREPARE_NO_BOUND;
        if (k < N) {
            const double s_pred = __dadd_rn(xp[4], __dmul_rn(__dmul_rn((double)k, 0.1), xp[0]));           // :295 / :315
            for (int side = c - 1; side <= c; side++) {                                                    // left, then right rival
                if (side < 0 || side >= nv) continue;
                double so = rivals[(size_t)side * 2 * N1 + k];
                while (so > p.lap_length) so = __dsub_rn(so, p.lap_length);                                // :291-292
                if (s_pred >= __dsub_rn(__dsub_rn(so, p.veh_length), p.safety_margin) &&
                    s_pred <= __dadd_rn(__dadd_rn(so, p.veh_length), p.safety_margin))                     // :296-300
                    lo_ey = fmax(lo_ey, __dadd_rn(__dadd_rn(rivals[((size_t)side * 2 + 1) * N1 + k], p.veh_width), p.safety_margin));
            }
        }
        bd[4 * k] = -PREPARE_NO_BOUND;
        bd[4 * k + 1] = lo_ey;
        bd[4 * k + 2] = (k >= 1) ? p.vx_max : PREPARE_NO_BOUND;                                            // vx_{k+1} <= 5 (:276)
        bd[4 * k + 3] = hi_ey;
        if (k == 0) ok0[c] = (lo_ey <= xe[5] && xe[5] <= hi_ey) ? 1 : 0;
    }
    double *wd = rec + kp.wd_off;                                                                         // :325-327
    for (int k = 0; k < ((N1) & ~1); k++) wd[k] = (k >= 1 && k <= N - 2) ? p.w_ey_rate : 0.0;
    region[c] = c;
    if (err) *err_flag = 1;
}

__global__ void __launch_bounds__(PREPARE_NT) planner_prepare_kernel(const __grid_constant__ PrepareKParams kp,
                                                                      const double *__restrict__ ego,      // ego.xcurv 6, xcurv_ego 6
                                                                      const double *__restrict__ rivals,   // num_veh x 2 x (N+1), sorted
                                                                      const double *__restrict__ rival_vx,
                                                                      const int32_t *__restrict__ insertion,
                                                                      const double *__restrict__ opt,      // num_opt x 2 (s, ey)
                                                                      double *cand, double *__restrict__ heur,
                                                                      int32_t *__restrict__ ok0, int32_t *__restrict__ region,
                                                                      double *__restrict__ offset, double *__restrict__ ctrl,
                                                                      double *bezier, int32_t *__restrict__ err_out) {
    const int C = kp.p.num_veh + 1, tid = threadIdx.x;
    __shared__ int s_err;
    if (tid == 0) s_err = 0;
    __syncthreads();
    int err = 0;
    for (int c = tid; c < C; c += PREPARE_NT)
        prepare_region(kp, c, ego, rivals, rival_vx, insertion, opt, cand, heur, ok0, region, offset, ctrl, bezier, &err);
    if (err) atomicOr(&s_err, 1);
    __syncthreads();
    if (tid == 0 && err_out != nullptr) *err_out = s_err;
}

}  // namespace b200mpc

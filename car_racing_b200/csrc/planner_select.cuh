// b200mpc: the planner's selection step and the hand-over to the tracking MPC, on the device.
//
// Reference: OvertakeTrajPlanner.solve_optimization_problem, the part after the candidates are gathered
// (car_racing/planning/overtake_traj_planner.py:205-246) -- selection cost
//     -10 (s_N - s_0) + 100 #{steps inside a neighbouring rival's L^2+W^2 disc} + 100 [index != old_direction_flag],
// direction_flag = first argmin -- and the first lines of control.mpc_multi_agents (control/control.py:277, 373-382),
// which turn the chosen trajectory into the per-stage targets x_t(i) = [vx, 0, 0, 0, 0, f_traj(s_i)],
// f_traj = scipy interp1d(traj s, traj ey) (linear = numpy.interp), s_i = clip(vx 0.1 i + s_0, traj s range).
//
// One CTA: thread c owns candidate c (strided), a shared-memory first-min argmin picks the region, then the threads
// copy the chosen trajectory and fill the target block of the tracking MPC's packed record in place, so the
// candidate solve -> selection -> tracking solve chain runs on one stream without a host round trip.
// HBM-bound in principle (C x 528 B in, 0.6 KB out); in practice a launch-latency kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200mpc.h"

namespace b200mpc {

struct SelectKParams {
    b200mpc_planner_select_params p;
    int32_t track_xt_off;   // offset of the per-stage target block inside the tracking record (doubles)
};

constexpr int SELECT_NT = 128;

__global__ void __launch_bounds__(SELECT_NT) planner_select_kernel(const __grid_constant__ SelectKParams kp,
                                                                    const b200mpc_record *__restrict__ rec,
                                                                    const double *__restrict__ xpred,
                                                                    const double *__restrict__ heur,
                                                                    const int32_t *__restrict__ ok0,
                                                                    const int32_t *__restrict__ region,
                                                                    const double *__restrict__ rivals,
                                                                    double *__restrict__ sel_cost, int32_t *__restrict__ flag,
                                                                    double *__restrict__ traj, double *__restrict__ track_rec) {
    const b200mpc_planner_select_params &p = kp.p;
    const int C = p.C, N1 = p.N + 1, tid = threadIdx.x;
    __shared__ double s_cost[SELECT_NT];
    __shared__ int s_idx[SELECT_NT];
    __shared__ int s_best;
    const double r2 = p.veh_length * p.veh_length + p.veh_width * p.veh_width;
    double best = 0.0;
    int besti = -1;
    for (int c = tid; c < C; c += SELECT_NT) {
        // the candidate's trajectory: its QP solution, or the reference's heuristic fallback when IPOPT "failed"
        // (overtake_traj_planner.py:365-374) -- here: x_0 violates a stage-0 row, or the solve did not converge
        const bool solved = ok0[c] != 0 && rec[c].status == 0;
        const double *x = (solved ? xpred : heur) + (size_t)c * 6 * N1;
        double cost = -10.0 * (x[6 * (N1 - 1) + 4] - x[4]);                       // :207
        const int reg = region[c];
        for (int side = reg - 1; side <= reg; side++) {                           // :208-237, left then right neighbour
            if (side < 0 || side >= p.num_veh) continue;
            const double *os = rivals + (size_t)side * 2 * N1, *oe = os + N1;
            for (int j = 0; j < N1; j++) {
                double so = os[j];
                while (so > p.lap_length) so -= p.lap_length;                      // :214-215 / :229-230
                const double ds = x[6 * j + 4] - so, de = x[6 * j + 5] - oe[j];
                if (ds * ds + de * de - r2 < 0.0) cost += 100.0;                   // :218-222
            }
        }
        if (p.old_direction_flag >= 0 && p.old_direction_flag != c) cost += 100.0;   // :238-243
        if (sel_cost != nullptr) sel_cost[c] = cost;
        if (besti < 0 || cost < best) { best = cost; besti = c; }                 // strided: lower c first within a thread
    }
    s_cost[tid] = best;
    s_idx[tid] = besti;
    __syncthreads();
    if (tid == 0) {   // list.index(min(list)): the lowest index among equal costs (:244)
        double b = 0.0;
        int bi = -1;
        for (int t = 0; t < SELECT_NT; t++) {
            const int i = s_idx[t];
            if (i < 0) continue;
            if (bi < 0 || s_cost[t] < b || (s_cost[t] == b && i < bi)) { b = s_cost[t]; bi = i; }
        }
        s_best = bi;
        flag[0] = bi;
        flag[1] = bi >= 0 ? region[bi] : -1;
    }
    __syncthreads();
    const int bi = s_best;
    if (bi < 0) return;
    const bool solved = ok0[bi] != 0 && rec[bi].status == 0;
    const double *x = (solved ? xpred : heur) + (size_t)bi * 6 * N1;
    if (traj != nullptr)
        for (int e = tid; e < 6 * N1; e += SELECT_NT) traj[e] = x[e];              // traj_xcurv (N+1, 6) (:245)
    if (track_rec != nullptr) {
        // control.py:277 + :373-382: x_t(i) = [vx, 0, 0, 0, 0, f_traj(s_i)] for the tracking MPC's stages
        const double vx = track_rec[0], s0 = track_rec[4];
        const double s_lo = x[4], s_hi = x[6 * (N1 - 1) + 4];
        for (int i = tid; i <= p.N_ctrl; i += SELECT_NT) {
            double st = vx * 0.1 * (double)i + s0;
            st = fmax(st, s_lo);
            if (st >= s_hi) st = s_hi;
            // scipy interp1d(kind="linear") on float64 data delegates to numpy.interp: j with xs[j] <= st < xs[j+1];
            // knots (and the last point) return their ordinate exactly
            int j = 0;
            while (j + 1 < N1 && x[6 * (j + 1) + 4] <= st) j++;
            double ey;
            if (j >= N1 - 1 || x[6 * j + 4] == st)
                ey = x[6 * j + 5];
            else {
                const double x_lo = x[6 * j + 4], x_hi = x[6 * (j + 1) + 4], y_lo = x[6 * j + 5], y_hi = x[6 * (j + 1) + 5];
                const double slope = (y_hi - y_lo) / (x_hi - x_lo);
                ey = slope * (st - x_lo) + y_lo;
            }
            double *t = track_rec + kp.track_xt_off + 6 * i;
            t[0] = vx; t[1] = 0.0; t[2] = 0.0; t[3] = 0.0; t[4] = 0.0; t[5] = ey;
        }
    }
}

}  // namespace b200mpc

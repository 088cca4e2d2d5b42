// b200mpc: prediction of rivals that have dynamics -- offboard.DynamicBicycleModel.get_trajectory_nsteps
// (car_racing/racing/offboard.py:80-94): n explicit-Euler steps of the zero-input Frenet kinematics of get_estimation
// (offboard.py:51-77) with the track curvature looked up every step (utils/racing_env.py:225-246), s wrapped into the lap
// after every step (:89-90).  The MPC-CBF controller (control.py:505-507, realtime_flag) and the planner
// (overtake_traj_planner.py:84-86) call it once per rival and control step.
//
// One thread per rival: a strictly sequential n-long recurrence (n = N+1 = 11 or 21), embarrassingly parallel across
// rivals / scenarios.  96 B in, n x 96 B out per rival, written in the reference's (6, n) layout; rows 4, 5 of the
// curvilinear block are what the solver's records take.  HBM-bound in principle, launch latency in practice.
// Reference quirks kept: xglob_est[4] is assigned twice (the X update is overwritten by the Y update) and
// xglob_est[5] stays 0 (offboard.py:71-76).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200mpc.h"

namespace b200mpc {

struct RolloutKParams {
    b200mpc_rollout_params p;
    int32_t B;
};

__device__ __forceinline__ double segment_curvature(const double *__restrict__ segments, int num_segments, double lap_length, double s) {
    while (s > lap_length) s -= lap_length;          // racing_env.py:231-234
    while (s < 0.0) s += lap_length;
    for (int g = 0; g < num_segments; g++) {
        const double s0 = segments[3 * g], len = segments[3 * g + 1];
        if (s >= s0 && s <= s0 + len) return segments[3 * g + 2];
    }
    return 0.0;
}

// one rival: xc0, xg0 6 doubles each; out_c, out_g (6, n) row-major
__device__ __forceinline__ void rollout_one(const b200mpc_rollout_params &p, const double *__restrict__ xc0,
                                            const double *__restrict__ xg0, const double *__restrict__ segments,
                                            double *__restrict__ out_c, double *__restrict__ out_g) {
    const int n = p.n;
    const double dt = p.timestep;
    double vx = xc0[0], vy = xc0[1], wz = xc0[2], epsi = xc0[3], s = xc0[4], ey = xc0[5];
    double g0 = xg0[0], g1 = xg0[1], g2 = xg0[2], psi = xg0[3], g4 = xg0[4];
    for (int k = 0; k < n; k++) {
        const double cur = segment_curvature(segments, p.num_segments, p.lap_length, s);
        const double se = sin(epsi), ce = cos(epsi), sp = sin(psi), cp = cos(psi);
        const double vlon = vx * ce - vy * se;
        const double e3 = epsi + dt * (wz - vlon / (1.0 - cur * ey) * cur);
        double e4 = s + dt * (vlon / (1.0 - cur * ey));
        const double e5 = ey + dt * (vx * se + vy * ce);
        const double n3 = psi + dt * g2;
        const double n4 = g4 + dt * (g0 * sp + g1 * cp);
        while (e4 > p.lap_length) e4 -= p.lap_length;                              // offboard.py:89-90
        epsi = e3; s = e4; ey = e5; psi = n3; g4 = n4;
        out_c[0 * n + k] = vx; out_c[1 * n + k] = vy; out_c[2 * n + k] = wz; out_c[3 * n + k] = epsi; out_c[4 * n + k] = s;
        out_c[5 * n + k] = ey;
        if (out_g != nullptr) {
            out_g[0 * n + k] = g0; out_g[1 * n + k] = g1; out_g[2 * n + k] = g2; out_g[3 * n + k] = psi; out_g[4 * n + k] = g4;
            out_g[5 * n + k] = 0.0;
        }
    }
}

__global__ void __launch_bounds__(128) rival_rollout_kernel(const __grid_constant__ RolloutKParams kp, const double *__restrict__ xcurv,
                                                            const double *__restrict__ xglob, const double *__restrict__ segments,
                                                            double *__restrict__ xcurv_n, double *__restrict__ xglob_n) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= kp.B) return;
    const size_t o = (size_t)b * 6 * kp.p.n;
    rollout_one(kp.p, xcurv + (size_t)b * 6, xglob + (size_t)b * 6, segments, xcurv_n + o, xglob_n != nullptr ? xglob_n + o : nullptr);
}

}  // namespace b200mpc

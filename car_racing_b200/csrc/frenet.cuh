// b200mpc: curvilinear -> global-frame conversion of predicted trajectories (last clause) --
// racing_env.get_global_position / get_orientation (car_racing/utils/racing_env.py:6-127, wrap / sign :268-283), which
// the controllers call per predicted stage to log x_pred in the global frame (utils/base.py:500-509, 573-580) and the
// planner per point of every candidate (planning/planner_helper.py:208-220: (num_veh+1) x 2 x (N+1) + 2 (N+1) calls per step).
//
// One thread per point: segment search over the track table (<= a dozen rows, L1-resident), then a straight-line blend
// or an arc.  16 B in, 24 B out per point; HBM-bound in principle, launch latency at the planner's sizes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200mpc.h"

namespace b200mpc {

// pat: num_segments x 6 rows (x, y, psi, start s, length, curvature) = ClosedTrack.point_and_tangent
__device__ __forceinline__ void curv_to_glob_one(const double *__restrict__ pat, int num_segments, double lap_length, double s,
                                                 double ey, double *x, double *y, double *psi_out) {
    const double PI = 3.141592653589793;
    while (s > lap_length) s -= lap_length;                                    // :12-15
    while (s < 0.0) s += lap_length;
    int i = num_segments - 1;
    for (int g = 0; g < num_segments; g++) {                                   // first match, s_tolerance = 0.001 (:11, 18-25)
        if (s >= pat[6 * g + 3] && s < pat[6 * g + 3] + pat[6 * g + 4] + 0.001) { i = g; break; }
    }
    const int ip = (i + num_segments - 1) % num_segments;                      // point_and_tangent[i - 1] (row -1 = last row)
    const double *row = pat + 6 * i, *prev = pat + 6 * ip;
    if (row[5] == 0.0) {                                                       // straight segment (:27-41)
        const double psi = row[2], rel = (s - row[3]) / row[4];
        *x = (1.0 - rel) * prev[0] + rel * row[0] + ey * cos(psi + PI / 2);
        *y = (1.0 - rel) * prev[1] + rel * row[1] + ey * sin(psi + PI / 2);
        *psi_out = psi;
        return;
    }
    const double r = 1.0 / row[5], ar = fabs(r), ang = prev[2];                // arc (:42-68)
    const double dir = (r >= 0.0) ? 1.0 : -1.0;
    const double cx = prev[0] + ar * cos(ang + dir * PI / 2), cy = prev[1] + ar * sin(ang + dir * PI / 2);
    const double span = (s - row[3]) / (PI * ar) * PI;
    double an = dir * PI / 2 + ang;
    an = (an < -PI) ? 2 * PI + an : ((an > PI) ? an - 2 * PI : an);            // wrap (:268-275)
    const double angle = -(PI - fabs(an)) * ((an >= 0.0) ? 1.0 : -1.0);        // sign (:278-283)
    *x = cx + (ar - dir * ey) * cos(angle + dir * span);
    *y = cy + (ar - dir * ey) * sin(angle + dir * span);
    *psi_out = angle + dir * span + PI / 2;                                    // get_orientation (:125)
}

// s, ey: P values each with strides (in doubles) so that columns 4, 5 of (.., 6) trajectories can be read in place;
// out: P x 3 (x, y, psi)
__global__ void __launch_bounds__(128) curv_to_glob_kernel(int P, int num_segments, double lap_length, const double *__restrict__ pat,
                                                           const double *__restrict__ s, int s_stride, const double *__restrict__ ey,
                                                           int ey_stride, double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    curv_to_glob_one(pat, num_segments, lap_length, s[(size_t)k * s_stride], ey[(size_t)k * ey_stride], out + 3 * (size_t)k,
                     out + 3 * (size_t)k + 1, out + 3 * (size_t)k + 2);
}

}  // namespace b200mpc

// b200mpc: batched plant step -- DynamicBicycleModel.forward_dynamics (car_racing/utils/base.py:897-942): n_sub explicit
// Euler sub-steps of vehicle_dynamics (car_racing/system/vehicle_dynamics.py:4-49: Pacejka tyres, Frenet + global
// kinematics) with the track curvature looked up every sub-step (utils/racing_env.py:225-246), then the clipped process
// noise (base.py:927-939) and the lap wrap of update_memory (base.py:804-809).  One thread per vehicle: the step is a
// strictly sequential 100-long recurrence per vehicle and embarrassingly parallel across the batch.  The next state can be
// written straight into the x0 slot of the MPC records (out stride/offset), so solve -> plant -> solve chains on the device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200mpc.h"

namespace b200mpc {

struct PlantKParams {
    b200mpc_plant_params p;
    int32_t B, xcurv_stride, xcurv_offset;
};

__global__ void __launch_bounds__(128) plant_kernel(const __grid_constant__ PlantKParams kp, double *__restrict__ xcurv,
                                                    double *__restrict__ xglob, const double *__restrict__ u, int u_stride,
                                                    const double *__restrict__ draws, const double *__restrict__ segments,
                                                    int32_t *__restrict__ laps) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= kp.B) return;
    const b200mpc_plant_params &p = kp.p;
    double *xc = xcurv + (size_t)b * kp.xcurv_stride + kp.xcurv_offset, *xg = xglob + (size_t)b * 6;
    double vx = xc[0], vy = xc[1], wz = xc[2], epsi = xc[3], s = xc[4], ey = xc[5];
    double psi = xg[3], X = xg[4], Y = xg[5];
    const double delta = u[(size_t)b * u_stride], a = u[(size_t)b * u_stride + 1];
    const double sd = sin(delta), cd = cos(delta), dt = p.delta_t;
    for (int it = 0; it < p.n_sub; it++) {
        double sw = s;
        while (sw > p.lap_length) sw -= p.lap_length;
        while (sw < 0.0) sw += p.lap_length;
        double cur = 0.0;
        for (int g = 0; g < p.num_segments; g++) {
            double s0 = __ldg(segments + 3 * g), len = __ldg(segments + 3 * g + 1);
            if (sw >= s0 && sw <= s0 + len) { cur = __ldg(segments + 3 * g + 2); break; }
        }
        double alpha_f = delta - atan2(vy + p.lf * wz, vx);
        double alpha_r = -atan2(vy - p.lf * wz, vx);            // lf, not lr: reference quirk (vehicle_dynamics.py:26)
        double Fyf = 2.0 * p.Df * sin(p.Cf * atan(p.Bf * alpha_f));
        double Fyr = 2.0 * p.Dr * sin(p.Cr * atan(p.Br * alpha_r));
        double se = sin(epsi), ce = cos(epsi), sp = sin(psi), cp = cos(psi);
        double n0 = vx + dt * (a - 1.0 / p.m * Fyf * sd + wz * vy);
        double n1 = vy + dt * (1.0 / p.m * (Fyf * cd + Fyr) - wz * vx);
        double n2 = wz + dt * (1.0 / p.Iz * (p.lf * Fyf * cd - p.lr * Fyr));
        double g3 = psi + dt * wz;
        double g4 = X + dt * (vx * cp - vy * sp);
        double g5 = Y + dt * (vx * sp + vy * cp);
        double c3 = epsi + dt * (wz - (vx * ce - vy * se) / (1.0 - cur * ey) * cur);
        double c4 = s + dt * ((vx * ce - vy * se) / (1.0 - cur * ey));
        double c5 = ey + dt * (vx * se + vy * ce);
        vx = n0; vy = n1; wz = n2; psi = g3; X = g4; Y = g5; epsi = c3; s = c4; ey = c5;
    }
    xg[0] = vx; xg[1] = vy; xg[2] = wz; xg[3] = psi; xg[4] = X; xg[5] = Y;
    if (draws != nullptr) {   // clipped noise enters the curvilinear state only (base.py:927-938)
        const double *d = draws + (size_t)b * 3;
        vx += 0.5 * fmax(-0.05, fmin(d[0] * 0.01, 0.05));
        vy += 0.5 * fmax(-0.1, fmin(d[1] * 0.01, 0.1));
        wz += 0.5 * fmax(-0.05, fmin(d[2] * 0.005, 0.05));
    }
    if (p.wrap_lap && s > p.lap_length) {   // update_memory (base.py:804-809)
        s -= p.lap_length;
        if (laps != nullptr) laps[b] += 1;
    }
    xc[0] = vx; xc[1] = vy; xc[2] = wz; xc[3] = epsi; xc[4] = s; xc[5] = ey;
}

}  // namespace b200mpc

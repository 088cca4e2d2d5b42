// TEST INFRASTRUCTURE: the per-point body of frenet.cuh compiled for the host (see cuda_runtime.h beside this file).
#include "cuda_runtime.h"
#include "../../car_racing_b200/csrc/frenet.cuh"

extern "C" void emu_curv_to_glob(int P, int num_segments, double lap_length, const double *pat, const double *s, const double *ey,
                                 double *out) {
    for (int k = 0; k < P; k++)
        b200mpc::curv_to_glob_one(pat, num_segments, lap_length, s[k], ey[k], out + 3 * k, out + 3 * k + 1, out + 3 * k + 2);
}

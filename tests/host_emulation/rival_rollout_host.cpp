// TEST INFRASTRUCTURE: the per-rival body of rival_rollout.cuh compiled for the host (see cuda_runtime.h beside this file).
#include "cuda_runtime.h"
#include "../../car_racing_b200/csrc/rival_rollout.cuh"

extern "C" void emu_rival_rollout(const b200mpc_rollout_params *prm, int B, const double *xcurv, const double *xglob,
                                  const double *segments, double *xcurv_n, double *xglob_n) {
    for (int b = 0; b < B; b++)
        b200mpc::rollout_one(*prm, xcurv + 6 * b, xglob + 6 * b, segments, xcurv_n + (size_t)b * 6 * prm->n,
                             xglob_n + (size_t)b * 6 * prm->n);
}

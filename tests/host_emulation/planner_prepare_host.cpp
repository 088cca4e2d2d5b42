// TEST INFRASTRUCTURE: the per-region body of planner_prepare.cuh compiled for the host (see cuda_runtime.h beside
// this file).  Built by tests/test_planner_prepare.py into tests/host_emulation/_build/ with
//   g++ -O1 -ffp-contract=off -shared -fPIC -I tests/host_emulation planner_prepare_host.cpp
#include "cuda_runtime.h"
#include "../../car_racing_b200/csrc/planner_prepare.cuh"

extern "C" int emu_planner_prepare(const b200mpc_planner_prepare_params *prm, int stride, int xt_off, int bnd_off, int wd_off,
                                   const double *ego, const double *rivals, const double *rival_vx, const int32_t *insertion,
                                   const double *opt, double *cand, double *heur, int32_t *ok0, int32_t *region, double *offset,
                                   double *ctrl, double *bezier) {
    b200mpc::PrepareKParams kp;
    kp.p = *prm;
    kp.stride = stride;
    kp.xt_off = xt_off;
    kp.bnd_off = bnd_off;
    kp.wd_off = wd_off;
    int err = 0;
    for (int c = 0; c <= prm->num_veh; c++)
        b200mpc::prepare_region(kp, c, ego, rivals, rival_vx, insertion, opt, cand, heur, ok0, region, offset, ctrl, bezier, &err);
    return err;
}

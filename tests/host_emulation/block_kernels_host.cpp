// TEST INFRASTRUCTURE: block-level kernels (one thread per item, shared memory, __syncthreads, no warp intrinsics) run on
// the host with one std::thread per CUDA thread (emu_launch in cuda_runtime.h beside this file): planner_select_kernel and
// plant_kernel exactly as the CUDA build compiles them.
#define EMU_WITH_LAUNCH
#include "cuda_runtime.h"
#include "../../car_racing_b200/csrc/planner_select.cuh"
#include "../../car_racing_b200/csrc/plant.cuh"

extern "C" void emu_planner_select(const b200mpc_planner_select_params *prm, int track_xt_off, const b200mpc_record *rec,
                                   const double *xpred, const double *heur, const int32_t *ok0, const int32_t *region,
                                   const double *rivals, double *sel_cost, int32_t *flag, double *traj, double *track_rec) {
    b200mpc::SelectKParams kp;
    kp.p = *prm;
    kp.track_xt_off = track_xt_off;
    emu_launch(1, b200mpc::SELECT_NT, 0, [&]() {
        b200mpc::planner_select_kernel(kp, rec, xpred, heur, ok0, region, rivals, sel_cost, flag, traj, track_rec);
    });
}

extern "C" void emu_plant_step(const b200mpc_plant_params *prm, int B, double *xcurv, int xcurv_stride, int xcurv_offset,
                               double *xglob, const double *u, int u_stride, const double *draws, const double *segments,
                               int32_t *laps) {
    b200mpc::PlantKParams kp;
    kp.p = *prm;
    kp.B = B;
    kp.xcurv_stride = xcurv_stride;
    kp.xcurv_offset = xcurv_offset;
    emu_launch((B + 127) / 128, 128, 0, [&]() { b200mpc::plant_kernel(kp, xcurv, xglob, u, u_stride, draws, segments, laps); });
}

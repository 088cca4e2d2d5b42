// TEST INFRASTRUCTURE: ocp_ipm_kernel -- the hot kernel's own source -- compiled by g++ and run with one host thread per
// lane (cuda_runtime.h beside this file: std::barrier per warp for __syncwarp and the shuffle / reduce collectives, a
// synchronous memcpy for the TMA bulk copy, exact seeds for the rcp / rsqrt approximations).  It checks the kernel's LOGIC
// (and that no phase relies on lock-step execution without a __syncwarp) against the oracle on machines without a GPU;
// the CUDA build of the same source is what the -m gpu tests and the bench run.
#define B200MPC_HOST_EMULATION
#define EMU_WITH_LAUNCH
#define EMU_WARPS
#define EMU_DYNAMIC_SMEM_ONLY
#include "cuda_runtime.h"
#include "../../include/b200mpc.h"
namespace b200mpc {
alignas(16) double sm[32768];   // the dynamic shared memory of the one block in flight (256 KB), NOT cleared between blocks
}
#include "../../car_racing_b200/csrc/ocp_ipm.cuh"

using namespace b200mpc;

// the KParams the C-ABI builds (car_racing_b200/csrc/capi.cu, make_kp)
static KParams make_kp(const b200mpc_cbf_params *p, const b200mpc_ipm_options *o, int B) {
    KParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.p = *p;
    kp.o = *o;
    kp.B = B;
    kp.in_stride = cbf_record_doubles(p->N, p->M, p->xt_per_stage, p->flags);
    kp.hdr = cbf_hdr_doubles(p->M);
    kp.obs_off = kp.hdr + (p->xt_per_stage ? 6 * (p->N + 1) : 6);
    kp.bnd_off = cbf_base_doubles(p->N, p->M, p->xt_per_stage);
    kp.wd_off = kp.bnd_off + ((p->flags & B200MPC_FLAG_STAGE_BOUNDS) ? 4 * (p->N + 1) : 0);
    kp.sz_off = kp.wd_off + ((p->flags & B200MPC_FLAG_EY_RATE) ? ((p->N + 1) & ~1) : 0);
    double L2 = p->L * p->L, W2 = p->W * p->W;
    kp.iL6 = 1.0 / (L2 * L2 * L2);
    kp.iW6 = 1.0 / (W2 * W2 * W2);
    kp.q_diag = 1;
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            kp.Q2[6 * a + b] = p->Q[6 * a + b] + p->Q[6 * b + a];
            if (a != b && p->Q[6 * a + b] != 0.0) kp.q_diag = 0;
        }
    return kp;
}

template <int M, int FL, int NT>
static void run(const KParams &kp, int B, const double *in, b200mpc_record *rec, double *aux, double *xpred, double *upred,
                double *sigma) {
    emu_launch(B, 32, 0, [&]() { ocp_ipm_kernel<M, FL, NT>(kp, in, rec, aux, xpred, upred, sigma); });
}

// the dispatch of b200mpc_cbf_solve_device; `specialised` = 0 forces the runtime-horizon instantiation
extern "C" int emu_cbf_solve(const b200mpc_cbf_params *prm, const b200mpc_ipm_options *opt, int B, int specialised, const double *in,
                             b200mpc_record *rec, double *aux, double *xpred, double *upred, double *sigma) {
    KParams kp = make_kp(prm, opt, B);
    if (prm->flags == (B200MPC_FLAG_STAGE_BOUNDS | B200MPC_FLAG_EY_RATE) && prm->M == 0) {
        run<0, 3, 0>(kp, B, in, rec, aux, xpred, upred, sigma);
        return 0;
    }
    if (prm->flags == B200MPC_FLAG_RIVAL_SIZE && prm->M == 2) {
        run<2, 4, 0>(kp, B, in, rec, aux, xpred, upred, sigma);
        return 0;
    }
    if (prm->flags != 0) return -1;
    if (specialised && prm->N == 20 && prm->M == 3 && !prm->xt_per_stage) {
        if (kp.q_diag) run<3, OCP_FL_QDIAG, 20>(kp, B, in, rec, aux, xpred, upred, sigma);   // as capi.cu dispatches
        else run<3, 0, 20>(kp, B, in, rec, aux, xpred, upred, sigma);
        return 0;
    }
    switch (prm->M) {
        case 0: run<0, 0, 0>(kp, B, in, rec, aux, xpred, upred, sigma); return 0;
        case 2: run<2, 0, 0>(kp, B, in, rec, aux, xpred, upred, sigma); return 0;
        case 3: run<3, 0, 0>(kp, B, in, rec, aux, xpred, upred, sigma); return 0;
    }
    return -1;
}

// b200mpc: the instantiation sets of ocp_ipm_kernel (see ocp_launch.cuh).  Compiled with OCP_INST_SET = k defined (one set per
// translation unit, CUDA build) or -DOCP_INST_EMU (tests/host_emulation: a reduced list inside capi's translation unit).
#pragma once
#include "ocp_launch.cuh"

namespace b200mpc {

template <int M, int FL, int NT>
static int launch_cbf_one(const CbfLaunch &l, const KParams &kp) {
    SmemPlan<M> pl(kp.p.N, kp.in_stride, FL);
    size_t smem = pl.bytes() + l.smem_pad;
    if ((int)smem > l.max_smem_optin) return CBF_LAUNCH_SMEM;
    // the opt-in shared-memory size is a per-device function attribute: set it once per (instantiation, device), and again
    // only when a longer horizon needs more
    static int granted[64];
    const int d = l.device & 63;
    if ((int)smem > granted[d]) {
        cudaError_t e = cudaFuncSetAttribute(ocp_ipm_kernel<M, FL, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        granted[d] = (int)smem;
    }
    ocp_ipm_kernel<M, FL, NT><<<kp.B, 32, smem, l.stream>>>(kp, l.in, l.rec, l.aux, l.x, l.u, l.sig, l.xa);
    return (int)cudaGetLastError();
}

#define OCP_CASE(M_, FL_, NT_) \
    if (M == M_ && FL == FL_ && NT == NT_) return launch_cbf_one<M_, FL_, NT_>(l, kp);

#if defined(OCP_INST_EMU)
// host emulation: every code path once (runtime horizon, compile-time horizon, planner blocks, per-rival sizes, M > 4),
// not every rival count -- g++ needs ~20 s per instantiation
int launch_cbf_set0(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT) {
    OCP_CASE(3, OCP_FL_QDIAG, 20) OCP_CASE(3, 0, 20) OCP_CASE(0, 0, 0) OCP_CASE(0, 3, 0)
    return CBF_LAUNCH_NOT_HERE;
}
int launch_cbf_set1(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT) {
    OCP_CASE(1, 0, 0) OCP_CASE(2, 0, 0) OCP_CASE(3, 0, 0) OCP_CASE(4, 0, 0)
    return CBF_LAUNCH_NOT_HERE;
}
int launch_cbf_set2(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT) {
    OCP_CASE(5, 0, 0)
    return CBF_LAUNCH_NOT_HERE;
}
int launch_cbf_set3(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT) {
    OCP_CASE(2, 4, 0)
    return CBF_LAUNCH_NOT_HERE;
}
int launch_cbf_set4(const CbfLaunch &, const KParams &, int, int, int) { return CBF_LAUNCH_NOT_HERE; }
#else
#if OCP_INST_SET == 0
int launch_cbf_set0(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT) {
    OCP_CASE(3, OCP_FL_QDIAG, 20) OCP_CASE(3, 0, 20) OCP_CASE(0, 0, 0) OCP_CASE(0, 3, 0)
    return CBF_LAUNCH_NOT_HERE;
}
#elif OCP_INST_SET == 1
int launch_cbf_set1(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT) {
    OCP_CASE(1, 0, 0) OCP_CASE(2, 0, 0) OCP_CASE(3, 0, 0) OCP_CASE(4, 0, 0)
    return CBF_LAUNCH_NOT_HERE;
}
#elif OCP_INST_SET == 2
int launch_cbf_set2(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT) {
    OCP_CASE(5, 0, 0) OCP_CASE(6, 0, 0) OCP_CASE(7, 0, 0) OCP_CASE(8, 0, 0)
    return CBF_LAUNCH_NOT_HERE;
}
#elif OCP_INST_SET == 3
int launch_cbf_set3(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT) {
    OCP_CASE(1, 4, 0) OCP_CASE(2, 4, 0) OCP_CASE(3, 4, 0) OCP_CASE(4, 4, 0)
    return CBF_LAUNCH_NOT_HERE;
}
#elif OCP_INST_SET == 4
int launch_cbf_set4(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT) {
    OCP_CASE(5, 4, 0) OCP_CASE(6, 4, 0) OCP_CASE(7, 4, 0) OCP_CASE(8, 4, 0)
    return CBF_LAUNCH_NOT_HERE;
}
#else
#error "OCP_INST_SET must be 0..4"
#endif
#endif
#undef OCP_CASE

}  // namespace b200mpc

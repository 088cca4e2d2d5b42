// b200mpc: launch interface of ocp_ipm_kernel's template instantiations.
//
// The instantiations are compiled in several translation units (generated wrappers "#define OCP_INST_SET k + #include ocp_inst.cuh", built in parallel by
// __graft_entry__.build()) because each one is ~8 k SASS instructions; capi.cu only sees this header.
//   set 0: <3,QDIAG,20> (BASELINE north star: diagonal Q at compile time), <3,0,20> (the same shape, any Q), <0,0,0> (mpc_lti),
//          <0,3,0> (planner candidate QP)
//   set 1: <1..4,0,0>      set 2: <5..8,0,0>
//   set 3: <1..4,4,0>      set 4: <5..8,4,0>     (flag RIVAL_SIZE: per-rival (L, W) in the record)
#pragma once
#include "ocp_ipm.cuh"

namespace b200mpc {

struct CbfLaunch {
    cudaStream_t stream;
    int device;
    int max_smem_optin;
    size_t smem_pad;       // measurement hook (B200MPC_SMEM_PAD), 0 in production
    const double *in;
    b200mpc_record *rec;
    double *aux, *x, *u, *sig;
    XchgArgs xa;           // tab == nullptr: no exchange (exchange.cuh)
};

enum { CBF_LAUNCH_NOT_HERE = -1000, CBF_LAUNCH_SMEM = -1001 };

// Each returns cudaSuccess (0) / a cudaError_t (> 0) / CBF_LAUNCH_SMEM / CBF_LAUNCH_NOT_HERE (instantiation not in this set)
int launch_cbf_set0(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT);
int launch_cbf_set1(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT);
int launch_cbf_set2(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT);
int launch_cbf_set3(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT);
int launch_cbf_set4(const CbfLaunch &l, const KParams &kp, int M, int FL, int NT);

}  // namespace b200mpc

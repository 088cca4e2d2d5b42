// b200mpc: C-ABI (include/b200mpc.h) over the sm_100a kernels.  No torch types, no CPU
// fallback: every entry point either launches the CUDA kernels or fails with an error code.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "../../include/b200mpc.h"
#include "exchange.cuh"
#include "frenet.cuh"
#include "ilqr.cuh"
#include "lmpc.cuh"
#include "ocp_launch.cuh"
#ifdef B200MPC_HOST_EMULATION   // tests/host_emulation compiles one translation unit: a reduced instantiation list inline
#define OCP_INST_EMU
#include "ocp_inst.cuh"
#endif
#include "plant.cuh"
#include "planner_prepare.cuh"
#include "planner_select.cuh"
#include "rival_rollout.cuh"
#include "sysid.cuh"

using namespace b200mpc;

struct b200mpc_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int max_batch = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    size_t smem_pad = 0;   // B200MPC_SMEM_PAD, read once at create (occupancy sweeps only)
    // staging buffers for the host-pointer API (grown on demand)
    void *d_in = nullptr, *d_rec = nullptr, *d_aux = nullptr, *d_x = nullptr, *d_u = nullptr, *d_sig = nullptr;
    void *d_laps = nullptr, *d_seg = nullptr, *d_idx = nullptr, *d_stat = nullptr, *d_chain = nullptr;
    size_t c_in = 0, c_rec = 0, c_aux = 0, c_x = 0, c_u = 0, c_sig = 0, c_laps = 0, c_seg = 0, c_idx = 0, c_stat = 0, c_chain = 0;
    uint64_t launches = 0;
    std::string err;
    // one-shot: the next solver launch also publishes its records into this slot of the exchange window
    b200mpc_comm *xchg_comm = nullptr;
    int xchg_slot = 0;
};

// exchange window of one rank (csrc/exchange.cuh)
struct b200mpc_comm {
    int device = 0, rank = 0, world = 1, max_batch = 0, slots = 0;
    void *window = nullptr;                      // own window: rec | cnt | ack
    size_t window_bytes = 0;
    void *peer_base[XCHG_MAX_WORLD] = {};        // every rank's window in this process' address space
    bool peer_opened[XCHG_MAX_WORLD] = {};
    XchgTable *d_table = nullptr;
    bool connected = false;
    struct Use { int B; unsigned long long cum; };
    static constexpr int RING = 8;
    unsigned long long published[64] = {}, consumed[64] = {}, cum[64] = {};
    Use ring[64][RING];
};

static thread_local std::string g_create_err;

static int fail(b200mpc_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    else g_create_err = msg;
    return code;
}
#define CK(h, call)                                                                                          \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(h, B200MPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

static int grow(b200mpc_handle *h, void **p, size_t *cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) return fail(h, B200MPC_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    *cap = need;
    return 0;
}

extern "C" {

int b200mpc_version(void) { return B200MPC_VERSION; }

void b200mpc_default_ipm_options(b200mpc_ipm_options *o) {
    o->tol = 1e-8;
    o->max_iter = 200;
    o->acceptable_iter = 15;
    o->acceptable_tol = 1e-6;
    o->mu_init = 0.1;
    o->rho = 1e3;
    o->bound_push = 1e-2;
    o->bound_frac = 1e-2;
    o->max_grad = 100.0;
    o->start = B200MPC_START_ROLLOUT;
    o->max_reset = 5;
}

int b200mpc_create(int device, int max_batch, b200mpc_handle **out) { return b200mpc_create_ex(device, max_batch, 0, out); }

int b200mpc_create_ex(int device, int max_batch, int high_priority, b200mpc_handle **out) {
    if (!out || max_batch < 1) return fail(nullptr, B200MPC_ERR_ARG, "b200mpc_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, B200MPC_ERR_NODEVICE,
                    std::string("b200mpc_create: no CUDA device (") + cudaGetErrorString(e) + ") -- there is no CPU fallback");
    if (device < 0) {
        e = cudaGetDevice(&device);
        if (e != cudaSuccess) return fail(nullptr, B200MPC_ERR_CUDA, cudaGetErrorString(e));
    }
    if (device >= ndev) return fail(nullptr, B200MPC_ERR_ARG, "b200mpc_create: device index out of range");
    b200mpc_handle *h = new (std::nothrow) b200mpc_handle();
    if (!h) return fail(nullptr, B200MPC_ERR_NOMEM, "out of host memory");
    h->device = device;
    h->max_batch = max_batch;
    int prio_lo = 0, prio_hi = 0;   // numerically lower = higher priority; pending blocks of a high-priority stream are placed first
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi)) != cudaSuccess ||
        (e = cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, high_priority ? prio_hi : 0)) != cudaSuccess) {
        std::string m = cudaGetErrorString(e);
        delete h;
        return fail(nullptr, B200MPC_ERR_CUDA, "b200mpc_create: " + m);
    }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    // B200MPC_SMEM_PAD=<bytes>: measurement hook only (occupancy sweep) -- pads the dynamic shared
    // memory of ocp_ipm_kernel so that fewer CTAs fit an SM; results are unaffected
    if (const char *pad = getenv("B200MPC_SMEM_PAD")) h->smem_pad = (size_t)atoi(pad);
    *out = h;
    return B200MPC_OK;
}

void b200mpc_destroy(b200mpc_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void *bufs[] = {h->d_in, h->d_rec, h->d_aux, h->d_x, h->d_u, h->d_sig, h->d_laps, h->d_seg, h->d_idx, h->d_stat, h->d_chain};
    for (void *b : bufs)
        if (b) cudaFree(b);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char *b200mpc_last_error(const b200mpc_handle *h) { return h ? h->err.c_str() : g_create_err.c_str(); }
uint64_t b200mpc_stream(const b200mpc_handle *h) { return h ? (uint64_t)(uintptr_t)h->stream : 0; }
uint64_t b200mpc_launch_count(const b200mpc_handle *h) { return h ? h->launches : 0; }

int b200mpc_cbf_record_doubles(int N, int M, int xt_per_stage) {
    if (N < 1 || N > B200MPC_NMAX || M < 0 || M > B200MPC_MMAX) return B200MPC_ERR_ARG;
    return cbf_record_doubles(N, M, xt_per_stage);
}
int b200mpc_cbf_record_doubles_ex(int N, int M, int xt_per_stage, int flags) {
    if (N < 1 || N > B200MPC_NMAX || M < 0 || M > B200MPC_MMAX) return B200MPC_ERR_ARG;
    return cbf_record_doubles(N, M, xt_per_stage, flags);
}
int b200mpc_ilqr_record_doubles(int N) {
    if (N < 1 || N > B200MPC_NMAX) return B200MPC_ERR_ARG;
    return ilqr_record_doubles(N);
}
int b200mpc_lmpc_record_doubles(int N, int K) {
    if (N < 2 || N > B200MPC_LMPC_NMAX || K < 1 || K > B200MPC_LMPC_KMAX) return B200MPC_ERR_ARG;
    return lmpc_record_doubles(N, K);
}

}  // extern "C"

// The one-shot exchange request of b200mpc_comm_publish_next: consumed by the next solver launch on the handle.
static int take_xchg(b200mpc_handle *h, int B, XchgArgs *xa) {
    *xa = XchgArgs{nullptr, 0, 0, 0};
    b200mpc_comm *c = h->xchg_comm;
    if (!c) return B200MPC_OK;
    h->xchg_comm = nullptr;
    if (B > c->max_batch) return fail(h, B200MPC_ERR_ARG, "exchange: B exceeds the max_batch of the communicator");
    const int s = h->xchg_slot;
    if (c->published[s] - c->consumed[s] >= (unsigned long long)b200mpc_comm::RING)
        return fail(h, B200MPC_ERR_ARG, "exchange: too many unconsumed uses of one slot (call b200mpc_comm_argmin)");
    c->published[s]++;
    c->cum[s] += (unsigned long long)B;
    c->ring[s][c->published[s] % b200mpc_comm::RING] = {B, c->cum[s]};
    *xa = XchgArgs{c->d_table, s, B, c->published[s]};
    if (c->world > 1 && c->published[s] > 1) {   // gate: every rank has consumed the previous use of this slot
        xchg_wait_ack_kernel<<<1, 32, 0, h->stream>>>(c->d_table, s, c->published[s]);
        CK(h, cudaGetLastError());
        h->launches++;
    }
    return B200MPC_OK;
}

static int check_cbf(b200mpc_handle *h, const b200mpc_cbf_params *p, const b200mpc_ipm_options *o, int B, const void *in,
                     const void *rec) {
    if (!h) return B200MPC_ERR_ARG;
    if (!p || !o || !in || !rec || B < 1) return fail(h, B200MPC_ERR_ARG, "b200mpc_cbf_solve: null argument or B < 1");
    if (p->N < 1 || p->N > B200MPC_NMAX || p->M < 0 || p->M > B200MPC_MMAX)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_cbf_solve: N or M out of range");
    if (!(p->alpha >= 0.0 && p->alpha <= 1.0) || !(p->L > 0.0) || !(p->W > 0.0) || !(o->tol > 0.0) || o->max_iter < 1 ||
        o->max_reset < 0 || (o->start != B200MPC_START_ROLLOUT && o->start != B200MPC_START_ZERO))
        return fail(h, B200MPC_ERR_ARG, "b200mpc_cbf_solve: bad parameter value");
    return B200MPC_OK;
}

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
#ifdef B200MPC_NO_QDIAG   // A/B switch (tools/variants.sh): always the general path
    kp.q_diag = 0;
#endif
    return kp;
}

extern "C" {

int b200mpc_cbf_solve_device(b200mpc_handle *h, const b200mpc_cbf_params *prm, const b200mpc_ipm_options *opt, int B,
                             const double *d_in, b200mpc_record *d_rec, double *d_aux, double *d_xpred, double *d_upred,
                             double *d_sigma) {
    int rc = check_cbf(h, prm, opt, B, d_in, d_rec);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    KParams kp = make_kp(prm, opt, B);
    // which instantiation <M, FLAGS, NT> (ocp_launch.cuh): the planner-candidate blocks are compiled in only where the reference
    // uses them (no rival rows there); the per-rival sizes only with rival rows; the BASELINE.json north-star shape (N = 20,
    // M = 3, one target) has a horizon-specialised instantiation (per-stage xtarget excluded: the record stride differs)
    int M = prm->M, FL = prm->flags, NT = 0;
    if (M == 0) FL &= ~B200MPC_FLAG_RIVAL_SIZE;
    if (FL == (B200MPC_FLAG_STAGE_BOUNDS | B200MPC_FLAG_EY_RATE)) {
        if (M != 0) return fail(h, B200MPC_ERR_ARG, "b200mpc_cbf_solve: STAGE_BOUNDS|EY_RATE is supported with M = 0 only");
    } else if (FL != 0 && FL != B200MPC_FLAG_RIVAL_SIZE)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_cbf_solve: flags must be 0, STAGE_BOUNDS|EY_RATE (M = 0) or RIVAL_SIZE");
    if (FL == 0 && prm->N == 20 && M == 3 && !prm->xt_per_stage) NT = 20;
    if (NT == 20 && kp.q_diag) FL = OCP_FL_QDIAG;   // the north star with the reference's diagonal Q: no general path in the kernel
    XchgArgs xa;
    if ((rc = take_xchg(h, B, &xa))) return rc;
    CbfLaunch l{h->stream, h->device, h->max_smem_optin, h->smem_pad, d_in, d_rec, d_aux, d_xpred, d_upred, d_sigma, xa};
    int e = (FL == B200MPC_FLAG_RIVAL_SIZE) ? (M <= 4 ? launch_cbf_set3(l, kp, M, FL, NT) : launch_cbf_set4(l, kp, M, FL, NT))
            : (M == 0 || NT) ? launch_cbf_set0(l, kp, M, FL, NT)
            : (M <= 4)       ? launch_cbf_set1(l, kp, M, FL, NT)
                             : launch_cbf_set2(l, kp, M, FL, NT);
    if (e == CBF_LAUNCH_SMEM) return fail(h, B200MPC_ERR_ARG, "b200mpc_cbf_solve: horizon too long for one CTA's shared memory");
    if (e == CBF_LAUNCH_NOT_HERE) return fail(h, B200MPC_ERR_ARG, "b200mpc_cbf_solve: no kernel instantiation for this (M, flags)");
    if (e != 0) return fail(h, B200MPC_ERR_CUDA, std::string("ocp_ipm_kernel launch: ") + cudaGetErrorString((cudaError_t)e));
    h->launches++;
    return B200MPC_OK;
}

void *b200mpc_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void b200mpc_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int b200mpc_synchronize(b200mpc_handle *h) {
    if (!h) return B200MPC_ERR_ARG;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

int b200mpc_cbf_solve(b200mpc_handle *h, const b200mpc_cbf_params *prm, const b200mpc_ipm_options *opt, int B,
                      const double *in, b200mpc_record *rec, double *aux, double *xpred, double *upred, double *sigma) {
    int rc = b200mpc_cbf_solve_async(h, prm, opt, B, in, rec, aux, xpred, upred, sigma);
    if (rc) return rc;
    return b200mpc_synchronize(h);
}

int b200mpc_cbf_solve_async(b200mpc_handle *h, const b200mpc_cbf_params *prm, const b200mpc_ipm_options *opt, int B,
                            const double *in, b200mpc_record *rec, double *aux, double *xpred, double *upred, double *sigma) {
    int rc = check_cbf(h, prm, opt, B, in, rec);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    const int N = prm->N, M = prm->M;
    const size_t stride = (size_t)cbf_record_doubles(N, M, prm->xt_per_stage, prm->flags);
    const size_t b_in = stride * 8 * B, b_rec = sizeof(b200mpc_record) * (size_t)B, b_aux = 32 * (size_t)B;
    const size_t b_x = 48 * (size_t)(N + 1) * B, b_u = 16 * (size_t)N * B, b_sig = 8 * (size_t)M * (N + 1) * B;
    if ((rc = grow(h, &h->d_in, &h->c_in, b_in))) return rc;
    if ((rc = grow(h, &h->d_rec, &h->c_rec, b_rec))) return rc;
    if (aux && (rc = grow(h, &h->d_aux, &h->c_aux, b_aux))) return rc;
    if (xpred && (rc = grow(h, &h->d_x, &h->c_x, b_x))) return rc;
    if (upred && (rc = grow(h, &h->d_u, &h->c_u, b_u))) return rc;
    if (sigma && M > 0 && (rc = grow(h, &h->d_sig, &h->c_sig, b_sig))) return rc;
    CK(h, cudaMemcpyAsync(h->d_in, in, b_in, cudaMemcpyHostToDevice, h->stream));
    rc = b200mpc_cbf_solve_device(h, prm, opt, B, (const double *)h->d_in, (b200mpc_record *)h->d_rec,
                                  aux ? (double *)h->d_aux : nullptr, xpred ? (double *)h->d_x : nullptr,
                                  upred ? (double *)h->d_u : nullptr, (sigma && M > 0) ? (double *)h->d_sig : nullptr);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(rec, h->d_rec, b_rec, cudaMemcpyDeviceToHost, h->stream));
    if (aux) CK(h, cudaMemcpyAsync(aux, h->d_aux, b_aux, cudaMemcpyDeviceToHost, h->stream));
    if (xpred) CK(h, cudaMemcpyAsync(xpred, h->d_x, b_x, cudaMemcpyDeviceToHost, h->stream));
    if (upred) CK(h, cudaMemcpyAsync(upred, h->d_u, b_u, cudaMemcpyDeviceToHost, h->stream));
    if (sigma && M > 0) CK(h, cudaMemcpyAsync(sigma, h->d_sig, b_sig, cudaMemcpyDeviceToHost, h->stream));
    return B200MPC_OK;
}

static int check_ilqr(b200mpc_handle *h, const b200mpc_ilqr_params *p, int B, const void *in, const void *rec) {
    if (!h) return B200MPC_ERR_ARG;
    if (!p || !in || !rec || B < 1) return fail(h, B200MPC_ERR_ARG, "b200mpc_ilqr_solve: null argument or B < 1");
    if (p->N < 1 || p->N > B200MPC_NMAX || p->max_iter < 1 || !(p->L > 0.0) || !(p->W > 0.0))
        return fail(h, B200MPC_ERR_ARG, "b200mpc_ilqr_solve: bad parameter value");
    return B200MPC_OK;
}

int b200mpc_ilqr_solve_device(b200mpc_handle *h, const b200mpc_ilqr_params *prm, int B, const double *d_in,
                              b200mpc_record *d_rec, double *d_xpred, double *d_upred) {
    int rc = check_ilqr(h, prm, B, d_in, d_rec);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    IlqrKParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.p = *prm;
    kp.B = B;
    kp.in_stride = ilqr_record_doubles(prm->N);
    IlqrPlan pl(prm->N, kp.in_stride);
    size_t smem = pl.bytes();
    CK(h, cudaFuncSetAttribute(ilqr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XchgArgs xa;
    if ((rc = take_xchg(h, B, &xa))) return rc;
    ilqr_kernel<<<B, 32, smem, h->stream>>>(kp, d_in, d_rec, d_xpred, d_upred, xa);
    CK(h, cudaGetLastError());
    h->launches++;
    return B200MPC_OK;
}

int b200mpc_ilqr_solve(b200mpc_handle *h, const b200mpc_ilqr_params *prm, int B, const double *in, b200mpc_record *rec,
                       double *xpred, double *upred) {
    int rc = b200mpc_ilqr_solve_async(h, prm, B, in, rec, xpred, upred);
    if (rc) return rc;
    return b200mpc_synchronize(h);
}

int b200mpc_ilqr_solve_async(b200mpc_handle *h, const b200mpc_ilqr_params *prm, int B, const double *in, b200mpc_record *rec,
                             double *xpred, double *upred) {
    int rc = check_ilqr(h, prm, B, in, rec);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    const int N = prm->N;
    const size_t b_in = (size_t)ilqr_record_doubles(N) * 8 * B, b_rec = sizeof(b200mpc_record) * (size_t)B;
    const size_t b_x = 48 * (size_t)(N + 1) * B, b_u = 16 * (size_t)N * B;
    if ((rc = grow(h, &h->d_in, &h->c_in, b_in))) return rc;
    if ((rc = grow(h, &h->d_rec, &h->c_rec, b_rec))) return rc;
    if (xpred && (rc = grow(h, &h->d_x, &h->c_x, b_x))) return rc;
    if (upred && (rc = grow(h, &h->d_u, &h->c_u, b_u))) return rc;
    CK(h, cudaMemcpyAsync(h->d_in, in, b_in, cudaMemcpyHostToDevice, h->stream));
    rc = b200mpc_ilqr_solve_device(h, prm, B, (const double *)h->d_in, (b200mpc_record *)h->d_rec,
                                   xpred ? (double *)h->d_x : nullptr, upred ? (double *)h->d_u : nullptr);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(rec, h->d_rec, b_rec, cudaMemcpyDeviceToHost, h->stream));
    if (xpred) CK(h, cudaMemcpyAsync(xpred, h->d_x, b_x, cudaMemcpyDeviceToHost, h->stream));
    if (upred) CK(h, cudaMemcpyAsync(upred, h->d_u, b_u, cudaMemcpyDeviceToHost, h->stream));
    return B200MPC_OK;
}

static int check_lmpc(b200mpc_handle *h, const b200mpc_lmpc_params *p, const b200mpc_ipm_options *o, int B, const void *in,
                      const void *rec) {
    if (!h) return B200MPC_ERR_ARG;
    if (!p || !o || !in || !rec || B < 1) return fail(h, B200MPC_ERR_ARG, "b200mpc_lmpc_solve: null argument or B < 1");
    if (p->N < 2 || p->N > B200MPC_LMPC_NMAX || p->K < 1 || p->K > B200MPC_LMPC_KMAX)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_lmpc_solve: N or K out of range");
    if (!(p->umax[0] > 0.0) || !(p->umax[1] > 0.0) || !(p->width > 0.0) || !(o->tol > 0.0) || o->max_iter < 1)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_lmpc_solve: bad parameter value");
    return B200MPC_OK;
}

int b200mpc_lmpc_solve_device(b200mpc_handle *h, const b200mpc_lmpc_params *prm, const b200mpc_ipm_options *opt, int B,
                              const double *d_in, b200mpc_record *d_rec, double *d_aux, double *d_xpred, double *d_upred,
                              double *d_lambda) {
    int rc = check_lmpc(h, prm, opt, B, d_in, d_rec);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    LmpcKParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.p = *prm;
    kp.o = *opt;
    kp.B = B;
    kp.in_stride = lmpc_record_doubles(prm->N, prm->K);
    LmpcPlan pl(prm->N, prm->K, kp.in_stride);
    size_t smem = pl.bytes();
    if ((int)smem > h->max_smem_optin)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_lmpc_solve: N, K too large for one CTA's shared memory");
    CK(h, cudaFuncSetAttribute(lmpc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XchgArgs xa;
    if ((rc = take_xchg(h, B, &xa))) return rc;
    lmpc_kernel<<<B, LMPC_NT, smem, h->stream>>>(kp, d_in, d_rec, d_aux, d_xpred, d_upred, d_lambda, xa);
    CK(h, cudaGetLastError());
    h->launches++;
    return B200MPC_OK;
}

int b200mpc_lmpc_solve(b200mpc_handle *h, const b200mpc_lmpc_params *prm, const b200mpc_ipm_options *opt, int B,
                       const double *in, b200mpc_record *rec, double *aux, double *xpred, double *upred, double *lambda) {
    int rc = b200mpc_lmpc_solve_async(h, prm, opt, B, in, rec, aux, xpred, upred, lambda);
    if (rc) return rc;
    return b200mpc_synchronize(h);
}

int b200mpc_lmpc_solve_async(b200mpc_handle *h, const b200mpc_lmpc_params *prm, const b200mpc_ipm_options *opt, int B,
                             const double *in, b200mpc_record *rec, double *aux, double *xpred, double *upred, double *lambda) {
    int rc = check_lmpc(h, prm, opt, B, in, rec);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    const int N = prm->N, K = prm->K;
    const size_t b_in = (size_t)lmpc_record_doubles(N, K) * 8 * B, b_rec = sizeof(b200mpc_record) * (size_t)B, b_aux = 32 * (size_t)B;
    const size_t b_x = 48 * (size_t)(N + 1) * B, b_u = 16 * (size_t)N * B, b_l = 8 * (size_t)K * B;
    if ((rc = grow(h, &h->d_in, &h->c_in, b_in))) return rc;
    if ((rc = grow(h, &h->d_rec, &h->c_rec, b_rec))) return rc;
    if (aux && (rc = grow(h, &h->d_aux, &h->c_aux, b_aux))) return rc;
    if (xpred && (rc = grow(h, &h->d_x, &h->c_x, b_x))) return rc;
    if (upred && (rc = grow(h, &h->d_u, &h->c_u, b_u))) return rc;
    if (lambda && (rc = grow(h, &h->d_sig, &h->c_sig, b_l))) return rc;
    CK(h, cudaMemcpyAsync(h->d_in, in, b_in, cudaMemcpyHostToDevice, h->stream));
    rc = b200mpc_lmpc_solve_device(h, prm, opt, B, (const double *)h->d_in, (b200mpc_record *)h->d_rec,
                                   aux ? (double *)h->d_aux : nullptr, xpred ? (double *)h->d_x : nullptr,
                                   upred ? (double *)h->d_u : nullptr, lambda ? (double *)h->d_sig : nullptr);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(rec, h->d_rec, b_rec, cudaMemcpyDeviceToHost, h->stream));
    if (aux) CK(h, cudaMemcpyAsync(aux, h->d_aux, b_aux, cudaMemcpyDeviceToHost, h->stream));
    if (xpred) CK(h, cudaMemcpyAsync(xpred, h->d_x, b_x, cudaMemcpyDeviceToHost, h->stream));
    if (upred) CK(h, cudaMemcpyAsync(upred, h->d_u, b_u, cudaMemcpyDeviceToHost, h->stream));
    if (lambda) CK(h, cudaMemcpyAsync(lambda, h->d_sig, b_l, cudaMemcpyDeviceToHost, h->stream));
    return B200MPC_OK;
}

static size_t sysid_smem(const b200mpc_sysid_params *p) {
    size_t d = (size_t)((p->lap_stride + 1) & ~1) + (size_t)((p->num_laps * p->max_num_point + 1) & ~1) / 2 +
               (size_t)p->num_laps * p->max_num_point + 40 + 96 + 56;
    return d * sizeof(double);
}

static int check_sysid(b200mpc_handle *h, const b200mpc_sysid_params *p, int B, const void *lin, const void *laps,
                       const void *seg, const void *out, int out_stride, int out_offset) {
    if (!h) return B200MPC_ERR_ARG;
    if (!p || !lin || !laps || !seg || !out || B < 1) return fail(h, B200MPC_ERR_ARG, "b200mpc_lmpc_sysid: null argument or B < 1");
    if (p->N < 1 || p->N > B200MPC_NMAX || p->num_laps < 1 || p->num_laps > B200MPC_SYSID_LMAX || p->max_num_point < 1 ||
        p->max_num_point > B200MPC_SYSID_PMAX || p->num_segments < 1 || p->lap_stride < 2)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_lmpc_sysid: size out of range");
    for (int l = 0; l < p->num_laps; l++)
        if (p->lap_rows[l] < 2 || p->lap_rows[l] >= p->lap_stride)
            return fail(h, B200MPC_ERR_ARG, "b200mpc_lmpc_sysid: lap_rows must satisfy 2 <= time_ss < lap_stride");
    if (!(p->dt > 0.0) || !(p->h > 0.0) || !(p->lap_length > 0.0) || out_offset < 0 || out_stride < out_offset + 54 * p->N)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_lmpc_sysid: bad parameter value");
    if ((int)sysid_smem(p) > h->max_smem_optin) return fail(h, B200MPC_ERR_ARG, "b200mpc_lmpc_sysid: laps too long for shared memory");
    return B200MPC_OK;
}

int b200mpc_lmpc_sysid_device(b200mpc_handle *h, const b200mpc_sysid_params *prm, int B, const double *d_lin,
                              const double *d_laps, const double *d_segments, double *d_out, int out_stride, int out_offset,
                              int32_t *d_idx, int32_t *d_status) {
    int rc = check_sysid(h, prm, B, d_lin, d_laps, d_segments, d_out, out_stride, out_offset);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    SysidKParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.p = *prm;
    kp.B = B;
    kp.out_stride = out_stride;
    kp.out_offset = out_offset;
    size_t smem = sysid_smem(prm);
    CK(h, cudaFuncSetAttribute(sysid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sysid_kernel<<<B * prm->N, 32, smem, h->stream>>>(kp, d_lin, d_laps, d_segments, d_out, d_idx, d_status);
    CK(h, cudaGetLastError());
    h->launches++;
    return B200MPC_OK;
}

int b200mpc_lmpc_sysid(b200mpc_handle *h, const b200mpc_sysid_params *prm, int B, const double *lin, const double *laps,
                       const double *segments, double *out, int out_stride, int out_offset, int32_t *idx, int32_t *status) {
    int rc = check_sysid(h, prm, B, lin, laps, segments, out, out_stride, out_offset);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    const int N = prm->N;
    const size_t b_lin = (size_t)B * N * 64, b_laps = (size_t)prm->num_laps * 5 * prm->lap_stride * 8, b_seg = (size_t)prm->num_segments * 24;
    const size_t b_out = (size_t)B * out_stride * 8, b_idx = (size_t)B * N * prm->num_laps * prm->max_num_point * 4, b_st = (size_t)B * N * 4;
    if ((rc = grow(h, &h->d_in, &h->c_in, b_lin))) return rc;
    if ((rc = grow(h, &h->d_laps, &h->c_laps, b_laps))) return rc;
    if ((rc = grow(h, &h->d_seg, &h->c_seg, b_seg))) return rc;
    if ((rc = grow(h, &h->d_x, &h->c_x, b_out))) return rc;
    if (idx && (rc = grow(h, &h->d_idx, &h->c_idx, b_idx))) return rc;
    if (status && (rc = grow(h, &h->d_stat, &h->c_stat, b_st))) return rc;
    CK(h, cudaMemcpyAsync(h->d_in, lin, b_lin, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_laps, laps, b_laps, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_seg, segments, b_seg, cudaMemcpyHostToDevice, h->stream));
    rc = b200mpc_lmpc_sysid_device(h, prm, B, (const double *)h->d_in, (const double *)h->d_laps, (const double *)h->d_seg,
                                   (double *)h->d_x, out_stride, out_offset, idx ? (int32_t *)h->d_idx : nullptr,
                                   status ? (int32_t *)h->d_stat : nullptr);
    if (rc) return rc;
    // only the model block of each record comes back: the caller's other fields stay untouched
    CK(h, cudaMemcpy2DAsync(out + out_offset, (size_t)out_stride * 8, (const double *)h->d_x + out_offset, (size_t)out_stride * 8,
                            (size_t)54 * N * 8, B, cudaMemcpyDeviceToHost, h->stream));
    if (idx) CK(h, cudaMemcpyAsync(idx, h->d_idx, b_idx, cudaMemcpyDeviceToHost, h->stream));
    if (status) CK(h, cudaMemcpyAsync(status, h->d_stat, b_st, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

static int check_plant(b200mpc_handle *h, const b200mpc_plant_params *p, int B, const void *xc, int stride, int offset,
                       const void *xg, const void *u, int u_stride, const void *seg) {
    if (!h) return B200MPC_ERR_ARG;
    if (!p || !xc || !xg || !u || !seg || B < 1) return fail(h, B200MPC_ERR_ARG, "b200mpc_plant_step: null argument or B < 1");
    if (p->n_sub < 1 || p->n_sub > 100000 || p->num_segments < 1 || offset < 0 || stride < offset + 6 || u_stride < 2)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_plant_step: size out of range");
    if (!(p->delta_t > 0.0) || !(p->lap_length > 0.0) || !(p->m > 0.0) || !(p->Iz > 0.0))
        return fail(h, B200MPC_ERR_ARG, "b200mpc_plant_step: bad parameter value");
    return B200MPC_OK;
}

int b200mpc_plant_step_device(b200mpc_handle *h, const b200mpc_plant_params *prm, int B, double *d_xcurv, int xcurv_stride,
                              int xcurv_offset, double *d_xglob, const double *d_u, int u_stride, const double *d_draws,
                              const double *d_segments, int32_t *d_laps) {
    int rc = check_plant(h, prm, B, d_xcurv, xcurv_stride, xcurv_offset, d_xglob, d_u, u_stride, d_segments);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    PlantKParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.p = *prm;
    kp.B = B;
    kp.xcurv_stride = xcurv_stride;
    kp.xcurv_offset = xcurv_offset;
    plant_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(kp, d_xcurv, d_xglob, d_u, u_stride, d_draws, d_segments, d_laps);
    CK(h, cudaGetLastError());
    h->launches++;
    return B200MPC_OK;
}

int b200mpc_plant_step(b200mpc_handle *h, const b200mpc_plant_params *prm, int B, double *xcurv, int xcurv_stride,
                       int xcurv_offset, double *xglob, const double *u, int u_stride, const double *draws,
                       const double *segments, int32_t *laps) {
    int rc = check_plant(h, prm, B, xcurv, xcurv_stride, xcurv_offset, xglob, u, u_stride, segments);
    if (rc) return rc;
    CK(h, cudaSetDevice(h->device));
    const size_t b_xc = (size_t)B * xcurv_stride * 8, b_xg = (size_t)B * 48, b_u = ((size_t)(B - 1) * u_stride + 2) * 8;
    const size_t b_d = (size_t)B * 24, b_seg = (size_t)prm->num_segments * 24, b_l = (size_t)B * 4;
    if ((rc = grow(h, &h->d_in, &h->c_in, b_xc))) return rc;
    if ((rc = grow(h, &h->d_x, &h->c_x, b_xg))) return rc;
    if ((rc = grow(h, &h->d_u, &h->c_u, b_u))) return rc;
    if (draws && (rc = grow(h, &h->d_aux, &h->c_aux, b_d))) return rc;
    if ((rc = grow(h, &h->d_seg, &h->c_seg, b_seg))) return rc;
    if (laps && (rc = grow(h, &h->d_stat, &h->c_stat, b_l))) return rc;
    CK(h, cudaMemcpyAsync(h->d_in, xcurv, b_xc, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_x, xglob, b_xg, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_u, u, b_u, cudaMemcpyHostToDevice, h->stream));
    if (draws) CK(h, cudaMemcpyAsync(h->d_aux, draws, b_d, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_seg, segments, b_seg, cudaMemcpyHostToDevice, h->stream));
    if (laps) CK(h, cudaMemcpyAsync(h->d_stat, laps, b_l, cudaMemcpyHostToDevice, h->stream));
    rc = b200mpc_plant_step_device(h, prm, B, (double *)h->d_in, xcurv_stride, xcurv_offset, (double *)h->d_x,
                                   (const double *)h->d_u, u_stride, draws ? (const double *)h->d_aux : nullptr,
                                   (const double *)h->d_seg, laps ? (int32_t *)h->d_stat : nullptr);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(xcurv, h->d_in, b_xc, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(xglob, h->d_x, b_xg, cudaMemcpyDeviceToHost, h->stream));
    if (laps) CK(h, cudaMemcpyAsync(laps, h->d_stat, b_l, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

int b200mpc_planner_select_device(b200mpc_handle *h, const b200mpc_planner_select_params *prm, const b200mpc_record *d_rec,
                                  const double *d_xpred, const double *d_heur, const int32_t *d_ok0, const int32_t *d_region,
                                  const double *d_rivals, double *d_sel_cost, int32_t *d_flag, double *d_traj,
                                  double *d_track_rec) {
    if (!h) return B200MPC_ERR_ARG;
    if (!prm || !d_rec || !d_xpred || !d_heur || !d_ok0 || !d_region || !d_flag)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_planner_select: null argument");
    if (prm->C < 1 || prm->N < 1 || prm->N > B200MPC_NMAX || prm->num_veh < 0 || (prm->num_veh > 0 && !d_rivals) ||
        (d_track_rec && (prm->N_ctrl < 1 || prm->N_ctrl > B200MPC_NMAX || prm->M_ctrl < 0 || prm->M_ctrl > B200MPC_MMAX)))
        return fail(h, B200MPC_ERR_ARG, "b200mpc_planner_select: bad parameter value");
    CK(h, cudaSetDevice(h->device));
    SelectKParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.p = *prm;
    kp.track_xt_off = d_track_rec ? cbf_hdr_doubles(prm->M_ctrl) : 0;
    planner_select_kernel<<<1, SELECT_NT, 0, h->stream>>>(kp, d_rec, d_xpred, d_heur, d_ok0, d_region, d_rivals, d_sel_cost, d_flag,
                                                         d_traj, d_track_rec);
    CK(h, cudaGetLastError());
    h->launches++;
    return B200MPC_OK;
}

int b200mpc_curv_to_glob_device(b200mpc_handle *h, int P, int num_segments, double lap_length, const double *d_track,
                                const double *d_s, int s_stride, const double *d_ey, int ey_stride, double *d_out) {
    if (!h) return B200MPC_ERR_ARG;
    if (!d_track || !d_s || !d_ey || !d_out || P < 1 || num_segments < 1 || s_stride < 1 || ey_stride < 1 || !(lap_length > 0.0))
        return fail(h, B200MPC_ERR_ARG, "b200mpc_curv_to_glob: bad arguments");
    CK(h, cudaSetDevice(h->device));
    curv_to_glob_kernel<<<(P + 127) / 128, 128, 0, h->stream>>>(P, num_segments, lap_length, d_track, d_s, s_stride, d_ey, ey_stride,
                                                              d_out);
    CK(h, cudaGetLastError());
    h->launches++;
    return B200MPC_OK;
}

int b200mpc_curv_to_glob(b200mpc_handle *h, int P, int num_segments, double lap_length, const double *track, const double *s,
                         int s_stride, const double *ey, int ey_stride, double *out) {
    if (!h) return B200MPC_ERR_ARG;
    if (!track || !s || !ey || !out || P < 1 || num_segments < 1 || s_stride < 1 || ey_stride < 1)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_curv_to_glob: bad arguments");
    CK(h, cudaSetDevice(h->device));
    const size_t n_s = (size_t)(P - 1) * s_stride + 1, n_e = (size_t)(P - 1) * ey_stride + 1;
    int rc;
    if ((rc = grow(h, &h->d_in, &h->c_in, 8 * (n_s + n_e)))) return rc;
    if ((rc = grow(h, &h->d_seg, &h->c_seg, 48 * (size_t)num_segments))) return rc;
    if ((rc = grow(h, &h->d_x, &h->c_x, 24 * (size_t)P))) return rc;
    double *din = (double *)h->d_in;
    CK(h, cudaMemcpyAsync(din, s, 8 * n_s, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(din + n_s, ey, 8 * n_e, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_seg, track, 48 * (size_t)num_segments, cudaMemcpyHostToDevice, h->stream));
    rc = b200mpc_curv_to_glob_device(h, P, num_segments, lap_length, (const double *)h->d_seg, din, s_stride, din + n_s, ey_stride,
                                     (double *)h->d_x);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(out, h->d_x, 24 * (size_t)P, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

int b200mpc_rival_rollout_device(b200mpc_handle *h, const b200mpc_rollout_params *prm, int B, const double *d_xcurv,
                                 const double *d_xglob, const double *d_segments, double *d_xcurv_n, double *d_xglob_n) {
    if (!h) return B200MPC_ERR_ARG;
    if (!prm || !d_xcurv || !d_xglob || !d_segments || !d_xcurv_n || B < 1)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_rival_rollout: null argument or B < 1");
    if (prm->n < 1 || prm->num_segments < 1 || !(prm->lap_length > 0.0) || !(prm->timestep > 0.0))
        return fail(h, B200MPC_ERR_ARG, "b200mpc_rival_rollout: bad parameter value");
    CK(h, cudaSetDevice(h->device));
    RolloutKParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.p = *prm;
    kp.B = B;
    rival_rollout_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(kp, d_xcurv, d_xglob, d_segments, d_xcurv_n, d_xglob_n);
    CK(h, cudaGetLastError());
    h->launches++;
    return B200MPC_OK;
}

int b200mpc_rival_rollout(b200mpc_handle *h, const b200mpc_rollout_params *prm, int B, const double *xcurv, const double *xglob,
                          const double *segments, double *xcurv_n, double *xglob_n) {
    if (!h) return B200MPC_ERR_ARG;
    if (!prm || !xcurv || !xglob || !segments || !xcurv_n || B < 1 || prm->n < 1 || prm->num_segments < 1)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_rival_rollout: null argument, B < 1 or bad parameter value");
    CK(h, cudaSetDevice(h->device));
    const size_t b_in = 48 * (size_t)B, b_seg = 24 * (size_t)prm->num_segments, b_out = 48 * (size_t)B * prm->n;
    int rc;
    if ((rc = grow(h, &h->d_in, &h->c_in, 2 * b_in))) return rc;
    if ((rc = grow(h, &h->d_seg, &h->c_seg, b_seg))) return rc;
    if ((rc = grow(h, &h->d_x, &h->c_x, b_out))) return rc;
    if (xglob_n && (rc = grow(h, &h->d_laps, &h->c_laps, b_out))) return rc;
    double *din = (double *)h->d_in;
    CK(h, cudaMemcpyAsync(din, xcurv, b_in, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(din + 6 * (size_t)B, xglob, b_in, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_seg, segments, b_seg, cudaMemcpyHostToDevice, h->stream));
    rc = b200mpc_rival_rollout_device(h, prm, B, din, din + 6 * (size_t)B, (const double *)h->d_seg, (double *)h->d_x,
                                      xglob_n ? (double *)h->d_laps : nullptr);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(xcurv_n, h->d_x, b_out, cudaMemcpyDeviceToHost, h->stream));
    if (xglob_n) CK(h, cudaMemcpyAsync(xglob_n, h->d_laps, b_out, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

static int check_prepare(b200mpc_handle *h, const b200mpc_planner_prepare_params *prm, const void *ego, const void *rivals,
                         const void *rival_vx, const void *insertion, const void *opt_traj) {
    if (!h) return B200MPC_ERR_ARG;
    if (!prm || !ego || !rivals || !rival_vx || !insertion || !opt_traj)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_planner_prepare: null argument");
    if (prm->N < 3 || prm->N > B200MPC_NMAX || prm->num_veh < 1 || prm->num_opt < 2 || !(prm->lap_length > 0.0) ||
        !(prm->w_track > 0.0) || !(prm->veh_width > 0.0) || !(prm->veh_length > 0.0))
        return fail(h, B200MPC_ERR_ARG, "b200mpc_planner_prepare: bad parameter value");
    return B200MPC_OK;
}

int b200mpc_planner_prepare_device(b200mpc_handle *h, const b200mpc_planner_prepare_params *prm, const double *d_ego,
                                   const double *d_rivals, const double *d_rival_vx, const int32_t *d_insertion,
                                   const double *d_opt_traj, double *d_cand, double *d_heur, int32_t *d_ok0, int32_t *d_region,
                                   double *d_offset, double *d_ctrl, double *d_bezier, int32_t *d_err) {
    int rc = check_prepare(h, prm, d_ego, d_rivals, d_rival_vx, d_insertion, d_opt_traj);
    if (rc) return rc;
    if (!d_cand || !d_heur || !d_ok0 || !d_region) return fail(h, B200MPC_ERR_ARG, "b200mpc_planner_prepare: null output");
    CK(h, cudaSetDevice(h->device));
    PrepareKParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.p = *prm;
    const int fl = B200MPC_FLAG_STAGE_BOUNDS | B200MPC_FLAG_EY_RATE;
    kp.stride = cbf_record_doubles(prm->N, 0, 1, fl);
    kp.xt_off = cbf_hdr_doubles(0);
    kp.bnd_off = cbf_base_doubles(prm->N, 0, 1);
    kp.wd_off = kp.bnd_off + 4 * (prm->N + 1);
    planner_prepare_kernel<<<1, PREPARE_NT, 0, h->stream>>>(kp, d_ego, d_rivals, d_rival_vx, d_insertion, d_opt_traj, d_cand, d_heur,
                                                           d_ok0, d_region, d_offset, d_ctrl, d_bezier, d_err);
    CK(h, cudaGetLastError());
    h->launches++;
    return B200MPC_OK;
}

// staging of the raw planner inputs: d_u = [ego 12][rival_vx nv][opt 2 num_opt][insertion nv ints], d_seg = rivals
struct PrepStage {
    const double *ego, *rival_vx, *opt;
    const int32_t *insertion;
};
static int stage_prepare_inputs(b200mpc_handle *h, const b200mpc_planner_prepare_params *prm, const double *ego,
                                const double *rivals, const double *rival_vx, const int32_t *insertion, const double *opt_traj,
                                PrepStage *st) {
    const int nv = prm->num_veh, N1 = prm->N + 1;
    const size_t n_d = 12 + (size_t)nv + 2 * (size_t)prm->num_opt;
    int rc;
    if ((rc = grow(h, &h->d_u, &h->c_u, 8 * n_d + 4 * (size_t)nv))) return rc;
    if ((rc = grow(h, &h->d_seg, &h->c_seg, 16 * (size_t)N1 * nv))) return rc;
    double *d = (double *)h->d_u;
    CK(h, cudaMemcpyAsync(d, ego, 96, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(d + 12, rival_vx, 8 * (size_t)nv, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(d + 12 + nv, opt_traj, 16 * (size_t)prm->num_opt, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(d + n_d, insertion, 4 * (size_t)nv, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_seg, rivals, 16 * (size_t)N1 * nv, cudaMemcpyHostToDevice, h->stream));
    st->ego = d;
    st->rival_vx = d + 12;
    st->opt = d + 12 + nv;
    st->insertion = (const int32_t *)(d + n_d);
    return B200MPC_OK;
}

int b200mpc_planner_prepare(b200mpc_handle *h, const b200mpc_planner_prepare_params *prm, const double *ego,
                            const double *rivals, const double *rival_vx, const int32_t *insertion, const double *opt_traj,
                            double *cand, double *heur, int32_t *ok0, int32_t *region, double *offset, double *ctrl,
                            double *bezier, int32_t *err) {
    int rc = check_prepare(h, prm, ego, rivals, rival_vx, insertion, opt_traj);
    if (rc) return rc;
    if (!cand || !heur || !ok0 || !region) return fail(h, B200MPC_ERR_ARG, "b200mpc_planner_prepare: null output");
    CK(h, cudaSetDevice(h->device));
    const int C_ = prm->num_veh + 1, N1 = prm->N + 1;
    const size_t cs = (size_t)cbf_record_doubles(prm->N, 0, 1, B200MPC_FLAG_STAGE_BOUNDS | B200MPC_FLAG_EY_RATE);
    const size_t b_in = cs * 8 * C_, b_x = 48 * (size_t)N1 * C_, b_int = 4 * (size_t)C_;
    // d_sig: [offset C][ctrl 8 C][bezier 2 N1 C][err]
    const size_t o_ctrl = (size_t)C_, o_bez = o_ctrl + 8 * (size_t)C_, o_err = o_bez + 2 * (size_t)N1 * C_;
    PrepStage st;
    if ((rc = stage_prepare_inputs(h, prm, ego, rivals, rival_vx, insertion, opt_traj, &st))) return rc;
    if ((rc = grow(h, &h->d_in, &h->c_in, b_in))) return rc;
    if ((rc = grow(h, &h->d_laps, &h->c_laps, b_x))) return rc;
    if ((rc = grow(h, &h->d_idx, &h->c_idx, b_int))) return rc;
    if ((rc = grow(h, &h->d_stat, &h->c_stat, b_int))) return rc;
    if ((rc = grow(h, &h->d_sig, &h->c_sig, 8 * (o_err + 1)))) return rc;
    double *sg = (double *)h->d_sig;
    rc = b200mpc_planner_prepare_device(h, prm, st.ego, (const double *)h->d_seg, st.rival_vx, st.insertion, st.opt, (double *)h->d_in,
                                        (double *)h->d_laps, (int32_t *)h->d_idx, (int32_t *)h->d_stat, sg, sg + o_ctrl, sg + o_bez,
                                        (int32_t *)(sg + o_err));
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(cand, h->d_in, b_in, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(heur, h->d_laps, b_x, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(ok0, h->d_idx, b_int, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(region, h->d_stat, b_int, cudaMemcpyDeviceToHost, h->stream));
    if (offset) CK(h, cudaMemcpyAsync(offset, sg, 8 * (size_t)C_, cudaMemcpyDeviceToHost, h->stream));
    if (ctrl) CK(h, cudaMemcpyAsync(ctrl, sg + o_ctrl, 64 * (size_t)C_, cudaMemcpyDeviceToHost, h->stream));
    if (bezier) CK(h, cudaMemcpyAsync(bezier, sg + o_bez, 16 * (size_t)N1 * C_, cudaMemcpyDeviceToHost, h->stream));
    if (err) CK(h, cudaMemcpyAsync(err, sg + o_err, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

// candidate solve -> selection -> tracking solve -> D2H on the handle's stream; the candidates (d_in), heuristic
// trajectories (d_laps), ok0 (d_idx), regions (d_stat), rivals (d_seg) and the tracking record (chain buffer) are in place
static int plan_chain(b200mpc_handle *h, const b200mpc_cbf_params *plan_prm, const b200mpc_cbf_params *track_prm,
                      const b200mpc_ipm_options *opt, const b200mpc_planner_select_params *sel, b200mpc_record *cand_rec,
                      double *cand_xpred, double *sel_cost, int32_t *flag, double *traj, b200mpc_record *track_rec,
                      double *track_xpred, double *track_upred) {
    const bool tracking = track_prm != nullptr;
    const int C_ = sel->C, N = plan_prm->N, Nc = tracking ? track_prm->N : 0, Mc = tracking ? track_prm->M : 0;
    const size_t ts = tracking ? (size_t)cbf_record_doubles(Nc, Mc, 1, 0) : 0;
    const size_t b_rec = sizeof(b200mpc_record) * (size_t)C_, b_x = 48 * (size_t)(N + 1) * C_;
    const size_t o_traj = 2, o_trk = o_traj + 6 * (size_t)(N + 1), o_out = o_trk + ts, o_tx = o_out + 4,
                 o_tu = o_tx + 6 * (size_t)(Nc + 1);
    double *ch = (double *)h->d_chain;
    // (1) all candidates, one launch
    int rc = b200mpc_cbf_solve_device(h, plan_prm, opt, C_, (const double *)h->d_in, (b200mpc_record *)h->d_rec, nullptr,
                                      (double *)h->d_x, nullptr, nullptr);
    if (rc) return rc;
    // (2) selection cost, first argmin, chosen trajectory, per-stage targets of the tracking record
    rc = b200mpc_planner_select_device(h, sel, (const b200mpc_record *)h->d_rec, (const double *)h->d_x, (const double *)h->d_laps,
                                       (const int32_t *)h->d_idx, (const int32_t *)h->d_stat, (const double *)h->d_seg,
                                       (double *)h->d_aux, (int32_t *)ch, ch + o_traj, tracking ? ch + o_trk : nullptr);
    if (rc) return rc;
    // (3) the tracking MPC on the record the selection kernel completed
    if (tracking) {
        rc = b200mpc_cbf_solve_device(h, track_prm, opt, 1, ch + o_trk, (b200mpc_record *)(ch + o_out), nullptr, ch + o_tx,
                                      ch + o_tu, nullptr);
        if (rc) return rc;
    }
    if (cand_rec) CK(h, cudaMemcpyAsync(cand_rec, h->d_rec, b_rec, cudaMemcpyDeviceToHost, h->stream));
    if (cand_xpred) CK(h, cudaMemcpyAsync(cand_xpred, h->d_x, b_x, cudaMemcpyDeviceToHost, h->stream));
    if (sel_cost) CK(h, cudaMemcpyAsync(sel_cost, h->d_aux, 8 * (size_t)C_, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(flag, ch, 8, cudaMemcpyDeviceToHost, h->stream));
    if (traj) CK(h, cudaMemcpyAsync(traj, ch + o_traj, 48 * (size_t)(N + 1), cudaMemcpyDeviceToHost, h->stream));
    if (tracking) {
        CK(h, cudaMemcpyAsync(track_rec, ch + o_out, sizeof(b200mpc_record), cudaMemcpyDeviceToHost, h->stream));
        if (track_xpred) CK(h, cudaMemcpyAsync(track_xpred, ch + o_tx, 48 * (size_t)(Nc + 1), cudaMemcpyDeviceToHost, h->stream));
        if (track_upred) CK(h, cudaMemcpyAsync(track_upred, ch + o_tu, 16 * (size_t)Nc, cudaMemcpyDeviceToHost, h->stream));
    }
    return B200MPC_OK;
}

static int plan_chain_buffers(b200mpc_handle *h, const b200mpc_cbf_params *plan_prm, const b200mpc_cbf_params *track_prm,
                              const b200mpc_planner_select_params *sel, const double *track_in) {
    const bool tracking = track_prm != nullptr;
    const int C_ = sel->C, N = plan_prm->N, Nc = tracking ? track_prm->N : 0, Mc = tracking ? track_prm->M : 0;
    const size_t cs = (size_t)cbf_record_doubles(N, plan_prm->M, plan_prm->xt_per_stage, plan_prm->flags);
    const size_t ts = tracking ? (size_t)cbf_record_doubles(Nc, Mc, 1, 0) : 0;
    const size_t b_in = cs * 8 * C_, b_rec = sizeof(b200mpc_record) * (size_t)C_, b_x = 48 * (size_t)(N + 1) * C_;
    const size_t b_riv = 16 * (size_t)(N + 1) * (sel->num_veh > 0 ? sel->num_veh : 1), b_int = 4 * (size_t)C_;
    // chain buffer: [flag 2 ints + pad][traj 6(N+1)][tracking record][tracking result record][x_pred][u_pred]
    const size_t o_traj = 2, o_trk = o_traj + 6 * (size_t)(N + 1), o_out = o_trk + ts, o_tx = o_out + 4,
                 o_tu = o_tx + 6 * (size_t)(Nc + 1), n_chain = o_tu + 2 * (size_t)Nc;
    int rc;
    if ((rc = grow(h, &h->d_in, &h->c_in, b_in))) return rc;
    if ((rc = grow(h, &h->d_rec, &h->c_rec, b_rec))) return rc;
    if ((rc = grow(h, &h->d_x, &h->c_x, b_x))) return rc;
    if ((rc = grow(h, &h->d_laps, &h->c_laps, b_x))) return rc;
    if ((rc = grow(h, &h->d_seg, &h->c_seg, b_riv))) return rc;
    if ((rc = grow(h, &h->d_idx, &h->c_idx, b_int))) return rc;
    if ((rc = grow(h, &h->d_stat, &h->c_stat, b_int))) return rc;
    if ((rc = grow(h, &h->d_aux, &h->c_aux, 8 * (size_t)C_))) return rc;
    if ((rc = grow(h, &h->d_chain, &h->c_chain, 8 * n_chain))) return rc;
    if (tracking) CK(h, cudaMemcpyAsync((double *)h->d_chain + o_trk, track_in, ts * 8, cudaMemcpyHostToDevice, h->stream));
    return B200MPC_OK;
}

static int check_plan(b200mpc_handle *h, const b200mpc_cbf_params *plan_prm, const b200mpc_cbf_params *track_prm,
                      const b200mpc_ipm_options *opt, const b200mpc_planner_select_params *sel, const void *track_in,
                      const void *flag, const void *track_rec, bool allow_plan_only) {
    if (!h) return B200MPC_ERR_ARG;
    if (!plan_prm || !opt || !sel || !flag) return fail(h, B200MPC_ERR_ARG, "b200mpc_plan_and_track: null argument");
    if (!track_prm) {   // planning only: no tracking record, no tracking outputs
        if (!allow_plan_only || track_in) return fail(h, B200MPC_ERR_ARG, "b200mpc_plan_and_track: null argument");
        if (sel->C < 1 || sel->N != plan_prm->N)
            return fail(h, B200MPC_ERR_ARG, "b200mpc_plan_and_track: inconsistent parameters");
        return check_cbf(h, plan_prm, opt, sel->C, flag, flag);
    }
    if (!track_in || !track_rec) return fail(h, B200MPC_ERR_ARG, "b200mpc_plan_and_track: null argument");
    if (sel->C < 1 || sel->N != plan_prm->N || sel->N_ctrl != track_prm->N || sel->M_ctrl != track_prm->M ||
        !track_prm->xt_per_stage || track_prm->flags != 0)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_plan_and_track: inconsistent parameters (the tracking record has per-stage targets)");
    return check_cbf(h, plan_prm, opt, sel->C, track_in, flag);
}

int b200mpc_plan_and_track(b200mpc_handle *h, const b200mpc_cbf_params *plan_prm, const b200mpc_cbf_params *track_prm,
                           const b200mpc_ipm_options *opt, const b200mpc_planner_select_params *sel, const double *cand_in,
                           const double *heur, const int32_t *ok0, const int32_t *region, const double *rivals,
                           const double *track_in, b200mpc_record *cand_rec, double *cand_xpred, double *sel_cost, int32_t *flag,
                           double *traj, b200mpc_record *track_rec, double *track_xpred, double *track_upred) {
    int rc = check_plan(h, plan_prm, track_prm, opt, sel, track_in, flag, track_rec, false);
    if (rc) return rc;
    if (!cand_in || !heur || !ok0 || !region || (sel->num_veh > 0 && !rivals))
        return fail(h, B200MPC_ERR_ARG, "b200mpc_plan_and_track: null argument");
    CK(h, cudaSetDevice(h->device));
    const int C_ = sel->C, N = plan_prm->N;
    const size_t cs = (size_t)cbf_record_doubles(N, plan_prm->M, plan_prm->xt_per_stage, plan_prm->flags);
    const size_t b_in = cs * 8 * C_, b_x = 48 * (size_t)(N + 1) * C_, b_int = 4 * (size_t)C_;
    const size_t b_riv = 16 * (size_t)(N + 1) * (sel->num_veh > 0 ? sel->num_veh : 1);
    if ((rc = plan_chain_buffers(h, plan_prm, track_prm, sel, track_in))) return rc;
    CK(h, cudaMemcpyAsync(h->d_in, cand_in, b_in, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_laps, heur, b_x, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_idx, ok0, b_int, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->d_stat, region, b_int, cudaMemcpyHostToDevice, h->stream));
    if (sel->num_veh > 0) CK(h, cudaMemcpyAsync(h->d_seg, rivals, b_riv, cudaMemcpyHostToDevice, h->stream));
    rc = plan_chain(h, plan_prm, track_prm, opt, sel, cand_rec, cand_xpred, sel_cost, flag, traj, track_rec, track_xpred, track_upred);
    if (rc) return rc;
    CK(h, cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

int b200mpc_plan_and_track_prepared(b200mpc_handle *h, const b200mpc_cbf_params *plan_prm, const b200mpc_cbf_params *track_prm,
                                    const b200mpc_ipm_options *opt, const b200mpc_planner_select_params *sel,
                                    const b200mpc_planner_prepare_params *prep, const double *ego, const double *rivals,
                                    const double *rival_vx, const int32_t *insertion, const double *opt_traj, int n_extra,
                                    const double *extra_cand, const double *extra_heur, const int32_t *extra_ok0,
                                    const int32_t *extra_region, const double *track_in, b200mpc_record *cand_rec,
                                    double *cand_xpred, double *sel_cost, int32_t *flag, double *traj, b200mpc_record *track_rec,
                                    double *track_xpred, double *track_upred, double *heur_out, int32_t *ok0_out, double *offset,
                                    double *bezier, int32_t *err) {
    int rc = check_plan(h, plan_prm, track_prm, opt, sel, track_in, flag, track_rec, true);
    if (rc) return rc;
    if ((rc = check_prepare(h, prep, ego, rivals, rival_vx, insertion, opt_traj))) return rc;
    const int fl = B200MPC_FLAG_STAGE_BOUNDS | B200MPC_FLAG_EY_RATE;
    const int C0 = prep->num_veh + 1, C_ = sel->C, N = plan_prm->N, N1 = N + 1;
    if (n_extra < 0 || C_ != C0 + n_extra || prep->N != N || sel->num_veh != prep->num_veh || plan_prm->M != 0 ||
        !plan_prm->xt_per_stage || plan_prm->flags != fl || (n_extra > 0 && (!extra_cand || !extra_heur || !extra_ok0 || !extra_region)))
        return fail(h, B200MPC_ERR_ARG, "b200mpc_plan_and_track_prepared: inconsistent parameters (C = num_veh + 1 + n_extra, "
                                        "planner records: M = 0, per-stage targets, STAGE_BOUNDS|EY_RATE)");
    CK(h, cudaSetDevice(h->device));
    const size_t cs = (size_t)cbf_record_doubles(N, 0, 1, fl);
    if ((rc = plan_chain_buffers(h, plan_prm, track_prm, sel, track_in))) return rc;
    PrepStage st;
    if ((rc = stage_prepare_inputs(h, prep, ego, rivals, rival_vx, insertion, opt_traj, &st))) return rc;
    // d_sig: [offset C0][bezier 2 N1 C0][err]
    const size_t o_bez = (size_t)C0, o_err = o_bez + 2 * (size_t)N1 * C0;
    if ((rc = grow(h, &h->d_sig, &h->c_sig, 8 * (o_err + 1)))) return rc;
    double *sg = (double *)h->d_sig;
    rc = b200mpc_planner_prepare_device(h, prep, st.ego, (const double *)h->d_seg, st.rival_vx, st.insertion, st.opt, (double *)h->d_in,
                                        (double *)h->d_laps, (int32_t *)h->d_idx, (int32_t *)h->d_stat, sg, nullptr, sg + o_bez,
                                        (int32_t *)(sg + o_err));
    if (rc) return rc;
    if (n_extra > 0) {   // additional candidates behind the reference's regions
        CK(h, cudaMemcpyAsync((double *)h->d_in + cs * C0, extra_cand, cs * 8 * n_extra, cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaMemcpyAsync((double *)h->d_laps + 6 * (size_t)N1 * C0, extra_heur, 48 * (size_t)N1 * n_extra, cudaMemcpyHostToDevice,
                              h->stream));
        CK(h, cudaMemcpyAsync((int32_t *)h->d_idx + C0, extra_ok0, 4 * (size_t)n_extra, cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaMemcpyAsync((int32_t *)h->d_stat + C0, extra_region, 4 * (size_t)n_extra, cudaMemcpyHostToDevice, h->stream));
    }
    rc = plan_chain(h, plan_prm, track_prm, opt, sel, cand_rec, cand_xpred, sel_cost, flag, traj, track_rec, track_xpred, track_upred);
    if (rc) return rc;
    if (heur_out) CK(h, cudaMemcpyAsync(heur_out, h->d_laps, 48 * (size_t)N1 * C_, cudaMemcpyDeviceToHost, h->stream));
    if (ok0_out) CK(h, cudaMemcpyAsync(ok0_out, h->d_idx, 4 * (size_t)C_, cudaMemcpyDeviceToHost, h->stream));
    if (offset) CK(h, cudaMemcpyAsync(offset, sg, 8 * (size_t)C0, cudaMemcpyDeviceToHost, h->stream));
    if (bezier) CK(h, cudaMemcpyAsync(bezier, sg + o_bez, 16 * (size_t)N1 * C0, cudaMemcpyDeviceToHost, h->stream));
    if (err) CK(h, cudaMemcpyAsync(err, sg + o_err, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return B200MPC_OK;
}

int b200mpc_argmin_cost_device(b200mpc_handle *h, const b200mpc_record *d_rec, int B, int max_status, int32_t *d_out) {
    if (!h) return B200MPC_ERR_ARG;
    if (!d_rec || !d_out || B < 1) return fail(h, B200MPC_ERR_ARG, "b200mpc_argmin_cost_device: bad arguments");
    CK(h, cudaSetDevice(h->device));
    argmin_cost_kernel<<<1, 256, 0, h->stream>>>(d_rec, B, max_status, d_out);
    CK(h, cudaGetLastError());
    h->launches++;
    return B200MPC_OK;
}


// ---- exchange window (include/b200mpc.h, csrc/exchange.cuh)
#ifdef B200MPC_HOST_EMULATION   // tests/host_emulation: "device" memory is host memory of this process, the handle is the pointer
static int ipc_export(void *base, void *out64) { memset(out64, 0, 64); memcpy(out64, &base, sizeof(base)); return 0; }
static int ipc_open(const void *in64, void **base) { memcpy(base, in64, sizeof(*base)); return 0; }
static void ipc_close(void *) {}
#else
static_assert(sizeof(cudaIpcMemHandle_t) <= B200MPC_COMM_HANDLE_BYTES, "IPC handle size");
static int ipc_export(void *base, void *out64) {
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, base);
    if (e != cudaSuccess) return (int)e;
    memset(out64, 0, 64);
    memcpy(out64, &hd, sizeof(hd));
    return 0;
}
static int ipc_open(const void *in64, void **base) {
    cudaIpcMemHandle_t hd;
    memcpy(&hd, in64, sizeof(hd));
    return (int)cudaIpcOpenMemHandle(base, hd, cudaIpcMemLazyEnablePeerAccess);
}
static void ipc_close(void *base) { cudaIpcCloseMemHandle(base); }
#endif

static size_t comm_rec_bytes(const b200mpc_comm *c) { return (size_t)c->slots * c->world * c->max_batch * sizeof(b200mpc_record); }
static size_t comm_ctr_bytes(const b200mpc_comm *c) { return (size_t)c->slots * c->world * sizeof(unsigned long long); }

int b200mpc_comm_create(b200mpc_handle *h, int rank, int world, int max_batch, int slots, b200mpc_comm **out) {
    if (!h) return B200MPC_ERR_ARG;
    if (!out || world < 1 || world > XCHG_MAX_WORLD || rank < 0 || rank >= world || max_batch < 1 || slots < 1 || slots > 64)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_comm_create: bad arguments");
    *out = nullptr;
    CK(h, cudaSetDevice(h->device));
    b200mpc_comm *c = new (std::nothrow) b200mpc_comm();
    if (!c) return fail(h, B200MPC_ERR_NOMEM, "out of host memory");
    c->device = h->device; c->rank = rank; c->world = world; c->max_batch = max_batch; c->slots = slots;
    c->window_bytes = comm_rec_bytes(c) + 2 * comm_ctr_bytes(c);
    cudaError_t e = cudaMalloc(&c->window, c->window_bytes);
    if (e == cudaSuccess) e = cudaMemset(c->window, 0, c->window_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_table, sizeof(XchgTable));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        std::string m = cudaGetErrorString(e);
        b200mpc_comm_destroy(c);
        return fail(h, B200MPC_ERR_NOMEM, "b200mpc_comm_create: " + m);
    }
    c->peer_base[rank] = c->window;
    *out = c;
    return B200MPC_OK;
}

int b200mpc_comm_export(b200mpc_comm *c, void *handle_out) {
    if (!c || !handle_out) return B200MPC_ERR_ARG;
    cudaSetDevice(c->device);
    return ipc_export(c->window, handle_out) == 0 ? B200MPC_OK : B200MPC_ERR_CUDA;
}

int b200mpc_comm_connect(b200mpc_comm *c, const void *handles) {
    if (!c || (!handles && c->world > 1)) return B200MPC_ERR_ARG;
    if (cudaSetDevice(c->device) != cudaSuccess) return B200MPC_ERR_CUDA;
    for (int p = 0; p < c->world; p++) {
        if (p == c->rank || c->peer_opened[p]) continue;
        void *base = nullptr;
        if (ipc_open((const char *)handles + (size_t)p * B200MPC_COMM_HANDLE_BYTES, &base) != 0 || !base) {
            g_create_err = "b200mpc_comm_connect: cannot map the window of rank " + std::to_string(p) +
                           " (CUDA IPC / peer access unavailable: " + cudaGetErrorString(cudaGetLastError()) + ")";
            return B200MPC_ERR_CUDA;
        }
        c->peer_base[p] = base;
        c->peer_opened[p] = true;
    }
    XchgTable t;
    memset(&t, 0, sizeof(t));
    t.rank = c->rank; t.world = c->world; t.max_batch = c->max_batch; t.slots = c->slots;
    for (int p = 0; p < c->world; p++) {
        char *b = (char *)c->peer_base[p];
        t.rec[p] = (b200mpc_record *)b;
        t.cnt[p] = (unsigned long long *)(b + comm_rec_bytes(c));
        t.ack[p] = (unsigned long long *)(b + comm_rec_bytes(c) + comm_ctr_bytes(c));
    }
    if (cudaMemcpy(c->d_table, &t, sizeof(t), cudaMemcpyHostToDevice) != cudaSuccess) return B200MPC_ERR_CUDA;
    c->connected = true;
    return B200MPC_OK;
}

void b200mpc_comm_destroy(b200mpc_comm *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < c->world; p++)
        if (c->peer_opened[p]) ipc_close(c->peer_base[p]);
    if (c->d_table) cudaFree(c->d_table);
    if (c->window) cudaFree(c->window);
    delete c;
}

int b200mpc_comm_publish_next(b200mpc_handle *h, b200mpc_comm *c, int slot) {
    if (!h) return B200MPC_ERR_ARG;
    if (!c || !c->connected || slot < 0 || slot >= c->slots || c->device != h->device)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_comm_publish_next: communicator not connected, slot out of range or wrong device");
    h->xchg_comm = c;
    h->xchg_slot = slot;
    return B200MPC_OK;
}

int b200mpc_comm_argmin(b200mpc_handle *h, b200mpc_comm *c, int slot, int max_status, int32_t *d_out, b200mpc_record *d_all) {
    if (!h) return B200MPC_ERR_ARG;
    if (!c || !c->connected || slot < 0 || slot >= c->slots || c->device != h->device)
        return fail(h, B200MPC_ERR_ARG, "b200mpc_comm_argmin: communicator not connected, slot out of range or wrong device");
    if (c->consumed[slot] >= c->published[slot])
        return fail(h, B200MPC_ERR_ARG, "b200mpc_comm_argmin: nothing published into this slot");
    CK(h, cudaSetDevice(h->device));
    const unsigned long long use = ++c->consumed[slot];
    const b200mpc_comm::Use u = c->ring[slot][use % b200mpc_comm::RING];
    xchg_argmin_kernel<<<1, 32, 0, h->stream>>>(c->d_table, slot, u.B, u.cum, use, max_status, d_out, d_all);
    CK(h, cudaGetLastError());
    h->launches++;
    return B200MPC_OK;
}

}  // extern "C"

// b200mpc: batched interior-point solver for the reference's MPC-LTI / MPC-CBF /
// mpc_multi_agents NLP (car_racing/control/control.py:198-248, 476-607, 251-473).
//
// One warp (= one CTA of 32 threads) solves one problem instance.  All solver state lives in
// shared memory (~31 KB at N=20, M=3) so 7 instances are resident per SM; HBM is touched
// twice: a TMA bulk copy (cp.async.bulk -> UBLKCP) of the packed 1136-byte instance record
// at the start and the coalesced result record at the end.  FP64 throughout; no tensor cores
// (KKT blocks are 6..14 wide and banded -- see DESIGN.md).
//
// Algorithm (DESIGN.md "Solver definition"; same definition as oracle/ocp_oracle.c, different
// linear algebra): primal-dual barrier method with IPOPT's conventions -- monotone mu,
// fraction-to-boundary, filter line search, inertia correction, gradient-based scaling,
// E_0 <= tol termination -- l1-elastic CBF rows instead of a restoration phase.  The Newton
// step is a Riccati recursion over the horizon on the augmented stage state
// X_k = (x_k, sigma_k), U_k = (u_k, sigma_{k+1}); a non-positive pivot in any stage's
// Cholesky is the inertia test.
//
// Execution style (measured: a lone warp pays ~4 issue cycles per shared-memory load and 2 per
// DFMA, so the design minimises LDS count and keeps the model out of shared memory):
//  * O(N) phases (residuals, derivatives, optimality error, assembly, line-search evaluation,
//    updates) run with LANE = STAGE: each lane does all the work of one stage as straight-line
//    code; the model matrices A, B, Q, R are compile-time-indexed kernel parameters, i.e.
//    constant-bank operands of the DFMAs, not loads.
//  * the backward Riccati sweep runs with LANE = COLUMN of the stage Hessian: P lives in
//    registers (lane a holds row a), phases exchange transposes through small shared buffers.
//  * the forward and costate sweeps carry the state redundantly in every lane's registers.
//  * CODE SIZE is a first-order performance parameter here: an SM that streams code past its 32 KB instruction
//    cache is capped near one warp-instruction per clock (tools/icache_probe.cu, profiles/r01t_icache_probe.txt), and
//    7 resident warps sit in 7 different phases.  Hence: every once-per-iteration phase has ONE inlined call site,
//    per-rival loops are rolled where the data is indexed in shared memory (no register arrays), sequential stage
//    loops are rolled, math-library calls (log) are funnelled through one site.  Keep it that way.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200mpc.h"
#include "exchange.cuh"

namespace b200mpc {

enum { OCP_FL_QDIAG = 8 };   // internal template flag (not a b200mpc_cbf_params flag): Q is diagonal, known at compile time

struct KParams {
    b200mpc_cbf_params p;
    b200mpc_ipm_options o;
    int32_t B;
    int32_t in_stride;  // doubles per instance record
    int32_t hdr;        // offset of xtarget inside the record
    int32_t obs_off;    // offset of the rival block inside the record
    int32_t bnd_off;    // offset of the per-stage bound block (flag STAGE_BOUNDS)
    int32_t wd_off;     // offset of the ey-rate weights (flag EY_RATE)
    int32_t sz_off;     // offset of the per-rival (L_j, W_j) block (flag RIVAL_SIZE)
    int32_t q_diag;     // Q is diagonal (every configuration of the reference: base.py:231,277,384): the O(N) phases take one
                        // product per state row instead of six -- same bits, the skipped terms are exact zeros
    double iL6, iW6;    // 1/L^6, 1/W^6
    double Q2[36];      // Q + Q' (the objective's Hessian block and gradient matrix), formed once on the host
};

__host__ __device__ inline int cbf_hdr_doubles(int M) { return (6 + M + 1) & ~1; }
__host__ __device__ inline int cbf_record_doubles(int N, int M, int xt_per_stage, int flags = 0) {
    int n = cbf_hdr_doubles(M) + (xt_per_stage ? 6 * (N + 1) : 6) + 2 * M * (N + 1);
    n = (n + 1) & ~1;
    if (flags & B200MPC_FLAG_STAGE_BOUNDS) n += 4 * (N + 1);
    if (flags & B200MPC_FLAG_EY_RATE) n += (N + 1) & ~1;
    if (flags & B200MPC_FLAG_RIVAL_SIZE) n += 2 * M;
    return n;
}
__host__ __device__ inline int cbf_base_doubles(int N, int M, int xt_per_stage) {
    return (cbf_hdr_doubles(M) + (xt_per_stage ? 6 * (N + 1) : 6) + 2 * M * (N + 1) + 1) & ~1;
}

// ---------------------------------------------------------------- shared memory plan
template <int M>
struct SmemPlan {
    static constexpr int NXA = 6 + M, NUA = 2 + M, NZ = NXA + NUA;
    static constexpr int NXAP = (NXA + 1) & ~1, NUAP = (NUA + 1) & ~1, NC = 8 + M;
    // row of the feedback law in KFB: NXA gains, then the feed-forward term, padded to an even length
    static constexpr int NKP = (NXA + 2) & ~1;
    int N, R, NB, NW, OU, OS;
    int oIN, oW, oD, oHD, oZL, oZU, oS, oT, oY, oZ, oV, oDG, oG, oSIGE, oYHAT, oJD, oJA, oLAM, oCRES, oKFB, oPT, oQV,
        oGUU, oGVU, oYF, oYG, oS0, oJDC, oAB, total;
    __host__ __device__ SmemPlan(int N_, int in_stride, int flags) {
        N = N_;
        R = M * N;
        NB = 4 * N + M * (N + 1);
        OU = 6 * (N + 1);
        OS = OU + 2 * N;
        NW = OS + M * (N + 1);
        int o = 2;  // doubles 0..1: mbarrier (8 B) + pad, keeps everything after 16-byte aligned
        auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
        oIN = take(in_stride);
        oW = take(NW + 2); oD = take(NW + 2); oHD = take(NW + 2);
        oZL = take(NB); oZU = take(4 * N);
        oS = take(R); oT = take(R); oY = take(R); oZ = take(R); oV = take(R);
        oDG = take(R); oG = take(R); oSIGE = take(R); oYHAT = take(R); oJD = take(R);
        oJA = take(4 * R);
        oLAM = take(6 * N + 6); oCRES = take(6 * N + 6);
        oKFB = take(N * NUA * NKP);
        oPT = take((NC + 1) * NXAP); oQV = take(NXAP);
        oGUU = take(NUA * NUAP); oGVU = take(NUAP); oYF = take(NXA * NUAP); oYG = take(NUAP);
        oS0 = take((M > 0 ? M : 1) * (M + 2));
        oJDC = take((flags & B200MPC_FLAG_EY_RATE) ? N + 2 : 0);   // ey-rate differences of the direction (planner QP only)
        oAB = take(48);   // rows of [A | B] for the Riccati sweep (16-byte loads instead of constant-bank loads)
        total = o;
    }
    __host__ __device__ size_t bytes() const { return (size_t)total * sizeof(double); }
};

// The once-per-iteration phases keep their loops over the six state rows (OCP_ROLL_A) and over the rivals (OCP_ROLL_J) rolled
// for code size (instruction cache); the unroll factors are build switches so that the trade against the
// dependent-chain latency of a rolled iteration can be measured (tools/variants.sh; profiles/r06_variants_3.txt: x2 / x3 of the
// row loops +0.7 % / -0.2 %, x3 of the rival loops -12 % on a full SM; r06_variants_4.txt: x2 of the row loops +2 %, the default)
#ifndef B200MPC_UNROLL_A
#define B200MPC_UNROLL_A 2
#endif
#ifndef B200MPC_UNROLL_J
#define B200MPC_UNROLL_J 1
#endif
#define OCP_STR2_(x) #x
#define OCP_STR_(x) OCP_STR2_(x)
#define OCP_ROLL_A _Pragma(OCP_STR_(unroll B200MPC_UNROLL_A))
#define OCP_ROLL_J _Pragma(OCP_STR_(unroll B200MPC_UNROLL_J))

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// max / min over the warp of NON-NEGATIVE doubles (error norms, step lengths): for those the IEEE bit pattern orders
// like the value, so two hardware integer reductions (REDUX: high word, then the low words of the lanes that hold the
// winning high word) replace five shuffle + DSETP/FSEL/SEL rounds.  A NaN anywhere wins the max (the caller's
// comparisons then fail and the solve ends at max_iter), where fmax would have dropped it silently.
__device__ __forceinline__ double warp_max_nn(double v) {
    unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ double warp_min_nn(double v) {
    unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return __hiloint2double((int)mh, (int)ml);
}
// branch-free FP64 reciprocal / reciprocal square root: hardware seed (MUFU.RCP64H / RSQ64H, relative error < 1e-6 measured
// over 60 binades, profiles/r01k_rsq_accuracy.txt) + ONE third-order step: r (1 + e + e^2) with e = 1 - x r, and
// y (1 + e/2 + 3 e^2/8) with e = 1 - x y^2 -- error e^3 ~ 1e-18, and a dependent chain of 3 / 4 instructions where two
// Newton steps have 4 / 6 (+4.7 % solves/s, profiles/r06_variants_1.txt).
// The compiler's IEEE division carries a slow-path call per site (~20 instructions, 2 branches); the solver's operands
// are positive, finite and far from the denormal range, and 1-ulp differences are irrelevant here.
// tests/host_emulation (the kernel source compiled by g++) has no hardware seed: it takes the exact value off by 9e-7,
// so that the correction step is what the host tests check.
__device__ __forceinline__ double rcp(double x) {
    double r;
#ifdef B200MPC_HOST_EMULATION
    r = (1.0 / x) * (1.0 + 9e-7);
#else
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#endif
    double e = fma(-x, r, 1.0);
    return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double rsq(double x) {
    double y;
#ifdef B200MPC_HOST_EMULATION
    y = (1.0 / sqrt(x)) * (1.0 - 9e-7);
#else
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#endif
    double e = fma(-(x * y), y, 1.0);
    return fma(y * e, fma(e, 0.375, 0.5), y);
}
__device__ __forceinline__ double p4(double a) { double b = a * a; return b * b; }
__device__ __forceinline__ double p5(double a) { double b = a * a; return b * b * a; }
__device__ __forceinline__ double p6(double a) { double b = a * a; return b * b * b; }

// sum of logs of positive numbers without one log() per term: multiply the values (raw() -- a product of a few
// terms cannot leave the double range), pull the exponent out of the running product once per group (norm(): integer
// ops on the exponent field; the mantissa bits are those of the un-normalised product, so the result equals the
// term-by-term frexp version bit for bit)
struct LogAcc {
    double m = 1.0;
    int e = 0;
    __device__ __forceinline__ void raw(double v) { m *= v; }
    __device__ __forceinline__ void norm() {   // m is positive and normal here
        int hi = __double2hiint(m);
        e += (hi >> 20) - 1022;
        m = __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(m));
    }
    __device__ __forceinline__ void mul(double v) { raw(v); norm(); }
    __device__ __forceinline__ double value() const { return log(m) + (double)e * 0.693147180559945309417232; }
};

// 16-byte shared-memory accesses (all array bases and the offsets used are even)
__device__ __forceinline__ void ld6(const double *p, double (&v)[6]) {
    double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2),
            c = *reinterpret_cast<const double2 *>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y;
}
__device__ __forceinline__ void st6(double *p, const double (&v)[6]) {
    *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]);
    *reinterpret_cast<double2 *>(p + 4) = make_double2(v[4], v[5]);
}
template <int K>
__device__ __forceinline__ void ldv(const double *p, double (&v)[K]) {  // K values, storage padded to even
#pragma unroll
    for (int i = 0; i + 1 < K; i += 2) {
        double2 a = *reinterpret_cast<const double2 *>(p + i);
        v[i] = a.x;
        v[i + 1] = a.y;
    }
    if (K & 1) v[K - 1] = p[K - 1];
}
template <int K>
__device__ __forceinline__ void stv(double *p, const double (&v)[K]) {
#pragma unroll
    for (int i = 0; i + 1 < K; i += 2) *reinterpret_cast<double2 *>(p + i) = make_double2(v[i], v[i + 1]);
    if (K & 1) p[K - 1] = v[K - 1];
}
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }

// ---------------------------------------------------------------- TMA (bulk async copy) + mbarrier
#ifdef B200MPC_HOST_EMULATION
// tests/host_emulation: the issuing lane copies synchronously; the __syncwarp() that follows mbar_wait() publishes the data
__device__ __forceinline__ void mbar_init(uint64_t *, int) {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *, uint32_t) {}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t *, uint32_t) {}
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
#endif

// ---------------------------------------------------------------- the solver
// NT > 0: horizon known at compile time (all shared-memory offsets become immediates); NT == 0: runtime horizon
template <int M, int FL, int NT>
struct Ipm {
    static constexpr int NXA = 6 + M, NUA = 2 + M, NZ = NXA + NUA;
    static constexpr int NXAP = (NXA + 1) & ~1, NUAP = (NUA + 1) & ~1, NC = 8 + M, MM = (M > 0 ? M : 1);
    static constexpr int NKP = SmemPlan<M>::NKP;
    const KParams &kp;
    const int lane;
    // role lane of the backward sweep: when the sweep's NZ + 1 lanes fit a half-warp (M <= 3), lanes 16..31 mirror lanes 0..15
    // (same column, same arithmetic: free in SIMT) and take the second half of the columns of [A B] in step (1)
#if !defined(B200MPC_NO_MIRROR)
    static constexpr bool kMirror = (NZ + 1 <= 16);
#else
    static constexpr bool kMirror = false;
#endif
    const int rl;
    static constexpr bool kStaticN = NT > 0;
    // stage loops: with a compile-time horizon below 32 every stage has its own lane and the loop runs once -- a step
    // that overshoots any horizon lets the compiler drop the loop (back edge, loop-carried moves)
    static constexpr int KSTEP = (NT > 0 && NT < 32) ? (1 << 20) : 32;
    const int N, R, NB, NW, OU, OS;
    double *IN, *W, *D, *HD, *ZL, *ZU, *S, *T, *Y, *Z, *V, *DG, *GR, *SIGE, *YHAT, *JD, *JA, *LAM, *CRES, *KFB, *PT,
        *QVs, *GUU, *GVU, *YF, *YG, *S0, *JDC, *ABs;
    const double *xt, *obs, *lapoff, *bnd, *wdp, *szp;
    // per-stage bounds / ey-rate cost present (planner QP): compile-time so that the MPC-CBF path pays nothing
    static constexpr bool psb = (FL & B200MPC_FLAG_STAGE_BOUNDS) != 0, hwd = (FL & B200MPC_FLAG_EY_RATE) != 0;
    // per-rival sizes (control.py:530-535 reads length / width of every rival): the record carries (L_j, W_j), the kernel
    // turns them into 1/L_j^6, 1/W_j^6 in place once after staging; without the flag they are kernel-parameter constants
    static constexpr bool prs = (FL & B200MPC_FLAG_RIVAL_SIZE) != 0;
    // diagonal Q: instantiations with the internal bit OCP_FL_QDIAG know it at compile time and carry no general path at all
    // (the north star: its hot loop body must stay small, the second path cost 5 points of instruction-fetch stalls); the
    // others ask the kernel parameter
    static constexpr bool kQD = (FL & OCP_FL_QDIAG) != 0;
    __device__ __forceinline__ bool qdiag() const { return kQD || kp.q_diag != 0; }
    int nb_count;        // number of bound + row multipliers (for the error scaling)
    double df, mu, rho, a1;  // a1 = 1 - alpha
    // lane = column role of the Riccati sweep (set once)
    bool isx, iss, isu, isp;
    int jrole, cmap;
    int idx_base, idx_step, kmin;   // lane's diagonal / gradient slot at stage k: idx_base + k*idx_step, live for k >= kmin
    double e4, e5;                  // one-hot of lanes 4 (s) and 5 (ey) as doubles
    double crow[MM], psel[MM];      // d(row j)/d(own sigma variable): (1-alpha) for sigma_k, -1 for sigma_{k+1}; one-hot of sigma_{k+1}
    double tcol[6], qqcol[6], rrcol[2];
    double arow[6], brow[2];   // row `lane` of A and B (lanes < 6), for the forward sweep

    __device__ Ipm(const KParams &kp_, const SmemPlan<M> &pl, double *sm, int lane_)
        : kp(kp_), lane(lane_), rl(kMirror ? (lane_ & 15) : lane_), N(NT ? NT : pl.N), R(M * (NT ? NT : pl.N)), NB(NT ? 4 * NT + M * (NT + 1) : pl.NB),
          NW(NT ? 8 * NT + 6 + M * (NT + 1) : pl.NW), OU(NT ? 6 * (NT + 1) : pl.OU), OS(NT ? 8 * NT + 6 : pl.OS) {
        IN = sm + pl.oIN; W = sm + pl.oW; D = sm + pl.oD; HD = sm + pl.oHD; ZL = sm + pl.oZL; ZU = sm + pl.oZU;
        S = sm + pl.oS; T = sm + pl.oT; Y = sm + pl.oY; Z = sm + pl.oZ; V = sm + pl.oV;
        DG = sm + pl.oDG; GR = sm + pl.oG; SIGE = sm + pl.oSIGE; YHAT = sm + pl.oYHAT; JD = sm + pl.oJD;
        JA = sm + pl.oJA; LAM = sm + pl.oLAM; CRES = sm + pl.oCRES; KFB = sm + pl.oKFB;
        PT = sm + pl.oPT; QVs = sm + pl.oQV; GUU = sm + pl.oGUU; GVU = sm + pl.oGVU; YF = sm + pl.oYF; YG = sm + pl.oYG;
        S0 = sm + pl.oS0;
        JDC = sm + pl.oJDC;
        ABs = sm + pl.oAB;
        bnd = IN + kp.bnd_off;
        wdp = IN + kp.wd_off;
        szp = IN + kp.sz_off;
        nb_count = 0;
        lapoff = IN + 6;
        xt = IN + kp.hdr;
        obs = IN + kp.obs_off;
        a1 = 1.0 - kp.p.alpha;
        rho = kp.o.rho;
        df = 1.0;
        mu = kp.o.mu_init;
        const int l = rl;
        isx = l < 6;
        iss = l >= 6 && l < NXA;
        isu = l >= NXA && l < NXA + 2;
        isp = l >= NXA + 2 && l < NZ;
        jrole = iss ? l - 6 : (isp ? l - NXA - 2 : (isu ? l - NXA : 0));
        cmap = isx ? l : (isu ? 6 + (l - NXA) : (isp ? 8 + (l - NXA - 2) : NC));  // NC = the all-zero row of PT
        idx_base = isx ? l : (iss ? OS + jrole * (N + 1) : (isu ? OU + jrole : 0));
        idx_step = isx ? 6 : (iss ? 1 : (isu ? 2 : 0));
        kmin = isx ? 1 : ((iss || isu) ? 0 : (1 << 30));
        e4 = (l == 4) ? 1.0 : 0.0;
        e5 = (l == 5) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < MM; j++) {
            crow[j] = (iss && jrole == j) ? a1 : ((isp && jrole == j) ? -1.0 : 0.0);
            psel[j] = (isp && jrole == j) ? 1.0 : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 6; q++) {
            double va = 0.0, vb = 0.0;
#pragma unroll
            for (int c = 0; c < 6; c++) va = (l == c) ? kp.p.A[6 * q + c] : va;
#pragma unroll
            for (int c = 0; c < 2; c++) vb = (l == NXA + c) ? kp.p.B[2 * q + c] : vb;
            tcol[q] = isx ? va : vb;
            qqcol[q] = 0.0;
        }
        rrcol[0] = rrcol[1] = 0.0;
#pragma unroll
        for (int b = 0; b < 6; b++) {
            double v = 0.0;
#pragma unroll
            for (int a = 0; a < 6; a++) v = (l == a) ? kp.p.A[6 * a + b] : v;
            arow[b] = v;
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            double v = 0.0;
#pragma unroll
            for (int a = 0; a < 6; a++) v = (l == a) ? kp.p.B[2 * a + c] : v;
            brow[c] = v;
        }
    }
    // after the objective scaling df is known
    __device__ void set_scaled_columns() {
#pragma unroll
        for (int q = 0; q < 6; q++) {
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < 6; c++) v = (rl == c) ? Q2(q, c) : v;
            qqcol[q] = df * v;
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            double v = 0.0;
#pragma unroll
            for (int c = 0; c < 2; c++) v = (rl == NXA + c) ? R2(q, c) : v;
            rrcol[q] = df * v;
        }
    }

    // model constants with compile-time indices -> constant-bank operands
    __device__ __forceinline__ double Am(int a, int b) const { return kp.p.A[6 * a + b]; }
    __device__ __forceinline__ double Bm(int a, int c) const { return kp.p.B[2 * a + c]; }
    __device__ __forceinline__ double Q2(int a, int b) const { return kp.p.Q[6 * a + b] + kp.p.Q[6 * b + a]; }
    __device__ __forceinline__ double R2(int a, int b) const { return kp.p.R[2 * a + b] + kp.p.R[2 * b + a]; }

    // ---- bounds on (vx_i, ey_i): c = 0 -> vx, c = 1 -> ey.  |bound| >= 1e299 (or inf) means "none".
    __device__ __forceinline__ double xlb(int i, int c) const { return psb ? bnd[4 * i + c] : (c ? -kp.p.width : kp.p.vmin); }
    __device__ __forceinline__ double xub(int i, int c) const { return psb ? bnd[4 * i + 2 + c] : (c ? kp.p.width : kp.p.vmax); }
    static __device__ __forceinline__ bool has(double b) { return !psb || fabs(b) < 1e299; }
    __device__ __forceinline__ double wdk(int k) const { return hwd ? wdp[k] : 0.0; }
    // barrier gradient terms -mu/(x-l) + mu/(u-x) of the bounded components of x_i, added to g[0], g[5]
    __device__ __forceinline__ void barrier_grad_x(int i, const double (&x)[6], double (&g)[6]) const {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            double lb = xlb(i, c), ub = xub(i, c), w = c ? x[5] : x[0], t = 0.0;
            if (has(lb)) t -= mu * rcp(w - lb);
            if (has(ub)) t += mu * rcp(ub - w);
            g[c ? 5 : 0] += t;
        }
    }

    // ---- indexing
    __device__ __forceinline__ const double *xtp(int i) const { return kp.p.xt_per_stage ? xt + 6 * i : xt; }
    __device__ __forceinline__ double obs_s(int j, int i) const { return obs[(2 * j) * (N + 1) + i]; }
    __device__ __forceinline__ double obs_e(int j, int i) const { return obs[(2 * j + 1) * (N + 1) + i]; }
    __device__ __forceinline__ int isg(int j, int i) const { return OS + j * (N + 1) + i; }
    // bound-multiplier slots: x_i (i>=1): 2(i-1) (vx), 2(i-1)+1 (ey); u_k: 2N+2k, 2N+2k+1; sigma_{j,i}: 4N + j(N+1)+i
    __device__ __forceinline__ int bsx(int i) const { return 2 * (i - 1); }
    __device__ __forceinline__ int bsu(int k) const { return 2 * N + 2 * k; }
    __device__ __forceinline__ int bss(int j, int i) const { return 4 * N + j * (N + 1) + i; }

    // ---- stage evaluation helpers (lane = stage)
    struct RowV { double ds, de, dsn, den, sg, sgn, iL, iW; };
    __device__ __forceinline__ double iL6(int j) const { return prs ? szp[2 * j] : kp.iL6; }
    __device__ __forceinline__ double iW6(int j) const { return prs ? szp[2 * j + 1] : kp.iW6; }
    __device__ __forceinline__ RowV row_vals(int j, int i, const double (&x)[6], const double (&xn)[6], double sg, double sgn) const {
        RowV v;
        v.sg = sg;
        v.sgn = sgn;
        v.ds = x[4] - obs_s(j, i) - lapoff[j];   // control.py:539-540 (with lap offset)
        v.de = x[5] - obs_e(j, i);
        v.dsn = xn[4] - obs_s(j, i + 1);         // control.py:542 (quirk: no lap offset)
        v.den = xn[5] - obs_e(j, i + 1);
        v.iL = iL6(j);
        v.iW = iW6(j);
        return v;
    }
    __device__ __forceinline__ double row_g(const RowV &v) const {  // unscaled h_next - (1-alpha) h  (control.py:558)
        double h = p6(v.ds) * v.iL + p6(v.de) * v.iW - 1.0 - kp.p.margin - v.sg;
        double hn = p6(v.dsn) * v.iL + p6(v.den) * v.iW - 1.0 - kp.p.margin - v.sgn;
        return hn - a1 * h;
    }
    // x_k, u_k at W + al*D
    template <bool useD>
    __device__ __forceinline__ void load_x(int k, double al, double (&x)[6]) const {
        ld6(W + 6 * k, x);
        if (useD) {
            double d[6];
            ld6(D + 6 * k, d);
#pragma unroll
            for (int a = 0; a < 6; a++) x[a] += al * d[a];
        }
    }
    template <bool useD>
    __device__ __forceinline__ void load_u(int k, double al, double (&u)[2]) const {
        double2 a = ld2(W + OU + 2 * k);
        u[0] = a.x; u[1] = a.y;
        if (useD) {
            double2 d = ld2(D + OU + 2 * k);
            u[0] += al * d.x; u[1] += al * d.y;
        }
    }
    // The once-per-iteration phases keep their matrix loops ROLLED (#pragma unroll 1 over the output row):
    // fully unrolled they were ~20k instructions of straight-line code and the instruction cache, not the
    // arithmetic, set their speed (ncu: stall_no_inst 43 % of samples).  The row index is uniform across the
    // warp, so A, B, Q are read with register-indexed constant-bank loads.
    // c_k = x_{k+1} - A x_k - B u_k ; x_{k+1} is read from shared memory (+al*D); returns sum_a |c_a|, optionally stores c
    template <bool useD, bool store>
    __device__ __forceinline__ double dyn_res(int k, double al, const double (&x)[6], const double (&u)[2]) const {
        double acc = 0.0;
OCP_ROLL_A
        for (int a = 0; a < 6; a++) {
            double s = W[6 * (k + 1) + a];
            if (useD) s += al * D[6 * (k + 1) + a];
            const double *Ar = kp.p.A + 6 * a;
#pragma unroll
            for (int b = 0; b < 6; b++) s -= Ar[b] * x[b];
            s -= kp.p.B[2 * a] * u[0] + kp.p.B[2 * a + 1] * u[1];
            if (store) CRES[6 * k + a] = s;
            acc += fabs(s);
        }
        return acc;
    }
    // (x-xt)'Q(x-xt)
    __device__ __forceinline__ double stage_cost(int i, const double (&x)[6]) const {
        const double *t = xtp(i);
        double d[6];
#pragma unroll
        for (int a = 0; a < 6; a++) d[a] = x[a] - t[a];
        double f = 0.0;
        if (qdiag()) {   // warp-uniform
#pragma unroll
            for (int a = 0; a < 6; a++) f += d[a] * (kp.p.Q[7 * a] * d[a]);
            return f;
        }
OCP_ROLL_A
        for (int a = 0; a < 6; a++) {
            const double *Qr = kp.p.Q + 6 * a;
            double acc = 0.0, da = 0.0;
#pragma unroll
            for (int b = 0; b < 6; b++) {
                acc += Qr[b] * d[b];
                da = (b == a) ? d[b] : da;
            }
            f += da * acc;
        }
        return f;
    }
    __device__ __forceinline__ double input_cost(const double (&u)[2]) const {
        return u[0] * (kp.p.R[0] * u[0] + kp.p.R[1] * u[1]) + u[1] * (kp.p.R[2] * u[0] + kp.p.R[3] * u[1]);
    }
    // Component a of the scaled objective gradient wrt x_i; d = x_i - xt_i.  Called from rolled loops with a warp-uniform
    // runtime a (register-indexed constant-bank loads of Q + Q').  The gradient is recomputed where it is needed (error,
    // assembly, costate residual: 36 FMAs per stage each) instead of being cached in shared memory -- those 126 doubles,
    // with the feed-forward terms folded into KFB's padding, are what lets an 8th CTA fit an SM.
    __device__ __forceinline__ double grad_x_comp(int i, int a, const double (&d)[6], double ey) const {
        const double *q2 = kp.Q2 + 6 * a;
        double acc = 0.0;
        if (qdiag()) {   // warp-uniform; d_a re-formed from shared memory (a is a runtime index in the callers' rolled loops)
            acc = q2[a] * (W[6 * i + a] - xtp(i)[a]);
        } else {
#pragma unroll
            for (int b = 0; b < 6; b++) acc += q2[b] * d[b];
        }
        acc *= df;
        if (hwd && a == 5) {   // d/d ey_i of  wd_{i-1}(ey_i-ey_{i-1})^2 + wd_i(ey_{i+1}-ey_i)^2
            double tt = wdp[i - 1] * (ey - W[6 * (i - 1) + 5]);
            if (i < N) tt -= wdp[i] * (W[6 * (i + 1) + 5] - ey);
            acc += 2.0 * df * tt;
        }
        return acc;
    }
    __device__ __forceinline__ void diff_target(int i, const double (&x)[6], double (&d)[6]) const {
        const double *t = xtp(i);
#pragma unroll
        for (int a = 0; a < 6; a++) d[a] = x[a] - t[a];
    }

    // ---- Newton steps of the row slacks / multipliers from JD (all at the current iterate)
    __device__ __forceinline__ void row_step(int r, double &ds, double &dt, double &dy, double &dz, double &dv) const {
        double s = S[r], t = T[r], z = Z[r], v = V[r];
        double is = rcp(s), it = rcp(t);
        double sig_s = z * is, sig_t = v * it;
        double rg = GR[r] + t - s;
        double jd = JD[r];
        dy = YHAT[r] - SIGE[r] * jd - Y[r];
        dt = (mu * is + mu * it - rho - sig_s * rg - sig_s * jd) * rcp(sig_s + sig_t);
        dv = mu * it - v - sig_t * dt;
        ds = jd + dt + rg;
        dz = mu * is - z - sig_s * ds;
    }

    // ---- constraint violation theta (1-norm) and barrier function phi at W + al*D, slacks moved by al
    template <bool useD>
    __device__ void theta_phi(double al, double &theta, double &phi) const {
        double th = 0.0, f = 0.0, tsum = 0.0, ss = 0.0;
        LogAcc la;
        for (int k = lane; k <= N; k += KSTEP) {
            double x[6];
            load_x<useD>(k, al, x);
            f += stage_cost(k, x);                                   // control.py:588-591
            if (k >= 1) {                                            // bounds on vx_k, ey_k (:582-586)
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    double lb = xlb(k, c), ub = xub(k, c), w = c ? x[5] : x[0];
                    if (has(lb)) la.raw(w - lb);
                    if (has(ub)) la.raw(ub - w);
                }
                la.norm();
            }
            const bool kn = k < N;
            double xn[6], u[2];
            if (kn) {
                load_x<useD>(k + 1, al, xn);
                load_u<useD>(k, al, u);
                f += input_cost(u);                                  // :578-579
                la.raw(u[0] + kp.p.umax[0]); la.raw(kp.p.umax[0] - u[0]);
                la.raw(u[1] + kp.p.umax[1]); la.raw(kp.p.umax[1] - u[1]);
                la.norm();
                th += dyn_res<useD, false>(k, al, x, u);
                if (hwd) f += wdp[k] * (xn[5] - x[5]) * (xn[5] - x[5]);   // overtake_traj_planner.py:325-327
            }
            // one rolled loop over the rivals (the kernel is instruction-fetch bound: code size counts)
OCP_ROLL_J
            for (int j = 0; j < M; j++) {
                double sv = W[isg(j, k)];
                if (useD) sv += al * D[isg(j, k)];
                ss += sv;
                la.raw(sv);                                          // sigma >= 0 (:559,561)
                if (kn) {
                    int r = j * N + k;
                    double sgn = W[isg(j, k + 1)];
                    if (useD) sgn += al * D[isg(j, k + 1)];
                    double g = DG[r] * row_g(row_vals(j, k, x, xn, sv, sgn));
                    double s = S[r], tt = T[r];
                    if (useD) {
                        double ds, dt, dy, dz, dv;
                        row_step(r, ds, dt, dy, dz, dv);
                        s += al * ds;
                        tt += al * dt;
                    }
                    th += fabs(g + tt - s);
                    la.raw(s);
                    la.raw(tt);
                    tsum += tt;
                }
                la.norm();
            }
        }
        f += kp.p.slack_w * ss;
        theta = warp_sum(th);
        phi = warp_sum(df * f + rho * tsum - mu * la.value());   // lane-local partial of phi, one reduction
    }

    // constraint violation at the start point: the rows start with s = g + t, i.e. zero residual, so only the
    // dynamics count (the bound push may have moved the rolled-out states)
    __device__ double theta0() const {
        double th = 0.0;
        for (int k = lane; k < N; k += KSTEP) {
            double x[6], u[2];
            load_x<false>(k, 0.0, x);
            load_u<false>(k, 0.0, u);
            th += dyn_res<false, false>(k, 0.0, x, u);
        }
        return warp_sum(th);
    }

    // unscaled objective at W
    __device__ double objective() const {
        double f = 0.0, ss = 0.0;
        for (int k = lane; k <= N; k += KSTEP) {
            double x[6];
            load_x<false>(k, 0.0, x);
            f += stage_cost(k, x);
            if (k < N) {
                double u[2];
                load_u<false>(k, 0.0, u);
                f += input_cost(u);
                if (hwd) { double de = W[6 * (k + 1) + 5] - x[5]; f += wdp[k] * de * de; }
            }
#pragma unroll
            for (int j = 0; j < M; j++) ss += W[isg(j, k)];
        }
        return warp_sum(f + kp.p.slack_w * ss);
    }

    // ---- evaluate rows (GR, JA) and dynamics residual at the current iterate (lane = stage)
    __device__ void eval_point() {
        for (int k = lane; k < N; k += KSTEP) {
            double x[6], xn[6], u[2];
            load_x<false>(k, 0.0, x);
            load_x<false>(k + 1, 0.0, xn);
            load_u<false>(k, 0.0, u);
            dyn_res<false, true>(k, 0.0, x, u);
OCP_ROLL_J
            for (int j = 0; j < M; j++) {
                int r = j * N + k;
                RowV v = row_vals(j, k, x, xn, W[isg(j, k)], W[isg(j, k + 1)]);
                double sc = DG[r];
                GR[r] = sc * row_g(v);
                double ja[4];
                ja[0] = sc * (-a1 * 6.0 * p5(v.ds) * v.iL);
                ja[1] = sc * (-a1 * 6.0 * p5(v.de) * v.iW);
                ja[2] = sc * (6.0 * p5(v.dsn) * v.iL);
                ja[3] = sc * (6.0 * p5(v.den) * v.iW);
                stv<4>(JA + 4 * r, ja);
            }
        }
        __syncwarp();
    }

    // ---- optimality error pieces that do not depend on mu (lane = stage; stage k owns x_k, sigma_k, u_k, rows (.,k))
    struct Err { double dual, prim, ysum, zsum; };
    __device__ Err error_base() const {
        double dual = 0.0, prim = 0.0, ysum = 0.0, zsum = 0.0;
        for (int k = lane; k <= N; k += KSTEP) {
            double lamn[6];  // lam_{k+1} (multiplier of c_k), zero for k = N
            if (k < N) ld6(LAM + 6 * k, lamn);
            else {
#pragma unroll
                for (int a = 0; a < 6; a++) lamn[a] = 0.0;
            }
            if (k >= 1) {
                double jy4 = 0.0, jy5 = 0.0;
OCP_ROLL_J
                for (int j = 0; j < M; j++) {  // J'y on (s, ey)
                    if (k < N) {
                        int r = j * N + k;
                        jy4 += JA[4 * r + 0] * Y[r];
                        jy5 += JA[4 * r + 1] * Y[r];
                    }
                    int r = j * N + k - 1;
                    jy4 += JA[4 * r + 2] * Y[r];
                    jy5 += JA[4 * r + 3] * Y[r];
                }
                double2 zl = ld2(ZL + bsx(k)), zu = ld2(ZU + bsx(k));
                zsum += zl.x + zl.y + zu.x + zu.y;
                double xk[6], dk[6];
                load_x<false>(k, 0.0, xk);
                diff_target(k, xk, dk);
OCP_ROLL_A
                for (int a = 0; a < 6; a++) {
                    double s = grad_x_comp(k, a, dk, xk[5]) + LAM[6 * (k - 1) + a];
#pragma unroll
                    for (int b = 0; b < 6; b++) s -= kp.p.A[6 * b + a] * lamn[b];
                    if (a == 0) s += -zl.x + zu.x;
                    if (a == 4) s -= jy4;
                    if (a == 5) s += -jy5 - zl.y + zu.y;
                    dual = fmax(dual, fabs(s));
                }
            }
OCP_ROLL_J
            for (int j = 0; j < M; j++) {  // sigma_{j,k}
                double zl = ZL[bss(j, k)];
                double rw = df * kp.p.slack_w - zl;
                if (k < N) rw -= DG[j * N + k] * a1 * Y[j * N + k];
                if (k >= 1) rw += DG[j * N + k - 1] * Y[j * N + k - 1];
                dual = fmax(dual, fabs(rw));
                zsum += zl;
            }
            if (k < N) {
                double u[2];
                load_u<false>(k, 0.0, u);
                double2 zl = ld2(ZL + bsu(k)), zu = ld2(ZU + bsu(k));
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    double s = df * (R2(c, 0) * u[0] + R2(c, 1) * u[1]);
#pragma unroll
                    for (int b = 0; b < 6; b++) s -= Bm(b, c) * lamn[b];
                    s += c ? (-zl.y + zu.y) : (-zl.x + zu.x);
                    dual = fmax(dual, fabs(s));
                }
                zsum += zl.x + zl.y + zu.x + zu.y;
                double c6[6];
                ld6(CRES + 6 * k, c6);
#pragma unroll
                for (int a = 0; a < 6; a++) {
                    prim = fmax(prim, fabs(c6[a]));
                    ysum += fabs(lamn[a]);
                }
OCP_ROLL_J
                for (int j = 0; j < M; j++) {
                    int r = j * N + k;
                    dual = fmax(dual, fabs(Y[r] - Z[r]));
                    dual = fmax(dual, fabs(rho - Y[r] - V[r]));
                    prim = fmax(prim, fabs(GR[r] + T[r] - S[r]));
                    zsum += Z[r] + V[r];
                    ysum += fabs(Y[r]);
                }
            }
        }
        Err e;
        e.dual = warp_max_nn(dual);
        e.prim = warp_max_nn(prim);
        e.ysum = warp_sum(ysum);
        e.zsum = warp_sum(zsum);
        return e;
    }
    // complementarity error for m = 0 (-> c0) and m = mu (-> cm) in ONE pass over the iterate: one copy of the code in
    // the instruction stream instead of two (the crowded kernel is instruction-fetch bound)
    __device__ void comp_err2(double m, double &c0, double &cm) const {
        double a0 = 0.0, am = 0.0;
        auto acc = [&](double v) { a0 = fmax(a0, fabs(v)); am = fmax(am, fabs(v - m)); };
        for (int k = lane; k <= N; k += KSTEP) {
            if (k >= 1) {
                double vx = W[6 * k], ey = W[6 * k + 5];
                double2 zl = ld2(ZL + bsx(k)), zu = ld2(ZU + bsx(k));
                if (has(xlb(k, 0))) acc((vx - xlb(k, 0)) * zl.x);
                if (has(xub(k, 0))) acc((xub(k, 0) - vx) * zu.x);
                if (has(xlb(k, 1))) acc((ey - xlb(k, 1)) * zl.y);
                if (has(xub(k, 1))) acc((xub(k, 1) - ey) * zu.y);
            }
OCP_ROLL_J
            for (int j = 0; j < M; j++) acc(W[isg(j, k)] * ZL[bss(j, k)]);
            if (k < N) {
                double u[2];
                load_u<false>(k, 0.0, u);
                double2 zl = ld2(ZL + bsu(k)), zu = ld2(ZU + bsu(k));
                acc((u[0] + kp.p.umax[0]) * zl.x);
                acc((kp.p.umax[0] - u[0]) * zu.x);
                acc((u[1] + kp.p.umax[1]) * zl.y);
                acc((kp.p.umax[1] - u[1]) * zu.y);
OCP_ROLL_J
                for (int j = 0; j < M; j++) {
                    int r = j * N + k;
                    acc(S[r] * Z[r]);
                    acc(T[r] * V[r]);
                }
            }
        }
        c0 = warp_max_nn(a0);
        cm = warp_max_nn(am);
    }
    __device__ double total_err(const Err &e, double comp) const {
        const double s_max = 100.0;
        int nb = nb_count;  // lower + upper bound multipliers + row (z, v)
        int nmul = 6 * N + R + nb;
        double sd = fmax(s_max, (e.ysum + e.zsum) / (double)(nmul > 0 ? nmul : 1)) / s_max;
        double sc = fmax(s_max, e.zsum / (double)(nb > 0 ? nb : 1)) / s_max;
        return fmax(e.dual / sd, fmax(e.prim, comp / sc));
    }

    // ---- per-iteration assembly (lane = stage): HD (diag Hessian additions), base gradient (into D), SIGE, YHAT
    __device__ void assemble() {
        for (int k = lane; k <= N; k += KSTEP) {
            if (k >= 1) {
                double x[6], dk[6], hd[6], gb0, gb5;
                load_x<false>(k, 0.0, x);
                diff_target(k, x, dk);
#pragma unroll
                for (int a = 0; a < 6; a++) hd[a] = 0.0;
                double2 zl = ld2(ZL + bsx(k)), zu = ld2(ZU + bsx(k));
                {
                    double lb = xlb(k, 0), ub = xub(k, 0);
                    double il = has(lb) ? rcp(x[0] - lb) : 0.0, iu = has(ub) ? rcp(ub - x[0]) : 0.0;
                    hd[0] = zl.x * il + zu.x * iu;
                    gb0 = -mu * il + mu * iu;
                    lb = xlb(k, 1);
                    ub = xub(k, 1);
                    il = has(lb) ? rcp(x[5] - lb) : 0.0;
                    iu = has(ub) ? rcp(ub - x[5]) : 0.0;
                    hd[5] = zl.y * il + zu.y * iu;
                    gb5 = -mu * il + mu * iu;
                }
                // base gradient of the barrier problem -> D (overwritten by the forward pass)
OCP_ROLL_A
                for (int a = 0; a < 6; a++) {
                    double gv = grad_x_comp(k, a, dk, x[5]);
                    if (a == 0) gv += gb0;
                    if (a == 5) gv += gb5;
                    D[6 * k + a] = gv;
                }
OCP_ROLL_J
                for (int j = 0; j < M; j++) {  // Hessian of -y_r g_r: diagonal on (s, ey) (control.py:544-557, degree 6)
                    if (k < N) {
                        int r = j * N + k;
                        double yd = Y[r] * DG[r] * a1 * 30.0;
                        hd[4] += yd * p4(x[4] - obs_s(j, k) - lapoff[j]) * iL6(j);
                        hd[5] += yd * p4(x[5] - obs_e(j, k)) * iW6(j);
                    }
                    int r = j * N + k - 1;
                    double yd = Y[r] * DG[r] * 30.0;
                    hd[4] -= yd * p4(x[4] - obs_s(j, k)) * iL6(j);
                    hd[5] -= yd * p4(x[5] - obs_e(j, k)) * iW6(j);
                }
                st6(HD + 6 * k, hd);
            }
OCP_ROLL_J
            for (int j = 0; j < M; j++) {
                double is = rcp(W[isg(j, k)]);
                HD[isg(j, k)] = ZL[bss(j, k)] * is;
                D[isg(j, k)] = df * kp.p.slack_w - mu * is;
            }
            if (k < N) {
                double u[2];
                load_u<false>(k, 0.0, u);
                double2 zl = ld2(ZL + bsu(k)), zu = ld2(ZU + bsu(k));
                double il0 = rcp(u[0] + kp.p.umax[0]), iu0 = rcp(kp.p.umax[0] - u[0]);
                double il1 = rcp(u[1] + kp.p.umax[1]), iu1 = rcp(kp.p.umax[1] - u[1]);
                st2(HD + OU + 2 * k, zl.x * il0 + zu.x * iu0, zl.y * il1 + zu.y * iu1);
                st2(D + OU + 2 * k, df * (R2(0, 0) * u[0] + R2(0, 1) * u[1]) - mu * il0 + mu * iu0,
                    df * (R2(1, 0) * u[0] + R2(1, 1) * u[1]) - mu * il1 + mu * iu1);
OCP_ROLL_J
                for (int j = 0; j < M; j++) {
                    int r = j * N + k;
                    double s = S[r], tt = T[r];
                    double is = rcp(s), it = rcp(tt);
                    double sig_s = Z[r] * is, sig_t = V[r] * it;
                    double beta = sig_t * rcp(sig_s + sig_t);
                    double rg = GR[r] + tt - s;
                    SIGE[r] = beta * sig_s;
                    YHAT[r] = (1.0 - beta) * (rho - mu * it) + beta * (mu * is - sig_s * rg);
                }
            }
        }
        __syncwarp();
    }

    // ---- Riccati backward sweep with primal regularisation dw.  Returns false if a pivot <= 0.
    // Stage variables zeta = (dx 6, dsigma M | du 2, dsigma+ M); lane l < NZ owns column l of the stage Hessian G,
    // lane a < NXA additionally owns row a of the value function P (registers).
#ifdef B200MPC_PHASE_CLOCKS
    long long bc[5] = {0, 0, 0, 0, 0};
#define BCLK(k) { long long t_ = clock64(); bc[k] += t_ - tb0; tb0 = t_; }
#else
#define BCLK(k)
#endif
    __device__ bool riccati_backward(double dw) {
#ifdef B200MPC_PHASE_CLOCKS
        long long tb0 = clock64();
#endif
        double Pr[NXA], pv;
        {   // terminal value function: stage-N state block
            int idx = isx ? 6 * N + rl : (iss ? isg(jrole, N) : 0);
            double dN = (isx || iss) ? HD[idx] + dw : 0.0;
            pv = (isx || iss) ? D[idx] : 0.0;
#pragma unroll
            for (int b = 0; b < NXA; b++) Pr[b] = ((b < 6) ? qqcol[b < 6 ? b : 0] : 0.0) + ((b == rl) ? dN : 0.0);
        }
        bool ok = true;
        for (int k = N - 1; k >= 0; k--) {
            // (1) row a of P*[A B] (constants as immediates), q = p - P[:,0:6] c_k; transpose through PT
            double c6[6];
            ld6(CRES + 6 * k, c6);
            {
                constexpr int NCOL = kMirror ? 4 : 8;   // columns of [A B] per lane
                const int c0 = kMirror ? (lane >> 4) * NCOL : 0;
                double ptx[NCOL];
#pragma unroll
                for (int c = 0; c < NCOL; c++) ptx[c] = 0.0;
                double qv = pv;
#pragma unroll
                for (int b = 0; b < 6; b++) {
                    double ab[NCOL];
                    ldv<NCOL>(ABs + 8 * b + c0, ab);
#pragma unroll
                    for (int c = 0; c < NCOL; c++) ptx[c] += Pr[b] * ab[c];
                    qv -= Pr[b] * c6[b];
                }
                if (rl < NXA) {
#pragma unroll
                    for (int c = 0; c < NCOL; c++) PT[(c0 + c) * NXAP + rl] = ptx[c];
                }
                if (lane < NXA) {   // the mirror half would store the same values to the same addresses
#pragma unroll
                    for (int j = 0; j < M; j++) PT[(8 + j) * NXAP + lane] = Pr[6 + j];
                    QVs[lane] = qv;
                }
            }
            __syncwarp();
            BCLK(0)
            // (2) column l of G = base + T'(P T)[:, l] + sum_j SIGE_j ct_j ct_j[l].  The row vectors are ct_j = T'c_j with
            //     c_j = (ja0, ja1) on (s, ey) of x_k and (ja2, ja3) on (s, ey) of x_{k+1}, so the whole sum over the rivals
            //     folds into four scalars S0..S3 = sum_j ja._j w_j (w_j = SIGE_j ct_j[l]): S2, S3 join (P T)[4:6, l] BEFORE
            //     the product with [A B]', S0, S1 land on rows s, ey.  The input rows (G_uu feeds the Cholesky) and the
            //     gradient come before the barrier, the six state rows after it: they are needed only in (4) and issue
            //     in the shadow of the factorisation's dependent chain (one basic block with (3)).
            double g[NZ], gv, dsg, cv4, cv5, S0 = 0.0, S1 = 0.0;
            {
                double colv[NXAP], qvs[NXAP];
                ldv<NXAP>(PT + cmap * NXAP, colv);
                ldv<NXAP>(QVs, qvs);
                int idx = idx_base + k * idx_step;
                bool live = k >= kmin;
                dsg = live ? HD[idx] + dw : 0.0;
                gv = live ? D[idx] : 0.0;
#pragma unroll
                for (int q = 0; q < 6; q++) gv += tcol[q] * qvs[q];
#pragma unroll
                for (int j = 0; j < M; j++) gv += psel[j] * qvs[6 + j];
                double S2 = 0.0, S3 = 0.0;
#pragma unroll
                for (int j = 0; j < M; j++) {
                    int r = j * N + k;
                    double ja[4];
                    ldv<4>(JA + 4 * r, ja);
                    double dgr = DG[r], sg = SIGE[r], yh = YHAT[r];
                    double own = tcol[4] * ja[2] + tcol[5] * ja[3];
                    own += e4 * ja[0] + e5 * ja[1];
                    own += crow[j] * dgr;
                    double w = sg * own;
                    S0 += ja[0] * w;
                    S1 += ja[1] * w;
                    S2 += ja[2] * w;
                    S3 += ja[3] * w;
                    g[6 + j] = (((6 + j) == rl) ? dsg : 0.0) + (dgr * a1) * w;
                    g[NXA + 2 + j] = (colv[6 + j] + (((NXA + 2 + j) == rl) ? dsg : 0.0)) - dgr * w;
                    gv -= yh * own;
                }
                if (hwd) {   // curvature of wd_k (ey_{k+1}-ey_k)^2: a pure "cost row" J = e5'(dx_{k+1}-dx_k), weight 2 df wd_k
                    double w = (2.0 * df * wdp[k]) * (tcol[5] - ((rl == 5) ? 1.0 : 0.0));
                    S3 += w;
                    S1 -= w;
                }
                gv -= S2 * c6[4] + S3 * c6[5];
                cv4 = colv[4] + S2;
                cv5 = colv[5] + S3;
                colv[4] = cv4;
                colv[5] = cv5;
#pragma unroll
                for (int c = 0; c < 2; c++) g[NXA + c] = rrcol[c] + (((NXA + c) == rl) ? dsg : 0.0);
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    double ab[2];
                    ldv<2>(ABs + 8 * q + 6, ab);
#pragma unroll
                    for (int c = 0; c < 2; c++) g[NXA + c] += ab[c] * colv[q];
                }
                if (lane >= NXA && lane < NZ) {
                    double gu[NUAP];
#pragma unroll
                    for (int m = 0; m < NUA; m++) gu[m] = g[NXA + m];
                    if (NUAP > NUA) gu[NUAP - 1] = 0.0;
                    stv<NUAP>(GUU + (lane - NXA) * NUAP, gu);
                    GVU[lane - NXA] = gv;
                }
            }
            __syncwarp();
            BCLK(1)
            {   // state rows of the column (PT is rewritten only after the next barrier)
                double c4[4];
                ldv<4>(PT + cmap * NXAP, c4);
                const double colv[6] = {c4[0], c4[1], c4[2], c4[3], cv4, cv5};
#pragma unroll
                for (int a = 0; a < 6; a++) g[a] = qqcol[a] + ((a == rl) ? dsg : 0.0);
                g[4] += S0;
                g[5] += S1;
#pragma unroll
                for (int q = 0; q < 6; q++) {
                    double ab[6];
                    ld6(ABs + 8 * q, ab);
#pragma unroll
                    for (int a = 0; a < 6; a++) g[a] += ab[a] * colv[q];
                }
            }
            // (3) root-free factorisation G_uu = L D L' (L unit lower) redundantly in every lane, column solves.  Against
            //     the Cholesky form: a reciprocal instead of a reciprocal square root per pivot, and the substitutions carry
            //     no scaling on their dependent chains (one FMA per step): s = L^-1 b forward, k = L^-T (D^-1 s) backward.
            //     YF / YG keep the unscaled s; (4) subtracts s_a' D^-1 s_own.
            double Lt[NUA][NUA], rd[NUA], sv[NUA], tv[NUA];
            {
                double Vr[NUA][NUA];   // unscaled entries v_iq = l_iq d_q
#pragma unroll
                for (int a = 0; a < NUA; a++) {
                    double row[NUAP];
                    ldv<NUAP>(GUU + a * NUAP, row);
#pragma unroll
                    for (int b = 0; b <= a; b++) Vr[a][b] = row[b];
                }
#pragma unroll
                for (int j = 0; j < NUA; j++) {
                    double d = Vr[j][j];
#pragma unroll
                    for (int q = 0; q < j; q++) d -= Lt[j][q] * Vr[j][q];
                    if (!(d > 0.0)) ok = false;
                    double r = rcp(d);
                    rd[j] = r;
#pragma unroll
                    for (int i = j + 1; i < NUA; i++) {
                        double v = Vr[i][j];
#pragma unroll
                        for (int q = 0; q < j; q++) v -= Lt[i][q] * Vr[j][q];
                        Vr[i][j] = v;
                        Lt[i][j] = v * r;
                    }
                }
                double gvu[NUAP], kv[NUA];
                ldv<NUAP>(GVU, gvu);
                const bool gcol = (rl == NZ);  // this lane solves the gradient column
#pragma unroll
                for (int a = 0; a < NUA; a++) {
                    double v = gcol ? gvu[a] : g[NXA + a];
#pragma unroll
                    for (int q = 0; q < a; q++) v -= Lt[a][q] * sv[q];
                    sv[a] = v;
                }
#pragma unroll
                for (int a = 0; a < NUA; a++) tv[a] = sv[a] * rd[a];
#pragma unroll
                for (int a = NUA - 1; a >= 0; a--) {
                    double v = tv[a];
#pragma unroll
                    for (int q = a + 1; q < NUA; q++) v -= Lt[q][a] * kv[q];
                    kv[a] = v;
                }
                if (!ok) {   // checked after the solves so that they overlap the factorisation's latency
                    // the state rows above read PT after the last barrier, and the retry's first stage rewrites it: order them
                    // (found by ThreadSanitizer on the host emulation; the warp is converged here in practice, not by contract)
                    __syncwarp();
                    return false;
                }
                if (lane < NXA) {
#pragma unroll
                    for (int m = 0; m < NUA; m++) KFB[(k * NUA + m) * NKP + lane] = -kv[m];
                    double yp[NUAP];
#pragma unroll
                    for (int m = 0; m < NUA; m++) yp[m] = sv[m];
                    if (NUAP > NUA) yp[NUAP - 1] = 0.0;
                    stv<NUAP>(YF + lane * NUAP, yp);
                } else if (lane == NZ) {
                    double yp[NUAP];
#pragma unroll
                    for (int m = 0; m < NUA; m++) { yp[m] = sv[m]; KFB[(k * NUA + m) * NKP + NXA] = -kv[m]; }   // feed-forward term
                    if (NUAP > NUA) yp[NUAP - 1] = 0.0;
                    stv<NUAP>(YG, yp);
                }
            }
            __syncwarp();
            BCLK(2)
            // (4) row b of P = G_xx - S' D^-1 S, p = g_x - S' D^-1 s_g (state lanes keep them in registers)
            {
                double yg[NUAP];
                ldv<NUAP>(YG, yg);
                double pn = gv;
#pragma unroll
                for (int m = 0; m < NUA; m++) pn -= tv[m] * yg[m];
                pv = pn;
#pragma unroll
                for (int a = 0; a < NXA; a++) {
                    double ya[NUAP];
                    ldv<NUAP>(YF + a * NUAP, ya);
                    double sa = g[a];
#pragma unroll
                    for (int m = 0; m < NUA; m++) sa -= ya[m] * tv[m];
                    Pr[a] = sa;
                }
            }
            BCLK(3)
        }
        // stage 0: x_0 is fixed (control.py:497), sigma_{.,0} is free: d sigma_0 = -P_ss^-1 p_s  -> QVs[6+j]
        if (M > 0) {
            if (iss && lane == rl) {   // not the mirror half: it would store the same values to the same addresses
#pragma unroll
                for (int j = 0; j < M; j++) S0[jrole * (M + 2) + j] = Pr[6 + j];
                S0[jrole * (M + 2) + M] = pv;
            }
            __syncwarp();
            double Ls[MM][MM], ri[MM], yv[MM], xv[MM];
#pragma unroll
            for (int a = 0; a < M; a++)
#pragma unroll
                for (int b = 0; b <= a; b++) Ls[a][b] = S0[a * (M + 2) + b];
#pragma unroll
            for (int j = 0; j < M; j++) {
                double d = Ls[j][j];
#pragma unroll
                for (int q = 0; q < j; q++) d -= Ls[j][q] * Ls[j][q];
                if (!(d > 0.0)) ok = false;
                ri[j] = rsq(d);
#pragma unroll
                for (int i = j + 1; i < M; i++) {
                    double s = Ls[i][j];
#pragma unroll
                    for (int q = 0; q < j; q++) s -= Ls[i][q] * Ls[j][q];
                    Ls[i][j] = s * ri[j];
                }
            }
            if (!ok) return false;
#pragma unroll
            for (int a = 0; a < M; a++) {
                double s = -S0[a * (M + 2) + M];
#pragma unroll
                for (int q = 0; q < a; q++) s -= Ls[a][q] * yv[q];
                yv[a] = s * ri[a];
            }
#pragma unroll
            for (int a = M - 1; a >= 0; a--) {
                double s = yv[a];
#pragma unroll
                for (int q = a + 1; q < M; q++) s -= Ls[q][a] * xv[q];
                xv[a] = s * ri[a];
            }
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int a = 0; a < M; a++) QVs[6 + a] = xv[a];
            }
            __syncwarp();
        }
        return true;
    }

    // ---- forward sweep: D <- Newton direction.  Every lane carries dX_k in registers.  Lane m < NUA evaluates
    //      row m of the feedback law (K row from shared memory), lane a < 6 evaluates component a of the state
    //      update with its private rows of A and B; results travel by shuffles -- no barrier, no model loads.
    __device__ void riccati_forward() {
        double dx[NXAP];
#pragma unroll
        for (int a = 0; a < NXAP; a++) dx[a] = 0.0;
#pragma unroll
        for (int j = 0; j < M; j++) dx[6 + j] = QVs[6 + j];
        __syncwarp();
        if (lane < 6) D[lane] = 0.0;
        if (M > 0 && lane < M) D[isg(lane, 0)] = QVs[6 + lane];
        const int mrow = (lane < NUA) ? lane : 0;
        const int arow_i = (lane < 6) ? lane : 0;
        // slot of this lane's input component in D: u_k at OU + 2k + lane, sigma_{j,k+1} at isg(j, k + 1) -- base + k * step, no branch
        const int dst_base = (lane < 2) ? OU + lane : isg((lane < NUA) ? lane - 2 : 0, 1);
        const int dst_step = (lane < 2) ? 2 : 1;
        for (int k = 0; k < N; k++) {
            double kr[NKP];
            ldv<NKP>(KFB + (k * NUA + mrow) * NKP, kr);
            // three / two partial sums instead of one chain of NXA / 6 dependent FMAs (the sweep is a dependent chain)
            double s = kr[NXA], s1 = 0.0, s2 = 0.0;
            double t = -CRES[6 * k + arow_i], t1 = 0.0;
#pragma unroll
            for (int c = 0; c < NXA; c += 3) {
                s += kr[c] * dx[c];
                if (c + 1 < NXA) s1 += kr[c + 1] * dx[c + 1];
                if (c + 2 < NXA) s2 += kr[c + 2] * dx[c + 2];
            }
            s += s1 + s2;
#pragma unroll
            for (int b = 0; b < 6; b += 2) {
                t += arow[b] * dx[b];
                t1 += arow[b + 1] * dx[b + 1];
            }
            t += t1;
            double du[NUA];
#pragma unroll
            for (int m = 0; m < NUA; m++) du[m] = __shfl_sync(0xffffffffu, s, m);
            t += brow[0] * du[0] + brow[1] * du[1];
            if (lane < NUA) D[dst_base + k * dst_step] = s;
            if (lane < 6) D[6 * (k + 1) + lane] = t;
#pragma unroll
            for (int a = 0; a < 6; a++) dx[a] = __shfl_sync(0xffffffffu, t, a);
#pragma unroll
            for (int j = 0; j < M; j++) dx[6 + j] = du[2 + j];
        }
        __syncwarp();
    }
};

// ---------------------------------------------------------------- kernel
// 248 registers, not 255: the register file is granted per warp in units of 256, so 8 resident solver warps take 8 x 7936 and
// leave room for the one-warp exchange kernels (csrc/exchange.cuh) on a full SM; above 248 they would take all 65536.
// __maxnreg__ and __launch_bounds__ exclude each other; measured the same speed either way (profiles/r06_variants_3.txt).
#ifdef B200MPC_HOST_EMULATION
#define OCP_KERNEL_BOUNDS
#elif defined(B200MPC_LAUNCH_BOUNDS)
#define OCP_KERNEL_BOUNDS __launch_bounds__(32)
#else
#define OCP_KERNEL_BOUNDS __maxnreg__(248)
#endif
template <int M, int FL, int NT>
__global__ void OCP_KERNEL_BOUNDS ocp_ipm_kernel(const __grid_constant__ KParams kp, const double *__restrict__ in,
                                                     b200mpc_record *__restrict__ rec, double *__restrict__ aux,
                                                     double *__restrict__ xpred, double *__restrict__ upred,
                                                     double *__restrict__ sigma, const XchgArgs xa = XchgArgs{nullptr, 0, 0, 0}) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x;
    const int inst = blockIdx.x;
    const SmemPlan<M> pl(NT ? NT : kp.p.N, NT ? cbf_record_doubles(NT, M, 0, FL) : kp.in_stride, FL);
    Ipm<M, FL, NT> S_(kp, pl, sm, lane);
    Ipm<M, FL, NT> &q = S_;
    using IP = Ipm<M, FL, NT>;
    constexpr int NXAP = IP::NXAP, NC = IP::NC, KSTEP = IP::KSTEP;
    const int N = q.N, R = q.R, NW = q.NW, OU = q.OU, OS = q.OS;
    const b200mpc_ipm_options &o = kp.o;

    // ---- stage the instance record with one TMA bulk copy
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
    const uint32_t in_bytes = (uint32_t)kp.in_stride * 8u;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(bar, in_bytes);
        bulk_g2s(q.IN, in + (size_t)inst * kp.in_stride, in_bytes, bar);
    }
    for (int e = lane; e < NW + 2; e += 32) { q.W[e] = 0.0; q.D[e] = 0.0; q.HD[e] = 0.0; }
    for (int e = lane; e < (NC + 1) * NXAP; e += 32) q.PT[e] = 0.0;   // includes the all-zero row NC
    for (int e = lane; e < 6 * N + 6; e += 32) { q.LAM[e] = 0.0; q.CRES[e] = 0.0; }
    for (int e = lane; e < 48; e += 32) { int r_ = e >> 3, c_ = e & 7; q.ABs[e] = (c_ < 6) ? kp.p.A[6 * r_ + c_] : kp.p.B[2 * r_ + (c_ - 6)]; }
    mbar_wait(bar, 0);
    __syncwarp();
    if (IP::prs) {   // (L_j, W_j) -> (1/L_j^6, 1/W_j^6), in place in the staged record
        if (lane < 2 * M) {
            double v = q.IN[kp.sz_off + lane];
            q.IN[kp.sz_off + lane] = 1.0 / p6(v);
        }
        __syncwarp();
    }

    // ---- start point.  start = B200MPC_START_ROLLOUT: u = 0 roll-out from x_0 (every lane redundantly, constants as
    //      immediates); B200MPC_START_ZERO: x_1..x_N = 0, what Opti/IPOPT start from (the reference never calls
    //      opti.set_initial in control.py:476-607).  u = 0, sigma = 0; everything is then pushed into the bounds.
    {
        double x[6];
        ld6(q.IN, x);
        if (lane == 0) st6(q.W, x);
        for (int i = 1; o.start == B200MPC_START_ROLLOUT && i <= N; i++) {
            double xn[6];
#pragma unroll
            for (int a = 0; a < 6; a++) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < 6; b++) s += q.Am(a, b) * x[b];
                xn[a] = s;
            }
#pragma unroll
            for (int a = 0; a < 6; a++) x[a] = xn[a];
            if (lane == (i & 31)) st6(q.W + 6 * i, x);
        }
    }
    __syncwarp();
    int nbc = 0;
    for (int k = lane; k <= N; k += KSTEP) {
        if (k >= 1) {   // bound_push / bound_frac on vx_k, ey_k
#pragma unroll
            for (int c = 0; c < 2; c++) {
                int wi = 6 * k + (c ? 5 : 0);
                double w = q.W[wi];
                double lb = q.xlb(k, c), ub = q.xub(k, c);
                bool hl = IP::has(lb), hu = IP::has(ub);
                if (hl) {
                    double pl_ = o.bound_push * fmax(1.0, fabs(lb));
                    if (hu) pl_ = fmin(pl_, o.bound_frac * (ub - lb));
                    if (w < lb + pl_) w = lb + pl_;
                }
                if (hu) {
                    double pu = o.bound_push * fmax(1.0, fabs(ub));
                    if (hl) pu = fmin(pu, o.bound_frac * (ub - lb));
                    if (w > ub - pu) w = ub - pu;
                }
                q.W[wi] = w;
                q.ZL[q.bsx(k) + c] = hl ? 1.0 : 0.0;
                q.ZU[q.bsx(k) + c] = hu ? 1.0 : 0.0;
                nbc += (hl ? 1 : 0) + (hu ? 1 : 0);
            }
        }
#pragma unroll
        for (int j = 0; j < M; j++) {
            q.W[q.isg(j, k)] = o.bound_push;   // sigma = 0 pushed off its lower bound 0: push*max(1,|0|)
            q.ZL[q.bss(j, k)] = 1.0;
            nbc += 1;
        }
        if (k < N) {
#pragma unroll
            for (int c = 0; c < 2; c++) {
                double ub = kp.p.umax[c], lb = -ub;
                double w = 0.0;
                double pl_ = fmin(o.bound_push * fmax(1.0, fabs(lb)), o.bound_frac * (ub - lb));
                if (w < lb + pl_) w = lb + pl_;
                if (w > ub - pl_) w = ub - pl_;
                q.W[OU + 2 * k + c] = w;
                q.ZL[q.bsu(k) + c] = 1.0;
                q.ZU[q.bsu(k) + c] = 1.0;
            }
            nbc += 4 + 2 * M;   // u bounds + (z, v) of the rows of this stage
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nbc += __shfl_xor_sync(0xffffffffu, nbc, off);
    q.nb_count = nbc;
    __syncwarp();
    // ---- gradient-based scaling at the start (nlp_scaling_max_gradient)
    {
        double gm = (M > 0) ? kp.p.slack_w : 0.0;
        for (int k = lane; k <= N; k += KSTEP) {
            if (k >= 1) {
                double x[6], dk[6];
                q.template load_x<false>(k, 0.0, x);
                q.diff_target(k, x, dk);
OCP_ROLL_A
                for (int a = 0; a < 6; a++) gm = fmax(gm, fabs(q.grad_x_comp(k, a, dk, x[5])));   // df == 1 here
            }
            if (k < N) {
                double u[2];
                q.template load_u<false>(k, 0.0, u);
                gm = fmax(gm, fabs(q.R2(0, 0) * u[0] + q.R2(0, 1) * u[1]));
                gm = fmax(gm, fabs(q.R2(1, 0) * u[0] + q.R2(1, 1) * u[1]));
            }
        }
        gm = warp_max_nn(gm);
        q.df = gm > o.max_grad ? o.max_grad / gm : 1.0;
        q.set_scaled_columns();
        for (int k = lane; k < N; k += KSTEP) {
            double x[6], xn[6];
            q.template load_x<false>(k, 0.0, x);
            q.template load_x<false>(k + 1, 0.0, xn);
#pragma unroll
            for (int j = 0; j < M; j++) {
                int r = j * N + k;
                typename IP::RowV v = q.row_vals(j, k, x, xn, q.W[q.isg(j, k)], q.W[q.isg(j, k + 1)]);
                double rm = fmax(q.a1, 1.0);  // |d/dsigma_i| = (1-alpha), |d/dsigma_{i+1}| = 1
                rm = fmax(rm, fmax(fabs(6.0 * p5(v.dsn) * v.iL), fabs(6.0 * p5(v.den) * v.iW)));
                if (k > 0) rm = fmax(rm, q.a1 * fmax(fabs(6.0 * p5(v.ds) * v.iL), fabs(6.0 * p5(v.de) * v.iW)));
                double dgr = rm > o.max_grad ? o.max_grad / rm : 1.0;
                q.DG[r] = dgr;
                double g = dgr * q.row_g(v);
                double t = fmax(0.0, -g) + o.bound_push;
                q.T[r] = t;
                q.S[r] = g + t;
                q.Z[r] = 1.0;
                q.V[r] = 1.0;
                q.Y[r] = 0.0;
                q.JD[r] = 0.0;
                q.SIGE[r] = 0.0;
                q.YHAT[r] = 0.0;
                q.GR[r] = g;
            }
        }
        __syncwarp();
    }
    double th0 = q.theta0();
    const double theta_max = 1e4 * fmax(1.0, th0), theta_min = 1e-4 * fmax(1.0, th0);

    // filter: entry f lives in lane f%32, slot f/32 (capacity 64; the oldest entry is overwritten)
    double f_th0 = 0.0, f_ph0 = 0.0, f_th1 = 0.0, f_ph1 = 0.0;
    int nfilt = 0, fpos = 0;
    int iter = 0, status = B200MPC_MAX_ITER, n_acc = 0, n_refac = 0, n_back = 0, n_reset = 0;
    double dw_last = 0.0, E0 = 0.0;
    bool last_needed = false;
    // phi at the current point = phi of the trial point accepted by the previous iteration, as long as mu and the
    // slacks were not touched in between (IPOPT caches it the same way)
    bool ph_valid = false;
    double ph_cache = 0.0;

    const double kappa_eps = 10.0, kappa_mu = 0.2, tau_min = 0.99;
    const double gamma_theta = 1e-5, gamma_phi = 1e-8, s_theta = 1.1, s_phi = 2.3, eta_phi = 1e-8;
    const double gamma_alpha = 0.05, kappa_sigma = 1e10;
    const bool one_round = (N < 32);   // every stage has its own lane

#ifdef B200MPC_PHASE_CLOCKS
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tc0 = clock64();
#define PCLK(k) { long long t_ = clock64(); pc[k] += t_ - tc0; tc0 = t_; }
#else
#define PCLK(k)
#endif
    for (;;) {
        PCLK(7)
        q.eval_point();
        typename IP::Err eb = q.error_base();
        bool stop = false;
        // termination test (first pass) and the monotone Fiacco-McCormick barrier update share one evaluation site
        for (bool first = true;; first = false) {
            double c0, cm;
            q.comp_err2(q.mu, c0, cm);
            if (first) {
                E0 = q.total_err(eb, c0);
                if (E0 <= o.tol) { status = B200MPC_SOLVED; stop = true; break; }
                if (E0 <= o.acceptable_tol) {
                    if (++n_acc >= o.acceptable_iter) { status = B200MPC_SOLVED; stop = true; break; }
                } else
                    n_acc = 0;
                if (iter >= o.max_iter) { status = B200MPC_MAX_ITER; stop = true; break; }
            }
            double em = q.total_err(eb, cm);
            if (em <= kappa_eps * q.mu && q.mu > o.tol / 11.0) {
                q.mu = fmax(o.tol / 11.0, fmin(kappa_mu * q.mu, q.mu * sqrt(q.mu)));
                nfilt = 0;
                fpos = 0;
                ph_valid = false;
            } else
                break;
        }
        if (stop) break;
        const double mu = q.mu, rho = q.rho;
        const double tau = fmax(tau_min, 1.0 - mu);
        PCLK(0)
        // ---- Newton step by Riccati, with inertia correction
        q.assemble();
        PCLK(1)
        // grad(phi)'d needs the base gradient of this lane's stage, which the forward sweep overwrites in D
        double gbx[6] = {0, 0, 0, 0, 0, 0}, gbu[2] = {0, 0};
        if (one_round && lane <= N) {
            ld6(q.D + 6 * lane, gbx);
            if (lane < N) { double2 t2 = ld2(q.D + OU + 2 * lane); gbu[0] = t2.x; gbu[1] = t2.y; }
        }
        // IPOPT always retries dw = 0 first; when the previous iteration needed a correction we start from
        // dw_last/3 instead (saves one full sweep per iteration on the non-convex stragglers)
        double dw_try = last_needed ? fmax(1e-20, dw_last / 3.0) : 0.0;
        bool fail = false;
        for (;;) {
            if (q.riccati_backward(dw_try)) break;
            n_refac++;
            if (dw_try == 0.0) dw_try = (dw_last == 0.0) ? 1e-4 : fmax(1e-20, dw_last / 3.0);
            else dw_try *= (dw_last == 0.0) ? 100.0 : 8.0;
            if (dw_try > 1e40) { fail = true; break; }
        }
        if (fail) { status = B200MPC_INERTIA; break; }
        if (dw_try > 0.0) dw_last = dw_try;
        last_needed = dw_try > 0.0;
        PCLK(2)
        q.riccati_forward();
        PCLK(3)
        // ---- rows: J d, step bounds, directional derivative (lane = stage)
        for (int k = lane; k < N; k += KSTEP) {
            double dxk[6], dxn[6];
            ld6(q.D + 6 * k, dxk);
            ld6(q.D + 6 * (k + 1), dxn);
OCP_ROLL_J
            for (int j = 0; j < M; j++) {
                int r = j * N + k;
                double ja[4];
                ldv<4>(q.JA + 4 * r, ja);
                q.JD[r] = ja[0] * dxk[4] + ja[1] * dxk[5] + q.DG[r] * q.a1 * q.D[q.isg(j, k)] + ja[2] * dxn[4] + ja[3] * dxn[5] -
                          q.DG[r] * q.D[q.isg(j, k + 1)];
            }
            if (q.hwd) q.JDC[k] = dxn[5] - dxk[5];
        }
        __syncwarp();
        double a_max = 1.0, a_z = 1.0, gphi = 0.0, th = 0.0;
        for (int k = lane; k <= N; k += KSTEP) {
            if (k >= 1) {
                double x[6], d[6];
                q.template load_x<false>(k, 0.0, x);
                ld6(q.D + 6 * k, d);
                double2 zl = ld2(q.ZL + q.bsx(k)), zu = ld2(q.ZU + q.bsx(k));
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    double w = c ? x[5] : x[0], dd = c ? d[5] : d[0];
                    double zlc = c ? zl.y : zl.x, zuc = c ? zu.y : zu.x;
                    double lb = q.xlb(k, c), ub = q.xub(k, c);
                    double idd = rcp(dd);   // +-inf for dd == 0 is never selected below
                    if (IP::has(lb)) {
                        double dl = w - lb;
                        double dzl = (mu - zlc * dd) * rcp(dl) - zlc;
                        if (dd < 0.0) a_max = fmin(a_max, -tau * dl * idd);
                        if (dzl < 0.0) a_z = fmin(a_z, -tau * zlc * rcp(dzl));
                    }
                    if (IP::has(ub)) {
                        double du = ub - w;
                        double dzu = (mu + zuc * dd) * rcp(du) - zuc;
                        if (dd > 0.0) a_max = fmin(a_max, tau * du * idd);
                        if (dzu < 0.0) a_z = fmin(a_z, -tau * zuc * rcp(dzu));
                    }
                }
                if (one_round) {
#pragma unroll
                    for (int a = 0; a < 6; a++) gphi += gbx[a] * d[a];
                } else {   // long horizons: recompute the base gradient from the iterate
                    double dk[6], gb[6] = {0, 0, 0, 0, 0, 0};
                    q.diff_target(k, x, dk);
                    q.barrier_grad_x(k, x, gb);
                    gphi += gb[0] * d[0] + gb[5] * d[5];
OCP_ROLL_A
                    for (int a = 0; a < 6; a++) gphi += q.grad_x_comp(k, a, dk, x[5]) * q.D[6 * k + a];
                }
            }
OCP_ROLL_J
            for (int j = 0; j < M; j++) {
                double w = q.W[q.isg(j, k)], dd = q.D[q.isg(j, k)], zlc = q.ZL[q.bss(j, k)];
                double dzl = (mu - zlc * dd) * rcp(w) - zlc;
                if (dd < 0.0) a_max = fmin(a_max, -tau * w * rcp(dd));
                if (dzl < 0.0) a_z = fmin(a_z, -tau * zlc * rcp(dzl));
                gphi += (q.df * kp.p.slack_w - mu * rcp(w)) * dd;
            }
            if (k < N) {
                double u[2];
                q.template load_u<false>(k, 0.0, u);
                double2 dd2 = ld2(q.D + OU + 2 * k);
                double2 zl = ld2(q.ZL + q.bsu(k)), zu = ld2(q.ZU + q.bsu(k));
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    double w = u[c], dd = c ? dd2.y : dd2.x, zlc = c ? zl.y : zl.x, zuc = c ? zu.y : zu.x;
                    double dl = w + kp.p.umax[c], du = kp.p.umax[c] - w;
                    double idl = rcp(dl), idu = rcp(du), idd = rcp(dd);
                    double dzl = (mu - zlc * dd) * idl - zlc, dzu = (mu + zuc * dd) * idu - zuc;
                    if (dd < 0.0) a_max = fmin(a_max, -tau * dl * idd);
                    if (dd > 0.0) a_max = fmin(a_max, tau * du * idd);
                    if (dzl < 0.0) a_z = fmin(a_z, -tau * zlc * rcp(dzl));
                    if (dzu < 0.0) a_z = fmin(a_z, -tau * zuc * rcp(dzu));
                    double gb = one_round ? gbu[c] : (q.df * (q.R2(c, 0) * u[0] + q.R2(c, 1) * u[1]) - mu * idl + mu * idu);
                    gphi += gb * dd;
                }
                double c6[6];
                ld6(q.CRES + 6 * k, c6);
#pragma unroll
                for (int a = 0; a < 6; a++) th += fabs(c6[a]);
OCP_ROLL_J
                for (int j = 0; j < M; j++) {
                    int r = j * N + k;
                    double ds, dt, dy, dz, dv;
                    q.row_step(r, ds, dt, dy, dz, dv);
                    double s = q.S[r], tt = q.T[r];
                    if (ds < 0.0) a_max = fmin(a_max, -tau * s * rcp(ds));
                    if (dt < 0.0) a_max = fmin(a_max, -tau * tt * rcp(dt));
                    if (dz < 0.0) a_z = fmin(a_z, -tau * q.Z[r] * rcp(dz));
                    if (dv < 0.0) a_z = fmin(a_z, -tau * q.V[r] * rcp(dv));
                    gphi += rho * dt - mu * (ds * rcp(s) + dt * rcp(tt));
                    th += fabs(q.GR[r] + tt - s);
                }
            }
        }
        a_max = warp_min_nn(a_max);
        a_z = warp_min_nn(a_z);
        gphi = warp_sum(gphi);
        th = warp_sum(th);
        PCLK(4)
        // ---- filter line search.  One evaluation site: the first pass (alpha = 0) yields phi at the current point,
        //      the following passes are the trial points (same code, warm in the instruction cache).
        // switching condition a (-gphi)^s_phi > delta th^s_theta and the a_min term delta th^s_theta / (-gphi)^s_phi in
        // the log domain: lsw = s_phi log(-gphi) - log(delta) - s_theta log(th); a = a_max 2^-nls
        double lsw = 0.0, la_max = 0.0;
        if (gphi < 0.0) {   // the three logarithms through ONE call site: lane 0/1/2 take -gphi, th, a_max (delta_sw = 1)
            double lg = log(lane == 0 ? -gphi : (lane == 1 ? th : a_max));
            lsw = s_phi * __shfl_sync(0xffffffffu, lg, 0) - s_theta * __shfl_sync(0xffffffffu, lg, 1);
            la_max = __shfl_sync(0xffffffffu, lg, 2);
        }
        double amin;
        if (gphi < 0.0 && th <= theta_min)
            amin = gamma_alpha * fmin(gamma_theta, fmin(gamma_phi * th / (-gphi), exp(-lsw)));
        else if (gphi < 0.0)
            amin = gamma_alpha * fmin(gamma_theta, gamma_phi * th / (-gphi));
        else
            amin = gamma_alpha * gamma_theta;
        double a = a_max, ph = ph_cache, ph_acc = 0.0;
        bool accepted = false, ftype = false;
        int nls = 0;
        for (bool base = !ph_valid;;) {
            double tht, pht;
            q.template theta_phi<true>(base ? 0.0 : a, tht, pht);
            if (base) { ph = pht; base = false; continue; }
            ph_acc = pht;
            bool dom = false;
            if (lane < nfilt && tht >= f_th0 && pht >= f_ph0) dom = true;
            if (lane + 32 < nfilt && tht >= f_th1 && pht >= f_ph1) dom = true;
            bool okf = (tht < theta_max) && !__any_sync(0xffffffffu, dom);
            if (okf) {
                bool sw = gphi < 0.0 && la_max - (double)nls * 0.693147180559945309417232 + lsw > 0.0;
                if (th <= theta_min && sw) {
                    if (pht <= ph + eta_phi * a * gphi) { accepted = true; ftype = true; }
                } else if (tht <= (1.0 - gamma_theta) * th || pht <= ph - gamma_phi * th)
                    accepted = true;
            }
            if (accepted) break;
            a *= 0.5;
            nls++;
            n_back++;
            if (!(a >= amin)) break;
        }
        if (!accepted) {
            // IPOPT would call its restoration phase; the rows are elastic, so remove their residual by
            // enlarging the slacks (t' = max(t, s-g), s' = g+t') and restart the filter.
            if (n_reset >= o.max_reset) { status = B200MPC_LINESEARCH; break; }
            n_reset++;
            for (int r = lane; r < R; r += 32) {
                double g = q.GR[r];
                double tn = fmax(q.T[r], q.S[r] - g);
                q.T[r] = tn;
                q.S[r] = g + tn;
            }
            nfilt = 0;
            fpos = 0;
            ph_valid = false;
            iter++;
            __syncwarp();
            continue;
        }
        ph_cache = ph_acc;
        ph_valid = true;
        if (!ftype) {
            double nth = (1.0 - gamma_theta) * th, nph = ph - gamma_phi * th;
            int slot = fpos >> 5, ln = fpos & 31;
            if (lane == ln) {
                if (slot == 0) { f_th0 = nth; f_ph0 = nph; }
                else { f_th1 = nth; f_ph1 = nph; }
            }
            fpos = (fpos + 1) & 63;
            if (nfilt < 64) nfilt++;
        }
        PCLK(5)
        // ---- accept: multipliers of the dynamics by the costate recursion, from
        //      K d + Jc' lam+ = rhs  =>  lam+_i = (rhs - K d)_{x_i} + A' lam+_{i+1}
        // (a) residuals res_i = (rhs - K d)_{x_i} for all stages in parallel -> CRES (dead until the next eval)
        for (int i = lane + 1; i <= N; i += KSTEP) {
            double x[6], d[6], dk[6], g[6] = {0, 0, 0, 0, 0, 0};
            q.template load_x<false>(i, 0.0, x);
            ld6(q.D + 6 * i, d);
            q.diff_target(i, x, dk);
            q.barrier_grad_x(i, x, g);        // barrier terms of vx_i, ey_i only (g[0], g[5])
            double r4 = 0.0, r5 = 0.0;
OCP_ROLL_J
            for (int j = 0; j < M; j++) {
                if (i < N) {
                    int r = j * N + i;
                    double yp = q.YHAT[r] - q.SIGE[r] * q.JD[r];
                    r4 += q.JA[4 * r + 0] * yp;
                    r5 += q.JA[4 * r + 1] * yp;
                }
                int r = j * N + i - 1;
                double yp = q.YHAT[r] - q.SIGE[r] * q.JD[r];
                r4 += q.JA[4 * r + 2] * yp;
                r5 += q.JA[4 * r + 3] * yp;
            }
            if (q.hwd) {   // -(K d) of the ey-rate curvature rows (i-1 as "next", i as "current")
                r5 -= 2.0 * q.df * q.wdp[i - 1] * q.JDC[i - 1];
                if (i < N) r5 += 2.0 * q.df * q.wdp[i] * q.JDC[i];
            }
            const double g0 = g[0], g5 = g[5];
OCP_ROLL_A
            for (int a2 = 0; a2 < 6; a2++) {
                double kd = (q.HD[6 * i + a2] + dw_try) * q.D[6 * i + a2];
                if (q.qdiag()) kd += q.df * kp.Q2[7 * a2] * q.D[6 * i + a2];
                else {
#pragma unroll
                    for (int b = 0; b < 6; b++) kd += q.df * kp.Q2[6 * a2 + b] * d[b];
                }
                double gg = q.grad_x_comp(i, a2, dk, x[5]) + ((a2 == 0) ? g0 : ((a2 == 5) ? g5 : 0.0));
                double res = -gg - kd;
                if (a2 == 4) res += r4;
                if (a2 == 5) res += r5;
                q.CRES[6 * (i - 1) + a2] = res;
            }
        }
        __syncwarp();
        // (b) the recursion itself: lane a < 6 owns component a (its column of A is q.tcol), lam+_{i+1} travels by
        //     shuffles; no barrier inside the loop
        {
            double ln[6];
#pragma unroll
            for (int a2 = 0; a2 < 6; a2++) ln[a2] = 0.0;
            const int comp = (lane < 6) ? lane : 0;
#pragma unroll 1
            for (int i = N; i >= 1; i--) {
                double sres = q.CRES[6 * (i - 1) + comp];
#pragma unroll
                for (int b = 0; b < 6; b++) sres += q.tcol[b] * ln[b];
                if (lane < 6) {
                    double lo = q.LAM[6 * (i - 1) + lane];
                    q.LAM[6 * (i - 1) + lane] = lo + a * (sres - lo);
                }
#pragma unroll
                for (int a2 = 0; a2 < 6; a2++) ln[a2] = __shfl_sync(0xffffffffu, sres, a2);
            }
        }
        // bound multipliers (old point), primal step, kappa_sigma safeguard (new point): lane = stage
        for (int k = lane; k <= N; k += KSTEP) {
            if (k >= 1) {
                double x[6], d[6];
                q.template load_x<false>(k, 0.0, x);
                ld6(q.D + 6 * k, d);
                double2 zl = ld2(q.ZL + q.bsx(k)), zu = ld2(q.ZU + q.bsx(k));
                double zln[2], zun[2];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    double w = c ? x[5] : x[0], dd = c ? d[5] : d[0];
                    double zlc = c ? zl.y : zl.x, zuc = c ? zu.y : zu.x;
                    double lb = q.xlb(k, c), ub = q.xub(k, c);
                    double wn = w + a * dd;
                    zln[c] = 0.0;
                    zun[c] = 0.0;
                    if (IP::has(lb)) {
                        double dl = w - lb, muidln = mu * rcp(wn - lb);
                        zlc += a_z * ((mu - zlc * dd) * rcp(dl) - zlc);
                        zln[c] = fmax(fmin(zlc, kappa_sigma * muidln), muidln * (1.0 / kappa_sigma));
                    }
                    if (IP::has(ub)) {
                        double du = ub - w, muidun = mu * rcp(ub - wn);
                        zuc += a_z * ((mu + zuc * dd) * rcp(du) - zuc);
                        zun[c] = fmax(fmin(zuc, kappa_sigma * muidun), muidun * (1.0 / kappa_sigma));
                    }
                }
                st2(q.ZL + q.bsx(k), zln[0], zln[1]);
                st2(q.ZU + q.bsx(k), zun[0], zun[1]);
#pragma unroll
                for (int a2 = 0; a2 < 6; a2++) x[a2] += a * d[a2];
                st6(q.W + 6 * k, x);
            }
OCP_ROLL_J
            for (int j = 0; j < M; j++) {
                double w = q.W[q.isg(j, k)], dd = q.D[q.isg(j, k)], zlc = q.ZL[q.bss(j, k)];
                zlc += a_z * ((mu - zlc * dd) * rcp(w) - zlc);
                double wn = w + a * dd, muiwn = mu * rcp(wn);
                q.ZL[q.bss(j, k)] = fmax(fmin(zlc, kappa_sigma * muiwn), muiwn * (1.0 / kappa_sigma));
                q.W[q.isg(j, k)] = wn;
            }
            if (k < N) {
                double u[2];
                q.template load_u<false>(k, 0.0, u);
                double2 dd2 = ld2(q.D + OU + 2 * k);
                double2 zl = ld2(q.ZL + q.bsu(k)), zu = ld2(q.ZU + q.bsu(k));
                double zln[2], zun[2], un[2];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    double w = u[c], dd = c ? dd2.y : dd2.x, zlc = c ? zl.y : zl.x, zuc = c ? zu.y : zu.x;
                    double dl = w + kp.p.umax[c], du = kp.p.umax[c] - w;
                    zlc += a_z * ((mu - zlc * dd) * rcp(dl) - zlc);
                    zuc += a_z * ((mu + zuc * dd) * rcp(du) - zuc);
                    double wn = w + a * dd, muidln = mu * rcp(wn + kp.p.umax[c]), muidun = mu * rcp(kp.p.umax[c] - wn);
                    zln[c] = fmax(fmin(zlc, kappa_sigma * muidln), muidln * (1.0 / kappa_sigma));
                    zun[c] = fmax(fmin(zuc, kappa_sigma * muidun), muidun * (1.0 / kappa_sigma));
                    un[c] = wn;
                }
                st2(q.ZL + q.bsu(k), zln[0], zln[1]);
                st2(q.ZU + q.bsu(k), zun[0], zun[1]);
                st2(q.W + OU + 2 * k, un[0], un[1]);
OCP_ROLL_J
                for (int j = 0; j < M; j++) {
                    int r = j * N + k;
                    double ds, dt, dy, dz, dv;
                    q.row_step(r, ds, dt, dy, dz, dv);
                    double s = q.S[r] + a * ds, tt = q.T[r] + a * dt;
                    q.Y[r] += a * dy;
                    double z = q.Z[r] + a_z * dz, v = q.V[r] + a_z * dv;
                    double muis = mu * rcp(s), muit = mu * rcp(tt);
                    q.Z[r] = fmax(fmin(z, kappa_sigma * muis), muis * (1.0 / kappa_sigma));
                    q.V[r] = fmax(fmin(v, kappa_sigma * muit), muit * (1.0 / kappa_sigma));
                    q.S[r] = s;
                    q.T[r] = tt;
                }
            }
        }
        __syncwarp();
        iter++;
        PCLK(6)
    }
#ifdef B200MPC_PHASE_CLOCKS
    // profiling build: phase cycle counters replace x_pred (first 8 doubles of the instance's slot)
    if (xpred != nullptr && lane == 0)
        for (int k = 0; k < 13; k++) xpred[(size_t)inst * 6 * (N + 1) + k] = (k < 8) ? (double)pc[k] : (double)q.bc[k - 8];
    xpred = nullptr;
#endif

    // ---- results
    // control.py:582-586 (and :228-232) impose the bound rows on stage 0 as well, where x_0 is fixed (:497): an x_0 outside
    // them (beyond IPOPT's constr_viol_tol, 1e-4) makes the reference's NLP infeasible.  What was solved above is the problem
    // without the stage-0 rows; its solution is returned, flagged.
    {
        const double ctol = 1e-4, vx0 = q.IN[0], ey0 = q.IN[5];
        if (vx0 < q.xlb(0, 0) - ctol || vx0 > q.xub(0, 0) + ctol || ey0 < q.xlb(0, 1) - ctol || ey0 > q.xub(0, 1) + ctol)
            status = B200MPC_INFEASIBLE_X0;
    }
    double cost = q.objective();
    double tm = 0.0;
    for (int r = lane; r < R; r += 32) tm = fmax(tm, q.T[r]);
    tm = warp_max_nn(tm);
    {
        b200mpc_record rc;
        rc.cost = cost;
        rc.u0[0] = q.W[OU];
        rc.u0[1] = q.W[OU + 1];
        rc.status = status;
        rc.iters = iter;
        if (lane == 0) rec[inst] = rc;
        xchg_publish(xa, lane, inst, rc);   // sharded batch: lane p stores the record into rank p's gathered buffer (exchange.cuh)
    }
    if (aux != nullptr && lane < 4) {
        double v = (lane == 0) ? E0 : (lane == 1) ? tm : (lane == 2) ? (double)n_refac : (double)n_back;
        aux[(size_t)inst * 4 + lane] = v;
    }
    if (xpred != nullptr)
        for (int e = lane; e < 6 * (N + 1); e += 32) xpred[(size_t)inst * 6 * (N + 1) + e] = q.W[e];
    if (upred != nullptr)
        for (int e = lane; e < 2 * N; e += 32) upred[(size_t)inst * 2 * N + e] = q.W[OU + e];
    if (sigma != nullptr)
        for (int e = lane; e < M * (N + 1); e += 32) sigma[(size_t)inst * M * (N + 1) + e] = q.W[OS + e];
}

}  // namespace b200mpc

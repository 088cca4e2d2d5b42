// b200mpc: the exchange step of the sharded solve, fused into the solver kernels' epilogue.
//
// The reference's planner forks one process per candidate, joins them and takes the argmin of the gathered costs
// (planning/overtake_traj_planner.py:177-204, :244).  With the batch sharded over G GPUs (one process per GPU) the join is
// an all-gather of the 32-byte result records.  Here that is not a collective call after the solve: every rank owns a
// WINDOW in its HBM that its peers map (CUDA IPC over NVLink/NVSwitch peer access), and the warp that finishes an instance
// stores the record straight into every rank's gathered buffer (lane p -> peer p: one 32-byte store over NVLink each) and
// bumps an arrival counter there.  No NCCL kernel competes with the solver grids for SM slots and nothing waits for the
// slowest instance of the slowest rank before the transfer starts.
//
// Window of one rank (device memory, S slots so that S steps can be in flight):
//   rec[S][world * max_batch]   gathered records, rank r's shard at [r * B, (r + 1) * B)
//   cnt[S][world]  (u64)        records of rank r that have arrived for slot s, cumulative over the uses of the slot
//   ack[S][world]  (u64)        uses of slot s that rank r has consumed (its argmin kernel has read its own copy)
// Protocol for use number n = 1, 2, ... of slot s with B records per rank:
//   gate (xchg_wait_ack_kernel, one warp, on the solver's stream BEFORE the solver grid of use n):
//                                        wait own.ack[s][p] >= n - 1 for all p  (every rank has consumed the previous use of the
//                                        slot: flow control, a fast rank cannot lap a slow one)
//   producer warp (rank q, any peer p):  peer[p].rec[s][q * B + i] = record;  peer[p].cnt[s][q] += 1 (release, system scope)
//   consumer (xchg_argmin_kernel, one warp, rank p): wait own.cnt[s][q] >= cumulative count for all q;  argmin (+ copy-out);
//                                        peer[q].ack[s][p] = n (release)  for all q
// Dead-lock freedom: the gate of use n only waits for consumers of use n - 1, which only wait for producers of use n - 1;
// the uses of one slot are issued in stream order on every rank.  Solver CTAs never spin (a spinning CTA would hold its SM
// slot), and the two spinning kernels are ONE warp each with < 40 registers, which fits beside a full complement of solver
// CTAs on an SM (8 x 7936 + 1280 registers <= 65536, 4.7 KB of shared memory to spare): they can always be placed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200mpc.h"

namespace b200mpc {

constexpr int XCHG_MAX_WORLD = 16;

// per-rank table in device memory: where every rank's window lives in THIS rank's address space
struct XchgTable {
    int32_t rank, world, max_batch, slots;
    b200mpc_record *rec[XCHG_MAX_WORLD];           // peer p's rec[0][0]
    unsigned long long *cnt[XCHG_MAX_WORLD];       // peer p's cnt[0][0]
    unsigned long long *ack[XCHG_MAX_WORLD];       // peer p's ack[0][0]
};

// kernel argument of the solver kernels (by value); tab == nullptr: no exchange
struct XchgArgs {
    const XchgTable *tab;
    int32_t slot;
    int32_t B;                      // records per rank in this use
    unsigned long long use;         // n
};

#ifdef B200MPC_HOST_EMULATION
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ void add_release_sys(unsigned long long *p, unsigned long long v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ void xchg_backoff() {}
__device__ __forceinline__ unsigned long long xchg_now_ns() { return 0ull; }
#else
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void add_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void xchg_backoff() { __nanosleep(200); }
__device__ __forceinline__ unsigned long long xchg_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif
// A peer that never arrives (a crashed rank, streams aliased onto one hardware queue -- see b200mpc_comm_create) must not
// hang the GPU: every spin gives up after this long; the consumer then reports -2 instead of an index.
constexpr unsigned long long XCHG_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

// Epilogue of a solver warp: the lanes of one warp call this convergently; `rc` is the instance's record (valid in every
// lane).  Lane p < world stores it into rank p's gathered buffer and signals its arrival there.
__device__ __forceinline__ void xchg_publish(const XchgArgs &xa, int lane, int inst, const b200mpc_record &rc) {
    if (xa.tab == nullptr) return;
    const XchgTable &t = *xa.tab;
    for (int p = lane; p < t.world; p += 32) {
        b200mpc_record *dst = t.rec[p] + ((size_t)xa.slot * t.world + t.rank) * t.max_batch + inst;
#ifdef B200MPC_HOST_EMULATION
        *dst = rc;
#else
        // one 32-byte record = two 16-byte stores (the window is 32-byte aligned)
        double2 lo = make_double2(rc.cost, rc.u0[0]);
        double2 hi = make_double2(rc.u0[1], __hiloint2double(rc.iters, rc.status));
        reinterpret_cast<double2 *>(dst)[0] = lo;
        reinterpret_cast<double2 *>(dst)[1] = hi;
#endif
        add_release_sys(t.cnt[p] + (size_t)xa.slot * t.world + t.rank, 1ull);   // release: the record is visible before the count
    }
}

// Consumer: one warp.  Waits until every rank's `B` records of this use have arrived in this rank's window, takes the
// first-min argmin over the world * B gathered records (rank-major = instance order of the global batch; lowest index
// wins ties, overtake_traj_planner.py:244), optionally copies them out, then acknowledges the use to every peer.
// want[q] = cumulative arrival count expected from rank q (host-side bookkeeping: sum of B over the uses of the slot).
// Gate of a use: one warp on the solver's stream ahead of the solver grid.
static __global__ void __launch_bounds__(32) xchg_wait_ack_kernel(const XchgTable *__restrict__ tab, int slot, unsigned long long use) {
    const XchgTable &t = *tab;
    for (int p = threadIdx.x; p < t.world; p += 32) {
        const unsigned long long *ack = t.ack[t.rank] + (size_t)slot * t.world + p;   // rank p's acks land in OUR window
        const unsigned long long t0 = xchg_now_ns();
        while (ld_acquire_sys(ack) + 1 < use) {
            xchg_backoff();
            if (xchg_now_ns() - t0 > XCHG_TIMEOUT_NS) break;
        }
    }
}

static __global__ void __launch_bounds__(32) xchg_argmin_kernel(const XchgTable *__restrict__ tab, int slot, int B, unsigned long long want, unsigned long long use,
                                   int max_status, int32_t *__restrict__ out, b200mpc_record *__restrict__ copy_out) {
    const XchgTable &t = *tab;
    const int world = t.world;
    bool late = false;
    for (int q = threadIdx.x; q < world; q += 32) {
        const unsigned long long *c = t.cnt[t.rank] + (size_t)slot * world + q;
        const unsigned long long t0 = xchg_now_ns();
        while (ld_acquire_sys(c) < want) {
            xchg_backoff();
            if (xchg_now_ns() - t0 > XCHG_TIMEOUT_NS) { late = true; break; }
        }
    }
    late = __any_sync(0xffffffffu, late);
    const b200mpc_record *base = t.rec[t.rank] + (size_t)slot * world * t.max_batch;
    double best = 1e300 * 1e300;
    int bi = 0x7fffffff;
    // one warp reads world * B records: 4 independent 32-byte loads in flight per lane (the loop is latency bound)
#pragma unroll 4
    for (int i = threadIdx.x; i < world * B; i += 32) {
        const int q = i / B, k = i - q * B;
        const b200mpc_record r = base[(size_t)q * t.max_batch + k];
        if (copy_out != nullptr) copy_out[i] = r;
        if (r.status <= max_status && (r.cost < best || (r.cost == best && i < bi))) { best = r.cost; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    __syncwarp();      // every lane has finished reading the slot (and its copy-out stores are issued)
    if (threadIdx.x == 0 && out != nullptr) *out = late ? -2 : ((bi == 0x7fffffff) ? -1 : bi);   // -2: a rank's records never arrived
    for (int q = threadIdx.x; q < world; q += 32) st_release_sys(t.ack[q] + (size_t)slot * world + t.rank, use);
}

}  // namespace b200mpc

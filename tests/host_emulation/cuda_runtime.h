/* TEST INFRASTRUCTURE: a stand-in for <cuda_runtime.h> that lets g++ compile the per-region body of
 * car_racing_b200/csrc/planner_prepare.cuh as host code (tests/test_planner_prepare.py checks its LOGIC against the
 * reference golden on machines without a GPU).  Nothing on the product path includes this file. */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#ifdef EMU_DYNAMIC_SMEM_ONLY
/* translation units whose kernels use `extern __shared__ double sm[]`: the declaration becomes `extern double sm[]` and the
 * harness defines b200mpc::sm (one block in flight at a time) */
#define __shared__
#define __align__(x)
#else
#define __shared__ static
#endif
#define __launch_bounds__(...)
struct EmuIdx { int x; };
/* one host thread per CUDA thread of a block (emu_launch below); single-threaded callers leave the defaults */
static thread_local EmuIdx threadIdx = {0}, blockIdx = {0};
static EmuIdx blockDim = {1}, gridDim = {1};
static void (*emu_barrier_fn)() = nullptr;
static inline void __syncthreads() { if (emu_barrier_fn) emu_barrier_fn(); }
static inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
/* compiled with -ffp-contract=off: plain operators are the separately rounded operations */
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }

#ifdef EMU_WITH_LAUNCH
#include <barrier>
#include <cstdlib>
#include <functional>
#include <thread>
#include <vector>
#endif

#ifdef EMU_WARPS
/* warp collectives for kernels whose CTA is one or more full warps executing convergently: every lane is a host thread,
 * a collective = publish, barrier, read, barrier on the warp's own std::barrier.  A collective inside divergent code would
 * dead-lock here (the tests run under a timeout); on the GPU the same code would be undefined behaviour with a full mask. */
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct EmuWarp {
    std::barrier<> bar{32};
    uint64_t xch[32];
};
static EmuWarp *emu_warps = nullptr;    // one per warp of the block in flight
static inline EmuWarp &emu_warp() { return emu_warps[threadIdx.x >> 5]; }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp().bar.arrive_and_wait(); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    EmuWarp &w = emu_warp();
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.xch[threadIdx.x & 31] = raw;
    w.bar.arrive_and_wait();
    raw = w.xch[src & 31];
    w.bar.arrive_and_wait();
    T r;
    memcpy(&r, &raw, sizeof(T));
    return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int lanemask) { return __shfl_sync(m, v, (threadIdx.x & 31) ^ lanemask); }
template <typename F> static inline uint64_t emu_warp_fold(uint64_t mine, F f) {
    EmuWarp &w = emu_warp();
    w.xch[threadIdx.x & 31] = mine;
    w.bar.arrive_and_wait();
    uint64_t acc = w.xch[0];
    for (int l = 1; l < 32; l++) acc = f(acc, w.xch[l]);
    w.bar.arrive_and_wait();
    return acc;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) { return (unsigned)emu_warp_fold(v, [](uint64_t a, uint64_t b) { return a > b ? a : b; }); }
static inline unsigned __reduce_min_sync(unsigned, unsigned v) { return (unsigned)emu_warp_fold(v, [](uint64_t a, uint64_t b) { return a < b ? a : b; }); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    return (unsigned)emu_warp_fold(pred ? (1ull << (threadIdx.x & 31)) : 0ull, [](uint64_t a, uint64_t b) { return a | b; });
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __double2hiint(double d) { uint64_t r; memcpy(&r, &d, 8); return (int)(r >> 32); }
static inline int __double2loint(double d) { uint64_t r; memcpy(&r, &d, 8); return (int)(r & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo) { uint64_t r = ((uint64_t)(unsigned)hi << 32) | (unsigned)lo; double d; memcpy(&d, &r, 8); return d; }
#endif

#ifdef EMU_WITH_LAUNCH
/* blocks run one after the other; the threads of a block are real host threads, __syncthreads() is a std::barrier and
 * __shared__ variables are function-local statics (shared by the block's threads, reused by the next block) */
static std::barrier<> *emu_block_barrier = nullptr;
static void emu_barrier_wait() { emu_block_barrier->arrive_and_wait(); }
/* dynamic shared memory of the block in flight: allocated per block with exactly the launch's size (AddressSanitizer then
 * sees out-of-bounds shared-memory accesses) and filled with 0xFF bytes = NaN doubles, so that a kernel that reads shared
 * memory it has not written -- garbage on the GPU -- cannot pass a parity test here by luck */
static double *emu_dynamic_smem = nullptr;
static inline void emu_launch(int grid, int block, size_t smem_bytes, const std::function<void()> &kernel) {
    gridDim.x = grid;
    blockDim.x = block;
    for (int b = 0; b < grid; b++) {
        const size_t bytes = (smem_bytes + 15) & ~(size_t)15;
        emu_dynamic_smem = bytes ? (double *)aligned_alloc(16, bytes) : nullptr;
        if (emu_dynamic_smem) memset(emu_dynamic_smem, 0xFF, bytes);
        std::barrier<> bar(block);
        emu_block_barrier = &bar;
        emu_barrier_fn = emu_barrier_wait;
#ifdef EMU_WARPS
        std::vector<EmuWarp> warps((block + 31) / 32);
        emu_warps = warps.data();
#endif
        std::vector<std::thread> ts;
        for (int t = 0; t < block; t++)
            ts.emplace_back([&, t]() {
                threadIdx.x = t;
                blockIdx.x = b;
                kernel();
            });
        for (auto &th : ts) th.join();
        emu_barrier_fn = nullptr;
        free(emu_dynamic_smem);
        emu_dynamic_smem = nullptr;
    }
    gridDim.x = 1;
    blockDim.x = 1;
}
#endif


#ifdef EMU_RUNTIME_API
/* the slice of the CUDA runtime API that car_racing_b200/csrc/capi.cu uses, on host memory: "device" buffers are malloc'ed,
 * copies are memcpy, streams are tokens (everything is synchronous), there is exactly one "device" */
#include <cstdlib>
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, int attr, int) {
    *v = (attr == cudaDevAttrMultiProcessorCount) ? 148 : 227 * 1024;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height,
                                            cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < height; r++) memcpy((char *)d + r * dpitch, (const char *)s + r * spitch, width);
    return cudaSuccess;
}
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
/* `extern __shared__ double sm[]` of the kernels is rewritten to `double *sm = emu_dynamic_smem;` by build_emu_library.py
 * (emu_launch above allocates it per block) */
#endif

/* TEST INFRASTRUCTURE: a stand-in for <cuda_runtime.h> that lets g++ compile the per-region body of
 * car_racing_b200/csrc/planner_prepare.cuh as host code (tests/test_planner_prepare.py checks its LOGIC against the
 * reference golden on machines without a GPU).  Nothing on the product path includes this file. */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#define __shared__ static
#define __launch_bounds__(x)
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
#include <functional>
#include <thread>
#include <vector>
/* blocks run one after the other; the threads of a block are real host threads, __syncthreads() is a std::barrier and
 * __shared__ variables are function-local statics (shared by the block's threads, reused by the next block) */
static std::barrier<> *emu_block_barrier = nullptr;
static void emu_barrier_wait() { emu_block_barrier->arrive_and_wait(); }
static inline void emu_launch(int grid, int block, const std::function<void()> &kernel) {
    gridDim.x = grid;
    blockDim.x = block;
    for (int b = 0; b < grid; b++) {
        std::barrier<> bar(block);
        emu_block_barrier = &bar;
        emu_barrier_fn = emu_barrier_wait;
        std::vector<std::thread> ts;
        for (int t = 0; t < block; t++)
            ts.emplace_back([&, t]() {
                threadIdx.x = t;
                blockIdx.x = b;
                kernel();
            });
        for (auto &th : ts) th.join();
        emu_barrier_fn = nullptr;
    }
    gridDim.x = 1;
    blockDim.x = 1;
}
#endif

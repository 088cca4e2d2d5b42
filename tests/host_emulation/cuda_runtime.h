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
static EmuIdx threadIdx = {0}, blockIdx = {0}, blockDim = {1};
static inline void __syncthreads() {}
static inline int atomicOr(int *p, int v) { int o = *p; *p |= v; return o; }
/* compiled with -ffp-contract=off: plain operators are the separately rounded operations */
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }

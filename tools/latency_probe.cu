// Dependent-issue latency probe (evidence for DESIGN.md section 5): ONE warp on one SM runs chains of dependent
// instructions of the kinds the solver's critical paths are made of; clock64() around 512 links gives cycles per link.
// A second set of numbers runs 2 / 4 independent chains in the same warp (how much ILP the FP64 pipe takes from one
// warp), and K warps on one scheduler each running one chain (what a second resident warp adds).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_probe tools/latency_probe.cu && ./latency_probe
#include <cstdio>
#include <cuda_runtime.h>

#define LINKS 512

__device__ __forceinline__ double rcp_seed(double x) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ __forceinline__ double rsq_seed(double x) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }

// kind: 0 DFMA, 1 DMUL, 2 DADD, 3 MUFU.RCP64H seed, 4 MUFU.RSQ64H seed, 5 shuffle of a double, 6 STS -> syncwarp -> LDS,
//       7 LDS (pointer chase through shared memory), 8 REDUX (integer warp max)
template <int KIND, int CHAINS>
__global__ void probe(double *out, long long *cycles, double a, double b) {
    __shared__ double sm[64];
    __shared__ int nxt[32];
    const int lane = threadIdx.x & 31;
    sm[lane] = a + lane;
    sm[32 + lane] = b;
    nxt[lane] = (lane * 5 + 3) & 31;
    __syncthreads();
    double r[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) r[c] = a + c + lane * 1e-3;
    int p = lane;
    unsigned u = lane + 1;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < LINKS / 16; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                if (KIND == 0) r[c] = fma(r[c], b, a);
                if (KIND == 1) r[c] = r[c] * b;
                if (KIND == 2) r[c] = r[c] + b;
                if (KIND == 3) r[c] = rcp_seed(r[c]);
                if (KIND == 4) r[c] = rsq_seed(r[c]);
                if (KIND == 5) r[c] = __shfl_sync(0xffffffffu, r[c], (lane + 1) & 31);
                if (KIND == 6) { sm[lane] = r[c]; __syncwarp(); r[c] = sm[(lane + 1) & 31]; __syncwarp(); }
                if (KIND == 7) p = nxt[p];
                if (KIND == 8) u = __reduce_max_sync(0xffffffffu, u) + lane;
            }
        }
    }
    long long t1 = clock64();
    double s = p + u;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += r[c];
    if (s == 12345.678) out[0] = s;
    if (lane == 0) cycles[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
}

template <int KIND, int CHAINS>
static void run(const char *name, int warps) {
    double *out;
    long long *cyc, h[64];
    cudaMalloc(&out, 8);
    cudaMalloc(&cyc, sizeof(h));
    // warps of one CTA: 4 per scheduler round-robin, so `warps` = 4 k puts k warps on every scheduler
    probe<KIND, CHAINS><<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999);
    probe<KIND, CHAINS><<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999);
    cudaMemcpy(h, cyc, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < warps; i++) mx = h[i] > mx ? h[i] : mx;
    printf("%-34s chains/warp %d  warps %2d : %7.2f cycles per link  (%.3f links/clk/SM)\n", name, CHAINS, warps,
           (double)mx / LINKS, (double)LINKS * CHAINS * warps / mx);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0, 1>("DFMA dependent", 1);
    run<0, 2>("DFMA", 1);
    run<0, 4>("DFMA", 1);
    run<0, 8>("DFMA", 1);
    run<0, 1>("DFMA dependent", 4);
    run<0, 1>("DFMA dependent", 8);
    run<0, 4>("DFMA", 8);
    run<0, 8>("DFMA", 8);
    run<1, 1>("DMUL dependent", 1);
    run<2, 1>("DADD dependent", 1);
    run<3, 1>("MUFU.RCP64H dependent", 1);
    run<4, 1>("MUFU.RSQ64H dependent", 1);
    run<5, 1>("SHFL (double = 2 x SHFL) dependent", 1);
    run<6, 1>("STS.64 syncwarp LDS.64 syncwarp", 1);
    run<7, 1>("LDS pointer chase", 1);
    run<8, 1>("REDUX.MAX dependent", 1);
    return 0;
}

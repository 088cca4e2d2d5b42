// Instruction-delivery probe for the SM (evidence for DESIGN.md section 5): a loop whose body is BODY independent FP64
// FMAs (straight-line, 16 B per instruction) run by W warps per SM.  If the body fits the per-scheduler L0 instruction
// cache every scheduler issues from its own L0; once it does not, all schedulers of the SM fetch through the shared
// L1.5 and the SM's total issue rate is capped by that path.  Prints warp-instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_probe tools/icache_probe.cu && ./icache_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int BODY>
__global__ void probe(double *out, int iters, double a, double b, int desync) {
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = a + i + threadIdx.x;
    // de-synchronise the warps of an SM (in the solver they sit in different phases of the iteration)
    long long t0 = clock64();
    while (desync && clock64() - t0 < 1777LL * (threadIdx.x >> 5)) { }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < BODY; k++) r[k & 7] = fma(r[k & 7], b, a);   // 8 independent chains
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += r[i];
    if (s == 12345.678) out[0] = s;
}

template <int BODY>
static void run(int warps_per_sm, int sms, double clock_ghz, int desync = 1) {
    double *out;
    cudaMalloc(&out, 8);
    int iters = (1 << 22) / BODY;
    dim3 grid(sms), block(32 * warps_per_sm);
    probe<BODY><<<grid, block>>>(out, 16, 1.0, 0.999, desync);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<BODY><<<grid, block>>>(out, iters, 1.0, 0.999, desync);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double instr = (double)iters * BODY * warps_per_sm;   // per SM
    printf("%s body %5d instr (%6.1f KB)  warps/SM %2d  : %.3f warp-instr/clk/SM  (%.3f per warp)\n",
           desync ? "desync" : "lockstep", BODY, BODY * 16 / 1024.0, warps_per_sm, instr / (ms * 1e-3 * clock_ghz * 1e9), instr / (ms * 1e-3 * clock_ghz * 1e9) / warps_per_sm);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double ghz = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz (nominal)\n", p.name, p.multiProcessorCount, ghz);
    int sms = p.multiProcessorCount;
    for (int w : {1, 4, 7}) {
        run<1024>(w, sms, ghz);
        run<2048>(w, sms, ghz);
        run<2560>(w, sms, ghz);
        run<3072>(w, sms, ghz);
        run<3584>(w, sms, ghz);
        run<4096>(w, sms, ghz);
        run<5120>(w, sms, ghz);
        run<6144>(w, sms, ghz);
        run<8192>(w, sms, ghz);
    }
    for (int w : {4, 7}) {   // warps in lockstep share every fetched line
        run<4096>(w, sms, ghz, 0);
        run<6144>(w, sms, ghz, 0);
        run<8192>(w, sms, ghz, 0);
    }
    return 0;
}

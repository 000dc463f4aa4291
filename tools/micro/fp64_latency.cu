// fp64_latency.cu -- dependent-issue latency of the FP64 instructions the trajectory kernels are made of (B200, sm_100a).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu ; run: ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double* out, long long* cycles, int iters, double a, double b) {
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            if (OP == 0) x = fma(x, b, a);
            if (OP == 1) x = x + b;
            if (OP == 2) x = x * b;
            if (OP == 3) x = 1.0 / x + a;      // division (+1 add)
            if (OP == 4) x = sqrt(x) + a;      // sqrt (+1 add)
            if (OP == 5) x = rsqrt(x) + a;     // rsqrt (+1 add)
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cycles[0] = t1 - t0; }
    if (x == 123.456) out[0] = x;
}

// W warps on one SM issuing independent DFMA chains: throughput per SM vs warps
template <int ILP>
__global__ void tput(double* out, long long* cycles, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + threadIdx.x * 1e-9 + k;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], b, a);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    double s = 0; for (int k = 0; k < ILP; ++k) s += x[k];
    if (s == 123.456) out[0] = s;
}

int main() {
    double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    long long h;
    const char* names[] = {"DFMA", "DADD", "DMUL", "1/x + add", "sqrt + add", "rsqrt + add"};
    const int iters = 256;
#define RUN(OP) chain<OP><<<1, 32>>>(d, c, iters, 1.0000001, 0.9999999); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("%-12s dependent latency: %.1f cycles\n", names[OP], (double)h / (iters * 32.0));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
#define TP(ILP, WARPS) tput<ILP><<<1, 32 * WARPS>>>(d, c, iters, 1.0000001, 0.9999999); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("warps/SM %2d  ILP %d : %.2f DFMA warp-instr / cycle / SM (peak 2.0)\n", WARPS, ILP, (double)iters * 16 * ILP * WARPS / (double)h);
    TP(1, 4) TP(2, 4) TP(4, 4) TP(8, 4) TP(1, 8) TP(2, 8) TP(4, 8) TP(1, 16) TP(2, 16) TP(4, 16) TP(1, 32)
    return 0;
}

// dmma_peak.cu -- FP64 tensor-core (DMMA) vs DFMA throughput microbenchmark on B200 (sm_100a).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_peak dmma_peak.cu ; run: ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9 + i; c[i][1] = 0.5 * i; }
    const double a = 1.0 + 1e-12 * threadIdx.x, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9 + i; c[i][1] = 0.5 * i; c[i][2] = 1.0; c[i][3] = 2.0; }
    const double a[4] = {1.0, 1.0 + 1e-12 * threadIdx.x, 0.5, 0.25};
    const double b[2] = {1e-9, 2e-9};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0 + i * 1e-3 + threadIdx.x * 1e-6;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
}

template <class F>
double time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    double* d; cudaMalloc(&d, 8);
    const int sms = prop.multiProcessorCount, iters = 1 << 13;
    for (int wps : {4, 8, 16, 32}) {     // warps per SM
        const int blocks = sms * (wps * 32 / 256 > 0 ? wps * 32 / 256 : 1), threads = wps * 32 >= 256 ? 256 : wps * 32;
        const double warps = (double)blocks * threads / 32;
        double ms = time_ms([&] { k_dmma884<8><<<blocks, threads>>>(d, iters); });
        printf("warps/SM %2d  dmma m8n8k4  x8 acc : %7.2f TFLOP/s\n", wps, warps * iters * 8 * 512.0 / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_dmma884<16><<<blocks, threads>>>(d, iters); });
        printf("warps/SM %2d  dmma m8n8k4  x16 acc: %7.2f TFLOP/s\n", wps, warps * iters * 16 * 512.0 / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_dmma1688<4><<<blocks, threads>>>(d, iters); });
        printf("warps/SM %2d  dmma m16n8k8 x4 acc : %7.2f TFLOP/s\n", wps, warps * iters * 4 * 2048.0 / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_dmma1688<8><<<blocks, threads>>>(d, iters); });
        printf("warps/SM %2d  dmma m16n8k8 x8 acc : %7.2f TFLOP/s\n", wps, warps * iters * 8 * 2048.0 / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_dfma<<<blocks, threads>>>(d, iters); });
        printf("warps/SM %2d  dfma x16 chains     : %7.2f TFLOP/s\n", wps, (double)blocks * threads * iters * 16 * 2.0 / (ms * 1e-3) / 1e12);
    }
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
    return 0;
}

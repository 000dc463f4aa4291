// mathcheck.cu -- accuracy of the branch-free device helpers (exp_nb, rsqrt_pos, div_nb) against the CUDA math library.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -I nqcdynamics.jl_b200/csrc -o gpurun_out/mathcheck tools/micro/mathcheck.cu
#include <cstdio>
#include <cstdint>
#include "linalg.cuh"
#include "philox.cuh"

using namespace nq;

__device__ double ulps(double got, double ref) {
    if (ref == 0.0) return fabs(got) > 0 ? 1e300 : 0.0;
    const double u = fabs(ref) * 1.1102230246251565e-16;   // half an ulp at most
    return fabs(got - ref) / (2.0 * u);
}

__global__ void check(double* out, int reps) {
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double e_exp = 0, e_rs = 0, e_div = 0;
    for (int i = 0; i < reps; ++i) {
        const double u0 = philox_uniform(12345ull, id, (uint64_t)i, 0u), u1 = philox_uniform(12345ull, id, (uint64_t)i, 1u);
        const double x = (u0 - 0.5) * 1400.0;                 // exp argument in [-700, 700]
        e_exp = fmax(e_exp, ulps(exp_nb(x), exp(x)));
        const double xs = (u0 - 0.5) * 4.0;                   // and where the models live
        e_exp = fmax(e_exp, ulps(exp_nb(xs), exp(xs)));
        const double y = exp((u1 - 0.5) * 600.0);             // rsqrt argument over 260 decades
        e_rs = fmax(e_rs, ulps(rsqrt_pos(y), rsqrt(y)));
        const double b = 1.0 + 40000.0 * u1, a = (u0 - 0.5) * 10.0;
        e_div = fmax(e_div, ulps(div_nb(a, b, 1.0 / b), a / b));
    }
    out[3 * id + 0] = e_exp; out[3 * id + 1] = e_rs; out[3 * id + 2] = e_div;
}

int main() {
    const int blocks = 592, threads = 128, reps = 2000;
    double* d;
    cudaMalloc(&d, sizeof(double) * 3 * blocks * threads);
    check<<<blocks, threads>>>(d, reps);
    double* h = new double[3 * blocks * threads];
    if (cudaMemcpy(h, d, sizeof(double) * 3 * blocks * threads, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
    double m[3] = {0, 0, 0};
    for (int i = 0; i < blocks * threads; ++i) for (int k = 0; k < 3; ++k) m[k] = h[3 * i + k] > m[k] ? h[3 * i + k] : m[k];
    printf("max error in ulps of the library result over %d samples: exp_nb %.3f  rsqrt_pos %.3f  div_nb %.3f\n", blocks * threads * reps, m[0], m[1], m[2]);
    return (m[0] < 2.0 && m[1] < 2.0 && m[2] < 1.0) ? 0 : 2;
}

// philox.cuh -- counter-based Philox4x32-10 (Salmon et al., SC'11).
// One uniform draw per (seed; global trajectory id, step, purpose): replaces the reference's
// task-local `rand()` in select_new_state (fssh.jl:112) / iesh_check_hop! (iesh.jl:393) with a
// stream that does not depend on how trajectories are sharded over blocks or GPUs.
#pragma once
#include "common.cuh"

namespace nq {

NQ_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Raw block: counter (gid, step, purpose) under key seed -> four 32-bit words.
NQ_HD void philox4x32(uint64_t seed, uint64_t gid, uint64_t step, uint32_t purpose, uint32_t (&out)[4]) {
    uint32_t c0 = (uint32_t)gid, c1 = (uint32_t)(gid >> 32), c2 = (uint32_t)step,
             c3 = ((uint32_t)(step >> 32) & 0x00FFFFFFu) | (purpose << 24);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

NQ_HD double philox_uniform(uint64_t seed, uint64_t gid, uint64_t step, uint32_t purpose) {
    uint32_t w[4];
    philox4x32(seed, gid, step, purpose, w);
    uint64_t bits = ((uint64_t)w[0] << 32) | w[1];
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);  // [0, 1) with 53 random bits
}

// Two independent standard normals per block (Box-Muller), used by the device-side initial-condition sampler
// (nqcb200_sample_state): counter = (global trajectory id, component, purpose 2).
NQ_HD void philox_normal2(uint64_t seed, uint64_t gid, uint64_t comp, double& z0, double& z1, uint32_t purpose = 2u) {
    uint32_t w[4];
    philox4x32(seed, gid, comp, purpose, w);
    const double u1 = (double)(((((uint64_t)w[0] << 32) | w[1]) >> 11) + 1ull) * (1.0 / 9007199254740992.0);   // (0, 1]
    const double u2 = (double)((((uint64_t)w[2] << 32) | w[3]) >> 11) * (1.0 / 9007199254740992.0);           // [0, 1)
    const double rad = sqrt(-2.0 * log(u1));
#if defined(__CUDA_ARCH__)
    double s, c;
    sincospi(2.0 * u2, &s, &c);
#else
    const double s = sin(6.283185307179586476925286766559 * u2), c = cos(6.283185307179586476925286766559 * u2);
#endif
    z0 = rad * c; z1 = rad * s;
}

}  // namespace nq

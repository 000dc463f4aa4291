// kernel_nrpmd.cuh -- NRPMD (MMST mapping variables per bead + ring polymer), beads on lanes.
//
// Reference restated:
//   RingPolymerMInt.perform_step!  src/DynamicsMethods/IntegrationAlgorithms/ringpolymer_mint.jl:28-78
//     half Cayley (built with half=true, :22) -> update_cache! + Vbar, Dbar, traceless adiabatic derivative (:45-48)
//     -> propagate_mapping_variables! (:80-94, C = Z cos(l dt) Z', D = Z sin(-l dt) Z')
//     -> nuclear kick from Gamma / Xi (:52-70, :107-130) -> half Cayley (:72-76)
//   estimators: diabatic_population nrpmd.jl:111-122, classical_potential_energy nrpmd.jl:124-139
//
// The reference builds C, D, E = Z Gamma Z', F = Z Xi Z' as dense n x n matrices per (dof, atom, bead);
// here the mapping variables are rotated into the adiabatic basis once per bead (qa = Z'q, pa = Z'p),
// where C, D are diagonal and the force is 1/2 (qa'G qa + pa'G pa) - qa' Xi pa -- same numbers, no
// similarity transforms.  Results do not depend on the eigenvector gauge, so none is carried.
#pragma once
#include "kernel_ring.cuh"

namespace nq {

#if defined(__CUDACC__)

template <int N, int NB>
NQ_D void nrpmd_record_save(const KParams& p, Emitter& em, int lane, int group_base, double r, double v, double mass,
                            const double (&q)[N], const double (&pm)[N], const double (&Vp)[sym_size(N)]) {
    const uint32_t obs = p.observables;
    const int64_t T = p.ntraj;
    double dia[N];
#pragma unroll
    for (int j = 0; j < N; ++j) dia[j] = lane_sum<NB>(0.5 * (q[j] * q[j] + pm[j] * pm[j]) - p.nrpmd_gamma) / NB;
    if (em.isave == 0 && em.active && (obs & (1u << NQCB200_OBS_POPCORR_DIABATIC))) {
#pragma unroll
        for (int i = 0; i < N; ++i) p.pop0[(int64_t)i * T + em.traj] = dia[i];
    }
    if (obs & (1u << NQCB200_OBS_DIABATIC_POP)) {
#pragma unroll
        for (int i = 0; i < N; ++i) em.emit(NQCB200_OBS_DIABATIC_POP, i, dia[i]);
    }
    if (obs & (1u << NQCB200_OBS_POPCORR_DIABATIC)) {
        double p0[N];
#pragma unroll
        for (int i = 0; i < N; ++i) p0[i] = (em.isave == 0) ? dia[i] : p.pop0[(int64_t)i * T + em.traj];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) em.emit(NQCB200_OBS_POPCORR_DIABATIC, i + N * j, p0[i] * dia[j]);
    }
    if (obs & ((1u << NQCB200_OBS_KINETIC) | (1u << NQCB200_OBS_POTENTIAL) | (1u << NQCB200_OBS_TOTAL_ENERGY))) {
        const double kin = 0.5 * lane_sum<NB>(mass * v * v);
        // 0.5 (p'Vt p + q'Vt q) + Vbar per bead, Vt = V - Vbar I  (nrpmd.jl:124-139)
        double vbar = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) vbar += Vp[sidx(N, i, i)];
        vbar /= N;
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double vt = ((i <= j) ? Vp[sidx(N, i, j)] : Vp[sidx(N, j, i)]) - ((i == j) ? vbar : 0.0);
                s += pm[i] * vt * pm[j] + q[i] * vt * q[j];
            }
        const double pot = lane_sum<NB>(0.5 * s + vbar);
        if (obs & (1u << NQCB200_OBS_KINETIC)) em.emit(NQCB200_OBS_KINETIC, 0, kin);
        if (obs & (1u << NQCB200_OBS_POTENTIAL)) em.emit(NQCB200_OBS_POTENTIAL, 0, pot);
        if (obs & (1u << NQCB200_OBS_TOTAL_ENERGY)) {
            const double spr = spring_energy<NB>(r, mass, p.omega_n, lane, group_base);
            em.emit(NQCB200_OBS_TOTAL_ENERGY, 0, kin + pot + spr);
        }
    }
    if (obs & ((1u << NQCB200_OBS_POSITION) | (1u << NQCB200_OBS_VELOCITY))) {
        const double rc = lane_sum<NB>(r) / NB, vc = lane_sum<NB>(v) / NB;
        if (obs & (1u << NQCB200_OBS_POSITION)) em.emit(NQCB200_OBS_POSITION, 0, rc);
        if (obs & (1u << NQCB200_OBS_VELOCITY)) em.emit(NQCB200_OBS_VELOCITY, 0, vc);
    }
    if (obs & ((1u << NQCB200_OBS_MAPPING_Q) | (1u << NQCB200_OBS_MAPPING_P))) {
        // OutputMappingPosition / OutputMappingMomentum (DynamicsOutputs.jl:157,165): bead b's variables travel to the
        // group's lane 0, which carries the trajectory's values into the stream ((nstates, nbeads) column-major)
#pragma unroll 1
        for (int b = 0; b < NB; ++b) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double qb = __shfl_sync(0xffffffffu, q[i], group_base + b);
                const double pb = __shfl_sync(0xffffffffu, pm[i], group_base + b);
                if (obs & (1u << NQCB200_OBS_MAPPING_Q)) em.emit(NQCB200_OBS_MAPPING_Q, i + N * b, qb);
                if (obs & (1u << NQCB200_OBS_MAPPING_P)) em.emit(NQCB200_OBS_MAPPING_P, i + N * b, pb);
            }
        }
    }
}

template <class M>
NQ_D void nrpmd_potential(const KParams& p, double q, double (&Vp)[sym_size(M::NS)]) {
    const double rr[1] = {q}, zz[1] = {0.0};
    M::template potential_partial<1>(p.params, rr, zz, zz, true, Vp);
}

template <class M, int NB>
__global__ void __launch_bounds__(kBlockThreads) nrpmd_step_kernel(const __grid_constant__ KParams p) {
    constexpr int N = M::NS;
    __shared__ double smem[2 * (kBlockThreads / 32)];
    const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t traj = gthread / NB;
    const int lane = (int)(gthread % NB);
    const int group_base = (threadIdx.x & 31) & ~(NB - 1);
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const int64_t T = p.ntraj;
    double r = p.r[(int64_t)lane * T + traj], v = p.v[(int64_t)lane * T + traj];
    const double mass = p.masses[0];
    double q[N], pm[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        q[j] = p.qmap[((int64_t)lane * N + j) * T + traj];
        pm[j] = p.pmap[((int64_t)lane * N + j) * T + traj];
    }
    FreeRingPolymer<NB> frp;   // half-step Cayley (the engine uploads sqrt(M) for NRPMD)
    frp.init(p, lane, group_base);
    const double dt = p.dt;
    double Vp[sym_size(N)];

#pragma unroll 1
    for (int is = 0; is < p.nsteps; ++is) {
        const int64_t step = p.step0 + is;
        __syncthreads();      // keeps the block's warps in the same part of the step's code (instruction fetch: ncu showed 0.64
                              // `no_instruction` stalls per issue; 8.20e8 -> 8.50e8 with the barrier, profiles/r02/SUMMARY.md)
        frp.step(r, v);
        double dVp[sym_size(N)];
        model_value_and_derivative<M>(p.params, r, Vp, dVp);    // V and dV/dr at the same r: shared transcendentals
        Eig<N> e;
        sym_eigh<N>(Vp, e);
        double vbar = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) vbar += Vp[sidx(N, i, i)];
        vbar /= N;
        double lam[N], qa[N], pa[N];
#pragma unroll
        for (int i = 0; i < N; ++i) lam[i] = e.w[i] - vbar;
        // adiabatic mapping variables, rotated by the diagonal propagators C = cos(l dt), D = sin(-l dt)
#pragma unroll
        for (int a = 0; a < N; ++a) {
            double sq = 0.0, sp = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j) { sq += e.Z[j][a] * q[j]; sp += e.Z[j][a] * pm[j]; }
            double sn, cs;
            sincos(lam[a] * dt, &sn, &cs);
            qa[a] = cs * sq + sn * sp;      // C q - D p with D = -sin(l dt)
            pa[a] = cs * sp - sn * sq;      // C p + D q
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double sq = 0.0, sp = 0.0;
#pragma unroll
            for (int a = 0; a < N; ++a) { sq += e.Z[j][a] * qa[a]; sp += e.Z[j][a] * pa[a]; }
            q[j] = sq; pm[j] = sp;
        }
        // nuclear kick: W = Z'(dV - Dbar I)Z ; Gamma, Xi (ringpolymer_mint.jl:107-121)
        double Ap[sym_size(N)];
        double dbar = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) dbar += dVp[sidx(N, i, i)];
        dbar /= N;
        similarity<N>(dVp, e.Z, Ap);
        double force = 0.0;
#pragma unroll
        for (int a = 0; a < N; ++a) {
            force += 0.5 * (Ap[sidx(N, a, a)] - dbar) * dt * (qa[a] * qa[a] + pa[a] * pa[a]);
#pragma unroll
            for (int b = a + 1; b < N; ++b) {
                const double dl = lam[a] - lam[b], W = Ap[sidx(N, a, b)];
                double sn, cs;
                sincos(dl * dt, &sn, &cs);
                const double G = sn * W / dl;                 // symmetric
                const double X = (1.0 - cs) * W / dl;         // Xi[b,a] = X, Xi[a,b] = -X
                force += G * (qa[a] * qa[b] + pa[a] * pa[b]);
                force -= X * (qa[b] * pa[a] - qa[a] * pa[b]);
            }
        }
        v -= force / mass;
        v -= dbar / mass * dt;
        frp.step(r, v);

        if ((step + 1) % p.save_every == 0) {
            const int64_t isave = (step + 1) / p.save_every;
            if (isave < p.nsave) {
                Emitter em{p, traj, valid && lane == 0, (int)isave, smem, 0};
                nrpmd_potential<M>(p, r, Vp);
                nrpmd_record_save<N, NB>(p, em, lane, group_base, r, v, mass, q, pm, Vp);
            }
        }
    }
    if (valid) {
        p.r[(int64_t)lane * T + traj] = r;
        p.v[(int64_t)lane * T + traj] = v;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            p.qmap[((int64_t)lane * N + j) * T + traj] = q[j];
            p.pmap[((int64_t)lane * N + j) * T + traj] = pm[j];
        }
    }
}

// save point 0 (called after nqcb200_set_mapping); initialize! is empty for RingPolymerMInt (:26)
template <class M, int NB>
__global__ void __launch_bounds__(kBlockThreads) nrpmd_init_kernel(const __grid_constant__ KParams p, int, int, const double*) {
    constexpr int N = M::NS;
    __shared__ double smem[2 * (kBlockThreads / 32)];
    const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t traj = gthread / NB;
    const int lane = (int)(gthread % NB);
    const int group_base = (threadIdx.x & 31) & ~(NB - 1);
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const int64_t T = p.ntraj;
    const double r = p.r[(int64_t)lane * T + traj], v = p.v[(int64_t)lane * T + traj];
    double q[N], pm[N], Vp[sym_size(N)];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        q[j] = p.qmap[((int64_t)lane * N + j) * T + traj];
        pm[j] = p.pmap[((int64_t)lane * N + j) * T + traj];
    }
    nrpmd_potential<M>(p, r, Vp);
    Emitter em{p, traj, valid && lane == 0, 0, smem, 0};
    nrpmd_record_save<N, NB>(p, em, lane, group_base, r, v, p.masses[0], q, pm, Vp);
}

#endif

}  // namespace nq

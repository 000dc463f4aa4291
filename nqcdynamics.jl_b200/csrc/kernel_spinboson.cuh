// kernel_spinboson.cuh -- Simulation{FSSH} / Simulation{Ehrenfest} on SpinBoson (linear-coupling harmonic bath):
// LPT (1 or 2) THREADS per trajectory, the D bath coordinates and velocities of a block's 128 trajectories resident in
// shared memory ([mode][trajectory slot], conflict-free), the 2x2 electronic problem in registers (replicated on the
// LPT lanes, which execute identical instructions on identical data), persistent over the steps of a launch.  The
// bath fills the SM's shared memory at 128 trajectories, i.e. one warp per scheduler with one thread per trajectory:
// every fixed-latency dependency was exposed (ncu: 43 % "wait" stalls, FP64 pipe 42 % active).  With LPT = 2 the two
// lanes of a trajectory split the modes (even / odd) and each scheduler owns two warps.  Same reference path as kernel_density.cuh (BABwithTsit5, bab_electronics.jl:61-91; HoppingCallback,
// surface_hopping.jl:2-168; fssh.jl:67-121; ehrenfest.jl:50-68), specialised to the structure of the model
// (NQCModels SpinBoson, docs/src/NQCModels/systembathmodels.md:20-26):
//
//     V = (sum_j w_j^2 r_j^2 / 2) I + (eps + sum_j c_j r_j) sigma_z + Delta sigma_x ,   dV/dr_j = w_j^2 r_j I + c_j sigma_z
//
//  * Z' (dV/dr_j) Z = w_j^2 r_j I + c_j S with ONE 2x2 matrix S = Z' sigma_z Z per step, so the force on mode j is
//    -(w_j^2 r_j A + c_j B) / m_j with two per-trajectory scalars (FSSH: A = 1, B = S_ss; Ehrenfest: A = tr sigma,
//    B = sum_ab Re sigma_ab S_ab), and d_j[0,1] = -c_j S_01 / (w_0 - w_1): the D similarity transforms of the generic
//    kernel collapse to a few flops.
//  * v.d = -S_01/(w_0 - w_1) sum_j c_j v_j,  and the rescaling coefficients a = 1/2 sum d_j^2/m_j, b = sum d_j v_j are
//    closed forms of sum_j c_j^2/m_j and sum_j c_j v_j: a hop costs O(1).
//  * Because the force is linear in two scalars, the second half kick of step k and the first half kick of step k+1 use
//    the same per-mode acceleration; the modes are therefore swept ONCE per step (half kick, half kick, drift) and the
//    velocity kept in shared memory is the half-kicked one.  sum_j c_j v_j after the full kick follows from sums
//    accumulated during the sweep.  The velocity change of an accepted hop (v_j -= gamma d_j / m_j) or of a reflected
//    frustrated hop is applied inside the next sweep through two per-trajectory scalars.
//  * Launch-fused initialisation: when KParams.r_aos / v_aos are set (nqcb200_run_from_host) the block reads its
//    128 x D tile of the caller's trajectory-major arrays directly (pinned host memory is read over PCIe by the kernel
//    itself, so the upload of later blocks overlaps the dynamics of earlier ones), evaluates the t0 eigenproblem,
//    transforms rho, samples the initial state and records save point 0 -- no transposition / init kernels.
//  * The harmonic shift sum_j w_j^2 r_j^2 / 2 is a multiple of the identity: it is accumulated only when an energy
//    output or the diagnostics need the absolute eigenvalues.
//  * Quirks Q1-Q4 of the reference are kept: zeroed electronic buffer on the first step, force not refreshed after a
//    hop (the carried A, B are the pre-hop ones), buffered v.d from the pre-rescale velocity.
#pragma once
#include "kernel_density.cuh"

namespace nq {

#if defined(__CUDACC__)


constexpr int kSbTraj = 128;   // trajectories per block (the bath of 128 trajectories fills shared memory at D = 100)

struct SbSmem {
    double* rs;      // [D][kSbTraj]
    double* vs;      // [D][kSbTraj]  half-kicked velocity (true velocity before the first step of a launch)
    double2 *k1, *k2, *k3;   // [D] {dt w^2/m, c/m}, {c, c w^2/m}, {w^2/2, m}  (k3 only for energies / kinetic sums)
    NQ_D void carve(double* base, int D) {
        rs = base; vs = rs + (size_t)D * kSbTraj;
        k1 = reinterpret_cast<double2*>(vs + (size_t)D * kSbTraj); k2 = k1 + D; k3 = k2 + D;
    }
};
NQ_HD size_t sb_smem_bytes(int D) { return ((size_t)2 * D * kSbTraj + 6 * (size_t)D) * sizeof(double); }

// Thread -> (trajectory slot, part).  A warp holds 32/LPT slots; lanes [0, 32/LPT) are part 0, the next 32/LPT lanes
// part 1 of the SAME slots, so the 16 lanes of one shared-memory phase touch 16 consecutive slots (conflict-free) and
// the partner of a lane is lane ^ (32/LPT).
template <int LPT>
struct SbMap {
    static constexpr int kSlotsPerWarp = 32 / LPT;
    NQ_D static int slot(int tid) { return (tid >> 5) * kSlotsPerWarp + (tid & (kSlotsPerWarp - 1)); }
    NQ_D static int part(int tid) { return (tid & 31) / kSlotsPerWarp; }
    NQ_D static double sum(double x) {      // sum over the LPT lanes of a trajectory, bit-identical on all of them
#pragma unroll
        for (int o = kSlotsPerWarp; o < 32; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        return x;
    }
};

// Observable accumulation without a block barrier: warp sum, one atomicAdd per warp into one of the replicas.
struct SbEmitter {
    const KParams& p;
    int64_t traj;
    bool active;     // valid trajectory, part 0
    int isave;
    NQ_D void emit(int obs_id, int k, double val) {
        const int64_t off = p.layout.offset[obs_id] + (int64_t)isave * p.layout.width[obs_id] + k;
        if (p.obs_traj != nullptr && active) p.obs_traj[off * p.ntraj + traj] = val;
        const double ws = warp_sum(active ? val : 0.0);
        if ((threadIdx.x & 31) == 0)
            atomicAdd(&p.obs_sum[(int64_t)((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % kObsReplicas) * p.layout.total + off], ws);
    }
    // Four consecutive values (k0 .. k0+3) in one transposing reduction: after two exchange rounds each quarter of
    // the warp owns one value, three more rounds finish it -- 6 double shuffles instead of 20, one atomic per value.
    NQ_D void emit4(int obs_id, int k0, const double (&val)[4]) {
        const int64_t off = p.layout.offset[obs_id] + (int64_t)isave * p.layout.width[obs_id] + k0;
        if (p.obs_traj != nullptr && active) {
#pragma unroll
            for (int q = 0; q < 4; ++q) p.obs_traj[(off + q) * p.ntraj + traj] = val[q];
        }
        const int lane = threadIdx.x & 31;
        const double v0 = active ? val[0] : 0.0, v1 = active ? val[1] : 0.0, v2 = active ? val[2] : 0.0, v3 = active ? val[3] : 0.0;
        const bool hi = (lane & 16) != 0;
        double a = hi ? v2 : v0, b = hi ? v3 : v1;
        a += __shfl_xor_sync(0xffffffffu, hi ? v0 : v2, 16);
        b += __shfl_xor_sync(0xffffffffu, hi ? v1 : v3, 16);
        const bool mid = (lane & 8) != 0;
        double c = mid ? b : a;
        c += __shfl_xor_sync(0xffffffffu, mid ? a : b, 8);
        c += __shfl_xor_sync(0xffffffffu, c, 4);
        c += __shfl_xor_sync(0xffffffffu, c, 2);
        c += __shfl_xor_sync(0xffffffffu, c, 1);
        if ((lane & 7) == 0) {      // lanes 0, 8, 16, 24 hold the sums of values 0, 1, 2, 3
            const int q = lane >> 3;
            atomicAdd(&p.obs_sum[(int64_t)((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % kObsReplicas) * p.layout.total + off + q], c);
        }
    }
};

// S = Z' sigma_z Z (symmetric 2x2): s00, s01, s11
NQ_D void sb_sz(const Eig<2>& e, double& s00, double& s01, double& s11) {
    s00 = e.Z[0][0] * e.Z[0][0] - e.Z[1][0] * e.Z[1][0];
    s01 = e.Z[0][0] * e.Z[0][1] - e.Z[1][0] * e.Z[1][1];
    s11 = e.Z[0][1] * e.Z[0][1] - e.Z[1][1] * e.Z[1][1];
}
// force scalars (A, B): acceleration of mode j = -(w_j^2 r_j A + c_j B) / m_j
template <int METHOD>
NQ_D void sb_force_scalars(const Herm<2>& s, int st, double tr0, double s00, double s01, double s11, double& A, double& B) {
    if (METHOD == NQCB200_METHOD_FSSH) { A = 1.0; B = st ? s11 : s00; }                       // fssh.jl:67-74
    else { A = tr0; B = s.x[0] * s00 + 2.0 * s.x[1] * s01 + s.x[2] * s11; }                    // ehrenfest.jl:57-65 (tr sigma is conserved)
}

struct SbTraj {
    Herm<2> s;
    int st;
    double Zref[2][2];
    ElecParams<2> cur;
    double A, B;          // force scalars carried between steps (quirk Q2)
    double gm, gd;        // pending velocity change: v_j -= c_j (gm / m_j + gd)
    double p0d[2], p0a[2];   // initial diabatic / adiabatic populations (correlation functions), cached from pop0
};

struct SbSums { double harm, lin, cvt, cwr, msv2; };

// One sweep over this lane's modes (j = part, part + LPT, ...).  first: vs holds the true velocity (launch entry)
// instead of the half-kicked one.  DRIFT: perform the drift of a new step (false = only finish the pending kicks,
// used at save points that need velocities and at launch exit).  HARM: also accumulate sum w^2 r^2 / 2.
// PEND (warp-uniform): some trajectory of the warp has a pending hop rescaling.  The sums are over ALL modes of the
// trajectory (pair-summed), at the new positions.  Four modes are advanced in lock step so that their dependent FMA
// chains interleave.
// LIN: accumulate sum c r at the new positions (otherwise the caller advances it: sum c r += dt sum c vt, exact in real
// arithmetic because every mode drifts by dt vt).  k1.x carries dt w^2/m, so "dacc" below is dt times the acceleration
// and the ordinary drift sweep (both half kicks, FSSH) turns v into vt with two FMAs.
template <int LPT, int METHOD, bool DRIFT, bool HARM, bool LIN = true>
NQ_D void sb_sweep(const SbSmem& M, int D, int slot, int part, bool first, bool pend, double A, double B, double gm,
                   double gd, double dt, double hdt, SbSums& S) {
    constexpr int W = 4;
    double h4[W], l4[W], c4[W], w4[W], m4[W];
#pragma unroll
    for (int q = 0; q < W; ++q) { h4[q] = 0.0; l4[q] = 0.0; c4[q] = 0.0; w4[q] = 0.0; m4[q] = 0.0; }
    // second half kick of the previous step (not on launch entry) + first half kick of this one, in units of dt
    const double nA = -A, nBdt = -B * dt, kfac_2 = first ? 0.0 : 0.5, kfac_both = kfac_2 + 0.5;
    // Ehrenfest: A = tr sigma is a constant of the motion, 1 up to the rounding of Z' rho Z for a normalised initial
    // state: within 4 ulp it is taken as exactly 1 (a relative change of the force below 1e-15) -> same fast path
    const bool unitA = (METHOD == NQCB200_METHOD_FSSH) || __all_sync(0xffffffffu, fabs(A - 1.0) <= 1e-15);
    auto body = [&](const int (&jj)[W], const bool (&ok)[W]) {
        double2 k1[W], k2[W], k3[W];
        double r[W], v[W];
#pragma unroll
        for (int q = 0; q < W; ++q) {
            k1[q] = M.k1[jj[q]]; k2[q] = M.k2[jj[q]];
            if (HARM || !DRIFT) k3[q] = M.k3[jj[q]];
            r[q] = M.rs[(size_t)jj[q] * kSbTraj + slot];
            v[q] = M.vs[(size_t)jj[q] * kSbTraj + slot];
        }
        if (pend) {
#pragma unroll
            for (int q = 0; q < W; ++q) v[q] = fma(-gm, k1[q].y, fma(-gd, k2[q].x, v[q]));   // hop rescaling / reflection
        }
        double vt[W];
        if (DRIFT && !first && unitA) {
            // vt = v + dt a,  dt a = -(dt w^2/m) r - dt B (c/m)   (A = 1, fssh.jl:67-74; step_B! twice, steps.jl:3-5)
#pragma unroll
            for (int q = 0; q < W; ++q) vt[q] = fma(nBdt, k1[q].y, fma(-k1[q].x, r[q], v[q]));
        } else {
#pragma unroll
            for (int q = 0; q < W; ++q) {
                const double dacc = (METHOD == NQCB200_METHOD_FSSH) ? fma(-k1[q].x, r[q], nBdt * k1[q].y)
                                                                     : fma(nA * k1[q].x, r[q], nBdt * k1[q].y);
                vt[q] = fma(DRIFT ? kfac_both : kfac_2, dacc, v[q]);
            }
        }
        if (DRIFT) {
            double rn[W];
#pragma unroll
            for (int q = 0; q < W; ++q) rn[q] = fma(dt, vt[q], r[q]);                   // step_A!  steps.jl:6-8
#pragma unroll
            for (int q = 0; q < W; ++q) {
                if (ok[q]) {
                    M.rs[(size_t)jj[q] * kSbTraj + slot] = rn[q];
                    M.vs[(size_t)jj[q] * kSbTraj + slot] = vt[q];
                    if (HARM) h4[q] = fma(k3[q].x * rn[q], rn[q], h4[q]);
                    if (LIN) l4[q] = fma(k2[q].x, rn[q], l4[q]);
                    c4[q] = fma(k2[q].x, vt[q], c4[q]);
                    w4[q] = fma(k2[q].y, rn[q], w4[q]);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < W; ++q) {
                if (ok[q]) {
                    M.vs[(size_t)jj[q] * kSbTraj + slot] = vt[q];                       // true velocity (previous step's kick finished)
                    m4[q] = fma(k3[q].y * vt[q], vt[q], m4[q]);                         // sum m v^2
                }
            }
        }
    };
    int j = part;
    for (; j + (W - 1) * LPT < D; j += W * LPT) {
        const int jj[W] = {j, j + LPT, j + 2 * LPT, j + 3 * LPT};
        const bool ok[W] = {true, true, true, true};
        body(jj, ok);
    }
    if (j < D) {
        const int jl = j;   // clamp the tail to a mode this lane owns
        const int jj[W] = {j, (j + LPT < D) ? j + LPT : jl, (j + 2 * LPT < D) ? j + 2 * LPT : jl, (j + 3 * LPT < D) ? j + 3 * LPT : jl};
        const bool ok[W] = {true, j + LPT < D, j + 2 * LPT < D, j + 3 * LPT < D};
        body(jj, ok);
    }
    S.harm = SbMap<LPT>::sum((h4[0] + h4[1]) + (h4[2] + h4[3]));
    S.lin = SbMap<LPT>::sum((l4[0] + l4[1]) + (l4[2] + l4[3]));
    S.cvt = SbMap<LPT>::sum((c4[0] + c4[1]) + (c4[2] + c4[3]));
    S.cwr = SbMap<LPT>::sum((w4[0] + w4[1]) + (w4[2] + w4[3]));
    S.msv2 = DRIFT ? 0.0 : SbMap<LPT>::sum((m4[0] + m4[1]) + (m4[2] + m4[3]));
    __syncwarp();   // the partner lane's modes are read at save points / launch exit
}

template <int METHOD>
NQ_D void sb_record_save(const KParams& p, SbEmitter& em, const SbSmem& M, int slot, SbTraj& R, const Eig<2>& e,
                         double msv2) {
    constexpr int N = 2;
    const uint32_t obs = p.observables;
    const int64_t T = p.ntraj;
    double adi[N], dia[N];
    adiabatic_population<N, METHOD>(R.s, R.st, adi);
    const bool need_dia = obs & ((1u << NQCB200_OBS_DIABATIC_POP) | (1u << NQCB200_OBS_POPCORR_DIABATIC) |
                                 (1u << NQCB200_OBS_SCATTERING_DIABATIC));
    if (need_dia) diabatic_population<N, METHOD>(R.s, R.st, e, dia);
    else { dia[0] = 0.0; dia[1] = 0.0; }
    if (em.isave == 0 && em.active && (obs & ((1u << NQCB200_OBS_POPCORR_DIABATIC) | (1u << NQCB200_OBS_POPCORR_ADIABATIC)))) {
#pragma unroll
        for (int i = 0; i < N; ++i) { p.pop0[(int64_t)i * T + em.traj] = dia[i]; p.pop0[(int64_t)(N + i) * T + em.traj] = adi[i]; }
    }
    if (em.isave == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) { R.p0d[i] = dia[i]; R.p0a[i] = adi[i]; }
    }
    if (obs & (1u << NQCB200_OBS_ADIABATIC_POP)) {
#pragma unroll
        for (int i = 0; i < N; ++i) em.emit(NQCB200_OBS_ADIABATIC_POP, i, adi[i]);
    }
    if (obs & (1u << NQCB200_OBS_DIABATIC_POP)) {
#pragma unroll
        for (int i = 0; i < N; ++i) em.emit(NQCB200_OBS_DIABATIC_POP, i, dia[i]);
    }
    if (obs & (1u << NQCB200_OBS_POPCORR_DIABATIC)) {
        double p0[N];
#pragma unroll
        for (int i = 0; i < N; ++i) p0[i] = R.p0d[i];
        const double c4[4] = {p0[0] * dia[0], p0[1] * dia[0], p0[0] * dia[1], p0[1] * dia[1]};   // i + N * j
        em.emit4(NQCB200_OBS_POPCORR_DIABATIC, 0, c4);
    }
    if (obs & (1u << NQCB200_OBS_POPCORR_ADIABATIC)) {
        double p0[N];
#pragma unroll
        for (int i = 0; i < N; ++i) p0[i] = R.p0a[i];
        const double c4[4] = {p0[0] * adi[0], p0[1] * adi[0], p0[0] * adi[1], p0[1] * adi[1]};
        em.emit4(NQCB200_OBS_POPCORR_ADIABATIC, 0, c4);
    }
    if (obs & ((1u << NQCB200_OBS_KINETIC) | (1u << NQCB200_OBS_POTENTIAL) | (1u << NQCB200_OBS_TOTAL_ENERGY))) {
        const double kin = 0.5 * msv2;
        double pot;
        if (METHOD == NQCB200_METHOD_FSSH) pot = R.st ? e.w[1] : e.w[0];                       // fssh.jl:150-154
        else pot = R.s.x[0] * e.w[0] + R.s.x[2] * e.w[1];                                     // ehrenfest.jl:85-95
        if (obs & (1u << NQCB200_OBS_KINETIC)) em.emit(NQCB200_OBS_KINETIC, 0, kin);
        if (obs & (1u << NQCB200_OBS_POTENTIAL)) em.emit(NQCB200_OBS_POTENTIAL, 0, pot);
        if (obs & (1u << NQCB200_OBS_TOTAL_ENERGY)) em.emit(NQCB200_OBS_TOTAL_ENERGY, 0, kin + pot);
    }
    if (obs & ((1u << NQCB200_OBS_POSITION) | (1u << NQCB200_OBS_VELOCITY))) {
        for (int dof = 0; dof < p.D; ++dof) {
            if (obs & (1u << NQCB200_OBS_POSITION)) em.emit(NQCB200_OBS_POSITION, dof, M.rs[(size_t)dof * kSbTraj + slot]);
            if (obs & (1u << NQCB200_OBS_VELOCITY)) em.emit(NQCB200_OBS_VELOCITY, dof, M.vs[(size_t)dof * kSbTraj + slot]);
        }
    }
    if (obs & (1u << NQCB200_OBS_DISCRETE_STATE)) em.emit(NQCB200_OBS_DISCRETE_STATE, 0, (double)(R.st + 1));
    const bool last = (em.isave == p.nsave - 1);
    if (obs & ((1u << NQCB200_OBS_SCATTERING) | (1u << NQCB200_OBS_SCATTERING_DIABATIC))) {
        const bool trans = M.rs[slot] > 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (obs & (1u << NQCB200_OBS_SCATTERING)) {
                em.emit(NQCB200_OBS_SCATTERING, i, (last && !trans) ? adi[i] : 0.0);
                em.emit(NQCB200_OBS_SCATTERING, N + i, (last && trans) ? adi[i] : 0.0);
            }
            if (obs & (1u << NQCB200_OBS_SCATTERING_DIABATIC)) {
                em.emit(NQCB200_OBS_SCATTERING_DIABATIC, i, (last && !trans) ? dia[i] : 0.0);
                em.emit(NQCB200_OBS_SCATTERING_DIABATIC, N + i, (last && trans) ? dia[i] : 0.0);
            }
        }
    }
    if (obs & (1u << NQCB200_OBS_SIGMA)) {
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                em.emit(NQCB200_OBS_SIGMA, j + N * k, R.s.X(j, k));
                em.emit(NQCB200_OBS_SIGMA, N * N + j + N * k, R.s.Y(j, k));
            }
    }
}


template <int METHOD, int LPT>
__global__ void __launch_bounds__(kSbTraj * LPT, 1) spinboson_step_kernel(const __grid_constant__ KParams p) {
    constexpr int N = 2;
    extern __shared__ __align__(16) double sb_sm[];
    SbSmem M;
    M.carve(sb_sm, p.D);
    const int tid = threadIdx.x, D = p.D;
    const int slot = SbMap<LPT>::slot(tid), part = SbMap<LPT>::part(tid);
    const int64_t block_base = (int64_t)blockIdx.x * kSbTraj;
    int64_t traj = block_base + slot;
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const bool lead = valid && part == 0;
    const int64_t T = p.ntraj;
    const double dt = p.dt, hdt = 0.5 * p.dt;
    const bool need_harm = p.diagnostics || (p.observables & ((1u << NQCB200_OBS_POTENTIAL) | (1u << NQCB200_OBS_TOTAL_ENERGY)));
    const bool fused = (p.r_aos != nullptr) && p.step0 == 0;

    // per-mode constants
    for (int j = tid; j < D; j += blockDim.x) {
        const double w = p.bath_a[j], c = p.bath_b[j], m = p.masses[j];
        M.k1[j] = make_double2(dt * w * w / m, c / m); M.k2[j] = make_double2(c, c * w * w / m); M.k3[j] = make_double2(0.5 * w * w, m);
    }
    if (fused) {
        // this block's [128][D] tile of the caller's trajectory-major r, v (device staging or pinned host memory)
        const int64_t nt = min((int64_t)kSbTraj, T - block_base);
        const int n = (int)(nt * D), ntot = kSbTraj * D;
        const double* __restrict__ sr = p.r_aos + block_base * D;
        const double* __restrict__ sv = p.v_aos + block_base * D;
        constexpr int U = 8;
        for (int base = 0; base < ntot; base += U * (int)blockDim.x) {
            double tr[U], tv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = base + u * (int)blockDim.x + tid;
                tr[u] = (idx < n) ? sr[idx] : 0.0;
                tv[u] = (idx < n) ? sv[idx] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = base + u * (int)blockDim.x + tid;
                if (idx < ntot) {
                    const int t = idx / D, j = idx - t * D;
                    M.rs[(size_t)j * kSbTraj + t] = tr[u];
                    M.vs[(size_t)j * kSbTraj + t] = tv[u];
                }
            }
        }
    } else {
        for (int j = part; j < D; j += LPT) {
            M.rs[(size_t)j * kSbTraj + slot] = p.r[(int64_t)j * T + traj];
            M.vs[(size_t)j * kSbTraj + slot] = p.v[(int64_t)j * T + traj];
        }
    }
    __syncthreads();
    double C2 = 0.0, Cc = 0.0;      // sum c^2/m, sum c^2
    for (int j = 0; j < D; ++j) { const double c = M.k2[j].x; C2 = fma(c, M.k1[j].y, C2); Cc = fma(c, c, Cc); }

    SbTraj R;
    R.s.x[0] = p.sig_re[(int64_t)0 * T + traj]; R.s.x[1] = p.sig_re[(int64_t)2 * T + traj]; R.s.x[2] = p.sig_re[(int64_t)3 * T + traj];
    R.s.y[0] = p.sig_im[(int64_t)2 * T + traj];
    R.st = p.state ? p.state[traj] : 0;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = 0; k < N; ++k) R.Zref[j][k] = p.Zprev[(int64_t)(j + N * k) * T + traj];
    R.gm = 0.0; R.gd = 0.0;
    {
        const bool corr = p.observables & ((1u << NQCB200_OBS_POPCORR_DIABATIC) | (1u << NQCB200_OBS_POPCORR_ADIABATIC));
#pragma unroll
        for (int i = 0; i < N; ++i) {      // written by save point 0 (init kernel / fused initialisation)
            R.p0d[i] = (corr && !fused) ? p.pop0[(int64_t)i * T + traj] : 0.0;
            R.p0a[i] = (corr && !fused) ? p.pop0[(int64_t)(N + i) * T + traj] : 0.0;
        }
    }
    Eig<N> e;
    if (fused) {
        // DynamicsVariables at t0 (fssh.jl:47-65, ehrenfest.jl:43-48): eigenproblem at r0, sigma = Z' rho Z, initial
        // state ~ diag(sigma), zeroed electronic buffer (Q1), save point 0
        double lin = 0.0, harm = 0.0, msv2 = 0.0;
        for (int j = part; j < D; j += LPT) {
            const double r = M.rs[(size_t)j * kSbTraj + slot], v = M.vs[(size_t)j * kSbTraj + slot];
            lin = fma(M.k2[j].x, r, lin);
            harm = fma(M.k3[j].x * r, r, harm);
            msv2 = fma(M.k3[j].y * v, v, msv2);
        }
        lin = SbMap<LPT>::sum(lin); harm = need_harm ? SbMap<LPT>::sum(harm) : 0.0; msv2 = SbMap<LPT>::sum(msv2);
        const double l = p.params[0] + lin;
        double Vp[3] = {harm + l, p.params[1], harm - l};
        sym_eigh<N>(Vp, e);
        fix_gauge<N>(e, R.Zref);
        if (p.init_basis == 1) {
            Herm<N> o;
#pragma unroll
            for (int i = 0; i < N; ++i)
#pragma unroll
                for (int j = i; j < N; ++j) {
                    double sx = 0.0, sy = 0.0;
#pragma unroll
                    for (int a = 0; a < N; ++a)
#pragma unroll
                        for (int b = 0; b < N; ++b) {
                            sx += e.Z[a][i] * R.s.X(a, b) * e.Z[b][j];
                            sy += e.Z[a][i] * R.s.Y(a, b) * e.Z[b][j];
                        }
                    o.x[sidx(N, i, j)] = sx;
                    if (j > i) o.y[aidx(N, i, j)] = sy;
                }
            R.s = o;
        }
        if (METHOD == NQCB200_METHOD_FSSH && p.init_sample_state) {
            const double xi = p.init_state_draw ? p.init_state_draw[traj]
                                                : philox_uniform(p.seed, (uint64_t)(p.traj_offset + traj), 0ull, 1u);
            const double target = xi * (R.s.x[0] + R.s.x[2]);
            R.st = (R.s.x[0] < target) ? 1 : 0;
        }
        R.cur.E[0] = 0.0; R.cur.E[1] = 0.0; R.cur.g[0] = 0.0;
        SbEmitter em{p, traj, lead, 0};
        sb_record_save<METHOD>(p, em, M, slot, R, e, msv2);
    } else {
        R.cur.E[0] = p.ecur[(int64_t)0 * T + traj]; R.cur.E[1] = p.ecur[(int64_t)1 * T + traj];
        R.cur.g[0] = p.ecur[(int64_t)(N + 0 + N * 1) * T + traj];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            e.w[i] = R.cur.E[i];
#pragma unroll
            for (int k = 0; k < N; ++k) e.Z[i][k] = R.Zref[i][k];
        }
    }
    const double tr0 = R.s.x[0] + R.s.x[2];     // tr sigma: constant of the motion (the similarity transform keeps it too)
    if (p.step0 == 0) {
        // first step after set_state: no hop has happened, sigma is sigma(t0), Zref the eigenvectors at r0
        double s00, s01, s11;
        sb_sz(e, s00, s01, s11);
        sb_force_scalars<METHOD>(R.s, R.st, tr0, s00, s01, s11, R.A, R.B);
    } else {
        R.A = p.sb_carry[traj]; R.B = p.sb_carry[T + traj];
    }
    unsigned long long nhops = 0, nfrus = 0;
    bool first = true;
    double lin_run = 0.0;     // sum_j c_j r_j at the positions in shared memory (re-summed on every launch entry)

#pragma unroll 1
    for (int is = 0; is < p.nsteps; ++is) {
        const int64_t step = p.step0 + is;
        const double t = p.t0 + dt * (double)step;
        const double tcur = (step == 0) ? 0.0 : t;   // Q1
        SbSums S;
        const bool pend = __any_sync(0xffffffffu, R.gm != 0.0 || R.gd != 0.0);
        if (need_harm) sb_sweep<LPT, METHOD, true, true>(M, D, slot, part, first, pend, R.A, R.B, R.gm, R.gd, dt, hdt, S);
        else if (first) sb_sweep<LPT, METHOD, true, false>(M, D, slot, part, first, pend, R.A, R.B, R.gm, R.gd, dt, hdt, S);
        else {
            sb_sweep<LPT, METHOD, true, false, false>(M, D, slot, part, first, pend, R.A, R.B, R.gm, R.gd, dt, hdt, S);
            S.lin = fma(dt, S.cvt, lin_run);      // sum c r advances by dt sum c vt
        }
        lin_run = S.lin;
        first = false; R.gm = 0.0; R.gd = 0.0;
        // update_cache!: V -> eigen (gauge-fixed)
        {
            const double l = p.params[0] + S.lin;
            double Vp[3] = {S.harm + l, p.params[1], S.harm - l};
            sym_eigh<N>(Vp, e);
            fix_gauge<N>(e, R.Zref);
        }
        double s00, s01, s11;
        sb_sz(e, s00, s01, s11);
        sb_force_scalars<METHOD>(R.s, R.st, tr0, s00, s01, s11, R.A, R.B);      // pre-hop state, sigma_prev
        // sum_j c_j v_j after the second half kick: v_j = vt_j - hdt (w_j^2 r_j A + c_j B) / m_j
        const double cv = S.cvt - hdt * (R.A * S.cwr + R.B * C2);
        const double dfac = -s01 / (e.w[0] - e.w[1]);                       // d_j[0,1] = c_j dfac
        ElecParams<N> nxt;
        nxt.E[0] = e.w[0]; nxt.E[1] = e.w[1];
        nxt.g[0] = dfac * cv;
        propagate_density<N>(R.cur, tcur, nxt, t + dt, t, dt, R.s, p.tsit5_ha);

        if (METHOD == NQCB200_METHOD_FSSH) {
            const double xi = (p.rng == NQCB200_RNG_INJECTED)
                                  ? p.draws[(step - p.draws_step0) * T + traj]
                                  : philox_uniform(p.seed, (uint64_t)(p.traj_offset + traj), (uint64_t)step, 0u);
            const int s0 = R.st, m = 1 - s0;
            // fewest_switches_probability! fssh.jl:96-108 (Q4) + select_new_state :110-121 for two states
            double g = 2.0 * (R.s.x[1] / (s0 ? R.s.x[2] : R.s.x[0])) * (s0 ? -nxt.g[0] : nxt.g[0]) * dt;   // G(s0, m)
            g = fmin(1.0, fmax(0.0, g));
            if (g > xi) {
                bool accept = true;
                if (p.rescaling != NQCB200_RESCALE_OFF) {                   // surface_hopping.jl:64-99
                    const double wn = m ? e.w[1] : e.w[0], wo = s0 ? e.w[1] : e.w[0];
                    const double df = -s01 / (wn - wo);                     // d_j[new, old] = c_j df
                    const double a = 0.5 * df * df * C2, b = df * cv, c = wn - wo;
                    const double disc = b * b - 4.0 * a * c;
                    if (disc < 0.0) {
                        accept = false;
                        nfrus += lead;
                        if (p.rescaling == NQCB200_RESCALE_VINVERSION) {    // v -= 2 (v.dhat) dhat
                            const double nrm = sqrt(df * df * Cc);
                            const double gam = b / nrm;
                            R.gd = 2.0 * gam * df / nrm;
                        }
                    } else {
                        const double root = sqrt(disc);
                        const double gam = (b < 0.0) ? (b + root) / (2.0 * a) : (b - root) / (2.0 * a);
                        R.gm = gam * df;
                    }
                }
                if (accept) { R.st = m; nhops += lead; }
            }
        }
        R.cur = nxt;

        if ((step + 1) % p.save_every == 0) {
            const int64_t isave = (step + 1) / p.save_every;
            if (isave < p.nsave) {
                double msv2 = 0.0;
                if (p.observables & ((1u << NQCB200_OBS_KINETIC) | (1u << NQCB200_OBS_TOTAL_ENERGY) | (1u << NQCB200_OBS_VELOCITY))) {
                    // finish the pending kicks so that shared memory holds the true velocities
                    SbSums S2;
                    const bool pend2 = __any_sync(0xffffffffu, R.gm != 0.0 || R.gd != 0.0);
                    sb_sweep<LPT, METHOD, false, false>(M, D, slot, part, false, pend2, R.A, R.B, R.gm, R.gd, dt, hdt, S2);
                    msv2 = S2.msv2;
                    first = true; R.gm = 0.0; R.gd = 0.0;
                }
                SbEmitter em{p, traj, lead, (int)isave};
                sb_record_save<METHOD>(p, em, M, slot, R, e, msv2);
            }
        }
    }

    // launch exit: true velocities, state back to global memory
    if (!first) {
        SbSums S2;
        const bool pend2 = __any_sync(0xffffffffu, R.gm != 0.0 || R.gd != 0.0);
        sb_sweep<LPT, METHOD, false, false>(M, D, slot, part, false, pend2, R.A, R.B, R.gm, R.gd, dt, hdt, S2);
    }
    if (valid) {
        for (int j = part; j < D; j += LPT) {
            p.r[(int64_t)j * T + traj] = M.rs[(size_t)j * kSbTraj + slot];
            p.v[(int64_t)j * T + traj] = M.vs[(size_t)j * kSbTraj + slot];
        }
    }
    if (lead) {
        p.sb_carry[traj] = R.A; p.sb_carry[T + traj] = R.B;
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = 0; k < N; ++k) {
                p.sig_re[(int64_t)(j + N * k) * T + traj] = R.s.X(j, k);
                p.sig_im[(int64_t)(j + N * k) * T + traj] = R.s.Y(j, k);
                p.Zprev[(int64_t)(j + N * k) * T + traj] = R.Zref[j][k];
            }
        if (p.state) p.state[traj] = R.st;
        p.ecur[(int64_t)0 * T + traj] = R.cur.E[0]; p.ecur[(int64_t)1 * T + traj] = R.cur.E[1];
        p.ecur[(int64_t)(N + 0 + N * 1) * T + traj] = R.cur.g[0];
        if (p.diagnostics) {
            double s00, s01, s11;
            sb_sz(e, s00, s01, s11);
            const double dfac = -s01 / (e.w[0] - e.w[1]);
#pragma unroll
            for (int i = 0; i < N; ++i) p.diag_eig[(int64_t)i * T + traj] = e.w[i];
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int k = 0; k < N; ++k) p.diag_Z[(int64_t)(j + N * k) * T + traj] = e.Z[j][k];
            for (int dof = 0; dof < D; ++dof) {
                const double2 k1 = M.k1[dof];
                const double c = M.k2[dof].x;
                const double r = M.rs[(size_t)dof * kSbTraj + slot];
                p.acc[(int64_t)dof * T + traj] = fma(-R.A * (k1.x / dt), r, -R.B * k1.y);
                p.diag_nac[((int64_t)dof * N * N + 0) * T + traj] = 0.0;
                p.diag_nac[((int64_t)dof * N * N + 1) * T + traj] = -c * dfac;       // d[1,0]
                p.diag_nac[((int64_t)dof * N * N + 2) * T + traj] = c * dfac;        // d[0,1]
                p.diag_nac[((int64_t)dof * N * N + 3) * T + traj] = 0.0;
            }
        }
    }
    const unsigned long long wh = __reduce_add_sync(0xffffffffu, (unsigned)nhops);
    const unsigned long long wf = __reduce_add_sync(0xffffffffu, (unsigned)nfrus);
    if ((threadIdx.x & 31) == 0) {
        if (wh) atomicAdd(&p.counters[0], wh);
        if (wf) atomicAdd(&p.counters[1], wf);
    }
}

#endif  // __CUDACC__

}  // namespace nq

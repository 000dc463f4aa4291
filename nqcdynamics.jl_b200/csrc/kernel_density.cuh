// kernel_density.cuh -- persistent multi-step kernel for Simulation{FSSH} / Simulation{Ehrenfest}.
//
// One launch advances every trajectory by `nsteps` nuclear steps with the whole trajectory state
// (r, v, acceleration, sigma, active state, eigenvector gauge, electronic double buffer) held in
// registers; global memory is touched only at launch entry/exit and at save points.
//
// Reference restated (B = 1):
//   BABwithTsit5.perform_step!   src/DynamicsMethods/IntegrationAlgorithms/bab_electronics.jl:61-91
//   update_cache! (V, eigen+gauge, Z'dVZ, NAC)   NQCCalculators (external), see linalg.cuh
//   acceleration!   fssh.jl:67-74 ; ehrenfest.jl:50-68
//   update_parameters! / DensityMatrixODEProblem   electronic_dynamics.jl:38-130 (density.cuh)
//   HoppingCallback: check_hop!/execute_hop!   surface_hopping.jl:2-16, fssh.jl:86-121,
//                    rescale_velocity! surface_hopping.jl:64-99,115-168
//   save after the callback, estimators fssh.jl:132-154, ehrenfest.jl:70-95
//
// Work decomposition: L lanes (L | 32) cooperate on one trajectory; nuclear dofs are strided over
// the lanes (dof = lane + L*jj, DPL per lane) and the few cross-dof sums (V, v.d, a, b, kinetic
// energy) are xor-butterfly reductions, which leave bit-identical values on all L lanes, so the
// replicated electronic state never diverges.  L = 1 for the one-dimensional scattering models.
#pragma once
#include "common.cuh"
#include "density.cuh"
#include "linalg.cuh"
#include "models.cuh"
#include "philox.cuh"
#include "kernels.h"

namespace nq {

#if defined(__CUDACC__)

// Sum over the L lanes that share one trajectory.  The shuffle mask names only that lane group, so
// the reduction is legal inside branches taken by whole groups (the hop / rescale path) while other
// trajectories of the same warp are elsewhere.
template <int L>
NQ_D unsigned group_mask() {
    if (L >= 32) return 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    return ((1u << L) - 1u) << (lane & ~(unsigned)(L - 1));
}
template <int L>
NQ_D double lane_sum(double x) {
    const unsigned mask = group_mask<L>();
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o);
    return x;
}
NQ_D double warp_sum(double x) { return lane_sum<32>(x); }

// kObsReplicas: see kernels.h

// Block-wide sum of one observable value into the shard accumulator.  Warp-collective and
// block-collective: every thread of the block must call it the same number of times.
struct Emitter {
    const KParams& p;
    int64_t traj;
    bool active;     // this thread carries the trajectory's value (valid trajectory, lane 0)
    int isave;
    double* smem;    // [2][kBlockThreads/32]
    int slot;
    bool fresh = true;   // first emit of this save point
    int nw = kBlockThreads / 32;   // warps per block
    NQ_D void emit(int obs_id, int k, double val) {
        // The two scratch rows alternate, so an emit never overwrites the row thread 0 is still summing from the
        // previous emit -- except across save points (every Emitter starts at row 0, and the previous save may have
        // ended on row 0 after an odd number of emits): one barrier per save point closes that window (racecheck).
        if (fresh) { __syncthreads(); fresh = false; }
        const int64_t off = p.layout.offset[obs_id] + (int64_t)isave * p.layout.width[obs_id] + k;
        if (p.obs_traj != nullptr && active) p.obs_traj[off * p.ntraj + traj] = val;
        const double ws = warp_sum(active ? val : 0.0);
        const int warp = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) smem[slot * nw + warp] = ws;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < nw; ++w) tot += smem[slot * nw + w];
            atomicAdd(&p.obs_sum[(int64_t)(blockIdx.x % kObsReplicas) * p.layout.total + off], tot);
        }
        slot ^= 1;
    }
};

template <int N>
NQ_D double select(const double (&a)[N], int i) {
    double out = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) out = (k == i) ? a[k] : out;
    return out;
}

// Per-trajectory view used by the estimators at a save point.
template <int N, int DPL, int L, int METHOD>
struct TrajView {
    const Herm<N>& s;
    int st;
    const Eig<N>& e;
    const double (&r)[DPL];
    const double (&v)[DPL];
    const double (&mass)[DPL];
};

template <int N, int METHOD>
NQ_D void adiabatic_population(const Herm<N>& s, int st, double (&pop)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) pop[i] = (METHOD == NQCB200_METHOD_FSSH) ? ((i == st) ? 1.0 : 0.0) : s.x[sidx(N, i, i)];
}
// FSSH: diag(U (Re sigma with diag -> delta_{i,st}) U') fssh.jl:132-142 ; Ehrenfest: Re diag(U sigma U')
template <int N, int METHOD>
NQ_D void diabatic_population(const Herm<N>& s, int st, const Eig<N>& e, double (&pop)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < N; ++a)
#pragma unroll
            for (int b = 0; b < N; ++b) {
                double rho = s.X(a, b);
                if (METHOD == NQCB200_METHOD_FSSH && a == b) rho = (a == st) ? 1.0 : 0.0;
                acc += e.Z[i][a] * rho * e.Z[i][b];
            }
        pop[i] = acc;
    }
}

template <int N, int DPL, int L, int METHOD>
NQ_D void record_save(const KParams& p, Emitter& em, int lane, const Herm<N>& s, int st, const Eig<N>& e,
                      const double (&r)[DPL], const double (&v)[DPL], const double (&mass)[DPL]) {
    const uint32_t obs = p.observables;
    const int64_t T = p.ntraj;
    double adi[N], dia[N];
    adiabatic_population<N, METHOD>(s, st, adi);
    const bool need_dia = obs & ((1u << NQCB200_OBS_DIABATIC_POP) | (1u << NQCB200_OBS_POPCORR_DIABATIC) |
                                 (1u << NQCB200_OBS_SCATTERING_DIABATIC));
    if (need_dia) diabatic_population<N, METHOD>(s, st, e, dia);
    else {
#pragma unroll
        for (int i = 0; i < N; ++i) dia[i] = 0.0;
    }
    if (em.isave == 0 && em.active && (obs & ((1u << NQCB200_OBS_POPCORR_DIABATIC) | (1u << NQCB200_OBS_POPCORR_ADIABATIC)))) {
#pragma unroll
        for (int i = 0; i < N; ++i) { p.pop0[(int64_t)i * T + em.traj] = dia[i]; p.pop0[(int64_t)(N + i) * T + em.traj] = adi[i]; }
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
        for (int i = 0; i < N; ++i) p0[i] = (em.isave == 0) ? dia[i] : p.pop0[(int64_t)i * T + em.traj];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) em.emit(NQCB200_OBS_POPCORR_DIABATIC, i + N * j, p0[i] * dia[j]);
    }
    if (obs & (1u << NQCB200_OBS_POPCORR_ADIABATIC)) {
        double p0[N];
#pragma unroll
        for (int i = 0; i < N; ++i) p0[i] = (em.isave == 0) ? adi[i] : p.pop0[(int64_t)(N + i) * T + em.traj];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int i = 0; i < N; ++i) em.emit(NQCB200_OBS_POPCORR_ADIABATIC, i + N * j, p0[i] * adi[j]);
    }
    if (obs & ((1u << NQCB200_OBS_KINETIC) | (1u << NQCB200_OBS_POTENTIAL) | (1u << NQCB200_OBS_TOTAL_ENERGY))) {
        double kin = 0.0;
#pragma unroll
        for (int jj = 0; jj < DPL; ++jj) kin += mass[jj] * v[jj] * v[jj];
        kin = 0.5 * lane_sum<L>(kin);
        double pot = 0.0;
        if (METHOD == NQCB200_METHOD_FSSH) pot = select<N>(e.w, st);                     // fssh.jl:150-154
        else {
#pragma unroll
            for (int i = 0; i < N; ++i) pot += s.x[sidx(N, i, i)] * e.w[i];             // ehrenfest.jl:85-95
        }
        if (obs & (1u << NQCB200_OBS_KINETIC)) em.emit(NQCB200_OBS_KINETIC, 0, kin);
        if (obs & (1u << NQCB200_OBS_POTENTIAL)) em.emit(NQCB200_OBS_POTENTIAL, 0, pot);
        if (obs & (1u << NQCB200_OBS_TOTAL_ENERGY)) em.emit(NQCB200_OBS_TOTAL_ENERGY, 0, kin + pot);
    }
    if (obs & ((1u << NQCB200_OBS_POSITION) | (1u << NQCB200_OBS_VELOCITY))) {
        // every lane owns different dofs: broadcast each dof's value to lane 0 of the group in turn
        for (int dof = 0; dof < p.D; ++dof) {
            const int owner = dof % L, jj = dof / L;
            double rv = 0.0, vv = 0.0;
#pragma unroll
            for (int k = 0; k < DPL; ++k) { rv = (k == jj) ? r[k] : rv; vv = (k == jj) ? v[k] : vv; }
            if (L > 1) {
                const int src = (threadIdx.x & 31 & ~(L - 1)) + owner;
                rv = __shfl_sync(0xffffffffu, rv, src);
                vv = __shfl_sync(0xffffffffu, vv, src);
            }
            if (obs & (1u << NQCB200_OBS_POSITION)) em.emit(NQCB200_OBS_POSITION, dof, rv);
            if (obs & (1u << NQCB200_OBS_VELOCITY)) em.emit(NQCB200_OBS_VELOCITY, dof, vv);
        }
    }
    if (obs & (1u << NQCB200_OBS_DISCRETE_STATE)) em.emit(NQCB200_OBS_DISCRETE_STATE, 0, (double)(st + 1));
    const bool last = (em.isave == p.nsave - 1);
    if (obs & ((1u << NQCB200_OBS_SCATTERING) | (1u << NQCB200_OBS_SCATTERING_DIABATIC))) {
        // OutputStateResolvedScattering1D, DynamicsOutputs.jl:313-338: final frame only
        double r0 = r[0];
        if (L > 1) r0 = __shfl_sync(0xffffffffu, r0, threadIdx.x & 31 & ~(L - 1));
        const bool trans = r0 > 0.0;
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
                em.emit(NQCB200_OBS_SIGMA, j + N * k, s.X(j, k));
                em.emit(NQCB200_OBS_SIGMA, N * N + j + N * k, s.Y(j, k));
            }
    }
}

// ---- trajectory state <-> global memory (SoA [component][traj]) --------------------------------
template <int N, int DPL, int L>
struct Regs {
    double r[DPL], v[DPL], acc[DPL], mass[DPL], ba[DPL], bb[DPL];
    Herm<N> s;
    int st;
    double Zref[N][N];
    ElecParams<N> cur;
};

template <int N, int DPL, int L>
NQ_D void load_regs(const KParams& p, int64_t traj, int lane, Regs<N, DPL, L>& R, bool with_dynamics) {
    const int64_t T = p.ntraj;
#pragma unroll
    for (int jj = 0; jj < DPL; ++jj) {
        const int dof = lane + L * jj;
        const bool ok = dof < p.D;
        R.r[jj] = ok ? p.r[(int64_t)dof * T + traj] : 0.0;
        R.v[jj] = ok ? p.v[(int64_t)dof * T + traj] : 0.0;
        R.acc[jj] = (ok && with_dynamics) ? p.acc[(int64_t)dof * T + traj] : 0.0;
        R.mass[jj] = ok ? p.masses[dof] : 1.0;
        R.ba[jj] = (ok && p.bath_a) ? p.bath_a[dof] : 0.0;
        R.bb[jj] = (ok && p.bath_b) ? p.bath_b[dof] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = j; k < N; ++k) {
            R.s.x[sidx(N, j, k)] = p.sig_re[(int64_t)(j + N * k) * T + traj];
            if (k > j) R.s.y[aidx(N, j, k)] = p.sig_im[(int64_t)(j + N * k) * T + traj];
        }
    R.st = p.state ? p.state[traj] : 0;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = 0; k < N; ++k) R.Zref[j][k] = p.Zprev[(int64_t)(j + N * k) * T + traj];
    if (with_dynamics) {
#pragma unroll
        for (int i = 0; i < N; ++i) R.cur.E[i] = p.ecur[(int64_t)i * T + traj];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = j + 1; k < N; ++k) R.cur.g[aidx(N, j, k)] = p.ecur[(int64_t)(N + j + N * k) * T + traj];
    }
}

template <int N, int DPL, int L>
NQ_D void store_regs(const KParams& p, int64_t traj, int lane, const Regs<N, DPL, L>& R) {
    const int64_t T = p.ntraj;
#pragma unroll
    for (int jj = 0; jj < DPL; ++jj) {
        const int dof = lane + L * jj;
        if (dof < p.D) {
            p.r[(int64_t)dof * T + traj] = R.r[jj];
            p.v[(int64_t)dof * T + traj] = R.v[jj];
            p.acc[(int64_t)dof * T + traj] = R.acc[jj];
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = 0; k < N; ++k) {
                p.sig_re[(int64_t)(j + N * k) * T + traj] = R.s.X(j, k);
                p.sig_im[(int64_t)(j + N * k) * T + traj] = R.s.Y(j, k);
                p.Zprev[(int64_t)(j + N * k) * T + traj] = R.Zref[j][k];
            }
        if (p.state) p.state[traj] = R.st;
#pragma unroll
        for (int i = 0; i < N; ++i) p.ecur[(int64_t)i * T + traj] = R.cur.E[i];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = j + 1; k < N; ++k) p.ecur[(int64_t)(N + j + N * k) * T + traj] = R.cur.g[aidx(N, j, k)];
    }
}

// V(r) -> eigen (gauge-fixed) ; returns e
template <class M, int DPL, int L>
NQ_D void eval_eigen(const KParams& p, const double (&r)[DPL], const double (&ba)[DPL], const double (&bb)[DPL],
                     bool lane0, double (&Zref)[M::NS][M::NS], Eig<M::NS>& e) {
    constexpr int N = M::NS;
    double Vp[sym_size(N)];
    M::template potential_partial<DPL>(p.params, r, ba, bb, lane0, Vp);
    if (L > 1) {
#pragma unroll
        for (int i = 0; i < sym_size(N); ++i) Vp[i] = lane_sum<L>(Vp[i]);
    }
    sym_eigh<N>(Vp, e);
    fix_gauge<N>(e, Zref);
}

// acceleration contribution of one dof from its adiabatic derivative (packed symmetric)
template <int N, int METHOD>
NQ_D double force_from_adiab(const double (&Ap)[sym_size(N)], int st, const Herm<N>& s) {
    double f = 0.0;
    if (METHOD == NQCB200_METHOD_FSSH) {
#pragma unroll
        for (int i = 0; i < N; ++i) f = (i == st) ? -Ap[sidx(N, i, i)] : f;
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = i; j < N; ++j) f -= ((i == j) ? 1.0 : 2.0) * Ap[sidx(N, i, j)] * s.x[sidx(N, i, j)];
    }
    return f;
}

// TERM: TerminatingCallback instantiation (thread-per-trajectory kernels only).  A terminated trajectory skips the
// step body (its registers keep the final state) but still takes part in the block-collective save points, so the
// fixed-shape outputs carry its final state from the termination on; term_step tells the host where the series ends.
template <class M, int DPL, int L, int METHOD, bool TERM = false>
__global__ void __launch_bounds__(kBlockThreads) density_step_kernel(const __grid_constant__ KParams p) {
    static_assert(!TERM || L == 1, "termination masks exist for the thread-per-trajectory layout");
    constexpr int N = M::NS;
    __shared__ double smem[2 * (kBlockThreads / 32)];
    const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t traj = gthread / L;
    const int lane = (int)(gthread % L);
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const bool lane0 = (lane == 0);
    const int64_t T = p.ntraj;

    Regs<N, DPL, L> R;
    load_regs<N, DPL, L>(p, traj, lane, R, true);
    Eig<N> e;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        e.w[i] = R.cur.E[i];   // eigenvalues/vectors at the current position (for a save before any step)
#pragma unroll
        for (int k = 0; k < N; ++k) e.Z[i][k] = R.Zref[i][k];
    }
    unsigned long long nhops = 0, nfrus = 0;
    const double dt = p.dt, hdt = 0.5 * p.dt;
    long long term_step = -1;
    if (TERM) term_step = p.term_step[traj];

#pragma unroll 1
    for (int is = 0; is < p.nsteps; ++is) {
        const int64_t step = p.step0 + is;
        if (!TERM || term_step < 0) {
        const double t = p.t0 + dt * (double)step;
        const double tcur = (step == 0) ? 0.0 : t;   // Q1: electronic buffer starts at t = 0, all zero
        double vt[DPL];
#pragma unroll
        for (int jj = 0; jj < DPL; ++jj) {
            vt[jj] = fma(hdt, R.acc[jj], R.v[jj]);        // step_B!  steps.jl:3-5
            R.r[jj] = fma(dt, vt[jj], R.r[jj]);           // step_A!  steps.jl:6-8
        }
        eval_eigen<M, DPL, L>(p, R.r, R.ba, R.bb, lane0, R.Zref, e);
        ElecParams<N> nxt;
#pragma unroll
        for (int i = 0; i < N; ++i) nxt.E[i] = e.w[i];
#pragma unroll
        for (int i = 0; i < asym_size(N); ++i) nxt.g[i] = 0.0;
#pragma unroll
        for (int jj = 0; jj < DPL; ++jj) {
            double dVp[sym_size(N)], Ap[sym_size(N)];
            M::derivative_dof(p.params, R.r[jj], R.ba[jj], R.bb[jj], dVp);
            similarity<N>(dVp, e.Z, Ap);                                     // Z' dV Z
            const double f = force_from_adiab<N, METHOD>(Ap, R.st, R.s);     // pre-hop state / sigma_prev
            R.acc[jj] = div_fast(f, R.mass[jj]);
            R.v[jj] = fma(hdt, R.acc[jj], vt[jj]);
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int k = j + 1; k < N; ++k)   // d[j,k] = -adiab[j,k] / (w_j - w_k)
                    nxt.g[aidx(N, j, k)] += div_fast(-Ap[sidx(N, j, k)], e.w[j] - e.w[k]) * R.v[jj];
        }
        if (L > 1) {
#pragma unroll
            for (int i = 0; i < asym_size(N); ++i) nxt.g[i] = lane_sum<L>(nxt.g[i]);
        }
        propagate_density<N>(R.cur, tcur, nxt, t + dt, t, dt, R.s, p.tsit5_ha);

        if (METHOD == NQCB200_METHOD_FSSH) {
            const double xi = (p.rng == NQCB200_RNG_INJECTED)
                                  ? p.draws[(step - p.draws_step0) * T + traj]
                                  : philox_uniform(p.seed, (uint64_t)(p.traj_offset + traj), (uint64_t)step, 0u);
            // fewest_switches_probability! fssh.jl:96-108 (Q4) + select_new_state :110-121
            const int s0 = R.st;
            const double inv_ss = rcp_nb(R.s.Xsel(s0, s0));
            double cum = 0.0;
            int new_state = s0;
#pragma unroll
            for (int m = 0; m < N; ++m) {
                double g = 0.0;
                if (m != s0) g = 2.0 * (R.s.Xsel(m, s0) * inv_ss) * nxt.Gsel(s0, m) * dt;
                g = fmin(1.0, fmax(0.0, g));
                cum += g;
                if (new_state == s0 && m != s0 && cum > xi) new_state = m;
            }
            if (new_state != s0) {
                // execute_hop! -> rescale_velocity!  surface_hopping.jl:64-99
                bool accept = true;
                if (p.rescaling != NQCB200_RESCALE_OFF) {
                    double dvec[DPL];
                    double a = 0.0, b = 0.0, nrm2 = 0.0;
                    const double wn = select<N>(e.w, new_state), wo = select<N>(e.w, s0);
#pragma unroll
                    for (int jj = 0; jj < DPL; ++jj) {
                        double dVp[sym_size(N)], Ap[sym_size(N)];
                        M::derivative_dof(p.params, R.r[jj], R.ba[jj], R.bb[jj], dVp);
                        similarity<N>(dVp, e.Z, Ap);
                        double ano = 0.0;   // adiab[new, old]
#pragma unroll
                        for (int j = 0; j < N; ++j)
#pragma unroll
                            for (int k = j + 1; k < N; ++k)
                                ano = ((j == new_state && k == s0) || (k == new_state && j == s0)) ? Ap[sidx(N, j, k)] : ano;
                        const int dof = lane + L * jj;
                        const double d = (dof < p.D) ? -ano / (wn - wo) : 0.0;
                        dvec[jj] = d;
                        a += d * d / R.mass[jj];
                        b += d * R.v[jj];
                        nrm2 += d * d;
                    }
                    a = 0.5 * lane_sum<L>(a); b = lane_sum<L>(b); nrm2 = lane_sum<L>(nrm2);
                    const double c = wn - wo;
                    const double disc = b * b - 4.0 * a * c;
                    if (disc < 0.0) {
                        accept = false;
                        nfrus += (lane0 && valid);
                        if (p.rescaling == NQCB200_RESCALE_VINVERSION) {   // :155-164
                            const double nrm = sqrt(nrm2);
                            const double gam = b / nrm;
#pragma unroll
                            for (int jj = 0; jj < DPL; ++jj) R.v[jj] -= 2.0 * gam * dvec[jj] / nrm;
                        }
                    } else {
                        const double root = sqrt(disc);
                        const double gam = (b < 0.0) ? (b + root) / (2.0 * a) : (b - root) / (2.0 * a);
#pragma unroll
                        for (int jj = 0; jj < DPL; ++jj) R.v[jj] -= gam * dvec[jj] / R.mass[jj];
                    }
                }
                if (accept) { R.st = new_state; nhops += (lane0 && valid); }
                // Q2: R.acc keeps the pre-hop force; Q3: nxt.g keeps the pre-rescale velocity.
            }
        }
        R.cur = nxt;
        if (TERM) {   // DiscreteCallback(condition, terminate!) after the hopping callback, on the new u
            double x = R.r[0], vx = R.v[0];   // L == 1: DPL == D, every dof lives in this thread
#pragma unroll
            for (int jj = 1; jj < DPL; ++jj) { x = (jj == p.term_dof) ? R.r[jj] : x; vx = (jj == p.term_dof) ? R.v[jj] : vx; }
            const bool og = p.term_outgoing != 0;
            if ((x < p.term_lo && (!og || vx < 0.0)) || (x > p.term_hi && (!og || vx > 0.0)) || p.t0 + dt * (double)(step + 1) > p.term_tcut)
                term_step = step + 1;
        }
        }

        if ((step + 1) % p.save_every == 0) {
            const int64_t isave = (step + 1) / p.save_every;
            if (isave < p.nsave) {
                Emitter em{p, traj, valid && lane0, (int)isave, smem, 0};
                record_save<N, DPL, L, METHOD>(p, em, lane, R.s, R.st, e, R.r, R.v, R.mass);
            }
        }
    }

    if (valid) {
        store_regs<N, DPL, L>(p, traj, lane, R);
        if (TERM) p.term_step[traj] = term_step;
        if (p.diagnostics) {
            if (lane0) {
#pragma unroll
                for (int i = 0; i < N; ++i) p.diag_eig[(int64_t)i * T + traj] = e.w[i];
#pragma unroll
                for (int j = 0; j < N; ++j)
#pragma unroll
                    for (int k = 0; k < N; ++k) p.diag_Z[(int64_t)(j + N * k) * T + traj] = e.Z[j][k];
            }
#pragma unroll
            for (int jj = 0; jj < DPL; ++jj) {
                const int dof = lane + L * jj;
                if (dof < p.D) {
                    double dVp[sym_size(N)], Ap[sym_size(N)];
                    M::derivative_dof(p.params, R.r[jj], R.ba[jj], R.bb[jj], dVp);
                    similarity<N>(dVp, e.Z, Ap);
#pragma unroll
                    for (int j = 0; j < N; ++j)
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            double d = 0.0;
                            if (j != k) d = -((j < k) ? Ap[sidx(N, j, k)] : Ap[sidx(N, k, j)]) / (e.w[j] - e.w[k]);
                            p.diag_nac[((int64_t)dof * N * N + j + N * k) * T + traj] = d;
                        }
                }
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

// Initialisation: update_cache!(r0), optional diabatic -> adiabatic rotation of the density matrix
// (density_matrix_dynamics.jl:37-75), FSSH active-state sampling (fssh.jl:53-54), initial
// acceleration (bab_electronics.jl:48-59), zeroed electronic buffer (Q1), save point 0.
//   basis: 0 adiabatic, 1 diabatic ; sample_state: draw the FSSH state from Re diag(sigma)
template <class M, int DPL, int L, int METHOD>
__global__ void __launch_bounds__(kBlockThreads) density_init_kernel(const __grid_constant__ KParams p, int basis,
                                                                     int sample_state, const double* state_draw) {
    constexpr int N = M::NS;
    __shared__ double smem[2 * (kBlockThreads / 32)];
    const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t traj = gthread / L;
    const int lane = (int)(gthread % L);
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const bool lane0 = (lane == 0);

    Regs<N, DPL, L> R;
    load_regs<N, DPL, L>(p, traj, lane, R, false);
    Eig<N> e;
    eval_eigen<M, DPL, L>(p, R.r, R.ba, R.bb, lane0, R.Zref, e);
    if (basis == 1) {
        // sigma = Z' rho Z : X' = Z' X Z (symmetric), Y' = Z' Y Z (antisymmetric)
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
    if (METHOD == NQCB200_METHOD_FSSH && sample_state) {
        // StatsBase.sample(Weights(w)): t = rand() * sum(w); first i with cumsum(w)[i] >= t
        const double xi = state_draw ? state_draw[traj]
                                     : philox_uniform(p.seed, (uint64_t)(p.traj_offset + traj), 0ull, 1u);
        double tot = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) tot += R.s.x[sidx(N, i, i)];
        const double target = xi * tot;
        double cw = R.s.x[sidx(N, 0, 0)];
        int st = 0;
#pragma unroll
        for (int i = 1; i < N; ++i) {
            if (cw < target && st == i - 1) { st = i; cw += R.s.x[sidx(N, i, i)]; }
        }
        R.st = st;
    }
#pragma unroll
    for (int jj = 0; jj < DPL; ++jj) {
        double dVp[sym_size(N)], Ap[sym_size(N)];
        M::derivative_dof(p.params, R.r[jj], R.ba[jj], R.bb[jj], dVp);
        similarity<N>(dVp, e.Z, Ap);
        R.acc[jj] = force_from_adiab<N, METHOD>(Ap, R.st, R.s) / R.mass[jj];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) R.cur.E[i] = 0.0;
#pragma unroll
    for (int i = 0; i < asym_size(N); ++i) R.cur.g[i] = 0.0;
    {
        Emitter em{p, traj, valid && lane0, 0, smem, 0};
        record_save<N, DPL, L, METHOD>(p, em, lane, R.s, R.st, e, R.r, R.v, R.mass);
    }
    if (valid) {
        store_regs<N, DPL, L>(p, traj, lane, R);
        // the step kernel reads eigenvalues at the current position from ecur only for saves that
        // precede the first step, which cannot happen (save 0 is recorded here): keep Q1 zeros.
    }
}

#endif  // __CUDACC__

}  // namespace nq

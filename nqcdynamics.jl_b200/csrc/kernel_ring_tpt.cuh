// kernel_ring_tpt.cuh -- RPSH / RP-Ehrenfest with ONE THREAD per trajectory.
//
// Same reference path as kernel_ring.cuh (BCBwithTsit5, bcb_electronics.jl:53-97; centroid hopping,
// SurfaceHoppingMethods.jl:85-103, rpsh.jl:30-64; ehrenfest_rpmd.jl:23-51), different decomposition.  With the
// beads on lanes, the centroid eigenproblem and the 31-RHS Tsit5 electronic propagation -- 3/4 of the work for
// ThreeStateMorse -- were replicated on all NB lanes of a trajectory (ncu, profiles/r01: 125 k FP64 instructions
// per trajectory-step, of which ~95 k redundant).  Here a thread owns the whole ring polymer:
//   * bead coordinates, velocities and accelerations live in shared memory, [field][bead][thread]
//     (conflict-free), and are visited bead by bead (model, NxN Jacobi, force);
//   * the free ring-polymer step is an in-register radix-2 FFT of z_b = r_b + i v_b (same algebra as
//     FreeRingPolymer in kernel_ring.cuh, coefficient tables in shared memory);
//   * the centroid eigenproblem, NAC, Tsit5, hop and rescaling are done once per trajectory in registers.
// FSSH needs no per-bead eigenvector gauge (the bead force is the diagonal element of Z' dV Z, which does not
// depend on the column signs); RP-Ehrenfest uses off-diagonal elements, so its per-bead gauge references are
// carried in shared memory as well.
#pragma once
#include "kernel_ring.cuh"

namespace nq {

#if defined(__CUDACC__)

constexpr int kRtThreads = kBlockThreads;   // 128 (the Emitter's block reduction is sized for it)
constexpr int kRpshMaxThreads = 384;   // ring_tpt_step_kernel: ONE block per SM, up to 12 warps at 168 registers

// fft: nbeads is a power of two (compile-time NBT); otherwise the dense normal-mode product through two scratch rows
// threads = trajectory slots per block (block size / lanes per trajectory); lanes > 1 adds the [bead][N] energy scratch
NQ_HD constexpr size_t ring_tpt_smem_bytes(int N, int NB, bool ehrenfest, bool fft, int threads = kRtThreads, bool lanes = false) {
    return ((size_t)(3 + (ehrenfest ? N * N : 0) + (fft ? 0 : 2) + (lanes ? N : 0)) * NB * threads + (lanes ? threads : 0) +
            (fft ? 6 * NB : NB * NB + 4 * NB)) * sizeof(double);
}

// Block size of ring_tpt_step_kernel for a batch of ntraj trajectories on `sms` SMs: the kernel runs one block per SM and
// a trajectory cannot be split, so the batch takes ceil(blocks / sms) waves whatever the block size -- pick the size
// (a multiple of 32 up to kRpshMaxThreads and up to what shared memory holds) that minimises waves x time per wave, with
// the time per wave of a block of w warps ~ (w + 4) (FP64 latency hiding saturates slowly: 8 warps 1.52e9, 12 warps
// 1.50e9 trajectory-steps/s on whole waves).  BASELINE config 5, 10^5 trajectories on 148 SMs: 352 threads, 2 waves
// (96 % full) instead of 3 waves of 256 (88 %).
inline int ring_tpt_block_threads(int64_t ntraj, int sms, int max_threads) {
    int best = 32;
    double best_cost = 1e300;
    for (int b = 32; b <= max_threads; b += 32) {
        const int64_t blocks = (ntraj + b - 1) / b;
        const int64_t waves = (blocks + sms - 1) / sms;
        const double cost = (double)waves * (b / 32 + 4);
        if (cost < best_cost - 1e-12) { best_cost = cost; best = b; }
    }
    return best;
}

template <int NB>
struct RtTables {      // per block, thread-uniform
    const double* twr;  // [NB/2]  Re W^j, W = exp(-2 pi i / NB)
    const double* twi;  // [NB/2]
    const double* al;   // [2 NB]  alpha_k (re, im) with 1/NB folded in
    const double* be;   // [2 NB]  beta_k
};

template <int NB>
NQ_D constexpr int rt_log2() { return (NB >= 32) ? 5 : (NB >= 16) ? 4 : (NB >= 8) ? 3 : (NB >= 4) ? 2 : (NB >= 2) ? 1 : 0; }
template <int NB>
NQ_D constexpr int rt_bitrev(int x) {
    int r = 0;
    for (int b = 0; b < rt_log2<NB>(); ++b) r |= ((x >> b) & 1) << (rt_log2<NB>() - 1 - b);
    return r;
}

// to-normal-modes -> Cayley -> back for one ring polymer held in registers (see FreeRingPolymer for the algebra)
template <int NB, bool NOISE = false>
NQ_D void rt_free_step(const RtTables<NB>& tb, double (&zr)[NB], double (&zi)[NB], const double* xi_ = nullptr,
                       const double* bn = nullptr, const double* dn = nullptr) {
    constexpr int LOG = rt_log2<NB>();
#pragma unroll
    for (int s = 0; s < LOG; ++s) {          // DIF: natural in, bit-reversed out
        const int h = NB >> (s + 1);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if ((i & h) == 0) {
                const int j = i & (h - 1), tw = j * (NB / (2 * h));
                const double ur = zr[i], ui = zi[i], tr = zr[i + h], ti = zi[i + h];
                zr[i] = ur + tr; zi[i] = ui + ti;
                const double dr = ur - tr, di = ui - ti;
                if (tw == 0) { zr[i + h] = dr; zi[i + h] = di; }
                else {
                    const double wr = tb.twr[tw], wi = tb.twi[tw];
                    zr[i + h] = fma(dr, wr, -di * wi); zi[i + h] = fma(dr, wi, di * wr);
                }
            }
        }
    }
    // Z'_k = alpha_k Z_k + beta_k conj(Z_{NB-k}), in place: the pair (k, NB-k) is updated together (k = 0 and NB/2
    // pair with themselves), so no second copy of the spectrum is live
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int k = rt_bitrev<NB>(i), ip = rt_bitrev<NB>((NB - k) % NB);
        if (ip < i) continue;                 // handled with its partner
        const double ar = tb.al[2 * k], ai = tb.al[2 * k + 1], br = tb.be[2 * k], bi = tb.be[2 * k + 1];
        const double xr = zr[i], xi = zi[i], yr = zr[ip], yi = zi[ip];
        zr[i] = fma(ar, xr, fma(-ai, xi, fma(br, yr, bi * yi)));          // W = conj(Z_partner) = (yr, -yi)
        zi[i] = fma(ar, xi, fma(ai, xr, fma(-br, yi, bi * yr)));
        if (ip != i) {
            const int kp = (NB - k) % NB;
            const double cr = tb.al[2 * kp], ci = tb.al[2 * kp + 1], dr = tb.be[2 * kp], di = tb.be[2 * kp + 1];
            zr[ip] = fma(cr, yr, fma(-ci, yi, fma(dr, xr, di * xi)));
            zi[ip] = fma(cr, yi, fma(ci, yr, fma(-dr, xi, di * xr)));
        }
        if (NOISE) {      // Z'_k += (b_k + i d_k) N_k, N_k = n_lo + i n_hi for the pair (lo, NB - lo), real for k = 0, NB/2
            const int lo = (k <= NB - k) ? k : NB - k;          // mode with the cosine: min(k, NB - k); k = 0 -> 0
            const int hi = (NB - lo) % NB;
            const double b0 = bn[lo], d0 = dn[lo];
            if (ip == i) {
                zr[i] = fma(b0, xi_[k], zr[i]); zi[i] = fma(d0, xi_[k], zi[i]);
            } else {
                const double nc = xi_[lo], ns = xi_[hi];
                const double sgn = (k == lo) ? 1.0 : -1.0;      // position i holds mode lo (+) or NB - lo (conjugate)
                zr[i] += b0 * nc - sgn * d0 * ns;  zi[i] += sgn * b0 * ns + d0 * nc;
                zr[ip] += b0 * nc + sgn * d0 * ns; zi[ip] += -sgn * b0 * ns + d0 * nc;
            }
        }
    }
#pragma unroll
    for (int s = LOG - 1; s >= 0; --s) {     // DIT: bit-reversed in, natural out
        const int h = NB >> (s + 1);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if ((i & h) == 0) {
                const int j = i & (h - 1), tw = j * (NB / (2 * h));
                double tr = zr[i + h], ti = zi[i + h];
                if (tw != 0) {
                    const double wr = tb.twr[tw], wi = tb.twi[tw];
                    const double xr = fma(tr, wr, ti * wi), xi = fma(ti, wr, -tr * wi);   // conj(twiddle)
                    tr = xr; ti = xi;
                }
                const double ur = zr[i], ui = zi[i];
                zr[i] = ur + tr; zi[i] = ui + ti;
                zr[i + h] = ur - tr; zi[i + h] = ui - ti;
            }
        }
    }
}

// r_b = s_r[b * stride + idx] (shared memory: stride 128, idx = thread; global memory: stride T, idx = trajectory)
template <int N, int METHOD>
NQ_D void rt_record_save(const KParams& p, Emitter& em, int NB, const double* s_r, const double* s_v, int64_t stride,
                         int64_t idx, const Herm<N>& s, int st, const Eig<N>& ec, double pot, double mass) {
    const uint32_t obs = p.observables;
    const int64_t T = p.ntraj;
    double adi[N], dia[N];
    adiabatic_population<N, METHOD>(s, st, adi);
    diabatic_population<N, METHOD>(s, st, ec, dia);   // centroid transformation (density_matrix_dynamics.jl:83-87)
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
    double rsum = 0.0, vsum = 0.0, mv2 = 0.0, spr = 0.0;
    {
        double rprev = s_r[(NB - 1) * stride + idx];
        for (int b = 0; b < NB; ++b) {
            const double rb = s_r[b * stride + idx], vb = s_v[b * stride + idx];
            rsum += rb; vsum += vb;
            mv2 = fma(mass * vb, vb, mv2);
            const double d = rprev - rb;       // every neighbouring pair once (ring_polymer.jl:89-107)
            spr = fma(mass * d, d, spr);
            rprev = rb;
        }
    }
    if (obs & ((1u << NQCB200_OBS_KINETIC) | (1u << NQCB200_OBS_POTENTIAL) | (1u << NQCB200_OBS_TOTAL_ENERGY))) {
        const double kin = 0.5 * mv2;
        if (obs & (1u << NQCB200_OBS_KINETIC)) em.emit(NQCB200_OBS_KINETIC, 0, kin);
        if (obs & (1u << NQCB200_OBS_POTENTIAL)) em.emit(NQCB200_OBS_POTENTIAL, 0, pot);
        if (obs & (1u << NQCB200_OBS_TOTAL_ENERGY))
            em.emit(NQCB200_OBS_TOTAL_ENERGY, 0, kin + pot + 0.5 * p.omega_n * p.omega_n * spr);
    }
    if (obs & (1u << NQCB200_OBS_POSITION)) em.emit(NQCB200_OBS_POSITION, 0, rsum / NB);
    if (obs & (1u << NQCB200_OBS_VELOCITY)) em.emit(NQCB200_OBS_VELOCITY, 0, vsum / NB);
    if (obs & (1u << NQCB200_OBS_DISCRETE_STATE)) em.emit(NQCB200_OBS_DISCRETE_STATE, 0, (double)(st + 1));
    const bool last = (em.isave == p.nsave - 1);
    if (obs & ((1u << NQCB200_OBS_SCATTERING) | (1u << NQCB200_OBS_SCATTERING_DIABATIC))) {
        const bool trans = s_r[idx] > 0.0;   // get_positions(final)[1]: first dof of the first bead (DynamicsOutputs.jl:332)
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

// NBT > 0: nbeads = NBT, a power of two (FFT in registers).  NBT == 0: any nbeads (p.B), dense normal-mode product.
// TERM: TerminatingCallback instantiation (nqcb200_set_termination): the position-window predicate on the CENTROID of the
// chosen dof (and its centroid velocity for the `outgoing` clause), tested after the hopping callback on the new u
// (callbacks.jl:29); a terminated trajectory skips the step body but keeps taking part in the barriers and save points.
// LPT > 1 (FSSH, register FFT, no TERM): warp-specialised phases for shards smaller than one wave of threads (BASELINE
// config 5 on 8 GPUs: 12 500 trajectories per GPU), where a thread-per-trajectory launch leaves most of every SM empty and
// runs at the latency of one dependent chain.  The block holds KS = blockDim / LPT trajectories; thread t works for
// trajectory t % KS as member t / KS of its group, so the OWNERS (member 0) fill the first KS / 32 warps.  Per step:
// owners run the free ring-polymer step; barrier; ALL warps visit bead pairs (member g the beads [g NB/LPT, (g+1) NB/LPT),
// the bead loop is ~55 % of a step); barrier; owners take the bead sums in bead order from shared memory and run the
// centroid eigenproblem, Tsit5 and the hop test while the helper warps wait at the next barrier (whole warps: no issue
// slots).  A first version with the LPT members on adjacent lanes repeating the owner's work was slower than one lane
// (profiles/r02/SUMMARY.md).  Same bead pairs and the same summation order as LPT = 1: results do not depend on LPT.
template <class M, int NBT, int METHOD, bool TERM = false, int LPT = 1>
__global__ void __launch_bounds__(kRpshMaxThreads, 1) ring_tpt_step_kernel(const __grid_constant__ KParams p) {
    constexpr int N = M::NS;
    constexpr bool EHR = (METHOD == NQCB200_METHOD_EHRENFEST);
    constexpr bool FFT = NBT > 0;
    static_assert(NBT == 0 || (NBT >= 2 && (NBT & (NBT - 1)) == 0), "ring polymer FFT: nbeads must be a power of two");
    constexpr int NBF = FFT ? NBT : 2;      // array extent of the FFT path (unused when dense)
    constexpr bool LANES = LPT > 1;
    static_assert(!LANES || (FFT && !EHR && !TERM && NBT % (2 * LPT) == 0), "lanes per trajectory: FSSH, register FFT, whole bead pairs per lane");
    const int NB = FFT ? NBT : p.B;
    extern __shared__ __align__(16) double rt_sm[];
    __shared__ double red[2 * (kRpshMaxThreads / 32)];
    const int KT = blockDim.x;      // chosen on the host (ring_tpt_block_threads): whole waves, at most 12 warps
    double* s_r = rt_sm;
    const int KSd = blockDim.x / LPT;
    double* s_v = s_r + NB * KSd;
    double* s_a = s_v + NB * KSd;
    double* s_Z = s_a + NB * KSd;                      // EHR only: [bead][N*N][thread]
    double* s_t = s_Z + (EHR ? N * N * NB * KSd : 0);   // dense only: two scratch rows [2][bead][thread]
    double* s_w = s_t + (FFT ? 0 : 2 * NB * KSd);       // LANES only: per-bead adiabatic energies of a saving step [bead][N][slot]
    int* s_st = reinterpret_cast<int*>(s_w + (LANES ? NB * N * KSd : 0));     // LANES only: occupied state per trajectory (owner -> helpers)
    double* s_tab = s_w + (LANES ? NB * N * KSd + KSd : 0);   // FFT: twr[NB/2] twi[NB/2] al[2NB] be[2NB]; dense: U[NB*NB] cay[4NB]
    const int tid = threadIdx.x;
    const int KS = KT / LPT;                // trajectories per block
    const int slot = LANES ? tid % KS : tid, sub = LANES ? tid / KS : 0;
    int64_t traj = (int64_t)blockIdx.x * KS + slot;
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const int64_t T = p.ntraj;

    // coefficient tables (thread-uniform)
    if (!FFT) {
        for (int i = tid; i < NB * NB; i += KT) s_tab[i] = p.nm_to[i];            // U[j,k] at j*NB + k
        for (int i = tid; i < 4 * NB; i += KT) s_tab[NB * NB + i] = p.cayley[i];
    }
    for (int j = tid; FFT && j < NB; j += KT) {
        if (j < NB / 2) {
            double si, co;
            sincospi(-2.0 * (double)j / (double)NB, &si, &co);
            s_tab[j] = co; s_tab[NB / 2 + j] = si;
        }
        const double a = p.cayley[4 * j + 0], b = p.cayley[4 * j + 1], c = p.cayley[4 * j + 2], d = p.cayley[4 * j + 3];
        const double inv = 0.5 / NB;
        s_tab[NB + 2 * j] = (a + d) * inv; s_tab[NB + 2 * j + 1] = (c - b) * inv;
        s_tab[3 * NB + 2 * j] = (a - d) * inv; s_tab[3 * NB + 2 * j + 1] = (c + b) * inv;
    }
    RtTables<NBF> tb{s_tab, s_tab + NB / 2, s_tab + NB, s_tab + 3 * NB};

    for (int b = sub; b < NB; b += LPT) {
        s_r[b * KS + slot] = p.r[(int64_t)b * T + traj];
        s_v[b * KS + slot] = p.v[(int64_t)b * T + traj];
        s_a[b * KS + slot] = p.acc[(int64_t)b * T + traj];
        if (EHR) {
            for (int jk = 0; jk < N * N; ++jk)
                s_Z[(b * N * N + jk) * KS + slot] = p.Zprev[((int64_t)b * N * N + jk) * T + traj];
        }
    }
    __syncthreads();

    const double mass = p.masses[0], rmass = 1.0 / mass;
    Herm<N> s;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = j; k < N; ++k) {
            s.x[sidx(N, j, k)] = p.sig_re[(int64_t)(j + N * k) * T + traj];
            if (k > j) s.y[aidx(N, j, k)] = p.sig_im[(int64_t)(j + N * k) * T + traj];
        }
    int st = p.state ? p.state[traj] : 0;
    if (LANES && sub == 0) s_st[slot] = st;
    double Zc[N][N];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = 0; k < N; ++k) Zc[j][k] = p.Zprev[((int64_t)NB * N * N + j + N * k) * T + traj];
    ElecParams<N> cur;
#pragma unroll
    for (int i = 0; i < N; ++i) cur.E[i] = p.ecur[(int64_t)i * T + traj];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = j + 1; k < N; ++k) cur.g[aidx(N, j, k)] = p.ecur[(int64_t)(N + j + N * k) * T + traj];

    Eig<N> ec = {};
    double Ac[sym_size(N)];
    unsigned long long nhops = 0, nfrus = 0;
    const double dt = p.dt, hdt = 0.5 * p.dt;
    long long term_step = -1;
    if (TERM) term_step = p.term_step[traj];
    double wsum[N];      // sum over beads of the adiabatic energies (potential outputs use the post-hop state)
    if (TERM && term_step >= 0) {
        // terminated in an earlier launch: the estimators of its frozen state for this launch's save points
#pragma unroll
        for (int i = 0; i < N; ++i) { wsum[i] = 0.0; ec.w[i] = cur.E[i]; }
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k2 = 0; k2 < N; ++k2) ec.Z[j][k2] = Zc[j][k2];
        for (int b = 0; b < NB; ++b) {
            double Vp[sym_size(N)], dVp[sym_size(N)];
            Eig<N> eb;
            model_value_and_derivative<M>(p.params, s_r[b * KS + slot], Vp, dVp);
            sym_eigh<N>(Vp, eb);
#pragma unroll
            for (int i = 0; i < N; ++i) wsum[i] += eb.w[i];
        }
    }

#pragma unroll 1
    for (int is = 0; is < p.nsteps; ++is) {
        const int64_t step = p.step0 + is;
        const double t = p.t0 + dt * (double)step;
        const double tcur = (step == 0) ? 0.0 : t;   // Q1
        // one barrier per step keeps the block's warps inside the same ~15 KB of code (the step loop is ~60 KB of SASS
        // against a 32 KB instruction cache: ncu showed `no_instruction` as the largest stall, 25 % of all samples,
        // spread evenly; with the barrier +13 %, profiles/r02/SUMMARY.md)
        __syncthreads();
        if (!TERM || term_step < 0) {
        // B (half kick) + C (free ring polymer)  bcb_electronics.jl:62-71
        if (FFT && (!LANES || sub == 0)) {
            double zr[NBF], zi[NBF];
#pragma unroll
            for (int b = 0; b < NBF; ++b) {
                zr[b] = s_r[b * KS + slot];
                zi[b] = fma(hdt, s_a[b * KS + slot], s_v[b * KS + slot]);
            }
            rt_free_step<NBF>(tb, zr, zi);
#pragma unroll
            for (int b = 0; b < NBF; ++b) { s_r[b * KS + slot] = zr[b]; s_v[b * KS + slot] = zi[b]; }
        } else if (!FFT) {
            // dense U' .. Cayley .. U (RingPolymerArrays transform!, steps.jl:10-17)
            const double* U = s_tab;
            const double* cay = s_tab + NB * NB;
            for (int b = 0; b < NB; ++b) s_v[b * KS + slot] = fma(hdt, s_a[b * KS + slot], s_v[b * KS + slot]);
            for (int k = 0; k < NB; ++k) {
                double a = 0.0, c = 0.0;
                for (int j = 0; j < NB; ++j) {
                    const double u = U[j * NB + k];
                    a = fma(u, s_r[j * KS + slot], a);
                    c = fma(u, s_v[j * KS + slot], c);
                }
                s_t[k * KS + slot] = cay[4 * k + 0] * a + cay[4 * k + 1] * c;
                s_t[(NB + k) * KS + slot] = cay[4 * k + 2] * a + cay[4 * k + 3] * c;
            }
            for (int j = 0; j < NB; ++j) {
                double a = 0.0, c = 0.0;
                for (int k = 0; k < NB; ++k) {
                    const double u = U[j * NB + k];
                    a = fma(u, s_t[k * KS + slot], a);
                    c = fma(u, s_t[(NB + k) * KS + slot], c);
                }
                s_r[j * KS + slot] = a;
                s_v[j * KS + slot] = c;
            }
        }
        if (LANES) { __syncthreads(); st = s_st[slot]; }      // the owners' beads and occupied state for every member
        // update_cache! on every bead (bcb_electronics.jl:73), force, second half kick
        double rsum = 0.0, vsum = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) wsum[i] = 0.0;
        if constexpr (!EHR) {
            // FSSH: the bead force is -(Z' dV Z)[st, st] (fssh.jl:67-74) -- one eigenvector column, no gauge -- so the beads
            // are visited two at a time with their Jacobi chains in lock step and without sorting the eigenvectors
            // (ncu, profiles/r02: the kernel waits on fixed-latency dependencies 39 % of its idle issue slots; the
            // sorting network was 14 % of the executed instructions)
            const bool saving = TERM || ((step + 1) % p.save_every == 0);    // TERM: this step may be the trajectory's last
            const int nb_lane = NB / LPT, b_lo = sub * nb_lane;        // this lane's beads (all of them when LPT == 1)
#pragma unroll 1
            for (int bb = 0; bb < nb_lane; bb += 2) {
                const int b = b_lo + bb;
                const int b1 = FFT ? b + 1 : ((b + 1 < NB) ? b + 1 : b);     // odd bead counts (dense path): the last bead twice
                const double q0 = s_r[b * KS + slot], q1 = s_r[b1 * KS + slot];
                double V0[sym_size(N)], dV0[sym_size(N)], V1[sym_size(N)], dV1[sym_size(N)];
                model_value_and_derivative<M>(p.params, q0, V0, dV0);
                model_value_and_derivative<M>(p.params, q1, V1, dV1);
                double w0[N], w1[N], Z0[N][N], Z1[N][N];
                sym_eigh_pair_unsorted<N>(V0, V1, w0, Z0, w1, Z1);
                int rk0[N], rk1[N];
                eig_rank<N>(w0, rk0);
                eig_rank<N>(w1, rk1);
                double z0[N], z1[N];
#pragma unroll
                for (int a = 0; a < N; ++a) {
                    z0[a] = 0.0; z1[a] = 0.0;
#pragma unroll
                    for (int i = 0; i < N; ++i) { z0[a] = (rk0[i] == st) ? Z0[a][i] : z0[a]; z1[a] = (rk1[i] == st) ? Z1[a][i] : z1[a]; }
                }
                double f0 = 0.0, f1 = 0.0;
#pragma unroll
                for (int a = 0; a < N; ++a) {
                    double row0 = 0.0, row1 = 0.0;
#pragma unroll
                    for (int c = 0; c < N; ++c) {
                        row0 += dV0[(a <= c) ? sidx(N, a, c) : sidx(N, c, a)] * z0[c];
                        row1 += dV1[(a <= c) ? sidx(N, a, c) : sidx(N, c, a)] * z1[c];
                    }
                    f0 += z0[a] * row0; f1 += z1[a] * row1;
                }
                if (saving) {
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        double a0 = 0.0, a1 = 0.0;
#pragma unroll
                        for (int i = 0; i < N; ++i) { a0 = (rk0[i] == k) ? w0[i] : a0; a1 = (rk1[i] == k) ? w1[i] : a1; }
                        if (LANES) { s_w[(b * N + k) * KS + slot] = a0; s_w[(b1 * N + k) * KS + slot] = a1; }
                        else {
                            wsum[k] += a0;
                            if (FFT || b1 != b) wsum[k] += a1;
                        }
                    }
                }
                {
                    const double acc = div_nb(-f0, mass, rmass);
                    const double vb = fma(hdt, acc, s_v[b * KS + slot]);
                    s_a[b * KS + slot] = acc;
                    s_v[b * KS + slot] = vb;
                    if (!LANES) { rsum += q0; vsum += vb; }
                }
                if (FFT || b1 != b) {
                    const double acc = div_nb(-f1, mass, rmass);
                    const double vb = fma(hdt, acc, s_v[b1 * KS + slot]);
                    s_a[b1 * KS + slot] = acc;
                    s_v[b1 * KS + slot] = vb;
                    if (!LANES) { rsum += q1; vsum += vb; }
                }
            }
            if (LANES) {       // bead sums in bead order (the order of the single-lane kernel), from shared memory
                __syncthreads();
                for (int b = 0; b < NB; ++b) { rsum += s_r[b * KS + slot]; vsum += s_v[b * KS + slot]; }
                if (saving) {
                    for (int b = 0; b < NB; ++b)
#pragma unroll
                        for (int k = 0; k < N; ++k) wsum[k] += s_w[(b * N + k) * KS + slot];
                }
            }
        } else {
#pragma unroll 1
        for (int b = 0; b < NB; ++b) {
            const double q = s_r[b * KS + slot];
            double Vp[sym_size(N)], dVp[sym_size(N)];
            Eig<N> eb;
            model_value_and_derivative<M>(p.params, q, Vp, dVp);
            sym_eigh<N>(Vp, eb);
            double Zb[N][N];
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int k = 0; k < N; ++k) Zb[j][k] = s_Z[(b * N * N + j + N * k) * KS + slot];
            fix_gauge<N>(eb, Zb);
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int k = 0; k < N; ++k) s_Z[(b * N * N + j + N * k) * KS + slot] = Zb[j][k];
            double Ab[sym_size(N)];
            similarity<N>(dVp, eb.Z, Ab);
            const double f = force_from_adiab<N, METHOD>(Ab, st, s);
#pragma unroll
            for (int i = 0; i < N; ++i) wsum[i] += eb.w[i];
            const double acc = f / mass;
            const double vb = fma(hdt, acc, s_v[b * KS + slot]);
            s_a[b * KS + slot] = acc;
            s_v[b * KS + slot] = vb;
            rsum += q; vsum += vb;
        }
        }
        if (!LANES || sub == 0) {          // LANES: the owner warps; the helper warps go on to the next barrier
        const double rcent = rsum / NB, vcent = vsum / NB;
        eval_point<M>(p, rcent, Zc, ec, Ac);
        ElecParams<N> nxt;
#pragma unroll
        for (int i = 0; i < N; ++i) nxt.E[i] = ec.w[i];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = j + 1; k < N; ++k) nxt.g[aidx(N, j, k)] = div_fast(-Ac[sidx(N, j, k)], ec.w[j] - ec.w[k]) * vcent;
        propagate_density<N>(cur, tcur, nxt, t + dt, t, dt, s, p.tsit5_ha);
        double dv_hop = 0.0;      // velocity change of the hop callback (every bead, hence the centroid)

        if (METHOD == NQCB200_METHOD_FSSH) {
            const double xi = (p.rng == NQCB200_RNG_INJECTED)
                                  ? p.draws[(step - p.draws_step0) * T + traj]
                                  : philox_uniform(p.seed, (uint64_t)(p.traj_offset + traj), (uint64_t)step, 0u);
            const int s0 = st;
            const double inv_ss = rcp_nb(s.Xsel(s0, s0));
            double cum = 0.0;
            int new_state = s0;
#pragma unroll
            for (int m = 0; m < N; ++m) {
                double g = 0.0;
                if (m != s0) g = 2.0 * (s.Xsel(m, s0) * inv_ss) * nxt.Gsel(s0, m) * dt;
                g = fmin(1.0, fmax(0.0, g));
                cum += g;
                if (new_state == s0 && m != s0 && cum > xi) new_state = m;
            }
            if (new_state != s0) {
                bool accept = true;
                double dv = 0.0;     // the same velocity change on every bead (rpsh.jl:30-50)
                if (p.rescaling != NQCB200_RESCALE_OFF) {
                    const double wn = select<N>(ec.w, new_state), wo = select<N>(ec.w, s0);
                    double ano = 0.0;
#pragma unroll
                    for (int j = 0; j < N; ++j)
#pragma unroll
                        for (int k = j + 1; k < N; ++k)
                            ano = ((j == new_state && k == s0) || (k == new_state && j == s0)) ? Ac[sidx(N, j, k)] : ano;
                    const double d = -ano / (wn - wo);
                    const double a = 0.5 * d * d / mass, b = d * vcent, c = wn - wo;
                    const double disc = b * b - 4.0 * a * c;
                    if (disc < 0.0) {
                        accept = false;
                        nfrus += (valid && sub == 0);
                        if (p.rescaling == NQCB200_RESCALE_VINVERSION) {
                            const double dn = d / fabs(d);
                            dv = -2.0 * (vcent * dn) * dn;
                        }
                    } else {
                        const double root = sqrt(disc);
                        const double gam = (b < 0.0) ? (b + root) / (2.0 * a) : (b - root) / (2.0 * a);
                        dv = -(gam * d / mass);
                    }
                }
                if (dv != 0.0) {
                    for (int b = 0; b < NB; ++b) s_v[b * KS + slot] += dv;
                    dv_hop = dv;
                }
                if (accept) { st = new_state; nhops += (valid && sub == 0); if (LANES) s_st[slot] = st; }
            }
        }
        cur = nxt;
        if (TERM) {
            const double vx = vcent + dv_hop;
            const bool og = p.term_outgoing != 0;
            if ((rcent < p.term_lo && (!og || vx < 0.0)) || (rcent > p.term_hi && (!og || vx > 0.0)) || p.t0 + dt * (double)(step + 1) > p.term_tcut)
                term_step = step + 1;
        }
        }
        }

        if ((step + 1) % p.save_every == 0) {
            const int64_t isave = (step + 1) / p.save_every;
            if (isave < p.nsave) {
                double pot = 0.0;     // rpsh.jl:52-64, ehrenfest_rpmd.jl:45-51
                if (EHR) {
#pragma unroll
                    for (int i = 0; i < N; ++i) pot += s.x[sidx(N, i, i)] * wsum[i];
                } else pot = select<N>(wsum, st);
                Emitter em{p, traj, valid && sub == 0, (int)isave, red, 0, true, KT / 32};
                rt_record_save<N, METHOD>(p, em, NB, s_r, s_v, KS, slot, s, st, ec, pot, mass);
            }
        }
    }

    if (LANES) __syncthreads();
    if (valid && sub == 0) {
        for (int b = 0; b < NB; ++b) {
            p.r[(int64_t)b * T + traj] = s_r[b * KS + slot];
            p.v[(int64_t)b * T + traj] = s_v[b * KS + slot];
            p.acc[(int64_t)b * T + traj] = s_a[b * KS + slot];
            if (EHR) {
                for (int jk = 0; jk < N * N; ++jk)
                    p.Zprev[((int64_t)b * N * N + jk) * T + traj] = s_Z[(b * N * N + jk) * KS + slot];
            }
        }
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = 0; k < N; ++k) {
                p.sig_re[(int64_t)(j + N * k) * T + traj] = s.X(j, k);
                p.sig_im[(int64_t)(j + N * k) * T + traj] = s.Y(j, k);
                p.Zprev[((int64_t)NB * N * N + j + N * k) * T + traj] = Zc[j][k];
            }
        if (p.state) p.state[traj] = st;
        if (TERM) p.term_step[traj] = term_step;
#pragma unroll
        for (int i = 0; i < N; ++i) p.ecur[(int64_t)i * T + traj] = cur.E[i];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = j + 1; k < N; ++k) p.ecur[(int64_t)(N + j + N * k) * T + traj] = cur.g[aidx(N, j, k)];
        if (p.diagnostics) {
#pragma unroll
            for (int i = 0; i < N; ++i) p.diag_eig[(int64_t)i * T + traj] = ec.w[i];
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    p.diag_Z[(int64_t)(j + N * k) * T + traj] = ec.Z[j][k];
                    double d = 0.0;
                    if (j != k) d = -((j < k) ? Ac[sidx(N, j, k)] : Ac[sidx(N, k, j)]) / (ec.w[j] - ec.w[k]);
                    p.diag_nac[(int64_t)(j + N * k) * T + traj] = d;
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


template <class M>
NQ_D void classical_tpt_record_save(const KParams& p, Emitter& em, int NB, const double* s_r, const double* s_v,
                                    int64_t stride, int64_t idx, double mass) {
    const uint32_t obs = p.observables;
    double rsum = 0.0, vsum = 0.0, mv2 = 0.0, spr = 0.0, pot = 0.0;
    double rprev = s_r[(NB - 1) * stride + idx];
    for (int b = 0; b < NB; ++b) {
        const double rb = s_r[b * stride + idx], vb = s_v[b * stride + idx];
        rsum += rb; vsum += vb;
        mv2 = fma(mass * vb, vb, mv2);
        pot += M::potential_dof(p.params, rb);
        const double d = rprev - rb;
        spr = fma(mass * d, d, spr);
        rprev = rb;
    }
    if (NB == 1) spr = 0.0;
    const double kin = 0.5 * mv2;
    if (obs & (1u << NQCB200_OBS_KINETIC)) em.emit(NQCB200_OBS_KINETIC, 0, kin);
    if (obs & (1u << NQCB200_OBS_POTENTIAL)) em.emit(NQCB200_OBS_POTENTIAL, 0, pot);
    if (obs & (1u << NQCB200_OBS_TOTAL_ENERGY)) em.emit(NQCB200_OBS_TOTAL_ENERGY, 0, kin + pot + 0.5 * p.omega_n * p.omega_n * spr);
    if (obs & (1u << NQCB200_OBS_POSITION)) em.emit(NQCB200_OBS_POSITION, 0, rsum / NB);
    if (obs & (1u << NQCB200_OBS_VELOCITY)) em.emit(NQCB200_OBS_VELOCITY, 0, vsum / NB);
}

// DynamicsVariables at t0 for any nbeads, one thread per trajectory: same steps as ring_init_kernel (kernel_ring.cuh).
template <class M, int METHOD>
__global__ void __launch_bounds__(kRtThreads) ring_tpt_init_kernel(const __grid_constant__ KParams p, int basis,
                                                                   int sample_state, const double* state_draw) {
    constexpr int N = M::NS;
    __shared__ double red[2 * (kRtThreads / 32)];
    const int NB = p.B;
    int64_t traj = (int64_t)blockIdx.x * kRtThreads + threadIdx.x;
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const int64_t T = p.ntraj;
    const double mass = p.masses[0];
    Herm<N> s;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = j; k < N; ++k) {
            s.x[sidx(N, j, k)] = p.sig_re[(int64_t)(j + N * k) * T + traj];
            if (k > j) s.y[aidx(N, j, k)] = p.sig_im[(int64_t)(j + N * k) * T + traj];
        }
    int st = p.state ? p.state[traj] : 0;
    double Zc[N][N];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = 0; k < N; ++k) Zc[j][k] = p.Zprev[((int64_t)NB * N * N + j + N * k) * T + traj];
    double rsum = 0.0;
    for (int b = 0; b < NB; ++b) rsum += p.r[(int64_t)b * T + traj];
    Eig<N> ec;
    double Ac[sym_size(N)];
    eval_point<M>(p, rsum / NB, Zc, ec, Ac);
    if (basis == 1) {   // centroid transformation (density_matrix_dynamics.jl:83-87)
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
                        sx += ec.Z[a][i] * s.X(a, b) * ec.Z[b][j];
                        sy += ec.Z[a][i] * s.Y(a, b) * ec.Z[b][j];
                    }
                o.x[sidx(N, i, j)] = sx;
                if (j > i) o.y[aidx(N, i, j)] = sy;
            }
        s = o;
    }
    if (METHOD == NQCB200_METHOD_FSSH && sample_state) {
        const double xi = state_draw ? state_draw[traj] : philox_uniform(p.seed, (uint64_t)(p.traj_offset + traj), 0ull, 1u);
        double tot = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) tot += s.x[sidx(N, i, i)];
        const double target = xi * tot;
        double cw = s.x[sidx(N, 0, 0)];
        st = 0;
#pragma unroll
        for (int i = 1; i < N; ++i) {
            if (cw < target && st == i - 1) { st = i; cw += s.x[sidx(N, i, i)]; }
        }
    }
    double wsum[N];
#pragma unroll
    for (int i = 0; i < N; ++i) wsum[i] = 0.0;
    for (int b = 0; b < NB; ++b) {
        double Zb[N][N], Ab[sym_size(N)];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = 0; k < N; ++k) Zb[j][k] = p.Zprev[((int64_t)b * N * N + j + N * k) * T + traj];
        Eig<N> eb;
        eval_point<M>(p, p.r[(int64_t)b * T + traj], Zb, eb, Ab);
        const double acc = force_from_adiab<N, METHOD>(Ab, st, s) / mass;
#pragma unroll
        for (int i = 0; i < N; ++i) wsum[i] += eb.w[i];
        if (valid) {
            p.acc[(int64_t)b * T + traj] = acc;
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int k = 0; k < N; ++k) p.Zprev[((int64_t)b * N * N + j + N * k) * T + traj] = Zb[j][k];
        }
    }
    {
        double pot = 0.0;
        if (METHOD == NQCB200_METHOD_EHRENFEST) {
#pragma unroll
            for (int i = 0; i < N; ++i) pot += s.x[sidx(N, i, i)] * wsum[i];
        } else pot = select<N>(wsum, st);
        Emitter em{p, traj, valid, 0, red, 0};
        rt_record_save<N, METHOD>(p, em, NB, p.r, p.v, T, traj, s, st, ec, pot, mass);
    }
    if (valid) {
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = 0; k < N; ++k) {
                p.sig_re[(int64_t)(j + N * k) * T + traj] = s.X(j, k);
                p.sig_im[(int64_t)(j + N * k) * T + traj] = s.Y(j, k);
                p.Zprev[((int64_t)NB * N * N + j + N * k) * T + traj] = Zc[j][k];
            }
        if (p.state) p.state[traj] = st;
#pragma unroll
        for (int i = 0; i < N + N * N; ++i) p.ecur[(int64_t)i * T + traj] = 0.0;   // Q1
    }
}

// Classical RPMD (BCB, bcb.jl:81-116) for any nbeads: one thread per trajectory, dense normal-mode product.
// LANGEVIN: BCOCB (bcocb.jl:95-120) -- the Cayley table then holds the HALF step and the O-step of the PILE thermostat
// acts on the normal-mode velocities between the two halves.
template <class M, bool LANGEVIN = false>
__global__ void __launch_bounds__(kRtThreads) classical_tpt_step_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) double rt_sm[];
    __shared__ double red[2 * (kRtThreads / 32)];
    const int NB = p.B, tid = threadIdx.x;
    double* s_r = rt_sm;
    double* s_v = s_r + NB * kRtThreads;
    double* s_a = s_v + NB * kRtThreads;
    double* s_t = s_a + NB * kRtThreads;
    double* U = s_t + 2 * NB * kRtThreads;
    double* cay = U + NB * NB;
    int64_t traj = (int64_t)blockIdx.x * kRtThreads + tid;
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const int64_t T = p.ntraj;
    for (int i = tid; i < NB * NB; i += kRtThreads) U[i] = p.nm_to[i];
    for (int i = tid; i < 4 * NB; i += kRtThreads) cay[i] = p.cayley[i];
    for (int b = 0; b < NB; ++b) {
        s_r[b * kRtThreads + tid] = p.r[(int64_t)b * T + traj];
        s_v[b * kRtThreads + tid] = p.v[(int64_t)b * T + traj];
        s_a[b * kRtThreads + tid] = p.acc[(int64_t)b * T + traj];
    }
    __syncthreads();
    const double mass = p.masses[0], dt = p.dt, hdt = 0.5 * p.dt;
#pragma unroll 1
    for (int is = 0; is < p.nsteps; ++is) {
        const int64_t step = p.step0 + is;
        for (int b = 0; b < NB; ++b) s_v[b * kRtThreads + tid] = fma(hdt, s_a[b * kRtThreads + tid], s_v[b * kRtThreads + tid]);
        for (int k = 0; k < NB; ++k) {
            double a = 0.0, c = 0.0;
            for (int j = 0; j < NB; ++j) {
                const double u = U[j * NB + k];
                a = fma(u, s_r[j * kRtThreads + tid], a);
                c = fma(u, s_v[j * kRtThreads + tid], c);
            }
            double rn = cay[4 * k + 0] * a + cay[4 * k + 1] * c;
            double vn = cay[4 * k + 2] * a + cay[4 * k + 3] * c;
            if (LANGEVIN) {
                const double wk = 2.0 * p.omega_n * sin(k * 3.14159265358979323846 / NB);
                const double gam = (k == 0) ? p.langevin_gamma : 2.0 * wk;
                const double c1 = exp(-gam * dt), c2 = sqrt(1.0 - c1 * c1);
                double xi;
                if (p.rng == NQCB200_RNG_INJECTED) xi = p.noise[((step - p.noise_step0) * T + traj) * NB + k];
                else {
                    double z0, z1;
                    philox_normal2(p.seed, (uint64_t)(p.traj_offset + traj), (uint64_t)step * ((NB + 1) / 2) + k / 2, z0, z1, 3u);
                    xi = (k & 1) ? z1 : z0;
                }
                vn = c1 * vn + c2 * sqrt(p.omega_n / mass) * xi;                     // step_O!  steps.jl:109-124
                const double r2 = cay[4 * k + 0] * rn + cay[4 * k + 1] * vn;         // second half Cayley step
                vn = cay[4 * k + 2] * rn + cay[4 * k + 3] * vn;
                rn = r2;
            }
            s_t[k * kRtThreads + tid] = rn;
            s_t[(NB + k) * kRtThreads + tid] = vn;
        }
        for (int j = 0; j < NB; ++j) {
            double a = 0.0, c = 0.0;
            for (int k = 0; k < NB; ++k) {
                const double u = U[j * NB + k];
                a = fma(u, s_t[k * kRtThreads + tid], a);
                c = fma(u, s_t[(NB + k) * kRtThreads + tid], c);
            }
            const double acc = -M::gradient_dof(p.params, a) / mass;     // classical.jl:63-67
            s_r[j * kRtThreads + tid] = a;
            s_a[j * kRtThreads + tid] = acc;
            s_v[j * kRtThreads + tid] = fma(hdt, acc, c);
        }
        if ((step + 1) % p.save_every == 0) {
            const int64_t isave = (step + 1) / p.save_every;
            if (isave < p.nsave) {
                Emitter em{p, traj, valid, (int)isave, red, 0};
                classical_tpt_record_save<M>(p, em, NB, s_r, s_v, kRtThreads, tid, mass);
            }
        }
    }
    if (valid) {
        for (int b = 0; b < NB; ++b) {
            p.r[(int64_t)b * T + traj] = s_r[b * kRtThreads + tid];
            p.v[(int64_t)b * T + traj] = s_v[b * kRtThreads + tid];
            p.acc[(int64_t)b * T + traj] = s_a[b * kRtThreads + tid];
        }
    }
}

// Classical RPMD, power-of-two bead count, whole ring polymer in REGISTERS: one thread per trajectory, in-register FFT
// for the free ring-polymer step, no shuffles and no shared-memory traffic in the step loop (the beads-on-lanes kernel
// spends 69 % of the LSU wavefront budget on its 22 double shuffles per step).  The acceleration is not carried: a
// single-surface model's force is a few flops, so it is re-evaluated at the start of the step.
template <class M, int NB>
__global__ void __launch_bounds__(kRtThreads, 2) classical_tpt_fft_kernel(const __grid_constant__ KParams p) {
    static_assert(NB >= 2 && (NB & (NB - 1)) == 0, "power-of-two bead count");
    __shared__ double red[2 * (kRtThreads / 32)];
    __shared__ double s_tab[5 * NB];
    const int tid = threadIdx.x;
    int64_t traj = (int64_t)blockIdx.x * kRtThreads + tid;
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const int64_t T = p.ntraj;
    for (int j = tid; j < NB; j += kRtThreads) {
        if (j < NB / 2) {
            double si, co;
            sincospi(-2.0 * (double)j / (double)NB, &si, &co);
            s_tab[j] = co; s_tab[NB / 2 + j] = si;
        }
        const double a = p.cayley[4 * j + 0], b = p.cayley[4 * j + 1], c = p.cayley[4 * j + 2], d = p.cayley[4 * j + 3];
        const double inv = 0.5 / NB;
        s_tab[NB + 2 * j] = (a + d) * inv; s_tab[NB + 2 * j + 1] = (c - b) * inv;
        s_tab[3 * NB + 2 * j] = (a - d) * inv; s_tab[3 * NB + 2 * j + 1] = (c + b) * inv;
    }
    __syncthreads();
    RtTables<NB> tb{s_tab, s_tab + NB / 2, s_tab + NB, s_tab + 3 * NB};
    double r[NB], v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) { r[b] = p.r[(int64_t)b * T + traj]; v[b] = p.v[(int64_t)b * T + traj]; }
    const double mass = p.masses[0], hdt = 0.5 * p.dt, hm = hdt / mass;
#pragma unroll 1
    for (int is = 0; is < p.nsteps; ++is) {
        const int64_t step = p.step0 + is;
#pragma unroll
        for (int b = 0; b < NB; ++b) v[b] = fma(-hm, M::gradient_dof(p.params, r[b]), v[b]);     // B
        rt_free_step<NB>(tb, r, v);                                                               // C
#pragma unroll
        for (int b = 0; b < NB; ++b) v[b] = fma(-hm, M::gradient_dof(p.params, r[b]), v[b]);     // B (classical.jl:63-67)
        if ((step + 1) % p.save_every == 0) {
            const int64_t isave = (step + 1) / p.save_every;
            if (isave < p.nsave) {
                const uint32_t obs = p.observables;
                double rsum = 0.0, vsum = 0.0, mv2 = 0.0, spr = 0.0, pot = 0.0;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    rsum += r[b]; vsum += v[b];
                    mv2 = fma(mass * v[b], v[b], mv2);
                    pot += M::potential_dof(p.params, r[b]);
                    const double d = r[(b + NB - 1) % NB] - r[b];
                    spr = fma(mass * d, d, spr);
                }
                Emitter em{p, traj, valid, (int)isave, red, 0};
                const double kin = 0.5 * mv2;
                if (obs & (1u << NQCB200_OBS_KINETIC)) em.emit(NQCB200_OBS_KINETIC, 0, kin);
                if (obs & (1u << NQCB200_OBS_POTENTIAL)) em.emit(NQCB200_OBS_POTENTIAL, 0, pot);
                if (obs & (1u << NQCB200_OBS_TOTAL_ENERGY)) em.emit(NQCB200_OBS_TOTAL_ENERGY, 0, kin + pot + 0.5 * p.omega_n * p.omega_n * spr);
                if (obs & (1u << NQCB200_OBS_POSITION)) em.emit(NQCB200_OBS_POSITION, 0, rsum / NB);
                if (obs & (1u << NQCB200_OBS_VELOCITY)) em.emit(NQCB200_OBS_VELOCITY, 0, vsum / NB);
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            p.r[(int64_t)b * T + traj] = r[b];
            p.v[(int64_t)b * T + traj] = v[b];
            p.acc[(int64_t)b * T + traj] = -M::gradient_dof(p.params, r[b]) / mass;
        }
    }
}

// RingPolymerSimulation{ThermalLangevin} + BCOCB (bcocb.jl:95-120): B, to normal modes, C(1/2), O, C(1/2), back, force, B,
// on the register-resident ring polymer of classical_tpt_fft_kernel.  In the spectrum Z_k of z = r + i v (see
// FreeRingPolymer) the deterministic part C(1/2) diag(1, c1_k) C(1/2) is again one real 2x2 matrix per frequency, i.e.
// another (alpha_k, beta_k) pair, and the O-step noise n_k = c2_k sigma xi_k of the ORTHONORMAL real normal modes enters
// as Z'_k += (b_k + i d_k) N_k with [[a,b],[c,d]] the half Cayley step and
//     N_0 = sqrt(NB) n_0,  N_{NB/2} = sqrt(NB) n_{NB/2},  N_k = sqrt(NB/2) (n_k + i n_{NB-k}) = conj(N_{NB-k})  (0 < k < NB/2)
// (R_k = sqrt(NB/2) (y_k + i y_{NB-k}) relates the DFT to the orthonormal cos / sin modes of RingPolymerArrays).
// PILE friction (FrictionCache, bcocb.jl:78-87): gamma_0 = gamma, gamma_k = 2 omega_k, c1 = exp(-gamma_k dt),
// c2 = sqrt(1 - c1^2), sigma = sqrt(NB kT / m) (steps.jl:109-124).
template <class M, int NB>
__global__ void __launch_bounds__(kRtThreads, 2) langevin_tpt_fft_kernel(const __grid_constant__ KParams p) {
    static_assert(NB >= 2 && (NB & (NB - 1)) == 0, "power-of-two bead count");
    __shared__ double red[2 * (kRtThreads / 32)];
    __shared__ double s_tab[7 * NB];      // twr[NB/2] twi[NB/2] al[2NB] be[2NB] bn[NB] dn[NB]
    const int tid = threadIdx.x;
    int64_t traj = (int64_t)blockIdx.x * kRtThreads + tid;
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const int64_t T = p.ntraj;
    const double mass = p.masses[0], hdt = 0.5 * p.dt, hm = hdt / mass;
    const double sigma = sqrt(p.omega_n / mass);          // omega_n = NB kT: ring-polymer temperature
    const double pi = 3.14159265358979323846;
    for (int j = tid; j < NB; j += kRtThreads) {
        if (j < NB / 2) {
            double si, co;
            sincospi(-2.0 * (double)j / (double)NB, &si, &co);
            s_tab[j] = co; s_tab[NB / 2 + j] = si;
        }
        const double a = p.cayley[4 * j + 0], b = p.cayley[4 * j + 1], c = p.cayley[4 * j + 2], d = p.cayley[4 * j + 3];   // half step
        const double wk = 2.0 * p.omega_n * sin(j * pi / NB);
        const double gam = (j == 0) ? p.langevin_gamma : 2.0 * wk;
        const double c1 = exp(-gam * p.dt), c2 = sqrt(1.0 - c1 * c1);
        // S diag(1, c1) S
        const double ma = a * a + b * c1 * c, mb = a * b + b * c1 * d, mc = c * a + d * c1 * c, md = c * b + d * c1 * d;
        const double inv = 0.5 / NB;
        s_tab[NB + 2 * j] = (ma + md) * inv; s_tab[NB + 2 * j + 1] = (mc - mb) * inv;
        s_tab[3 * NB + 2 * j] = (ma - md) * inv; s_tab[3 * NB + 2 * j + 1] = (mc + mb) * inv;
        const double scale = ((j == 0 || 2 * j == NB) ? sqrt((double)NB) : sqrt(0.5 * NB)) * c2 * sigma / NB;
        s_tab[5 * NB + j] = b * scale; s_tab[6 * NB + j] = d * scale;
    }
    __syncthreads();
    RtTables<NB> tb{s_tab, s_tab + NB / 2, s_tab + NB, s_tab + 3 * NB};
    const double* bn = s_tab + 5 * NB;
    const double* dn = s_tab + 6 * NB;
    double r[NB], v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) { r[b] = p.r[(int64_t)b * T + traj]; v[b] = p.v[(int64_t)b * T + traj]; }
#pragma unroll 1
    for (int is = 0; is < p.nsteps; ++is) {
        const int64_t step = p.step0 + is;
        double xi[NB];      // one standard normal per normal mode (mode index k)
        if (p.rng == NQCB200_RNG_INJECTED) {
#pragma unroll
            for (int k = 0; k < NB; ++k) xi[k] = p.noise[((step - p.noise_step0) * T + traj) * NB + k];
        } else {
#pragma unroll
            for (int k = 0; k < NB; k += 2)
                philox_normal2(p.seed, (uint64_t)(p.traj_offset + traj), (uint64_t)step * (NB / 2) + k / 2, xi[k], xi[k + 1], 3u);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) v[b] = fma(-hm, M::gradient_dof(p.params, r[b]), v[b]);     // B
        rt_free_step<NB, true>(tb, r, v, xi, bn, dn);                                             // C O C
#pragma unroll
        for (int b = 0; b < NB; ++b) v[b] = fma(-hm, M::gradient_dof(p.params, r[b]), v[b]);     // B
        if ((step + 1) % p.save_every == 0) {
            const int64_t isave = (step + 1) / p.save_every;
            if (isave < p.nsave) {
                const uint32_t obs = p.observables;
                double rsum = 0.0, vsum = 0.0, mv2 = 0.0, spr = 0.0, pot = 0.0;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    rsum += r[b]; vsum += v[b];
                    mv2 = fma(mass * v[b], v[b], mv2);
                    pot += M::potential_dof(p.params, r[b]);
                    const double d = r[(b + NB - 1) % NB] - r[b];
                    spr = fma(mass * d, d, spr);
                }
                Emitter em{p, traj, valid, (int)isave, red, 0};
                const double kin = 0.5 * mv2;
                if (obs & (1u << NQCB200_OBS_KINETIC)) em.emit(NQCB200_OBS_KINETIC, 0, kin);
                if (obs & (1u << NQCB200_OBS_POTENTIAL)) em.emit(NQCB200_OBS_POTENTIAL, 0, pot);
                if (obs & (1u << NQCB200_OBS_TOTAL_ENERGY)) em.emit(NQCB200_OBS_TOTAL_ENERGY, 0, kin + pot + 0.5 * p.omega_n * p.omega_n * spr);
                if (obs & (1u << NQCB200_OBS_POSITION)) em.emit(NQCB200_OBS_POSITION, 0, rsum / NB);
                if (obs & (1u << NQCB200_OBS_VELOCITY)) em.emit(NQCB200_OBS_VELOCITY, 0, vsum / NB);
            }
        }
    }
    if (valid) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            p.r[(int64_t)b * T + traj] = r[b];
            p.v[(int64_t)b * T + traj] = v[b];
            p.acc[(int64_t)b * T + traj] = -M::gradient_dof(p.params, r[b]) / mass;
        }
    }
}

template <class M>
__global__ void __launch_bounds__(kRtThreads) classical_tpt_init_kernel(const __grid_constant__ KParams p, int, int,
                                                                        const double*) {
    __shared__ double red[2 * (kRtThreads / 32)];
    const int NB = p.B;
    int64_t traj = (int64_t)blockIdx.x * kRtThreads + threadIdx.x;
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const int64_t T = p.ntraj;
    const double mass = p.masses[0];
    if (valid)
        for (int b = 0; b < NB; ++b) p.acc[(int64_t)b * T + traj] = -M::gradient_dof(p.params, p.r[(int64_t)b * T + traj]) / mass;
    Emitter em{p, traj, valid, 0, red, 0};
    classical_tpt_record_save<M>(p, em, NB, p.r, p.v, T, traj, mass);
}

#endif  // __CUDACC__

}  // namespace nq

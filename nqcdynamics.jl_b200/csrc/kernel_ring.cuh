// kernel_ring.cuh -- ring-polymer kernels: beads on lanes.
//
// NB beads of one trajectory occupy NB adjacent lanes of a warp (NB | 32, so 32/NB trajectories per
// warp); every lane owns one bead's (r, v, acceleration, eigenvector gauge) in registers for the
// whole launch.  The free ring-polymer step (to-normal-modes -> Cayley 2x2 per mode -> back) is ONE
// complex FFT of r + i v over the lanes (FreeRingPolymer below).
//
// Role after round 1: these kernels initialise every power-of-two ring polymer (ring_init_kernel,
// classical_ring_init_kernel), step classical MD (NB == 1) and NRPMD (kernel_nrpmd.cuh), and remain the A/B
// fallback of the thread-per-trajectory step kernels of kernel_ring_tpt.cuh (RPSH / RP-Ehrenfest / RPMD), which
// are 3-6x faster because they do not replicate the centroid electronic work / do not shuffle.
//
// Reference restated:
//   BCB.perform_step!          src/DynamicsMethods/IntegrationAlgorithms/bcb.jl:81-116   (RPMD)
//   BCBwithTsit5.perform_step! .../bcb_electronics.jl:53-97                              (RPSH, RP-Ehrenfest)
//   step_C!                    .../steps.jl:10-17 ; cayley_propagator  src/RingPolymers/ring_polymer.jl:71-82
//   NormalModeTransformation   RingPolymerArrays (external; docs/src/api/RingPolymerArrays/ringpolymerarrays.md:93-133)
//   centroid hopping quantities  SurfaceHoppingMethods.jl:85-103 ; RP rescale rpsh.jl:30-50
//   acceleration! fssh.jl:67-74 (3-index), ehrenfest_rpmd.jl:23-43, classical.jl:63-67
//   energies rpsh.jl:52-64, ehrenfest_rpmd.jl:45-51, DynamicsUtils.jl:108-151, ring_polymer.jl:89-107
// Ring-polymer kernels are instantiated for one nuclear dof per bead (ndofs*natoms == 1), which is
// what every ring-polymer config in BASELINE.json uses.
#pragma once
#include "kernel_density.cuh"

namespace nq {

#if defined(__CUDACC__)

// Free ring-polymer step (to-normal-modes -> Cayley 2x2 per mode -> back, steps.jl:10-17) as ONE
// complex FFT over the lanes of a bead group instead of four dense NB x NB mat-vecs.
//
// With z_j = r_j + i v_j and Z_k = sum_j z_j e^{-2 pi i jk/NB}: the transforms of the two real
// sequences are R_k = (Z_k + W_k)/2, V_k = (Z_k - W_k)/(2i), W_k = conj(Z_{NB-k}); the real normal
// modes k and NB-k (cos / sin pair, same frequency, same Cayley matrix [[a,b],[c,d]]) are the real and
// imaginary parts of R_k, so the propagated spectrum is
//     Z'_k = R'_k + i V'_k = alpha_k Z_k + beta_k W_k,
//     alpha = ((a+d) + i(c-b))/2,  beta = ((a-d) + i(c+b))/2,
// and r', v' are the real / imaginary parts of the inverse FFT (1/NB folded into alpha, beta; the
// orthogonal-U normalisation cancels).  Forward = radix-2 DIF (natural in, bit-reversed out), inverse =
// radix-2 DIT (bit-reversed in, natural out), so no reordering: 2 log2(NB) + 1 shuffle rounds of one
// complex number instead of 4 NB shuffles + 4 NB shared-memory loads (the dense version was 99.8 %
// LSU-bound, profiles/r01).  Agrees with the dense U'..U product to rounding (1e-14).
template <int NB>
struct FreeRingPolymer {
    static constexpr int LOG = (NB >= 32) ? 5 : (NB >= 16) ? 4 : (NB >= 8) ? 3 : (NB >= 4) ? 2 : (NB >= 2) ? 1 : 0;
    static constexpr int NS = LOG > 0 ? LOG : 1;
    double twr[NS], twi[NS];   // stage twiddle W_{2h}^{lane mod h} on the upper lanes, 1 on the lower ones
    double sg[NS];             // -1 on the upper lanes of a stage, +1 on the lower ones
    double ar, ai, br, bi;     // alpha, beta of the mode this lane holds after the forward pass
    int partner;               // warp lane holding mode NB - k

    NQ_D void init(const KParams& p, int lane, int group_base) {
        int k = 0;
#pragma unroll
        for (int b = 0; b < LOG; ++b) k |= ((lane >> b) & 1) << (LOG - 1 - b);
        const int kc = (NB - k) % NB;
        int lp = 0;
#pragma unroll
        for (int b = 0; b < LOG; ++b) lp |= ((kc >> b) & 1) << (LOG - 1 - b);
        partner = group_base + lp;
#pragma unroll
        for (int s = 0; s < LOG; ++s) {
            const int h = NB >> (s + 1);
            const bool upper = (lane & h) != 0;
            double si = 0.0, co = 1.0;
            if (upper) sincospi(-(double)(lane & (h - 1)) / (double)h, &si, &co);
            twr[s] = co; twi[s] = si; sg[s] = upper ? -1.0 : 1.0;
        }
        const double a = p.cayley[4 * k + 0], b = p.cayley[4 * k + 1], c = p.cayley[4 * k + 2], d = p.cayley[4 * k + 3];
        if (NB == 1) { ar = a; ai = b; br = c; bi = d; return; }
        const double inv = 0.5 / NB;
        ar = (a + d) * inv; ai = (c - b) * inv; br = (a - d) * inv; bi = (c + b) * inv;
    }

    NQ_D void step(double& r, double& v) const {
        if (NB == 1) {
            const double rt = ar * r + ai * v, vt = br * r + bi * v;
            r = rt; v = vt;
            return;
        }
        double zr = r, zi = v;
#pragma unroll
        for (int s = 0; s < LOG; ++s) {
            const int h = NB >> (s + 1);
            const double yr = __shfl_xor_sync(0xffffffffu, zr, h), yi = __shfl_xor_sync(0xffffffffu, zi, h);
            const double tr = fma(sg[s], zr, yr), ti = fma(sg[s], zi, yi);   // lower: z + y, upper: y - z
            if (h > 1) { zr = fma(tr, twr[s], -ti * twi[s]); zi = fma(tr, twi[s], ti * twr[s]); }
            else { zr = tr; zi = ti; }
        }
        const double wr = __shfl_sync(0xffffffffu, zr, partner), wi = -__shfl_sync(0xffffffffu, zi, partner);
        double nr = fma(ar, zr, fma(-ai, zi, fma(br, wr, -bi * wi)));
        double ni = fma(ar, zi, fma(ai, zr, fma(br, wi, bi * wr)));
#pragma unroll
        for (int s = LOG - 1; s >= 0; --s) {
            const int h = NB >> (s + 1);
            double xr = nr, xi = ni;
            if (h > 1) { xr = fma(nr, twr[s], ni * twi[s]); xi = fma(ni, twr[s], -nr * twi[s]); }   // conj(twiddle)
            const double yr = __shfl_xor_sync(0xffffffffu, xr, h), yi = __shfl_xor_sync(0xffffffffu, xi, h);
            nr = fma(sg[s], xr, yr); ni = fma(sg[s], xi, yi);                   // lower: x + y, upper: y - x
        }
        r = nr; v = ni;
    }
};

template <int NB>
NQ_D double spring_energy(double r, double mass, double omega_n, int lane, int group_base) {
    if (NB == 1) return 0.0;
    const double rnext = __shfl_sync(0xffffffffu, r, group_base + ((lane + 1) % NB));
    const double d = r - rnext;
    return 0.5 * omega_n * omega_n * lane_sum<NB>(mass * d * d);
}

// ---------------------------------------------------------------------------------------------
// RPSH / RP-Ehrenfest
// ---------------------------------------------------------------------------------------------
template <int N, int NB, int METHOD>
NQ_D void ring_record_save(const KParams& p, Emitter& em, int lane, int group_base, const Herm<N>& s, int st,
                           const Eig<N>& ec, const Eig<N>& eb, double r, double v, double mass) {
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
    if (obs & ((1u << NQCB200_OBS_KINETIC) | (1u << NQCB200_OBS_POTENTIAL) | (1u << NQCB200_OBS_TOTAL_ENERGY))) {
        const double kin = 0.5 * lane_sum<NB>(mass * v * v);
        double pb = 0.0;   // this bead's potential
        if (METHOD == NQCB200_METHOD_FSSH) pb = select<N>(eb.w, st);
        else {
#pragma unroll
            for (int i = 0; i < N; ++i) pb += s.x[sidx(N, i, i)] * eb.w[i];
        }
        const double pot = lane_sum<NB>(pb);
        if (obs & (1u << NQCB200_OBS_KINETIC)) em.emit(NQCB200_OBS_KINETIC, 0, kin);
        if (obs & (1u << NQCB200_OBS_POTENTIAL)) em.emit(NQCB200_OBS_POTENTIAL, 0, pot);
        if (obs & (1u << NQCB200_OBS_TOTAL_ENERGY)) {
            const double spr = spring_energy<NB>(r, mass, p.omega_n, lane, group_base);
            em.emit(NQCB200_OBS_TOTAL_ENERGY, 0, kin + pot + spr);
        }
    }
    const double rc = lane_sum<NB>(r) / NB, vc = lane_sum<NB>(v) / NB;
    if (obs & (1u << NQCB200_OBS_POSITION)) em.emit(NQCB200_OBS_POSITION, 0, rc);
    if (obs & (1u << NQCB200_OBS_VELOCITY)) em.emit(NQCB200_OBS_VELOCITY, 0, vc);
    if (obs & (1u << NQCB200_OBS_DISCRETE_STATE)) em.emit(NQCB200_OBS_DISCRETE_STATE, 0, (double)(st + 1));
    const bool last = (em.isave == p.nsave - 1);
    if (obs & ((1u << NQCB200_OBS_SCATTERING) | (1u << NQCB200_OBS_SCATTERING_DIABATIC))) {
        // get_positions(final)[1]: first dof of the first bead (DynamicsOutputs.jl:332)
        const double r0 = __shfl_sync(0xffffffffu, r, group_base);
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

template <int N, int NB>
struct RingRegs {
    double r, v, acc, mass;
    double Zb[N][N];   // this bead's eigenvector gauge
    double Zc[N][N];   // centroid gauge (replicated)
    Herm<N> s;
    int st;
    ElecParams<N> cur;
};

template <int N, int NB>
NQ_D void ring_load(const KParams& p, int64_t traj, int lane, RingRegs<N, NB>& R, bool with_dynamics) {
    const int64_t T = p.ntraj;
    R.r = p.r[(int64_t)lane * T + traj];
    R.v = p.v[(int64_t)lane * T + traj];
    R.acc = with_dynamics ? p.acc[(int64_t)lane * T + traj] : 0.0;
    R.mass = p.masses[0];
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
        for (int k = 0; k < N; ++k) {
            R.Zb[j][k] = p.Zprev[((int64_t)lane * N * N + j + N * k) * T + traj];
            R.Zc[j][k] = p.Zprev[((int64_t)NB * N * N + j + N * k) * T + traj];
        }
    if (with_dynamics) {
#pragma unroll
        for (int i = 0; i < N; ++i) R.cur.E[i] = p.ecur[(int64_t)i * T + traj];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = j + 1; k < N; ++k) R.cur.g[aidx(N, j, k)] = p.ecur[(int64_t)(N + j + N * k) * T + traj];
    }
}

template <int N, int NB>
NQ_D void ring_store(const KParams& p, int64_t traj, int lane, const RingRegs<N, NB>& R) {
    const int64_t T = p.ntraj;
    p.r[(int64_t)lane * T + traj] = R.r;
    p.v[(int64_t)lane * T + traj] = R.v;
    p.acc[(int64_t)lane * T + traj] = R.acc;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = 0; k < N; ++k) p.Zprev[((int64_t)lane * N * N + j + N * k) * T + traj] = R.Zb[j][k];
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = 0; k < N; ++k) {
                p.sig_re[(int64_t)(j + N * k) * T + traj] = R.s.X(j, k);
                p.sig_im[(int64_t)(j + N * k) * T + traj] = R.s.Y(j, k);
                p.Zprev[((int64_t)NB * N * N + j + N * k) * T + traj] = R.Zc[j][k];
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

// eigen + adiabatic derivative at one position (bead or centroid)
template <class M>
NQ_D void eval_point(const KParams& p, double q, double (&Zref)[M::NS][M::NS], Eig<M::NS>& e, double (&Ap)[sym_size(M::NS)]) {
    constexpr int N = M::NS;
    double Vp[sym_size(N)], dVp[sym_size(N)];
    model_value_and_derivative<M>(p.params, q, Vp, dVp);
    sym_eigh<N>(Vp, e);
    fix_gauge<N>(e, Zref);
    similarity<N>(dVp, e.Z, Ap);
}

template <class M, int NB, int METHOD>
__global__ void __launch_bounds__(kBlockThreads) ring_step_kernel(const __grid_constant__ KParams p) {
    constexpr int N = M::NS;
    __shared__ double smem[2 * (kBlockThreads / 32)];
    const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t traj = gthread / NB;
    const int lane = (int)(gthread % NB);
    const int group_base = (threadIdx.x & 31) & ~(NB - 1);
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const bool lane0 = (lane == 0);
    const int64_t T = p.ntraj;

    RingRegs<N, NB> R;
    ring_load<N, NB>(p, traj, lane, R, true);
    FreeRingPolymer<NB> frp;
    frp.init(p, lane, group_base);
    Eig<N> eb, ec;
    double Ab[sym_size(N)], Ac[sym_size(N)];
    unsigned long long nhops = 0, nfrus = 0;
    const double dt = p.dt, hdt = 0.5 * p.dt;

#pragma unroll 1
    for (int is = 0; is < p.nsteps; ++is) {
        const int64_t step = p.step0 + is;
        const double t = p.t0 + dt * (double)step;
        const double tcur = (step == 0) ? 0.0 : t;   // Q1
        double vt = fma(hdt, R.acc, R.v);
        double rt = R.r;
        frp.step(rt, vt);
        // update_cache!: every bead and the centroid (bcb_electronics.jl:73)
        eval_point<M>(p, rt, R.Zb, eb, Ab);
        const double rcent = lane_sum<NB>(rt) / NB;
        eval_point<M>(p, rcent, R.Zc, ec, Ac);
        R.acc = force_from_adiab<N, METHOD>(Ab, R.st, R.s) / R.mass;
        R.v = fma(hdt, R.acc, vt);
        R.r = rt;
        const double vcent = lane_sum<NB>(R.v) / NB;
        ElecParams<N> nxt;
#pragma unroll
        for (int i = 0; i < N; ++i) nxt.E[i] = ec.w[i];
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = j + 1; k < N; ++k) nxt.g[aidx(N, j, k)] = (-Ac[sidx(N, j, k)] / (ec.w[j] - ec.w[k])) * vcent;
        propagate_density<N>(R.cur, tcur, nxt, t + dt, t, dt, R.s, p.tsit5_ha);

        if (METHOD == NQCB200_METHOD_FSSH) {
            const double xi = (p.rng == NQCB200_RNG_INJECTED)
                                  ? p.draws[(step - p.draws_step0) * T + traj]
                                  : philox_uniform(p.seed, (uint64_t)(p.traj_offset + traj), (uint64_t)step, 0u);
            const int s0 = R.st;
            const double inv_ss = 1.0 / R.s.Xsel(s0, s0);
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
                bool accept = true;
                if (p.rescaling != NQCB200_RESCALE_OFF) {
                    const double wn = select<N>(ec.w, new_state), wo = select<N>(ec.w, s0);
                    double ano = 0.0;
#pragma unroll
                    for (int j = 0; j < N; ++j)
#pragma unroll
                        for (int k = j + 1; k < N; ++k)
                            ano = ((j == new_state && k == s0) || (k == new_state && j == s0)) ? Ac[sidx(N, j, k)] : ano;
                    const double d = -ano / (wn - wo);
                    const double a = 0.5 * d * d / R.mass, b = d * vcent, c = wn - wo;
                    const double disc = b * b - 4.0 * a * c;
                    if (disc < 0.0) {
                        accept = false;
                        nfrus += (lane0 && valid);
                        if (p.rescaling == NQCB200_RESCALE_VINVERSION) {   // rpsh.jl:39-50
                            const double dn = d / fabs(d);
                            R.v -= 2.0 * (vcent * dn) * dn;
                        }
                    } else {
                        const double root = sqrt(disc);
                        const double gam = (b < 0.0) ? (b + root) / (2.0 * a) : (b - root) / (2.0 * a);
                        R.v -= gam * d / R.mass;                           // rpsh.jl:30-37: every bead
                    }
                }
                if (accept) { R.st = new_state; nhops += (lane0 && valid); }
            }
        }
        R.cur = nxt;

        if ((step + 1) % p.save_every == 0) {
            const int64_t isave = (step + 1) / p.save_every;
            if (isave < p.nsave) {
                Emitter em{p, traj, valid && lane0, (int)isave, smem, 0};
                ring_record_save<N, NB, METHOD>(p, em, lane, group_base, R.s, R.st, ec, eb, R.r, R.v, R.mass);
            }
        }
    }

    if (valid) {
        ring_store<N, NB>(p, traj, lane, R);
        if (p.diagnostics && lane0) {
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

template <class M, int NB, int METHOD>
__global__ void __launch_bounds__(kBlockThreads) ring_init_kernel(const __grid_constant__ KParams p, int basis,
                                                                  int sample_state, const double* state_draw) {
    constexpr int N = M::NS;
    __shared__ double smem[2 * (kBlockThreads / 32)];
    const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t traj = gthread / NB;
    const int lane = (int)(gthread % NB);
    const int group_base = (threadIdx.x & 31) & ~(NB - 1);
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const bool lane0 = (lane == 0);

    RingRegs<N, NB> R;
    ring_load<N, NB>(p, traj, lane, R, false);
    Eig<N> eb, ec;
    double Ab[sym_size(N)], Ac[sym_size(N)];
    eval_point<M>(p, R.r, R.Zb, eb, Ab);
    const double rcent = lane_sum<NB>(R.r) / NB;
    eval_point<M>(p, rcent, R.Zc, ec, Ac);
    if (basis == 1) {   // centroid transformation (evaluate_transformation, density_matrix_dynamics.jl:83-87)
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
                        sx += ec.Z[a][i] * R.s.X(a, b) * ec.Z[b][j];
                        sy += ec.Z[a][i] * R.s.Y(a, b) * ec.Z[b][j];
                    }
                o.x[sidx(N, i, j)] = sx;
                if (j > i) o.y[aidx(N, i, j)] = sy;
            }
        R.s = o;
    }
    if (METHOD == NQCB200_METHOD_FSSH && sample_state) {
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
    R.acc = force_from_adiab<N, METHOD>(Ab, R.st, R.s) / R.mass;
#pragma unroll
    for (int i = 0; i < N; ++i) R.cur.E[i] = 0.0;
#pragma unroll
    for (int i = 0; i < asym_size(N); ++i) R.cur.g[i] = 0.0;
    {
        Emitter em{p, traj, valid && lane0, 0, smem, 0};
        ring_record_save<N, NB, METHOD>(p, em, lane, group_base, R.s, R.st, ec, eb, R.r, R.v, R.mass);
    }
    if (valid) ring_store<N, NB>(p, traj, lane, R);
}

// ---------------------------------------------------------------------------------------------
// Classical MD (NB == 1, VelocityVerlet) and RPMD (NB > 1, BCB) on a single-surface model
// ---------------------------------------------------------------------------------------------
template <class M, int NB>
NQ_D void classical_record_save(const KParams& p, Emitter& em, int lane, int group_base, double r, double v, double mass) {
    const uint32_t obs = p.observables;
    if (obs & ((1u << NQCB200_OBS_KINETIC) | (1u << NQCB200_OBS_POTENTIAL) | (1u << NQCB200_OBS_TOTAL_ENERGY))) {
        const double kin = 0.5 * lane_sum<NB>(mass * v * v);
        const double pot = lane_sum<NB>(M::potential_dof(p.params, r));
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
}

template <class M, int NB>
__global__ void __launch_bounds__(kBlockThreads) classical_ring_step_kernel(const __grid_constant__ KParams p) {
    __shared__ double smem[2 * (kBlockThreads / 32)];
    const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t traj = gthread / NB;
    const int lane = (int)(gthread % NB);
    const int group_base = (threadIdx.x & 31) & ~(NB - 1);
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const int64_t T = p.ntraj;
    double r = p.r[(int64_t)lane * T + traj], v = p.v[(int64_t)lane * T + traj], acc = p.acc[(int64_t)lane * T + traj];
    const double mass = p.masses[0];
    FreeRingPolymer<NB> frp;
    frp.init(p, lane, group_base);
    const double dt = p.dt, hdt = 0.5 * p.dt;
#pragma unroll 1
    for (int is = 0; is < p.nsteps; ++is) {
        const int64_t step = p.step0 + is;
        if (NB == 1) {
            // OrdinaryDiffEq VelocityVerlet: u = uprev + dt*duprev + dt^2/2*ku ; du = duprev + dt/2*(ku + kdu)
            const double a_old = acc;
            r = r + dt * v + dt * dt * 0.5 * a_old;
            acc = -M::gradient_dof(p.params, r) / mass;
            v = v + dt * (0.5 * a_old + 0.5 * acc);
        } else {
            double vt = fma(hdt, acc, v);
            frp.step(r, vt);
            acc = -M::gradient_dof(p.params, r) / mass;     // classical.jl:63-67
            v = fma(hdt, acc, vt);
        }
        if ((step + 1) % p.save_every == 0) {
            const int64_t isave = (step + 1) / p.save_every;
            if (isave < p.nsave) {
                Emitter em{p, traj, valid && lane == 0, (int)isave, smem, 0};
                classical_record_save<M, NB>(p, em, lane, group_base, r, v, mass);
            }
        }
    }
    if (valid) {
        p.r[(int64_t)lane * T + traj] = r;
        p.v[(int64_t)lane * T + traj] = v;
        p.acc[(int64_t)lane * T + traj] = acc;
    }
}

template <class M, int NB>
__global__ void __launch_bounds__(kBlockThreads) classical_ring_init_kernel(const __grid_constant__ KParams p, int, int,
                                                                            const double*) {
    __shared__ double smem[2 * (kBlockThreads / 32)];
    const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t traj = gthread / NB;
    const int lane = (int)(gthread % NB);
    const int group_base = (threadIdx.x & 31) & ~(NB - 1);
    const bool valid = traj < p.ntraj;
    if (!valid) traj = p.ntraj - 1;
    const int64_t T = p.ntraj;
    const double r = p.r[(int64_t)lane * T + traj], v = p.v[(int64_t)lane * T + traj];
    const double mass = p.masses[0];
    const double acc = -M::gradient_dof(p.params, r) / mass;
    {
        Emitter em{p, traj, valid && lane == 0, 0, smem, 0};
        classical_record_save<M, NB>(p, em, lane, group_base, r, v, mass);
    }
    if (valid) p.acc[(int64_t)lane * T + traj] = acc;
}

#endif  // __CUDACC__

}  // namespace nq

// kernel_spinboson_epoch.cuh -- SpinBoson FSSH / Ehrenfest as two throughput kernels per EPOCH of E nuclear steps.
//
// Same reference path and model algebra as kernel_spinboson.cuh (BABwithTsit5, bab_electronics.jl:61-91;
// HoppingCallback, surface_hopping.jl:2-168; fssh.jl:67-121; ehrenfest.jl:50-68) -- read that header first.
//
// Why.  kernel_spinboson.cuh keeps the bath of 128 trajectories resident in shared memory and steps it every dt; the SM
// then holds 128 serial electronic chains (eigen -> Tsit5 -> hop, ~1100 mostly dependent FP64 instructions per step)
// and nothing else to hide their latency: FP64 pipe 44 % active, 62 % of the step spent in the chain, the rest in a
// shared-memory-bound sweep (profiles/r01 sb_v5).  A shared-memory-resident K-step variant of the idea below (round 2,
// profiles/r02/SUMMARY.md) confirmed the diagnosis: with one chain per trajectory the SM issues 0.3 instructions per
// cycle per scheduler whatever the sweep costs.  The bound is the NUMBER OF CHAINS in flight, and shared memory caps it.
//
// What.  Each bath mode is an oscillator driven by a drive common to all modes up to a factor: with u the half-kicked
// velocity one nuclear step of mode j is
//     vt = u - a_j r - (c_j/m_j) f1 - c_j f2 ,   r' = r + dt vt ,   u' = vt ,         a_j = dt w_j^2 / m_j ,
// f1 = dt B + gamma_m (force scalar B of fssh.jl:67-74 / ehrenfest.jl:57-65 and the velocity change of an accepted
// hop, surface_hopping.jl:139-146), f2 the reflection of a frustrated hop under :vinversion (:155-164).  The electronic
// step needs three bath sums only, L = sum c r', C = sum c vt, W = sum (c w^2/m) r', and they are LINEAR in the impulses:
//     sums(step k) = [free evolution from the epoch's first state] + sum_{0 < i <= k} f_i kappa(k - i)
// with trajectory-independent lag tables kappa (host-built at create, KParams::sb_kap).  Hence per epoch
//   1. sb_bath_kernel  (thread per trajectory, bath in HBM as [mode][trajectory]): loads a mode once, REPLAYS the E steps
//      of the previous epoch with their now known impulses (the FMA sequence of the step-by-step kernel), stores it, and
//      free-evolves it E steps further in registers accumulating the 3E sums.  No dependent chain longer than one mode,
//      four modes in flight per thread: FP64-throughput bound.  HBM traffic: the 1.6 KB of bath state per trajectory
//      once per EPOCH in each direction.
//   2. sb_elec_kernel  (thread per trajectory, nothing but ~25 doubles of electronic state): E electronic steps --
//      eigenproblem, force scalars, Tsit5, hop test, estimators -- with full occupancy, like the 1-D kernels.
// plus sb_prep_kernel at launch entry (force scalars -> first impulse, entry half kick) and a last bath pass at launch
// exit (replay + second half kick: true velocities back in KParams::v).
//
// Covers the outputs that need no bath coordinates at the save points (populations, correlation functions, sigma,
// discrete state); energies / positions / velocities / diagnostics keep the step-by-step kernel
// (select_density_spinboson).  tr(sigma) != 1 (Ehrenfest force scalar A != 1: the oscillator frequency itself becomes
// trajectory dependent, the lag tables do not apply) makes the engine run epochs of ONE step, which is exact.
// Rounding: r, v follow the FMA sequence of kernel_spinboson.cuh; the bath sums are assembled in a different order
// (free part + lag table), i.e. they differ in the last bits -- far inside the 1e-10 parity tolerance.
#pragma once
#include <cuda_pipeline.h>

#include "kernel_spinboson.cuh"

namespace nq {

#if defined(__CUDACC__)

constexpr int kSeLagStride = 32;   // KParams::sb_kap: [2 shapes][3 sums][32 lags], then sum c^2/m, sum c^2
constexpr int kSeThreads = 128;
constexpr int kSeGroup = 4;        // modes a thread advances in lock step (independent FMA chains)
constexpr int kSeStages = 4;       // depth of the bath pass's cp.async ring

// impulses of one epoch: sb_f[(s * (E + 1) + i) * T + traj], s = 0 (f1) | 1 (f2), i = 0 .. E (entry kb: the carry)
template <int E>
NQ_D int64_t se_f_index(int s, int i, int64_t T, int64_t traj) { return ((int64_t)(s * (E + 1) + i)) * T + traj; }

NQ_D double4 se_ldkc(const double* __restrict__ kc, int j) {      // uniform address: one broadcast transaction
    const double2* q = reinterpret_cast<const double2*>(kc) + 2 * j;
    const double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// ---- launch entry: force scalars -> first impulse and entry half kick -----------------------------------------------
template <int METHOD, int E>
__global__ void __launch_bounds__(kSeThreads) sb_prep_kernel(const __grid_constant__ KParams p) {
    constexpr int N = 2;
    const int64_t T = p.ntraj, traj = p.tlo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (traj >= p.thi) return;
    double A, B;
    if (p.step0 == 0) {
        // first step after set_state: no hop has happened, sigma is sigma(t0), Zprev the eigenvectors at r0
        Eig<N> e;
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = 0; k < N; ++k) e.Z[j][k] = p.Zprev[(int64_t)(j + N * k) * T + traj];
        Herm<N> s;
        s.x[0] = p.sig_re[(int64_t)0 * T + traj]; s.x[1] = p.sig_re[(int64_t)2 * T + traj]; s.x[2] = p.sig_re[(int64_t)3 * T + traj];
        s.y[0] = p.sig_im[(int64_t)2 * T + traj];
        const int st = p.state ? p.state[traj] : 0;
        double s00, s01, s11;
        sb_sz(e, s00, s01, s11);
        sb_force_scalars<METHOD>(s, st, s.x[0] + s.x[2], s00, s01, s11, A, B);
        p.sb_carry[traj] = A; p.sb_carry[T + traj] = B;
    } else {
        A = p.sb_carry[traj]; B = p.sb_carry[T + traj];
    }
    p.sb_aux[traj] = 0.5 * p.dt * B;                                   // xk: u = v + 1/2 (A a r + dt B c/m) on entry
    p.sb_f[se_f_index<E>(0, 0, T, traj)] = p.dt * B;                   // impulse of the first step
    p.sb_f[se_f_index<E>(1, 0, T, traj)] = 0.0;
}

// ---- DynamicsVariables at t0 (fssh.jl:47-65, ehrenfest.jl:43-48): eigenproblem at r0, sigma = Z' rho Z, initial state
// ~ diag(sigma), zeroed electronic buffer (quirk Q1), save point 0 -- thread per trajectory -------------------------------
template <int METHOD>
__global__ void __launch_bounds__(kSeThreads) sb_init_kernel(const __grid_constant__ KParams p) {
    constexpr int N = 2;
    const int64_t T = p.ntraj;
    int64_t traj = p.tlo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = traj < p.thi;
    if (!valid) traj = p.thi - 1;
    double lin = 0.0;
    for (int j = 0; j < p.D; ++j) lin = fma(se_ldkc(p.sb_kc, j).z, p.r[(int64_t)j * T + traj], lin);
    SbTraj R;
    R.s.x[0] = p.sig_re[(int64_t)0 * T + traj]; R.s.x[1] = p.sig_re[(int64_t)2 * T + traj]; R.s.x[2] = p.sig_re[(int64_t)3 * T + traj];
    R.s.y[0] = p.sig_im[(int64_t)2 * T + traj];
    R.st = p.state ? p.state[traj] : 0;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = 0; k < N; ++k) R.Zref[j][k] = p.Zprev[(int64_t)(j + N * k) * T + traj];
    Eig<N> e;
    const double l = p.params[0] + lin;
    double Vp[3] = {l, p.params[1], -l};
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
    SbSmem Mdummy; Mdummy.rs = nullptr; Mdummy.vs = nullptr; Mdummy.k1 = Mdummy.k2 = Mdummy.k3 = nullptr;
    SbEmitter em{p, traj, valid, 0};
    sb_record_save<METHOD>(p, em, Mdummy, 0, R, e, 0.0);      // also writes pop0 (correlation functions)
    if (valid) {
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = 0; k < N; ++k) {
                p.sig_re[(int64_t)(j + N * k) * T + traj] = R.s.X(j, k);
                p.sig_im[(int64_t)(j + N * k) * T + traj] = R.s.Y(j, k);
                p.Zprev[(int64_t)(j + N * k) * T + traj] = R.Zref[j][k];
            }
        if (p.state) p.state[traj] = R.st;
        p.ecur[(int64_t)0 * T + traj] = 0.0; p.ecur[(int64_t)1 * T + traj] = 0.0;
        p.ecur[(int64_t)(N + 0 + N * 1) * T + traj] = 0.0;
    }
}

// ---- bath pass -------------------------------------------------------------------------------------------------------
//   sb_entry : KParams::v holds true velocities (launch entry): u = v + 1/2 (A a r + dt B c/m) first
//   sb_nrep  : replay steps (the previous epoch's, impulses f[0 .. nrep-1])
//   sb_nfree : free-evolution steps of the new epoch; its first step carries the impulse f[nrep]
//   sb_exit  : launch exit: after the replay, finish the last half kick and store TRUE velocities
//   FULL     : nrep == nfree == E, no entry / exit work (every epoch of a long run but the first and the last): without
//              the per-step conditions the E steps of a group form ONE basic block, which lets ptxas interleave the
//              dependent accumulator FMAs of step k with the coordinate FMAs of step k + 1
template <int E, bool VINV, bool GEN, bool FULL>
NQ_D void se_bath_body(const KParams& p, int64_t traj, double* ring) {
    constexpr int W = kSeGroup;
    const int64_t T = p.ntraj;
    const int D = p.D, nrep = FULL ? E : p.sb_nrep, nfree = FULL ? E : p.sb_nfree;
    const bool entry = !FULL && p.sb_entry, exitk = !FULL && p.sb_exit;
    const double dt = p.dt;
    const double* __restrict__ kcg = p.sb_kc;
    // L before the first free step and after each of them (C follows from their differences: every mode drifts by
    // dt vt, so sum c vt_k = (L_k - L_{k-1}) / dt exactly in real arithmetic), and W
    double aL[E + 1], aW[E];
#pragma unroll
    for (int k = 0; k < E; ++k) { aL[k] = 0.0; aW[k] = 0.0; }
    aL[E] = 0.0;
    double f1[E], f2[VINV ? E : 1];
#pragma unroll
    for (int k = 0; k < E; ++k) {
        f1[k] = (FULL || k < nrep) ? p.sb_f[se_f_index<E>(0, k, T, traj)] : 0.0;
        if (VINV) f2[k] = (FULL || k < nrep) ? p.sb_f[se_f_index<E>(1, k, T, traj)] : 0.0;
    }
    const double f1c = p.sb_f[se_f_index<E>(0, nrep, T, traj)];                      // the next step's impulse
    const double f2c = VINV ? p.sb_f[se_f_index<E>(1, nrep, T, traj)] : 0.0;
    const double A = GEN ? p.sb_carry[traj] : 1.0;
    const double xk = (entry || exitk) ? p.sb_aux[traj] : 0.0;
    const bool store = FULL || nrep > 0 || entry || exitk;
    // The coordinates of the next kSeStages - 1 groups of modes are in flight (cp.async into this thread's own slots of
    // a shared-memory ring: no registers, no barriers) while this group's arithmetic runs -- two or three warps per
    // scheduler do not hide an HBM round trip on their own.
    const int ngroups = (D + W - 1) / W;
    auto issue = [&](int g) {
        if (g < ngroups) {
            double* dst = ring + ((size_t)(g % kSeStages) * 2 * W) * kSeThreads + threadIdx.x;
#pragma unroll
            for (int q = 0; q < W; ++q) {
                const int j = (g * W + q < D) ? g * W + q : g * W;
                __pipeline_memcpy_async(dst + (size_t)q * kSeThreads, p.r + (int64_t)j * T + traj, sizeof(double));
                __pipeline_memcpy_async(dst + (size_t)(W + q) * kSeThreads, p.v + (int64_t)j * T + traj, sizeof(double));
            }
        }
        __pipeline_commit();
    };
#pragma unroll
    for (int g = 0; g < kSeStages - 1; ++g) issue(g);
    for (int j0 = 0, g = 0; j0 < D; j0 += W, ++g) {
        double r[W], u[W];
        double4 kc[W];
        bool ok[W];
        issue(g + kSeStages - 1);
        __pipeline_wait_prior(kSeStages - 1);
        const double* src = ring + ((size_t)(g % kSeStages) * 2 * W) * kSeThreads + threadIdx.x;
#pragma unroll
        for (int q = 0; q < W; ++q) {
            const int j = j0 + q;
            ok[q] = j < D;
            r[q] = src[(size_t)q * kSeThreads]; u[q] = src[(size_t)(W + q) * kSeThreads];
            kc[q] = se_ldkc(kcg, ok[q] ? j : j0);
            if (!ok[q]) kc[q] = make_double4(0.0, 0.0, 0.0, 0.0);      // a tail mode contributes nothing (and is not stored)
            if (GEN) kc[q].x *= A;
        }
        if (entry) {
#pragma unroll
            for (int q = 0; q < W; ++q) u[q] = fma(xk, kc[q].y, fma(0.5 * kc[q].x, r[q], u[q]));
        }
#pragma unroll
        for (int k = 0; k < E; ++k) {
            if (FULL || k < nrep) {
#pragma unroll
                for (int q = 0; q < W; ++q) {
                    double v = u[q];
                    if (VINV) v = fma(-f2[k], kc[q].z, v);
                    const double vt = fma(-f1[k], kc[q].y, fma(-kc[q].x, r[q], v));       // step_B! twice (steps.jl:3-5)
                    r[q] = fma(dt, vt, r[q]);                                              // step_A! (steps.jl:6-8)
                    u[q] = vt;
                }
            }
        }
        if (exitk) {
            // true velocity: the second half kick of the last step and whatever its hop left pending
#pragma unroll
            for (int q = 0; q < W; ++q) {
                double v = u[q];
                if (VINV) v = fma(-f2c, kc[q].z, v);
                u[q] = fma(-(f1c - xk), kc[q].y, fma(-0.5 * kc[q].x, r[q], v));           // f1c - xk = 1/2 dt B + gamma_m
            }
        }
        if (store) {
#pragma unroll
            for (int q = 0; q < W; ++q)
                if (ok[q]) { p.r[(int64_t)(j0 + q) * T + traj] = r[q]; p.v[(int64_t)(j0 + q) * T + traj] = u[q]; }
        }
        if (FULL || nfree > 0) {
#pragma unroll
            for (int q = 0; q < W; ++q) aL[0] = fma(kc[q].z, r[q], aL[0]);
        }
#pragma unroll
        for (int k = 0; k < E; ++k) {
            if (FULL || k < nfree) {
#pragma unroll
                for (int q = 0; q < W; ++q) {
                    double v = u[q];
                    if (k == 0) {
                        if (VINV) v = fma(-f2c, kc[q].z, v);
                        v = fma(-f1c, kc[q].y, v);
                    }
                    const double vt = fma(-kc[q].x, r[q], v);
                    r[q] = fma(dt, vt, r[q]);
                    u[q] = vt;
                    aL[k + 1] = fma(kc[q].z, r[q], aL[k + 1]);
                    aW[k] = fma(kc[q].w, r[q], aW[k]);
                }
            }
        }
    }
    const double rdt = 1.0 / dt;
#pragma unroll
    for (int k = 0; k < E; ++k) {
        if (FULL || k < nfree) {
            p.sb_sums[(int64_t)(3 * k + 0) * T + traj] = aL[k + 1];
            p.sb_sums[(int64_t)(3 * k + 1) * T + traj] = (aL[k + 1] - aL[k]) * rdt;
            p.sb_sums[(int64_t)(3 * k + 2) * T + traj] = aW[k];
        }
    }
}

template <int E>
__global__ void __launch_bounds__(kSeThreads, 3) sb_bath_kernel(const __grid_constant__ KParams p) {
    __shared__ double ring[kSeStages * 2 * kSeGroup * kSeThreads];
    const int64_t traj = p.tlo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (traj >= p.thi) return;
    const bool vinv = p.rescaling == NQCB200_RESCALE_VINVERSION;
    const bool full = p.sb_nrep == E && p.sb_nfree == E && !p.sb_entry && !p.sb_exit;
    if (p.sb_gen) {        // epochs of one step
        if (vinv) se_bath_body<E, true, true, false>(p, traj, ring); else se_bath_body<E, false, true, false>(p, traj, ring);
    } else if (full) {
        if (vinv) se_bath_body<E, true, false, true>(p, traj, ring); else se_bath_body<E, false, false, true>(p, traj, ring);
    } else {
        if (vinv) se_bath_body<E, true, false, false>(p, traj, ring); else se_bath_body<E, false, false, false>(p, traj, ring);
    }
}

// ---- electronic steps of one epoch ------------------------------------------------------------------------------------
template <int METHOD, int E>
__global__ void __launch_bounds__(kSeThreads, 4) sb_elec_kernel(const __grid_constant__ KParams p) {
    constexpr int N = 2;
    __shared__ double fs[2][E + 1][kSeThreads];      // this epoch's impulses (f1, f2), per thread
    __shared__ double kaps[6][E];                    // lag tables: (c/m | c) x (L, C, W) x lag
    const int tid = threadIdx.x;
    const int64_t T = p.ntraj;
    int64_t traj = p.tlo + (int64_t)blockIdx.x * blockDim.x + tid;
    const bool valid = traj < p.thi;
    if (!valid) traj = p.thi - 1;
    const double dt = p.dt, hdt = 0.5 * p.dt;
    const bool vinv = p.rescaling == NQCB200_RESCALE_VINVERSION;
    const double* __restrict__ kap = p.sb_kap;
    const double C2 = kap[6 * kSeLagStride], Cc = kap[6 * kSeLagStride + 1];     // sum c^2/m, sum c^2
    const int kb = p.nsteps;
    for (int idx = tid; idx < 6 * E; idx += kSeThreads) kaps[idx / E][idx % E] = kap[(idx / E) * kSeLagStride + idx % E];
    __syncthreads();

    SbTraj R;
    Eig<N> e;
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
        for (int i = 0; i < N; ++i) {
            R.p0d[i] = corr ? p.pop0[(int64_t)i * T + traj] : 0.0;
            R.p0a[i] = corr ? p.pop0[(int64_t)(N + i) * T + traj] : 0.0;
        }
    }
    R.cur.E[0] = p.ecur[(int64_t)0 * T + traj]; R.cur.E[1] = p.ecur[(int64_t)1 * T + traj];
    R.cur.g[0] = p.ecur[(int64_t)(N + 0 + N * 1) * T + traj];
    R.A = p.sb_carry[traj]; R.B = p.sb_carry[T + traj];
    const double tr0 = R.s.x[0] + R.s.x[2];
    fs[0][0][tid] = p.sb_f[se_f_index<E>(0, p.sb_nrep, T, traj)];      // carried impulse = this epoch's first
    fs[1][0][tid] = p.sb_f[se_f_index<E>(1, p.sb_nrep, T, traj)];
    unsigned long long nhops = 0, nfrus = 0;
    SbSmem Mdummy; Mdummy.rs = nullptr; Mdummy.vs = nullptr; Mdummy.k1 = Mdummy.k2 = Mdummy.k3 = nullptr;
    int64_t next_save = ((p.step0 / p.save_every) + 1) * (int64_t)p.save_every;     // first step count that is a save point

#pragma unroll 1
    for (int k = 0; k < kb; ++k) {
        const int64_t step = p.step0 + k;
        const double t = p.t0 + dt * (double)step;
        const double tcur = (step == 0) ? 0.0 : t;   // Q1
        double L = p.sb_sums[(int64_t)(3 * k + 0) * T + traj], Cv = p.sb_sums[(int64_t)(3 * k + 1) * T + traj],
               Wr = p.sb_sums[(int64_t)(3 * k + 2) * T + traj];
        {
            // impulses of the epoch's later steps through the lag tables (shared-memory copy); two interleaved partial
            // sums per bath sum halve the dependent FMA chains
            double L2 = 0.0, C2b = 0.0, W2 = 0.0;
            int i = 1;
            for (; i + 1 <= k; i += 2) {
                const double fa = fs[0][i][tid], fb = fs[0][i + 1][tid];
                L = fma(fa, kaps[0][k - i], L);          L2 = fma(fb, kaps[0][k - i - 1], L2);
                Cv = fma(fa, kaps[1][k - i], Cv);        C2b = fma(fb, kaps[1][k - i - 1], C2b);
                Wr = fma(fa, kaps[2][k - i], Wr);        W2 = fma(fb, kaps[2][k - i - 1], W2);
            }
            if (i <= k) {
                const double fa = fs[0][i][tid];
                L = fma(fa, kaps[0][k - i], L); Cv = fma(fa, kaps[1][k - i], Cv); Wr = fma(fa, kaps[2][k - i], Wr);
            }
            if (vinv) {
                for (int j = 1; j <= k; ++j) {
                    const double gi = fs[1][j][tid];
                    L2 = fma(gi, kaps[3][k - j], L2); C2b = fma(gi, kaps[4][k - j], C2b); W2 = fma(gi, kaps[5][k - j], W2);
                }
            }
            L += L2; Cv += C2b; Wr += W2;
        }
        // update_cache!: V -> eigen (gauge-fixed); the harmonic shift is a multiple of the identity
        {
            const double l = p.params[0] + L;
            double Vp[3] = {l, p.params[1], -l};
            sym_eigh<N>(Vp, e);
            fix_gauge<N>(e, R.Zref);
        }
        double s00, s01, s11;
        sb_sz(e, s00, s01, s11);
        sb_force_scalars<METHOD>(R.s, R.st, tr0, s00, s01, s11, R.A, R.B);      // pre-hop state, sigma_prev
        const double cv = Cv - hdt * (R.A * Wr + R.B * C2);
        const double dfac = div_fast(-s01, e.w[0] - e.w[1]);
        ElecParams<N> nxt;
        nxt.E[0] = e.w[0]; nxt.E[1] = e.w[1];
        nxt.g[0] = dfac * cv;
        propagate_density<N>(R.cur, tcur, nxt, t + dt, t, dt, R.s, p.tsit5_ha);
        double gm = 0.0, gd = 0.0;
        if (METHOD == NQCB200_METHOD_FSSH) {
            const double xi = (p.rng == NQCB200_RNG_INJECTED)
                                  ? p.draws[(step - p.draws_step0) * T + traj]
                                  : philox_uniform(p.seed, (uint64_t)(p.traj_offset + traj), (uint64_t)step, 0u);
            const int s0 = R.st, m = 1 - s0;
            double g = 2.0 * div_fast(R.s.x[1], s0 ? R.s.x[2] : R.s.x[0]) * (s0 ? -nxt.g[0] : nxt.g[0]) * dt;   // fssh.jl:96-121 (Q4)
            g = fmin(1.0, fmax(0.0, g));
            if (g > xi) {
                bool accept = true;
                if (p.rescaling != NQCB200_RESCALE_OFF) {                   // surface_hopping.jl:64-99
                    const double wn = m ? e.w[1] : e.w[0], wo = s0 ? e.w[1] : e.w[0];
                    const double df = -s01 / (wn - wo);
                    const double a = 0.5 * df * df * C2, b = df * cv, c = wn - wo;
                    const double disc = b * b - 4.0 * a * c;
                    if (disc < 0.0) {
                        accept = false;
                        nfrus += valid;
                        if (vinv) {                                         // v -= 2 (v.dhat) dhat
                            const double nrm = sqrt(df * df * Cc);
                            const double gam = b / nrm;
                            gd = 2.0 * gam * df / nrm;
                        }
                    } else {
                        const double root = sqrt(disc);
                        const double gam = (b < 0.0) ? (b + root) / (2.0 * a) : (b - root) / (2.0 * a);
                        gm = gam * df;
                    }
                }
                if (accept) { R.st = m; nhops += valid; }
            }
        }
        R.cur = nxt;
        fs[0][k + 1][tid] = fma(dt, R.B, gm);      // impulse of the next step
        fs[1][k + 1][tid] = gd;
        if (step + 1 == next_save) {
            const int64_t isave = next_save / p.save_every;
            next_save += p.save_every;
            if (isave < p.nsave) {
                SbEmitter em{p, traj, valid, (int)isave};
                sb_record_save<METHOD>(p, em, Mdummy, 0, R, e, 0.0);
            }
        }
    }

    if (valid) {
        for (int k = 0; k <= kb; ++k) {
            p.sb_f[se_f_index<E>(0, k, T, traj)] = fs[0][k][tid];
            p.sb_f[se_f_index<E>(1, k, T, traj)] = fs[1][k][tid];
        }
        p.sb_aux[traj] = hdt * R.B;
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
    }
    const unsigned long long wh = __reduce_add_sync(0xffffffffu, (unsigned)nhops);
    const unsigned long long wf = __reduce_add_sync(0xffffffffu, (unsigned)nfrus);
    if ((tid & 31) == 0) {
        if (wh) atomicAdd(&p.counters[0], wh);
        if (wf) atomicAdd(&p.counters[1], wf);
    }
}

#endif  // __CUDACC__

}  // namespace nq

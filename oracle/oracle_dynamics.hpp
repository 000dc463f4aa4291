// oracle_dynamics.hpp -- CPU ORACLE (test infrastructure, NOT product code; see oracle_core.hpp).
//
// Per-trajectory restatement of the reference's fixed-step integrators and hop callbacks:
//   BABwithTsit5   src/DynamicsMethods/IntegrationAlgorithms/bab_electronics.jl:15-91
//   BCBwithTsit5   .../bcb_electronics.jl:21-97
//   BCB            .../bcb.jl:81-116          (+ OrdinaryDiffEq VelocityVerlet for B = 1)
//   VerletwithElectronics .../verlet_with_electronics.jl:42-69
//   RingPolymerMInt .../ringpolymer_mint.jl:28-130
//   DensityMatrixODEProblem + double buffer   src/DynamicsMethods/electronic_dynamics.jl:15-140
//   FSSH hop / rescale  SurfaceHoppingMethods/{fssh.jl:86-130, surface_hopping.jl:2-168, rpsh.jl:30-50}
//   IESH  SurfaceHoppingMethods/iesh.jl:190-411, DynamicsUtils/wavefunction_dynamics.jl:15-58
//   Tsit5 (OrdinaryDiffEq, external): tableau of Tsitouras 2011, fixed dt/5 sub-steps.
// Reference quirks Q1-Q7 are reproduced on purpose and marked "Q<n>".
#pragma once
#include "oracle_core.hpp"

namespace nqco {

// ---- Tsit5 tableau (Tsitouras 2011; OrdinaryDiffEq Tsit5ConstantCache) -----------------------
namespace tsit5 {
constexpr double c1 = 0.161, c2 = 0.327, c3 = 0.9, c4 = 0.9800255409045097;
constexpr double a21 = 0.161;
constexpr double a31 = -0.008480655492356989, a32 = 0.335480655492357;
constexpr double a41 = 2.8971530571054935, a42 = -6.359448489975075, a43 = 4.3622954328695815;
constexpr double a51 = 5.325864828439257, a52 = -11.748883564062828, a53 = 7.4955393428898365,
                 a54 = -0.09249506636175525;
constexpr double a61 = 5.86145544294642, a62 = -12.92096931784711, a63 = 8.159367898576159,
                 a64 = -0.071584973281401, a65 = -0.028269050394068383;
constexpr double a71 = 0.09646076681806523, a72 = 0.01, a73 = 0.4798896504144996, a74 = 1.379008574103742,
                 a75 = -3.290069515436081, a76 = 2.324710524099774;
}  // namespace tsit5

// electronic_dynamics.jl:15-36 -- one half of the DoubleBuffer
struct ElectronicParameters {
    cvec vd;  // dynamical_coupling  n*n (stored complex in the reference)
    cvec E;   // eigenvalues n
    double t = 0.0;
};

struct Counters {
    int64_t steps = 0, hops = 0, frustrated = 0, nonfinite = 0, hop_searches = 0;
};

struct Setup {
    nqcb200_config cfg;
    Model model;
    vec masses;           // [D]
    vec U;                // normal-mode matrix B*B
    vec cayley;           // 4*B (full step; half step for NRPMD)
    double omega_n = 0.0;
    int n, D, B, ne;
};

// One trajectory's DynamicsVariables + integrator cache + calculator caches.
struct Trajectory {
    vec r, v, k;          // [B][D]  (bead-major: index d + D*b) ; k = acceleration
    cvec sigma;           // n*n density matrix (FSSH/Ehrenfest) or n*ne psi (IESH)
    int state = 0;        // 0-based active state (FSSH)
    std::vector<int> occ; // IESH occupied states (0-based, sorted)
    vec qmap, pmap;       // NRPMD [B][n]
    std::vector<Cache> bead;  // B caches
    Cache centroid;       // RP only
    ElectronicParameters cur, nxt;   // DoubleBuffer (current, next)
    vec pop0;             // initial population for the correlation function
    int64_t step = 0;
    int64_t term_step = -1;   // TerminatingCallback (callbacks.jl:29): step count at which terminate! fired, -1 = running
    Counters cnt;
    vec last_nac, last_eig, last_Z;
    vec noise;            // ThermalLangevin: the B*D standard normals of the current step (W.dW / sqrt(dt))
};

inline void centroid_of(const Setup& S, const vec& x, double* out) {
    for (int d = 0; d < S.D; ++d) {
        double s = 0.0;
        for (int b = 0; b < S.B; ++b) s += x[d + (size_t)S.D * b];
        out[d] = s / S.B;
    }
}

// hopping quantities: plain Simulation -> the single cache; ring polymer -> centroid cache
// (SurfaceHoppingMethods.jl:81-103)
inline const Cache& hop_cache(const Setup& S, const Trajectory& T) { return S.B > 1 ? T.centroid : T.bead[0]; }
inline void hop_velocity(const Setup& S, const Trajectory& T, double* vh) {
    if (S.B > 1) centroid_of(S, T.v, vh); else std::copy(T.v.begin(), T.v.begin() + S.D, vh);
}

inline void update_all_caches(const Setup& S, Trajectory& T, const vec& r) {
    for (int b = 0; b < S.B; ++b) T.bead[b].update(S.model, &r[(size_t)S.D * b]);
    if (S.B > 1) {
        vec rc(S.D);
        centroid_of(S, r, rc.data());
        T.centroid.update(S.model, rc.data());
    }
}

// ---- accelerations ---------------------------------------------------------------------------
// FSSH fssh.jl:67-74 ; Ehrenfest ehrenfest.jl:50-68, ehrenfest_rpmd.jl:23-43 ;
// IESH iesh.jl:190-207 ; classical classical.jl:63-67
inline void acceleration(const Setup& S, Trajectory& T, const vec& r, const cvec& sigma_prev) {
    const int n = S.n, D = S.D;
    for (int b = 0; b < S.B; ++b) {
        const Cache& c = T.bead[b];
        vec g0(D, 0.0);
        for (int I = 0; I < D; ++I) {
            double f = 0.0;
            const double* a = S.model.classical() ? nullptr : &c.adiab[(size_t)I * n * n];
            switch (S.cfg.method) {
                case NQCB200_METHOD_FSSH: f = -a[T.state + (size_t)n * T.state]; break;
                case NQCB200_METHOD_EHRENFEST: {
                    if (I == 0) S.model.dU0(&r[(size_t)D * b], g0.data());
                    f = -g0[I];
                    for (int m = 0; m < n; ++m)
                        for (int nn = 0; nn < n; ++nn) f -= a[nn + (size_t)n * m] * sigma_prev[nn + (size_t)n * m].real();
                } break;
                case NQCB200_METHOD_IESH: {
                    if (I == 0) S.model.dU0(&r[(size_t)D * b], g0.data());
                    f = -g0[I];
                    for (int kk : T.occ) f -= a[kk + (size_t)n * kk];
                } break;
                case NQCB200_METHOD_EHRENFEST_NA: {   // ehrenfest_na.jl:72-90 (psi = sigma_prev of the wavefunction integrator)
                    if (I == 0) S.model.dU0(&r[(size_t)D * b], g0.data());
                    f = -g0[I];
                    for (int e = 0; e < S.ne; ++e)
                        for (int m = 0; m < n; ++m)
                            for (int nn = 0; nn < n; ++nn)
                                f -= a[nn + (size_t)n * m] * (T.sigma[nn + (size_t)n * e] * std::conj(T.sigma[m + (size_t)n * e])).real();
                } break;
                case NQCB200_METHOD_CLASSICAL:
                case NQCB200_METHOD_THERMAL_LANGEVIN: f = -c.dV[I]; break;
                default: throw std::runtime_error("acceleration: method");
            }
            T.k[I + (size_t)D * b] = f / S.masses[I];
        }
    }
}

// ---- electronic density-matrix propagation ---------------------------------------------------
// RHS of DensityMatrixODEProblem, electronic_dynamics.jl:106-116 with interpolate_* :55-79 and
// commutator! density_matrix_dynamics.jl:29-33.
inline void density_rhs(int n, const ElectronicParameters& cur, const ElectronicParameters& nxt, double t,
                        const cd* u, cd* du, cd* A) {
    double loc = (t - cur.t) / (nxt.t - cur.t);
    if (std::isnan(loc)) loc = 0.0;
    const cd mi(0.0, -1.0);
    for (int i = 0; i < n * n; ++i) A[i] = (cur.vd[i] + (nxt.vd[i] - cur.vd[i]) * loc) * mi;
    for (int i = 0; i < n; ++i) A[i + (size_t)n * i] = cur.E[i] + (nxt.E[i] - cur.E[i]) * loc;
    // du = A u - u A ; du *= -i
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            cd s = 0.0;
            for (int k = 0; k < n; ++k) s += A[i + (size_t)n * k] * u[k + (size_t)n * j] - u[i + (size_t)n * k] * A[k + (size_t)n * j];
            du[i + (size_t)n * j] = s * mi;
        }
}

// set_ut!(integrator, sigma, t); step!(integrator, dt, true) with Tsit5, adaptive=false,
// dt_sub = dt/5 (bab_electronics.jl:40-43,88-89): 5 sub-steps, 31 RHS evaluations (FSAL is
// recomputed after set_ut!, then reused: identical values either way).
inline void propagate_density(int n, const ElectronicParameters& cur, const ElectronicParameters& nxt,
                              double t, double dt, cvec& sigma) {
    using namespace tsit5;
    const int nn = n * n;
    cvec k1(nn), k2(nn), k3(nn), k4(nn), k5(nn), k6(nn), k7(nn), tmp(nn), A(nn);
    const double h = dt / 5.0;
    double ts = t;
    density_rhs(n, cur, nxt, ts, sigma.data(), k1.data(), A.data());
    for (int sub = 0; sub < 5; ++sub) {
        double hh = h;
        if (sub == 4) hh = (t + dt) - ts;  // tstop snapping of the last sub-step
        for (int i = 0; i < nn; ++i) tmp[i] = sigma[i] + hh * (a21 * k1[i]);
        density_rhs(n, cur, nxt, ts + c1 * hh, tmp.data(), k2.data(), A.data());
        for (int i = 0; i < nn; ++i) tmp[i] = sigma[i] + hh * (a31 * k1[i] + a32 * k2[i]);
        density_rhs(n, cur, nxt, ts + c2 * hh, tmp.data(), k3.data(), A.data());
        for (int i = 0; i < nn; ++i) tmp[i] = sigma[i] + hh * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
        density_rhs(n, cur, nxt, ts + c3 * hh, tmp.data(), k4.data(), A.data());
        for (int i = 0; i < nn; ++i) tmp[i] = sigma[i] + hh * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
        density_rhs(n, cur, nxt, ts + c4 * hh, tmp.data(), k5.data(), A.data());
        for (int i = 0; i < nn; ++i)
            tmp[i] = sigma[i] + hh * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
        density_rhs(n, cur, nxt, ts + hh, tmp.data(), k6.data(), A.data());
        for (int i = 0; i < nn; ++i)
            sigma[i] = sigma[i] + hh * (a71 * k1[i] + a72 * k2[i] + a73 * k3[i] + a74 * k4[i] + a75 * k5[i] + a76 * k6[i]);
        ts = (sub == 4) ? (t + dt) : ts + hh;
        density_rhs(n, cur, nxt, ts, sigma.data(), k7.data(), A.data());
        k1.swap(k7);
    }
}

// update_parameters!, electronic_dynamics.jl:38-53 : swap, then fill `next`
inline void update_parameters(const Setup& S, Trajectory& T, const Cache& c, const double* vhop, double tnext) {
    std::swap(T.cur, T.nxt);
    const int n = S.n;
    for (int i = 0; i < n; ++i) T.nxt.E[i] = c.w[i];
    std::fill(T.nxt.vd.begin(), T.nxt.vd.end(), cd(0.0));
    for (int I = 0; I < S.D; ++I)
        for (int J = 0; J < n * n; ++J) T.nxt.vd[J] += c.nac[(size_t)I * n * n + J] * vhop[I];
    T.nxt.t = tnext;
}

// ---- nuclear propagation ----------------------------------------------------------------------
inline void to_normal_modes(const Setup& S, vec& x) {  // x_k = sum_j U[j,k] x_j
    const int B = S.B, D = S.D;
    vec tmp(B);
    for (int d = 0; d < D; ++d) {
        for (int k = 0; k < B; ++k) {
            double s = 0.0;
            for (int j = 0; j < B; ++j) s += S.U[j + (size_t)B * k] * x[d + (size_t)D * j];
            tmp[k] = s;
        }
        for (int k = 0; k < B; ++k) x[d + (size_t)D * k] = tmp[k];
    }
}
inline void from_normal_modes(const Setup& S, vec& x) {  // x_j = sum_k U[j,k] x_k
    const int B = S.B, D = S.D;
    vec tmp(B);
    for (int d = 0; d < D; ++d) {
        for (int j = 0; j < B; ++j) {
            double s = 0.0;
            for (int k = 0; k < B; ++k) s += S.U[j + (size_t)B * k] * x[d + (size_t)D * k];
            tmp[j] = s;
        }
        for (int j = 0; j < B; ++j) x[d + (size_t)D * j] = tmp[j];
    }
}
inline void step_C(const Setup& S, vec& v, vec& r) {  // steps.jl:10-17
    for (int b = 0; b < S.B; ++b)
        for (int d = 0; d < S.D; ++d) {
            size_t i = d + (size_t)S.D * b;
            const double* c = &S.cayley[4 * b];
            double rt = c[0] * r[i] + c[1] * v[i];
            double vt = c[2] * r[i] + c[3] * v[i];
            r[i] = rt; v[i] = vt;
        }
}

// ---- FSSH hop ---------------------------------------------------------------------------------
// evaluate_hopping_probability! / fewest_switches_probability! fssh.jl:86-108 (Q4),
// select_new_state fssh.jl:110-121, rescale_velocity! surface_hopping.jl:64-99 (+ rpsh.jl:30-50)
// fewest_switches_probability! fssh.jl:96-108: clamped, then cumulative
inline void fssh_probabilities(const Setup& S, const Trajectory& T, vec& prob) {
    const int n = S.n, D = S.D, s = T.state;
    const Cache& c = hop_cache(S, T);
    vec vh(D);
    hop_velocity(S, T, vh.data());
    prob.assign(n, 0.0);
    for (int m = 0; m < n; ++m) {
        if (m == s) continue;
        for (int I = 0; I < D; ++I) {
            cd ratio = T.sigma[m + (size_t)n * s] / T.sigma[s + (size_t)n * s];
            prob[m] += 2.0 * vh[I] * ratio.real() * c.nac[(size_t)I * n * n + s + (size_t)n * m] * S.cfg.dt;
        }
    }
    for (int m = 0; m < n; ++m) prob[m] = std::min(1.0, std::max(0.0, prob[m]));
    for (int m = 1; m < n; ++m) prob[m] += prob[m - 1];
}
// select_new_state fssh.jl:110-121 (prob is the cumulative vector)
inline int fssh_select(const vec& prob, int s, double xi) {
    for (int m = 0; m < (int)prob.size(); ++m)
        if (m != s && prob[m] > xi) return m;
    return s;
}
// rescale_velocity! surface_hopping.jl:64-99 (+ RP variants rpsh.jl:30-50); true = hop accepted
inline bool rescale_velocity(const Setup& S, Trajectory& T, int new_state, int old_state) {
    if (S.cfg.rescaling == NQCB200_RESCALE_OFF) return true;
    const int n = S.n, D = S.D;
    const Cache& c = hop_cache(S, T);
    vec vh(D), d(D);
    hop_velocity(S, T, vh.data());
    for (int I = 0; I < D; ++I) d[I] = c.nac[(size_t)I * n * n + new_state + (size_t)n * old_state];  // d[I][new, old]
    double a = 0.0, b = 0.0;
    for (int I = 0; I < D; ++I) { a += d[I] * d[I] / S.masses[I]; b += d[I] * vh[I]; }
    a /= 2.0;
    double cc = c.w[new_state] - c.w[old_state];
    double disc = b * b - 4.0 * a * cc;
    if (disc < 0.0) {
        if (S.cfg.rescaling == NQCB200_RESCALE_VINVERSION) {
            double nrm = 0.0;
            for (int I = 0; I < D; ++I) nrm += d[I] * d[I];
            nrm = std::sqrt(nrm);
            double gam = 0.0;
            for (int I = 0; I < D; ++I) gam += vh[I] * d[I] / nrm;  // (RP: bead-average velocity)
            for (int b2 = 0; b2 < S.B; ++b2)
                for (int I = 0; I < D; ++I) T.v[I + (size_t)D * b2] -= 2.0 * gam * d[I] / nrm;
        }
        return false;
    }
    double root = std::sqrt(disc);
    double gam = (b < 0.0) ? (b + root) / (2.0 * a) : (b - root) / (2.0 * a);
    for (int b2 = 0; b2 < S.B; ++b2)
        for (int I = 0; I < D; ++I) T.v[I + (size_t)D * b2] -= gam * d[I] / S.masses[I];
    return true;
}
inline void fssh_hop(const Setup& S, Trajectory& T, double xi) {
    vec prob;
    fssh_probabilities(S, T, prob);
    const int s = T.state;
    const int new_state = fssh_select(prob, s, xi);
    if (new_state == s) return;
    // execute_hop! surface_hopping.jl:9-16
    if (rescale_velocity(S, T, new_state, s)) { T.state = new_state; T.cnt.hops++; }
    else T.cnt.frustrated++;
    // Q2: T.k (acceleration) is NOT refreshed; Q3: T.nxt.vd keeps the pre-rescale velocity.
}

// ---- one step: FSSH / Ehrenfest, B = 1 (BABwithTsit5) or B > 1 (BCBwithTsit5) -----------------
inline void step_density_method(const Setup& S, Trajectory& T, double xi) {
    const double dt = S.cfg.dt, t = S.cfg.t0 + dt * (double)T.step;
    const size_t N = (size_t)S.B * S.D;
    cvec sigma_prev = T.sigma;
    vec vtmp(N), rtmp(T.r);
    for (size_t i = 0; i < N; ++i) vtmp[i] = std::fma(dt / 2, T.k[i], T.v[i]);       // step_B!
    if (S.B == 1) {
        for (size_t i = 0; i < N; ++i) rtmp[i] = std::fma(dt, vtmp[i], T.r[i]);     // step_A!
    } else {
        to_normal_modes(S, rtmp); to_normal_modes(S, vtmp);
        step_C(S, vtmp, rtmp);
        from_normal_modes(S, rtmp); from_normal_modes(S, vtmp);
    }
    update_all_caches(S, T, rtmp);
    acceleration(S, T, rtmp, sigma_prev);  // FSSH: current (pre-hop) state; Ehrenfest: sigma_prev
    for (size_t i = 0; i < N; ++i) T.v[i] = std::fma(dt / 2, T.k[i], vtmp[i]);
    T.r = rtmp;
    vec vh(S.D);
    hop_velocity(S, T, vh.data());
    update_parameters(S, T, hop_cache(S, T), vh.data(), t + dt);
    propagate_density(S.n, T.cur, T.nxt, t, dt, T.sigma);
    if (S.cfg.method == NQCB200_METHOD_FSSH) fssh_hop(S, T, xi);
}

// ---- one step: classical MD (B = 1, OrdinaryDiffEq VelocityVerlet) / RPMD (BCB) ---------------
inline void step_classical(const Setup& S, Trajectory& T) {
    const double dt = S.cfg.dt;
    const size_t N = (size_t)S.B * S.D;
    cvec none;
    if (S.B == 1) {
        // VelocityVerlet (OrdinaryDiffEq symplectic_perform_step): u = uprev + dt*duprev + dt^2/2*ku
        vec a_old = T.k;
        for (size_t i = 0; i < N; ++i) T.r[i] = T.r[i] + dt * T.v[i] + dt * dt * 0.5 * a_old[i];
        update_all_caches(S, T, T.r);
        acceleration(S, T, T.r, none);
        for (size_t i = 0; i < N; ++i) T.v[i] = T.v[i] + dt * (0.5 * a_old[i] + 0.5 * T.k[i]);
        return;
    }
    vec vtmp(N), r(T.r);
    for (size_t i = 0; i < N; ++i) vtmp[i] = std::fma(dt / 2, T.k[i], T.v[i]);
    to_normal_modes(S, vtmp); to_normal_modes(S, r);
    step_C(S, vtmp, r);
    from_normal_modes(S, vtmp); from_normal_modes(S, r);
    update_all_caches(S, T, r);
    acceleration(S, T, r, none);
    for (size_t i = 0; i < N; ++i) T.v[i] = std::fma(dt / 2, T.k[i], vtmp[i]);
    T.r = r;
}

// ---- one step: RingPolymerSimulation{ThermalLangevin}, BCOCB  bcocb.jl:95-120 ---------------------
// B, to normal modes, C(1/2), O, C(1/2), from normal modes, force, B.  FrictionCache bcocb.jl:78-87: gamma = [gamma_0,
// 2 sqrt(2 springs_k)] = [gamma_0, 2 omega_k], c1 = exp(-dt gamma), c2 = sqrt(1 - c1^2); O-step steps.jl:109-124:
// v_mode = c1 v_mode + c2 sqrt(T_rp / m) xi with T_rp = nbeads kT (get_ring_polymer_temperature, simulations.jl:104).
inline void step_langevin_bcocb(const Setup& S, Trajectory& T) {
    const double dt = S.cfg.dt;
    const int B = S.B, D = S.D;
    const size_t N = (size_t)B * D;
    const double pi = 3.14159265358979323846;
    cvec none;
    vec vtmp(N), r(T.r);
    for (size_t i = 0; i < N; ++i) vtmp[i] = std::fma(dt / 2, T.k[i], T.v[i]);
    to_normal_modes(S, vtmp); to_normal_modes(S, r);
    step_C(S, vtmp, r);                       // S.cayley holds the half step for this method
    for (int b = 0; b < B; ++b) {
        const double wk = 2.0 * S.omega_n * std::sin(b * pi / B);
        const double gam = (b == 0) ? S.cfg.nrpmd_gamma : 2.0 * wk;
        const double c1 = std::exp(-gam * dt), c2 = std::sqrt(1.0 - c1 * c1);
        for (int d = 0; d < D; ++d) {
            const double sigma = std::sqrt(S.omega_n / S.masses[d]);      // omega_n = nbeads kT
            vtmp[d + (size_t)D * b] = c1 * vtmp[d + (size_t)D * b] + c2 * sigma * T.noise[d + (size_t)D * b];
        }
    }
    step_C(S, vtmp, r);
    from_normal_modes(S, vtmp); from_normal_modes(S, r);
    update_all_caches(S, T, r);
    acceleration(S, T, r, none);
    for (size_t i = 0; i < N; ++i) T.v[i] = std::fma(dt / 2, T.k[i], vtmp[i]);
    T.r = r;
}

// ---- one step: NRPMD (RingPolymerMInt) ringpolymer_mint.jl:28-130 -----------------------------
inline void step_nrpmd(const Setup& S, Trajectory& T) {
    const int n = S.n, D = S.D, B = S.B;
    const double dt = S.cfg.dt;
    if (B > 1) {
        to_normal_modes(S, T.r); to_normal_modes(S, T.v);
        step_C(S, T.v, T.r);  // half-step Cayley (built with half=true, :22)
        from_normal_modes(S, T.r); from_normal_modes(S, T.v);
    } else {
        step_C(S, T.v, T.r);
    }
    for (int b = 0; b < B; ++b) T.bead[b].update(S.model, &T.r[(size_t)D * b], false);
    // propagate_mapping_variables! :80-94
    vec Vbar(B);
    for (int b = 0; b < B; ++b) {
        const Cache& c = T.bead[b];
        double tr = 0.0;
        for (int i = 0; i < n; ++i) tr += c.V[i + (size_t)n * i];
        Vbar[b] = tr / n;
        vec lam(n), Cm((size_t)n * n, 0.0), Dm((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i) lam[i] = c.w[i] - Vbar[b];
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                double sc = 0.0, ss = 0.0;
                for (int k = 0; k < n; ++k) {
                    double zz = c.Z[i + (size_t)n * k] * c.Z[j + (size_t)n * k];
                    sc += zz * std::cos(lam[k] * dt);
                    ss += zz * std::sin(-lam[k] * dt);
                }
                Cm[i + (size_t)n * j] = sc; Dm[i + (size_t)n * j] = ss;
            }
        vec q(n), p(n);
        double* qm = &T.qmap[(size_t)n * b];
        double* pm = &T.pmap[(size_t)n * b];
        for (int i = 0; i < n; ++i) {
            double sq = 0.0, sp = 0.0;
            for (int j = 0; j < n; ++j) {
                sq += Cm[i + (size_t)n * j] * qm[j] - Dm[i + (size_t)n * j] * pm[j];
                sp += Cm[i + (size_t)n * j] * pm[j] + Dm[i + (size_t)n * j] * qm[j];
            }
            q[i] = sq; p[i] = sp;
        }
        for (int i = 0; i < n; ++i) { qm[i] = q[i]; pm[i] = p[i]; }
    }
    // nuclear kick :52-70
    for (int b = 0; b < B; ++b) {
        const Cache& c = T.bead[b];
        vec lam(n);
        for (int i = 0; i < n; ++i) lam[i] = c.w[i] - Vbar[b];
        const double* qm = &T.qmap[(size_t)n * b];
        const double* pm = &T.pmap[(size_t)n * b];
        for (int I = 0; I < D; ++I) {
            // traceless adiabatic derivative W = Z' (dV - Dbar I) Z = adiab - Dbar I
            const double* dv = &c.dV[(size_t)I * n * n];
            double Dbar = 0.0;
            for (int i = 0; i < n; ++i) Dbar += dv[i + (size_t)n * i];
            Dbar /= n;
            vec W(&c.adiab[(size_t)I * n * n], &c.adiab[(size_t)I * n * n] + (size_t)n * n);
            for (int i = 0; i < n; ++i) W[i + (size_t)n * i] -= Dbar;
            // get_gamma / get_xi :107-121.  SMatrix{n,n}(f(i,j) for j=1:n, i=1:n) fills column-major
            // with j fastest: element [row=j, col=i] = f(i,j).
            vec G((size_t)n * n), X((size_t)n * n);
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    double gi, xi_;
                    if (i != j) {
                        double dl = lam[i] - lam[j];
                        gi = std::sin(dl * dt) * W[i + (size_t)n * j] / dl;
                        xi_ = (1.0 - std::cos(dl * dt)) * W[i + (size_t)n * j] / dl;
                    } else { gi = W[i + (size_t)n * j] * dt; xi_ = 0.0; }
                    G[j + (size_t)n * i] = gi; X[j + (size_t)n * i] = xi_;
                }
            // E = Z G Z', F = Z X Z' ; force = 0.5 (q'Eq + p'Ep) - q'F p
            auto transform = [&](const vec& M, vec& out) {
                vec tmp((size_t)n * n);
                for (int j = 0; j < n; ++j)
                    for (int i = 0; i < n; ++i) {
                        double s = 0.0;
                        for (int k = 0; k < n; ++k) s += c.Z[i + (size_t)n * k] * M[k + (size_t)n * j];
                        tmp[i + (size_t)n * j] = s;
                    }
                for (int j = 0; j < n; ++j)
                    for (int i = 0; i < n; ++i) {
                        double s = 0.0;
                        for (int k = 0; k < n; ++k) s += tmp[i + (size_t)n * k] * c.Z[j + (size_t)n * k];
                        out[i + (size_t)n * j] = s;
                    }
            };
            vec E((size_t)n * n), F((size_t)n * n);
            transform(G, E); transform(X, F);
            double qEq = 0.0, pEp = 0.0, qFp = 0.0;
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    qEq += qm[i] * E[i + (size_t)n * j] * qm[j];
                    pEp += pm[i] * E[i + (size_t)n * j] * pm[j];
                    qFp += qm[i] * F[i + (size_t)n * j] * pm[j];
                }
            double force = 0.5 * (qEq + pEp) - qFp;
            size_t idx = I + (size_t)D * b;
            T.v[idx] -= force / S.masses[I];
            T.v[idx] -= Dbar / S.masses[I] * dt;
        }
    }
    if (B > 1) {
        to_normal_modes(S, T.r); to_normal_modes(S, T.v);
        step_C(S, T.v, T.r);
        from_normal_modes(S, T.r); from_normal_modes(S, T.v);
    } else {
        step_C(S, T.v, T.r);
    }
}

// ---- IESH -------------------------------------------------------------------------------------
// get_quantum_propagator / propagate_wavefunction! wavefunction_dynamics.jl:15-58
// propagate_wavefunction!(sigma_final, sigma, v, r, sim, dt), wavefunction_dynamics.jl:15-58: c = the cache evaluated at r
// (ring polymer: centroid), vh = get_hopping_velocity(sim, v)
inline void iesh_propagate_wavefunction(const Setup& S, Trajectory& T, const Cache& c, const double* vh) {
    const int n = S.n, ne = S.ne;
    cvec H((size_t)n * n, cd(0.0));
    for (int i = 0; i < n; ++i) H[i + (size_t)n * i] = c.w[i];
    for (int I = 0; I < S.D; ++I)
        for (int J = 0; J < n * n; ++J) H[J] -= cd(0.0, 1.0) * c.nac[(size_t)I * n * n + J] * vh[I];
    // tmp1 .= Hermitian(prop): upper triangle defines the matrix
    for (int j = 0; j < n; ++j)
        for (int i = j + 1; i < n; ++i) H[i + (size_t)n * j] = std::conj(H[j + (size_t)n * i]);
    vec lam(n);
    cvec vecs((size_t)n * n);
    jacobi_heig(n, H.data(), lam.data(), vecs.data());
    // U = vecs * diag(exp(-i lam dt)) * vecs'
    cvec U((size_t)n * n, cd(0.0));
    for (int k = 0; k < n; ++k) {
        cd ph = std::exp(cd(0.0, -lam[k] * S.cfg.dt));
        for (int j = 0; j < n; ++j) {
            cd f = ph * std::conj(vecs[j + (size_t)n * k]);
            for (int i = 0; i < n; ++i) U[i + (size_t)n * j] += vecs[i + (size_t)n * k] * f;
        }
    }
    cvec out((size_t)n * ne, cd(0.0));
    for (int e = 0; e < ne; ++e)
        for (int k = 0; k < n; ++k) {
            cd x = T.sigma[k + (size_t)n * e];
            for (int i = 0; i < n; ++i) out[i + (size_t)n * e] += U[i + (size_t)n * k] * x;
        }
    T.sigma = out;
}

inline void iesh_unoccupied(int n, const std::vector<int>& occ, std::vector<int>& un) {  // DynamicsUtils.jl:162-171
    un.clear();
    for (int i = 0; i < n; ++i)
        if (std::find(occ.begin(), occ.end(), i) == occ.end()) un.push_back(i);
}

// iesh_check_hop! / evaluate_hopping_probability! / select_new_state / iesh_execute_hop!
// iesh.jl:231-335,390-407 ; Q6: one draw for both pruning and selection.
inline void iesh_hop(const Setup& S, Trajectory& T, double xi) {
    if (S.cfg.disable_hopping) return;
    const int n = S.n, ne = S.ne, D = S.D;
    const Cache& c = hop_cache(S, T);          // ring polymer: centroid (SurfaceHoppingMethods.jl:85-103)
    vec vh(D);
    hop_velocity(S, T, vh.data());
    std::vector<int> un;
    iesh_unoccupied(n, T.occ, un);
    cvec Sm((size_t)ne * ne);
    auto overlap = [&](const std::vector<int>& st) {
        for (int i = 0; i < ne; ++i)
            for (int j = 0; j < ne; ++j) Sm[j + (size_t)ne * i] = T.sigma[st[j] + (size_t)n * i];
    };
    overlap(T.occ);
    cd det_current = complex_det(ne, Sm.data());
    double Akk = std::norm(det_current);
    double prefactor = 2.0 * S.cfg.dt / Akk;
    // evaluate_v_dot_d! :285-298   v_dot_d[m,n] -= v[I] * d[I][m, state[n]]
    vec vdd((size_t)n * ne, 0.0);
    for (int I = 0; I < D; ++I)
        for (int e = 0; e < ne; ++e)
            for (int m : un) vdd[m + (size_t)n * e] -= vh[I] * c.nac[(size_t)I * n * n + m + (size_t)n * T.occ[e]];
    vec prob((size_t)n * ne, 0.0);
    bool pruned = false;
    if (S.cfg.estimate_probability) {
        double sabs = 0.0;
        for (double x : vdd) sabs += std::fabs(x);
        double estimate = prefactor * sabs * (std::fabs(det_current.real()) + std::fabs(det_current.imag()));
        if (estimate < xi) pruned = true;
    }
    if (!pruned) {
        T.cnt.hop_searches++;
        std::vector<int> prop(ne);
        for (int e = 0; e < ne; ++e)
            for (int m : un) {
                prop = T.occ; prop[e] = m;
                overlap(prop);
                cd det_new = complex_det(ne, Sm.data());
                cd Akj = det_current * std::conj(det_new);
                double pr = prefactor * Akj.real() * vdd[m + (size_t)n * e];
                prob[m + (size_t)n * e] = std::min(1.0, std::max(0.0, pr));
            }
    }
    // select_new_state :318-335
    double cumulative = 0.0;
    int he = -1, hm = -1;
    for (int e = 0; e < ne && he < 0; ++e)
        for (int m : un) {
            cumulative += prob[m + (size_t)n * e];
            if (xi < cumulative) { he = e; hm = m; break; }
        }
    if (he < 0) return;
    // rescale_velocity! with (new_state, old_state) = symdiff(new, old): the differing orbitals
    int new_state = hm, old_state = T.occ[he];
    bool accept = true;
    if (S.cfg.rescaling != NQCB200_RESCALE_OFF) {
        vec d(D);
        for (int I = 0; I < D; ++I) d[I] = c.nac[(size_t)I * n * n + new_state + (size_t)n * old_state];
        double a = 0.0, b = 0.0;
        for (int I = 0; I < D; ++I) { a += d[I] * d[I] / S.masses[I]; b += d[I] * vh[I]; }
        a /= 2.0;
        double cc = c.w[new_state] - c.w[old_state];
        double disc = b * b - 4.0 * a * cc;
        if (disc < 0.0) {
            accept = false;
            T.cnt.frustrated++;
            if (S.cfg.rescaling == NQCB200_RESCALE_VINVERSION) {
                double nrm = 0.0;
                for (int I = 0; I < D; ++I) nrm += d[I] * d[I];
                nrm = std::sqrt(nrm);
                double gam = 0.0;
                for (int I = 0; I < D; ++I) gam += vh[I] * d[I] / nrm;
                for (int b2 = 0; b2 < S.B; ++b2)      // every bead (rpsh.jl:39-50)
                    for (int I = 0; I < D; ++I) T.v[I + (size_t)D * b2] -= 2.0 * gam * d[I] / nrm;
            }
        } else {
            double root = std::sqrt(disc);
            double gam = (b < 0.0) ? (b + root) / (2.0 * a) : (b - root) / (2.0 * a);
            for (int b2 = 0; b2 < S.B; ++b2)          // rpsh.jl:30-37
                for (int I = 0; I < D; ++I) T.v[I + (size_t)D * b2] -= gam * d[I] / S.masses[I];
        }
    }
    if (accept) {
        // set_state!(..., new_state::Vector): the proposed vector is copied verbatim -- it is NOT
        // re-sorted (iesh.jl:399-407, surface_hopping.jl:23-28)
        T.occ[he] = hm;
        T.cnt.hops++;
    }
}

// EDC decoherence, decoherence_corrections.jl:21-38 via iesh.jl:430-441
inline void iesh_edc(const Setup& S, Trajectory& T) {
    const int n = S.n, ne = S.ne;
    const Cache& c = T.bead[0];
    double Ekin = 0.0;
    for (int I = 0; I < S.D; ++I) Ekin += S.masses[I] * T.v[I] * T.v[I];
    Ekin /= 2.0;
    for (int e = 0; e < ne; ++e) {
        int occ = T.occ[e];
        double un_norm = 0.0;
        for (int i = 0; i < n; ++i) {
            if (i == occ) continue;
            double tau = (1.0 + S.cfg.edc_C / Ekin) / std::fabs(c.w[i] - c.w[occ]);
            T.sigma[i + (size_t)n * e] *= std::exp(-S.cfg.dt / tau);
            un_norm += std::norm(T.sigma[i + (size_t)n * e]);
        }
        cd Cm = T.sigma[occ + (size_t)n * e];
        T.sigma[occ + (size_t)n * e] = Cm * std::sqrt((1.0 - un_norm) / std::norm(Cm));
    }
}

inline void step_iesh(const Setup& S, Trajectory& T, double xi) {  // verlet_with_electronics.jl:42-69
    const double dt = S.cfg.dt;
    const int D = S.D;
    cvec none;
    vec vtmp(D);
    for (int i = 0; i < D; ++i) vtmp[i] = std::fma(dt / 2, T.k[i], T.v[i]);
    for (int i = 0; i < D; ++i) T.r[i] = std::fma(dt, vtmp[i], T.r[i]);
    T.bead[0].update(S.model, T.r.data());
    acceleration(S, T, T.r, none);
    for (int i = 0; i < D; ++i) T.v[i] = std::fma(dt / 2, T.k[i], vtmp[i]);
    iesh_propagate_wavefunction(S, T, T.bead[0], T.v.data());   // uses (vfinal, rfinal): Q5
    if (S.cfg.method == NQCB200_METHOD_EHRENFEST_NA) return;   // no callback (ehrenfest_na.jl has none)
    iesh_hop(S, T, xi);
    if (S.cfg.edc_C > 0.0) iesh_edc(S, T);
}

// RingPolymerSimulation{AdiabaticIESH} / {EhrenfestNA}: BCBWavefunction perform_step!, bcb_wavefunction.jl:37-69.
// Q5: psi is propagated with (vprev, rprev) -- get_hopping_eigenvalues / get_hopping_nonadiabatic_coupling evaluate the
// centroid at the position they are handed (NQCCalculators' position-keyed getters, external) -- i.e. with the centroid
// cache and centroid velocity from BEFORE this step's nuclear update; the hop callback then sees the new geometry.
inline void step_rpiesh(const Setup& S, Trajectory& T, double xi) {
    const double dt = S.cfg.dt;
    const size_t nd = (size_t)S.B * S.D;
    const Cache cprev = T.centroid;
    vec vprev(S.D);
    centroid_of(S, T.v, vprev.data());
    vec vtmp(nd);
    for (size_t i = 0; i < nd; ++i) vtmp[i] = std::fma(dt / 2, T.k[i], T.v[i]);     // step_B!
    to_normal_modes(S, T.r); to_normal_modes(S, vtmp);
    step_C(S, vtmp, T.r);
    from_normal_modes(S, T.r); from_normal_modes(S, vtmp);
    update_all_caches(S, T, T.r);
    acceleration(S, T, T.r, T.sigma);                                                // sigma_prev / method.state
    for (size_t i = 0; i < nd; ++i) T.v[i] = std::fma(dt / 2, T.k[i], vtmp[i]);
    iesh_propagate_wavefunction(S, T, cprev, vprev.data());
    if (S.cfg.method == NQCB200_METHOD_EHRENFEST_NA) return;
    iesh_hop(S, T, xi);
}

inline void step(const Setup& S, Trajectory& T, double xi) {
    switch (S.cfg.method) {
        case NQCB200_METHOD_FSSH:
        case NQCB200_METHOD_EHRENFEST: step_density_method(S, T, xi); break;
        case NQCB200_METHOD_CLASSICAL: step_classical(S, T); break;
        case NQCB200_METHOD_THERMAL_LANGEVIN: step_langevin_bcocb(S, T); break;
        case NQCB200_METHOD_NRPMD: step_nrpmd(S, T); break;
        case NQCB200_METHOD_IESH:
        case NQCB200_METHOD_EHRENFEST_NA: if (S.B > 1) step_rpiesh(S, T, xi); else step_iesh(S, T, xi); break;
        default: throw std::runtime_error("step: method");
    }
    T.step++;
    T.cnt.steps++;
}

// ---- initialisation (alg_cache + initialize!) -------------------------------------------------
inline void initialise(const Setup& S, Trajectory& T, const double* Zref) {
    const int n = S.n, D = S.D, B = S.B;
    T.bead.assign(B, Cache());
    for (int b = 0; b < B; ++b) T.bead[b].init(n, D, Zref ? Zref + (size_t)n * n * b : nullptr);
    if (B > 1) T.centroid.init(n, D, Zref ? Zref + (size_t)n * n * B : nullptr);
    T.k.assign((size_t)B * D, 0.0);
    T.step = 0;
    T.term_step = -1;
    T.cnt = Counters();
    // Q1: both halves of the double buffer all-zero with t = 0.0 (tspan[1] of a (0, dt) problem)
    T.cur.vd.assign((size_t)n * n, cd(0.0)); T.cur.E.assign(n, cd(0.0)); T.cur.t = 0.0;
    T.nxt = T.cur;
    if (S.cfg.method == NQCB200_METHOD_NRPMD) return;  // initialize! is empty (ringpolymer_mint.jl:26)
    update_all_caches(S, T, T.r);
}
inline void initial_acceleration(const Setup& S, Trajectory& T) {
    if (S.cfg.method == NQCB200_METHOD_NRPMD) return;
    acceleration(S, T, T.r, T.sigma);
}

// ---- observables (Estimators / DynamicsOutputs) -----------------------------------------------
inline void adiabatic_population(const Setup& S, const Trajectory& T, double* pop) {
    const int n = S.n;
    std::fill(pop, pop + n, 0.0);
    switch (S.cfg.method) {
        case NQCB200_METHOD_FSSH: pop[T.state] = 1.0; break;                                   // fssh.jl:144-148
        case NQCB200_METHOD_EHRENFEST: for (int i = 0; i < n; ++i) pop[i] = T.sigma[i + (size_t)n * i].real(); break;
        case NQCB200_METHOD_IESH: for (int kk : T.occ) pop[kk] = 1.0; break;                   // iesh.jl:371-375
        case NQCB200_METHOD_EHRENFEST_NA:                                                      // ehrenfest_na.jl:116-125
            for (int e = 0; e < S.ne; ++e)
                for (int m = 0; m < n; ++m) pop[m] += std::norm(T.sigma[m + (size_t)n * e]);
            break;
        default: break;
    }
}
inline void diabatic_population(const Setup& S, const Trajectory& T, double* pop) {
    const int n = S.n;
    std::fill(pop, pop + n, 0.0);
    if (S.cfg.method == NQCB200_METHOD_NRPMD) {  // nrpmd.jl:111-122
        for (int b = 0; b < S.B; ++b)
            for (int j = 0; j < n; ++j) {
                double q = T.qmap[j + (size_t)n * b], p = T.pmap[j + (size_t)n * b];
                pop[j] += (q * q + p * p) / 2.0 - S.cfg.nrpmd_gamma;
            }
        for (int j = 0; j < n; ++j) pop[j] /= S.B;
        return;
    }
    const Cache& c = hop_cache(S, T);
    const vec& U = c.Z;
    if (S.cfg.method == NQCB200_METHOD_IESH) {  // iesh.jl:337-369
        for (int e = 0; e < S.ne; ++e) {
            vec rho((size_t)n * n);
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i)
                    rho[i + (size_t)n * j] = T.sigma[i + (size_t)n * e].real() * T.sigma[j + (size_t)n * e].real();
            for (int i = 0; i < n; ++i) rho[i + (size_t)n * i] = 0.0;
            rho[T.occ[e] + (size_t)n * T.occ[e]] = 1.0;
            for (int i = 0; i < n; ++i) {
                double s = 0.0;
                for (int a = 0; a < n; ++a)
                    for (int b = 0; b < n; ++b) s += U[i + (size_t)n * a] * rho[a + (size_t)n * b] * U[i + (size_t)n * b];
                pop[i] += s;
            }
        }
        return;
    }
    vec rho((size_t)n * n);
    for (int i = 0; i < n * n; ++i) rho[i] = T.sigma[i].real();
    if (S.cfg.method == NQCB200_METHOD_FSSH) {  // fssh.jl:132-142
        for (int i = 0; i < n; ++i) rho[i + (size_t)n * i] = 0.0;
        rho[T.state + (size_t)n * T.state] = 1.0;
    }
    for (int i = 0; i < n; ++i) {
        double s = 0.0;
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < n; ++b) s += U[i + (size_t)n * a] * rho[a + (size_t)n * b] * U[i + (size_t)n * b];
        pop[i] = s;
    }
}
inline double kinetic_energy(const Setup& S, const Trajectory& T) {  // DynamicsUtils.jl:108-135
    double kin = 0.0;
    for (int b = 0; b < S.B; ++b)
        for (int I = 0; I < S.D; ++I) kin += S.masses[I] * T.v[I + (size_t)S.D * b] * T.v[I + (size_t)S.D * b];
    return kin / 2.0;
}
inline double potential_energy(const Setup& S, const Trajectory& T) {
    const int n = S.n;
    double pot = 0.0;
    switch (S.cfg.method) {
        case NQCB200_METHOD_FSSH:  // fssh.jl:150-154, rpsh.jl:52-57
            for (int b = 0; b < S.B; ++b) pot += T.bead[b].w[T.state];
            break;
        case NQCB200_METHOD_EHRENFEST:  // ehrenfest.jl:85-95, ehrenfest_rpmd.jl:45-51
            for (int b = 0; b < S.B; ++b) {
                if (S.B == 1) pot += S.model.U0(&T.r[0]);
                for (int i = 0; i < n; ++i) pot += T.sigma[i + (size_t)n * i].real() * T.bead[b].w[i];
            }
            break;
        case NQCB200_METHOD_IESH:  // iesh.jl:380-388, rpiesh.jl:38-52
            for (int b = 0; b < S.B; ++b) {
                pot += S.model.U0(&T.r[(size_t)S.D * b]);
                for (int kk : T.occ) pot += T.bead[b].w[kk];
            }
            break;
        case NQCB200_METHOD_EHRENFEST_NA:  // ehrenfest_na.jl:103-114, rpehrenfest_na.jl:37-52
            for (int b = 0; b < S.B; ++b) {
                pot += S.model.U0(&T.r[(size_t)S.D * b]);
                for (int e = 0; e < S.ne; ++e)
                    for (int i = 0; i < n; ++i) pot += T.bead[b].w[i] * std::norm(T.sigma[i + (size_t)n * e]);
            }
            break;
        case NQCB200_METHOD_THERMAL_LANGEVIN:
        case NQCB200_METHOD_CLASSICAL:  // DynamicsUtils.jl:141-151
            for (int b = 0; b < S.B; ++b) pot += T.bead[b].V[0];
            break;
        case NQCB200_METHOD_NRPMD: {  // nrpmd.jl:124-139 (potential re-evaluated at r)
            vec V((size_t)n * n);
            for (int b = 0; b < S.B; ++b) {
                S.model.potential(&T.r[(size_t)S.D * b], V.data());
                double vbar = 0.0;
                for (int i = 0; i < n; ++i) vbar += V[i + (size_t)n * i];
                vbar /= n;
                const double* q = &T.qmap[(size_t)n * b];
                const double* p = &T.pmap[(size_t)n * b];
                double s = 0.0;
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < n; ++j) {
                        double vt = V[i + (size_t)n * j] - (i == j ? vbar : 0.0);
                        s += p[i] * vt * p[j] + q[i] * vt * q[j];
                    }
                pot += 0.5 * s + vbar;
            }
        } break;
    }
    return pot;
}
inline double spring_energy(const Setup& S, const Trajectory& T) {  // ring_polymer.jl:89-107
    if (S.B == 1) return 0.0;
    double E = 0.0;
    for (int I = 0; I < S.D; ++I) {
        double d = T.r[I + (size_t)S.D * (S.B - 1)] - T.r[I];
        E += S.masses[I] * d * d;
    }
    for (int b = 0; b < S.B - 1; ++b)
        for (int I = 0; I < S.D; ++I) {
            double d = T.r[I + (size_t)S.D * b] - T.r[I + (size_t)S.D * (b + 1)];
            E += S.masses[I] * d * d;
        }
    return E * S.omega_n * S.omega_n / 2.0;
}

}  // namespace nqco

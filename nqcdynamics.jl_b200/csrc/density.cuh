// density.cuh -- electronic density-matrix propagation in registers.
//
// Reference: DensityMatrixODEProblem (src/DynamicsMethods/electronic_dynamics.jl:104-130),
//   dsigma/dt = -i [A(t), sigma],  A(t) = diag(lerp(E_cur, E_next)) - i lerp(vd_cur, vd_next)
// integrated with OrdinaryDiffEq Tsit5 at fixed dt/5 (bab_electronics.jl:40-43,88-89): 5 sub-steps,
// 31 RHS evaluations per nuclear step (FSAL reused inside the step).
//
// B200 restatement.  sigma is Hermitian and vd = sum_I d_I v_I is real antisymmetric, so with
// sigma = X + iY (X symmetric, Y antisymmetric, both real) and G = vd, dE_jk = E_j - E_k:
//     dX/dt =  dE o Y - [G, X]        dY/dt = -dE o X - [G, Y]
// Only the packed upper triangles (n(n+1)/2 + n(n-1)/2 = n^2 doubles) are carried, every index is a
// compile-time constant after unrolling, and the whole Tsit5 stage set lives in registers
// (n = 2: 4 state variables x 6 stages).  This executes ~8x fewer flops than the reference's dense
// complex commutator (16 n^3 + 10 n^2 per RHS) for the same result up to rounding.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace nq {

template <int N>
struct Herm {
    double x[sym_size(N)];                     // Re sigma, packed upper
    double y[asym_size(N) > 0 ? asym_size(N) : 1];  // Im sigma, packed strict upper
    NQ_HD double X(int j, int k) const { return (j <= k) ? x[sidx(N, j, k)] : x[sidx(N, k, j)]; }
    NQ_HD double Y(int j, int k) const { return (j == k) ? 0.0 : ((j < k) ? y[aidx(N, j, k)] : -y[aidx(N, k, j)]); }
    // X(j, k) for RUN-TIME indices as a chain of selects over compile-time indices: a dynamically indexed x[] would
    // push the whole density matrix into local memory (21 local loads + 17 stores per trajectory-step in the
    // TullyModelOne kernel, profiles/r02/SUMMARY.md)
    NQ_HD double Xsel(int j, int k) const {
        double v = 0.0;
#pragma unroll
        for (int a = 0; a < N; ++a)
#pragma unroll
            for (int b = a; b < N; ++b) v = ((a == j && b == k) || (a == k && b == j)) ? x[sidx(N, a, b)] : v;
        return v;
    }
};

// One half of the reference's DoubleBuffer (electronic_dynamics.jl:15-36): eigenvalues and the
// strict upper triangle of the antisymmetric dynamical coupling vd.
template <int N>
struct ElecParams {
    double E[N];
    double g[asym_size(N) > 0 ? asym_size(N) : 1];
    NQ_HD double G(int j, int k) const { return (j == k) ? 0.0 : ((j < k) ? g[aidx(N, j, k)] : -g[aidx(N, k, j)]); }
    NQ_HD double Gsel(int j, int k) const {      // run-time indices, see Herm::Xsel
        double v = 0.0;
#pragma unroll
        for (int a = 0; a < N; ++a)
#pragma unroll
            for (int b = a + 1; b < N; ++b) v = (a == j && b == k) ? g[aidx(N, a, b)] : ((a == k && b == j) ? -g[aidx(N, a, b)] : v);
        return v;
    }
};

template <int N>
NQ_HD void density_rhs(const ElecParams<N>& cur, const ElecParams<N>& nxt, double loc, const Herm<N>& u, Herm<N>& du) {
    ElecParams<N> a;
#pragma unroll
    for (int i = 0; i < N; ++i) a.E[i] = cur.E[i] + (nxt.E[i] - cur.E[i]) * loc;
#pragma unroll
    for (int i = 0; i < asym_size(N); ++i) a.g[i] = cur.g[i] + (nxt.g[i] - cur.g[i]) * loc;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = j; k < N; ++k) {
            // [G,X]_jk = sum_l G_jl X_lk - X_jl G_lk
            double cx = 0.0;
#pragma unroll
            for (int l = 0; l < N; ++l) cx += a.G(j, l) * u.X(l, k) - u.X(j, l) * a.G(l, k);
            du.x[sidx(N, j, k)] = (a.E[j] - a.E[k]) * u.Y(j, k) - cx;
            if (k > j) {
                double cy = 0.0;
#pragma unroll
                for (int l = 0; l < N; ++l) cy += a.G(j, l) * u.Y(l, k) - u.Y(j, l) * a.G(l, k);
                du.y[aidx(N, j, k)] = -(a.E[j] - a.E[k]) * u.X(j, k) - cy;
            }
        }
}

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

// The same coefficients in the constant bank: an FP64 immediate cannot be encoded in a DFMA, so compile-time
// constants cost a UMOV pair per use (9 % of the issued instructions of the SpinBoson kernel, profiles/r01); a
// constant-bank operand is free.
#if defined(__CUDACC__)
namespace tsit5 {
enum { I_c1, I_c2, I_c3, I_c4, I_a21, I_a31, I_a32, I_a41, I_a42, I_a43, I_a51, I_a52, I_a53, I_a54, I_a61, I_a62, I_a63, I_a64, I_a65, I_a71, I_a72, I_a73, I_a74, I_a75, I_a76, I_COUNT };
}
static __constant__ double kTsit5Dev[tsit5::I_COUNT] = {tsit5::c1, tsit5::c2, tsit5::c3, tsit5::c4, tsit5::a21, tsit5::a31, tsit5::a32, tsit5::a41, tsit5::a42, tsit5::a43, tsit5::a51, tsit5::a52, tsit5::a53, tsit5::a54, tsit5::a61, tsit5::a62, tsit5::a63, tsit5::a64, tsit5::a65, tsit5::a71, tsit5::a72, tsit5::a73, tsit5::a74, tsit5::a75, tsit5::a76};
#endif
#if defined(__CUDA_ARCH__)
#define NQ_TS(name) kTsit5Dev[tsit5::I_##name]
#define NQ_RCP(x) rcp_nb(x)      // one-ulp reciprocal without the IEEE slow path (common.cuh)
#else
#define NQ_TS(name) tsit5::name
#define NQ_RCP(x) (1.0 / (x))
#endif

#define NQ_FOR_HERM(expr)                                              \
    _Pragma("unroll") for (int i_ = 0; i_ < sym_size(N); ++i_) { auto& o = tmp.x[i_]; const int i = i_; const bool isx = true; expr; } \
    _Pragma("unroll") for (int i_ = 0; i_ < asym_size(N); ++i_) { auto& o = tmp.y[i_]; const int i = i_; const bool isx = false; expr; }

// sigma(t) -> sigma(t+dt): set_ut! + step!(integrator, dt, true) of the reference.
// tcur / tnext are the time stamps of the two buffer halves (quirk Q1: tcur = 0 and cur = 0 on the
// very first step of a trajectory).
template <int N>
NQ_HD void propagate_density(const ElecParams<N>& cur, double tcur, const ElecParams<N>& nxt, double tnext,
                             double t, double dt, Herm<N>& s, const double (&ha)[21]) {
    using namespace tsit5;
    const double h = dt / 5.0;
    const double inv_span = NQ_RCP(tnext - tcur);
    auto loc_of = [&](double tau) {
        double l = (tau - tcur) * inv_span;
        return (l != l) ? 0.0 : l;   // isnan -> 0 (electronic_dynamics.jl:62,75)
    };
    Herm<N> k1, k2, k3, k4, k5, k6, tmp;
    double ts = t;
    density_rhs<N>(cur, nxt, loc_of(ts), s, k1);
#pragma unroll 1
    for (int sub = 0; sub < 5; ++sub) {
        const double hh = (sub == 4) ? (t + dt) - ts : h;   // tstop snapping of the last sub-step
        // stage argument = s + sum_j (h a_ij) k_j with h a_ij = ha[] (kernel-parameter constant bank), FMAs only
#define NQ_STAGE(EXPRX, EXPRY)                                                              \
        _Pragma("unroll") for (int i = 0; i < sym_size(N); ++i) tmp.x[i] = (EXPRX);  \
        _Pragma("unroll") for (int i = 0; i < asym_size(N); ++i) tmp.y[i] = (EXPRY);
#define NQ_F1(A, K, REST) fma(ha[A], K, REST)
        NQ_STAGE(NQ_F1(0, k1.x[i], s.x[i]), NQ_F1(0, k1.y[i], s.y[i]))
        density_rhs<N>(cur, nxt, loc_of(ts + NQ_TS(c1) * hh), tmp, k2);
        NQ_STAGE(NQ_F1(2, k2.x[i], NQ_F1(1, k1.x[i], s.x[i])), NQ_F1(2, k2.y[i], NQ_F1(1, k1.y[i], s.y[i])))
        density_rhs<N>(cur, nxt, loc_of(ts + NQ_TS(c2) * hh), tmp, k3);
        NQ_STAGE(NQ_F1(5, k3.x[i], NQ_F1(4, k2.x[i], NQ_F1(3, k1.x[i], s.x[i]))),
                 NQ_F1(5, k3.y[i], NQ_F1(4, k2.y[i], NQ_F1(3, k1.y[i], s.y[i]))))
        density_rhs<N>(cur, nxt, loc_of(ts + NQ_TS(c3) * hh), tmp, k4);
        NQ_STAGE(NQ_F1(9, k4.x[i], NQ_F1(8, k3.x[i], NQ_F1(7, k2.x[i], NQ_F1(6, k1.x[i], s.x[i])))),
                 NQ_F1(9, k4.y[i], NQ_F1(8, k3.y[i], NQ_F1(7, k2.y[i], NQ_F1(6, k1.y[i], s.y[i])))))
        density_rhs<N>(cur, nxt, loc_of(ts + NQ_TS(c4) * hh), tmp, k5);
        NQ_STAGE(NQ_F1(14, k5.x[i], NQ_F1(13, k4.x[i], NQ_F1(12, k3.x[i], NQ_F1(11, k2.x[i], NQ_F1(10, k1.x[i], s.x[i]))))),
                 NQ_F1(14, k5.y[i], NQ_F1(13, k4.y[i], NQ_F1(12, k3.y[i], NQ_F1(11, k2.y[i], NQ_F1(10, k1.y[i], s.y[i]))))))
        density_rhs<N>(cur, nxt, loc_of(ts + hh), tmp, k6);
#pragma unroll
        for (int i = 0; i < sym_size(N); ++i)
            s.x[i] = NQ_F1(20, k6.x[i], NQ_F1(19, k5.x[i], NQ_F1(18, k4.x[i], NQ_F1(17, k3.x[i], NQ_F1(16, k2.x[i], NQ_F1(15, k1.x[i], s.x[i]))))));
#pragma unroll
        for (int i = 0; i < asym_size(N); ++i)
            s.y[i] = NQ_F1(20, k6.y[i], NQ_F1(19, k5.y[i], NQ_F1(18, k4.y[i], NQ_F1(17, k3.y[i], NQ_F1(16, k2.y[i], NQ_F1(15, k1.y[i], s.y[i]))))));
#undef NQ_F1
#undef NQ_STAGE
        ts = (sub == 4) ? (t + dt) : ts + hh;
        if (sub < 4) density_rhs<N>(cur, nxt, loc_of(ts), s, k1);  // FSAL: k7 of this sub-step = k1 of the next
    }
}
#undef NQ_FOR_HERM

// Two-state specialisation (every BASELINE config except ThreeStateMorse / IESH): G = g [[0,1],[-1,0]], so
//     dx00 = -2 g x01 = -dx11 ,   dx01 = dE y01 + g (x00 - x11) ,   dy01 = -dE x01 ,
// 10 FP64 instructions per RHS instead of the generic commutator's ~40, and the x11 stage sums are the negated
// x00 ones (bit-identical to carrying them separately).  Same Tsit5 staging as the generic version above.
// `between(sub, std::integral_constant<int, stage>)` is called once per Tsit5 stage (6 per sub-step, 30 per nuclear
// step) right after that stage's RHS has been issued: independent work placed there (the SpinBoson kernel runs the
// next step's mode sweep in these slots) fills the fixed-latency bubbles of the serial stage chain.
template <class Between>
NQ_HD void propagate_density_2state(const ElecParams<2>& cur, double tcur, const ElecParams<2>& nxt, double tnext,
                                    double t, double dt, Herm<2>& s, const double (&ha)[21], Between&& between) {
    // Carried variables: z = x00 - x11 (the trace x00 + x11 is a constant of the motion), x01, y01:
    //     dz = -4 g x01 ,   dx01 = dE y01 + g z ,   dy01 = -dE x01 .
    // FP64 latency, not throughput, bounds the kernels that call this (one serial chain per thread, few warps per
    // scheduler), so every stage is written with the NEWEST slope last: the partial sum over the older slopes is off
    // the critical path and one FMA turns a fresh slope into the next stage's argument
    // (3 dependent operations from d z to d x01 and 2 back, instead of 5 and 3).
    using namespace tsit5;
    const double h = dt / 5.0;
    const double inv_span = NQ_RCP(tnext - tcur);
    const bool flat = !(fabs(inv_span) <= 1.0e300);   // zero span: the reference's isnan(loc) -> 0 branch
    const double dEc = cur.E[0] - cur.E[1], dEd = (nxt.E[0] - nxt.E[1]) - dEc;
    const double gc = cur.g[0], gd = nxt.g[0] - gc;
    struct K { double a, b, c; };   // d z, d x01, d y01
    const double tr = s.x[0] + s.x[2];
    double z = s.x[0] - s.x[2], x01 = s.x[1], y01 = s.y[0];
    // generator at stage time ts + c hh: loc = (ts - tcur) inv_span + c (hh inv_span), so dE and g are ONE FMA each from
    // their values at ts (dE0, g0) and their increments over the sub-step (ddE, ddg) -- the same linear interpolation of
    // electronic_dynamics.jl:55-79 with the division of labour moved out of the 31 RHS evaluations
    double dE0, g0, ddE, ddg;
    auto rhs = [&](double c, double uz, double u01, double w01, K& k) {
        const double dE = fma(c, ddE, dE0), g = fma(c, ddg, g0);
        k.a = -((4.0 * g) * u01);
        k.b = fma(dE, w01, g * uz);
        k.c = -(dE * u01);
    };
    auto at_substep = [&](double ts_, double hh_) {
        const double l0 = flat ? 0.0 : (ts_ - tcur) * inv_span, dl = flat ? 0.0 : hh_ * inv_span;
        dE0 = fma(dEd, l0, dEc); g0 = fma(gd, l0, gc);
        ddE = dEd * dl; ddg = gd * dl;
    };
    K k1, k2, k3, k4, k5, k6;
    double ts = t;
    at_substep(ts, h);
    rhs(0.0, z, x01, y01, k1);
#pragma unroll 1
    for (int sub = 0; sub < 5; ++sub) {
        const double hh = (sub == 4) ? (t + dt) - ts : h;   // tstop snapping of the last sub-step
        at_substep(ts, hh);
        // stage i+1 argument = [x + h sum_{j<i} a_{i+1,j} k_j] + (h a_{i+1,i}) k_i with h a_ij = ha[] from the constant
        // bank (the last sub-step's snapped length differs from h = dt/5 by rounding only: it enters the stage times)
#define NQ_ARG(P, C, KN) fma((C), (KN), (P))
        {
            const double c21 = ha[0];
            rhs(NQ_TS(c1), NQ_ARG(z, c21, k1.a), NQ_ARG(x01, c21, k1.b), NQ_ARG(y01, c21, k1.c), k2);
        }
        between(sub, std::integral_constant<int, 0>{});
        {
            const double c31 = ha[1], c32 = ha[2];
            rhs(NQ_TS(c2), NQ_ARG(fma(c31, k1.a, z), c32, k2.a), NQ_ARG(fma(c31, k1.b, x01), c32, k2.b),
                NQ_ARG(fma(c31, k1.c, y01), c32, k2.c), k3);
        }
        between(sub, std::integral_constant<int, 1>{});
        {
            const double c41 = ha[3], c42 = ha[4], c43 = ha[5];
            rhs(NQ_TS(c3), NQ_ARG(fma(c42, k2.a, fma(c41, k1.a, z)), c43, k3.a),
                NQ_ARG(fma(c42, k2.b, fma(c41, k1.b, x01)), c43, k3.b),
                NQ_ARG(fma(c42, k2.c, fma(c41, k1.c, y01)), c43, k3.c), k4);
        }
        between(sub, std::integral_constant<int, 2>{});
        {
            const double c51 = ha[6], c52 = ha[7], c53 = ha[8], c54 = ha[9];
            rhs(NQ_TS(c4), NQ_ARG(fma(c53, k3.a, fma(c52, k2.a, fma(c51, k1.a, z))), c54, k4.a),
                NQ_ARG(fma(c53, k3.b, fma(c52, k2.b, fma(c51, k1.b, x01))), c54, k4.b),
                NQ_ARG(fma(c53, k3.c, fma(c52, k2.c, fma(c51, k1.c, y01))), c54, k4.c), k5);
        }
        between(sub, std::integral_constant<int, 3>{});
        {
            const double c61 = ha[10], c62 = ha[11], c63 = ha[12], c64 = ha[13],
                         c65 = ha[14];
            rhs(1.0, NQ_ARG(fma(c64, k4.a, fma(c63, k3.a, fma(c62, k2.a, fma(c61, k1.a, z)))), c65, k5.a),
                NQ_ARG(fma(c64, k4.b, fma(c63, k3.b, fma(c62, k2.b, fma(c61, k1.b, x01)))), c65, k5.b),
                NQ_ARG(fma(c64, k4.c, fma(c63, k3.c, fma(c62, k2.c, fma(c61, k1.c, y01)))), c65, k5.c), k6);
        }
        between(sub, std::integral_constant<int, 4>{});
        {
            const double c71 = ha[15], c72 = ha[16], c73 = ha[17], c74 = ha[18],
                         c75 = ha[19], c76 = ha[20];
            z = NQ_ARG(fma(c75, k5.a, fma(c74, k4.a, fma(c73, k3.a, fma(c72, k2.a, fma(c71, k1.a, z))))), c76, k6.a);
            x01 = NQ_ARG(fma(c75, k5.b, fma(c74, k4.b, fma(c73, k3.b, fma(c72, k2.b, fma(c71, k1.b, x01))))), c76, k6.b);
            y01 = NQ_ARG(fma(c75, k5.c, fma(c74, k4.c, fma(c73, k3.c, fma(c72, k2.c, fma(c71, k1.c, y01))))), c76, k6.c);
        }
#undef NQ_ARG
        ts = (sub == 4) ? (t + dt) : ts + hh;
        if (sub < 4) rhs(1.0, z, x01, y01, k1);   // FSAL: the slope at the end of this sub-step = at the start of the next
        between(sub, std::integral_constant<int, 5>{});
    }
    s.x[0] = 0.5 * (tr + z); s.x[1] = x01; s.x[2] = 0.5 * (tr - z); s.y[0] = y01;
}

template <>
NQ_HD void propagate_density<2>(const ElecParams<2>& cur, double tcur, const ElecParams<2>& nxt, double tnext,
                                double t, double dt, Herm<2>& s, const double (&ha)[21]) {
    propagate_density_2state(cur, tcur, nxt, tnext, t, dt, s, ha, [](int, auto) {});
}

}  // namespace nq

// models.cuh -- device model table: diabatic potential V(r) and derivative dV/dr_j as packed
// symmetric matrices.  B200-side restatement of NQCModels.jl `potential!` / `derivative!`
// (external package; reference call sites fssh.jl:44, ehrenfest.jl:52, simulations.jl:75-83;
// formulas docs/src/NQCModels/analyticmodels.md, systembathmodels.md:20-26 and SURVEY.md 8c/A.5).
//
// Interface (all static, inlined into the step kernels; P = KParams::params):
//   NS                      number of electronic states
//   kBath                   true if the model carries per-dof arrays (bath_a/bath_b) and V is a sum
//                           over dofs that needs a cross-lane reduction
//   potential_partial<DPL>  this lane's contribution to V (packed upper, row-wise)
//   derivative_dof          dV/dr_j for ONE local dof
#pragma once
#include "common.cuh"

namespace nq {

// device: the branch-free exponential of common.cuh (several of them overlap in one basic block); host: libm
#if defined(__CUDA_ARCH__)
#define NQ_EXP exp_nb
#else
#define NQ_EXP exp
#endif

template <int KIND> struct ModelT;

template <> struct ModelT<NQCB200_MODEL_TULLY_ONE> {
    static constexpr int NS = 2;
    static constexpr bool kBath = false;
    template <int DPL>
    NQ_HD static void potential_partial(const double* P, const double (&r)[DPL], const double (&)[DPL],
                                        const double (&)[DPL], bool lane0, double (&V)[3]) {
        const double q = r[0];
        const double e = NQ_EXP(-P[1] * fabs(q));
        const double v11 = (q > 0.0) ? P[0] * (1.0 - e) : -P[0] * (1.0 - e);
        V[0] = lane0 ? v11 : 0.0; V[2] = -V[0];
        V[1] = lane0 ? P[2] * NQ_EXP(-P[3] * q * q) : 0.0;
    }
    NQ_HD static void derivative_dof(const double* P, double q, double, double, double (&dV)[3]) {
        const double d11 = P[0] * P[1] * NQ_EXP(-P[1] * fabs(q));
        dV[0] = d11; dV[2] = -d11;
        dV[1] = -2.0 * P[2] * P[3] * q * NQ_EXP(-P[3] * q * q);
    }
};

template <> struct ModelT<NQCB200_MODEL_TULLY_TWO> {
    static constexpr int NS = 2;
    static constexpr bool kBath = false;
    template <int DPL>
    NQ_HD static void potential_partial(const double* P, const double (&r)[DPL], const double (&)[DPL],
                                        const double (&)[DPL], bool lane0, double (&V)[3]) {
        const double q = r[0];
        V[0] = 0.0;
        V[2] = lane0 ? -P[0] * NQ_EXP(-P[1] * q * q) + P[4] : 0.0;
        V[1] = lane0 ? P[2] * NQ_EXP(-P[3] * q * q) : 0.0;
    }
    NQ_HD static void derivative_dof(const double* P, double q, double, double, double (&dV)[3]) {
        dV[0] = 0.0;
        dV[2] = 2.0 * P[0] * P[1] * q * NQ_EXP(-P[1] * q * q);
        dV[1] = -2.0 * P[2] * P[3] * q * NQ_EXP(-P[3] * q * q);
    }
};

template <> struct ModelT<NQCB200_MODEL_TULLY_THREE> {
    static constexpr int NS = 2;
    static constexpr bool kBath = false;
    template <int DPL>
    NQ_HD static void potential_partial(const double* P, const double (&r)[DPL], const double (&)[DPL],
                                        const double (&)[DPL], bool lane0, double (&V)[3]) {
        const double q = r[0];
        const double e = NQ_EXP(-P[2] * fabs(q));
        V[0] = lane0 ? P[0] : 0.0; V[2] = -V[0];
        V[1] = lane0 ? ((q < 0.0) ? P[1] * e : P[1] * (2.0 - e)) : 0.0;
    }
    NQ_HD static void derivative_dof(const double* P, double q, double, double, double (&dV)[3]) {
        dV[0] = 0.0; dV[2] = 0.0;
        dV[1] = P[1] * P[2] * NQ_EXP(-P[2] * fabs(q));
    }
};

template <> struct ModelT<NQCB200_MODEL_DOUBLE_WELL> {
    static constexpr int NS = 2;
    static constexpr bool kBath = false;
    template <int DPL>
    NQ_HD static void potential_partial(const double* P, const double (&r)[DPL], const double (&)[DPL],
                                        const double (&)[DPL], bool lane0, double (&V)[3]) {
        const double q = r[0];
        const double v0 = 0.5 * P[0] * P[1] * P[1] * q * q, vv = 1.4142135623730951 * P[2] * q;
        V[0] = lane0 ? v0 + vv : 0.0; V[2] = lane0 ? v0 - vv : 0.0; V[1] = lane0 ? 0.5 * P[3] : 0.0;
    }
    NQ_HD static void derivative_dof(const double* P, double q, double, double, double (&dV)[3]) {
        const double d0 = P[0] * P[1] * P[1] * q, dv = 1.4142135623730951 * P[2];
        dV[0] = d0 + dv; dV[2] = d0 - dv; dV[1] = 0.0;
    }
};

// V11/22 = sum_j 1/2 w_j^2 r_j^2 +- (eps + sum_j c_j r_j), V12 = Delta.  bath_a = w_j, bath_b = c_j.
template <> struct ModelT<NQCB200_MODEL_SPIN_BOSON> {
    static constexpr int NS = 2;
    static constexpr bool kBath = true;
    template <int DPL>
    NQ_HD static void potential_partial(const double* P, const double (&r)[DPL], const double (&w)[DPL],
                                        const double (&c)[DPL], bool lane0, double (&V)[3]) {
        double harm = 0.0, lin = lane0 ? P[0] : 0.0;
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            harm += 0.5 * w[j] * w[j] * r[j] * r[j];
            lin += c[j] * r[j];
        }
        V[0] = harm + lin; V[2] = harm - lin; V[1] = lane0 ? P[1] : 0.0;
    }
    NQ_HD static void derivative_dof(const double*, double q, double w, double c, double (&dV)[3]) {
        const double h = w * w * q;
        dV[0] = h + c; dV[2] = h - c; dV[1] = 0.0;
    }
};

// V_ii = d_i (1 - exp(-alpha_i (q - r_i)))^2 + c_i ; V_ij = a_ij exp(-alpha_ij (q - r_ij)^2)
// params: d[0:3] alpha[3:6] r[6:9] c[9:12] a[12:15] alphac[15:18] rc[18:21], pairs (12, 13, 23)
template <> struct ModelT<NQCB200_MODEL_THREE_STATE_MORSE> {
    static constexpr int NS = 3;
    static constexpr bool kBath = false;
    template <int DPL>
    NQ_HD static void potential_partial(const double* P, const double (&r)[DPL], const double (&)[DPL],
                                        const double (&)[DPL], bool lane0, double (&V)[6]) {
        const double q = r[0];
        const double s = lane0 ? 1.0 : 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double e = 1.0 - NQ_EXP(-P[3 + i] * (q - P[6 + i]));
            V[sidx(3, i, i)] = s * (P[i] * e * e + P[9 + i]);
        }
        const double d01 = q - P[18], d02 = q - P[19], d12 = q - P[20];
        V[sidx(3, 0, 1)] = s * P[12] * NQ_EXP(-P[15] * d01 * d01);
        V[sidx(3, 0, 2)] = s * P[13] * NQ_EXP(-P[16] * d02 * d02);
        V[sidx(3, 1, 2)] = s * P[14] * NQ_EXP(-P[17] * d12 * d12);
    }
    NQ_HD static void derivative_dof(const double* P, double q, double, double, double (&dV)[6]) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double ex = NQ_EXP(-P[3 + i] * (q - P[6 + i]));
            dV[sidx(3, i, i)] = 2.0 * P[i] * P[3 + i] * ex * (1.0 - ex);
        }
        const double d01 = q - P[18], d02 = q - P[19], d12 = q - P[20];
        dV[sidx(3, 0, 1)] = -2.0 * P[15] * d01 * P[12] * NQ_EXP(-P[15] * d01 * d01);
        dV[sidx(3, 0, 2)] = -2.0 * P[16] * d02 * P[13] * NQ_EXP(-P[16] * d02 * d02);
        dV[sidx(3, 1, 2)] = -2.0 * P[17] * d12 * P[14] * NQ_EXP(-P[17] * d12 * d12);
    }
};

// Classical (single-surface) models: V packed has one entry, derivative is the force gradient.
template <> struct ModelT<NQCB200_MODEL_HARMONIC> {
    static constexpr int NS = 1;
    static constexpr bool kBath = false;
    NQ_HD static double potential_dof(const double* P, double q) { return 0.5 * P[0] * P[1] * P[1] * (q - P[2]) * (q - P[2]); }
    NQ_HD static double gradient_dof(const double* P, double q) { return P[0] * P[1] * P[1] * (q - P[2]); }
};
template <> struct ModelT<NQCB200_MODEL_FREE> {
    static constexpr int NS = 1;
    static constexpr bool kBath = false;
    NQ_HD static double potential_dof(const double*, double) { return 0.0; }
    NQ_HD static double gradient_dof(const double*, double) { return 0.0; }
};

// V(q) and dV/dq of a one-dof model at once; the default evaluates them separately, models whose derivative shares
// its transcendentals with the potential specialise it (ThreeStateMorse: 6 exponentials instead of 12).
template <class M>
NQ_HD void model_value_and_derivative(const double* P, double q, double (&V)[sym_size(M::NS)], double (&dV)[sym_size(M::NS)]) {
    const double rr[1] = {q}, zz[1] = {0.0};
    M::template potential_partial<1>(P, rr, zz, zz, true, V);
    M::derivative_dof(P, q, 0.0, 0.0, dV);
}
template <>
NQ_HD void model_value_and_derivative<ModelT<NQCB200_MODEL_THREE_STATE_MORSE>>(const double* P, double q, double (&V)[6], double (&dV)[6]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double ex = NQ_EXP(-P[3 + i] * (q - P[6 + i]));
        const double e = 1.0 - ex;
        V[sidx(3, i, i)] = 1.0 * (P[i] * e * e + P[9 + i]);
        dV[sidx(3, i, i)] = 2.0 * P[i] * P[3 + i] * ex * (1.0 - ex);
    }
    const double d01 = q - P[18], d02 = q - P[19], d12 = q - P[20];
    const double g01 = NQ_EXP(-P[15] * d01 * d01), g02 = NQ_EXP(-P[16] * d02 * d02), g12 = NQ_EXP(-P[17] * d12 * d12);
    V[sidx(3, 0, 1)] = 1.0 * P[12] * g01; V[sidx(3, 0, 2)] = 1.0 * P[13] * g02; V[sidx(3, 1, 2)] = 1.0 * P[14] * g12;
    dV[sidx(3, 0, 1)] = -2.0 * P[15] * d01 * P[12] * g01;
    dV[sidx(3, 0, 2)] = -2.0 * P[16] * d02 * P[13] * g02;
    dV[sidx(3, 1, 2)] = -2.0 * P[17] * d12 * P[14] * g12;
}

#undef NQ_EXP

}  // namespace nq

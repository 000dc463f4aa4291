// oracle_core.hpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// A plain, scalar C++ restatement of the arithmetic NQCDynamics.jl's ensemble hot path performs
// per trajectory.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
// arm may build, load or call anything under oracle/.  The product (nqcdynamics.jl_b200/) never
// links or imports it.
//
// The reference is pure Julia and cannot run in the build container (no julia binary, and half
// of the arithmetic lives in unvendored registry packages: NQCModels, NQCCalculators,
// RingPolymerArrays, OrdinaryDiffEq).  Parity status: PINNED against every in-tree golden
// vector / known-answer test listed in SURVEY.md section 8c (tests/test_oracle_kats.py) and
// against independent physics checks (diabatic-representation Ehrenfest, analytic harmonic ring
// polymer, Gao/Saller spin-boson curve); "PARITY UNPINNED" for the items the reference tree
// itself does not pin: eigenvector sign choice at t0, Tsit5 staging, ThreeStateMorse and
// MiaoSubotnik numeric parameters (DESIGN.md, section "Oracle").
//
// This file: dense helpers, eigensolvers, the model table (NQCModels restatement) and the
// potential/eigen/adiabatic-derivative/NAC cache (NQCCalculators restatement).
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "../include/nqcb200.h"

namespace nqco {

using cd = std::complex<double>;
using vec = std::vector<double>;
using cvec = std::vector<cd>;

// ---------------------------------------------------------------------------------------------
// Dense helpers (column-major, like Julia).
// ---------------------------------------------------------------------------------------------
inline double& at(vec& A, int n, int i, int j) { return A[i + (size_t)n * j]; }
inline double at(const vec& A, int n, int i, int j) { return A[i + (size_t)n * j]; }

// Cyclic Jacobi eigensolver for a real symmetric matrix; ascending eigenvalues, orthonormal
// columns.  Stands in for LAPACK syevr/syev behind Julia's `eigen(Hermitian(V))`
// (NQCCalculators, external; semantics pinned by test/Core/calculators.jl:99-108:
// w == eigvals(V), |Z| == |eigvecs(V)|).
inline void jacobi_eigh(int n, const double* Ain, double* w, double* Z) {
    vec A(Ain, Ain + (size_t)n * n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) Z[i + (size_t)n * j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                double a = A[i + (size_t)n * j];
                if (i == j) diag += a * a; else off += a * a;
            }
        if (off == 0.0 || off <= 1e-32 * (diag + off)) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double apq = A[p + (size_t)n * q];
                if (apq == 0.0) continue;
                double app = A[p + (size_t)n * p], aqq = A[q + (size_t)n * q];
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {  // columns p,q
                    double akp = A[k + (size_t)n * p], akq = A[k + (size_t)n * q];
                    A[k + (size_t)n * p] = c * akp - s * akq;
                    A[k + (size_t)n * q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {  // rows p,q
                    double apk = A[p + (size_t)n * k], aqk = A[q + (size_t)n * k];
                    A[p + (size_t)n * k] = c * apk - s * aqk;
                    A[q + (size_t)n * k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    double zkp = Z[k + (size_t)n * p], zkq = Z[k + (size_t)n * q];
                    Z[k + (size_t)n * p] = c * zkp - s * zkq;
                    Z[k + (size_t)n * q] = s * zkp + c * zkq;
                }
            }
    }
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return A[a + (size_t)n * a] < A[b + (size_t)n * b]; });
    vec Zs((size_t)n * n);
    for (int j = 0; j < n; ++j) {
        w[j] = A[idx[j] + (size_t)n * idx[j]];
        for (int i = 0; i < n; ++i) Zs[i + (size_t)n * j] = Z[i + (size_t)n * idx[j]];
    }
    std::memcpy(Z, Zs.data(), sizeof(double) * n * n);
}

// Householder tridiagonalisation + implicit-shift QL (classic symmetric-QR algorithm), used for
// large n (IESH baths) where Jacobi is too slow on the CPU.  Same contract as jacobi_eigh.
inline void tridiag_ql_eigh(int n, const double* Ain, double* w, double* Z) {
    vec a(Ain, Ain + (size_t)n * n);  // a[i + n*j]
    vec d(n), e(n);
    auto A = [&](int i, int j) -> double& { return a[i + (size_t)n * j]; };
    for (int i = n - 1; i > 0; --i) {
        int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; ++k) scale += std::fabs(A(i, k));
            if (scale == 0.0) e[i] = A(i, l);
            else {
                for (int k = 0; k <= l; ++k) { A(i, k) /= scale; h += A(i, k) * A(i, k); }
                double f = A(i, l);
                double g = (f >= 0.0 ? -std::sqrt(h) : std::sqrt(h));
                e[i] = scale * g; h -= f * g; A(i, l) = f - g; f = 0.0;
                for (int j = 0; j <= l; ++j) {
                    A(j, i) = A(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k <= j; ++k) g += A(j, k) * A(i, k);
                    for (int k = j + 1; k <= l; ++k) g += A(k, j) * A(i, k);
                    e[j] = g / h; f += e[j] * A(i, j);
                }
                double hh = f / (h + h);
                for (int j = 0; j <= l; ++j) {
                    f = A(i, j); e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; ++k) A(j, k) -= (f * e[k] + g * A(i, k));
                }
            }
        } else e[i] = A(i, l);
        d[i] = h;
    }
    d[0] = 0.0; e[0] = 0.0;
    for (int i = 0; i < n; ++i) {
        int l = i - 1;
        if (d[i] != 0.0) {
            for (int j = 0; j <= l; ++j) {
                double g = 0.0;
                for (int k = 0; k <= l; ++k) g += A(i, k) * A(k, j);
                for (int k = 0; k <= l; ++k) A(k, j) -= g * A(k, i);
            }
        }
        d[i] = A(i, i); A(i, i) = 1.0;
        for (int j = 0; j <= l; ++j) A(j, i) = A(i, j) = 0.0;
    }
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 200) throw std::runtime_error("tridiag_ql_eigh: no convergence");
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i], b = c * e[i];
                    e[i + 1] = (r = std::hypot(f, g));
                    if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
                    s = f / r; c = g / r; g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r); g = c * r - b;
                    for (int k = 0; k < n; ++k) {
                        f = A(k, i + 1);
                        A(k, i + 1) = s * A(k, i) + c * f;
                        A(k, i) = c * A(k, i) - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p; e[l] = g; e[m] = 0.0;
            }
        } while (m != l);
    }
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return d[x] < d[y]; });
    for (int j = 0; j < n; ++j) {
        w[j] = d[idx[j]];
        for (int i = 0; i < n; ++i) Z[i + (size_t)n * j] = A(i, idx[j]);
    }
}

inline void sym_eigh(int n, const double* A, double* w, double* Z) {
    if (n <= 12) jacobi_eigh(n, A, w, Z); else tridiag_ql_eigh(n, A, w, Z);
}

// Cyclic Jacobi for a complex Hermitian matrix (stands in for LAPACK zheevr behind
// `LAPACK.syevr!` on a complex matrix, wavefunction_dynamics.jl:45).
inline void jacobi_heig(int n, const cd* Ain, double* w, cd* Zc) {
    cvec A(Ain, Ain + (size_t)n * n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) Zc[i + (size_t)n * j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                double a = std::norm(A[i + (size_t)n * j]);
                if (i == j) diag += a; else off += a;
            }
        if (off == 0.0 || off <= 1e-32 * (diag + off)) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                cd apq = A[p + (size_t)n * q];
                double mag = std::abs(apq);
                if (mag == 0.0) continue;
                cd phase = apq / mag;  // e^{i phi}
                double app = A[p + (size_t)n * p].real(), aqq = A[q + (size_t)n * q].real();
                double theta = (aqq - app) / (2.0 * mag);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                // Unitary G: columns p,q -> [c*p - s*conj(phase)*q , s*phase*p + c*q]
                cd sp = s * phase, spc = s * std::conj(phase);
                for (int k = 0; k < n; ++k) {  // A <- A G
                    cd akp = A[k + (size_t)n * p], akq = A[k + (size_t)n * q];
                    A[k + (size_t)n * p] = c * akp - spc * akq;
                    A[k + (size_t)n * q] = sp * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {  // A <- G^H A
                    cd apk = A[p + (size_t)n * k], aqk = A[q + (size_t)n * k];
                    A[p + (size_t)n * k] = c * apk - sp * aqk;
                    A[q + (size_t)n * k] = spc * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    cd zkp = Zc[k + (size_t)n * p], zkq = Zc[k + (size_t)n * q];
                    Zc[k + (size_t)n * p] = c * zkp - spc * zkq;
                    Zc[k + (size_t)n * q] = sp * zkp + c * zkq;
                }
            }
    }
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(),
                     [&](int a, int b) { return A[a + (size_t)n * a].real() < A[b + (size_t)n * b].real(); });
    cvec Zs((size_t)n * n);
    for (int j = 0; j < n; ++j) {
        w[j] = A[idx[j] + (size_t)n * idx[j]].real();
        for (int i = 0; i < n; ++i) Zs[i + (size_t)n * j] = Zc[i + (size_t)n * idx[j]];
    }
    std::copy(Zs.begin(), Zs.end(), Zc);
}

// Complex LU determinant with partial pivoting (LAPACK getrf + det(LU), FastDeterminant.jl:20-22).
inline cd complex_det(int n, const cd* Ain) {
    cvec A(Ain, Ain + (size_t)n * n);
    cd det = 1.0;
    for (int k = 0; k < n; ++k) {
        int piv = k; double best = std::abs(A[k + (size_t)n * k]);
        for (int i = k + 1; i < n; ++i) {
            double m = std::abs(A[i + (size_t)n * k]);
            if (m > best) { best = m; piv = i; }
        }
        if (best == 0.0) return 0.0;
        if (piv != k) {
            for (int j = 0; j < n; ++j) std::swap(A[k + (size_t)n * j], A[piv + (size_t)n * j]);
            det = -det;
        }
        cd akk = A[k + (size_t)n * k];
        det *= akk;
        for (int i = k + 1; i < n; ++i) {
            cd f = A[i + (size_t)n * k] / akk;
            if (f == cd(0.0)) continue;
            for (int j = k + 1; j < n; ++j) A[i + (size_t)n * j] -= f * A[k + (size_t)n * j];
        }
    }
    return det;
}

// ---------------------------------------------------------------------------------------------
// Model table -- restatement of NQCModels.jl `potential!` / `derivative!` (EXTERNAL package;
// call sites in the reference: fssh.jl:44, ehrenfest.jl:52, iesh.jl:78-79,192,
// simulations.jl:75-83).  Formulas: docs/src/NQCModels/{analyticmodels,systembathmodels}.md and
// SURVEY.md section 8c / appendix A.5.  One replica: r[D] -> V[n*n], dV[D][n*n] (column-major).
// ---------------------------------------------------------------------------------------------
struct Model {
    int kind = 0, n = 1, D = 1;
    double p[NQCB200_MAX_PARAMS] = {0};
    vec ba, bb;  // bath arrays

    bool classical() const { return kind == NQCB200_MODEL_HARMONIC || kind == NQCB200_MODEL_FREE; }

    // state-independent part (NQCModels.state_independent_potential / _derivative!):
    // only AndersonHolstein carries one (U0); every other quantum model folds it into V.
    double U0(const double* r) const {
        if (kind == NQCB200_MODEL_ANDERSON_HOLSTEIN_MIAO_SUBOTNIK) return 0.5 * p[0] * p[1] * p[1] * r[0] * r[0];
        if (kind == NQCB200_MODEL_ANDERSON_HOLSTEIN_ERPENBECK_THOSS) {
            const double e = std::exp(-p[1] * (r[0] - p[2])) - 1.0;
            return p[0] * e * e + p[3];
        }
        return 0.0;
    }
    void dU0(const double* r, double* g) const {
        for (int i = 0; i < D; ++i) g[i] = 0.0;
        if (kind == NQCB200_MODEL_ANDERSON_HOLSTEIN_MIAO_SUBOTNIK) g[0] = p[0] * p[1] * p[1] * r[0];
        if (kind == NQCB200_MODEL_ANDERSON_HOLSTEIN_ERPENBECK_THOSS) {
            const double ex = std::exp(-p[1] * (r[0] - p[2]));
            g[0] = -2.0 * p[0] * p[1] * ex * (ex - 1.0);
        }
    }
    // ErpenbeckThoss: U1(x), U1'(x) and the coupling scale f(x), f'(x)
    void erpenbeck(double x, double& u1, double& du1, double& f, double& df) const {
        const double ex = std::exp(-p[6] * (x - p[7]));
        u1 = p[4] * ex * ex - p[5] * ex + p[8];
        du1 = -2.0 * p[6] * p[4] * ex * ex + p[6] * p[5] * ex;
        const double t = std::tanh((x - p[11]) / p[10]);
        f = 0.5 * (1.0 - p[9]) * (1.0 - t) + p[9];
        df = -0.5 * (1.0 - p[9]) * (1.0 - t * t) / p[10];
    }

    void potential(const double* r, double* V) const {
        std::fill(V, V + (size_t)n * n, 0.0);
        const double q = r[0];
        switch (kind) {
            case NQCB200_MODEL_TULLY_ONE: {
                double a = p[0], b = p[1], c = p[2], d = p[3];
                double v11 = q > 0 ? a * (1.0 - std::exp(-b * q)) : -a * (1.0 - std::exp(b * q));
                V[0] = v11; V[3] = -v11; V[1] = V[2] = c * std::exp(-d * q * q);
            } break;
            case NQCB200_MODEL_TULLY_TWO: {
                double a = p[0], b = p[1], c = p[2], d = p[3], e = p[4];
                V[0] = 0.0; V[3] = -a * std::exp(-b * q * q) + e; V[1] = V[2] = c * std::exp(-d * q * q);
            } break;
            case NQCB200_MODEL_TULLY_THREE: {
                double a = p[0], b = p[1], c = p[2];
                V[0] = a; V[3] = -a;
                V[1] = V[2] = q < 0 ? b * std::exp(c * q) : b * (2.0 - std::exp(-c * q));
            } break;
            case NQCB200_MODEL_DOUBLE_WELL: {
                double m = p[0], om = p[1], g = p[2], dl = p[3];
                double v0 = 0.5 * m * om * om * q * q, vv = std::sqrt(2.0) * g * q;
                V[0] = v0 + vv; V[3] = v0 - vv; V[1] = V[2] = dl / 2.0;
            } break;
            case NQCB200_MODEL_SPIN_BOSON: {
                double eps = p[0], dl = p[1], harm = 0.0, lin = 0.0;
                for (int j = 0; j < D; ++j) { harm += 0.5 * ba[j] * ba[j] * r[j] * r[j]; lin += bb[j] * r[j]; }
                V[0] = harm + eps + lin; V[3] = harm - eps - lin; V[1] = V[2] = dl;
            } break;
            case NQCB200_MODEL_THREE_STATE_MORSE: {
                for (int i = 0; i < 3; ++i) {
                    double e = 1.0 - std::exp(-p[3 + i] * (q - p[6 + i]));
                    V[i + 3 * i] = p[i] * e * e + p[9 + i];
                }
                const int pi[3] = {0, 0, 1}, pj[3] = {1, 2, 2};
                for (int k = 0; k < 3; ++k) {
                    double dq = q - p[18 + k];
                    double v = p[12 + k] * std::exp(-p[15 + k] * dq * dq);
                    V[pi[k] + 3 * pj[k]] = V[pj[k] + 3 * pi[k]] = v;
                }
            } break;
            case NQCB200_MODEL_HARMONIC: {
                double s = 0.0;
                for (int j = 0; j < D; ++j) s += 0.5 * p[0] * p[1] * p[1] * (r[j] - p[2]) * (r[j] - p[2]);
                V[0] = s;
            } break;
            case NQCB200_MODEL_FREE: V[0] = 0.0; break;
            case NQCB200_MODEL_ANDERSON_HOLSTEIN_MIAO_SUBOTNIK: {
                double m = p[0], om = p[1], g = p[2], dG = p[3];
                double u0 = 0.5 * m * om * om * q * q, u1 = 0.5 * m * om * om * (q - g) * (q - g) + dG;
                V[0] = u1 - u0;
                for (int k = 1; k < n; ++k) {
                    V[k + (size_t)n * k] = ba[k - 1];
                    V[0 + (size_t)n * k] = V[k] = bb[k - 1];
                }
            } break;
            case NQCB200_MODEL_ANDERSON_HOLSTEIN_ERPENBECK_THOSS: {
                double u1, du1, f, df;
                erpenbeck(q, u1, du1, f, df);
                V[0] = u1 - U0(r);
                for (int k = 1; k < n; ++k) {
                    V[k + (size_t)n * k] = ba[k - 1];
                    V[0 + (size_t)n * k] = V[k] = bb[k - 1] * f;
                }
            } break;
            default: throw std::runtime_error("oracle: unknown model");
        }
    }

    void derivative(const double* r, double* dV) const {
        std::fill(dV, dV + (size_t)D * n * n, 0.0);
        const double q = r[0];
        switch (kind) {
            case NQCB200_MODEL_TULLY_ONE: {
                double a = p[0], b = p[1], c = p[2], d = p[3];
                double d11 = q > 0 ? a * b * std::exp(-b * q) : a * b * std::exp(b * q);
                dV[0] = d11; dV[3] = -d11; dV[1] = dV[2] = -2.0 * c * d * q * std::exp(-d * q * q);
            } break;
            case NQCB200_MODEL_TULLY_TWO: {
                double a = p[0], b = p[1], c = p[2], d = p[3];
                dV[3] = 2.0 * a * b * q * std::exp(-b * q * q);
                dV[1] = dV[2] = -2.0 * c * d * q * std::exp(-d * q * q);
            } break;
            case NQCB200_MODEL_TULLY_THREE: {
                double b = p[1], c = p[2];
                dV[1] = dV[2] = q < 0 ? b * c * std::exp(c * q) : b * c * std::exp(-c * q);
            } break;
            case NQCB200_MODEL_DOUBLE_WELL: {
                double m = p[0], om = p[1], g = p[2];
                double d0 = m * om * om * q, dv = std::sqrt(2.0) * g;
                dV[0] = d0 + dv; dV[3] = d0 - dv;
            } break;
            case NQCB200_MODEL_SPIN_BOSON: {
                for (int j = 0; j < D; ++j) {
                    double h = ba[j] * ba[j] * r[j];
                    dV[(size_t)j * 4 + 0] = h + bb[j];
                    dV[(size_t)j * 4 + 3] = h - bb[j];
                }
            } break;
            case NQCB200_MODEL_THREE_STATE_MORSE: {
                for (int i = 0; i < 3; ++i) {
                    double ex = std::exp(-p[3 + i] * (q - p[6 + i]));
                    dV[i + 3 * i] = 2.0 * p[i] * p[3 + i] * ex * (1.0 - ex);
                }
                const int pi[3] = {0, 0, 1}, pj[3] = {1, 2, 2};
                for (int k = 0; k < 3; ++k) {
                    double dq = q - p[18 + k];
                    double v = -2.0 * p[15 + k] * dq * p[12 + k] * std::exp(-p[15 + k] * dq * dq);
                    dV[pi[k] + 3 * pj[k]] = dV[pj[k] + 3 * pi[k]] = v;
                }
            } break;
            case NQCB200_MODEL_HARMONIC:
                for (int j = 0; j < D; ++j) dV[j] = p[0] * p[1] * p[1] * (r[j] - p[2]);
                break;
            case NQCB200_MODEL_FREE: break;
            case NQCB200_MODEL_ANDERSON_HOLSTEIN_MIAO_SUBOTNIK: {
                double m = p[0], om = p[1], g = p[2];
                dV[0] = m * om * om * (q - g) - m * om * om * q;  // d(U1-U0)/dq
            } break;
            case NQCB200_MODEL_ANDERSON_HOLSTEIN_ERPENBECK_THOSS: {
                double u1, du1, f, df, g0[8] = {0};
                erpenbeck(q, u1, du1, f, df);
                dU0(r, g0);
                dV[0] = du1 - g0[0];
                for (int k = 1; k < n; ++k) dV[0 + (size_t)n * k] = dV[k] = bb[k - 1] * df;
            } break;
            default: throw std::runtime_error("oracle: unknown model");
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Cache -- restatement of NQCCalculators.update_cache! for one replica (EXTERNAL package; call
// sites bab_electronics.jl:53,75; bcb_electronics.jl:45,73; verlet_with_electronics.jl:33,62;
// classical.jl:64).  Semantics pinned by test/Core/calculators.jl:99-108 (adiabatic_derivative
// = Z' dV Z, w = eigvals(V)) and test/Dynamics/fssh.jl:32-33 (NAC antisymmetric).
//   eigen: ascending w, columns sign-corrected against the previous Z:
//          dot(Z_new[:,i], Z_old[:,i]) < 0 -> flip          (recalled)
//   nac[I][j,i] = -adiab[I][j,i] / (w_j - w_i), zero diagonal   (<phi_j| d/dR phi_i>)
// ---------------------------------------------------------------------------------------------
struct Cache {
    int n = 1, D = 1;
    vec V, dV, w, Z, adiab, nac;
    void init(int n_, int D_, const double* Zref) {
        n = n_; D = D_;
        V.assign((size_t)n * n, 0.0); dV.assign((size_t)D * n * n, 0.0); w.assign(n, 0.0);
        Z.assign((size_t)n * n, 0.0); adiab.assign((size_t)D * n * n, 0.0); nac.assign((size_t)D * n * n, 0.0);
        if (Zref) std::copy(Zref, Zref + (size_t)n * n, Z.begin());
        else for (int i = 0; i < n; ++i) Z[i + (size_t)n * i] = 1.0;  // gauge reference = identity
    }
    void update(const Model& m, const double* r, bool need_nac = true) {
        m.potential(r, V.data());
        m.derivative(r, dV.data());
        if (m.classical()) return;
        vec Zn((size_t)n * n), wn(n);
        sym_eigh(n, V.data(), wn.data(), Zn.data());
        for (int i = 0; i < n; ++i) {
            double dot = 0.0;
            for (int k = 0; k < n; ++k) dot += Zn[k + (size_t)n * i] * Z[k + (size_t)n * i];
            if (dot < 0.0) for (int k = 0; k < n; ++k) Zn[k + (size_t)n * i] = -Zn[k + (size_t)n * i];
        }
        Z = Zn; w = wn;
        vec tmp((size_t)n * n);
        for (int I = 0; I < D; ++I) {
            const double* d = &dV[(size_t)I * n * n];
            double* a = &adiab[(size_t)I * n * n];
            // tmp = dV Z ; a = Z' tmp
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    double s = 0.0;
                    for (int k = 0; k < n; ++k) s += d[i + (size_t)n * k] * Z[k + (size_t)n * j];
                    tmp[i + (size_t)n * j] = s;
                }
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    double s = 0.0;
                    for (int k = 0; k < n; ++k) s += Z[k + (size_t)n * i] * tmp[k + (size_t)n * j];
                    a[i + (size_t)n * j] = s;
                }
            if (need_nac) {
                double* c = &nac[(size_t)I * n * n];
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < n; ++j)
                        c[j + (size_t)n * i] = (i == j) ? 0.0 : -a[j + (size_t)n * i] / (w[j] - w[i]);
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Ring-polymer normal modes -- restatement of RingPolymerArrays.NormalModeTransformation
// (EXTERNAL; pinned by docs/src/api/RingPolymerArrays/ringpolymerarrays.md:93-133 and
// test/Core/ring_polymers.jl:22-27) and RingPolymers.cayley_propagator (ring_polymer.jl:71-82).
// ---------------------------------------------------------------------------------------------
inline vec normal_mode_matrix(int B) {  // U[j + B*k]
    vec U((size_t)B * B);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < B; ++k)
        for (int j = 0; j < B; ++j) {
            double u;
            if (k == 0) u = 1.0 / std::sqrt((double)B);
            else if (2 * k < B) u = std::sqrt(2.0 / B) * std::cos(2.0 * pi * j * k / B);
            else if (2 * k == B) u = ((j % 2) ? -1.0 : 1.0) / std::sqrt((double)B);
            else u = std::sqrt(2.0 / B) * std::sin(2.0 * pi * j * k / B);
            U[j + (size_t)B * k] = u;
        }
    return U;
}
inline vec matsubara_frequencies(int B, double omega_n) {  // ring_polymer.jl:60
    vec w(B);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < B; ++k) w[k] = 2.0 * omega_n * std::sin(k * pi / B);
    return w;
}
// cay(dt*A) = inv(I - dt*A/2) (I + dt*A/2), A = [0 1; -w^2 0]; returns [c11,c12,c21,c22] per mode.
// half=true: principal square root of that 2x2 (ring_polymer.jl:79, `real.(sqrt(cay(dt.*A)))`).
inline vec cayley_propagator(int B, double omega_n, double dt, bool half) {
    vec wk = matsubara_frequencies(B, omega_n), out((size_t)4 * B);
    for (int k = 0; k < B; ++k) {
        double w2 = wk[k] * wk[k];
        // M = I - dt/2 A = [1, -dt/2; w2 dt/2, 1];  N = I + dt/2 A = [1, dt/2; -w2 dt/2, 1]
        double h = dt / 2.0, det = 1.0 + w2 * h * h;
        // inv(M) = 1/det [1, h; -w2 h, 1]
        double i11 = 1.0 / det, i12 = h / det, i21 = -w2 * h / det, i22 = 1.0 / det;
        double c11 = i11 * 1.0 + i12 * (-w2 * h), c12 = i11 * h + i12 * 1.0;
        double c21 = i21 * 1.0 + i22 * (-w2 * h), c22 = i21 * h + i22 * 1.0;
        if (half) {
            // principal sqrt of a 2x2 with det 1: sqrt(M) = (M + I)/sqrt(tr(M) + 2)
            double s = std::sqrt(c11 + c22 + 2.0);
            c11 = (c11 + 1.0) / s; c22 = (c22 + 1.0) / s; c12 /= s; c21 /= s;
        }
        out[4 * k + 0] = c11; out[4 * k + 1] = c12; out[4 * k + 2] = c21; out[4 * k + 3] = c22;
    }
    return out;
}

}  // namespace nqco

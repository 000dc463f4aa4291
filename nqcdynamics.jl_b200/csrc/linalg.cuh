// linalg.cuh -- register-resident symmetric eigensolver for the tiny diabatic matrices (n <= 6).
//
// Replaces LAPACK syevr behind NQCCalculators' `eigen(Hermitian(V))` (external; semantics pinned by
// test/Core/calculators.jl:99-108).  All loops have compile-time bounds so that, after unrolling,
// every matrix element is a named register: no local memory, no shared memory.
//   n == 2 : one Jacobi rotation is exact (closed form).
//   n >= 3 : cyclic Jacobi sweeps until the off-diagonal mass is below 1e-32 of the total.
// Eigenvalues ascending; the column signs are then fixed by continuity with the previous
// eigenvectors: dot(Z_new[:,i], Z_old[:,i]) < 0 -> flip   (NQCCalculators correct_phase!).
#pragma once
#include "common.cuh"

namespace nq {

template <int N>
struct Eig {
    double w[N];
    double Z[N][N];  // Z[row][col], column i = eigenvector i
};

#if defined(__CUDACC__)
// 1 / sqrt(x) for a normal positive x: hardware seed (MUFU.RSQ64H) + one third-order correction
// y (1 + e/2 + 3 e^2/8), e = 1 - x y^2 -- no special-case path, i.e. no branch: two of these can be in flight in one
// basic block (the library rsqrt() carries a slow-path branch that ends the block).
NQ_D double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);
    return fma(y * e, fma(0.375, e, 0.5), y);
}

// jacobi_rotate without branches (same rotation, the zero-element case as selects)
template <int N>
NQ_D void jacobi_rotate_bf(double (&A)[N][N], double (&Z)[N][N], int p, int q) {
    const double apq = A[p][q];
    const double a = A[q][q] - A[p][p], b = 2.0 * apq;
    const double x = fma(a, a, b * b);
    const bool zero = (apq == 0.0) || !(x > 1.0e-280);
    const double ih = rsqrt_pos(fmax(x, 1.0e-280));
    const double c2 = fma(0.5 * fabs(a), ih, 0.5);
    const double rc = rsqrt_pos(c2);
    const double c = zero ? 1.0 : c2 * rc;
    const double s = zero ? 0.0 : (a >= 0.0 ? 0.5 : -0.5) * (b * ih) * rc;
    const double t = s * rc;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (k != p && k != q) {
            const double akp = A[k][p], akq = A[k][q];
            const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
            A[k][p] = np_; A[p][k] = np_;
            A[k][q] = nq_; A[q][k] = nq_;
        }
    }
    A[p][p] -= t * apq;
    A[q][q] += t * apq;
    A[p][q] = 0.0; A[q][p] = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double zkp = Z[k][p], zkq = Z[k][q];
        Z[k][p] = c * zkp - s * zkq;
        Z[k][q] = s * zkp + c * zkq;
    }
}

#endif

template <int N>
NQ_HD void jacobi_rotate(double (&A)[N][N], double (&Z)[N][N], int p, int q) {
    const double apq = A[p][q];
    if (N > 2 && apq == 0.0) return;
#if defined(__CUDA_ARCH__)
    jacobi_rotate_bf<N>(A, Z, p, q);     // same rotation, straight-line code (two reciprocal square roots without slow paths)
#else
    if (apq == 0.0) return;
    // Rotation angle |theta| <= pi/4 with tan 2 theta = b / a (a = A_qq - A_pp, b = 2 A_pq):
    //     cos 2theta = |a| / hyp,  c = sqrt((1 + cos 2theta) / 2),  s = sgn(a) (b / hyp) / (2 c),  t = s / c.
    // FP64 sqrt and division are ~90 / ~70-cycle dependent sequences on B200 (tools/micro/fp64_latency.cu) and the
    // three rotations of a sweep are serial, so the rotation is written with two reciprocal square roots and no
    // division: 1/hyp = rsqrt(a^2 + b^2), 1/c = rsqrt(c^2).
    const double a = A[q][q] - A[p][p], b = 2.0 * apq;
    const double ih = 1.0 / sqrt(fma(a, a, b * b));
    const double c2 = fma(0.5 * fabs(a), ih, 0.5);
    const double rc = 1.0 / sqrt(c2);
    const double c = c2 * rc;
    const double s = (a >= 0.0 ? 0.5 : -0.5) * (b * ih) * rc;
    const double t = s * rc;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if (k != p && k != q) {
            const double akp = A[k][p], akq = A[k][q];
            const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
            A[k][p] = np_; A[p][k] = np_;
            A[k][q] = nq_; A[q][k] = nq_;
        }
    }
    A[p][p] -= t * apq;
    A[q][q] += t * apq;
    A[p][q] = 0.0; A[q][p] = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double zkp = Z[k][p], zkq = Z[k][q];
        Z[k][p] = c * zkp - s * zkq;
        Z[k][q] = s * zkp + c * zkq;
    }
#endif
}

// Vp: packed upper triangle (row-wise) of the symmetric matrix.
template <int N>
NQ_HD void sym_eigh(const double (&Vp)[sym_size(N)], Eig<N>& e) {
    double A[N][N];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = j; k < N; ++k) { A[j][k] = Vp[sidx(N, j, k)]; A[k][j] = A[j][k]; }
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = 0; k < N; ++k) e.Z[j][k] = (j == k) ? 1.0 : 0.0;

    if (N == 2) {
        jacobi_rotate<N>(A, e.Z, 0, 1);
    } else if (N > 2) {
        for (int sweep = 0; sweep < 30; ++sweep) {
            double off = 0.0, diag = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j)
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    if (j == k) diag += A[j][k] * A[j][k]; else off += A[j][k] * A[j][k];
                }
            if (off <= 1e-32 * (diag + off)) break;
#pragma unroll
            for (int p = 0; p < N - 1; ++p)
#pragma unroll
                for (int q = p + 1; q < N; ++q) jacobi_rotate<N>(A, e.Z, p, q);
        }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) e.w[j] = A[j][j];
    // ascending order (stable odd-even transposition network: N passes)
#pragma unroll
    for (int pass = 0; pass < N; ++pass)
#pragma unroll
        for (int j = 0; j < N - 1; ++j) {
            if (e.w[j + 1] < e.w[j]) {
                const double tw = e.w[j]; e.w[j] = e.w[j + 1]; e.w[j + 1] = tw;
#pragma unroll
                for (int k = 0; k < N; ++k) { const double tz = e.Z[k][j]; e.Z[k][j] = e.Z[k][j + 1]; e.Z[k][j + 1] = tz; }
            }
        }
}

#if defined(__CUDACC__)
// Two independent eigenproblems in lock step (instruction-level parallelism for the serial rotation chains: the
// ring-polymer kernel visits its beads two at a time).  Eigenvalues / eigenvectors are left in Jacobi order --
// `eig_rank` gives the ascending position of each column, so that a caller that needs one column (the occupied state's
// bead force, fssh.jl:67-74) or order-independent sums does not pay for sorting Z.
template <int N>
NQ_D void sym_eigh_pair_unsorted(const double (&V0)[sym_size(N)], const double (&V1)[sym_size(N)], double (&w0)[N],
                                 double (&Z0)[N][N], double (&w1)[N], double (&Z1)[N][N]) {
    double A0[N][N], A1[N][N];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = j; k < N; ++k) {
            A0[j][k] = V0[sidx(N, j, k)]; A0[k][j] = A0[j][k];
            A1[j][k] = V1[sidx(N, j, k)]; A1[k][j] = A1[j][k];
        }
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int k = 0; k < N; ++k) { Z0[j][k] = (j == k) ? 1.0 : 0.0; Z1[j][k] = Z0[j][k]; }
    if (N == 2) { jacobi_rotate_bf<N>(A0, Z0, 0, 1); jacobi_rotate_bf<N>(A1, Z1, 0, 1); }   // one rotation is exact
    for (int sweep = 0; N > 2 && sweep < 30; ++sweep) {
        double off0 = 0.0, diag0 = 0.0, off1 = 0.0, diag1 = 0.0;
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int k = j; k < N; ++k) {
                if (j == k) { diag0 = fma(A0[j][k], A0[j][k], diag0); diag1 = fma(A1[j][k], A1[j][k], diag1); }
                else { off0 = fma(2.0 * A0[j][k], A0[j][k], off0); off1 = fma(2.0 * A1[j][k], A1[j][k], off1); }
            }
        if (off0 <= 1e-32 * (diag0 + off0) && off1 <= 1e-32 * (diag1 + off1)) break;
#pragma unroll
        for (int p = 0; p < N - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < N; ++q) { jacobi_rotate_bf<N>(A0, Z0, p, q); jacobi_rotate_bf<N>(A1, Z1, p, q); }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) { w0[j] = A0[j][j]; w1[j] = A1[j][j]; }
}

// ascending position of every eigenvalue (ties in index order, like the stable sorting network of sym_eigh)
template <int N>
NQ_D void eig_rank(const double (&w)[N], int (&rank)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        int r = 0;
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (j != i) r += (w[j] < w[i] || (w[j] == w[i] && j < i)) ? 1 : 0;
        rank[i] = r;
    }
}
#endif

// Column-sign continuity with the previous eigenvectors; Zref is updated to the new vectors.
template <int N>
NQ_HD void fix_gauge(Eig<N>& e, double (&Zref)[N][N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double dot = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) dot += e.Z[k][i] * Zref[k][i];
        const double sgn = dot < 0.0 ? -1.0 : 1.0;
#pragma unroll
        for (int k = 0; k < N; ++k) { e.Z[k][i] *= sgn; Zref[k][i] = e.Z[k][i]; }
    }
}

// A = Z' * S * Z for packed symmetric S; result packed symmetric.
template <int N>
NQ_HD void similarity(const double (&Sp)[sym_size(N)], const double (&Z)[N][N], double (&Ap)[sym_size(N)]) {
    double T[N][N];  // T = S * Z
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) s += Sp[(i <= k) ? sidx(N, i, k) : sidx(N, k, i)] * Z[k][j];
            T[i][j] = s;
        }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = i; j < N; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) s += Z[k][i] * T[k][j];
            Ap[sidx(N, i, j)] = s;
        }
}

}  // namespace nq

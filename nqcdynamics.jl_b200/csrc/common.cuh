// common.cuh -- shared definitions for the sm_100a trajectory kernels.
#pragma once
#include <cstdint>
#include <cmath>

#include "../../include/nqcb200.h"

#if defined(__CUDACC__)
#define NQ_HD __host__ __device__ __forceinline__
#define NQ_D __device__ __forceinline__
#else
#define NQ_HD inline
#define NQ_D inline
#endif

namespace nq {

constexpr int kBlockThreads = 128;

// Packed layout of the enabled observables: each observable occupies [nsave][width] doubles at
// obs_offset[id] in the per-shard accumulator (and, per trajectory, [nsave][width][ntraj] at the
// same offset scaled by ntraj in the output stream).
struct ObsLayout {
    int32_t width[NQCB200_OBS_COUNT];
    int64_t offset[NQCB200_OBS_COUNT];  // -1 when disabled
    int64_t total;                      // doubles in the accumulator
};

// AdiabaticIESH (CTA per trajectory, kernel_iesh.cuh): tile / shared-memory plan computed on the host.
struct IeshLayout {
    int32_t threads;    // block size (12 warps: 3 per SM sub-partition)
    int32_t nrt;        // row tiles of 8 states
    int32_t ldg;        // leading dimension of G (column-major), >= 8*nrt and = 4 (mod 16): conflict-free DMMA fragments
    int32_t nct;        // column tiles (8 doubles = 4 electrons, re/im interleaved) per psi chunk
    int32_t ldb;        // row length of the psi chunk, >= 8*nct and = 4 (mod 16)
    int32_t nchunks;    // psi chunks per trajectory
    int32_t resident;   // 1: G lives in shared memory; 0: streamed from global in slabs of kb columns
    int32_t kb;
    int32_t nslab;
    int32_t lds;        // leading dimension of the overlap matrix S in the hop phase
    int32_t lr;         // lanes per secular-equation root (power of two)
    int32_t off_b;      // psi-chunk offset (doubles) inside the work region
    int32_t off_hop;    // hop-phase offset inside the work region
    int32_t work_doubles;
    int32_t smem_bytes;
    int32_t rounds;     // full rows of 8 states per warp (0, 1 or 2); the other rows are dealt out tile by tile
};

// Kernel parameter block (passed by value as a __grid_constant__).
struct KParams {
    // sizes
    int64_t ntraj;
    int64_t traj_offset;
    int32_t n, D, B, ne;
    // run control
    int64_t step0;       // global step index of the first step of this launch
    int32_t nsteps;      // steps in this launch
    int32_t save_every, nsave;
    int32_t rescaling, rng, diagnostics, per_trajectory;
    int32_t estimate_probability, disable_hopping;
    int32_t mean_field;   // EhrenfestNA on the IESH kernel: force and estimators from psi, no occupations
    uint32_t observables;
    uint64_t seed;
    double dt, t0, omega_n, nrpmd_gamma, edc_C;
    double tsit5_ha[21];   // (dt/5) a_ij of Tsit5, row by row (a21; a31 a32; ...; a71..a76): constant-bank operands
    double params[NQCB200_MAX_PARAMS];
    // model / system arrays (device)
    const double* masses;   // [D]
    const double* bath_a;   // [nbath]
    const double* bath_b;
    // ring polymer: normal-mode tables (lane-fastest, see kernel_ring.cuh) and Cayley 2x2 per mode
    const double* nm_to;    // [j*B + k] = U[j,k]
    const double* nm_from;  // [k*B + j] = U[j,k]
    const double* cayley;   // [4*k + {c11,c12,c21,c22}]
    // state, SoA: [component][traj]
    double* r;      // [B*D][T]
    double* v;      // [B*D][T]
    double* acc;    // [B*D][T]   acceleration k carried between steps (quirk Q2)
    double* sig_re; // [n*n or n*ne][T]
    double* sig_im;
    int32_t* state; // [1 or ne][T]  0-based on the device
    double* Zprev;  // [(B (+1 centroid)) * n*n][T] eigenvector gauge reference
    double* ecur;   // [n + n*n][T]  "current" half of the electronic double buffer: E then v.d
    double* pop0;   // [2n][T] initial diabatic / adiabatic population (correlation functions)
    double* qmap;   // [B*n][T]
    double* pmap;
    double* sb_carry; // [2][T]  SpinBoson thread-per-trajectory kernel: force scalars (A, B) carried between launches
    // kernel_spinboson_epoch.cuh: per-mode constants {dt w^2/m, c/m, c, c w^2/m}[D] and the bath's lag tables
    // [2 shapes][3 sums][32 lags] followed by sum c^2/m and sum c^2 (host-built at create); bath sums of the current
    // epoch [3E][T], impulses [2][E+1][T], entry / exit half kick [T]; sb_gen: some trajectory has tr sigma != 1; and
    // the shape of the bath pass being launched
    const double* sb_kc;
    const double* sb_kap;
    double* sb_sums;
    double* sb_f;
    double* sb_aux;
    int32_t sb_entry, sb_nrep, sb_nfree, sb_exit, sb_gen;
    int64_t tlo, thi;        // trajectory range [tlo, thi) this launch works on (chunked nqcb200_run_from_host); 0, ntraj otherwise
    // launch-fused initialisation (nqcb200_run_from_host, kernels with KernelSet::fused_init): trajectory-major
    // [T][B*D] sources read by the step kernel itself at step0 == 0 (device staging or pinned host memory)
    const double* r_aos;
    const double* v_aos;
    int32_t init_basis, init_sample_state;
    const double* init_state_draw;
    // AdiabaticIESH: psi / occupations are TRAJECTORY-major ([T][n*ne], [T][ne]); see kernel_iesh.cuh
    IeshLayout iesh;
    int32_t iesh_impurity;   // 0: MiaoSubotnik (constant coupling), 1: ErpenbeckThoss (coupling scaled by f(x)), kernel_iesh.cuh
    double* iesh_lam;   // [T][n]   adiabatic energies of the last step (warm start of the root finder)
    double* iesh_sgn;   // [T][n]   eigenvector column signs (gauge), constant along a trajectory
    double* iesh_orth;  // [T]      1.0 when the initial orbitals are orthonormal (determinant-free pruning bound)
    double* iesh_G;     // [grid][ldg*kb*nslab] v.d scratch when it does not fit in shared memory
    // ThermalLangevin: friction gamma_0 and injected normals xi[((step - noise_step0) * T + traj) * B + mode]
    double langevin_gamma;
    const double* noise;
    int64_t noise_step0;
    // TerminatingCallback(u -> r[term_dof] < term_lo || r[term_dof] > term_hi) (callbacks.jl:29): step count at which
    // terminate! fired per trajectory (-1 = running); read by the TERM instantiations and by the IESH kernel
    int32_t term_dof;
    int32_t term_outgoing;   // != 0: the window test also asks for an outward velocity (v < 0 below lo, v > 0 above hi)
    double term_lo, term_hi;
    double term_tcut;        // ... || t > term_tcut (+inf: no time clause)
    long long* term_step;   // [T]
    // draws (injected): xi[(step - draws_step0) * T + traj]
    const double* draws;
    int64_t draws_step0;
    // outputs
    double* obs_sum;    // [layout.total]
    double* obs_traj;   // [layout.total][T] or nullptr
    ObsLayout layout;
    // diagnostics of the last step (optional)
    double* diag_eig;   // [n][T]
    double* diag_nac;   // [D*n*n][T]
    double* diag_Z;     // [n*n][T]
    // counters: [0]=hops [1]=frustrated
    unsigned long long* counters;
};

NQ_HD constexpr int sym_size(int n) { return n * (n + 1) / 2; }
NQ_HD constexpr int asym_size(int n) { return n * (n - 1) / 2; }
// packed upper-triangular index for j <= k (row-wise)
NQ_HD constexpr int sidx(int n, int j, int k) { return j * n - j * (j - 1) / 2 + (k - j); }
// packed strict upper index for j < k
NQ_HD constexpr int aidx(int n, int j, int k) { return j * n - j * (j + 1) / 2 + (k - j - 1); }

#if defined(__CUDACC__)
// exp(x) without a special-case path.  The library exp() ends its basic block with a range branch, so the six
// exponentials of a ThreeStateMorse evaluation ran as six serial dependency chains (ncu, profiles/r02: 21 % of the RPSH
// kernel's samples); this one is straight-line code and several of them overlap.  n = rint(x log2 e) by the 1.5 * 2^52 shift (the
// exponent is clamped to the normal range: 2^-1022 instead of a subnormal or zero, 2^1023 x 1.. instead of +inf),
// Cody-Waite reduction r = x - n ln2 with the fdlibm constants, Taylor polynomial of degree 13 on |r| <= ln2 / 2
// (truncation 6e-18 relative), 2^n through the exponent field.  Within 1 ulp of the library function.
NQ_D double exp_nb(double x) {
    const double shift = 6755399441055744.0;
    const double t = fma(x, 1.4426950408889634074, shift);
    const int n = min(max(__double2loint(t), -1022), 1023);
    const double fn = t - shift;
    double r = fma(fn, -6.93147180369123816490e-01, x);
    r = fma(fn, -1.90821492927058770002e-10, r);
    double q = 1.0 / 6227020800.0;
    q = fma(q, r, 1.0 / 479001600.0);
    q = fma(q, r, 1.0 / 39916800.0);
    q = fma(q, r, 1.0 / 3628800.0);
    q = fma(q, r, 1.0 / 362880.0);
    q = fma(q, r, 1.0 / 40320.0);
    q = fma(q, r, 1.0 / 5040.0);
    q = fma(q, r, 1.0 / 720.0);
    q = fma(q, r, 1.0 / 120.0);
    q = fma(q, r, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    return q * __hiloint2double((n + 1023) << 20, 0);
}
// 1 / x and a / b for NORMAL operands without the IEEE slow path: hardware seed (rcp.approx.ftz.f64, ~20 bits), two
// Newton steps, and for the quotient one residual correction (correctly rounded in all but rare cases; the library
// division is ~15 instructions with a branch that ends the basic block, this is 5 / 8 straight-line FMAs).
NQ_D double rcp_nb(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
NQ_D double div_fast(double a, double b) {
    const double rb = rcp_nb(b);
    const double q = a * rb;
    return fma(fma(-q, b, a), rb, q);
}
// a / b given rb = 1 / b: product + one residual correction (no slow-path branch)
NQ_D double div_nb(double a, double b, double rb) {
    const double q = a * rb;
    return fma(fma(-q, b, a), rb, q);
}
#endif

}  // namespace nq

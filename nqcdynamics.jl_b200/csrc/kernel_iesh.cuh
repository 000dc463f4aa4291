// kernel_iesh.cuh -- Simulation{AdiabaticIESH} on the Newns-Anderson (AndersonHolstein) Hamiltonian:
// one CTA per trajectory, persistent over the steps of a launch.
//
// Reference restated (per nuclear step, D = 1):
//   VerletwithElectronics.perform_step!   src/DynamicsMethods/IntegrationAlgorithms/verlet_with_electronics.jl:42-69
//   update_cache! (eigen, Z'dVZ, NAC)     NQCCalculators (external)
//   acceleration!                         SurfaceHoppingMethods/iesh.jl:190-207
//   propagate_wavefunction!               DynamicsUtils/wavefunction_dynamics.jl:15-58   psi' = exp(-i (W - i v.d) dt) psi
//   IESHCallback                          iesh.jl:231-335, 390-407 (one draw for pruning and selection, Q6)
//   rescale_velocity!                     surface_hopping.jl:64-99,115-168
//   EDC decoherence                       decoherence_corrections.jl:21-38 via iesh.jl:416-441
//   estimators                            iesh.jl:337-388
//
// How the same numbers are produced with far less work than the reference's dense formulation:
//  * The diabatic matrix is an ARROWHEAD: H[0,0] = h(q), H[k,k] = eps_k, H[0,k] = V_k.  Its eigenvalues
//    are the roots of the secular equation  h - l - sum_k V_k^2 / (eps_k - l) = 0, one per interval
//    between consecutive bath energies (interlacing), found by a safeguarded Newton iteration on
//    mu*f(mu) in the offset mu from the nearest pole (warm-started from the previous step); the
//    eigenvectors are z_i ~ (1, V_k / (l_i - eps_k)).  O(n^2) instead of 9 n^3.
//  * dV/dq = h'(q) e0 e0', so Z'dVZ = h' z0 z0' and d_ij = -h' z0_i z0_j / (w_i - w_j): no similarity
//    transform.  (w_i - w_j is formed from pole offsets, not from the rounded eigenvalues.)
//  * Column-sign continuity (NQCCalculators: flip when dot(Z_new[:,i], Z_old[:,i]) < 0): both the new and
//    the old root i lie in the same pole interval, so the dot product of the un-normalised secular
//    eigenvectors is 1 + sum_k V_k^2/((l-eps_k)(l'-eps_k)) > 0 -- the sign fixed at t0 never changes.
//  * exp(-i (W - i G) dt) psi  (G = v.d real antisymmetric) is applied as a Taylor series in the shifted
//    generator A = -i dt (W - s) - dt G acting on the n x 2ne real matrix [Re psi | Im psi]: one real
//    GEMM  G * [X|Y]  per Horner stage on the FP64 tensor cores (DMMA m8n8k4; G in shared memory or streamed from
//    L2 in slabs).
//    The truncation order is chosen per step from rho = dt (max|w - s| + ||G||_F) so that the remainder
//    is below 1e-17; eigensolver-independent, agrees with V exp(-i lambda dt) V' psi to rounding.
//  * Hop test: det S (ne x ne complex overlap) from an LU factorisation held in registers across the CTA; the
//    reference's pruning estimate and single draw are kept.  Only when the estimate does not rule a hop out, one
//    in-place Gauss-Jordan inversion of S gives, through the matrix-determinant lemma, every
//    det S_{e->m} / det S = (psi_m . S^-1)_e -- instead of ne (n - ne) separate LU factorisations.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "kernels.h"

namespace nq {

#if defined(__CUDACC__)

// ---- small helpers ------------------------------------------------------------------------------
NQ_D void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
NQ_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
NQ_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// 1/x to within one ulp: hardware seed (rcp.approx.ftz.f64, ~20 bits) and two Newton steps.  Used where the quotient
// feeds an iteration or a 1e-10-tolerance quantity (secular sums, v.d); an IEEE division costs about twice as much
// and the root finder / propagator are insensitive to the last bit.  |x| must be a normal number.
NQ_D double iesh_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// block-wide sum / max, result broadcast to every thread (red: >= 34 doubles of shared memory)
NQ_D double iesh_block_sum(double x, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    const int warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[warp] = x;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}
NQ_D double iesh_block_max(double x, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    const int warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[warp] = x;
    __syncthreads();
    double s = red[0];
    for (int w = 1; w < nw; ++w) s = fmax(s, red[w]);
    return s;
}

// Shared-memory carve-up common to the step / init / diagnostics kernels.
struct IeshSmem {
    double *eps, *V2, *Vb, *lam, *mu, *z0, *ws, *sgn, *red, *pop;
    double* fsc;        // [0] coupling scale f(x) of the current geometry (1 for MiaoSubotnik), [1] h(x)
    double* rp;         // ring polymer (RPIESH / RP-EhrenfestNA): r[32] v[32] acc[32] + 64 doubles of normal-mode scratch
    int *pole, *occ, *un, *flag, *ctl;
    double* work;
    int np;
    NQ_D void carve(double* base, int n) {
        np = (n + 3) & ~3;
        eps = base; V2 = eps + np; Vb = V2 + np; lam = Vb + np; mu = lam + np; z0 = mu + np; ws = z0 + np;
        sgn = ws + np; pop = sgn + np; red = pop + np; fsc = red + 56;
        rp = red + 64;
        pole = (int*)(rp + 160); occ = pole + np; un = occ + np; flag = un + np; ctl = flag + np;
        work = (double*)(ctl + 16);
    }
};
// doubles needed in front of the work region (host mirror of carve)
NQ_HD int iesh_small_doubles(int n) {
    const int np = (n + 3) & ~3;
    return 9 * np + 64 + 160 + (4 * np + 16) / 2;
}

// Secular function at offset mu from pole p: f = (h - eps_p) - mu - sum_k V_k^2 / ((eps_k - eps_p) - mu),
// fp = -f' = 1 + sum_k V_k^2 / (...)^2, s3 = sum_k V_k^2 / (...)^3 (so d fp / d mu = 2 s3).
// The lr lanes of a group split the sum (xor butterfly: bit-identical result on all of them).
NQ_D void iesh_secular(const double* eps, const double* V2, double f2, int M, int p, double hp, double mu, int sub, int lr,
                       unsigned mask, double& f, double& fp, double& s3) {
    const double ep = eps[p];
    double s1 = 0.0, s2 = 0.0, s3l = 0.0;
#pragma unroll 4
    for (int k = sub; k < M; k += lr) {
        const double d = (eps[k] - ep) - mu;
        const double inv = iesh_rcp(d);
        const double t = (V2[k] * f2) * inv;
        const double u = t * inv;
        s1 += t;
        s2 += u;
        s3l = fma(u, inv, s3l);
    }
    for (int o = lr >> 1; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(mask, s1, o);
        s2 += __shfl_xor_sync(mask, s2, o);
        s3l += __shfl_xor_sync(mask, s3l, o);
    }
    f = (hp - mu) - s1;
    fp = 1.0 + s2;
    s3 = s3l;
}

// Root i of the secular equation (i = 0..M, ascending): pole index p, offset mu = lambda - eps_p, fp = -f'(lambda).
// Roots interlace the poles: root i lies in (eps[i-1], eps[i]) (eps[-1] = -inf, eps[M] = +inf).  The offset is taken
// from the pole nearest to the warm start lam_prev (the previous step's eigenvalue; NaN = cold start: the pole
// is then chosen by the sign of f at the interval midpoint).  Safeguarded Newton on phi(mu) = mu f(mu), which
// removes the pole at mu = 0 and converges quadratically; the last step is accepted when it is below 3e-8 |mu|
// (error after it ~ 1e-15 |mu|) and fp is corrected to first order with s3.
NQ_D void iesh_root(const double* eps, const double* V2, double f2, int M, int i, double h, double vnorm, double lam_prev,
                    int sub, int lr, unsigned mask, int& p_out, double& mu_out, double& fp_out) {
    int p;
    double lo, hi;
    const bool cold = !(lam_prev == lam_prev);
    if (i == 0) {
        p = 0; hi = 0.0;
        lo = (fmin(h, eps[0]) - vnorm) - eps[0];
        lo -= 1e-6 * fabs(lo) + 1e-300;
    } else if (i == M) {
        p = M - 1; lo = 0.0;
        hi = (fmax(h, eps[M - 1]) + vnorm) - eps[M - 1];
        hi += 1e-6 * fabs(hi) + 1e-300;
    } else {
        const double gap = eps[i] - eps[i - 1];
        bool right;
        if (cold) {
            double f, fp, s3;
            iesh_secular(eps, V2, f2, M, i - 1, h - eps[i - 1], 0.5 * gap, sub, lr, mask, f, fp, s3);
            right = f > 0.0;                                  // f decreases: root right of the midpoint
        } else right = (eps[i] - lam_prev) < (lam_prev - eps[i - 1]);
        if (right) { p = i; lo = -gap; hi = 0.0; }
        else { p = i - 1; lo = 0.0; hi = gap; }
    }
    const double hp = h - eps[p];
    double mu = lam_prev - eps[p];
    if (!(mu > lo && mu < hi)) {
        if (i == 0 || i == M) mu = 0.5 * (lo + hi);
        else mu = (p == i) ? 0.25 * lo : 0.25 * hi;          // a quarter of the gap away from the chosen pole
    }
    double fpv = 1.0;
    for (int it = 0; it < 100; ++it) {
        double f, fp, s3;
        iesh_secular(eps, V2, f2, M, p, hp, mu, sub, lr, mask, f, fp, s3);
        fpv = fp;
        if (f == 0.0) break;
        if (f > 0.0) lo = mu; else hi = mu;
        // Newton on phi(mu) = mu f(mu): phi' = f + mu f' = f - mu fp
        double munew = mu - mu * f / (f - mu * fp);
        const bool newton = (munew > lo && munew < hi);
        if (!newton) munew = 0.5 * (lo + hi);
        const double step = munew - mu;
        mu = munew;
        if (newton && fabs(step) <= 3e-8 * fabs(munew)) { fpv = fma(2.0 * s3, step, fp); break; }
        if (hi - lo <= 4.5e-16 * fmax(fabs(lo), fabs(hi))) break;
    }
    p_out = p; mu_out = mu; fp_out = fpv;
}

// w_i - w_j from pole offsets
NQ_D double iesh_wdiff(const IeshSmem& S, int i, int j) {
    return (S.eps[S.pole[i]] - S.eps[S.pole[j]]) + (S.mu[i] - S.mu[j]);
}
// eigenvector entry Z[k, i] (k = 0 impurity, k >= 1 bath state k-1)
NQ_D double iesh_Z(const IeshSmem& S, int k, int i) {
    if (k == 0) return S.z0[i];
    const double d = (S.eps[S.pole[i]] - S.eps[k - 1]) + S.mu[i];   // lambda_i - eps_{k-1}
    return S.z0[i] * (S.Vb[k - 1] * S.fsc[0]) / d;
}

// Impurity of the AndersonHolstein model: h = U1 - U0, the state-independent U0, and the common scale f(x) of the couplings
// H[0,k] = Vbar_k f(x).  MiaoSubotnik: f = 1.  ErpenbeckThoss (NQCModels, external; the impurity of test/Dynamics/iesh.jl:23
// and iesh.md:85-105): f = (1-q)/2 (1 - tanh((x - xt)/at)) + q > 0.  With a common scale the matrix stays an arrowhead and
// dH/dx = h' e0 e0' + f' (e0 Vbar' + Vbar e0'); the eigen equation's first row gives Vbar . z_b = (w_i - h) z0_i / f, hence
//     (Z' dH Z)_ij = z0_i z0_j [ h' + (f'/f) (w_i + w_j - 2h) ]
// -- still an outer product with a two-term weight: force, NAC and v.d keep their O(n^2) forms.
struct IeshModel {
    int erp;
    double a[12], mass;
    NQ_D void load(const KParams& p) {
        erp = p.iesh_impurity;
        for (int i = 0; i < 12; ++i) a[i] = p.params[i];
        mass = p.masses[0];
        if (!erp) a[0] = p.params[0] * p.params[1] * p.params[1];      // m w^2
    }
    NQ_D void eval(double q, double& h, double& dh, double& u0, double& du0, double& f, double& df) const {
        if (!erp) {
            const double mw2 = a[0], g = a[2], dG = a[3];
            u0 = 0.5 * mw2 * q * q;
            const double u1 = 0.5 * mw2 * (q - g) * (q - g) + dG;
            h = u1 - u0;
            dh = mw2 * (q - g) - mw2 * q;
            du0 = mw2 * q;
            f = 1.0; df = 0.0;
        } else {
            const double e0 = exp(-a[1] * (q - a[2]));
            u0 = a[0] * (e0 - 1.0) * (e0 - 1.0) + a[3];
            du0 = -2.0 * a[0] * a[1] * e0 * (e0 - 1.0);
            const double e1 = exp(-a[6] * (q - a[7]));
            const double u1 = a[4] * e1 * e1 - a[5] * e1 + a[8];
            const double du1 = -2.0 * a[6] * a[4] * e1 * e1 + a[6] * a[5] * e1;
            h = u1 - u0; dh = du1 - du0;
            const double t = tanh((q - a[11]) / a[10]);
            f = 0.5 * (1.0 - a[9]) * (1.0 - t) + a[9];
            df = -0.5 * (1.0 - a[9]) * (1.0 - t * t) / a[10];
        }
    }
};

// Load the bath into shared memory (once per CTA).  Returns ||V||_2.
NQ_D double iesh_load_bath(const KParams& p, IeshSmem& S) {
    const int M = p.n - 1;
    double part = 0.0;
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        const double v = p.bath_b[k];
        S.eps[k] = p.bath_a[k]; S.Vb[k] = v; S.V2[k] = v * v;
        part += v * v;
    }
    return sqrt(iesh_block_sum(part, S.red));
}

// All roots for impurity level h; fills pole, mu, lam, z0 (signed with S.sgn).  Ends with a barrier.
NQ_D void iesh_eigen(const KParams& p, IeshSmem& S, double h, double f, double vnorm, bool cold) {
    const int n = p.n, M = n - 1, lr = p.iesh.lr;
    const int lane = threadIdx.x & 31;
    const int root = threadIdx.x / lr, sub = threadIdx.x % lr;
    const unsigned mask = (lr >= 32) ? 0xffffffffu : (((1u << lr) - 1u) << (lane & ~(lr - 1)));
    if (threadIdx.x == 0) { S.fsc[0] = f; S.fsc[1] = h; }      // read after the closing barrier (iesh_Z, estimators)
    vnorm *= fabs(f);
    if (root < n) {
        int pp; double mu, fp;
        const double guess = cold ? nan("") : S.lam[root];
        iesh_root(S.eps, S.V2, f * f, M, root, h, vnorm, guess, sub, lr, mask, pp, mu, fp);
        __syncwarp(mask);   // every lane of the group has read its warm start S.lam[root]
        if (sub == 0) {
            S.pole[root] = pp; S.mu[root] = mu; S.lam[root] = S.eps[pp] + mu;
            S.z0[root] = S.sgn[root] / sqrt(fp);
        }
    }
    __syncthreads();
}

// occupied flags and the ascending list of unoccupied states (DynamicsUtils.jl:162-171); thread 0 + barrier
NQ_D void iesh_refresh_unoccupied(const KParams& p, IeshSmem& S) {
    __syncthreads();
    for (int i = threadIdx.x; i < p.n; i += blockDim.x) S.flag[i] = -1;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int e = 0; e < p.ne; ++e) S.flag[S.occ[e]] = e;
        int c = 0;
        for (int i = 0; i < p.n; ++i) if (S.flag[i] < 0) S.un[c++] = i;
    }
    __syncthreads();
}

// EhrenfestNA force weight (ehrenfest_na.jl:72-90 with Z' dV Z = h' z0 z0'): sum_e |sum_n z0[n] psi[n,e]|^2.
// Warps take electrons, lanes split the states; result on every thread.
NQ_D double iesh_mean_field_weight(const KParams& p, const IeshSmem& S, const double* psi_re, const double* psi_im, int erp,
                                   double h, double& weight2) {
    // second weight (position-dependent coupling): sum_e Re conj(s_e) t_e, s_e = sum_i z0_i psi_ie, t_e = sum_i (w_i - h) z0_i psi_ie
    const int n = p.n, ne = p.ne, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double part = 0.0, part2 = 0.0;
    for (int e = warp; e < ne; e += nwarps) {
        double cr = 0.0, ci = 0.0, tr = 0.0, ti = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double z = S.z0[i], xr = psi_re[(int64_t)n * e + i], xi = psi_im[(int64_t)n * e + i];
            cr = fma(z, xr, cr);
            ci = fma(z, xi, ci);
            if (erp) { const double zw = z * (S.lam[i] - h); tr = fma(zw, xr, tr); ti = fma(zw, xi, ti); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { cr += __shfl_xor_sync(0xffffffffu, cr, o); ci += __shfl_xor_sync(0xffffffffu, ci, o); }
        if (erp) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { tr += __shfl_xor_sync(0xffffffffu, tr, o); ti += __shfl_xor_sync(0xffffffffu, ti, o); }
        }
        if (lane == 0) { part += cr * cr + ci * ci; part2 += cr * tr + ci * ti; }
    }
    weight2 = erp ? iesh_block_sum(part2, S.red) : 0.0;
    return iesh_block_sum(part, S.red);
}

NQ_D void iesh_emit(const KParams& p, int64_t traj, int isave, int obs_id, int k, double val) {
    const int64_t off = p.layout.offset[obs_id] + (int64_t)isave * p.layout.width[obs_id] + k;
    if (p.obs_traj != nullptr) p.obs_traj[off * p.ntraj + traj] = val;
    atomicAdd(&p.obs_sum[(int64_t)(blockIdx.x % kObsReplicas) * p.layout.total + off], val);
}

// Estimators at a save point (iesh.jl:337-388).  Every thread of the CTA must call it.
// zt: zt_cap doubles of scratch shared memory (the psi-chunk region is free at a save point), >= (warps + 9) * roundup(n, 4)
// Ring polymers (rp_pot >= -inf passed with p.B > 1): r, v are the centroid, S holds the centroid eigenproblem (populations
// use the centroid transformation like rt_record_save), the kinetic energy sums the beads in S.rp, rp_pot is the bead sum
// of the potential (rpiesh.jl:38-52, rpehrenfest_na.jl:37-52) and the total adds the spring term (ring_polymer.jl:89-107).
__device__ __noinline__ void iesh_record_save(const KParams& p, IeshSmem& S, int64_t traj, int isave, double r, double v,
                           const IeshModel& mdl, const double* psi_re, const double* psi_im, double* zt, int zt_cap,
                           double rp_pot = 0.0) {
    const uint32_t obs = p.observables;
    const int n = p.n, ne = p.ne, tid = threadIdx.x, nt = blockDim.x;
    const bool last = (isave == p.nsave - 1);
    const bool trans = ((p.B > 1) ? S.rp[0] : r) > 0.0;      // get_positions(final)[1]: first dof of the first bead
    if (obs & ((1u << NQCB200_OBS_ADIABATIC_POP) | (1u << NQCB200_OBS_SCATTERING))) {
        for (int i = tid; i < n; i += nt) {
            double a = (S.flag[i] >= 0) ? 1.0 : 0.0;                             // iesh.jl:371-375
            if (p.mean_field) {                                                   // ehrenfest_na.jl:116-125
                a = 0.0;
                for (int e = 0; e < ne; ++e) {
                    const double x = psi_re[(int64_t)n * e + i], y = psi_im[(int64_t)n * e + i];
                    a += x * x + y * y;
                }
            }
            if (obs & (1u << NQCB200_OBS_ADIABATIC_POP)) iesh_emit(p, traj, isave, NQCB200_OBS_ADIABATIC_POP, i, a);
            if (obs & (1u << NQCB200_OBS_SCATTERING)) {
                iesh_emit(p, traj, isave, NQCB200_OBS_SCATTERING, i, (last && !trans) ? a : 0.0);
                iesh_emit(p, traj, isave, NQCB200_OBS_SCATTERING, n + i, (last && trans) ? a : 0.0);
            }
        }
    }
    if (obs & ((1u << NQCB200_OBS_DIABATIC_POP) | (1u << NQCB200_OBS_SCATTERING_DIABATIC))) {
        // pop_i = sum_e [ (sum_a Z_ia x_ae)^2 - sum_a Z_ia^2 x_ae^2 + Z_{i,occ_e}^2 ],  x = Re psi  (iesh.jl:337-369)
        // One warp per diabatic row i: the lanes hold Z[i, a] for a = lane, lane + 32, ... (one division each), the row
        // also sits in the warp's scratch row for Z[i, occ_e]; x[., e] is read coalesced; two warp sums per (i, e).
        //   second term: sum_e sum_a Z_ia^2 x_ae^2 = sum_a Z_ia^2 s_a with s_a = sum_e x_ae^2 (O(n^2) instead of O(n^2 ne))
        constexpr int QA = 8;                                   // n <= 256 states
        constexpr int EB = 8;                                   // electrons per inner iteration (independent reduction chains)
        const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
        const int n4 = (n + 3) & ~3;
        double* zrow = zt + (size_t)warp * n4;
        double* sa = zt + (size_t)nwarps * n4;                  // s_a
        double* xs = sa + n4;                                   // Re psi of a chunk of electrons, [e][n4]
        const int ecap = max(EB, ((zt_cap - (nwarps + 1) * n4) / n4) & ~(EB - 1));
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            S.pop[i] = 0.0;
            double sq = 0.0;
            for (int e = 0; e < ne; ++e) { const double x = psi_re[(int64_t)n * e + i]; sq = fma(x, x, sq); }
            sa[i] = sq;
        }
        for (int ec0 = 0; ec0 < ne; ec0 += ecap) {
            const int ecn = min(ecap, ne - ec0);
            __syncthreads();
            for (int idx = tid; idx < ecn * n; idx += nt) {
                const int e = idx / n, a = idx - e * n;
                xs[e * n4 + a] = psi_re[(int64_t)n * (ec0 + e) + a];
            }
            __syncthreads();
            for (int i = warp; i < n; i += nwarps) {
                double zr[QA];
                double acc = 0.0;
#pragma unroll
                for (int qa = 0; qa < QA; ++qa) {
                    const int a = lane + 32 * qa;
                    zr[qa] = (a < n) ? iesh_Z(S, i, a) : 0.0;
                    if (a < n) {
                        zrow[a] = zr[qa];
                        if (ec0 == 0) acc = fma(-zr[qa] * zr[qa], sa[a], acc);
                    }
                }
                __syncwarp();      // zrow is read below by lanes that did not write the element (shuffles order execution, not memory)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                for (int eb = 0; eb < ecn; eb += EB) {
                    double y[EB];
#pragma unroll
                    for (int k = 0; k < EB; ++k) y[k] = 0.0;
#pragma unroll
                    for (int qa = 0; qa < QA; ++qa) {
                        const int a = lane + 32 * qa;
                        if (a < n) {
#pragma unroll
                            for (int k = 0; k < EB; ++k)
                                if (eb + k < ecn) y[k] = fma(zr[qa], xs[(eb + k) * n4 + a], y[k]);
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int k = 0; k < EB; ++k) y[k] += __shfl_xor_sync(0xffffffffu, y[k], o);
                    }
#pragma unroll
                    for (int k = 0; k < EB; ++k) {
                        if (eb + k < ecn) {
                            const double zo = zrow[S.occ[ec0 + eb + k]];
                            acc += y[k] * y[k] + zo * zo;
                        }
                    }
                }
                if (lane == 0) S.pop[i] += acc;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            const double d = S.pop[i];
            if (obs & (1u << NQCB200_OBS_DIABATIC_POP)) iesh_emit(p, traj, isave, NQCB200_OBS_DIABATIC_POP, i, d);
            if (obs & (1u << NQCB200_OBS_SCATTERING_DIABATIC)) {
                iesh_emit(p, traj, isave, NQCB200_OBS_SCATTERING_DIABATIC, i, (last && !trans) ? d : 0.0);
                iesh_emit(p, traj, isave, NQCB200_OBS_SCATTERING_DIABATIC, n + i, (last && trans) ? d : 0.0);
            }
        }
    }
    if (tid == 0) {
        if (obs & ((1u << NQCB200_OBS_KINETIC) | (1u << NQCB200_OBS_POTENTIAL) | (1u << NQCB200_OBS_TOTAL_ENERGY))) {
            double kin = (mdl.mass * v * v) / 2.0;
            double h, dh, u0, du0, fs, dfs;
            mdl.eval(r, h, dh, u0, du0, fs, dfs);
            double pot = u0;                                                      // iesh.jl:380-388
            double spring = 0.0;
            if (p.B > 1) {
                const double* rb = S.rp; const double* vb = S.rp + 32;
                double mv2 = 0.0, spr = 0.0, rprev = rb[p.B - 1];
                for (int b = 0; b < p.B; ++b) {
                    mv2 = fma(mdl.mass * vb[b], vb[b], mv2);
                    const double d = rprev - rb[b];
                    spr = fma(mdl.mass * d, d, spr);
                    rprev = rb[b];
                }
                kin = 0.5 * mv2; pot = rp_pot; spring = 0.5 * p.omega_n * p.omega_n * spr;
            } else if (p.mean_field) {                                            // ehrenfest_na.jl:103-114
                for (int e = 0; e < ne; ++e)
                    for (int i = 0; i < n; ++i) {
                        const double x = psi_re[(int64_t)n * e + i], y = psi_im[(int64_t)n * e + i];
                        pot += S.lam[i] * (x * x + y * y);
                    }
            } else
                for (int e = 0; e < ne; ++e) pot += S.lam[S.occ[e]];
            if (obs & (1u << NQCB200_OBS_KINETIC)) iesh_emit(p, traj, isave, NQCB200_OBS_KINETIC, 0, kin);
            if (obs & (1u << NQCB200_OBS_POTENTIAL)) iesh_emit(p, traj, isave, NQCB200_OBS_POTENTIAL, 0, pot);
            if (obs & (1u << NQCB200_OBS_TOTAL_ENERGY)) iesh_emit(p, traj, isave, NQCB200_OBS_TOTAL_ENERGY, 0, kin + pot + spring);
        }
        if (obs & (1u << NQCB200_OBS_POSITION)) iesh_emit(p, traj, isave, NQCB200_OBS_POSITION, 0, r);
        if (obs & (1u << NQCB200_OBS_VELOCITY)) iesh_emit(p, traj, isave, NQCB200_OBS_VELOCITY, 0, v);
    }
    if (obs & (1u << NQCB200_OBS_DISCRETE_STATE))
        for (int e = tid; e < ne; e += nt) iesh_emit(p, traj, isave, NQCB200_OBS_DISCRETE_STATE, e, (double)(S.occ[e] + 1));
    if (obs & (1u << NQCB200_OBS_SIGMA))
        for (int idx = tid; idx < n * ne; idx += nt) {
            iesh_emit(p, traj, isave, NQCB200_OBS_SIGMA, idx, psi_re[idx]);
            iesh_emit(p, traj, isave, NQCB200_OBS_SIGMA, n * ne + idx, psi_im[idx]);
        }
    __syncthreads();
}

// rows-done bit mask helpers (static register indexing)
NQ_D bool iesh_bit(const unsigned (&m)[4], int i) {
    unsigned w = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) w = ((i >> 5) == q) ? m[q] : w;
    return (w >> (i & 31)) & 1u;
}
NQ_D void iesh_set_bit(unsigned (&m)[4], int i) {
#pragma unroll
    for (int q = 0; q < 4; ++q) if ((i >> 5) == q) m[q] |= 1u << (i & 31);
}

// ---- FP64 tensor-core tile product -------------------------------------------------------------------
// D(8x8) += A(8x4) B(4x8), fragments per lane T: a = A[T/4][T%4], b = B[T%4][T/4], c = C[T/4][2(T%4) + {0,1}].
NQ_D void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// psi' = exp(-i (W - i G) dt) psi  (wavefunction_dynamics.jl:15-58) for every chunk of electrons:
// Horner form of the Taylor polynomial of the shifted generator A = -i dts (W - sigma) - dts G,
//     y_K+1 = psi0 ;  y_j = psi0 + (A y_j+1) / j  (j = K..1) ;  psi' = y_1 ,  psi0 = e^{-i sigma dts} psi,
// acting on the real matrix [X | Y] (columns 2e, 2e+1 = Re, Im of electron e).  Stages j > Kg drop the G part
// (its contribution to psi' is below 1e-18) and are purely diagonal.  Every thread keeps the y elements it owns in
// the DMMA accumulators for the whole polynomial: before stage j the accumulators are set to
// psi0 + (D y_j+1)/j (D = diagonal part) and shared memory holds u = -(dts/j) y_j+1, so that the tensor-core product
// c += G u leaves y_j in the accumulators.  A fragments come from G (column-major, leading dimension = 4 mod 16:
// conflict-free), B fragments from the u chunk (row-major, same rule).  Tile ownership (12 warps = 3 per SM
// sub-partition, so the tensor pipes stay balanced): warp w owns the R full rows of 8 states {w, w+12, ..} x all NT
// column tiles of the chunk; the tiles of the remaining rows are dealt out one by one ("extra" tiles, <= NX per warp).
// Returns (to every thread) the leakage  sum_{m unoccupied, e} |psi'_me|^2.
template <int R, int NT, int NX>
__device__ __noinline__ double iesh_propagate(const KParams& p, const IeshSmem& S, double* __restrict__ psi_re,
                                              double* __restrict__ psi_im, double* Gs, double* Bs, const double* Gglob,
                                              double sigma, double dts, int nsub, int K, int Kg) {
    const IeshLayout& L = p.iesh;
    const int n = p.n, ne = p.ne, tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int lq = lane >> 2, lr4 = lane & 3;
    const int n4 = (n + 3) & ~3;
    const int ecap = 4 * L.nct;                       // electrons per chunk (nct column tiles of 8 doubles)
    const double cph = cos(sigma * dts), sph = sin(sigma * dts);
    const int rem_rows = L.nrt - R * nwarps;          // rows of 8 states dealt out tile by tile
    double leak = 0.0;

    for (int sub_i = 0; sub_i < nsub; ++sub_i) {
        const bool last_sub = (sub_i == nsub - 1);
        for (int ch = 0; ch < L.nchunks; ++ch) {
            const int e0 = ch * ecap, e1 = min(ne, e0 + ecap);
            const int nt_act = (e1 - e0 + 3) / 4;
            int xrow[NX], xcol[NX];
            bool xok[NX];
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                const int g = warp + nwarps * x;
                xok[x] = g < rem_rows * nt_act;
                xrow[x] = xok[x] ? R * nwarps + g / nt_act : 0;
                xcol[x] = xok[x] ? g % nt_act : 0;
            }
            double c[R > 0 ? R : 1][NT][2], cx[NX][2];       // y (DMMA accumulators)
            double q[R > 0 ? R : 1][NT][2], qx[NX][2];       // psi0 = e^{-i sigma dts} psi of the same elements
            double wr[R > 0 ? R : 1], wx[NX];                // w_i - sigma of the owned rows
            // f(row i, electron e, column tile t, y.re, y.im, psi0.re, psi0.im, ws) on every element pair this thread owns
            // which of its element pairs this thread really owns (inside the matrix, inside the chunk): one bit each,
            // evaluated once per chunk instead of three compares per element per stage
            unsigned own[R > 0 ? R : 1], ownx = 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                own[r] = 0u;
                const int i = 8 * (warp + r * nwarps) + lq;
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    if (t < nt_act && i < n && e0 + 4 * t + lr4 < e1) own[r] |= 1u << t;
            }
#pragma unroll
            for (int x = 0; x < NX; ++x)
                if (xok[x] && 8 * xrow[x] + lq < n && e0 + 4 * xcol[x] + lr4 < e1) ownx |= 1u << x;
            auto for_own = [&](auto&& f) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int i = 8 * (warp + r * nwarps) + lq;
#pragma unroll
                    for (int t = 0; t < NT; ++t)
                        if ((own[r] >> t) & 1u) f(i, e0 + 4 * t + lr4, t, c[r][t][0], c[r][t][1], q[r][t][0], q[r][t][1], wr[r]);
                }
#pragma unroll
                for (int x = 0; x < NX; ++x)
                    if ((ownx >> x) & 1u) f(8 * xrow[x] + lq, e0 + 4 * xcol[x] + lr4, xcol[x], cx[x][0], cx[x][1], qx[x][0], qx[x][1], wx[x]);
            };
            // accumulators <- psi0 + ck (ws Y, -ws X) of the current y
            auto diag_stage = [&](double ck) {
                for_own([&](int, int, int, double& yr, double& yi, double& pr, double& pi, double& w) {
                    const double nr = fma(ck * w, yi, pr), ni = fma(-ck * w, yr, pi);
                    yr = nr; yi = ni;
                });
            };
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = 8 * (warp + r * nwarps) + lq;
                wr[r] = (i < n) ? S.ws[i] : 0.0;
#pragma unroll
                for (int t = 0; t < NT; ++t) { c[r][t][0] = 0.0; c[r][t][1] = 0.0; q[r][t][0] = 0.0; q[r][t][1] = 0.0; }
            }
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                const int i = 8 * xrow[x] + lq;
                wx[x] = (i < n) ? S.ws[i] : 0.0;
                cx[x][0] = 0.0; cx[x][1] = 0.0; qx[x][0] = 0.0; qx[x][1] = 0.0;
            }
            for_own([&](int i, int e, int, double& yr, double& yi, double& pr, double& pi, double&) {
                const double a = psi_re[i + (int64_t)n * e], b = psi_im[i + (int64_t)n * e];
                pr = a * cph + b * sph; pi = b * cph - a * sph;
                yr = pr; yi = pi;                                  // y_K+1 = psi0
            });
            for (int j = K; j > Kg; --j) diag_stage(dts / j);      // diagonal-only stages
            if (Kg >= 1) {
                // rows / columns of the u chunk that nobody owns must be zero: the k padding rows and the electron padding
                for (int idx = tid; idx < (n4 - n) * L.ldb; idx += nt) Bs[n * L.ldb + idx] = 0.0;
                if (e1 - e0 < 4 * nt_act)
                    for (int idx = tid; idx < n * 8; idx += nt) Bs[(idx >> 3) * L.ldb + 8 * (nt_act - 1) + (idx & 7)] = 0.0;
                __syncthreads();
                for (int j = Kg; j >= 1; --j) {
                    // u = -(dts/j) y_j+1 -> shared ; accumulators = psi0 + (D y_j+1) / j
                    const double sj = -dts / j;
                    for_own([&](int i, int, int t, double& yr, double& yi, double&, double&, double&) {
                        *reinterpret_cast<double2*>(Bs + i * L.ldb + 8 * t + 2 * lr4) = make_double2(sj * yr, sj * yi);
                    });
                    diag_stage(dts / j);
                    __syncthreads();
                    // one k-slab [k0, k1) of c += G u, G columns read from gbase (column k0 at offset 0)
                    auto slab = [&](const double* gbase, int k0, int k1) {
                        const double* gp = gbase + lq + L.ldg * lr4;
                        const double* bp = Bs + lr4 * L.ldb + lq;
#pragma unroll 2
                        for (int k = k0; k < k1; k += 4) {
                            const double* gk = gp + (k - k0) * L.ldg;
                            const double* bk = bp + k * L.ldb;
                            double a[R > 0 ? R : 1];
#pragma unroll
                            for (int r = 0; r < R; ++r) a[r] = gk[8 * (warp + r * nwarps)];
#pragma unroll
                            for (int t = 0; t < NT; ++t) {
                                if (R > 0 && t < nt_act) {
                                    const double b = bk[8 * t];
#pragma unroll
                                    for (int r = 0; r < R; ++r) dmma884(c[r][t][0], c[r][t][1], a[r], b);
                                }
                            }
#pragma unroll
                            for (int x = 0; x < NX; ++x)
                                if (xok[x]) dmma884(cx[x][0], cx[x][1], gk[8 * xrow[x]], bk[8 * xcol[x]]);
                        }
                    };
                    if (L.resident) {
                        slab(Gs, 0, n4);
                        __syncthreads();
                    } else {
                        const int pieces = L.ldg * L.kb / 2;   // 16-byte pieces per slab
                        for (int cc = tid; cc < pieces; cc += nt) cp_async16(Gs + 2 * cc, Gglob + 2 * cc);
                        cp_async_commit();
                        for (int s = 0; s < L.nslab; ++s) {
                            if (s + 1 < L.nslab) {
                                const double* src = Gglob + (int64_t)(s + 1) * L.ldg * L.kb;
                                double* dst = Gs + ((s + 1) & 1) * L.ldg * L.kb;
                                for (int cc = tid; cc < pieces; cc += nt) cp_async16(dst + 2 * cc, src + 2 * cc);
                            }
                            cp_async_commit();
                            cp_async_wait<1>();
                            __syncthreads();
                            slab(Gs + (s & 1) * L.ldg * L.kb, s * L.kb, min(n4, (s + 1) * L.kb));
                            __syncthreads();
                        }
                    }
                }
            }
            // y_1 = psi' of this chunk: every thread stores the elements it owns (+ leakage out of the occupied orbitals)
            for_own([&](int i, int e, int, double& yr, double& yi, double&, double&, double&) {
                psi_re[i + (int64_t)n * e] = yr; psi_im[i + (int64_t)n * e] = yi;
                if (last_sub && S.flag[i] < 0) leak += yr * yr + yi * yi;
            });
            __syncthreads();
        }
    }
    return iesh_block_sum(leak, S.red);
}

// det S for S[j][i] = psi[occ_j, i] (iesh.jl:309-316; FastDeterminant.det!, FastDeterminant.jl:20-22): LU with
// partial (row) pivoting on a 16 x (threads/16) thread grid with cyclic distribution.  Each thread owns RL x CL
// elements, held in REGISTERS (REG = true, ne <= 52) or in its own shared-memory slots (REG = false); the pivot
// column and pivot row travel through shared memory (double buffered: two barriers per column).  Rows are never
// swapped (implicit pivoting); the permutation sign is accumulated from the rank of each pivot among the
// remaining rows.   buf: 8 * roundup(ne, 4) doubles ; store (REG = false): 2 * ne * lds doubles.
template <int RL, int CL, bool REG>
__device__ __noinline__ void iesh_det_lu(const KParams& p, const IeshSmem& S, const double* __restrict__ psi_re,
                      const double* __restrict__ psi_im, double* buf, double* store, double& det_re, double& det_im) {
    const int n = p.n, ne = p.ne, tid = threadIdx.x, nt = blockDim.x;
    const int PC = nt >> 4;
    const int tr = tid & 15, tc = tid >> 4, lane = tid & 31;
    double are[REG ? RL : 1][REG ? CL : 1], aim[REG ? RL : 1][REG ? CL : 1];
    const int lds = p.iesh.lds;                       // REG = false: S in the standard [ne][lds] layout
    double* sre = store;
    double* sim = store + ne * lds;
    auto slot = [&](int a, int b) { return (tr + 16 * a) * lds + tc + PC * b; };
#pragma unroll
    for (int a = 0; a < RL; ++a)
#pragma unroll
        for (int b = 0; b < CL; ++b) {
            const int i = tr + 16 * a, j = tc + PC * b;
            const bool ok = (i < ne && j < ne);
            const double vr = ok ? psi_re[S.occ[i] + (int64_t)n * j] : 0.0;
            const double vi = ok ? psi_im[S.occ[i] + (int64_t)n * j] : 0.0;
            if (REG) { are[a][b] = vr; aim[a][b] = vi; } else if (ok) { sre[slot(a, b)] = vr; sim[slot(a, b)] = vi; }
        }
    const int nep = (ne + 3) & ~3;
    double* cb_re = buf; double* cb_im = buf + 2 * nep; double* rb_re = buf + 4 * nep; double* rb_im = buf + 6 * nep;
    unsigned done[4] = {0u, 0u, 0u, 0u};
    double dr = 1.0, di = 0.0;
    for (int k = 0; k < ne; ++k) {
        const int par = (k & 1) * nep;
        if (tc == k % PC) {
            const int bk = k / PC;
#pragma unroll
            for (int a = 0; a < RL; ++a) {
                const int i = tr + 16 * a;
                if (i < ne) {
                    double vr = 0.0, vi = 0.0;
                    if (REG) {
#pragma unroll
                        for (int b = 0; b < CL; ++b) if (b == bk) { vr = are[a][b]; vi = aim[a][b]; }
                    } else { vr = sre[slot(a, bk)]; vi = sim[slot(a, bk)]; }
                    cb_re[par + i] = vr; cb_im[par + i] = vi;
                }
            }
        }
        __syncthreads();
        // pivot: largest |S[i][k]| among the rows not used yet (ties: smallest i), found redundantly by every warp
        double best = -1.0; int bi = 0;
        for (int i = lane; i < ne; i += 32) {
            const bool dn = iesh_bit(done, i);
            const double m2 = cb_re[par + i] * cb_re[par + i] + cb_im[par + i] * cb_im[par + i];
            if (!dn && m2 > best) { best = m2; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        const int pr = bi;
        if (tr == (pr & 15)) {
            const int ap = pr >> 4;
#pragma unroll
            for (int b = 0; b < CL; ++b) {
                const int j = tc + PC * b;
                if (j < ne) {
                    double vr = 0.0, vi = 0.0;
                    if (REG) {
#pragma unroll
                        for (int a = 0; a < RL; ++a) if (a == ap) { vr = are[a][b]; vi = aim[a][b]; }
                    } else { vr = sre[slot(ap, b)]; vi = sim[slot(ap, b)]; }
                    rb_re[par + j] = vr; rb_im[par + j] = vi;
                }
            }
        }
        __syncthreads();
        const double pre = cb_re[par + pr], pim = cb_im[par + pr];
        const double pm2 = pre * pre + pim * pim;
        const double ire = pre / pm2, iim = -pim / pm2;
        int rank = 0;   // rank of pr among the remaining rows -> sign of the permutation
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const unsigned remaining = ~done[w];
            if (32 * w + 31 < pr) rank += __popc(remaining);
            else if (32 * w <= pr) rank += __popc(remaining & ((1u << (pr & 31)) - 1u));
        }
        {
            const double nr = dr * pre - di * pim, ni = dr * pim + di * pre;
            dr = (rank & 1) ? -nr : nr; di = (rank & 1) ? -ni : ni;
        }
        iesh_set_bit(done, pr);
#pragma unroll
        for (int a = 0; a < RL; ++a) {
            const int i = tr + 16 * a;
            const bool dn = iesh_bit(done, i);     // includes pr itself
            if (i < ne && !dn) {
                const double cr = cb_re[par + i], ci = cb_im[par + i];
                const double lre = cr * ire - ci * iim, lim = cr * iim + ci * ire;
#pragma unroll
                for (int b = 0; b < CL; ++b) {
                    const int j = tc + PC * b;
                    if (j > k && j < ne) {
                        const double ur = rb_re[par + j], ui = rb_im[par + j];
                        if (REG) {
                            are[a][b] -= lre * ur - lim * ui;
                            aim[a][b] -= lre * ui + lim * ur;
                        } else {
                            sre[slot(a, b)] -= lre * ur - lim * ui;
                            sim[slot(a, b)] -= lre * ui + lim * ur;
                        }
                    }
                }
            }
        }
    }
    det_re = dr; det_im = di;
}

// TerminatingCallback predicate (nqcb200_set_termination): outside the position window [, moving outwards]
// [, or t > tcut]; `steps_done` = steps the trajectory has completed, t = t0 + dt * steps_done
NQ_D bool iesh_outside(const KParams& p, double r, double v, int64_t steps_done) {
    return (r < p.term_lo && (!p.term_outgoing || v < 0.0)) || (r > p.term_hi && (!p.term_outgoing || v > 0.0)) ||
           p.t0 + p.dt * (double)steps_done > p.term_tcut;
}

// Eigen-decomposition at the frozen position of a trajectory that terminated in an earlier launch (cold path).
__device__ __noinline__ void iesh_eigen_frozen(const KParams& p, IeshSmem& S, const IeshModel& mdl, double r, double vnorm) {
    double h, dh, u0, du0, fs, dfs;
    mdl.eval(r, h, dh, u0, du0, fs, dfs);
    iesh_eigen(p, S, h, fs, vnorm, false);
}

// ---- ring polymers: RingPolymerSimulation{AdiabaticIESH} / {EhrenfestNA} with BCBWavefunction ---------------------
// (rpiesh.jl:21-52, rpehrenfest_na.jl:13-52, bcb_wavefunction.jl:37-69).  The CTA still owns one trajectory; its beads
// (positions, velocities, accelerations: CTA-uniform scalars) sit in S.rp.  Per step: half kick + free ring-polymer
// step (dense normal-mode product on warp 0), one arrowhead eigenproblem per BEAD for the bead force, psi propagated
// with the centroid generator of the PREVIOUS geometry and velocity (quirk Q5: propagate_wavefunction!(.., vprev, rprev, ..),
// bcb_wavefunction.jl:67 -- the centroid eigenproblem and G = v.d built at the end of the previous step are exactly that
// generator, so it is kept in S.ws / G across the bead solves), then the centroid eigenproblem at the new geometry for
// the hop test (SurfaceHoppingMethods.jl:85-103: centroid eigenvalues, couplings and velocity; the rescaling moves
// every bead by the same amount, rpsh.jl:30-50).  All geometries share the column signs S.sgn: for the arrowhead
// matrix the sign that "continuity with the identity" picks is sign(V_{i-1}) whatever the impurity level.

// acceleration of the geometry whose eigenproblem is in S (iesh.jl:190-207 / ehrenfest_na.jl:72-90); every thread
NQ_D double iesh_acceleration(const KParams& p, const IeshSmem& S, const IeshModel& mdl, const double* psi_re,
                              const double* psi_im, double h, double dh, double du0, double phi) {
    double w1, w2 = 0.0;
    if (p.mean_field) w1 = iesh_mean_field_weight(p, S, psi_re, psi_im, mdl.erp, h, w2);
    else {
        double part = 0.0, part2 = 0.0;
        for (int e = threadIdx.x; e < p.ne; e += blockDim.x) {
            const int o = S.occ[e];
            const double z = S.z0[o];
            part += z * z;
            if (mdl.erp) part2 = fma(z * z, S.lam[o] - h, part2);
        }
        w1 = iesh_block_sum(part, S.red);
        if (mdl.erp) w2 = iesh_block_sum(part2, S.red);
    }
    return ((-du0 - dh * w1) - 2.0 * phi * w2) / mdl.mass;
}

// G = v.d of the eigenproblem in S into Gdst (same element formula as the step kernel's build), shifted eigenvalues
// into S.ws; returns the shift, the half spectral width, ||G||_F and sum |v.d[m, e]| (unoccupied m, occupied e)
NQ_D void iesh_build_generator(const KParams& p, IeshSmem& S, const IeshModel& mdl, double* Gdst, double h, double dh,
                               double phi, double v, double& sigma, double& wspan, double& gnorm, double& sabs_out) {
    const int n = p.n, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int ldg = p.iesh.ldg;
    const double wmin = S.lam[0], wmax = S.lam[n - 1];
    sigma = 0.5 * (wmin + wmax);
    wspan = 0.5 * (wmax - wmin);
    __syncthreads();                 // the previous generator's S.ws / G have been consumed
    for (int i = tid; i < n; i += nt) S.ws[i] = S.lam[i] - sigma;
    const double gpref = -v * dh;
    double g2 = 0.0, sabs = 0.0;
    for (int j = warp; j < n; j += nwarps) {
        const double zj = S.z0[j] * gpref, ej = S.eps[S.pole[j]], mj = S.mu[j];
        const double zje = -v * S.z0[j] * phi, wj = S.lam[j] - h;
        const bool occ_j = S.flag[j] >= 0;
        for (int i = lane; i < n; i += 32) {
            double g = 0.0;
            if (i != j) {
                const double num = mdl.erp ? fma(zje, (S.lam[i] - h) + wj, zj) : zj;
                g = num * S.z0[i] * iesh_rcp((S.eps[S.pole[i]] - ej) + (S.mu[i] - mj));
                if (occ_j && S.flag[i] < 0) sabs += fabs(g);
            }
            g2 = fma(g, g, g2);
            Gdst[i + (int64_t)ldg * j] = g;
        }
    }
    gnorm = sqrt(iesh_block_sum(g2, S.red));
    sabs_out = iesh_block_sum(sabs, S.red);
}

// Taylor plan of the propagator (same rule as the step kernel): nsub sub-steps, K stages, the first Kg of them with G
NQ_D void iesh_taylor_plan(double dt, double wspan, double gnorm, int& nsub, double& dts, int& K, int& Kg) {
    const double rho_full = dt * (wspan + gnorm);
    nsub = max(1, (int)ceil(rho_full / 4.0));
    dts = dt / nsub;
    const double rho = rho_full / nsub, rho_g = dts * gnorm;
    K = 1; Kg = (rho_g >= 1e-18) ? 1 : 0;
    double term = rho;
    while (term > 1e-17 && K < 90) {
        ++K;
        if (rho_g * term / K >= 1e-18) Kg = K;
        term *= rho / K;
    }
}

// B (half kick) + C (free ring polymer) of BCBWavefunction on the beads in S.rp: dense U' . Cayley . U on warp 0
// (RingPolymerArrays transform!, steps.jl:10-17; the engine's nm_to / cayley tables, full step).  Ends with a barrier.
NQ_D void iesh_rp_free_step(const KParams& p, IeshSmem& S, double hdt) {
    const int B = p.B, lane = threadIdx.x & 31;
    double* rb = S.rp; double* vb = S.rp + 32; double* ab = S.rp + 64; double* tn = S.rp + 96;
    if (threadIdx.x < 32) {
        if (lane < B) vb[lane] = fma(hdt, ab[lane], vb[lane]);
        __syncwarp();
        if (lane < B) {
            double a = 0.0, c = 0.0;
            for (int j = 0; j < B; ++j) { const double u = p.nm_to[j * B + lane]; a = fma(u, rb[j], a); c = fma(u, vb[j], c); }
            tn[lane] = p.cayley[4 * lane + 0] * a + p.cayley[4 * lane + 1] * c;
            tn[32 + lane] = p.cayley[4 * lane + 2] * a + p.cayley[4 * lane + 3] * c;
        }
        __syncwarp();
        if (lane < B) {
            double a = 0.0, c = 0.0;
            for (int k = 0; k < B; ++k) { const double u = p.nm_to[lane * B + k]; a = fma(u, tn[k], a); c = fma(u, tn[32 + k], c); }
            rb[lane] = a; vb[lane] = c;
        }
    }
    __syncthreads();
}

// bead sum of the potential (rpiesh.jl:38-52 / rpehrenfest_na.jl:37-52).  Overwrites S with the last bead's
// eigenproblem: the caller re-solves the centroid afterwards.  Every thread gets the sum.
NQ_D double iesh_rp_potential(const KParams& p, IeshSmem& S, const IeshModel& mdl, const double* psi_re,
                              const double* psi_im, double vnorm) {
    const int n = p.n, ne = p.ne, tid = threadIdx.x, nt = blockDim.x;
    double pot = 0.0;
    for (int b = 0; b < p.B; ++b) {
        double h, dh, u0, du0, fs, dfs;
        mdl.eval(S.rp[b], h, dh, u0, du0, fs, dfs);
        iesh_eigen(p, S, h, fs, vnorm, false);
        double part = 0.0;
        if (p.mean_field) {
            for (int idx = tid; idx < n * ne; idx += nt) {
                const int i = idx % n;
                part = fma(S.lam[i], psi_re[idx] * psi_re[idx] + psi_im[idx] * psi_im[idx], part);
            }
        } else
            for (int e = tid; e < ne; e += nt) part += S.lam[S.occ[e]];
        pot += u0 + iesh_block_sum(part, S.red);
    }
    return pot;
}

// centroid of the beads in S.rp (same summation order on every thread)
NQ_D void iesh_rp_centroid(const KParams& p, const IeshSmem& S, double& r, double& v) {
    double sr = 0.0, sv = 0.0;
    for (int b = 0; b < p.B; ++b) { sr += S.rp[b]; sv += S.rp[32 + b]; }
    r = sr / p.B; v = sv / p.B;
}

// ---- the step kernel -----------------------------------------------------------------------------
// RP = false: Simulation{AdiabaticIESH / EhrenfestNA} (VerletwithElectronics); RP = true: the ring-polymer variants
// (BCBWavefunction).  The RP = false instantiation contains none of the ring-polymer code.
template <bool RP>
__global__ void __launch_bounds__(384, 1) iesh_step_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) double iesh_sm[];
    IeshSmem S;
    S.carve(iesh_sm, p.n);
    const IeshLayout& L = p.iesh;
    const int n = p.n, ne = p.ne, tid = threadIdx.x, nt = blockDim.x;
    const int nun = n - ne;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int n4 = (n + 3) & ~3;
    const double dt = p.dt, hdt = 0.5 * p.dt;
    IeshModel mdl;
    mdl.load(p);
    const double vnorm = iesh_load_bath(p, S);

    double* Gs = S.work;                                   // resident: ldg x n4 ; streaming: 2 slabs of ldg x kb
    double* Bs = S.work + L.off_b;                         // y chunk, row-major [n4][ldb]
    double* Hs = S.work + L.off_hop;                       // hop phase
    double* Gglob = L.resident ? nullptr : p.iesh_G + (int64_t)blockIdx.x * L.ldg * L.kb * L.nslab;
    if (L.resident) for (int idx = tid; idx < L.ldg * n4; idx += nt) Gs[idx] = 0.0;
    unsigned long long nhops = 0, nfrus = 0, nunpruned = 0, ndet = 0, nstages = 0, ngemm = 0;

    for (int64_t traj = blockIdx.x; traj < p.ntraj; traj += gridDim.x) {
        double* psi_re = p.sig_re + traj * (int64_t)n * ne;
        double* psi_im = p.sig_im + traj * (int64_t)n * ne;
        __syncthreads();
        for (int i = tid; i < n; i += nt) { S.lam[i] = p.iesh_lam[traj * n + i]; S.sgn[i] = p.iesh_sgn[traj * n + i]; }
        for (int e = tid; e < ne; e += nt) S.occ[e] = p.state[traj * ne + e];
        iesh_refresh_unoccupied(p, S);
        double r = p.r[traj], v = p.v[traj], acc = p.acc[traj];
        const bool orth_ok = p.iesh_orth[traj] != 0.0;
        // RP: generator of the next propagation = centroid eigenproblem and G = v.d of the current geometry and velocity
        double g_sigma = 0.0, g_wspan = 0.0, g_gnorm = 0.0, g_sabs = 0.0;
        if constexpr (RP) {
            if (tid < p.B) {
                S.rp[tid] = p.r[(int64_t)tid * p.ntraj + traj];
                S.rp[32 + tid] = p.v[(int64_t)tid * p.ntraj + traj];
                S.rp[64 + tid] = p.acc[(int64_t)tid * p.ntraj + traj];
            }
            __syncthreads();
            iesh_rp_centroid(p, S, r, v);
            double h, dh, u0, du0, fs, dfs;
            mdl.eval(r, h, dh, u0, du0, fs, dfs);
            iesh_eigen(p, S, h, fs, vnorm, false);
            iesh_build_generator(p, S, mdl, L.resident ? Gs : Gglob, h, dh, dfs / fs, v, g_sigma, g_wspan, g_gnorm, g_sabs);
        }

#pragma unroll 1
        for (int is = 0; is < p.nsteps; ++is) {
            const int64_t step = p.step0 + is;
            // TerminatingCallback mask (nqcb200_set_termination).  Stateless: a terminated trajectory is frozen, so the
            // predicate that ended it still holds on its (r, v) (and t > tcut stays true); DiffEq does not test the condition at t0, hence
            // step > 0.  r and v are uniform over the CTA, so the branch is too.
            if (p.term_dof >= 0 && step > 0 && iesh_outside(p, r, v, step)) {
                if (is == 0) iesh_eigen_frozen(p, S, mdl, r, vnorm);   // terminated in an earlier launch: eigenvectors for the saves
                goto save_point;
            }
            {
            double h, dh, u0, du0, fs, dfs, phi, sabs, sigma, dts;
            int nsub, K, Kg;
            if constexpr (!RP) {
            // ---- nuclei + eigen + force (verlet_with_electronics.jl:55-66) -------------------------
            const double vt = fma(hdt, acc, v);
            r = fma(dt, vt, r);
            mdl.eval(r, h, dh, u0, du0, fs, dfs);
            iesh_eigen(p, S, h, fs, vnorm, false);
            phi = dfs / fs;              // f'/f: (Z' dH Z)_ij = z0_i z0_j [h' + phi (w_i + w_j - 2h)]
            if (p.mean_field) {
                double wsum2 = 0.0;
                const double wsum = iesh_mean_field_weight(p, S, psi_re, psi_im, mdl.erp, h, wsum2);  // sigma_prev: psi before this step's propagation
                acc = ((-du0 - dh * wsum) - 2.0 * phi * wsum2) / mdl.mass;
            } else {
                double part = 0.0, part2 = 0.0;
                for (int e = tid; e < ne; e += nt) {
                    const int o = S.occ[e];
                    const double z = S.z0[o];
                    part += z * z;
                    if (mdl.erp) part2 = fma(z * z, S.lam[o] - h, part2);
                }
                const double occsum = iesh_block_sum(part, S.red);
                const double occsum2 = mdl.erp ? iesh_block_sum(part2, S.red) : 0.0;
                acc = ((-du0 - dh * occsum) - 2.0 * phi * occsum2) / mdl.mass;      // iesh.jl:190-207
            }
            v = fma(hdt, acc, vt);

            // ---- G = v.d, shift, norms ------------------------------------------------------------
            const double wmin = S.lam[0], wmax = S.lam[n - 1];
            sigma = 0.5 * (wmin + wmax);
            for (int i = tid; i < n; i += nt) S.ws[i] = S.lam[i] - sigma;
            const double gpref = -v * dh;
            double g2 = 0.0;
            sabs = 0.0;
            {
                double* Gdst = L.resident ? Gs : Gglob;
                // thread -> (row i fastest); one division per element, w_i - w_j from pole offsets
                for (int j = warp; j < n; j += nwarps) {
                    const double zj = S.z0[j] * gpref, ej = S.eps[S.pole[j]], mj = S.mu[j];
                    const double zje = -v * S.z0[j] * phi, wj = S.lam[j] - h;      // S.lam is final (barrier of iesh_eigen); S.ws is not yet
                    const bool occ_j = S.flag[j] >= 0;
                    for (int i = lane; i < n; i += 32) {
                        double g = 0.0;
                        if (i != j) {
                            const double num = mdl.erp ? fma(zje, (S.lam[i] - h) + wj, zj) : zj;
                            g = num * S.z0[i] * iesh_rcp((S.eps[S.pole[i]] - ej) + (S.mu[i] - mj));
                            if (occ_j && S.flag[i] < 0) sabs += fabs(g);         // |v_dot_d[m, e]|, iesh.jl:285-298
                        }
                        g2 = fma(g, g, g2);
                        Gdst[i + (int64_t)L.ldg * j] = g;
                    }
                }
            }
            const double gnorm = sqrt(iesh_block_sum(g2, S.red));
            sabs = iesh_block_sum(sabs, S.red);
            const double wspan = 0.5 * (wmax - wmin);
            // Taylor plan: nsub sub-steps of dt/nsub; K stages (remainder < 1e-17), the first Kg of them with G
            const double rho_full = dt * (wspan + gnorm);
            nsub = max(1, (int)ceil(rho_full / 4.0));
            dts = dt / nsub;
            const double rho = rho_full / nsub, rho_g = dts * gnorm;
            K = 1; Kg = (rho_g >= 1e-18) ? 1 : 0;
            {
                double term = rho;                       // rho^K / K!
                while (term > 1e-17 && K < 90) {
                    ++K;
                    if (rho_g * term / K >= 1e-18) Kg = K;   // stage K contributes rho_g rho^(K-1) / K!
                    term *= rho / K;
                }
            }
            } else {
                // ---- BCBWavefunction (bcb_wavefunction.jl:49-66): B, C, update_cache! on every bead, acceleration!, B
                iesh_rp_free_step(p, S, hdt);
                for (int b = 0; b < p.B; ++b) {
                    double hb, dhb, u0b, du0b, fsb, dfsb;
                    mdl.eval(S.rp[b], hb, dhb, u0b, du0b, fsb, dfsb);
                    iesh_eigen(p, S, hb, fsb, vnorm, false);
                    const double ab = iesh_acceleration(p, S, mdl, psi_re, psi_im, hb, dhb, du0b, dfsb / fsb);   // psi: sigma_prev
                    if (tid == 0) { S.rp[64 + b] = ab; S.rp[32 + b] = fma(hdt, ab, S.rp[32 + b]); }
                }
                // propagate_wavefunction!(.., vprev, rprev, ..): the generator left by the previous step (Q5)
                sigma = g_sigma;
                iesh_taylor_plan(dt, g_wspan, g_gnorm, nsub, dts, K, Kg);
            }
            nstages += (tid == 0) ? (unsigned long long)K * nsub : 0ull;
            ngemm += (tid == 0) ? (unsigned long long)Kg * nsub : 0ull;
            __syncthreads();

            double leak;
            if (L.rounds == 0) leak = iesh_propagate<0, 1, 2>(p, S, psi_re, psi_im, Gs, Bs, Gglob, sigma, dts, nsub, K, Kg);
            else if (L.rounds == 1) leak = iesh_propagate<1, 14, 2>(p, S, psi_re, psi_im, Gs, Bs, Gglob, sigma, dts, nsub, K, Kg);
            else leak = iesh_propagate<2, 8, 2>(p, S, psi_re, psi_im, Gs, Bs, Gglob, sigma, dts, nsub, K, Kg);
            __syncthreads();
            if constexpr (RP) {
                // hopping quantities of a ring polymer: centroid eigenvalues, couplings, velocity (SurfaceHoppingMethods.jl:85-103)
                iesh_rp_centroid(p, S, r, v);
                mdl.eval(r, h, dh, u0, du0, fs, dfs);
                phi = dfs / fs;
                iesh_eigen(p, S, h, fs, vnorm, false);
                iesh_build_generator(p, S, mdl, L.resident ? Gs : Gglob, h, dh, phi, v, g_sigma, g_wspan, g_gnorm, g_sabs);
                sabs = g_sabs;
            }
            const double v_before_hop = v;

            // ---- IESHCallback: hop test (iesh.jl:231-335,390-407) ----------------------------------
            if (!p.disable_hopping) {
                const double xi = (p.rng == NQCB200_RNG_INJECTED)
                                      ? p.draws[(step - p.draws_step0) * p.ntraj + traj]
                                      : philox_uniform(p.seed, (uint64_t)(p.traj_offset + traj), (uint64_t)step, 0u);
                // Pruning without the determinant: for orthonormal orbitals |det S|^2 = det(I - U'U) >= 1 - tr U'U = 1 - leak
                // (U = unoccupied rows of psi), and |Re det| + |Im det| <= sqrt(2) |det|, hence
                //   estimate <= 2 sqrt(2) dt sum|v.d| / sqrt(1 - leak);  if even this bound is below xi the reference's
                // test (iesh.jl:251-254) prunes the step and det S is not needed at all.
                bool certainly_pruned = false;
                if (p.estimate_probability && orth_ok && p.edc_C <= 0.0) {
                    const double d2lo = (1.0 - leak) * (1.0 - 1e-9) - 1e-9;
                    if (d2lo > 0.0) certainly_pruned = 2.8284271247461903 * dt * sabs / sqrt(d2lo) * (1.0 + 1e-9) < xi;
                }
                if (!certainly_pruned) {
                ndet += (tid == 0);
                double det_re, det_im;
                const int nep8 = 8 * ((ne + 3) & ~3);
                if (ne <= 64) iesh_det_lu<4, 3, true>(p, S, psi_re, psi_im, Hs, Hs + nep8, det_re, det_im);      // 16 x 24 threads
                else iesh_det_lu<7, 5, false>(p, S, psi_re, psi_im, Hs, Hs + nep8, det_re, det_im);
                const double Akk = det_re * det_re + det_im * det_im;
                const double prefactor = 2.0 * dt / Akk;
                bool pruned = false;
                if (p.estimate_probability) {                                     // iesh.jl:251-254
                    const double estimate = prefactor * sabs * (fabs(det_re) + fabs(det_im));
                    pruned = estimate < xi;
                }
                __syncthreads();
                if (!pruned) {
                    nunpruned += (tid == 0);
                    // rare path: all ne (n - ne) probabilities from S^-1 (matrix-determinant lemma):
                    //   det S_{e->m} / det S = (psi_m . S^-1)_e ;  g = 2 dt Re(.) v_dot_d[m, e]
                    const int lds = L.lds;
                    double* Sre = Hs; double* Sim = Sre + ne * lds;
                    double* rk_re = Sim + ne * lds; double* rk_im = rk_re + ne;
                    double* ro_re = rk_im + ne; double* ro_im = ro_re + ne;
                    double* ck_re = ro_im + ne; double* ck_im = ck_re + ne;
                    double* sum_e = ck_im + ne; double* probs = sum_e + ne;
                    int* perm = (int*)(probs + n); int* srcc = perm + ne;
                    for (int idx = tid; idx < ne * ne; idx += nt) {
                        const int j = idx % ne, i = idx / ne;
                        Sre[j * lds + i] = psi_re[S.occ[j] + (int64_t)n * i];
                        Sim[j * lds + i] = psi_im[S.occ[j] + (int64_t)n * i];
                    }
                    __syncthreads();
                    // in-place Gauss-Jordan inversion with row pivoting
                    for (int k = 0; k < ne; ++k) {
                        if (warp == 0) {
                            double best = -1.0; int bi = k;
                            for (int i = k + lane; i < ne; i += 32) {
                                const double m2 = Sre[i * lds + k] * Sre[i * lds + k] + Sim[i * lds + k] * Sim[i * lds + k];
                                if (m2 > best) { best = m2; bi = i; }
                            }
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                            }
                            if (lane == 0) S.ctl[0] = bi;
                        }
                        __syncthreads();
                        const int pr = S.ctl[0];
                        const double pre = Sre[pr * lds + k], pim = Sim[pr * lds + k];
                        const double pm2 = pre * pre + pim * pim;
                        const double ire = pre / pm2, iim = -pim / pm2;          // 1 / pivot
                        for (int j = tid; j < ne; j += nt) {
                            const double xr = Sre[pr * lds + j], xi_ = Sim[pr * lds + j];
                            rk_re[j] = (j == k) ? ire : xr * ire - xi_ * iim;
                            rk_im[j] = (j == k) ? iim : xr * iim + xi_ * ire;
                            ro_re[j] = Sre[k * lds + j]; ro_im[j] = Sim[k * lds + j];
                            ck_re[j] = Sre[j * lds + k]; ck_im[j] = Sim[j * lds + k];
                        }
                        __syncthreads();
                        for (int idx = tid; idx < ne * ne; idx += nt) {
                            const int i = idx / ne, j = idx % ne;
                            double nr, ni;
                            if (i == k) { nr = rk_re[j]; ni = rk_im[j]; }
                            else {
                                const bool sw = (i == pr);
                                const double cr = sw ? ro_re[k] : ck_re[i], ci = sw ? ro_im[k] : ck_im[i];
                                double br = sw ? ro_re[j] : Sre[i * lds + j], bi_ = sw ? ro_im[j] : Sim[i * lds + j];
                                if (j == k) { br = 0.0; bi_ = 0.0; }
                                nr = br - (cr * rk_re[j] - ci * rk_im[j]);
                                ni = bi_ - (cr * rk_im[j] + ci * rk_re[j]);
                            }
                            Sre[i * lds + j] = nr; Sim[i * lds + j] = ni;
                        }
                        if (tid == 0) perm[k] = pr;
                        __syncthreads();
                    }
                    if (tid == 0) {
                        for (int j = 0; j < ne; ++j) srcc[j] = j;
                        for (int k = ne - 1; k >= 0; --k) { const int t = srcc[k]; srcc[k] = srcc[perm[k]]; srcc[perm[k]] = t; }
                        S.ctl[1] = -1; S.ctl[2] = -1;
                    }
                    __syncthreads();
                    const double* Gsrc = L.resident ? Gs : Gglob;
                    auto prob_of = [&](int m, int e) -> double {
                        const int ce = srcc[e];
                        double rr = 0.0;
                        for (int i = 0; i < ne; ++i) {
                            rr = fma(psi_re[m + (int64_t)n * i], Sre[i * lds + ce], rr);
                            rr = fma(-psi_im[m + (int64_t)n * i], Sim[i * lds + ce], rr);
                        }
                        const double vdd = -Gsrc[m + (int64_t)L.ldg * S.occ[e]];
                        return fmin(1.0, fmax(0.0, 2.0 * dt * rr * vdd));
                    };
                    for (int e = warp; e < ne; e += nwarps) {
                        double part = 0.0;
                        for (int u = lane; u < nun; u += 32) part += prob_of(S.un[u], e);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                        if (lane == 0) sum_e[e] = part;
                    }
                    __syncthreads();
                    // select_new_state (iesh.jl:318-335): electrons outer, unoccupied states inner
                    double cum = 0.0;
                    int e_scan = 0;
                    while (true) {
                        while (e_scan < ne && !(xi < cum + sum_e[e_scan] * (1.0 + 1e-12) + 1e-300)) { cum += sum_e[e_scan]; ++e_scan; }
                        if (e_scan >= ne) break;
                        for (int u = tid; u < nun; u += nt) probs[u] = prob_of(S.un[u], e_scan);
                        __syncthreads();
                        if (tid == 0) {
                            double c = cum; int found = -1;
                            for (int u = 0; u < nun; ++u) { c += probs[u]; if (xi < c) { found = u; break; } }
                            S.ctl[1] = found;
                            S.red[40] = c;
                        }
                        __syncthreads();
                        const int found = S.ctl[1];
                        cum = S.red[40];
                        __syncthreads();
                        if (found >= 0) { if (tid == 0) { S.ctl[2] = e_scan; } break; }
                        ++e_scan;
                    }
                    __syncthreads();
                    const int he = S.ctl[2];
                    if (he >= 0 && e_scan < ne) {
                        const int hm = S.un[S.ctl[1]];
                        const int old_state = S.occ[he];
                        bool accept = true;
                        if (p.rescaling != NQCB200_RESCALE_OFF) {                 // surface_hopping.jl:64-99
                            const double wd = iesh_wdiff(S, hm, old_state);
                            const double dhw = mdl.erp ? dh + phi * ((S.lam[hm] - h) + (S.lam[old_state] - h)) : dh;
                            const double d = -dhw * S.z0[hm] * S.z0[old_state] / wd;
                            const double aa = (d * d / mdl.mass) / 2.0, bb = d * v, cc = wd;
                            const double disc = bb * bb - 4.0 * aa * cc;
                            if (disc < 0.0) {
                                accept = false;
                                nfrus += (tid == 0);
                                if (p.rescaling == NQCB200_RESCALE_VINVERSION) {
                                    const double nrm = sqrt(d * d);
                                    const double gam = v * d / nrm;
                                    v -= 2.0 * gam * d / nrm;
                                }
                            } else {
                                const double root = sqrt(disc);
                                const double gam = (bb < 0.0) ? (bb + root) / (2.0 * aa) : (bb - root) / (2.0 * aa);
                                v -= gam * d / mdl.mass;
                            }
                        }
                        if (accept) {
                            nhops += (tid == 0);
                            __syncthreads();
                            if (tid == 0) S.occ[he] = hm;                          // not re-sorted (iesh.jl:399-407)
                            iesh_refresh_unoccupied(p, S);
                        }
                    }
                }
                }
                __syncthreads();
            }

            if constexpr (RP) {
                if (v != v_before_hop) {      // rescaling / velocity inversion: the same change on every bead (rpsh.jl:30-50)
                    const double dv = v - v_before_hop;
                    __syncthreads();
                    if (tid < p.B) S.rp[32 + tid] += dv;
                    __syncthreads();
                    iesh_rp_centroid(p, S, r, v);
                    iesh_build_generator(p, S, mdl, L.resident ? Gs : Gglob, h, dh, phi, v, g_sigma, g_wspan, g_gnorm, g_sabs);
                }
            }
            // ---- EDC decoherence (decoherence_corrections.jl:21-38) -----------------------------------
            if (p.edc_C > 0.0) {
                const double Ekin = (mdl.mass * v * v) / 2.0;
                const double fac = 1.0 + p.edc_C / Ekin;
                for (int e = warp; e < ne; e += nwarps) {
                    const int oc = S.occ[e];
                    double un_norm = 0.0;
                    for (int i = lane; i < n; i += 32) {
                        if (i == oc) continue;
                        const double tau = fac / fabs(iesh_wdiff(S, i, oc));
                        const double damp = exp(-dt / tau);
                        const double a = psi_re[i + (int64_t)n * e] * damp, b = psi_im[i + (int64_t)n * e] * damp;
                        psi_re[i + (int64_t)n * e] = a; psi_im[i + (int64_t)n * e] = b;
                        un_norm += a * a + b * b;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) un_norm += __shfl_xor_sync(0xffffffffu, un_norm, o);
                    if (lane == 0) {
                        const double a = psi_re[oc + (int64_t)n * e], b = psi_im[oc + (int64_t)n * e];
                        const double sc = sqrt((1.0 - un_norm) / (a * a + b * b));
                        psi_re[oc + (int64_t)n * e] = a * sc; psi_im[oc + (int64_t)n * e] = b * sc;
                    }
                }
                __syncthreads();
            }

            // ---- TerminatingCallback: after the hopping callback, on the new u (callbacks.jl:29) -------------
            if (p.term_dof >= 0 && tid == 0 && iesh_outside(p, r, v, step + 1)) p.term_step[traj] = step + 1;
            }
        save_point:

            // ---- save (after the callback) ---------------------------------------------
            if ((step + 1) % p.save_every == 0) {
                const int64_t isave = (step + 1) / p.save_every;
                if (isave < p.nsave) {
                    double rp_pot = 0.0;
                    if constexpr (RP) {
                        if (p.observables & ((1u << NQCB200_OBS_POTENTIAL) | (1u << NQCB200_OBS_TOTAL_ENERGY))) {
                            rp_pot = iesh_rp_potential(p, S, mdl, psi_re, psi_im, vnorm);
                            double h, dh, u0, du0, fs, dfs;
                            mdl.eval(r, h, dh, u0, du0, fs, dfs);
                            iesh_eigen(p, S, h, fs, vnorm, false);       // back to the centroid (estimators, next generator's roots)
                        }
                    }
                    iesh_record_save(p, S, traj, (int)isave, r, v, mdl, psi_re, psi_im, Bs, L.work_doubles - L.off_b, rp_pot);
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt) p.iesh_lam[traj * n + i] = S.lam[i];
        for (int e = tid; e < ne; e += nt) p.state[traj * ne + e] = S.occ[e];
        if constexpr (RP) {
            if (tid < p.B) {
                p.r[(int64_t)tid * p.ntraj + traj] = S.rp[tid];
                p.v[(int64_t)tid * p.ntraj + traj] = S.rp[32 + tid];
                p.acc[(int64_t)tid * p.ntraj + traj] = S.rp[64 + tid];
            }
        } else if (tid == 0) { p.r[traj] = r; p.v[traj] = v; p.acc[traj] = acc; }
        if (p.diagnostics && p.diag_eig) {
            // eigenvalues, NAC d[j,i] (column-major j + n i), eigenvectors of the last evaluated geometry
            double h, dh, u0, du0, fs, dfs;
            mdl.eval(r, h, dh, u0, du0, fs, dfs);
            const double phi = dfs / fs;
            for (int i = tid; i < n; i += nt) p.diag_eig[traj * n + i] = S.lam[i];
            for (int idx = tid; idx < n * n; idx += nt) {
                const int j = idx % n, i = idx / n;
                const double dhw = dh + phi * ((S.lam[i] - h) + (S.lam[j] - h));
                p.diag_nac[traj * (int64_t)n * n + idx] = (i == j) ? 0.0 : -dhw * S.z0[j] * S.z0[i] / iesh_wdiff(S, j, i);
                p.diag_Z[traj * (int64_t)n * n + idx] = iesh_Z(S, j, i);
            }
        }
    }
    if (tid == 0) {
        if (nhops) atomicAdd(&p.counters[0], nhops);
        if (nfrus) atomicAdd(&p.counters[1], nfrus);
        if (nunpruned) atomicAdd(&p.counters[3], nunpruned);
        atomicAdd(&p.counters[4], ndet);
        atomicAdd(&p.counters[5], nstages);
        atomicAdd(&p.counters[6], ngemm);
    }
}

// Initialisation: update_cache!(r0) with a cold root search, gauge signs against the identity
// (or a user reference Z through p.Zprev, [T][n*n] trajectory-major), initial acceleration
// (verlet_with_electronics.jl:30-40), save point 0.
template <bool RP>
__global__ void __launch_bounds__(384, 1) iesh_init_kernel(const __grid_constant__ KParams p, int user_gauge, int,
                                                          const double*) {
    extern __shared__ __align__(16) double iesh_sm[];
    IeshSmem S;
    S.carve(iesh_sm, p.n);
    const int n = p.n, ne = p.ne, tid = threadIdx.x, nt = blockDim.x;
    IeshModel mdl;
    mdl.load(p);
    const double vnorm = iesh_load_bath(p, S);
    for (int64_t traj = blockIdx.x; traj < p.ntraj; traj += gridDim.x) {
        const double* psi_re = p.sig_re + traj * (int64_t)n * ne;
        const double* psi_im = p.sig_im + traj * (int64_t)n * ne;
        __syncthreads();
        for (int i = tid; i < n; i += nt) S.sgn[i] = 1.0;
        for (int e = tid; e < ne; e += nt) S.occ[e] = p.state[traj * ne + e];
        iesh_refresh_unoccupied(p, S);
        double r = p.r[traj], v = p.v[traj];
        if constexpr (RP) {      // gauge, roots and estimators of a ring polymer: the centroid
            if (tid < p.B) {
                S.rp[tid] = p.r[(int64_t)tid * p.ntraj + traj];
                S.rp[32 + tid] = p.v[(int64_t)tid * p.ntraj + traj];
            }
            __syncthreads();
            iesh_rp_centroid(p, S, r, v);
        }
        double h, dh, u0, du0, fs, dfs;
        mdl.eval(r, h, dh, u0, du0, fs, dfs);
        const double phi = dfs / fs;
        iesh_eigen(p, S, h, fs, vnorm, true);
        // gauge: flip column i when dot(Z_new[:,i], Z_ref[:,i]) < 0 ; identity reference -> sign of Z[i,i]
        for (int i = tid; i < n; i += nt) {
            double dot;
            if (user_gauge) {
                dot = 0.0;
                for (int k = 0; k < n; ++k) dot += iesh_Z(S, k, i) * p.Zprev[traj * (int64_t)n * n + k + (int64_t)n * i];
            } else dot = iesh_Z(S, i, i);
            S.sgn[i] = (dot < 0.0) ? -1.0 : 1.0;
        }
        __syncthreads();
        for (int i = tid; i < n; i += nt) { S.z0[i] *= S.sgn[i]; p.iesh_sgn[traj * n + i] = S.sgn[i]; p.iesh_lam[traj * n + i] = S.lam[i]; }
        __syncthreads();
        double rp_pot = 0.0;
        if constexpr (RP) {
            // initial acceleration of every bead (bcb_wavefunction.jl:26-35) and the bead sum of the potential
            for (int b = 0; b < p.B; ++b) {
                double hb, dhb, u0b, du0b, fsb, dfsb;
                mdl.eval(S.rp[b], hb, dhb, u0b, du0b, fsb, dfsb);
                iesh_eigen(p, S, hb, fsb, vnorm, false);
                const double ab = iesh_acceleration(p, S, mdl, psi_re, psi_im, hb, dhb, du0b, dfsb / fsb);
                if (tid == 0) p.acc[(int64_t)b * p.ntraj + traj] = ab;
            }
            if (p.observables & ((1u << NQCB200_OBS_POTENTIAL) | (1u << NQCB200_OBS_TOTAL_ENERGY)))
                rp_pot = iesh_rp_potential(p, S, mdl, psi_re, psi_im, vnorm);
            iesh_eigen(p, S, h, fs, vnorm, false);      // back to the centroid
            for (int i = tid; i < n; i += nt) p.iesh_lam[traj * n + i] = S.lam[i];
            __syncthreads();
        }
        double occsum, occsum2 = 0.0;
        if (p.mean_field) occsum = iesh_mean_field_weight(p, S, psi_re, psi_im, mdl.erp, h, occsum2);
        else {
            double part = 0.0, part2 = 0.0;
            for (int e = tid; e < ne; e += nt) {
                const int o = S.occ[e];
                const double z = S.z0[o];
                part += z * z;
                if (mdl.erp) part2 = fma(z * z, S.lam[o] - h, part2);
            }
            occsum = iesh_block_sum(part, S.red);
            if (mdl.erp) occsum2 = iesh_block_sum(part2, S.red);
        }
        if (!RP && tid == 0) p.acc[traj] = ((-du0 - dh * occsum) - 2.0 * phi * occsum2) / mdl.mass;
        iesh_record_save(p, S, traj, 0, r, v, mdl, psi_re, psi_im, S.work + p.iesh.off_b, p.iesh.work_doubles - p.iesh.off_b, rp_pot);
        {
            // are the orbitals orthonormal?  (enables the determinant-free pruning bound of the step kernel)
            double dev = 0.0;
            for (int idx = tid; idx < ne * ne; idx += nt) {
                const int e1 = idx % ne, e2 = idx / ne;
                if (e1 > e2) continue;
                double sr = 0.0, si = 0.0;
                for (int i = 0; i < n; ++i) {
                    const double ar = psi_re[i + (int64_t)n * e1], ai = psi_im[i + (int64_t)n * e1];
                    const double br = psi_re[i + (int64_t)n * e2], bi = psi_im[i + (int64_t)n * e2];
                    sr += ar * br + ai * bi; si += ar * bi - ai * br;
                }
                dev = fmax(dev, fmax(fabs(sr - (e1 == e2 ? 1.0 : 0.0)), fabs(si)));
            }
            dev = iesh_block_max(dev, S.red);
            if (tid == 0) p.iesh_orth[traj] = (dev < 1e-11) ? 1.0 : 0.0;
        }
        if (p.diagnostics && p.diag_eig) {
            for (int i = tid; i < n; i += nt) p.diag_eig[traj * n + i] = S.lam[i];
            for (int idx = tid; idx < n * n; idx += nt) {
                const int j = idx % n, i = idx / n;
                const double dhw = dh + phi * ((S.lam[i] - h) + (S.lam[j] - h));
                p.diag_nac[traj * (int64_t)n * n + idx] = (i == j) ? 0.0 : -dhw * S.z0[j] * S.z0[i] / iesh_wdiff(S, j, i);
                p.diag_Z[traj * (int64_t)n * n + idx] = iesh_Z(S, j, i);
            }
        }
    }
}

#endif  // __CUDACC__

}  // namespace nq

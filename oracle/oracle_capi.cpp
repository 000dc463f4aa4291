// oracle_capi.cpp -- CPU ORACLE (test infrastructure, NOT product code; see oracle_core.hpp).
//
// C entry points `nqco_*` mirroring include/nqcb200.h one to one so that a parity test drives the
// CUDA engine and this oracle with the same config struct, the same host buffers and the same
// injected (or Philox) draws.  OpenMP over trajectories (they are independent, exactly like the
// reference's EnsembleThreads; docs/src/ensemble_simulations.md:25-28).
#include <cstdio>
#include <cstdlib>
#include <string>

#include "oracle_dynamics.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace nqco;

namespace {

inline bool iesh_family(int method) { return method == NQCB200_METHOD_IESH || method == NQCB200_METHOD_EHRENFEST_NA; }

// Philox4x32-10 (Salmon et al. 2011).  key = seed, counter = (global trajectory id, step, purpose)
inline void philox4x32_10(uint32_t ctr[4], uint32_t key0, uint32_t key1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)M0 * ctr[0], p1 = (uint64_t)M1 * ctr[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ ctr[1] ^ key0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ key1, n3 = lo0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        key0 += W0; key1 += W1;
    }
}
inline double philox_uniform(uint64_t seed, uint64_t gid, uint64_t step, uint32_t purpose) {
    uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)step,
                     ((uint32_t)(step >> 32) & 0x00FFFFFFu) | (purpose << 24)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t bits = ((uint64_t)c[0] << 32) | c[1];
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}

// Two standard normals per Philox block (Box-Muller): the stream of nqcb200_sample_state (include/nqcb200.h),
// counter = (global trajectory id, component, purpose 2); first normal -> position, second -> velocity.
inline void philox_normal2(uint64_t seed, uint64_t gid, uint64_t comp, double& z0, double& z1, uint32_t purpose = 2u) {
    uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)comp, ((uint32_t)(comp >> 32) & 0x00FFFFFFu) | (purpose << 24)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = (double)(((((uint64_t)c[0] << 32) | c[1]) >> 11) + 1ull) * (1.0 / 9007199254740992.0);
    const double u2 = (double)((((uint64_t)c[2] << 32) | c[3]) >> 11) * (1.0 / 9007199254740992.0);
    const double rad = std::sqrt(-2.0 * std::log(u1)), ang = 6.283185307179586476925286766559 * u2;
    z0 = rad * std::cos(ang); z1 = rad * std::sin(ang);
}

struct ObsLayout {
    int width[NQCB200_OBS_COUNT];
    int64_t offset[NQCB200_OBS_COUNT];
    int64_t total = 0;       // doubles over all enabled observables ([nsave][width] each)
    int64_t per_save = 0;    // sum of enabled widths
};

}  // namespace

struct nqco_handle {
    Setup S;
    std::vector<Trajectory> traj;
    ObsLayout L;
    vec obs_sum;                 // [obs][nsave][width] packed by offset
    vec obs_traj;                // per trajectory: [traj][obs-packed as obs_sum]
    vec draws;                   // injected: [step][traj]
    int64_t draws_first_step = 0, draws_nsteps = 0;
    vec noise;                   // ThermalLangevin injected normals: [step][traj][B*D]
    int64_t noise_first_step = 0, noise_nsteps = 0;
    vec Zref;                    // optional gauge reference
    int64_t zref_per_traj = 0;
    int64_t nsave_done = 0;
    bool has_state = false;
    // TerminatingCallback(u -> r[dof] < lo || r[dof] > hi), callbacks.jl:29; dof < 0: none
    int term_dof = -1, term_outgoing = 0;
    double term_lo = 0.0, term_hi = 0.0, term_tcut = INFINITY;
    std::string err;
};

static std::string g_err;

static int obs_width(const nqcb200_config& c, int id) {
    const int n = c.nstates, D = c.ndofs;
    switch (id) {
        case NQCB200_OBS_ADIABATIC_POP: case NQCB200_OBS_DIABATIC_POP: return n;
        case NQCB200_OBS_POPCORR_DIABATIC: case NQCB200_OBS_POPCORR_ADIABATIC: return n * n;
        case NQCB200_OBS_KINETIC: case NQCB200_OBS_POTENTIAL: case NQCB200_OBS_TOTAL_ENERGY: return 1;
        case NQCB200_OBS_POSITION: case NQCB200_OBS_VELOCITY: return D;
        case NQCB200_OBS_DISCRETE_STATE: return iesh_family(c.method) ? c.nelectrons : 1;
        case NQCB200_OBS_SCATTERING: case NQCB200_OBS_SCATTERING_DIABATIC: return 2 * n;
        case NQCB200_OBS_SIGMA: return iesh_family(c.method) ? 2 * n * c.nelectrons : 2 * n * n;
        case NQCB200_OBS_MAPPING_Q: case NQCB200_OBS_MAPPING_P: return c.method == NQCB200_METHOD_NRPMD ? n * c.nbeads : 0;
    }
    return 0;
}

static void build_layout(nqco_handle* h) {
    const nqcb200_config& c = h->S.cfg;
    int64_t off = 0;
    h->L.per_save = 0;
    for (int id = 0; id < NQCB200_OBS_COUNT; ++id) {
        h->L.width[id] = obs_width(c, id);
        h->L.offset[id] = -1;
        if (c.observables & (1u << id)) {
            h->L.offset[id] = off;
            off += (int64_t)c.nsave * h->L.width[id];
            h->L.per_save += h->L.width[id];
        }
    }
    h->L.total = off;
}

static void record_save(nqco_handle* h, int64_t isave) {
    const Setup& S = h->S;
    const nqcb200_config& c = S.cfg;
    if (isave >= c.nsave) return;
    const int n = S.n;
    const int64_t T = (int64_t)h->traj.size();
    const bool last = (isave == c.nsave - 1);
    vec local;  // per trajectory values, then summed sequentially (deterministic)
    std::vector<vec> vals(T);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
        Trajectory& tr = h->traj[t];
        vec& out = vals[t];
        out.assign(h->L.per_save, 0.0);
        int64_t p = 0;
        vec adi(n), dia(n);
        bool need_adi = c.observables & ((1u << NQCB200_OBS_ADIABATIC_POP) | (1u << NQCB200_OBS_POPCORR_ADIABATIC) |
                                         (1u << NQCB200_OBS_SCATTERING));
        bool need_dia = c.observables & ((1u << NQCB200_OBS_DIABATIC_POP) | (1u << NQCB200_OBS_POPCORR_DIABATIC) |
                                         (1u << NQCB200_OBS_SCATTERING_DIABATIC));
        if (need_adi) adiabatic_population(S, tr, adi.data());
        if (need_dia) diabatic_population(S, tr, dia.data());
        if (isave == 0) {  // initial value of the correlation functions (TimeCorrelationFunctions.jl:43-44)
            tr.pop0.assign(2 * n, 0.0);
            for (int i = 0; i < n; ++i) { tr.pop0[i] = dia[i]; tr.pop0[n + i] = adi[i]; }
        }
        for (int id = 0; id < NQCB200_OBS_COUNT; ++id) {
            if (!(c.observables & (1u << id))) continue;
            double* o = &out[p];
            switch (id) {
                case NQCB200_OBS_ADIABATIC_POP: for (int i = 0; i < n; ++i) o[i] = adi[i]; break;
                case NQCB200_OBS_DIABATIC_POP: for (int i = 0; i < n; ++i) o[i] = dia[i]; break;
                case NQCB200_OBS_POPCORR_DIABATIC:
                    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) o[i + n * j] = tr.pop0[i] * dia[j];
                    break;
                case NQCB200_OBS_POPCORR_ADIABATIC:
                    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) o[i + n * j] = tr.pop0[n + i] * adi[j];
                    break;
                case NQCB200_OBS_KINETIC: o[0] = kinetic_energy(S, tr); break;
                case NQCB200_OBS_POTENTIAL: o[0] = potential_energy(S, tr); break;
                case NQCB200_OBS_TOTAL_ENERGY:
                    o[0] = kinetic_energy(S, tr) + potential_energy(S, tr) + spring_energy(S, tr);
                    break;
                case NQCB200_OBS_POSITION: centroid_of(S, tr.r, o); break;
                case NQCB200_OBS_VELOCITY: centroid_of(S, tr.v, o); break;
                case NQCB200_OBS_DISCRETE_STATE:
                    if (c.method == NQCB200_METHOD_IESH) for (int e = 0; e < S.ne; ++e) o[e] = tr.occ[e] + 1;
                    else o[0] = tr.state + 1;
                    break;
                case NQCB200_OBS_SCATTERING:
                case NQCB200_OBS_SCATTERING_DIABATIC:
                    if (last) {
                        const vec& pp = (id == NQCB200_OBS_SCATTERING) ? adi : dia;
                        int base = tr.r[0] > 0.0 ? n : 0;
                        for (int i = 0; i < n; ++i) o[base + i] = pp[i];
                    }
                    break;
                case NQCB200_OBS_SIGMA: {
                    int len = h->L.width[id] / 2;
                    for (int i = 0; i < len; ++i) { o[i] = tr.sigma[i].real(); o[len + i] = tr.sigma[i].imag(); }
                } break;
                case NQCB200_OBS_MAPPING_Q:      // OutputMappingPosition / Momentum, DynamicsOutputs.jl:157,165
                case NQCB200_OBS_MAPPING_P: {
                    const vec& m = (id == NQCB200_OBS_MAPPING_Q) ? tr.qmap : tr.pmap;
                    for (int i = 0; i < h->L.width[id]; ++i) o[i] = m[i];
                } break;
            }
            p += h->L.width[id];
        }
    }
    for (int64_t t = 0; t < T; ++t) {
        int64_t p = 0;
        for (int id = 0; id < NQCB200_OBS_COUNT; ++id) {
            if (!(c.observables & (1u << id))) continue;
            const int w = h->L.width[id];
            double* dst = &h->obs_sum[h->L.offset[id] + isave * w];
            for (int i = 0; i < w; ++i) dst[i] += vals[t][p + i];
            if (c.per_trajectory) {
                double* dt_ = &h->obs_traj[(size_t)t * h->L.total + h->L.offset[id] + isave * w];
                for (int i = 0; i < w; ++i) dt_[i] = vals[t][p + i];
            }
            p += w;
        }
    }
    h->nsave_done = isave + 1;
}

extern "C" {

int nqco_version(void) { return NQCB200_ABI_VERSION; }

const char* nqco_last_error(const nqco_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int nqco_set_num_threads(int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    return omp_get_max_threads();
#else
    (void)nthreads;
    return 1;
#endif
}

int nqco_create(const nqcb200_config* cfg, nqco_handle** out) {
    if (!cfg || !out) { g_err = "null argument"; return NQCB200_ERR_INVALID; }
    if (cfg->abi_version != NQCB200_ABI_VERSION) { g_err = "abi version mismatch"; return NQCB200_ERR_INVALID; }
    if (cfg->nstates < 1 || cfg->ndofs < 1 || cfg->nbeads < 1 || cfg->ntraj < 0 || cfg->save_every < 1 || cfg->nsave < 1 ||
        !cfg->masses) { g_err = "invalid sizes"; return NQCB200_ERR_INVALID; }
    nqco_handle* h = new nqco_handle();
    Setup& S = h->S;
    S.cfg = *cfg;
    S.n = cfg->nstates; S.D = cfg->ndofs; S.B = cfg->nbeads; S.ne = cfg->nelectrons;
    S.model.kind = cfg->model; S.model.n = S.n; S.model.D = S.D;
    std::memcpy(S.model.p, cfg->params, sizeof(cfg->params));
    if (cfg->nbath > 0 && cfg->bath_a) S.model.ba.assign(cfg->bath_a, cfg->bath_a + cfg->nbath);
    if (cfg->nbath > 0 && cfg->bath_b) S.model.bb.assign(cfg->bath_b, cfg->bath_b + cfg->nbath);
    S.masses.assign(cfg->masses, cfg->masses + S.D);
    S.cfg.masses = nullptr; S.cfg.bath_a = nullptr; S.cfg.bath_b = nullptr;
    S.omega_n = S.B * cfg->temperature;
    S.U = normal_mode_matrix(S.B);
    S.cayley = cayley_propagator(S.B, S.omega_n, cfg->dt,
                                 cfg->method == NQCB200_METHOD_NRPMD || cfg->method == NQCB200_METHOD_THERMAL_LANGEVIN);
    bool quantum = !S.model.classical();
    if ((cfg->method == NQCB200_METHOD_CLASSICAL || cfg->method == NQCB200_METHOD_THERMAL_LANGEVIN) == quantum) {
        // classical method on a quantum model would route through the Fermi-weighted force
        // (rpmdef.jl:30-57) which is out of scope; quantum method on a classical model is invalid
        g_err = "method/model combination unsupported"; delete h; return NQCB200_ERR_UNSUPPORTED;
    }
    if (iesh_family(cfg->method) && (S.ne < 1 || S.ne >= S.n || (S.B > 1 && cfg->edc_C > 0.0))) {
        g_err = "IESH needs 1 <= nelectrons < nstates (EDC: nbeads == 1)"; delete h; return NQCB200_ERR_INVALID;
    }
    build_layout(h);
    h->obs_sum.assign(h->L.total, 0.0);
    if (cfg->per_trajectory) h->obs_traj.assign((size_t)h->L.total * cfg->ntraj, 0.0);
    h->traj.resize(cfg->ntraj);
    *out = h;
    return NQCB200_OK;
}

int nqco_destroy(nqco_handle* h) { delete h; return NQCB200_OK; }

int nqco_observable_width(const nqco_handle* h, int obs_id) {
    if (!h || obs_id < 0 || obs_id >= NQCB200_OBS_COUNT) return NQCB200_ERR_INVALID;
    return h->L.width[obs_id];
}

int nqco_set_gauge_reference(nqco_handle* h, const double* Z, int64_t count_per_traj) {
    if (!h || !Z) return NQCB200_ERR_INVALID;
    h->zref_per_traj = count_per_traj;
    h->Zref.assign(Z, Z + (size_t)count_per_traj * h->S.n * h->S.n * h->traj.size());
    return NQCB200_OK;
}

// basis: 0 = sigma given in the adiabatic basis (PureState(i, Adiabatic())), 1 = diabatic density
// matrix rho, transformed sigma = Z' rho Z at r0 (density_matrix_dynamics.jl:37-75).
// state == NULL for FSSH: the active state is sampled with weights Re diag(sigma)
// (fssh.jl:53-54, StatsBase.sample(Weights)) from state_draw[traj] or Philox(purpose=1).
static int set_state_impl(nqco_handle* h, const double* r, const double* v, const double* sre, const double* sim,
                          const int32_t* state, int basis, const double* state_draw) {
    if (!h || !r || !v) return NQCB200_ERR_INVALID;
    Setup& S = h->S;
    const int n = S.n, D = S.D, B = S.B, ne = S.ne;
    const int method = S.cfg.method;
    const int64_t T = (int64_t)h->traj.size();
    const bool density = (method == NQCB200_METHOD_FSSH || method == NQCB200_METHOD_EHRENFEST);
    const size_t nsig = iesh_family(method) ? (size_t)n * ne : (size_t)n * n;
    if (density && !sre) { h->err = "sigma required"; return NQCB200_ERR_INVALID; }
    if (method == NQCB200_METHOD_IESH && sre && !state) { h->err = "state required"; return NQCB200_ERR_INVALID; }
    int rc = NQCB200_OK;
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
        Trajectory& tr = h->traj[t];
        tr.r.assign(r + (size_t)t * B * D, r + (size_t)(t + 1) * B * D);
        tr.v.assign(v + (size_t)t * B * D, v + (size_t)(t + 1) * B * D);
        tr.sigma.assign(nsig, cd(0.0));
        if ((density || iesh_family(method)) && sre)
            for (size_t i = 0; i < nsig; ++i) tr.sigma[i] = cd(sre[t * nsig + i], sim ? sim[t * nsig + i] : 0.0);
        if (iesh_family(method) && !sre)      // iesh.jl:89-128: electron e starts in the adiabatic orbital state[e] (1..ne without state)
            for (int e = 0; e < ne; ++e) tr.sigma[(size_t)(state ? state[t * ne + e] - 1 : e) + (size_t)n * e] = cd(1.0);
        tr.occ.clear();
        if (method == NQCB200_METHOD_IESH) for (int e = 0; e < ne; ++e) tr.occ.push_back(state ? state[t * ne + e] - 1 : e);
        const double* zr = h->Zref.empty() ? nullptr : &h->Zref[(size_t)t * h->zref_per_traj * n * n];
        initialise(S, tr, zr);
        if (density && basis == 1) {
            const vec& Z = hop_cache(S, tr).Z;
            cvec rho = tr.sigma, tmp((size_t)n * n);
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    cd s = 0.0;
                    for (int k = 0; k < n; ++k) s += rho[i + (size_t)n * k] * Z[k + (size_t)n * j];
                    tmp[i + (size_t)n * j] = s;
                }
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    cd s = 0.0;
                    for (int k = 0; k < n; ++k) s += Z[k + (size_t)n * i] * tmp[k + (size_t)n * j];
                    tr.sigma[i + (size_t)n * j] = s;
                }
        }
        if (method == NQCB200_METHOD_FSSH) {
            if (state) tr.state = state[t] - 1;
            else {
                double xi = state_draw ? state_draw[t] : philox_uniform(S.cfg.seed, S.cfg.traj_offset + t, 0, 1);
                double tot = 0.0;
                for (int i = 0; i < n; ++i) tot += tr.sigma[i + (size_t)n * i].real();
                double target = xi * tot, cw = tr.sigma[0].real();
                int i = 0;
                while (cw < target && i < n - 1) { ++i; cw += tr.sigma[i + (size_t)n * i].real(); }
                tr.state = i;
            }
        }
        initial_acceleration(S, tr);
    }
    std::fill(h->obs_sum.begin(), h->obs_sum.end(), 0.0);
    std::fill(h->obs_traj.begin(), h->obs_traj.end(), 0.0);
    h->nsave_done = 0;
    h->has_state = true;
    if (S.cfg.method != NQCB200_METHOD_NRPMD) record_save(h, 0);
    return rc;
}

int nqco_set_state(nqco_handle* h, const double* r, const double* v, const double* sre, const double* sim,
                   const int32_t* state) {
    return set_state_impl(h, r, v, sre, sim, state, 0, nullptr);
}
int nqco_set_state_diabatic(nqco_handle* h, const double* r, const double* v, const double* rho_re,
                            const double* rho_im, const int32_t* state, const double* state_draw) {
    return set_state_impl(h, r, v, rho_re, rho_im, state, 1, state_draw);
}

// CPU restatement of nqcb200_sample_state: same Philox / Box-Muller stream, then the ordinary set_state path.
int nqco_sample_state(nqco_handle* h, const nqcb200_dist* r_dist, const nqcb200_dist* v_dist, int normal_modes,
                      const double* rho_re, const double* rho_im, int diabatic, int32_t state) {
    if (!h || !r_dist || !v_dist) return NQCB200_ERR_INVALID;
    const Setup& S = h->S;
    const int n = S.n, D = S.D, B = S.B;
    const int method = S.cfg.method;
    if (iesh_family(method)) { h->err = "device-side sampling is not available for AdiabaticIESH / EhrenfestNA"; return NQCB200_ERR_UNSUPPORTED; }
    const bool density = (method == NQCB200_METHOD_FSSH || method == NQCB200_METHOD_EHRENFEST);
    if (density && !rho_re) return NQCB200_ERR_INVALID;
    const int64_t T = (int64_t)h->traj.size();
    const int C = B * D;
    std::vector<double> r((size_t)T * C), v((size_t)T * C);
    const vec U = normal_mode_matrix(B);    // U[j + B*k]
    for (int64_t t = 0; t < T; ++t) {
        for (int c = 0; c < C; ++c) {
            double z0 = 0.0, z1 = 0.0;
            if (r_dist[c].kind == 1 || v_dist[c].kind == 1) philox_normal2(S.cfg.seed, (uint64_t)(S.cfg.traj_offset + t), (uint64_t)c, z0, z1);
            r[(size_t)t * C + c] = r_dist[c].kind == 1 ? std::fma(r_dist[c].b, z0, r_dist[c].a) : r_dist[c].a;
            v[(size_t)t * C + c] = v_dist[c].kind == 1 ? std::fma(v_dist[c].b, z1, v_dist[c].a) : v_dist[c].a;
        }
        if (normal_modes && B > 1) {
            for (std::vector<double>* x : {&r, &v}) {
                std::vector<double> tmp(C);
                for (int d = 0; d < D; ++d)
                    for (int j = 0; j < B; ++j) {
                        double s = 0.0;
                        for (int k = 0; k < B; ++k) s = std::fma(U[j + (size_t)B * k], (*x)[(size_t)t * C + (size_t)k * D + d], s);
                        tmp[(size_t)j * D + d] = s;
                    }
                std::copy(tmp.begin(), tmp.end(), x->begin() + (size_t)t * C);
            }
        }
    }
    std::vector<double> sre, sim;
    std::vector<int32_t> st;
    if (density) {
        sre.resize((size_t)T * n * n); sim.assign((size_t)T * n * n, 0.0);
        for (int64_t t = 0; t < T; ++t)
            for (int i = 0; i < n * n; ++i) { sre[(size_t)t * n * n + i] = rho_re[i]; if (rho_im) sim[(size_t)t * n * n + i] = rho_im[i]; }
        if (method == NQCB200_METHOD_FSSH && state > 0) st.assign((size_t)T, state);
    }
    return set_state_impl(h, r.data(), v.data(), density ? sre.data() : nullptr, density ? sim.data() : nullptr,
                          st.empty() ? nullptr : st.data(), diabatic ? 1 : 0, nullptr);
}

int nqco_set_mapping(nqco_handle* h, const double* qmap, const double* pmap) {
    if (!h || !qmap || !pmap || !h->has_state) return NQCB200_ERR_INVALID;
    const size_t per = (size_t)h->S.n * h->S.B;
    for (size_t t = 0; t < h->traj.size(); ++t) {
        h->traj[t].qmap.assign(qmap + t * per, qmap + (t + 1) * per);
        h->traj[t].pmap.assign(pmap + t * per, pmap + (t + 1) * per);
    }
    std::fill(h->obs_sum.begin(), h->obs_sum.end(), 0.0);
    record_save(h, 0);
    return NQCB200_OK;
}

// CPU restatement of nqcb200_sample_occupations (same Philox stream and the same list bookkeeping as the device kernel):
// sample_fermi_dirac_distribution, DynamicsUtils.jl:194-208 (Boltzmann-factor variant), on the adiabatic energies at r0.
int nqco_sample_occupations(nqco_handle* h, double beta) {
    if (!h || !h->has_state || h->S.cfg.method != NQCB200_METHOD_IESH || !(beta >= 0.0)) return NQCB200_ERR_INVALID;
    Setup& S = h->S;
    const int n = S.n, ne = S.ne, nun = n - ne, BD = S.B * S.D;
    const int64_t T = (int64_t)h->traj.size();
    std::vector<double> r((size_t)T * BD), v((size_t)T * BD);
    std::vector<int32_t> state((size_t)T * ne);
    const bool cold = !(beta < 1e300);
    for (int64_t t = 0; t < T; ++t) {
        const Trajectory& tr = h->traj[t];
        std::copy(tr.r.begin(), tr.r.end(), r.begin() + t * BD);
        std::copy(tr.v.begin(), tr.v.end(), v.begin() + t * BD);
        const vec& E = hop_cache(S, tr).w;
        std::vector<int> lst(n);
        for (int i = 0; i < n; ++i) lst[i] = i;
        int* occ = lst.data();
        int* un = occ + ne;
        const uint64_t gid = (uint64_t)(S.cfg.traj_offset + t);
        for (int64_t it = 0; it < (int64_t)n * ne; ++it) {
            const int k = std::min(ne - 1, (int)(philox_uniform(S.cfg.seed, gid, 3 * it + 0, 5u) * ne));
            const int u = std::min(nun - 1, (int)(philox_uniform(S.cfg.seed, gid, 3 * it + 1, 5u) * nun));
            const double de = E[un[u]] - E[occ[k]];
            const double prob = cold ? (de <= 0.0 ? 1.0 : 0.0) : std::exp(std::min(700.0, -beta * de));
            if (prob > philox_uniform(S.cfg.seed, gid, 3 * it + 2, 5u)) std::swap(occ[k], un[u]);
        }
        std::sort(occ, occ + ne);
        for (int e = 0; e < ne; ++e) state[t * ne + e] = occ[e] + 1;
    }
    return set_state_impl(h, r.data(), v.data(), nullptr, nullptr, state.data(), 0, nullptr);
}

// CPU restatement of nqcb200_sample_mapping (nrpmd.jl:47-65), same Philox stream
int nqco_sample_mapping(nqco_handle* h, int32_t state) {
    if (!h || !h->has_state || h->S.cfg.method != NQCB200_METHOD_NRPMD || state < 1 || state > h->S.n) return NQCB200_ERR_INVALID;
    const Setup& S = h->S;
    const size_t per = (size_t)S.n * S.B;
    const double g = S.cfg.nrpmd_gamma;
    std::vector<double> q(per * h->traj.size()), p(per * h->traj.size());
    for (size_t t = 0; t < h->traj.size(); ++t)
        for (size_t comp = 0; comp < per; ++comp) {
            const int s = (int)(comp % S.n);
            const double theta = 6.283185307179586 * philox_uniform(S.cfg.seed, (uint64_t)(S.cfg.traj_offset + t), (uint64_t)comp, 4u);
            const double R = (s == state - 1) ? std::sqrt(2.0 + 2.0 * g) : std::sqrt(2.0 * g);
            q[t * per + comp] = R * std::cos(theta);
            p[t * per + comp] = R * std::sin(theta);
        }
    return nqco_set_mapping(h, q.data(), p.data());
}

int nqco_set_draws(nqco_handle* h, const double* xi, int64_t nsteps) {
    if (!h || !xi || nsteps < 0) return NQCB200_ERR_INVALID;
    h->draws.assign(xi, xi + (size_t)nsteps * h->traj.size());
    h->draws_first_step = h->traj.empty() ? 0 : h->traj[0].step;
    h->draws_nsteps = nsteps;
    return NQCB200_OK;
}

int nqco_set_noise(nqco_handle* h, const double* xi, int64_t nsteps) {
    if (!h || !xi || nsteps < 0) return NQCB200_ERR_INVALID;
    if (h->S.cfg.method != NQCB200_METHOD_THERMAL_LANGEVIN) return NQCB200_ERR_INVALID;
    h->noise.assign(xi, xi + (size_t)nsteps * h->traj.size() * h->S.B * h->S.D);
    h->noise_first_step = h->traj.empty() ? 0 : h->traj[0].step;
    h->noise_nsteps = nsteps;
    return NQCB200_OK;
}

int nqco_run(nqco_handle* h, int64_t nsteps) {
    if (!h) return NQCB200_ERR_INVALID;
    if (!h->has_state) { h->err = "run before set_state"; return NQCB200_ERR_STATE; }
    const Setup& S = h->S;
    const int64_t T = (int64_t)h->traj.size();
    if (T == 0) return NQCB200_OK;
    const bool needs_draws = (S.cfg.method == NQCB200_METHOD_FSSH || S.cfg.method == NQCB200_METHOD_IESH);
    int64_t step0 = h->traj[0].step;
    if (needs_draws && S.cfg.rng == NQCB200_RNG_INJECTED) {
        if (step0 < h->draws_first_step || step0 + nsteps > h->draws_first_step + h->draws_nsteps) {
            h->err = "not enough injected draws"; return NQCB200_ERR_STATE;
        }
    }
    int64_t done = 0;
    std::string err;
    while (done < nsteps) {
        // advance to the next save point (or the end) in one parallel region
        int64_t cur = step0 + done;
        int64_t to_save = S.cfg.save_every - (cur % S.cfg.save_every);
        int64_t chunk = std::min(nsteps - done, to_save);
#pragma omp parallel for schedule(static)
        for (int64_t t = 0; t < T; ++t) {
            Trajectory& tr = h->traj[t];
            try {
                for (int64_t s = 0; s < chunk; ++s) {
                    double xi = 0.0;
                    // terminate!(integrator): the trajectory is over; its frozen final state is what the later save
                    // points of the fixed-shape output carry (the step counter still advances: it indexes the draws)
                    if (tr.term_step >= 0) { tr.step++; continue; }
                    if (needs_draws) {
                        if (S.cfg.rng == NQCB200_RNG_INJECTED)
                            xi = h->draws[(size_t)(tr.step - h->draws_first_step) * T + t];
                        else
                            xi = philox_uniform(S.cfg.seed, S.cfg.traj_offset + t, (uint64_t)tr.step, 0);
                    }
                    if (S.cfg.method == NQCB200_METHOD_THERMAL_LANGEVIN) {
                        const size_t nb = (size_t)S.B * S.D;
                        tr.noise.resize(nb);
                        if (S.cfg.rng == NQCB200_RNG_INJECTED) {
                            if (tr.step < h->noise_first_step || tr.step >= h->noise_first_step + h->noise_nsteps)
                                throw std::runtime_error("not enough injected noise");
                            const double* src = &h->noise[((size_t)(tr.step - h->noise_first_step) * T + t) * nb];
                            std::copy(src, src + nb, tr.noise.begin());
                        } else {
                            for (size_t k = 0; k < nb; k += 2) {     // same stream as the CUDA kernel (purpose 3)
                                double z0, z1;
                                philox_normal2(S.cfg.seed, (uint64_t)(S.cfg.traj_offset + t), (uint64_t)tr.step * ((nb + 1) / 2) + k / 2, z0, z1, 3u);
                                tr.noise[k] = z0; if (k + 1 < nb) tr.noise[k + 1] = z1;
                            }
                        }
                    }
                    step(S, tr, xi);
                    // DiscreteCallback(condition, terminate!) runs after perform_step! and after the problem's own
                    // hopping callback (CallbackSet(prob callbacks, solve callbacks)); condition on the new u
                    if (h->term_dof >= 0) {
                        double x = tr.r[h->term_dof], vx = tr.v[h->term_dof];
                        if (S.B > 1) {      // ring polymers: the predicate sees the centroid of that dof
                            vec rc(S.D), vc(S.D);
                            centroid_of(S, tr.r, rc.data()); centroid_of(S, tr.v, vc.data());
                            x = rc[h->term_dof]; vx = vc[h->term_dof];
                        }
                        const bool og = h->term_outgoing != 0;
                        if ((x < h->term_lo && (!og || vx < 0.0)) || (x > h->term_hi && (!og || vx > 0.0)) ||
                            S.cfg.t0 + S.cfg.dt * (double)tr.step > h->term_tcut) tr.term_step = tr.step;
                    }
                }
            } catch (const std::exception& e) {
#pragma omp critical
                err = e.what();
            }
        }
        if (!err.empty()) { h->err = err; return NQCB200_ERR_INVALID; }
        done += chunk;
        if ((step0 + done) % S.cfg.save_every == 0) record_save(h, (step0 + done) / S.cfg.save_every);
    }
    return NQCB200_OK;
}

int nqco_get_state(nqco_handle* h, double* r, double* v, double* sre, double* sim, int32_t* state) {
    if (!h || !h->has_state) return NQCB200_ERR_STATE;
    const Setup& S = h->S;
    const size_t N = (size_t)S.B * S.D;
    for (size_t t = 0; t < h->traj.size(); ++t) {
        const Trajectory& tr = h->traj[t];
        if (r) std::copy(tr.r.begin(), tr.r.end(), r + t * N);
        if (v) std::copy(tr.v.begin(), tr.v.end(), v + t * N);
        const size_t ns = tr.sigma.size();
        if (sre) for (size_t i = 0; i < ns; ++i) sre[t * ns + i] = tr.sigma[i].real();
        if (sim) for (size_t i = 0; i < ns; ++i) sim[t * ns + i] = tr.sigma[i].imag();
        if (state) {
            if (S.cfg.method == NQCB200_METHOD_IESH) for (int e = 0; e < S.ne; ++e) state[t * S.ne + e] = tr.occ[e] + 1;
            else if (S.cfg.method == NQCB200_METHOD_FSSH) state[t] = tr.state + 1;
        }
    }
    return NQCB200_OK;
}

int nqco_get_mapping(nqco_handle* h, double* qmap, double* pmap) {
    if (!h || !h->has_state) return NQCB200_ERR_STATE;
    const size_t per = (size_t)h->S.n * h->S.B;
    for (size_t t = 0; t < h->traj.size(); ++t) {
        if (qmap) std::copy(h->traj[t].qmap.begin(), h->traj[t].qmap.end(), qmap + t * per);
        if (pmap) std::copy(h->traj[t].pmap.begin(), h->traj[t].pmap.end(), pmap + t * per);
    }
    return NQCB200_OK;
}

int nqco_get_observable_sum(nqco_handle* h, int obs_id, double* out, int64_t len) {
    if (!h || obs_id < 0 || obs_id >= NQCB200_OBS_COUNT || h->L.offset[obs_id] < 0) return NQCB200_ERR_INVALID;
    int64_t need = (int64_t)h->S.cfg.nsave * h->L.width[obs_id];
    if (len < need) return NQCB200_ERR_INVALID;
    std::copy(h->obs_sum.begin() + h->L.offset[obs_id], h->obs_sum.begin() + h->L.offset[obs_id] + need, out);
    return NQCB200_OK;
}

int nqco_get_observable_per_trajectory(nqco_handle* h, int obs_id, double* out, int64_t len) {
    if (!h || obs_id < 0 || obs_id >= NQCB200_OBS_COUNT || h->L.offset[obs_id] < 0 || !h->S.cfg.per_trajectory)
        return NQCB200_ERR_INVALID;
    int64_t per = (int64_t)h->S.cfg.nsave * h->L.width[obs_id];
    if (len < per * (int64_t)h->traj.size()) return NQCB200_ERR_INVALID;
    for (size_t t = 0; t < h->traj.size(); ++t)
        std::copy(h->obs_traj.begin() + t * h->L.total + h->L.offset[obs_id],
                  h->obs_traj.begin() + t * h->L.total + h->L.offset[obs_id] + per, out + t * per);
    return NQCB200_OK;
}

int nqco_get_diagnostics(nqco_handle* h, double* eig, double* nac, double* accel, double* Z) {
    if (!h || !h->has_state) return NQCB200_ERR_STATE;
    const Setup& S = h->S;
    const int n = S.n, D = S.D;
    for (size_t t = 0; t < h->traj.size(); ++t) {
        const Trajectory& tr = h->traj[t];
        const Cache& c = hop_cache(S, tr);
        if (eig) std::copy(c.w.begin(), c.w.end(), eig + t * n);
        if (nac) std::copy(c.nac.begin(), c.nac.end(), nac + t * (size_t)D * n * n);
        if (accel) std::copy(tr.k.begin(), tr.k.end(), accel + t * (size_t)S.B * D);
        if (Z) std::copy(c.Z.begin(), c.Z.end(), Z + t * (size_t)n * n);
    }
    return NQCB200_OK;
}

int nqco_get_counters(nqco_handle* h, int64_t* steps, int64_t* hops, int64_t* frustrated, int64_t* nonfinite) {
    if (!h) return NQCB200_ERR_INVALID;
    int64_t s = 0, hp = 0, fr = 0, nf = 0;
    for (const Trajectory& tr : h->traj) {
        s += tr.cnt.steps; hp += tr.cnt.hops; fr += tr.cnt.frustrated;
        bool bad = false;
        for (double x : tr.r) bad |= !std::isfinite(x);
        for (double x : tr.v) bad |= !std::isfinite(x);
        nf += bad;
    }
    if (steps) *steps = s;
    if (hops) *hops = hp;
    if (frustrated) *frustrated = fr;
    if (nonfinite) *nonfinite = nf;
    return NQCB200_OK;
}

int nqco_get_iesh_stats(nqco_handle* h, int64_t* hop_searches, int64_t* determinants, int64_t* taylor_stages,
                        int64_t* gemm_stages) {
    if (!h) return NQCB200_ERR_INVALID;
    int64_t s = 0, st = 0;
    for (const Trajectory& tr : h->traj) { s += tr.cnt.hop_searches; st += tr.cnt.steps; }
    if (hop_searches) *hop_searches = s;
    if (determinants) *determinants = (h->S.cfg.method == NQCB200_METHOD_IESH && !h->S.cfg.disable_hopping) ? st : 0;
    if (taylor_stages) *taylor_stages = 0;   // the oracle follows the reference: dense Hermitian eigendecomposition
    if (gemm_stages) *gemm_stages = 0;
    return NQCB200_OK;
}

/* TerminatingCallback with a position-window predicate (see nqcb200_set_termination). */
int nqco_set_termination(nqco_handle* h, int dof, double lo, double hi, int outgoing, double tcut) {
    if (!h) return NQCB200_ERR_INVALID;
    if (dof >= 0 && dof >= h->S.D) { h->err = "termination: dof < D"; return NQCB200_ERR_INVALID; }
    h->term_dof = dof; h->term_lo = lo; h->term_hi = hi; h->term_outgoing = outgoing ? 1 : 0;
    h->term_tcut = (tcut == tcut) ? tcut : INFINITY;
    return NQCB200_OK;
}
int nqco_get_termination(nqco_handle* h, int64_t* term_step) {
    if (!h || !term_step) return NQCB200_ERR_INVALID;
    for (size_t t = 0; t < h->traj.size(); ++t) term_step[t] = h->traj[t].term_step;
    return NQCB200_OK;
}

int nqco_get_progress(nqco_handle* h, int64_t* nsave_done, int64_t* step_count) {
    if (!h) return NQCB200_ERR_INVALID;
    if (nsave_done) *nsave_done = h->nsave_done;
    if (step_count) *step_count = h->traj.empty() ? 0 : h->traj[0].step;
    return NQCB200_OK;
}

// ---- unit-level entry points used by the known-answer tests ----------------------------------
int nqco_normal_mode_matrix(int B, double* U) { vec u = normal_mode_matrix(B); std::copy(u.begin(), u.end(), U); return 0; }
int nqco_cayley(int B, double omega_n, double dt, int half, double* out) {
    vec c = cayley_propagator(B, omega_n, dt, half != 0); std::copy(c.begin(), c.end(), out); return 0;
}
int nqco_sym_eigh(int n, const double* A, double* w, double* Z, int algo) {
    try {
        if (algo == 1) jacobi_eigh(n, A, w, Z); else if (algo == 2) tridiag_ql_eigh(n, A, w, Z); else sym_eigh(n, A, w, Z);
    } catch (const std::exception& e) { g_err = e.what(); return NQCB200_ERR_INVALID; }
    return 0;
}
int nqco_herm_eigh(int n, const double* Are, const double* Aim, double* w, double* Zre, double* Zim) {
    cvec A((size_t)n * n), Zc((size_t)n * n);
    for (int i = 0; i < n * n; ++i) A[i] = cd(Are[i], Aim[i]);
    jacobi_heig(n, A.data(), w, Zc.data());
    for (int i = 0; i < n * n; ++i) { Zre[i] = Zc[i].real(); Zim[i] = Zc[i].imag(); }
    return 0;
}
int nqco_complex_det(int n, const double* Are, const double* Aim, double* det_re, double* det_im) {
    cvec A((size_t)n * n);
    for (int i = 0; i < n * n; ++i) A[i] = cd(Are[i], Aim[i]);
    cd d = complex_det(n, A.data());
    *det_re = d.real(); *det_im = d.imag();
    return 0;
}
// model + calculator cache at one geometry: V, dV, w, Z (gauge reference identity), adiab, nac
int nqco_evaluate_model(const nqcb200_config* cfg, const double* r, double* V, double* dV, double* w, double* Z,
                        double* adiab, double* nac) {
    Model m;
    m.kind = cfg->model; m.n = cfg->nstates; m.D = cfg->ndofs;
    std::memcpy(m.p, cfg->params, sizeof(cfg->params));
    if (cfg->nbath > 0 && cfg->bath_a) m.ba.assign(cfg->bath_a, cfg->bath_a + cfg->nbath);
    if (cfg->nbath > 0 && cfg->bath_b) m.bb.assign(cfg->bath_b, cfg->bath_b + cfg->nbath);
    Cache c;
    c.init(m.n, m.D, nullptr);
    try { c.update(m, r); } catch (const std::exception& e) { g_err = e.what(); return NQCB200_ERR_INVALID; }
    const size_t nn = (size_t)m.n * m.n;
    if (V) std::copy(c.V.begin(), c.V.end(), V);
    if (dV) std::copy(c.dV.begin(), c.dV.end(), dV);
    if (w) std::copy(c.w.begin(), c.w.end(), w);
    if (Z) std::copy(c.Z.begin(), c.Z.end(), Z);
    if (adiab) std::copy(c.adiab.begin(), c.adiab.begin() + nn * m.D, adiab);
    if (nac) std::copy(c.nac.begin(), c.nac.begin() + nn * m.D, nac);
    return 0;
}
// Density-matrix sub-integration alone (electronic_dynamics.jl + Tsit5): buffers given explicitly.
int nqco_propagate_density(int n, const double* E0, const double* vd0, double t0, const double* E1, const double* vd1,
                           double t1, double t, double dt, double* sre, double* sim) {
    ElectronicParameters cur, nxt;
    cur.E.resize(n); nxt.E.resize(n); cur.vd.resize((size_t)n * n); nxt.vd.resize((size_t)n * n);
    for (int i = 0; i < n; ++i) { cur.E[i] = E0[i]; nxt.E[i] = E1[i]; }
    for (int i = 0; i < n * n; ++i) { cur.vd[i] = vd0[i]; nxt.vd[i] = vd1[i]; }
    cur.t = t0; nxt.t = t1;
    cvec s((size_t)n * n);
    for (int i = 0; i < n * n; ++i) s[i] = cd(sre[i], sim[i]);
    propagate_density(n, cur, nxt, t, dt, s);
    for (int i = 0; i < n * n; ++i) { sre[i] = s[i].real(); sim[i] = s[i].imag(); }
    return 0;
}
// select_new_state (fssh.jl:110-121): cumulative probabilities, 1-based states
int nqco_select_new_state(int n, const double* cumprob, int state, double xi) {
    vec p(cumprob, cumprob + n);
    return fssh_select(p, state - 1, xi) + 1;
}
// rescale_velocity! for one configuration: caches evaluated at r, velocities updated in place.
// returns 1 (accepted) / 0 (frustrated); eig receives the hopping eigenvalues.
int nqco_unit_rescale(const nqcb200_config* cfg, const double* r, double* v, int new_state, int old_state, double* eig) {
    nqco_handle* h = nullptr;
    nqcb200_config c = *cfg;
    c.ntraj = 1;
    if (nqco_create(&c, &h) != 0) return -1;
    Trajectory& tr = h->traj[0];
    const size_t N = (size_t)h->S.B * h->S.D;
    tr.r.assign(r, r + N); tr.v.assign(v, v + N);
    tr.sigma.assign((size_t)h->S.n * h->S.n, cd(0.0));
    initialise(h->S, tr, nullptr);
    bool ok = rescale_velocity(h->S, tr, new_state - 1, old_state - 1);
    std::copy(tr.v.begin(), tr.v.end(), v);
    if (eig) { const Cache& cc = hop_cache(h->S, tr); std::copy(cc.w.begin(), cc.w.end(), eig); }
    nqco_destroy(h);
    return ok ? 1 : 0;
}
int nqco_unoccupied(int n, int ne, const int32_t* occ, int32_t* out) {   // 1-based (DynamicsUtils.jl:162-171)
    std::vector<int> o(ne), un;
    for (int i = 0; i < ne; ++i) o[i] = occ[i] - 1;
    iesh_unoccupied(n, o, un);
    for (size_t i = 0; i < un.size(); ++i) out[i] = un[i] + 1;
    return (int)un.size();
}
// apply_decoherence_correction! (decoherence_corrections.jl:21-38) on one wavefunction column
int nqco_edc(int n, double* psi_re, double* psi_im, int occupied, double dt, const double* E, double Ekin, double C) {
    int occ = occupied - 1;
    double un = 0.0;
    for (int i = 0; i < n; ++i) {
        if (i == occ) continue;
        double tau = (1.0 + C / Ekin) / std::fabs(E[i] - E[occ]);
        double f = std::exp(-dt / tau);
        psi_re[i] *= f; psi_im[i] *= f;
        un += psi_re[i] * psi_re[i] + psi_im[i] * psi_im[i];
    }
    double nrm = psi_re[occ] * psi_re[occ] + psi_im[occ] * psi_im[occ];
    double f = std::sqrt((1.0 - un) / nrm);
    psi_re[occ] *= f; psi_im[occ] *= f;
    return 0;
}
int nqco_philox_raw(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox4x32_10(c, key[0], key[1]);
    for (int i = 0; i < 4; ++i) out[i] = c[i];
    return 0;
}
double nqco_philox_uniform(uint64_t seed, uint64_t gid, uint64_t step, uint32_t purpose) {
    return philox_uniform(seed, gid, step, purpose);
}

}  // extern "C"

/*
 * nqcb200.h -- C ABI of the B200 ensemble-trajectory engine.
 *
 * This is the drop-in boundary for NQCDynamics.jl's ensemble hot path.  The reference has NO
 * FFI for this path (it is 100 % Julia); the seam these entry points replace is the call
 *
 *     SciMLBase.solve(ensemble_problem, algorithm, ensemble_algorithm; trajectories, kwargs...)
 *                                              reference: src/Ensembles/run_dynamics.jl:91-97
 *
 * i.e. "step `trajectories` independent copies of one Simulation for nsteps of dt, evaluate the
 * outputs at the save points, hand the per-trajectory (or reduced) outputs back".  A Julia
 * `EnsembleB200 <: SciMLBase.EnsembleAlgorithm` binds these symbols with `@ccall` (stub in
 * INTEGRATION.md); the Python mirror in `nqcdynamics.jl_b200/` binds them with ctypes.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in any signature; all pointers are HOST pointers unless the
 *     function name ends in `_device`.
 *   - the caller owns every host buffer; nothing is retained after a call returns.
 *   - every function returns NQCB200_OK (0) or a negative error code; the message is available
 *     from nqcb200_last_error().  No exception, abort or sticky CUDA error crosses the boundary.
 *   - Host layout is the Julia layout, trajectory-major: for trajectory t
 *         r, v   : (ndofs, natoms, nbeads) column-major  -> flat [dof + D*bead + D*B*t],  D = ndofs*natoms
 *         sigma  : (nstates, nstates) column-major, real and imaginary parts separate
 *                  (reference: SurfaceHoppingVariables.jl:10-25 sigma_real / sigma_imag)
 *         psi    : (nstates, nelectrons) column-major (IESH; reference: iesh.jl:89-97)
 *         qmap,pmap : (nstates, nbeads) column-major (NRPMD; reference: nrpmd.jl:47-65)
 *     The library transposes to SoA ([field][component][trajectory]) on upload.
 *   - State indices are 1-based on the host side, exactly as the reference stores them
 *     (u.state as Float64, sim.method.state as Int; fssh.jl:53-63).
 *   - There is NO CPU fallback: nqcb200_create fails with NQCB200_ERR_NO_DEVICE when no sm_100
 *     class GPU is visible, and with NQCB200_ERR_UNSUPPORTED for (method, model, size)
 *     combinations that have no kernel.
 */
#ifndef NQCB200_H
#define NQCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NQCB200_ABI_VERSION 1

/* ---- error codes ------------------------------------------------------------------------- */
#define NQCB200_OK               0
#define NQCB200_ERR_INVALID     -1   /* bad argument / config                                   */
#define NQCB200_ERR_UNSUPPORTED -2   /* no kernel for this (method, model, n, D, B) combination */
#define NQCB200_ERR_NO_DEVICE   -3   /* no usable GPU (the library never computes on the CPU)   */
#define NQCB200_ERR_CUDA        -4   /* CUDA runtime error, message holds cudaGetErrorString    */
#define NQCB200_ERR_STATE       -5   /* call order violated (e.g. run before set_state)         */
#define NQCB200_ERR_NOMEM       -6

/* ---- dynamics method: reference type -> enum ---------------------------------------------- */
/* Simulation{FSSH} + BABwithTsit5          fssh.jl:23-45, bab_electronics.jl:61-91
 * RingPolymerSimulation{FSSH}+BCBwithTsit5 rpsh.jl:8-10,  bcb_electronics.jl:53-97   (nbeads>1)
 * Simulation{Ehrenfest}                    ehrenfest.jl:27-41                        (+RP: nbeads>1)
 * Simulation{AdiabaticIESH}+VerletwithElectronics  iesh.jl:26-87, verlet_with_electronics.jl:42-69
 * RingPolymerSimulation{AdiabaticIESH} / {EhrenfestNA} + BCBWavefunction  rpiesh.jl:3-52, rpehrenfest_na.jl:3-52,
 *   bcb_wavefunction.jl:37-69  (METHOD_IESH / METHOD_EHRENFEST_NA with 2 <= nbeads <= 32, ndofs == 1: bead forces from one
 *   eigenproblem per bead, psi propagated with the centroid eigenvalues / couplings / velocity of the previous geometry,
 *   centroid hop test, rescaling on every bead; no EDC, no termination mask, no gauge reference)
 * RingPolymerSimulation{Classical}+BCB     classical.jl:40-86, bcb.jl:81-116         (RPMD)
 *   (nbeads==1: Simulation{Classical} + VelocityVerlet, classical.jl:86)
 * RingPolymerSimulation{NRPMD}+RingPolymerMInt  nrpmd.jl:34-45, ringpolymer_mint.jl:28-78     */
enum nqcb200_method {
    NQCB200_METHOD_FSSH      = 1,
    NQCB200_METHOD_EHRENFEST = 2,
    NQCB200_METHOD_IESH      = 3,
    NQCB200_METHOD_CLASSICAL = 4,
    NQCB200_METHOD_NRPMD     = 5,
    /* Simulation{EhrenfestNA} + VerletwithElectronics (ehrenfest_na.jl:5-126, verlet_with_electronics.jl:76-79):
     * the AdiabaticIESH state (psi: n x ne) without occupations or hops; mean-field force
     * -sum_e <psi_e| Z' dV Z |psi_e>.  set_state: sigma = psi, state = NULL.                     */
    NQCB200_METHOD_EHRENFEST_NA = 6,
    /* RingPolymerSimulation{ThermalLangevin} + BCOCB (langevin.jl:67-82, bcocb.jl:78-120, steps.jl:109-124): thermostatted
     * ring-polymer dynamics, the sampler of the RPMD / RPSH thermal distributions.  PILE friction
     * gamma_0 = cfg.nrpmd_gamma (the struct's one gamma field), gamma_k = 2 omega_k; O-step in normal modes between two
     * half Cayley steps.  Classical (single-surface) models, nbeads >= 2.  Noise: Philox normals
     * (purpose 3) keyed by (seed; trajectory, step, mode), or injected with nqcb200_set_noise.       */
    NQCB200_METHOD_THERMAL_LANGEVIN = 7
};

/* ---- analytic model Hamiltonians (NQCModels.jl, external to the reference tree) ------------ */
/* params[] meaning per model (defaults are the NQCModels defaults):
 *  TULLY_ONE   {a,b,c,d}           V11=sgn(q) a (1-exp(-b|q|)), V22=-V11, V12=c exp(-d q^2)
 *  TULLY_TWO   {a,b,c,d,e}         V11=0, V22=-a exp(-b q^2)+e, V12=c exp(-d q^2)
 *  TULLY_THREE {a,b,c}             V11=a, V22=-a, V12= b exp(c q) (q<0) | b (2-exp(-c q)) (q>=0)
 *  DOUBLE_WELL {mass,omega,gamma,delta}  V11/22 = 1/2 m w^2 q^2 +- sqrt(2) gamma q, V12 = delta/2
 *  SPIN_BOSON  {epsilon,delta}; bath_a=omega_j[D], bath_b=c_j[D]
 *              V11/22 = +-(eps + sum c_j r_j) + sum 1/2 w_j^2 r_j^2, V12 = delta
 *  THREE_STATE_MORSE {d1,d2,d3, alpha1..3, r1..3, c1..3, a12,a13,a23, alpha12,alpha13,alpha23,
 *                     r12,r13,r23}  V_ii = d_i (1-exp(-alpha_i (q-r_i)))^2 + c_i,
 *                                   V_ij = a_ij exp(-alpha_ij (q-r_ij)^2)
 *  HARMONIC    {m,omega,r0}        classical, V = sum_dof 1/2 m w^2 (q-r0)^2
 *  FREE        {}                  classical, V = 0
 *  ANDERSON_HOLSTEIN_MIAO_SUBOTNIK {m,omega,g,DeltaG}; bath_a=eps_k[M], bath_b=V_k[M]
 *              U0 = 1/2 m w^2 q^2 (state independent), h = U1-U0, U1 = 1/2 m w^2 (q-g)^2 + DeltaG
 *              H[0,0]=h(q), H[k,k]=eps_k, H[0,k]=H[k,0]=V_k ; nstates = M+1   (iesh.md:71-76)
 *  ANDERSON_HOLSTEIN_ERPENBECK_THOSS {De,a,x0,c, D1,D2,a1,x01,Vinf, q,atilde,xtilde}; bath_a=eps_k[M], bath_b=Vbar_k[M]
 *              the impurity of the reference's own IESH tests and example (test/Dynamics/iesh.jl:23, iesh.md:85-105):
 *              U0 = De (exp(-a (x-x0)) - 1)^2 + c,  U1 = D1 exp(-2 a1 (x-x01)) - D2 exp(-a1 (x-x01)) + Vinf,  h = U1-U0,
 *              and a POSITION-DEPENDENT coupling H[0,k] = Vbar_k f(x), f = (1-q)/2 (1 - tanh((x-xtilde)/atilde)) + q
 *              (NQCModels ErpenbeckThoss, external; formula and defaults recalled, carried as explicit params)      */
enum nqcb200_model {
    NQCB200_MODEL_TULLY_ONE         = 1,
    NQCB200_MODEL_TULLY_TWO         = 2,
    NQCB200_MODEL_TULLY_THREE       = 3,
    NQCB200_MODEL_DOUBLE_WELL       = 4,
    NQCB200_MODEL_SPIN_BOSON        = 5,
    NQCB200_MODEL_THREE_STATE_MORSE = 6,
    NQCB200_MODEL_HARMONIC          = 7,
    NQCB200_MODEL_FREE              = 8,
    NQCB200_MODEL_ANDERSON_HOLSTEIN_MIAO_SUBOTNIK = 9,
    NQCB200_MODEL_ANDERSON_HOLSTEIN_ERPENBECK_THOSS = 10
};

/* frustrated-hop policy: surface_hopping.jl:65,79-91 */
enum nqcb200_rescaling {
    NQCB200_RESCALE_STANDARD   = 0,
    NQCB200_RESCALE_VINVERSION = 1,
    NQCB200_RESCALE_OFF        = 2
};

/* Random numbers for the hop test (reference: one rand() per trajectory per step,
 * fssh.jl:112, iesh.jl:393).  PHILOX: counter-based Philox4x32-10 keyed by (seed; global
 * trajectory id, step) -> results independent of sharding.  INJECTED: draws supplied by
 * nqcb200_set_draws (parity mode, identical hop sequences in engine and oracle).              */
enum nqcb200_rng {
    NQCB200_RNG_PHILOX   = 0,
    NQCB200_RNG_INJECTED = 1
};

/* ---- observables evaluated on the device at every save point ------------------------------ */
/* Each maps to a reference output functor / estimator; `width` doubles per save point.
 *  ADIABATIC_POP   n        Estimators.adiabatic_population  fssh.jl:144-148, ehrenfest.jl:70-73, iesh.jl:371-375
 *  DIABATIC_POP    n        Estimators.diabatic_population   fssh.jl:132-142, ehrenfest.jl:75-83, iesh.jl:337-369, nrpmd.jl:111-122
 *  POPCORR_DIABATIC  n*n    PopulationCorrelationFunction{Diabatic}  TimeCorrelationFunctions.jl:26-40,86-88
 *                           out[i + n*j] = P_i(0) * P_j(t)
 *  POPCORR_ADIABATIC n*n    PopulationCorrelationFunction{Adiabatic}
 *  KINETIC         1        OutputKineticEnergy  DynamicsOutputs.jl:98  (DynamicsUtils.jl:108-135)
 *  POTENTIAL       1        OutputPotentialEnergy :82 (classical_potential_energy per method)
 *  TOTAL_ENERGY    1        OutputTotalEnergy :90 (classical_hamiltonian, includes RP spring energy)
 *  POSITION        D        OutputPosition :39 / OutputCentroidPosition :47 for ring polymers
 *  VELOCITY        D        OutputVelocity :66 / OutputCentroidVelocity :74
 *  DISCRETE_STATE  1 (FSSH) / ne (IESH)   OutputDiscreteState :178
 *  SCATTERING      2n       OutputStateResolvedScattering1D(:adiabatic) :313-338 -- final state only:
 *                           [reflection(n), transmission(n)], transmission iff r[0] > 0
 *  SCATTERING_DIABATIC 2n   same with type=:diabatic
 *  SIGMA           2*n*n    OutputQuantumSubsystem :149 (re then im, column-major)
 *  MAPPING_Q       n*B      OutputMappingPosition :157 (NRPMD; (nstates, nbeads) column-major, as nqcb200_set_mapping)
 *  MAPPING_P       n*B      OutputMappingMomentum :165                                             */
enum nqcb200_observable {
    NQCB200_OBS_ADIABATIC_POP       = 0,
    NQCB200_OBS_DIABATIC_POP        = 1,
    NQCB200_OBS_POPCORR_DIABATIC    = 2,
    NQCB200_OBS_POPCORR_ADIABATIC   = 3,
    NQCB200_OBS_KINETIC             = 4,
    NQCB200_OBS_POTENTIAL           = 5,
    NQCB200_OBS_TOTAL_ENERGY        = 6,
    NQCB200_OBS_POSITION            = 7,
    NQCB200_OBS_VELOCITY            = 8,
    NQCB200_OBS_DISCRETE_STATE      = 9,
    NQCB200_OBS_SCATTERING          = 10,
    NQCB200_OBS_SCATTERING_DIABATIC = 11,
    NQCB200_OBS_SIGMA               = 12,
    NQCB200_OBS_MAPPING_Q           = 13,
    NQCB200_OBS_MAPPING_P           = 14,
    NQCB200_OBS_COUNT               = 15
};

#define NQCB200_MAX_PARAMS 32

/* Flat POD description of one ensemble run.  Mirrors what the reference spreads over
 * Simulation / RingPolymerSimulation (simulations.jl:12-73), the method constructor kwargs
 * (fssh.jl:41, iesh.jl:72-74, nrpmd.jl:43) and the run_dynamics kwargs (run_dynamics.jl:42-56). */
typedef struct nqcb200_config {
    int32_t  abi_version;       /* must be NQCB200_ABI_VERSION                                   */
    int32_t  method;            /* enum nqcb200_method                                           */
    int32_t  model;             /* enum nqcb200_model                                            */
    int32_t  nstates;           /* n  (1 for classical models)                                   */
    int32_t  ndofs;             /* D = ndofs*natoms, nuclear degrees of freedom per replica      */
    int32_t  nbeads;            /* B  (1 = plain Simulation)                                     */
    int32_t  nelectrons;        /* ne (IESH only)                                                */
    int32_t  rescaling;         /* enum nqcb200_rescaling                                        */
    int32_t  estimate_probability; /* IESH pruning, iesh.jl:251-254 (default 1)                  */
    int32_t  disable_hopping;   /* IESH, iesh.jl:392                                             */
    int32_t  rng;               /* enum nqcb200_rng                                              */
    int32_t  device;            /* CUDA device ordinal this handle runs on                       */
    int32_t  save_every;        /* save point every `save_every` steps (saveat = k*dt); >=1      */
    int32_t  nsave;             /* capacity: number of save points incl. t0 (steps/save_every+1) */
    int32_t  per_trajectory;    /* 0: observables summed over trajectories on device
                                   1: additionally keep every trajectory's values (output stream) */
    int32_t  diagnostics;       /* 1: keep eigenvalues / NAC / acceleration of the last step     */
    uint32_t observables;       /* bitmask of (1u << enum nqcb200_observable)                    */
    uint32_t reserved0;
    int64_t  ntraj;             /* trajectories owned by this handle                             */
    int64_t  traj_offset;       /* global index of local trajectory 0 (RNG key, sharding)        */
    uint64_t seed;
    double   dt;
    double   t0;                /* tspan[1]                                                      */
    double   temperature;       /* ring polymer: omega_n = nbeads*temperature (ring_polymer.jl:18) */
    double   nrpmd_gamma;       /* nrpmd.jl:43                                                   */
    double   edc_C;             /* >0: EDC decoherence constant (decoherence_corrections.jl:14)  */
    double   params[NQCB200_MAX_PARAMS];
    const double* masses;       /* [D] mass per nuclear degree of freedom                        */
    const double* bath_a;       /* model array a (see enum nqcb200_model), may be NULL           */
    const double* bath_b;       /* model array b                                                 */
    int32_t  nbath;             /* length of bath_a / bath_b                                     */
    int32_t  reserved1;
} nqcb200_config;

typedef struct nqcb200_handle nqcb200_handle;

/* Library / ABI version (NQCB200_ABI_VERSION). */
int nqcb200_version(void);

/* Number of CUDA devices visible to the library (0 when none: every other call then fails). */
int nqcb200_device_count(void);

/* Create an engine for one shard of trajectories on cfg->device.  Copies everything it needs
 * out of *cfg (masses, bath arrays).  Replaces: Simulation construction + create_problem +
 * alg_cache (simulations.jl:28-44, SurfaceHoppingMethods.jl:75-79, bab_electronics.jl:15-46). */
int nqcb200_create(const nqcb200_config* cfg, nqcb200_handle** out);
int nqcb200_destroy(nqcb200_handle* h);

/* Message of the last failing call on this handle (h may be NULL: last create error). */
const char* nqcb200_last_error(const nqcb200_handle* h);

/* Width (doubles per save point per trajectory) of an observable for this handle's config. */
int nqcb200_observable_width(const nqcb200_handle* h, int obs_id);

/* Upload initial DynamicsVariables for all ntraj trajectories and (re)initialise the integrator
 * caches: update_cache!(r), acceleration k0, all-zero electronic double buffer (quirk Q1,
 * electronic_dynamics.jl:118-127), eigenvector gauge reference, step counter = 0, observable
 * accumulators = 0, save point 0 recorded.
 *   sig_re/sig_im: density matrix (FSSH/Ehrenfest, n*n per trajectory) or psi (IESH, n*ne);
 *                  NULL for CLASSICAL/NRPMD.  AdiabaticIESH / EhrenfestNA with sig_re == NULL: electron e starts in the
 *                  adiabatic orbital state[e] (psi[state[e], e] = 1; state == NULL: orbitals 1..ne), which is what
 *                  DynamicsVariables(sim, v, r) and DynamicsVariables(sim, v, r, FermiDiracState{Adiabatic}) build
 *                  (iesh.jl:89-128) -- psi is then filled on the device and only the occupations are uploaded.
 *   state: 1-based active state (FSSH: 1 per trajectory; IESH: ne sorted occupied states);
 *          NULL for methods without a discrete state.
 * Replaces: prob_func/sample_distribution output -> integrator init
 *           (selections.jl:38-101, bab_electronics.jl:48-59).                                   */
int nqcb200_set_state(nqcb200_handle* h, const double* r, const double* v,
                      const double* sig_re, const double* sig_im, const int32_t* state);

/* Same, but the electronic state is given in the DIABATIC basis (PureState(i) / MixedState,
 * the default basis of NQCDistributions): rho is rotated on the device, sigma = Z' rho Z at r0
 * (density_matrix_dynamics.jl:37-46,64-75).  FSSH: if state == NULL the active state is sampled
 * with weights Re diag(sigma) (fssh.jl:53-54, StatsBase.sample(Weights)) from state_draw[traj]
 * (uniform [0,1)) or, when state_draw == NULL, from Philox (purpose 1).                          */
int nqcb200_set_state_diabatic(nqcb200_handle* h, const double* r, const double* v,
                               const double* rho_re, const double* rho_im, const int32_t* state,
                               const double* state_draw);

/* NRPMD mapping variables (nstates, nbeads) per trajectory; call after nqcb200_set_state. */
int nqcb200_set_mapping(nqcb200_handle* h, const double* qmap, const double* pmap);

/* Optional: eigenvector gauge reference Z_ref (n*n column-major per trajectory [, per bead and
 * centroid]) replacing the identity default in the column-sign continuity rule of
 * NQCCalculators (dot(Z_new[:,i], Z_old[:,i]) < 0 -> flip).  Lets the Julia shim hand over
 * sim.cache.eigen.Z so the engine continues in LAPACK's gauge.  Call before nqcb200_set_state. */
int nqcb200_set_gauge_reference(nqcb200_handle* h, const double* Z, int64_t count_per_traj);

/* Parity mode (cfg.rng == INJECTED): xi[step*ntraj + traj], uniform [0,1) draws consumed one per
 * trajectory per step starting at the current step counter.                                    */
int nqcb200_set_draws(nqcb200_handle* h, const double* xi, int64_t nsteps);

/* ThermalLangevin parity mode (cfg.rng == INJECTED): standard normals xi[(step*ntraj + traj)*nbeads + mode], one per
 * ring-polymer normal mode per step (W.dW / sqrt(dt) of steps.jl:116-119), starting at the current step counter.   */
int nqcb200_set_noise(nqcb200_handle* h, const double* xi, int64_t nsteps);

/* TerminatingCallback(func) = DiscreteCallback(func, terminate!) (src/DynamicsUtils/callbacks.jl:29, exported
 * DynamicsUtils.jl:156; the scattering examples pass it as `callback=` to run_dynamics) for the predicate family a
 * device kernel can evaluate: func(u, t, integrator) = r[dof] < lo || r[dof] > hi (0-based dof; "the particle has left
 * the interaction region"); outgoing != 0 additionally asks for an outward velocity, (r < lo && v < 0) || (r > hi &&
 * v > 0), the form of the IESH scattering example (docs/src/dynamicssimulations/dynamicsmethods/iesh.md:127-138:
 * mean(r) > 5.5 A && mean(v) > 0), and `|| t > tcut` is its time clause (t = t0 + steps*dt; +INFINITY or NaN: none;
 * lo = -INFINITY, hi = +INFINITY leaves only the time clause).  Checked after every step, after the method's own hopping callback (DiffEq merges the
 * callbacks as CallbackSet(problem callbacks, solve callbacks)).  A terminated trajectory stops stepping; its final
 * state is what nqcb200_get_state returns and what every later save point of the fixed-shape observable arrays
 * carries, so final-state outputs (OutputFinal, OutputStateResolvedScattering1D, DynamicsOutputs.jl:205-338) are the
 * reference's, and the host trims per-trajectory series with nqcb200_get_termination (OutputFinalTime :226-231 =
 * t0 + term_step*dt).  dof < 0 removes the callback.  Call before nqcb200_run; the flags are reset by set_state.
 * Available for the thread-per-trajectory FSSH / Ehrenfest kernels (1-D models; ring polymers -- RPSH / RP-Ehrenfest --
 * evaluate the predicate on the centroid of that dof and its centroid velocity) and for the AdiabaticIESH / EhrenfestNA
 * kernel with nbeads == 1 (whose CTA moves on to its next trajectory), otherwise NQCB200_ERR_UNSUPPORTED (SpinBoson bath
 * kernels, classical RPMD / NRPMD / Langevin, ring-polymer IESH). */
int nqcb200_set_termination(nqcb200_handle* h, int dof, double lo, double hi, int outgoing, double tcut);
/* term_step[traj]: number of steps the trajectory took before terminate! fired, -1 while it is still running. */
int nqcb200_get_termination(nqcb200_handle* h, int64_t* term_step);

/* Advance every trajectory by nsteps of dt (blocking).  Order inside a step follows the
 * reference: perform_step! -> hop callback -> save.                              */
int nqcb200_run(nqcb200_handle* h, int64_t nsteps);

/* One batch of the ensemble in one call: nqcb200_set_state (diabatic == 0: sigma / state in the adiabatic basis) or
 * nqcb200_set_state_diabatic (diabatic != 0) followed by nqcb200_run(nsteps) -- what EnsembleB200's __solve does per
 * batch with the arrays prob_func produced (selections.jl:38-101 -> solve, run_dynamics.jl:91-97).  For kernel
 * families with a launch-fused initialisation (SpinBoson FSSH / Ehrenfest) the step kernel itself reads r and v in
 * the caller's trajectory-major layout: from PINNED (cudaHostAlloc / cudaHostRegister) memory in place over PCIe,
 * block by block, overlapping the dynamics of the blocks already loaded; pageable memory is staged with one
 * cudaMemcpy.  Blocking; r and v must stay valid until it returns.  Other kernel families: identical to the two
 * separate calls.                                                                                */
int nqcb200_run_from_host(nqcb200_handle* h, const double* r, const double* v, const double* rho_re,
                          const double* rho_im, const int32_t* state, const double* state_draw, int diabatic,
                          int64_t nsteps);

/* Device-side initial conditions: sample_distribution (selections.jl:51-101) for the
 * distributions the ensemble configs use -- every nuclear component either fixed or Normal(mean, sd), which covers
 * DynamicalDistribution(v, r, size) of numbers / Normal / VelocityBoltzmann / harmonic Wigner -- times
 * PureState(i) in either basis.  Nothing but the specification crosses PCIe.
 *   r_dist, v_dist: nbeads*ndofs entries, [bead][dof]; kind 0: value a; kind 1: a + b * z, z ~ N(0, 1).
 *   normal_modes != 0: the entries describe ring-polymer NORMAL-MODE coordinates (the exact thermal sample of a
 *                      free / harmonic ring polymer); the engine transforms to beads with U (x_j = sum_k U[j,k] y_k).
 *   rho_re, rho_im:  ONE n x n column-major matrix shared by all trajectories (NULL for methods without sigma);
 *                    diabatic != 0 as in nqcb200_set_state_diabatic.
 *   state:           1-based active state, or 0 to sample it from diag(sigma) (FSSH, fssh.jl:53-54; Philox purpose 1).
 * The normals come from Philox4x32-10 keyed by (seed; global trajectory id, component, purpose 2) through Box-Muller
 * (first normal -> position, second -> velocity), so the sample does not depend on the sharding; the CPU oracle
 * implements the same stream.  Not available for AdiabaticIESH (NQCB200_ERR_UNSUPPORTED).                          */
typedef struct nqcb200_dist {
    int32_t kind;       /* 0 = fixed value a, 1 = Normal(mean a, standard deviation b) */
    int32_t reserved;
    double  a, b;
} nqcb200_dist;
int nqcb200_sample_state(nqcb200_handle* h, const nqcb200_dist* r_dist, const nqcb200_dist* v_dist, int normal_modes,
                         const double* rho_re, const double* rho_im, int diabatic, int32_t state);

/* Device-side electronic initial conditions (second half).  Both are called AFTER the nuclei are in
 * place (nqcb200_set_state / nqcb200_sample_state) and re-record save point 0; the CPU oracle implements the same streams.
 *
 * nqcb200_sample_occupations -- AdiabaticIESH with FermiDiracState{Adiabatic} (iesh.jl:99-128): the occupied adiabatic
 *   orbitals of every trajectory are drawn by the reference's Metropolis walk over orbital swaps
 *   (sample_fermi_dirac_distribution, DynamicsUtils.jl:194-208, Boltzmann-factor variant: nstates * nelectrons proposals
 *   "occupied k <-> unoccupied u", accepted when exp(-beta (E_u - E_k)) > rand()) on the adiabatic energies at r0, then
 *   sorted; psi[state[e], e] = 1.  beta = 1 / kT in atomic units (INFINITY: only downhill / level swaps are accepted).
 *   Uniforms: Philox4x32-10 keyed by (seed; global trajectory id, 3 * proposal + {0, 1, 2}, purpose 5): index of the
 *   occupied orbital, index in the CURRENT list of unoccupied orbitals (a swap exchanges the two list entries), acceptance.
 *
 * nqcb200_sample_mapping -- NRPMD with PureState{Diabatic}(state) (nrpmd.jl:47-65): theta ~ U[0, 2 pi) per (state, bead),
 *   (q, p) = R (cos theta, sin theta), R = sqrt(2 + 2 gamma) on the occupied (1-based) state and sqrt(2 gamma) elsewhere.
 *   Uniforms: Philox keyed by (seed; global trajectory id, state_index + nstates * bead, purpose 4).                       */
int nqcb200_sample_occupations(nqcb200_handle* h, double beta);
int nqcb200_sample_mapping(nqcb200_handle* h, int32_t state);

/* Download the current DynamicsVariables (any pointer may be NULL to skip that field). */
int nqcb200_get_state(nqcb200_handle* h, double* r, double* v,
                      double* sig_re, double* sig_im, int32_t* state);
int nqcb200_get_mapping(nqcb200_handle* h, double* qmap, double* pmap);

/* Observable summed over this handle's trajectories: out[isave*width + k], isave < nsave_done.
 * (SumReduction, reductions.jl:12-31; MeanReduction divides by the global trajectory count.)   */
int nqcb200_get_observable_sum(nqcb200_handle* h, int obs_id, double* out, int64_t len);

/* Same accumulator as a DEVICE pointer (ntotal doubles, all enabled observables packed in enum
 * order, each [nsave][width]) so a host can all-reduce it in place with NCCL across shards.    */
int nqcb200_observable_sum_device(nqcb200_handle* h, double** dev_ptr, int64_t* ntotal);
int nqcb200_observable_offset(const nqcb200_handle* h, int obs_id, int64_t* offset);

/* Per-trajectory values (cfg.per_trajectory=1): out[(traj*nsave + isave)*width + k]
 * (SortByTrajectoryReduction, reductions.jl:54-55).                                             */
int nqcb200_get_observable_per_trajectory(nqcb200_handle* h, int obs_id, double* out, int64_t len);

/* Diagnostics of the most recent step (cfg.diagnostics=1), trajectory-major:
 *   eig   [traj][n]          adiabatic energies w (centroid for ring polymers)
 *   nac   [traj][D][n*n]     nonadiabatic coupling d_I, column-major (centroid for ring polymers)
 *   accel [traj][B][D]       acceleration k used for the next half kick
 *   Z     [traj][n*n]        eigenvectors in the engine's gauge                                 */
int nqcb200_get_diagnostics(nqcb200_handle* h, double* eig, double* nac, double* accel, double* Z);

/* Counters summed over this handle's trajectories (SURVEY.md section 5 metrics row). */
int nqcb200_get_counters(nqcb200_handle* h, int64_t* steps, int64_t* hops, int64_t* frustrated,
                         int64_t* nonfinite);

/* AdiabaticIESH work counters, summed over this handle's trajectories since set_state (any pointer may be NULL):
 *   hop_searches   trajectory-steps on which the pruning estimate (iesh.jl:251-254) did NOT rule out a hop, i.e. on
 *                  which all ne*(n-ne) hopping probabilities were evaluated (iesh.jl:256-266)
 *   determinants   trajectory-steps on which det S was evaluated (the others were pruned by a rigorous bound)
 *   taylor_stages  polynomial stages spent on propagate_wavefunction! (wavefunction_dynamics.jl:15-58)
 *   gemm_stages    of which stages that needed the dense v.d product
 * All zero for the other methods.                                                                  */
int nqcb200_get_iesh_stats(nqcb200_handle* h, int64_t* hop_searches, int64_t* determinants,
                           int64_t* taylor_stages, int64_t* gemm_stages);

/* Number of save points recorded so far, and device time (ms, CUDA events on the launch stream)
 * spent in step kernels by the last nqcb200_run, with the number of kernel launches it made.   */
int nqcb200_get_progress(nqcb200_handle* h, int64_t* nsave_done, int64_t* step_count);
int nqcb200_get_last_run_timing(nqcb200_handle* h, double* kernel_ms, int64_t* launches);
/* The last per-trajectory download (get_state / get_observable_per_trajectory / ...): device time of the
 * SoA -> trajectory-major transposition kernel (HBM-bound) and of the device-to-host copy, and the bytes moved. */
int nqcb200_get_last_download_timing(nqcb200_handle* h, double* transpose_ms, double* copy_ms, int64_t* bytes);
/* Every kernel this handle has launched since nqcb200_create (uploads, sampling, init, step, fold, ...). */
int nqcb200_get_launch_count(nqcb200_handle* h, int64_t* launches_total);

/* Measured FP64 FMA peak of `device` in TFLOP/s (a DFMA-saturating microbenchmark; the roofline
 * denominator for this FP64 path -- MEASURED_PEAKS.json only carries HBM and bf16 numbers).      */
int nqcb200_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* NQCB200_H */

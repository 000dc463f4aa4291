"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (north_star): per-step deterministic quantities within 1e-10 relative in FP64; identical hop
sequences when the same uniform draws are injected into both implementations.
"""
import numpy as np
import pytest

import nqcdynamics_jl_b200 as nq
from helpers import A, ALL_POP_OBS, CLASSICAL_OBS, engine_factory, make_pair, model_config, oracle_factory, rel_err

pytestmark = pytest.mark.gpu

STEP_TOL = 1e-10


def _pure_state(T, n, i):
    rho = np.zeros((T, n, n))
    rho[:, i, i] = 1.0
    return rho


def _compare_state(e, o, tol, what=""):
    se, so = e.get_state(), o.get_state()
    for key in ("r", "v"):
        assert rel_err(se[key], so[key]) < tol, f"{what} {key}"
    if "sigma" in so:
        assert np.max(np.abs(se["sigma"] - so["sigma"])) < tol, f"{what} sigma"
    if "state" in so:
        assert np.array_equal(se["state"], so["state"]), f"{what} discrete state"


def _compare_observables(e, o, obs_mask, tol, T):
    for oid in range(A.OBS_COUNT):
        if obs_mask & (1 << oid):
            a, b = e.observable_sum(oid), o.observable_sum(oid)
            assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b))), f"observable {oid}"


SCATTER_MODELS = [
    ("tully1", nq.TullyModelOne(), 2000.0, -5.0, 10.0 / 2000),
    ("tully2", nq.TullyModelTwo(), 2000.0, -8.0, 16.0 / 2000),
    ("tully3", nq.TullyModelThree(), 2000.0, -10.0, 10.0 / 2000),
    ("doublewell", nq.DoubleWell(), 1.0, -0.5, 0.7),
    ("morse3", nq.ThreeStateMorse(), 20000.0, 2.1, 0.0),
]


@pytest.mark.parametrize("method", [A.METHOD_FSSH, A.METHOD_EHRENFEST])
@pytest.mark.parametrize("name,model,mass,r0,v0", SCATTER_MODELS)
def test_per_step_parity(method, name, model, mass, r0, v0):
    """Step by step: r, v, sigma, state, eigenvalues, NAC, acceleration, eigenvectors within 1e-10."""
    T, nsteps = 96, 40
    rng = np.random.default_rng(7)
    dt = 1.0 if mass > 100 else 0.05
    kw = model_config(model, method=method, masses=[mass], ntraj=T, dt=dt, rng=A.RNG_INJECTED, diagnostics=1,
                      save_every=1, nsave=nsteps + 1, observables=ALL_POP_OBS, per_trajectory=1)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    r = r0 + 0.3 * rng.standard_normal(T)
    v = v0 * (1 + 0.1 * rng.standard_normal(T)) + (0.0 if v0 else 1e-4 * rng.standard_normal(T))
    rho = _pure_state(T, model.nstates, 0)
    draws = rng.random((nsteps, T))
    sdraw = rng.random(T)
    for h in (e, o):
        h.set_state_diabatic(r, v, rho, None, None, sdraw)
        h.set_draws(draws)
    _compare_state(e, o, STEP_TOL, "t0")
    for step in range(nsteps):
        e.run(1); o.run(1)
        _compare_state(e, o, STEP_TOL, f"step {step}")
        de, do = e.diagnostics(), o.diagnostics()
        assert rel_err(de["eig"], do["eig"]) < STEP_TOL
        assert np.max(np.abs(de["Z"] - do["Z"])) < STEP_TOL
        assert rel_err(de["nac"], do["nac"]) < STEP_TOL
        assert rel_err(de["accel"], do["accel"]) < STEP_TOL
    _compare_observables(e, o, ALL_POP_OBS, 1e-9, T)
    for oid in (A.OBS_DIABATIC_POP, A.OBS_TOTAL_ENERGY):
        assert np.max(np.abs(e.observable_per_trajectory(oid) - o.observable_per_trajectory(oid))) < 1e-9
    assert e.counters()["hops"] == o.counters()["hops"]
    assert e.counters()["frustrated"] == o.counters()["frustrated"]


def test_tully_fssh_long_run_identical_hops():
    """BASELINE config 1 (TullyModelOne FSSH, 1000 trajectories, 3000 steps): same draws -> same hop sequence."""
    T, nsteps, save_every = 1000, 3000, 10
    rng = np.random.default_rng(11)
    model = nq.TullyModelOne()
    obs = ALL_POP_OBS | (1 << A.OBS_DISCRETE_STATE)
    kw = model_config(model, method=A.METHOD_FSSH, masses=[2000.0], ntraj=T, dt=1.0, rng=A.RNG_INJECTED,
                      save_every=save_every, nsave=nsteps // save_every + 1, observables=obs, per_trajectory=1)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    r = rng.normal(-8.0, 1.0, T)
    v = np.full(T, 10.0 / 2000)
    rho = _pure_state(T, 2, 1)
    draws = rng.random((nsteps, T))
    sdraw = rng.random(T)
    for h in (e, o):
        h.set_state_diabatic(r, v, rho, None, None, sdraw)
        h.set_draws(draws)
        h.run(nsteps)
    se = e.observable_per_trajectory(A.OBS_DISCRETE_STATE)
    so = o.observable_per_trajectory(A.OBS_DISCRETE_STATE)
    assert np.array_equal(se, so), "hop sequences differ"
    assert e.counters()["hops"] == o.counters()["hops"] > 0
    _compare_state(e, o, 1e-8, "final")
    _compare_observables(e, o, obs, 1e-9, T)
    scat = e.observable_sum(A.OBS_SCATTERING)[-1] / T
    assert abs(scat.sum() - 1.0) < 1e-12


def test_philox_streams_match_oracle():
    """Production RNG: engine and oracle implement the same Philox4x32-10 keyed by (seed, trajectory, step)."""
    T, nsteps = 512, 1500
    model = nq.TullyModelOne()
    obs = (1 << A.OBS_DISCRETE_STATE) | (1 << A.OBS_ADIABATIC_POP)
    kw = model_config(model, method=A.METHOD_FSSH, masses=[2000.0], ntraj=T, dt=1.0, rng=A.RNG_PHILOX, seed=20261017,
                      traj_offset=12345, save_every=25, nsave=nsteps // 25 + 1, observables=obs, per_trajectory=1)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    rng = np.random.default_rng(3)
    r = rng.normal(-6.0, 0.5, T); v = np.full(T, 12.0 / 2000)
    rho = _pure_state(T, 2, 0)
    for h in (e, o):
        h.set_state_diabatic(r, v, rho)
        h.run(nsteps)
    assert np.array_equal(e.observable_per_trajectory(A.OBS_DISCRETE_STATE), o.observable_per_trajectory(A.OBS_DISCRETE_STATE))
    assert e.counters()["hops"] == o.counters()["hops"] > 0


def test_sharding_independence():
    """Philox keyed by the global trajectory id: two shards reproduce the single-shard result exactly."""
    T, nsteps = 256, 800
    model = nq.TullyModelOne()
    obs = (1 << A.OBS_DISCRETE_STATE)
    rng = np.random.default_rng(5)
    r = rng.normal(-6.0, 0.5, T); v = np.full(T, 12.0 / 2000)
    rho = _pure_state(T, 2, 0)
    mk = engine_factory()

    def run(lo, hi):
        kw = model_config(model, method=A.METHOD_FSSH, masses=[2000.0], ntraj=hi - lo, dt=1.0, seed=99, traj_offset=lo,
                          save_every=nsteps, nsave=2, observables=obs, per_trajectory=1)
        cfg, keep = A.make_config(**kw)
        h = mk(cfg, keep)
        h.set_state_diabatic(r[lo:hi], v[lo:hi], rho[lo:hi])
        h.run(nsteps)
        return h.get_state()

    full = run(0, T)
    a, b = run(0, 100), run(100, T)
    for key in ("r", "v", "state"):
        assert np.array_equal(np.concatenate([a[key], b[key]]), full[key])


@pytest.mark.parametrize("method", [A.METHOD_FSSH, A.METHOD_EHRENFEST])
@pytest.mark.parametrize("nmodes", [3, 8, 100])
def test_spin_boson_parity(method, nmodes):
    """BASELINE config 2: SpinBoson with a Debye bath, lanes-over-modes kernels."""
    T, nsteps = 48, 60
    rng = np.random.default_rng(13)
    model = nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), nmodes, 0.0, 1.0)
    obs = ALL_POP_OBS & ~((1 << A.OBS_SCATTERING) | (1 << A.OBS_SCATTERING_DIABATIC))
    kw = model_config(model, method=method, masses=np.ones(nmodes), ntraj=T, dt=0.1, rng=A.RNG_INJECTED, diagnostics=1,
                      save_every=5, nsave=nsteps // 5 + 1, observables=obs)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    w = model.bath_a
    beta = 5.0
    sr = np.sqrt(1.0 / (2 * w * np.tanh(beta * w / 2))); sv = np.sqrt(w / (2 * np.tanh(beta * w / 2)))
    r = rng.standard_normal((T, nmodes)) * sr
    v = rng.standard_normal((T, nmodes)) * sv
    rho = _pure_state(T, 2, 0)
    draws = rng.random((nsteps, T)); sdraw = rng.random(T)
    for h in (e, o):
        h.set_state_diabatic(r, v, rho, None, None, sdraw)
        h.set_draws(draws)
    for chunk in range(nsteps // 5):
        e.run(5); o.run(5)
        _compare_state(e, o, 1e-9, f"chunk {chunk}")
        de, do = e.diagnostics(), o.diagnostics()
        assert rel_err(de["eig"], do["eig"]) < 1e-9
        assert rel_err(de["nac"], do["nac"]) < 1e-9
        assert rel_err(de["accel"], do["accel"]) < 1e-9
    _compare_observables(e, o, obs, 1e-9, T)


@pytest.mark.parametrize("method", [A.METHOD_FSSH, A.METHOD_EHRENFEST])
@pytest.mark.parametrize("nmodes,pinned", [(3, False), (37, True), (100, True), (100, False)])
def test_spin_boson_run_from_host(method, nmodes, pinned):
    """nqcb200_run_from_host: launch-fused initialisation straight from the caller's trajectory-major arrays
    (pinned: read in place over PCIe; pageable: staged) == set_state_diabatic + run == the oracle.  Population
    observables only (the kernel then skips the harmonic shift of the eigenvalues), ragged last block."""
    T, nsteps = 300, 45
    rng = np.random.default_rng(17)
    model = nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), nmodes, 0.1, 1.0)
    obs = (1 << A.OBS_POPCORR_DIABATIC) | (1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_SIGMA)
    kw = model_config(model, method=method, masses=np.ones(nmodes), ntraj=T, dt=0.1, rng=A.RNG_INJECTED,
                      save_every=3, nsave=nsteps // 3 + 1, observables=obs, per_trajectory=1)
    mk = engine_factory()
    cfgs = [A.make_config(**kw) for _ in range(3)]
    fused, plain = mk(*cfgs[0]), mk(*cfgs[1])
    o = oracle_factory()(*cfgs[2])
    w = model.bath_a
    sr = np.sqrt(1.0 / (2 * w * np.tanh(2.5 * w))); sv = np.sqrt(w / (2 * np.tanh(2.5 * w)))
    r = rng.standard_normal((T, nmodes)) * sr
    v = rng.standard_normal((T, nmodes)) * sv
    if pinned:
        import torch
        r_h = torch.from_numpy(r.copy()).pin_memory().numpy()
        v_h = torch.from_numpy(v.copy()).pin_memory().numpy()
    else:
        r_h, v_h = r, v
    rho = _pure_state(T, 2, 0)
    draws = rng.random((nsteps, T)); sdraw = rng.random(T)
    fused.set_draws(draws)
    fused.run_from_host(r_h, v_h, rho, None, None, sdraw, diabatic=True, nsteps=30)
    fused.run(nsteps - 30)
    for h in (plain, o):
        h.set_state_diabatic(r, v, rho, None, None, sdraw)
        h.set_draws(draws)
        h.run(nsteps)
    _compare_state(fused, o, 1e-9, "fused vs oracle")
    _compare_state(fused, plain, 1e-12, "fused vs two-call")
    _compare_observables(fused, o, obs, 1e-9, T)
    for oid in (A.OBS_DIABATIC_POP, A.OBS_SIGMA):
        assert np.max(np.abs(fused.observable_per_trajectory(oid) - o.observable_per_trajectory(oid))) < 1e-9
    assert fused.counters()["hops"] == o.counters()["hops"]
    assert fused.counters()["frustrated"] == o.counters()["frustrated"]


@pytest.mark.parametrize("method", [A.METHOD_FSSH, A.METHOD_EHRENFEST])
@pytest.mark.parametrize("nbeads", [3, 4, 10, 16, 32])      # 10 beads: test/Dynamics/bcbwithtsit5.jl:10-37
@pytest.mark.parametrize("name,model,mass,r0,v0,temp", [
    ("tully1", nq.TullyModelOne(), 2000.0, -4.0, 10.0 / 2000, 1e-3),
    ("morse3", nq.ThreeStateMorse(), 20000.0, 2.1, 0.0, 9.5e-4),
])
def test_ring_polymer_parity(method, nbeads, name, model, mass, r0, v0, temp):
    """RPSH / RP-Ehrenfest (BASELINE config 5): beads-on-lanes kernel vs oracle BCBwithTsit5."""
    T, nsteps = 40, 50
    rng = np.random.default_rng(17)
    kw = model_config(model, method=method, masses=[mass], ntraj=T, dt=1.0, nbeads=nbeads, temperature=temp,
                      rng=A.RNG_INJECTED, diagnostics=1, save_every=5, nsave=nsteps // 5 + 1, observables=ALL_POP_OBS)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    r = r0 + 0.05 * rng.standard_normal((T, nbeads))
    v = v0 + np.sqrt(temp * nbeads / mass) * rng.standard_normal((T, nbeads))
    rho = _pure_state(T, model.nstates, 0)
    draws = rng.random((nsteps, T)); sdraw = rng.random(T)
    for h in (e, o):
        h.set_state_diabatic(r, v, rho, None, None, sdraw)
        h.set_draws(draws)
    for chunk in range(nsteps // 5):
        e.run(5); o.run(5)
        _compare_state(e, o, 1e-9, f"chunk {chunk}")
        de, do = e.diagnostics(), o.diagnostics()
        assert rel_err(de["eig"], do["eig"]) < 1e-9
        assert rel_err(de["nac"], do["nac"]) < 1e-9
        assert rel_err(de["accel"], do["accel"]) < 1e-9
    _compare_observables(e, o, ALL_POP_OBS, 1e-9, T)


@pytest.mark.parametrize("nbeads", [1, 2, 5, 8, 10, 32])
def test_rpmd_parity(nbeads):
    """BASELINE config 3: RPMD on Harmonic, normal-mode Cayley propagation."""
    T, nsteps = 64, 200
    rng = np.random.default_rng(19)
    model = nq.Harmonic(m=1837.0, ω=0.005, r0=0.1)
    temp = 9.5e-4
    kw = model_config(model, method=A.METHOD_CLASSICAL, masses=[1837.0], ntraj=T, dt=2.5, nbeads=nbeads, temperature=temp,
                      save_every=10, nsave=nsteps // 10 + 1, observables=CLASSICAL_OBS, per_trajectory=1)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    r = 0.1 + 0.2 * rng.standard_normal((T, nbeads))
    v = np.sqrt(temp * nbeads / 1837.0) * rng.standard_normal((T, nbeads))
    for h in (e, o):
        h.set_state(r, v)
        h.run(nsteps)
    _compare_state(e, o, 1e-10, "final")
    _compare_observables(e, o, CLASSICAL_OBS, 1e-10, T)
    E = e.observable_per_trajectory(A.OBS_TOTAL_ENERGY)[:, :, 0]
    assert np.max(np.abs(E - E[:, :1])) < 1e-3 * np.max(np.abs(E))   # symplectic: bounded energy error


@pytest.mark.parametrize("nbeads", [1, 4, 16])
@pytest.mark.parametrize("name,model,mass,r0,temp", [
    ("doublewell", nq.DoubleWell(), 1.0, 0.2, 0.7),
    ("morse3", nq.ThreeStateMorse(), 20000.0, 2.6, 9.5e-4),
])
def test_nrpmd_parity(nbeads, name, model, mass, r0, temp):
    """BASELINE config 5 (NRPMD): RingPolymerMInt with mapping variables in the adiabatic basis vs the oracle's
    dense C/D/E/F matrices (ringpolymer_mint.jl:28-130)."""
    T, nsteps = 48, 120
    n = model.nstates
    rng = np.random.default_rng(23)
    obs = ((1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_POPCORR_DIABATIC) | (1 << A.OBS_KINETIC) | (1 << A.OBS_POTENTIAL) |
           (1 << A.OBS_TOTAL_ENERGY) | (1 << A.OBS_POSITION) | (1 << A.OBS_VELOCITY) | (1 << A.OBS_MAPPING_Q) | (1 << A.OBS_MAPPING_P))
    dt = 0.01 if mass < 100 else 1.0
    kw = model_config(model, method=A.METHOD_NRPMD, masses=[mass], ntraj=T, dt=dt, nbeads=nbeads, temperature=temp,
                      save_every=10, nsave=nsteps // 10 + 1, observables=obs, per_trajectory=1, nrpmd_gamma=0.5)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    r = r0 + 0.1 * rng.standard_normal((T, nbeads, 1))
    v = np.sqrt(temp * nbeads / mass) * rng.standard_normal((T, nbeads, 1))
    th = rng.random((T, nbeads, n)) * 2 * np.pi
    R = np.full(n, np.sqrt(2 * 0.5)); R[0] = np.sqrt(2 + 2 * 0.5)
    q, p = np.cos(th) * R, np.sin(th) * R
    for h in (e, o):
        h.set_state(r, v)
        h.set_mapping(q, p)
    for chunk in range(nsteps // 10):
        e.run(10); o.run(10)
        _compare_state(e, o, 1e-10, f"chunk {chunk}")
        qe, pe = e.get_mapping(); qo, po = o.get_mapping()
        assert np.max(np.abs(qe - qo)) < 1e-10 and np.max(np.abs(pe - po)) < 1e-10
    _compare_observables(e, o, obs, 1e-9, T)
    assert np.max(np.abs(e.observable_per_trajectory(A.OBS_TOTAL_ENERGY) - o.observable_per_trajectory(A.OBS_TOTAL_ENERGY))) < 1e-9
    # OutputMappingPosition / OutputMappingMomentum streams: frame 0 is what set_mapping uploaded, the last frame is get_mapping
    for oid, first, last in ((A.OBS_MAPPING_Q, q, e.get_mapping()[0]), (A.OBS_MAPPING_P, p, e.get_mapping()[1])):
        se_, so_ = e.observable_per_trajectory(oid), o.observable_per_trajectory(oid)
        assert np.max(np.abs(se_ - so_)) < 1e-9
        assert np.array_equal(se_[:, 0].reshape(T, nbeads, n), first)
        assert np.array_equal(se_[:, -1].reshape(T, nbeads, n), np.asarray(last).reshape(T, nbeads, n))


def test_unsupported_configuration_fails_loudly():
    """No CPU fallback: a configuration without a kernel is an error, not a silent slow path."""
    kw = model_config(nq.TullyModelOne(), method=A.METHOD_IESH, masses=[2000.0], ntraj=4, dt=1.0, nelectrons=1)
    cfg, keep = A.make_config(**kw)
    with pytest.raises(nq.EngineError) as ei:
        engine_factory()(cfg, keep)
    assert ei.value.code == -2


# ---- AdiabaticIESH on the Newns-Anderson model (BASELINE config 4) ---------------------------------------
IESH_OBS = ((1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_KINETIC) | (1 << A.OBS_POTENTIAL) |
            (1 << A.OBS_TOTAL_ENERGY) | (1 << A.OBS_POSITION) | (1 << A.OBS_VELOCITY) | (1 << A.OBS_DISCRETE_STATE) |
            (1 << A.OBS_SIGMA))


def _iesh_model(M, width=0.0192):
    return nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(M, -width, width))


def _iesh_ground_state(T, n, ne):
    psi = np.zeros((T, ne, n))
    psi[:, np.arange(ne), np.arange(ne)] = 1.0            # DynamicsVariables(sim, v, r): iesh.jl:89-97
    state = np.tile(np.arange(1, ne + 1, dtype=np.int32), (T, 1))
    return psi, state


def _iesh_random_state(rng, T, n, ne):
    """Orthonormal complex orbitals near (but not at) the adiabatic ground state, so that det S is not tiny."""
    re = np.empty((T, ne, n)); im = np.empty((T, ne, n))
    for t in range(T):
        q, _ = np.linalg.qr(rng.standard_normal((n, ne)) + 1j * rng.standard_normal((n, ne)))
        q, _ = np.linalg.qr(np.eye(n, ne) + 0.3 * q)
        re[t], im[t] = q.T.real, q.T.imag
    state = np.tile(np.arange(1, ne + 1, dtype=np.int32), (T, 1))
    return re, im, state


def _iesh_pair(M, T, dt, nsave, save_every=1, **extra):
    model = _iesh_model(M)
    kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], ntraj=T, dt=dt, rng=A.RNG_INJECTED, diagnostics=1,
                      save_every=save_every, nsave=nsave, observables=IESH_OBS, per_trajectory=1)
    kw.update(extra)
    return model, make_pair(engine_factory(), oracle_factory(), **kw)


def _iesh_compare(e, o, tol, what):
    se, so = e.get_state(), o.get_state()
    for key in ("r", "v"):
        assert rel_err(se[key], so[key]) < tol, f"{what} {key}"
    assert np.max(np.abs(se["sigma"] - so["sigma"])) < tol, f"{what} psi"
    assert np.array_equal(se["state"], so["state"]), f"{what} occupations"
    de, do = e.diagnostics(), o.diagnostics()
    assert rel_err(de["eig"], do["eig"]) < tol, f"{what} eigenvalues"
    assert rel_err(de["accel"], do["accel"]) < tol, f"{what} acceleration"
    assert np.max(np.abs(de["Z"] - do["Z"])) < tol, f"{what} eigenvectors"
    assert np.max(np.abs(de["nac"] - do["nac"])) < tol * max(1.0, np.max(np.abs(do["nac"]))), f"{what} NAC"


@pytest.mark.parametrize("start", ["ground", "random"])
def test_iesh_per_step_parity(start):
    """n = 31, ne = 15 (test/Dynamics/iesh.jl:19,30): r, v, psi, occupations, w, Z, NAC, force, every estimator."""
    M, T, nsteps = 30, 6, 12
    rng = np.random.default_rng(21)
    model, (e, o) = _iesh_pair(M, T, 1.0, nsteps + 1)
    n, ne = model.nstates, model.nelectrons
    r = 21.0 * rng.random(T)            # both wells and the crossing region
    v = rng.standard_normal(T) * np.sqrt(9.5e-4 / 2000.0) * 5
    if start == "ground":
        re, state = _iesh_ground_state(T, n, ne); im = None
    else:
        re, im, state = _iesh_random_state(rng, T, n, ne)
    xi = rng.random((nsteps, T))
    for h in (e, o):
        h.set_state(r, v, re, im, state)
        h.set_draws(xi)
    _iesh_compare(e, o, STEP_TOL, "t0")
    for chunk in range(nsteps // 4):
        e.run(4); o.run(4)
        _iesh_compare(e, o, STEP_TOL, f"chunk {chunk}")
    _compare_observables(e, o, IESH_OBS, 1e-9, T)
    for oid in (A.OBS_DIABATIC_POP, A.OBS_TOTAL_ENERGY, A.OBS_SIGMA):
        assert np.max(np.abs(e.observable_per_trajectory(oid) - o.observable_per_trajectory(oid))) < 1e-9
    psi = e.get_state()["sigma"]        # sanity: the orbitals stay normalised
    assert np.allclose(np.einsum("tie,tie->te", psi.conj(), psi).real, 1.0, atol=1e-12)


@pytest.mark.parametrize("rescaling", [A.RESCALE_STANDARD, A.RESCALE_VINVERSION])
def test_iesh_identical_hop_sequences(rescaling):
    """Same injected draws -> same hops, frustrated hops and occupations (small draws force the unpruned branch)."""
    M, T, nsteps = 30, 8, 30
    rng = np.random.default_rng(22)
    model, (e, o) = _iesh_pair(M, T, 5.0, nsteps + 1, rescaling=rescaling)
    n, ne = model.nstates, model.nelectrons
    r = 5.0 + 12.0 * rng.random(T)
    v = -np.abs(rng.standard_normal(T)) * 6e-3
    v[::2] *= 0.05                      # slow trajectories: frustrated hops
    re, im, state = _iesh_random_state(rng, T, n, ne)
    xi = rng.random((nsteps, T)) * 5e-4
    for h in (e, o):
        h.set_state(r, v, re, im, state)
        h.set_draws(xi)
    e.run(nsteps); o.run(nsteps)
    ce, co = e.counters(), o.counters()
    assert ce["hops"] == co["hops"] and ce["frustrated"] == co["frustrated"], (ce, co)
    assert ce["hops"] > 0 and ce["frustrated"] > 0, ce
    assert e.hop_search_count() == o.hop_search_count() > 0
    de, do = e.observable_per_trajectory(A.OBS_DISCRETE_STATE), o.observable_per_trajectory(A.OBS_DISCRETE_STATE)
    assert np.array_equal(de, do)
    _iesh_compare(e, o, 1e-9, "after hops")


@pytest.mark.parametrize("M,T,nsteps,dt", [(100, 2, 3, 1.0), (200, 1, 2, 1.0), (100, 1, 2, 10.0)])
def test_iesh_large_bath_parity(M, T, nsteps, dt):
    """BASELINE config 4 sizes: n = 101 (G resident in shared memory) and n = 201 (G streamed in slabs)."""
    rng = np.random.default_rng(23)
    model, (e, o) = _iesh_pair(M, T, dt, nsteps + 1)
    n, ne = model.nstates, model.nelectrons
    r = 8.0 + 10.0 * rng.random(T)
    v = -np.abs(rng.standard_normal(T)) * 2e-3
    re, state = _iesh_ground_state(T, n, ne)
    xi = rng.random((nsteps, T)) * 0.05
    for h in (e, o):
        h.set_state(r, v, re, None, state)
        h.set_draws(xi)
    e.run(nsteps); o.run(nsteps)
    _iesh_compare(e, o, STEP_TOL, f"n={n}")
    _compare_observables(e, o, IESH_OBS, 1e-9, T)


def test_iesh_edc_and_philox_sharding():
    """EDC decoherence matches the oracle; Philox draws keyed by the global trajectory id are shard independent."""
    M, T, nsteps = 30, 6, 10
    rng = np.random.default_rng(24)
    model, (e, o) = _iesh_pair(M, T, 10.0, nsteps + 1, edc_C=0.1)
    n, ne = model.nstates, model.nelectrons
    r = 5.0 + 10.0 * rng.random(T); v = -np.abs(rng.standard_normal(T)) * 2e-3
    re, im, state = _iesh_random_state(rng, T, n, ne)
    xi = rng.random((nsteps, T))
    for h in (e, o):
        h.set_state(r, v, re, im, state); h.set_draws(xi)
    e.run(nsteps); o.run(nsteps)
    _iesh_compare(e, o, 1e-9, "edc")
    # sharding: the same six trajectories as one handle or as two handles with traj_offset
    kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], dt=5.0, rng=A.RNG_PHILOX, seed=99, save_every=1,
                      nsave=nsteps + 1, observables=(1 << A.OBS_DISCRETE_STATE), per_trajectory=1)
    outs = []
    for parts in ([(0, T)], [(0, 2), (2, T)]):
        acc = []
        for lo, hi in parts:
            cfg, keep = A.make_config(ntraj=hi - lo, traj_offset=lo, **kw)
            h = engine_factory()(cfg, keep)
            h.set_state(r[lo:hi], 3 * v[lo:hi], re[lo:hi], im[lo:hi], state[lo:hi])
            h.run(nsteps)
            acc.append((h.get_state(), h.observable_per_trajectory(A.OBS_DISCRETE_STATE)))
            h.close()
        outs.append((np.concatenate([a[0]["sigma"] for a in acc]), np.concatenate([a[1] for a in acc])))
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][0], outs[1][0])


NA_OBS = ((1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_KINETIC) | (1 << A.OBS_POTENTIAL) | (1 << A.OBS_TOTAL_ENERGY) |
          (1 << A.OBS_POSITION) | (1 << A.OBS_VELOCITY) | (1 << A.OBS_SIGMA))


@pytest.mark.parametrize("M,dt,nsteps", [(30, 10.0, 24), (100, 1.0, 8)])
def test_ehrenfest_na_parity(M, dt, nsteps):
    """Simulation{EhrenfestNA} (ehrenfest_na.jl) on the IESH kernel family: mean-field force from psi, no hops."""
    T = 5
    rng = np.random.default_rng(33)
    model, (e, o) = _iesh_pair(M, T, dt, nsteps + 1, method=A.METHOD_EHRENFEST_NA, observables=NA_OBS)
    n, ne = model.nstates, model.nelectrons
    r = 21.0 * rng.random(T)
    v = rng.standard_normal(T) * np.sqrt(9.5e-4 / 2000.0) * 3
    re, im, _ = _iesh_random_state(rng, T, n, ne)
    for h in (e, o):
        h.set_state(r, v, re, im, None)
    for chunk in range(nsteps // 4):
        e.run(4); o.run(4)
        se, so = e.get_state(), o.get_state()
        assert rel_err(se["r"], so["r"]) < STEP_TOL and rel_err(se["v"], so["v"]) < STEP_TOL
        assert np.max(np.abs(se["sigma"] - so["sigma"])) < STEP_TOL
        de, do = e.diagnostics(), o.diagnostics()
        assert rel_err(de["accel"], do["accel"]) < STEP_TOL and rel_err(de["eig"], do["eig"]) < STEP_TOL
    _compare_observables(e, o, NA_OBS, 1e-9, T)
    adi = e.observable_per_trajectory(A.OBS_ADIABATIC_POP)
    assert np.allclose(adi.sum(axis=2), ne, atol=1e-9)
    assert e.counters()["hops"] == 0

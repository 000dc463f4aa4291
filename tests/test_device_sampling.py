"""Device-side initial conditions (nqcb200_sample_state).

CPU: the oracle's restatement of the Philox / Box-Muller stream has the requested moments, honours fixed components,
the normal-mode transform and the sharding offset.  GPU: the CUDA sampler reproduces the oracle's stream (libm vs
CUDA log / sincos: 1e-13) and the dynamics that follow agree like any other parity run."""
import numpy as np
import pytest

import nqcdynamics_jl_b200 as nq
from helpers import A, engine_factory, make_pair, model_config, oracle_factory, rel_err


def _tully_cfg(T, **kw):
    base = dict(method=A.METHOD_FSSH, masses=[2000.0], ntraj=T, dt=1.0, seed=77, save_every=10, nsave=21,
                observables=(1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_DISCRETE_STATE), per_trajectory=1)
    base.update(kw)
    return model_config(nq.TullyModelOne(), **base)


def test_oracle_sampler_moments_and_sharding():
    T = 20000
    o = oracle_factory()(*A.make_config(**_tully_cfg(T)))
    rho = np.zeros((2, 2)); rho[1, 1] = 1.0
    o.sample_state([(-8.0, 1.5)], [10.0 / 2000], rho, diabatic=True, state=0)
    st = o.get_state()
    r, v = st["r"].reshape(-1), st["v"].reshape(-1)
    assert np.all(v == 10.0 / 2000)
    assert abs(r.mean() + 8.0) < 5 * 1.5 / np.sqrt(T) and abs(r.std() - 1.5) < 0.03
    assert abs(np.mean((r + 8.0) ** 3)) < 0.2 and abs(np.mean(((r + 8.0) / 1.5) ** 4) - 3.0) < 0.15
    assert set(np.unique(st["state"])) <= {1, 2}
    # the stream is keyed by the global trajectory index
    lo = 12345
    a = oracle_factory()(*A.make_config(**_tully_cfg(T - lo, traj_offset=lo)))
    a.sample_state([(-8.0, 1.5)], [10.0 / 2000], rho, diabatic=True, state=0)
    assert np.array_equal(a.get_state()["r"].reshape(-1), r[lo:])
    assert np.array_equal(a.get_state()["state"], st["state"][lo:])


def test_oracle_sampler_normal_modes():
    """Ring polymer: entries given per normal mode, beads = U y (RingPolymerArrays transform)."""
    B, T = 8, 4000
    model = nq.Harmonic(m=1837.0, ω=0.005, r0=0.0)
    kw = model_config(model, method=A.METHOD_CLASSICAL, masses=[1837.0], ntraj=T, dt=2.5, nbeads=B, temperature=9.5e-4,
                      seed=5, save_every=10, nsave=3, observables=(1 << A.OBS_POSITION))
    o = oracle_factory()(*A.make_config(**kw))
    sd = [0.5 / (k + 1) for k in range(B)]
    o.sample_state([(0.0, s) for s in sd], [0.0] * B, None, state=0, normal_modes=True)
    x = o.get_state()["r"].reshape(T, B)
    import oracle
    U = oracle.normal_mode_matrix(B)           # U[j, k]
    y = x @ U                                  # back to normal modes
    assert np.max(np.abs(y.std(axis=0) - np.array(sd))) < 0.03
    assert np.max(np.abs(np.corrcoef(y.T) - np.eye(B))) < 0.08


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tully_fssh", "spinboson_ehrenfest", "rpmd_nm", "rpsh10"])
def test_engine_sampler_matches_oracle(case):
    T = 500
    rng = np.random.default_rng(3)
    rho = None
    kwargs = dict(diabatic=True, state=0, normal_modes=False)
    if case == "tully_fssh":
        kw = _tully_cfg(T, rng=A.RNG_INJECTED, traj_offset=999)
        rspec, vspec = [(-8.0, 1.0)], [10.0 / 2000]
        rho = np.zeros((2, 2)); rho[1, 1] = 1.0
        nsteps, draws = 200, rng.random((200, T))
    elif case == "spinboson_ehrenfest":
        N = 16
        model = nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), N, 0.0, 1.0)
        w = model.bath_a
        kw = model_config(model, method=A.METHOD_EHRENFEST, masses=np.ones(N), ntraj=T, dt=0.1, seed=9, save_every=10,
                          nsave=7, observables=(1 << A.OBS_POPCORR_DIABATIC), per_trajectory=1)
        rspec = [(0.0, float(np.sqrt(1 / (2 * x * np.tanh(2.5 * x))))) for x in w]
        vspec = [(0.0, float(np.sqrt(x / (2 * np.tanh(2.5 * x))))) for x in w]
        rho = np.zeros((2, 2)); rho[0, 0] = 1.0
        nsteps, draws = 60, None
    elif case == "rpmd_nm":
        B = 16
        model = nq.Harmonic(m=1837.0, ω=0.005, r0=0.0)
        kw = model_config(model, method=A.METHOD_CLASSICAL, masses=[1837.0], ntraj=T, dt=2.5, nbeads=B, temperature=9.5e-4,
                          seed=5, save_every=10, nsave=11, observables=(1 << A.OBS_TOTAL_ENERGY), per_trajectory=1)
        rspec = [(0.1, 0.3 / (k + 1)) for k in range(B)]
        vspec = [(0.0, 1e-3)] * B
        kwargs.update(normal_modes=True)
        nsteps, draws = 100, None
    else:
        B = 10
        kw = model_config(nq.ThreeStateMorse(), method=A.METHOD_FSSH, masses=[20000.0], ntraj=T, dt=1.0, nbeads=B,
                          temperature=9.5e-4, seed=21, rng=A.RNG_INJECTED, save_every=10, nsave=11,
                          observables=(1 << A.OBS_POPCORR_DIABATIC), per_trajectory=1)
        rspec, vspec = [(2.1, 0.1)] * B, [(0.0, float(np.sqrt(9.5e-4 * B / 20000.0)))] * B
        rho = np.zeros((3, 3)); rho[0, 0] = 1.0
        nsteps, draws = 100, rng.random((100, T))
    e = engine_factory()(*A.make_config(**kw))
    o = oracle_factory()(*A.make_config(**kw))
    for h in (e, o):
        h.sample_state(rspec, vspec, rho, **kwargs)
        if draws is not None:
            h.set_draws(draws)
    se, so = e.get_state(), o.get_state()
    assert rel_err(se["r"], so["r"]) < 1e-13 and rel_err(se["v"], so["v"]) < 1e-13
    if "state" in so:
        assert np.array_equal(se["state"], so["state"])
    for h in (e, o):
        h.run(nsteps)
    se, so = e.get_state(), o.get_state()
    assert rel_err(se["r"], so["r"]) < 1e-9 and rel_err(se["v"], so["v"]) < 1e-9
    if "sigma" in so:
        assert np.max(np.abs(se["sigma"] - so["sigma"])) < 1e-9
    if "state" in so:
        assert np.array_equal(se["state"], so["state"])


@pytest.mark.gpu
def test_run_dynamics_device_sampling():
    """Host API: EnsembleB200(device_sampling=True) agrees statistically with numpy-sampled initial conditions."""
    sim = nq.Simulation[nq.FSSH](nq.Atoms(2000), nq.TullyModelOne())
    dist = nq.DynamicalDistribution(10.0 / 2000, nq.Normal(-5.0, 0.5), sim.size) * nq.PureState(2)
    T = 20000
    outs = (nq.OutputDiabaticPopulation, nq.OutputPosition)
    a = nq.run_dynamics(sim, (0.0, 2000.0), dist, output=outs, trajectories=T, saveat=50.0, seed=11, reduction=nq.MeanReduction())
    b = nq.run_dynamics(sim, (0.0, 2000.0), dist, output=outs, trajectories=T, saveat=50.0, seed=12, reduction=nq.MeanReduction(),
                        ensemble_algorithm=nq.EnsembleB200(1, device_sampling=True))
    se = np.sqrt(0.25 / T)
    assert np.max(np.abs(a["OutputDiabaticPopulation"] - b["OutputDiabaticPopulation"])) < 6 * se
    assert abs(a["OutputPosition"][0].item() - b["OutputPosition"][0].item()) < 6 * 0.5 / np.sqrt(T)


def _fd_model():
    return nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(30, -0.0192, 0.0192))


def test_oracle_fermi_dirac_sampler_statistics():
    """CPU: the oracle's restatement of sample_fermi_dirac_distribution (DynamicsUtils.jl:194-208) with the Philox stream of
    nqcb200_sample_occupations: sorted distinct occupations, ne electrons, the mean occupation follows the Fermi function of
    the adiabatic energies (canonical vs grand-canonical: within 0.08 here), beta = inf gives the ground state."""
    import oracle
    model = _fd_model()
    n, ne, T = model.nstates, model.nelectrons, 1500
    kT = 3e-3
    kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], ntraj=T, dt=1.0, seed=77, traj_offset=5,
                      observables=1 << A.OBS_DISCRETE_STATE, per_trajectory=1, nsave=1)
    cfg, keep = A.make_config(**kw)
    o = oracle.OracleEngine(cfg, keep)
    r = np.full(T, 21.0); v = np.zeros(T)
    o.set_state(r, v, None, None, None)
    assert np.array_equal(o.get_state()["state"][0], np.arange(1, ne + 1))
    o.sample_occupations(1.0 / kT)
    occ = o.get_state()["state"]
    assert occ.shape == (T, ne) and np.all(np.diff(occ, axis=1) > 0) and occ.min() >= 1 and occ.max() <= n
    psi = o.get_state()["sigma"]                                   # psi[state[e], e] = 1
    assert np.array_equal(np.argmax(np.abs(psi), axis=1) + 1, occ) and np.allclose(np.abs(psi).sum(axis=1), 1.0)
    E = model.adiabatic_energies([21.0])
    mean_occ = np.bincount((occ - 1).ravel(), minlength=n) / T
    mu = 0.5 * (E[ne - 1] + E[ne])
    fermi = 1.0 / (1.0 + np.exp((E - mu) / kT))
    assert np.max(np.abs(mean_occ - fermi)) < 0.08
    o.sample_occupations(float("inf"))
    assert np.array_equal(o.get_state()["state"], np.tile(np.arange(1, ne + 1), (T, 1)))


@pytest.mark.gpu
def test_device_fermi_dirac_occupations_match_oracle():
    """nqcb200_sample_occupations: the device's Metropolis walk gives the oracle's occupations trajectory by trajectory
    (same Philox stream, keyed by the global trajectory index -> also shard independent), psi is rebuilt from them and save
    point 0 / the initial force are re-recorded."""
    model = _fd_model()
    n, ne, T = model.nstates, model.nelectrons, 96
    obs = (1 << A.OBS_DISCRETE_STATE) | (1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_TOTAL_ENERGY)
    kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], ntraj=T, dt=1.0, seed=123, traj_offset=40, rng=A.RNG_INJECTED,
                      observables=obs, per_trajectory=1, nsave=3, save_every=2, diagnostics=1)
    rng = np.random.default_rng(8)
    r = 10.0 + 12.0 * rng.random(T); v = rng.standard_normal(T) * 1e-3
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    for h in (e, o):
        h.set_state(r, v, None, None, None)
        h.sample_occupations(1.0 / 2e-3)
        h.set_draws(np.ones((4, T)))
    se, so = e.get_state(), o.get_state()
    assert np.array_equal(se["state"], so["state"]) and np.any(se["state"] != np.arange(1, ne + 1))
    assert np.array_equal(se["sigma"], so["sigma"])
    for h in (e, o):
        h.run(4)
    for oid in (A.OBS_DISCRETE_STATE, A.OBS_ADIABATIC_POP):
        assert np.array_equal(e.observable_per_trajectory(oid), o.observable_per_trajectory(oid))
    assert np.max(np.abs(e.observable_per_trajectory(A.OBS_TOTAL_ENERGY) - o.observable_per_trajectory(A.OBS_TOTAL_ENERGY))) < 1e-10
    # two shards == one
    halves = []
    for lo, hi in ((0, 40), (40, T)):
        cfg, keep = A.make_config(**{**kw, "ntraj": hi - lo, "traj_offset": 40 + lo})
        h = engine_factory()(cfg, keep)
        h.set_state(r[lo:hi], v[lo:hi], None, None, None)
        h.sample_occupations(1.0 / 2e-3)
        halves.append(h.get_state()["state"])
    assert np.array_equal(np.concatenate(halves), se["state"])


@pytest.mark.gpu
def test_device_nrpmd_mapping_matches_oracle():
    """nqcb200_sample_mapping (nrpmd.jl:47-65): radii sqrt(2 + 2 gamma) / sqrt(2 gamma), angles from the shared Philox stream."""
    model, B, T, g = nq.ThreeStateMorse(), 4, 64, 0.5
    obs = (1 << A.OBS_MAPPING_Q) | (1 << A.OBS_DIABATIC_POP)
    kw = model_config(model, method=A.METHOD_NRPMD, masses=[20000.0], ntraj=T, dt=1.0, nbeads=B, temperature=9.5e-4, seed=9,
                      traj_offset=3, observables=obs, per_trajectory=1, nsave=3, save_every=5, nrpmd_gamma=g)
    rng = np.random.default_rng(2)
    r = 2.6 + 0.1 * rng.standard_normal((T, B, 1)); v = 1e-4 * rng.standard_normal((T, B, 1))
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    for h in (e, o):
        h.set_state(r, v)
        h.sample_mapping(2)
    (qe, pe), (qo, po) = e.get_mapping(), o.get_mapping()
    assert np.max(np.abs(qe - qo)) < 1e-14 and np.max(np.abs(pe - po)) < 1e-14
    rad = np.sqrt(qe ** 2 + pe ** 2)
    assert np.allclose(rad[:, :, 1], np.sqrt(2 + 2 * g)) and np.allclose(rad[:, :, [0, 2]], np.sqrt(2 * g))
    assert np.allclose(e.observable_sum(A.OBS_DIABATIC_POP)[0] / T, [0.0, 1.0, 0.0], atol=1e-12)      # nrpmd.jl:111-122 at t0
    for h in (e, o):
        h.run(10)
    assert np.max(np.abs(e.get_mapping()[0] - o.get_mapping()[0])) < 1e-10


@pytest.mark.gpu
def test_run_dynamics_device_side_electronic_initial_conditions():
    """EnsembleB200(device_sampling=True) for AdiabaticIESH x FermiDiracState (occupations on the device) and NRPMD x
    PureState (nuclei and mapping variables on the device): same shapes and invariants as the host-sampled path."""
    model = _fd_model()
    ne, T = model.nelectrons, 24
    sim = nq.Simulation[nq.AdiabaticIESH](nq.Atoms(2000), model)
    dist = nq.DynamicalDistribution(nq.Normal(0.0, 7e-4), nq.Normal(21.0, 1.0), (1, 1)) * nq.FermiDiracState(0.0, 2e-3)
    out = nq.run_dynamics(sim, (0.0, 20.0), dist, output=(nq.OutputOccupations, nq.OutputAdiabaticPopulation), trajectories=T,
                          dt=5.0, seed=4, ensemble_algorithm=nq.EnsembleB200(1, device_sampling=True))
    occ0 = np.array([tr["OutputOccupations"][0] for tr in out])
    assert occ0.shape == (T, ne) and np.all(np.diff(occ0, axis=1) > 0) and np.any(occ0 != np.arange(1, ne + 1))
    assert all(np.allclose(tr["OutputAdiabaticPopulation"].sum(axis=1), ne) for tr in out)
    sim = nq.RingPolymerSimulation[nq.NRPMD](nq.Atoms(20000), nq.ThreeStateMorse(), 4, γ=0.5, temperature=9.5e-4)
    dist = nq.DynamicalDistribution(nq.Normal(0.0, 1e-4), nq.Normal(2.6, 0.1), sim.size) * nq.PureState(1, nq.Diabatic())
    res = nq.run_dynamics(sim, (0.0, 20.0), dist, output=(nq.OutputDiabaticPopulation, nq.OutputMappingPosition), trajectories=T,
                          dt=1.0, saveat=10.0, seed=5, ensemble_algorithm=nq.EnsembleB200(1, device_sampling=True))
    assert np.allclose(res[0]["OutputDiabaticPopulation"][0], [1.0, 0.0, 0.0], atol=1e-12)
    assert res[3]["OutputMappingPosition"].shape == (3, 3, 4)

"""Device-side initial conditions (nqcb200_sample_state).

CPU: the oracle's restatement of the Philox / Box-Muller stream has the requested moments, honours fixed components,
the normal-mode transform and the sharding offset.  GPU: the CUDA sampler reproduces the oracle's stream (libm vs
CUDA log / sincos: 1e-13) and the dynamics that follow agree like any other parity run."""
import numpy as np
import pytest

import nqcdynamics_jl_b200 as nq
from helpers import A, engine_factory, model_config, oracle_factory, rel_err


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

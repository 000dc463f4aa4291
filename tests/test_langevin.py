"""RingPolymerSimulation{ThermalLangevin} + BCOCB (thermal ring-polymer sampling on the GPU).

CPU: the oracle's restatement thermalises a harmonic ring polymer to the exact discretised path-integral distribution
(test/Dynamics/langevin.jl-style check: the sampled variances match the analytic normal-mode variances).
GPU: the register-resident FFT kernel reproduces the oracle step by step with the same injected normals, reproduces the
oracle's Philox stream, and thermalises to the same distribution."""
import numpy as np
import pytest

import nqcdynamics_jl_b200 as nq
from helpers import A, engine_factory, model_config, oracle_factory, rel_err

M_H, W, KT, GAMMA = 1837.0, 0.005, 9.5e-4, 2e-3


def _cfg(T, B, nsteps, save_every, **kw):
    model = nq.Harmonic(m=M_H, ω=W, r0=0.0)
    base = dict(method=A.METHOD_THERMAL_LANGEVIN, masses=[M_H], ntraj=T, dt=5.0, nbeads=B, temperature=KT, nrpmd_gamma=GAMMA,
                save_every=save_every, nsave=nsteps // save_every + 1, seed=31,
                observables=(1 << A.OBS_KINETIC) | (1 << A.OBS_POTENTIAL) | (1 << A.OBS_POSITION), per_trajectory=0)
    base.update(kw)
    return model_config(model, **base)


def _exact_variances(B):
    """<r_bead^2> and <KE> of the B-bead harmonic ring polymer at ring-polymer temperature B kT."""
    beta_B = 1.0 / (KT * B)
    wk = 2.0 * B * KT * np.sin(np.arange(B) * np.pi / B)
    var_modes = 1.0 / (beta_B * M_H * (wk ** 2 + W ** 2))
    return var_modes.sum() / B, 0.5 * B * B * KT       # per-bead <r^2>; total <KE> = B * (B kT) / 2


def test_oracle_bcocb_thermalises_harmonic_ring_polymer():
    B, T, nsteps = 4, 256, 6000
    o = oracle_factory()(*A.make_config(**_cfg(T, B, nsteps, 50)))
    o.set_state(np.zeros((T, B, 1)), np.zeros((T, B, 1)))
    o.run(nsteps)
    st = o.get_state()
    r2, ke = _exact_variances(B)
    n = T * B
    assert abs(np.mean(st["r"] ** 2) / r2 - 1.0) < 6.0 * np.sqrt(2.0 / n) + 0.02
    kin = o.observable_sum(A.OBS_KINETIC)[-40:, 0] / T         # time-averaged over the equilibrated tail
    assert abs(kin.mean() / ke - 1.0) < 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("B", [2, 5, 8, 10, 32])      # 5, 10: the dense any-bead-count kernel
def test_engine_bcocb_matches_oracle_with_injected_noise(B):
    T, nsteps = 70, 40
    rng = np.random.default_rng(5)
    kw = _cfg(T, B, nsteps, 5, rng=A.RNG_INJECTED, per_trajectory=1)
    e = engine_factory()(*A.make_config(**kw)); o = oracle_factory()(*A.make_config(**kw))
    r = 0.1 * rng.standard_normal((T, B, 1)); v = 1e-3 * rng.standard_normal((T, B, 1))
    xi = rng.standard_normal((nsteps, T, B))
    for h in (e, o):
        h.set_state(r, v)
        h.set_noise(xi)
    for chunk in range(4):
        e.run(10); o.run(10)
        se, so = e.get_state(), o.get_state()
        assert rel_err(se["r"], so["r"]) < 1e-10 and rel_err(se["v"], so["v"]) < 1e-10, chunk
    for oid in (A.OBS_KINETIC, A.OBS_POTENTIAL, A.OBS_POSITION):
        a, b = e.observable_sum(oid), o.observable_sum(oid)
        assert np.max(np.abs(a - b)) <= 1e-9 * max(1.0, np.max(np.abs(b)))


@pytest.mark.gpu
def test_engine_bcocb_philox_stream_and_thermal_distribution():
    B, T, nsteps = 16, 4096, 4000
    kw = _cfg(T, B, nsteps, 50)
    e = engine_factory()(*A.make_config(**kw))
    e.set_state(np.zeros((T, B, 1)), np.zeros((T, B, 1)))
    e.run(nsteps)
    r2, ke = _exact_variances(B)
    st = e.get_state()
    assert abs(np.mean(st["r"] ** 2) / r2 - 1.0) < 6.0 * np.sqrt(2.0 / (T * B)) + 0.02
    kin = e.observable_sum(A.OBS_KINETIC)[-40:, 0] / T
    assert abs(kin.mean() / ke - 1.0) < 0.02
    # the production stream is the oracle's: a short run agrees trajectory by trajectory (FFT kernel and dense kernel)
    for B2 in (16, 5):
        kw2 = _cfg(64, B2, 20, 5)
        e2 = engine_factory()(*A.make_config(**kw2)); o2 = oracle_factory()(*A.make_config(**kw2))
        for h in (e2, o2):
            h.set_state(np.full((64, B2, 1), 0.05), np.zeros((64, B2, 1)))
            h.run(20)
        assert rel_err(e2.get_state()["r"], o2.get_state()["r"]) < 1e-10
        assert rel_err(e2.get_state()["v"], o2.get_state()["v"]) < 1e-10


@pytest.mark.gpu
def test_thermal_langevin_through_run_dynamics():
    """Host API: RingPolymerSimulation[ThermalLangevin] thermalises to the ring-polymer temperature (equipartition)."""
    B = 8
    sim = nq.RingPolymerSimulation[nq.ThermalLangevin](nq.Atoms(M_H), nq.Harmonic(m=M_H, ω=W, r0=0.0), B, temperature=KT, γ=GAMMA)
    dist = nq.DynamicalDistribution(0.0, 0.0, sim.size)
    out = nq.run_dynamics(sim, (0.0, 20000.0), dist, output=(nq.OutputKineticEnergy, nq.OutputCentroidPosition),
                          trajectories=2048, dt=5.0, saveat=250.0, seed=3, reduction=nq.MeanReduction())
    ke = out["OutputKineticEnergy"][-30:].mean()
    assert abs(ke / (0.5 * B * B * KT) - 1.0) < 0.03

"""RingPolymerSimulation{AdiabaticIESH} / {EhrenfestNA} with BCBWavefunction (rpiesh.jl:21-52, rpehrenfest_na.jl:13-52,
bcb_wavefunction.jl:37-69): the ring-polymer instantiation of the AdiabaticIESH kernel against the
oracle's dense restatement -- bead forces from one arrowhead eigenproblem per bead, psi propagated with the centroid
generator of the previous geometry and velocity (quirk Q5), centroid hop test, the rescaling applied to every bead."""
import numpy as np
import pytest

import nqcdynamics_jl_b200 as nq
from helpers import engine_factory, make_pair, model_config, oracle_factory, rel_err
from test_parity_gpu import (IESH_OBS, NA_OBS, STEP_TOL, _compare_observables, _iesh_ground_state, _iesh_model,
                             _iesh_random_state)

A = nq._abi
pytestmark = pytest.mark.gpu
KT = 9.5e-4


def _pair(M, T, B, dt, nsave, save_every=1, **extra):
    model = _iesh_model(M)
    kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], ntraj=T, dt=dt, nbeads=B, temperature=KT, rng=A.RNG_INJECTED,
                      diagnostics=1, save_every=save_every, nsave=nsave, observables=IESH_OBS, per_trajectory=1)
    kw.update(extra)
    return model, make_pair(engine_factory(), oracle_factory(), **kw)


def _compare(e, o, tol, what, occupations=True):
    se, so = e.get_state(), o.get_state()
    for key in ("r", "v"):
        assert rel_err(se[key], so[key]) < tol, f"{what} {key}"
    assert np.max(np.abs(se["sigma"] - so["sigma"])) < tol, f"{what} psi"
    if occupations:
        assert np.array_equal(se["state"], so["state"]), f"{what} occupations"
    de, do = e.diagnostics(), o.diagnostics()
    assert rel_err(de["eig"], do["eig"]) < tol, f"{what} centroid eigenvalues"
    assert rel_err(de["accel"], do["accel"]) < tol, f"{what} bead accelerations"
    assert np.max(np.abs(de["Z"] - do["Z"])) < tol, f"{what} centroid eigenvectors"
    assert np.max(np.abs(de["nac"] - do["nac"])) < tol * max(1.0, np.max(np.abs(do["nac"]))), f"{what} centroid NAC"


def _beads(rng, T, B, r_lo, r_hi, vscale):
    rc = r_lo + (r_hi - r_lo) * rng.random((T, 1))
    r = rc + 0.3 * rng.standard_normal((T, B))
    v = rng.standard_normal((T, B)) * vscale
    return r, v


@pytest.mark.parametrize("start", ["ground", "random"])
@pytest.mark.parametrize("B", [3, 4, 8])
def test_rpiesh_per_step_parity(start, B):
    """n = 31, ne = 15, the reference's RPIESH test system (test/Dynamics/rpiesh.jl:11-22 uses 4 beads)."""
    M, T, nsteps = 30, 5, 12
    rng = np.random.default_rng(71)
    model, (e, o) = _pair(M, T, B, 1.0, nsteps + 1)
    n, ne = model.nstates, model.nelectrons
    r, v = _beads(rng, T, B, 2.0, 19.0, np.sqrt(KT * B / 2000.0) * 3)
    if start == "ground":
        re, state = _iesh_ground_state(T, n, ne); im = None
    else:
        re, im, state = _iesh_random_state(rng, T, n, ne)
    xi = rng.random((nsteps, T))
    for h in (e, o):
        h.set_state(r, v, re, im, state)
        h.set_draws(xi)
    _compare(e, o, STEP_TOL, "t0")
    for chunk in range(nsteps // 4):
        e.run(4); o.run(4)
        _compare(e, o, STEP_TOL, f"chunk {chunk}")
    _compare_observables(e, o, IESH_OBS, 1e-9, T)
    psi = e.get_state()["sigma"]
    assert np.allclose(np.einsum("tie,tie->te", psi.conj(), psi).real, 1.0, atol=1e-12)


@pytest.mark.parametrize("rescaling", [A.RESCALE_STANDARD, A.RESCALE_VINVERSION])
def test_rpiesh_identical_hop_sequences(rescaling):
    """Small injected draws force the unpruned hop search: same hops / frustrated hops, the velocity change on every bead."""
    M, T, B, nsteps = 30, 8, 4, 30
    rng = np.random.default_rng(72)
    model, (e, o) = _pair(M, T, B, 5.0, nsteps + 1, rescaling=rescaling)
    n, ne = model.nstates, model.nelectrons
    r, v = _beads(rng, T, B, 5.0, 17.0, 1e-4)
    v += -np.abs(rng.standard_normal((T, 1))) * 6e-3
    v[::2] *= 0.05
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
    assert np.array_equal(e.observable_per_trajectory(A.OBS_DISCRETE_STATE), o.observable_per_trajectory(A.OBS_DISCRETE_STATE))
    _compare(e, o, 1e-9, "after hops")


@pytest.mark.parametrize("M,B,T,nsteps,dt", [(100, 4, 2, 3, 1.0), (200, 2, 1, 2, 1.0)])
def test_rpiesh_large_bath_parity(M, B, T, nsteps, dt):
    """n = 101 (G resident in shared memory across the bead solves) and n = 201 (G kept in global memory)."""
    rng = np.random.default_rng(73)
    model, (e, o) = _pair(M, T, B, dt, nsteps + 1)
    n, ne = model.nstates, model.nelectrons
    r, v = _beads(rng, T, B, 8.0, 16.0, 2e-4)
    v -= 2e-3
    re, state = _iesh_ground_state(T, n, ne)
    xi = rng.random((nsteps, T)) * 0.05
    for h in (e, o):
        h.set_state(r, v, re, None, state)
        h.set_draws(xi)
    e.run(nsteps); o.run(nsteps)
    _compare(e, o, STEP_TOL, f"n={n}")
    _compare_observables(e, o, IESH_OBS, 1e-9, T)


def test_rp_ehrenfest_na_parity_and_energy():
    """RingPolymerSimulation{EhrenfestNA} on the reference's test system (test/Dynamics/rp_ehrenfest_na.jl:10-41: M = 30,
    4 beads, v = 0, r ~ 21 + N(0,1), dt = 10, var(E) < 1e-6): parity with the oracle and the reference's own assertion."""
    M, B, T, nsteps, dt = 30, 4, 4, 200, 10.0
    rng = np.random.default_rng(74)
    model = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(M, -0.0192, 0.0192), fermi_level=0.001)
    kw = model_config(model, method=A.METHOD_EHRENFEST_NA, masses=[2000.0], ntraj=T, dt=dt, nbeads=B, temperature=KT,
                      diagnostics=1, save_every=1, nsave=nsteps + 1, observables=NA_OBS, per_trajectory=1)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    n, ne = model.nstates, model.nelectrons
    r = 21.0 + rng.standard_normal((T, B))
    v = np.zeros((T, B))
    re, _ = _iesh_ground_state(T, n, ne)
    for h in (e, o):
        h.set_state(r, v, re, None, None)
    for chunk in range(4):
        e.run(10); o.run(10)
        _compare(e, o, 1e-9, f"chunk {chunk}", occupations=False)
    e.run(nsteps - 40); o.run(nsteps - 40)
    _compare(e, o, 1e-8, "final", occupations=False)
    _compare_observables(e, o, NA_OBS, 1e-8, T)
    E = e.observable_per_trajectory(A.OBS_TOTAL_ENERGY)[:, :, 0]
    assert np.all(np.var(E, axis=1) < 1e-6), np.var(E, axis=1)


def test_rpiesh_launch_boundaries():
    """The generator kept across steps is rebuilt at a launch boundary: run(3) x 4 == run(12) to rounding."""
    M, T, B, nsteps = 30, 4, 4, 12
    rng = np.random.default_rng(75)
    model = _iesh_model(M)
    n, ne = model.nstates, model.nelectrons
    r, v = _beads(rng, T, B, 5.0, 17.0, np.sqrt(KT * B / 2000.0) * 3)
    re, im, state = _iesh_random_state(rng, T, n, ne)
    xi = rng.random((nsteps, T)) * 0.01
    outs = []
    for chunks in ([12], [3, 3, 3, 3]):
        kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], ntraj=T, dt=2.0, nbeads=B, temperature=KT, rng=A.RNG_INJECTED,
                          save_every=1, nsave=nsteps + 1, observables=IESH_OBS, per_trajectory=1)
        cfg, keep = A.make_config(**kw)
        h = engine_factory()(cfg, keep)
        h.set_state(r, v, re, im, state); h.set_draws(xi)
        for c in chunks:
            h.run(c)
        outs.append((h.get_state(), h.observable_per_trajectory(A.OBS_TOTAL_ENERGY)))
        h.close()
    a, b = outs
    assert np.array_equal(a[0]["state"], b[0]["state"])
    assert np.max(np.abs(a[0]["sigma"] - b[0]["sigma"])) < 1e-12
    assert rel_err(a[0]["r"], b[0]["r"]) < 1e-12 and rel_err(a[0]["v"], b[0]["v"]) < 1e-12
    assert np.max(np.abs(a[1] - b[1])) < 1e-11


def test_rpiesh_rejects_what_is_not_built():
    """EDC decoherence and termination masks exist for nbeads == 1: errors, not fallbacks."""
    model = _iesh_model(30)
    kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], ntraj=2, dt=1.0, nbeads=4, temperature=KT, save_every=1, nsave=2,
                      observables=(1 << A.OBS_POSITION), edc_C=0.1)
    cfg, keep = A.make_config(**kw)
    with pytest.raises(nq.EngineError) as ei:
        engine_factory()(cfg, keep)
    assert ei.value.code == -2


def test_rpiesh_through_run_dynamics():
    """RingPolymerSimulation{AdiabaticIESH}(atoms, model, n_beads) and {EhrenfestNA} through run_dynamics: ground-state and
    FermiDiracState initial conditions (centroid eigenvalues, test/Dynamics/rpiesh.jl:28-56), centroid outputs."""
    M, B, T = 30, 4, 64
    model = _iesh_model(M)
    n, ne = model.nstates, model.nelectrons
    sim = nq.RingPolymerSimulation[nq.AdiabaticIESH](nq.Atoms(2000.0), model, B, temperature=KT)
    assert sim.size == (1, 1, B)
    dist = nq.DynamicalDistribution(nq.VelocityBoltzmann(KT * B, [2000.0], (1, 1)), nq.Normal(10.0, 0.5), sim.size)
    out = nq.run_dynamics(sim, (0.0, 50.0), dist, trajectories=T, dt=1.0, saveat=10.0, seed=5,
                          output=(nq.OutputAdiabaticPopulation, nq.OutputDiabaticPopulation, nq.OutputPosition,
                                  nq.OutputTotalEnergy), reduction=nq.MeanReduction())
    pop = np.asarray(out["OutputAdiabaticPopulation"])
    assert pop.shape[0] == 6 and abs(pop[0].sum() - ne) < 1e-12 and np.allclose(pop.sum(axis=-1), ne)
    dia = np.asarray(out["OutputDiabaticPopulation"])
    assert np.allclose(dia.sum(axis=-1), ne, atol=1e-9)
    E = np.asarray(out["OutputTotalEnergy"]).reshape(6)
    assert np.max(np.abs(E - E[0])) < 1e-4 * abs(E[0])
    out2 = nq.run_dynamics(sim, (0.0, 20.0), dist * nq.FermiDiracState(0.0, 300 * 3.166811563e-6), trajectories=T, dt=1.0, seed=6,
                           output=(nq.OutputAdiabaticPopulation,), reduction=nq.MeanReduction())
    p2 = np.asarray(out2["OutputAdiabaticPopulation"])
    assert abs(p2[0].sum() - ne) < 1e-9 and p2[0][0] > 0.9 and p2[0][-1] < 0.1      # thermal occupations around the Fermi level
    na = nq.RingPolymerSimulation[nq.EhrenfestNA](nq.Atoms(2000.0), model, B, temperature=KT)
    out3 = nq.run_dynamics(na, (0.0, 100.0), dist, trajectories=8, dt=10.0, seed=7,
                           output=(nq.OutputTotalEnergy, nq.OutputAdiabaticPopulation), reduction=nq.SortByTrajectoryReduction())
    E3 = np.asarray([tr["OutputTotalEnergy"] for tr in out3]).reshape(8, -1)
    assert E3.shape[1] == 11
    assert np.all(np.var(E3, axis=1) < 1e-6)

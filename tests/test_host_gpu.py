"""GPU: the reference-facing host API (run_dynamics) -- return shapes of every reduction (test/Ensembles/ensembles.jl:15-45,
reductions.jl:36-66), outputs, and statistical agreement between independent seeds."""
import numpy as np
import pytest

import nqcdynamics_jl_b200 as nq

pytestmark = pytest.mark.gpu


def _tully():
    sim = nq.Simulation[nq.FSSH](nq.Atoms(2000), nq.TullyModelOne())
    dist = nq.DynamicalDistribution(10.0 / 2000, nq.Normal(-5.0, 0.5), sim.size) * nq.PureState(2)
    return sim, dist


def test_reduction_shapes():
    sim, dist = _tully()
    out = (nq.OutputDiabaticPopulation, nq.OutputAdiabaticPopulation, nq.OutputTotalEnergy, nq.OutputPosition,
           nq.OutputDiscreteState, nq.OutputSurfaceHops, nq.OutputStateResolvedScattering1D(sim, "adiabatic"),
           nq.PopulationCorrelationFunction(sim, nq.Diabatic()), nq.OutputQuantumSubsystem)
    T = 10
    res = nq.run_dynamics(sim, (0.0, 100.0), dist, output=out, trajectories=T, dt=1.0, saveat=10.0, seed=1)
    assert isinstance(res, list) and len(res) == T                       # SortByTrajectoryReduction
    tr = res[0]
    assert list(tr.keys())[0] == "Time" and np.allclose(tr["Time"], np.arange(0, 101, 10))
    assert tr["OutputDiabaticPopulation"].shape == (11, 2) and tr["OutputTotalEnergy"].shape == (11,)
    assert tr["OutputPosition"].shape == (11, 1, 1) and tr["PopulationCorrelationFunction"].shape == (11, 2, 2)
    assert tr["OutputQuantumSubsystem"].shape == (11, 2, 2) and np.iscomplexobj(tr["OutputQuantumSubsystem"])
    assert tr["OutputDiscreteState"].dtype == np.int64 and isinstance(tr["OutputSurfaceHops"], int)
    sc = tr["OutputStateResolvedScattering1D"]
    assert set(sc) == {"reflection", "transmission"} and abs(sc["reflection"].sum() + sc["transmission"].sum() - 1) < 1e-12
    one = nq.run_dynamics(sim, (0.0, 100.0), dist, output=nq.OutputDiabaticPopulation, trajectories=1, saveat=10.0, seed=1)
    assert isinstance(one, dict)                                          # trajectories == 1 -> unwrapped
    out2 = (nq.OutputDiabaticPopulation, nq.OutputTotalEnergy)
    s = nq.run_dynamics(sim, (0.0, 100.0), dist, output=out2, trajectories=T, saveat=10.0, seed=1, reduction=nq.SumReduction())
    m = nq.run_dynamics(sim, (0.0, 100.0), dist, output=out2, trajectories=T, saveat=10.0, seed=1, reduction=nq.MeanReduction())
    assert np.allclose(s["Time"], T * np.arange(0, 101, 10))              # `:Time` is summed too (ensembles.jl:21)
    assert np.allclose(m["Time"], np.arange(0, 101, 10))
    assert np.allclose(s["OutputDiabaticPopulation"], T * m["OutputDiabaticPopulation"])
    assert np.allclose(m["OutputDiabaticPopulation"].sum(axis=1), 1.0)
    o = nq.run_dynamics(sim, (0.0, 100.0), dist, output=out2, trajectories=T, saveat=10.0, seed=1, reduction=nq.SortByOutputReduction())
    assert set(o) == {"Time", "OutputDiabaticPopulation", "OutputTotalEnergy"} and len(o["OutputTotalEnergy"]) == T
    per = np.mean([t["OutputDiabaticPopulation"] for t in res], axis=0)
    assert np.allclose(per, m["OutputDiabaticPopulation"], atol=1e-12)     # same seed -> same trajectories


def test_independent_seeds_agree_statistically():
    """north_star level 2: population / transmission curves from independent seeds agree within statistical error."""
    sim, dist = _tully()
    T = 20000
    outs = (nq.OutputDiabaticPopulation, nq.OutputStateResolvedScattering1D(sim, "adiabatic"))
    a = nq.run_dynamics(sim, (0.0, 2500.0), dist, output=outs, trajectories=T, saveat=50.0, seed=11, reduction=nq.MeanReduction())
    b = nq.run_dynamics(sim, (0.0, 2500.0), dist, output=outs, trajectories=T, saveat=50.0, seed=12, reduction=nq.MeanReduction())
    se = np.sqrt(0.25 / T)
    assert np.max(np.abs(a["OutputDiabaticPopulation"] - b["OutputDiabaticPopulation"])) < 6 * se
    ta, tb = a["OutputStateResolvedScattering1D"]["transmission"], b["OutputStateResolvedScattering1D"]["transmission"]
    assert np.max(np.abs(ta - tb)) < 6 * se and abs(ta.sum() - 1.0) < 1e-3    # k = 10: everything transmits


def test_ehrenfest_and_rpmd_through_host_api():
    sb = nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), 100, 0.0, 1.0)
    sim = nq.Simulation[nq.Ehrenfest](nq.Atoms(np.ones(100)), sb)
    w = sb.bath_a
    pos = type("W", (), {"sample": staticmethod(lambda rng, shape: rng.standard_normal(shape) * np.sqrt(1 / (2 * w * np.tanh(2.5 * w))))})
    vel = type("W", (), {"sample": staticmethod(lambda rng, shape: rng.standard_normal(shape) * np.sqrt(w / (2 * np.tanh(2.5 * w))))})
    dist = nq.DynamicalDistribution(vel, pos, sim.size) * nq.PureState(1)
    res = nq.run_dynamics(sim, (0.0, 5.0), dist, output=nq.PopulationCorrelationFunction(sim, nq.Diabatic()),
                          trajectories=2000, dt=0.1, reduction=nq.MeanReduction(), seed=3)
    pc = res["PopulationCorrelationFunction"]
    assert pc.shape == (51, 2, 2) and abs(pc[0, 0, 0] - 1.0) < 1e-12 and np.allclose(pc[:, 0, :].sum(axis=1), 1.0, atol=1e-9)
    rp = nq.RingPolymerSimulation[nq.Classical](nq.Atoms(1837.47), nq.Harmonic(m=1837.47, ω=0.005), 32, temperature=9.5e-4)
    d2 = nq.DynamicalDistribution(nq.VelocityBoltzmann(32 * 9.5e-4, [1837.47], (1, 1)), nq.Normal(0.0, 0.2), rp.size)
    r2 = nq.run_dynamics(rp, (0.0, 500.0), d2, output=(nq.OutputTotalEnergy, nq.OutputCentroidPosition), trajectories=64,
                         dt=2.5, saveat=25.0, seed=4)
    E = np.array([t["OutputTotalEnergy"] for t in r2])
    assert np.max(np.abs(E - E[:, :1])) < 1e-3 * np.max(np.abs(E))
    assert r2[0]["OutputCentroidPosition"].shape == (21, 1, 1)


def test_run_dynamics_iesh_ground_state_and_fermi_dirac():
    """Simulation{AdiabaticIESH} through run_dynamics: ground-state DynamicsVariables (iesh.jl:89-97) and
    FermiDiracState occupations (iesh.jl:99-128); output shapes as in the reference (psi is n x ne)."""
    model = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(30, -0.0192, 0.0192))
    sim = nq.Simulation[nq.AdiabaticIESH](nq.Atoms(2000), model)
    n, ne, T = model.nstates, model.nelectrons, 12
    dist = nq.DynamicalDistribution(nq.Normal(0.0, 7e-4), nq.Normal(21.0, 1.0), (1, 1))
    out = nq.run_dynamics(sim, (0.0, 50.0), dist, output=(nq.OutputQuantumSubsystem, nq.OutputDiscreteState,
                                                          nq.OutputTotalEnergy, nq.OutputSurfaceHops),
                          trajectories=T, dt=5.0, seed=3)
    assert len(out) == T and out[0]["OutputQuantumSubsystem"].shape == (11, n, ne)
    assert np.array_equal(out[0]["OutputDiscreteState"][0], np.arange(1, ne + 1))
    psi = out[3]["OutputQuantumSubsystem"][-1]
    assert np.allclose(np.sum(np.abs(psi) ** 2, axis=0), 1.0, atol=1e-12)
    E = np.array([tr["OutputTotalEnergy"] for tr in out])
    assert np.max(np.abs(E - E[:, :1])) < 1e-5           # no hops accepted without energy conservation
    hot = nq.run_dynamics(sim, (0.0, 20.0), dist * nq.FermiDiracState(0.0, 9.5e-4), output=nq.OutputDiscreteState,
                          trajectories=T, dt=5.0, seed=4)
    occ0 = np.array([tr["OutputDiscreteState"][0] for tr in hot])
    assert occ0.shape == (T, ne) and np.all(np.diff(occ0, axis=1) > 0) and np.any(occ0 != np.arange(1, ne + 1))
    mean = nq.run_dynamics(sim, (0.0, 20.0), dist, output=(nq.OutputDiabaticPopulation, nq.OutputAdiabaticPopulation),
                           reduction=nq.MeanReduction(), trajectories=T, dt=5.0, seed=5)
    assert mean["OutputAdiabaticPopulation"].shape == (5, n)
    assert np.allclose(mean["OutputAdiabaticPopulation"].sum(axis=1), ne)
    assert np.allclose(mean["OutputDiabaticPopulation"].sum(axis=1), ne, atol=1e-9)

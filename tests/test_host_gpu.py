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
                                                          nq.OutputTotalEnergy, nq.OutputSurfaceHops, nq.OutputOccupations),
                          trajectories=T, dt=5.0, seed=3)
    assert len(out) == T and out[0]["OutputQuantumSubsystem"].shape == (11, n, ne)
    assert np.array_equal(out[0]["OutputOccupations"][0], np.arange(1, ne + 1))
    assert out[0]["OutputDiscreteState"].shape == (11,) and out[0]["OutputDiscreteState"][0] == 1    # first(u.state), DynamicsOutputs.jl:178
    psi = out[3]["OutputQuantumSubsystem"][-1]
    assert np.allclose(np.sum(np.abs(psi) ** 2, axis=0), 1.0, atol=1e-12)
    E = np.array([tr["OutputTotalEnergy"] for tr in out])
    assert np.max(np.abs(E - E[:, :1])) < 1e-5           # no hops accepted without energy conservation
    hot = nq.run_dynamics(sim, (0.0, 20.0), dist * nq.FermiDiracState(0.0, 9.5e-4), output=nq.OutputOccupations,
                          trajectories=T, dt=5.0, seed=4)
    occ0 = np.array([tr["OutputOccupations"][0] for tr in hot])
    assert occ0.shape == (T, ne) and np.all(np.diff(occ0, axis=1) > 0) and np.any(occ0 != np.arange(1, ne + 1))
    mean = nq.run_dynamics(sim, (0.0, 20.0), dist, output=(nq.OutputDiabaticPopulation, nq.OutputAdiabaticPopulation),
                           reduction=nq.MeanReduction(), trajectories=T, dt=5.0, seed=5)
    assert mean["OutputAdiabaticPopulation"].shape == (5, n)
    assert np.allclose(mean["OutputAdiabaticPopulation"].sum(axis=1), ne)
    assert np.allclose(mean["OutputDiabaticPopulation"].sum(axis=1), ne, atol=1e-9)


def test_derived_outputs():
    """Outputs assembled on the host from the device streams (DynamicsOutputs.jl:101-141,192-227,387-397)."""
    sim, dist = _tully()
    out = (nq.OutputKineticEnergy, nq.OutputFinalKineticEnergy, nq.OutputPosition, nq.OutputFirstPosition,
           nq.OutputFinalPosition, nq.OutputVelocity, nq.OutputFinalVelocity, nq.OutputDiabaticPopulation,
           nq.OutputTotalDiabaticPopulation, nq.OutputTotalAdiabaticPopulation, nq.OutputFinalTime,
           nq.OutputDynamicsVariables, nq.OutputInitial, nq.OutputFinal, nq.OutputDiscreteState,
           nq.OutputQuantumSubsystem)
    T = 6
    res = nq.run_dynamics(sim, (0.0, 200.0), dist, output=out, trajectories=T, dt=1.0, saveat=20.0, seed=3)
    for tr in res:
        assert tr["OutputFinalKineticEnergy"] == tr["OutputKineticEnergy"][-1]
        assert np.array_equal(tr["OutputFirstPosition"], tr["OutputPosition"][0])
        assert np.array_equal(tr["OutputFinalPosition"], tr["OutputPosition"][-1])
        assert np.array_equal(tr["OutputFinalVelocity"], tr["OutputVelocity"][-1])
        assert np.allclose(tr["OutputTotalDiabaticPopulation"], 1.0) and np.allclose(tr["OutputTotalAdiabaticPopulation"], 1.0)
        assert tr["OutputFinalTime"] == 200.0
        frames = tr["OutputDynamicsVariables"]
        assert len(frames) == 11 and set(frames[0]) == {"v", "r", "σreal", "σimag", "state"}
        assert np.array_equal(frames[3]["r"], tr["OutputPosition"][3]) and frames[3]["σreal"].shape == (2, 2)
        assert np.array_equal(frames[-1]["state"], np.atleast_1d(tr["OutputDiscreteState"][-1]))
        assert np.array_equal(tr["OutputInitial"]["r"], frames[0]["r"]) and np.array_equal(tr["OutputFinal"]["v"], frames[-1]["v"])
        assert np.allclose(frames[5]["σreal"] + 1j * frames[5]["σimag"], tr["OutputQuantumSubsystem"][5])
    # ring polymer: spring energy and centroid kinetic energy
    kT = 9.5e-4
    rp = nq.RingPolymerSimulation[nq.Classical](nq.Atoms(1837.0), nq.Harmonic(m=1837.0, ω=0.005, r0=0.0), 8, temperature=kT)
    d2 = nq.DynamicalDistribution(nq.Normal(0.0, 1e-3), nq.Normal(0.0, 0.2), rp.size)
    o2 = (nq.OutputTotalEnergy, nq.OutputKineticEnergy, nq.OutputPotentialEnergy, nq.OutputSpringEnergy,
          nq.OutputCentroidKineticEnergy, nq.OutputCentroidVelocity)
    rr = nq.run_dynamics(rp, (0.0, 100.0), d2, output=o2, trajectories=4, dt=2.5, saveat=10.0, seed=5)
    for tr in rr:
        assert np.all(tr["OutputSpringEnergy"] > 0)
        assert np.allclose(tr["OutputSpringEnergy"] + tr["OutputKineticEnergy"] + tr["OutputPotentialEnergy"], tr["OutputTotalEnergy"])
        vc = np.asarray(tr["OutputCentroidVelocity"]).reshape(len(tr["OutputCentroidKineticEnergy"]), -1)
        assert np.allclose(tr["OutputCentroidKineticEnergy"], 0.5 * 1837.0 * vc[:, 0] ** 2)


def test_ehrenfest_na_through_run_dynamics():
    """test/Dynamics/ehrenfest_na.jl:24-52 through the host API: energy conservation, conserved electron count."""
    G = 6.4e-3
    model = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=G), nq.TrapezoidalRule(30, -3 * G, 3 * G))
    sim = nq.Simulation[nq.EhrenfestNA](nq.Atoms(2000), model)
    dist = nq.DynamicalDistribution(0.0, 21.0, sim.size)
    out = (nq.OutputTotalEnergy, nq.OutputKineticEnergy, nq.OutputPotentialEnergy, nq.OutputPosition,
           nq.OutputAdiabaticPopulation, nq.OutputQuantumSubsystem)
    tr = nq.run_dynamics(sim, (0.0, 2000.0), dist, output=out, trajectories=1, dt=10.0, saveat=10.0)
    assert np.var(tr["OutputTotalEnergy"]) < 1e-6
    assert np.allclose(tr["OutputAdiabaticPopulation"].sum(axis=1), model.nelectrons)
    assert tr["OutputQuantumSubsystem"].shape == (201, model.nstates, model.nelectrons)
    assert abs(tr["OutputPosition"][-1].item() - 21.0) > 1e-3


def test_subset_kinetic_outputs_through_host_api():
    """OutputSubsetKineticEnergy / OutputFinalSubsetKineticEnergy / OutputKineticTemperature on a 100-atom bath:
    complementary subsets add up to the device's OutputKineticEnergy."""
    sb = nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), 100, 0.0, 1.0)
    sim = nq.Simulation[nq.Ehrenfest](nq.Atoms(np.ones(100)), sb)
    dist = nq.DynamicalDistribution(nq.Normal(0.0, 0.7), nq.Normal(0.0, 0.5), sim.size) * nq.PureState(1)
    lo, hi = list(range(1, 41)), list(range(41, 101))
    outs = (nq.OutputKineticEnergy, nq.OutputSubsetKineticEnergy(lo), nq.OutputKineticTemperature(hi),
            nq.OutputFinalSubsetKineticEnergy(hi), nq.OutputVelocity)
    res = nq.run_dynamics(sim, (0.0, 2.0), dist, output=outs, trajectories=5, dt=0.1, saveat=0.5, seed=9)
    for tr in res:
        ke = tr["OutputKineticEnergy"]
        sub_lo = tr["OutputSubsetKineticEnergy"]
        v = tr["OutputVelocity"].reshape(len(ke), 100)
        assert np.allclose(sub_lo, 0.5 * (v[:, :40] ** 2).sum(axis=1), rtol=1e-13)
        ke_hi = 0.5 * (v[:, 40:] ** 2).sum(axis=1)
        assert np.allclose(sub_lo + ke_hi, ke, rtol=1e-12)
        assert tr["OutputFinalSubsetKineticEnergy"] == pytest.approx(ke_hi[-1], rel=1e-13)
        assert np.allclose(tr["OutputKineticTemperature"], 2 * ke_hi / 3.166811563455546e-06 / 60, rtol=1e-13)
    with pytest.raises(ValueError):
        nq.run_dynamics(sim, (0.0, 1.0), dist, output=nq.OutputKineticTemperature(None), trajectories=4, dt=0.1, reduction=nq.MeanReduction())


def test_fermi_dirac_diabatic_state_matches_engine_gauge():
    """FermiDiracState{Diabatic} (iesh.jl:138-184): psi is built on the host from eigenvectors in the engine's default
    gauge, so Z_engine psi_e must be a single diabatic level; and the state runs through run_dynamics."""
    from helpers import A, engine_factory, model_config
    model = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(30, -0.0192, 0.0192))
    n, ne = model.nstates, model.nelectrons
    T = 6
    rng = np.random.default_rng(1)
    r = 18.0 + 6.0 * rng.random(T)
    fd = nq.FermiDiracState(0.0, 9.5e-4, nq.Diabatic())
    psi = np.zeros((T, ne, n)); occ = np.zeros((T, ne), dtype=np.int32)
    for t in range(T):
        psi[t], occ[t] = fd.sample_diabatic(rng, model.diabatic_hamiltonian([r[t]]), ne)
    kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], ntraj=T, dt=1.0, diagnostics=1, nsave=2,
                      observables=1 << A.OBS_KINETIC)
    e = engine_factory()(*A.make_config(**kw))
    e.set_state(r, np.zeros(T), psi, None, occ)
    Z = e.diagnostics()["Z"]                           # (T, n, n), Z[t][d, k] = <diabatic d | adiabatic k>
    for t in range(T):
        dia = psi[t] @ Z[t].T                          # <d | psi_e> = sum_k Z[d, k] psi_e[k]
        d = np.argmax(np.abs(dia), axis=1)
        assert np.allclose(np.abs(dia[np.arange(ne), d]), 1.0, atol=1e-9), "one diabatic level per electron"
        assert len(set(d.tolist())) == ne
    sim = nq.Simulation[nq.AdiabaticIESH](nq.Atoms(2000), model)
    dist = nq.DynamicalDistribution(nq.Normal(0.0, 7e-4), nq.Normal(21.0, 1.0), (1, 1)) * fd
    out = nq.run_dynamics(sim, (0.0, 10.0), dist, output=(nq.OutputOccupations, nq.OutputDiabaticPopulation), trajectories=4,
                          dt=1.0, seed=4)
    for tr in out:
        assert tr["OutputOccupations"].shape == (11, ne)
        assert np.allclose(tr["OutputDiabaticPopulation"].sum(axis=1), ne, atol=1e-8)


def test_mixed_state_distribution():
    """MixedState(populations, basis): sigma(0) = Diagonal(populations) in that basis (density_matrix_dynamics.jl:37-62);
    FSSH draws the active state with weights diag(sigma) (fssh.jl:53-54)."""
    sim = nq.Simulation[nq.Ehrenfest](nq.Atoms(2000), nq.TullyModelTwo())
    nuc = nq.DynamicalDistribution(16.0 / 2000, nq.Normal(-2.0, 0.3), (1, 1))
    res = nq.run_dynamics(sim, (0.0, 50.0), nuc * nq.MixedState([0.3, 0.7], nq.Diabatic()), output=(nq.OutputDiabaticPopulation,
                          nq.OutputAdiabaticPopulation), trajectories=16, dt=1.0, saveat=10.0, seed=6)
    for tr in res:
        assert np.allclose(tr["OutputDiabaticPopulation"][0], [0.3, 0.7], atol=1e-12)      # U (U' rho U) U' = rho
        assert np.allclose(tr["OutputDiabaticPopulation"].sum(axis=1), 1.0, atol=1e-9)
    res = nq.run_dynamics(sim, (0.0, 50.0), nuc * nq.MixedState([0.2, 0.8], nq.Adiabatic()), output=nq.OutputAdiabaticPopulation,
                          trajectories=4, dt=1.0, saveat=10.0, seed=6)
    for tr in res:
        assert np.allclose(tr["OutputAdiabaticPopulation"][0], [0.2, 0.8], atol=1e-12)
    fssh = nq.Simulation[nq.FSSH](nq.Atoms(2000), nq.TullyModelTwo())
    T = 4000
    for basis in (nq.Adiabatic(), nq.Diabatic()):
        out = nq.run_dynamics(fssh, (0.0, 10.0), nq.DynamicalDistribution(16.0 / 2000, -9.0, (1, 1)) * nq.MixedState([1.0, 3.0], basis),
                              output=nq.OutputDiscreteState, trajectories=T, dt=1.0, saveat=10.0, seed=7, reduction=nq.SortByOutputReduction())
        first = np.array([s[0] for s in out["OutputDiscreteState"]])
        frac2 = np.mean(first == 2)                        # at r = -9 the bases coincide: weights 1 : 3 either way
        assert abs(frac2 - 0.75) < 5 * np.sqrt(0.75 * 0.25 / T), frac2
    with pytest.raises(ValueError):
        nq.run_dynamics(sim, (0.0, 5.0), nuc * nq.MixedState([1.0], nq.Diabatic()), output=nq.OutputDiabaticPopulation, trajectories=2)


def test_nrpmd_through_host_api():
    """RingPolymerSimulation{NRPMD} through run_dynamics (nrpmd.jl:47-65 initial mapping variables): populations start
    on the PureState, the mapping-variable outputs have the reference's (nstates, nbeads) frames and radii."""
    B, g = 4, 0.5
    sim = nq.RingPolymerSimulation[nq.NRPMD](nq.Atoms(1.0), nq.DoubleWell(), B, γ=g, temperature=1.0)
    dist = nq.DynamicalDistribution(nq.Normal(0.0, 0.5), nq.Normal(0.0, 0.3), sim.size) * nq.PureState(1, nq.Diabatic())
    outs = (nq.OutputDiabaticPopulation, nq.OutputMappingPosition, nq.OutputMappingMomentum, nq.OutputTotalEnergy)
    res = nq.run_dynamics(sim, (0.0, 1.0), dist, output=outs, trajectories=6, dt=0.01, saveat=0.1, seed=8)
    for tr in res:
        q, p = tr["OutputMappingPosition"], tr["OutputMappingMomentum"]
        assert q.shape == (11, 2, B) and p.shape == (11, 2, B)
        rad2 = q[0] ** 2 + p[0] ** 2
        assert np.allclose(rad2[0], 2 + 2 * g) and np.allclose(rad2[1], 2 * g)
        assert np.allclose(tr["OutputDiabaticPopulation"][0], [1.0, 0.0], atol=1e-12)       # (q^2 + p^2)/2 - gamma per bead
        E = tr["OutputTotalEnergy"]
        assert np.max(np.abs(E - E[0])) < 1e-3 * max(1.0, abs(E[0]))
        assert not np.array_equal(q[0], q[-1])
    with pytest.raises(TypeError):
        nq.run_dynamics(sim, (0.0, 0.1), dist.nuclear * nq.PureState(1, nq.Adiabatic()), output=nq.OutputDiabaticPopulation, dt=0.01)


def test_run_dynamics_file_reduction(tmp_path):
    """test/Ensembles/reductions.jl:20-34: run_dynamics(sim, (0, 10), u0; dt = 0.1, reduction = FileReduction("test.h5"),
    output = (OutputPosition, OutputTotalEnergy)) on Simulation(Atoms(1), Harmonic()): trajectory_1/OutputPosition is a
    3-index array, trajectory_1/OutputTotalEnergy a vector."""
    sim = nq.Simulation[nq.Classical](nq.Atoms(1), nq.Harmonic())
    dist = nq.DynamicalDistribution(nq.Normal(0.0, 1.0), nq.Normal(0.0, 1.0), sim.size)
    red = nq.FileReduction(str(tmp_path / "test.h5"))
    msg = nq.run_dynamics(sim, (0.0, 10.0), dist, dt=0.1, reduction=red, output=(nq.OutputPosition, nq.OutputTotalEnergy),
                          trajectories=3, seed=1)
    assert msg == f"Output written to {red.target}."
    if red.backend == "h5py":
        import h5py
        with h5py.File(red.target, "r") as f:
            pos, ene, t = f["trajectory_1/OutputPosition"][()], f["trajectory_1/OutputTotalEnergy"][()], f["trajectory_3/Time"][()]
    else:
        with np.load(red.target) as f:
            pos, ene, t = f["trajectory_1/OutputPosition"], f["trajectory_1/OutputTotalEnergy"], f["trajectory_3/Time"]
    assert pos.shape == (1, 1, 101) and ene.shape == (101,) and np.allclose(t, np.arange(101) * 0.1)
    assert np.max(np.abs(ene - ene[0])) < 2e-2 * abs(ene[0]) + 1e-9      # velocity Verlet at omega dt = 0.1

"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol the header declares, refuses
to compute without a GPU, and the host mirror reproduces the reference's layouts / argument handling."""
import ctypes
import os
import re

import numpy as np
import pytest

import nqcdynamics_jl_b200 as nq
from helpers import A, model_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "nqcb200.h")).read()
    return sorted(set(re.findall(r"\b(nqcb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(A.engine_library_path())
    names = _header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/nqcb200.h but not exported"
    assert sorted("nqcb200_" + s for s in A.HEADER_SYMBOLS) == names, "python binding list out of sync with the header"
    assert lib.nqcb200_version() == A.ABI_VERSION


def test_config_struct_matches_header_layout():
    """Field order / sizes of the ctypes mirror against the C compiler's view of the header."""
    import subprocess, tempfile, textwrap
    fields = [f[0] for f in A.Config._fields_]
    src = textwrap.dedent("""
        #include <stdio.h>
        #include <stddef.h>
        #include "nqcb200.h"
        int main(void) {
        %s
            printf("sizeof %%zu\\n", sizeof(nqcb200_config));
            return 0;
        }""") % "\n".join(f'    printf("{f} %zu\\n", offsetof(nqcb200_config, {f}));' for f in fields)
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-I", os.path.join(ROOT, "include"),
                               os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = dict(line.split() for line in subprocess.check_output([os.path.join(d, "t")], text=True).splitlines())
    for f in fields:
        assert int(out[f]) == getattr(A.Config, f).offset, f
    assert int(out["sizeof"]) == ctypes.sizeof(A.Config)


def test_no_gpu_means_error_not_fallback():
    lib = A.load_engine_library()
    if lib.nqcb200_device_count() > 0:
        pytest.skip("a GPU is visible")
    cfg, keep = A.make_config(**model_config(nq.TullyModelOne(), method=A.METHOD_FSSH, masses=[2000.0], ntraj=4, dt=1.0))
    from nqcdynamics_jl_b200.engine import Engine
    with pytest.raises(nq.EngineError) as ei:
        Engine(cfg, keep)
    assert ei.value.code == -3
    sim = nq.Simulation[nq.FSSH](nq.Atoms(2000), nq.TullyModelOne())
    dist = nq.DynamicalDistribution(0.005, -5.0, sim.size) * nq.PureState(1)
    with pytest.raises(RuntimeError):
        nq.run_dynamics(sim, (0.0, 10.0), dist, output=nq.OutputDiabaticPopulation, trajectories=2)


def test_unsupported_and_invalid_configs_are_rejected_before_touching_a_device():
    from nqcdynamics_jl_b200.engine import Engine
    cfg, keep = A.make_config(**model_config(nq.TullyModelOne(), method=A.METHOD_FSSH, masses=[2000.0], ntraj=4, dt=1.0))
    cfg.abi_version = 99
    with pytest.raises(nq.EngineError) as ei:
        Engine(cfg, keep)
    assert ei.value.code == -1
    cfg, keep = A.make_config(**model_config(nq.Harmonic(), method=A.METHOD_FSSH, masses=[1.0], ntraj=4, dt=1.0))
    with pytest.raises(nq.EngineError) as ei:
        Engine(cfg, keep)
    assert ei.value.code == -2        # no kernel for that combination -> unsupported, never a CPU path


def test_model_table_matches_reference_docs():
    """Bath discretisations: docs/src/NQCModels/systembathmodels.md:47-59 (Ohmic), :82-94 (Debye), :210-215."""
    N = 10
    w, c = nq.OhmicSpectralDensity(2.5, 0.1).discretize(N)
    j = np.arange(1, N + 1)
    assert np.allclose(w, -2.5 * np.log(1 - j / (N + 1)))
    assert np.allclose(c, np.sqrt(0.1 * 2.5 / (N + 1)) * w)
    w, c = nq.DebyeSpectralDensity(0.25, 0.5).discretize(N)
    assert np.allclose(w, 0.25 * np.tan(np.pi / 2 * (1 - j / (N + 1))))
    assert np.allclose(c, np.sqrt(2 * 0.5 / (N + 1)) * w)
    m = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(30, -0.0192, 0.0192))
    assert m.nstates == 31 and m.nelectrons == 15          # test/Dynamics/iesh.jl:19,30,78-79
    sb = nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), 100, 0.0, 1.0)
    sim = nq.Simulation[nq.FSSH](nq.Atoms(np.ones(100)), sb)
    assert sim.size == (1, 100) and sim.ndofs_total == 100


def test_distribution_layouts():
    """Julia (ndofs, natoms, nbeads) column-major -> engine [traj][bead][dof + ndofs*atom]."""
    sim = nq.RingPolymerSimulation[nq.FSSH](nq.Atoms([1.0, 2.0]), nq.TullyModelOne(), 3, temperature=1e-3)
    assert sim.size == (1, 2, 3)
    x = np.arange(6.0).reshape(1, 2, 3)            # x[dof, atom, bead]
    d = nq.DynamicalDistribution(x, x * 10, sim.size)
    r, v = d.sample(np.random.default_rng(0), 4)
    assert r.shape == (4, 3, 2)
    for b in range(3):
        for a in range(2):
            assert r[1, b, a] == 10 * x[0, a, b] and v[1, b, a] == x[0, a, b]
    sims = nq.Simulation[nq.FSSH](nq.Atoms(2000), nq.TullyModelOne())
    d2 = nq.DynamicalDistribution(nq.VelocityBoltzmann(1e-3, [2000.0], (1, 1)), nq.Normal(-8, 1), sims.size)
    r, v = d2.sample(np.random.default_rng(1), 20000)
    assert abs(r.mean() + 8) < 0.05 and abs(v.std() - np.sqrt(1e-3 / 2000)) < 2e-5
    # OrderedSelection: 1-based indices into a vector of configurations (selections.jl:38-42)
    d3 = nq.DynamicalDistribution([np.array([[0.1]]), np.array([[0.2]]), np.array([[0.3]])],
                                  [np.array([[1.0]]), np.array([[2.0]]), np.array([[3.0]])], sims.size)
    r, v = d3.sample(np.random.default_rng(2), 2, selection=[3, 1])
    assert list(r.ravel()) == [3.0, 1.0] and list(v.ravel()) == [0.3, 0.1]


def test_shenvi_gauss_legendre_bath():
    """docs/src/NQCModels/systembathmodels.md:236-262: two Gauss-Legendre halves around the Fermi level; sum V_n^2 =
    (band width) x coupling^2 (the weights integrate 1 over each half), knots denser towards the band centre, M/2
    electrons at eF = 0; the model goes through the same AndersonHolstein table as TrapezoidalRule."""
    M, W = 40, 0.9
    bath = nq.ShenviGaussLegendre(M, -W, W)
    eps, V = bath.discretize(0.3)
    assert eps.shape == V.shape == (M,) and np.all(np.diff(eps) > 0)
    assert abs(np.sum(V ** 2) - 2 * W * 0.3 ** 2) < 1e-13
    assert np.allclose(eps, -eps[::-1]) and np.allclose(V, V[::-1])
    x, w = np.polynomial.legendre.leggauss(M // 2)
    assert np.allclose(eps[:M // 2], 0.5 * W * x - 0.5 * W) and np.allclose(V[M // 2:] ** 2, 0.5 * W * w * 0.09)
    model = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), bath)
    assert model.nstates == M + 1 and model.nelectrons == M // 2
    assert np.array_equal(model.bath_a, eps) and np.allclose(model.bath_b, V / 0.3 * np.sqrt(6.4e-3 / (2 * np.pi)))
    shifted = nq.ShenviGaussLegendre(M, -W, W, fermi_level=0.2).discretize(1.0)[0]
    assert np.count_nonzero(shifted < 0.2) == M // 2
    with pytest.raises(ValueError):
        nq.ShenviGaussLegendre(7, -1.0, 1.0).discretize(1.0)


def test_paired_configurations_stay_paired():
    """rand(distribution) draws ONE index for velocity and position (selections.jl:70-73 -> `u = rand(distribution)`;
    `u.v`, `u.r`): correlated phase-space samples from a previous run must not be re-paired at random."""
    sims = nq.Simulation[nq.FSSH](nq.Atoms(2000), nq.TullyModelOne())
    n = 50
    pos = [np.array([[float(i)]]) for i in range(n)]
    vel = [np.array([[10.0 * i]]) for i in range(n)]
    d = nq.DynamicalDistribution(vel, pos, sims.size)
    r, v = d.sample(np.random.default_rng(4), 400)
    assert np.array_equal(v.ravel(), 10.0 * r.ravel()) and len(set(r.ravel())) > 20
    with pytest.raises(ValueError):
        nq.DynamicalDistribution(vel[:3], pos, sims.size).sample(np.random.default_rng(0), 4)
    # a vector of configurations next to a samplable entry: still one index for the vector
    d2 = nq.DynamicalDistribution(nq.Normal(0.0, 1.0), pos, sims.size)
    r, v = d2.sample(np.random.default_rng(5), 10)
    assert r.shape == v.shape == (10, 1, 1)


def test_run_dynamics_argument_checks():
    sim = nq.Simulation[nq.FSSH](nq.Atoms(2000), nq.TullyModelOne(), rescaling="vinversion")
    assert sim.method.rescaling == "vinversion"
    dist = nq.DynamicalDistribution(0.005, -5.0, sim.size) * nq.PureState(1)
    with pytest.raises(TypeError):
        nq.run_dynamics(sim, (0.0, 10.0), dist, output=lambda sol, i: 0, trajectories=2)
    with pytest.raises(ValueError):
        nq.run_dynamics(sim, (0.0, 10.0), dist, output=nq.OutputDiabaticPopulation, trajectories=2, dt=1.0, saveat=2.5)
    with pytest.raises(ValueError):
        nq.OutputStateResolvedScattering1D(sim, "blah")


def test_shard_bounds_cover_range():
    from nqcdynamics_jl_b200.distributed import shard_bounds
    for T in (0, 1, 7, 1000, 1001):
        for W in (1, 2, 3, 8):
            blocks = [shard_bounds(T, W, r) for r in range(W)]
            assert blocks[0][0] == 0 and blocks[-1][1] == T
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_subset_kinetic_outputs_host_logic():
    """OutputSubsetKineticEnergy / OutputFinalSubsetKineticEnergy / OutputKineticTemperature (DynamicsOutputs.jl:111-130,
    501-523) assembled from the velocity stream: classical_kinetic_energy(masses[idx], v[:, idx]) per frame."""
    import numpy as np
    import nqcdynamics_jl_b200 as nq
    from nqcdynamics_jl_b200 import api
    rng = np.random.default_rng(4)
    masses = np.array([1.0, 3.0, 5.0, 7.0])
    sim = nq.Simulation[nq.Ehrenfest](nq.Atoms(masses), nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), 4, 0.0, 1.0))
    nsave = 6
    v = rng.standard_normal((nsave, 4))                       # (nsave, ndofs*natoms), ndofs = 1
    arrs = {api.A.OBS_VELOCITY: v}
    ke = api._finalise(sim, nq.OutputSubsetKineticEnergy([2, 4]), arrs, True)
    want = 0.5 * (3.0 * v[:, 1] ** 2 + 7.0 * v[:, 3] ** 2)
    assert np.allclose(ke, want, rtol=1e-14)
    assert api._finalise(sim, nq.OutputFinalSubsetKineticEnergy([2, 4]), arrs, True) == pytest.approx(want[-1], rel=1e-14)
    temp = api._finalise(sim, nq.OutputKineticTemperature(None), arrs, True)
    full = 0.5 * (v ** 2 * masses).sum(axis=1)
    assert np.allclose(temp, 2 * full / 3.166811563455546e-06 / 1 / 4, rtol=1e-13)
    t2 = api._finalise(sim, nq.OutputKineticTemperature([1]), arrs, True)
    assert np.allclose(t2, 2 * (0.5 * v[:, 0] ** 2) / 3.166811563455546e-06, rtol=1e-13)
    with pytest.raises(ValueError):
        api._finalise(sim, nq.OutputSubsetKineticEnergy([1]), arrs, False)
    with pytest.raises(IndexError):
        api._finalise(sim, nq.OutputSubsetKineticEnergy([5]), arrs, True)


def test_fermi_dirac_diabatic_initial_conditions_host_logic():
    """DynamicsVariables(sim, v, r, FermiDiracState{Diabatic}) (iesh.jl:138-184): every electron starts in ONE diabatic
    level, drawn from the T = 0 Fermi-Dirac filling of diag(H); distinct sorted adiabatic occupations."""
    import numpy as np
    import nqcdynamics_jl_b200 as nq
    model = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(30, -0.0192, 0.0192))
    n, ne = model.nstates, model.nelectrons
    H = model.diabatic_hamiltonian([21.0])
    fd = nq.FermiDiracState(0.0, 0.0, nq.Diabatic())
    rng = np.random.default_rng(8)
    psi, occ = fd.sample_diabatic(rng, H, ne)
    assert psi.shape == (ne, n) and occ.shape == (ne,)
    assert np.all(np.diff(occ) > 0) and occ.min() >= 1 and occ.max() <= n
    assert np.allclose((psi ** 2).sum(axis=1), 1.0, atol=1e-13)
    w, U = np.linalg.eigh(H)
    U = U * np.where(np.diag(U) < 0, -1.0, 1.0)[None, :]
    back = psi @ U.T                                   # diabatic coefficients of every electron: +-e_d
    d = np.argmax(np.abs(back), axis=1)
    assert np.allclose(np.abs(back[np.arange(ne), d]), 1.0, atol=1e-12)
    lowest = np.sort(np.argsort(np.diag(H), kind="stable")[:ne])
    assert np.array_equal(np.sort(d), lowest), "T = 0: the ne lowest diabatic levels are filled"
    assert np.allclose(w, model.adiabatic_energies([21.0]))


def test_nrpmd_initial_mapping_kat():
    """test/Dynamics/nrpmd.jl:27-40 ("Population correlation"): with DynamicsVariables(sim, v, r, PureState(1)) the
    averaged K(0) K(0)' is [1 0; 0 0] (atol 0.1 there; exact here, because every sample has populations (1, 0))."""
    import numpy as np
    from nqcdynamics_jl_b200 import api
    rng = np.random.default_rng(1)
    q, p = api.sample_nrpmd_mapping(rng, 10_000, 10, 2, 1, 0.5)
    pop = ((q ** 2 + p ** 2) / 2 - 0.5).mean(axis=1)               # Estimators.diabatic_population, nrpmd.jl:111-122
    out = np.einsum("ti,tj->ij", pop, pop) / len(pop)
    assert np.allclose(out, [[1, 0], [0, 0]], atol=1e-12)
    th = np.arctan2(p, q) % (2 * np.pi)
    assert abs(th.mean() - np.pi) < 0.02 and abs(th.var() - (2 * np.pi) ** 2 / 12) < 0.05      # uniform angles
    q2, p2 = api.sample_nrpmd_mapping(rng, 4, 3, 3, 2, 0.0)         # gamma = 0: unoccupied states sit at the origin
    assert np.all(q2[:, :, [0, 2]] == 0) and np.allclose(q2[:, :, 1] ** 2 + p2[:, :, 1] ** 2, 2.0)


def test_terminated_series_trimming_host_logic():
    """TerminatingCallback host side: sol.t / sol.u of a terminated trajectory from the fixed-shape device stream."""
    import numpy as np
    from nqcdynamics_jl_b200 import api
    se, nsave = 5, 9
    time = 2.0 + 0.5 * se * np.arange(nsave)                       # t0 = 2, dt = 0.5
    stream = np.arange(nsave, dtype=float).reshape(nsave, 1) * 10   # frame k holds 10 k
    scat = np.zeros((nsave, 4)); scat[-1] = [0, 0, 1, 0]
    arrs = {api.A.OBS_POSITION: stream, api.A.OBS_SCATTERING: scat}
    t, a = api._trim_terminated(time, arrs, -1, se, time[-1])
    assert t is time and a is arrs                                   # still running: untouched
    # terminated after 17 steps (between save 3 and save 4): 4 saveat frames, then the terminal frame twice
    t, a = api._trim_terminated(time, arrs, 17, se, 2.0 + 0.5 * 17)
    assert np.array_equal(t, [2.0, 4.5, 7.0, 9.5, 10.5, 10.5])
    assert np.array_equal(a[api.A.OBS_POSITION][:, 0], [0, 10, 20, 30, 40, 40])
    assert a[api.A.OBS_SCATTERING] is scat
    # terminated exactly on a saveat point (step 20 = save 4): that frame once from saveat, once after the affect
    t, a = api._trim_terminated(time, arrs, 20, se, 2.0 + 0.5 * 20)
    assert np.array_equal(t, [2.0, 4.5, 7.0, 9.5, 12.0, 12.0])
    assert np.array_equal(a[api.A.OBS_POSITION][:, 0], [0, 10, 20, 30, 40, 40])
    # terminated at the very last step of the span
    t, a = api._trim_terminated(time, arrs, 40, se, time[-1])
    assert len(t) == nsave + 1 and t[-1] == t[-2] == time[-1]
    # first step
    t, a = api._trim_terminated(time, arrs, 1, se, 2.5)
    assert np.array_equal(t, [2.0, 2.5, 2.5]) and np.array_equal(a[api.A.OBS_POSITION][:, 0], [0, 10, 10])


def test_file_reduction_layout(tmp_path):
    """FileReduction (reductions.jl:57-91; test/Ensembles/reductions.jl:20-34): extension forced to .h5, one group per
    trajectory, one dataset per output, a vector of (ndofs, natoms) frames becomes a 3-index array with the frames on the
    LAST axis, a vector of numbers stays a vector.  Host logic only (the writer takes the per-trajectory dictionaries
    run_dynamics builds)."""
    red = nq.FileReduction(str(tmp_path / "test.out"))
    assert red.filename.endswith("test.h5")
    assert nq.FileReduction("a.hdf5").filename == "a.hdf5" and nq.FileReduction("b.h5").filename == "b.h5"
    nsave = 11
    trajs = [{"Time": np.linspace(0, 10, nsave), "OutputPosition": np.arange(nsave * 3 * 2, dtype=float).reshape(nsave, 3, 2) + t,
              "OutputTotalEnergy": np.full(nsave, 0.5 + t),
              "OutputStateResolvedScattering1D": {"reflection": np.array([1.0, 0.0]), "transmission": np.array([0.0, 0.0])},
              "OutputSurfaceHops": 2} for t in range(3)]
    msg = red.write(trajs)
    assert msg == f"Output written to {red.target}."
    if red.backend == "h5py":
        import h5py
        with h5py.File(red.target, "r") as f:
            pos = f["trajectory_1"]["OutputPosition"][()]; ene = f["trajectory_3"]["OutputTotalEnergy"][()]
            keys = sorted(f.keys())
    else:
        with np.load(red.target) as f:
            pos = f["trajectory_1/OutputPosition"]; ene = f["trajectory_3/OutputTotalEnergy"]
            keys = sorted({k.split("/")[0] for k in f.files})
            assert f["trajectory_2/OutputStateResolvedScattering1D"].shape == (4,) and f["trajectory_2/OutputSurfaceHops"] == 2
    assert keys == ["trajectory_1", "trajectory_2", "trajectory_3"]
    assert pos.ndim == 3 and pos.shape == (3, 2, nsave) and ene.shape == (nsave,)      # Array{<:Real,3}, Vector{<:Real}
    assert np.array_equal(pos[:, :, 4], trajs[0]["OutputPosition"][4]) and np.all(ene == 2.5)
    with pytest.raises(Exception):       # the reference opens with "cw" and create_group fails on an existing group
        red.write(trajs[:1])
    assert red.write(trajs[:1], first_id=10).startswith("Output written")

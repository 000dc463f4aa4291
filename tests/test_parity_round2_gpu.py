"""GPU parity, round 2: the branches and sizes VERDICT r01 found untested on the device.

 * frustrated-hop policies :vinversion / :off (surface_hopping.jl:65,79-91,155-164; ring polymers rpsh.jl:39-50) for the
   FSSH, SpinBoson and ring-polymer kernels, on slow trajectories so that frustrated hops occur;
 * AdiabaticIESH at the BASELINE config-4 sizes (n = 101 / 201) over many steps, from random orthonormal orbitals,
   with small draws so that the unpruned hop search runs;
 * the reference's own literature pin (test/Dynamics/ehrenfest.jl:110-144) through the CUDA Ehrenfest kernel;
 * run_dynamics(EnsembleB200(2)) == EnsembleB200(1), per trajectory, bit for bit.
"""
import os

import numpy as np
import pytest

import nqcdynamics_jl_b200 as nq
from helpers import A, ALL_POP_OBS, engine_factory, make_pair, model_config, oracle_factory, rel_err
from test_parity_gpu import (IESH_OBS, _compare_observables, _compare_state, _iesh_compare, _iesh_pair, _iesh_random_state,
                             _pure_state)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POLICIES = [A.RESCALE_STANDARD, A.RESCALE_VINVERSION, A.RESCALE_OFF]


def _counters_match(e, o, need_hops=True, need_frustrated=True):
    ce, co = e.counters(), o.counters()
    assert ce["hops"] == co["hops"] and ce["frustrated"] == co["frustrated"], (ce, co)
    if need_hops:
        assert ce["hops"] > 0, ce
    if need_frustrated:
        assert ce["frustrated"] > 0, ce
    return ce


@pytest.mark.parametrize("rescaling", POLICIES)
@pytest.mark.parametrize("name,model,mass,r0,v0,dt", [
    ("tully1", nq.TullyModelOne(), 2000.0, -0.6, 4.0 / 2000, 1.0),      # KE = 0.004 < gap, sigma coherent at t0: frustrated up-hops
    ("tully2", nq.TullyModelTwo(), 2000.0, -4.0, 9.0 / 2000, 1.0),
    ("morse3", nq.ThreeStateMorse(), 20000.0, 3.2, 2e-4, 1.0),
])
def test_fssh_frustrated_hop_policies(rescaling, name, model, mass, r0, v0, dt):
    """density_step_kernel: hop attempts with too little kinetic energy -> :standard keeps v, :vinversion reflects
    v along d (surface_hopping.jl:155-164), :off accepts without rescaling (:65); every step within 1e-10, same hops."""
    T, nsteps = 64, 400
    rng = np.random.default_rng(31)
    obs = ALL_POP_OBS | (1 << A.OBS_DISCRETE_STATE)
    kw = model_config(model, method=A.METHOD_FSSH, masses=[mass], ntraj=T, dt=dt, rng=A.RNG_INJECTED, diagnostics=1,
                      save_every=10, nsave=nsteps // 10 + 1, observables=obs, per_trajectory=1, rescaling=rescaling)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    r = r0 + 0.3 * rng.standard_normal(T)
    v = v0 * (1 + 0.2 * rng.standard_normal(T))
    rho = _pure_state(T, model.nstates, 0)
    draws = rng.random((nsteps, T)) * 0.05          # small draws: many hop attempts
    sdraw = rng.random(T)
    for h in (e, o):
        h.set_state_diabatic(r, v, rho, None, None, sdraw)
        h.set_draws(draws)
    for chunk in range(nsteps // 50):
        e.run(50); o.run(50)
        _compare_state(e, o, 1e-9, f"{name} chunk {chunk}")
        de, do = e.diagnostics(), o.diagnostics()
        assert rel_err(de["accel"], do["accel"]) < 1e-9 and rel_err(de["nac"], do["nac"]) < 1e-9
    assert np.array_equal(e.observable_per_trajectory(A.OBS_DISCRETE_STATE), o.observable_per_trajectory(A.OBS_DISCRETE_STATE))
    _counters_match(e, o, need_frustrated=rescaling != A.RESCALE_OFF)
    _compare_observables(e, o, obs, 1e-9, T)


@pytest.mark.parametrize("rescaling", POLICIES)
@pytest.mark.parametrize("method", [A.METHOD_FSSH])
@pytest.mark.parametrize("nmodes,kernel", [(8, "step"), (100, "step"), (8, "kblock"), (37, "kblock"), (100, "kblock")])
def test_spin_boson_frustrated_hop_policies(rescaling, method, nmodes, kernel):
    """A cold bath (beta = 200) makes most up-hops frustrated.  kernel = "step": spinboson_step_kernel
    (kernel_spinboson.cuh, selected by energy outputs / diagnostics); "kblock": spinboson_kblock_kernel
    (kernel_spinboson_kblock.cuh, electronic outputs only, K steps per pass over the bath, 70 trajectories = ragged
    last block, 123 steps = ragged last K-block, two launches)."""
    step_kernel = kernel == "step"
    T, nsteps = (64, 120) if step_kernel else (70, 123)
    rng = np.random.default_rng(37)
    model = nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), nmodes, 0.5, 1.0)
    obs = (1 << A.OBS_POPCORR_DIABATIC) | (1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_SIGMA) | \
          (1 << A.OBS_DISCRETE_STATE)
    if step_kernel:
        obs |= (1 << A.OBS_KINETIC) | (1 << A.OBS_TOTAL_ENERGY)
    kw = model_config(model, method=method, masses=np.ones(nmodes) * (1.0 if step_kernel else 1.3), ntraj=T, dt=0.1,
                      rng=A.RNG_INJECTED, diagnostics=int(step_kernel),
                      save_every=4, nsave=nsteps // 4 + 1, observables=obs, per_trajectory=1, rescaling=rescaling)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    w = model.bath_a
    beta = 200.0
    sr = np.sqrt(1.0 / (2 * w * np.tanh(beta * w / 2))); sv = np.sqrt(w / (2 * np.tanh(beta * w / 2)))
    r = 0.2 * rng.standard_normal((T, nmodes)) * sr
    v = 0.2 * rng.standard_normal((T, nmodes)) * sv
    rho = _pure_state(T, 2, 0)
    draws = rng.random((nsteps, T)) * 0.02
    sdraw = rng.random(T)
    for h in (e, o):
        h.set_state_diabatic(r, v, rho, None, None, sdraw)
        h.set_draws(draws)
    done = 0
    for chunk in ([20] * (nsteps // 20) + ([nsteps % 20] if nsteps % 20 else [])):
        e.run(chunk); o.run(chunk)
        done += chunk
        _compare_state(e, o, 1e-9, f"after {done} steps")
        if step_kernel:
            de, do = e.diagnostics(), o.diagnostics()
            assert rel_err(de["accel"], do["accel"]) < 1e-9
    assert np.array_equal(e.observable_per_trajectory(A.OBS_DISCRETE_STATE), o.observable_per_trajectory(A.OBS_DISCRETE_STATE))
    _counters_match(e, o, need_frustrated=rescaling != A.RESCALE_OFF)
    _compare_observables(e, o, obs, 1e-9, T)


@pytest.mark.parametrize("method", [A.METHOD_FSSH, A.METHOD_EHRENFEST])
def test_spin_boson_kblock_long_run_and_unnormalised_sigma(method):
    """spinboson_kblock_kernel over a whole BASELINE config-2 job (200 steps of dt = 0.1, 100 modes) against the oracle,
    with Philox draws, plus an Ehrenfest density of trace 0.8 (force scalar A != 1: the K = 1 fallback of a block)."""
    T, nsteps = 96, 200
    rng = np.random.default_rng(53)
    model = nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), 100, 0.0, 1.0)
    obs = (1 << A.OBS_POPCORR_DIABATIC) | (1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_DISCRETE_STATE) | (1 << A.OBS_SIGMA)
    kw = model_config(model, method=method, masses=np.ones(100), ntraj=T, dt=0.1, rng=A.RNG_PHILOX, seed=4242, traj_offset=1000,
                      save_every=1, nsave=nsteps + 1, observables=obs, per_trajectory=1)
    w = model.bath_a
    sr = np.sqrt(1.0 / (2 * w * np.tanh(2.5 * w))); sv = np.sqrt(w / (2 * np.tanh(2.5 * w)))
    r = rng.standard_normal((T, 100)) * sr
    v = rng.standard_normal((T, 100)) * sv
    rho = _pure_state(T, 2, 0)
    if method == A.METHOD_EHRENFEST:
        rho[40:50] *= 0.8                      # unnormalised rows: one block of 32 trajectories takes the general-A path
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    for h in (e, o):
        h.set_state_diabatic(r, v, rho)
        h.run(nsteps)
    _compare_state(e, o, 1e-9, "final")
    _compare_observables(e, o, obs, 1e-9, T)
    assert np.max(np.abs(e.observable_per_trajectory(A.OBS_SIGMA) - o.observable_per_trajectory(A.OBS_SIGMA))) < 1e-9
    if method == A.METHOD_FSSH:
        assert np.array_equal(e.observable_per_trajectory(A.OBS_DISCRETE_STATE), o.observable_per_trajectory(A.OBS_DISCRETE_STATE))
        assert e.counters()["hops"] == o.counters()["hops"] > 0


@pytest.mark.parametrize("rescaling", POLICIES)
@pytest.mark.parametrize("nbeads", [4, 10, 16])       # 4 / 16: register FFT kernel; 10: dense normal-mode path
def test_rpsh_frustrated_hop_policies(rescaling, nbeads):
    """ring_tpt_step_kernel / ring_step_kernel (kernel_ring_tpt.cuh:413-427, kernel_ring.cuh:350-364): the ring-polymer
    rescaling and velocity inversion act on every bead with the centroid coupling (rpsh.jl:30-50)."""
    T, nsteps = 48, 300
    rng = np.random.default_rng(41)
    model, mass, temp = nq.TullyModelOne(), 2000.0, 1e-4
    obs = ALL_POP_OBS | (1 << A.OBS_DISCRETE_STATE)
    kw = model_config(model, method=A.METHOD_FSSH, masses=[mass], ntraj=T, dt=1.0, nbeads=nbeads, temperature=temp,
                      rng=A.RNG_INJECTED, diagnostics=1, save_every=10, nsave=nsteps // 10 + 1, observables=obs,
                      per_trajectory=1, rescaling=rescaling)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    r = -0.6 + 0.3 * rng.standard_normal((T, 1)) + 0.02 * rng.standard_normal((T, nbeads))
    v = 4.0 / 2000 * (1 + 0.2 * rng.standard_normal((T, 1))) + np.sqrt(temp * nbeads / mass) * rng.standard_normal((T, nbeads))
    rho = _pure_state(T, 2, 0)
    draws = rng.random((nsteps, T)) * 0.05
    sdraw = rng.random(T)
    for h in (e, o):
        h.set_state_diabatic(r, v, rho, None, None, sdraw)
        h.set_draws(draws)
    for chunk in range(nsteps // 50):
        e.run(50); o.run(50)
        _compare_state(e, o, 1e-9, f"chunk {chunk}")
        de, do = e.diagnostics(), o.diagnostics()
        assert rel_err(de["accel"], do["accel"]) < 1e-9
    assert np.array_equal(e.observable_per_trajectory(A.OBS_DISCRETE_STATE), o.observable_per_trajectory(A.OBS_DISCRETE_STATE))
    _counters_match(e, o, need_frustrated=rescaling != A.RESCALE_OFF)
    _compare_observables(e, o, obs, 1e-9, T)


# The CPU oracle needs ~1 s per n = 201 base step and ~20 s per unpruned hop search (ne (n - ne) complex LUs), so the
# full-depth n = 201 cases (4 x 50 and 2 x 50 steps; 250 s and 320 s on the GPU box's host, both green at this commit,
# profiles/r02/SUMMARY.md) run only with NQCB200_SLOW_TESTS=1; the default suite keeps a 12-step n = 201 case.
_SLOW = os.environ.get("NQCB200_SLOW_TESTS", "0") not in ("", "0")
_DEPTH = [(100, 8, 60, 1.0, 3), (100, 8, 50, 10.0, 4), (200, 1, 12, 1.0, 6)]
if _SLOW:
    _DEPTH += [(200, 4, 50, 1.0, 10), (200, 2, 50, 10.0, 10)]


@pytest.mark.parametrize("M,T,nsteps,dt,small_every", _DEPTH)
def test_iesh_config4_depth(M, T, nsteps, dt, small_every):
    """BASELINE config 4 sizes over >= 50 steps and >= 4-8 trajectories from random orthonormal orbitals; the draws are
    small enough for the pruning estimate to fail on some steps, so the full hop search (one LU + determinant lemma on
    the device vs ne (n - ne) LUs in the oracle) is exercised at n = 101 / 201."""
    rng = np.random.default_rng(43 + M + int(dt))
    chunk_len = 10 if nsteps % 10 == 0 else 6
    model, (e, o) = _iesh_pair(M, T, dt, nsteps // chunk_len + 1, save_every=chunk_len)
    n, ne = model.nstates, model.nelectrons
    r = 6.0 + 12.0 * rng.random(T)
    v = -np.abs(rng.standard_normal(T)) * 4e-3
    re, im, state = _iesh_random_state(rng, T, n, ne)
    xi = rng.random((nsteps, T))
    xi[::small_every] *= 2e-3                          # small draws: the pruning bound fails, the full search runs
    for h in (e, o):
        h.set_state(r, v, re, im, state)
        h.set_draws(xi)
    for chunk in range(nsteps // chunk_len):
        e.run(chunk_len); o.run(chunk_len)
        _iesh_compare(e, o, 1e-9, f"n={n} chunk {chunk}")
    assert e.hop_search_count() == o.hop_search_count() > 0
    ce, co = e.counters(), o.counters()
    assert ce["hops"] == co["hops"] and ce["frustrated"] == co["frustrated"], (ce, co)
    _compare_observables(e, o, IESH_OBS, 1e-9, T)
    psi = e.get_state()["sigma"]
    assert np.allclose(np.einsum("tie,tie->te", psi.conj(), psi).real, 1.0, atol=1e-11)


def test_spin_boson_ehrenfest_cuda_vs_gao_saller_curve():
    """The reference's literature pin (test/Dynamics/ehrenfest.jl:110-144: Ohmic(2.5, 0.09), N = 100, beta = 5, eps = 0,
    Delta = 1, dt = 0.1, 500 trajectories, |<sigma_z>(t) - Gao/Saller Fig. 2b| < 0.2) checked on the CUDA kernel itself,
    through run_dynamics and the C ABI -- not on the oracle."""
    N, beta, T = 100, 5.0, 500
    model = nq.SpinBoson(nq.OhmicSpectralDensity(2.5, 0.09), N, 0.0, 1.0)
    w = model.bath_a
    rng = np.random.default_rng(2020)
    sr = np.sqrt(1 / (2 * w * np.tanh(beta * w / 2))); sv = np.sqrt(w / (2 * np.tanh(beta * w / 2)))
    positions = [(rng.standard_normal(N) * sr).reshape(1, N) for _ in range(T)]
    velocities = [(rng.standard_normal(N) * sv).reshape(1, N) for _ in range(T)]
    sim = nq.Simulation[nq.Ehrenfest](nq.Atoms(np.ones(N)), model)
    dist = nq.DynamicalDistribution(velocities, positions, sim.size) * nq.PureState(1)
    out = nq.run_dynamics(sim, (0.0, 20.0), dist, selection=list(range(1, T + 1)), trajectories=T, dt=0.1,
                          output=nq.PopulationCorrelationFunction(sim, nq.Diabatic()), reduction=nq.MeanReduction())
    pc = out["PopulationCorrelationFunction"]                   # (201, 2, 2): [t][i, j] = P_i(0) P_j(t)
    result = pc[:, 0, 0] - pc[:, 0, 1]
    data = np.loadtxt(os.path.join(GOLDEN, "gao_saller_jctc_2020_fig2b.csv"), delimiter=",")
    t = 0.1 * np.arange(201)
    ref = np.interp(t, data[:, 0], data[:, 1])
    lo = t < data[0, 0]
    ref[lo] = data[0, 1] + (t[lo] - data[0, 0]) * (data[1, 1] - data[0, 1]) / (data[1, 0] - data[0, 0])
    hi = t > data[-1, 0]
    ref[hi] = data[-1, 1] + (t[hi] - data[-1, 0]) * (data[-1, 1] - data[-2, 1]) / (data[-1, 0] - data[-2, 0])
    assert np.max(np.abs(result - ref)) < 0.2


@pytest.mark.parametrize("case", ["tully_fssh", "spinboson_fssh", "rpsh", "iesh"])
def test_run_dynamics_two_shards_equal_one(case):
    """run_dynamics(..., EnsembleB200(2)) -- two handles driven from two host threads, here both on device 0 -- returns
    per trajectory exactly what EnsembleB200(1) returns (Philox keyed by the global trajectory index, one sharding rule)."""
    T = 37
    if case == "tully_fssh":
        sim = nq.Simulation[nq.FSSH](nq.Atoms(2000), nq.TullyModelOne())
        dist = nq.DynamicalDistribution(10 / 2000, nq.Normal(-4, 0.5), sim.size) * nq.PureState(2)
        kw = dict(tspan=(0.0, 600.0), dt=1.0, saveat=20.0, output=(nq.OutputDiabaticPopulation, nq.OutputDiscreteState, nq.OutputPosition))
    elif case == "spinboson_fssh":
        model = nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), 100, 0.0, 1.0)
        sim = nq.Simulation[nq.FSSH](nq.Atoms(np.ones(100)), model)
        rng = np.random.default_rng(5)
        dist = nq.DynamicalDistribution([rng.standard_normal((1, 100)) * 0.3 for _ in range(T)],
                                        [rng.standard_normal((1, 100)) for _ in range(T)], sim.size) * nq.PureState(1)
        kw = dict(tspan=(0.0, 8.0), dt=0.1, saveat=0.5, selection=list(range(1, T + 1)),
                  output=(nq.PopulationCorrelationFunction(sim, nq.Diabatic()), nq.OutputDiscreteState))
    elif case == "rpsh":
        sim = nq.RingPolymerSimulation[nq.FSSH](nq.Atoms(2000), nq.TullyModelOne(), 4, temperature=1e-3)
        dist = nq.DynamicalDistribution(nq.Normal(10 / 2000, 1e-3), nq.Normal(-3, 0.3), sim.size) * nq.PureState(1)
        kw = dict(tspan=(0.0, 400.0), dt=1.0, saveat=20.0, output=(nq.OutputDiabaticPopulation, nq.OutputDiscreteState))
    else:
        model = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(30, -0.0192, 0.0192))
        sim = nq.Simulation[nq.AdiabaticIESH](nq.Atoms(2000), model)
        dist = nq.DynamicalDistribution(nq.Normal(0.0, 2e-3), nq.Normal(12.0, 3.0), (1, 1)) * nq.FermiDiracState(0.0, 9.5e-4)
        kw = dict(tspan=(0.0, 100.0), dt=5.0, saveat=10.0, output=(nq.OutputAdiabaticPopulation, nq.OutputOccupations, nq.OutputKineticEnergy))
    tspan = kw.pop("tspan")
    one = nq.run_dynamics(sim, tspan, dist, trajectories=T, seed=77, ensemble_algorithm=nq.EnsembleB200(1), **kw)
    two = nq.run_dynamics(sim, tspan, dist, trajectories=T, seed=77,
                          ensemble_algorithm=nq.EnsembleB200(2, device_ids=[0, 0]), **kw)
    assert len(one) == len(two) == T
    for a, b in zip(one, two):
        assert a.keys() == b.keys()
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (case, k)
    # and the reduced path: the sum of two shard accumulators equals the single accumulator to rounding
    m1 = nq.run_dynamics(sim, tspan, dist, trajectories=T, seed=77, reduction=nq.MeanReduction(),
                         ensemble_algorithm=nq.EnsembleB200(1), **{**kw, "output": kw["output"][0]})
    m2 = nq.run_dynamics(sim, tspan, dist, trajectories=T, seed=77, reduction=nq.MeanReduction(),
                         ensemble_algorithm=nq.EnsembleB200(3, device_ids=[0, 0, 0]), **{**kw, "output": kw["output"][0]})
    for k in m1:
        assert np.allclose(np.asarray(m1[k]), np.asarray(m2[k]), rtol=1e-12, atol=1e-13), (case, k)


@pytest.mark.parametrize("method", [A.METHOD_IESH, A.METHOD_EHRENFEST_NA])
def test_iesh_device_side_orbitals_and_gauss_legendre_bath(method):
    """nqcb200_set_state with sig_re == NULL: psi[state[e], e] = 1 is built on the device (iesh.jl:89-128) -- bit-identical to
    uploading the identity columns -- on an AndersonHolstein model discretised with ShenviGaussLegendre (iesh.md:98-105),
    and the run agrees with the oracle."""
    T, nsteps = 6, 20
    rng = np.random.default_rng(61)
    model = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.ShenviGaussLegendre(30, -0.0192, 0.0192))
    n, ne = model.nstates, model.nelectrons
    obs = (1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_KINETIC) | (1 << A.OBS_SIGMA)
    kw = model_config(model, method=method, masses=[2000.0], ntraj=T, dt=5.0, rng=A.RNG_INJECTED, diagnostics=1,
                      save_every=5, nsave=nsteps // 5 + 1, observables=obs, per_trajectory=1)
    r = 6.0 + 12.0 * rng.random(T)
    v = -np.abs(rng.standard_normal(T)) * 3e-3
    occ = np.stack([np.sort(rng.choice(n, ne, replace=False)) + 1 for _ in range(T)]).astype(np.int32)
    psi = np.zeros((T, ne, n)); psi[np.arange(T)[:, None], np.arange(ne)[None, :], occ - 1] = 1.0
    xi = rng.random((nsteps, T)) * 0.01
    outs = []
    for dev_psi in (True, False):
        cfg, keep = A.make_config(**kw)
        e = engine_factory()(cfg, keep)
        e.set_state(r, v, None if dev_psi else psi, None, occ)
        if method == A.METHOD_IESH:
            e.set_draws(xi)
        e.run(nsteps)
        outs.append((e.get_state(), e.observable_per_trajectory(A.OBS_SIGMA)))
    assert np.array_equal(outs[0][0]["sigma"], outs[1][0]["sigma"]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][0]["r"], outs[1][0]["r"])
    cfg, keep = A.make_config(**kw)
    o = oracle_factory()(cfg, keep)
    o.set_state(r, v, psi, None, occ if method == A.METHOD_IESH else None)
    if method == A.METHOD_IESH:
        o.set_draws(xi)
    o.run(nsteps)
    so = o.get_state()
    assert rel_err(outs[0][0]["r"], so["r"]) < 1e-9 and np.max(np.abs(outs[0][0]["sigma"] - so["sigma"])) < 1e-9
    if method == A.METHOD_IESH:
        assert np.array_equal(outs[0][0]["state"], so["state"])


@pytest.mark.parametrize("method", [A.METHOD_IESH, A.METHOD_EHRENFEST_NA])
@pytest.mark.parametrize("bath", ["trapezoidal", "gauss_legendre"])
def test_iesh_erpenbeck_thoss_parity(method, bath):
    """AndersonHolstein(ErpenbeckThoss(; Γ), bath): the model of the reference's own IESH tests (test/Dynamics/iesh.jl:17-25:
    Γ = 6.4e-3, M = 30, W = 3Γ, Atoms(2000)) and of its scattering example (iesh.md:85-105, ShenviGaussLegendre).  The
    coupling depends on the position, so dH/dx is rank two; the arrowhead kernel uses
    (Z' dH Z)_ij = z0_i z0_j [h' + (f'/f)(w_i + w_j - 2h)] and must reproduce the oracle's dense Z' dV Z path step by step."""
    T, nsteps = 6, 24
    rng = np.random.default_rng(71)
    bath_obj = nq.TrapezoidalRule(30, -0.0192, 0.0192) if bath == "trapezoidal" else nq.ShenviGaussLegendre(30, -0.0192, 0.0192)
    model = nq.AndersonHolstein(nq.ErpenbeckThoss(Γ=6.4e-3), bath_obj)
    n, ne = model.nstates, model.nelectrons
    obs = IESH_OBS if method == A.METHOD_IESH else (IESH_OBS & ~((1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_DISCRETE_STATE)))
    kw = model_config(model, method=method, masses=[2000.0], ntraj=T, dt=2.0, rng=A.RNG_INJECTED, diagnostics=1,
                      save_every=4, nsave=nsteps // 4 + 1, observables=obs, per_trajectory=1)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    r = 3.0 + 5.0 * rng.random(T)                      # from the chemisorption well out to the coupling's switching region
    v = rng.standard_normal(T) * 3e-3
    re, im, state = _iesh_random_state(rng, T, n, ne)
    xi = rng.random((nsteps, T)) * 0.05
    for h in (e, o):
        h.set_state(r, v, re, im, state if method == A.METHOD_IESH else None)
        if method == A.METHOD_IESH:
            h.set_draws(xi)
    _iesh_compare(e, o, 1e-10, "t0") if method == A.METHOD_IESH else None
    for chunk in range(nsteps // 4):
        e.run(4); o.run(4)
        se, so = e.get_state(), o.get_state()
        assert rel_err(se["r"], so["r"]) < 1e-10 and rel_err(se["v"], so["v"]) < 1e-10, chunk
        assert np.max(np.abs(se["sigma"] - so["sigma"])) < 1e-10, chunk
        de, do = e.diagnostics(), o.diagnostics()
        assert rel_err(de["eig"], do["eig"]) < 1e-10 and rel_err(de["accel"], do["accel"]) < 1e-10
        assert np.max(np.abs(de["Z"] - do["Z"])) < 1e-10
        assert np.max(np.abs(de["nac"] - do["nac"])) < 1e-10 * max(1.0, np.max(np.abs(do["nac"])))
        if method == A.METHOD_IESH:
            assert np.array_equal(se["state"], so["state"])
    _compare_observables(e, o, obs, 1e-9, T)
    if method == A.METHOD_IESH:
        assert e.counters() == o.counters() or (e.counters()["hops"] == o.counters()["hops"])


def test_iesh_erpenbeck_thoss_energy_conservation():
    """test/Dynamics/iesh.jl:155-170 in spirit: a single AdiabaticIESH trajectory on the ErpenbeckThoss model conserves the
    total energy between hops (no hops here: draws of 1) -- an independent check of the rank-two force."""
    model = nq.AndersonHolstein(nq.ErpenbeckThoss(Γ=6.4e-3), nq.TrapezoidalRule(30, -0.0192, 0.0192))
    n, ne, T, nsteps = model.nstates, model.nelectrons, 4, 400
    kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], ntraj=T, dt=1.0, rng=A.RNG_INJECTED,
                      save_every=10, nsave=nsteps // 10 + 1, observables=(1 << A.OBS_TOTAL_ENERGY) | (1 << A.OBS_POSITION), per_trajectory=1)
    cfg, keep = A.make_config(**kw)
    e = engine_factory()(cfg, keep)
    r = np.array([3.2, 4.0, 5.0, 6.5]); v = np.array([2e-3, -3e-3, 1e-3, -4e-3])
    e.set_state(r, v, None, None, np.tile(np.arange(1, ne + 1, dtype=np.int32), (T, 1)))
    e.set_draws(np.ones((nsteps, T)))
    e.run(nsteps)
    E = e.observable_per_trajectory(A.OBS_TOTAL_ENERGY)[:, :, 0]
    x = e.observable_per_trajectory(A.OBS_POSITION)[:, :, 0]
    assert np.max(np.abs(x - x[:, :1])) > 0.5                       # the trajectories actually move
    kin0 = 0.5 * 2000.0 * v * v
    assert np.max(np.abs(E - E[:, :1])) < 5e-4 * np.max(kin0)       # velocity Verlet at dt = 1: 9e-7 on a kinetic energy of 0.009 (oracle: same)


@pytest.mark.parametrize("nbeads", [8, 16])
def test_rpsh_launch_shapes_agree_bit_for_bit(nbeads, monkeypatch):
    """ring_tpt_step_kernel: thread per trajectory (LPT = 1) and the warp-specialised variants the engine selects for shards
    smaller than one wave (LPT = 2, 4: owner warps + helper warps over the bead pairs) visit the same bead pairs and take the
    bead sums in the same order -- identical results, so a run does not depend on how it was sharded; parity with the oracle
    for every shape."""
    from nqcdynamics_jl_b200.engine import Engine
    import oracle
    T, nsteps = 70, 600
    rng = np.random.default_rng(41)
    model = nq.TullyModelOne()          # scattering through the crossing: hops, frustrated hops, the rescaling on every bead
    mass, temp = 2000.0, 1e-3
    kw = model_config(model, method=A.METHOD_FSSH, masses=[mass], ntraj=T, dt=1.0, nbeads=nbeads, temperature=temp, rng=A.RNG_INJECTED,
                      diagnostics=1, save_every=50, nsave=nsteps // 50 + 1, observables=ALL_POP_OBS, per_trajectory=1)
    r = -3.0 + 0.5 * rng.standard_normal((T, 1)) + 0.05 * rng.standard_normal((T, nbeads))
    v = (6.0 + 14.0 * rng.random((T, 1))) / mass + np.sqrt(temp * nbeads / mass) * 0.2 * rng.standard_normal((T, nbeads))
    rho = np.zeros((T, 2, 2)); rho[:, 0, 0] = 1.0
    draws = rng.random((nsteps, T)); sdraw = rng.random(T)
    outs = {}
    for lpt in ("1", "2", "4"):
        monkeypatch.setenv("NQCB200_RING_LPT", lpt)
        e = Engine(*A.make_config(**kw))
        e.set_state_diabatic(r, v, rho, None, None, sdraw); e.set_draws(draws)
        for n in (7, 593):
            e.run(n)
        outs[lpt] = (e.get_state(), e.observable_per_trajectory(A.OBS_TOTAL_ENERGY), e.observable_sum(A.OBS_POPCORR_DIABATIC), e.counters())
        e.close()
    monkeypatch.delenv("NQCB200_RING_LPT")
    o = oracle.OracleEngine(*A.make_config(**kw))
    o.set_state_diabatic(r, v, rho, None, None, sdraw); o.set_draws(draws)
    o.run(nsteps)
    so = o.get_state()
    for lpt, (st, E, pc, cnt) in outs.items():
        for key in ("r", "v", "sigma", "state"):
            assert np.array_equal(st[key], outs["1"][0][key]), (lpt, key)
        assert np.array_equal(E, outs["1"][1]) and cnt == outs["1"][3]
        assert np.max(np.abs(pc - outs["1"][2])) < 1e-12 * T            # block sums in a different order
        assert rel_err(st["r"], so["r"]) < 1e-9 and rel_err(st["v"], so["v"]) < 1e-9
        assert np.array_equal(st["state"], so["state"])
    assert outs["1"][3]["hops"] == o.counters()["hops"] > 0


def test_rpsh_32_beads_mid_size_shard(monkeypatch):
    """150 trajectories per SM with 32 beads: the two-member warp-specialised shape is selected and its owner count is cut
    to what shared memory holds (128 owners x 32 beads x 6 arrays = 198 KB); same bits as thread per trajectory."""
    from nqcdynamics_jl_b200.engine import Engine
    T, nsteps, B = 148 * 150, 12, 32
    rng = np.random.default_rng(43)
    model = nq.TullyModelOne()
    kw = model_config(model, method=A.METHOD_FSSH, masses=[2000.0], ntraj=T, dt=1.0, nbeads=B, temperature=1e-3, seed=17,
                      save_every=4, nsave=nsteps // 4 + 1, observables=(1 << A.OBS_POPCORR_DIABATIC) | (1 << A.OBS_TOTAL_ENERGY))
    r = -1.0 + 0.5 * rng.standard_normal((T, 1)) + 0.05 * rng.standard_normal((T, B))
    v = 12.0 / 2000.0 + np.sqrt(1e-3 * B / 2000.0) * 0.2 * rng.standard_normal((T, B))
    rho = np.zeros((T, 2, 2)); rho[:, 0, 0] = 1.0
    outs = []
    for lpt in (None, "1"):
        if lpt is None:
            monkeypatch.delenv("NQCB200_RING_LPT", raising=False)
        else:
            monkeypatch.setenv("NQCB200_RING_LPT", lpt)
        e = Engine(*A.make_config(**kw))
        e.set_state_diabatic(r, v, rho)
        e.run(nsteps)
        outs.append((e.get_state(), e.observable_sum(A.OBS_TOTAL_ENERGY), e.counters()))
        e.close()
    (sa, ea, ca), (sb, eb, cb) = outs
    assert ca == cb and ca["nonfinite"] == 0
    for key in ("r", "v", "sigma", "state"):
        assert np.array_equal(sa[key], sb[key]), key
    assert np.max(np.abs(ea - eb)) <= 1e-9 * np.max(np.abs(eb))

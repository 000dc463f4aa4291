"""GPU: BASELINE.json's FULL sizes, checked through size-independent properties of the path (the oracle cannot run
10^6 trajectories in seconds): probability conservation of every estimator, reflection + transmission = 1,
shard additivity of the observable accumulators (the multi-GPU reduction), hop counters, bounded energy drift."""
import numpy as np
import pytest

import nqcdynamics_jl_b200 as nq
from nqcdynamics_jl_b200 import workloads
from helpers import A, engine_factory

pytestmark = pytest.mark.gpu


def _run(wl, T, nsteps, *, lo=0, n=None, seed=5, fused=False, ic=None):
    n = T if n is None else n
    kw = wl.config_kwargs(n, seed=seed, traj_offset=lo)
    cfg, keep = A.make_config(**kw)
    e = engine_factory()(cfg, keep)
    sub = {k: v[lo:lo + n] for k, v in ic.items()}
    if fused:
        e.run_from_host(sub["r"], sub["v"], wl.initial_density(n), diabatic=True, nsteps=nsteps)
    else:
        wl.upload(e, sub)
        e.run(nsteps)
    return e


@pytest.mark.parametrize("name", ["spinboson_debye100_fssh", "spinboson_debye100_ehrenfest"])
def test_spinboson_million_trajectories(name):
    """configs[1]: 10^6 trajectories x 200 steps.  sum_ij C_ij(t) = sum_i P_i(0) sum_j P_j(t) = 1 per trajectory."""
    wl = workloads.get(name)
    T = 1_000_000
    ic = wl.sample(np.random.default_rng(1), T)
    e = _run(wl, T, wl.nsteps, ic=ic, fused=True)
    C = e.observable_sum(A.OBS_POPCORR_DIABATIC).reshape(wl.nsave, 2, 2)
    assert np.max(np.abs(C.sum(axis=(1, 2)) - T)) < 1e-6 * T
    if wl.method == A.METHOD_EHRENFEST:      # (the FSSH estimator mixes the active state with coherences: fssh.jl:132-142)
        assert np.all(C > -1e-6 * T)                                       # populations stay in [0, 1]
        assert abs(C[0, 0, 0] - T) < 1e-6 * T                              # PureState(1): P(0) = (1, 0)
    c = e.counters()
    assert c["steps"] == T * wl.nsteps and c["nonfinite"] == 0
    if wl.method == A.METHOD_FSSH:
        assert 0 < c["hops"] < c["steps"] and 0 <= c["frustrated"] < c["steps"]
    # shard additivity: the two halves (keyed by the global trajectory index) add up to the whole
    a = _run(wl, T, wl.nsteps, lo=0, n=T // 2, ic=ic, fused=True)
    b = _run(wl, T, wl.nsteps, lo=T // 2, n=T // 2, ic=ic, fused=False)    # also: fused and two-call paths agree
    Cab = (a.observable_sum(A.OBS_POPCORR_DIABATIC) + b.observable_sum(A.OBS_POPCORR_DIABATIC)).reshape(wl.nsave, 2, 2)
    assert np.max(np.abs(Cab - C)) < 1e-7 * T
    if wl.method == A.METHOD_FSSH:
        assert a.counters()["hops"] + b.counters()["hops"] == c["hops"]


def test_tully_scattering_full_size():
    """configs[0] at 2^20 trajectories: every trajectory ends reflected or transmitted on exactly one surface."""
    wl = workloads.get("tully1_fssh")
    T = 1 << 20
    ic = wl.sample(np.random.default_rng(2), T)
    e = _run(wl, T, wl.nsteps, ic=ic)
    scat = e.observable_sum(A.OBS_SCATTERING)
    assert np.all(scat[:-1] == 0.0) and abs(scat[-1].sum() - T) < 1e-9 * T
    pop = e.observable_sum(A.OBS_DIABATIC_POP)
    assert np.max(np.abs(pop.sum(axis=1) - T)) < 1e-8 * T
    assert scat[-1, 2:].sum() > 0.99 * T                                   # k = 10 a.u.: transmission


def test_rpmd_energy_conservation_full_size():
    """configs[2]: 32-bead RPMD on the harmonic model, 2^17 trajectories: the ring-polymer Hamiltonian is conserved."""
    wl = workloads.get("rpmd_harmonic32")
    T = 1 << 17
    ic = wl.sample(np.random.default_rng(3), T)
    e = _run(wl, T, 2000, ic=ic)
    E = e.observable_sum(A.OBS_TOTAL_ENERGY)[: 2000 // wl.save_every + 1, 0]
    assert np.max(np.abs(E - E[0])) < 1e-6 * abs(E[0])


def test_rpsh_population_conservation_full_size():
    """configs[4]: RPSH, 16 beads, ThreeStateMorse, 10^5 trajectories."""
    wl = workloads.get("rpsh_morse3_16")
    T = 100_000
    ic = wl.sample(np.random.default_rng(4), T)
    e = _run(wl, T, 1000, ic=ic)
    ns = 1000 // wl.save_every + 1
    C = e.observable_sum(A.OBS_POPCORR_DIABATIC)[:ns].reshape(ns, 3, 3)
    assert np.max(np.abs(C.sum(axis=(1, 2)) - T)) < 1e-6 * T
    assert e.counters()["nonfinite"] == 0


def test_rpsh_one_eighth_shard_matches_thread_per_trajectory(monkeypatch):
    """configs[4] as one of 8 GPUs sees it (12 500 trajectories, less than one wave of threads): the engine selects the
    warp-specialised launch shape; hop counts and the correlation function equal those of the thread-per-trajectory shape."""
    wl = workloads.get("rpsh_morse3_16")
    T, nsteps = 12_500, 1500
    ic = wl.sample(np.random.default_rng(6), T)
    ns = nsteps // wl.save_every + 1
    res = []
    for lpt in (None, "1"):
        if lpt is None:
            monkeypatch.delenv("NQCB200_RING_LPT", raising=False)
        else:
            monkeypatch.setenv("NQCB200_RING_LPT", lpt)
        e = _run(wl, T, nsteps, ic=ic)
        res.append((e.observable_sum(A.OBS_POPCORR_DIABATIC)[:ns].copy(), e.counters(), e.get_state()))
    (ca, na, sa), (cb, nb, sb) = res
    assert na == nb and na["hops"] > 0 and na["nonfinite"] == 0
    for key in ("r", "v", "sigma", "state"):
        assert np.array_equal(sa[key], sb[key]), key
    assert np.max(np.abs(ca - cb)) < 1e-9 * T
    assert np.max(np.abs(ca.reshape(ns, 3, 3).sum(axis=(1, 2)) - T)) < 1e-6 * T


def test_tully_scattering_with_termination_full_size():
    """configs[0] the way the reference's scattering scripts run it: TerminatingCallback once the particle has left the
    interaction region.  The scattering probabilities are those of the run to the end of tspan (beyond r = 4 the
    coupling is ~exp(-16): about one hop in 10^6 trajectories happens out there), in fewer steps."""
    wl = workloads.get("tully1_fssh")
    T = 1 << 20
    ic = wl.sample(np.random.default_rng(2), T)
    full = _run(wl, T, wl.nsteps, ic=ic)
    cfg, keep = A.make_config(**wl.config_kwargs(T, seed=5))
    term = engine_factory()(cfg, keep)
    term.set_termination(0, -6.0, 4.0, True)
    wl.upload(term, ic)
    term.run(wl.nsteps)
    ts = term.termination()
    assert np.count_nonzero(ts < 0) < 0.01 * T, "all but the 3-sigma late starters leave the window within tspan"
    a, b = full.observable_sum(A.OBS_SCATTERING)[-1], term.observable_sum(A.OBS_SCATTERING)[-1]
    assert abs(b.sum() - T) < 1e-9 * T
    assert np.max(np.abs(a - b)) <= 5.0, (a, b)
    cf, ct = full.counters(), term.counters()
    assert ct["steps"] == int(np.where(ts >= 0, ts, wl.nsteps).sum()) and ct["steps"] < 0.85 * cf["steps"]
    assert abs(ct["hops"] - cf["hops"]) <= 5
    print("masked / full kernel ms:", term.last_run_timing()[0], full.last_run_timing()[0])


def test_iesh_termination_frees_the_cta():
    """AdiabaticIESH scattering (iesh.md:127-138): one CTA per trajectory, so a terminated trajectory hands its CTA to the
    next one -- the masked run takes the steps the oracle semantics say and correspondingly less device time."""
    from test_parity_gpu import _iesh_model
    from helpers import model_config
    model = _iesh_model(30)
    n, ne = model.nstates, model.nelectrons
    T, nsteps = 148 * 16, 120
    rng = np.random.default_rng(12)
    r = 7.0 + 2.0 * rng.random(T)                     # half start beyond the window edge at 8 and leave after one step
    v = -np.abs(rng.standard_normal(T)) * 2e-3 - 1e-4
    psi = np.zeros((T, ne, n)); psi[:, np.arange(ne), np.arange(ne)] = 1.0
    occ = np.tile(np.arange(1, ne + 1, dtype=np.int32), (T, 1))
    kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], ntraj=T, dt=5.0, seed=3, save_every=10, nsave=nsteps // 10 + 1,
                      observables=(1 << A.OBS_KINETIC) | (1 << A.OBS_POSITION) | (1 << A.OBS_ADIABATIC_POP))
    runs = {}
    for masked in (False, True):
        e = engine_factory()(*A.make_config(**kw))
        if masked:
            e.set_termination(0, 8.0, 1e9, True)
        e.set_state(r, v, psi, None, occ)
        e.run(nsteps)
        runs[masked] = (e.last_run_timing()[0], e.counters()["steps"], e.termination(), e.observable_sum(A.OBS_ADIABATIC_POP))
    ms_full, steps_full, _, pop_full = runs[False]
    ms_mask, steps_mask, ts, pop_mask = runs[True]
    assert steps_full == T * nsteps
    assert steps_mask == int(np.where(ts >= 0, ts, nsteps).sum()) and steps_mask < 0.7 * steps_full
    assert np.count_nonzero(ts == 1) > 0.3 * T
    assert np.allclose(pop_mask.sum(axis=1), T * ne, rtol=1e-12) and np.allclose(pop_full.sum(axis=1), T * ne, rtol=1e-12)
    print("IESH masked / full kernel ms:", ms_mask, ms_full, "steps", steps_mask, steps_full)
    assert ms_mask < 0.85 * ms_full

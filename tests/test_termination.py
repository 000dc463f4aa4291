"""TerminatingCallback (src/DynamicsUtils/callbacks.jl:29) as an in-kernel termination mask.

CPU: the oracle's restatement against what the callback means (the step after which the predicate first holds ends the
trajectory; the state is frozen there; nothing else changes for the trajectories that never leave the window).
GPU: the TERM instantiation of the thread-per-trajectory kernels against the oracle with injected draws (identical
termination steps and hop sequences, states and observables within 1e-10), and the host API's ragged series.
"""
import numpy as np
import pytest

import nqcdynamics_jl_b200 as nq
from helpers import A, ALL_POP_OBS, engine_factory, make_pair, model_config, oracle_factory, rel_err

STEP_TOL = 1e-10


def _scatter_setup(T, nsteps, method=A.METHOD_FSSH, save_every=5, seed=3):
    rng = np.random.default_rng(seed)
    model = nq.TullyModelOne()
    kw = model_config(model, method=method, masses=[2000.0], ntraj=T, dt=1.0, rng=A.RNG_INJECTED, save_every=save_every,
                      nsave=nsteps // save_every + 1, observables=ALL_POP_OBS, per_trajectory=1)
    r = -3.0 + 0.5 * rng.standard_normal(T)
    v = (8.0 + 14.0 * rng.random(T)) / 2000.0            # slow ones never reach the window's edge
    rho = np.zeros((T, 2, 2)); rho[:, 0, 0] = 1.0
    return kw, r, v, rho, rng.random((nsteps, T)), rng.random(T)


def _drive(h, r, v, rho, draws, sdraw, window, nsteps, pieces=(None,)):
    if window is not None:
        h.set_termination(0, *window)
    h.set_state_diabatic(r, v, rho, None, None, sdraw)
    h.set_draws(draws)
    for n in pieces:
        h.run(nsteps if n is None else n)


def test_oracle_termination_semantics():
    T, nsteps, se = 24, 1500, 5
    kw, r, v, rho, draws, sdraw = _scatter_setup(T, nsteps, save_every=se)
    make = oracle_factory()
    free = make(*A.make_config(**kw))
    _drive(free, r, v, rho, draws, sdraw, None, nsteps)
    assert np.all(free.termination() == -1)
    pos = free.observable_per_trajectory(A.OBS_POSITION)[:, :, 0]          # (T, nsave)
    lo, hi = -4.5, 4.0
    term = make(*A.make_config(**kw))
    _drive(term, r, v, rho, draws, sdraw, (lo, hi), nsteps)
    ts = term.termination()
    assert (ts >= 0).any() and (ts < 0).any(), "the case must mix terminated and running trajectories"
    pt = term.observable_per_trajectory(A.OBS_POSITION)[:, :, 0]
    st_free, st_term = free.get_state(), term.get_state()
    for t in range(T):
        if ts[t] < 0:            # never left the window: bit-identical to the run without the callback
            assert np.all((pos[t] >= lo) & (pos[t] <= hi))
            assert np.array_equal(pt[t], pos[t])
            assert st_term["r"][t] == st_free["r"][t] and st_term["state"][t] == st_free["state"][t]
            continue
        k = ts[t] // se
        assert np.array_equal(pt[t, :k + 1], pos[t, :k + 1]), "identical up to the termination"
        assert np.all((pos[t, :k + 1] >= lo) & (pos[t, :k + 1] <= hi)) or ts[t] % se == 0
        x = st_term["r"][t].item()
        assert x < lo or x > hi, "terminated on the predicate"
        assert np.all(pt[t, k + 1:] == x), "later save points carry the terminal state"
        # the step before was still inside: advance a fresh oracle by ts - 1 steps
    one = make(*A.make_config(**kw))
    tmax = int(ts.max())
    _drive(one, r, v, rho, draws, sdraw, None, tmax - 1)
    xb = one.get_state()["r"].reshape(T)
    for t in np.nonzero(ts == tmax)[0]:
        assert lo <= xb[t] <= hi
    # counters: a terminated trajectory stops counting steps
    assert term.counters()["steps"] == int(np.where(ts >= 0, ts, nsteps).sum())
    # scattering probabilities of the terminated ensemble: transmission iff the terminal position is positive
    sc = term.observable_sum(A.OBS_SCATTERING)[-1]
    assert abs(sc.sum() - T) < 1e-12
    assert abs(sc[2:].sum() - np.count_nonzero(st_term["r"].reshape(T) > 0)) < 1e-12


def _ring_scatter_setup(T, B, nsteps, method=A.METHOD_FSSH, save_every=5, seed=8):
    rng = np.random.default_rng(seed)
    model = nq.TullyModelOne()
    temp = 1e-3
    kw = model_config(model, method=method, masses=[2000.0], ntraj=T, dt=1.0, nbeads=B, temperature=temp, rng=A.RNG_INJECTED,
                      save_every=save_every, nsave=nsteps // save_every + 1, observables=ALL_POP_OBS, per_trajectory=1)
    r = -3.0 + 0.5 * rng.standard_normal((T, 1)) + 0.05 * rng.standard_normal((T, B))
    v = (8.0 + 14.0 * rng.random((T, 1))) / 2000.0 + np.sqrt(temp * B / 2000.0) * 0.2 * rng.standard_normal((T, B))
    rho = np.zeros((T, 2, 2)); rho[:, 0, 0] = 1.0
    return kw, r, v, rho, rng.random((nsteps, T)), rng.random(T)


def test_oracle_termination_ring_polymer_centroid():
    """Ring polymers: the position-window predicate is evaluated on the centroid (a condition on
    get_centroid(get_positions(u)), the form the reference's ring-polymer scattering scripts use)."""
    T, B, nsteps, se = 12, 4, 1200, 5
    kw, r, v, rho, draws, sdraw = _ring_scatter_setup(T, B, nsteps, save_every=se)
    o = oracle_factory()(*A.make_config(**kw))
    _drive(o, r, v, rho, draws, sdraw, (-4.5, 4.0), nsteps)
    ts = o.termination()
    assert (ts >= 0).any() and (ts < 0).any()
    rc = o.get_state()["r"].reshape(T, B).mean(axis=1)
    for t in range(T):
        assert (ts[t] >= 0) == (rc[t] < -4.5 or rc[t] > 4.0)
    pos = o.observable_per_trajectory(A.OBS_POSITION)[:, :, 0]
    for t in np.nonzero(ts >= 0)[0]:
        assert np.all(pos[t, ts[t] // se + 1:] == pos[t, -1]) and abs(pos[t, -1] - rc[t]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("method", [A.METHOD_FSSH, A.METHOD_EHRENFEST])
@pytest.mark.parametrize("B,pieces", [(4, (None,)), (16, (7, 400, 293, 500)), (10, (600, 600))])
def test_ring_polymer_termination_parity(method, B, pieces):
    """TERM instantiation of ring_tpt_step_kernel (register FFT and dense normal-mode paths): identical termination steps,
    hop sequences, frozen states and observables; terminated trajectories re-entering a later launch keep their estimators."""
    T, nsteps, se = 200, 1200, 5
    kw, r, v, rho, draws, sdraw = _ring_scatter_setup(T, B, nsteps, method=method, save_every=se, seed=12)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    for h in (e, o):
        _drive(h, r, v, rho, draws, sdraw, (-4.5, 4.0), nsteps, pieces)
    te, to = e.termination(), o.termination()
    assert (to >= 0).any() and (to < 0).any()
    assert np.array_equal(te, to), "identical termination steps"
    se_, so_ = e.get_state(), o.get_state()
    for key in ("r", "v"):
        assert rel_err(se_[key], so_[key]) < 1e-9
    assert np.max(np.abs(se_["sigma"] - so_["sigma"])) < 1e-9
    if method == A.METHOD_FSSH:
        assert np.array_equal(se_["state"], so_["state"])
    for oid in range(A.OBS_COUNT):
        if ALL_POP_OBS & (1 << oid):
            a, b = e.observable_sum(oid), o.observable_sum(oid)
            assert np.max(np.abs(a - b)) <= 1e-9 * max(1.0, np.max(np.abs(b))), f"observable {oid}"
    ce, co = e.counters(), o.counters()
    assert (ce["steps"], ce["hops"], ce["frustrated"]) == (co["steps"], co["hops"], co["frustrated"])


@pytest.mark.gpu
@pytest.mark.parametrize("method", [A.METHOD_FSSH, A.METHOD_EHRENFEST])
@pytest.mark.parametrize("pieces", [(None,), (7, 400, 593, 500)])
def test_termination_parity(method, pieces):
    T, nsteps, se = 300, 1500, 5
    kw, r, v, rho, draws, sdraw = _scatter_setup(T, nsteps, method=method, save_every=se, seed=11)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    for h in (e, o):
        _drive(h, r, v, rho, draws, sdraw, (-4.5, 4.0), nsteps, pieces)
    te, to = e.termination(), o.termination()
    assert (to >= 0).any() and (to < 0).any()
    assert np.array_equal(te, to), "identical termination steps"
    se_, so_ = e.get_state(), o.get_state()
    for key in ("r", "v"):
        assert rel_err(se_[key], so_[key]) < STEP_TOL
    assert np.max(np.abs(se_["sigma"] - so_["sigma"])) < STEP_TOL
    if method == A.METHOD_FSSH:
        assert np.array_equal(se_["state"], so_["state"])
    for oid in range(A.OBS_COUNT):
        if ALL_POP_OBS & (1 << oid):
            a, b = e.observable_sum(oid), o.observable_sum(oid)
            assert np.max(np.abs(a - b)) <= 1e-9 * max(1.0, np.max(np.abs(b))), f"observable {oid}"
    for oid in (A.OBS_POSITION, A.OBS_SCATTERING, A.OBS_DIABATIC_POP):
        assert np.max(np.abs(e.observable_per_trajectory(oid) - o.observable_per_trajectory(oid))) < 1e-9
    ce, co = e.counters(), o.counters()
    assert (ce["steps"], ce["hops"], ce["frustrated"]) == (co["steps"], co["hops"], co["frustrated"])


@pytest.mark.gpu
def test_termination_reset_and_unsupported():
    T, nsteps = 64, 1200
    kw, r, v, rho, draws, sdraw = _scatter_setup(T, nsteps, seed=5)
    e = engine_factory()(*A.make_config(**kw))
    _drive(e, r, v, rho, draws, sdraw, (-4.5, 4.0), nsteps)
    first = e.termination().copy()
    assert (first >= 0).any()
    _drive(e, r, v, rho, draws, sdraw, (-4.5, 4.0), nsteps)      # set_state resets the flags: same answer again
    assert np.array_equal(e.termination(), first)
    for bad in ((3, -1.0, 1.0), (0, 2.0, 1.0), (0, float("nan"), 1.0)):     # dof out of range, empty window, NaN bound
        with pytest.raises(nq.EngineError):
            e.set_termination(*bad)
    e.set_termination(-1, 0.0, 0.0)                               # callback removed
    _drive(e, r, v, rho, draws, sdraw, None, nsteps)
    assert np.all(e.termination() == -1)
    sb = nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), 8, 0.0, 1.0)       # lane-cooperative / bath kernels: no mask
    kw2 = model_config(sb, method=A.METHOD_FSSH, masses=[1.0] * 8, ntraj=4, dt=0.1, nsave=2, observables=1 << A.OBS_KINETIC)
    e2 = engine_factory()(*A.make_config(**kw2))
    with pytest.raises(nq.EngineError):
        e2.set_termination(0, -1.0, 1.0)


@pytest.mark.gpu
def test_run_dynamics_terminating_callback():
    """Host mirror: ragged per-trajectory series, OutputFinalTime, final-state outputs (the scattering example)."""
    T, dt, saveat, tmax = 40, 1.0, 10.0, 3000.0
    sim = nq.Simulation[nq.FSSH](nq.Atoms(2000.0), nq.TullyModelOne())
    dist = nq.DynamicalDistribution(nq.Normal(15.0 / 2000, 2.0 / 2000), -5.0, (1, 1)) * nq.PureState(1, nq.Adiabatic())
    cb = nq.TerminatingCallback(nq.PositionOutside(-5.5, 5.0))
    outs = (nq.OutputPosition, nq.OutputFinalTime, nq.OutputFinalPosition, nq.OutputStateResolvedScattering1D(sim, "adiabatic"))
    res = nq.run_dynamics(sim, (0.0, tmax), dist, output=outs, trajectories=T, dt=dt, saveat=saveat, seed=2, callback=cb)
    full = nq.run_dynamics(sim, (0.0, tmax), dist, output=(nq.OutputPosition,), trajectories=T, dt=dt, saveat=saveat, seed=2)
    nterm = 0
    for tr, fr in zip(res, full):
        t_end = tr["OutputFinalTime"]
        time, pos = tr["Time"], tr["OutputPosition"].reshape(-1)
        assert time[-1] == t_end and len(time) == len(pos)
        if t_end < tmax:
            nterm += 1
            x = tr["OutputFinalPosition"].item()
            assert x < -5.5 or x > 5.0
            assert time[-2] == t_end and pos[-2] == pos[-1] == x, "terminal state saved before and after terminate!"
            k = int(t_end // saveat)
            assert np.array_equal(pos[:k + 1], fr["OutputPosition"].reshape(-1)[:k + 1])
            assert np.all(np.diff(time[:k + 1]) == saveat)
            assert len(time) == k + 1 + (2 if t_end % saveat else 1)
        sc = tr["OutputStateResolvedScattering1D"]
        assert abs(sc["reflection"].sum() + sc["transmission"].sum() - 1.0) < 1e-12
        assert (sc["transmission"].sum() > 0.5) == (tr["OutputFinalPosition"].item() > 0)
    assert nterm > T // 2
    mean = nq.run_dynamics(sim, (0.0, tmax), dist, output=(nq.OutputFinalTime, nq.OutputStateResolvedScattering1D(sim, "adiabatic")),
                           trajectories=T, dt=dt, saveat=saveat, seed=2, callback=cb, reduction=nq.MeanReduction())
    assert abs(mean["OutputFinalTime"] - np.mean([tr["OutputFinalTime"] for tr in res])) < 1e-9
    tsum = sum(tr["OutputStateResolvedScattering1D"]["transmission"] for tr in res) / T
    assert np.max(np.abs(mean["OutputStateResolvedScattering1D"]["transmission"] - tsum)) < 1e-12
    with pytest.raises(TypeError):
        nq.TerminatingCallback(lambda u, t, integ: False)


# ---- AdiabaticIESH: the CTA-per-trajectory kernel moves on to its next trajectory once terminate! has fired ---------
def _iesh_case(T=16, nsteps=30, seed=22):
    from test_parity_gpu import IESH_OBS, _iesh_model, _iesh_random_state
    rng = np.random.default_rng(seed)
    model = _iesh_model(30)
    kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], ntraj=T, dt=5.0, rng=A.RNG_INJECTED, diagnostics=1,
                      save_every=3, nsave=nsteps // 3 + 1, observables=IESH_OBS, per_trajectory=1)
    r = 7.0 + 2.5 * rng.random(T)
    v = -np.abs(rng.standard_normal(T)) * 6e-3 - 1e-3
    re, im, state = _iesh_random_state(rng, T, model.nstates, model.nelectrons)
    xi = rng.random((nsteps, T)) * 5e-4
    return kw, (r, v, re, im, state), xi, IESH_OBS


def _iesh_drive(h, st, xi, window, pieces):
    if window is not None:
        h.set_termination(0, *window)
    h.set_state(*st)
    h.set_draws(xi)
    for n in pieces:
        h.run(n)


def test_oracle_iesh_termination_outgoing():
    """Predicate of the IESH scattering example: beyond the window AND moving outwards; not tested at t0."""
    kw, st, xi, _ = _iesh_case()
    o = oracle_factory()(*A.make_config(**kw))
    _iesh_drive(o, st, xi, (8.0, 1e9, True), (30,))
    ts = o.termination()
    r0, v0 = st[0], st[1]
    assert np.all(ts[r0 + 5.0 * v0 < 7.9] == 1), "already beyond the window: ends after the FIRST step, not at t0"
    assert (ts > 1).any() and (ts < 0).any()
    inward = oracle_factory()(*A.make_config(**kw))
    _iesh_drive(inward, st, xi, (-1e9, 4.0, True), (30,))          # above hi but moving inwards (v < 0): never fires
    assert np.all(inward.termination() == -1)
    plain = oracle_factory()(*A.make_config(**kw))
    _iesh_drive(plain, st, xi, (-1e9, 4.0, False), (30,))           # without the velocity clause it fires at once
    assert np.all(plain.termination() == 1)


@pytest.mark.gpu
@pytest.mark.parametrize("pieces", [(30,), (4, 11, 15)])
def test_iesh_termination_parity(pieces):
    kw, st, xi, obs = _iesh_case()
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    for h in (e, o):
        _iesh_drive(h, st, xi, (8.0, 1e9, True), pieces)
    te, to = e.termination(), o.termination()
    assert (to == 1).any() and (to > 1).any() and (to < 0).any()
    assert np.array_equal(te, to)
    se_, so_ = e.get_state(), o.get_state()
    for key in ("r", "v"):
        assert rel_err(se_[key], so_[key]) < 1e-9
    assert np.max(np.abs(se_["sigma"] - so_["sigma"])) < 1e-9
    assert np.array_equal(se_["state"], so_["state"])
    for oid in range(A.OBS_COUNT):
        if obs & (1 << oid):
            a, b = e.observable_per_trajectory(oid), o.observable_per_trajectory(oid)
            assert np.max(np.abs(a - b)) <= 1e-8 * max(1.0, np.max(np.abs(b))), f"observable {oid}"
    ce, co = e.counters(), o.counters()
    assert (ce["steps"], ce["hops"], ce["frustrated"]) == (co["steps"], co["hops"], co["frustrated"])


def test_oracle_time_clause():
    """`|| t > tcut`: whoever is still running takes the first step that ends beyond tcut and stops there."""
    T, nsteps = 12, 200
    kw, r, v, rho, draws, sdraw = _scatter_setup(T, nsteps, save_every=5)
    o = oracle_factory()(*A.make_config(**kw))
    _drive(o, r, v, rho, draws, sdraw, (-np.inf, np.inf, False, 37.3), nsteps)
    assert np.all(o.termination() == 38)
    assert o.counters()["steps"] == 38 * T
    o2 = oracle_factory()(*A.make_config(**kw))
    _drive(o2, r, v, rho, draws, sdraw, None, 38)
    assert np.array_equal(o.get_state()["r"], o2.get_state()["r"])


@pytest.mark.gpu
def test_time_clause_parity():
    T, nsteps = 200, 1500
    kw, r, v, rho, draws, sdraw = _scatter_setup(T, nsteps, seed=17)
    e, o = make_pair(engine_factory(), oracle_factory(), **kw)
    for h in (e, o):
        _drive(h, r, v, rho, draws, sdraw, (-4.5, 4.0, True, 1203.5), nsteps, (600, 900))
    te, to = e.termination(), o.termination()
    assert np.array_equal(te, to) and to.max() == 1204 and (to < 1204).any() and (to >= 0).all()
    assert rel_err(e.get_state()["r"], o.get_state()["r"]) < STEP_TOL
    a, b = e.observable_per_trajectory(A.OBS_POSITION), o.observable_per_trajectory(A.OBS_POSITION)
    assert np.max(np.abs(a - b)) < 1e-9
    kw2, st, xi, obs = _iesh_case()
    e2, o2 = make_pair(engine_factory(), oracle_factory(), **kw2)
    for h in (e2, o2):
        _iesh_drive(h, st, xi, (8.0, 1e9, True, 61.0), (7, 23))          # dt = 5: t > 61 first holds after 13 steps
    assert np.array_equal(e2.termination(), o2.termination()) and o2.termination().max() == 13
    assert np.max(np.abs(e2.observable_per_trajectory(A.OBS_DIABATIC_POP) - o2.observable_per_trajectory(A.OBS_DIABATIC_POP))) < 1e-8

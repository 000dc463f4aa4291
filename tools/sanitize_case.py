"""Small ragged cases of every kernel family for compute-sanitizer (memcheck / racecheck).
usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import nqcdynamics_jl_b200 as nq
from nqcdynamics_jl_b200 import workloads
from nqcdynamics_jl_b200.engine import Engine
A = nq._abi
rng = np.random.default_rng(0)
for name, T, nsteps in [("spinboson_debye100_fssh", 300, 7), ("spinboson_debye100_ehrenfest", 131, 5), ("tully1_fssh", 333, 50),
                        ("rpmd_harmonic32", 70, 20), ("rpsh_morse3_16", 150, 12), ("nrpmd_morse3_16", 40, 10), ("langevin_harmonic32", 90, 12)]:
    wl = workloads.get(name)
    obs = wl.observables | (1 << A.OBS_KINETIC) | (1 << A.OBS_TOTAL_ENERGY) if name.startswith("spinboson") else wl.observables
    for per_traj in (0, 1):
        kw = wl.config_kwargs(T, seed=3, nsave=nsteps // wl.save_every + 1, observables=obs, per_trajectory=per_traj)
        e = Engine(*A.make_config(**kw))
        ic = wl.sample(rng, T)
        if wl.method in (A.METHOD_FSSH, A.METHOD_EHRENFEST):
            e.run_from_host(ic["r"], ic["v"], wl.initial_density(T), diabatic=True, nsteps=nsteps // 2)
            e.run(nsteps - nsteps // 2)
        else:
            wl.upload(e, ic)
            e.run(nsteps)
        if wl.device_spec is not None:
            rho1 = None
            if wl.method in (A.METHOD_FSSH, A.METHOD_EHRENFEST):
                rho1 = np.zeros((wl.model.nstates,) * 2); rho1[0, 0] = 1.0
            e.sample_state(wl.device_spec[0], wl.device_spec[1], rho1, diabatic=True, state=0, normal_modes=wl.device_spec[2])
            e.run(nsteps)
        e.get_state(); e.counters()
        for o in range(A.OBS_COUNT):
            if (obs >> o) & 1:
                e.observable_sum(o)
                if per_traj: e.observable_per_trajectory(o)
        e.close()
    print("ok", name)
# odd bead counts (thread-per-trajectory dense path)
for method, model, B in [(A.METHOD_FSSH, nq.ThreeStateMorse(), 10), (A.METHOD_EHRENFEST, nq.TullyModelOne(), 3), (A.METHOD_CLASSICAL, nq.Harmonic(m=1837.0, ω=0.005), 5)]:
    T = 77
    D = 1
    kw = dict(method=method, model=model.kind, nstates=model.nstates, ndofs=1, masses=[2000.0], ntraj=T, dt=1.0, nbeads=B,
              params=model.params, temperature=1e-3, save_every=2, nsave=6, nelectrons=0,
              observables=(1 << A.OBS_KINETIC) | (1 << A.OBS_TOTAL_ENERGY) | (1 << A.OBS_POSITION))
    e = Engine(*A.make_config(**kw))
    r = 2.0 + 0.1 * rng.standard_normal((T, B, 1)); v = 1e-3 * rng.standard_normal((T, B, 1))
    if method == A.METHOD_CLASSICAL:
        e.set_state(r, v)
    else:
        rho = np.zeros((T, model.nstates, model.nstates)); rho[:, 0, 0] = 1.0
        e.set_state_diabatic(r, v, rho)
    e.run(10); e.get_state(); e.close()
    print("ok beads", B)
# TerminatingCallback masks: TERM instantiation of the thread-per-trajectory kernel, and the IESH kernel's frozen path
# (terminated trajectories re-entering a later launch rebuild their eigenvectors before the save points)
T, nsteps = 333, 900
model = nq.TullyModelOne()
for method in (A.METHOD_FSSH, A.METHOD_EHRENFEST):
    kw = dict(method=method, model=model.kind, nstates=2, ndofs=1, masses=[2000.0], ntraj=T, dt=1.0, params=model.params, save_every=7,
              nsave=nsteps // 7 + 1, per_trajectory=1, seed=5,
              observables=(1 << A.OBS_POSITION) | (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_SCATTERING) | (1 << A.OBS_TOTAL_ENERGY))
    e = Engine(*A.make_config(**kw))
    e.set_termination(0, -4.5, 1.0, True, 700.5)
    rho = np.zeros((T, 2, 2)); rho[:, 0, 0] = 1.0
    e.set_state_diabatic(-3.0 + 0.5 * rng.standard_normal(T), (8.0 + 14.0 * rng.random(T)) / 2000.0, rho)
    for n in (3, 400, 497):
        e.run(n)
    ts = e.termination(); e.counters(); e.observable_per_trajectory(A.OBS_POSITION); e.close()
    assert (ts >= 0).all() and ts.max() == 701 and (ts < 701).any()
    print("ok termination", method)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from test_termination import _iesh_case, _iesh_drive
kw, st, xi, obs = _iesh_case()
for method in (A.METHOD_IESH, A.METHOD_EHRENFEST_NA):
    kw["method"] = method
    if method == A.METHOD_EHRENFEST_NA:
        kw["observables"] = (1 << A.OBS_KINETIC) | (1 << A.OBS_POSITION) | (1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_TOTAL_ENERGY)
    e = Engine(*A.make_config(**kw))
    if method == A.METHOD_EHRENFEST_NA:
        e.set_termination(0, 8.0, 1e9, True, 61.0)
        e.set_state(st[0], st[1], st[2], st[3], None)
        for n in (7, 23):
            e.run(n)
    else:
        _iesh_drive(e, st, xi, (8.0, 1e9, True, 61.0), (7, 23))
    ts = e.termination(); e.observable_per_trajectory(A.OBS_ADIABATIC_POP); e.close()
    assert ts.max() == 13 and (ts == 1).any()
    print("ok iesh termination", method)

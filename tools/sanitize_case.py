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
# round 2, last third: ring-polymer IESH / EhrenfestNA (bead arrays + kept generator in shared memory), TERM instantiation of
# the ring-polymer kernel (register FFT and dense bead counts), lanes-per-trajectory variant (subprocess: the switch is
# read at create)
from test_parity_gpu import IESH_OBS, NA_OBS, _iesh_model, _iesh_random_state
from helpers import model_config
for method, obs_, M, B in ((A.METHOD_IESH, IESH_OBS, 30, 4), (A.METHOD_EHRENFEST_NA, NA_OBS, 30, 3), (A.METHOD_IESH, IESH_OBS, 100, 2)):
    model = _iesh_model(M)
    T, nsteps = 3, 6
    kw = model_config(model, method=method, masses=[2000.0], ntraj=T, dt=5.0, nbeads=B, temperature=9.5e-4, seed=11, save_every=2,
                      nsave=nsteps // 2 + 1, observables=obs_, per_trajectory=1, diagnostics=1)
    e = Engine(*A.make_config(**kw))
    re, im, state = _iesh_random_state(rng, T, model.nstates, model.nelectrons)
    r = 10.0 + rng.standard_normal((T, B)); v = -3e-3 + 1e-4 * rng.standard_normal((T, B))
    e.set_state(r, v, re, im, state if method == A.METHOD_IESH else None)
    for n in (2, 4):
        e.run(n)
    e.get_state(); e.diagnostics(); e.observable_per_trajectory(A.OBS_TOTAL_ENERGY); e.close()
    print("ok ring-polymer iesh", method, M, B)
from test_termination import _ring_scatter_setup, _drive
for B in (4, 10):
    kw, r, v, rho, draws, sdraw = _ring_scatter_setup(70, B, 900, save_every=5)
    e = Engine(*A.make_config(**kw))
    _drive(e, r, v, rho, draws, sdraw, (-4.5, 4.0), 900, (300, 600))
    assert (e.termination() >= 0).any()
    e.observable_per_trajectory(A.OBS_POSITION); e.close()
    print("ok ring termination", B)
if os.environ.get("NQCB200_RING_LPT") is None:
    import subprocess
    env = dict(os.environ, NQCB200_RING_LPT="4", NQCB200_SANITIZE_LANES_ONLY="1")
    code = ("import sys; sys.path[:0]=[%r, %r, %r]; import numpy as np; import nqcdynamics_jl_b200 as nq; from nqcdynamics_jl_b200 import workloads; "
            "from nqcdynamics_jl_b200.engine import Engine; A=nq._abi; wl=workloads.get('rpsh_morse3_16'); T=150; "
            "e=Engine(*A.make_config(**wl.config_kwargs(T, seed=3, nsave=12//wl.save_every+1, per_trajectory=1))); ic=wl.sample(np.random.default_rng(0),T); "
            "e.run_from_host(ic['r'], ic['v'], wl.initial_density(T), diabatic=True, nsteps=6); e.run(6); e.get_state(); e.close(); print('ok lanes')"
            % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")))
    subprocess.run([sys.executable, "-c", code], env=env, check=True)

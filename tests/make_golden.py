"""Generate tests/golden/engine_*.npz: outputs of the CPU oracle (oracle/) on small seeded cases.

    python tests/make_golden.py

The GPU parity tests replay the same inputs through the CUDA engine and compare with these files, so a golden
mismatch is visible even if the oracle library were rebuilt differently on the GPU box.  The Julia reference cannot
run here (no julia binary), so these are oracle outputs; the oracle itself is pinned by tests/test_oracle_kats.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle")); sys.path.insert(0, HERE)

import nqcdynamics_jl_b200 as nq  # noqa: E402
import oracle  # noqa: E402
from helpers import A, ALL_POP_OBS, CLASSICAL_OBS, model_config  # noqa: E402


def cases():
    rng = np.random.default_rng(2026)
    T = 24
    out = {}
    # C1: TullyModelOne FSSH
    out["tully1_fssh"] = dict(model=nq.TullyModelOne(), kw=dict(method=A.METHOD_FSSH, masses=[2000.0], dt=1.0, save_every=20,
                              nsave=41, observables=ALL_POP_OBS | (1 << A.OBS_DISCRETE_STATE)), nsteps=800,
                              r=rng.normal(-4.0, 0.5, (T, 1, 1)), v=np.full((T, 1, 1), 10.0 / 2000), state0=1)
    # C2: SpinBoson (Debye, 100 modes) FSSH + Ehrenfest
    sb = nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), 100, 0.0, 1.0)
    w = sb.bath_a
    sr = np.sqrt(1 / (2 * w * np.tanh(2.5 * w))); sv = np.sqrt(w / (2 * np.tanh(2.5 * w)))
    for name, method in (("spinboson_fssh", A.METHOD_FSSH), ("spinboson_ehrenfest", A.METHOD_EHRENFEST)):
        out[name] = dict(model=sb, kw=dict(method=method, masses=np.ones(100), dt=0.1, save_every=10, nsave=11,
                         observables=(1 << A.OBS_POPCORR_DIABATIC) | (1 << A.OBS_TOTAL_ENERGY)), nsteps=100,
                         r=(rng.standard_normal((T, 100)) * sr).reshape(T, 1, 100),
                         v=(rng.standard_normal((T, 100)) * sv).reshape(T, 1, 100), state0=0)
    # C3: RPMD 32 beads, Harmonic
    out["rpmd_harmonic32"] = dict(model=nq.Harmonic(m=1837.47, ω=0.005), kw=dict(method=A.METHOD_CLASSICAL, masses=[1837.47],
                                  dt=2.5, nbeads=32, temperature=9.5e-4, save_every=20, nsave=11, observables=CLASSICAL_OBS),
                                  nsteps=200, r=0.2 * rng.standard_normal((T, 32, 1)),
                                  v=np.sqrt(9.5e-4 * 32 / 1837.47) * rng.standard_normal((T, 32, 1)), state0=None)
    # C5: RPSH 16 beads, ThreeStateMorse
    out["rpsh_morse3_16"] = dict(model=nq.ThreeStateMorse(), kw=dict(method=A.METHOD_FSSH, masses=[20000.0], dt=1.0, nbeads=16,
                                 temperature=9.5e-4, save_every=20, nsave=11, observables=ALL_POP_OBS), nsteps=200,
                                 r=rng.normal(2.1, 0.1, (T, 16, 1)), v=np.sqrt(9.5e-4 * 16 / 20000) * rng.standard_normal((T, 16, 1)),
                                 state0=0)
    # C4 (test-sized, test/Dynamics/iesh.jl:17-30): AdiabaticIESH, 31 states / 15 electrons, ground-state orbitals
    ah = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(30, -0.0192, 0.0192))
    out["iesh_m30"] = dict(model=ah, kw=dict(method=A.METHOD_IESH, masses=[2000.0], dt=5.0, save_every=5, nsave=9,
                           observables=(1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_TOTAL_ENERGY) |
                                       (1 << A.OBS_DISCRETE_STATE)), nsteps=40,
                           r=rng.normal(12.0, 4.0, (T, 1, 1)), v=-np.abs(rng.standard_normal((T, 1, 1))) * 8e-3, state0="iesh",
                           draw_scale=2e-4)
    # SURVEY 8f additions (appended: the cases above keep their seeded inputs)
    out["ehrenfest_na_m30"] = dict(model=ah, kw=dict(method=A.METHOD_EHRENFEST_NA, masses=[2000.0], dt=5.0, save_every=5, nsave=9,
                                   observables=(1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_TOTAL_ENERGY) | (1 << A.OBS_POSITION)),
                                   nsteps=40, r=rng.normal(12.0, 4.0, (T, 1, 1)), v=-np.abs(rng.standard_normal((T, 1, 1))) * 8e-3,
                                   state0="na")
    out["rpsh_morse3_10"] = dict(model=nq.ThreeStateMorse(), kw=dict(method=A.METHOD_FSSH, masses=[20000.0], dt=1.0, nbeads=10,
                                 temperature=9.5e-4, save_every=20, nsave=11, observables=ALL_POP_OBS), nsteps=200,
                                 r=rng.normal(2.1, 0.1, (T, 10, 1)), v=np.sqrt(9.5e-4 * 10 / 20000) * rng.standard_normal((T, 10, 1)),
                                 state0=0)
    out["langevin_harmonic8"] = dict(model=nq.Harmonic(m=1837.47, ω=0.005), kw=dict(method=A.METHOD_THERMAL_LANGEVIN, masses=[1837.47],
                                     dt=2.5, nbeads=8, temperature=9.5e-4, nrpmd_gamma=2e-3, save_every=20, nsave=11,
                                     observables=CLASSICAL_OBS), nsteps=200, r=0.2 * rng.standard_normal((T, 8, 1)),
                                     v=np.sqrt(9.5e-4 * 8 / 1837.47) * rng.standard_normal((T, 8, 1)), state0=None, noise=True)
    # TerminatingCallback masks (callbacks.jl:29): scattering off TullyModelOne with the window / outward-velocity / tcut
    # predicate, and the IESH scattering form of iesh.md:127-138
    out["tully1_fssh_terminating"] = dict(model=nq.TullyModelOne(), kw=dict(method=A.METHOD_FSSH, masses=[2000.0], dt=1.0, save_every=20,
                                          nsave=61, observables=ALL_POP_OBS | (1 << A.OBS_DISCRETE_STATE)), nsteps=1200,
                                          r=rng.normal(-3.0, 0.5, (T, 1, 1)), v=(8.0 + 14.0 * rng.random((T, 1, 1))) / 2000, state0=1,
                                          termination=(0, -4.5, 3.0, True, 1100.5))
    out["iesh_m30_terminating"] = dict(model=ah, kw=dict(method=A.METHOD_IESH, masses=[2000.0], dt=5.0, save_every=5, nsave=9,
                                       observables=(1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_TOTAL_ENERGY) |
                                                   (1 << A.OBS_POSITION)), nsteps=40,
                                       r=7.0 + 2.5 * rng.random((T, 1, 1)), v=-np.abs(rng.standard_normal((T, 1, 1))) * 6e-3 - 1e-3,
                                       state0="iesh", draw_scale=2e-4, termination=(0, 8.0, 1e9, True, 161.0))
    return out, T, rng


def run_case(make, case, T, draws, sdraw, noise=None):
    kw = model_config(case["model"], ntraj=T, rng=A.RNG_INJECTED, **case["kw"])
    cfg, keep = A.make_config(**kw)
    h = make(cfg, keep)
    if "termination" in case:
        h.set_termination(*case["termination"])
    if case["state0"] is None:
        h.set_state(case["r"], case["v"])
        if noise is not None:
            h.set_noise(noise)
    elif case["state0"] == "na":
        n, ne = case["model"].nstates, case["model"].nelectrons
        psi = np.zeros((T, ne, n)); psi[:, np.arange(ne), np.arange(ne)] = 1.0
        h.set_state(case["r"], case["v"], psi, None, None)
    elif case["state0"] == "iesh":
        n, ne = case["model"].nstates, case["model"].nelectrons
        psi = np.zeros((T, ne, n)); psi[:, np.arange(ne), np.arange(ne)] = 1.0
        h.set_state(case["r"], case["v"], psi, None, np.tile(np.arange(1, ne + 1, dtype=np.int32), (T, 1)))
        h.set_draws(draws)
    else:
        n = case["model"].nstates
        rho = np.zeros((T, n, n)); rho[:, case["state0"], case["state0"]] = 1.0
        h.set_state_diabatic(case["r"], case["v"], rho, None, None, sdraw)
        h.set_draws(draws)
    h.run(case["nsteps"])
    res = {"r": h.get_state()["r"], "v": h.get_state()["v"]}
    if "sigma" in h.get_state():
        res["sigma_re"] = h.get_state()["sigma"].real; res["sigma_im"] = h.get_state()["sigma"].imag
    for oid in range(A.OBS_COUNT):
        if case["kw"]["observables"] & (1 << oid):
            res[f"obs{oid}"] = h.observable_sum(oid)
    if "termination" in case:
        res["term_step"] = h.termination().astype(np.float64)
    return res


if __name__ == "__main__":
    cs, T, rng = cases()
    only = set(sys.argv[1:])
    for name, case in cs.items():
        draws = rng.random((case["nsteps"], T)) * case.get("draw_scale", 1.0); sdraw = rng.random(T)
        noise = rng.standard_normal((case["nsteps"], T, case["kw"].get("nbeads", 1))) if case.get("noise") else None
        if only and name not in only:
            continue
        res = run_case(lambda c, k: oracle.OracleEngine(c, k), case, T, draws, sdraw, noise)
        extra = {"noise": noise} if noise is not None else {}
        np.savez_compressed(os.path.join(HERE, "golden", f"engine_{name}.npz"), draws=draws, sdraw=sdraw,
                            r0=case["r"], v0=case["v"], **extra, **res)
        print(name, {k: v.shape for k, v in res.items()})

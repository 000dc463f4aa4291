"""Pin the oracle (CPU) and the CUDA engine (GPU) to per-step dumps of the REAL reference.

``baseline/julia/dump_reference.jl`` (run wherever Julia + NQCDynamics.jl are installed) writes
``tests/golden/julia_<cfg>.json``; every such file present is replayed here -- same t0 state, LAPACK's t0 eigenvectors
as the gauge reference, same uniform draws -- and r, v, sigma, w, Z, d, the carried acceleration must agree within 1e-10
relative at every step, the discrete state exactly (north_star correctness levels 1 and 2).  No file present -> those
tests are skipped ("parity unpinned"); the harness itself is always tested with a dump written in
the same schema from the oracle.
"""
import glob
import json
import os

import numpy as np
import pytest

import nqcdynamics_jl_b200 as nq
import julia_golden as jg
from helpers import A, engine_factory, oracle_factory

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
JULIA_FILES = sorted(glob.glob(os.path.join(GOLDEN, "julia_*.json")))
TOL = 1e-10


def _handle(factory, doc):
    cfg, keep = A.make_config(**jg.config_kwargs(doc))
    return factory(cfg, keep)


@pytest.mark.parametrize("path", JULIA_FILES or [None])
def test_oracle_matches_reference_dump(path):
    if path is None:
        pytest.skip("no tests/golden/julia_*.json: run baseline/julia/dump_reference.jl where Julia is installed")
    doc = jg.load(path)
    jg.compare(_handle(oracle_factory(), doc), doc, TOL)


@pytest.mark.gpu
@pytest.mark.parametrize("path", JULIA_FILES or [None])
def test_engine_matches_reference_dump(path):
    if path is None:
        pytest.skip("no tests/golden/julia_*.json: run baseline/julia/dump_reference.jl where Julia is installed")
    doc = jg.load(path)
    jg.compare(_handle(engine_factory(), doc), doc, TOL)


# ---- the harness itself: a dump of the same schema written from the oracle must replay, a perturbed one must not ----
def _tully_header():
    return {"config": "selftest_tully1_fssh", "method": "FSSH", "model": "TullyModelOne",
            "model_params": {"a": 0.01, "b": 1.6, "c": 0.005, "d": 1.0}, "masses": [2000.0], "size": [1, 1],
            "dt": 1.0, "t0": 0.0, "nstates": 2, "rescaling": "standard"}


def _selftest_dump(tmp_path, header, nsteps, T, seed, r0, v0, n, state0=1):
    rng = np.random.default_rng(seed)
    doc0 = dict(header, nsteps=nsteps, trajectories=[{}] * T)
    h = _handle(oracle_factory(), doc0)
    B = doc0.get("nbeads", 1)
    r = r0 + 0.2 * rng.standard_normal((T, B)); v = v0 * (1 + 0.1 * rng.standard_normal((T, B)))
    sre = np.zeros((T, n, n)); sre[:, state0 - 1, state0 - 1] = 1.0
    draws = rng.random((nsteps, T)) * 0.2
    path = os.path.join(tmp_path, "dump.json")
    doc = jg.write_dump(path, h, header, r, v, sre, np.zeros_like(sre), np.full(T, state0, dtype=np.int32), draws, nsteps)
    return path, doc


def test_harness_roundtrip_and_sensitivity(tmp_path):
    path, _ = _selftest_dump(str(tmp_path), _tully_header(), 300, 6, 3, -1.5, 9.0 / 2000, 2)
    doc = jg.load(path)
    worst = jg.compare(_handle(oracle_factory(), doc), doc, TOL)
    assert set(worst) >= {"r", "v", "sigma_re", "sigma_im", "w", "Z", "nac", "accel"}
    assert max(worst.values()) < 1e-13
    states = np.array([[s["state"][0] for s in tr["steps"]] for tr in doc["trajectories"]])
    assert (states != 1).any(), "the self-test dump should contain hops"
    # a reference that differs in the 9th digit must be rejected
    doc["trajectories"][2]["steps"][40]["v"][0] *= 1 + 1e-9
    with pytest.raises(AssertionError):
        jg.compare(_handle(oracle_factory(), doc), doc, TOL)
    # a flipped eigenvector gauge in the dump is followed (set_gauge_reference), not a failure
    doc = jg.load(path)
    for tr in doc["trajectories"]:
        for snap in [tr["t0"]] + tr["steps"]:
            Z = np.asarray(snap["Z"]).reshape(2, 2); Z[1] *= -1.0          # column 2 (column-major rows of the flat array)
            snap["Z"] = Z.reshape(-1).tolist()
            d = np.asarray(snap["nac"]); snap["nac"] = (-d).tolist()       # d_12 changes sign with one column
            s_re, s_im = np.asarray(snap["sigma_re"]).reshape(2, 2), np.asarray(snap["sigma_im"]).reshape(2, 2)
            s_re[0, 1] *= -1; s_re[1, 0] *= -1; s_im[0, 1] *= -1; s_im[1, 0] *= -1
            snap["sigma_re"], snap["sigma_im"] = s_re.reshape(-1).tolist(), s_im.reshape(-1).tolist()
    jg.compare(_handle(oracle_factory(), doc), doc, TOL)


@pytest.mark.gpu
def test_engine_replays_oracle_dump(tmp_path):
    """The same replay path on the GPU: the CUDA engine through the C ABI against a dump written by the oracle."""
    path, _ = _selftest_dump(str(tmp_path), _tully_header(), 300, 6, 5, -1.5, 9.0 / 2000, 2)
    doc = jg.load(path)
    jg.compare(_handle(engine_factory(), doc), doc, TOL)


# ---- ring-polymer AdiabaticIESH / EhrenfestNA dumps (BCBWavefunction): no gauge reference, the dump is canonicalised ----
def _rpiesh_selftest(tmp_path, method, seed):
    M, B, T, nsteps = 30, 4, 3, 24
    model = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(M, -0.0192, 0.0192))
    n, ne = model.nstates, model.nelectrons
    header = {"config": f"selftest_rp_{method}", "method": method, "model": "AndersonHolstein",
              "model_params": {"m": 2000.0, "omega": 2e-4, "g": 20.6097, "DeltaG": -3.8e-3, "Gamma": 6.4e-3,
                               "eps": np.asarray(model.bath_a).tolist(), "V": np.asarray(model.bath_b).tolist()},
              "masses": [2000.0], "size": [1, 1, B], "nbeads": B, "temperature": 9.5e-4, "dt": 5.0, "t0": 0.0, "nstates": n,
              "nelectrons": ne, "rescaling": "standard"}
    imp = nq.MiaoSubotnik(Γ=6.4e-3)
    header["model_params"].update(m=imp.m, omega=imp.ω, g=imp.g, DeltaG=imp.ΔG)
    rng = np.random.default_rng(seed)
    doc0 = dict(header, nsteps=nsteps, trajectories=[{}] * T)
    h = _handle(oracle_factory(), doc0)
    r = 10.0 + rng.standard_normal((T, B)); v = -3e-3 + 1e-4 * rng.standard_normal((T, B))
    psi = np.zeros((T, ne, n)); psi[:, np.arange(ne), np.arange(ne)] = 1.0
    state = np.tile(np.arange(1, ne + 1, dtype=np.int32), (T, 1)) if method == "AdiabaticIESH" else None
    draws = rng.random((nsteps, T)) * 0.05 if method == "AdiabaticIESH" else None
    path = os.path.join(tmp_path, f"dump_{method}.json")
    jg.write_dump(path, h, header, r, v, psi, None, state, draws, nsteps)
    doc = jg.load(path)
    # what LAPACK may do: arbitrary column signs of the centroid eigenvectors at t0 (kept by continuity afterwards)
    flip = np.where(rng.random(n) < 0.4, -1.0, 1.0)
    for tr in doc["trajectories"]:
        for snap in [tr["t0"]] + tr["steps"]:
            snap["Z"] = (np.asarray(snap["Z"]).reshape(n, n) * flip[:, None]).reshape(-1).tolist()
            for key in ("sigma_re", "sigma_im"):
                snap[key] = (np.asarray(snap[key]).reshape(-1, n) * flip[None, :]).reshape(-1).tolist()
            snap["nac"] = (np.asarray(snap["nac"]).reshape(-1, n, n) * flip[None, :, None] * flip[None, None, :]).reshape(-1).tolist()
    return doc


@pytest.mark.parametrize("method", ["AdiabaticIESH", "EhrenfestNA"])
def test_harness_replays_ring_polymer_wavefunction_dump(tmp_path, method):
    doc = _rpiesh_selftest(str(tmp_path), method, 7)
    worst = jg.compare(_handle(oracle_factory(), doc), doc, TOL)
    assert set(worst) >= {"r", "v", "sigma_re", "sigma_im", "w", "Z", "nac", "accel"} and max(worst.values()) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["AdiabaticIESH", "EhrenfestNA"])
def test_engine_replays_ring_polymer_wavefunction_dump(tmp_path, method):
    doc = _rpiesh_selftest(str(tmp_path), method, 8)
    jg.compare(_handle(engine_factory(), doc), doc, TOL)

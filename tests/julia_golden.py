"""Reader / comparer for the per-step dumps of the REAL reference written by ``baseline/julia/dump_reference.jl``
(``tests/golden/julia_<cfg>.json``), and a writer of the same schema from any engine handle (used to test the reader
itself where Julia is not available -- files written that way are never committed as ``julia_*.json``).

Schema (one file per config):
  config, method ("FSSH" | "Ehrenfest" | "Classical" | "AdiabaticIESH" | "NRPMD"), model (NQCModels type name),
  model_params {..}, masses [natoms], size [ndofs, natoms(, nbeads)], dt, t0, nsteps, nstates,
  optional: rescaling, nbeads, temperature, nelectrons
  trajectories: [{t0: SNAP, draws: [nsteps], steps: [SNAP ...]}]
  SNAP: r, v (Julia column-major, flat), sigma_re / sigma_im (n x n or n x ne column-major), state [1 | ne] (1-based),
        w [n], Z [n*n column-major], Z_beads [B*n*n] (ring polymers), nac [D*n*n], accel [B*D]
"""
import json

import numpy as np

import nqcdynamics_jl_b200 as nq

A = nq._abi

METHODS = {"FSSH": A.METHOD_FSSH, "Ehrenfest": A.METHOD_EHRENFEST, "Classical": A.METHOD_CLASSICAL,
           "AdiabaticIESH": A.METHOD_IESH, "NRPMD": A.METHOD_NRPMD, "EhrenfestNA": A.METHOD_EHRENFEST_NA}
RESCALING = {"standard": A.RESCALE_STANDARD, "vinversion": A.RESCALE_VINVERSION, "off": A.RESCALE_OFF}


def model_from_doc(doc):
    """NQCModels type name + dumped parameters -> this package's model table entry."""
    p, name = doc.get("model_params", {}), doc["model"]
    if name == "TullyModelOne":
        return nq.TullyModelOne(p["a"], p["b"], p["c"], p["d"])
    if name == "TullyModelTwo":
        return nq.TullyModelTwo(p["a"], p["b"], p["c"], p["d"], p["e"])
    if name == "DoubleWell":
        return nq.DoubleWell(p["mass"], p["omega"], p["gamma"], p["delta"])
    if name == "SpinBoson":
        m = nq.SpinBoson(nq.DebyeSpectralDensity(1.0, 1.0), len(p["omega"]), p["epsilon"], p["delta"])
        m.bath_a, m.bath_b = np.asarray(p["omega"], dtype=float), np.asarray(p["c"], dtype=float)
        return m
    if name == "Harmonic":
        return nq.Harmonic(m=p["m"], ω=p["omega"], r0=p["r0"], dofs=doc["size"][0])
    if name == "ThreeStateMorse":
        g = lambda *ks: tuple(float(p[k]) for k in ks)
        return nq.ThreeStateMorse(d=g("d1", "d2", "d3"), α=g("α1", "α2", "α3"), r=g("r1", "r2", "r3"), c=g("c1", "c2", "c3"),
                                  a=g("a12", "a13", "a23"), αc=g("α12", "α13", "α23"), rc=g("r12", "r13", "r23"))
    if name == "AndersonHolstein":
        M = len(p["eps"])
        m = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=p["Gamma"], m=p["m"], ω=p["omega"], g=p["g"], ΔG=p["DeltaG"]),
                                nq.TrapezoidalRule(M, -1.0, 1.0))
        m.bath_a, m.bath_b = np.asarray(p["eps"], dtype=float), np.asarray(p["V"], dtype=float)
        m.nelectrons = int(doc.get("nelectrons", m.nelectrons))
        return m
    raise KeyError(f"no model table entry for {name}")


def config_kwargs(doc, **extra):
    model = model_from_doc(doc)
    size = doc["size"]
    masses = np.repeat(np.asarray(doc["masses"], dtype=float), size[0])
    T = len(doc["trajectories"])
    kw = dict(method=METHODS[doc["method"]], model=model.kind, nstates=model.nstates, ndofs=len(masses), masses=masses,
              ntraj=T, dt=float(doc["dt"]), nbeads=int(doc.get("nbeads", size[2] if len(size) > 2 else 1)),
              nelectrons=model.nelectrons, params=model.params, bath_a=model.bath_a, bath_b=model.bath_b,
              rescaling=RESCALING[doc.get("rescaling", "standard")], rng=A.RNG_INJECTED, diagnostics=1, save_every=1,
              nsave=int(doc["nsteps"]) + 1, t0=float(doc.get("t0", 0.0)), temperature=float(doc.get("temperature", 0.0)),
              nrpmd_gamma=float(doc.get("nrpmd_gamma", 0.5)))
    kw.update(extra)
    return kw


def _ring_polymer_wavefunction(doc):
    size = doc["size"]
    return doc["method"] in ("AdiabaticIESH", "EhrenfestNA") and int(doc.get("nbeads", size[2] if len(size) > 2 else 1)) > 1


def canonical_gauge(doc):
    """Ring-polymer AdiabaticIESH / EhrenfestNA dumps: the engine takes no gauge reference there (its arrowhead solver uses
    continuity with the identity for every geometry), so the DUMP is moved into that gauge instead -- column i of the
    centroid eigenvectors is flipped when Z[i, i] < 0 at t0, together with row i of psi and row / column i of the couplings,
    in every frame (both sides then follow sign continuity in time).  Returns a transformed copy."""
    import copy
    doc = copy.deepcopy(doc)
    n = int(doc["nstates"])
    for tr in doc["trajectories"]:
        Z0 = np.asarray(tr["t0"]["Z"], dtype=np.float64).reshape(n, n)          # [col][row]
        sgn = np.where(np.diagonal(Z0) < 0.0, -1.0, 1.0)
        for snap in [tr["t0"]] + tr["steps"]:
            if "Z" in snap:
                snap["Z"] = (np.asarray(snap["Z"], dtype=np.float64).reshape(n, n) * sgn[:, None]).reshape(-1).tolist()
            for key in ("sigma_re", "sigma_im"):
                if key in snap:
                    a = np.asarray(snap[key], dtype=np.float64).reshape(-1, n)         # [electron][row]
                    snap[key] = (a * sgn[None, :]).reshape(-1).tolist()
            if "nac" in snap:
                d = np.asarray(snap["nac"], dtype=np.float64).reshape(-1, n, n)       # [dof][col][row]
                snap["nac"] = (d * sgn[None, :, None] * sgn[None, None, :]).reshape(-1).tolist()
    return doc


def _field(doc, key, where):
    """Stack one SNAP field over trajectories: where = 't0' or a step index."""
    rows = []
    for tr in doc["trajectories"]:
        snap = tr["t0"] if where == "t0" else tr["steps"][where]
        if key not in snap:
            return None
        rows.append(np.asarray(snap[key], dtype=np.float64))
    return np.stack(rows)


def upload_initial_state(h, doc):
    """The dumped t0 frame -> nqcb200_set_gauge_reference + nqcb200_set_state (+ set_mapping, set_draws)."""
    T = len(doc["trajectories"])
    Zb, Z = _field(doc, "Z_beads", "t0"), _field(doc, "Z", "t0")
    if _ring_polymer_wavefunction(doc):
        pass                                        # identity continuity; the dump was moved into that gauge (canonical_gauge)
    elif Zb is not None and Z is not None:          # ring polymer: per bead, then the centroid
        h.set_gauge_reference(np.concatenate([Zb, Z], axis=1), h.B + 1)
    elif Z is not None:
        h.set_gauge_reference(Z, 1)
    r, v = _field(doc, "r", "t0"), _field(doc, "v", "t0")
    sre, sim, st = _field(doc, "sigma_re", "t0"), _field(doc, "sigma_im", "t0"), _field(doc, "state", "t0")
    h.set_state(r, v, sre, sim, None if st is None else st.astype(np.int32))
    q, p = _field(doc, "qmap", "t0"), _field(doc, "pmap", "t0")
    if q is not None:
        h.set_mapping(q, p)
    if doc["method"] in ("FSSH", "AdiabaticIESH"):
        xi = np.stack([np.asarray(tr["draws"], dtype=np.float64) for tr in doc["trajectories"]], axis=1)
        h.set_draws(xi)
    return T


def _snapshot(h, doc):
    """Current state + diagnostics of a handle in the dump's flat layouts, stacked over trajectories."""
    st = h.get_state()
    T = h.T
    out = {"r": st["r"].reshape(T, -1), "v": st["v"].reshape(T, -1)}
    if "sigma" in st:
        s = st["sigma"].transpose(0, 2, 1).reshape(T, -1)          # [t, row, col] -> column-major flat
        out["sigma_re"], out["sigma_im"] = s.real, s.imag
    if "state" in st:
        out["state"] = st["state"].astype(np.float64)
    if h.n > 1:
        d = h.diagnostics()
        out["w"] = d["eig"]
        out["Z"] = d["Z"].transpose(0, 2, 1).reshape(T, -1)
        out["nac"] = d["nac"].transpose(0, 1, 3, 2).reshape(T, -1)
        out["accel"] = d["accel"].reshape(T, -1)
    return out


def compare(h, doc, tol=1e-10, stride=1):
    """Step ``h`` through the dump and return the worst relative deviation per field; raises on a hop mismatch."""
    if _ring_polymer_wavefunction(doc):
        doc = canonical_gauge(doc)
    upload_initial_state(h, doc)
    worst = {}

    def check(where):
        snap = _snapshot(h, doc)
        for key, mine in snap.items():
            ref = _field(doc, key, where)
            if ref is None:
                continue
            if where == "t0" and key in ("w", "Z", "nac", "accel"):
                continue        # the handles report the cache of the most recent STEP; the t0 Z went in as the gauge reference
            if key == "state":
                if not np.array_equal(np.rint(mine), np.rint(ref)):
                    raise AssertionError(f"{doc['config']}: discrete state differs at {where}")
                continue
            scale = max(1.0, float(np.max(np.abs(ref)))) if key in ("sigma_re", "sigma_im", "Z") else max(float(np.max(np.abs(ref))), 1e-300)
            worst[key] = max(worst.get(key, 0.0), float(np.max(np.abs(mine - ref))) / scale)

    check("t0")
    nsteps = int(doc["nsteps"])
    for k in range(0, nsteps, stride):
        n = min(stride, nsteps - k)
        h.run(n)
        check(k + n - 1)
    bad = {k: v for k, v in worst.items() if not v < tol}
    if bad:
        raise AssertionError(f"{doc['config']}: deviation from the reference dump above {tol:g}: {bad}")
    return worst


def write_dump(path, h, doc_header, r, v, sigma_re=None, sigma_im=None, state=None, draws=None, nsteps=10):
    """Write a file of the dump's schema from a handle (harness self-test; NOT a reference dump)."""
    T = h.T
    h.set_state(r, v, sigma_re, sigma_im, state)
    if draws is not None:
        h.set_draws(draws)
    snaps = [_snapshot(h, doc_header)]
    for _ in range(nsteps):
        h.run(1)
        snaps.append(_snapshot(h, doc_header))
    trajs = []
    for t in range(T):
        pick = lambda s: {k: (np.rint(a[t]).astype(int).tolist() if k == "state" else a[t].tolist()) for k, a in s.items()}
        trajs.append({"t0": pick(snaps[0]), "steps": [pick(s) for s in snaps[1:]],
                      "draws": (draws[:, t].tolist() if draws is not None else [])})
    doc = dict(doc_header, nsteps=nsteps, trajectories=trajs)
    with open(path, "w") as f:
        json.dump(doc, f)
    return doc


def load(path):
    with open(path) as f:
        return json.load(f)

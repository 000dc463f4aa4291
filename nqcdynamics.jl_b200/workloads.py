"""The BASELINE.json configurations as seeded synthetic workloads.

Each workload fixes the model, method, time step, run length, save grid, observables and a sampler for the
initial conditions, plus the ALGORITHMIC flop count per trajectory-step from SURVEY.md 8d (the reference's
dense formulation), which is the numerator of the reported roofline.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import numpy as np

from . import _abi as A
from . import models


def fssh_flops(n: int, D: int, f_model: float, eig: Optional[float] = None) -> float:
    """F_model + E(n) + (4n^3+3n^2+2n+9) D + 496 n^3 + 730 n^2 + 10 n."""
    E = 40.0 if n == 2 else 9.0 * n ** 3
    if eig is not None:
        E = eig
    return f_model + E + (4 * n ** 3 + 3 * n ** 2 + 2 * n + 9) * D + 496 * n ** 3 + 730 * n ** 2 + 10 * n


def ehrenfest_flops(n: int, D: int, f_model: float) -> float:
    """FSSH count minus the hop test (2nD + 10n) plus the mean-field force (2 n^2 D instead of D)."""
    return fssh_flops(n, D, f_model) - (2 * n * D + 10 * n) + (2 * n * n - 1) * D


def rpmd_flops(B: int, D: int, f_model: float) -> float:
    """RP nuclear part (BCB): 8 B^2 D dense normal-mode transforms + 14 B D, plus the model per bead."""
    return 8 * B * B * D + 14 * B * D + B * f_model


def rpsh_flops(n: int, D: int, B: int, f_model: float) -> float:
    E = 40.0 if n == 2 else 9.0 * n ** 3
    return (B + 1) * (f_model + E + (4 * n ** 3 + n * n) * D) + 8 * B * B * D + 14 * B * D + 496 * n ** 3 + 730 * n * n


def iesh_flops(n: int, ne: int, D: int = 1) -> float:
    """SURVEY.md 8d, IESH base step (no unpruned hop search): F_model + 9n^3 + 4n^3 D + n^2 D + [36n^3 + 8n^3 + 6n^2]
    + 8 n^2 ne + (8/3) ne^3 + 2 ne (n - ne) D -- the reference's dense formulation (real symmetric eigen, two
    similarity products, complex Hermitian eigen of H_eff, U = V e^{-i lambda dt} V', psi' = U psi, one complex LU)."""
    return 4.0 * n + 9.0 * n ** 3 + 4.0 * n ** 3 * D + n * n * D + 44.0 * n ** 3 + 6.0 * n * n + 8.0 * n * n * ne \
        + (8.0 / 3.0) * ne ** 3 + 2.0 * ne * (n - ne) * D


def iesh_hop_search_flops(n: int, ne: int) -> float:
    """Extra flops of a step whose hop search is not pruned: ne (n - ne) complex LUs of size ne (iesh.jl:256-266)."""
    return ne * (n - ne) * (8.0 / 3.0) * ne ** 3


@dataclass
class Workload:
    name: str
    description: str
    model: models.Model
    method: int
    masses: np.ndarray
    dt: float
    nsteps: int                  # nuclear steps of one job (tspan / dt)
    save_every: int
    observables: int
    ntraj_default: int
    flops_per_traj_step: float
    sample: Callable[[np.random.Generator, int], Dict[str, np.ndarray]]
    nbeads: int = 1
    temperature: float = 0.0
    initial_diabatic_state: int = 0      # 0-based PureState index (diabatic basis)
    rescaling: int = A.RESCALE_STANDARD
    # the same initial-condition distribution as `sample`, as (r_spec, v_spec, normal_modes) for nqcb200_sample_state
    device_spec: Optional[tuple] = None

    @property
    def nsave(self) -> int:
        return self.nsteps // self.save_every + 1

    def config_kwargs(self, ntraj: int, **extra):
        D = len(self.masses)
        kw = dict(method=self.method, model=self.model.kind, nstates=self.model.nstates, ndofs=D, masses=self.masses,
                  ntraj=ntraj, dt=self.dt, nbeads=self.nbeads, params=self.model.params, bath_a=self.model.bath_a,
                  bath_b=self.model.bath_b, save_every=self.save_every, nsave=self.nsave,
                  observables=self.observables, temperature=self.temperature, rescaling=self.rescaling,
                  nelectrons=self.model.nelectrons)
        if hasattr(self, "gamma"):
            kw["nrpmd_gamma"] = self.gamma
        kw.update(extra)
        return kw

    def upload(self, h, ic, rho=None) -> int:
        """Hand one batch of initial DynamicsVariables to an engine / oracle handle; returns the bytes uploaded."""
        T = ic["r"].shape[0]
        if self.method in (A.METHOD_FSSH, A.METHOD_EHRENFEST):
            rho = self.initial_density(T) if rho is None else rho
            h.set_state_diabatic(ic["r"], ic["v"], rho)
            return ic["r"].nbytes + ic["v"].nbytes + rho.nbytes
        if self.method == A.METHOD_IESH:
            psi, state = ic.get("psi"), ic.get("state")
            if psi is None:      # ground-state orbitals (iesh.jl:89-97) are built on the device from the occupations
                state = np.tile(np.arange(1, self.model.nelectrons + 1, dtype=np.int32), (T, 1)) if state is None else state
                h.set_state(ic["r"], ic["v"], None, None, state)
                return ic["r"].nbytes + ic["v"].nbytes + state.nbytes
            h.set_state(ic["r"], ic["v"], psi, None, state)
            return ic["r"].nbytes + ic["v"].nbytes + psi.nbytes + state.nbytes
        h.set_state(ic["r"], ic["v"])
        if self.method == A.METHOD_NRPMD:
            h.set_mapping(ic["qmap"], ic["pmap"])
            return ic["r"].nbytes + ic["v"].nbytes + ic["qmap"].nbytes + ic["pmap"].nbytes
        return ic["r"].nbytes + ic["v"].nbytes

    def iesh_ground_state(self, ntraj: int):
        """DynamicsVariables(sim, v, r) for AdiabaticIESH (iesh.jl:89-97): the ne lowest adiabatic orbitals."""
        n, ne = self.model.nstates, self.model.nelectrons
        psi = np.zeros((ntraj, ne, n))
        psi[:, np.arange(ne), np.arange(ne)] = 1.0
        return psi, np.tile(np.arange(1, ne + 1, dtype=np.int32), (ntraj, 1))

    def initial_density(self, ntraj: int) -> np.ndarray:
        n = self.model.nstates
        rho = np.zeros((ntraj, n, n))
        rho[:, self.initial_diabatic_state, self.initial_diabatic_state] = 1.0
        return rho


def _tully1_fssh() -> Workload:
    # C1: docs/src/ensemble_simulations.md:39-55 -- Atoms(2000), k = 10, r ~ Normal(-8, 1), PureState(2)
    def sample(rng, T):
        return {"r": rng.normal(-8.0, 1.0, (T, 1, 1)), "v": np.full((T, 1, 1), 10.0 / 2000.0)}
    obs = (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_SCATTERING)
    return Workload("tully1_fssh", "C1 TullyModelOne FSSH, n=2, D=1, mass 2000, k0=10, dt=1, tspan (0,3000), saveat 10",
                    models.TullyModelOne(), A.METHOD_FSSH, np.array([2000.0]), 1.0, 3000, 10, obs, 1 << 22,
                    fssh_flops(2, 1, 30.0), sample, initial_diabatic_state=1,
                    device_spec=([(-8.0, 1.0)], [10.0 / 2000.0], False))


def _spinboson(method: int, name: str) -> Workload:
    # C2: SpinBoson(DebyeSpectralDensity(0.25, 0.5), 100, 0, 1), beta = 5, dt = 0.1, tspan (0, 20)
    N = 100
    model = models.SpinBoson(models.DebyeSpectralDensity(0.25, 0.5), N, 0.0, 1.0)
    w = model.bath_a
    beta = 5.0
    sr = np.sqrt(1.0 / (2.0 * w * np.tanh(beta * w / 2.0)))   # harmonic Wigner, m = 1
    sv = np.sqrt(w / (2.0 * np.tanh(beta * w / 2.0)))

    def sample(rng, T):
        return {"r": (rng.standard_normal((T, N)) * sr).reshape(T, 1, N),
                "v": (rng.standard_normal((T, N)) * sv).reshape(T, 1, N)}
    obs = (1 << A.OBS_POPCORR_DIABATIC)
    flops = fssh_flops(2, N, 6.0 * N) if method == A.METHOD_FSSH else ehrenfest_flops(2, N, 6.0 * N)
    return Workload(name, "C2 SpinBoson (Debye bath, 100 modes), n=2, D=100, beta=5, dt=0.1, tspan (0,20), saveat 0.1",
                    model, method, np.ones(N), 0.1, 200, 1, obs, 1_000_000, flops, sample,
                    device_spec=([(0.0, float(x)) for x in sr], [(0.0, float(x)) for x in sv], False))


def _rpmd_harmonic() -> Workload:
    # C3: RingPolymerSimulation{Classical}(Atoms(:H), Harmonic(m, w), 32; T = 300 K), exact thermal sample
    B, m, w = 32, 1837.4715941070515, 0.005
    kT = 300.0 * 3.166811563e-6
    beta_B = 1.0 / (kT * B)
    omega_n = B * kT
    wk = 2.0 * omega_n * np.sin(np.arange(B) * np.pi / B)
    U = np.empty((B, B))
    for k in range(B):
        j = np.arange(B)
        if k == 0: U[:, k] = 1 / np.sqrt(B)
        elif 2 * k < B: U[:, k] = np.sqrt(2 / B) * np.cos(2 * np.pi * j * k / B)
        elif 2 * k == B: U[:, k] = (-1.0) ** j / np.sqrt(B)
        else: U[:, k] = np.sqrt(2 / B) * np.sin(2 * np.pi * j * k / B)
    sr = np.sqrt(1.0 / (beta_B * m * (wk ** 2 + w ** 2)))
    sv = np.sqrt(1.0 / (beta_B * m))

    def sample(rng, T):
        rn = rng.standard_normal((T, B)) * sr
        vn = rng.standard_normal((T, B)) * sv
        return {"r": (rn @ U.T).reshape(T, B, 1), "v": (vn @ U.T).reshape(T, B, 1)}
    obs = (1 << A.OBS_POSITION) | (1 << A.OBS_KINETIC) | (1 << A.OBS_TOTAL_ENERGY)
    return Workload("rpmd_harmonic32", "C3 RPMD 32 beads, Harmonic(m_H, w=0.005), 300 K thermal sample, dt=2.5, 10^4 steps, saveat 100",
                    models.Harmonic(m=m, ω=w), A.METHOD_CLASSICAL, np.array([m]), 2.5, 10000, 100, obs, 1 << 18,
                    rpmd_flops(B, 1, 3.0), sample, nbeads=B, temperature=kT,
                    device_spec=([(0.0, float(x)) for x in sr], [(0.0, float(sv))] * B, True))


def _langevin_harmonic() -> Workload:
    # RingPolymerSimulation{ThermalLangevin}(Atoms(:H), Harmonic(m, w), 32; gamma, T = 300 K) -- the
    # thermostatted run that GENERATES the C3 / C5 thermal distributions (BCOCB), started cold at the minimum
    base = _rpmd_harmonic()
    B = base.nbeads

    def sample(rng, T):
        return {"r": np.zeros((T, B, 1)), "v": np.zeros((T, B, 1))}
    wl = Workload("langevin_harmonic32", "ThermalLangevin (BCOCB, PILE) 32 beads, Harmonic(m_H, w=0.005), 300 K, gamma=0.002, dt=2.5, 10^4 steps, saveat 100",
                  base.model, A.METHOD_THERMAL_LANGEVIN, base.masses, 2.5, 10000, 100, base.observables, 1 << 18,
                  base.flops_per_traj_step + 4 * B, sample, nbeads=B, temperature=base.temperature,
                  device_spec=([0.0] * B, [0.0] * B, False))
    wl.gamma = 0.002
    return wl


def _rpsh_morse() -> Workload:
    # C5: RingPolymerSimulation{FSSH}(Atoms(20000), ThreeStateMorse(), 16; T = 300 K)
    B, m = 16, 20000.0
    kT = 300.0 * 3.166811563e-6

    def sample(rng, T):
        return {"r": rng.normal(2.1, 1.0 / np.sqrt(m * 0.005), (T, B, 1)),
                "v": rng.standard_normal((T, B, 1)) * np.sqrt(kT * B / m)}
    obs = (1 << A.OBS_POPCORR_DIABATIC)
    return Workload("rpsh_morse3_16", "C5 RPSH 16 beads, ThreeStateMorse, mass 20000, 300 K, dt=1, tspan (0,3000), saveat 50",
                    models.ThreeStateMorse(), A.METHOD_FSSH, np.array([m]), 1.0, 3000, 50, obs, 100_000,
                    rpsh_flops(3, 1, B, 60.0), sample, nbeads=B, temperature=kT,
                    device_spec=([(2.1, 1.0 / np.sqrt(m * 0.005))] * B, [(0.0, float(np.sqrt(kT * B / m)))] * B, False))


def nrpmd_flops(n: int, D: int, B: int, f_model: float) -> float:
    """16 B^2 D + B [F_model + E(n) + 4n^3 D + 4n^3 + 2n^2 + 8n^2 + D (8n^3 + 2n^2 + 6n^2)]."""
    E = 40.0 if n == 2 else 9.0 * n ** 3
    return 16 * B * B * D + B * (f_model + E + 4 * n ** 3 * D + 4 * n ** 3 + 10 * n * n + D * (8 * n ** 3 + 8 * n * n))


def _nrpmd_morse() -> Workload:
    # C5 (second half): RingPolymerSimulation{NRPMD}(Atoms(20000), ThreeStateMorse(), 16; T = 300 K), gamma = 0.5;
    # mapping variables on the focused-sampling circles of nrpmd.jl:47-65 (state 1 occupied), uniform angles
    B, m, n, gamma = 16, 20000.0, 3, 0.5
    kT = 300.0 * 3.166811563e-6
    radius = np.full(n, np.sqrt(2 * gamma)); radius[0] = np.sqrt(2 + 2 * gamma)

    def sample(rng, T):
        th = rng.random((T, B, n)) * 2 * np.pi
        return {"r": rng.normal(2.1, 1.0 / np.sqrt(m * 0.005), (T, B, 1)),
                "v": rng.standard_normal((T, B, 1)) * np.sqrt(kT * B / m),
                "qmap": radius * np.cos(th), "pmap": radius * np.sin(th)}
    obs = (1 << A.OBS_POPCORR_DIABATIC)
    return Workload("nrpmd_morse3_16", "C5 NRPMD 16 beads, ThreeStateMorse, mass 20000, 300 K, gamma=0.5, dt=1, tspan (0,3000), saveat 50",
                    models.ThreeStateMorse(), A.METHOD_NRPMD, np.array([m]), 1.0, 3000, 50, obs, 100_000,
                    nrpmd_flops(n, 1, B, 60.0), sample, nbeads=B, temperature=kT)


def _iesh(M: int, nsteps: int, ntraj: int) -> Workload:
    # C4: Simulation{AdiabaticIESH}(Atoms(2000), AndersonHolstein(MiaoSubotnik(G = 6.4e-3), TrapezoidalRule(M, -W, W))),
    # W = 6G/2 (test/Dynamics/iesh.jl:17-25), kT = 9.5e-4, dt = 1 (iesh.jl:158); thermal sample in the U1 well.
    imp = models.MiaoSubotnik(Γ=6.4e-3)
    W = 3.0 * imp.Γ
    model = models.AndersonHolstein(imp, models.TrapezoidalRule(M, -W, W))
    kT, m = 9.5e-4, 2000.0
    sr, sv = np.sqrt(kT / (m * imp.ω ** 2)), np.sqrt(kT / m)

    def sample(rng, T):
        return {"r": rng.normal(imp.g, sr, (T, 1, 1)), "v": rng.normal(0.0, sv, (T, 1, 1))}
    obs = (1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_KINETIC)
    n, ne = model.nstates, model.nelectrons
    return Workload(f"iesh_anderson_holstein_m{M}",
                    f"C4 AdiabaticIESH, AndersonHolstein(MiaoSubotnik, TrapezoidalRule({M}, -3G, 3G)), n={n}, ne={ne}, mass 2000, "
                    f"kT=9.5e-4 thermal sample around x=g, ground-state orbitals, dt=1, tspan (0,{nsteps}), saveat 10",
                    model, A.METHOD_IESH, np.array([m]), 1.0, nsteps, 10, obs, ntraj, iesh_flops(n, ne), sample)


def _iesh_scattering(M: int, nsteps: int, ntraj: int) -> Workload:
    # The same model driven hard (not a BASELINE config; VERDICT r01: the thermal well is the propagator's best case):
    # every trajectory starts at x = g moving towards the crossing with 0.05 hartree of kinetic energy (~50 kT), so v.d is
    # large (more Horner stages with G), the orbitals leave the ground state and the unpruned hop search runs more often.
    wl = _iesh(M, nsteps, ntraj)
    imp = models.MiaoSubotnik(Γ=6.4e-3)
    m, ke = 2000.0, 0.05
    v0 = -np.sqrt(2.0 * ke / m)

    def sample(rng, T):
        return {"r": rng.normal(imp.g, 0.05, (T, 1, 1)), "v": np.full((T, 1, 1), v0) * (1.0 + 0.05 * rng.standard_normal((T, 1, 1)))}
    return Workload(f"iesh_scattering_m{M}", wl.description.replace("kT=9.5e-4 thermal sample around x=g", "incident kinetic energy 0.05 Eh (~50 kT) from x=g towards the crossing"),
                    wl.model, wl.method, wl.masses, wl.dt, nsteps, wl.save_every, wl.observables, ntraj, wl.flops_per_traj_step, sample)


def _rpiesh(M: int, B: int, nsteps: int, ntraj: int) -> Workload:
    # RingPolymerSimulation{AdiabaticIESH}(Atoms(2000), AndersonHolstein(MiaoSubotnik, TrapezoidalRule(M)), B beads) with
    # BCBWavefunction (rpiesh.jl, test/Dynamics/rpiesh.jl:11-22), thermal ring polymers around x = g
    wl = _iesh(M, nsteps, ntraj)
    imp = models.MiaoSubotnik(Γ=6.4e-3)
    kT, m = 9.5e-4, 2000.0
    sr, sv = np.sqrt(kT / (m * imp.ω ** 2)), np.sqrt(kT * B / m)

    def sample(rng, T):
        return {"r": rng.normal(imp.g, sr, (T, 1, 1)) + 0.1 * sr * rng.standard_normal((T, B, 1)), "v": rng.normal(0.0, sv, (T, B, 1))}
    return Workload(f"rpiesh_anderson_holstein_m{M}_b{B}", f"RPIESH (SURVEY 8f rank 3): {B}-bead ring polymer, " + wl.description,
                    wl.model, wl.method, wl.masses, wl.dt, nsteps, wl.save_every, wl.observables, ntraj,
                    wl.flops_per_traj_step * (1.0 + 0.15 * B), sample, nbeads=B, temperature=kT)


def get(name: str) -> Workload:
    table = {
        "iesh_scattering_m100": lambda: _iesh_scattering(100, 1000, 10_000),
        "rpiesh_anderson_holstein_m100_b4": lambda: _rpiesh(100, 4, 1000, 10_000),
        "iesh_anderson_holstein_m100": lambda: _iesh(100, 1000, 10_000),
        "iesh_anderson_holstein_m200": lambda: _iesh(200, 200, 10_000),
        "iesh_anderson_holstein_m30": lambda: _iesh(30, 1000, 10_000),
        "tully1_fssh": _tully1_fssh,
        "spinboson_debye100_fssh": lambda: _spinboson(A.METHOD_FSSH, "spinboson_debye100_fssh"),
        "spinboson_debye100_ehrenfest": lambda: _spinboson(A.METHOD_EHRENFEST, "spinboson_debye100_ehrenfest"),
        "rpmd_harmonic32": _rpmd_harmonic,
        "rpsh_morse3_16": _rpsh_morse,
        "nrpmd_morse3_16": _nrpmd_morse,
        "langevin_harmonic32": _langevin_harmonic,
    }
    if name not in table:
        raise KeyError(f"unknown workload {name!r}; available: {sorted(table)}")
    return table[name]()


NAMES = ["tully1_fssh", "spinboson_debye100_fssh", "spinboson_debye100_ehrenfest", "rpmd_harmonic32", "rpsh_morse3_16",
         "nrpmd_morse3_16", "langevin_harmonic32",
         "iesh_anderson_holstein_m100", "iesh_anderson_holstein_m200", "iesh_anderson_holstein_m30", "iesh_scattering_m100",
         "rpiesh_anderson_holstein_m100_b4"]

"""Model table: NQCModels.jl constructors -> (model enum, params[], bath arrays) of the C ABI.

The reference's models live in the external package NQCModels.jl (not under /root/reference); the
constructor names, keyword names and defaults below follow its documentation in the reference tree
(docs/src/NQCModels/analyticmodels.md, systembathmodels.md) and SURVEY.md section 8c / A.5.  The numeric
evaluation of V(r) and dV/dr happens on the device (csrc/models.cuh) -- nothing here is numerics.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _abi


@dataclass
class Model:
    kind: int
    nstates: int
    params: Sequence[float] = ()
    bath_a: Optional[np.ndarray] = None
    bath_b: Optional[np.ndarray] = None
    ndofs: int = 1            # dofs per atom (reference: NQCModels.ndofs)
    natoms: Optional[int] = None  # fixed by the model (SpinBoson: one atom per bath mode)
    nelectrons: int = 0
    name: str = ""
    classical: bool = False
    fermi_level: float = 0.0

    def adiabatic_energies(self, r) -> np.ndarray:
        """Eigenvalues of the diabatic Hamiltonian at one configuration -- host-side, used ONLY to draw Fermi-Dirac
        initial occupations (the reference does the same inside sample_distribution, iesh.jl:114-120)."""
        if self.kind not in _abi.ANDERSON_HOLSTEIN_FAMILY:
            raise NotImplementedError("adiabatic_energies: only needed for AndersonHolstein initial conditions")
        return np.linalg.eigvalsh(self.diabatic_hamiltonian(r))

    def diabatic_hamiltonian(self, r) -> np.ndarray:
        """The n x n electronic Hamiltonian the IESH cache diagonalises (state-independent U0 removed): host-side, for
        the Fermi-Dirac initial conditions only (iesh.jl:114-120, 153-160)."""
        if self.kind not in _abi.ANDERSON_HOLSTEIN_FAMILY:
            raise NotImplementedError("diabatic_hamiltonian: only needed for AndersonHolstein initial conditions")
        q = float(np.asarray(r).reshape(-1)[0])
        n = self.nstates
        H = np.zeros((n, n))
        if self.kind == _abi.MODEL_ANDERSON_HOLSTEIN_MIAO_SUBOTNIK:
            m, w, g, dG = self.params
            H[0, 0] = 0.5 * m * w * w * (q - g) ** 2 + dG - 0.5 * m * w * w * q * q
            f = 1.0
        else:
            De, a, x0, c, D1, D2, a1, x01, Vinf, qq, at, xt = self.params
            u0 = De * (math.exp(-a * (q - x0)) - 1.0) ** 2 + c
            e1 = math.exp(-a1 * (q - x01))
            H[0, 0] = D1 * e1 * e1 - D2 * e1 + Vinf - u0
            f = 0.5 * (1.0 - qq) * (1.0 - math.tanh((q - xt) / at)) + qq
        H[np.arange(1, n), np.arange(1, n)] = self.bath_a
        H[0, 1:] = H[1:, 0] = self.bath_b * f
        return H


def TullyModelOne(a=0.01, b=1.6, c=0.005, d=1.0) -> Model:
    return Model(_abi.MODEL_TULLY_ONE, 2, (a, b, c, d), name="TullyModelOne")


def TullyModelTwo(a=0.1, b=0.28, c=0.015, d=0.06, e=0.05) -> Model:
    return Model(_abi.MODEL_TULLY_TWO, 2, (a, b, c, d, e), name="TullyModelTwo")


def TullyModelThree(a=6e-4, b=0.1, c=0.9) -> Model:
    return Model(_abi.MODEL_TULLY_THREE, 2, (a, b, c), name="TullyModelThree")


def DoubleWell(mass=1.0, ω=1.0, γ=1.0, Δ=1.0) -> Model:
    return Model(_abi.MODEL_DOUBLE_WELL, 2, (mass, ω, γ, Δ), name="DoubleWell")


@dataclass
class OhmicSpectralDensity:
    """J(w) = pi/2 alpha w exp(-w/wc); discretisation docs/src/NQCModels/systembathmodels.md:47-59."""
    ωᶜ: float
    α: float

    def discretize(self, N: int):
        j = np.arange(1, N + 1)
        ω = -self.ωᶜ * np.log(1.0 - j / (N + 1.0))
        c = math.sqrt(self.α * self.ωᶜ / (N + 1.0)) * ω
        return ω, c


@dataclass
class DebyeSpectralDensity:
    """J(w) = 2 lambda wc w/(wc^2+w^2); discretisation systembathmodels.md:82-94."""
    ωᶜ: float
    λ: float

    def discretize(self, N: int):
        j = np.arange(1, N + 1)
        ω = self.ωᶜ * np.tan(np.pi / 2.0 * (1.0 - j / (N + 1.0)))
        c = math.sqrt(2.0 * self.λ / (N + 1.0)) * ω
        return ω, c


def SpinBoson(density, N: int, ϵ: float, Δ: float) -> Model:
    ω, c = density.discretize(N)
    return Model(_abi.MODEL_SPIN_BOSON, 2, (ϵ, Δ), bath_a=ω, bath_b=c, natoms=N, name="SpinBoson")


def ThreeStateMorse(d=(0.02, 0.02, 0.003), α=(0.4, 0.65, 0.65), r=(4.0, 4.5, 6.0), c=(0.02, 0.0, 0.02),
                    a=(0.005, 0.005, 0.0), αc=(32.0, 32.0, 0.0), rc=(3.40, 4.97, 0.0)) -> Model:
    """Coronado/Xing/Miller 2001 three-state Morse model (pairs ordered 12, 13, 23); parameters recalled
    from NQCModels ('NOT in tree')."""
    return Model(_abi.MODEL_THREE_STATE_MORSE, 3, (*d, *α, *r, *c, *a, *αc, *rc), name="ThreeStateMorse")


def Harmonic(m=1.0, ω=1.0, r0=0.0, dofs=1) -> Model:
    return Model(_abi.MODEL_HARMONIC, 1, (m, ω, r0), ndofs=dofs, name="Harmonic", classical=True)


def Free(dofs=1) -> Model:
    return Model(_abi.MODEL_FREE, 1, (), ndofs=dofs, name="Free", classical=True)


@dataclass
class TrapezoidalRule:
    """eps_n = a + (n-1)(b-a)/(M-1), V_n = sqrt((b-a)/(M-1)) V(eps)  (systembathmodels.md:210-215)."""
    M: int
    bandmin: float
    bandmax: float

    def discretize(self, coupling: float):
        eps = self.bandmin + np.arange(self.M) * (self.bandmax - self.bandmin) / (self.M - 1)
        V = np.full(self.M, math.sqrt((self.bandmax - self.bandmin) / (self.M - 1)) * coupling)
        return eps, V


@dataclass
class ShenviGaussLegendre:
    """Gauss-Legendre discretisation of Shenvi, Roy, Tully (2009) in two halves around the Fermi level
    (docs/src/NQCModels/systembathmodels.md:236-262): with x_n, w_n the M/2-point Gauss-Legendre knots and weights,
        eps_n = (eF - a)/2 x_n + (a + eF)/2 , w~_n = (eF - a)/2 w_n     (lower half)
        eps_n = (b - eF)/2 x_n + (eF + b)/2 , w~_n = (b - eF)/2 w_n     (upper half)
    and V_n = sqrt(w~_n) V(eps).  The reference's iesh.md:85-105 example uses it for AdiabaticIESH; the engine only ever
    sees the resulting (eps_k, V_k)."""
    M: int
    bandmin: float
    bandmax: float
    fermi_level: float = 0.0

    def discretize(self, coupling: float):
        if self.M % 2:
            raise ValueError("ShenviGaussLegendre needs an even number of states")
        x, w = np.polynomial.legendre.leggauss(self.M // 2)
        a, b, eF = self.bandmin, self.bandmax, self.fermi_level
        eps = np.concatenate([0.5 * (eF - a) * x + 0.5 * (a + eF), 0.5 * (b - eF) * x + 0.5 * (eF + b)])
        wt = np.concatenate([0.5 * (eF - a) * w, 0.5 * (b - eF) * w])
        return eps, np.sqrt(wt) * coupling


@dataclass
class MiaoSubotnik:
    """U0 = 1/2 m w^2 x^2, U1 = 1/2 m w^2 (x-g)^2 + DeltaG, coupling sqrt(Gamma/2pi) (recalled)."""
    Γ: float = 6.4e-3
    m: float = 2000.0
    ω: float = 2e-4
    g: float = 20.6097
    ΔG: float = -3.8e-3


_EV = 1.0 / 27.211386245988          # hartree per eV
_ANG = 1.0 / 0.529177210903         # bohr per angstrom


@dataclass
class ErpenbeckThoss:
    """``ErpenbeckThoss(; Γ)`` (NQCModels, external; Erpenbeck & Thoss 2018): the impurity of the reference's own IESH tests
    and example (test/Dynamics/iesh.jl:23, iesh.md:85-105).  U0 = Morse(De, a, x0) + c, U1 = D1 e^{-2a'(x-x0')} - D2 e^{-a'(x-x0')}
    + V_inf, and a position-dependent coupling V_k(x) = Vbar_k [(1-q)/2 (1 - tanh((x - xt)/at)) + q], Vbar = sqrt(Γ/2π).
    Formula and defaults are RECALLED from NQCModels (not in the reference tree); they travel as explicit parameters."""
    Γ: float
    Dₑ: float = 3.52 * _EV
    a: float = 1.7361 / _ANG
    x0: float = 1.78 * _ANG
    c: float = -1.5 * _EV
    D1: float = 4.52 * _EV
    D2: float = 0.79 * _EV
    a1: float = 1.379 / _ANG
    x01: float = 1.78 * _ANG
    Vinf: float = -1.5 * _EV
    q: float = 0.05
    at: float = 0.5 * _ANG
    xt: float = 3.5 * _ANG


def AndersonHolstein(impurity, bath, fermi_level: float = 0.0) -> Model:
    """``AndersonHolstein(impurity_model, bath; fermi_level)``: ``impurity`` is a MiaoSubotnik or an ErpenbeckThoss, ``bath`` a
    TrapezoidalRule or a ShenviGaussLegendre."""
    eps, V = bath.discretize(math.sqrt(impurity.Γ / (2.0 * math.pi)))
    ne = int(np.count_nonzero(eps <= fermi_level))
    if isinstance(impurity, ErpenbeckThoss):
        i = impurity
        return Model(_abi.MODEL_ANDERSON_HOLSTEIN_ERPENBECK_THOSS, bath.M + 1,
                     (i.Dₑ, i.a, i.x0, i.c, i.D1, i.D2, i.a1, i.x01, i.Vinf, i.q, i.at, i.xt), bath_a=eps, bath_b=V,
                     nelectrons=ne, name="AndersonHolstein", fermi_level=fermi_level)
    return Model(_abi.MODEL_ANDERSON_HOLSTEIN_MIAO_SUBOTNIK, bath.M + 1,
                 (impurity.m, impurity.ω, impurity.g, impurity.ΔG), bath_a=eps, bath_b=V, nelectrons=ne,
                 name="AndersonHolstein", fermi_level=fermi_level)

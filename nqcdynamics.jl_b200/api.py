"""Host-side mirror of the reference's user interface for the ensemble hot path.

Same names, argument meaning and return shapes as NQCDynamics.jl (Julia cannot run in this image, so the host
side above the C ABI is Python; the Julia shim that binds the same ABI is in INTEGRATION.md):

    sim = Simulation[FSSH](Atoms(2000), TullyModelOne())             # Simulation{FSSH}(atoms, model)
    out = run_dynamics(sim, (0.0, 3000.0), distribution; output=..., trajectories=..., dt=1.0,
                       ensemble_algorithm=EnsembleB200(ngpus))        # src/Ensembles/run_dynamics.jl:42-136

Every trajectory is stepped on the GPU by ``libnqcb200.so``; there is no CPU path -- unsupported
(method, model, output) combinations raise.
"""
from __future__ import annotations

import math
import os
import threading
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _abi as A
from . import models as _models
from .distributed import shard_bounds
from .engine import Engine, device_count

# ---- atoms -------------------------------------------------------------------------------------
_AMU = 1822.888486209          # atomic mass unit in electron masses
_SYMBOL_MASS = {"H": 1.008 * _AMU, "D": 2.014 * _AMU, "C": 12.011 * _AMU, "N": 14.007 * _AMU, "O": 15.999 * _AMU}


class Atoms:
    """``Atoms(2000)`` = one atom of mass 2000 a.u.; ``Atoms([:H, :H])`` -> periodic-table masses
    (docs/src/atoms.md:16-21)."""

    def __init__(self, spec):
        if np.isscalar(spec) and not isinstance(spec, str):
            self.types, self.masses = ["X"], np.array([float(spec)])
        elif isinstance(spec, str):
            self.types, self.masses = [spec], np.array([_SYMBOL_MASS[spec]])
        else:
            spec = list(spec)
            if spec and isinstance(spec[0], str):
                self.types, self.masses = spec, np.array([_SYMBOL_MASS[s] for s in spec])
            else:
                self.types, self.masses = ["X"] * len(spec), np.asarray(spec, dtype=np.float64)

    def __len__(self):
        return len(self.masses)


# ---- dynamics methods (reference: DynamicsMethods.Method subtypes) -------------------------------
@dataclass
class FSSH:
    rescaling: str = "standard"          # :standard | :vinversion | :off  (fssh.jl:41)
    method_id: int = A.METHOD_FSSH


@dataclass
class Ehrenfest:
    method_id: int = A.METHOD_EHRENFEST


@dataclass
class Classical:
    method_id: int = A.METHOD_CLASSICAL


@dataclass
class AdiabaticIESH:
    rescaling: str = "standard"
    estimate_probability: bool = True
    disable_hopping: bool = False
    decoherence_C: float = 0.0           # DecoherenceCorrectionEDC(C) when > 0 (iesh.jl:72-74)
    method_id: int = A.METHOD_IESH


@dataclass
class EhrenfestNA:
    """``Simulation{EhrenfestNA}`` (ehrenfest_na.jl): independent-electron mean-field dynamics on the IESH state."""
    method_id: int = A.METHOD_EHRENFEST_NA


@dataclass
class ThermalLangevin:
    """``RingPolymerSimulation{ThermalLangevin}(atoms, model, n_beads; γ, temperature)`` (langevin.jl:67-82), BCOCB."""
    γ: float = 1.0
    method_id: int = A.METHOD_THERMAL_LANGEVIN


@dataclass
class NRPMD:
    γ: float = 0.5                        # nrpmd.jl:43
    method_id: int = A.METHOD_NRPMD


_RESCALE = {"standard": A.RESCALE_STANDARD, "vinversion": A.RESCALE_VINVERSION, "off": A.RESCALE_OFF}


class _Parametric(type):
    """``Simulation[FSSH](atoms, model, **kw)`` mirrors Julia's ``Simulation{FSSH}(atoms, model; kw...)``."""

    def __getitem__(cls, method_type):
        def construct(atoms, model, *args, **kw):
            method_kw = {k: kw.pop(k) for k in list(kw) if k in getattr(method_type, "__dataclass_fields__", {})}
            return cls(atoms, model, method_type(**method_kw), *args, **kw)
        return construct


class Simulation(metaclass=_Parametric):
    """simulations.jl:12-44.  ``size(sim) = (ndofs, natoms)``."""

    def __init__(self, atoms: Atoms, model: _models.Model, method=None, temperature: float = 0.0):
        self.atoms, self.model, self.method = atoms, model, method or Classical()
        self.temperature = float(temperature)
        if model.natoms is not None and model.natoms != len(atoms):
            raise ValueError(f"{model.name} expects {model.natoms} atoms, got {len(atoms)}")
        self.beads = 1

    @property
    def size(self) -> Tuple[int, ...]:
        return (self.model.ndofs, len(self.atoms))

    @property
    def ndofs_total(self) -> int:
        return self.model.ndofs * len(self.atoms)

    @property
    def dof_masses(self) -> np.ndarray:
        return np.repeat(self.atoms.masses, self.model.ndofs)


class RingPolymerSimulation(Simulation):
    """simulations.jl:46-73.  ``size(sim) = (ndofs, natoms, nbeads)``; omega_n = nbeads * temperature."""

    def __init__(self, atoms, model, method=None, n_beads: int = 1, temperature: float = 0.0):
        super().__init__(atoms, model, method, temperature)
        self.beads = int(n_beads)

    @property
    def size(self):
        return (self.model.ndofs, len(self.atoms), self.beads)


# ---- distributions (NQCDistributions.jl, external) ------------------------------------------------
@dataclass
class Normal:
    μ: float = 0.0
    σ: float = 1.0

    def sample(self, rng, shape):
        return rng.normal(self.μ, self.σ, shape)


@dataclass
class VelocityBoltzmann:
    """v ~ Normal(0, sqrt(T/m)) per atom (src/NQCDistributions-convenience.jl:16-42)."""
    temperature: float
    masses: Sequence[float]
    dims: Tuple[int, int]

    def sample(self, rng, shape):
        sd = np.sqrt(self.temperature / np.repeat(np.asarray(self.masses, dtype=float), self.dims[0]))
        return rng.standard_normal(shape) * sd


class Diabatic:
    pass


class Adiabatic:
    pass


@dataclass
class PureState:
    """``PureState(i)`` is diabatic by default; 1-based state index."""
    state: int
    statetype: Any = field(default_factory=Diabatic)

    def __mul__(self, other):
        return ProductDistribution(other, self)

    __rmul__ = __mul__


@dataclass
class MixedState:
    """``MixedState(populations, statetype)`` (NQCDistributions; docs/src/NQCDistributions/overview.md:141-155): the
    initial density matrix is ``Diagonal(populations)`` in the given basis (density_matrix_dynamics.jl:37-62); FSSH
    draws the active state with weights Re diag(sigma) (fssh.jl:53-54)."""
    populations: Sequence[float]
    statetype: Any = field(default_factory=Diabatic)

    def __mul__(self, other):
        return ProductDistribution(other, self)

    __rmul__ = __mul__


def _electronic_density(electronic, n: int) -> np.ndarray:
    """density_matrix(electronics, nstates) of NQCDistributions for PureState / MixedState (n x n, real diagonal)."""
    rho = np.zeros((n, n))
    if isinstance(electronic, MixedState):
        pops = np.asarray(electronic.populations, dtype=np.float64)
        if pops.shape != (n,):
            raise ValueError(f"MixedState needs one population per state ({n})")
        rho[np.arange(n), np.arange(n)] = pops
    elif isinstance(electronic, PureState):
        rho[electronic.state - 1, electronic.state - 1] = 1.0
    else:
        raise TypeError("FSSH / Ehrenfest take PureState or MixedState electronic distributions")
    return rho


@dataclass
class FermiDiracState:
    """``FermiDiracState(fermi_level, temperature)`` in the adiabatic basis (NQCDistributions; used by AdiabaticIESH,
    iesh.jl:99-128): occupations drawn by the reference's Metropolis walk over orbital swaps
    (``sample_fermi_dirac_distribution``, DynamicsUtils.jl:194-208, Boltzmann-factor variant)."""
    fermi_level: float = 0.0
    temperature: float = 0.0            # k_B T in hartree (atomic units, k_B = 1)
    statetype: Any = field(default_factory=Adiabatic)

    def __mul__(self, other):
        return ProductDistribution(other, self)

    __rmul__ = __mul__

    @property
    def β(self) -> float:
        return math.inf if self.temperature == 0.0 else 1.0 / self.temperature

    def sample_occupations(self, rng, energies: np.ndarray, nelectrons: int) -> np.ndarray:
        """1-based sorted occupied adiabatic orbitals for one trajectory."""
        nstates = len(energies)
        state = list(range(nelectrons))
        beta = self.β
        for _ in range(nstates * nelectrons):
            k = int(rng.integers(0, nelectrons))
            i = state[k]
            free = np.setdiff1d(np.arange(nstates), state, assume_unique=True)
            j = int(free[rng.integers(0, len(free))])
            de = energies[j] - energies[i]
            prob = (1.0 if de <= 0 else 0.0) if math.isinf(beta) else math.exp(min(700.0, -beta * de))
            if prob > rng.random():
                state[k] = j
        return np.sort(np.asarray(state, dtype=np.int32)) + 1

    def sample_diabatic(self, rng, H: np.ndarray, nelectrons: int):
        """``DynamicsVariables(sim, v, r, ::FermiDiracState{Diabatic})`` (iesh.jl:138-184): Fermi-Dirac occupations of
        the DIABATIC levels diag(H); electron i starts in its diabatic state d_i, i.e. with adiabatic coefficients
        U[d_i, :] (up to the sign of U[d_i, 1], as the reference normalises by sqrt of the first population), and its
        discrete adiabatic state is drawn with weights |U[d_i, :]|^2 without repetition.  U is taken in the engine's
        default gauge (U[k, k] >= 0: column-sign continuity against the identity), so psi means the same thing on the
        device.  Returns (psi (ne, n), sorted 1-based adiabatic occupations (ne,))."""
        n = H.shape[0]
        diab = self.sample_occupations(rng, np.diag(H).copy(), nelectrons) - 1
        _, U = np.linalg.eigh(H)
        U = U * np.where(np.diag(U) < 0.0, -1.0, 1.0)[None, :]
        psi = np.zeros((nelectrons, n))
        chosen: List[int] = []
        for i, d in enumerate(diab):
            row = U[d, :]
            pop = row * row
            while True:
                s = int(rng.choice(n, p=pop / pop.sum()))
                if s not in chosen:
                    chosen.append(s)
                    break
            psi[i] = row * (row[0] / abs(row[0]))       # adiabatic_density[:, 1] ./ sqrt(adiabatic_population[1])
        return psi, np.sort(np.asarray(chosen, dtype=np.int32)) + 1


@dataclass
class DynamicalDistribution:
    """``DynamicalDistribution(velocity, position, size)``: each entry a number, an array (broadcast over
    trajectories or indexed by trajectory on the leading axis) or a samplable distribution."""
    velocity: Any
    position: Any
    size: Tuple[int, ...]

    def __mul__(self, electronic):
        return ProductDistribution(self, electronic)

    __rmul__ = __mul__

    def _to_flat(self, x):
        """One configuration in the Julia layout (ndofs, natoms[, nbeads]) -> (B, D), D index = dof + ndofs*atom."""
        x = np.asarray(x, dtype=np.float64)
        D = int(np.prod(self.size[:2]))
        B = self.size[2] if len(self.size) > 2 else 1
        if x.ndim == 3:
            return x.transpose(2, 1, 0).reshape(B, D)
        if x.ndim == 2 and x.shape == tuple(self.size[:2]):
            return np.broadcast_to(x.T.reshape(1, D), (B, D))
        return np.broadcast_to(x.reshape(-1), (B, D)) if x.size == D else x.reshape(B, D)

    @staticmethod
    def _is_configurations(spec):
        return isinstance(spec, (list, tuple)) and len(spec) and np.ndim(spec[0]) >= 1

    def _draw(self, spec, rng, T, idx):
        D = int(np.prod(self.size[:2]))
        B = self.size[2] if len(self.size) > 2 else 1
        shape = (T, B, D)
        if hasattr(spec, "sample"):
            return np.ascontiguousarray(spec.sample(rng, shape))
        if self._is_configurations(spec):
            # a vector of configurations, indexed by the ONE index drawn per trajectory (see sample)
            return np.ascontiguousarray(np.stack([self._to_flat(spec[i]) for i in idx]))
        arr = np.asarray(spec, dtype=np.float64)
        if arr.ndim == 0:
            return np.full(shape, float(arr))
        return np.ascontiguousarray(np.broadcast_to(self._to_flat(arr), shape))

    def sample(self, rng, T, selection=None):
        """``distribution[j]`` for the selected indices (OrderedSelection, selections.jl:38-42) or ``rand(distribution)``
        (RandomSelection, :70-73).  Either way ONE index per trajectory picks the entry of every vector-of-configurations
        field, so paired (r_i, v_i) samples -- e.g. the output of a previous Langevin run -- stay paired."""
        lengths = {len(s) for s in (self.position, self.velocity) if self._is_configurations(s)}
        if len(lengths) > 1:
            raise ValueError("DynamicalDistribution: velocity and position sample vectors differ in length")
        idx = None
        if lengths:
            n = lengths.pop()
            idx = [j - 1 for j in selection] if selection is not None else rng.integers(0, n, T)
        return self._draw(self.position, rng, T, idx), self._draw(self.velocity, rng, T, idx)

    def _component_spec(self, spec):
        """Per-component (fixed | Normal) description of one entry for nqcb200_sample_state, or None if the entry is
        not of that form (a vector of configurations, an arbitrary sampler)."""
        D = int(np.prod(self.size[:2]))
        B = self.size[2] if len(self.size) > 2 else 1
        if isinstance(spec, Normal):
            return [(spec.μ, spec.σ)] * (B * D)
        if isinstance(spec, VelocityBoltzmann):
            sd = np.sqrt(spec.temperature / np.repeat(np.asarray(spec.masses, dtype=float), spec.dims[0]))
            return [(0.0, float(x)) for x in np.broadcast_to(sd, (B, D)).reshape(-1)]
        if hasattr(spec, "sample") or (isinstance(spec, (list, tuple)) and len(spec) and np.ndim(spec[0]) >= 1):
            return None
        arr = np.asarray(spec, dtype=np.float64)
        if arr.ndim == 0:
            return [float(arr)] * (B * D)
        return [float(x) for x in np.broadcast_to(self._to_flat(arr), (B, D)).reshape(-1)]

    def device_spec(self):
        """(r_spec, v_spec) for the device-side sampler, or None when host sampling is needed."""
        r, v = self._component_spec(self.position), self._component_spec(self.velocity)
        return None if r is None or v is None else (r, v)


@dataclass
class ProductDistribution:
    nuclear: DynamicalDistribution
    electronic: Any                      # PureState | FermiDiracState


# ---- outputs (src/DynamicsOutputs.jl, src/TimeCorrelationFunctions.jl) ----------------------------
@dataclass(frozen=True)
class _Output:
    name: str
    obs: int
    kind: str = "series"      # series | first | last | sum | hops | scatter | spring | centroid_ke | variables[_first|_last] | final_time
    deps: Tuple[int, ...] = ()   # device observables this output is built from (default: just `obs`)

    @property
    def needs(self) -> Tuple[int, ...]:
        return self.deps or (self.obs,)


OutputDiabaticPopulation = _Output("OutputDiabaticPopulation", A.OBS_DIABATIC_POP)
OutputAdiabaticPopulation = _Output("OutputAdiabaticPopulation", A.OBS_ADIABATIC_POP)
OutputKineticEnergy = _Output("OutputKineticEnergy", A.OBS_KINETIC)
OutputPotentialEnergy = _Output("OutputPotentialEnergy", A.OBS_POTENTIAL)
OutputTotalEnergy = _Output("OutputTotalEnergy", A.OBS_TOTAL_ENERGY)
OutputPosition = _Output("OutputPosition", A.OBS_POSITION)
OutputVelocity = _Output("OutputVelocity", A.OBS_VELOCITY)
OutputCentroidPosition = _Output("OutputCentroidPosition", A.OBS_POSITION)
OutputCentroidVelocity = _Output("OutputCentroidVelocity", A.OBS_VELOCITY)
OutputDiscreteState = _Output("OutputDiscreteState", A.OBS_DISCRETE_STATE)
OutputQuantumSubsystem = _Output("OutputQuantumSubsystem", A.OBS_SIGMA)
# build-defined (not in DynamicsOutputs.jl): all ne occupied orbitals of an AdiabaticIESH trajectory per frame.  The
# reference's OutputDiscreteState keeps only `first(u.state)` when the state is a vector (DynamicsOutputs.jl:178).
OutputOccupations = _Output("OutputOccupations", A.OBS_DISCRETE_STATE, "occupations")
OutputSurfaceHops = _Output("OutputSurfaceHops", A.OBS_DISCRETE_STATE, "hops")
# outputs assembled on the host from the same device streams (DynamicsOutputs.jl:101-141,192-227,387-397)
OutputFinalKineticEnergy = _Output("OutputFinalKineticEnergy", A.OBS_KINETIC, "last")
OutputFirstPosition = _Output("OutputFirstPosition", A.OBS_POSITION, "first")
OutputFirstVelocity = _Output("OutputFirstVelocity", A.OBS_VELOCITY, "first")
OutputFinalPosition = _Output("OutputFinalPosition", A.OBS_POSITION, "last")
OutputFinalVelocity = _Output("OutputFinalVelocity", A.OBS_VELOCITY, "last")
OutputTotalDiabaticPopulation = _Output("OutputTotalDiabaticPopulation", A.OBS_DIABATIC_POP, "sum")
OutputTotalAdiabaticPopulation = _Output("OutputTotalAdiabaticPopulation", A.OBS_ADIABATIC_POP, "sum")
OutputSpringEnergy = _Output("OutputSpringEnergy", A.OBS_TOTAL_ENERGY, "spring",
                             (A.OBS_TOTAL_ENERGY, A.OBS_KINETIC, A.OBS_POTENTIAL))
OutputCentroidKineticEnergy = _Output("OutputCentroidKineticEnergy", A.OBS_VELOCITY, "centroid_ke")
OutputFinalTime = _Output("OutputFinalTime", A.OBS_KINETIC, "final_time")
OutputMappingPosition = _Output("OutputMappingPosition", A.OBS_MAPPING_Q)      # DynamicsOutputs.jl:157 (NRPMD)
OutputMappingMomentum = _Output("OutputMappingMomentum", A.OBS_MAPPING_P)      # DynamicsOutputs.jl:165
_VARS = (A.OBS_POSITION, A.OBS_VELOCITY, A.OBS_SIGMA, A.OBS_DISCRETE_STATE)
OutputDynamicsVariables = _Output("OutputDynamicsVariables", A.OBS_POSITION, "variables", _VARS)
OutputInitial = _Output("OutputInitial", A.OBS_POSITION, "variables_first", _VARS)
OutputFinal = _Output("OutputFinal", A.OBS_POSITION, "variables_last", _VARS)


def _atom_indices(indices, natoms):
    """Julia-style atom selection: ``None`` / ``slice(None)`` = ``:``, otherwise 1-based indices."""
    if indices is None or (isinstance(indices, slice) and indices == slice(None)):
        return np.arange(natoms)
    idx = np.atleast_1d(np.asarray(indices, dtype=np.int64)) - 1
    if idx.min() < 0 or idx.max() >= natoms:
        raise IndexError("atom index out of range (indices are 1-based)")
    return idx


@dataclass(frozen=True)
class _SubsetOutput(_Output):
    indices: Any = None


def OutputSubsetKineticEnergy(indices):                           # DynamicsOutputs.jl:111-117
    return _SubsetOutput("OutputSubsetKineticEnergy", A.OBS_VELOCITY, "subset_ke", (), indices)


def OutputFinalSubsetKineticEnergy(indices):                      # DynamicsOutputs.jl:124-130
    return _SubsetOutput("OutputFinalSubsetKineticEnergy", A.OBS_VELOCITY, "subset_ke_last", (), indices)


def OutputKineticTemperature(indices=None):                       # DynamicsOutputs.jl:501-523 (kelvin)
    return _SubsetOutput("OutputKineticTemperature", A.OBS_VELOCITY, "kinetic_temperature", (), indices)


K_AU = 3.166811563455546e-06      # Boltzmann constant in hartree / K (UnitfulAtomic k_au)


def OutputStateResolvedScattering1D(sim, type="adiabatic"):      # DynamicsOutputs.jl:313-338
    if type not in ("adiabatic", "diabatic"):
        raise ValueError(f"{type} not recognised. Only `:diabatic` or `:adiabatic` accepted.")
    return _Output("OutputStateResolvedScattering1D",
                   A.OBS_SCATTERING if type == "adiabatic" else A.OBS_SCATTERING_DIABATIC, "scatter")


def PopulationCorrelationFunction(sim, statetype=None):          # TimeCorrelationFunctions.jl:61-88
    adiabatic = isinstance(statetype, Adiabatic) or statetype is Adiabatic
    return _Output("PopulationCorrelationFunction", A.OBS_POPCORR_ADIABATIC if adiabatic else A.OBS_POPCORR_DIABATIC)


# ---- callbacks (src/DynamicsUtils/callbacks.jl) ------------------------------------------------------
@dataclass(frozen=True)
class PositionOutside:
    """Termination predicate a device kernel can evaluate: ``(u, t, integrator) -> r[dof] < lo || r[dof] > hi || t > tcut`` with
    ``r = get_positions(u)`` flattened column-major and ``dof`` 1-based (the scattering examples of the reference
    documentation terminate when the particle has left the interaction region).  Ring polymers: ``r`` and the velocity of
    the ``outgoing`` clause are the CENTROID of that dof (a condition on ``get_centroid(get_positions(u))``)."""
    lo: float
    hi: float
    dof: int = 1
    outgoing: bool = False          # ... && the velocity points outwards (iesh.md:127-138: mean(r) > x && mean(v) > 0)
    tcut: float = float("inf")      # ... || t > tcut


class TerminatingCallback:
    """``TerminatingCallback(func) = DiscreteCallback(func, terminate!)`` (callbacks.jl:29).  ``func`` must be a
    :class:`PositionOutside` (arbitrary closures cannot run in the kernel).  Passed as ``callback=`` to
    :func:`run_dynamics`, it is evaluated after every step, after the method's own hopping callback."""

    def __init__(self, func):
        if not isinstance(func, PositionOutside):
            raise TypeError("EnsembleB200 evaluates the termination predicate in the step kernel: pass PositionOutside(lo, hi, dof)")
        self.func = func


# ---- reductions (src/Ensembles/reductions.jl) ------------------------------------------------------
class SortByTrajectoryReduction:
    pass


class SortByOutputReduction:
    pass


class SumReduction:
    pass


class MeanReduction:
    pass


class _NpzStore:
    """Group / dataset container with the layout of the reference's HDF5 files, used when h5py is not importable: one
    ``.npz`` archive whose member names are the HDF5 paths (``trajectory_<i>/<OutputName>``)."""

    def __init__(self, path):
        self.path, self.members = path, {}
        if os.path.exists(path):
            with np.load(path) as old:
                self.members = {k: old[k] for k in old.files}

    def create_dataset(self, group, name, value):
        key = f"{group}/{name}"
        if key in self.members:
            raise ValueError(f"{key} already exists in {self.path}")      # HDF5 "cw" mode: create_group fails on duplicates
        self.members[key] = np.asarray(value)

    def close(self):
        with open(self.path, "wb") as f:          # a file object keeps numpy from appending another extension
            np.savez(f, **self.members)


class _H5Store:
    def __init__(self, path, h5py):
        self.file = h5py.File(path, "a")

    def create_dataset(self, group, name, value):
        g = self.file.require_group(group)
        g.create_dataset(name, data=np.asarray(value))

    def close(self):
        self.file.close()


class FileReduction:
    """``FileReduction(filename)`` (reductions.jl:57-91): every trajectory's outputs go to the group ``trajectory_<id>`` of one
    file, one dataset per output (``Time`` included); the extension is forced to ``.h5`` unless it is ``.h5`` / ``.hdf5``.
    Values follow the reference's conversion: numbers and numeric arrays as they are; a vector of arrays is stacked on a
    NEW LAST axis (``reshape(reduce(hcat, value), size(value[1])..., :)``) -- numpy's leading frame axis is moved to the
    end so that the on-disk dataset has the reference's shape.  ``run_dynamics`` then returns the reference's message.
    Backend: h5py when importable (real HDF5); otherwise an ``.npz`` archive with the same group/dataset paths as member
    names, written next to the requested name (``<name>.h5.npz``) -- the image this package is built in has no HDF5 library."""

    def __init__(self, filename: str):
        root, ext = os.path.splitext(filename)
        self.filename = filename if ext in (".h5", ".hdf5") else root + ".h5"
        try:
            import h5py                      # noqa: F401
            self.backend = "h5py"
        except ImportError:
            self.backend = "npz"
        self.target = self.filename if self.backend == "h5py" else self.filename + ".npz"

    def _open(self):
        if self.backend == "h5py":
            import h5py
            return _H5Store(self.target, h5py)
        return _NpzStore(self.target)

    @staticmethod
    def _convert(value):
        if isinstance(value, dict):          # ComponentVector(reflection=..., transmission=...): flat vector, fields in order
            return np.concatenate([np.ravel(np.asarray(v, dtype=np.float64)) for v in value.values()])
        if isinstance(value, list):
            if value and isinstance(value[0], dict):
                raise TypeError("Cannot convert output type to HDF5 format")      # reductions.jl:82-84
            value = np.asarray(value)
        arr = np.asarray(value)
        if arr.dtype == object:
            raise TypeError("Cannot convert output type to HDF5 format")
        return np.moveaxis(arr, 0, -1) if arr.ndim >= 2 else arr      # frames last, as the reference stacks them

    def write(self, trajectories, first_id: int = 1):
        store = self._open()
        try:
            for i, traj in enumerate(trajectories):
                for key, value in traj.items():
                    store.create_dataset(f"trajectory_{first_id + i}", str(key), self._convert(value))
        finally:
            store.close()
        return f"Output written to {self.target}."


@dataclass
class EnsembleB200:
    """``ensemble_algorithm=EnsembleB200(ngpus)``: shard trajectories over ``ngpus`` B200s of this node.
    ``device_sampling=True``: initial conditions of the form (number | Normal | VelocityBoltzmann) x PureState are
    drawn on the device (nqcb200_sample_state; Philox stream keyed by the global trajectory index) instead of by
    numpy on the host -- nothing but the specification is uploaded."""
    ngpus: int = 1
    device_sampling: bool = False
    device_ids: Optional[Sequence[int]] = None      # CUDA ordinal of each shard (default 0 .. ngpus-1; may repeat)


# ---- run_dynamics ----------------------------------------------------------------------------------
def _trim_terminated(time: np.ndarray, arrs: Dict[int, np.ndarray], term_step: int, save_every: int, t_end: float):
    """Fixed-shape device stream of one trajectory -> the frames a DiffEq solution holds after ``terminate!``.

    sol.t of a terminated trajectory: the saveat points <= t_term, then the terminal state saved before and after the
    affect (DiscreteCallback ``save_positions = (true, true)``; the pre-affect copy is skipped when t_term is itself a
    saveat point).  The device stream carries the terminal state from save index ``term_step // save_every + 1`` on, so
    that index provides the appended frames.  The scattering observables are final-frame quantities and stay whole."""
    if term_step < 0:
        return time, arrs
    kf, rem = divmod(term_step, save_every)
    idx = list(range(kf + 1)) + ([kf + 1, kf + 1] if rem else [kf])
    ti = np.concatenate([time[:kf + 1], np.full(len(idx) - kf - 1, t_end)])
    keep_whole = (A.OBS_SCATTERING, A.OBS_SCATTERING_DIABATIC)
    return ti, {k: (a if k in keep_whole else a[idx]) for k, a in arrs.items()}


def sample_nrpmd_mapping(rng, T: int, nbeads: int, nstates: int, state: int, γ: float):
    """Initial NRPMD mapping variables (nrpmd.jl:47-65): theta ~ U[0, 2 pi) per state and bead, (q, p) = R (cos, sin) with
    R = sqrt(2 + 2 gamma) on the occupied (1-based) state and sqrt(2 gamma) elsewhere.  Returns (T, nbeads, nstates) arrays."""
    theta = rng.random((T, nbeads, nstates)) * 2.0 * np.pi
    radius = np.full(nstates, math.sqrt(2.0 * γ))
    radius[state - 1] = math.sqrt(2.0 + 2.0 * γ)
    return np.cos(theta) * radius, np.sin(theta) * radius


def _shape_series(sim, out: _Output, arr: np.ndarray):
    """arr: (nsave, width) -> the reference's per-frame value shape."""
    n = sim.model.nstates
    if out.obs in (A.OBS_POPCORR_DIABATIC, A.OBS_POPCORR_ADIABATIC):
        return arr.reshape(-1, n, n).transpose(0, 2, 1)          # column-major (i, j) per frame
    if out.obs == A.OBS_SIGMA:
        ncol = sim.model.nelectrons if sim.method.method_id in A.IESH_FAMILY else n      # psi is (n, ne) for IESH
        c = arr.reshape(-1, 2, ncol, n)
        return (c[:, 0] + 1j * c[:, 1]).transpose(0, 2, 1)
    if out.obs in (A.OBS_MAPPING_Q, A.OBS_MAPPING_P):
        return arr.reshape(-1, sim.beads, n).transpose(0, 2, 1)                          # (nsave, nstates, nbeads) per frame
    if out.obs in (A.OBS_POSITION, A.OBS_VELOCITY):
        return arr.reshape((-1,) + tuple(reversed(sim.size[:2]))).transpose(0, 2, 1)   # (nsave, ndofs, natoms)
    if arr.shape[1] == 1:
        return arr[:, 0]
    return arr


def _finalise(sim, out: _Output, arrs: Dict[int, np.ndarray], per_trajectory: bool, t_final: float = 0.0):
    """Engine arrays -> reference output value.  arrs[obs]: (nsave, width) (reduced) or that of one trajectory."""
    arr = arrs[out.obs]
    if out.kind == "scatter":
        last = arr[-1]
        n = sim.model.nstates
        return {"reflection": last[:n].copy(), "transmission": last[n:].copy()}
    if out.kind == "hops":
        if not per_trajectory:
            raise ValueError("OutputSurfaceHops needs per-trajectory states (use SortByTrajectoryReduction)")
        st = np.rint(arr).astype(np.int64)
        if sim.method.method_id == A.METHOD_IESH:
            st = np.sort(st, axis=1)        # the occupation vector is a set (an accepted hop is not re-sorted, iesh.jl:399-407)
        return int(np.count_nonzero(np.any(st[1:] != st[:-1], axis=1)))
    if out.kind == "final_time":
        return t_final
    if out.kind == "spring":                # classical_hamiltonian = kinetic + potential + spring (DynamicsUtils.jl:108-151)
        return (arrs[A.OBS_TOTAL_ENERGY] - arrs[A.OBS_KINETIC] - arrs[A.OBS_POTENTIAL])[:, 0]
    if out.kind == "centroid_ke":           # DynamicsOutputs.jl:50-59: sum_i m_i v_centroid,i^2 / 2
        return 0.5 * np.sum(sim.dof_masses * arr * arr, axis=1)
    if out.kind in ("subset_ke", "subset_ke_last", "kinetic_temperature"):
        # classical_kinetic_energy(masses[indices], v[:, indices]) per saved frame (DynamicsUtils.jl:108-135)
        if not per_trajectory:
            raise ValueError(f"{out.name} is quadratic in the velocity stream: use a per-trajectory reduction")
        if sim.beads > 1:
            raise ValueError(f"{out.name}: bead-resolved velocities are not streamed for ring polymers")
        natoms = len(sim.atoms)
        idx = _atom_indices(out.indices, natoms)
        v = _shape_series(sim, OutputVelocity, arr)[:, :, idx]                      # (nsave, ndofs, subset)
        ke = 0.5 * np.einsum("a,kda->k", np.asarray(sim.atoms.masses, dtype=np.float64)[idx], v * v)
        if out.kind == "subset_ke_last":
            return float(ke[-1])
        if out.kind == "kinetic_temperature":
            return 2.0 * ke / K_AU / sim.size[0] / len(idx)
        return ke
    if out.kind.startswith("variables"):
        if not per_trajectory:
            raise ValueError(f"{out.name} is a per-trajectory output (use SortByTrajectoryReduction)")
        if sim.beads > 1:
            raise ValueError(f"{out.name}: bead-resolved frames are not streamed for ring polymers (centroids only)")
        method = sim.method.method_id
        r = _shape_series(sim, OutputPosition, arrs[A.OBS_POSITION])
        v = _shape_series(sim, OutputVelocity, arrs[A.OBS_VELOCITY])
        frames = []
        for k in range(r.shape[0]):
            u = {"v": v[k], "r": r[k]}
            if method in (A.METHOD_FSSH, A.METHOD_EHRENFEST) + A.IESH_FAMILY:
                sig = _shape_series(sim, OutputQuantumSubsystem, arrs[A.OBS_SIGMA][k:k + 1])[0]
                u["σreal"], u["σimag"] = sig.real.copy(), sig.imag.copy()
            if method in (A.METHOD_FSSH, A.METHOD_IESH):
                u["state"] = np.rint(arrs[A.OBS_DISCRETE_STATE][k]).astype(np.int64)
            frames.append(u)
        return frames[0] if out.kind == "variables_first" else frames[-1] if out.kind == "variables_last" else frames
    val = _shape_series(sim, out, arr)
    if out.kind == "occupations":
        return np.rint(val).astype(np.int64).reshape(arr.shape[0], -1) if per_trajectory else val
    if out.obs == A.OBS_DISCRETE_STATE:
        if val.ndim > 1:                     # IESH: `length(u.state) > 1 ? round(Int, first(u.state))` (DynamicsOutputs.jl:178)
            val = val[:, 0]
        if per_trajectory:                   # reduced (Sum / Mean) series stay real-valued
            val = np.rint(val).astype(np.int64)
    if out.kind == "first":
        return val[0]
    if out.kind == "last":
        return val[-1]
    if out.kind == "sum":
        return val.sum(axis=1)
    return val


def run_dynamics(sim: Simulation, tspan, distribution, *, output, selection: Optional[Sequence[int]] = None,
                 reduction=None, ensemble_algorithm: Optional[EnsembleB200] = None, trajectories: int = 1,
                 dt: float = 1.0, saveat: Optional[float] = None, savetime: bool = True, seed: Optional[int] = None,
                 draws: Optional[np.ndarray] = None, callback: Optional[TerminatingCallback] = None, **kwargs):
    """Run ``trajectories`` trajectories over ``tspan`` sampling ``distribution`` (run_dynamics.jl:42-136).

    Keywords follow the reference; ``saveat`` must be a multiple of ``dt`` (the engine's integrators are
    fixed-step, like the reference's custom algorithms).  Extra: ``seed`` (Philox key / IC sampling) and
    ``draws`` (parity mode: uniform hop draws of shape (nsteps, trajectories)).  The reference's
    ``precompile_dynamics`` pass is skipped (nothing to JIT).

    ``callback=TerminatingCallback(PositionOutside(lo, hi, dof))``: a terminated trajectory's series end the way a
    DiffEq solution does -- the saveat points up to the termination time, then the terminal state saved before and
    after ``terminate!`` (``save_positions = (true, true)``, the DiscreteCallback default: the terminal time appears
    twice, once if it coincides with a saveat point plus the post-affect copy).  That rule restates DiffEqBase's
    ``apply_discrete_callback!`` (a dependency outside the reference tree; unpinned).  ``OutputFinal*``,
    ``OutputStateResolvedScattering1D`` and ``OutputFinalTime`` see the terminal state / time.  Reduced (Sum / Mean)
    series keep their fixed shape: a terminated trajectory contributes its terminal state to the later save points
    (the reference cannot add ragged series at all)."""
    kwargs.pop("precompile_dynamics", None)
    if kwargs:
        raise TypeError(f"unsupported keyword(s) for the B200 ensemble path: {sorted(kwargs)}")
    reduction = reduction or SortByTrajectoryReduction()
    alg = ensemble_algorithm or EnsembleB200(1)
    outputs = output if isinstance(output, (tuple, list)) else (output,)
    for o in outputs:
        if not isinstance(o, _Output):
            raise TypeError("EnsembleB200 evaluates outputs on the device: pass Output* objects of this package "
                            "(arbitrary f(sol, i) closures are not supported on this path)")
    T = int(trajectories)
    t0, t1 = float(tspan[0]), float(tspan[1])
    nsteps = int(round((t1 - t0) / dt))
    save_every = 1 if saveat is None else int(round(float(saveat) / dt))
    if save_every < 1 or (saveat is not None and abs(save_every * dt - float(saveat)) > 1e-9 * max(1.0, abs(float(saveat)))):
        raise ValueError("saveat must be a positive multiple of dt")
    nsave = nsteps // save_every + 1
    if callback is not None:
        if not isinstance(callback, TerminatingCallback):
            raise TypeError("callback must be a TerminatingCallback(PositionOutside(...))")
        if nsteps % save_every:
            raise ValueError("with a TerminatingCallback the time span must be a whole number of saveat intervals")
        if not 1 <= callback.func.dof <= sim.ndofs_total:
            raise ValueError("PositionOutside.dof out of range")
    per_traj = isinstance(reduction, (SortByTrajectoryReduction, SortByOutputReduction, FileReduction))
    obs_mask = 0
    method_id = sim.method.method_id
    for o in outputs:
        for d in o.needs:
            if d == A.OBS_SIGMA and method_id not in (A.METHOD_FSSH, A.METHOD_EHRENFEST) + A.IESH_FAMILY:
                continue
            if d == A.OBS_DISCRETE_STATE and method_id not in (A.METHOD_FSSH, A.METHOD_IESH):
                continue
            obs_mask |= 1 << d
    obs_ids = [d for d in range(A.OBS_COUNT) if (obs_mask >> d) & 1]

    method, model = sim.method, sim.model
    rng = np.random.default_rng(seed)
    if isinstance(distribution, ProductDistribution):
        nuclear, electronic = distribution.nuclear, distribution.electronic
    else:
        nuclear, electronic = distribution, None
    density = method.method_id in (A.METHOD_FSSH, A.METHOD_EHRENFEST)
    dev_spec = None
    iesh_family = method.method_id in A.IESH_FAMILY
    device_sampling = bool(getattr(alg, "device_sampling", False))
    if device_sampling and not iesh_family:      # AdiabaticIESH: two nuclear numbers per trajectory stay on the host, the
        dev_spec = nuclear.device_spec() if selection is None else None      # occupations are drawn on the device (below)
        if dev_spec is None:
            raise ValueError("device_sampling needs number / Normal / VelocityBoltzmann entries and no selection")
        if (density and method.method_id == A.METHOD_FSSH and isinstance(electronic, MixedState)
                and (isinstance(electronic.statetype, Adiabatic) or electronic.statetype is Adiabatic)):
            raise ValueError("device_sampling: an adiabatic MixedState needs the active state drawn per trajectory on the "
                             "host (fssh.jl:53-54); use device_sampling=False or a diabatic MixedState")
        r = v = None
    else:
        r, v = nuclear.sample(rng, T, selection)
    if density and electronic is None:
        raise ValueError("FSSH / Ehrenfest need an electronic distribution: nuclear * PureState(i)")
    iesh = method.method_id in A.IESH_FAMILY
    mean_field = method.method_id == A.METHOD_EHRENFEST_NA
    psi0 = occ0 = None
    if iesh:
        # DynamicsVariables(sim, v, r) -> the ne lowest adiabatic orbitals (iesh.jl:89-97);
        # DynamicsVariables(sim, v, r, FermiDiracState) -> sampled occupations (iesh.jl:99-128)
        n, ne = model.nstates, model.nelectrons
        if electronic is None:
            occ0 = np.tile(np.arange(1, ne + 1, dtype=np.int32), (T, 1))
        elif isinstance(electronic, FermiDiracState):
            if abs(electronic.fermi_level - getattr(model, "fermi_level", 0.0)) > 1e-12:
                raise ValueError("Fermi level of model and distribution do not match")          # iesh.jl:105-112
            occ0 = np.empty((T, ne), dtype=np.int32)
            diabatic_fd = isinstance(electronic.statetype, Diabatic) or electronic.statetype is Diabatic
            fd_on_device = device_sampling and not diabatic_fd and not mean_field
            if diabatic_fd:
                if mean_field:
                    raise TypeError("EhrenfestNA: FermiDiracState{Diabatic} is defined for AdiabaticIESH only (iesh.jl:138)")
                psi0 = np.zeros((T, ne, n))
            for t in range(T):
                # ring polymers: the centroid's eigenvalues / eigenvectors (get_centroid_eigen, test/Dynamics/rpiesh.jl:52)
                r0 = r[t].reshape(sim.beads, -1).mean(axis=0) if r is not None else None
                if diabatic_fd:
                    psi0[t], occ0[t] = electronic.sample_diabatic(rng, model.diabatic_hamiltonian(r0), ne)
                elif fd_on_device:       # nqcb200_sample_occupations draws them from the device's own eigenvalues at r0
                    occ0[t] = np.arange(1, ne + 1)
                else:
                    occ0[t] = electronic.sample_occupations(rng, model.adiabatic_energies(r0), ne)
        else:
            raise TypeError("AdiabaticIESH / EhrenfestNA take no electronic distribution (ground state) or a FermiDiracState")
        # psi0 stays None for the adiabatic cases: electron e starts in orbital occ0[e], built on the device from the
        # occupations alone (nqcb200_set_state with sig_re == NULL)

    fd_beta = None
    if iesh and isinstance(electronic, FermiDiracState) and device_sampling and not mean_field and not (
            isinstance(electronic.statetype, Diabatic) or electronic.statetype is Diabatic):
        fd_beta = electronic.β
    qmap0 = pmap0 = None
    if method.method_id == A.METHOD_NRPMD:
        # DynamicsVariables(sim::RingPolymerSimulation{<:NRPMD}, v, r, ::PureState{Diabatic}) (nrpmd.jl:47-65): one random
        # angle per state and bead; radius sqrt(2 + 2 gamma) on the occupied state, sqrt(2 gamma) on the others
        if not isinstance(electronic, PureState) or isinstance(electronic.statetype, Adiabatic) or electronic.statetype is Adiabatic:
            raise TypeError("NRPMD takes nuclear * PureState(i, Diabatic())")
        if not device_sampling:
            qmap0, pmap0 = sample_nrpmd_mapping(rng, T, sim.beads, model.nstates, electronic.state, float(method.γ))

    ngpus = max(1, int(alg.ngpus))
    device_ids = list(range(ngpus)) if alg.device_ids is None else [int(d) for d in alg.device_ids]
    if len(device_ids) != ngpus:
        raise ValueError("EnsembleB200: device_ids needs one CUDA ordinal per shard")
    if max(device_ids) >= device_count() or min(device_ids) < 0:
        raise RuntimeError(f"EnsembleB200(device_ids={device_ids}) but only {device_count()} CUDA device(s) visible")
    bounds = [shard_bounds(T, ngpus, g) for g in range(ngpus)]      # the one sharding rule (distributed.py)
    engine_seed = int(rng.integers(0, 2 ** 63 - 1)) if seed is None else int(seed)
    results: List[Any] = [None] * ngpus
    errors: List[BaseException] = []

    def shard(g):
        try:
            lo, hi = bounds[g]
            Tg = hi - lo
            cfg, keep = A.make_config(
                method=method.method_id, model=model.kind, nstates=model.nstates, ndofs=sim.ndofs_total,
                masses=sim.dof_masses, ntraj=Tg, dt=dt, nbeads=sim.beads, nelectrons=model.nelectrons,
                params=model.params, bath_a=model.bath_a, bath_b=model.bath_b,
                rescaling=_RESCALE[getattr(method, "rescaling", "standard")],
                estimate_probability=int(getattr(method, "estimate_probability", True)),
                disable_hopping=int(getattr(method, "disable_hopping", False)),
                rng=A.RNG_INJECTED if draws is not None else A.RNG_PHILOX, device=device_ids[g], save_every=save_every,
                nsave=nsave, per_trajectory=int(per_traj), observables=obs_mask, traj_offset=lo, seed=engine_seed,
                t0=t0, temperature=sim.temperature, nrpmd_gamma=getattr(method, "γ", 0.5),
                edc_C=getattr(method, "decoherence_C", 0.0))
            with Engine(cfg, keep) as eng:
                ran = False
                if callback is not None:
                    eng.set_termination(callback.func.dof - 1, callback.func.lo, callback.func.hi, callback.func.outgoing, callback.func.tcut)
                if dev_spec is not None:
                    rho1 = None
                    adiabatic = True
                    if density:
                        n = model.nstates
                        rho1 = _electronic_density(electronic, n)
                        adiabatic = isinstance(electronic.statetype, Adiabatic) or electronic.statetype is Adiabatic
                    st = electronic.state if (adiabatic and method.method_id == A.METHOD_FSSH and isinstance(electronic, PureState)) else 0
                    eng.sample_state(dev_spec[0], dev_spec[1], rho1, diabatic=not adiabatic, state=st)
                    if method.method_id == A.METHOD_NRPMD:
                        eng.sample_mapping(electronic.state)
                    rg = vg = None
                else:
                    rg, vg = r[lo:hi], v[lo:hi]
                if dev_spec is not None:
                    pass
                elif density:
                    n = model.nstates
                    rho = np.tile(_electronic_density(electronic, n), (Tg, 1, 1))
                    adiabatic = isinstance(electronic.statetype, Adiabatic) or electronic.statetype is Adiabatic
                    state = None
                    if adiabatic and method.method_id == A.METHOD_FSSH:
                        if isinstance(electronic, PureState):
                            state = np.full(Tg, electronic.state, dtype=np.int32)
                        else:     # sample(Weights(diag(sigma))) per trajectory (fssh.jl:53-54), keyed by the global index
                            w = np.diag(rho[0]) / np.trace(rho[0])
                            state = np.array([np.random.default_rng([engine_seed, lo + i]).choice(n, p=w) + 1 for i in range(Tg)], dtype=np.int32)
                    if draws is None:
                        # one call per batch (nqcb200_run_from_host): kernels with a launch-fused initialisation read
                        # r, v in place; everything else behaves like set_state[_diabatic] + run
                        eng.run_from_host(rg, vg, rho, None, state, None, diabatic=not adiabatic, nsteps=nsteps)
                        ran = True
                    elif adiabatic:
                        eng.set_state(rg, vg, rho, None, state)
                    else:
                        eng.set_state_diabatic(rg, vg, rho)
                elif iesh:
                    eng.set_state(rg, vg, None if psi0 is None else psi0[lo:hi], None, None if (mean_field and psi0 is not None) else occ0[lo:hi])
                    if fd_beta is not None:
                        eng.sample_occupations(fd_beta)
                elif method.method_id == A.METHOD_NRPMD:
                    eng.set_state(rg, vg)
                    eng.set_mapping(qmap0[lo:hi], pmap0[lo:hi])
                else:
                    eng.set_state(rg, vg)
                if draws is not None:
                    eng.set_draws(np.ascontiguousarray(draws[:, lo:hi]))
                if not ran:
                    eng.run(nsteps)
                results[g] = {d: (eng.observable_per_trajectory(d) if per_traj else eng.observable_sum(d)) for d in obs_ids}
                if callback is not None:
                    results[g]["term"] = eng.termination()
        except BaseException as exc:   # re-raised on the caller's thread
            errors.append(exc)

    if ngpus == 1:
        shard(0)
    else:
        threads = [threading.Thread(target=shard, args=(g,)) for g in range(ngpus)]
        for th in threads: th.start()
        for th in threads: th.join()
    if errors:
        raise errors[0]

    time = t0 + dt * save_every * np.arange(nsave)
    term = np.concatenate([res["term"] for res in results]) if callback is not None else np.full(T, -1, dtype=np.int64)
    t_end = np.where(term >= 0, t0 + dt * term, time[-1])      # OutputFinalTime: last(sol.t)
    if per_traj:
        per_obs = {k: np.concatenate([res[k] for res in results], axis=0) for k in obs_ids}    # (T, nsave, w)
        trajs = []
        for i in range(T):
            ti, arrs = _trim_terminated(time, {k: per_obs[k][i] for k in obs_ids}, int(term[i]), save_every, float(t_end[i]))
            d: Dict[str, Any] = {"Time": ti.copy()} if savetime else {}
            for o in outputs:
                d[o.name] = _finalise(sim, o, arrs, True, float(t_end[i]))
            trajs.append(d)
        if isinstance(reduction, FileReduction):
            return reduction.write(trajs)
        if isinstance(reduction, SortByOutputReduction):
            keys = list(trajs[0].keys())
            return {k: [tr[k] for tr in trajs] for k in keys}
        return trajs[0] if T == 1 else trajs
    scale = 1.0 / T if isinstance(reduction, MeanReduction) else 1.0
    summed = {k: sum(res[k] for res in results) * scale for k in obs_ids}
    d = {"Time": time * (T * scale)} if savetime else {}     # `:Time` is reduced too (test/Ensembles/ensembles.jl:21)
    for o in outputs:
        if o.kind == "centroid_ke":
            raise ValueError("OutputCentroidKineticEnergy is not linear in the stream: use a per-trajectory reduction")
        if o.kind in ("subset_ke", "subset_ke_last", "kinetic_temperature"):
            raise ValueError(f"{o.name} is quadratic in the velocity stream: use a per-trajectory reduction")
        d[o.name] = _finalise(sim, o, summed, False, float(t_end.sum()) * scale)
    return d

"""Shared helpers for the parity tests: build the same config for the CUDA engine and the CPU oracle."""
import numpy as np

import nqcdynamics_jl_b200 as nq

A = nq._abi

ALL_POP_OBS = ((1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_POPCORR_DIABATIC) |
               (1 << A.OBS_POPCORR_ADIABATIC) | (1 << A.OBS_KINETIC) | (1 << A.OBS_POTENTIAL) |
               (1 << A.OBS_TOTAL_ENERGY) | (1 << A.OBS_POSITION) | (1 << A.OBS_VELOCITY) |
               (1 << A.OBS_SCATTERING) | (1 << A.OBS_SCATTERING_DIABATIC) | (1 << A.OBS_SIGMA))
CLASSICAL_OBS = ((1 << A.OBS_KINETIC) | (1 << A.OBS_POTENTIAL) | (1 << A.OBS_TOTAL_ENERGY) |
                 (1 << A.OBS_POSITION) | (1 << A.OBS_VELOCITY))


def model_config(model, **kw):
    """kwargs for make_config from a nq.models.Model."""
    natoms = model.natoms or kw.pop("natoms", 1)
    D = model.ndofs * natoms
    out = dict(model=model.kind, nstates=model.nstates, ndofs=D, params=model.params,
               bath_a=model.bath_a, bath_b=model.bath_b, nelectrons=model.nelectrons)
    out.update(kw)
    return out


def make_pair(make_engine, make_oracle, **cfg_kw):
    cfg_e, keep_e = A.make_config(**cfg_kw)
    cfg_o, keep_o = A.make_config(**cfg_kw)
    return make_engine(cfg_e, keep_e), make_oracle(cfg_o, keep_o)


def engine_factory():
    from nqcdynamics_jl_b200.engine import Engine
    return lambda cfg, keep: Engine(cfg, keep)


def oracle_factory():
    import oracle
    return lambda cfg, keep: oracle.OracleEngine(cfg, keep)


def rel_err(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), floor)
    return float(np.max(np.abs(a - b)) / scale)


def gauge_align(Ze, Zo):
    """Column signs s_i = sign(<Ze_i, Zo_i>) relating two eigenvector sets (T, n, n)."""
    return np.sign(np.einsum("tki,tki->ti", Ze, Zo))

"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/oracle_core.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_here))
import nqcdynamics_jl_b200 as _pkg  # noqa: E402  (ABI struct definitions only)

_abi = _pkg._abi
_LIB = None
_dp = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(_here, "libnqcd_oracle.so")
    srcs = [os.path.join(_here, f) for f in ("oracle_capi.cpp", "oracle_dynamics.hpp", "oracle_core.hpp")]
    srcs.append(os.path.join(_here, "..", "include", "nqcb200.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _here, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        _abi.bind(L, "nqco_")
        L.nqco_set_num_threads.argtypes, L.nqco_set_num_threads.restype = [C.c_int], C.c_int
        L.nqco_normal_mode_matrix.argtypes = [C.c_int, _dp]
        L.nqco_cayley.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, _dp]
        L.nqco_sym_eigh.argtypes = [C.c_int, _dp, _dp, _dp, C.c_int]
        L.nqco_herm_eigh.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.nqco_complex_det.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        L.nqco_evaluate_model.argtypes = [C.POINTER(_abi.Config), _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.nqco_propagate_density.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp, _dp, C.c_double, C.c_double,
                                             C.c_double, _dp, _dp]
        L.nqco_philox_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
        L.nqco_philox_uniform.restype = C.c_double
        L.nqco_select_new_state.argtypes = [C.c_int, _dp, C.c_int, C.c_double]
        L.nqco_unit_rescale.argtypes = [C.POINTER(_abi.Config), _dp, _dp, C.c_int, C.c_int, _dp]
        _ip = C.POINTER(C.c_int32)
        L.nqco_unoccupied.argtypes = [C.c_int, C.c_int, _ip, _ip]
        L.nqco_edc.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_double, _dp, C.c_double, C.c_double]
        _up = C.POINTER(C.c_uint32)
        L.nqco_philox_raw.argtypes = [_up, _up, _up]
        _LIB = L
    return _LIB


def set_num_threads(n):
    return lib().nqco_set_num_threads(int(n))


class OracleEngine(_abi.CHandle):
    """Same call surface as the CUDA engine's handle, executed by the CPU restatement."""

    def __init__(self, cfg, keepalive=()):
        super().__init__(lib(), "nqco_", cfg, keepalive)


def _p(a):
    return a.ctypes.data_as(_dp)


def normal_mode_matrix(B):
    U = np.empty((B, B))
    lib().nqco_normal_mode_matrix(B, _p(U))
    return U.T.copy()  # column-major U[j + B*k] -> U[j, k]


def cayley(B, omega_n, dt, half):
    out = np.empty((B, 2, 2))
    lib().nqco_cayley(B, omega_n, dt, int(half), _p(out))
    return out  # [k] = [[c11, c12], [c21, c22]]


def sym_eigh(A, algo=0):
    A = np.asfortranarray(A, dtype=np.float64)
    n = A.shape[0]
    w = np.empty(n); Z = np.empty((n, n), order="F")
    rc = lib().nqco_sym_eigh(n, _p(A), _p(w), _p(Z), algo)
    assert rc == 0
    return w, np.array(Z)


def herm_eigh(A):
    A = np.asarray(A, dtype=np.complex128)
    n = A.shape[0]
    Are, Aim = np.asfortranarray(A.real), np.asfortranarray(A.imag)
    w = np.empty(n); Zre = np.empty((n, n), order="F"); Zim = np.empty((n, n), order="F")
    lib().nqco_herm_eigh(n, _p(Are), _p(Aim), _p(w), _p(Zre), _p(Zim))
    return w, np.array(Zre) + 1j * np.array(Zim)


def complex_det(A):
    A = np.asarray(A, dtype=np.complex128)
    n = A.shape[0]
    Are, Aim = np.asfortranarray(A.real), np.asfortranarray(A.imag)
    re, im = C.c_double(), C.c_double()
    lib().nqco_complex_det(n, _p(Are), _p(Aim), C.byref(re), C.byref(im))
    return complex(re.value, im.value)


def evaluate_model(cfg, r):
    n, D = cfg.nstates, cfg.ndofs
    r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1)
    V = np.empty((n, n), order="F"); dV = np.empty((D, n, n)); w = np.empty(n); Z = np.empty((n, n), order="F")
    ad = np.empty((D, n, n)); nac = np.empty((D, n, n))
    rc = lib().nqco_evaluate_model(C.byref(cfg), _p(r), _p(V), _p(dV), _p(w), _p(Z), _p(ad), _p(nac))
    assert rc == 0
    tr = lambda x: x.transpose(0, 2, 1).copy()
    return dict(V=np.array(V), dV=tr(dV), w=w, Z=np.array(Z), adiab=tr(ad), nac=tr(nac))


def propagate_density(E0, vd0, t0, E1, vd1, t1, t, dt, sigma):
    n = len(E0)
    f = lambda x: np.asfortranarray(np.asarray(x, dtype=np.float64))
    E0, E1, vd0, vd1 = map(f, (E0, E1, vd0, vd1))
    sre = np.asfortranarray(np.real(sigma).astype(np.float64)); sim = np.asfortranarray(np.imag(sigma).astype(np.float64))
    lib().nqco_propagate_density(n, _p(E0), _p(vd0), t0, _p(E1), _p(vd1), t1, t, dt, _p(sre), _p(sim))
    return np.array(sre) + 1j * np.array(sim)


def philox_uniform(seed, gid, step, purpose=0):
    return lib().nqco_philox_uniform(seed, gid, step, purpose)


def select_new_state(cumprob, state, xi):
    p = np.ascontiguousarray(cumprob, dtype=np.float64)
    return lib().nqco_select_new_state(len(p), _p(p), int(state), float(xi))


def unit_rescale(cfg, r, v, new_state, old_state):
    r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1)
    v = np.array(v, dtype=np.float64).reshape(-1)
    eig = np.empty(cfg.nstates)
    ok = lib().nqco_unit_rescale(C.byref(cfg), _p(r), _p(v), new_state, old_state, _p(eig))
    assert ok >= 0
    return bool(ok), v, eig


def unoccupied(n, occ):
    occ = np.ascontiguousarray(occ, dtype=np.int32)
    out = np.empty(n - len(occ), dtype=np.int32)
    ip = C.POINTER(C.c_int32)
    lib().nqco_unoccupied(n, len(occ), occ.ctypes.data_as(ip), out.ctypes.data_as(ip))
    return out


def edc(psi, occupied, dt, E, Ekin, Cc=0.1):
    re = np.ascontiguousarray(np.real(psi), dtype=np.float64).copy(); im = np.ascontiguousarray(np.imag(psi), dtype=np.float64).copy()
    E = np.ascontiguousarray(E, dtype=np.float64)
    lib().nqco_edc(len(re), _p(re), _p(im), occupied, dt, _p(E), Ekin, Cc)
    return re + 1j * im


def philox_raw(ctr, key):
    up = C.POINTER(C.c_uint32)
    c = np.array(ctr, dtype=np.uint32); k = np.array(key, dtype=np.uint32); o = np.empty(4, dtype=np.uint32)
    lib().nqco_philox_raw(c.ctypes.data_as(up), k.ctypes.data_as(up), o.ctypes.data_as(up))
    return o

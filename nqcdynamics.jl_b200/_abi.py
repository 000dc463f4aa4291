"""ctypes mirror of ``include/nqcb200.h`` (the C ABI of the B200 ensemble-trajectory engine).

This module only describes the ABI (config struct, enums, argument types) and wraps an opaque
handle; it contains no numerics.  ``load_engine_library`` loads the CUDA shared library built
from ``csrc/`` and raises if it is missing -- there is no CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

ABI_VERSION = 1
MAX_PARAMS = 32

# enums (include/nqcb200.h)
METHOD_FSSH, METHOD_EHRENFEST, METHOD_IESH, METHOD_CLASSICAL, METHOD_NRPMD, METHOD_EHRENFEST_NA = 1, 2, 3, 4, 5, 6
METHOD_THERMAL_LANGEVIN = 7
IESH_FAMILY = (METHOD_IESH, METHOD_EHRENFEST_NA)      # psi: n x ne, trajectory-major
(MODEL_TULLY_ONE, MODEL_TULLY_TWO, MODEL_TULLY_THREE, MODEL_DOUBLE_WELL, MODEL_SPIN_BOSON,
 MODEL_THREE_STATE_MORSE, MODEL_HARMONIC, MODEL_FREE, MODEL_ANDERSON_HOLSTEIN_MIAO_SUBOTNIK,
 MODEL_ANDERSON_HOLSTEIN_ERPENBECK_THOSS) = range(1, 11)
ANDERSON_HOLSTEIN_FAMILY = (MODEL_ANDERSON_HOLSTEIN_MIAO_SUBOTNIK, MODEL_ANDERSON_HOLSTEIN_ERPENBECK_THOSS)
RESCALE_STANDARD, RESCALE_VINVERSION, RESCALE_OFF = 0, 1, 2
RNG_PHILOX, RNG_INJECTED = 0, 1
(OBS_ADIABATIC_POP, OBS_DIABATIC_POP, OBS_POPCORR_DIABATIC, OBS_POPCORR_ADIABATIC, OBS_KINETIC,
 OBS_POTENTIAL, OBS_TOTAL_ENERGY, OBS_POSITION, OBS_VELOCITY, OBS_DISCRETE_STATE, OBS_SCATTERING,
 OBS_SCATTERING_DIABATIC, OBS_SIGMA, OBS_MAPPING_Q, OBS_MAPPING_P) = range(15)
OBS_COUNT = 15

ERRORS = {0: "ok", -1: "invalid argument", -2: "unsupported configuration (no kernel, no CPU fallback)",
          -3: "no CUDA device", -4: "CUDA error", -5: "call order violated", -6: "out of memory"}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class Config(C.Structure):
    """``nqcb200_config`` -- field order and types must match the header exactly."""
    _fields_ = [
        ("abi_version", C.c_int32), ("method", C.c_int32), ("model", C.c_int32), ("nstates", C.c_int32),
        ("ndofs", C.c_int32), ("nbeads", C.c_int32), ("nelectrons", C.c_int32), ("rescaling", C.c_int32),
        ("estimate_probability", C.c_int32), ("disable_hopping", C.c_int32), ("rng", C.c_int32),
        ("device", C.c_int32), ("save_every", C.c_int32), ("nsave", C.c_int32), ("per_trajectory", C.c_int32),
        ("diagnostics", C.c_int32), ("observables", C.c_uint32), ("reserved0", C.c_uint32),
        ("ntraj", C.c_int64), ("traj_offset", C.c_int64), ("seed", C.c_uint64),
        ("dt", C.c_double), ("t0", C.c_double), ("temperature", C.c_double), ("nrpmd_gamma", C.c_double),
        ("edc_C", C.c_double), ("params", C.c_double * MAX_PARAMS),
        ("masses", _dp), ("bath_a", _dp), ("bath_b", _dp), ("nbath", C.c_int32), ("reserved1", C.c_int32),
    ]


class EngineError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}: {ERRORS.get(code, '?')}] {message}")
        self.code = code


def _as_f64(a, size: Optional[int] = None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    if size is not None and out.size != size:
        raise ValueError(f"expected {size} doubles, got {out.size}")
    return out


def _ptr(a: Optional[np.ndarray], typ=_dp):
    return a.ctypes.data_as(typ) if a is not None else None


class Dist(C.Structure):
    """``nqcb200_dist``: kind 0 = fixed value a, kind 1 = Normal(mean a, sd b)."""
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("a", C.c_double), ("b", C.c_double)]


DIST_FIXED, DIST_NORMAL = 0, 1


def bind(lib: C.CDLL, prefix: str) -> None:
    """Declare argument/return types for every entry point of the header on ``lib``."""
    H = C.c_void_p

    def f(name, args, res=C.c_int, required=True):
        try:
            fn = getattr(lib, prefix + name)
        except AttributeError:
            if required:
                raise
            return
        fn.argtypes, fn.restype = args, res

    f("version", [])
    f("device_count", [], required=False)
    f("create", [C.POINTER(Config), C.POINTER(H)])
    f("destroy", [H])
    f("last_error", [H], C.c_char_p)
    f("observable_width", [H, C.c_int])
    f("set_state", [H, _dp, _dp, _dp, _dp, _ip])
    f("set_state_diabatic", [H, _dp, _dp, _dp, _dp, _ip, _dp])
    f("set_mapping", [H, _dp, _dp])
    f("set_gauge_reference", [H, _dp, C.c_int64])
    f("set_draws", [H, _dp, C.c_int64])
    f("set_noise", [H, _dp, C.c_int64])
    f("set_termination", [H, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double])
    f("get_termination", [H, _lp])
    f("run", [H, C.c_int64])
    f("run_from_host", [H, _dp, _dp, _dp, _dp, _ip, _dp, C.c_int, C.c_int64], required=False)
    f("sample_state", [H, C.POINTER(Dist), C.POINTER(Dist), C.c_int, _dp, _dp, C.c_int, C.c_int32])
    f("sample_occupations", [H, C.c_double])
    f("sample_mapping", [H, C.c_int32])
    f("get_state", [H, _dp, _dp, _dp, _dp, _ip])
    f("get_mapping", [H, _dp, _dp])
    f("get_observable_sum", [H, C.c_int, _dp, C.c_int64])
    f("observable_sum_device", [H, C.POINTER(C.c_void_p), _lp], required=False)
    f("observable_offset", [H, C.c_int, _lp], required=False)
    f("get_observable_per_trajectory", [H, C.c_int, _dp, C.c_int64])
    f("get_diagnostics", [H, _dp, _dp, _dp, _dp])
    f("get_counters", [H, _lp, _lp, _lp, _lp])
    f("get_iesh_stats", [H, _lp, _lp, _lp, _lp])
    f("get_progress", [H, _lp, _lp])
    f("get_last_run_timing", [H, _dp, _lp], required=False)
    f("get_launch_count", [H, _lp], required=False)
    f("get_last_download_timing", [H, _dp, _dp, _lp], required=False)
    f("measure_fp64_peak", [C.c_int, _dp], required=False)


HEADER_SYMBOLS = [
    "version", "device_count", "create", "destroy", "last_error", "observable_width", "set_state",
    "set_state_diabatic", "set_mapping", "set_gauge_reference", "set_draws", "set_noise", "set_termination", "get_termination", "run", "run_from_host", "sample_state", "sample_occupations", "sample_mapping", "get_state", "get_mapping",
    "get_observable_sum", "observable_sum_device", "observable_offset", "get_observable_per_trajectory",
    "get_diagnostics", "get_counters", "get_iesh_stats", "get_progress", "get_last_run_timing", "get_launch_count", "get_last_download_timing", "measure_fp64_peak",
]

_ENGINE_LIB: Optional[C.CDLL] = None


def engine_library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libnqcb200.so")


def load_engine_library() -> C.CDLL:
    """Load ``csrc/libnqcb200.so`` (built by ``__graft_entry__.build()`` / ``csrc/Makefile``)."""
    global _ENGINE_LIB
    if _ENGINE_LIB is None:
        path = engine_library_path()
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()' "
                "or make -C nqcdynamics.jl_b200/csrc).  There is no CPU fallback.")
        lib = C.CDLL(path)
        bind(lib, "nqcb200_")
        _ENGINE_LIB = lib
    return _ENGINE_LIB


class CHandle:
    """Thin object wrapper over an opaque ``<prefix>handle*``; every method is one C call."""

    def __init__(self, lib: C.CDLL, prefix: str, cfg: Config, keepalive=()):
        self._lib, self._p = lib, prefix
        self.cfg = cfg
        self._keepalive = keepalive
        self._h = C.c_void_p()
        rc = getattr(lib, prefix + "create")(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = getattr(lib, prefix + "last_error")(None)
            self._h = C.c_void_p()
            raise EngineError(rc, (msg or b"").decode())
        self.n, self.D, self.B, self.ne = cfg.nstates, cfg.ndofs, cfg.nbeads, cfg.nelectrons
        self.T = int(cfg.ntraj)
        self.nsig = self.n * (self.ne if cfg.method in IESH_FAMILY else self.n)
        self.nstate = self.ne if cfg.method in IESH_FAMILY else 1

    # -- plumbing ------------------------------------------------------------------------------
    def _call(self, name, *args):
        rc = getattr(self._lib, self._p + name)(self._h, *args)
        if rc < 0:
            msg = getattr(self._lib, self._p + "last_error")(self._h)
            raise EngineError(rc, (msg or b"").decode())
        return rc

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            getattr(self._lib, self._p + "destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- state upload / download ---------------------------------------------------------------
    def _state_args(self, r, v, sre, sim, state):
        nrv = self.T * self.B * self.D
        r_, v_ = _as_f64(r, nrv), _as_f64(v, nrv)
        sre_ = _as_f64(sre, self.T * self.nsig) if sre is not None else None
        sim_ = _as_f64(sim, self.T * self.nsig) if sim is not None else None
        st_ = np.ascontiguousarray(state, dtype=np.int32).reshape(-1) if state is not None else None
        if st_ is not None and st_.size != self.T * self.nstate:
            raise ValueError("state has the wrong size")
        return r_, v_, sre_, sim_, st_

    def set_state(self, r, v, sigma_re=None, sigma_im=None, state=None):
        r_, v_, sre_, sim_, st_ = self._state_args(r, v, sigma_re, sigma_im, state)
        self._call("set_state", _ptr(r_), _ptr(v_), _ptr(sre_), _ptr(sim_), _ptr(st_, _ip))

    def set_state_diabatic(self, r, v, rho_re, rho_im=None, state=None, state_draw=None):
        r_, v_, sre_, sim_, st_ = self._state_args(r, v, rho_re, rho_im, state)
        dr_ = _as_f64(state_draw, self.T) if state_draw is not None else None
        self._call("set_state_diabatic", _ptr(r_), _ptr(v_), _ptr(sre_), _ptr(sim_), _ptr(st_, _ip), _ptr(dr_))

    def set_mapping(self, qmap, pmap):
        q, p = _as_f64(qmap, self.T * self.n * self.B), _as_f64(pmap, self.T * self.n * self.B)
        self._call("set_mapping", _ptr(q), _ptr(p))

    def set_gauge_reference(self, Z, count_per_traj=1):
        z = _as_f64(Z, self.T * count_per_traj * self.n * self.n)
        self._call("set_gauge_reference", _ptr(z), C.c_int64(count_per_traj))

    def set_draws(self, xi):
        xi_ = np.ascontiguousarray(xi, dtype=np.float64)
        if xi_.ndim != 2 or xi_.shape[1] != self.T:
            raise ValueError("draws must have shape (nsteps, ntraj)")
        self._call("set_draws", _ptr(xi_.reshape(-1)), C.c_int64(xi_.shape[0]))

    def set_noise(self, xi):
        """ThermalLangevin parity mode: standard normals of shape (nsteps, ntraj, nbeads), one per normal mode."""
        xi_ = np.ascontiguousarray(xi, dtype=np.float64)
        if xi_.ndim != 3 or xi_.shape[1] != self.T or xi_.shape[2] != self.B * self.D:
            raise ValueError("noise must have shape (nsteps, ntraj, nbeads*ndofs)")
        self._call("set_noise", _ptr(xi_.reshape(-1)), C.c_int64(xi_.shape[0]))

    def set_termination(self, dof: int, lo: float, hi: float, outgoing: bool = False, tcut: float = float("inf")):
        """TerminatingCallback(u -> r[dof] < lo || r[dof] > hi || t > tcut) (callbacks.jl:29); ``outgoing``: the position
        clauses also ask for an outward velocity; dof < 0 removes it."""
        self._call("set_termination", C.c_int(int(dof)), C.c_double(float(lo)), C.c_double(float(hi)), C.c_int(int(bool(outgoing))),
                   C.c_double(float(tcut)))

    def termination(self) -> np.ndarray:
        """Steps taken before terminate! fired, per trajectory (-1: still running)."""
        out = np.full(max(self.T, 1), -1, dtype=np.int64)
        self._call("get_termination", out.ctypes.data_as(_lp))
        return out[:self.T]

    def run(self, nsteps: int):
        self._call("run", C.c_int64(int(nsteps)))

    def run_from_host(self, r, v, rho_re=None, rho_im=None, state=None, state_draw=None, diabatic=True, nsteps=0):
        """set_state[_diabatic] + run in one call (nqcb200_run_from_host): pinned r / v are read in place by the kernel."""
        r_, v_, sre_, sim_, st_ = self._state_args(r, v, rho_re, rho_im, state)
        dr_ = _as_f64(state_draw, self.T) if state_draw is not None else None
        if not hasattr(self._lib, self._p + "run_from_host"):      # the CPU oracle mirrors the two separate calls
            if diabatic:
                self._call("set_state_diabatic", _ptr(r_), _ptr(v_), _ptr(sre_), _ptr(sim_), _ptr(st_, _ip), _ptr(dr_))
            else:
                self._call("set_state", _ptr(r_), _ptr(v_), _ptr(sre_), _ptr(sim_), _ptr(st_, _ip))
            self._call("run", C.c_int64(int(nsteps)))
            return
        self._call("run_from_host", _ptr(r_), _ptr(v_), _ptr(sre_), _ptr(sim_), _ptr(st_, _ip), _ptr(dr_),
                   C.c_int(1 if diabatic else 0), C.c_int64(int(nsteps)))

    def sample_state(self, r_spec, v_spec, rho=None, diabatic=True, state=0, normal_modes=False):
        """Device-side initial conditions (nqcb200_sample_state).  r_spec / v_spec: nbeads*ndofs entries, each a number
        (fixed) or a (mean, sd) pair (Normal); rho: ONE n x n matrix (numpy [row, col]) shared by all trajectories."""
        def pack(spec):
            spec = list(spec)
            if len(spec) != self.B * self.D:
                raise ValueError(f"expected {self.B * self.D} component specifications")
            arr = (Dist * len(spec))()
            for i, x in enumerate(spec):
                if isinstance(x, (tuple, list)):
                    arr[i].kind, arr[i].a, arr[i].b = DIST_NORMAL, float(x[0]), float(x[1])
                else:
                    arr[i].kind, arr[i].a, arr[i].b = DIST_FIXED, float(x), 0.0
            return arr
        rd, vd = pack(r_spec), pack(v_spec)
        re = im = None
        if rho is not None:
            m = np.asarray(rho, dtype=np.complex128).reshape(self.n, self.n)
            re = np.ascontiguousarray(m.real.T).reshape(-1)      # column-major
            im = np.ascontiguousarray(m.imag.T).reshape(-1)
        self._call("sample_state", rd, vd, C.c_int(1 if normal_modes else 0), _ptr(re), _ptr(im),
                   C.c_int(1 if diabatic else 0), C.c_int32(int(state)))

    def sample_occupations(self, beta: float):
        """AdiabaticIESH: FermiDiracState{Adiabatic} occupations drawn on the device (nqcb200_sample_occupations)."""
        self._call("sample_occupations", C.c_double(float(beta)))

    def sample_mapping(self, state: int):
        """NRPMD: initial mapping variables for PureState{Diabatic}(state) drawn on the device (nqcb200_sample_mapping)."""
        self._call("sample_mapping", C.c_int32(int(state)))

    def get_state(self):
        T, B, D, n = self.T, self.B, self.D, self.n
        r = np.empty((T, B, D)); v = np.empty((T, B, D))
        has_sig = self.cfg.method in (METHOD_FSSH, METHOD_EHRENFEST) + IESH_FAMILY
        has_state = self.cfg.method in (METHOD_FSSH, METHOD_IESH)
        ncol = self.ne if self.cfg.method in IESH_FAMILY else n
        sre = np.empty((T, ncol, n)) if has_sig else None
        sim = np.empty((T, ncol, n)) if has_sig else None
        st = np.empty((T, self.nstate), dtype=np.int32) if has_state else None
        self._call("get_state", _ptr(r), _ptr(v), _ptr(sre), _ptr(sim), _ptr(st, _ip))
        out = {"r": r, "v": v}
        if has_sig:
            # column-major (n, ncol) per trajectory -> numpy [t, col, row]; expose as [t, row, col]
            out["sigma"] = (sre + 1j * sim).transpose(0, 2, 1)
        if has_state:
            out["state"] = st
        return out

    def get_mapping(self):
        q = np.empty((self.T, self.B, self.n)); p = np.empty((self.T, self.B, self.n))
        self._call("get_mapping", _ptr(q), _ptr(p))
        return q, p

    # -- outputs -------------------------------------------------------------------------------
    def observable_width(self, obs_id: int) -> int:
        return self._call("observable_width", C.c_int(obs_id))

    def observable_sum(self, obs_id: int) -> np.ndarray:
        w = self.observable_width(obs_id)
        out = np.empty((self.cfg.nsave, w))
        self._call("get_observable_sum", C.c_int(obs_id), _ptr(out), C.c_int64(out.size))
        return out

    def observable_per_trajectory(self, obs_id: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """(ntraj, nsave, width); pass a pinned ``out`` of that shape to receive the copy at PCIe speed."""
        w = self.observable_width(obs_id)
        if out is None:
            out = np.empty((self.T, self.cfg.nsave, w))
        elif out.shape != (self.T, self.cfg.nsave, w) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape (ntraj, nsave, width)")
        self._call("get_observable_per_trajectory", C.c_int(obs_id), _ptr(out), C.c_int64(out.size))
        return out

    def diagnostics(self):
        T, n, D, B = self.T, self.n, self.D, self.B
        eig = np.empty((T, n)); nac = np.empty((T, D, n, n)); acc = np.empty((T, B, D)); Z = np.empty((T, n, n))
        self._call("get_diagnostics", _ptr(eig), _ptr(nac), _ptr(acc), _ptr(Z))
        # matrices are column-major: numpy [.., col, row] -> [.., row, col]
        return {"eig": eig, "nac": nac.transpose(0, 1, 3, 2), "accel": acc, "Z": Z.transpose(0, 2, 1)}

    def counters(self):
        vals = [C.c_int64() for _ in range(4)]
        self._call("get_counters", *[C.byref(x) for x in vals])
        return dict(zip(("steps", "hops", "frustrated", "nonfinite"), (int(x.value) for x in vals)))

    def iesh_stats(self) -> dict:
        vals = [C.c_int64() for _ in range(4)]
        self._call("get_iesh_stats", *[C.byref(x) for x in vals])
        return dict(zip(("hop_searches", "determinants", "taylor_stages", "gemm_stages"), (int(x.value) for x in vals)))

    def hop_search_count(self) -> int:
        return self.iesh_stats()["hop_searches"]

    def progress(self):
        a, b = C.c_int64(), C.c_int64()
        self._call("get_progress", C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def last_download_timing(self):
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        self._call("get_last_download_timing", C.byref(a), C.byref(b), C.byref(n))
        return {"transpose_ms": a.value, "copy_ms": b.value, "bytes": int(n.value)}

    def launch_count(self) -> int:
        n = C.c_int64()
        self._call("get_launch_count", C.byref(n))
        return int(n.value)

    def last_run_timing(self):
        ms, n = C.c_double(), C.c_int64()
        self._call("get_last_run_timing", C.byref(ms), C.byref(n))
        return float(ms.value), int(n.value)

    def observable_sum_device(self):
        p, n = C.c_void_p(), C.c_int64()
        self._call("observable_sum_device", C.byref(p), C.byref(n))
        return int(p.value or 0), int(n.value)

    def observable_offset(self, obs_id: int) -> int:
        o = C.c_int64()
        self._call("observable_offset", C.c_int(obs_id), C.byref(o))
        return int(o.value)


def make_config(*, method, model, nstates, ndofs, masses, ntraj, dt, nbeads=1, nelectrons=0, params=(),
                bath_a=None, bath_b=None, rescaling=RESCALE_STANDARD, estimate_probability=1, disable_hopping=0,
                rng=RNG_PHILOX, device=0, save_every=1, nsave=1, per_trajectory=0, diagnostics=0, observables=0,
                traj_offset=0, seed=0, t0=0.0, temperature=0.0, nrpmd_gamma=0.5, edc_C=0.0):
    """Build a :class:`Config`; returns ``(cfg, keepalive)`` -- keep ``keepalive`` referenced until create returns."""
    cfg = Config()
    cfg.abi_version = ABI_VERSION
    cfg.method, cfg.model, cfg.nstates, cfg.ndofs, cfg.nbeads = method, model, nstates, ndofs, nbeads
    cfg.nelectrons, cfg.rescaling = nelectrons, rescaling
    cfg.estimate_probability, cfg.disable_hopping, cfg.rng, cfg.device = estimate_probability, disable_hopping, rng, device
    cfg.save_every, cfg.nsave, cfg.per_trajectory, cfg.diagnostics = save_every, nsave, per_trajectory, diagnostics
    cfg.observables = observables
    cfg.ntraj, cfg.traj_offset, cfg.seed = ntraj, traj_offset, seed
    cfg.dt, cfg.t0, cfg.temperature, cfg.nrpmd_gamma, cfg.edc_C = dt, t0, temperature, nrpmd_gamma, edc_C
    if len(params) > MAX_PARAMS:
        raise ValueError("too many model parameters")
    for i, p in enumerate(params):
        cfg.params[i] = float(p)
    m = _as_f64(masses, ndofs)
    keep = [m]
    cfg.masses = _ptr(m)
    if bath_a is not None:
        a, b = _as_f64(bath_a), _as_f64(bath_b)
        if a.size != b.size:
            raise ValueError("bath arrays differ in length")
        keep += [a, b]
        cfg.bath_a, cfg.bath_b, cfg.nbath = _ptr(a), _ptr(b), a.size
    return cfg, keep

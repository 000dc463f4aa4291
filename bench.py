#!/usr/bin/env python
"""bench.py -- trajectory-steps/s of the ensemble hot path on N B200s (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--trajectories T] [--impl reference]

A "step" is one whole ensemble job of the named workload: every trajectory advanced over the config's full
tspan (e.g. 200 nuclear steps of dt = 0.1 for the spin-boson config) with its observables accumulated on the
device at every save point.  Every step is the same job on a fresh batch: the initial conditions are re-drawn on the
device from the resident distribution parameters (nqcb200_sample_state) before each run; only AdiabaticIESH / NRPMD
(no device sampler) continue the same trajectories across steps.

  value     trajectory-steps/s with the step's inputs resident in HBM: the K timed regions bracket the device-side draw of a
            fresh batch (nqcb200_sample_state: sampling, t0 eigenproblem, save point 0) and the blocking,
            stream-synchronised nqcb200_run of each step (max over ranks); the engine's CUDA-event time of the step
            kernels is kernel_ms_total and feeds the roofline.
  e2e       the same metric through the public C-ABI call sequence with HOST buffers: each of the K steps hands over fresh
            initial conditions in pinned host memory (nqcb200_run_from_host / set_state), runs, and reads the reduced
            observable back.
  roofline  FP64: frac = the flops the kernels EXECUTE (committed ncu instruction mix x the live rate) over the DFMA peak
            measured in this run (MEASURED_PEAKS.json has no FP64 entry); algorithmic_frac = the reference's dense
            formulation, reported next to it because the kernels do not execute that work.
  other_configs  short legs of the other BASELINE configs (C1, C3, C4, C5) measured by the same code in the same run.
  cpu_baseline  the CPU oracle (a C++ restatement of the reference algorithm, NOT Julia) on the host cores,
            on a bounded sample of the same workload.

`--impl reference` times that CPU restatement alone (the Julia reference cannot run here: no julia binary).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_WORKLOAD = "spinboson_debye100_fssh"   # BASELINE.json configs[1]

# Evidence tables, refreshed from the ncu passes committed under profiles/r02/ (tools/profile.sh; SUMMARY.md there):
#   NCU_DRAM_BYTES_PER_TRAJ_STEP   dram__bytes_read.sum + dram__bytes_write.sum of the step kernel(s) per trajectory-step
#   NCU_EXECUTED_FLOPS_PER_TRAJ_STEP  FP64 flops the kernels EXECUTE per trajectory-step (DFMA = 2, DADD = DMUL = 1; thread-level
#                                  counts of the instruction-mix pass / (T x steps)), summed over the kernels of a step
PROFILE_ROUND = "r02"
NCU_DRAM_BYTES_PER_TRAJ_STEP = {}
NCU_EXECUTED_FLOPS_PER_TRAJ_STEP = {}
try:
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", PROFILE_ROUND, "executed_flops.json")) as _f:
        _tab = json.load(_f)
    NCU_EXECUTED_FLOPS_PER_TRAJ_STEP = {k: (v["flops_per_traj_step"], v["source"]) for k, v in _tab.items() if "flops_per_traj_step" in v}
    NCU_DRAM_BYTES_PER_TRAJ_STEP = {k: (v["dram_bytes_per_traj_step"], v["source"]) for k, v in _tab.items() if "dram_bytes_per_traj_step" in v}
except (OSError, ValueError):
    pass


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--trajectories", type=int, default=0, help="trajectories PER GPU (default: the config's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --trajectories (default: the config's count) is the TOTAL over all GPUs (BASELINE config 5: "
                         "10^5 RPSH trajectories on 8 GPUs), sharded with distributed.shard_bounds")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target duration of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short legs of the other BASELINE configs appended under other_configs")
    ap.add_argument("--stream", action="store_true",
                    help="also measure the per-trajectory output-streaming path (SortByTrajectory / FileReduction)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_model_name():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_cpu_oracle(wl, seconds, seed=1):
    """Time the CPU restatement (oracle) on a bounded sample of the workload; returns (traj-steps/s, info)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    import nqcdynamics_jl_b200 as nq
    A = nq._abi
    # all host cores the process may use (torchrun exports OMP_NUM_THREADS=1, which is not what is measured here)
    cores = oracle.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    rng = np.random.default_rng(seed)

    # the large-bath IESH oracle takes ~0.1-10 s per trajectory-step: bound the sample by shortening the run
    nsteps = wl.nsteps if wl.method != A.METHOD_IESH else max(2, min(wl.nsteps, int(2000 // wl.model.nstates)))

    def job(T):
        cfg, keep = A.make_config(**wl.config_kwargs(T, seed=seed))
        h = oracle.OracleEngine(cfg, keep)
        ic = wl.sample(rng, T)
        t0 = time.perf_counter()
        wl.upload(h, ic)
        h.run(nsteps)
        dt = time.perf_counter() - t0
        h.close()
        return dt
    T = max(cores * 2, 16) if wl.method != A.METHOD_IESH else cores
    dt = job(T)                               # calibration sample
    rate = T * nsteps / dt
    T2 = int(max(T, min(rate * seconds / nsteps, 4_000_000)))
    T2 = max(cores, (T2 // cores) * cores)
    dt2 = job(T2) if T2 > T or wl.method != A.METHOD_IESH else dt
    value = T2 * nsteps / dt2
    info = {"value": value, "unit": "trajectory-steps/s", "cores": cores, "kind": "port",
            "sample": f"{T2} trajectories x {nsteps} steps of {wl.name} ({dt2:.1f} s wall), OpenMP over "
                      f"trajectories, g++ -O2 -ffp-contract=off, CPU: {cpu_model_name()}; C++ restatement of the "
                      f"reference algorithm (oracle/), not the Julia reference (no julia binary in this image)"}
    return value, info, T2 * nsteps / wl.nsteps, dt2


def workload_config(wl, T, world, resample, A):
    """The `config` object of the JSON line: identical in the b200 arm and the reference arm (which times a bounded sample
    of the SAME workload and says so under cpu_baseline.sample, not here)."""
    return {"workload": wl.name, "description": wl.description, "trajectories_per_gpu": T,
            "nuclear_steps_per_step": wl.nsteps, "save_every": wl.save_every,
            "batch": ("every step draws a fresh batch from the config's initial-condition distribution and runs the full tspan"
                      if resample else "trajectories continue across steps"),
            "observables_on_device": [o for o in range(A.OBS_COUNT) if (wl.observables >> o) & 1],
            "l2": "trajectory state larger than L2" if T * 8 * 3 * len(wl.masses) * wl.nbeads > 126e6
                  else "state register-resident for the whole launch; no reuse of cached inputs between steps",
            "parallelism": f"trajectories sharded over {world} GPU(s), one NCCL all-reduce of observables"}


def reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import nqcdynamics_jl_b200 as nq
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    times, units = [], []
    info = None
    for i in range(args.warmup + args.steps):
        value, info, T2, dt2 = run_cpu_oracle(wl, per_step, seed=100 + i)
        if i >= args.warmup:
            times.append(dt2); units.append(T2 * wl.nsteps)
    value = sum(units) / sum(times)
    info["value"] = value
    T = args.trajectories or wl.ntraj_default
    line = {"impl": "reference", "metric": "trajectory-steps/sec (FP64)", "value": value, "unit": "trajectory-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(wl, T, max(1, args.gpus), wl.device_spec is not None, nq._abi),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (for the NCCL all-reduce)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class Bench:
    """One process = one GPU.  measure() runs W + K jobs of one workload and returns the numbers of the JSON line."""

    def __init__(self):
        import nqcdynamics_jl_b200 as nq
        from nqcdynamics_jl_b200 import workloads
        self.nq, self.A, self.workloads = nq, nq._abi, workloads
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = self.torch = None
        self.stdout_fd = None
        self.strong_total = None         # --scaling strong: total trajectories over all ranks, this rank's first global index
        self.strong_offset = 0
        self.cpu_affinity = None
        if self.world > 1:
            # pin this rank to the host cores next to ITS GPU before any pinned buffer is allocated (first touch decides the
            # NUMA node of the 1.6 GB of host input the e2e leg streams per job; unbound ranks share one node's memory and
            # its PCIe root).  Purely host-side; skipped when NVML or the affinity call is unavailable.
            try:
                import pynvml
                pynvml.nvmlInit()
                hdl = pynvml.nvmlDeviceGetHandleByIndex(self.local_rank)
                ncpu = os.cpu_count() or 1
                words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (ncpu + 63) // 64)
                cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < ncpu]
                if cpus:
                    os.sched_setaffinity(0, cpus)
                    self.cpu_affinity = f"{len(cpus)} cores ({min(cpus)}-{max(cpus)})"
            except Exception:
                pass
            # stdout carries exactly one JSON line: whatever libraries write to fd 1 meanwhile (NCCL prints its
            # "NCCL version ..." banner there when the communicator is created) goes to stderr until the line is printed
            sys.stdout.flush()
            self.stdout_fd = os.dup(1)
            os.dup2(2, 1)
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist, self.torch = dist, torch
        self.lib = self.A.load_engine_library()
        if self.lib.nqcb200_device_count() <= 0:
            raise SystemExit("bench.py: no CUDA device visible -- the engine has no CPU path")
        peak = ctypes.c_double()
        self.lib.nqcb200_measure_fp64_peak(self.local_rank, ctypes.byref(peak))
        self.peak = float(peak.value)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def allreduce_observables(self, h):
        if self.dist is None:
            return
        ptr, n = h.observable_sum_device()
        if n:
            t = self.torch.as_tensor(_DevArray(ptr, n), device=f"cuda:{self.local_rank}")
            self.dist.all_reduce(t)
            self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.dist is None:
            return vals
        tt = self.torch.tensor(list(vals), device=f"cuda:{self.local_rank}", dtype=self.torch.float64)
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return tuple(float(x) for x in tt)

    def pin(self, a):
        try:
            import torch as _t
            return _t.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        except Exception:
            return np.ascontiguousarray(a)

    # ------------------------------------------------------------------------------------------
    def measure(self, wl, T, K, W, e2e_steps, sample_clocks=True):
        A, Engine = self.A, __import__("nqcdynamics_jl_b200.engine", fromlist=["Engine"]).Engine
        world, rank, local_rank = self.world, self.rank, self.local_rank
        density = wl.method in (A.METHOD_FSSH, A.METHOD_EHRENFEST)
        # Every step is the SAME job.  Where the device-side sampler covers the workload's distribution, each step draws a
        # fresh batch from the resident distribution parameters (nqcb200_sample_state) INSIDE the timed region and runs the
        # config's full tspan; otherwise (AdiabaticIESH, NRPMD) the trajectories keep running across steps.
        resample = wl.device_spec is not None
        nsave_total = wl.nsave if resample else (W + K) * wl.nsteps // wl.save_every + 1
        toff = self.strong_offset if self.strong_total else rank * T
        kw = wl.config_kwargs(T, seed=20261017, device=local_rank, traj_offset=toff)
        kw["nsave"] = nsave_total
        cfg, keep = A.make_config(**kw)
        eng = Engine(cfg, keep)
        rng = np.random.default_rng(1234 + rank)
        ic = wl.sample(rng, T)
        rho = wl.initial_density(T) if density else None
        if wl.method == A.METHOD_IESH:      # ground-state orbitals: psi is built on the device from the occupations
            ic["state"] = wl.iesh_ground_state(T)[1]

        def upload(h, r, v, psi=None):
            return wl.upload(h, {**ic, "r": r, "v": v}, rho)

        upload(eng, ic["r"], ic["v"])
        rho1 = None
        if resample and density:
            rho1 = np.zeros((wl.model.nstates, wl.model.nstates))
            rho1[wl.initial_diabatic_state, wl.initial_diabatic_state] = 1.0

        def fresh_batch(h):
            if resample:
                h.sample_state(wl.device_spec[0], wl.device_spec[1], rho1, diabatic=True, state=0, normal_modes=wl.device_spec[2])

        for _ in range(W):
            fresh_batch(eng)
            eng.run(wl.nsteps)
        self.allreduce_observables(eng)
        self.barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        kernel_ms, launches = 0.0, 0
        launches_before = eng.launch_count()
        wall, prep = 0.0, 0.0
        for k in range(K):
            # timed region of one step: barrier + fresh batch (device-side sampling, t0 eigenproblem, save point 0) +
            # the (blocking) run [+ the job's only exchange after the last step]; the engine synchronises its stream
            # before returning and times its step kernels with CUDA events on that stream
            self.barrier()
            t0 = time.perf_counter()
            fresh_batch(eng)
            t1 = time.perf_counter()
            eng.run(wl.nsteps)
            if k == K - 1:
                self.allreduce_observables(eng)          # one all-reduce of the accumulators
            wall += time.perf_counter() - t0
            prep += t1 - t0
            ms, nl = eng.last_run_timing()
            kernel_ms += ms; launches += nl
        launches_all = eng.launch_count() - launches_before      # sampling / init / step / fold kernels of the K steps
        clocks = sampler.stop() if sampler else None
        self.barrier()
        dev_s, wall = self.max_over_ranks(kernel_ms * 1e-3, wall)
        ntot = float(self.strong_total) if self.strong_total else float(T) * world
        units = ntot * wl.nsteps * K
        value = units / wall
        counters = eng.counters()
        flops_alg = wl.flops_per_traj_step
        flops_note = ""
        executed = None
        if wl.method == A.METHOD_IESH:
            st = eng.iesh_stats()
            counters.update(st)
            frac = st["hop_searches"] / max(1, counters["steps"])
            extra = self.workloads.iesh_hop_search_flops(wl.model.nstates, wl.model.nelectrons)
            flops_alg = wl.flops_per_traj_step + frac * extra
            flops_note = (f"; IESH: base step {wl.flops_per_traj_step:.4g} flops + unpruned hop search {extra:.4g} flops on "
                          f"{frac:.4f} of the steps (measured), both counted in the reference's formulation")
            # executed: the Horner GEMMs of the Taylor propagator on the DMMA pipe, 2 n^2 (2 ne) flops per stage, counted live
            n, ne = wl.model.nstates, wl.model.nelectrons
            executed = (2.0 * n * n * 2 * ne * st["gemm_stages"] / max(1, counters["steps"]),
                        "live: nqcb200_get_iesh_stats gemm_stages x 2 n^2 (2 ne) DMMA flops (the vector-pipe work of the secular "
                        "solver, norms and LU is not counted)")
        elif wl.name in NCU_EXECUTED_FLOPS_PER_TRAJ_STEP:
            executed = NCU_EXECUTED_FLOPS_PER_TRAJ_STEP[wl.name]
        obs_check = (float(np.sum(eng.observable_sum(A.OBS_POPCORR_DIABATIC)[0]))
                     if (wl.observables >> A.OBS_POPCORR_DIABATIC) & 1 else None)

        # ---- end-to-end through the C ABI with HOST buffers: K' jobs, each handing over fresh pinned host arrays ----
        e2e = None
        if e2e_steps > 0:
            kw2 = wl.config_kwargs(T, seed=7, device=local_rank, traj_offset=toff)
            cfg2, keep2 = A.make_config(**kw2)
            eng.close()
            eng2 = Engine(cfg2, keep2)
            r_h, v_h = self.pin(ic["r"]), self.pin(ic["v"])
            rho_h = self.pin(rho) if rho is not None else None
            first_obs = next(o for o in range(A.OBS_COUNT) if (wl.observables >> o) & 1)
            d2h = 0

            def job(h):
                """one ensemble batch through the public C-ABI calls with HOST buffers; returns the bytes handed over"""
                if density:     # set_state_diabatic + run in one call (nqcb200_run_from_host)
                    h.run_from_host(r_h, v_h, rho_h, None, None, None, diabatic=True, nsteps=wl.nsteps)
                    return r_h.nbytes + v_h.nbytes + rho_h.nbytes
                nbytes = upload(h, r_h, v_h)
                h.run(wl.nsteps)
                return nbytes
            h2d = job(eng2)                                                     # warm-up job
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                job(eng2)
                self.allreduce_observables(eng2)
                out = eng2.observable_sum(first_obs)
                d2h = out.nbytes
            e2e_s = time.perf_counter() - t0
            self.barrier()
            (e2e_s,) = self.max_over_ranks(e2e_s)
            e2e = {"value": ntot * wl.nsteps * e2e_steps / e2e_s, "unit": "trajectory-steps/s",
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                   "h2d_GBps_per_rank": h2d * e2e_steps / e2e_s / 1e9, "host_affinity": self.cpu_affinity,
                   "path": ("nqcb200_run_from_host (pinned host r, v: chunked cudaMemcpyAsync on a copy stream under the "
                            "previous chunk's kernels where the kernel family supports it; rho uploaded)" if density
                            else "nqcb200_set_state (pinned host r, v, psi) -> nqcb200_run") + " -> nqcb200_get_observable_sum"}
            eng2.close()
        else:
            eng.close()

        per_launch_units = float(T) * wl.nsteps * K          # trajectory-steps behind kernel_ms_total (this rank)
        kernel_s = max(kernel_ms * 1e-3, 1e-12)
        alg_tf = flops_alg * per_launch_units / kernel_s / 1e12
        roof = {"bound": "fp64", "peak": self.peak, "unit": "TFLOP/s",
                "peak_source": "DFMA microbenchmark measured in this run (nqcb200_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "algorithmic_achieved": alg_tf, "algorithmic_frac": alg_tf / self.peak if self.peak else None,
                "flops_per_trajectory_step_algorithmic": flops_alg,
                "note": "achieved / frac = FP64 flops the kernels EXECUTE (ncu instruction mix of the committed capture x the live "
                        "rate) over the measured DFMA peak; algorithmic_* = the reference's dense complex formulation (SURVEY.md "
                        "8d), which the kernels do not execute (Hermitian / antisymmetric / model structure), see DESIGN.md" + flops_note}
        if executed is not None:
            ex_tf = executed[0] * per_launch_units / kernel_s / 1e12
            roof.update({"achieved": ex_tf, "frac": ex_tf / self.peak if self.peak else None,
                         "flops_per_trajectory_step_executed": executed[0], "executed_source": executed[1]})
        else:
            roof.update({"achieved": None, "frac": None, "executed_source": "no instruction-mix capture committed for this workload"})
        if wl.name in NCU_DRAM_BYTES_PER_TRAJ_STEP:
            b, src = NCU_DRAM_BYTES_PER_TRAJ_STEP[wl.name]
            roof.update({"traffic": b * float(T) * wl.nsteps, "traffic_source": f"{src}: DRAM read + write bytes per trajectory-step "
                         f"of that capture x {T} trajectories x {wl.nsteps} steps (bytes per job)"})
        else:
            roof["traffic"] = None
        return {"value": value, "ms_per_step": 1e3 * wall / K, "e2e": e2e, "gpu_launches": int(launches_all),
                "gpu_step_kernel_launches": int(launches), "clocks": clocks, "roofline": roof, "counters": counters,
                "kernel_ms_total": kernel_ms, "wall_s_timed_region": wall, "batch_prepare_ms": 1e3 * prep / K,
                "observable_checksum": obs_check, "config": workload_config(wl, T, world, resample, A)}


# short legs of the other BASELINE configs appended to the headline line (trajectories per GPU, steps, warm-up)
OTHER_CONFIGS = [("tully1_fssh", 1 << 21, 2, 1), ("rpmd_harmonic32", 1 << 17, 2, 1), ("rpsh_morse3_16", 113664, 2, 1),
                 ("iesh_anderson_holstein_m100", 2960, 2, 1)]


def main():
    args = parse_args()
    import nqcdynamics_jl_b200 as nq
    from nqcdynamics_jl_b200 import workloads
    A = nq._abi
    wl = workloads.get(args.workload)
    if args.impl == "reference":
        reference_arm(args, wl)
        return
    B = Bench()
    T = args.trajectories or wl.ntraj_default
    if args.scaling == "strong":
        from nqcdynamics_jl_b200.distributed import shard_bounds
        lo, hi = shard_bounds(T, B.world, B.rank)
        B.strong_total, B.strong_offset = T, lo
        T = hi - lo
    K, W = args.steps, args.warmup
    res = B.measure(wl, T, K, W, 0 if args.no_e2e else K)

    other = None
    if not args.no_other_configs and args.workload == DEFAULT_WORKLOAD and args.scaling == "weak":
        # the north_star's other targets (C1 >= 1e9 on 8 GPUs, C4 as a fraction of the FP64 peak, C3, C5), measured by the
        # same code in the same run so that the driver sees them: short legs, no e2e / clocks
        other = {}
        for name, To, Ko, Wo in OTHER_CONFIGS:
            try:
                r = B.measure(workloads.get(name), To, Ko, Wo, 0, sample_clocks=False)
                other[name] = {"value": r["value"], "unit": "trajectory-steps/s", "ms_per_step": r["ms_per_step"],
                               "trajectories_per_gpu": To, "steps": Ko, "warmup": Wo,
                               "roofline": {k: r["roofline"].get(k) for k in ("achieved", "frac", "algorithmic_frac", "peak",
                                                                                 "flops_per_trajectory_step_executed", "executed_source")},
                               "counters": r["counters"]}
            except Exception as exc:      # a leg must never take the headline down
                other[name] = {"error": str(exc)[:200]}

    # ---- output-streaming path: per-trajectory observables written at every save point, transposed to the
    # reference's trajectory-major layout at HBM speed and copied to pinned host memory ---------------------
    stream = None
    if args.stream and B.rank == 0:
        stream = measure_stream(B, wl, T)

    cpu = None
    if B.rank == 0 and B.world == 1 and not args.no_cpu_baseline:
        _, cpu, _, _ = run_cpu_oracle(wl, args.cpu_seconds)

    if B.rank == 0:
        line = {"metric": "trajectory-steps/sec (FP64)", "value": res["value"], "unit": "trajectory-steps/s", "n_gpus": B.world,
                "steps": K, "warmup": W, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": res["config"], "e2e": res["e2e"],
                "gpu_launches": res["gpu_launches"], "gpu_step_kernel_launches": res["gpu_step_kernel_launches"],
                "clocks": res["clocks"], "roofline": res["roofline"], "cpu_baseline": cpu, "stream": stream,
                "counters": res["counters"], "kernel_ms_total": res["kernel_ms_total"],
                "wall_s_timed_region": res["wall_s_timed_region"], "batch_prepare_ms": res["batch_prepare_ms"],
                "observable_checksum": res["observable_checksum"], "other_configs": other}
        if B.stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(B.stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if B.dist is not None:
        B.dist.destroy_process_group()


def measure_stream(B, wl, T):
    import torch as _t
    A, Engine = B.A, __import__("nqcdynamics_jl_b200.engine", fromlist=["Engine"]).Engine
    obs_ids = [o for o in range(A.OBS_COUNT) if (wl.observables >> o) & 1]
    probe = Engine(*A.make_config(**wl.config_kwargs(1, device=B.local_rank)))
    widths = {o: probe.observable_width(o) for o in obs_ids}
    probe.close()
    per_traj_bytes = 8 * wl.nsave * sum(widths.values())
    Ts = int(max(1024, min(T, (6 << 30) // max(1, per_traj_bytes))))        # <= 6 GiB of output
    kw3 = wl.config_kwargs(Ts, seed=11, device=B.local_rank, per_trajectory=1)
    es = Engine(*A.make_config(**kw3))
    ics = wl.sample(np.random.default_rng(99), Ts)
    wl.upload(es, ics, wl.initial_density(Ts) if wl.method in (A.METHOD_FSSH, A.METHOD_EHRENFEST) else None)
    es.run(wl.nsteps)
    ms_stream, _ = es.last_run_timing()
    tr_ms = cp_ms = 0.0
    nbytes = 0
    for o in obs_ids:
        buf = _t.empty((Ts, wl.nsave, widths[o]), dtype=_t.float64).pin_memory().numpy()
        es.observable_per_trajectory(o, out=buf)
        tm = es.last_download_timing()
        tr_ms += tm["transpose_ms"]; cp_ms += tm["copy_ms"]; nbytes += tm["bytes"]
    es.close()
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    tr_gbs = 2.0 * nbytes / (tr_ms * 1e-3) / 1e9 if tr_ms > 0 else None      # read + write
    return {"trajectories": Ts, "output_bytes": nbytes,
            "step_kernel_traj_steps_per_s": float(Ts) * wl.nsteps / (ms_stream * 1e-3),
            "transpose": {"ms": tr_ms, "GB/s": tr_gbs, "hbm_peak_GB/s": hbm_peak,
                          "frac": (tr_gbs / hbm_peak) if (tr_gbs and hbm_peak) else None,
                          "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "unavailable"},
            "d2h": {"ms": cp_ms, "GB/s": nbytes / (cp_ms * 1e-3) / 1e9 if cp_ms > 0 else None}}


if __name__ == "__main__":
    main()

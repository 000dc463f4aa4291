#!/usr/bin/env python
"""bench.py -- trajectory-steps/s of the ensemble hot path on N B200s (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--trajectories T] [--impl reference]

A "step" is one whole ensemble job of the named workload: every trajectory advanced over the config's full
tspan (e.g. 200 nuclear steps of dt = 0.1 for the spin-boson config) with its observables accumulated on the
device at every save point.  Every step is the same job on a fresh batch: the initial conditions are re-drawn on the
device from the resident distribution parameters (nqcb200_sample_state) before each run; only AdiabaticIESH / NRPMD
(no device sampler) continue the same trajectories across steps.

  value     trajectory-steps/s with the step's inputs resident in HBM: the K timed regions bracket the blocking,
            stream-synchronised nqcb200_run of each step (max over ranks); the engine's CUDA-event time of the same
            launches is kernel_ms_total and feeds the roofline.
  e2e       the same metric through the public C-ABI call sequence with HOST buffers: every step hands over fresh
            initial conditions in pinned host memory (nqcb200_run_from_host / set_state), runs, and reads the reduced
            observable back.
  roofline  FP64: algorithmic flops per trajectory-step x trajectory-steps per launch / kernel
            time, against the DFMA peak measured in this run (MEASURED_PEAKS.json has no FP64 entry).
  cpu_baseline  the CPU oracle (a C++ restatement of the reference algorithm, NOT Julia) on the host cores,
            on a bounded sample of the same workload.

`--impl reference` times that CPU restatement alone (the Julia reference cannot run here: no julia binary).
"""
import argparse
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

# dram__bytes_read.sum + dram__bytes_write.sum of the step kernel per TRAJECTORY per launch, from the `ncu --set full`
# captures committed under profiles/r01/ (prof_r01_<tag>_raw.csv; bytes / trajectories of the capture).  The kernels
# touch HBM only at launch entry / exit (state in, state out) and at save points, so the traffic of a launch scales
# with the number of trajectories, not with the number of steps.
NCU_DRAM_BYTES_PER_TRAJ = {
    "spinboson_debye100_fssh": (2696.0, "sb_v5"), "spinboson_debye100_ehrenfest": (2696.0, "sb_v5"),
    "tully1_fssh": (236.0, "tully1_v3"), "rpmd_harmonic32": (1356.0, "rpmd_fft"), "rpsh_morse3_16": (664.0, "rpsh_tpt2"),
}
# FP64 flops the kernels EXECUTE per trajectory-step (DFMA = 2), from the committed instruction-mix passes
# profiles/r01/instmix_r01_<tag>.csv (thread-level DFMA/DADD/DMUL counts / (T x steps)); see profiles/r01/SUMMARY.md
NCU_EXECUTED_FLOPS_PER_TRAJ_STEP = {
    "spinboson_debye100_fssh": (3630.0, "sb_v5"), "tully1_fssh": (1306.0, "tully1_v3"), "rpmd_harmonic32": (1940.0, "rpmd_tpt"),
    "rpsh_morse3_16": (12700.0, "rpsh_tpt2"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--trajectories", type=int, default=0, help="trajectories PER GPU (default: the config's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target duration of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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


def reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    times, units = [], []
    info = None
    for i in range(args.warmup + args.steps):
        value, info, T2, dt2 = run_cpu_oracle(wl, per_step, seed=100 + i)
        if i >= args.warmup:
            times.append(dt2); units.append(T2 * wl.nsteps)
    value = sum(units) / sum(times)
    info["value"] = value
    line = {"impl": "reference", "metric": "trajectory-steps/sec (FP64)", "value": value, "unit": "trajectory-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "description": wl.description,
                       "note": "CPU restatement of the reference algorithm on the host cores; bounded sample per step"},
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (for the NCCL all-reduce)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def main():
    args = parse_args()
    import nqcdynamics_jl_b200 as nq
    from nqcdynamics_jl_b200 import workloads
    A = nq._abi
    wl = workloads.get(args.workload)
    if args.impl == "reference":
        reference_arm(args, wl)
        return

    from nqcdynamics_jl_b200.engine import Engine
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch = None
    stdout_fd = None
    if world > 1:
        # stdout carries exactly one JSON line: whatever libraries write to fd 1 meanwhile (NCCL prints its
        # "NCCL version ..." banner there when the communicator is created) is sent to stderr until the line is printed
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = A.load_engine_library()
    if lib.nqcb200_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device visible -- the engine has no CPU path")

    T = args.trajectories or wl.ntraj_default
    K, W = args.steps, args.warmup
    density = wl.method in (A.METHOD_FSSH, A.METHOD_EHRENFEST)
    # one handle.  Every step is the SAME job: where the device-side sampler covers the workload's distribution, each
    # step re-draws the batch from the resident distribution parameters (nqcb200_sample_state) and runs the config's
    # full tspan; otherwise (AdiabaticIESH, NRPMD) the trajectories keep running and the save capacity covers all steps.
    resample = wl.device_spec is not None
    nsave_total = wl.nsave if resample else (W + K) * wl.nsteps // wl.save_every + 1
    kw = wl.config_kwargs(T, seed=20261017, device=local_rank, traj_offset=rank * T)
    kw["nsave"] = nsave_total
    cfg, keep = A.make_config(**kw)
    eng = Engine(cfg, keep)
    rng = np.random.default_rng(1234 + rank)
    ic = wl.sample(rng, T)
    rho = wl.initial_density(T) if density else None

    if wl.method == A.METHOD_IESH:
        ic["psi"], ic["state"] = wl.iesh_ground_state(T)

    def upload(h, r, v, psi=None):
        return wl.upload(h, {**ic, "r": r, "v": v, "psi": ic.get("psi") if psi is None else psi}, rho)

    upload(eng, ic["r"], ic["v"])
    peak = __import__("ctypes").c_double()
    lib.nqcb200_measure_fp64_peak(local_rank, __import__("ctypes").byref(peak))

    def barrier():
        if dist is not None:
            dist.barrier()

    def allreduce_observables(h):
        if dist is None:
            return
        ptr, n = h.observable_sum_device()
        if n:
            t = torch.as_tensor(_DevArray(ptr, n), device=f"cuda:{local_rank}")
            dist.all_reduce(t)
            torch.cuda.synchronize()

    rho1 = None
    if resample and density:
        rho1 = np.zeros((wl.model.nstates, wl.model.nstates))
        rho1[wl.initial_diabatic_state, wl.initial_diabatic_state] = 1.0

    def fresh_batch():
        if resample:
            eng.sample_state(wl.device_spec[0], wl.device_spec[1], rho1, diabatic=True, state=0, normal_modes=wl.device_spec[2])

    for _ in range(W):
        fresh_batch()
        eng.run(wl.nsteps)
    allreduce_observables(eng)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    kernel_ms, launches = 0.0, 0
    launches_before = eng.launch_count()
    wall, prep = 0.0, 0.0
    for k in range(K):
        # the step's inputs are made resident first (untimed, reported as batch_prepare_ms): a fresh batch drawn on the
        # device from the distribution parameters, gauge reference, t0 eigenproblem / save point 0
        tp = time.perf_counter()
        fresh_batch()
        prep += time.perf_counter() - tp
        # timed region of one step: barrier + (blocking) run [+ the job's only exchange after the last step] ; the engine
        # synchronises its stream before returning, and times its kernels with CUDA events on that stream
        barrier()
        t0 = time.perf_counter()
        eng.run(wl.nsteps)
        if k == K - 1:
            allreduce_observables(eng)          # one all-reduce of the accumulators
        wall += time.perf_counter() - t0
        ms, nl = eng.last_run_timing()
        kernel_ms += ms; launches += nl
    launches_all = eng.launch_count() - launches_before      # sampling / init / step / fold kernels of the K steps
    clocks = sampler.stop()
    barrier()
    # value: the K timed regions (host clock around blocking, stream-synchronised calls), max over ranks; the CUDA-event
    # kernel time of the same launches feeds the roofline
    dev_s = kernel_ms * 1e-3
    if dist is not None:
        tt = torch.tensor([dev_s, wall], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s, wall = float(tt[0]), float(tt[1])
    region_s = wall
    units = float(T) * world * wl.nsteps * K
    value = units / region_s
    counters = eng.counters()
    flops_step = wl.flops_per_traj_step
    flops_note = ""
    if wl.method == A.METHOD_IESH:
        counters.update(eng.iesh_stats())
        frac = counters["hop_searches"] / max(1, counters["steps"])
        extra = workloads.iesh_hop_search_flops(wl.model.nstates, wl.model.nelectrons)
        flops_step = wl.flops_per_traj_step + frac * extra
        flops_note = (f"; IESH: base step {wl.flops_per_traj_step:.4g} flops + unpruned hop search {extra:.4g} flops on "
                      f"{frac:.4f} of the steps (measured), both counted in the reference's formulation")
    obs_check = float(np.sum(eng.observable_sum(A.OBS_POPCORR_DIABATIC)[0])) if (wl.observables >> A.OBS_POPCORR_DIABATIC) & 1 else None

    # ---- end-to-end through the C ABI with host buffers -------------------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            import torch as _t
            pin = lambda a: _t.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        except Exception:
            pin = np.ascontiguousarray
        kw2 = wl.config_kwargs(T, seed=7, device=local_rank, traj_offset=rank * T)
        cfg2, keep2 = A.make_config(**kw2)
        eng.close()
        eng2 = Engine(cfg2, keep2)
        r_h, v_h = pin(ic["r"]), pin(ic["v"])
        rho_h = pin(rho) if rho is not None else None
        psi_h = pin(ic["psi"]) if "psi" in ic else None
        first_obs = next(o for o in range(A.OBS_COUNT) if (wl.observables >> o) & 1)
        d2h = 0
        ke = max(1, min(K, 3))
        def job(h):
            """one ensemble batch through the public C-ABI calls with HOST buffers; returns the bytes handed over"""
            if density:     # set_state_diabatic + run in one call; the SpinBoson kernels read pinned r, v in place
                h.run_from_host(r_h, v_h, rho_h, None, None, None, diabatic=True, nsteps=wl.nsteps)
                return r_h.nbytes + v_h.nbytes + rho_h.nbytes
            nbytes = upload(h, r_h, v_h, psi_h)
            h.run(wl.nsteps)
            return nbytes
        h2d = job(eng2)                                                     # warm-up job
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            job(eng2)
            allreduce_observables(eng2)
            out = eng2.observable_sum(first_obs)
            d2h = out.nbytes
        e2e_s = time.perf_counter() - t0
        barrier()
        if dist is not None:
            tt = torch.tensor([e2e_s], device=f"cuda:{local_rank}", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt[0])
        e2e = {"value": float(T) * world * wl.nsteps * ke / e2e_s, "unit": "trajectory-steps/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": ke,
               "path": ("nqcb200_run_from_host (pinned host r, v read in place by the step kernel; rho uploaded)" if density else "nqcb200_set_state (pinned host r, v, psi) -> nqcb200_run") + " -> nqcb200_get_observable_sum"}
        # informational: the same job with DEVICE-side initial conditions (nqcb200_sample_state: only the
        # distribution parameters cross PCIe) -- not the contract's e2e, which keeps host buffers
        if wl.device_spec is not None:
            rs, vs, nm = wl.device_spec
            rho1 = None
            if density:
                rho1 = np.zeros((wl.model.nstates, wl.model.nstates))
                rho1[wl.initial_diabatic_state, wl.initial_diabatic_state] = 1.0
            def job_dev(h):
                h.sample_state(rs, vs, rho1, diabatic=True, state=0, normal_modes=nm)
                h.run(wl.nsteps)
            job_dev(eng2)
            barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                job_dev(eng2)
                allreduce_observables(eng2)
                eng2.observable_sum(first_obs)
            dev_s2 = time.perf_counter() - t0
            barrier()
            e2e["device_sampled_ic"] = {"value": float(T) * world * wl.nsteps * ke / dev_s2, "unit": "trajectory-steps/s",
                                        "path": "nqcb200_sample_state -> nqcb200_run -> nqcb200_get_observable_sum (max over ranks not taken)"}
        eng2.close()

    # ---- output-streaming path: per-trajectory observables written at every save point, transposed to the
    # reference's trajectory-major layout at HBM speed and copied to pinned host memory ---------------------
    stream = None
    if args.stream and rank == 0:
        import torch as _t
        obs_ids = [o for o in range(A.OBS_COUNT) if (wl.observables >> o) & 1]
        probe = Engine(*A.make_config(**wl.config_kwargs(1, device=local_rank)))
        widths = {o: probe.observable_width(o) for o in obs_ids}
        probe.close()
        per_traj_bytes = 8 * wl.nsave * sum(widths.values())
        Ts = int(max(1024, min(T, (6 << 30) // max(1, per_traj_bytes))))        # <= 6 GiB of output
        kw3 = wl.config_kwargs(Ts, seed=11, device=local_rank, per_trajectory=1)
        es = Engine(*A.make_config(**kw3))
        ics = {k: v[:Ts] for k, v in ic.items()} if Ts <= T else wl.sample(rng, Ts)
        wl.upload(es, ics, rho[:Ts] if rho is not None else None)
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
        stream = {"trajectories": Ts, "output_bytes": nbytes,
                  "step_kernel_traj_steps_per_s": float(Ts) * wl.nsteps / (ms_stream * 1e-3),
                  "transpose": {"ms": tr_ms, "GB/s": tr_gbs, "hbm_peak_GB/s": hbm_peak,
                                "frac": (tr_gbs / hbm_peak) if (tr_gbs and hbm_peak) else None,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "unavailable"},
                  "d2h": {"ms": cp_ms, "GB/s": nbytes / (cp_ms * 1e-3) / 1e9 if cp_ms > 0 else None}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, cpu, _, _ = run_cpu_oracle(wl, args.cpu_seconds)

    if rank == 0:
        per_launch_units = float(T) * wl.nsteps            # one run() = one launch (<= 65536 steps)
        kernel_s_per_launch = (kernel_ms * 1e-3) / max(1, launches)
        achieved = flops_step * per_launch_units / kernel_s_per_launch / 1e12
        line = {
            "metric": "trajectory-steps/sec (FP64)", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * region_s / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "description": wl.description, "trajectories_per_gpu": T,
                       "nuclear_steps_per_step": wl.nsteps, "save_every": wl.save_every,
                       "batch": ("every step re-draws the batch on the device (nqcb200_sample_state) and runs the full tspan"
                                 if resample else "trajectories continue across steps"),
                       "observables_on_device": [o for o in range(A.OBS_COUNT) if (wl.observables >> o) & 1],
                       "l2": "trajectory state larger than L2" if T * 8 * 3 * len(wl.masses) * wl.nbeads > 126e6
                             else "state register-resident for the whole launch; no reuse of cached inputs between steps",
                       "parallelism": f"trajectories sharded over {world} GPU(s), one NCCL all-reduce of observables"},
            "e2e": e2e,
            "gpu_launches": int(launches_all), "gpu_step_kernel_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": float(peak.value), "unit": "TFLOP/s",
                         "frac": achieved / float(peak.value) if peak.value else None,
                         "traffic": (NCU_DRAM_BYTES_PER_TRAJ[wl.name][0] * T if wl.name in NCU_DRAM_BYTES_PER_TRAJ else None),
                         "traffic_source": (f"profiles/r01/prof_r01_{NCU_DRAM_BYTES_PER_TRAJ[wl.name][1]}_raw.csv: DRAM read + write "
                                            f"bytes per trajectory of that capture x {T} trajectories (bytes per launch)"
                                            if wl.name in NCU_DRAM_BYTES_PER_TRAJ else None),
                         "flops_per_trajectory_step_algorithmic": flops_step,
                         "executed": ({"flops_per_trajectory_step": NCU_EXECUTED_FLOPS_PER_TRAJ_STEP[wl.name][0],
                                       "achieved": NCU_EXECUTED_FLOPS_PER_TRAJ_STEP[wl.name][0] * per_launch_units / kernel_s_per_launch / 1e12,
                                       "frac": (NCU_EXECUTED_FLOPS_PER_TRAJ_STEP[wl.name][0] * per_launch_units / kernel_s_per_launch / 1e12
                                                / float(peak.value)) if peak.value else None,
                                       "source": f"profiles/r01/instmix_r01_{NCU_EXECUTED_FLOPS_PER_TRAJ_STEP[wl.name][1]}.csv"}
                                      if wl.name in NCU_EXECUTED_FLOPS_PER_TRAJ_STEP else None),
                         "peak_source": "DFMA microbenchmark measured in this run (nqcb200_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "note": "algorithmic = the reference's dense complex formulation (SURVEY.md 8d); the kernel "
                                 "executes fewer flops (Hermitian/antisymmetric structure), see DESIGN.md" + flops_note},
            "cpu_baseline": cpu, "stream": stream,
            "counters": counters, "kernel_ms_total": kernel_ms, "wall_s_timed_region": wall,
            "batch_prepare_ms": 1e3 * prep / K,
            "observable_checksum": obs_check,
        }
        if stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

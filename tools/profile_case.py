"""Run a few launches of one workload's step kernel (for ncu).  usage: profile_case.py WORKLOAD [T] [NSTEPS] [LAUNCHES]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import nqcdynamics_jl_b200 as nq
from nqcdynamics_jl_b200 import workloads
from nqcdynamics_jl_b200.engine import Engine
A = nq._abi
wl = workloads.get(sys.argv[1])
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else wl.nsteps
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 3
kw = wl.config_kwargs(T, seed=1, nsave=launches * nsteps // wl.save_every + 1)
cfg, keep = A.make_config(**kw)
e = Engine(cfg, keep)
ic = wl.sample(np.random.default_rng(0), T)
wl.upload(e, ic)
for i in range(launches):
    e.run(nsteps)
    ms, nl = e.last_run_timing()
    print(f"launch {i}: {ms:.3f} ms, {T * nsteps / ms * 1e3:.4g} traj-steps/s")
if wl.method == A.METHOD_IESH:
    print(e.iesh_stats(), e.counters())

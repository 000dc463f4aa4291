"""world_size-2 gloo test of the N>1 host logic on CPU: contiguous trajectory shards keyed by the global
trajectory id + one all-reduce(sum) of the observable accumulators reproduce the single-shard ensemble.
The per-shard compute here is the CPU oracle (test infrastructure); on the GPU box bench.py drives the same
sharding with the CUDA engine and NCCL."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, T, nsteps, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import nqcdynamics_jl_b200 as nq
    import oracle
    from nqcdynamics_jl_b200.distributed import allreduce_sum, shard_bounds
    from helpers import A, model_config
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle.set_num_threads(2)
    lo, hi = shard_bounds(T, world, rank)
    rng = np.random.default_rng(0)
    r = rng.normal(-6.0, 0.5, T); v = np.full(T, 12.0 / 2000)
    obs = (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_SCATTERING)
    kw = model_config(nq.TullyModelOne(), method=A.METHOD_FSSH, masses=[2000.0], ntraj=hi - lo, dt=1.0, seed=77,
                      traj_offset=lo, save_every=50, nsave=nsteps // 50 + 1, observables=obs)
    cfg, keep = A.make_config(**kw)
    h = oracle.OracleEngine(cfg, keep)
    rho = np.zeros((hi - lo, 2, 2)); rho[:, 0, 0] = 1
    h.set_termination(0, -7.5, -3.5, True)          # TerminatingCallback mask: shard-independent like everything else
    h.set_state_diabatic(r[lo:hi], v[lo:hi], rho)
    h.run(nsteps)
    ts = h.termination()
    acc = np.concatenate([h.observable_sum(A.OBS_DIABATIC_POP).ravel(), h.observable_sum(A.OBS_SCATTERING).ravel(),
                          [float(ts[ts >= 0].sum()), float(np.count_nonzero(ts >= 0)), float(h.counters()["steps"])]])
    allreduce_sum(acc)
    if rank == 0:
        q.put(acc)
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_shard():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nqcdynamics_jl_b200 as nq
    import oracle
    from helpers import A, model_config
    T, nsteps = 101, 600
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, T, nsteps, q)) for r in range(2)]
    for p in procs: p.start()
    acc = q.get(timeout=240)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    rng = np.random.default_rng(0)
    r = rng.normal(-6.0, 0.5, T); v = np.full(T, 12.0 / 2000)
    obs = (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_SCATTERING)
    kw = model_config(nq.TullyModelOne(), method=A.METHOD_FSSH, masses=[2000.0], ntraj=T, dt=1.0, seed=77,
                      save_every=50, nsave=nsteps // 50 + 1, observables=obs)
    cfg, keep = A.make_config(**kw)
    h = oracle.OracleEngine(cfg, keep)
    rho = np.zeros((T, 2, 2)); rho[:, 0, 0] = 1
    h.set_termination(0, -7.5, -3.5, True)
    h.set_state_diabatic(r, v, rho)
    h.run(nsteps)
    ts = h.termination()
    assert 0 < np.count_nonzero(ts >= 0) < T, "the case mixes terminated and running trajectories"
    ref = np.concatenate([h.observable_sum(A.OBS_DIABATIC_POP).ravel(), h.observable_sum(A.OBS_SCATTERING).ravel(),
                          [float(ts[ts >= 0].sum()), float(np.count_nonzero(ts >= 0)), float(h.counters()["steps"])]])
    assert np.max(np.abs(acc[:-3] - ref[:-3])) < 1e-10 * T
    assert np.array_equal(acc[-3:], ref[-3:]), "termination steps and step counters are shard-independent"

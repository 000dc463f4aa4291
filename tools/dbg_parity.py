import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
t0 = time.time()
def log(*a):
    print(f"[{time.time()-t0:7.2f}s]", *a, flush=True)
log('nproc', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)), 'OMP', os.environ.get('OMP_NUM_THREADS'))
import numpy as np
import nqcdynamics_jl_b200 as nq
from helpers import *
log('imported')
import oracle
log('oracle threads', oracle.set_num_threads(0))
T = 96
kw = model_config(nq.TullyModelOne(), method=A.METHOD_FSSH, masses=[2000.0], ntraj=T, dt=1.0, rng=A.RNG_INJECTED, diagnostics=1,
                  save_every=1, nsave=41, observables=ALL_POP_OBS, per_trajectory=1)
cfg, keep = A.make_config(**kw)
o = oracle.OracleEngine(cfg, keep); log('oracle created')
cfg2, keep2 = A.make_config(**kw)
e = engine_factory()(cfg2, keep2); log('engine created')
rng = np.random.default_rng(0)
r = -5 + 0.3*rng.standard_normal(T); v = np.full(T, 0.005)
rho = np.zeros((T,2,2)); rho[:,0,0] = 1
draws = rng.random((40, T)); sd = rng.random(T)
o.set_state_diabatic(r, v, rho, None, None, sd); o.set_draws(draws); log('oracle set_state')
e.set_state_diabatic(r, v, rho, None, None, sd); log('engine set_state'); e.set_draws(draws); log('engine set_draws')
for i in range(3):
    o.run(1); log('oracle step')
    e.run(1); log('engine step', e.last_run_timing())
se, so = e.get_state(), o.get_state(); log('get_state')
for k in ('r','v'):
    print(k, np.abs(se[k]-so[k]).max())
print('sigma', np.abs(se['sigma']-so['sigma']).max(), 'state', (se['state']!=so['state']).sum())
de, do = e.diagnostics(), o.diagnostics()
for k in de: print(k, np.abs(de[k]-do[k]).max())
log('done')

"""Run-to-run spread of the device sampler's step time: `ens_repeat.py <mode> <walkers> <steps> <repeats>` (config-3 model)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, models
from golemflavor_b200 import llh, mcmc
from golemflavor_b200.enums import Texture
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
mode, k, steps, reps = (int(x) for x in sys.argv[1:5])
a3, as3, ps3 = models.bsm_model_c3(g['asimov_angles'], dim=6, texture=Texture.OET)
f3 = llh.LnProb(a3, as3, ps3)
np.random.seed(25)
p3 = mcmc.flat_seed(ps3, k)
s = mcmc.DeviceEnsembleSampler(k, f3.ndim, f3, seed=25, mode=mode)
s.run_mcmc(p3, 100, store=False)
out = []
for r in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run_mcmc(None, steps, store=False, return_tensor=True)
    torch.cuda.synchronize()
    out.append((time.perf_counter() - t0) / steps * 1e6)
print('mode %d walkers %d: us/step min %.2f median %.2f max %.2f  all %s' % (mode, k, min(out), float(np.median(out)), max(out), ' '.join('%.1f' % x for x in out)), flush=True)

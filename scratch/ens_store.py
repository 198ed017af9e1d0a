"""Config 2 with chain storage (1024 walkers x 10^4 steps, chain + lnprob chain stored)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, models
from golemflavor_b200 import llh, mcmc
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
a2, as2, ps2 = models.notebook_model(g['asimov_angles'])
f2 = llh.LnProb(a2, as2, ps2)
np.random.seed(25)
p2 = mcmc.flat_seed(ps2, 1024)
p2[:, 4], p2[:, 5] = np.random.uniform(.9, 1, 1024), np.random.uniform(.8, 1, 1024)
for store in (False, True, True):
    s = mcmc.DeviceEnsembleSampler(1024, 6, f2, seed=25)
    s.run_mcmc(p2, 200, store=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.run_mcmc(None, 10000, store=store, return_tensor=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('store=%s: %.4f s (%.2f us/step)' % (store, dt, dt * 100))

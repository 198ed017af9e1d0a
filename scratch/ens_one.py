"""One config-2 run (1024 walkers, 6-D) of the device sampler for profiling: 100 warm-up steps, then N steps in one launch."""
import os, sys
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
s = mcmc.DeviceEnsembleSampler(1024, 6, f2, seed=25)
s.run_mcmc(p2, 100, store=False)
s.run_mcmc(None, int(sys.argv[1]) if len(sys.argv) > 1 else 1000, store=False)
torch.cuda.synchronize()
print('acceptance', float(np.mean(s.acceptance_fraction)))

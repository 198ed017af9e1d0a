"""One config-3 run (4096 walkers, BSM dim-6) of the device sampler: 100 warm-up steps, then N steps in one launch (with the
-DGF_ENS_PROFILE build the cluster kernel prints its clock64 anatomy)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, models
from golemflavor_b200 import llh, mcmc
from golemflavor_b200.enums import Texture
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 20
a3, as3, ps3 = models.bsm_model_c3(g['asimov_angles'], dim=6, texture=Texture.OET)
a3.binning = np.logspace(np.log10(6e4), np.log10(1e7), nb + 1)
f3 = llh.LnProb(a3, as3, ps3)
np.random.seed(25)
p3 = mcmc.flat_seed(ps3, k)
s = mcmc.DeviceEnsembleSampler(k, f3.ndim, f3, seed=25)
s.run_mcmc(p3, 100, store=False)
torch.cuda.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
t0 = time.perf_counter()
s.run_mcmc(None, n, store=False, return_tensor=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print('C3 walkers %d nbins %d: %.2f us/step, acceptance %.3f' % (k, nb, dt / n * 1e6, float(np.mean(s.acceptance_fraction))), flush=True)

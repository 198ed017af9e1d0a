"""Latency anatomy of the cluster sampler on the BSM model: step time vs number of energy bins and vs ensemble size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, models
from golemflavor_b200 import llh, mcmc
from golemflavor_b200.enums import Texture
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
def run(nb, k, steps=1000):
    a3, as3, ps3 = models.bsm_model_c3(g['asimov_angles'], dim=6, texture=Texture.OET)
    a3.binning = np.logspace(np.log10(6e4), np.log10(1e7), nb + 1)
    f3 = llh.LnProb(a3, as3, ps3)
    np.random.seed(25)
    p3 = mcmc.flat_seed(ps3, k)
    s = mcmc.DeviceEnsembleSampler(k, f3.ndim, f3, seed=25)
    s.run_mcmc(p3, 200, store=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run_mcmc(None, steps, store=False, return_tensor=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('nbins %2d walkers %5d: %.2f us/step acc %.3f' % (nb, k, dt / steps * 1e6, float(np.mean(s.acceptance_fraction))), flush=True)
for nb in (1, 4, 8, 20, 40):
    run(nb, 4096)
for k in (64, 256, 1024, 4096):
    run(20, k)

"""Developer timing of the device-resident ensemble sampler on the config-2 / config-3 shapes for every launch shape."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, models
from golemflavor_b200 import llh, mcmc
from golemflavor_b200.enums import Texture
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
a2, as2, ps2 = models.notebook_model(g['asimov_angles'])
f2 = llh.LnProb(a2, as2, ps2)
a3, as3, ps3 = models.bsm_model_c3(g['asimov_angles'], dim=6, texture=Texture.OET)
f3 = llh.LnProb(a3, as3, ps3)
np.random.seed(25)
p2 = mcmc.flat_seed(ps2, 1024)
p2[:, 4], p2[:, 5] = np.random.uniform(.9, 1, 1024), np.random.uniform(.8, 1, 1024)
p3 = mcmc.flat_seed(ps3, 4096)
def run(name, fn, p0, k, steps, mode, nc):
    s = mcmc.DeviceEnsembleSampler(k, fn.ndim, fn, seed=25, mode=mode, cluster_blocks=nc)
    try:
        s.run_mcmc(p0, 100, store=False)
    except ValueError as e:
        print('%-4s mode %d nc %2d: %s' % (name, mode, nc, str(e)[:80])); return
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run_mcmc(None, steps, store=False, return_tensor=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('%-4s mode %d nc %2d: %.4f s  %.2f us/step  acc %.3f' % (name, mode, nc, dt, dt / steps * 1e6, float(np.mean(s.acceptance_fraction))))
for mode, nc in ((1, 0), (2, 0), (3, 0), (3, 2), (3, 4), (3, 8), (3, 16)):
    run('C2', f2, p2, 1024, 10000, mode, nc)
for mode, nc in ((1, 0), (3, 0), (3, 16)):
    run('C3', f3, p3, 4096, 2000, mode, nc)

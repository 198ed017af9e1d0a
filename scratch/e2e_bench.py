"""e2e rate of the host pipeline (pinned host theta -> lnprob on the host) for the bench workload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, models
from golemflavor_b200 import _lib, llh
from golemflavor_b200.enums import Texture
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
args, asimov, pset = models.bsm_model_c3(g['asimov_angles'], dim=6, texture=Texture.OET)
fn = llh.LnProb(args, asimov, pset)
n = 1 << 22
th = torch.as_tensor(models.draw_in_ranges(pset, n, np.random.default_rng(25))).pin_memory()
out = torch.empty(n, dtype=torch.float64).pin_memory()
a, b = th.numpy(), out.numpy()
for _ in range(3): fn.evaluate_host(a, out=b)
t0 = time.perf_counter()
for _ in range(10): fn.evaluate_host(a, out=b)
dt = (time.perf_counter() - t0) / 10
print('%s  e2e %.4g evals/s  (%.3f ms, %.1f GB/s in)' % (os.path.basename(_lib.LIB_PATH), n / dt, dt * 1e3, n * 56 / dt / 1e9))

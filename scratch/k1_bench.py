"""Developer timing of the SM-only log-posterior (K1) on 2^24 points, row-major and SoA."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, models
from golemflavor_b200 import _lib, llh
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
args, asimov, pset = models.notebook_model(g['asimov_angles'])
fn = llh.LnProb(args, asimov, pset)
lib = _lib.load()
n = 1 << 24
th = torch.as_tensor(models.draw_in_ranges(pset, 1 << 20, np.random.default_rng(3))).cuda().repeat(16, 1)
soa = th.t().contiguous()
out = torch.empty(n, dtype=torch.float64, device='cuda')
for name, t, ldp, ldd in (('row-major', th, 6, 1), ('soa', soa, 1, n)):
    run = lambda: _lib.check(lib.gf_lnprob(fn.model.ref, _lib.ptr(t), n, ldp, ldd, _lib.ptr(out), None, None, None))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('%-9s %.3f ms  %.3g evals/s  %.0f GB/s' % (name, ms, n / ms * 1e3, 56 * n / ms / 1e6))

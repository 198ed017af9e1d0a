"""Developer timing of the BSM log-posterior kernel (K2) on 2^22 points; A/B via GOLEMFLAVOR_B200_LIB."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, models
from golemflavor_b200 import _lib, llh
from golemflavor_b200.enums import Texture
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
args, asimov, pset = models.bsm_model_c3(g['asimov_angles'], dim=6, texture=Texture.OET)
fn = llh.LnProb(args, asimov, pset)
lib = _lib.load()
n = 1 << 22
th = torch.as_tensor(models.draw_in_ranges(pset, n, np.random.default_rng(25))).cuda()
out = torch.empty(n, dtype=torch.float64, device='cuda')
run = lambda: _lib.check(lib.gf_lnprob(fn.model.ref, _lib.ptr(th), n, fn.ndim, 1, _lib.ptr(out), None, None, None))
for _ in range(5): run()
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): run()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 200)
o = out.cpu().numpy()
print('%s  K2 %.4f ms  %.4g evals/s  checksum %.12e finite %.4f' % (os.path.basename(_lib.LIB_PATH), best, n / best * 1e3, float(np.nansum(o[np.isfinite(o)])), np.isfinite(o).mean()))
if len(sys.argv) > 1:
    ref = sys.argv[1]
    if os.path.exists(ref):
        r = np.load(ref)
        both = np.isfinite(r) & np.isfinite(o)
        print('   vs %s: max rel diff %.3e, finite mismatch %d' % (ref, float(np.max(np.abs(o[both] - r[both]) / np.abs(r[both]))), int((np.isfinite(r) != np.isfinite(o)).sum())))
    else:
        np.save(ref, o)

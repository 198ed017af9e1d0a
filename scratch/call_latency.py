"""Per-call latency of the Python entry point for emcee-sized batches (host numpy in, numpy out)."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, models
from golemflavor_b200 import llh
from golemflavor_b200.enums import Texture
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
a2, as2, ps2 = models.notebook_model(g['asimov_angles'])
f2 = llh.LnProb(a2, as2, ps2)
a3, as3, ps3 = models.bsm_model_c3(g['asimov_angles'], dim=6, texture=Texture.OET)
f3 = llh.LnProb(a3, as3, ps3)
rng = np.random.default_rng(1)
for name, fn, ps, n in (('SM 512', f2, ps2, 512), ('BSM 2048', f3, ps3, 2048), ('SM 50', f2, ps2, 50)):
    th = models.draw_in_ranges(ps, n, rng, seeds=True)
    for _ in range(20): fn(th)
    t0 = time.perf_counter()
    for _ in range(2000): fn(th)
    dt = (time.perf_counter() - t0) / 2000
    print('%-9s %.1f us per call' % (name, dt * 1e6))
th = models.draw_in_ranges(ps2, 512, rng, seeds=True)
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): f2(th)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
# device tensors in, device tensors out: torch.ops.golemflavor.lnprob against the ctypes binding of the same entry point
from golemflavor_b200 import _lib
lib = _lib.load()
tht = torch.as_tensor(th).cuda()
out = torch.empty(512, dtype=torch.float64, device='cuda')
for label, call in (('torch.ops', lambda: f2.evaluate(tht)),
                    ('ctypes   ', lambda: _lib.check(lib.gf_lnprob(f2.model.ref, _lib.ptr(tht), 512, 6, 1, _lib.ptr(out), None, None, _lib.stream_ptr(torch))))):
    for _ in range(50): call()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5000): call()
    torch.cuda.synchronize()
    print('device tensor call through %s: %.1f us' % (label, (time.perf_counter() - t0) / 5000 * 1e6))

"""Developer timing of the SM-only log-posterior (K1) on 2^24 points, row-major and SoA: `k1_bench.py [saved.npy]`
(GOLEMFLAVOR_B200_LIB selects a library variant; the first run saves 2^20 outputs, later runs compare bit for bit)."""
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
    o = out.cpu().numpy()
    print('%-9s %.4f ms  %.4g evals/s  %.0f GB/s  checksum %.12e finite %.4f' % (name, ms, n / ms * 1e3, 56 * n / ms / 1e6, float(np.where(np.isfinite(o), o, 0.0).sum()), float(np.isfinite(o).mean())))
    if len(sys.argv) > 1 and name == 'row-major':      # bit-compare against (or create) a saved output: A/B of library variants
        if os.path.exists(sys.argv[1]):
            ref = np.load(sys.argv[1])
            print('   vs %s: identical = %s' % (sys.argv[1], bool(np.array_equal(ref, o[: len(ref)], equal_nan=True))))
        else:
            np.save(sys.argv[1], o[: 1 << 20])

"""Parity at scale (run on the GPU box): CUDA flux-averaged compositions vs the scaled-LAPACK truth of
oracle/truth.py on 2^18 points per (texture, dimension), status-bit statistics included.
Writes gpurun_out/parity_scale.json (copied to profiles/ by hand)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import models
from golemflavor_b200 import _lib, llh, model
from golemflavor_b200.enums import Texture
from oracle import truth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
rng = np.random.default_rng(2026)
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
rows = []
for tex in ('OET', 'OUT', 'OEU', 'NONE'):
    for dim in (3, 4, 5, 6, 7, 8):
        if tex == 'NONE':
            pset = models.bsm11_paramset(dim)
            theta = models.draw_in_ranges(pset, n, rng)
            theta[:, 6:9] = rng.uniform(0, 1, (n, 3))
            theta[:, 9] = rng.uniform(0, 2 * np.pi, n)
            np_ang = theta[:, 6:10]
            args = models.bsm_args(dim, Texture.NONE)
            ll = theta[:, 10]
        else:
            pset = models.bsm7_paramset(dim)
            theta = models.draw_in_ranges(pset, n, rng)
            np_ang = np.broadcast_to(model.TEXTURE_ANGLES[tex], (n, 4))
            args = models.bsm_args(dim, Texture[tex])
            ll = theta[:, 6]
        args.injected_ratio, args.smearing = [1 / 3, 1 / 3, 1 / 3], 0.02
        fn = llh.LnProb(args, None, pset)
        t0 = time.perf_counter()
        lnp, fr, st = (x.cpu().numpy() for x in fn.evaluate(theta, want_fr=True, want_status=True))
        ref = truth.eigh_flux_averaged_fr(theta[:, :4], theta[:, 4:6], np_ang, ll, dim, models.BINNING, args.source_ratio)
        err = np.abs(fr - ref).max(axis=1)
        rows.append(dict(texture=tex, dimension=dim, points=n, max_abs_err=float(err.max()), p999=float(np.quantile(err, 0.999)),
                         median=float(np.median(err)), refined_points=int(np.count_nonzero(st & _lib.ST_REFINED)),
                         ill_cond=int(np.count_nonzero(st & _lib.ST_ILL_COND)), non_unitary=int(np.count_nonzero(st & _lib.ST_NON_UNITARY)),
                         non_finite=int(np.count_nonzero(st & _lib.ST_NON_FINITE)), sum_err=float(np.abs(fr.sum(1) - 1).max()),
                         seconds=time.perf_counter() - t0))
        print(rows[-1], flush=True)
out = dict(tolerance=1e-10, worst=max(r['max_abs_err'] for r in rows), rows=rows)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'parity_scale.json'), 'w'), indent=1)
print('WORST', out['worst'])

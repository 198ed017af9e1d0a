#!/usr/bin/env python
"""Golden vectors for the Bayes-factor limit extraction ``plot.get_limit`` (``golemflavor/plot.py:149-213``),
generated from the UNMODIFIED reference function (build container only).

``golemflavor/plot.py`` imports a plotting stack that is absent here (matplotlib, getdist, python-ternary, shapely);
``get_limit`` itself uses NumPy and ``scipy.interpolate`` only, so those modules are replaced by inert stand-ins FOR THE
IMPORT, nothing in the reference is patched.  Output: ``ref_limit.npz`` -- evidence curves (scale grid of
``scripts/sens.py:199-201`` + lnZ), the limit the reference returns (NaN for ``None``, +inf for 'Discovered LV!') and
the splined reduced-evidence curves of ``return_interp=True``.
"""
import collections
import collections.abc
import contextlib
import fractions
import io
import math
import os
import sys
from argparse import Namespace
from unittest import mock

import numpy as np

fractions.gcd = math.gcd
collections.Sequence = collections.abc.Sequence
sys.path.insert(0, os.environ.get('GOLEM_REFERENCE', '/root/reference'))
for name in ('matplotlib', 'matplotlib.patches', 'matplotlib.gridspec', 'matplotlib.pyplot', 'matplotlib.offsetbox', 'matplotlib.lines',
             'matplotlib.colors', 'matplotlib.cm', 'mpl_toolkits', 'mpl_toolkits.axes_grid1', 'mpl_toolkits.mplot3d', 'getdist', 'getdist.plots',
             'getdist.mcsamples', 'ternary', 'ternary.heatmapping', 'shapely', 'shapely.geometry', 'shapely.ops'):
    sys.modules.setdefault(name, mock.MagicMock())
import scipy.ndimage  # noqa: E402
if not hasattr(scipy.ndimage, 'filters'):
    sys.modules['scipy.ndimage.filters'] = scipy.ndimage

import golemflavor.plot as rplot  # noqa: E402
from golemflavor.enums import StatCateg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SCALE_BOUNDARIES = {3: (-32, -20), 4: (-40, -24), 5: (-48, -27), 6: (-56, -30), 7: (-64, -33), 8: (-72, -36)}


def curve(rng, dim, segments, kind):
    lo, hi = SCALE_BOUNDARIES[dim]
    scales = np.concatenate([[-100.0], np.linspace(lo, hi, segments - 1)])
    x = (scales - lo) / (hi - lo)
    null = rng.uniform(-330, -300)
    noise = rng.normal(0, 0.03, len(scales))
    if kind == 'sigmoid':      # evidence drops once the operator becomes visible: a limit exists
        mid, width, depth = rng.uniform(0.3, 0.7), rng.uniform(0.03, 0.1), rng.uniform(4, 40)
        stat = null - depth / (1 + np.exp(-(x - mid) / width)) + noise
    elif kind == 'flat':       # no sensitivity
        stat = null + noise
    elif kind == 'peaked':     # disfavoured in a window only
        mid, width, depth = rng.uniform(0.3, 0.6), rng.uniform(0.03, 0.06), rng.uniform(4, 10)
        stat = null - depth * np.exp(-0.5 * ((x - mid) / width) ** 2) + noise
    elif kind == 'edge':       # only the last grid point crosses the threshold
        stat = null + noise
        stat[-1] -= rng.uniform(3, 6)
    else:                      # 'discovery': the null point is disfavoured
        stat = null + rng.uniform(3, 8) / (1 + np.exp(-(x - 0.5) / 0.05)) + noise
    stat[0] = null + noise[0] * (kind != 'discovery')
    return scales, stat


def main():
    rng = np.random.default_rng(27)
    args = Namespace(stat_method=StatCateg.BAYESIAN, dimension=6, source_ratio=(1, 2, 0))
    rows = []
    for kind in ('sigmoid',) * 10 + ('flat',) * 2 + ('peaked',) * 3 + ('edge',) * 2 + ('discovery',) * 2:
        dim = int(rng.integers(3, 9))
        segments = int(rng.choice([10, 20, 100]))
        args.dimension = dim
        scales, stat = curve(rng, dim, segments, kind)
        for mask_initial in (False, True):
            with contextlib.redirect_stdout(io.StringIO()):
                try:
                    lim = rplot.get_limit(scales.copy(), stat.copy(), args, mask_initial=mask_initial)
                    lim = np.nan if lim is None else float(lim)
                    interp = rplot.get_limit(scales.copy(), stat.copy(), args, mask_initial=mask_initial, return_interp=True)
                except AssertionError:
                    lim, interp = np.inf, None
            rows.append((kind, dim, mask_initial, scales, stat, lim, interp))
    out = {'n': len(rows), 'bayes_k': rplot.BAYES_K}
    for k, (kind, dim, mi, sc, st, lim, interp) in enumerate(rows):
        out['kind_%d' % k], out['dim_%d' % k], out['mask_%d' % k] = kind, dim, mi
        out['scales_%d' % k], out['stat_%d' % k], out['limit_%d' % k] = sc, st, lim
        if interp is not None:
            out['interp_sc_%d' % k], out['interp_ev_%d' % k] = np.asarray(interp[0]), np.asarray(interp[1])
    np.savez_compressed(os.path.join(HERE, "ref_limit.npz"), **out)
    lims = np.array([r[5] for r in rows])
    print('ref_limit.npz: %d curves, %d limits, %d None, %d discoveries' % (len(rows), np.isfinite(lims).sum(), np.isnan(lims).sum(), np.isposinf(lims).sum()))


if __name__ == '__main__':
    main()

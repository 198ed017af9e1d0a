#!/usr/bin/env python
"""Round-2 golden fixtures, generated from the UNMODIFIED reference (same two-line shim as
``make_golden.py``; build container only): the models with a SAMPLED SOURCE.

``ref_src.npz``
  * ``c1_*``  BASELINE config 1 -- three raw source flavor ratios, PMNS fixed at ``fr.NUFIT_U``
    (``fr.py:313``), composed like the tutorial model (``examples/tutorial.ipynb`` cells 13-15:
    ``lnprior`` + ``multi_gaussian``) with ``u_to_fr(theta, NUFIT_U)`` as the measured composition
    (``fr.py:502-536`` normalises the raw ratios).
  * ``sb_*``  sampled source on the BSM path -- the composition of ``llh.py:94-112``
    (``flux_averaged_BSMu`` with ``args.source_ratio`` taken from the SRCANGLES values) with
    ``multi_gaussian`` in place of GolemFit, for the three source parametrisations the product
    accepts: two angles (``angles_to_fr``, ``fr.py:82-113``), x with source (x, 1-x, 0)
    (``scripts/mc_x.py:187``) and three raw ratios.
  * ``nb_*``  the ``args.no_bsm`` branch of ``flux_averaged_BSMu`` (``fr.py:437-438``): the unmodified
    reference RAISES there (2-D source table into ``u_to_fr``'s einsum); recorded as such, together
    with ``u_to_fr(source_ratio, sm_u)`` -- the evident intent, which the product and the oracle compute.
"""

import argparse
import collections
import collections.abc
import fractions
import math
import os
import sys
from copy import deepcopy

import numpy as np

fractions.gcd = math.gcd                       # golemflavor/misc.py:15
collections.Sequence = collections.abc.Sequence  # golemflavor/param.py:15
sys.path.insert(0, os.environ.get('GOLEM_REFERENCE', '/root/reference'))

import golemflavor.fr as rfr          # noqa: E402
import golemflavor.llh as rllh        # noqa: E402
from golemflavor.enums import ParamTag, Texture  # noqa: E402
from golemflavor.param import Param, ParamSet  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import BINNING, TEX, bsm_paramset  # noqa: E402  (the round-1 helpers, unchanged)

SMEARING = 0.02


def injected():
    measured = rfr.u_to_fr(rfr.normalize_fr((1, 0, 0)), rfr.NUFIT_U)
    return np.asarray(measured, dtype=np.float64)


def gen_c1(rng, bf):
    tag = ParamTag.SRCANGLES
    pset = ParamSet([Param(name='f_%s' % n, value=1. / 3, ranges=[1e-6, 1.], tag=tag) for n in ('e', 'mu', 'tau')])
    n = 200
    theta = rng.uniform(1e-6, 1.0, size=(n, 3))
    theta[:60] = np.array([1.0, 0.0, 0.0]) * rng.uniform(0.3, 1.0, (60, 1)) + rng.uniform(1e-6, 0.05, (60, 3))  # near the injection
    theta[60:80, 1] = rng.uniform(-0.1, 1.1, 20)                                                             # some outside the box
    lnp, fr = [], []
    for t in theta:
        lp = rllh.lnprior(list(t), paramset=deepcopy(pset))
        if not np.isfinite(lp):
            lnp.append(-np.inf)
            fr.append(np.full(3, np.nan))
            continue
        f = np.asarray(rfr.u_to_fr(list(t), rfr.NUFIT_U), dtype=np.float64)
        with np.errstate(divide='ignore'):
            lnp.append(float(lp + rllh.multi_gaussian(f, bf, SMEARING)))
        fr.append(f)
    return dict(c1_theta=theta, c1_lnprob=np.array(lnp), c1_fr=np.array(fr))


def gen_sampled_source_bsm(rng, bf):
    rows = []
    for kind in ('angles', 'x', 'ratios'):
        for tex, dim in (('OET', 6), ('OUT', 6), ('OEU', 3), ('OET', 8)):
            lo, hi = rfr.SCALE_BOUNDARIES[dim]
            for _ in range(6):
                sm = [rng.uniform(0.26, 0.35), rng.uniform(0.950, 0.961), rng.uniform(0.31, 0.75), rng.uniform(0, 2 * np.pi),
                      rng.uniform(7.2e-23, 7.6e-23), rng.uniform(2.46e-21, 2.53e-21)]
                loglam = rng.uniform(lo, hi)
                if kind == 'angles':
                    srcv = [rng.uniform(0, 1), rng.uniform(-1, 1)]
                    source = rfr.angles_to_fr(srcv)
                elif kind == 'x':
                    srcv = [rng.uniform(0, 1)]
                    source = (srcv[0], 1.0 - srcv[0], 0.0)
                else:
                    srcv = list(rng.uniform(0.01, 1, 3))
                    source = tuple(srcv)
                theta_ref = sm + list(TEX[tex]) + [loglam]
                args = argparse.Namespace(binning=BINNING, source_ratio=np.array(source, dtype=np.float64), dimension=dim,
                                          texture=Texture.NONE, no_bsm=False)
                ok = True
                try:
                    fr = np.asarray(rfr.flux_averaged_BSMu(theta_ref, args, -2.0, bsm_paramset(dim, TEX[tex])), dtype=np.float64)
                    with np.errstate(divide='ignore'):
                        llh = float(rllh.multi_gaussian(fr, bf, SMEARING))
                except AssertionError:      # the reference's own unitarity assertion (fr.py:493-498)
                    ok, fr, llh = False, np.full(3, np.nan), np.nan
                rows.append(dict(kind=kind, tex=tex, dim=dim, sm=sm, src=srcv + [np.nan] * (3 - len(srcv)), loglam=loglam, ok=ok, fr=fr, llh=llh))
    return dict(sb_kind=np.array([r['kind'] for r in rows]), sb_tex=np.array([r['tex'] for r in rows]),
                sb_dim=np.array([r['dim'] for r in rows]), sb_sm=np.array([r['sm'] for r in rows]),
                sb_src=np.array([r['src'] for r in rows]), sb_loglam=np.array([r['loglam'] for r in rows]),
                sb_ok=np.array([r['ok'] for r in rows]), sb_fr=np.array([r['fr'] for r in rows]), sb_llh=np.array([r['llh'] for r in rows]))


def gen_no_bsm(rng):
    tag = ParamTag.SCALE
    pset = ParamSet([Param(name='logLam', value=-40., ranges=[-56, -30], std=3, tag=tag)])
    srcs = np.array([(1, 2, 0), (1, 0, 0), (0, 1, 0), (0.2, 0.5, 0.3)], dtype=np.float64)
    raised, intent = [], []
    for s in srcs:
        args = argparse.Namespace(binning=BINNING, source_ratio=rfr.normalize_fr(s), dimension=6, texture=Texture.NONE, no_bsm=True)
        try:
            rfr.flux_averaged_BSMu([-40.], args, -2.0, deepcopy(pset))
            raised.append('')
        except Exception as exc:  # noqa: BLE001
            raised.append(type(exc).__name__)
        intent.append(np.asarray(rfr.u_to_fr(rfr.normalize_fr(s), rfr.NUFIT_U), dtype=np.float64))
    return dict(nb_src=srcs, nb_raised=np.array(raised), nb_fr=np.array(intent))


def main():
    rng = np.random.default_rng(26)
    bf = injected()
    out = dict(bf=bf, smearing=SMEARING, binning=BINNING)
    out.update(gen_c1(rng, bf))
    out.update(gen_sampled_source_bsm(rng, bf))
    out.update(gen_no_bsm(rng))
    np.savez(os.path.join(HERE, 'ref_src.npz'), **out)
    print('ref_src.npz written:', {k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()

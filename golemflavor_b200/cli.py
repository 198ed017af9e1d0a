"""Command-line front ends with the flags, file names and ``.npy`` layouts of the reference scripts
(``scripts/mc_unitary.py``, ``mc_x.py``, ``mc_texture.py``, ``fr.py``), running on the CUDA path.

    python -m golemflavor_b200.cli mc_unitary --source-ratio 1 2 0 --nwalkers 60 --nsteps 1000 --datadir out
    python -m golemflavor_b200.cli mc_texture --dimension 6 --texture OET --datadir out
    python -m golemflavor_b200.cli fr --dimension 6 --texture OET --injected-ratio 1 1 1 --datadir out

The MC-scan scripts of the reference obtain ``nwalkers * nsteps`` prior samples by running emcee on a
flat likelihood; here the same number of samples is drawn directly (Philox, ``scan.scan_samples``) and
saved in the same column layout: ``frs`` for ``mc_unitary`` / ``mc_x`` (``mc_unitary.py:189-193``),
``(frs, samples)`` side by side for ``mc_texture`` (``mc_texture.py:216-223``).  ``fr`` runs the BSM
emcee fit of ``scripts/fr.py:162-242`` with the Gaussian flavor-ratio likelihood.
"""

import argparse
import os
from functools import reduce

import numpy as np

from . import llh, mcmc, scan
from . import model as _model
from .enums import ParamTag, Texture, enum_name
from .param import Param, ParamSet

__all__ = ['solve_ratio', 'gen_identifier', 'main']


def solve_ratio(fr):
    """'1_2_0'-style tag of a flavor ratio (``misc.py:34-41``)."""
    fr = [float(x) for x in fr]

    def fgcd(a, b):   # Euclid on floats, as Python 2's fractions.gcd did for the reference
        while abs(b) > 1e-9:
            a, b = b, a % b
        return a
    den = reduce(fgcd, fr) or 1.0
    f = [int(round(x / den)) for x in fr]
    if any(x not in (1, 2, 0) for x in f) or any(abs(x / den - round(x / den)) > 1e-6 for x in fr):
        return '{0:.2f}_{1:.2f}_{2:.2f}'.format(*fr)
    return '{0}_{1}_{2}'.format(*f)


def gen_identifier(args, kind):
    """Output-file identifiers of the reference (``mc_unitary.py:110-112``, ``mc_texture.py:140-144``,
    ``misc.py:44-51``)."""
    if kind in ('mc_unitary', 'mc_x'):
        return '_SRC_{0}'.format(solve_ratio(args.source_ratio)) if kind == 'mc_unitary' else ''
    if kind == 'mc_texture':
        return '_DIM{0}_SRC_{1}_{2}'.format(args.dimension, solve_ratio(args.source_ratio), enum_name(args.texture))
    f = '_DIM{0}_sfr_{1}'.format(args.dimension, solve_ratio(args.source_ratio))
    if getattr(args, 'injected_ratio', None) is not None:
        f += '_mfr_' + solve_ratio(args.injected_ratio)
    if enum_name(args.texture) != 'NONE':
        f += '_' + enum_name(args.texture)
    return f


def _texture(s):
    return Texture[s.upper()]


def _parser():
    p = argparse.ArgumentParser(prog='golemflavor_b200.cli', description='BSM flavor ratio analysis (B200 path)')
    sub = p.add_subparsers(dest='command', required=True)
    for name in ('mc_unitary', 'mc_x', 'mc_texture', 'fr'):
        s = sub.add_parser(name)
        s.add_argument('--source-ratio', type=float, nargs=3, default=[1, 2, 0])       # fr.py:266-269
        s.add_argument('--seed', type=int, default=26 if name.startswith('mc_') else 25)
        s.add_argument('--threads', default='1')                                      # accepted, unused
        s.add_argument('--datadir', type=str, default='./untitled')
        s.add_argument('--burnin', type=int, default=100)                              # mcmc.py:61-64
        s.add_argument('--nwalkers', type=int, default=60)
        s.add_argument('--nsteps', type=int, default=2000)
        s.add_argument('--run-mcmc', type=str, default='True')
        if name in ('mc_texture', 'fr'):
            s.add_argument('--dimension', type=int, default=3)                         # fr.py:274-277
            s.add_argument('--texture', type=_texture, default=Texture.NONE if name == 'fr' else Texture.OET)
            s.add_argument('--binning', type=float, nargs=3, default=[6e4, 1e7, 20])   # fr.py:282-285
            s.add_argument('--spectral-index', type=float, default=-2.0)
        if name == 'fr':
            s.add_argument('--injected-ratio', type=float, nargs=3, default=[1, 1, 1])
            s.add_argument('--smearing', type=float, default=0.02)
            s.add_argument('--outfile', type=str, default=None)
    return p


def _save(arr, path):
    mcmc.save_chains(arr, path)
    return path + '.npy'


def main(argv=None):
    args = _parser().parse_args(argv)
    args.source_ratio = np.asarray(args.source_ratio, dtype=np.float64) / np.sum(args.source_ratio)
    n = args.nwalkers * args.nsteps
    if args.command in ('mc_unitary', 'mc_x'):
        fm = scan.scan_model('unitary' if args.command == 'mc_unitary' else 'x', source_ratio=args.source_ratio)
        _, frs, _ = scan.scan_samples(fm, n, seed=args.seed)
        return _save(frs, os.path.join(args.datadir, args.command + gen_identifier(args, args.command)))
    if args.command == 'mc_texture':
        fm = scan.scan_model('texture', source_ratio=args.source_ratio, dimension=args.dimension, texture=args.texture,
                             binning=_model.binning_edges(args.binning))
        theta, frs, _ = scan.scan_samples(fm, n, seed=args.seed)
        return _save(np.hstack([frs, theta]), os.path.join(args.datadir, 'mc_texture' + gen_identifier(args, 'mc_texture')))
    # fr: BSM emcee fit (scripts/fr.py:62-104 parameter set without the GolemFit nuisance block)
    args.binning = _model.binning_edges(args.binning)
    args.no_bsm = False
    ps = scan.sm_paramset(with_mass=True)
    if enum_name(args.texture) == 'NONE':
        ps += [Param(name=nm, value=0.5, ranges=[0., 1.], std=0.2, tag=ParamTag.MMANGLES)
               for nm in ('np_s_12_2', 'np_c_13_4', 'np_s_23_2')]
        ps += [Param(name='np_dcp', value=np.pi, ranges=[0., 2 * np.pi], std=0.2, tag=ParamTag.MMANGLES)]
    b = _model.SCALE_BOUNDARIES[args.dimension]
    ps.append(Param(name='logLam', value=float(np.mean(b)), ranges=list(b), std=3, tag=ParamTag.SCALE))
    pset = ParamSet(ps)
    fn = llh.LnProb(args, None, pset)
    np.random.seed(args.seed)
    p0 = mcmc.flat_seed(pset, args.nwalkers)
    samples = mcmc.mcmc(p0, fn, len(pset), args.nwalkers, args.burnin, args.nsteps, seed=args.seed)
    out = args.outfile or os.path.join(args.datadir, 'chain')
    return _save(samples, out + gen_identifier(args, 'fr'))


if __name__ == '__main__':
    print(main())

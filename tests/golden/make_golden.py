#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  The reference is imported through
the two-line Python-3.12 shim (``fractions.gcd`` / ``collections.Sequence``),
nothing else is patched.  Outputs are small ``.npz`` files (float64 casts of the
reference's float128 results) committed next to this script.

Work-arounds for reference defects that are *not* on the arithmetic path
(SURVEY.md section 8c): fixed textures are requested as ``Texture.NONE`` with the
explicit angle tuple of ``fr.py:370-376`` (the texture branch builds a ragged
array under NumPy >= 1.24), and ``np.logspace`` gets an integer bin count.
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
from golemflavor.enums import ParamTag, PriorsCateg, Texture  # noqa: E402
from golemflavor.param import Param, ParamSet  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import truth  # noqa: E402  (mpmath truth only; not the restatement)

Z = 0. + 1e-9
TEX = {'OEU': (0.5, 1.0, Z, Z), 'OET': (Z, 0.25, Z, Z), 'OUT': (Z, 1.0, 0.5, Z)}
BINNING = np.logspace(np.log10(6e4), np.log10(1e7), 20 + 1)


def c128(x):
    return np.asarray(x).astype(np.complex128)


def sm_nuisance(with_mass=True, lg=True):
    """scripts/fr.py:30-49 / examples/inference.ipynb cell 17."""
    tag = ParamTag.SM_ANGLES
    lgp = PriorsCateg.LIMITEDGAUSS if lg else None
    e = 1e-9
    out = [
        Param(name='s_12_2', value=0.307, seed=[0.26, 0.35], ranges=[0., 1.], std=0.013, prior=lgp, tag=tag),
        Param(name='c_13_4', value=(1 - 0.02206) ** 2, seed=[0.950, 0.961], ranges=[0., 1.], std=0.00147, prior=lgp, tag=tag),
        Param(name='s_23_2', value=0.538, seed=[0.31, 0.75], ranges=[0., 1.], std=0.069, prior=lgp, tag=tag),
        Param(name='dcp', value=4.08404, seed=[0 + e, 2 * np.pi - e], ranges=[0., 2 * np.pi], std=2.0, tag=tag),
    ]
    if with_mass:
        g = PriorsCateg.GAUSSIAN
        out += [
            Param(name='m21_2', value=7.40E-23, seed=[7.2E-23, 7.6E-23], ranges=[6.80E-23, 8.02E-23], std=2.1E-24, prior=g, tag=tag),
            Param(name='m3x_2', value=2.494E-21, seed=[2.46E-21, 2.53E-21], ranges=[2.399E-21, 2.593E-21], std=3.3E-23, prior=g, tag=tag),
        ]
    return out


def gen_basic(rng):
    n = 256
    ang = np.column_stack([rng.uniform(0, 1, n), rng.uniform(0, 1, n),
                           rng.uniform(0, 1, n), rng.uniform(0, 2 * np.pi, n)])
    ang[0] = (0.2, 0.3, 0.5, 1.5)
    ang[1] = (0.307, (1 - 0.02195) ** 2, 0.565, 3.97935)
    u = np.array([c128(rfr.angles_to_u(a)) for a in ang])
    src_ang = np.column_stack([rng.uniform(0, 1, n), rng.uniform(-1, 1, n)])
    src_ang[0] = (0.3, 0.4)
    src_fr = np.array([rfr.angles_to_fr(a) for a in src_ang])
    back = np.array([[float(x) for x in rfr.fr_to_angles(f)] for f in src_fr])
    srcs = rng.uniform(0, 3, (n, 3))
    srcs[0] = (1, 2, 0)
    srcs[1] = (1, 0, 0)
    srcs[2] = (0, 1, 0)
    fr = np.array([np.asarray(rfr.u_to_fr(s, rfr.angles_to_u(a)), dtype=np.float64)
                   for s, a in zip(srcs, ang)])
    fr_nufit = np.array([np.asarray(rfr.u_to_fr(rfr.normalize_fr(s), rfr.NUFIT_U), dtype=np.float64)
                         for s in [(1, 0, 0), (0, 1, 0), (1, 2, 0)]])
    herm = []
    vecs = []
    for _ in range(64):
        a = rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3))
        h = (a + a.conj().T) / 2
        herm.append(h)
        vecs.append(c128(rfr.cardano_eqn(np.array(h, dtype=np.complex256))))
    np.savez(os.path.join(HERE, 'ref_basic.npz'),
             ang=ang, u=u, src_ang=src_ang, src_fr=src_fr, src_back=back,
             srcs=srcs, fr=fr, nufit_u=c128(rfr.NUFIT_U), fr_nufit=fr_nufit,
             herm=np.array(herm), herm_vecs=np.array(vecs))


def gen_bsm_u(rng):
    """params_to_BSMu / u_to_fr for single energies."""
    rows = []
    for tex in ['OEU', 'OET', 'OUT', 'NONE']:
        for dim in range(3, 9):
            lo, hi = rfr.SCALE_BOUNDARIES[dim]
            for _ in range(12):
                if tex == 'NONE':
                    npang = (rng.uniform(0, 1), rng.uniform(0, 1), rng.uniform(0, 1),
                             rng.uniform(0, 2 * np.pi))
                else:
                    npang = TEX[tex]
                loglam = rng.uniform(lo, hi)
                energy = 10 ** rng.uniform(np.log10(6e4), 7)
                sm = (rng.normal(0.307, 0.013), rng.normal((1 - 0.02206) ** 2, 0.00147),
                      rng.normal(0.538, 0.069), rng.uniform(0, 2 * np.pi))
                mass = (rng.normal(7.40e-23, 2.1e-24), rng.normal(2.494e-21, 3.3e-23))
                src = [(1, 2, 0), (1, 0, 0), (0, 1, 0)][rng.integers(3)]
                sm_u = rfr.angles_to_u(sm)
                v = rfr.params_to_BSMu(tuple(npang) + (loglam,), dim, energy,
                                       mass_eigenvalues=list(mass), sm_u=sm_u,
                                       texture=Texture.NONE, check_uni=False)
                resid_f = rfr.test_unitarity(v)
                resid = float(max(abs(np.trace(resid_f) - 3), abs(np.sum(resid_f) - 3)))
                fr = np.asarray(rfr.u_to_fr(src, v), dtype=np.float64)
                tr = [float(x) for x in truth.mp_bsm_fr_bin(sm, mass, npang, loglam, dim, energy, src)]
                rows.append(dict(tex=tex, dim=dim, npang=npang, loglam=loglam, energy=energy,
                                 sm=sm, mass=mass, src=src, v=c128(v), resid=resid, fr=fr, fr_mp=tr))
    np.savez(os.path.join(HERE, 'ref_bsm_u.npz'),
             tex=np.array([r['tex'] for r in rows]),
             dim=np.array([r['dim'] for r in rows]),
             npang=np.array([r['npang'] for r in rows]),
             loglam=np.array([r['loglam'] for r in rows]),
             energy=np.array([r['energy'] for r in rows]),
             sm=np.array([r['sm'] for r in rows]),
             mass=np.array([r['mass'] for r in rows]),
             src=np.array([r['src'] for r in rows], dtype=np.float64),
             v=np.array([r['v'] for r in rows]),
             resid=np.array([r['resid'] for r in rows]),
             fr=np.array([r['fr'] for r in rows]),
             fr_mp=np.array([r['fr_mp'] for r in rows]))


def bsm_paramset(dim, npang):
    """6 SM params + 4 MMANGLES (fixed-texture values) + logLam, MMANGLES before
    SCALE so that from_tag yields (np angles..., logLam) (fr.py:421-423)."""
    ps = sm_nuisance(with_mass=True)
    for k, nm in enumerate(['np_s12', 'np_c13', 'np_s23', 'np_dcp']):
        ps.append(Param(name=nm, value=npang[k], ranges=[0., 2 * np.pi], std=0.2, tag=ParamTag.MMANGLES))
    b = rfr.SCALE_BOUNDARIES[dim]
    ps.append(Param(name='logLam', value=np.mean(b), ranges=b, std=3, tag=ParamTag.SCALE))
    return ParamSet(ps)


def gen_flux(rng):
    """flux_averaged_BSMu on the 20-bin production binning."""
    rows = []
    for tex in ['OET', 'OUT', 'OEU', 'NONE']:
        for dim in [3, 6] if tex != 'NONE' else [6]:
            lo, hi = rfr.SCALE_BOUNDARIES[dim]
            for src in [(1, 2, 0), (1, 0, 0), (0, 1, 0)]:
                for _ in range(10):
                    npang = TEX[tex] if tex != 'NONE' else (
                        rng.uniform(0, 1), rng.uniform(0, 1), rng.uniform(0, 1), rng.uniform(0, 2 * np.pi))
                    sm = [rng.uniform(0.26, 0.35), rng.uniform(0.950, 0.961), rng.uniform(0.31, 0.75),
                          rng.uniform(0, 2 * np.pi), rng.uniform(7.2e-23, 7.6e-23),
                          rng.uniform(2.46e-21, 2.53e-21)]
                    loglam = rng.uniform(lo, hi)
                    theta = sm + list(npang) + [loglam]
                    pset = bsm_paramset(dim, npang)
                    args = argparse.Namespace(binning=BINNING, source_ratio=rfr.normalize_fr(src),
                                              dimension=dim, texture=Texture.NONE, no_bsm=False)
                    ok = True
                    try:
                        fr = np.asarray(rfr.flux_averaged_BSMu(theta, args, -2.0, pset), dtype=np.float64)
                    except AssertionError:
                        ok = False
                        fr = np.full(3, np.nan)
                    fr_mp = truth.mp_flux_averaged_fr(sm[:4], sm[4:6], npang, loglam, dim, BINNING,
                                                      rfr.normalize_fr(src))
                    rows.append(dict(tex=tex, dim=dim, src=src, theta=theta, ok=ok, fr=fr, fr_mp=fr_mp))
    # the survey's spot value: dim 6, OET, logLam = -43, nominal SM params, source (1,2,0)
    theta = [0.307, (1 - 0.02206) ** 2, 0.538, 4.08404, 7.40e-23, 2.494e-21] + list(TEX['OET']) + [-43.0]
    args = argparse.Namespace(binning=BINNING, source_ratio=rfr.normalize_fr((1, 2, 0)),
                              dimension=6, texture=Texture.NONE, no_bsm=False)
    for gamma in (-2.5, -2.0, 0.0):
        fr = np.asarray(rfr.flux_averaged_BSMu(theta, args, gamma, bsm_paramset(6, TEX['OET'])), dtype=np.float64)
        fr_mp = truth.mp_flux_averaged_fr(theta[:4], theta[4:6], TEX['OET'], -43.0, 6, BINNING, (1, 2, 0))
        rows.append(dict(tex='OET', dim=6, src=(1, 2, 0), theta=theta, ok=True, fr=fr, fr_mp=fr_mp))
    np.savez(os.path.join(HERE, 'ref_flux.npz'),
             binning=BINNING,
             tex=np.array([r['tex'] for r in rows]),
             dim=np.array([r['dim'] for r in rows]),
             src=np.array([r['src'] for r in rows], dtype=np.float64),
             theta=np.array([r['theta'] for r in rows]),
             ok=np.array([r['ok'] for r in rows]),
             fr=np.array([r['fr'] for r in rows]),
             fr_mp=np.array([r['fr_mp'] for r in rows]))


def notebook_model():
    """examples/inference.ipynb cells 5-23 (6-D SM fit)."""
    source = rfr.normalize_fr((1, 0, 0))
    measured = rfr.u_to_fr(source, rfr.NUFIT_U)
    smearing = 0.02
    angles = rfr.fr_to_angles(measured)
    tag = ParamTag.BESTFIT
    asimov = ParamSet([
        Param(name='measured_angle1', value=angles[0], ranges=[0., 1.], std=smearing, tag=tag),
        Param(name='measured_angle2', value=angles[1], ranges=[-1., 1.], std=smearing, tag=tag)])
    nuis = sm_nuisance(with_mass=False)
    nuis[3] = Param(name='dcp', value=4.08404, seed=[0, 2 * np.pi], ranges=[0., 2 * np.pi], std=2.0,
                    tag=ParamTag.SM_ANGLES)
    tag = ParamTag.SRCANGLES
    src = [Param(name='source_angle1', value=0, ranges=[0., 1.], tag=tag),
           Param(name='source_angle2', value=0, ranges=[-1., 1.], tag=tag)]
    return asimov, ParamSet(nuis + src)


def notebook_triangle_llh(theta, asimov, llh_paramset):
    for idx, p in enumerate(llh_paramset):
        p.value = theta[idx]
    sm_u = rfr.angles_to_u(llh_paramset.from_tag(ParamTag.SM_ANGLES, values=True))
    source = rfr.angles_to_fr(llh_paramset.from_tag(ParamTag.SRCANGLES, values=True))
    measured = rfr.u_to_fr(source, sm_u)
    bf = rfr.angles_to_fr(asimov.from_tag(ParamTag.BESTFIT, values=True))
    return rllh.multi_gaussian(measured, bf, asimov['measured_angle1'].std), measured


def notebook_ln_prob(theta, asimov, llh_paramset):
    lp = rllh.lnprior(theta, paramset=llh_paramset)
    if not np.isfinite(lp):
        return -np.inf, np.full(3, np.nan), lp
    llh, measured = notebook_triangle_llh(theta, asimov, llh_paramset)
    return lp + llh, np.asarray(measured, dtype=np.float64), lp


def gen_llh(rng):
    asimov, pset = notebook_model()
    n = 300
    lo = np.array([p.ranges[0] for p in pset])
    hi = np.array([p.ranges[1] for p in pset])
    seeds = np.array(pset.seeds)
    theta = rng.uniform(seeds[:, 0], seeds[:, 1], size=(n, len(pset)))
    # a third drawn across the full box, a few outside it
    theta[100:200] = rng.uniform(lo, hi, size=(100, len(pset)))
    theta[200:220, 0] = rng.uniform(-0.2, 1.2, 20)
    theta[220:240, 5] = rng.uniform(-1.3, 1.3, 20)
    theta[0] = [0.338311172296449, 0.9564050462153981, 0.4326891339084704, 1.1681147219086683,
                0.4111001279251132, -0.7652489057305689]
    # points close to the injected composition so that the LLH is finite
    theta[240:300, 4] = rng.uniform(0.9, 1.0, 60)
    theta[240:300, 5] = rng.uniform(0.8, 1.0, 60)
    out = [notebook_ln_prob(list(t), deepcopy(asimov), deepcopy(pset)) for t in theta]
    lnp = np.array([o[0] for o in out], dtype=np.float64)
    fr = np.array([o[1] for o in out])
    lp = np.array([o[2] for o in out], dtype=np.float64)

    # lnprior on the 7-D BSM set (3 LG + 1 uniform + 2 G + uniform scale)
    b = rfr.SCALE_BOUNDARIES[6]
    bsm = ParamSet(sm_nuisance(with_mass=True) + [
        Param(name='logLam', value=np.mean(b), ranges=b, std=3, tag=ParamTag.SCALE)])
    seeds7 = np.array(bsm.seeds)
    th7 = rng.uniform(seeds7[:, 0], seeds7[:, 1], size=(200, 7))
    th7[150:, 4] = rng.uniform(6.5e-23, 8.3e-23, 50)
    lp7 = np.array([rllh.lnprior(list(t), deepcopy(bsm)) for t in th7], dtype=np.float64)

    # multi_gaussian incl. the underflow edge
    frs = rng.dirichlet([1, 1, 1], size=200)
    bf = np.asarray(rfr.u_to_fr(rfr.normalize_fr((1, 0, 0)), rfr.NUFIT_U), dtype=np.float64)
    with np.errstate(divide='ignore'):
        mg = np.array([rllh.multi_gaussian(f, bf, 0.02) for f in frs], dtype=np.float64)
        mg_wide = np.array([rllh.multi_gaussian(f, bf, 0.2, offset=0) for f in frs], dtype=np.float64)
    np.savez(os.path.join(HERE, 'ref_llh.npz'),
             theta=theta, lnprob=lnp, fr=fr, lnprior=lp,
             asimov_angles=np.array(asimov.values, dtype=np.float64),
             theta7=th7, lnprior7=lp7,
             mg_fr=frs, mg_bf=bf, mg=mg, mg_wide=mg_wide,
             mg_spot=float(rllh.multi_gaussian([.3, .35, .35], [.55, .18, .27], .02)))


def main():
    rng = np.random.default_rng(25)
    gen_basic(rng)
    gen_llh(rng)
    gen_bsm_u(rng)
    gen_flux(rng)
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()

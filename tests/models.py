"""Model builders shared by the CPU and GPU tests: the reference's parameter sets
(``examples/inference.ipynb`` cells 5-23, ``scripts/fr.py:30-104``) expressed with the product's
``Param`` / ``ParamSet`` containers, mirroring ``tests/golden/make_golden.py``."""

from argparse import Namespace

import numpy as np

from golemflavor_b200.enums import ParamTag, Texture
from golemflavor_b200.model import SCALE_BOUNDARIES, TEXTURE_ANGLES
from golemflavor_b200.param import Param, ParamSet
from golemflavor_b200.scan import sm_paramset

BINNING = np.logspace(np.log10(6e4), np.log10(1e7), 20 + 1)
# u_to_fr((1,0,0), NUFIT_U) -> fr_to_angles, stored by make_golden.py (asimov_angles)


def notebook_model(asimov_angles, smearing=0.02):
    """6-D SM fit of examples/inference.ipynb: 4 mixing params (3 LG + flat dcp) + 2 source angles."""
    tag = ParamTag.BESTFIT
    asimov = ParamSet([
        Param(name='measured_angle1', value=float(asimov_angles[0]), ranges=[0., 1.], std=smearing, tag=tag),
        Param(name='measured_angle2', value=float(asimov_angles[1]), ranges=[-1., 1.], std=smearing, tag=tag)])
    nuis = sm_paramset(with_mass=False)
    nuis[3] = Param(name='dcp', value=4.08404, seed=[0, 2 * np.pi], ranges=[0., 2 * np.pi], std=2.0, tag=ParamTag.SM_ANGLES)
    tag = ParamTag.SRCANGLES
    src = [Param(name='source_angle1', value=0, ranges=[0., 1.], tag=tag),
           Param(name='source_angle2', value=0, ranges=[-1., 1.], tag=tag)]
    args = Namespace(source_ratio=[1, 2, 0], no_bsm=True)
    return args, asimov, ParamSet(nuis + src)


def bsm7_paramset(dim=6):
    """6 SM params + logLam (scripts/fr.py:62-104 without the GolemFit nuisance block)."""
    b = SCALE_BOUNDARIES[dim]
    return ParamSet(sm_paramset(with_mass=True) + [
        Param(name='logLam', value=float(np.mean(b)), ranges=list(b), std=3, tag=ParamTag.SCALE)])


def bsm11_paramset(dim=6, npang=TEXTURE_ANGLES['OET']):
    """6 SM + 4 MMANGLES + logLam (the Texture.NONE path; layout of ref_flux.npz theta)."""
    ps = sm_paramset(with_mass=True)
    for k, nm in enumerate(['np_s12', 'np_c13', 'np_s23', 'np_dcp']):
        ps.append(Param(name=nm, value=float(npang[k]), ranges=[0., 2 * np.pi], std=0.2, tag=ParamTag.MMANGLES))
    b = SCALE_BOUNDARIES[dim]
    ps.append(Param(name='logLam', value=float(np.mean(b)), ranges=list(b), std=3, tag=ParamTag.SCALE))
    return ParamSet(ps)


def bsm_args(dim=6, texture=Texture.OET, source=(1, 2, 0), binning=BINNING):
    s = np.asarray(source, dtype=np.float64)
    return Namespace(binning=np.asarray(binning), source_ratio=s / s.sum(), dimension=dim, texture=texture, no_bsm=False)


def bsm_model_c3(asimov_angles, dim=6, texture=Texture.OET, source=(1, 2, 0), smearing=0.02):
    """BASELINE config 3: BSM dim-6 operator fit (6 SM params + logLam, fixed texture, Gaussian LLH)."""
    tag = ParamTag.BESTFIT
    asimov = ParamSet([
        Param(name='astroFlavorAngle1', value=float(asimov_angles[0]), ranges=[0., 1.], std=smearing, tag=tag),
        Param(name='astroFlavorAngle2', value=float(asimov_angles[1]), ranges=[-1., 1.], std=smearing, tag=tag)])
    return bsm_args(dim, texture, source), asimov, bsm7_paramset(dim)


def draw_in_ranges(pset, n, rng, seeds=False):
    box = np.array(pset.seeds if seeds else pset.ranges, dtype=np.float64)
    return rng.uniform(box[:, 0], box[:, 1], size=(n, len(pset)))


def sm_fit_c1(asimov_angles, smearing=0.02):
    """BASELINE config 1: 3 source-flavor params (raw ratios in [0, 1]), fixed NuFIT PMNS, Gaussian LLH."""
    tag = ParamTag.BESTFIT
    asimov = ParamSet([
        Param(name='measured_angle1', value=float(asimov_angles[0]), ranges=[0., 1.], std=smearing, tag=tag),
        Param(name='measured_angle2', value=float(asimov_angles[1]), ranges=[-1., 1.], std=smearing, tag=tag)])
    tag = ParamTag.SRCANGLES
    eps = 1e-6
    src = [Param(name='f_%s' % n, value=1. / 3, ranges=[eps, 1.], tag=tag) for n in ('e', 'mu', 'tau')]
    return Namespace(source_ratio=[1, 2, 0], no_bsm=True), asimov, ParamSet(src)


SOURCE_KINDS = {'angles': 2, 'x': 1, 'ratios': 3}


def source_params(kind):
    """SRCANGLES-tagged source parametrisations the reference composes (llh.py:104-110: two angles;
    scripts/mc_x.py:187: x -> (x, 1-x, 0); BASELINE config 1: three raw ratios normalised by u_to_fr)."""
    tag = ParamTag.SRCANGLES
    if kind == 'angles':
        return [Param(name='astroFlavorAngle1', value=0.5, ranges=[0., 1.], tag=tag),
                Param(name='astroFlavorAngle2', value=0., ranges=[-1., 1.], tag=tag)]
    if kind == 'x':
        return [Param(name='astroX', value=0.5, seed=[0., 1.], ranges=[0., 1.], std=0.1, tag=tag)]
    if kind == 'ratios':
        return [Param(name='f_%s' % n, value=1. / 3, ranges=[1e-6, 1.], tag=tag) for n in ('e', 'mu', 'tau')]
    raise ValueError(kind)


def bsm_sampled_source(asimov_angles, kind='angles', dim=6, texture=Texture.OET, smearing=0.02, source_first=False):
    """The composition of llh.py:94-112 with a SAMPLED source on the binned BSM path: 6 SM params +
    the source parameters + logLam (source columns optionally ahead of the SM block: the column map is
    a runtime property of the model)."""
    tag = ParamTag.BESTFIT
    asimov = ParamSet([
        Param(name='measured_angle1', value=float(asimov_angles[0]), ranges=[0., 1.], std=smearing, tag=tag),
        Param(name='measured_angle2', value=float(asimov_angles[1]), ranges=[-1., 1.], std=smearing, tag=tag)])
    b = SCALE_BOUNDARIES[dim]
    scale = [Param(name='logLam', value=float(np.mean(b)), ranges=list(b), std=3, tag=ParamTag.SCALE)]
    sm, src = sm_paramset(with_mass=True), source_params(kind)
    return bsm_args(dim, texture), asimov, ParamSet((src + sm if source_first else sm + src) + scale)

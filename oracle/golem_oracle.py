"""NumPy (x87 ``longdouble``) restatement of GolemFlavor's log-posterior path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  The product package
never imports this module; a missing CUDA extension makes the product fail.

Two layers live here, both citing the reference lines they restate
(paths relative to the upstream repository root):

* ``*_scalar``-style functions with the reference's own names
  (``angles_to_u``, ``cardano_eqn``, ``params_to_BSMu``, ``u_to_fr``,
  ``flux_averaged_BSMu``, ``lnprior``, ``multi_gaussian``, ``ln_prob`` ...):
  one parameter point per call, same arithmetic type (80-bit ``np.longdouble`` /
  ``np.clongdouble``), same algorithmic steps and the same SciPy calls as the
  reference.  They are what ``bench.py`` times as the CPU "port" baseline, and
  what the golden fixtures pin.
* ``batch_*`` functions: the same formulas over a leading batch axis, used by
  the GPU parity tests at sizes (1e4..1e5 points) where a per-point Python loop
  would take minutes.  They are pinned against the scalar layer in
  ``tests/test_oracle_golden.py``.

Parity status: PINNED (docstring known answers of ``golemflavor/fr.py`` and
fixtures generated from the unmodified reference, ``tests/golden/``).
The emcee sampler is an un-vendored dependency of the reference with no golden
chains: sampler-level parity is UNPINNED and is defined on ``ln_prob(theta)``.
"""

from __future__ import annotations

import math
from copy import deepcopy

import numpy as np

LD = np.longdouble
CLD = np.clongdouble
PI_LD = np.arccos(LD(-1))

# golemflavor/fr.py:42
MASS_EIGENVALUES = [7.40e-23, 2.515e-21]
# golemflavor/fr.py:45-52
SCALE_BOUNDARIES = {3: (-32, -20), 4: (-40, -24), 5: (-48, -27),
                    6: (-56, -30), 7: (-64, -33), 8: (-72, -36)}
# golemflavor/fr.py:313
NUFIT_ANGLES = (0.307, (1 - 0.02195) ** 2, 0.565, 3.97935)

# golemflavor/fr.py:370-376 -- fixed-texture new-physics mixing angles
_Z = 0. + 1e-9
TEXTURE_ANGLES = {
    'OEU': (0.5, 1.0, _Z, _Z),
    'OET': (_Z, 0.25, _Z, _Z),
    'OUT': (_Z, 1.0, 0.5, _Z),
}


def _tagname(obj):
    """Name of an Enum-like tag/prior/texture (works for the reference's enums,
    the product's enums and plain strings)."""
    if obj is None:
        return 'NONE'
    return getattr(obj, 'name', str(obj)).upper()


# --------------------------------------------------------------------------
# scalar layer (one point per call)
# --------------------------------------------------------------------------

def determinant(x):
    """3x3 determinant by cofactor expansion along the first column.
    Restates golemflavor/fr.py:56-79."""
    m = x
    return (m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
            - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
            + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]))


def angles_to_fr(src_angles):
    """(sin^4 phi, cos 2psi) -> (f_e, f_mu, f_tau).  golemflavor/fr.py:82-113."""
    sphi4, c2psi = LD(src_angles[0]), LD(src_angles[1])
    psi = LD(0.5) * np.arccos(c2psi)
    sphi2 = np.sqrt(sphi4)
    cphi2 = LD(1) - sphi2
    spsi2 = np.sin(psi) ** 2
    cpsi2 = LD(1) - spsi2
    return (float(abs(sphi2 * cpsi2)), float(abs(sphi2 * spsi2)),
            float(abs(cphi2)))


def normalize_fr(fr):
    """x / sum(x).  golemflavor/fr.py:240-259."""
    return np.array(fr) / float(np.sum(fr))


def fr_to_angles(ratios):
    """Inverse of angles_to_fr.  golemflavor/fr.py:289-310."""
    f0, f1, f2 = normalize_fr(ratios)
    cphi2 = f2
    sphi2 = 1.0 - cphi2
    if sphi2 == 0.:
        return (0., 0.)
    cpsi2 = f0 / sphi2
    sphi4 = sphi2 ** 2
    c2psi = np.cos(np.arccos(np.sqrt(cpsi2)) * 2)
    return (sphi4, c2psi)


def angles_to_u(bsm_angles):
    """(s12^2, c13^4, s23^2, dcp) -> PMNS-like unitary R23 . R13(dcp) . R12.
    golemflavor/fr.py:116-162 (asin/acos round trip, then the product of the
    three rotation matrices)."""
    s12_2, c13_4, s23_2, dcp = [LD(v) for v in bsm_angles]
    dcp = CLD(dcp)
    c13_2 = np.sqrt(c13_4)
    th12 = np.arcsin(np.sqrt(s12_2))
    th13 = np.arccos(np.sqrt(c13_2))
    th23 = np.arcsin(np.sqrt(s23_2))
    c12, s12 = np.cos(th12), np.sin(th12)
    c13, s13 = np.cos(th13), np.sin(th13)
    c23, s23 = np.cos(th23), np.sin(th23)
    r23 = np.zeros((3, 3), dtype=CLD)
    r13 = np.zeros((3, 3), dtype=CLD)
    r12 = np.zeros((3, 3), dtype=CLD)
    r23[0, 0] = 1
    r23[1, 1] = c23
    r23[1, 2] = s23
    r23[2, 1] = -s23
    r23[2, 2] = c23
    r13[0, 0] = c13
    r13[0, 2] = s13 * np.exp(-1j * dcp)
    r13[1, 1] = 1
    r13[2, 0] = -s13 * np.exp(1j * dcp)
    r13[2, 2] = c13
    r12[0, 0] = c12
    r12[0, 1] = s12
    r12[1, 0] = -s12
    r12[1, 1] = c12
    r12[2, 2] = 1
    return np.dot(np.dot(r23, r13), r12)


NUFIT_U = angles_to_u(NUFIT_ANGLES)


def cardano_eqn(ham):
    """Analytic eigenvector matrix of a 3x3 Hermitian matrix (PRD 91, 052003):
    characteristic-polynomial coefficients -> trigonometric Cardano roots ->
    eigenvector columns (conj(B)C, AC, AB)/N.  golemflavor/fr.py:170-237."""
    if np.shape(ham) != (3, 3):
        raise ValueError('Input matrix should be a square and dimension 3, '
                         'got\n{0}'.format(ham))
    h = ham
    tr = np.trace(h)
    a = -tr
    b = LD(1) / 2 * (tr ** LD(2) - np.trace(np.dot(h, h)))
    c = -determinant(h)

    q = (LD(1) / 9) * (a ** LD(2) - LD(3) * b)
    r = (LD(1) / 54) * (LD(2) * a ** LD(3) - LD(9) * a * b + LD(27) * c)
    theta = np.arccos(r / np.sqrt(q ** LD(3)))

    energies = []
    for shift in (LD(0), -LD(2) * PI_LD, LD(2) * PI_LD):
        energies.append(-LD(2) * np.sqrt(q) * np.cos((theta + shift) / LD(3))
                        - (LD(1) / 3) * a)

    cols = []
    for e in energies:
        big_a = h[1][2] * (h[0][0] - e) - h[1][0] * h[0][2]
        big_b = h[2][0] * (h[1][1] - e) - h[2][1] * h[1][0]
        big_c = h[1][0] * (h[2][2] - e) - h[1][2] * h[2][0]
        norm = np.sqrt(np.abs(big_a * big_b) ** 2 + np.abs(big_a * big_c) ** 2
                       + np.abs(big_b * big_c) ** 2)
        cols.append((np.conjugate(big_b) * big_c / norm,
                     big_a * big_c / norm,
                     big_a * big_b / norm))
    return np.array([[cols[k][row] for k in range(3)] for row in range(3)])


def test_unitarity(x, prnt=False, rse=False, epsilon=None):
    """|x x^dagger|, optionally asserting trace and total are 3 within epsilon.
    golemflavor/fr.py:461-499."""
    f = np.abs(np.dot(x, x.conj().T), dtype=LD)
    if prnt:
        print('Unitarity test:\n{0}'.format(f))
    if rse:
        if not np.abs(np.trace(f) - 3.) < epsilon or \
           not np.abs(np.sum(f) - 3.) < epsilon:
            raise AssertionError('Matrix is not unitary!\nx\n{0}\ntest '
                                 'u\n{1}'.format(x, f))
    return f


test_unitarity.__test__ = False  # not a pytest test


def texture_tuple(bsm_angles, texture):
    """Resolve (np_s12_2, np_c13_4, np_s23_2, np_dcp, logLam) for a texture.
    golemflavor/fr.py:367-378.  (The reference builds a ragged array for the
    fixed textures under NumPy >= 1.24; unpacking the scalar here is the
    equivalent call on the identical code path.)"""
    name = _tagname(texture)
    if not isinstance(bsm_angles, (list, tuple, np.ndarray)):
        bsm_angles = [bsm_angles]
    if name in TEXTURE_ANGLES:
        sc = bsm_angles[-1] if len(bsm_angles) else bsm_angles
        return TEXTURE_ANGLES[name] + (sc,)
    return tuple(bsm_angles)


def params_to_BSMu(bsm_angles, dim, energy, mass_eigenvalues=MASS_EIGENVALUES,
                   sm_u=NUFIT_U, no_bsm=False, texture='NONE',
                   check_uni=True, epsilon=1e-7):
    """Eigenvector matrix of H = U diag(0,m21,m3x) U^+ /(2E)
    + E^(dim-3) NP_U diag(0, Lam/100, Lam) NP_U^+.  golemflavor/fr.py:317-400."""
    if np.shape(sm_u) != (3, 3):
        raise ValueError('Input matrix should be a square and dimension 3, '
                         'got\n{0}'.format(sm_u))
    np_s12_2, np_c13_4, np_s23_2, np_dcp, sc2 = texture_tuple(bsm_angles, texture)
    sc2 = np.power(10., sc2)
    sc1 = sc2 / 100.
    mass = np.diag([0, mass_eigenvalues[0], mass_eigenvalues[1]])
    sm_ham = (1. / (2 * energy)) * np.dot(sm_u, np.dot(mass, sm_u.conj().T))
    if no_bsm:
        vec = cardano_eqn(sm_ham)
    else:
        np_u = angles_to_u((np_s12_2, np_c13_4, np_s23_2, np_dcp))
        scales = np.diag([0, sc1, sc2])
        bsm_term = (energy ** (dim - 3)) * np.dot(np_u, np.dot(scales, np_u.conj().T))
        vec = cardano_eqn(sm_ham + bsm_term)
    if check_uni:
        test_unitarity(vec, rse=True, epsilon=epsilon)
    return vec


def u_to_fr(source_fr, matrix):
    """Decoherent flavor transition: fr_b = sum_{a,i} |U_ai|^2 |U_bi|^2 s_a / sum(s).
    golemflavor/fr.py:502-536."""
    try:
        p = np.abs(matrix) ** 2
        comp = np.einsum('ai, bi, a -> b', p, p, source_fr)
    except Exception:
        matrix = np.array(matrix, dtype=CLD)
        p = np.abs(matrix) ** 2
        comp = np.einsum('ai, bi, a -> b', p, p, source_fr)
    return comp / np.sum(source_fr)


def flux_averaged_BSMu(theta, args, spectral_index, llh_paramset):
    """Energy-bin-averaged measured flavor ratio.  golemflavor/fr.py:403-458.

    ``llh_paramset`` is duck-typed: an iterable of objects with ``name``,
    ``value`` and ``tag`` (Enum or str); ``args`` needs ``binning`` (bin edges),
    ``source_ratio``, ``dimension``, ``texture`` and ``no_bsm``."""
    if len(theta) != len(llh_paramset):
        raise AssertionError('Length of MCMC scan is not the same as the input '
                             'params\ntheta={0}\nparamset]{1}'.format(theta, llh_paramset))
    params = list(llh_paramset)
    for idx, prm in enumerate(params):
        prm.value = theta[idx]

    binning = np.asarray(args.binning)
    centers = np.sqrt(binning[:-1] * binning[1:])
    widths = np.abs(np.diff(binning))
    source_flux = np.array([f * np.power(centers, spectral_index)
                            for f in args.source_ratio]).T

    bsm_angles = tuple(p.value for p in params
                       if _tagname(p.tag) in ('SCALE', 'MMANGLES'))
    names = [p.name for p in params]
    m_names = ['m21_2', 'm3x_2']
    a_names = ['s_12_2', 'c_13_4', 's_23_2', 'dcp']
    if set(m_names + a_names).issubset(set(names)):
        mass_eigenvalues = [p.value for p in params if p.name in m_names]
        sm_u = angles_to_u([p.value for p in params if p.name in a_names])
    else:
        mass_eigenvalues = MASS_EIGENVALUES
        sm_u = NUFIT_U

    per_bin = []
    for ib in range(len(centers)):
        if getattr(args, 'no_bsm', False):
            # golemflavor/fr.py:437-438: `fr = u_to_fr(source_flux, sm_u)`.  The reference hands the
            # WHOLE [nbins, 3] flux table to u_to_fr, whose einsum 'ai,bi,a->b' rejects a 2-D source:
            # the unmodified reference raises ValueError on this branch for every binning (checked in
            # the build container, tests/golden/make_golden.py docstring).  The evident intent -- no
            # new physics, vacuum mixing with sm_u in every energy bin -- is restated per bin and sent
            # through the same width-weighted average; E^gamma cancels in u_to_fr's normalisation, so
            # the result is u_to_fr(args.source_ratio, sm_u) whatever the binning.
            u = np.array(sm_u, dtype=CLD)
        else:
            u = params_to_BSMu(bsm_angles=bsm_angles, dim=args.dimension,
                               energy=centers[ib], mass_eigenvalues=mass_eigenvalues,
                               sm_u=sm_u, no_bsm=args.no_bsm, texture=args.texture)
        per_bin.append(u_to_fr(source_flux[ib], u))
    measured = np.array(per_bin).T
    integrated = np.sum(measured * widths, axis=1)
    averaged = (1. / (binning[-1] - binning[0])) * integrated
    return averaged / np.sum(averaged)


def GaussianBoundedRV(loc=0., sigma=1., lower=-np.inf, upper=np.inf):
    """golemflavor/llh.py:25-29 (SciPy frozen truncnorm, as the reference)."""
    import scipy.stats
    low, up = (lower - loc) / sigma, (upper - loc) / sigma
    return scipy.stats.truncnorm(loc=loc, scale=sigma, a=low, b=up)


def multi_gaussian(fr, fr_bf, smearing, offset=-320):
    """log N_3(fr; fr_bf, smearing^2 I) + offset, via SciPy pdf then log
    (so it is -inf once the pdf underflows).  golemflavor/llh.py:32-54."""
    from scipy.stats import multivariate_normal
    cov = np.identity(3) * pow(smearing, 2)
    with np.errstate(divide='ignore'):
        return np.log(multivariate_normal.pdf(fr, mean=fr_bf, cov=cov)) + offset


def lnprior(theta, paramset):
    """Box prior in ``ranges`` plus (truncated-)Gaussian terms.
    golemflavor/llh.py:65-91."""
    if len(theta) != len(paramset):
        raise AssertionError('Length of MCMC scan is not the same as the input '
                             'params\ntheta={0}\nparamset={1}'.format(theta, paramset))
    params = list(paramset)
    for idx, prm in enumerate(params):
        prm.value = theta[idx]
    for value, prm in zip(theta, params):
        lo, hi = prm.ranges
        if not (lo <= value <= hi):
            return -np.inf
    total = 0
    for prm in params:
        kind = _tagname(prm.prior)
        if kind == 'GAUSSIAN':
            total += GaussianBoundedRV(loc=prm.nominal_value,
                                       sigma=prm.std).logpdf(prm.value)
        elif kind == 'LIMITEDGAUSS':
            total += GaussianBoundedRV(loc=prm.nominal_value, sigma=prm.std,
                                       lower=prm.ranges[0],
                                       upper=prm.ranges[1]).logpdf(prm.value)
    return total


def source_from_params(src, args):
    """Source composition from the SRCANGLES-tagged values of a parameter set:
    two values are the angles (sin^4 phi, cos 2psi) of ``angles_to_fr``
    (golemflavor/llh.py:104-110, examples/inference.ipynb cell 21); one value is
    x with source (x, 1-x, 0) (scripts/mc_x.py:187); three values are raw
    flavor ratios, normalised later by ``u_to_fr`` (golemflavor/fr.py:535 --
    BASELINE config 1, "3 source-flavor params, fixed PMNS"); none:
    ``args.source_ratio``."""
    if len(src) == 2:
        return angles_to_fr(src)
    if len(src) == 1:
        return (src[0], 1.0 - src[0], 0.0)
    if len(src) == 3:
        return tuple(src)
    if src:
        raise ValueError('expected one, two or three SRCANGLES params, got {0}'.format(len(src)))
    return args.source_ratio


def triangle_llh_gauss(theta, args, asimov_paramset, llh_paramset):
    """Gaussian flavor-ratio likelihood composed as in the reference notebooks
    (examples/inference.ipynb cell 21, examples/tutorial.ipynb), generalised to
    the BSM path of golemflavor/llh.py:94-112 with ``multi_gaussian`` in place
    of the proprietary GolemFit call (README.md:76-77).

    * source composition: SRCANGLES-tagged params if present, else
      ``args.source_ratio``;
    * measured composition: ``flux_averaged_BSMu`` if a SCALE-tagged param is
      present, else ``u_to_fr(source, angles_to_u(SM_ANGLES))``;
    * injected composition: ``angles_to_fr`` of the BESTFIT-tagged asimov params;
      smearing: ``std`` of the first BESTFIT param.
    """
    if len(theta) != len(llh_paramset):
        raise AssertionError('Length of MCMC scan is not the same as the input '
                             'params\ntheta={0}\nparamset]{1}'.format(theta, llh_paramset))
    params = list(llh_paramset)
    for idx, prm in enumerate(params):
        prm.value = theta[idx]
    bestfit = [p for p in asimov_paramset if _tagname(p.tag) == 'BESTFIT']
    fr_bf = angles_to_fr([p.value for p in bestfit])
    smearing = bestfit[0].std

    src = [p.value for p in params if _tagname(p.tag) == 'SRCANGLES']
    has_scale = any(_tagname(p.tag) == 'SCALE' for p in params)
    source = source_from_params(src, args)
    if has_scale:
        gamma = getattr(args, 'spectral_index', -2.0)
        if src:
            args = deepcopy(args)
            args.source_ratio = np.array(source)
        fr = flux_averaged_BSMu(theta, args, gamma, llh_paramset)
    else:
        names = ['s_12_2', 'c_13_4', 's_23_2', 'dcp']
        sm = [p.value for p in params if _tagname(p.tag) == 'SM_ANGLES'
              and p.name in names]
        sm_u = angles_to_u(sm) if len(sm) == 4 else NUFIT_U
        fr = u_to_fr(source, sm_u)
    return multi_gaussian(fr, fr_bf, smearing)


def ln_prob(theta, args, asimov_paramset, llh_paramset):
    """golemflavor/llh.py:121-130 with the Gaussian likelihood."""
    dc_asimov = deepcopy(asimov_paramset)
    dc_llh = deepcopy(llh_paramset)
    lp = lnprior(theta, paramset=dc_llh)
    if not np.isfinite(lp):
        return -np.inf
    return lp + triangle_llh_gauss(theta, args, dc_asimov, dc_llh)


# --------------------------------------------------------------------------
# batch layer (leading batch axis, same formulas, longdouble)
# --------------------------------------------------------------------------

def batch_angles_to_fr(src):
    """Vectorised angles_to_fr; src[..., 2] -> fr[..., 3] (float64 like the
    reference's float() casts).  golemflavor/fr.py:101-113."""
    src = np.asarray(src, dtype=LD)
    sphi4, c2psi = src[..., 0], src[..., 1]
    psi = LD(0.5) * np.arccos(c2psi)
    sphi2 = np.sqrt(sphi4)
    cphi2 = LD(1) - sphi2
    spsi2 = np.sin(psi) ** 2
    cpsi2 = LD(1) - spsi2
    out = np.stack([np.abs(sphi2 * cpsi2), np.abs(sphi2 * spsi2), np.abs(cphi2)],
                   axis=-1)
    return out.astype(np.float64)


def batch_angles_to_u(ang):
    """Vectorised angles_to_u; ang[..., 4] -> U[..., 3, 3] (clongdouble).
    Product R23.R13.R12 written out entry by entry.  golemflavor/fr.py:138-162."""
    ang = np.asarray(ang, dtype=LD)
    s12_2, c13_4, s23_2, dcp = (ang[..., k] for k in range(4))
    c13_2 = np.sqrt(c13_4)
    th12 = np.arcsin(np.sqrt(s12_2))
    th13 = np.arccos(np.sqrt(c13_2))
    th23 = np.arcsin(np.sqrt(s23_2))
    c12, s12 = np.cos(th12), np.sin(th12)
    c13, s13 = np.cos(th13), np.sin(th13)
    c23, s23 = np.cos(th23), np.sin(th23)
    ep = np.exp(1j * dcp.astype(CLD))
    em = np.exp(-1j * dcp.astype(CLD))
    u = np.zeros(ang.shape[:-1] + (3, 3), dtype=CLD)
    u[..., 0, 0] = c13 * c12
    u[..., 0, 1] = c13 * s12
    u[..., 0, 2] = s13 * em
    u[..., 1, 0] = -c23 * s12 - s23 * s13 * ep * c12
    u[..., 1, 1] = c23 * c12 - s23 * s13 * ep * s12
    u[..., 1, 2] = s23 * c13
    u[..., 2, 0] = s23 * s12 - c23 * s13 * ep * c12
    u[..., 2, 1] = -s23 * c12 - c23 * s13 * ep * s12
    u[..., 2, 2] = c23 * c13
    return u


def batch_cardano(h):
    """Vectorised cardano_eqn; h[..., 3, 3] clongdouble -> eigenvector matrices.
    golemflavor/fr.py:204-237.  Returns NaN where the reference would (A, B or C
    vanishing), no exception."""
    h = np.asarray(h, dtype=CLD)
    tr = h[..., 0, 0] + h[..., 1, 1] + h[..., 2, 2]
    h2 = np.matmul(h, h)
    tr2 = h2[..., 0, 0] + h2[..., 1, 1] + h2[..., 2, 2]
    det = (h[..., 0, 0] * (h[..., 1, 1] * h[..., 2, 2] - h[..., 2, 1] * h[..., 1, 2])
           - h[..., 1, 0] * (h[..., 0, 1] * h[..., 2, 2] - h[..., 2, 1] * h[..., 0, 2])
           + h[..., 2, 0] * (h[..., 0, 1] * h[..., 1, 2] - h[..., 1, 1] * h[..., 0, 2]))
    a = -tr
    b = LD(1) / 2 * (tr ** 2 - tr2)
    c = -det
    q = (LD(1) / 9) * (a ** 2 - LD(3) * b)
    r = (LD(1) / 54) * (LD(2) * a ** 3 - LD(9) * a * b + LD(27) * c)
    with np.errstate(all='ignore'):
        theta = np.arccos(r / np.sqrt(q ** 3))
        out = np.zeros(h.shape, dtype=CLD)
        for k, shift in enumerate((LD(0), -LD(2) * PI_LD, LD(2) * PI_LD)):
            e = -LD(2) * np.sqrt(q) * np.cos((theta + shift) / LD(3)) - (LD(1) / 3) * a
            big_a = h[..., 1, 2] * (h[..., 0, 0] - e) - h[..., 1, 0] * h[..., 0, 2]
            big_b = h[..., 2, 0] * (h[..., 1, 1] - e) - h[..., 2, 1] * h[..., 1, 0]
            big_c = h[..., 1, 0] * (h[..., 2, 2] - e) - h[..., 1, 2] * h[..., 2, 0]
            norm = np.sqrt(np.abs(big_a * big_b) ** 2 + np.abs(big_a * big_c) ** 2
                           + np.abs(big_b * big_c) ** 2)
            out[..., 0, k] = np.conjugate(big_b) * big_c / norm
            out[..., 1, k] = big_a * big_c / norm
            out[..., 2, k] = big_a * big_b / norm
    return out


def batch_unitarity_residual(v):
    """max(|tr f - 3|, |sum f - 3|) with f = |V V^+| (the quantity the reference
    asserts on, golemflavor/fr.py:489-494)."""
    f = np.abs(np.matmul(v, np.conjugate(np.swapaxes(v, -1, -2))))
    t = f[..., 0, 0] + f[..., 1, 1] + f[..., 2, 2]
    s = f.sum(axis=(-1, -2))
    return np.maximum(np.abs(t - 3), np.abs(s - 3)).astype(np.float64)


def batch_u_to_fr(source, u):
    """Vectorised u_to_fr; source[..., 3] (or [3]), u[..., 3, 3] -> fr[..., 3].
    golemflavor/fr.py:525-535."""
    p = np.abs(np.asarray(u)) ** 2
    source = np.broadcast_to(np.asarray(source, dtype=LD), p.shape[:-2] + (3,))
    w = np.einsum('...ai,...a->...i', p, source)
    comp = np.einsum('...bi,...i->...b', p, w)
    return comp / source.sum(axis=-1)[..., None]


def batch_bsm_hamiltonian(sm_u, mass, np_u, loglam, dim, energy):
    """H = sm_u diag(0,m21,m3x) sm_u^+ /(2E) + E^(dim-3) np_u diag(0,L/100,L) np_u^+
    (golemflavor/fr.py:380-395).  Shapes broadcast over leading axes; ``energy``
    may carry a trailing bin axis: the result is [..., nbins, 3, 3]."""
    sm_u = np.asarray(sm_u, dtype=CLD)
    np_u = np.asarray(np_u, dtype=CLD)
    mass = np.asarray(mass, dtype=np.float64)
    sc2 = np.power(10., np.asarray(loglam, dtype=np.float64))
    sc1 = sc2 / 100.
    energy = np.asarray(energy, dtype=np.float64)
    # U diag(0, a, b) U^+ = a u1 u1^+ + b u2 u2^+
    def outer(u, k):
        return u[..., :, k, None] * np.conjugate(u[..., None, :, k])
    h0 = mass[..., 0, None, None] * outer(sm_u, 1) + mass[..., 1, None, None] * outer(sm_u, 2)
    t = sc1[..., None, None] * outer(np_u, 1) + sc2[..., None, None] * outer(np_u, 2)
    e = energy[..., None, None]
    return (1. / (2 * e)) * h0[..., None, :, :] + (e ** (dim - 3)) * t[..., None, :, :]


def batch_flux_averaged_fr(sm_angles, mass, np_angles, loglam, dim, binning,
                           source, eig='cardano'):
    """Vectorised flux_averaged_BSMu (golemflavor/fr.py:413-457).

    sm_angles[N,4], mass[N,2], np_angles[N,4] or [4], loglam[N], source[3] or
    [N,3].  Returns (fr[N,3] float64, resid[N] = worst per-bin unitarity
    residual of the eigenvector matrix, which the reference asserts < 1e-7)."""
    sm_angles = np.asarray(sm_angles, dtype=np.float64)
    n = sm_angles.shape[0]
    binning = np.asarray(binning, dtype=np.float64)
    centers = np.sqrt(binning[:-1] * binning[1:])
    widths = np.abs(np.diff(binning))
    sm_u = batch_angles_to_u(sm_angles)
    np_angles = np.broadcast_to(np.asarray(np_angles, dtype=np.float64), (n, 4))
    np_u = batch_angles_to_u(np_angles)
    mass = np.broadcast_to(np.asarray(mass, dtype=np.float64), (n, 2))
    h = batch_bsm_hamiltonian(sm_u, mass, np_u, loglam, dim,
                              np.broadcast_to(centers, (n, len(centers))))
    if eig == 'cardano':
        v = batch_cardano(h)
    else:
        v = eig(h)
    resid = batch_unitarity_residual(v).max(axis=-1)
    source = np.broadcast_to(np.asarray(source, dtype=LD), (n, 3))
    fr_bin = batch_u_to_fr(source[:, None, :], v)            # [N, nbins, 3]
    integrated = np.sum(fr_bin * widths[None, :, None], axis=1)
    averaged = (1. / (binning[-1] - binning[0])) * integrated
    fr = averaged / averaged.sum(axis=-1)[:, None]
    return fr.astype(np.float64), resid


def truncnorm_lognorm(mu, sigma, lo=-np.inf, hi=np.inf):
    """log of the normalising constant of scipy.stats.truncnorm(a,b,loc,scale):
    logpdf(x) = -0.5 z^2 - log(sigma sqrt(2 pi)) - log(Phi(b) - Phi(a)).
    Closed form of golemflavor/llh.py:25-29, 82-90 (checked against SciPy in
    tests/test_oracle_golden.py)."""
    from scipy.special import ndtr, log_ndtr
    a, b = (lo - mu) / sigma, (hi - mu) / sigma
    if np.isinf(a) and np.isinf(b):
        logz = 0.0
    else:
        z = ndtr(b) - ndtr(a)
        if z > 1e-8:
            logz = math.log(z)
        else:  # deep tail: use log-space difference
            hi_l, lo_l = (log_ndtr(b), log_ndtr(a)) if b <= 0 else (log_ndtr(-a), log_ndtr(-b))
            logz = hi_l + math.log1p(-math.exp(lo_l - hi_l))
    return -math.log(sigma * math.sqrt(2 * math.pi)) - logz


def batch_lnprior(theta, lo, hi, kind, mu, sigma):
    """Vectorised lnprior (golemflavor/llh.py:74-90).  kind: 0 uniform,
    1 GAUSSIAN (unbounded normal), 2 LIMITEDGAUSS (normal truncated to [lo,hi])."""
    theta = np.asarray(theta, dtype=np.float64)
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    inside = np.all((theta >= lo) & (theta <= hi), axis=-1)
    total = np.zeros(theta.shape[:-1])
    for d in range(theta.shape[-1]):
        if kind[d] == 0:
            continue
        if kind[d] == 1:
            ln = truncnorm_lognorm(mu[d], sigma[d])
        else:
            ln = truncnorm_lognorm(mu[d], sigma[d], lo[d], hi[d])
        z = (theta[..., d] - mu[d]) / sigma[d]
        total = total + (-0.5 * z * z + ln)
    return np.where(inside, total, -np.inf)


# exp(x) rounds to +0 below log(2^-1075) = log(smallest subnormal / 2)
MG_UNDERFLOW_LOGPDF = -1075 * math.log(2.0)


def batch_multi_gaussian(fr, fr_bf, smearing, offset=-320.0, emulate_underflow=True):
    """Closed form of multi_gaussian (golemflavor/llh.py:53-54):
    -|fr-bf|^2/(2 s^2) - 1.5 log(2 pi s^2) + offset; with ``emulate_underflow``
    it returns -inf where the reference's pdf underflows to zero."""
    fr = np.asarray(fr, dtype=np.float64)
    d = fr - np.asarray(fr_bf, dtype=np.float64)
    logpdf = -0.5 * np.sum(d * d, axis=-1) / smearing ** 2 \
        - 1.5 * math.log(2 * math.pi * smearing ** 2)
    out = logpdf + offset
    if emulate_underflow:
        out = np.where(logpdf < MG_UNDERFLOW_LOGPDF, -np.inf, out)
    return out


# --------------------------------------------------------------------------
# scan support: Philox4x32-10, ternary histogram
# --------------------------------------------------------------------------

_PH_M0 = np.uint64(0xD2511F53)
_PH_M1 = np.uint64(0xCD9E8D57)
_PH_W0 = np.uint32(0x9E3779B9)
_PH_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11; Random123).  counter[..., 4] uint32,
    key[2] uint32 -> out[..., 4] uint32.  Known-answer vectors in
    tests/test_oracle_golden.py."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    c0, c1, c2, c3 = (c[..., k].copy() for k in range(4))
    k0 = np.uint32(key[0])
    k1 = np.uint32(key[1])
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PH_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PH_W1)) & 0xFFFFFFFF)
    return np.stack([c0, c1, c2, c3], axis=-1)


def philox_uniforms(seed, first_index, count, block=0):
    """The scan's draw convention: sample i uses counter (lo32(i), hi32(i), block, 0)
    and key (lo32(seed), hi32(seed)); each 32-bit word x maps to the open-interval
    uniform (x + 0.5) * 2^-32.  Returns u[count, 4] float64."""
    idx = np.arange(first_index, first_index + count, dtype=np.uint64)
    ctr = np.zeros((count, 4), dtype=np.uint32)
    ctr[:, 0] = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = (idx >> np.uint64(32)).astype(np.uint32)
    ctr[:, 2] = np.uint32(block)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    x = philox4x32_10(ctr, key)
    return (x.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


def ternary_histogram(frs, nb):
    """The scan's histogram definition: np.histogramdd over (f_e, f_mu, f_tau) with
    nb+1 bins per axis on [0,1]  (golemflavor/plot.py:364-370)."""
    frs = np.asarray(frs, dtype=np.float64)
    h, _ = np.histogramdd((frs[:, 0], frs[:, 1], frs[:, 2]),
                          bins=(nb + 1, nb + 1, nb + 1),
                          range=((0, 1), (0, 1), (0, 1)))
    return h.astype(np.int64)

"""Flattening of the reference's ``ln_prob`` closure into the ``gf_model`` struct.

The reference evaluates ``partial(ln_prob, args=..., asimov_paramset=..., llh_paramset=...)``
(``scripts/fr.py:182-187``, ``examples/inference.ipynb`` cell 23) one point at a time and looks
parameters up by name / tag on every call (``fr.py:421-435``, ``llh.py:65-112``).  Here that lookup
happens once, on the host: which theta column feeds which physical quantity, the fixed values of
everything that is not sampled, the energy binning, the priors and the likelihood constants end up
in one plain struct that the kernels take as a ``__grid_constant__`` parameter.
"""

import math

import numpy as np

from . import _lib
from .enums import enum_name

# golemflavor/fr.py:42, 45-52, 313, 370-376
MASS_EIGENVALUES = [7.40e-23, 2.515e-21]
SCALE_BOUNDARIES = {3: (-32, -20), 4: (-40, -24), 5: (-48, -27), 6: (-56, -30), 7: (-64, -33), 8: (-72, -36)}
NUFIT_ANGLES = (0.307, (1 - 0.02195) ** 2, 0.565, 3.97935)
_Z = 0. + 1e-9
TEXTURE_ANGLES = {'OEU': (0.5, 1.0, _Z, _Z), 'OET': (_Z, 0.25, _Z, _Z), 'OUT': (_Z, 1.0, 0.5, _Z)}

SM_ANGLE_NAMES = ('s_12_2', 'c_13_4', 's_23_2', 'dcp')
MASS_NAMES = ('m21_2', 'm3x_2')
_PRIOR_KIND = {'UNIFORM': _lib.PRIOR_UNIFORM, 'GAUSSIAN': _lib.PRIOR_GAUSSIAN, 'LIMITEDGAUSS': _lib.PRIOR_LIMITEDGAUSS}


def host_angles_to_fr(src_angles):
    """Closed form of ``fr.angles_to_fr`` (``fr.py:101-113``) for the *constants* of a model
    (injected composition); sampled source angles are converted on the device."""
    sphi4, c2psi = float(src_angles[0]), float(src_angles[1])
    sphi2 = math.sqrt(sphi4)
    spsi2 = 0.5 * (1.0 - c2psi)
    return (abs(sphi2 * (1.0 - spsi2)), abs(sphi2 * spsi2), abs(1.0 - sphi2))


def binning_edges(binning):
    """``args.binning`` is either the edge array (after ``process_args``, ``scripts/fr.py:122-124``)
    or the raw ``[lo, hi, nbins]`` triple of ``fr_argparse`` (``fr.py:283-285``)."""
    b = np.asarray(binning, dtype=np.float64).ravel()
    if b.size == 3 and b[2] < b[1] and float(b[2]).is_integer() and b[0] < b[1]:
        b = np.logspace(np.log10(b[0]), np.log10(b[1]), int(b[2]) + 1)
    return b


class FlatModel(object):
    """``gf_model`` plus the bookkeeping the host side needs (names, ndim)."""

    def __init__(self, struct, names=()):
        self.struct = struct
        self.names = tuple(names)
        self._blob = None
        _lib.check(_lib.load().gf_model_check(self.ref))

    @property
    def ref(self):
        import ctypes
        return ctypes.byref(self.struct)

    @property
    def ndim(self):
        return int(self.struct.ndim)

    @property
    def blob(self):
        """The struct's bytes as a CPU uint8 tensor (shares memory with ``struct``): the form in which the model is
        handed to the ``torch.ops.golemflavor.*`` operators."""
        if self._blob is None:
            import ctypes

            import torch
            self._blob = torch.frombuffer((ctypes.c_uint8 * ctypes.sizeof(self.struct)).from_buffer(self.struct), dtype=torch.uint8)
        return self._blob


def _new_struct():
    m = _lib.Model()
    for arr in (m.col_sm, m.col_mass, m.col_src, m.col_np, m.col_src3):
        for k in range(len(arr)):
            arr[k] = -1
    m.col_scale = m.col_x = -1
    m.fixed_sm[:] = NUFIT_ANGLES
    m.fixed_mass[:] = MASS_EIGENVALUES
    m.fixed_src[:] = (1.0, 2.0, 0.0)
    m.fixed_np[:] = TEXTURE_ANGLES['OET']
    m.fixed_loglam = -100.0
    m.no_bsm = 1
    m.dimension = 3
    m.nbins = 0
    m.llh_kind = _lib.LLH_GAUSSIAN
    m.emulate_underflow = 1
    m.fr_bf[:] = (1. / 3, 1. / 3, 1. / 3)
    m.smearing = 1.0
    m.offset = -320.0
    m.llh_const = 1.0
    m.epsilon = 1e-7
    return m


def _set_priors(m, paramset):
    for k, p in enumerate(paramset):
        d = m.prior[k]
        d.lo, d.hi = float(p.ranges[0]), float(p.ranges[1])
        kind = _PRIOR_KIND[enum_name(p.prior, 'UNIFORM')]
        d.kind = kind
        if kind != _lib.PRIOR_UNIFORM:
            d.mu, d.sigma = float(p.nominal_value), float(p.std)


def flatten(args, asimov_paramset, llh_paramset, likelihood=None):
    """Build the flat model of ``ln_prob(theta, args, asimov_paramset, llh_paramset)``.

    Column resolution follows the reference:
      * SM mixing angles / mass splittings are read from theta when the BSM path is active only
        if all six of ``s_12_2 c_13_4 s_23_2 dcp m21_2 m3x_2`` are present (``fr.py:425-435``),
        on the SM-only path when the four ``SM_ANGLES`` are present (``inference.ipynb`` cell 21);
        otherwise NuFIT values are used (``fr.py:42, 313``);
      * ``SRCANGLES``-tagged params give the source composition -- two: the angles (sin^4 phi, cos 2psi) of
        ``fr.angles_to_fr``; one: x with source (x, 1-x, 0) (``mc_x.py:187``); three: raw ratios normalised
        by ``u_to_fr`` -- else ``args.source_ratio``;
      * a ``SCALE``-tagged param switches on the binned BSM path (``fr.py:421-423``); with
        ``Texture.NONE`` the four ``MMANGLES`` params are the new-physics mixing angles, with a
        fixed texture they come from ``fr.py:370-376``;
      * the injected composition is ``angles_to_fr`` of the ``BESTFIT`` params of
        ``asimov_paramset`` and the smearing is their ``std`` (``inference.ipynb`` cells 9, 21).
    """
    params = list(llh_paramset)
    ndim = len(params)
    if not 1 <= ndim <= _lib.GF_MAX_DIM:
        raise ValueError('llh_paramset has {0} params, supported: 1..{1}'.format(ndim, _lib.GF_MAX_DIM))
    m = _new_struct()
    m.ndim = ndim
    names = [p.name for p in params]
    tags = [enum_name(p.tag) for p in params]
    _set_priors(m, params)

    scale_cols = [k for k, t in enumerate(tags) if t == 'SCALE']
    np_cols = [k for k, t in enumerate(tags) if t == 'MMANGLES']
    src_cols = [k for k, t in enumerate(tags) if t == 'SRCANGLES']
    fixed_scale = getattr(args, 'fixed_scale', None)   # sens.py:261-266: the scale is frozen at a grid value
    no_bsm = bool(getattr(args, 'no_bsm', False)) or (not scale_cols and fixed_scale is None)

    if len(src_cols) == 1:
        m.col_x = src_cols[0]  # scripts/mc_x.py:187: a single SRCANGLES param x, source = (x, 1-x, 0)
    elif len(src_cols) == 3:
        m.col_src3[:] = src_cols  # three raw source ratios, normalised by u_to_fr (fr.py:535)
    elif src_cols:
        if len(src_cols) != 2:
            raise ValueError('expected one, two or three SRCANGLES params, got {0}'.format(len(src_cols)))
        m.col_src[:] = src_cols
    else:
        m.fixed_src[:] = [float(x) for x in args.source_ratio]

    have_angles = all(n in names for n in SM_ANGLE_NAMES)
    have_masses = all(n in names for n in MASS_NAMES)
    if no_bsm:
        m.no_bsm = 1
        if have_angles:
            m.col_sm[:] = [k for k, n in enumerate(names) if n in SM_ANGLE_NAMES]
    else:
        m.no_bsm = 0
        if have_angles and have_masses:
            # paramset order, like the list comprehensions of fr.py:428-433
            m.col_sm[:] = [k for k, n in enumerate(names) if n in SM_ANGLE_NAMES]
            m.col_mass[:] = [k for k, n in enumerate(names) if n in MASS_NAMES]
        if fixed_scale is not None and not scale_cols:
            m.fixed_loglam = float(fixed_scale)
        elif len(scale_cols) != 1:
            raise ValueError('expected one SCALE param, got {0}'.format(len(scale_cols)))
        else:
            m.col_scale = scale_cols[0]
        texture = enum_name(getattr(args, 'texture', None))
        if texture in TEXTURE_ANGLES:
            m.fixed_np[:] = TEXTURE_ANGLES[texture]
        else:
            if len(np_cols) != 4:
                raise ValueError('Texture.NONE needs four MMANGLES params, got {0}'.format(len(np_cols)))
            m.col_np[:] = np_cols
        m.dimension = int(args.dimension)
        edges = binning_edges(args.binning)
        if not 2 <= edges.size <= _lib.GF_MAX_BINS + 1:
            raise ValueError('binning has {0} edges, supported: 2..{1}'.format(edges.size, _lib.GF_MAX_BINS + 1))
        m.nbins = edges.size - 1
        for k, e in enumerate(edges):
            m.bin_edges[k] = float(e)

    kind = enum_name(likelihood if likelihood is not None else getattr(args, 'likelihood', None), 'GAUSSIAN')
    if kind == 'FLAT':
        m.llh_kind = _lib.LLH_FLAT
        m.llh_const = float(getattr(args, 'llh_const', 1.0))
    elif kind == 'GAUSSIAN':
        m.llh_kind = _lib.LLH_GAUSSIAN
        bestfit = [p for p in (asimov_paramset or []) if enum_name(p.tag) == 'BESTFIT']
        if len(bestfit) == 2:
            m.fr_bf[:] = host_angles_to_fr([p.value for p in bestfit])
            m.smearing = float(bestfit[0].std)
        elif getattr(args, 'injected_ratio', None) is not None:
            inj = np.asarray(args.injected_ratio, dtype=np.float64)
            m.fr_bf[:] = inj / inj.sum()
            m.smearing = float(args.smearing)
        else:
            raise ValueError('Gaussian likelihood needs two BESTFIT params in asimov_paramset '
                             '(or args.injected_ratio and args.smearing)')
        m.offset = float(getattr(args, 'llh_offset', -320.0))
        m.emulate_underflow = int(bool(getattr(args, 'emulate_underflow', True)))
    else:
        raise ValueError('Likelihood.{0} needs the proprietary GolemFit fitter (gf.py:118-123), which is not part '
                         'of the hot path; use Likelihood.GAUSSIAN (README.md:76-77) or FLAT'.format(kind))
    return FlatModel(m, names)


def physics_model(source_ratio=(1, 2, 0), sm_angles=None, mass=None, no_bsm=True, dimension=3, texture='NONE',
                  binning=None, np_angles=None, loglam=None, ndim=1):
    """A model with no sampled columns at all -- used by the thin ``fr.*`` wrappers that take
    physical inputs directly rather than a theta vector."""
    m = _new_struct()
    m.ndim = ndim
    for k in range(ndim):
        m.prior[k].lo, m.prior[k].hi = -np.inf, np.inf
    m.fixed_src[:] = [float(x) for x in source_ratio]
    if sm_angles is not None:
        m.fixed_sm[:] = [float(x) for x in sm_angles]
    if mass is not None:
        m.fixed_mass[:] = [float(x) for x in mass]
    m.no_bsm = int(bool(no_bsm))
    m.llh_kind = _lib.LLH_FLAT
    if not no_bsm:
        m.dimension = int(dimension)
        tex = enum_name(texture)
        if tex in TEXTURE_ANGLES:
            m.fixed_np[:] = TEXTURE_ANGLES[tex]
        elif np_angles is not None:
            m.fixed_np[:] = [float(x) for x in np_angles]
        if loglam is not None:
            m.fixed_loglam = float(loglam)
        edges = binning_edges(binning)
        m.nbins = edges.size - 1
        for k, e in enumerate(edges):
            m.bin_edges[k] = float(e)
    return m

"""Likelihood and prior (drop-in for ``golemflavor/llh.py``), evaluated by CUDA kernels.

``ln_prob`` / ``lnprior`` / ``triangle_llh`` keep the reference's signatures and accept either one
parameter vector (returns a float, as emcee-2 / MultiNest expect) or a batch ``theta[N, ndim]``
(returns ``[N]``; this is what ``emcee``'s ``vectorize=True`` and ``mcmc.mcmc`` here use, so that a
whole stretch-move half-ensemble is scored by one kernel launch).  The likelihood is the Gaussian
flavor-ratio likelihood the reference documents as the stand-in for the proprietary GolemFit
(``README.md:76-77``), composed with the flavor functions exactly as in
``examples/inference.ipynb`` cells 21-23.
"""

import ctypes as C
from functools import partial  # noqa: F401  (parity with the reference's namespace)

import numpy as np

from . import _lib
from . import model as _model

__all__ = ['GaussianBoundedRV', 'multi_gaussian', 'lnprior', 'triangle_llh', 'ln_prob', 'LnProb', 'prepared']


def _is_tensor(x):
    return type(x).__module__.startswith('torch') and hasattr(x, 'data_ptr')


class GaussianBoundedRV(object):
    """Normalised Gaussian bounded to [lower, upper] (``llh.py:25-29``).

    The reference returns a frozen ``scipy.stats.truncnorm``; only ``logpdf`` is used on the hot
    path, so this small object exposes ``logpdf``/``pdf`` and evaluates them on the device through
    the same prior code as ``lnprior``."""

    def __init__(self, loc=0., sigma=1., lower=-np.inf, upper=np.inf):
        self.loc, self.sigma, self.lower, self.upper = float(loc), float(sigma), float(lower), float(upper)
        m = _model._new_struct()
        m.ndim = 1
        m.llh_kind = _lib.LLH_FLAT
        d = m.prior[0]
        d.lo, d.hi, d.mu, d.sigma, d.kind = self.lower, self.upper, self.loc, self.sigma, _lib.PRIOR_LIMITEDGAUSS
        self._model = _model.FlatModel(m, ('x',))

    def logpdf(self, x):
        torch = _lib.torch_cuda()
        t = _lib.to_device(x, torch).reshape(-1)
        out = torch.empty_like(t)
        _lib.check(_lib.load().gf_lnprior(self._model.ref, _lib.ptr(t), t.shape[0], 1, 1, _lib.ptr(out), _lib.stream_ptr(torch)))
        if _is_tensor(x):
            return out.reshape(x.shape)
        out = out.cpu().numpy().reshape(np.shape(x))
        return float(out) if np.ndim(x) == 0 else out

    def pdf(self, x):
        lp = self.logpdf(x)
        return lp.exp() if _is_tensor(lp) else np.exp(lp)


def multi_gaussian(fr, fr_bf, smearing, offset=-320):
    """log N_3(fr; fr_bf, smearing^2 I) + offset (``llh.py:32-54``); ``fr`` [3] or [N, 3].

    The reference takes the log of SciPy's pdf and therefore returns -inf once the pdf underflows;
    that behaviour is reproduced."""
    torch = _lib.torch_cuda()
    f = _lib.to_device(fr, torch, 3)
    batched = f.ndim > 1
    f2 = f.reshape(-1, 3)
    out = torch.empty((f2.shape[0],), dtype=torch.float64, device='cuda')
    bf = (C.c_double * 3)(*[float(x) for x in np.asarray(
        fr_bf.cpu().numpy() if _is_tensor(fr_bf) else fr_bf, dtype=np.float64)])
    _lib.check(_lib.load().gf_multi_gaussian(_lib.ptr(f2), f2.shape[0], bf, float(smearing), float(offset), 1,
                                             _lib.ptr(out), _lib.stream_ptr(torch)))
    if batched:
        out = out.reshape(f.shape[:-1])
        return out if _is_tensor(fr) else out.cpu().numpy()
    return out[0] if _is_tensor(fr) else float(out[0])


def _check_len(theta, paramset, closing='='):
    shape = tuple(theta.shape) if hasattr(theta, 'shape') else np.shape(theta)
    if not shape or shape[-1] != len(paramset):
        raise AssertionError('Length of MCMC scan is not the same as the input '
                             'params\ntheta={0}\nparamset{1}{2}'.format(theta, closing, paramset))


def _store_values(theta2, paramset):
    """The reference writes theta into Param.value (``llh.py:72-73``); keep that side effect
    (last point of a batch)."""
    last = theta2[-1].detach().cpu().numpy() if _is_tensor(theta2) else np.asarray(theta2)[-1]
    for k, p in enumerate(paramset):
        p.value = float(last[k])


class LnProb(object):
    """Prepared log-posterior: the model is flattened once, each call is one kernel launch.

    ``LnProb(args, asimov_paramset, llh_paramset)(theta)`` == ``ln_prob(theta, args, asimov_paramset,
    llh_paramset)``; picklable inputs are kept so that it can be handed to samplers."""

    def __init__(self, args, asimov_paramset, llh_paramset, likelihood=None):
        self.model = _model.flatten(args, asimov_paramset, llh_paramset, likelihood=likelihood)
        self.ndim = self.model.ndim

    def evaluate(self, theta, want_fr=False, want_status=False):
        """theta [N, ndim] (NumPy or CUDA tensor) -> lnprob [N] (+ fr [N, 3], status [N]) as CUDA tensors."""
        torch = _lib.torch_cuda()
        th = _lib.to_device(theta, torch, self.ndim).reshape(-1, self.ndim)
        # one operator call: torch.ops.golemflavor.lnprob -> gf_lnprob on torch's current stream
        out, fr, st = _lib.torch_ops().lnprob(th, self.model.blob, bool(want_fr), bool(want_status))
        res = (out,) + ((fr,) if want_fr else ()) + ((st,) if want_status else ())
        return res if len(res) > 1 else out

    def evaluate_host(self, theta, out=None, fr=None, status=None):
        """theta [N, ndim] float64 NumPy (ideally page-locked) -> lnprob [N] NumPy, through the
        chunked H2D / kernel / D2H pipeline of ``gf_lnprob_host``.  Optional output arrays
        ``fr`` [N, 3] float64 and ``status`` [N] uint8 are filled when given."""
        _lib.torch_cuda()
        th = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, self.ndim)
        n = th.shape[0]
        if out is None:
            out = np.empty((n,), dtype=np.float64)
        for name, arr, shape, dt in (('out', out, (n,), np.float64), ('fr', fr, (n, 3), np.float64), ('status', status, (n,), np.uint8)):
            if arr is not None and (arr.shape != shape or arr.dtype != dt or not arr.flags['C_CONTIGUOUS']):
                raise ValueError('{0} must be a C-contiguous {1} array of shape {2}'.format(name, np.dtype(dt).name, shape))
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _lib.check(_lib.load().gf_lnprob_host(self.model.ref, ptr(th), n, ptr(out), ptr(fr), ptr(status)))
        return out

    def __call__(self, theta):
        batched = (theta.ndim if hasattr(theta, 'ndim') else np.ndim(theta)) > 1
        if _is_tensor(theta):
            out = self.evaluate(theta)
            return out if batched else out[0]
        # host data in, host data out: one C call (staging, H2D, kernel, D2H inside the library) -- for the
        # emcee-sized batches of a host-side sampler the per-call latency is what counts
        out = self.evaluate_host(theta)
        return out if batched else float(out[0])


def lnprior(theta, paramset):
    """Box prior in ``Param.ranges`` plus (truncated-)Gaussian terms (``llh.py:65-91``)."""
    _check_len(theta, paramset)
    torch = _lib.torch_cuda()
    m = _model._new_struct()
    m.ndim = len(paramset)
    m.llh_kind = _lib.LLH_FLAT
    _model._set_priors(m, paramset)
    fm = _model.FlatModel(m, [p.name for p in paramset])
    th = _lib.to_device(theta, torch, fm.ndim)
    batched = th.ndim > 1
    th2 = th.reshape(-1, fm.ndim)
    _store_values(th2, paramset)
    out = _lib.torch_ops().lnprior(th2, fm.blob)
    if _is_tensor(theta):
        return out if batched else out[0]
    out = out.cpu().numpy()
    return out if batched else float(out[0])


def triangle_llh(theta, args, asimov_paramset, llh_paramset):
    """Log likelihood for theta (``llh.py:94-118`` with the Gaussian likelihood in place of
    GolemFit): llh = multi_gaussian(measured fr(theta), injected fr, smearing)."""
    _check_len(theta, llh_paramset, closing=']')
    torch = _lib.torch_cuda()
    fn = prepared(args, asimov_paramset, llh_paramset)
    th = _lib.to_device(theta, torch, fn.ndim)
    batched = th.ndim > 1
    th2 = th.reshape(-1, fn.ndim)
    _store_values(th2, llh_paramset)
    n = th2.shape[0]
    fr = torch.empty((n, 3), dtype=torch.float64, device='cuda')
    st = torch.empty((n,), dtype=torch.uint8, device='cuda')
    lib = _lib.load()
    _lib.check(lib.gf_flux_averaged_fr(fn.model.ref, _lib.ptr(th2), n, fn.ndim, 1, _lib.ptr(fr), _lib.ptr(st), _lib.stream_ptr(torch)))
    s = fn.model.struct
    if s.llh_kind == _lib.LLH_FLAT:
        out = torch.full((n,), float(s.llh_const), dtype=torch.float64, device='cuda')
    else:
        out = torch.empty((n,), dtype=torch.float64, device='cuda')
        _lib.check(lib.gf_multi_gaussian(_lib.ptr(fr), n, s.fr_bf, float(s.smearing), float(s.offset),
                                         int(s.emulate_underflow), _lib.ptr(out), _lib.stream_ptr(torch)))
    if _is_tensor(theta):
        return out if batched else out[0]
    out = out.cpu().numpy()
    return out if batched else float(out[0])


_PREPARED = {}


def _fingerprint(args, asimov_paramset, llh_paramset):
    """Everything ``model.flatten`` reads, as a hashable key: the reference calls ``ln_prob`` with the same three
    closure arguments millions of times (``scripts/fr.py:182-187``), so the flattened model is cached -- but keyed by
    CONTENT, because the reference's own code mutates ``Param.value`` / ``args`` between calls."""
    def num(x):
        if x is None:
            return None
        a = np.asarray(x, dtype=np.float64)
        return float(a) if a.ndim == 0 else a.tobytes()
    pkey = tuple((p.name, num(p.ranges), num(p.nominal_value), num(p.std), str(p.prior), str(p.tag)) for p in llh_paramset)
    akey = tuple((p.name, num(p.value), num(p.std), str(p.tag)) for p in (asimov_paramset or ()))
    names = ('source_ratio', 'no_bsm', 'fixed_scale', 'texture', 'dimension', 'binning', 'likelihood', 'llh_const',
             'injected_ratio', 'smearing', 'llh_offset', 'emulate_underflow')
    gkey = tuple(num(getattr(args, k)) if k in ('source_ratio', 'binning', 'injected_ratio', 'fixed_scale', 'smearing', 'llh_const', 'llh_offset')
                 and getattr(args, k, None) is not None else str(getattr(args, k, None)) for k in names)
    return pkey, akey, gkey


def prepared(args, asimov_paramset, llh_paramset):
    """The cached ``LnProb`` of a closure (a few entries, oldest dropped first)."""
    key = _fingerprint(args, asimov_paramset, llh_paramset)
    fn = _PREPARED.get(key)
    if fn is None:
        if len(_PREPARED) >= 16:
            _PREPARED.pop(next(iter(_PREPARED)))
        fn = _PREPARED[key] = LnProb(args, asimov_paramset, llh_paramset)
    return fn


def ln_prob(theta, args, asimov_paramset, llh_paramset):
    """lnprior + triangle_llh with the -inf short-circuit (``llh.py:121-130``), fused into one kernel.

    No deep copies are needed (``llh.py:122-123``): nothing on the device path mutates the ParamSets.  The flattened
    model is cached across calls (``prepared``), so a call costs one fingerprint of the closure plus one launch."""
    _check_len(theta, llh_paramset)
    return prepared(args, asimov_paramset, llh_paramset)(theta)

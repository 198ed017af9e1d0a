"""Flavor functions (drop-in for ``golemflavor/fr.py``), evaluated by CUDA kernels.

Every function keeps the reference's signature and scalar behaviour and additionally accepts a
leading batch axis.  NumPy / sequence input gives NumPy output; a CUDA ``torch.Tensor`` input gives
a CUDA tensor output (nothing leaves the device).  The arithmetic is IEEE fp64 on the GPU (the
reference uses x87 80-bit ``np.float128``, ``fr.py:22-23``); there is no CPU implementation.
"""

import ctypes as C

import numpy as np

from . import _lib
from . import model as _model
from .enums import Texture, enum_name
from .model import MASS_EIGENVALUES, SCALE_BOUNDARIES, NUFIT_ANGLES  # noqa: F401  (re-exported constants)

__all__ = ['angles_to_fr', 'angles_to_u', 'cardano_eqn', 'normalize_fr', 'normalise_fr', 'fr_to_angles',
           'params_to_BSMu', 'flux_averaged_BSMu', 'test_unitarity', 'u_to_fr', 'NUFIT_U',
           'MASS_EIGENVALUES', 'SCALE_BOUNDARIES']


def _is_tensor(x):
    return type(x).__module__.startswith('torch') and hasattr(x, 'data_ptr')


def _finish(out, like_tensor, batched):
    """Return convention: tensor in -> tensor out, else NumPy; scalar call -> no batch axis."""
    if not batched:
        out = out[0]
    return out if like_tensor else out.cpu().numpy()


def _view_c128(t, torch):
    return torch.view_as_complex(t.reshape(t.shape[:-1] + (3, 3, 2)))


def _u_to_f64(u, torch):
    """[..., 3, 3] complex (NumPy / tensor) -> contiguous float64 CUDA tensor [n, 18]."""
    if _is_tensor(u):
        t = u.to(device='cuda')
        if not t.is_complex():
            t = t.to(torch.float64).to(torch.complex128)
        t = t.to(torch.complex128)
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(u).astype(np.complex128))).cuda()
    if t.shape[-2:] != (3, 3):
        raise ValueError('Input matrix should be a square and dimension 3, got\n{0}'.format(u))
    return torch.view_as_real(t.contiguous()).reshape(-1, 18).contiguous()


def angles_to_fr(src_angles):
    """(sin^4 phi, cos 2psi) -> (f_e, f_mu, f_tau).  ``fr.py:82-113``.

    >>> angles_to_fr((0.3, 0.4))
    (0.38340579025361626, 0.1643167672515498, 0.45227744249483387)
    """
    torch = _lib.torch_cuda()
    a = _lib.to_device(src_angles, torch, 2)
    batched = a.ndim > 1
    a2 = a.reshape(-1, 2)
    out = torch.empty((a2.shape[0], 3), dtype=torch.float64, device='cuda')
    _lib.check(_lib.load().gf_angles_to_fr(_lib.ptr(a2), a2.shape[0], _lib.ptr(out), _lib.stream_ptr(torch)))
    if not batched and not _is_tensor(src_angles):
        return tuple(float(x) for x in out[0].cpu())
    return _finish(out.reshape(a.shape[:-1] + (3,)) if batched else out, _is_tensor(src_angles), batched)


def angles_to_u(bsm_angles):
    """(s12^2, c13^4, s23^2, dcp) -> 3x3 unitary R23 R13(dcp) R12.  ``fr.py:116-162``."""
    torch = _lib.torch_cuda()
    a = _lib.to_device(bsm_angles, torch, 4)
    batched = a.ndim > 1
    a2 = a.reshape(-1, 4)
    out = torch.empty((a2.shape[0], 18), dtype=torch.float64, device='cuda')
    _lib.check(_lib.load().gf_angles_to_u(_lib.ptr(a2), a2.shape[0], _lib.ptr(out), _lib.stream_ptr(torch)))
    u = _view_c128(out, torch)
    if batched:
        u = u.reshape(a.shape[:-1] + (3, 3))
    return _finish(u, _is_tensor(bsm_angles), batched)


def cardano_eqn(ham):
    """Eigenvector matrix (columns, ascending eigenvalue) of a 3x3 Hermitian matrix.

    The reference (``fr.py:170-237``) uses the analytic Cardano / cross-product formulas in 80-bit
    arithmetic, which lose accuracy for hierarchical spectra; this evaluates a cyclic complex
    Jacobi iteration in fp64.  Eigenvectors agree up to column order and a phase per column."""
    if np.shape(ham)[-2:] != (3, 3):
        raise ValueError('Input matrix should be a square and dimension 3, got\n{0}'.format(ham))
    torch = _lib.torch_cuda()
    h = _u_to_f64(ham, torch)
    batched = np.ndim(ham) > 2
    out = torch.empty_like(h)
    _lib.check(_lib.load().gf_eigvec_herm3(_lib.ptr(h), h.shape[0], _lib.ptr(out), None, None, _lib.stream_ptr(torch)))
    v = _view_c128(out, torch)
    if batched:
        v = v.reshape(tuple(np.shape(ham)[:-2]) + (3, 3))
    return _finish(v, _is_tensor(ham), batched)


def eigh3(ham):
    """(eigenvalues ascending [..., 3], eigenvectors [..., 3, 3], status [...]) of Hermitian 3x3 matrices."""
    torch = _lib.torch_cuda()
    h = _u_to_f64(ham, torch)
    n = h.shape[0]
    vec = torch.empty_like(h)
    val = torch.empty((n, 3), dtype=torch.float64, device='cuda')
    st = torch.empty((n,), dtype=torch.uint8, device='cuda')
    _lib.check(_lib.load().gf_eigvec_herm3(_lib.ptr(h), n, _lib.ptr(vec), _lib.ptr(val), _lib.ptr(st), _lib.stream_ptr(torch)))
    lead = tuple(np.shape(ham)[:-2])
    out = (val.reshape(lead + (3,)), _view_c128(vec, torch).reshape(lead + (3, 3)), st.reshape(lead))
    return out if _is_tensor(ham) else tuple(o.cpu().numpy() for o in out)


def normalize_fr(fr):
    """x / sum(x) (``fr.py:240-259``); host arithmetic on three numbers, or along the last axis of a tensor."""
    if _is_tensor(fr):
        return fr / fr.sum(dim=-1, keepdim=True)
    fr = np.array(fr)
    if fr.ndim > 1:
        return fr / fr.sum(axis=-1, keepdims=True)
    return fr / float(np.sum(fr))


normalise_fr = normalize_fr


def fr_to_angles(ratios):
    """Inverse of ``angles_to_fr`` (``fr.py:289-310``).  Host-side set-up helper (asimov params)."""
    f0, _, f2 = normalize_fr(ratios)
    sphi2 = 1.0 - f2
    if sphi2 == 0.:
        return (0., 0.)
    cpsi2 = f0 / sphi2
    return (sphi2 ** 2, float(np.cos(np.arccos(np.sqrt(cpsi2)) * 2)))


def test_unitarity(x, prnt=False, rse=False, epsilon=None):
    """|x x^dagger| with the optional assertion of ``fr.py:461-499`` (a 27-flop host check on one matrix)."""
    x = np.asarray(x.cpu().numpy() if _is_tensor(x) else x)
    f = np.abs(np.dot(x, x.conj().T))
    if prnt:
        print('Unitarity test:\n{0}'.format(f))
    if rse:
        if not abs(np.trace(f) - 3.) < epsilon or not abs(np.sum(f) - 3.) < epsilon:
            raise AssertionError('Matrix is not unitary!\nx\n{0}\ntest u\n{1}'.format(x, f))
    return f


test_unitarity.__test__ = False


def u_to_fr(source_fr, matrix):
    """Measured composition fr_b = sum_ai |U_ai|^2 |U_bi|^2 s_a / sum(s).  ``fr.py:502-536``.

    ``matrix`` may be [3, 3] or [N, 3, 3]; ``source_fr`` [3] (shared) or [N, 3]."""
    torch = _lib.torch_cuda()
    u = _u_to_f64(matrix, torch)
    n = u.shape[0]
    s = _lib.to_device(source_fr, torch, 3).reshape(-1, 3)
    if s.shape[0] not in (1, n):
        raise ValueError('source_fr has {0} rows, matrix has {1}'.format(s.shape[0], n))
    stride = 0 if s.shape[0] == 1 else 3
    out = torch.empty((n, 3), dtype=torch.float64, device='cuda')
    _lib.check(_lib.load().gf_u_to_fr(_lib.ptr(s), stride, _lib.ptr(u), n, _lib.ptr(out), _lib.stream_ptr(torch)))
    batched = np.ndim(matrix) > 2
    if batched:
        out = out.reshape(tuple(np.shape(matrix)[:-2]) + (3,))
    return _finish(out, _is_tensor(matrix), batched)


class _Lazy(object):
    """NUFIT_U is a device computation; evaluate it on first use so that importing the
    module does not need a GPU."""
    _value = None

    def _get(self):
        if _Lazy._value is None:
            _Lazy._value = angles_to_u(NUFIT_ANGLES)
        return _Lazy._value

    def __array__(self, dtype=None, copy=None):
        v = self._get()
        return v.astype(dtype) if dtype is not None else v

    def __getitem__(self, i):
        return self._get()[i]

    shape = (3, 3)
    ndim = 2

    def conj(self):
        return self._get().conj()


NUFIT_U = _Lazy()  # fr.py:313


def _texture_tuple(bsm_angles, texture):
    """(np_s12_2, np_c13_4, np_s23_2, np_dcp, logLam) for a texture (``fr.py:367-378``)."""
    name = enum_name(texture)
    if _is_tensor(bsm_angles):
        arr = bsm_angles.to(device='cuda').double()
    else:
        arr = np.asarray(bsm_angles, dtype=np.float64)
    if name in _model.TEXTURE_ANGLES:
        scale = arr[..., -1] if arr.ndim else arr
        tex = _model.TEXTURE_ANGLES[name]
        if _is_tensor(arr):
            import torch
            cols = [torch.full_like(scale, v) for v in tex] + [scale]
            return torch.stack(cols, dim=-1)
        return np.stack([np.full(np.shape(scale), v) for v in tex] + [scale], axis=-1)
    if arr.shape[-1] != 5:
        raise ValueError('bsm_angles needs 5 entries (4 mixing angles + scale) for Texture.NONE, got {0}'.format(arr.shape))
    return arr


def params_to_BSMu(bsm_angles, dim, energy, mass_eigenvalues=MASS_EIGENVALUES, sm_u=NUFIT_U, no_bsm=False,
                   texture=Texture.NONE, check_uni=True, epsilon=1e-7):
    """Eigenvector matrix of H = U diag(0, m21, m3x) U^+ / (2E) + E^(dim-3) N diag(0, L/100, L) N^+.

    ``fr.py:317-400``.  Batched over a leading axis of ``bsm_angles`` (and optionally ``energy``,
    ``mass_eigenvalues``, ``sm_u``).  Columns are ordered by ascending eigenvalue.  With ``check_uni``
    a non-finite result raises ``AssertionError`` like the reference's unitarity assertion."""
    if np.shape(sm_u)[-2:] != (3, 3):
        raise ValueError('Input matrix should be a square and dimension 3, got\n{0}'.format(sm_u))
    torch = _lib.torch_cuda()
    like_tensor = _is_tensor(bsm_angles)
    b = _lib.to_device(_texture_tuple(bsm_angles, texture), torch, 5)
    batched = b.ndim > 1
    b2 = b.reshape(-1, 5)
    n = b2.shape[0]
    e = _lib.to_device(energy, torch).reshape(-1)
    if e.shape[0] == 1 and n > 1:
        e = e.expand(n).contiguous()
    if e.shape[0] != n:
        raise ValueError('energy has {0} entries for {1} points'.format(e.shape[0], n))
    mass = _lib.to_device(mass_eigenvalues, torch, 2).reshape(-1, 2)
    smu = _u_to_f64(sm_u, torch)
    for name, t in (('mass_eigenvalues', mass), ('sm_u', smu)):
        if t.shape[0] not in (1, n):
            raise ValueError('{0} has {1} rows for {2} points'.format(name, t.shape[0], n))
    vec = torch.empty((n, 18), dtype=torch.float64, device='cuda')
    st = torch.empty((n,), dtype=torch.uint8, device='cuda')
    _lib.check(_lib.load().gf_params_to_bsmu(
        _lib.ptr(b2), int(dim), _lib.ptr(e), _lib.ptr(mass), 0 if mass.shape[0] == 1 else 2,
        _lib.ptr(smu), 0 if smu.shape[0] == 1 else 18, int(bool(no_bsm)), float(epsilon), n,
        _lib.ptr(vec), _lib.ptr(st), _lib.stream_ptr(torch)))
    if check_uni and bool((st & _lib.ST_NON_UNITARY).any()):
        raise AssertionError('Matrix is not unitary! (non-finite eigenvectors for {0} of {1} points)'.format(
            int((st & _lib.ST_NON_UNITARY).ne(0).sum()), n))
    v = _view_c128(vec, torch)
    if batched:
        v = v.reshape(b.shape[:-1] + (3, 3))
    return _finish(v, like_tensor, batched)


def flux_averaged_BSMu(theta, args, spectral_index, llh_paramset):
    """Energy-bin-averaged measured composition (``fr.py:403-458``); ``theta`` [ndim] or [N, ndim].

    ``spectral_index`` is accepted for signature compatibility: the source flux normalisation
    E^gamma cancels exactly in ``u_to_fr`` (``fr.py:416-419, 535``).  Like the reference this writes
    theta into ``llh_paramset`` (last point of a batch)."""
    del spectral_index
    th_np_shape = tuple(theta.shape) if hasattr(theta, 'shape') else np.shape(theta)
    if th_np_shape[-1] != len(llh_paramset):
        raise AssertionError('Length of MCMC scan is not the same as the input '
                             'params\ntheta={0}\nparamset]{1}'.format(theta, llh_paramset))
    torch = _lib.torch_cuda()
    fm = _model.flatten(args, None, llh_paramset, likelihood='FLAT')
    th = _lib.to_device(theta, torch, fm.ndim)
    batched = th.ndim > 1
    th2 = th.reshape(-1, fm.ndim)
    n = th2.shape[0]
    last = th2[-1].cpu().numpy()
    for k, p in enumerate(llh_paramset):
        p.value = float(last[k])
    out, st = _lib.torch_ops().flux_averaged_fr(th2, fm.blob)
    bad = st & (_lib.ST_NON_UNITARY | _lib.ST_NON_FINITE)
    if bool(bad.any()):
        raise AssertionError('Matrix is not unitary! ({0} of {1} points)'.format(int(bad.ne(0).sum()), n))
    if batched:
        out = out.reshape(th.shape[:-1] + (3,))
    return _finish(out, _is_tensor(theta), batched)

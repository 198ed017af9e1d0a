"""Parameter containers with the public interface of ``golemflavor/param.py:24-214``.

``Param`` carries the metadata of one sampled quantity (value, nominal value, box
``ranges``, prior category, seed box, Gaussian width, tag); ``ParamSet`` is an ordered,
name-addressable sequence of them.  They stay on the host: ``model.flatten`` turns a
pair of ParamSets plus the ``args`` namespace into the flat ``gf_model`` struct that the
CUDA kernels take as a kernel parameter, and ``ParamSet.prior_table`` is the per-column
prior description (``gf_prior_dim``) that goes into it.
"""

from collections.abc import Sequence
from copy import deepcopy

import numpy as np

from .enums import ParamTag, PriorsCateg, enum_name


def _member(enum_cls, value, default):
    """Our Enum members, the reference's (matched by name) or plain strings -> a member of ``enum_cls``."""
    if value is None:
        return default
    if isinstance(value, enum_cls):
        return value
    try:
        return enum_cls[enum_name(value)]
    except KeyError:
        raise AssertionError('{0!r} is not a member of {1}'.format(value, enum_cls.__name__))


# per-attribute normalisation applied on assignment (the reference does the same through a property per
# attribute, ``param.py:40-91``): boxes become lists, categories become Enum members, a missing TeX label falls
# back to the name
_NORMALISE = {
    'ranges': lambda p, v: list(v),
    'seed': lambda p, v: p._slots.get('seed') if v is None else list(v),
    'prior': lambda p, v: _member(PriorsCateg, v, PriorsCateg.UNIFORM),
    'tag': lambda p, v: _member(ParamTag, v, ParamTag.NONE),
    'tex': lambda p, v: v if v is not None else r'{\rm %s}' % p.name,
}
_PRESENT = {
    'ranges': tuple,
    'seed': lambda v: v,   # resolved in __getattr__ (falls back to the ranges)
    'tex': lambda v: r'{0}'.format(v),
}


class Param(object):
    """One parameter: ``Param(name, value, ranges, prior=None, seed=None, std=None, tex=None, tag=None)``.

    ``prior`` defaults to ``PriorsCateg.UNIFORM``, ``tag`` to ``ParamTag.NONE``, ``seed`` (the box walkers
    start from, ``mcmc.flat_seed``) to ``ranges``; ``nominal_value`` keeps a copy of the initial ``value``,
    which is what Gaussian priors are centred on (``llh.py:82-90``)."""

    def __init__(self, name, value, ranges, prior=None, seed=None, std=None, tex=None, tag=None):
        object.__setattr__(self, '_slots', {})
        self.name = name
        self.value = value
        self.nominal_value = deepcopy(value)
        for field, given in (('prior', prior), ('ranges', ranges), ('seed', seed), ('std', std), ('tex', tex), ('tag', tag)):
            setattr(self, field, given)

    def __setattr__(self, field, given):
        norm = _NORMALISE.get(field)
        self._slots[field] = norm(self, given) if norm else given

    def __getattr__(self, field):
        try:
            stored = object.__getattribute__(self, '_slots')[field]
        except (KeyError, AttributeError):
            raise AttributeError(field)
        if field == 'seed':
            return self.ranges if stored is None else tuple(stored)
        show = _PRESENT.get(field)
        return show(stored) if show else stored

    def __deepcopy__(self, memo):
        twin = Param.__new__(Param)
        object.__setattr__(twin, '_slots', deepcopy(self._slots, memo))
        return twin

    def __getstate__(self):
        return dict(self._slots)

    def __setstate__(self, state):
        object.__setattr__(self, '_slots', dict(state))

    def __repr__(self):
        return 'Param({0!r}, value={1!r}, ranges={2!r}, prior={3}, tag={4})'.format(
            self.name, self.value, self.ranges, self.prior.name, self.tag.name)


def _flatten_args(args):
    for arg in args:
        if isinstance(arg, Param):
            yield arg
        else:
            for p in arg:
                yield p


class ParamSet(Sequence):
    """Ordered container of ``Param``, addressable by position or by name.

    The tuple-valued views ``names, labels, values, nominal_values, seeds, ranges, stds, tags`` list one
    attribute over all parameters in order; ``from_tag`` selects by tag keeping that order (which is what
    identifies e.g. the four mixing coordinates to the physics, ``fr.py:421-435``)."""

    _VIEWS = {'names': 'name', 'labels': 'tex', 'values': 'value', 'nominal_values': 'nominal_value', 'seeds': 'seed',
              'ranges': 'ranges', 'stds': 'std', 'tags': 'tag'}

    def __init__(self, *args):
        members = list(_flatten_args(args))
        assert all(isinstance(p, Param) for p in members), 'All params must be of type "Param"'
        seen, twice = set(), []
        for p in members:
            if p.name in seen:
                twice.append(p.name)
            seen.add(p.name)
        if twice:
            raise ValueError('Duplicate definitions found for param(s): ' + ', '.join(map(str, sorted(set(twice)))))
        self._members = members

    # -- sequence protocol
    def __len__(self):
        return len(self._members)

    def __iter__(self):
        return iter(self._members)

    def __getitem__(self, key):
        if isinstance(key, str):
            for p in self._members:
                if p.name == key:
                    return p
            raise KeyError(key)
        return self._members[key]

    def __getattr__(self, view):
        attr = ParamSet._VIEWS.get(view)
        if attr is None:
            raise AttributeError(view)
        return tuple(getattr(p, attr) for p in self._members)

    def __str__(self):
        rows = ('== {0:<15} = {1:<15}, tag={2:<15}\n'.format(p.name, p.value, str(p.tag)) for p in self._members)
        return '\n' + ''.join(rows)

    @property
    def params(self):
        return self._members

    def to_dict(self):
        return dict(zip(self.names, self.values))

    # -- selections (all return new ParamSets sharing the Param objects, like the reference)
    def from_tag(self, tag, values=False, index=False, invert=False):
        """Parameters carrying -- with ``invert``: not carrying -- one of the given tag(s), in paramset order;
        ``values`` / ``index`` return their values / positions instead (``param.py:185-199``)."""
        assert not (values and index)
        wanted = {enum_name(t) for t in np.atleast_1d(tag)}
        picked = [(pos, p) for pos, p in enumerate(self._members) if (p.tag.name in wanted) is not bool(invert)]
        if values:
            return tuple(p.value for _, p in picked)
        if index:
            return tuple(pos for pos, _ in picked)
        return ParamSet(p for _, p in picked)

    def remove_params(self, params):
        gone = frozenset(params.names)
        return ParamSet(p for p in self._members if p.name not in gone)

    def extend(self, more):
        if isinstance(more, Param):
            more = [more]
        elif not isinstance(more, ParamSet):
            more = []
        return ParamSet(self._members, more)

    # -- what the kernels need
    def prior_table(self):
        """Per-column prior description ``(lo, hi, mu, sigma, kind)`` as arrays: the box, the centre and width of
        the (truncated) Gaussian where there is one (``llh.py:74-90``), and the ``PriorsCateg`` name."""
        lo = np.array([p.ranges[0] for p in self._members], dtype=np.float64)
        hi = np.array([p.ranges[1] for p in self._members], dtype=np.float64)
        mu = np.array([p.nominal_value for p in self._members], dtype=np.float64)
        sigma = np.array([np.nan if p.std is None else p.std for p in self._members], dtype=np.float64)
        kind = [p.prior.name for p in self._members]
        return lo, hi, mu, sigma, kind

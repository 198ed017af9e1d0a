"""Parameter containers (mirror of ``golemflavor/param.py:24-214``).

``Param`` carries the metadata of one sampled quantity (value, nominal value, box
``ranges``, prior category, seed box, Gaussian width, tag); ``ParamSet`` is an ordered,
name-addressable sequence of them.  They stay on the host: ``model.flatten`` turns a
pair of ParamSets plus the ``args`` namespace into the flat ``gf_model`` struct that the
CUDA kernels take as a kernel parameter.
"""

from collections.abc import Sequence
from copy import deepcopy

import numpy as np

from .enums import ParamTag, PriorsCateg, enum_name


def _coerce(enum_cls, value, default):
    """Accept our Enum members, the reference's (by name) or strings."""
    if value is None:
        return default
    if isinstance(value, enum_cls):
        return value
    try:
        return enum_cls[enum_name(value)]
    except KeyError:
        raise AssertionError('{0!r} is not a member of {1}'.format(value, enum_cls.__name__))


class Param(object):
    """One parameter (``param.py:24-91``)."""

    def __init__(self, name, value, ranges, prior=None, seed=None, std=None, tex=None, tag=None):
        self.name = name
        self.value = value
        self.nominal_value = deepcopy(value)
        self.prior = prior
        self.ranges = ranges
        self._seed = None
        self.seed = seed
        self.std = std
        self.tex = tex
        self.tag = tag

    ranges = property(lambda self: tuple(self._ranges))

    @ranges.setter
    def ranges(self, values):
        self._ranges = list(values)

    prior = property(lambda self: self._prior)

    @prior.setter
    def prior(self, value):
        self._prior = _coerce(PriorsCateg, value, PriorsCateg.UNIFORM)

    @property
    def seed(self):
        return self.ranges if self._seed is None else tuple(self._seed)

    @seed.setter
    def seed(self, values):
        if values is not None:
            self._seed = list(values)

    tex = property(lambda self: r'{0}'.format(self._tex))

    @tex.setter
    def tex(self, t):
        self._tex = t if t is not None else r'{\rm %s}' % self.name

    tag = property(lambda self: self._tag)

    @tag.setter
    def tag(self, t):
        self._tag = _coerce(ParamTag, t, ParamTag.NONE)

    def __repr__(self):
        return 'Param({0!r}, value={1!r}, ranges={2!r}, prior={3}, tag={4})'.format(
            self.name, self.value, self.ranges, self.prior.name, self.tag.name)


class ParamSet(Sequence):
    """Ordered container of ``Param`` (``param.py:94-214``)."""

    def __init__(self, *args):
        seq = []
        for arg in args:
            if isinstance(arg, Param):
                seq.append(arg)
            else:
                seq.extend(arg)
        names = [p.name for p in seq]
        dup = sorted({n for n in names if names.count(n) > 1})
        if dup:
            raise ValueError('Duplicate definitions found for param(s): ' + ', '.join(map(str, dup)))
        assert all(isinstance(p, Param) for p in seq), 'All params must be of type "Param"'
        self._params = seq

    def __len__(self):
        return len(self._params)

    def __getitem__(self, i):
        if isinstance(i, str):
            return self._by_name[i]
        return self._params[i]

    def __iter__(self):
        return iter(self._params)

    def __str__(self):
        return '\n' + ''.join('== {0:<15} = {1:<15}, tag={2:<15}\n'.format(p.name, p.value, str(p.tag))
                              for p in self._params)

    @property
    def _by_name(self):
        return {p.name: p for p in self._params}

    def _column(self, attr):
        return tuple(getattr(p, attr) for p in self._params)

    names = property(lambda self: self._column('name'))
    labels = property(lambda self: self._column('tex'))
    values = property(lambda self: self._column('value'))
    nominal_values = property(lambda self: self._column('nominal_value'))
    seeds = property(lambda self: self._column('seed'))
    ranges = property(lambda self: self._column('ranges'))
    stds = property(lambda self: self._column('std'))
    tags = property(lambda self: self._column('tag'))
    params = property(lambda self: self._params)

    def to_dict(self):
        return {p.name: p.value for p in self._params}

    def from_tag(self, tag, values=False, index=False, invert=False):
        """Sub-set carrying (or, with ``invert``, not carrying) ``tag`` (``param.py:185-199``)."""
        assert not (values and index)
        wanted = {enum_name(t) for t in np.atleast_1d(tag)}
        hits = [(i, p) for i, p in enumerate(self._params) if (p.tag.name in wanted) != bool(invert)]
        if values:
            return tuple(p.value for _, p in hits)
        if index:
            return tuple(i for i, _ in hits)
        return ParamSet([p for _, p in hits])

    def remove_params(self, params):
        drop = set(params.names)
        return ParamSet([p for p in self._params if p.name not in drop])

    def extend(self, p):
        seq = list(self._params)
        if isinstance(p, Param):
            seq.append(p)
        elif isinstance(p, ParamSet):
            seq.extend(p.params)
        return ParamSet(seq)

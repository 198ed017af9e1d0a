"""Enumerations of the flavor-ratio analysis (mirror of ``golemflavor/enums.py:19-63``).

Member names and values are those of the reference so that code written against it
(``Texture.OET``, ``ParamTag.SM_ANGLES``, ``PriorsCateg.LIMITEDGAUSS`` ...) keeps
working.  ``Likelihood.GAUSSIAN`` and ``Likelihood.FLAT`` are additions: the reference
enum only names the proprietary GolemFit likelihoods (``enums.py:23-25``) while its
documented stand-in (``README.md:76-77``) and its MC scans (``scripts/mc_unitary.py:121-131``)
use a Gaussian and a flat likelihood respectively.
"""

from enum import Enum


def str_enum(x):
    """'Texture.OET' -> 'OET' (``enums.py:15-16``)."""
    return str(x).rsplit('.', 1)[-1]


DataType = Enum('DataType', ['REAL', 'ASIMOV', 'REALISATION'])
Likelihood = Enum('Likelihood', ['GOLEMFIT', 'GF_FREQ', 'GAUSSIAN', 'FLAT'])
ParamTag = Enum('ParamTag', ['NUISANCE', 'SM_ANGLES', 'MMANGLES', 'SCALE', 'SRCANGLES', 'BESTFIT', 'NONE'])
PriorsCateg = Enum('PriorsCateg', ['UNIFORM', 'GAUSSIAN', 'LIMITEDGAUSS'])
MCMCSeedType = Enum('MCMCSeedType', ['UNIFORM', 'GAUSSIAN'])
StatCateg = Enum('StatCateg', ['BAYESIAN', 'FREQUENTIST'])
SteeringCateg = Enum('SteeringCateg', ['P2_0', 'P2_1'])
Texture = Enum('Texture', ['OEU', 'OET', 'OUT', 'NONE'])


def enum_name(obj, default='NONE'):
    """Upper-case member name of an Enum (ours or the reference's) or a plain string."""
    if obj is None:
        return default
    return str(getattr(obj, 'name', obj)).rsplit('.', 1)[-1].upper()

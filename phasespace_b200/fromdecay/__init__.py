"""``GenMultiDecay``: a container of decay modes, and the bridge to DecayLanguage dicts.

Mirror of the reference's ``phasespace.fromdecay`` (fromdecay/__init__.py:11-24).  Unlike the
reference this imports without ``zfit``, ``zfit_physics`` or ``particle``: the resonance shapes are
device samplers (mass_functions.py) and ``from_dict`` falls back to a small built-in PDG table when
``particle`` is not installed.
"""

from .genmultidecay import GenMultiDecay  # noqa: F401

__all__ = ("GenMultiDecay",)


def __dir__():
    return __all__

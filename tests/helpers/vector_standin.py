"""A minimal stand-in for the third-party ``vector`` package (scikit-hep/vector), for TESTS ONLY.

``vector`` is an optional dependency of the reference's interop path (phasespace.py:676-695 converts a ``vector``
momentum object given as ``boost_to``; phasespace.py:817-835 returns ``vector.array`` objects for
``as_vectors=True``) and it is not installed in this image.  This module implements exactly the surface those two
code paths and the reference's tests touch -- ``vector.Vector`` / ``vector.Momentum`` / ``vector.Lorentz`` as
``isinstance`` targets, ``vector.array({...})`` with ``px py pz energy`` (momentum, Lorentz), ``x y z t`` (Lorentz, not a
momentum) or ``px py pz`` (momentum, not Lorentz) coordinates, the attributes ``px py pz E energy`` and
``to_pxpypzenergy()`` -- so that the product's own conversion code runs in the tests instead of being skipped.  It is
installed as ``sys.modules["vector"]`` by a fixture only when the real package is missing.
"""

import numpy as np


class Vector:
    pass


class Momentum:
    pass


class Lorentz:
    pass


class _Array(Vector):
    def __init__(self, fields):
        self._fields = {k: np.asarray(v, dtype=np.float64) for k, v in fields.items()}

    def __len__(self):
        return len(next(iter(self._fields.values())))


class MomentumNumpy4D(_Array, Momentum, Lorentz):
    px = property(lambda self: self._fields["px"])
    py = property(lambda self: self._fields["py"])
    pz = property(lambda self: self._fields["pz"])
    energy = property(lambda self: self._fields["energy"])
    E = energy

    def to_pxpypzenergy(self):
        return self


class VectorNumpy4D(_Array, Lorentz):
    """x, y, z, t: Lorentz but not a momentum."""


class MomentumNumpy3D(_Array, Momentum):
    """px, py, pz: a momentum but not Lorentz."""


def array(fields):
    fields = dict(fields)
    for alias in ("E", "e"):  # the real package accepts these spellings of the energy (tests/test_physics.py:193-199)
        if alias in fields:
            fields["energy"] = fields.pop(alias)
    keys = set(fields)
    if keys == {"px", "py", "pz", "energy"}:
        return MomentumNumpy4D(fields)
    if keys == {"x", "y", "z", "t"}:
        return VectorNumpy4D(fields)
    if keys == {"px", "py", "pz"}:
        return MomentumNumpy3D(fields)
    raise TypeError(f"vector stand-in: unsupported coordinates {sorted(keys)}")

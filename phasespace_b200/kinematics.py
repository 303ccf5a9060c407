"""Lorentz-vector helpers on ``(N, 4)`` ``[px, py, pz, E]`` tensors (reference: kinematics.py).

Inside the generator these are ``__device__`` functions of the fused kernel (csrc/ps_kernel.cuh);
this module exposes the same operations on torch CUDA tensors for API parity.  ``lorentz_boost`` and
``mass`` run the library's own kernels (``ps_lorentz_boost`` / ``ps_mass``); the component accessors
are views.
"""

from __future__ import annotations

import ctypes

import torch

from . import _lib


def _cuda64(x):
    if not torch.cuda.is_available():
        raise RuntimeError("phasespace_b200.kinematics needs a CUDA device: there is no CPU implementation.")
    return torch.as_tensor(x, dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def scalar_product(vec1, vec2):
    """kinematics.py:14-24"""
    return torch.sum(vec1 * vec2, dim=1)


def spatial_component(vector):
    """kinematics.py:28-37"""
    return vector[..., 0:3]


def time_component(vector):
    """kinematics.py:41-50"""
    return vector[..., 3:4]


def x_component(vector):
    """kinematics.py:54-63"""
    return vector[..., 0:1]


def y_component(vector):
    """kinematics.py:67-76"""
    return vector[..., 1:2]


def z_component(vector):
    """kinematics.py:80-89"""
    return vector[..., 2:3]


def mass(vector):
    """Mass of Lorentz 4-momenta, ``(N,4) -> (N,1)`` (kinematics.py:93-104)."""
    v = _cuda64(vector).reshape(-1, 4).contiguous()
    out = torch.empty((v.shape[0], 1), dtype=torch.float64, device=v.device)
    _lib.check(_lib.lib.ps_mass(ctypes.c_void_p(v.data_ptr()), v.shape[0], ctypes.c_void_p(out.data_ptr()), _stream()))
    return out


def lorentz_vector(space, time):
    """kinematics.py:108-118"""
    return torch.cat([space, time], dim=-1)


def lorentz_boost(vector, boostvector):
    """Boost ``(N,4)`` vectors by ``(N or 1, 3 or 4)`` boost vectors (kinematics.py:122-149).

    A zero boost vector is the identity *per event* (the reference tests it over the whole batch and
    yields NaN for zero-boost events mixed with non-zero ones; see DESIGN.md "quirks").
    """
    v = _cuda64(vector).reshape(-1, 4).contiguous()
    b = _cuda64(boostvector)
    b = b.reshape(1, -1) if b.ndim == 1 else b
    b = b.contiguous()
    if b.shape[0] not in (1, v.shape[0]):
        raise ValueError("boost vector must have 1 or N rows")
    out = torch.empty_like(v)
    _lib.check(_lib.lib.ps_lorentz_boost(
        ctypes.c_void_p(v.data_ptr()), ctypes.c_void_p(b.data_ptr()), b.shape[0], b.shape[1], v.shape[0],
        ctypes.c_void_p(out.data_ptr()), _stream()))
    return out


def beta(vector):
    """kinematics.py:153-162"""
    return mass(vector) / time_component(_cuda64(vector).reshape(-1, 4))


def boost_components(vector):
    """kinematics.py:166-175"""
    return spatial_component(vector) / time_component(vector)


def metric_tensor():
    """kinematics.py:179-185"""
    return torch.tensor([-1.0, -1.0, -1.0, 1.0], dtype=torch.float64)

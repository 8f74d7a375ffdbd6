"""Transform helpers backed by the C ABI's yk_xf_* entry points (math/transforms.rs, math/transform.rs).

Pure host code (no GPU needed). The functions return `desc.Transform` pairs (m, m_inv).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .desc import Transform


def _f(v, n):
    return (C.c_float * n)(*[float(x) for x in v])


def _out(t):
    return capi.from_c_transform(t)


def identity() -> Transform:
    t = capi.Transform()
    capi.lib().yk_xf_identity(C.byref(t))
    return _out(t)


def translation(delta) -> Transform:
    t = capi.Transform()
    capi.lib().yk_xf_translation(_f(delta, 3), C.byref(t))
    return _out(t)


def scale(x, y, z) -> Transform:
    t = capi.Transform()
    capi.lib().yk_xf_scale(float(x), float(y), float(z), C.byref(t))
    return _out(t)


def rotation(theta, axis) -> Transform:
    t = capi.Transform()
    capi.lib().yk_xf_rotation(float(theta), _f(axis, 3), C.byref(t))
    return _out(t)


def new(m16) -> Transform:
    """Transform::new: the inverse comes from the reference's Gauss-Jordan routine."""
    t = capi.Transform()
    capi.check(capi.lib().yk_xf_new(_f(np.asarray(m16, np.float32).reshape(-1), 16), C.byref(t)))
    return _out(t)


def look_at(pos, target, up) -> Transform:
    t = capi.Transform()
    capi.check(capi.lib().yk_xf_look_at(_f(pos, 3), _f(target, 3), _f(up, 3), C.byref(t)))
    return _out(t)


def mul(a: Transform, b: Transform) -> Transform:
    t = capi.Transform()
    ca, cb = capi.to_c_transform(a), capi.to_c_transform(b)
    capi.lib().yk_xf_mul(C.byref(ca), C.byref(cb), C.byref(t))
    return _out(t)


def inverted(a: Transform) -> Transform:
    t = capi.Transform()
    ca = capi.to_c_transform(a)
    capi.lib().yk_xf_inverted(C.byref(ca), C.byref(t))
    return _out(t)


def _apply(fn, a: Transform, v):
    ca = capi.to_c_transform(a)
    out = (C.c_float * 3)()
    fn(C.byref(ca), _f(v, 3), out)
    return np.array(out, dtype=np.float32)


def point(a: Transform, p):
    return _apply(capi.lib().yk_xf_point, a, p)


def vec(a: Transform, v):
    return _apply(capi.lib().yk_xf_vec, a, v)


def normal(a: Transform, n):
    return _apply(capi.lib().yk_xf_normal, a, n)

"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes binding of oracle/liboracle.so (oracle/yk_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The descriptor structs are layout-identical to the product's, so the ctypes classes of yuki_b200.capi are
reused for marshalling (the product never imports anything from here).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from yuki_b200 import capi
from yuki_b200 import desc as D

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")


class OStats(C.Structure):
    _fields_ = [("ray_count", C.c_uint64), ("shadow_rays", C.c_uint64), ("samples", C.c_uint64),
                ("closest_nodes", C.c_uint64), ("closest_tris", C.c_uint64), ("any_nodes", C.c_uint64), ("any_tris", C.c_uint64),
                ("primary_hit_hash", C.c_uint64), ("seconds", C.c_double), ("threads", C.c_uint32), ("_pad", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "_pad"}


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, u32, fp = C.c_void_p, C.c_uint32, C.POINTER(C.c_float)
    T = C.POINTER(capi.Transform)
    L.yko_scene_create.argtypes = [C.POINTER(capi.HostSceneDesc)]
    L.yko_scene_create.restype = vp
    L.yko_scene_destroy.argtypes = [vp]
    L.yko_scene_destroy.restype = None
    L.yko_scene_node_count.argtypes = [vp]
    L.yko_scene_node_count.restype = u32
    L.yko_scene_shape_count.argtypes = [vp]
    L.yko_scene_shape_count.restype = u32
    L.yko_scene_copy_nodes.argtypes = [vp, vp]
    L.yko_scene_copy_nodes.restype = None
    L.yko_scene_copy_order.argtypes = [vp, vp]
    L.yko_scene_copy_order.restype = None
    L.yko_render.argtypes = [vp, C.POINTER(capi.CameraParams), C.POINTER(capi.FilmSettings), C.POINTER(capi.Sampler),
                             C.POINTER(capi.Integrator), vp, u32, u32, vp, vp, u32, C.POINTER(OStats)]
    L.yko_debug_ray_path.argtypes = [vp, C.POINTER(capi.CameraParams), C.POINTER(capi.FilmSettings), C.POINTER(capi.Sampler),
                                     C.POINTER(capi.Integrator), u32, u32, vp, u32, fp, C.POINTER(C.c_uint64)]
    L.yko_film_tiles.argtypes = [u32, u32, u32, vp, u32]
    L.yko_film_tiles.restype = u32
    L.yko_camera_make.argtypes = [C.POINTER(capi.CameraParams), u32, u32, fp, fp]
    L.yko_camera_rays.argtypes = [C.POINTER(capi.CameraParams), u32, u32, fp, u32, fp, fp]
    L.yko_camera_rays.restype = None
    L.yko_xf_identity.argtypes = [T]
    L.yko_xf_translation.argtypes = [fp, T]
    L.yko_xf_scale.argtypes = [C.c_float, C.c_float, C.c_float, T]
    L.yko_xf_rotation.argtypes = [C.c_float, fp, T]
    L.yko_xf_new.argtypes = [fp, T]
    L.yko_xf_look_at.argtypes = [fp, fp, fp, T]
    L.yko_xf_mul.argtypes = [T, T, T]
    L.yko_xf_inverted.argtypes = [T, T]
    for n in ("yko_xf_point", "yko_xf_vec", "yko_xf_normal"):
        getattr(L, n).argtypes = [T, fp, fp]
        getattr(L, n).restype = None
    for n in ("yko_xf_identity", "yko_xf_translation", "yko_xf_scale", "yko_xf_rotation", "yko_xf_mul", "yko_xf_inverted"):
        getattr(L, n).restype = None
    L.yko_cross.argtypes = [fp, fp, fp]
    L.yko_cross.restype = None
    L.yko_math_kat.argtypes = [u32, fp, fp]
    L.yko_math_kat.restype = C.c_int
    L.yko_lobe_eval.argtypes = [u32, fp, u32, fp, u32, fp]
    L.yko_lobe_eval.restype = C.c_int
    L.yko_light_sample.argtypes = [vp, u32, fp, u32, fp]
    L.yko_light_sample.restype = C.c_int
    L.yko_siphash13.argtypes = [C.c_char_p, C.c_uint64]
    L.yko_siphash13.restype = C.c_uint64
    L.yko_pcg32_sequence.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, u32, vp]
    L.yko_pcg32_sequence.restype = None
    L.yko_permutation_element.argtypes = [u32, u32, u32]
    L.yko_permutation_element.restype = u32
    L.yko_sampler_draws.argtypes = [C.POINTER(capi.Sampler), u32, u32, u32, u32, C.c_char_p, u32, fp]
    L.yko_sampler_draws.restype = None
    L.yko_trace.argtypes = [vp, fp, fp, fp, u32, C.c_int, fp, vp, vp]
    L.yko_trace.restype = None
    L.yko_occluded.argtypes = [vp, fp, fp, fp, u32, C.c_int, vp]
    L.yko_occluded.restype = None
    _lib = L
    return L


def _f(v, n):
    return (C.c_float * n)(*[float(x) for x in v])


class _Xf:
    """Transform backend with the same surface as yuki_b200.transforms, evaluated by the oracle."""

    @staticmethod
    def _out(t):
        return capi.from_c_transform(t)

    def identity(self):
        t = capi.Transform(); lib().yko_xf_identity(C.byref(t)); return self._out(t)

    def translation(self, d):
        t = capi.Transform(); lib().yko_xf_translation(_f(d, 3), C.byref(t)); return self._out(t)

    def scale(self, x, y, z):
        t = capi.Transform(); lib().yko_xf_scale(float(x), float(y), float(z), C.byref(t)); return self._out(t)

    def rotation(self, theta, axis):
        t = capi.Transform(); lib().yko_xf_rotation(float(theta), _f(axis, 3), C.byref(t)); return self._out(t)

    def new(self, m16):
        t = capi.Transform()
        if lib().yko_xf_new(_f(np.asarray(m16, np.float32).reshape(-1), 16), C.byref(t)) != 0:
            raise ValueError("singular matrix")
        return self._out(t)

    def look_at(self, pos, target, up):
        t = capi.Transform()
        if lib().yko_xf_look_at(_f(pos, 3), _f(target, 3), _f(up, 3), C.byref(t)) != 0:
            raise ValueError("singular matrix")
        return self._out(t)

    def mul(self, a, b):
        t = capi.Transform(); ca, cb = capi.to_c_transform(a), capi.to_c_transform(b)
        lib().yko_xf_mul(C.byref(ca), C.byref(cb), C.byref(t)); return self._out(t)

    def inverted(self, a):
        t = capi.Transform(); ca = capi.to_c_transform(a)
        lib().yko_xf_inverted(C.byref(ca), C.byref(t)); return self._out(t)

    def _apply(self, fn, a, v):
        ca = capi.to_c_transform(a); out = (C.c_float * 3)()
        fn(C.byref(ca), _f(v, 3), out); return np.array(out, dtype=np.float32)

    def point(self, a, p):
        return self._apply(lib().yko_xf_point, a, p)

    def vec(self, a, v):
        return self._apply(lib().yko_xf_vec, a, v)

    def normal(self, a, n):
        return self._apply(lib().yko_xf_normal, a, n)


transforms = _Xf()


def math_kat(op, values, n_out):
    """yko_math.h helper `op` (see yko_math_kat) on a flat list of floats; returns n_out floats."""
    vals = np.zeros(max(len(values), 16), np.float32)
    vals[:len(values)] = values
    out = np.zeros(8, np.float32)
    if lib().yko_math_kat(op, _f(vals, len(vals)), out.ctypes.data_as(C.POINTER(C.c_float))) != 0:
        raise ValueError("yko_math_kat failed")
    return out[:n_out].copy()


LOBE_LAMBERT, LOBE_OREN_NAYAR, LOBE_SPEC_REFL, LOBE_SPEC_TRANS, LOBE_GGX_CONDUCTOR, LOBE_GGX_SCHLICK = range(6)


def lobe_f_pdf(kind, params, wo, wi):
    """f(wo, wi) (n, 3) and pdf(wo, wi) (n,) of one BxDF in its local frame (z = normal)."""
    wo = np.ascontiguousarray(np.broadcast_to(np.asarray(wo, np.float32), np.asarray(wi).shape), np.float32)
    x = np.ascontiguousarray(np.concatenate([wo, np.asarray(wi, np.float32)], axis=1), np.float32)
    prm = np.zeros(16, np.float32)
    prm[:len(params)] = params
    out = np.zeros((x.shape[0], 4), np.float32)
    assert lib().yko_lobe_eval(kind, capi.fptr(prm), 0, capi.fptr(x), x.shape[0], capi.fptr(out)) == 0
    return out[:, :3], out[:, 3]


def lobe_sample(kind, params, wo, u):
    """sample_f(wo, u): (wi (n, 3), f (n, 3), pdf (n,), sample type (n,)) — type 0 = no sample."""
    u = np.asarray(u, np.float32)
    wo = np.ascontiguousarray(np.broadcast_to(np.asarray(wo, np.float32), (u.shape[0], 3)), np.float32)
    x = np.ascontiguousarray(np.concatenate([wo, u, np.zeros((u.shape[0], 1), np.float32)], axis=1), np.float32)
    prm = np.zeros(16, np.float32)
    prm[:len(params)] = params
    out = np.zeros((x.shape[0], 8), np.float32)
    assert lib().yko_lobe_eval(kind, capi.fptr(prm), 1, capi.fptr(x), x.shape[0], capi.fptr(out)) == 0
    return out[:, :3], out[:, 3:6], out[:, 6], out[:, 7].astype(np.int32)


def cross(a, b):
    out = (C.c_float * 3)()
    lib().yko_cross(_f(a, 3), _f(b, 3), out)
    return np.array(out, dtype=np.float32)


def siphash13(msg: bytes) -> int:
    return lib().yko_siphash13(msg, len(msg))


def pcg32_sequence(state, stream, advance, n):
    out = np.zeros(n, dtype=np.uint32)
    lib().yko_pcg32_sequence(state, stream, advance, n, out.ctypes.data)
    return out


def permutation_element(i, l, p):
    return lib().yko_permutation_element(i, l, p)


def sampler_draws(sampler: D.SamplerType, px, py, index, pattern, start_dim=0):
    pat = bytes(pattern)
    out = np.zeros(sum(pattern), dtype=np.float32)
    s = capi.sampler(sampler)
    lib().yko_sampler_draws(C.byref(s), px, py, index, start_dim, pat, len(pat), capi.fptr(out))
    return out


def film_tiles(film: D.FilmSettings):
    L = lib()
    n = L.yko_film_tiles(int(film.res[0]), int(film.res[1]), int(film.tile_dim), None, 0)
    out = np.zeros(n, dtype=capi.TILE_DTYPE)
    L.yko_film_tiles(int(film.res[0]), int(film.res[1]), int(film.tile_dim), out.ctypes.data, n)
    return out


def camera_matrices(params: D.CameraParameters, film: D.FilmSettings):
    c2w = np.zeros(16, np.float32)
    r2c = np.zeros(16, np.float32)
    cp = capi.camera_params(params)
    if lib().yko_camera_make(C.byref(cp), int(film.res[0]), int(film.res[1]), capi.fptr(c2w), capi.fptr(r2c)) != 0:
        raise ValueError("singular camera")
    return c2w, r2c


def camera_rays(params: D.CameraParameters, film: D.FilmSettings, p_film: np.ndarray):
    pf = np.ascontiguousarray(p_film, np.float32).reshape(-1, 2)
    o = np.zeros((pf.shape[0], 3), np.float32)
    d = np.zeros((pf.shape[0], 3), np.float32)
    cp = capi.camera_params(params)
    lib().yko_camera_rays(C.byref(cp), int(film.res[0]), int(film.res[1]), capi.fptr(pf), pf.shape[0], capi.fptr(o), capi.fptr(d))
    return o, d


class OracleScene:
    def __init__(self, scene: D.SceneDesc):
        hd, keep = capi.build_host_scene_desc(scene)
        self._h = lib().yko_scene_create(C.byref(hd))
        del keep
        if not self._h:
            raise RuntimeError("oracle: BVH build failed")

    def nodes(self) -> np.ndarray:
        n = lib().yko_scene_node_count(self._h)
        out = np.zeros(n, dtype=capi.NODE_DTYPE)
        lib().yko_scene_copy_nodes(self._h, out.ctypes.data)
        return out

    def order(self) -> np.ndarray:
        n = lib().yko_scene_shape_count(self._h)
        out = np.zeros(n, dtype=np.uint32)
        lib().yko_scene_copy_order(self._h, out.ctypes.data)
        return out

    def render(self, camera_params, film, sampler, integrator, tiles=None, threads=0, want_hit_ids=False, aux_sample=0,
               film_out=None):
        res_x, res_y = int(film.res[0]), int(film.res[1])
        out = film_out if film_out is not None else np.zeros((res_y, res_x, 3), np.float32)
        ids = np.full((res_y, res_x), -1, np.int32) if want_hit_ids else None
        cp, fs, sm, ig = capi.camera_params(camera_params), capi.film_settings(film), capi.sampler(sampler), capi.integrator(integrator)
        st = OStats()
        tptr, nt = None, 0
        if tiles is not None:
            tiles = np.ascontiguousarray(tiles, dtype=capi.TILE_DTYPE)
            tptr, nt = tiles.ctypes.data, len(tiles)
        rc = lib().yko_render(self._h, C.byref(cp), C.byref(fs), C.byref(sm), C.byref(ig), tptr, nt, threads, out.ctypes.data,
                              ids.ctypes.data if ids is not None else None, aux_sample, C.byref(st))
        if rc != 0:
            raise RuntimeError("oracle render failed")
        return out, ids, st

    def debug_ray(self, camera_params, film, sampler, integrator, film_px, max_rays=4096):
        """launch_debug_ray + li_debug (app/window.rs:812-905): (rays as capi.DEBUG_RAY_DTYPE, li, ray count)."""
        cp, fs, sm, ig = capi.camera_params(camera_params), capi.film_settings(film), capi.sampler(sampler), capi.integrator(integrator)
        rays = np.zeros(max_rays, dtype=capi.DEBUG_RAY_DTYPE)
        li = np.zeros(3, np.float32)
        count = C.c_uint64(0)
        n = lib().yko_debug_ray_path(self._h, C.byref(cp), C.byref(fs), C.byref(sm), C.byref(ig), int(film_px[0]), int(film_px[1]),
                                     rays.ctypes.data, max_rays, capi.fptr(li), C.byref(count))
        if n < 0:
            raise RuntimeError("oracle debug ray failed")
        if n > max_rays:
            return self.debug_ray(camera_params, film, sampler, integrator, film_px, max_rays=n)
        return rays[:n].copy(), li, int(count.value)

    def light_sample(self, light, p, n, u):
        """`Light::sample_li` for shading points p with normals n and samples u: dict of l, li, pdf, has_vis, vis_o, vis_d."""
        u = np.asarray(u, np.float32).reshape(-1, 2)
        m = u.shape[0]
        x = np.ascontiguousarray(np.concatenate([np.broadcast_to(np.asarray(p, np.float32), (m, 3)),
                                                 np.broadcast_to(np.asarray(n, np.float32), (m, 3)), u], axis=1), np.float32)
        out = np.zeros((m, 14), np.float32)
        assert lib().yko_light_sample(self._h, int(light), capi.fptr(x), m, capi.fptr(out)) == 0
        return {"l": out[:, 0:3], "li": out[:, 3:6], "pdf": out[:, 6], "has_vis": out[:, 7] != 0, "vis_o": out[:, 8:11], "vis_d": out[:, 11:14]}

    def trace(self, o, d, t_max=None, brute_force=False):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        t = np.zeros(n, np.float32)
        ids = np.zeros(n, np.int32)
        counts = np.zeros((n, 2), np.uint32)
        tm = None if t_max is None else capi.fptr(np.ascontiguousarray(t_max, np.float32))
        lib().yko_trace(self._h, capi.fptr(o), capi.fptr(d), tm, n, 1 if brute_force else 0, capi.fptr(t), ids.ctypes.data,
                        counts.ctypes.data)
        return t, ids, counts

    def occluded(self, o, d, t_max=None, brute_force=False):
        o = np.ascontiguousarray(o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        out = np.zeros(n, np.uint8)
        tm = None if t_max is None else capi.fptr(np.ascontiguousarray(t_max, np.float32))
        lib().yko_occluded(self._h, capi.fptr(o), capi.fptr(d), tm, n, 1 if brute_force else 0, out.ctypes.data)
        return out

    def close(self):
        if self._h:
            lib().yko_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

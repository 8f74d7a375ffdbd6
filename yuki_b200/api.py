"""Host-side mirror of the reference interface for the hot path, over the C ABI.

Names follow yuki: `Scene` (scene/mod.rs:41-49), `Camera` (camera.rs:52), `film_tiles` (film.rs:409),
`Renderer` (renderer/mod.rs:131-177, here blocking: the call returns when the film is complete),
`IntegratorType` / `SamplerType` / `FilmSettings` (desc.py). The GPU does the per-pixel work; this module
only marshals descriptors. It fails loudly when the CUDA library or a GPU is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import capi
from . import desc as D


class Context:
    """One CUDA device + stream (`yk_context`)."""

    def __init__(self, device_id: int = 0):
        self._h = C.c_void_p()
        capi.check(capi.lib().yk_context_create(device_id, C.byref(self._h)))
        self.device_id = device_id

    def close(self):
        if self._h:
            capi.lib().yk_context_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostScene:
    """Mesh::new + Triangle::new + BoundingVolumeHierarchy::new + flattening, on the host (no GPU)."""

    def __init__(self, scene: D.SceneDesc):
        hd, keep = capi.build_host_scene_desc(scene)
        self._h = C.c_void_p()
        capi.check(capi.lib().yk_host_scene_build(C.byref(hd), C.byref(self._h)))
        del keep
        self.flat = capi.SceneDescFlat()
        capi.lib().yk_host_scene_flat(self._h, C.byref(self.flat))

    @property
    def n_nodes(self):
        return self.flat.n_nodes

    @property
    def n_tris(self):
        return self.flat.n_tris

    def nodes(self) -> np.ndarray:
        buf = C.string_at(self.flat.nodes, self.flat.n_nodes * 32)
        return np.frombuffer(buf, dtype=capi.NODE_DTYPE).copy()

    def order(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.flat.tri_orig_id, shape=(self.flat.n_tris,)).copy()

    def tri_vertices(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.flat.tri_vertices, shape=(self.flat.n_tris, 3, 3)).copy()

    def close(self):
        if self._h:
            capi.lib().yk_host_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """Device-resident scene (`yk_scene`): SoA node / triangle / material / light / texture buffers in HBM."""

    def __init__(self, ctx: Context, scene: D.SceneDesc, host: Optional[HostScene] = None):
        self.ctx = ctx
        self.host = host or HostScene(scene)
        self._h = C.c_void_p()
        capi.check(capi.lib().yk_scene_create(ctx._h, C.byref(self.host.flat), C.byref(self._h)))

    def close(self):
        if self._h:
            capi.lib().yk_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_camera(params: D.CameraParameters, film: D.FilmSettings) -> capi.Camera:
    """Camera::new (camera.rs:52-102)."""
    cam = capi.Camera()
    cp = capi.camera_params(params)
    capi.check(capi.lib().yk_camera_make(C.byref(cp), int(film.res[0]), int(film.res[1]), C.byref(cam)))
    return cam


def film_tiles(film: D.FilmSettings) -> np.ndarray:
    """film_tiles (film.rs:409-475): all tiles of the film in outward-spiral order, as a TILE_DTYPE array."""
    L = capi.lib()
    n = L.yk_film_tiles(int(film.res[0]), int(film.res[1]), int(film.tile_dim), None, 0)
    out = np.zeros(n, dtype=capi.TILE_DTYPE)
    L.yk_film_tiles(int(film.res[0]), int(film.res[1]), int(film.tile_dim), out.ctypes.data, n)
    return out


def load_ply(path: str):
    """scene/ply.rs `load` up to the index buffer: returns (points (N,3), indices (T*3,), normals (N,3)|None, uvs (N,2)|None)."""
    h = C.c_void_p()
    capi.check(capi.lib().yk_ply_load(str(path).encode(), C.byref(h)))
    try:
        v = capi.PlyData()
        capi.lib().yk_ply_view(h, C.byref(v))
        pts = np.ctypeslib.as_array(v.points, shape=(v.n_points, 3)).copy() if v.n_points else np.zeros((0, 3), np.float32)
        idx = np.ctypeslib.as_array(v.indices, shape=(v.n_indices,)).copy() if v.n_indices else np.zeros((0,), np.uint32)
        nrm = np.ctypeslib.as_array(v.normals, shape=(v.n_points, 3)).copy() if v.normals and v.n_points else None
        uvs = np.ctypeslib.as_array(v.uvs, shape=(v.n_points, 2)).copy() if v.uvs and v.n_points else None
    finally:
        capi.lib().yk_ply_destroy(h)
    return pts, idx, nrm, uvs


def bvh_build(tri_vertices: np.ndarray, max_shapes_in_node=1, split_method=D.SPLIT_SAH):
    """BoundingVolumeHierarchy::new over world-space triangles (T,3,3) -> (nodes, order)."""
    v = np.ascontiguousarray(tri_vertices, np.float32).reshape(-1, 9)
    n = v.shape[0]
    nodes = np.zeros(max(2 * n - 1, 1), dtype=capi.NODE_DTYPE)
    order = np.zeros(n, dtype=np.uint32)
    n_nodes = C.c_uint32()
    capi.check(capi.lib().yk_bvh_build(capi.fptr(v), n, max_shapes_in_node, split_method, nodes.ctypes.data, C.byref(n_nodes),
                                        order.ctypes.data_as(C.POINTER(C.c_uint32))))
    return nodes[: n_nodes.value].copy(), order


class RenderResult:
    def __init__(self, film, hit_ids, stats):
        self.film = film
        self.hit_ids = hit_ids
        self.stats = stats


class Renderer:
    """Blocking counterpart of `Renderer::launch` + `check_status` (renderer/mod.rs:61-177): renders the given
    tiles (default: the whole spiral list) with the wavefront CUDA path and returns the film."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def render(self, scene: Scene, camera_params: D.CameraParameters, film: D.FilmSettings, sampler: D.SamplerType,
               integrator: D.IntegratorType, tiles: Optional[np.ndarray] = None, want_hit_ids: bool = False,
               aux_sample: int = 0, wavefront_paths: int = 0, film_out: Optional[np.ndarray] = None,
               device_film_ptr: Optional[int] = None, progress=None, pipes: int = 0) -> RenderResult:
        cam = make_camera(camera_params, film)
        if tiles is None:
            tiles = film_tiles(film)
        tiles = np.ascontiguousarray(tiles, dtype=capi.TILE_DTYPE)
        res_x, res_y = int(film.res[0]), int(film.res[1])
        opts = capi.RenderOpts()
        opts.wavefront_paths = wavefront_paths
        opts.aux_sample = aux_sample
        opts.pipes = pipes
        hit_ids = None
        if device_film_ptr is not None:
            opts.flags |= capi.RENDER_FILM_ON_DEVICE
            film_arr, film_ptr = None, C.c_void_p(device_film_ptr)
        else:
            film_arr = film_out if film_out is not None else np.zeros((res_y, res_x, 3), dtype=np.float32)
            assert film_arr.dtype == np.float32 and film_arr.flags["C_CONTIGUOUS"] and film_arr.size == res_x * res_y * 3
            film_ptr = C.c_void_p(film_arr.ctypes.data)
            if want_hit_ids:
                hit_ids = np.full((res_y, res_x), -1, dtype=np.int32)
                opts.hit_ids = hit_ids.ctypes.data
        cb = None
        if progress is not None:
            cb = capi.PROGRESS_FN(lambda user, done, total: int(bool(progress(done, total))))
            opts.progress = cb
        fs, sm, ig = capi.film_settings(film), capi.sampler(sampler), capi.integrator(integrator)
        stats = capi.Stats()
        capi.check(capi.lib().yk_render(self.ctx._h, scene._h, C.byref(cam), C.byref(fs), C.byref(sm), C.byref(ig),
                                        C.c_void_p(tiles.ctypes.data), len(tiles), C.byref(opts), film_ptr, C.byref(stats)))
        return RenderResult(film_arr, hit_ids, stats)

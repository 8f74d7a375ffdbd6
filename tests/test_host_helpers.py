"""Host-side helpers either side of the device path: spiral tile order, camera matrices, light constructors.
Product (C ABI level 2) against the oracle, plus the reference-described properties."""
import ctypes as C

import numpy as np
import pytest

from yuki_b200 import api, capi, desc as D


@pytest.mark.parametrize("res,dim", [((640, 480), 16), ((64, 48), 16), ((70, 50), 16), ((1024, 1024), 16), ((17, 9), 8),
                                      ((3840, 2160), 16), ((16, 16), 16), ((100, 30), 64)])
def test_spiral_tiles(oracle, res, dim):
    film = D.FilmSettings(res, dim)
    p, o = api.film_tiles(film), oracle.film_tiles(film)
    assert p.tobytes() == o.tobytes()
    cols, rows = -(-res[0] // dim), -(-res[1] // dim)
    assert len(p) == cols * rows
    assert sorted(p["index"].tolist()) == list(range(cols * rows))         # a permutation of all tiles
    area = sum(int(t["x1"] - t["x0"]) * int(t["y1"] - t["y0"]) for t in p)
    assert area == res[0] * res[1]                                           # tiles are clipped to the film, disjoint
    cx, cy = cols // 2 - (1 - cols % 2), rows // 2 - (1 - rows % 2)          # film.rs:340-341: first tile is the centre
    assert (p[0]["x0"] // dim, p[0]["y0"] // dim) == (cx, cy)


def test_bounds2_iteration_order_is_row_major():
    """tests/src/bounds.rs:400-409 pins the pixel order inside a tile: rows outer, x inner — the order pixel jobs are
    generated in yk_render; checked here through the tile geometry helper."""
    film = D.FilmSettings((32, 32), 16)
    t = api.film_tiles(film)[0]
    px = [(x, y) for y in range(t["y0"], t["y1"]) for x in range(t["x0"], t["x1"])]
    assert px[0] == (t["x0"], t["y0"]) and px[1] == (t["x0"] + 1, t["y0"]) and px[16] == (t["x0"], t["y0"] + 1)


@pytest.mark.parametrize("axis", [D.FOV_X, D.FOV_Y])
@pytest.mark.parametrize("res", [(640, 480), (1024, 1024), (1920, 1080)])
def test_camera_matrices_equal_oracle(oracle, axis, res):
    params = D.CameraParameters((0.278, 0.273, 0.8), (0.278, 0.273, -0.26), (0.0, 1.0, 0.0), axis, 40.0)
    film = D.FilmSettings(res, 16)
    cam = api.make_camera(params, film)
    c2w, r2c = oracle.camera_matrices(params, film)
    assert np.array_equal(np.array(cam.camera_to_world, np.float32).view(np.uint32), c2w.view(np.uint32))
    assert np.array_equal(np.array(cam.raster_to_camera, np.float32).view(np.uint32), r2c.view(np.uint32))


def test_camera_rays_geometry(oracle):
    params = D.CameraParameters((1.0, 2.0, 3.0), (1.0, 2.0, -5.0), (0.0, 1.0, 0.0), D.FOV_X, 60.0)
    film = D.FilmSettings((200, 100), 16)
    o, d = oracle.camera_rays(params, film, np.array([[100.0, 50.0], [0.0, 50.0], [200.0, 50.0], [100.0, 0.0]], np.float32))
    assert np.allclose(o, [1.0, 2.0, 3.0])
    assert np.allclose(d[0], [0.0, 0.0, -1.0], atol=1e-6)                     # film centre looks at the target
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-6)
    half = np.degrees(np.arccos(np.clip(d[1] @ d[0], -1, 1)))
    assert abs(half - 30.0) < 1e-3 and d[2][0] < 0 < d[1][0]   # fov X 60: +-30 deg; left-handed: looking down -z, raster x grows toward -x
    assert d[3][1] > 0                                                         # raster y grows downwards


def test_singular_camera_is_an_error():
    params = D.CameraParameters((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))            # default CameraParameters: target == position
    with pytest.raises(capi.YukiGpuError):
        api.make_camera(params, D.FilmSettings())


def test_light_constructors(xf):
    L = capi.lib()
    # spot: identity points down -Z; cos of the cone angles (spot_light.rs:23-36)
    ld = capi.LightDesc()
    ld.kind = D.LIGHT_SPOT
    ld.light_to_world = capi.to_c_transform(xf.translation((1.0, 2.0, 3.0)))
    ld.intensity = capi.f3((1, 2, 3))
    ld.total_width_deg, ld.falloff_start_deg = 30.0, 20.0
    out = capi.LightDev()
    capi.check(L.yk_light_make(C.byref(ld), C.byref(out)))
    assert list(out.p) == [1.0, 2.0, 3.0] and list(out.i) == [1.0, 2.0, 3.0]
    assert abs(out.cos_total_width - np.cos(np.radians(30.0))) < 1e-6 and abs(out.cos_falloff_start - np.cos(np.radians(20.0))) < 1e-6
    assert np.allclose(np.array(out.world_to_light).reshape(4, 4)[:3, 3], [-1.0, -2.0, -3.0])
    # rect: samples in [0,1)^2 map onto a size.x by size.y rectangle centred on the light (rectangular_light.rs:34-43)
    ld.kind = D.LIGHT_RECT
    ld.size = (C.c_float * 2)(0.5, 0.25)
    capi.check(L.yk_light_make(C.byref(ld), C.byref(out)))
    m = np.array(out.sample_to_world, np.float32).reshape(4, 4)
    corner0 = m @ np.array([0, 0, 0, 1], np.float32)
    corner1 = m @ np.array([1, 0, 1, 1], np.float32)
    assert np.allclose(corner0[:3], [0.75, 2.0, 2.875]) and np.allclose(corner1[:3], [1.25, 2.0, 3.125])
    assert out.area == np.float32(0.5) * np.float32(0.25)
    ld.kind = 9
    assert L.yk_light_make(C.byref(ld), C.byref(out)) == -1


def test_division_by_invariant_matches_integer_division():
    """csrc/yk_fastdiv.h replaces the `/` and `%` of stratified.rs:127-128,177 and the batch index arithmetic on the
    device; it must agree with n / d for every 32-bit numerator (edge values + random draws per divisor)."""
    from yuki_b200 import capi
    rng = np.random.default_rng(7)
    edge = np.array([0, 1, 2, 3, 2**16 - 1, 2**16, 2**31 - 1, 2**31, 2**31 + 1, 2**32 - 2, 2**32 - 1], dtype=np.uint64)
    divisors = list(range(1, 70)) + [255, 256, 257, 1000, 1023, 1024, 1025, 4095, 4096, 4097, 65535, 65536, 65537, 2**20, 2**22,
                                     2**22 + 1, 3 * 2**20 + 7, 2**31 - 1, 2**31, 2**31 + 1, 2**32 - 1]
    divisors += [int(x) for x in rng.integers(1, 2**32, size=200, dtype=np.uint64)]
    for d in divisors:
        near = np.array([k * d + o for k in (1, 2, 3, 1000, (2**32 - 1) // d) for o in (-1, 0, 1)], dtype=np.int64)
        near = near[(near >= 0) & (near < 2**32)].astype(np.uint64)
        n = np.concatenate([edge, near, rng.integers(0, 2**32, size=20000, dtype=np.uint64)]).astype(np.uint32)
        n = np.ascontiguousarray(n)
        assert capi.lib().yk_selftest_fastdiv(d, n.ctypes.data, len(n)) == 0, d


def test_host_scene_save_load_round_trip(tmp_path, xf):
    """HostScene.save / load: the flattened arrays a rank built are mapped back byte for byte by the other ranks of a node."""
    import ctypes as C
    from yuki_b200 import api, scenes
    for scene in (scenes.material_room(xf)[0], scenes.cornell(xf, light="rect", tall_box="glass", sphere=True, textured_back_wall=True)[0]):
        a = api.HostScene(scene)
        a.save(str(tmp_path / "hs"))
        b = api.HostScene.load(str(tmp_path / "hs"))
        fa, fb = a.flat, b.flat
        for k in ("n_nodes", "n_tris", "n_textures", "n_materials", "n_lights", "n_spheres"):
            assert getattr(fa, k) == getattr(fb, k), k
        assert list(fa.background) == list(fb.background)

        def raw(ptr, nbytes):
            return C.string_at(C.cast(ptr, C.c_void_p), nbytes) if ptr and nbytes else b""
        for name, count, per, dtype in api.HostScene._ARRAYS:
            n = getattr(fa, count) * per * np.dtype(dtype).itemsize
            assert bool(getattr(fa, name)) == bool(getattr(fb, name)), name
            assert raw(getattr(fa, name), n) == raw(getattr(fb, name), n), name
        assert raw(fa.materials, fa.n_materials * C.sizeof(capi.MaterialDesc)) == raw(fb.materials, fb.n_materials * C.sizeof(capi.MaterialDesc))
        assert raw(fa.lights, fa.n_lights * C.sizeof(capi.LightDev)) == raw(fb.lights, fb.n_lights * C.sizeof(capi.LightDev))
        assert raw(fa.spheres, fa.n_spheres * C.sizeof(capi.SphereDev)) == raw(fb.spheres, fb.n_spheres * C.sizeof(capi.SphereDev))
        for i in range(fa.n_textures):
            ta, tb = fa.textures[i], fb.textures[i]
            assert (ta.kind, ta.width, ta.height, list(ta.value)) == (tb.kind, tb.width, tb.height, list(tb.value))
            if ta.kind == 1:
                assert raw(ta.texels, ta.width * ta.height * 12) == raw(tb.texels, tb.width * tb.height * 12)
        a.close(); b.close()

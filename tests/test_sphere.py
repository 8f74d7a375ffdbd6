"""Sphere shape (shapes/sphere.rs) as the second primitive kind in BVH leaves: host build parity with the oracle, oracle
self-checks (brute force vs tree; analytic hit), and the CUDA path against the oracle. The hit (t, ids, counters) uses
only + - * / sqrt; uv / shading frame go through atan2 / acos, which the device evaluates with glibc's algorithms
(yk_libm.h) — so counters, normals and radiance are all held to bit equality."""
import numpy as np
import pytest

from conftest import rel_rmse
from yuki_b200 import api, desc as D, scenes

SPLITS = [D.SPLIT_SAH, D.SPLIT_MIDDLE, D.SPLIT_EQUAL_COUNTS]


def sphere_scene(xf, split=D.SPLIT_MIDDLE, light="rect"):
    return scenes.cornell(xf, light=light, tall_box="glass", sphere=True, split_method=split)


def sphere_field(xf, n=24, split=D.SPLIT_SAH, seed=5):
    """Many spheres (scaled / rotated / mirrored transforms) over a ground quad, point + rect light."""
    rng = np.random.default_rng(seed)
    s = D.SceneDesc(split_method=split, background=(0.1, 0.12, 0.15))
    zero = s.add_texture(D.Texture.constant(0.0))
    mats = [s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(*rng.uniform(0.2, 0.9, 3))), zero))) for _ in range(3)]
    mats.append(s.add_material(D.Material(D.MAT_GLASS, (s.add_texture(D.Texture.constant(1.0)), s.add_texture(D.Texture.constant(1.0))), eta=1.5)))
    mats.append(s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.from_image(scenes.checker_texture(64, 8))), zero))))
    p, i = scenes._quad([(-2, 0, -2), (-2, 0, 2), (2, 0, 2), (2, 0, -2)])
    s.meshes.append(D.Mesh(xf.identity(), p, i, mats[0]))
    for k in range(n):
        t = xf.translation(tuple(float(v) for v in (rng.uniform(-1.5, 1.5), rng.uniform(0.15, 0.9), rng.uniform(-1.5, 1.5))))
        if k % 3 == 1:
            t = xf.mul(t, xf.scale(*[float(v) for v in rng.uniform(0.6, 1.4, 3)]))
        if k % 3 == 2:
            t = xf.mul(t, xf.mul(xf.rotation(float(rng.uniform(0, 3)), (0.3, 1.0, 0.2)), xf.scale(1.0, -1.0, 1.0)))  # swaps handedness
        s.spheres.append(D.Sphere(t, float(rng.uniform(0.08, 0.22)), mats[k % len(mats)]))
    s.lights.append(D.Light(D.LIGHT_POINT, xf.translation((0.5, 2.5, 1.0)), (6.0, 6.0, 6.0)))
    cam = D.CameraParameters((0.0, 1.6, 4.0), (0.0, 0.4, 0.0), fov_axis=D.FOV_X, fov_deg=45.0)
    return s, cam


@pytest.mark.parametrize("split", SPLITS)
def test_product_bvh_with_spheres_equals_oracle(oracle, xf, split):
    for scene, _ in (sphere_scene(xf, split), sphere_field(xf, 40, split)):
        host, osc = api.HostScene(scene), oracle.OracleScene(scene)
        assert host.nodes().tobytes() == osc.nodes().tobytes()
        assert np.array_equal(host.order(), osc.order())
        assert host.n_tris == scene.n_triangles() + len(scene.spheres)


def test_oracle_sphere_hits_are_analytic(oracle, xf):
    """Unit sphere at the origin behind an identity transform: t = distance to the surface along an axis ray."""
    s = D.SceneDesc()
    zero = s.add_texture(D.Texture.constant(0.0))
    m = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(0.5)), zero)))
    s.spheres.append(D.Sphere(xf.translation((0.0, 0.0, 0.0)), 1.0, m))
    osc = oracle.OracleScene(s)
    o = np.array([[0, 0, 5], [0, 0, 0], [3, 0, 0], [0, 2, 5]], np.float32)
    d = np.array([[0, 0, -1], [0, 1, 0], [-2, 0, 0], [0, 0, -1]], np.float32)
    t, ids, _ = osc.trace(o, d)
    assert t[0] == 4.0 and t[1] == 1.0 and t[2] == 1.0 and ids[3] == -1
    assert ids[:3].tolist() == [0, 0, 0]


def test_oracle_bvh_equals_brute_force_with_spheres(oracle, xf):
    scene, cam = sphere_field(xf, 60)
    osc = oracle.OracleScene(scene)
    rng = np.random.default_rng(1)
    o = np.tile(np.array([[0.0, 1.6, 4.0]], np.float32), (4000, 1)) + rng.normal(0, 0.2, (4000, 3)).astype(np.float32)
    tgt = rng.uniform([-2, 0, -2], [2, 1, 2], (4000, 3)).astype(np.float32)
    d = tgt - o
    t0, i0, _ = osc.trace(o, d)
    t1, i1, _ = osc.trace(o, d, brute_force=True)
    assert np.array_equal(i0, i1) and np.array_equal(t0.view(np.uint32), t1.view(np.uint32))
    assert (i0 >= 2).sum() > 500   # plenty of sphere hits (ids 0, 1 are the ground)


@pytest.mark.gpu
@pytest.mark.parametrize("split", SPLITS)
def test_gpu_bvh_counters_with_spheres_bit_exact(gpu_ctx, oracle, xf, split):
    scene, cam = sphere_field(xf, 80, split)
    film = D.FilmSettings((160, 96), 16)
    smp, integ = D.SamplerType.uniform(1), D.IntegratorType.bvh_intersections()
    dev = api.Scene(gpu_ctx, scene)
    r = api.Renderer(gpu_ctx).render(dev, cam, film, smp, integ, want_hit_ids=True)
    o_img, o_ids, o_st = oracle.OracleScene(scene).render(cam, film, smp, integ, want_hit_ids=True)
    assert np.array_equal(r.hit_ids, o_ids)
    assert np.array_equal(r.film.view(np.uint32), o_img.view(np.uint32))
    assert r.stats.closest_nodes == o_st.closest_nodes and r.stats.closest_tris == o_st.closest_tris
    dev.close()


@pytest.mark.gpu
def test_gpu_geometry_normals_on_spheres(gpu_ctx, oracle, xf):
    scene, cam = sphere_field(xf, 80)
    film = D.FilmSettings((160, 96), 16)
    smp, integ = D.SamplerType.stratified(2, 2), D.IntegratorType.debug(D.INTEGRATOR_GEOMETRY_NORMALS)
    dev = api.Scene(gpu_ctx, scene)
    r = api.Renderer(gpu_ctx).render(dev, cam, film, smp, integ, want_hit_ids=True)
    o_img, o_ids, _ = oracle.OracleScene(scene).render(cam, film, smp, integ, want_hit_ids=True)
    assert np.array_equal(r.hit_ids, o_ids)
    assert np.array_equal(r.film.view(np.uint32), o_img.view(np.uint32))
    dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("integ", [D.IntegratorType.path(6), D.IntegratorType.whitted(3)])
def test_gpu_radiance_with_spheres(gpu_ctx, oracle, xf, integ):
    for scene, cam in (sphere_scene(xf), sphere_field(xf, 40)):
        film = D.FilmSettings((128, 96), 16)
        smp = D.SamplerType.stratified(3, 3)
        dev = api.Scene(gpu_ctx, scene)
        r = api.Renderer(gpu_ctx).render(dev, cam, film, smp, integ, want_hit_ids=True)
        o_img, o_ids, o_st = oracle.OracleScene(scene).render(cam, film, smp, integ, want_hit_ids=True)
        assert np.array_equal(r.hit_ids, o_ids)
        assert r.stats.primary_hit_hash == o_st.primary_hit_hash
        assert r.stats.ray_count == o_st.ray_count and r.stats.shadow_rays == o_st.shadow_rays
        assert rel_rmse(r.film, o_img) <= 1e-3
        assert np.array_equal(r.film.view(np.uint32), o_img.view(np.uint32))
        dev.close()

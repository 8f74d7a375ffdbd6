"""Analytic pins of the oracle's triangle path (shapes/triangle.rs, shapes/mesh.rs, the debug integrators): hit
distances on a known plane, barycentric uv interpolation, geometric and shading normals, and the image texture's
nearest-texel / flipped-v look-up (textures/image_texture.rs:81-111) seen through a Lambertian surface under a distant light."""
import numpy as np

from yuki_b200 import desc as D, scenes, transforms as xf


def _quad_scene(tex=None, normals=None):
    """The square [-1, 1]^2 in the plane z = 0 facing +z, uv = (x + 1) / 2, (y + 1) / 2."""
    s = D.SceneDesc(background=(0.0, 0.0, 0.0))
    zero = s.add_texture(D.Texture.constant(0.0))
    kd = s.add_texture(tex if tex is not None else D.Texture.constant(0.5))
    m = s.add_material(D.Material(D.MAT_MATTE, (kd, zero)))
    p = np.array([(-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0)], np.float32)
    uv = np.array([(0, 0), (1, 0), (1, 1), (0, 1)], np.float32)
    s.meshes.append(D.Mesh(xf.identity(), p, np.array([0, 1, 2, 0, 2, 3], np.uint32), m, uvs=uv, normals=normals))
    cam = D.CameraParameters((0.3, -0.2, 3.0), (0.1, 0.0, 0.0), fov_axis=D.FOV_X, fov_deg=50.0)
    return s, cam


def _pixel_centre_hits(oracle, cam, film):
    """Where the pixel-centre camera rays meet the plane z = 0 (the unjittered 1x1 stratified sample is the centre)."""
    ys, xs = np.mgrid[0:film.res[1], 0:film.res[0]]
    pf = np.stack([xs + 0.5, ys + 0.5], axis=-1).reshape(-1, 2).astype(np.float32)
    o, d = oracle.camera_rays(cam, film, pf)
    t = -o[:, 2].astype(np.float64) / d[:, 2].astype(np.float64)
    p = o.astype(np.float64) + t[:, None] * d.astype(np.float64)
    inside = (np.abs(p[:, 0]) < 1) & (np.abs(p[:, 1]) < 1)
    shape = (film.res[1], film.res[0])
    return o, d, t.reshape(shape), p.reshape(shape + (3,)), inside.reshape(shape)


def test_hit_distance_ids_and_misses_on_a_known_plane(oracle):
    scene, cam = _quad_scene()
    osc = oracle.OracleScene(scene)
    film = D.FilmSettings((64, 48), 16)
    o, d, t, p, inside = _pixel_centre_hits(oracle, cam, film)
    got_t, ids, _ = osc.trace(o, d)
    margin = (np.abs(np.abs(p[..., 0]) - 1) > 1e-4) & (np.abs(np.abs(p[..., 1]) - 1) > 1e-4)   # away from the outline
    hit = (ids >= 0).reshape(inside.shape)
    assert np.array_equal(hit[margin], inside[margin]) and inside.sum() > 500 and (~inside).sum() > 200
    assert np.allclose(got_t.reshape(inside.shape)[inside & margin], t[inside & margin], rtol=2e-6)
    assert np.isinf(got_t.reshape(inside.shape)[~inside & margin]).all()
    # the diagonal splits the quad: triangle 0 below it (y < x), triangle 1 above
    tri = ids.reshape(inside.shape)
    off_diag = np.abs(p[..., 1] - p[..., 0]) > 1e-3
    assert np.array_equal(tri[inside & margin & off_diag], (p[..., 1] > p[..., 0])[inside & margin & off_diag].astype(np.int32))


def test_uv_and_normal_debug_integrators_are_the_analytic_values(oracle):
    scene, cam = _quad_scene()
    osc = oracle.OracleScene(scene)
    film = D.FilmSettings((64, 48), 16)
    smp = D.SamplerType.stratified(1, 1, jitter=False)
    _, _, _, p, inside = _pixel_centre_hits(oracle, cam, film)
    margin = inside & (np.abs(np.abs(p[..., 0]) - 1) > 1e-3) & (np.abs(np.abs(p[..., 1]) - 1) > 1e-3)
    uv_img, ids, _ = osc.render(cam, film, smp, D.IntegratorType.debug(D.INTEGRATOR_SHADING_UVS), want_hit_ids=True)
    assert np.allclose(uv_img[margin][:, 0], (p[margin][:, 0] + 1) / 2, atol=2e-6)      # barycentric interpolation, triangle.rs:141-185
    assert np.allclose(uv_img[margin][:, 1], (p[margin][:, 1] + 1) / 2, atol=2e-6)
    assert (uv_img[~inside] == 0).all() and (ids[~inside & (np.abs(np.abs(p[..., 0]) - 1) > 1e-3) & (np.abs(np.abs(p[..., 1]) - 1) > 1e-3)] == -1).all()
    gn, _, _ = osc.render(cam, film, smp, D.IntegratorType.debug(D.INTEGRATOR_GEOMETRY_NORMALS))
    assert np.allclose(gn[margin], [0.5, 0.5, 1.0], atol=1e-6)                           # n / 2 + 0.5 with n = +z
    # shading normals: per-vertex normals tilted about y, interpolated and renormalised
    tilt = np.array([(-0.3, 0, 1), (0.3, 0, 1), (0.3, 0, 1), (-0.3, 0, 1)], np.float64)
    tilt /= np.linalg.norm(tilt, axis=1, keepdims=True)
    s2, _ = _quad_scene(normals=tilt.astype(np.float32))
    sn, _, _ = oracle.OracleScene(s2).render(cam, film, smp, D.IntegratorType.debug(D.INTEGRATOR_SHADING_NORMALS))
    x = p[margin][:, 0]
    nx = tilt[1, 0] * x                      # linear in x between -n and +n, z constant, then normalised
    want = np.stack([nx, np.zeros_like(x), np.full_like(x, tilt[0, 2])], axis=1)
    want /= np.linalg.norm(want, axis=1, keepdims=True)
    assert np.allclose(sn[margin], want / 2 + 0.5, atol=2e-6)


def test_image_texture_is_nearest_texel_with_flipped_v(oracle):
    """A Lambertian surface under a head-on distant light shows kd(uv) * L / pi: read the texel choice off the render."""
    rng = np.random.default_rng(0)
    w, h = 8, 5
    img = rng.uniform(0.1, 0.9, (h, w, 3)).astype(np.float32)
    scene, cam = _quad_scene(tex=D.Texture.from_image(img))
    scene.lights.append(D.Light(D.LIGHT_DISTANT, xf.identity(), (np.pi, np.pi, np.pi), direction=(0.0, 0.0, 1.0)))
    film = D.FilmSettings((96, 72), 16)
    out, _, _ = oracle.OracleScene(scene).render(cam, film, D.SamplerType.stratified(1, 1, jitter=False), D.IntegratorType.path(1))
    _, _, _, p, inside = _pixel_centre_hits(oracle, cam, film)
    u, v = (p[..., 0] + 1) / 2, (p[..., 1] + 1) / 2
    fx, fy = u * w - 0.5, (1 - v) * h - 0.5                    # image_texture.rs:85-110: st.y = 1 - st.y, ix = (st.x * W - 0.5) as usize
    ix, iy = np.clip(np.trunc(fx), 0, w - 1).astype(int), np.clip(np.trunc(fy), 0, h - 1).astype(int)
    # keep pixels whose texel choice does not hinge on rounding
    safe = inside & (np.abs(np.abs(p[..., 0]) - 1) > 1e-2) & (np.abs(np.abs(p[..., 1]) - 1) > 1e-2)
    safe &= (np.abs(fx - np.round(fx)) > 1e-3) & (np.abs(fy - np.round(fy)) > 1e-3)
    assert safe.sum() > 800
    assert np.allclose(out[safe], img[iy[safe], ix[safe]], rtol=1e-5)


def test_direct_lighting_and_shadow_of_a_point_light_are_analytic(oracle):
    """Whitted depth 1 (= Path depth 1) on a Lambertian plane: L = kd / pi * I / d^2 * cos(theta), and zero inside the
    shadow an occluding square casts from the light (its outline projected onto the plane)."""
    scene, cam = _quad_scene()
    light = np.array([0.4, 0.3, 2.0])
    scene.lights.append(D.Light(D.LIGHT_POINT, xf.translation(tuple(light)), (3.0, 2.0, 1.0)))
    # occluder: the square |x - 0.2|, |y + 0.1| < 0.25 at height z = 1, facing away from the camera's view of the plane below
    oc = np.array([(-0.05, -0.35, 1), (0.45, -0.35, 1), (0.45, 0.15, 1), (-0.05, 0.15, 1)], np.float32)
    zero = scene.add_texture(D.Texture.constant(0.0))
    black = scene.add_material(D.Material(D.MAT_MATTE, (zero, zero)))
    scene.meshes.append(D.Mesh(xf.identity(), oc, np.array([0, 1, 2, 0, 2, 3], np.uint32), black))
    film = D.FilmSettings((96, 72), 16)
    smp = D.SamplerType.stratified(1, 1, jitter=False)
    osc = oracle.OracleScene(scene)
    img_w, ids, _ = osc.render(cam, film, smp, D.IntegratorType.whitted(1), want_hit_ids=True)
    img_p, _, _ = osc.render(cam, film, smp, D.IntegratorType.path(1))
    assert np.array_equal(img_w.view(np.uint32), img_p.view(np.uint32))
    _, _, _, p, inside = _pixel_centre_hits(oracle, cam, film)
    on_plane = inside & (ids >= 0) & (ids < 2)               # pixels that see the plane (not the occluder in front of it)
    # shadow: the segment p -> light crosses z = 1 inside the occluder
    s = (1.0 - p[..., 2]) / (light[2] - p[..., 2])
    q = p + s[..., None] * (light - p)
    dx, dy = np.abs(q[..., 0] - 0.2), np.abs(q[..., 1] + 0.1)
    shadowed = (dx < 0.25 - 2e-3) & (dy < 0.25 - 2e-3)
    lit = (dx > 0.25 + 2e-3) | (dy > 0.25 + 2e-3)
    edge = (np.abs(np.abs(p[..., 0]) - 1) > 1e-2) & (np.abs(np.abs(p[..., 1]) - 1) > 1e-2)
    assert (on_plane & shadowed).sum() > 30 and (on_plane & lit & edge).sum() > 500
    assert (img_w[on_plane & shadowed & edge] == 0).all()
    to = light - p
    d2 = (to ** 2).sum(axis=-1)
    cos = to[..., 2] / np.sqrt(d2)
    want = (0.5 / np.pi) * np.array([3.0, 2.0, 1.0]) * (cos / d2)[..., None]
    sel = on_plane & lit & edge
    assert np.allclose(img_w[sel], want[sel], rtol=2e-5)


def test_deeper_paths_add_nothing_over_a_single_plane_and_the_background_is_weighted_by_the_throughput(oracle):
    """Every bounce ray off a lone plane escapes. With a black sky Path depth 8 is Path depth 1 bit for bit; with a sky of
    radiance B each path gains beta * B = kd * B exactly once (cosine sampling of a Lambertian: f cos / pdf = kd), and camera
    rays that miss return B (path.rs:155-160)."""
    scene, cam = _quad_scene()
    scene.lights.append(D.Light(D.LIGHT_POINT, xf.translation((0.4, 0.3, 2.0)), (3.0, 2.0, 1.0)))
    film = D.FilmSettings((64, 48), 16)
    smp = D.SamplerType.stratified(2, 2)
    osc = oracle.OracleScene(scene)
    p1, ids, st1 = osc.render(cam, film, smp, D.IntegratorType.path(1), want_hit_ids=True)
    p8, _, st8 = osc.render(cam, film, smp, D.IntegratorType.path(8))
    assert np.array_equal(p1.view(np.uint32), p8.view(np.uint32))
    n_samples = 64 * 48 * 4
    assert st1.ray_count == n_samples and n_samples < st8.ray_count < 2 * n_samples     # camera rays + one escaping bounce ray per hit
    assert st8.ray_count - n_samples > 0.2 * n_samples and st8.shadow_rays == st1.shadow_rays
    sky = (0.2, 0.4, 0.8)
    scene.background = sky
    osc2 = oracle.OracleScene(scene)
    b1, ids1, _ = osc2.render(cam, film, smp, D.IntegratorType.path(1), want_hit_ids=True)
    b8, _, _ = osc2.render(cam, film, smp, D.IntegratorType.path(8))
    _, _, _, p, inside = _pixel_centre_hits(oracle, cam, film)
    well_inside = (np.abs(p[..., 0]) < 0.9) & (np.abs(p[..., 1]) < 0.9)
    well_outside = (np.abs(p[..., 0]) > 1.1) | (np.abs(p[..., 1]) > 1.1)
    assert np.allclose(b8[well_outside], sky, rtol=1e-6) and np.allclose(b1[well_outside], sky, rtol=1e-6)
    assert np.allclose(b8[well_inside] - b1[well_inside], 0.5 * np.float32(sky), rtol=2e-4, atol=1e-6)

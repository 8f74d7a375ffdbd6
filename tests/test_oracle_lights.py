"""Analytic pins of the oracle's lights (lights/*.rs restated in oracle/yko_scene.h): inverse-square fall-off, the spot
light's cone and its delta^4 blend, the rectangular light's area density (its reciprocal averages to the subtended solid
angle), one-sidedness, the distant light, and the shadow-ray convention of interaction.rs:44-59 (offset origin,
unnormalised direction ending at the light). The reference has no tests for any of this."""
import numpy as np

from yuki_b200 import desc as D, scenes, transforms as xf


def _scene(lights):
    s = D.SceneDesc()
    zero = s.add_texture(D.Texture.constant(0.0))
    m = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(0.5)), zero)))
    p, i = scenes._quad([(-1, 0, -1), (-1, 0, 1), (1, 0, 1), (1, 0, -1)])
    s.meshes.append(D.Mesh(xf.identity(), p, i, m))
    s.lights.extend(lights)
    return s


def test_point_light_is_inverse_square_and_aims_its_shadow_ray_at_the_light(oracle):
    pos = (0.5, 2.0, -0.25)
    osc = oracle.OracleScene(_scene([D.Light(D.LIGHT_POINT, xf.translation(pos), (3.0, 2.0, 1.0))]))
    rng = np.random.default_rng(0)
    for _ in range(20):
        p = rng.uniform(-1, 1, 3).astype(np.float32) * np.float32([1, 0.2, 1])
        n = np.float32([0, 1, 0])
        r = osc.light_sample(0, p, n, [[0.3, 0.7]])
        to = np.float32(pos) - p
        d2 = float((to.astype(np.float64) ** 2).sum())
        assert np.allclose(r["li"][0], np.float32([3, 2, 1]) / d2, rtol=1e-6) and r["pdf"][0] == 1.0 and r["has_vis"][0]
        assert np.allclose(r["l"][0], to / np.sqrt(d2), rtol=1e-6)
        # spawn_ray_to: origin offset 1e-3 along the normal towards the light, d = target - origin (unnormalised)
        assert np.allclose(r["vis_o"][0], p + np.float32(1e-3) * n * np.sign(to @ n), atol=1e-7)
        assert np.allclose(r["vis_o"][0] + r["vis_d"][0], pos, atol=1e-6)


def test_spot_light_cone_and_blend(oracle):
    # identity points down -z (spot_light.rs:22); rotate so that it points down -y from (0, 2, 0)
    to_world = xf.mul(xf.translation((0.0, 2.0, 0.0)), xf.rotation(float(np.float32(np.pi / 2)), (1.0, 0.0, 0.0)))
    osc = oracle.OracleScene(_scene([D.Light(D.LIGHT_SPOT, to_world, (4.0, 4.0, 4.0), total_width_deg=30.0, falloff_start_deg=20.0)]))
    n = np.float32([0, 1, 0])
    for ang in np.linspace(0.0, 40.0, 81):
        x = 2.0 * np.tan(np.deg2rad(ang))
        r = osc.light_sample(0, np.float32([x, 0.0, 0.0]), n, [[0.5, 0.5]])
        d2 = x * x + 4.0
        c, ct, cs = np.cos(np.deg2rad(ang)), np.cos(np.deg2rad(30.0)), np.cos(np.deg2rad(20.0))
        want = 1.0 if c > cs else (0.0 if c < ct else ((c - ct) / (cs - ct)) ** 4)
        if abs(ang - 20.0) < 0.3 or abs(ang - 30.0) < 0.3:
            continue  # at the cone angles f32 rounding decides the branch
        assert np.allclose(r["li"][0], 4.0 * want / d2, rtol=2e-3, atol=1e-7), ang
        assert bool(r["has_vis"][0]) == (want > 0.0)            # no shadow ray outside the cone (spot_light.rs:61-72)


def test_rect_light_density_one_sidedness_and_solid_angle(oracle):
    # identity faces -y at the origin (rectangular_light.rs:18); a 1 x 0.5 light at height 2 above the point
    size = (1.0, 0.5)
    osc = oracle.OracleScene(_scene([D.Light(D.LIGHT_RECT, xf.translation((0.0, 2.0, 0.0)), (5.0, 5.0, 5.0), size=size)]))
    rng = np.random.default_rng(1)
    u = rng.uniform(0, 1, (200_000, 2)).astype(np.float32)
    n = np.float32([0, 1, 0])
    for p in (np.float32([0, 0, 0]), np.float32([0.8, 0.3, -0.6])):
        r = osc.light_sample(0, p, n, u)
        pts = r["vis_o"] + r["vis_d"]                                           # the shadow ray ends on the sampled point
        assert np.allclose(pts[:, 1], 2.0, atol=1e-6)
        assert pts[:, 0].min() >= -0.5 - 1e-6 and pts[:, 0].max() <= 0.5 + 1e-6 and np.abs(pts[:, 2]).max() <= 0.25 + 1e-6
        assert abs(pts[:, 0].mean()) < 5e-3 and abs(pts[:, 2].mean()) < 5e-3    # uniform over the rectangle
        assert (r["li"] == 5.0).all() and r["has_vis"].all()
        # pdf = d^2 / (|cos| A): its reciprocal averages to the solid angle the light subtends from p
        d = pts.astype(np.float64) - p
        d2 = (d ** 2).sum(axis=1)
        cos_l = np.abs(d[:, 1]) / np.sqrt(d2)
        assert np.allclose(r["pdf"], d2 / (cos_l * size[0] * size[1]), rtol=1e-4)
        omega_mc = float((1.0 / r["pdf"].astype(np.float64)).mean())
        # the same solid angle by quadrature over the rectangle
        g = (np.arange(400) + 0.5) / 400
        gx, gz = np.meshgrid((g - 0.5) * size[0], (g - 0.5) * size[1])
        q = np.stack([gx - p[0], np.full_like(gx, 2.0 - p[1]), gz - p[2]], axis=-1)
        q2 = (q ** 2).sum(axis=-1)
        omega = float((np.abs(q[..., 1]) / q2 ** 1.5).mean() * size[0] * size[1])
        assert abs(omega_mc - omega) < 2e-3 * omega
    above = osc.light_sample(0, np.float32([0, 3.0, 0]), np.float32([0, -1, 0]), u[:100])
    assert (above["li"] == 0.0).all()                                           # the back side emits nothing (rectangular_light.rs:57-61)


def test_distant_light(oracle):
    """`DistantLight::new` stores w as given (distant_light.rs:17-21; the pbrt loader normalises `from - to` before calling),
    `sample_li` returns it as l and aims the shadow ray at p + 10000 w."""
    for w in ((0.6, 0.8, 0.0), (0.3, 1.0, 0.2)):        # unit, and not: passed through either way
        osc = oracle.OracleScene(_scene([D.Light(D.LIGHT_DISTANT, xf.identity(), (2.0, 1.5, 1.0), direction=w)]))
        for p in ([0, 0, 0], [0.5, 0.1, -0.7]):
            r = osc.light_sample(0, np.float32(p), np.float32([0, 1, 0]), [[0.1, 0.9]])
            assert np.array_equal(r["l"][0], np.float32(w)) and np.allclose(r["li"][0], [2.0, 1.5, 1.0]) and r["pdf"][0] == 1.0
            end = r["vis_o"][0] + r["vis_d"][0]
            assert np.allclose(end, np.float32(p) + np.float32(w) * np.float32(10000.0), rtol=1e-6)   # distant_light.rs:35-40

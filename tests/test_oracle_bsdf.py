"""Analytic pins of the oracle's BxDFs (materials/bsdfs/*.rs restated in oracle/yko_bsdf.h). The reference has no tests
for them, so these check what any correct restatement must satisfy: sampled values agree with evaluated ones, densities
integrate to the probability of producing a sample, reflectance never exceeds one, the Fresnel terms hit their closed
forms, perfect glass conserves energy (the reference omits the eta^2 radiance scaling, specular.rs:83-84), and the
documented quirks behave as documented. The CUDA path reproduces the oracle bit for bit, so these hold for it too."""
import numpy as np
import pytest

N = 400_000


def _unit(v):
    v = np.asarray(v, np.float64)
    return (v / np.linalg.norm(v, axis=-1, keepdims=True)).astype(np.float32)


def _uniform_sphere(rng, n):
    z = rng.uniform(-1, 1, n)
    phi = rng.uniform(0, 2 * np.pi, n)
    r = np.sqrt(1 - z * z)
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)


LOBES = [("lambert", 0, [0.8, 0.5, 0.3]), ("oren_nayar_20deg", 1, [0.8, 0.5, 0.3, np.deg2rad(20.0)]),
         ("ggx_schlick_rough", 5, [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.4]), ("ggx_schlick_smooth", 5, [1.0, 1.0, 1.0, 0.9, 0.9, 0.9, 0.1]),
         ("ggx_copper", 4, [1.0, 1.0, 1.0, 0.27105, 0.67693, 1.31640, 3.60920, 2.62480, 2.29210, 0.25])]
WOS = [_unit([0.0, 0.0, 1.0]), _unit([0.6, 0.2, 0.5]), _unit([0.9, -0.3, 0.15])]


@pytest.mark.parametrize("name,kind,params", LOBES, ids=[l[0] for l in LOBES])
def test_sampling_is_consistent_with_evaluation(oracle, name, kind, params):
    rng = np.random.default_rng(1)
    for wo in WOS:
        u = rng.uniform(0, 1, (20000, 2)).astype(np.float32)
        wi, f, pdf, typ = oracle.lobe_sample(kind, params, wo, u)
        ok = typ != 0
        assert ok.mean() > 0.5
        f2, pdf2 = oracle.lobe_f_pdf(kind, params, wo, wi[ok])
        assert np.array_equal(f[ok].view(np.uint32), f2.view(np.uint32))          # sample_f returns f(wo, wi) ...
        if kind <= 1:
            assert np.array_equal(pdf[ok].view(np.uint32), pdf2.view(np.uint32))  # ... and pdf(wo, wi) (diffuse: the very call)
        else:                                                                      # GGX: pdf from the sampled wh, Lobe::pdf from normalize(wo + wi)
            assert np.allclose(pdf[ok], pdf2, rtol=2e-3)
        assert np.abs(np.linalg.norm(wi[ok].astype(np.float64), axis=1) - 1).max() < 1e-4
        assert (wi[ok][:, 2] * wo[2] > 0).all()                                    # reflection lobes stay in wo's hemisphere


@pytest.mark.parametrize("name,kind,params", LOBES, ids=[l[0] for l in LOBES])
def test_pdf_integrates_to_the_sampling_probability_and_energy_is_bounded(oracle, name, kind, params):
    rng = np.random.default_rng(2)
    for wo in WOS:
        # (1) the density over the sphere integrates to the probability that sample_f produces a direction
        w = _uniform_sphere(rng, N)
        _, pdf = oracle.lobe_f_pdf(kind, params, wo, w)
        integral = float(pdf.astype(np.float64).mean() * 4 * np.pi)
        u = rng.uniform(0, 1, (N, 2)).astype(np.float32)
        wi, f, spdf, typ = oracle.lobe_sample(kind, params, wo, u)
        ok = (typ != 0) & (spdf > 0)
        p_sample = float(ok.mean())
        if kind <= 1:
            assert abs(integral - 1.0) < 0.01 and p_sample == 1.0
        else:
            # full-distribution GGX sampling rejects half-vectors facing away and reflections below the horizon; the density
            # still describes the accepted ones (the smooth lobe's narrow peak needs a looser Monte-Carlo tolerance)
            assert integral <= 1.0 + 0.03 and abs(integral - p_sample) < (0.06 if params[-1] < 0.2 else 0.02)
        # (2) directional-hemispherical reflectance through the sampler: E[f cos / pdf] <= 1 (white furnace per lobe)
        est = (f[ok].astype(np.float64) * np.abs(wi[ok][:, 2:3]) / spdf[ok][:, None]).sum(axis=0) / N
        assert (est <= 1.0 + 0.02).all() and (est > 0.0).all()
        if kind == 0:
            assert np.allclose(est, params[:3], rtol=5e-3)                          # Lambertian: exactly the albedo


def test_diffuse_lobes_are_reciprocal_and_oren_nayar_reduces_to_lambert(oracle):
    rng = np.random.default_rng(3)
    a, b = _uniform_sphere(rng, 5000), _uniform_sphere(rng, 5000)
    a[:, 2], b[:, 2] = np.abs(a[:, 2]), np.abs(b[:, 2])
    for kind, params in ((0, [0.7, 0.7, 0.7]), (1, [0.7, 0.7, 0.7, 0.35])):
        fab, _ = oracle.lobe_f_pdf(kind, params, a, b)
        fba, _ = oracle.lobe_f_pdf(kind, params, b, a)
        assert np.allclose(fab, fba, rtol=1e-4, atol=1e-7)
    lam, _ = oracle.lobe_f_pdf(0, [0.7, 0.7, 0.7], a, b)
    assert np.allclose(lam, 0.7 / np.pi, rtol=1e-6)
    tiny, _ = oracle.lobe_f_pdf(1, [0.7, 0.7, 0.7, 1e-4], a, b)                     # sigma -> 0: A -> 1, B -> 0
    assert np.allclose(tiny, lam, rtol=1e-5)
    below, pdf = oracle.lobe_f_pdf(0, [0.7, 0.7, 0.7], a, -b)                        # other hemisphere: the density vanishes
    assert (pdf == 0).all()


def test_perfect_glass_fresnel_and_energy(oracle):
    eta = 1.5
    white = [1.0, 1.0, 1.0, eta]
    u = np.zeros((1, 2), np.float32)
    # normal incidence: R = ((eta - 1) / (eta + 1))^2 = 0.04, mirror / straight-through directions
    wi, f, pdf, typ = oracle.lobe_sample(2, white, [0.0, 0.0, 1.0], u)
    assert typ[0] != 0 and pdf[0] == 1.0 and np.allclose(wi[0], [0, 0, 1]) and np.allclose(f[0] * abs(wi[0, 2]), 0.04, rtol=1e-6)
    wi, f, pdf, typ = oracle.lobe_sample(3, white, [0.0, 0.0, 1.0], u)
    assert np.allclose(wi[0], [0, 0, -1]) and np.allclose(f[0] * abs(wi[0, 2]), 0.96, rtol=1e-6)
    # every angle, from outside and from inside: reflected + transmitted energy = 1 (no eta^2 scaling in the reference),
    # Snell's law for the refracted direction, total internal reflection beyond the critical angle
    for sign in (1.0, -1.0):
        for theta in np.linspace(0.02, 1.55, 40):
            wo = np.array([np.sin(theta), 0.0, sign * np.cos(theta)], np.float32)
            rwi, rf, _, rt = oracle.lobe_sample(2, white, wo, u)
            twi, tf, _, tt = oracle.lobe_sample(3, white, wo, u)
            assert rt[0] != 0 and np.allclose(rwi[0], [-wo[0], -wo[1], wo[2]])
            r_energy = rf[0, 0] * abs(rwi[0, 2])
            n_i, n_t = (1.0, eta) if sign > 0 else (eta, 1.0)
            if n_i * np.sin(theta) / n_t >= 1.0:                                       # TIR (inside, beyond asin(1 / 1.5) = 41.8 deg)
                assert sign < 0 and tt[0] == 0 and np.isclose(r_energy, 1.0, rtol=1e-6)
                continue
            assert tt[0] != 0 and twi[0, 2] * wo[2] < 0
            assert np.isclose(np.hypot(twi[0, 0], twi[0, 1]), n_i * np.sin(theta) / n_t, rtol=2e-4, atol=1e-5)    # Snell
            assert np.isclose(r_energy + tf[0, 0] * abs(twi[0, 2]), 1.0, rtol=2e-5)


def test_conductor_and_schlick_fresnel_closed_forms(oracle):
    """At normal incidence a conductor reflects ((eta - 1)^2 + k^2) / ((eta + 1)^2 + k^2) and Schlick's form returns Rs; at
    grazing incidence both go to one. Read off the GGX lobe through f * 4 cos_i cos_o / (D G) with wo = wi = wh = n."""
    eta, k = np.array([0.27105, 0.67693, 1.31640]), np.array([3.60920, 2.62480, 2.29210])
    n = np.array([[0.0, 0.0, 1.0]], np.float32)
    alpha = 0.5
    d_g = 1.0 / (np.pi * alpha * alpha)                    # D(n) = 1 / (pi alpha^2), Lambda(n) = 0 -> G = 1
    f, _ = oracle.lobe_f_pdf(4, [1, 1, 1, *eta, *k, alpha], n, n)
    want = ((eta - 1) ** 2 + k ** 2) / ((eta + 1) ** 2 + k ** 2)
    assert np.allclose(f[0] * 4.0 / d_g, want, rtol=1e-4)
    f, _ = oracle.lobe_f_pdf(5, [1, 1, 1, 0.2, 0.5, 0.8, alpha], n, n)
    assert np.allclose(f[0] * 4.0 / d_g, [0.2, 0.5, 0.8], rtol=1e-5)
    g = _unit([[1.0, 0.0, 1e-3]])
    gi = _unit([[-1.0, 0.0, 1e-3]])                         # mirror pair at grazing incidence: cos(wi, wh) ~ 1e-3
    f5, _ = oracle.lobe_f_pdf(5, [1, 1, 1, 0.2, 0.5, 0.8, alpha], g, gi)
    f5w, _ = oracle.lobe_f_pdf(5, [1, 1, 1, 1.0, 1.0, 1.0, alpha], g, gi)
    assert np.allclose(f5[0] / f5w[0], 1.0, atol=5e-3)      # Schlick -> 1 whatever Rs


def test_documented_quirks(oracle):
    """GGX pdf = D cos(theta_h) without abs (trowbridge_reitz.rs:76-78): negative below the surface side of wh; roughness
    floor alpha >= 1e-3 (trowbridge_reitz.rs:16-20)."""
    wo, wi = _unit([[0.3, 0.1, -0.8]]), _unit([[-0.2, 0.2, -0.9]])                    # both below: same hemisphere, wh.z < 0
    _, pdf = oracle.lobe_f_pdf(5, [1, 1, 1, 1, 1, 1, 0.3], wo, wi)
    assert pdf[0] < 0.0
    a, _ = oracle.lobe_f_pdf(5, [1, 1, 1, 1, 1, 1, 0.0], WOS[1][None], WOS[2][None])
    b, _ = oracle.lobe_f_pdf(5, [1, 1, 1, 1, 1, 1, 1e-3], WOS[1][None], WOS[2][None])
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))

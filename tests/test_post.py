"""Film output after the path: EXR writer (app/util.rs:89-110) and the tone-map / heat-map passes
(app/renderpasses/tonemap.rs:318-472) against the numpy restatement in oracle/post.py."""
import numpy as np
import pytest

from oracle import post as P
from yuki_b200 import api, capi, desc as D, scenes, transforms as xf


def _film(h=37, w=53, seed=0, scale=4.0):
    rng = np.random.default_rng(seed)
    f = (rng.random((h, w, 3)) ** 3 * scale).astype(np.float32)
    f[0, 0] = 0.0
    f[1, 1] = [1e-8, 3.5, 1e4]
    return f


def test_exr_round_trip(tmp_path):
    film = _film()
    film[2, 3] = [np.inf, -1.5, np.float32(1e-40)]   # specials and a denormal survive: the writer stores raw f32
    path = tmp_path / "out.exr"
    api.write_exr(path, film)
    back = P.read_exr_rgb(path)
    assert back.shape == film.shape
    assert np.array_equal(back.view(np.uint32), film.view(np.uint32))


def test_exr_header_is_standard(tmp_path):
    path = tmp_path / "o.exr"
    api.write_exr(path, np.zeros((2, 3, 3), np.float32))
    data = path.read_bytes()
    assert data[:4] == bytes([0x76, 0x2f, 0x31, 0x01]) and data[4:8] == bytes([2, 0, 0, 0])
    for needle in (b"channels\0chlist\0", b"compression\0compression\0", b"dataWindow\0box2i\0", b"displayWindow\0box2i\0",
                   b"lineOrder\0lineOrder\0", b"pixelAspectRatio\0float\0", b"screenWindowCenter\0v2f\0", b"screenWindowWidth\0float\0"):
        assert needle in data
    # header + 2-entry offset table + 2 scanline blocks of (4 + 4 + 3 px * 3 ch * 4 B)
    assert len(data) == data.index(b"screenWindowWidth") + len(b"screenWindowWidth\0float\0") + 4 + 4 + 1 + 2 * 8 + 2 * (8 + 36)


def test_exr_rejects_bad_arguments(tmp_path):
    with pytest.raises(capi.YukiGpuError):
        api.write_exr(tmp_path / "missing_dir" / "x.exr", np.zeros((2, 2, 3), np.float32))


def test_numpy_heatmap_gradient_endpoints():
    film = np.zeros((1, 3, 3), np.float32)
    film[0, :, 1] = [0.0, 5.0, 10.0]
    img = P.heatmap(film, 1, 0.0, 10.0)
    assert img[0, 0].tolist() == [0, 0, 1] and img[0, 1].tolist() == [0, 1, 0] and img[0, 2].tolist() == [1, 0, 0]
    assert P.find_min_max(film, 1) == (0.0, 10.0)


@pytest.mark.gpu
def test_tonemap_and_heatmap_kernels_match_the_restatement():
    ctx = api.Context(0)
    film = _film(67, 130)
    out = api.tonemap_filmic(ctx, film, exposure=1.7)
    ref = P.tonemap_filmic(film, 1.7)
    assert out.min() >= 0.0 and out.max() <= 1.0
    np.testing.assert_allclose(out, ref, rtol=0, atol=2e-6)
    # accumulating film: per-tile sample counts, 16-px tiles, the shader's truncated x tile count
    ts = np.arange(1, (130 // 16) * ((67 + 15) // 16) + 1, dtype=np.float32)
    ts[3] = 0.0
    out = api.tonemap_filmic(ctx, film * 5.0, exposure=0.9, tile_samples=ts, tile_dim=16)
    ref = P.tonemap_filmic(film * 5.0, 0.9, ts, 16)
    np.testing.assert_allclose(out, ref, rtol=0, atol=2e-6)
    for channel in range(4):
        img, (lo, hi) = api.heatmap(ctx, film, channel)
        assert (lo, hi) == P.find_min_max(film, channel)
        np.testing.assert_allclose(img, P.heatmap(film, channel, lo, hi), rtol=0, atol=2e-6)
    img, rng = api.heatmap(ctx, film, 2, value_range=(0.5, 2.5))
    assert rng == (0.5, 2.5)
    np.testing.assert_allclose(img, P.heatmap(film, 2, 0.5, 2.5), rtol=0, atol=2e-6)
    ctx.close()


@pytest.mark.gpu
def test_bvh_heatmap_of_a_render_round_trips_through_exr(tmp_path):
    """Config-3 flow end to end: BVHIntersections film -> EXR on disk -> heat map of the red (test count) channel."""
    scene, cam = scenes.heightfield(xf, 40, 32, seed=2)
    film = D.FilmSettings((96, 64), 16)
    ctx = api.Context(0)
    dev = api.Scene(ctx, scene)
    r = api.Renderer(ctx).render(dev, cam, film, D.SamplerType.uniform(1), D.IntegratorType.bvh_intersections())
    path = tmp_path / "bvh.exr"
    api.write_exr(path, r.film)
    back = P.read_exr_rgb(path)
    assert np.array_equal(back.view(np.uint32), r.film.view(np.uint32))
    img, (lo, hi) = api.heatmap(ctx, back, 1)
    assert lo == back[..., 1].min() and hi == back[..., 1].max() and hi > lo
    np.testing.assert_allclose(img, P.heatmap(back, 1, lo, hi), rtol=0, atol=2e-6)
    dev.close(); ctx.close()

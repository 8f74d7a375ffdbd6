"""Parity at BASELINE.json's scene sizes (the toy-size tests cannot reach the packing limits, the device-side repack of
17 M nodes, leaf refs beyond 2^20, or BVHs that do not fit the caches):
  configs[2]: the BVH-intersections debug integrator on the ~1 M-triangle heightfield, SAH / Middle / EqualCounts — counter
              image, primary-hit ids and the traversal totals bit-exact against the oracle on a 160x120 film;
  configs[4]: a handful of spiral tiles of the 10 M-triangle scene at 3840x2160, Path max_depth 8 — film bits, ray counts and
              traversal totals against the oracle, with the renderer's defaults (shadow rays one per lane on this scene) and
              with the ray sort on.
The oracle needs ~4 s per 1 M-triangle BVH and ~50 s for the 10 M-triangle one; the films are small so that it finishes."""
import numpy as np
import pytest

from yuki_b200 import api, desc as D, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("split", [D.SPLIT_SAH, D.SPLIT_MIDDLE, D.SPLIT_EQUAL_COUNTS])
def test_config3_one_million_triangles_counters_bit_exact(gpu_ctx, oracle, xf, split):
    scene, cam = scenes.heightfield(xf, 708, 708, split_method=split)   # 1 002 528 triangles
    film = D.FilmSettings((160, 120), 16)
    smp, integ = D.SamplerType.uniform(1), D.IntegratorType.bvh_intersections()
    dev = api.Scene(gpu_ctx, scene)
    assert dev.host.n_tris == 2 * 708 * 708
    osc = oracle.OracleScene(scene)
    r = api.Renderer(gpu_ctx).render(dev, cam, film, smp, integ, want_hit_ids=True)
    o_img, o_ids, o_st = osc.render(cam, film, smp, integ, want_hit_ids=True)
    assert np.array_equal(r.hit_ids, o_ids) and (o_ids >= 0).sum() > 1000
    assert np.array_equal(r.film.view(np.uint32), o_img.view(np.uint32))
    assert r.stats.closest_nodes == o_st.closest_nodes == int(o_img[..., 0].astype(np.float64).sum())
    assert r.stats.closest_tris == o_st.closest_tris
    assert r.stats.primary_hit_hash == o_st.primary_hit_hash and r.stats.ray_count == o_st.ray_count == 160 * 120
    # incoherent rays straight through the ABI on the same BVH: per-ray distances, ids and (tests, hits) counters
    rng = np.random.default_rng(21 + split)
    o = rng.uniform((-0.7, -0.3, -0.7), (0.7, 0.5, 0.7), (20000, 3)).astype(np.float32)
    d = (rng.uniform((-0.7, -0.3, -0.7), (0.7, 0.5, 0.7), (20000, 3)).astype(np.float32) - o)
    t, ids, cnt = dev.intersect(o, d)
    ot, oids, ocnt = osc.trace(o, d)
    assert np.array_equal(ids, oids) and np.array_equal(t.view(np.uint32), ot.view(np.uint32)) and np.array_equal(cnt, ocnt)
    dev.close()


def test_config5_ten_million_triangles_path_tiles_bit_exact(gpu_ctx, oracle, xf):
    scene, cam = scenes.terrain_room(xf)                               # 10 008 056 triangles, 3 lights, every material kind
    film = D.FilmSettings((3840, 2160), 16)
    smp, integ = D.SamplerType.stratified(2, 2), D.IntegratorType.path(8)
    tiles = api.film_tiles(film)
    assert len(tiles) == 240 * 135
    sel = np.ascontiguousarray(tiles[[0, 1, 777, 5000, 12345, 20000, len(tiles) - 1]])   # centre, mid-spiral, far corner
    dev = api.Scene(gpu_ctx, scene)
    assert dev.host.n_tris > 10_000_000 and dev.host.n_nodes > 17_000_000
    osc = oracle.OracleScene(scene)
    o_img, o_ids, o_st = osc.render(cam, film, smp, integ, tiles=sel, want_hit_ids=True)
    assert o_st.samples == len(sel) * 256 * 4 and (o_ids >= 0).sum() > 256
    rn = api.Renderer(gpu_ctx)
    for kw in ({}, {"ray_sort": 2}, {"ray_sort": 3 | 16, "pipes": 1}):
        r = rn.render(dev, cam, film, smp, integ, tiles=sel, want_hit_ids=True, **kw)
        assert np.array_equal(r.hit_ids, o_ids), kw
        assert np.array_equal(r.film.view(np.uint32), o_img.view(np.uint32)), kw
        for k in ("ray_count", "shadow_rays", "samples", "closest_nodes", "closest_tris", "any_nodes", "any_tris", "primary_hit_hash"):
            assert getattr(r.stats, k) == getattr(o_st, k), (kw, k)
    # the terrain's own hit ids reach past 2^23: the shape words keep all 32 bits of the original id
    big = rn.render(dev, cam, film, D.SamplerType.uniform(1), D.IntegratorType.bvh_intersections(), tiles=tiles[:64], want_hit_ids=True)
    o_big = osc.render(cam, film, D.SamplerType.uniform(1), D.IntegratorType.bvh_intersections(), tiles=tiles[:64], want_hit_ids=True)
    assert np.array_equal(big.hit_ids, o_big[1]) and np.array_equal(big.film.view(np.uint32), o_big[0].view(np.uint32))
    assert int(o_big[1].max()) > (1 << 22)
    dev.close()

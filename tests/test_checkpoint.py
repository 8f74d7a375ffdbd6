"""Checkpoint / resume of accumulating renders: the film's running sum and per-tile sample counts
(film.rs:74, 260-272) are a complete checkpoint because the samplers seek to any (pixel, sample index). A render stopped
after k sample indices, saved, restored and continued must equal the uninterrupted render bit for bit."""
import os

import numpy as np
import pytest

from yuki_b200 import api, desc as D, scenes


def _setup(xf):
    scene, cam = scenes.cornell(xf, light="rect", tall_box="glass")
    fs = D.FilmSettings((48, 40), 16, accumulate=True)
    return scene, cam, fs, D.SamplerType.stratified(3, 3), D.IntegratorType.path(5)


def _check_resume(tmp_path, renderer, render_fn_for, fs, smp, integ, scene_args):
    ck = str(tmp_path / "film.npz")
    whole = api.Film(fs)
    assert renderer.render_progressive(*scene_args, whole, smp, integ, samples_per_pass=100, render_fn=render_fn_for(whole)) == 9
    assert (whole.samples == 9).all()
    # stop after 2 passes of 2 sample indices, checkpointing each pass
    part = api.Film(fs)
    assert renderer.render_progressive(*scene_args, part, smp, integ, samples_per_pass=2, checkpoint_path=ck, max_passes=2,
                                       render_fn=render_fn_for(part)) == 4
    assert os.path.exists(ck) and not [f for f in os.listdir(tmp_path) if ".tmp." in f]
    assert not np.array_equal(part.pixels, whole.pixels)
    # a new process would start here
    back = api.Film.load(ck)
    assert (back.samples == 4).all() and np.array_equal(back.pixels.view(np.uint32), part.pixels.view(np.uint32))
    assert renderer.render_progressive(*scene_args, back, smp, integ, samples_per_pass=4, checkpoint_path=ck,
                                       render_fn=render_fn_for(back)) == 9
    assert (back.samples == 9).all()
    assert np.array_equal(back.pixels.view(np.uint32), whole.pixels.view(np.uint32))
    # finished film: nothing left to do, and the final checkpoint holds it
    assert renderer.render_progressive(*scene_args, back, smp, integ, render_fn=render_fn_for(back)) == 9
    assert np.array_equal(api.Film.load(ck).pixels.view(np.uint32), whole.pixels.view(np.uint32))
    return whole, ck


def test_resume_equals_uninterrupted_render_with_the_oracle_as_renderer(tmp_path, oracle, xf):
    scene, cam, fs, smp, integ = _setup(xf)
    osc = oracle.OracleScene(scene)

    def render_fn_for(film):
        return lambda tiles, film_out: osc.render(cam, fs, smp, integ, tiles=tiles, film_out=film_out, threads=1)

    rn = api.Renderer.__new__(api.Renderer)   # the host logic only: no GPU context on this path
    whole, ck = _check_resume(tmp_path, rn, render_fn_for, fs, smp, integ, (None, cam))
    # the mean of the accumulated film is the averaging render (integrators/mod.rs:172-182 divides once, in f32)
    mean, _, _ = osc.render(cam, D.FilmSettings(fs.res, fs.tile_dim), smp, integ)
    assert np.abs(whole.pixels / np.float32(9) - mean).max() <= 1e-5 * max(1.0, float(mean.max()))
    # guards
    with pytest.raises(ValueError, match="other settings"):
        rn.render_progressive(None, cam, api.Film.load(ck), D.SamplerType.stratified(3, 3, jitter=False), integ, render_fn=render_fn_for(whole))
    with pytest.raises(ValueError, match="other settings"):   # another indirect clamp / another camera is another render
        rn.render_progressive(None, cam, api.Film.load(ck), smp, D.IntegratorType.path(5, indirect_clamp=2.0), render_fn=render_fn_for(whole))
    other_cam = D.CameraParameters((0.1, 0.2, 0.9), cam.target, fov_axis=cam.fov_axis, fov_deg=cam.fov_deg)
    with pytest.raises(ValueError, match="other settings"):
        rn.render_progressive(None, other_cam, api.Film.load(ck), smp, integ, render_fn=render_fn_for(whole))
    with pytest.raises(ValueError, match="another render"):
        api.Film.load(ck, expect_meta={"spp": 16})
    # per-tile sample counts are u32 like the reference's Vec<u32> (film.rs:74): 65536 passes do not wrap to zero
    assert whole.samples.dtype == np.uint32
    big = api.Film(fs)
    big.samples[...] = 65535
    big.samples += 1
    assert int(big.samples.min()) == 65536
    uneven = api.Film.load(ck)
    uneven.samples[0] -= 1
    with pytest.raises(ValueError, match="different sample counts"):
        rn.render_progressive(None, cam, uneven, smp, integ, render_fn=render_fn_for(uneven))
    with pytest.raises(ValueError, match="accumulating"):
        rn.render_progressive(None, cam, api.Film(D.FilmSettings(fs.res, fs.tile_dim)), smp, integ, render_fn=render_fn_for(whole))
    over = api.Film.load(ck)
    with pytest.raises(ValueError, match="already holds"):
        rn.render_progressive(None, cam, over, D.SamplerType.stratified(2, 2), integ, render_fn=render_fn_for(over))


def test_film_save_is_atomic_and_round_trips(tmp_path):
    fs = D.FilmSettings((33, 17), 16, accumulate=True)
    f = api.Film(fs)
    f.pixels[...] = np.random.default_rng(0).normal(size=f.pixels.shape).astype(np.float32)
    f.samples[...] = 7
    p = str(tmp_path / "a.npz")
    f.save(p, meta={"k": [1, 2]})
    (tmp_path / "a.npz.tmp.999").write_bytes(b"torn")      # a crashed writer's leftover is never read
    g = api.Film.load(p, expect_meta={"k": (1, 2)})
    assert g.settings.res == (33, 17) and g.settings.tile_dim == 16 and g.settings.accumulate
    assert np.array_equal(g.pixels.view(np.uint32), f.pixels.view(np.uint32)) and np.array_equal(g.samples, f.samples)
    plain = api.Film(D.FilmSettings((8, 8), 16))
    plain.save(p)
    assert api.Film.load(p).samples is None


@pytest.mark.gpu
def test_gpu_resume_equals_uninterrupted_render(tmp_path, gpu_ctx, oracle, xf):
    scene, cam, fs, smp, integ = _setup(xf)
    dev = api.Scene(gpu_ctx, scene)
    rn = api.Renderer(gpu_ctx)
    whole, _ = _check_resume(tmp_path, rn, lambda film: None, fs, smp, integ, (dev, cam))
    tiles = np.concatenate([api.film_tiles(fs)] * 9)
    tiles["sample"] = np.repeat(np.arange(9, dtype=np.uint16), len(tiles) // 9)
    # one worker: with several, the reference (and the oracle) add the samples of a pixel in whatever order the workers
    # finish; the GPU adds them in tile-list order, which is the single-thread order (use_single_render_thread)
    o_img, _, _ = oracle.OracleScene(scene).render(cam, fs, smp, integ, tiles=tiles, threads=1)
    assert np.array_equal(whole.pixels.view(np.uint32), o_img.view(np.uint32))
    dev.close()

"""`python -m yuki_b200.headless`, the counterpart of `yuki --out=foo.exr` (app/headless.rs:24-111): config 1 of
BASELINE.json end to end — a pbrt-v3 Cornell box file, 512x512, Whitted, 16 spp stratified, headless EXR — against the
oracle rendering the same file, plus the scene-file dispatch of app/util.rs:15-66."""
import io

import numpy as np
import pytest

from yuki_b200 import api, desc as D, headless
from test_pbrt import cornell_pbrt


def test_scene_dispatch_follows_the_file_extension(tmp_path):
    with pytest.raises(FileNotFoundError):
        headless.try_load_scene(str(tmp_path / "missing.pbrt"))
    (tmp_path / "noext").write_text("x")
    with pytest.raises(ValueError, match="extension"):
        headless.try_load_scene(str(tmp_path / "noext"))
    (tmp_path / "scene.obj").write_text("x")
    with pytest.raises(ValueError, match="Unknown extension 'obj'"):
        headless.try_load_scene(str(tmp_path / "scene.obj"))
    path, ref, cam = cornell_pbrt(tmp_path)
    sc, lcam, film = headless.try_load_scene(str(path))
    assert film.res == (96, 96) and len(sc.meshes) == len(ref.meshes)
    sc, lcam, film = headless.try_load_scene("")                      # Scene::cornell()
    assert film.res == (640, 480) and len(sc.spheres) == 1 and sc.split_method == D.SPLIT_MIDDLE


def test_flags_map_to_the_reference_settings():
    a = headless.parse_args(["--out", "x.exr"])
    smp, integ = headless.settings_from_args(a)
    assert (integ.kind, integ.max_depth) == (D.INTEGRATOR_WHITTED, 3)                         # whitted.rs:21-25
    assert (smp.kind, smp.nx, smp.ny, smp.jitter) == (D.SAMPLER_STRATIFIED, 1, 1, True)      # stratified.rs:26-34
    a = headless.parse_args(["--out", "x.exr", "--integrator", "path", "--max-depth", "8", "--indirect-clamp", "2.5", "--sampler", "uniform", "8",
                             "--seed", "0x10"])
    smp, integ = headless.settings_from_args(a)
    assert (integ.kind, integ.max_depth, integ.indirect_clamp) == (D.INTEGRATOR_PATH, 8, 2.5)
    assert (smp.kind, smp.samples_per_pixel(), smp.seed) == (D.SAMPLER_UNIFORM, 8, 16)
    with pytest.raises(ValueError):
        headless.settings_from_args(headless.parse_args(["--out", "x.exr", "--sampler", "stratified", "4"]))


@pytest.mark.gpu
def test_config_1_headless_exr_equals_the_oracle(tmp_path, oracle):
    from oracle import post
    path, _, _ = cornell_pbrt(tmp_path, with_ply=True)
    out = tmp_path / "cornell.exr"
    log = io.StringIO()
    args = headless.parse_args(["--scene", str(path), "--out", str(out), "--integrator", "whitted", "--max-depth", "3", "--sampler", "stratified", "4", "4",
                                "--res", "512", "512", "--tone-map", "raw"])
    pixels = headless.render(args, out=log)
    assert "Render finished in" in log.getvalue() and "Msamples/s" in log.getvalue()   # (a render this short may finish before a progress line)
    exr = post.read_exr_rgb(str(out))
    assert exr.shape == (512, 512, 3) and np.array_equal(exr.view(np.uint32), pixels.view(np.uint32))
    sc, cam, _ = api.load_pbrt(str(path))
    o_img, _, _ = oracle.OracleScene(sc).render(cam, D.FilmSettings((512, 512), 16), D.SamplerType.stratified(4, 4), D.IntegratorType.whitted(3))
    assert np.array_equal(exr.view(np.uint32), o_img.view(np.uint32)) and float(o_img.mean()) > 0.01
    # the display passes: filmic and heat map outputs are written the same way
    for tm in ("filmic", "heatmap"):
        args = headless.parse_args(["--scene", str(path), "--out", str(tmp_path / f"{tm}.exr"), "--res", "64", "64", "--tone-map", tm])
        img = headless.render(args, out=io.StringIO())
        assert np.array_equal(post.read_exr_rgb(str(tmp_path / f"{tm}.exr")).view(np.uint32), img.view(np.uint32))
        assert float(img.min()) >= 0.0 and float(img.max()) <= 1.0 + 1e-6

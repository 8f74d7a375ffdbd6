"""Several GPUs: (a) `yk_multi_*` — all devices of one process behind the C ABI, tiles popped from one shared cursor, the film
assembled on the first device by peer stores (csrc/multi.inl; the reference's RenderManager + shared tile queue,
renderer/render_manager.rs:78-97,197-236); (b) the one-process-per-GPU path bench.py runs: each NCCL rank renders its
interleaved share of the spiral tile list and one ncclReduce(sum) assembles the film on rank 0. Both must give the
single-GPU film bit for bit (tiles are disjoint). Tests that need two devices are skipped on a one-GPU box; the host logic
of the N > 1 path is covered on CPU by tests/test_multi_rank.py (gloo)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from yuki_b200 import api, capi, desc as D, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


def test_multi_symbols_reject_bad_arguments():
    L = capi.lib()
    assert L.yk_multi_create(None, 0, None) != 0
    assert b"yk_multi_create" in L.yk_last_error()
    assert L.yk_multi_device_count(None) == 0
    assert L.yk_multi_render(None, None, None, None, None, None, None, 0, None, None, None, None) != 0


@pytest.mark.gpu
def test_group_of_one_device_equals_yk_render(gpu_ctx, xf):
    scene, cam = scenes.cornell(xf, light="rect", tall_box="glass")
    film = D.FilmSettings((96, 80), 16)
    smp, integ = D.SamplerType.stratified(2, 2), D.IntegratorType.path(6)
    dev = api.Scene(gpu_ctx, scene)
    ref = api.Renderer(gpu_ctx).render(dev, cam, film, smp, integ, want_hit_ids=True)
    mctx = api.MultiContext([0])
    ms = api.MultiScene(mctx, scene)
    r, per = api.multi_render(mctx, ms, cam, film, smp, integ, want_hit_ids=True)
    assert np.array_equal(r.film.view(np.uint32), ref.film.view(np.uint32)) and np.array_equal(r.hit_ids, ref.hit_ids)
    assert r.stats.ray_count == ref.stats.ray_count == per[0].ray_count
    with pytest.raises(capi.YukiGpuError):
        api.MultiContext([0, 0])
    ms.close(); mctx.close(); dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("run_tiles", [None, "5"])
def test_tile_cursor_runs_on_one_device_give_the_single_call_film(gpu_ctx, xf, monkeypatch, run_tiles):
    """The shared tile cursor on a one-GPU box: a render large enough for guided self-scheduling (remaining / 2G tiles per pop,
    never below ~16 Mi paths: 1024 tiles of 64 Ki paths -> runs of 512, 256, 256 tiles), and fixed runs of five tiles of the strided
    order, against one yk_render call: the same film, ids and counters whatever the partition into runs."""
    monkeypatch.setenv("YK_MULTI_FORCE_CURSOR", "1")   # a group of one device normally forwards to yk_render
    if run_tiles:
        monkeypatch.setenv("YK_MULTI_RUN_TILES", run_tiles)
    else:
        monkeypatch.delenv("YK_MULTI_RUN_TILES", raising=False)
        monkeypatch.setenv("YK_MULTI_STATIC_BELOW", "0")   # (jobs this short are split statically by default)
    scene, cam = scenes.cornell(xf, light="rect", tall_box="glass", sphere=True)
    film = D.FilmSettings((512, 512), 16) if run_tiles is None else D.FilmSettings((150, 100), 16)
    smp = D.SamplerType.stratified(16, 16) if run_tiles is None else D.SamplerType.stratified(2, 2)
    integ = D.IntegratorType.path(4)
    dev = api.Scene(gpu_ctx, scene)
    ref = api.Renderer(gpu_ctx).render(dev, cam, film, smp, integ, want_hit_ids=True)
    mctx = api.MultiContext([0])
    ms = api.MultiScene(mctx, scene)
    r, per = api.multi_render(mctx, ms, cam, film, smp, integ, want_hit_ids=True)
    assert np.array_equal(r.film.view(np.uint32), ref.film.view(np.uint32)) and np.array_equal(r.hit_ids, ref.hit_ids)
    for k in ("ray_count", "shadow_rays", "samples", "closest_nodes", "closest_tris", "any_nodes", "any_tris", "primary_hit_hash"):
        assert getattr(r.stats, k) == getattr(ref.stats, k) == getattr(per[0], k), k
    assert r.stats.kernel_launches > ref.stats.kernel_launches   # several runs: more launches than the single call
    ms.close(); mctx.close(); dev.close()


def _check_group(xf, oracle, monkeypatch, no_peer):
    if no_peer:
        monkeypatch.setenv("YK_MULTI_NO_PEER", "1")
    monkeypatch.setenv("YK_MULTI_RUN_TILES", "3")   # many small runs: every device takes part even on a small film
    n = min(_n_gpus(), 4)
    mctx = api.MultiContext(list(range(n)))
    assert mctx.peer_stores()[0] and (not no_peer or not any(mctx.peer_stores()[1:]))
    single = api.Context(0)
    try:
        for scene, cam in (scenes.material_room(xf), scenes.cornell(xf, light="rect", tall_box="glass", sphere=True)):
            film = D.FilmSettings((150, 100), 16)   # ragged: 10 x 7 tiles, the last column / row clipped
            smp, integ = D.SamplerType.stratified(2, 2), D.IntegratorType.path(8)
            dev = api.Scene(single, scene)
            ref = api.Renderer(single).render(dev, cam, film, smp, integ, want_hit_ids=True)
            ms = api.MultiScene(mctx, scene)
            r, per = api.multi_render(mctx, ms, cam, film, smp, integ, want_hit_ids=True)
            assert np.array_equal(r.film.view(np.uint32), ref.film.view(np.uint32))
            assert np.array_equal(r.hit_ids, ref.hit_ids)
            for k in ("ray_count", "shadow_rays", "samples", "closest_nodes", "closest_tris", "any_nodes", "any_tris", "primary_hit_hash"):
                want = getattr(ref.stats, k)
                if k == "primary_hit_hash":
                    assert sum(getattr(p, k) for p in per) % (1 << 64) == want == getattr(r.stats, k)
                else:
                    assert sum(getattr(p, k) for p in per) == want == getattr(r.stats, k), k
            assert sum(1 for p in per if p.samples > 0) == n   # the shared cursor fed every device
            # a tile subset into a film that already holds pixels: untouched pixels survive, and Whitted + a debug integrator
            tiles = api.film_tiles(film)[::3]
            base = np.full((100, 150, 3), 0.25, np.float32)
            a = api.Renderer(single).render(dev, cam, film, smp, D.IntegratorType.whitted(4), tiles=tiles, film_out=base.copy())
            b, _ = api.multi_render(mctx, ms, cam, film, smp, D.IntegratorType.whitted(4), tiles=tiles, film_out=base.copy())
            assert np.array_equal(a.film.view(np.uint32), b.film.view(np.uint32))
            # accumulating film: one tile list per sample index, added; the per-pixel add order is fixed (tile.index mod G)
            acc = D.FilmSettings(film.res, film.tile_dim, accumulate=True)
            t_all = api.film_tiles(film)
            rep = np.concatenate([t_all] * 4)
            rep["sample"] = np.repeat(np.arange(4, dtype=np.uint16), len(t_all))
            a = api.Renderer(single).render(dev, cam, acc, smp, integ, tiles=rep)
            b, _ = api.multi_render(mctx, ms, cam, acc, smp, integ, tiles=rep)
            assert np.array_equal(a.film.view(np.uint32), b.film.view(np.uint32))
            # ... also onto a film that already holds a sum (a resumed render): (film + s0) + s1 ... on every device
            a = api.Renderer(single).render(dev, cam, acc, smp, integ, tiles=rep, film_out=base.copy())
            b, _ = api.multi_render(mctx, ms, cam, acc, smp, integ, tiles=rep, film_out=base.copy())
            assert np.array_equal(a.film.view(np.uint32), b.film.view(np.uint32))
            o_img, _, _ = oracle.OracleScene(scene).render(cam, film, smp, integ)
            assert np.array_equal(r.film.view(np.uint32), o_img.view(np.uint32))
            ms.close(); dev.close()
    finally:
        single.close(); mctx.close()


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_device_group_film_is_bit_identical_to_one_device(xf, oracle, monkeypatch):
    _check_group(xf, oracle, monkeypatch, no_peer=False)


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_device_group_without_peer_stores_gathers_the_film(xf, oracle, monkeypatch):
    _check_group(xf, oracle, monkeypatch, no_peer=True)


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_two_nccl_ranks_reduce_to_the_single_gpu_film(tmp_path):
    """The path bench.py runs at N > 1: one process per GPU, tiles[rank::world], one NCCL sum-reduce to rank 0."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "nccl_film.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "tools", "nccl_film_check.py"), str(out)]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    got = np.load(out)
    assert bool(got["equal"]) and int(got["world"]) == 2
    assert int(got["ray_count_sum"]) == int(got["ray_count_single"])

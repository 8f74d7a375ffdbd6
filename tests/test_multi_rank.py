"""N > 1 host logic on CPU: two gloo ranks each take their interleaved share of the spiral tile list, fill a zeroed
film (the oracle stands in for the GPU renderer here), and a sum-reduce to rank 0 assembles the single-rank film bit
for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from yuki_b200 import api, desc as D, dist as ydist, scenes, transforms as xf
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene, cam = scenes.cornell(xf, light="rect", tall_box="glass")
    film = D.FilmSettings((70, 50), 16)
    smp, integ = D.SamplerType.stratified(2, 2), D.IntegratorType.path(5)
    tiles = api.film_tiles(film)
    mine = ydist.partition_tiles(tiles, rank, world)
    img, _, st = O.OracleScene(scene).render(cam, film, smp, integ, tiles=mine, threads=2)
    t = torch.from_numpy(img)
    ydist.reduce_film(t, dst=0)
    counts = ydist.reduce_stats([st.ray_count, st.samples], dst=0)
    if rank == 0:
        np.savez(out_path, film=t.numpy(), counts=counts.numpy(), n_mine=len(mine), n_all=len(tiles))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_disjoint_and_complete():
    from yuki_b200 import api, desc as D, dist as ydist
    tiles = api.film_tiles(D.FilmSettings((3840, 2160), 16))
    for world in (1, 2, 4, 8):
        parts = [ydist.partition_tiles(tiles, r, world) for r in range(world)]
        assert sum(len(p) for p in parts) == len(tiles) == 240 * 135
        assert sorted(np.concatenate([p["index"] for p in parts]).tolist()) == list(range(len(tiles)))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1       # balanced to one tile
        assert parts[0][0]["index"] == tiles[0]["index"]                            # rank 0 starts at the centre tile
    with pytest.raises(ValueError):
        ydist.partition_tiles(tiles, 2, 2)


def test_two_rank_gloo_reduce_equals_single_rank(tmp_path, oracle, xf):
    import torch.multiprocessing as mp
    from yuki_b200 import desc as D, scenes
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    scene, cam = scenes.cornell(xf, light="rect", tall_box="glass")
    ref, _, st = oracle.OracleScene(scene).render(cam, D.FilmSettings((70, 50), 16), D.SamplerType.stratified(2, 2), D.IntegratorType.path(5))
    assert np.array_equal(got["film"].view(np.uint32), ref.view(np.uint32))
    assert int(got["counts"][0]) == st.ray_count and int(got["counts"][1]) == st.samples == 70 * 50 * 4
    assert got["n_mine"] == (got["n_all"] + 1) // 2

"""Launched by tests/test_multi_gpu.py under torchrun with 2 ranks: every rank renders its interleaved share of the spiral
tile list on its own GPU into a zeroed device film, one NCCL sum-reduce assembles the film on rank 0, which compares it bit
for bit with the film it renders alone (the path bench.py times at N > 1)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from yuki_b200 import api, capi, desc as D, dist as ydist, scenes, transforms as xf  # noqa: E402


def main(out_path):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scene, cam = scenes.material_room(xf)
    film = D.FilmSettings((200, 120), 16)
    smp, integ = D.SamplerType.stratified(2, 2), D.IntegratorType.path(8)
    ctx = api.Context(local)
    dev = api.Scene(ctx, scene)
    rn = api.Renderer(ctx)
    stream = torch.cuda.ExternalStream(capi.lib().yk_context_stream(ctx._h), device=torch.device("cuda", local))
    tiles = api.film_tiles(film)
    mine = ydist.partition_tiles(tiles, rank, world)
    with torch.cuda.stream(stream):
        d_film = torch.zeros(film.res[0] * film.res[1] * 3, dtype=torch.float32, device="cuda")
        r = rn.render(dev, cam, film, smp, integ, tiles=mine, device_film_ptr=d_film.data_ptr())
        ydist.reduce_film(d_film, dst=0)
        counts = torch.tensor([float(r.stats.ray_count)], dtype=torch.float64, device="cuda")
        dist.reduce(counts, dst=0, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
    if rank == 0:
        single = rn.render(dev, cam, film, smp, integ)
        got = d_film.cpu().numpy().reshape(film.res[1], film.res[0], 3)
        np.savez(out_path, equal=np.array_equal(got.view(np.uint32), single.film.view(np.uint32)), world=world,
                 ray_count_sum=int(counts.item()), ray_count_single=int(single.stats.ray_count))
    dist.barrier()
    dev.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])

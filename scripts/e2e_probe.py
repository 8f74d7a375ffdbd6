"""Where the end-to-end time of the 10 M-triangle workload goes: times every call of bench.py's e2e step separately."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yuki_b200 import api, desc as D, scenes, transforms as xf
side = int(sys.argv[1]) if len(sys.argv) > 1 else 4
which = sys.argv[2] if len(sys.argv) > 2 else "terrain"
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # render tiles[0::stride]: one rank's share of an N-rank job
if which == "cornell":
    s, c = scenes.cornell(xf, light="rect", tall_box="glass")
    res = (1024, 1024)
else:
    s, c = scenes.terrain_room(xf)
    res = (3840, 2160)
t0 = time.perf_counter(); hs = api.HostScene(s); print(f"host scene build {time.perf_counter() - t0:.2f} s", flush=True)
ctx = api.Context(0)
dev = api.Scene(ctx, s, host=hs)
rn = api.Renderer(ctx)
film = D.FilmSettings(res, 16)
smp, integ = D.SamplerType.stratified(side, side), D.IntegratorType.path(8)
tiles = np.ascontiguousarray(api.film_tiles(film)[0::stride])
film_host = np.zeros((res[1], res[0], 3), np.float32)
for i in range(3):
    t0 = time.perf_counter(); r = rn.render(dev, c, film, smp, integ, tiles=tiles, film_out=film_host); t1 = time.perf_counter()
    print(f"resident scene, host film: wall {1e3*(t1-t0):.1f} ms, lib wall {1e3*r.stats.seconds:.1f} ms, device {r.stats.device_ms:.1f} ms", flush=True)
for pipes in (1, 2):
    for i in range(3):
        t0 = time.perf_counter(); r = rn.render(dev, c, film, smp, integ, tiles=tiles, film_out=film_host, pipes=pipes); t1 = time.perf_counter()
        print(f"pipes {pipes}: wall {1e3*(t1-t0):.1f} ms, lib wall {1e3*r.stats.seconds:.1f} ms, device {r.stats.device_ms:.1f} ms, launches {r.stats.kernel_launches}", flush=True)
for i in range(3):
    t0 = time.perf_counter(); d2 = api.Scene(ctx, s, host=hs); t1 = time.perf_counter()
    r = rn.render(d2, c, film, smp, integ, tiles=tiles, film_out=film_host); t2 = time.perf_counter()
    d2.close(); t3 = time.perf_counter()
    print(f"e2e step: scene create {1e3*(t1-t0):.1f} ms, render wall {1e3*(t2-t1):.1f} ms (lib {1e3*r.stats.seconds:.1f}, device {r.stats.device_ms:.1f}), "
          f"scene destroy {1e3*(t3-t2):.1f} ms", flush=True)

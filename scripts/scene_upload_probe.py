"""Times yk_scene_create (flattened host arrays -> device scene) on the 10 M-triangle scene."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yuki_b200 import api, scenes, transforms as xf
s, c = scenes.terrain_room(xf)
t0 = time.perf_counter(); hs = api.HostScene(s); t1 = time.perf_counter()
print(f"host scene (transform + BVH build + flatten): {t1 - t0:.2f} s, {hs.n_tris} triangles, {hs.n_nodes} nodes", flush=True)
ctx = api.Context(0)
for i in range(4):
    t0 = time.perf_counter(); dev = api.Scene(ctx, s, host=hs); t1 = time.perf_counter()
    print(f"yk_scene_create #{i}: {1e3 * (t1 - t0):.1f} ms ({(hs.n_nodes * 32 + hs.n_tris * 49) / (t1 - t0) / 1e9:.2f} GB/s of input arrays)", flush=True)
    dev.close()

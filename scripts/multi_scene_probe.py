"""Times yk_multi_scene_create of the 10 M-triangle scene on all devices of the box (YK_MULTI_NO_CLONE=1: every device uploads)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yuki_b200 import api, scenes, transforms as xf
n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
s, c = scenes.terrain_room(xf)
hs = api.HostScene(s)
mctx = api.MultiContext(list(range(n)))
for i in range(4):
    t0 = time.perf_counter(); ms = api.MultiScene(mctx, s, host=hs); t1 = time.perf_counter(); ms.close(); t2 = time.perf_counter()
    print(f"{n} devices: yk_multi_scene_create {1e3*(t1-t0):.1f} ms, destroy {1e3*(t2-t1):.1f} ms  (no_clone={os.environ.get('YK_MULTI_NO_CLONE')})", flush=True)
mctx.close()

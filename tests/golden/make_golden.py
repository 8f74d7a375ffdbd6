"""Regenerates tests/golden/oracle_renders.json from the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference (Rust) cannot be built or run in this image, so these are NOT reference outputs: they freeze the oracle's
own results on a fixed set of small scenes so that any later change to oracle/ that alters a single bit is caught."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from yuki_b200 import transforms as xf  # noqa: E402
import test_oracle_render as T  # noqa: E402

out = {}
for name, (scene, cam, film, smp, integ) in T.golden_cases(xf).items():
    img, ids, st = O.OracleScene(scene).render(cam, film, smp, integ, want_hit_ids=True)
    out[name] = T.digest(img, ids, st)
    print(name, out[name]["film_mean"], out[name]["ray_count"])
json.dump(out, open(T.GOLDEN, "w"), indent=1, sort_keys=True)

# sampler draws (Sampler::start_pixel_sample + get_1d / get_2d): a few (pixel, index) triples per sampler configuration
import numpy as np  # noqa: E402
import test_sampler_gpu as TS  # noqa: E402

draws = {}
for smp in TS.SAMPLERS:
    key = f"k{smp.kind}_{smp.nx}x{smp.ny}_j{int(smp.jitter)}_s{smp.seed}"
    rows = []
    for (px, py, idx) in TS.golden_triples(smp):
        rows.append({"pixel": [px, py], "index": idx,
                     "bits": [int(b) for b in O.sampler_draws(smp, px, py, idx, TS.GOLDEN_PATTERN).view(np.uint32)]})
    draws[key] = rows
json.dump({"pattern": TS.GOLDEN_PATTERN, "draws": draws}, open(TS.GOLDEN_DRAWS, "w"), indent=None, sort_keys=True)
print("sampler draws:", sum(len(v) for v in draws.values()), "rows")

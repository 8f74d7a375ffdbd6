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

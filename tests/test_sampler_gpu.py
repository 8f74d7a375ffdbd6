"""yk_sampler_draws: the samplers' trait surface (`start_pixel_sample`, `get_1d`, `get_2d`; sampling/mod.rs:46-57) evaluated
on the device and compared draw by draw with the oracle's samplers — extreme pixels and sample indices, the largest
sample counts, and hundreds of dimensions (beyond what a path reaches, so the per-dimension hash runs on the fly rather
than from the render's table)."""
import json
import os

import numpy as np
import pytest

from yuki_b200 import api, capi, desc as D

SAMPLERS = [D.SamplerType.uniform(1), D.SamplerType.uniform(8), D.SamplerType.uniform(65536), D.SamplerType.stratified(1, 1),
            D.SamplerType.stratified(4, 4), D.SamplerType.stratified(3, 2, jitter=False), D.SamplerType.stratified(7, 5),
            D.SamplerType.stratified(256, 256), D.SamplerType(D.SAMPLER_STRATIFIED, 32, 32, True, 12345)]


GOLDEN_DRAWS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler_draws.json")
GOLDEN_PATTERN = [2, 1, 2, 2, 1] * 3


def golden_triples(smp):
    spp = smp.samples_per_pixel()
    return [(0, 0, 0), (65535, 65535, spp - 1), (17, 4000, spp // 2), (1023, 7, min(spp - 1, 3))]


def _golden_key(smp):
    return f"k{smp.kind}_{smp.nx}x{smp.ny}_j{int(smp.jitter)}_s{smp.seed}"


def test_oracle_sampler_draws_match_the_committed_vectors(oracle):
    """tests/golden/sampler_draws.json (tests/golden/make_golden.py) freezes the oracle's draws: a regression pin that travels
    to the GPU box, not a reference output (the Rust samplers cannot run here)."""
    want = json.load(open(GOLDEN_DRAWS))
    assert want["pattern"] == GOLDEN_PATTERN
    for smp in SAMPLERS:
        rows = want["draws"][_golden_key(smp)]
        assert [tuple(r["pixel"]) + (r["index"],) for r in rows] == golden_triples(smp)
        for r in rows:
            got = oracle.sampler_draws(smp, r["pixel"][0], r["pixel"][1], r["index"], GOLDEN_PATTERN)
            assert [int(b) for b in got.view(np.uint32)] == r["bits"]


@pytest.mark.gpu
def test_gpu_sampler_draws_match_the_committed_vectors(gpu_ctx):
    want = json.load(open(GOLDEN_DRAWS))
    for smp in SAMPLERS:
        rows = want["draws"][_golden_key(smp)]
        got = api.sampler_draws(gpu_ctx, smp, [r["pixel"] + [r["index"]] for r in rows], GOLDEN_PATTERN)
        for g, r in zip(got, rows):
            assert [int(b) for b in g.view(np.uint32)] == r["bits"]


def test_sampler_draws_symbol_rejects_bad_arguments():
    L = capi.lib()
    assert L.yk_sampler_draws(None, None, None, 1, None, 0, None) != 0
    assert b"null argument" in L.yk_last_error()


@pytest.mark.gpu
@pytest.mark.parametrize("smp", SAMPLERS, ids=lambda s: f"k{s.kind}_{s.nx}x{s.ny}_{int(s.jitter)}")
def test_gpu_sampler_draws_equal_the_oracle(gpu_ctx, oracle, smp):
    spp = smp.samples_per_pixel()
    rng = np.random.default_rng(spp)
    n = 96
    px = rng.integers(0, 65536, n)
    py = rng.integers(0, 65536, n)
    idx = rng.integers(0, spp, n)
    px[:4], py[:4] = (0, 65535, 0, 65535), (0, 65535, 65535, 0)
    idx[:4] = (0, spp - 1, spp // 2, min(spp - 1, 1))
    pattern = [2, 1, 2, 2, 1, 1, 2] * 40          # 440 dimensions
    got = api.sampler_draws(gpu_ctx, smp, np.stack([px, py, idx], axis=1), pattern)
    assert got.shape == (n, sum(pattern))
    for i in range(n):
        want = oracle.sampler_draws(smp, int(px[i]), int(py[i]), int(idx[i]), pattern)
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), (i, px[i], py[i], idx[i])
    if smp.kind == D.SAMPLER_UNIFORM or smp.nx == smp.ny:   # (the reference's `y = stratum / ny` leaves [0, 1) when nx > ny, stratified.rs:128)
        assert (got >= 0).all() and (got < 1).all()
    # argument checks
    with pytest.raises(RuntimeError, match="fit u16"):
        api.sampler_draws(gpu_ctx, smp, [[70000, 0, 0]], [1])
    with pytest.raises(RuntimeError, match="pattern"):
        api.sampler_draws(gpu_ctx, smp, [[0, 0, 0]], [3])
    assert api.sampler_draws(gpu_ctx, smp, np.zeros((0, 3), np.uint32), [2, 1]).shape == (0, 3)

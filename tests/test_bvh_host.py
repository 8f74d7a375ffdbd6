"""BVH build: oracle invariants (the reference has no BVH tests), brute force vs tree traversal, and the product's host
builder against the oracle bit for bit for all three split methods."""
import numpy as np
import pytest

from yuki_b200 import api, desc as D, scenes

SPLITS = [D.SPLIT_SAH, D.SPLIT_MIDDLE, D.SPLIT_EQUAL_COUNTS]


def check_invariants(nodes, order, n_tris, max_in_node=None):
    assert sorted(order.tolist()) == list(range(n_tris))           # every primitive in exactly one leaf slot
    covered = np.zeros(n_tris, bool)
    stack = [(0, None)]
    seen = 0
    while stack:
        i, parent = stack.pop()
        seen += 1
        nd = nodes[i]
        if parent is not None:                                      # child bounds inside the parent's
            assert np.all(nd["p_min"] >= parent["p_min"]) and np.all(nd["p_max"] <= parent["p_max"])
        if nd["is_leaf"]:
            lo, cnt = int(nd["offset"]), int(nd["shape_count"])
            assert cnt >= 1 and not covered[lo:lo + cnt].any()
            covered[lo:lo + cnt] = True
        else:
            assert nd["split_axis"] in (0, 1, 2)
            assert int(nd["offset"]) > i + 1                        # pre-order: first child at i+1, second later
            stack.append((int(nd["offset"]), nd))
            stack.append((i + 1, nd))
    assert covered.all() and seen == len(nodes)


@pytest.mark.parametrize("split", SPLITS)
def test_oracle_bvh_invariants(oracle, xf, split):
    scene, _ = scenes.heightfield(xf, 40, 40, seed=2, split_method=split)
    osc = oracle.OracleScene(scene)
    check_invariants(osc.nodes(), osc.order(), scene.n_triangles())


@pytest.mark.parametrize("split", SPLITS)
@pytest.mark.parametrize("which", ["heightfield", "cornell", "room"])
def test_product_bvh_equals_oracle_bvh(oracle, xf, split, which):
    if which == "heightfield":
        scene, _ = scenes.heightfield(xf, 64, 48, seed=4, split_method=split)
    elif which == "cornell":
        scene, _ = scenes.cornell(xf, split_method=split)
    else:
        scene, _ = scenes.material_room(xf, split_method=split)
    host = api.HostScene(scene)
    osc = oracle.OracleScene(scene)
    pn, on = host.nodes(), osc.nodes()
    assert len(pn) == len(on)
    assert pn.tobytes() == on.tobytes()                              # bounds, offsets, axes, counts: all 32 bytes
    assert np.array_equal(host.order(), osc.order())
    check_invariants(pn, host.order(), scene.n_triangles())


def test_parallel_build_equals_serial_reference(oracle, xf):
    """Large enough (>= 2^16 primitives per subtree) to take the concurrent path of the product builder."""
    scene, _ = scenes.heightfield(xf, 300, 300, seed=9)
    host = api.HostScene(scene)
    osc = oracle.OracleScene(scene)
    assert host.nodes().tobytes() == osc.nodes().tobytes()
    assert np.array_equal(host.order(), osc.order())


@pytest.mark.parametrize("split,max_shapes", [(D.SPLIT_SAH, 1), (D.SPLIT_MIDDLE, 4)])
def test_wide_node_passes_equal_serial_reference(oracle, xf, split, max_shapes):
    """More than 2^19 primitives: the product builder runs the top nodes' bounds / bucket / partition passes on several
    threads. The parallel partition must reproduce itertools::partition's permutation (leaf order, multi-shape leaves)."""
    scene, _ = scenes.heightfield(xf, 560, 560, seed=11, split_method=split, max_shapes_in_node=max_shapes)
    assert scene.n_triangles() > (1 << 19)
    host = api.HostScene(scene)
    osc = oracle.OracleScene(scene)
    assert host.nodes().tobytes() == osc.nodes().tobytes()
    assert np.array_equal(host.order(), osc.order())


def test_max_shapes_in_node(oracle, xf):
    scene, _ = scenes.heightfield(xf, 32, 32, seed=5, split_method=D.SPLIT_EQUAL_COUNTS)
    scene.max_shapes_in_node = 4
    host = api.HostScene(scene)
    nodes = host.nodes()
    assert nodes[nodes["is_leaf"] == 1]["shape_count"].max() <= 4
    assert nodes.tobytes() == oracle.OracleScene(scene).nodes().tobytes()


def test_yk_bvh_build_entry_point(xf):
    scene, _ = scenes.heightfield(xf, 20, 20, seed=6)
    host = api.HostScene(scene)
    tv = host.tri_vertices()
    inv = np.argsort(host.order())
    nodes, order = api.bvh_build(tv[inv])                              # original (pre-BVH) triangle order
    assert nodes.tobytes() == host.nodes().tobytes()
    assert np.array_equal(order, host.order())


def test_degenerate_inputs(xf):
    from yuki_b200.capi import YukiGpuError
    with pytest.raises(YukiGpuError):
        api.bvh_build(np.zeros((0, 3, 3), np.float32))
    one = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], np.float32)
    nodes, order = api.bvh_build(one)
    assert len(nodes) == 1 and nodes[0]["is_leaf"] == 1 and order.tolist() == [0]
    same = np.repeat(one, 5, axis=0)                                     # identical centroids -> one leaf (bvh.rs:343)
    nodes, order = api.bvh_build(same)
    assert len(nodes) == 1 and nodes[0]["shape_count"] == 5


@pytest.mark.parametrize("split", SPLITS)
def test_traversal_equals_brute_force(oracle, xf, split):
    scene, cam = scenes.heightfield(xf, 24, 24, seed=8, split_method=split)
    osc = oracle.OracleScene(scene)
    rng = np.random.default_rng(3)
    n = 3000
    o = rng.uniform(-2.5, 2.5, (n, 3)).astype(np.float32)
    target = rng.uniform(-0.5, 0.5, (n, 3)).astype(np.float32) * np.array([1, 0.05, 1], np.float32)
    d = target - o
    t_bvh, id_bvh, counts = osc.trace(o, d)
    t_bf, id_bf, _ = osc.trace(o, d, brute_force=True)
    assert (id_bvh >= 0).sum() > n // 4
    assert np.array_equal(t_bvh.view(np.uint32), t_bf.view(np.uint32))
    # equal-t ties on shared edges may resolve to a different (equally close) triangle; everything else must agree
    differ = id_bvh != id_bf
    assert differ.sum() <= 0.01 * n
    assert np.all(counts[:, 0] >= counts[:, 1]) and np.all(counts[:, 0] >= 1)
    tm = np.full(n, 0.9999, np.float32)
    assert np.array_equal(osc.occluded(o, d, tm), osc.occluded(o, d, tm, brute_force=True))

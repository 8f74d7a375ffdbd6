"""PLY loader (csrc/host_ply.cpp) against the acceptance rules of yuki/src/scene/ply.rs: float vertex properties, int/uint
index lists under either name, fan triangulation, unknown content skipped, the reference's panics reported as errors."""
import numpy as np
import pytest

from yuki_b200 import api, capi, desc as D, scenes, transforms as xf


def _mesh(nx=5, nz=4):
    pts, idx = scenes.grid_mesh(nx, nz, seed=3)
    rng = np.random.default_rng(0)
    nrm = rng.normal(size=pts.shape).astype(np.float32)
    uvs = rng.random((len(pts), 2)).astype(np.float32)
    return pts, idx, nrm, uvs


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_round_trip_all_formats(tmp_path, fmt):
    pts, idx, nrm, uvs = _mesh()
    f = tmp_path / "m.ply"
    scenes.write_ply(f, pts, idx, nrm, uvs, fmt=fmt)
    p2, i2, n2, u2 = api.load_ply(f)
    assert np.array_equal(p2.view(np.uint32), pts.view(np.uint32))
    assert np.array_equal(i2, idx)
    assert np.array_equal(n2.view(np.uint32), nrm.view(np.uint32))
    assert np.array_equal(u2.view(np.uint32), uvs.view(np.uint32))


def test_positions_only_and_fan_triangulation(tmp_path):
    pts, idx, _, _ = _mesh()
    f = tmp_path / "q.ply"
    scenes.write_ply(f, pts, idx, fmt="binary_little_endian", quads=True)   # (a b c d) -> (a b c)(a c d), ply.rs:81-92
    p2, i2, n2, u2 = api.load_ply(f)
    assert n2 is None and u2 is None
    assert np.array_equal(i2, idx)
    (tmp_path / "pent.ply").write_text("ply\nformat ascii 1.0\nelement vertex 5\nproperty float x\nproperty float y\nproperty float z\n"
                                       "element face 1\nproperty list uchar uint vertex_index\nend_header\n"
                                       "0 0 0\n1 0 0\n2 1 0\n1 2 0\n0 1 0\n5 0 1 2 3 4\n")
    _, i3, _, _ = api.load_ply(tmp_path / "pent.ply")
    assert i3.tolist() == [0, 1, 2, 0, 2, 3, 0, 3, 4]


def test_unknown_elements_and_properties_are_skipped(tmp_path):
    body = ("ply\nformat ascii 1.0\ncomment made by hand\nobj_info x\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
            "property uchar red\nproperty double weight\nelement edge 2\nproperty int a\nproperty int b\n"
            "element face 1\nproperty uchar flag\nproperty list uchar int vertex_indices\nend_header\n"
            "0 0 0 255 0.5\n1 0 0 0 0.25\n0 1 0 7 1e-3\n0 1\n1 2\n9 3 0 1 2\n")
    f = tmp_path / "extra.ply"
    f.write_text(body)
    p, i, n, u = api.load_ply(f)
    assert p.tolist() == [[0, 0, 0], [1, 0, 0], [0, 1, 0]] and i.tolist() == [0, 1, 2] and n is None and u is None


def test_non_float_coordinates_are_ignored_like_the_reference(tmp_path):
    # ply.rs:253 only matches Property::Float: a `double` x stays at Point3::zeros()
    f = tmp_path / "dbl.ply"
    f.write_text("ply\nformat ascii 1.0\nelement vertex 3\nproperty double x\nproperty float y\nproperty float z\n"
                 "element face 1\nproperty list uchar int vertex_indices\nend_header\n5 0 0\n6 1 0\n7 0 1\n3 0 1 2\n")
    p, _, _, _ = api.load_ply(f)
    assert p.tolist() == [[0, 0, 0], [0, 1, 0], [0, 0, 1]]


@pytest.mark.parametrize("body,needle", [
    ("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nelement face 0\nproperty list uchar int vertex_indices\nend_header\n0 0\n",
     "missing property 'z'"),
    ("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nend_header\n0 0 0\n", "Missing element 'face'"),
    ("ply\nformat ascii 1.0\nelement face 0\nproperty list uchar int vertex_indices\nend_header\n", "Missing element 'vertex'"),
    ("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nelement face 1\nproperty list uchar int idx\nend_header\n0 0 0\n3 0 0 0\n",
     "vertex_index"),
    ("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n"
     "0 0 0\n1 0 0\n0 1 0\n3 0 -1 2\n", "Negative PLY index"),
    ("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nproperty float ny\nproperty float nx\nproperty float nz\n"
     "element face 0\nproperty list uchar int vertex_indices\nend_header\n0 0 0 0 1 0\n", "panics"),
    ("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\nelement face 1\nproperty list uchar int vertex_indices\nend_header\n"
     "0 0 0\n1 0 0\n", "truncated"),
    # element counts the payload cannot hold: a parse error, never a length_error / bad_alloc unwinding through the C ABI
    ("ply\nformat ascii 1.0\nelement vertex 9000000000000000000\nproperty float x\nproperty float y\nproperty float z\nelement face 1\n"
     "property list uchar int vertex_indices\nend_header\n0 0 0\n", "truncated"),
    ("ply\nformat binary_little_endian 1.0\nelement vertex 18446744073709551615\nproperty float x\nproperty float y\nproperty float z\nelement face 0\n"
     "property list uchar int vertex_indices\nend_header\n", "truncated"),
])
def test_rejected_files(tmp_path, body, needle):
    f = tmp_path / "bad.ply"
    f.write_text(body)
    with pytest.raises(capi.YukiGpuError) as e:
        api.load_ply(f)
    assert needle in str(e.value)


def test_missing_file(tmp_path):
    with pytest.raises(capi.YukiGpuError) as e:
        api.load_ply(tmp_path / "nope.ply")
    assert "Could not open" in str(e.value)


def test_scene_ply_matches_the_programmatic_heightfield(tmp_path):
    """Scene::ply over the file == the synthetic config-3 scene built in memory: same BVH, same flattened triangles."""
    pts, idx = scenes.grid_mesh(24, 20, seed=1)
    f = tmp_path / "hf.ply"
    scenes.write_ply(f, pts, idx, fmt="binary_big_endian")
    s_file, cam_file = scenes.ply(xf, f)
    s_mem, cam_mem = scenes.heightfield(xf, 24, 20, seed=1)
    a, b = api.HostScene(s_file), api.HostScene(s_mem)
    assert a.n_nodes == b.n_nodes and a.n_tris == b.n_tris
    assert a.nodes().tobytes() == b.nodes().tobytes()
    assert np.array_equal(a.order(), b.order())
    assert np.array_equal(a.tri_vertices().view(np.uint32), b.tri_vertices().view(np.uint32))
    assert cam_file == cam_mem


@pytest.mark.gpu
def test_ply_scene_bvh_counters_match_oracle(tmp_path):
    from oracle import oracle as O
    pts, idx = scenes.grid_mesh(40, 32, seed=2)
    f = tmp_path / "hf.ply"
    scenes.write_ply(f, pts, idx, fmt="binary_little_endian", quads=True)
    scene, cam = scenes.ply(xf, f, split_method=D.SPLIT_MIDDLE)
    film = D.FilmSettings((160, 120), 16)
    smp, integ = D.SamplerType.uniform(1), D.IntegratorType.bvh_intersections()
    ctx = api.Context(0)
    dev = api.Scene(ctx, scene)
    r = api.Renderer(ctx).render(dev, cam, film, smp, integ, want_hit_ids=True)
    o_img, o_ids, _ = O.OracleScene(scene).render(cam, film, smp, integ, want_hit_ids=True)
    assert np.array_equal(r.hit_ids, o_ids)
    assert np.array_equal(r.film.view(np.uint32), o_img.view(np.uint32))
    dev.close(); ctx.close()

"""Math KATs taken from the reference's own unit tests (tests/src/{matrix,transform,vector,normal,point,ray,bounds}.rs —
every test there that touches a helper the hot path uses), replayed against
both the oracle and the product's host helpers, plus product == oracle bit-for-bit on random inputs."""
import numpy as np
import pytest


@pytest.fixture(params=["product", "oracle"])
def backend(request, oracle, xf):
    return xf if request.param == "product" else oracle.transforms


T_ROWS = [16.0, 11.0, 6.0, 13.0, 12.0, 15.0, 10.0, 9.0, 8.0, 7.0, 14.0, 5.0, 4.0, 3.0, 2.0, 1.0]
TP_ROWS = [16.0, 11.0, 6.0, 13.0, 12.0, 15.0, 10.0, 9.0, 8.0, 7.0, 14.0, 5.0, 0.0, 0.0, 0.0, 1.0]


def test_transform_mul_kats(backend):
    """tests/src/transform.rs:102-128"""
    t, tp = backend.new(T_ROWS), backend.new(TP_ROWS)
    assert np.array_equal(backend.vec(t, (17.0, 18.0, 19.0)), np.array([584.0, 664.0, 528.0], np.float32))
    expect = np.array([597.0, 673.0, 533.0], np.float32) / np.float32(161.0)
    assert np.array_equal(backend.point(t, (17.0, 18.0, 19.0)), expect)           # homogeneous divide
    assert np.array_equal(backend.point(tp, (17.0, 18.0, 19.0)), np.array([597.0, 673.0, 533.0], np.float32))
    assert np.array_equal(backend.normal(t, (17.0, 18.0, 19.0)), np.array([-1.0694447, 0.5972222, 0.5972223], np.float32))
    assert np.array_equal(backend.vec(t, (20.0, 21.0, 22.0)), np.array([683.0, 775.0, 615.0], np.float32))  # ray direction


def test_matrix_product_kat(backend):
    """tests/src/matrix.rs:180-198 (through Transform * Transform, transform.rs:203-212)"""
    m = np.arange(1, 17, dtype=np.float32)
    ident = backend.identity()
    a = type(ident)(m.copy(), ident.m_inv.copy())
    prod = backend.mul(a, a).m.reshape(4, 4)
    assert np.array_equal(prod, np.array([[90, 100, 110, 120], [202, 228, 254, 280], [314, 356, 398, 440], [426, 484, 542, 600]], np.float32))


def test_matrix_inverse_kat(backend):
    """tests/src/matrix.rs:160-176: A * A^-1 = I and (A^-1)^-1 = A to 1e-5"""
    m = [9.2, 8.1, 8.0, -2.1, -8.3, 16.0, 3.0, 8.0, 0.5, 9.3, -4.0, 7.1, 3.0, -8.0, 2.0, 10.0]
    t = backend.new(m)
    a, ai = t.m.reshape(4, 4).astype(np.float64), t.m_inv.reshape(4, 4).astype(np.float64)
    assert np.allclose(a @ ai, np.eye(4), atol=1e-5)
    back = backend.new(t.m_inv)
    assert np.allclose(back.m_inv.reshape(4, 4), np.array(m, np.float32).reshape(4, 4), atol=1e-5)


def test_singular_matrix_is_an_error(oracle, xf):
    """matrix.rs:170 panics; the C ABI returns YK_ERR_SINGULAR instead."""
    from yuki_b200.capi import YukiGpuError
    with pytest.raises(YukiGpuError) as e:
        xf.new([1, 2, 3, 4, 2, 4, 6, 8, 0, 0, 1, 0, 0, 0, 0, 1])
    assert e.value.code == -6 and "singular" in str(e.value)
    with pytest.raises(ValueError):
        oracle.transforms.new([1, 2, 3, 4, 2, 4, 6, 8, 0, 0, 1, 0, 0, 0, 0, 1])


def test_look_at_kat(backend):
    """tests/src/transform.rs:254-278 (the reference checks in f64 to 1e-15; f32 here)"""
    m = np.array([[0.825307261249832, -0.322265731783557, 0.463694643754174, 1.0],
                  [0.0, 0.821157874256179, 0.570701100005137, 2.0],
                  [-0.564683915591990, -0.471003761837506, 0.677707556256101, 3.0],
                  [0.0, 0.0, 0.0, 1.0]])
    t = backend.look_at((1.0, 2.0, 3.0), (40.0, 50.0, 60.0), (0.0, 1.0, 0.0))
    assert np.allclose(t.m_inv.reshape(4, 4), m, atol=2e-7)


def test_rotation_kat(backend):
    """tests/src/transform.rs:225-251"""
    rm = np.array([[0.333333333333333, -0.244016935856292, 0.910683602522959, 0.0],
                   [0.910683602522959, 0.333333333333333, -0.244016935856292, 0.0],
                   [-0.244016935856292, 0.910683602522959, 0.333333333333333, 0.0],
                   [0.0, 0.0, 0.0, 1.0]])
    t = backend.rotation(np.float32(np.pi / 2), (1.0, 1.0, 1.0))
    assert np.allclose(t.m.reshape(4, 4), rm, atol=2e-7)
    assert np.allclose(t.m_inv.reshape(4, 4), rm.T, atol=2e-7)


def test_translation_and_scale(backend):
    t = backend.translation((1.0, 2.0, 3.0))
    assert np.array_equal(backend.point(t, (1.0, 1.0, 1.0)), np.array([2.0, 3.0, 4.0], np.float32))
    assert np.array_equal(backend.vec(t, (1.0, 1.0, 1.0)), np.array([1.0, 1.0, 1.0], np.float32))
    s = backend.scale(2.0, 3.0, 4.0)
    assert np.array_equal(backend.point(s, (1.0, 1.0, 1.0)), np.array([2.0, 3.0, 4.0], np.float32))
    assert np.array_equal(backend.normal(s, (1.0, 1.0, 1.0)), np.array([0.5, np.float32(1.0) / np.float32(3.0), 0.25], np.float32))


def test_cross_kat(oracle):
    """tests/src/vector.rs:99-104"""
    assert np.array_equal(oracle.cross((2.0, 3.0, 4.0), (5.0, 6.0, -7.0)), np.array([-45.0, 34.0, -3.0], np.float32))


def test_product_and_oracle_transforms_are_bit_identical(oracle, xf):
    rng = np.random.default_rng(7)
    ox = oracle.transforms
    for _ in range(50):
        d = rng.normal(size=3).astype(np.float32)
        s = (rng.uniform(0.1, 4.0, size=3)).astype(np.float32)
        axis = rng.normal(size=3).astype(np.float32)
        th = np.float32(rng.uniform(-6.0, 6.0))
        pa = xf.mul(xf.translation(d), xf.mul(xf.rotation(th, axis), xf.scale(*s)))
        oa = ox.mul(ox.translation(d), ox.mul(ox.rotation(th, axis), ox.scale(*s)))
        assert np.array_equal(pa.m.view(np.uint32), oa.m.view(np.uint32))
        assert np.array_equal(pa.m_inv.view(np.uint32), oa.m_inv.view(np.uint32))
        pn, on = xf.new(pa.m), ox.new(oa.m)      # Gauss-Jordan inverse
        assert np.array_equal(pn.m_inv.view(np.uint32), on.m_inv.view(np.uint32))
        v = rng.normal(size=3).astype(np.float32)
        for f in ("point", "vec", "normal"):
            assert np.array_equal(getattr(xf, f)(pa, v).view(np.uint32), getattr(ox, f)(oa, v).view(np.uint32))
        la_p = xf.look_at(d, d + v, (0.0, 1.0, 0.0))
        la_o = ox.look_at(d, d + v, (0.0, 1.0, 0.0))
        assert np.array_equal(la_p.m.view(np.uint32), la_o.m.view(np.uint32))


# ---- the rest of the reference's unit tests that touch helpers the hot path is built from ------------------------
# (vector / normal / point / ray / bounds / transform; the Vec2 / Vec4 / integer variants exercise the same derive-generated
# bodies and have no counterpart here.) Opcodes: oracle/yk_oracle.cpp:yko_math_kat.
def _k(oracle, op, values, n_out=1):
    out = oracle.math_kat(op, [float(v) for v in values], n_out)
    return float(out[0]) if n_out == 1 else [float(v) for v in out]


def test_vector_kats(oracle):
    """tests/src/vector.rs:81-211, normal.rs:42-73"""
    f32 = np.float32
    assert _k(oracle, 0, (2, 3, 4, 5, 6, 7)) == 2 * 5 + 3 * 6 + 4 * 7                   # Vec3::dot
    assert _k(oracle, 1, (2, 3, 4, 5, 6, 7)) == 2.0 * 5.0 + 3.0 * 6.0 + 4.0 * 7.0         # dot_n / Normal::dot / dot_v
    assert _k(oracle, 2, (2, 3, 4)) == 2 * 2 + 3 * 3 + 4 * 4                              # len_sqr
    assert abs(_k(oracle, 3, (2, 3, 4)) - float(np.sqrt(f32(29.0)))) <= 1.1920929e-07     # len (assert_abs_diff_eq, f32 epsilon)
    n = _k(oracle, 4, (1, 1, 1), 3)
    assert abs(_k(oracle, 3, n) - 1.0) <= 1.1920929e-07                                   # normalized().len() == 1
    assert _k(oracle, 5, (0, 2, 4, 3, 1, 5), 3) == [0, 1, 4] == _k(oracle, 5, (3, 1, 5, 0, 2, 4), 3)   # min, commutes
    assert _k(oracle, 6, (0, 2, 4, 3, 1, 5), 3) == [3, 2, 5] == _k(oracle, 6, (3, 1, 5, 0, 2, 4), 3)   # max
    assert _k(oracle, 7, (0, 1, 2)) == 0.0 and _k(oracle, 8, (0, 1, 2)) == 2.0            # min_comp / max_comp
    assert _k(oracle, 9, (0, 1, 2)) == 2                                                  # max_dimension
    assert _k(oracle, 10, (3, 4, 5, 1, 2, 0), 3) == [4, 5, 3]                             # permuted(1, 2, 0)


def test_point_and_ray_kats(oracle):
    """tests/src/point.rs:42-62, ray.rs:49-56"""
    p0 = np.array([1, 2, 3], np.float32)
    p1 = p0 + np.array(_k(oracle, 4, (4, 5, 6), 3), np.float32) * np.float32(3.0)
    assert abs(_k(oracle, 20, (*p0, *p1)) - 3.0) <= 1.1920929e-07 * 4                     # dist (the sum is rounded twice more)
    assert abs(_k(oracle, 21, (*p0, *p1)) - 9.0) <= 1e-5                                  # dist_sqr
    assert _k(oracle, 19, (1, 2, 3, 4, 5, 6, 1.0, 1.0), 3) == [5, 7, 9]                    # r.point(1) == o + d
    assert _k(oracle, 19, (1, 2, 3, 4, 5, 6, 1.0, 2.0), 3) == [9, 12, 15]                  # r.point(2) == o + d * 2


def test_bounds_kats(oracle):
    """tests/src/bounds.rs:25-35, 65-119, 272-381 (the Bounds3<f32> cases; the Bounds2 loops run on xy with z = 0)"""
    fmax = float(np.finfo(np.float32).max)
    assert _k(oracle, 17, (), 6) == [fmax] * 3 + [-fmax] * 3                              # default(): p_min = MAX, p_max = MIN
    bb = (0, 0, 0, 2, 2, 2)
    assert _k(oracle, 11, (*bb, 1, 1, 1), 6) == list(bb)                                  # union_p inside
    assert _k(oracle, 11, (0, 0, 0, 2, 2, 0, 3, 1, 0), 6) == [0, 0, 0, 3, 2, 0]
    assert _k(oracle, 11, (0, 0, 0, 2, 2, 0, 3, 4, 0), 6) == [0, 0, 0, 3, 4, 0]
    assert _k(oracle, 11, (0, 0, 0, 2, 2, 0, -3, -4, 0), 6) == [-3, -4, 0, 2, 2, 0]
    pts = [(0, 0, 0), (1, 1, 0), (2, 2, 0), (3, 3, 0)]
    import itertools
    for l, k, j, i in itertools.permutations(range(4)):                                   # union_b of any two disjoint pairs
        assert _k(oracle, 12, (*pts[l], *pts[k], *pts[j], *pts[i]), 6) == [0, 0, 0, 3, 3, 0]
    assert _k(oracle, 12, (*pts[1], *pts[2], *pts[0], *pts[3]), 6) == [0, 0, 0, 3, 3, 0]
    assert _k(oracle, 12, (*pts[0], *pts[3], *pts[1], *pts[2]), 6) == [0, 0, 0, 3, 3, 0]
    assert _k(oracle, 12, (*bb, 1, 1, 1, 1, 1, 1), 6) == list(bb) == _k(oracle, 12, (1, 1, 1, 1, 1, 1, *bb), 6)
    assert _k(oracle, 13, (1, 2, 3, 5, 8, 11), 3) == [4, 6, 8]                            # diagonal
    assert _k(oracle, 14, (1, 2, 3, 4, 5, 6, 2.5, 3.5, 4.5), 3) == [0.5, 0.5, 0.5]        # offset
    p0 = (1.0, 2.0, 0.0)
    for axis in range(2):                                                                 # the Bounds2 offset walk, on xy
        for step, want in ((0.0, 0.0), (1.0, 0.5), (2.0, 1.0), (4.0, 2.0), (-2.0, -1.0)):
            pp = list(p0)
            pp[axis] += step
            got = _k(oracle, 14, (1, 2, 0, 3, 4, 1, *pp), 3)
            assert got[axis] == want and got[1 - axis] == 0.0
    assert _k(oracle, 15, (1, 2, 3, 3, 5, 7)) == 52 == _k(oracle, 15, (-1, -2, -3, -3, -5, -7))   # surface_area
    assert _k(oracle, 16, (1, 2, 3, 4, 5, 7)) == 2                                        # maximum_extent
    assert _k(oracle, 16, (1, 2, 3, 4, 6, 5)) == 1
    assert _k(oracle, 16, (1, 2, 3, 7, 5, 6)) == 0


def test_swaps_handedness_kats(oracle):
    """tests/src/transform.rs:80-100"""
    ident = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
    swap_yz = [1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1]
    flip_z = [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1]
    assert _k(oracle, 18, ident) == 0.0
    assert _k(oracle, 18, swap_yz) == 1.0
    assert _k(oracle, 18, flip_z) == 1.0

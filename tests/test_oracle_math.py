"""Math KATs taken from the reference's own unit tests (tests/src/{matrix,transform,vector,bounds}.rs), replayed against
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

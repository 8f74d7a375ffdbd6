"""Pins the sampler arithmetic that lives in third-party crates (rand_pcg 0.3, rand 0.8, std SipHash-1-3) against the
published KATs and an independent pure-Python restatement, then checks yuki's sampler properties."""
import numpy as np

from yuki_b200 import desc as D

M64 = (1 << 64) - 1


def py_siphash13(msg: bytes) -> int:
    def rotl(x, b):
        return ((x << b) | (x >> (64 - b))) & M64
    v0, v1, v2, v3 = 0x736f6d6570736575, 0x646f72616e646f6d, 0x6c7967656e657261, 0x7465646279746573

    def rnd():
        nonlocal v0, v1, v2, v3
        v0 = (v0 + v1) & M64; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32)
        v2 = (v2 + v3) & M64; v3 = rotl(v3, 16); v3 ^= v2
        v0 = (v0 + v3) & M64; v3 = rotl(v3, 21); v3 ^= v0
        v2 = (v2 + v1) & M64; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32)
    n = len(msg)
    for i in range(0, n - n % 8, 8):
        m = int.from_bytes(msg[i:i + 8], "little")
        v3 ^= m; rnd(); v0 ^= m
    b = (n << 56) & M64 | int.from_bytes(msg[n - n % 8:], "little")
    v3 ^= b; rnd(); v0 ^= b
    v2 ^= 0xff
    rnd(); rnd(); rnd()
    return v0 ^ v1 ^ v2 ^ v3


class PyPcg32:
    MULT = 6364136223846793005

    def __init__(self, state, stream):
        self.inc = ((stream << 1) | 1) & M64
        self.state = (state + self.inc) & M64
        self.state = (self.state * self.MULT + self.inc) & M64

    def next_u32(self):
        old = self.state
        self.state = (old * self.MULT + self.inc) & M64
        xsh = (((old >> 18) ^ old) >> 27) & 0xffffffff
        rot = old >> 59
        return ((xsh >> rot) | (xsh << ((32 - rot) & 31))) & 0xffffffff

    def advance(self, delta):
        am, ap, cm, cp = 1, 0, self.MULT, self.inc
        while delta:
            if delta & 1:
                am = (am * cm) & M64
                ap = (ap * cm + cp) & M64
            cp = ((cm + 1) * cp) & M64
            cm = (cm * cm) & M64
            delta >>= 1
        self.state = (am * self.state + ap) & M64


def test_siphash13_kats(oracle):
    # Rust: DefaultHasher::new().finish() == 15130871412783076140 (SipHash-1-3, k = 0, empty input)
    assert oracle.siphash13(b"") == 15130871412783076140
    rng = np.random.default_rng(1)
    for n in list(range(0, 33)) + [100, 1000]:
        msg = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.siphash13(msg) == py_siphash13(msg)


def test_pcg32_kat_and_advance(oracle):
    # pcg32 demo / rand_pcg test vector: seed 42, stream 54
    assert [hex(x) for x in oracle.pcg32_sequence(42, 54, 0, 6)] == ["0xa15c02b7", "0x7b47f409", "0xba1d3330", "0x83d2f293", "0xbfa4784b", "0xcbed606e"]
    seq = oracle.pcg32_sequence(42, 54, 0, 1000)
    for adv in (1, 2, 7, 65536 % 1000, 999):
        assert np.array_equal(oracle.pcg32_sequence(42, 54, adv, 1000 - adv), seq[adv:])
    big = 1023 * 65536 + 5
    p = PyPcg32(0x73B9642E74AC471C, 0xDEADBEEFCAFEF00D)
    p.advance(big)
    assert list(oracle.pcg32_sequence(0x73B9642E74AC471C, 0xDEADBEEFCAFEF00D, big, 8)) == [p.next_u32() for _ in range(8)]


def test_uniform_sampler_matches_python_restatement(oracle):
    """uniform.rs:72-94: stream = SipHash(pixel), advance(index * 65536 + dim), f32 = (u32 >> 8) * 2^-24"""
    s = D.SamplerType.uniform(16, seed=0x1234567890ABCDEF)
    for (px, py, idx) in [(0, 0, 0), (17, 3, 5), (1023, 767, 15)]:
        h = py_siphash13(px.to_bytes(2, "little") + py.to_bytes(2, "little"))
        p = PyPcg32(s.seed, h)
        p.advance(idx * 65536)
        expect = np.array([np.float32(p.next_u32() >> 8) * np.float32(2.0 ** -24) for _ in range(5)], np.float32)
        got = oracle.sampler_draws(s, px, py, idx, [2, 1, 2])
        assert np.array_equal(got, expect)


def py_permutation_element(i, l, p):
    M = 0xffffffff
    w = l - 1
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16
    while True:
        i ^= p; i = (i * 0xe170893d) & M
        i ^= p >> 16
        i ^= (i & w) >> 4
        i ^= p >> 8; i = (i * 0x0929eb3f) & M
        i ^= p >> 23
        i ^= (i & w) >> 1; i = (i * (1 | p >> 27)) & M
        i = (i * 0x6935fa69) & M
        i ^= (i & w) >> 11; i = (i * 0x74dcb303) & M
        i ^= (i & w) >> 2; i = (i * 0x9e501cc3) & M
        i ^= (i & w) >> 2; i = (i * 0xc860a3df) & M
        i &= w
        i ^= i >> 5
        if i < l:
            break
    return ((i + p) & M) % l   # wrapping_add (stratified.rs:177)


def test_permutation_element(oracle):
    """stratified.rs:147-178. A permutation of 0..l whenever i + p does not wrap (or l is a power of two); the u32
    wrapping add of the reference is reproduced either way."""
    for l in (1, 2, 3, 16, 17, 64, 1000, 4096):
        for p in (0, 1, 0xdeadbeef, 0xffffffff, 0x7fffffff):
            out = [oracle.permutation_element(i, l, p) for i in range(l)]
            assert out == [py_permutation_element(i, l, p) for i in range(l)]
            if p + l < (1 << 32) or (l & (l - 1)) == 0:
                assert sorted(out) == list(range(l))


def test_stratified_covers_every_stratum_once(oracle):
    """stratified.rs:121-143. Note y = stratum / pixel_samples.y (reference quirk): with nx == ny every (x, y) cell is hit
    exactly once per dimension pair over the spp samples."""
    nx = ny = 4
    s = D.SamplerType.stratified(nx, ny, jitter=True)
    cells = set()
    for idx in range(nx * ny):
        u = oracle.sampler_draws(s, 10, 20, idx, [2])
        assert 0.0 <= u[0] < 1.0 and 0.0 <= u[1] < 1.0
        cells.add((int(u[0] * nx), int(u[1] * ny)))
    assert len(cells) == nx * ny
    nj = D.SamplerType.stratified(nx, ny, jitter=False)
    u = oracle.sampler_draws(nj, 10, 20, 3, [2, 1])
    assert np.all((u[:2] * 4) % 1 == 0.5) and (u[2] * 16) % 1 == 0.5


def test_stratified_non_square_quirk(oracle):
    """With nx != ny the reference's y index can reach beyond ny-1 (stratum / ny with stratum < nx*ny): reproduce, not fix."""
    s = D.SamplerType.stratified(6, 2, jitter=False)
    ys = [oracle.sampler_draws(s, 1, 1, idx, [2])[1] for idx in range(12)]
    assert max(ys) > 1.0

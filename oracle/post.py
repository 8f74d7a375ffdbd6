"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy float32 restatement of the display passes that follow the hot
path in yuki — the filmic tone map and the heat map fragment shaders (yuki/src/app/renderpasses/tonemap.rs:318-432) and
`find_min_max` (:447-472) — plus a minimal OpenEXR reader for checking the writer that stands in for app/util.rs:89-110.
Parity unpinned: the reference has no tests or golden images for these passes; GLSL leaves float contraction open, so
the GPU kernels are compared to this restatement with a small tolerance, not bit for bit."""
import struct

import numpy as np

F = np.float32


def _rrt_odt(v):  # tonemap.rs:353-358
    a = v * (v + F(0.0245786)) - F(0.000090537)
    b = v * (F(0.983729) * v + F(0.4329510)) + F(0.238081)
    return a / b


def tonemap_filmic(film, exposure=1.0, tile_samples=None, tile_dim=16):
    film = np.asarray(film, np.float32)
    h, w, _ = film.shape
    c = film.copy()
    if tile_samples is not None:  # tonemap.rs:383-392
        ts = np.asarray(tile_samples, np.float32).reshape(-1)
        ys, xs = np.mgrid[0:h, 0:w]
        flat = (ys // tile_dim) * (w // tile_dim) + xs // tile_dim
        n = np.where(flat < len(ts), ts[np.minimum(flat, len(ts) - 1)], F(0.0)).astype(np.float32)
        c = np.where((n > 0)[..., None], c / np.where(n > 0, n, F(1.0))[..., None], c).astype(np.float32)
    c = c * F(exposure)
    r, g, b = c[..., 0], c[..., 1], c[..., 2]
    ir = F(0.59719) * r + F(0.35458) * g + F(0.04823) * b
    ig = F(0.07600) * r + F(0.90834) * g + F(0.01566) * b
    ib = F(0.02840) * r + F(0.13383) * g + F(0.83777) * b
    fr, fg, fb = _rrt_odt(ir), _rrt_odt(ig), _rrt_odt(ib)
    out = np.stack([F(1.60475) * fr + F(-0.53108) * fg + F(-0.07367) * fb,
                    F(-0.10208) * fr + F(1.10813) * fg + F(-0.00605) * fb,
                    F(-0.00327) * fr + F(-0.07276) * fg + F(1.07602) * fb], axis=-1)
    return np.clip(out, F(0.0), F(1.0)).astype(np.float32)


def linear_to_srgb_shader(c):
    """The output pass's transfer function when the backbuffer is not sRGB (`gamma_before_output`, scale_output.rs:150-169):
    `x <= 0.0031308 ? 12.92 x : 1.055 pow(x, 1 / 2.2) - 0.055` — note the 2.2 where sRGB proper has 2.4."""
    c = np.asarray(c, np.float32)
    return np.where(c <= F(0.0031308), F(12.92) * c, F(1.055) * np.power(c, F(1.0 / 2.2)) - F(0.055)).astype(np.float32)


def _luminance(p):
    return F(0.2126) * p[..., 0] + F(0.7152) * p[..., 1] + F(0.0722) * p[..., 2]


def find_min_max(film, channel):  # tonemap.rs:447-472 (channel 0..2 = that component, 3 = luminance)
    film = np.asarray(film, np.float32)
    v = film[..., channel] if channel < 3 else _luminance(film)
    return float(v.min()), float(v.max())


def heatmap(film, channel, min_val, max_val):  # tonemap.rs:401-432
    film = np.asarray(film, np.float32)
    value = film[..., channel] if 0 < channel < 3 else _luminance(film)  # channel 0 reads luminance, as in the shader
    s = (value - F(min_val)) / (F(max_val) - F(min_val))
    t0 = np.clip(s * F(2.0), F(0.0), F(1.0))
    t1 = np.clip(s * F(2.0) - F(1.0), F(0.0), F(1.0))
    low_mid = np.stack([np.zeros_like(t0), t0, F(1.0) - t0], axis=-1)
    high = np.array([1.0, 0.0, 0.0], np.float32)
    return (low_mid * (F(1.0) - t1)[..., None] + high * t1[..., None]).astype(np.float32)


def read_exr_rgb(path):
    """Reads an uncompressed scanline OpenEXR with float channels; returns (H, W, 3) f32 in R, G, B order."""
    data = open(path, "rb").read()
    magic, version = struct.unpack_from("<II", data, 0)
    assert magic == 20000630 and (version & 0xff) == 2 and not (version & 0x200), "not a single-part scanline EXR"
    pos = 8
    attrs = {}
    while data[pos] != 0:
        end = data.index(b"\0", pos); name = data[pos:end].decode(); pos = end + 1
        end = data.index(b"\0", pos); typ = data[pos:end].decode(); pos = end + 1
        (size,) = struct.unpack_from("<I", data, pos); pos += 4
        attrs[name] = (typ, data[pos:pos + size]); pos += size
    pos += 1
    assert attrs["compression"][1] == b"\0"
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    names = []
    ch = attrs["channels"][1]
    p = 0
    while ch[p] != 0:
        end = ch.index(b"\0", p); names.append(ch[p:end].decode()); p = end + 1
        (ptype,) = struct.unpack_from("<i", ch, p); assert ptype == 2, "FLOAT channels only"
        p += 16
    offsets = struct.unpack_from(f"<{h}Q", data, pos)
    out = np.zeros((h, w, 3), np.float32)
    for off in offsets:
        y, nbytes = struct.unpack_from("<iI", data, off)
        line = np.frombuffer(data, "<f4", count=w * len(names), offset=off + 8).reshape(len(names), w)
        assert nbytes == w * len(names) * 4
        for k, nme in enumerate(names):
            out[y - y0, :, "RGB".index(nme)] = line[k]
    return out

"""Seeded synthetic scenes for the BASELINE.json configs (no external assets).

* `cornell`        — the geometry constants of `Scene::cornell()` (yuki/src/scene/mod.rs:157-172, 242-495) without the
                     sphere (spheres are a later row); configs 1 and 2.
* `heightfield`    — jittered grid mesh wrapped exactly like `Scene::ply` (scene/mod.rs:99-150, scene/ply.rs:99-108);
                     config 3 (and the terrain of config 5).
* `material_room`  — Matte / Oren-Nayar / Glass / copper Metal / Glossy objects, a textured wall, point + spot + rect
                     lights; config 4.
* `terrain_room`   — config 5: ~10 M-triangle terrain plus the config-4 objects.

Every function takes a transform backend `xf` (yuki_b200.transforms, or the oracle's for cross-checks) so transform
arithmetic is the reference's f32 sequence. All other arithmetic here uses np.float32 scalars on purpose.
"""
from __future__ import annotations

import numpy as np

from . import desc as D

F = np.float32


def _quad(points, indices=(0, 1, 2, 0, 2, 3)):
    return np.array(points, dtype=np.float32), np.array(indices, dtype=np.uint32)


def checker_texture(size=256, cells=8, a=(0.8, 0.8, 0.8), b=(0.2, 0.3, 0.7)) -> np.ndarray:
    """Stand-in for res/tiling_58-1K/tiling_58_basecolor-1K.png, which is missing from the reference checkout."""
    y, x = np.mgrid[0:size, 0:size]
    mask = (((x * cells) // size + (y * cells) // size) % 2).astype(bool)
    img = np.empty((size, size, 3), np.float32)
    img[mask] = np.array(a, np.float32)
    img[~mask] = np.array(b, np.float32)
    return img


def cornell(xf, light="rect", tall_box="glass", textured_back_wall=False, split_method=D.SPLIT_SAH, max_shapes_in_node=1,
            sphere=False, back_wall_albedo=None):
    """Cornell box in metres, camera looking down -z (scene/mod.rs:154-531). `sphere=True` adds the copper sphere of
    scene/mod.rs:497-501 (with `split_method=D.SPLIT_MIDDLE` this is `Scene::cornell()` up to the missing back-wall PNG)."""
    LEFT, RIGHT, BOTTOM, TOP, FRONT, BACK = F(555.0), F(0.0), F(0.0), F(550.0), F(0.0), F(560.0)
    X_CENTER = (LEFT + RIGHT) / F(2.0)
    Z_CENTER = (FRONT + BACK) / F(2.0)
    HEIGHT = TOP - BOTTOM
    LIGHT_WH = F(100.0)
    LIGHT_HALF = LIGHT_WH / F(2.0)
    LIGHT_FRONT, LIGHT_BACK = Z_CENTER - LIGHT_HALF, Z_CENTER + LIGHT_HALF
    LIGHT_LEFT, LIGHT_RIGHT = X_CENTER + LIGHT_HALF, X_CENTER - LIGHT_HALF
    HOLE_TOP = TOP + HEIGHT * F(0.025)

    swap = xf.new([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1])
    to_world = xf.mul(xf.scale(0.001, 0.001, 0.001), swap)

    s = D.SceneDesc(split_method=split_method, max_shapes_in_node=max_shapes_in_node)
    zero = s.add_texture(D.Texture.constant(0.0))
    one = s.add_texture(D.Texture.constant(1.0))
    g180 = float(F(180.0) / F(255.0))
    white = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(g180)), zero)))
    red = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(g180, 0.0, 0.0)), zero)))
    green = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(0.0, g180, 0.0)), zero)))
    black = s.add_material(D.Material(D.MAT_MATTE, (zero, zero)))
    glass = s.add_material(D.Material(D.MAT_GLASS, (one, one), eta=1.5))
    back = white
    if textured_back_wall:
        back = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.from_image(checker_texture())), zero)))
    elif back_wall_albedo is not None:  # a constant stand-in for the marble PNG missing from the reference checkout (scene/mod.rs:193-200)
        back = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(*back_wall_albedo)), zero)))

    area_light = -1
    if light == "rect":  # scene/mod.rs:230-240
        size = (float(LIGHT_WH / F(1000.0)), float(LIGHT_WH / F(1000.0)))
        area = F(size[0]) * F(size[1])
        radiance = float(F(2.0) / (area * F(np.pi)))
        pos = (float(X_CENTER / F(1000.0)), float(HOLE_TOP / F(1000.0)), float(-Z_CENTER / F(1000.0)))
        s.lights.append(D.Light(D.LIGHT_RECT, xf.translation(pos), (radiance,) * 3, size=size))
        area_light = 0
    elif light == "point":  # a point light just below the ceiling hole (pbrt-v3 subset: AreaLightSource is ignored)
        s.lights.append(D.Light(D.LIGHT_POINT, xf.translation((0.2775, 0.54, -0.28)), (0.35, 0.35, 0.35)))
    elif light is not None:
        raise ValueError(light)

    def add(points, indices, material, uvs=None, al=-1):
        p, i = _quad(points, indices)
        s.meshes.append(D.Mesh(to_world, p, i, material, uvs=None if uvs is None else np.array(uvs, np.float32), area_light=al))

    q = (0, 1, 2, 0, 2, 3)
    # emissive quad in the ceiling hole
    add([(LIGHT_RIGHT, HOLE_TOP, LIGHT_FRONT), (LIGHT_LEFT, HOLE_TOP, LIGHT_FRONT), (LIGHT_LEFT, HOLE_TOP, LIGHT_BACK),
         (LIGHT_RIGHT, HOLE_TOP, LIGHT_BACK)], q, black, al=area_light)
    walls = [
        ([(RIGHT, BOTTOM, BACK), (LEFT, BOTTOM, BACK), (LEFT, BOTTOM, FRONT), (RIGHT, BOTTOM, FRONT)], q, white, None),          # floor
        ([(RIGHT, TOP, FRONT), (LEFT, TOP, FRONT), (LEFT, TOP, LIGHT_FRONT), (RIGHT, TOP, LIGHT_FRONT)], q, white, None),       # ceiling front
        ([(RIGHT, TOP, LIGHT_BACK), (LEFT, TOP, LIGHT_BACK), (LEFT, TOP, BACK), (RIGHT, TOP, BACK)], q, white, None),           # ceiling back
        ([(LIGHT_LEFT, TOP, FRONT), (LEFT, TOP, FRONT), (LEFT, TOP, BACK), (LIGHT_LEFT, TOP, BACK)], q, white, None),           # ceiling left
        ([(RIGHT, TOP, FRONT), (LIGHT_RIGHT, TOP, FRONT), (LIGHT_RIGHT, TOP, BACK), (RIGHT, TOP, BACK)], q, white, None),       # ceiling right
        ([(LIGHT_RIGHT, HOLE_TOP, LIGHT_FRONT), (LIGHT_LEFT, HOLE_TOP, LIGHT_FRONT), (LIGHT_LEFT, TOP, LIGHT_FRONT),
          (LIGHT_RIGHT, TOP, LIGHT_FRONT)], (0, 2, 1, 0, 3, 2), white, None),                                                     # hole front
        ([(LIGHT_RIGHT, HOLE_TOP, LIGHT_BACK), (LIGHT_LEFT, HOLE_TOP, LIGHT_BACK), (LIGHT_LEFT, TOP, LIGHT_BACK),
          (LIGHT_RIGHT, TOP, LIGHT_BACK)], q, white, None),                                                                       # hole back
        ([(LIGHT_LEFT, TOP, LIGHT_FRONT), (LIGHT_LEFT, TOP, LIGHT_BACK), (LIGHT_LEFT, HOLE_TOP, LIGHT_BACK),
          (LIGHT_LEFT, HOLE_TOP, LIGHT_FRONT)], q, white, None),                                                                  # hole left
        ([(LIGHT_RIGHT, HOLE_TOP, LIGHT_FRONT), (LIGHT_RIGHT, HOLE_TOP, LIGHT_BACK), (LIGHT_RIGHT, TOP, LIGHT_BACK),
          (LIGHT_RIGHT, TOP, LIGHT_FRONT)], q, white, None),                                                                      # hole right
        ([(RIGHT, TOP, BACK), (LEFT, TOP, BACK), (LEFT, BOTTOM, BACK), (RIGHT, BOTTOM, BACK)], q, back,
         [(0.0, 0.0), (0.0, 1.0), (1.0, 1.0), (1.0, 0.0)]),                                                                       # back wall (uvs)
        ([(RIGHT, TOP, FRONT), (RIGHT, TOP, BACK), (RIGHT, BOTTOM, BACK), (RIGHT, BOTTOM, FRONT)], q, green, None),              # right wall
        ([(LEFT, BOTTOM, FRONT), (LEFT, BOTTOM, BACK), (LEFT, TOP, BACK), (LEFT, TOP, FRONT)], q, red, None),                    # left wall
    ]
    for pts, idx, mat, uvs in walls:
        add(pts, idx, mat, uvs)
    if tall_box is not None:
        box_mat = {"glass": glass, "matte": white}[tall_box]
        add([(423.0, 330.0, 247.0), (265.0, 330.0, 296.0), (314.0, 330.0, 456.0), (472.0, 330.0, 406.0), (423.0, 0.0, 247.0),
             (472.0, 0.0, 406.0), (314.0, 0.0, 456.0), (265.0, 0.0, 296.0)],
            (0, 1, 2, 0, 2, 3, 4, 0, 3, 4, 3, 5, 5, 3, 2, 5, 2, 6, 6, 2, 1, 6, 1, 7, 7, 1, 0, 7, 0, 4), box_mat)
    if sphere:  # scene/mod.rs:214-223, 497-501
        copper = s.add_material(D.Material(D.MAT_METAL, (s.add_texture(D.Texture.constant(0.27105, 0.67693, 1.31640)),
                                                         s.add_texture(D.Texture.constant(3.60920, 2.62480, 2.29210)),
                                                         s.add_texture(D.Texture.constant(0.01))), remap_roughness=True))
        s.spheres.append(D.Sphere(xf.translation((0.186, 0.082, -0.168)), 0.082, copper))
    cam = D.CameraParameters((0.278, 0.273, 0.800), (0.278, 0.273, -0.260), fov_axis=D.FOV_X, fov_deg=40.0)
    return s, cam


def grid_mesh(nx, nz, seed=1, jitter=0.35, height=0.08, extent=(1.0, 1.0)):
    """(nx x nz)-quad heightfield: smooth bumps + per-vertex jitter from a seeded generator, so every triangle centroid
    is distinct (EqualCounts parity needs distinct keys). Returns points (N,3) f32 and indices (T*3,) u32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xs = np.linspace(-extent[0], extent[0], nx + 1, dtype=np.float64)
    zs = np.linspace(-extent[1], extent[1], nz + 1, dtype=np.float64)
    gx, gz = np.meshgrid(xs, zs, indexing="xy")
    cell = min(2 * extent[0] / nx, 2 * extent[1] / nz)
    gx = gx + rng.uniform(-jitter, jitter, gx.shape) * cell
    gz = gz + rng.uniform(-jitter, jitter, gz.shape) * cell
    gy = height * (np.sin(3.1 * gx + 0.3) * np.cos(2.7 * gz - 0.2) + 0.5 * np.sin(7.3 * gx * gz + 1.0)) + rng.uniform(
        -0.2, 0.2, gx.shape) * cell
    pts = np.stack([gx, gy, gz], axis=-1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(nx, dtype=np.int64), np.arange(nz, dtype=np.int64), indexing="xy")
    v00 = (j * (nx + 1) + i).reshape(-1)
    v10, v01, v11 = v00 + 1, v00 + (nx + 1), v00 + (nx + 2)
    tris = np.stack([v00, v01, v11, v00, v11, v10], axis=-1).reshape(-1).astype(np.uint32)
    return pts, tris


def fit_to_unit(xf, points):
    """The transform `Scene::ply` applies when the file gives none (scene/ply.rs:99-108)."""
    lo = points.min(axis=0).astype(np.float32)
    hi = points.max(axis=0).astype(np.float32)
    diag = hi - lo
    center = lo + diag / F(2.0)
    mesh_scale = float(F(1.0) / max(diag[0], max(diag[1], diag[2])))
    return xf.mul(xf.scale(mesh_scale, mesh_scale, mesh_scale), xf.translation((-center[0], -center[1], -center[2])))


def heightfield(xf, nx=708, nz=708, seed=1, split_method=D.SPLIT_SAH, max_shapes_in_node=1):
    """`Scene::ply` wrapper around a synthetic mesh: white Matte, point light (5,5,0) I=600, camera (2,2,2) -> origin,
    fov X 40 (scene/mod.rs:99-150). 708 x 708 quads = 1 002 528 triangles (config 3)."""
    pts, idx = grid_mesh(nx, nz, seed)
    s = D.SceneDesc(split_method=split_method, max_shapes_in_node=max_shapes_in_node)
    white = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(1.0)), s.add_texture(D.Texture.constant(0.0)))))
    s.meshes.append(D.Mesh(fit_to_unit(xf, pts), pts, idx, white))
    s.lights.append(D.Light(D.LIGHT_POINT, xf.translation((5.0, 5.0, 0.0)), (600.0, 600.0, 600.0)))
    cam = D.CameraParameters((2.0, 2.0, 2.0), (0.0, 0.0, 0.0), fov_axis=D.FOV_X, fov_deg=40.0)
    return s, cam


def ply(xf, path, split_method=D.SPLIT_SAH, max_shapes_in_node=1):
    """`Scene::ply` (scene/mod.rs:99-150) on a PLY file: the mesh scaled to fit 2 units around the origin
    (scene/ply.rs:99-108), white Matte, point light (5,5,0) I=600, camera (2,2,2) -> origin, fov X 40."""
    from . import api
    pts, idx, nrm, uvs = api.load_ply(path)
    s = D.SceneDesc(split_method=split_method, max_shapes_in_node=max_shapes_in_node)
    white = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(1.0)), s.add_texture(D.Texture.constant(0.0)))))
    s.meshes.append(D.Mesh(fit_to_unit(xf, pts), pts, idx, white, normals=nrm, uvs=uvs))
    s.lights.append(D.Light(D.LIGHT_POINT, xf.translation((5.0, 5.0, 0.0)), (600.0, 600.0, 600.0)))
    cam = D.CameraParameters((2.0, 2.0, 2.0), (0.0, 0.0, 0.0), fov_axis=D.FOV_X, fov_deg=40.0)
    return s, cam


def write_ply(path, points, indices, normals=None, uvs=None, fmt="binary_little_endian", quads=False):
    """Test / tooling helper: writes a mesh the way the reference's inputs look (vertex x y z [nx ny nz] [u v] float,
    face list uchar int vertex_indices). `quads=True` merges consecutive triangle pairs (a b c)(a c d) into 4-gons."""
    points = np.asarray(points, np.float32)
    idx = np.asarray(indices, np.uint32).reshape(-1, 3)
    faces = [list(t) for t in idx]
    if quads:
        faces = []
        k = 0
        while k < len(idx):
            if k + 1 < len(idx) and idx[k][0] == idx[k + 1][0] and idx[k][2] == idx[k + 1][1]:
                faces.append([idx[k][0], idx[k][1], idx[k][2], idx[k + 1][2]])
                k += 2
            else:
                faces.append(list(idx[k]))
                k += 1
    cols = [points]
    names = ["x", "y", "z"]
    if normals is not None:
        cols.append(np.asarray(normals, np.float32)); names += ["nx", "ny", "nz"]
    if uvs is not None:
        cols.append(np.asarray(uvs, np.float32)); names += ["u", "v"]
    verts = np.concatenate(cols, axis=1).astype(np.float32)
    header = ["ply", f"format {fmt} 1.0", "comment yuki_b200 test mesh", f"element vertex {len(verts)}"]
    header += [f"property float {n}" for n in names]
    header += [f"element face {len(faces)}", "property list uchar int vertex_indices", "end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode())
        if fmt == "ascii":
            for v in verts:
                f.write((" ".join(repr(float(x)) for x in v) + "\n").encode())
            for fc in faces:
                f.write((str(len(fc)) + " " + " ".join(str(int(i)) for i in fc) + "\n").encode())
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            f.write(verts.astype(e + "f4").tobytes())
            for fc in faces:
                f.write(np.uint8(len(fc)).tobytes() + np.asarray(fc, e + "i4").tobytes())


def _box(lo, hi):
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    p = [(x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0), (x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)]
    i = (0, 2, 1, 0, 3, 2, 4, 5, 6, 4, 6, 7, 0, 1, 5, 0, 5, 4, 3, 6, 2, 3, 7, 6, 0, 4, 7, 0, 7, 3, 1, 2, 6, 1, 6, 5)
    return np.array(p, np.float32), np.array(i, np.uint32)


def _uv_sphere(radius, stacks=24, slices=48, with_normals=True):
    """Triangulated sphere with smooth vertex normals and uvs (exercises shading normals, triangle.rs:197-224)."""
    th = np.linspace(0.0, np.pi, stacks + 1)
    ph = np.linspace(0.0, 2.0 * np.pi, slices + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    n = np.stack([np.sin(T) * np.cos(P), np.cos(T), np.sin(T) * np.sin(P)], axis=-1).reshape(-1, 3)
    pts = (n * radius).astype(np.float32)
    uv = np.stack([P / (2.0 * np.pi), T / np.pi], axis=-1).reshape(-1, 2).astype(np.float32)
    idx = []
    for a in range(stacks):
        for b in range(slices):
            v0 = a * (slices + 1) + b
            v1, v2, v3 = v0 + 1, v0 + slices + 1, v0 + slices + 2
            if a != 0:
                idx += [v0, v1, v2]
            if a != stacks - 1:
                idx += [v1, v3, v2]
    return pts, np.array(idx, np.uint32), (n.astype(np.float32) if with_normals else None), uv


def add_material_objects(xf, s: D.SceneDesc, y_floor=0.0, spread=1.0):
    """The five config-4 objects: Matte sigma=0, Matte sigma=20deg (Oren-Nayar), Glass eta 1.5, copper Metal
    (eta/k of scene/mod.rs:214-223, roughness 0.01 remapped), Glossy (Rs 0.5, roughness 0.3)."""
    zero = s.add_texture(D.Texture.constant(0.0))
    one = s.add_texture(D.Texture.constant(1.0))
    lambert = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(0.7, 0.6, 0.3)), zero)))
    sigma = float(np.float32(np.deg2rad(20.0)))
    oren = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(0.3, 0.6, 0.7)), s.add_texture(D.Texture.constant(sigma)))))
    glass = s.add_material(D.Material(D.MAT_GLASS, (one, one), eta=1.5))
    copper = s.add_material(D.Material(D.MAT_METAL, (s.add_texture(D.Texture.constant(0.27105, 0.67693, 1.31640)),
                                                     s.add_texture(D.Texture.constant(3.60920, 2.62480, 2.29210)),
                                                     s.add_texture(D.Texture.constant(0.01))), remap_roughness=True))
    glossy = s.add_material(D.Material(D.MAT_GLOSSY, (s.add_texture(D.Texture.constant(0.5)), s.add_texture(D.Texture.constant(0.3))),
                                       remap_roughness=False))
    r = 0.16 * spread
    xs = [-0.72, -0.36, 0.0, 0.36, 0.72]
    mats = [lambert, oren, glass, copper, glossy]
    for k, (x, m) in enumerate(zip(xs, mats)):
        if k % 2 == 0:
            p, i, n, uv = _uv_sphere(r)
            t = xf.translation((x * spread, y_floor + r, (-0.1 + 0.1 * k) * spread))
            s.meshes.append(D.Mesh(t, p, i, m, normals=n, uvs=uv))
        else:
            p, i = _box((-r * 0.8, 0.0, -r * 0.8), (r * 0.8, 2.2 * r, r * 0.8))
            t = xf.mul(xf.translation((x * spread, y_floor, (-0.1 + 0.1 * k) * spread)), xf.rotation(0.5 + 0.3 * k, (0.0, 1.0, 0.0)))
            s.meshes.append(D.Mesh(t, p, i, m))
    return s


def add_room_lights(xf, s: D.SceneDesc, y_top=1.2):
    """One point, one spot (total width 30deg, falloff start 20deg) and one rect light with its emissive quad."""
    s.lights.append(D.Light(D.LIGHT_POINT, xf.translation((-0.6, y_top * 0.8, 0.6)), (0.6, 0.55, 0.5)))
    # Spot: identity points down -Z (spot_light.rs:22); rotate -90deg about X to aim down -Y.
    spot_xf = xf.mul(xf.translation((0.5, y_top * 0.95, 0.3)), xf.rotation(float(np.float32(np.pi / 2)), (1.0, 0.0, 0.0)))
    s.lights.append(D.Light(D.LIGHT_SPOT, spot_xf, (2.5, 2.5, 2.5), total_width_deg=30.0, falloff_start_deg=20.0))
    rect_index = len(s.lights)
    size = (0.5, 0.5)
    rect_xf = xf.translation((0.0, y_top, 0.0))  # identity faces -Y (rectangular_light.rs:18)
    s.lights.append(D.Light(D.LIGHT_RECT, rect_xf, (6.0, 6.0, 6.0), size=size))
    zero = s.add_texture(D.Texture.constant(0.0))
    black = s.add_material(D.Material(D.MAT_MATTE, (zero, zero)))
    h = 0.25
    # emissive quad coincident with the light, wound so that its geometric normal faces -Y
    p, i = _quad([(-h, 0.0, -h), (h, 0.0, -h), (h, 0.0, h), (-h, 0.0, h)], (0, 1, 2, 0, 2, 3))
    s.meshes.append(D.Mesh(rect_xf, p, i, black, area_light=rect_index))
    return s


def material_room(xf, split_method=D.SPLIT_SAH):
    """Config 4: closed room (textured back wall) + the five material objects + point/spot/rect lights."""
    s = D.SceneDesc(split_method=split_method)
    zero = s.add_texture(D.Texture.constant(0.0))
    wall = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(0.73)), zero)))
    tex = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.from_image(checker_texture(1024, 16))), zero)))
    ident = xf.identity()
    X, Y0, Y1, Z = 1.2, 0.0, 1.25, 1.2
    def q(points, mat, uvs=None):
        p, i = _quad(points)
        s.meshes.append(D.Mesh(ident, p, i, mat, uvs=None if uvs is None else np.array(uvs, np.float32)))
    q([(-X, Y0, -Z), (-X, Y0, Z), (X, Y0, Z), (X, Y0, -Z)], wall)                                       # floor (normal +y)
    q([(-X, Y1, -Z), (X, Y1, -Z), (X, Y1, Z), (-X, Y1, Z)], wall)                                       # ceiling
    q([(-X, Y0, -Z), (X, Y0, -Z), (X, Y1, -Z), (-X, Y1, -Z)], tex, [(0, 0), (2, 0), (2, 1), (0, 1)])    # back wall, uv repeats
    q([(-X, Y0, Z), (-X, Y0, -Z), (-X, Y1, -Z), (-X, Y1, Z)], wall)                                     # left
    q([(X, Y0, -Z), (X, Y0, Z), (X, Y1, Z), (X, Y1, -Z)], wall)                                         # right
    add_material_objects(xf, s, y_floor=Y0)
    add_room_lights(xf, s, y_top=Y1 - 0.01)
    cam = D.CameraParameters((0.0, 0.7, 3.2), (0.0, 0.45, 0.0), fov_axis=D.FOV_X, fov_deg=38.0)
    return s, cam


def open_scene(xf):
    """Objects of every material on a floor under a sky: rays leave the scene at every depth (background term), lit by a
    DistantLight (distant_light.rs:17-43) and a point light."""
    s = D.SceneDesc(background=(0.2, 0.25, 0.3))
    zero = s.add_texture(D.Texture.constant(0.0))
    floor = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(0.6, 0.6, 0.55)), zero)))
    p, i = _quad([(-2, 0, -2), (-2, 0, 2), (2, 0, 2), (2, 0, -2)])
    s.meshes.append(D.Mesh(xf.identity(), p, i, floor))
    add_material_objects(xf, s)
    s.lights.append(D.Light(D.LIGHT_DISTANT, xf.identity(), (2.0, 1.9, 1.7), direction=(0.3, 1.0, 0.2)))
    s.lights.append(D.Light(D.LIGHT_POINT, xf.translation((-0.8, 1.5, 1.0)), (1.5, 1.5, 1.8)))
    cam = D.CameraParameters((0.0, 0.9, 2.6), (0.0, 0.2, 0.0), fov_axis=D.FOV_X, fov_deg=40.0)
    return s, cam


def terrain_room(xf, nx=3163, nz=1581, seed=5, split_method=D.SPLIT_SAH):
    """Config 5: jittered terrain (3163 x 1581 quads = 10 001 406 triangles) with the config-4 objects floating above,
    lit by the same point / spot / rect lights; open sky (grey background)."""
    pts, idx = grid_mesh(nx, nz, seed, height=0.05, extent=(2.4, 1.2))
    s = D.SceneDesc(split_method=split_method, background=(0.25, 0.3, 0.4))
    zero = s.add_texture(D.Texture.constant(0.0))
    ground = s.add_material(D.Material(D.MAT_MATTE, (s.add_texture(D.Texture.constant(0.55, 0.5, 0.42)), zero)))
    s.meshes.append(D.Mesh(xf.identity(), pts, idx, ground))
    add_material_objects(xf, s, y_floor=0.2)
    add_room_lights(xf, s, y_top=1.4)
    cam = D.CameraParameters((0.0, 1.1, 3.0), (0.0, 0.25, 0.0), fov_axis=D.FOV_X, fov_deg=42.0)
    return s, cam

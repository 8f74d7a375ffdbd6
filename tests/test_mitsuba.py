"""Mitsuba 2.1.0 loader (csrc/host_mitsuba.cpp) against the rules of yuki/src/scene/mitsuba/*.rs: a generated scene file
(PLY shapes, twosided / diffuse / dielectric bsdfs, constant / point / spot emitters, a transformed sensor) must load to
the description those rules give, each error case must be the reference's, and the loaded scene renders bit-identically
on the GPU and the oracle (the only file format that yields spot lights)."""
import numpy as np
import pytest

from yuki_b200 import api, capi, desc as D, scenes, transforms as xf

F = np.float32


def mirror_x(t):
    return xf.mul(xf.scale(-1.0, 1.0, 1.0), t)


def write_scene(tmp_path, sensor_xml=None, extra="", version="2.1.0", bsdfs=None, shapes=None):
    """Two PLY quads (floor, glass pane) + emitters; returns the file path."""
    floor = np.array([[-1, 0, -1], [1, 0, -1], [1, 0, 1], [-1, 0, 1]], F)
    pane = np.array([[-0.3, 0.0, 0.0], [0.3, 0.0, 0.0], [0.3, 0.6, 0.0], [-0.3, 0.6, 0.0]], F)
    quad = np.array([0, 1, 2, 0, 2, 3], np.uint32)
    (tmp_path / "meshes").mkdir(exist_ok=True)
    scenes.write_ply(tmp_path / "meshes" / "floor.ply", floor, quad, fmt="ascii")
    scenes.write_ply(tmp_path / "meshes" / "pane.ply", pane, quad, fmt="binary_little_endian")
    if sensor_xml is None:
        sensor_xml = """
  <sensor type="perspective">
    <string name="fov_axis" value="y"/>
    <float name="fov" value="35.5"/>
    <float name="near_clip" value="0.1"/>
    <float name="far_clip" value="100"/>
    <transform name="to_world">
      <rotate y="1" angle="180"/>
      <rotate x="1" angle="-20"/>
      <translate value="0.5 1.5 4"/>
    </transform>
    <sampler type="independent"><integer name="sample_count" value="16"/></sampler>
    <film type="hdrfilm"><integer name="width" value="$resx"/><rfilter type="box"/></film>
  </sensor>"""
    if bsdfs is None:
        bsdfs = """
  <bsdf type="twosided" id="mat-floor"><bsdf type="diffuse"><rgb name="reflectance" value="0.8 0.25 0.125"/></bsdf></bsdf>
  <bsdf type="diffuse" id="mat-grey"/>
  <bsdf type="dielectric" id="mat-glass"><float name="int_ior" value="1.45"/><float name="ext_ior" value="1.000277"/>
    <rgb name="specular_transmittance" value="0.9 0.95 1"/></bsdf>"""
    if shapes is None:
        shapes = """
  <shape type="ply"><string name="filename" value="meshes/floor.ply"/><ref name="bsdf" id="mat-floor"/></shape>
  <shape type="ply"><string name="filename" value="meshes\\pane.ply"/>
    <transform name="to_world"><scale value="1.5"/><rotate y="1" angle="30"/><translate value="0.25 0 -0.5"/></transform>
    <ref name="bsdf" id="mat-glass"/></shape>"""
    text = f"""<?xml version="1.0" encoding="utf-8"?>
<!-- generated test scene -->
<scene version="{version}">
  <default name="spp" value="16"/>
  <default name="resx" value="96"/>
  <default name="resy" value="72"/>
  <integrator type="path"><integer name="max_depth" value="12"/></integrator>{sensor_xml}{bsdfs}
  <emitter type="constant"><rgb name="radiance" value="0.05 0.06 0.07"/></emitter>
  <emitter type="point"><point name="position" x="1.25" y="2" z="0.5"/><rgb name="intensity" value="6 5 4"/></emitter>
  <emitter type="spot"><float name="cutoff_angle" value="30"/><float name="beam_width" value="20"/>
    <transform name="to_world"><rotate x="1" angle="180"/><translate value="-0.5 2.5 0.25"/></transform>
    <rgb name="intensity" value="40 40 35"/></emitter>
  <emitter type="area"><rgb name="radiance" value="1 1 1"/></emitter>{shapes}{extra}
</scene>
"""
    path = tmp_path / "scene.xml"
    path.write_text(text)
    return path


def test_scene_file_loads_to_the_reference_description(tmp_path):
    sc, cam, film = api.load_mitsuba(write_scene(tmp_path))
    assert film.res == (96, 72) and film.tile_dim == 16                      # <default resx/resy>, mod.rs:73-83
    assert sc.background == tuple(float(F(v)) for v in (0.05, 0.06, 0.07))  # constant emitter
    # materials: twosided -> its nested diffuse; diffuse default 0.5 grey; dielectric -> Glass(R=1, T, int_ior)
    kinds = [m.kind for m in sc.materials]
    assert kinds == [D.MAT_MATTE, D.MAT_MATTE, D.MAT_GLASS]
    tex = lambda m, k: sc.textures[sc.materials[m].tex[k]].value
    assert tex(0, 0) == (float(F(0.8)), 0.25, 0.125) and tex(0, 1)[0] == 0.0
    assert tex(1, 0) == (0.5, 0.5, 0.5)
    assert tex(2, 0) == (1.0, 1.0, 1.0) and tex(2, 1) == (float(F(0.9)), float(F(0.95)), 1.0)
    assert sc.materials[2].eta == float(F(1.45))
    # lights: point with x mirrored; spot with scale(-1,1,1) * to_world, cutoff -> total width, beam -> falloff start
    assert [l.kind for l in sc.lights] == [D.LIGHT_POINT, D.LIGHT_SPOT]      # the area emitter is skipped
    assert np.array_equal(sc.lights[0].light_to_world.m, xf.translation((-1.25, 2.0, 0.5)).m)
    assert sc.lights[0].intensity == (6.0, 5.0, 4.0)
    spot = mirror_x(xf.mul(xf.translation((-0.5, 2.5, 0.25)), xf.rotation(np.deg2rad(F(180.0)), (1.0, 0.0, 0.0))))
    assert np.array_equal(sc.lights[1].light_to_world.m, spot.m) and np.array_equal(sc.lights[1].light_to_world.m_inv, spot.m_inv)
    assert (sc.lights[1].total_width_deg, sc.lights[1].falloff_start_deg) == (30.0, 20.0)
    # shapes: PLY meshes in file order, object_to_world = scale(-1,1,1) * (translate * rotate * scale), material by id
    assert len(sc.meshes) == 2 and sc.objects == [0, 1] and not sc.spheres
    assert np.array_equal(sc.meshes[0].object_to_world.m, xf.scale(-1.0, 1.0, 1.0).m) and sc.meshes[0].material == 0
    t = xf.mul(xf.translation((0.25, 0.0, -0.5)), xf.mul(xf.rotation(np.deg2rad(F(30.0)), (0.0, 1.0, 0.0)), xf.scale(1.5, 1.5, 1.5)))
    assert np.array_equal(sc.meshes[1].object_to_world.m, mirror_x(t).m) and sc.meshes[1].material == 2
    assert np.array_equal(sc.meshes[1].indices, [0, 1, 2, 0, 2, 3]) and sc.meshes[1].points.shape == (4, 3)
    # sensor: position = mirrored translation, fov axis / angle as given
    assert cam.fov_axis == D.FOV_Y and cam.fov_deg == 35.5
    assert np.allclose(cam.position, (-0.5, 1.5, 4.0), atol=1e-6)


def euler_camera(m):
    """sensor.rs:71-106 + Matrix4x4::decompose (math/matrix.rs:217-255) in float64."""
    m = np.diag([-1.0, 1.0, 1.0, 1.0]) @ m
    pos = m[:3, 3]
    r = m[:3, :3] / np.linalg.norm(m[:3, :3], axis=0)
    tx = np.arctan2(r[1, 2], r[2, 2])
    ty = np.arctan2(-r[0, 2], np.hypot(r[0, 0], r[0, 1]))
    s1, c1 = np.sin(tx), np.cos(tx)
    tz = np.arctan2(s1 * r[2, 0] - c1 * r[1, 0], c1 * r[1, 1] - s1 * r[2, 1])

    def rot(axis, a):
        c, s = np.cos(a), np.sin(a)
        i, j = [(1, 2), (2, 0), (0, 1)][axis]
        out = np.eye(3)
        out[i, i] = c; out[i, j] = -s; out[j, i] = s; out[j, j] = c
        return out
    rr = rot(0, -tx) @ rot(1, -ty) @ rot(2, tz)
    return pos, rr @ np.array([0.0, 0.0, 1.0]), rr @ np.array([0.0, 1.0, 0.0])


def test_sensor_matrix_decomposition_and_retargeting(tmp_path):
    def rot(axis, deg):
        a = np.deg2rad(deg)
        c, s = np.cos(a), np.sin(a)
        m = np.eye(4)
        i, j = [(1, 2), (2, 0), (0, 1)][axis]
        m[i, i] = c; m[i, j] = -s; m[j, i] = s; m[j, j] = c
        return m
    tr = np.eye(4)
    tr[:3, 3] = (0.5, 1.5, 4.0)
    to_world = tr @ rot(0, -20.0) @ rot(1, 180.0)          # entries pre-multiply in file order (transform.rs:46,55)
    sc, cam, _ = api.load_mitsuba(write_scene(tmp_path))
    pos, fwd, up = euler_camera(to_world)
    assert np.allclose(cam.position, pos, atol=1e-5)
    assert np.allclose(cam.up, up, atol=1e-5)
    d = np.asarray(cam.target, np.float64) - np.asarray(cam.position, np.float64)
    assert np.allclose(d / np.linalg.norm(d), fwd, atol=1e-5)
    # mod.rs:185-197: the target sits midway through the part of the scene bounds the view direction crosses
    pts = []
    for m in sc.meshes:
        mm = np.asarray(m.object_to_world.m, np.float64).reshape(4, 4)
        pts.append((mm[:3, :3] @ m.points.T.astype(np.float64)).T + mm[:3, 3])
    pts = np.concatenate(pts)
    lo, hi = pts.min(0), pts.max(0)
    with np.errstate(divide="ignore"):
        t0, t1 = (lo - pos) / fwd, (hi - pos) / fwd
    p0, p1 = max(np.minimum(t0, t1).max(), 0.0), np.maximum(t0, t1).min()
    assert p0 <= p1
    assert np.isclose(np.linalg.norm(d), (p0 + p1) / 2 if p0 > 0 else p1 / 2, rtol=1e-4)
    # the same camera given as one matrix entry
    vals = " ".join(repr(float(v)) for v in to_world.astype(F).reshape(-1))
    sensor = f"""
  <sensor type="perspective"><string name="fov_axis" value="x"/><float name="fov" value="40"/>
    <transform name="to_world"><matrix value="{vals}"/></transform></sensor>"""
    _, cam2, _ = api.load_mitsuba(write_scene(tmp_path, sensor_xml=sensor))
    assert cam2.fov_axis == D.FOV_X and np.allclose(cam2.position, cam.position, atol=1e-5) and np.allclose(cam2.up, cam.up, atol=1e-5)


@pytest.mark.parametrize("kw,msg", [
    (dict(version="3.0.0"), "Scene file version is not 2.1.0"),
    (dict(extra="<texture type='bitmap'/>"), "Unknown element: 'texture'"),
    (dict(bsdfs='<bsdf type="plastic" id="mat-floor"/>'), "Unknown bsdf type 'plastic'"),
    (dict(bsdfs='<bsdf type="diffuse"/>'), "Could not find element attribute 'id'"),
    (dict(bsdfs='<bsdf type="diffuse" id="a"><rgb name="albedo" value="1 1 1"/></bsdf>'), "Expected rgb to be 'reflectance', got 'albedo'"),
    (dict(bsdfs='<bsdf type="diffuse" id="a"><float name="x" value="1"/></bsdf>'), "Unknown light data type 'float'"),
    (dict(bsdfs='<bsdf type="dielectric" id="mat-glass"><float name="ext_ior" value="1.33"/></bsdf>'), "Only air supported"),
    (dict(bsdfs='<bsdf type="dielectric" id="mat-glass"><rgb name="tint" value="1 1 1"/></bsdf>'), "Unknown dielectric rgb data 'tint'"),
    (dict(bsdfs='<bsdf type="diffuse" id="other"/>'), "Unknown mesh material 'mat-floor'"),
    (dict(shapes='<shape type="sphere"/>'), "Unexpected shape type 'sphere'!"),
    (dict(shapes='<shape type="ply"><ref name="bsdf" id="mat-grey"/></shape>'), "Mesh with no ply"),
    (dict(shapes='<shape type="ply"><string name="filename" value="meshes/floor.ply"/></shape>'), "Mesh with no material"),
    (dict(shapes='<shape type="ply"><string name="filename" value="meshes/none.ply"/><ref name="bsdf" id="mat-grey"/></shape>'), "Error canonicalizing"),
    (dict(shapes='<shape type="ply"><string name="filename" value="meshes/floor.ply"/><ref name="bsdf" id="mat-grey"/>'
                 '<transform name="to_world"><lookat origin="0 0 0"/></transform></shape>'), "Unknown transformation data type 'lookat'"),
    (dict(sensor_xml='<sensor type="perspective"><string name="fov_axis" value="diagonal"/></sensor>'), "Unknown fov axis 'diagonal'"),
    (dict(sensor_xml='<sensor type="perspective"><string name="fov_axis" value="x"/><transform name="to_world"><scale value="2"/></transform></sensor>'),
     "Camera to world has scaling"),
    (dict(sensor_xml='<sensor type="perspective"><boolean name="x" value="true"/></sensor>'), "Unknown sensor data type 'boolean'"),
    (dict(extra="stray text"), "Unexpected characters outside tags: stray text"),
    (dict(extra="<![CDATA[data]]>"), "Unexpected CDATA: data"),
    (dict(extra="<?php echo ?>"), "Unexpected processing instruction: php"),
    (dict(extra='<emitter type="point"><point name="position" x="1" w="2"/></emitter>'), "Invalid point axis 'w'"),
    (dict(extra='<emitter type="spot"><float name="radius" value="2"/></emitter>'), "Unexpected spot light float 'name': 'radius'"),
])
def test_error_cases_are_the_references(tmp_path, kw, msg):
    with pytest.raises(capi.YukiGpuError) as e:
        api.load_mitsuba(write_scene(tmp_path, **kw))
    assert msg in str(e.value)


def test_ignored_and_default_pieces(tmp_path):
    """Skipped subtrees may hold anything; a twosided bsdf without children is white; later bsdfs override an id; entities in
    attribute values are decoded; scenes without a sensor keep the default camera (camera.rs:32-41)."""
    bsdfs = """
  <bsdf type="twosided" id="mat-floor"/>
  <bsdf type="diffuse" id="mat-glass"><rgb name="reflectance" value="0.1 0.2 0.3"/></bsdf>
  <bsdf type="diffuse" id="mat-glass"><rgb name="reflectance" value="0.25"/></bsdf>"""
    extra = """<integrator type="volpath"><weird><nested a="1 &lt; 2 &amp;&#38; 3"/></weird></integrator>
  <emitter type="envmap"><string name="filename" value="sky.exr"/><transform name="to_world"><lookat/></transform></emitter>"""
    sc, cam, film = api.load_mitsuba(write_scene(tmp_path, sensor_xml="", bsdfs=bsdfs, extra=extra))
    tex = lambda m, k: sc.textures[sc.materials[m].tex[k]].value
    assert tex(sc.meshes[0].material, 0) == (1.0, 1.0, 1.0)
    assert tex(sc.meshes[1].material, 0) == (0.25, 0.0, 0.0)      # parse_rgb fills only the components given (common.rs:11-16)
    assert cam.position == (0.0, 0.0, 0.0) and cam.up == (0.0, 1.0, 0.0) and cam.fov_axis == D.FOV_X and cam.fov_deg == 0.0


@pytest.mark.gpu
def test_mitsuba_scene_renders_bit_identically(tmp_path, gpu_ctx, oracle):
    sc, cam, film = api.load_mitsuba(write_scene(tmp_path))
    dev = api.Scene(gpu_ctx, sc)
    sampler, integ = D.SamplerType.stratified(3, 3), D.IntegratorType.path(6)
    r = api.Renderer(gpu_ctx).render(dev, cam, film, sampler, integ, want_hit_ids=True)
    o_img, o_ids, o_st = oracle.OracleScene(sc).render(cam, film, sampler, integ, want_hit_ids=True)
    dev.close()
    assert np.array_equal(r.hit_ids, o_ids) and (o_ids >= 0).mean() > 0.15
    assert np.array_equal(r.film.view(np.uint32), o_img.view(np.uint32))
    assert r.stats.shadow_rays == o_st.shadow_rays and r.stats.any_nodes == o_st.any_nodes
    assert float(o_img.max()) > 0.05   # the spot and point lights reach the floor

"""The C-ABI library loads, exports every symbol include/yuki_gpu.h declares, keeps its struct layouts, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from yuki_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "yuki_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yk_[a-z0-9_]+)\s*\(", text)) - {"yk_progress_fn"})


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in yuki_gpu.h but not exported"
    assert sorted(capi.EXPORTS) == names


def test_struct_layouts_match_the_header():
    assert C.sizeof(capi.BvhNode) == 32 == capi.NODE_DTYPE.itemsize          # bvh.rs:556
    assert C.sizeof(capi.Tile) == 16 == capi.TILE_DTYPE.itemsize
    assert C.sizeof(capi.Transform) == 128 and C.sizeof(capi.Camera) == 128
    assert C.sizeof(capi.CameraParams) == 44 and C.sizeof(capi.Sampler) == 24 and C.sizeof(capi.Integrator) == 16
    assert C.sizeof(capi.LightDev) == 4 + 12 + 12 + 8 + 3 * 64 + 4
    assert C.sizeof(capi.Stats) == 15 * 8


def test_no_cpu_fallback_without_a_device():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = capi.lib().yk_context_create(0, C.byref(h))
    assert rc == -2 and not h.value                                           # YK_ERR_CUDA
    msg = capi.lib().yk_last_error().decode()
    assert "no CUDA device" in msg and "no CPU fallback" in msg
    from yuki_b200 import api
    with pytest.raises(capi.YukiGpuError):
        api.Context(0)


def test_null_arguments_are_rejected_not_crashing():
    L = capi.lib()
    assert L.yk_context_create(0, None) == -1
    assert L.yk_host_scene_build(None, None) == -1
    assert L.yk_camera_make(None, 16, 16, None) == -1
    assert L.yk_film_tiles(0, 0, 16, None, 0) == 0
    assert L.yk_render(None, None, None, None, None, None, None, 0, None, None, None) == -1
    assert b"null" in L.yk_last_error()


def test_host_scene_validation(xf):
    from yuki_b200 import api, desc as D, scenes
    scene, _ = scenes.cornell(xf)
    scene.meshes[1].material = 99
    with pytest.raises(capi.YukiGpuError):
        api.HostScene(scene)
    scene, _ = scenes.cornell(xf, light="point")
    scene.meshes[0].area_light = 0                                            # a point light cannot be an area light
    with pytest.raises(capi.YukiGpuError):
        api.HostScene(scene)


C_CLIENT = r"""
/* A plain C99 client of include/yuki_gpu.h: what a Rust `extern "C"` / bindgen crate sees. */
#include <stdio.h>
#include <stdlib.h>
#include "yuki_gpu.h"
int main(void) {
    uint32_t n = yk_film_tiles(1024, 1024, 16, NULL, 0);
    yk_tile* tiles = (yk_tile*)malloc(sizeof(yk_tile) * n);
    if (yk_film_tiles(1024, 1024, 16, tiles, n) != n) return 2;
    yk_camera_params cp = {{0.f, 0.f, -3.f}, {0.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, YK_FOV_X, 40.f};
    yk_camera cam;
    if (yk_camera_make(&cp, 1024, 1024, &cam) != YK_OK) return 3;
    yk_context* ctx = NULL;
    int rc = yk_context_create(0, &ctx);
    printf("%u %u %u %d %d %zu %zu\n", n, (unsigned)tiles[0].x0, (unsigned)tiles[0].y0, rc, rc == YK_OK ? 0 : (int)(yk_last_error()[0] != 0),
           sizeof(yk_bvh_node), sizeof(yk_tile));
    if (ctx) yk_context_destroy(ctx);
    free(tiles);
    return 0;
}
"""


def test_header_is_plain_c_and_a_c_client_links(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C99 and a C program must link and call it."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "client.c"
    src.write_text(C_CLIENT)
    exe = tmp_path / "client"
    libdir = os.path.join(ROOT, "yuki_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lyuki_gpu", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    n, x0, y0, rc, has_msg, node_size, tile_size = (int(v) for v in out)
    assert n == 4096 and (x0, y0) == (496, 496)           # first spiral tile = the centre tile (film.rs:340-347)
    assert rc in (0, -2) and (rc == 0 or has_msg == 1)     # YK_OK with a GPU, YK_ERR_CUDA + message without one
    assert node_size == 32 and tile_size == 16

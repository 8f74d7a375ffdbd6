"""Diagnostic: per-pixel comparison of the CUDA path and the oracle on a small scene."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from yuki_b200 import api, desc as D, scenes, transforms as xf
from oracle import oracle as O

def run(name, scene, cam, film, sampler, integ):
    ctx = api.Context(0)
    dev = api.Scene(ctx, scene)
    r = api.Renderer(ctx).render(dev, cam, film, sampler, integ, want_hit_ids=True)
    o_img, o_ids, o_st = O.OracleScene(scene).render(cam, film, sampler, integ, want_hit_ids=True)
    g = r.film.astype(np.float64); o = o_img.astype(np.float64)
    diff = np.abs(g - o).max(axis=2)
    rel = diff / np.maximum(o.max(axis=2), 1e-6)
    print(f"== {name}: rays gpu {r.stats.ray_count} cpu {o_st.ray_count} | shadow gpu {r.stats.shadow_rays} cpu {o_st.shadow_rays}")
    print("   nodes gpu", r.stats.closest_nodes, "cpu", o_st.closest_nodes, "| any nodes", r.stats.any_nodes, o_st.any_nodes)
    print("   bit-equal pixels:", int((r.film.view(np.uint32) == o_img.view(np.uint32)).all(axis=2).sum()), "of", diff.size)
    print("   pixels rel>1e-5:", int((rel > 1e-5).sum()), " rel>1e-3:", int((rel > 1e-3).sum()), " rel>0.1:", int((rel > 0.1).sum()))
    print("   rmse", np.sqrt(np.mean((g - o) ** 2)) / o.mean(), "mean", o.mean(), "max", o.max())
    ys, xs = np.where(rel > 1e-3)
    for y, x in list(zip(ys, xs))[:8]:
        print("   px", x, y, "gpu", r.film[y, x], "cpu", o_img[y, x])
    dev.close(); ctx.close()

which = sys.argv[1] if len(sys.argv) > 1 else "all"
film = D.FilmSettings((128, 128), 16)
if which in ("all", "path_matte"):
    s, c = scenes.cornell(xf, light="rect", tall_box="matte")
    run("path matte box d3", s, c, film, D.SamplerType.stratified(2, 2), D.IntegratorType.path(3))
    run("path matte box d8", s, c, film, D.SamplerType.stratified(2, 2), D.IntegratorType.path(8))
if which in ("all", "path_glass"):
    s, c = scenes.cornell(xf, light="rect", tall_box="glass")
    run("path glass box d8", s, c, film, D.SamplerType.stratified(2, 2), D.IntegratorType.path(8))
    run("path glass box d2", s, c, film, D.SamplerType.stratified(2, 2), D.IntegratorType.path(2))
if which in ("all", "point"):
    s, c = scenes.cornell(xf, light="point", tall_box="matte")
    run("path point matte d8", s, c, film, D.SamplerType.stratified(2, 2), D.IntegratorType.path(8))
if which in ("all", "room"):
    s, c = scenes.material_room(xf)
    run("room d8", s, c, D.FilmSettings((160, 90), 16), D.SamplerType.stratified(2, 2), D.IntegratorType.path(8))

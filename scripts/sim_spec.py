"""Offline SIMT model of *postponed leaves* in the two-phase traversal (design aid, not product code).

Today a lane that reaches a leaf parks until its warp's triangle phase (wf_trace.cuh). Here a lane may keep walking boxes with
up to D leaves pending; pending leaves are served first-in first-out in the triangle phase, each behind a re-test of its own
box key against the *current* t_max. That re-test makes the sequence of triangle tests identical to the reference's (a child's
slab entry is never below its parent's, so a leaf the reference would have culled through any ancestor fails its own key), hence
the same hits, ties included; the price is box tests taken with a stale t_max. This script runs real rays through the real BVH
with a lane-exact engine and replays 32-lane warps to see what the trade buys.

usage: python scripts/sim_spec.py [cornell|room|hf] [closest|any]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from yuki_b200 import api, desc as Dsc, scenes, transforms as xf
from sim_warp import make_rays

INF = float("inf")


class Lane:
    __slots__ = ("o", "d", "inv", "neg", "t_max", "stack", "cur", "fifo", "pos", "anyhit", "n_box", "n_tri", "alive")

    def __init__(self):
        self.alive = False

    def start(self, S, o, d, t_max, anyhit):
        self.o, self.d = o, d
        with np.errstate(divide="ignore"):
            self.inv = 1.0 / d
        self.neg = self.inv < 0
        self.t_max = t_max
        self.stack, self.fifo, self.pos, self.cur = [], [], None, None
        self.anyhit = anyhit
        self.n_box = 1
        self.n_tri = 0
        self.alive = True
        lo, hi = S.slab(self, 0)
        if lo <= min(hi, self.t_max):
            self.enter(S, (0, lo))
        self.check_done()

    def pop_passing(self):
        while self.stack:
            ref, key = self.stack.pop()
            if key <= self.t_max:
                return (ref, key)
        return None

    def enter(self, S, take):
        while True:
            if take is None:
                self.cur = None
                return
            ref, key = take
            if not S.is_leaf[ref]:
                self.cur = ref
                return
            self.fifo.append((int(S.offset[ref]), int(S.offset[ref] + S.count[ref]), key))
            self.cur = None
            if len(self.fifo) > S.D:
                return  # parked
            take = self.pop_passing()

    def wants_box(self):
        return self.alive and self.cur is not None

    def has_leaf(self):
        return self.alive and bool(self.fifo)

    def check_done(self):
        if self.alive and self.cur is None and not self.fifo:
            # the walk may continue from the stack (a leaf was dropped or finished while parked)
            self.alive = False

    def box_step(self, S):
        i = self.cur
        c0, c1 = i + 1, int(S.offset[i])
        near, far = (c1, c0) if self.neg[S.axis[i]] else (c0, c1)
        lo_n, hi_n = S.slab(self, near)
        lo_f, hi_f = S.slab(self, far)
        self.n_box += 2
        hit_n = lo_n <= min(hi_n, self.t_max)
        ok_f = lo_f <= hi_f
        key_f = lo_f if ok_f else INF
        if hit_n:
            if ok_f:
                self.stack.append((far, key_f))
            take = (near, lo_n)
        elif key_f <= self.t_max:
            take = (far, key_f)
        else:
            take = self.pop_passing()
        self.enter(S, take)
        self.check_done()

    def tri_step(self, S):
        first, end, key = self.fifo[0]
        if self.pos is None:
            if not (key <= self.t_max):  # the leaf's own deferred box test, with the current t_max
                self.fifo.pop(0)
                self.after_leaf(S)
                return False
            self.pos = first
        t = S.tri_test(self, self.pos)
        self.n_tri += 1
        self.pos += 1
        if t is not None:
            if self.anyhit:
                self.alive = False
                return True
            self.t_max = t
        if self.pos == end:
            self.fifo.pop(0)
            self.pos = None
            self.after_leaf(S)
        return True

    def after_leaf(self, S):
        if self.cur is None and len(self.fifo) <= S.D:
            self.enter(S, self.pop_passing())
        self.check_done()


class Scene:
    def __init__(self, nodes, tris, D):
        self.p_min = nodes["p_min"].astype(np.float64)
        self.p_max = nodes["p_max"].astype(np.float64)
        self.is_leaf = nodes["is_leaf"].astype(bool)
        self.offset = nodes["offset"].astype(np.int64)
        self.count = nodes["shape_count"].astype(np.int64)
        self.axis = nodes["split_axis"].astype(np.int64)
        self.tris = tris
        self.D = D

    def slab(self, ln, i):
        with np.errstate(invalid="ignore"):
            t0 = (self.p_min[i] - ln.o) * ln.inv
            t1 = (self.p_max[i] - ln.o) * ln.inv
        lo = max(np.fmax.reduce(np.fmin(t0, t1)), 0.0)
        hi = np.fmin.reduce(np.fmax(t0, t1))
        return lo, hi

    def tri_test(self, ln, s):
        p0, p1, p2 = self.tris[s]
        e1 = p1 - p0
        e2 = p2 - p0
        pv = np.cross(ln.d, e2)
        det = e1.dot(pv)
        if abs(det) < 1e-14:
            return None
        tv = ln.o - p0
        u = tv.dot(pv) / det
        if u < 0 or u > 1:
            return None
        qv = np.cross(tv, e1)
        v = ln.d.dot(qv) / det
        if v < 0 or u + v > 1:
            return None
        t = e2.dot(qv) / det
        return t if 0 < t <= ln.t_max else None


def simulate(S, rays, anyhit, K=14, refill_below=22, CN=45, CT=55, CREFILL=60, box_steps_per_vote=3, tri_min=1):
    lanes = [Lane() for _ in range(32)]
    qi = 0
    cost = 0
    n_box_steps = n_tri_steps = lanes_box = lanes_tri = 0
    tot_box = tot_tri = 0
    extra = CN // 8 if S.D > 0 else 0  # fifo bookkeeping per box step

    def refill():
        nonlocal qi, cost, tot_box, tot_tri
        got = False
        for ln in lanes:
            if not ln.alive and qi < len(rays):
                if hasattr(ln, "n_box") and ln.n_box is not None:
                    pass
                o, d, tm = rays[qi]
                qi += 1
                ln.start(S, o, d, tm, anyhit)
                got = True
        if got:
            cost += CREFILL

    done_box = done_tri = 0

    def harvest():
        nonlocal done_box, done_tri
        for ln in lanes:
            if not ln.alive and getattr(ln, "n_box", None) is not None:
                done_box += ln.n_box
                done_tri += ln.n_tri
                ln.n_box = None

    for ln in lanes:
        ln.n_box = None
    refill()
    while True:
        harvest()
        live = [ln for ln in lanes if ln.alive]
        if not live:
            if qi >= len(rays):
                break
            refill()
            continue
        # box phase
        while True:
            wb = [ln for ln in lanes if ln.wants_box()]
            if not wb:
                break
            parked = [ln for ln in lanes if ln.alive and not ln.wants_box()]
            if len(wb) < K and parked:
                break
            for _ in range(box_steps_per_vote):
                wb = [ln for ln in lanes if ln.wants_box()]
                if not wb:
                    break
                cost += CN + extra
                n_box_steps += 1
                lanes_box += len(wb)
                for ln in wb:
                    ln.box_step(S)
            cost += 10  # vote
        # triangle phase: drain every pending leaf
        while True:
            wt = [ln for ln in lanes if ln.has_leaf()]
            if not wt:
                break
            cost += CT
            n_tri_steps += 1
            lanes_tri += len(wt)
            for ln in wt:
                ln.tri_step(S)
        harvest()
        nbusy = sum(ln.alive for ln in lanes)
        if qi < len(rays) and nbusy < refill_below:
            refill()
    n = len(rays)
    return dict(cost=cost / n, lanes_box=lanes_box / max(n_box_steps, 1), lanes_tri=lanes_tri / max(n_tri_steps, 1), box_per_ray=done_box / n,
                tri_per_ray=done_tri / n, box_steps=n_box_steps, tri_steps=n_tri_steps)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "cornell"
    mode = sys.argv[2] if len(sys.argv) > 2 else "closest"
    n_rays = int(sys.argv[3]) if len(sys.argv) > 3 else 3200
    if which == "cornell":
        sc, _ = scenes.cornell(xf, light="rect", tall_box="glass")
    elif which == "room":
        sc, _ = scenes.material_room(xf)
    else:
        sc, _ = scenes.heightfield(xf, 160, 160)
    hs = api.HostScene(sc)
    nodes, tris = hs.nodes(), hs.tri_vertices().astype(np.float64)
    rng = np.random.default_rng(1)
    base = make_rays(tris, n_rays, rng)
    if mode == "closest":
        rays = [(o, d, INF) for o, d in base]
    else:  # shadow rays towards a point below the ceiling / above the scene
        lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
        target = np.array([(lo[0] + hi[0]) / 2, hi[1] - 0.02 * (hi[1] - lo[1]), (lo[2] + hi[2]) / 2])
        rays = [(o, target - o, 0.9999) for o, _ in base]
    anyhit = mode != "closest"
    print(f"{which} / {mode}: {len(tris)} triangles, {len(rays)} rays")
    ref = None
    for D in (0, 1, 2, 3):
        for K in (10, 14, 18, 24):
            S = Scene(nodes, tris, D)
            r = simulate(S, rays, anyhit, K=K)
            if ref is None and D == 0 and K == 14:
                ref = r["cost"]
            print(f"  D={D} K={K:2d}: {r['cost']:7.1f} warp-instr/ray ({100 * (ref or r['cost']) / r['cost']:5.1f} % of today's speed) lanes box {r['lanes_box']:4.1f} "
                  f"tri {r['lanes_tri']:4.1f} | box tests/ray {r['box_per_ray']:5.1f} tri tests/ray {r['tri_per_ray']:4.2f}", flush=True)

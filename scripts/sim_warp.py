"""Offline model of SIMT scheduling policies for the BVH traversal kernel (design aid, not product code).
Traces random bounce rays through the host-built BVH in Python, records each ray's exact step sequence
(N = box test, T = triangle test), then replays 32-lane warps under different phase policies and reports
warp-instruction cost per ray."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yuki_b200 import api, desc as D, scenes, transforms as xf

def trace_seq(nodes, tris, o, d):
    inv = 1.0 / d
    neg = inv < 0
    tmax = np.inf
    seq = []
    stack = []
    cur = 0
    while True:
        n = nodes[cur]
        seq.append('N')
        t0 = (n['p_min'] - o) * inv; t1 = (n['p_max'] - o) * inv
        tmin = max(np.minimum(t0, t1).max(), 0.0); tmx = min(np.maximum(t0, t1).min(), tmax)
        if tmin <= tmx:
            if n['is_leaf']:
                for s in range(n['offset'], n['offset'] + n['shape_count']):
                    seq.append('T')
                    p0, p1, p2 = tris[s]
                    e1 = p1 - p0; e2 = p2 - p0
                    pv = np.cross(d, e2); det = e1.dot(pv)
                    if abs(det) < 1e-12: continue
                    tv = o - p0; u = tv.dot(pv) / det
                    if u < 0 or u > 1: continue
                    qv = np.cross(tv, e1); v = d.dot(qv) / det
                    if v < 0 or u + v > 1: continue
                    t = e2.dot(qv) / det
                    if 0 < t <= tmax: tmax = t
                if not stack: break
                cur = stack.pop()
            else:
                if neg[n['split_axis']]:
                    stack.append(cur + 1); cur = n['offset']
                else:
                    stack.append(n['offset']); cur = cur + 1
        else:
            if not stack: break
            cur = stack.pop()
    return ''.join(seq)

def make_rays(tris, n, rng):
    e1 = tris[:, 1] - tris[:, 0]; e2 = tris[:, 2] - tris[:, 0]
    nrm = np.cross(e1, e2); area = np.linalg.norm(nrm, axis=1) / 2
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    centre = tris.reshape(-1, 3).mean(axis=0)
    pick = rng.choice(len(tris), size=n, p=area / area.sum())
    rays = []
    for t in pick:
        a, b = rng.random(2)
        if a + b > 1: a, b = 1 - a, 1 - b
        p = tris[t, 0] + a * e1[t] + b * e2[t]
        nn = nrm[t] if nrm[t].dot(centre - p) > 0 else -nrm[t]
        u1, u2 = rng.random(2)
        r = np.sqrt(u1); th = 2 * np.pi * u2
        x, y, z = r * np.cos(th), r * np.sin(th), np.sqrt(max(0, 1 - u1))
        s = np.cross(nn, [1, 0, 0] if abs(nn[0]) < 0.9 else [0, 1, 0]); s /= np.linalg.norm(s)
        tt = np.cross(nn, s)
        dd = x * s + y * tt + z * nn
        rays.append((p + nn * 1e-3, dd))
    return rays

CN, CT = 45, 100   # warp instructions per box step / triangle step
CREFILL = 60       # per refill event (loads + setup)

def simulate(seqs, policy, K=16, refill_below=22):
    """seqs: list of step strings. Returns (warp-instr per ray, avg lane utilisation in N, in T)."""
    q = list(seqs)
    lanes = [None] * 32   # (seq, pos)
    cost = 0; nN = nT = 0; lanesN = lanesT = 0; done = 0
    qi = 0
    def refill():
        nonlocal qi, cost
        got = False
        for i in range(32):
            if lanes[i] is None and qi < len(q):
                lanes[i] = [q[qi], 0]; qi += 1; got = True
        if got: cost += CREFILL
    refill()
    while True:
        busy = [l for l in lanes if l is not None]
        if not busy:
            if qi >= len(q): break
            refill(); continue
        wantN = [l for l in busy if l[0][l[1]] == 'N']
        wantT = [l for l in busy if l[0][l[1]] == 'T']
        if policy == 'postpone_all':      # current v2: N until nobody wants N, then T until nobody wants T
            ph = 'N' if wantN and (simulate.phase == 'N' or not wantT) else 'T'
            if simulate.phase == 'N' and not wantN: ph = 'T'
            if simulate.phase == 'T' and not wantT: ph = 'N'
        elif policy == 'threshold':       # stay in N while >= K lanes want N, else serve T if any
            if len(wantN) >= K or not wantT: ph = 'N'
            else: ph = 'T'
            if ph == 'N' and not wantN: ph = 'T'
        elif policy == 'drain':           # N while >= K lanes want N; then T until no lane wants T
            if simulate.phase == 'T' and wantT: ph = 'T'
            elif len(wantN) >= K or not wantT: ph = 'N'
            else: ph = 'T'
            if ph == 'N' and not wantN: ph = 'T'
        elif policy == 'majority':        # whichever phase has more lanes (weighted)
            ph = 'N' if len(wantN) * K >= len(wantT) * 16 else 'T'
            if ph == 'N' and not wantN: ph = 'T'
            if ph == 'T' and not wantT: ph = 'N'
        simulate.phase = ph
        group = wantN if ph == 'N' else wantT
        cost += CN if ph == 'N' else CT
        if ph == 'N': nN += 1; lanesN += len(group)
        else: nT += 1; lanesT += len(group)
        for l in group:
            l[1] += 1
        for i in range(32):
            if lanes[i] is not None and lanes[i][1] >= len(lanes[i][0]):
                lanes[i] = None; done += 1
        nbusy = sum(l is not None for l in lanes)
        if qi < len(q) and nbusy < refill_below: refill()
    return cost / len(seqs), lanesN / max(nN, 1), lanesT / max(nT, 1)
simulate.phase = 'N'

if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'cornell'
    if which == 'cornell': sc, _ = scenes.cornell(xf, light='rect', tall_box='glass')
    else: sc, _ = scenes.material_room(xf)
    hs = api.HostScene(sc)
    nodes, tris = hs.nodes(), hs.tri_vertices().astype(np.float64)
    rng = np.random.default_rng(1)
    rays = make_rays(tris, 3200, rng)
    seqs = [trace_seq(nodes, tris, o, d) for o, d in rays]
    nn = np.mean([s.count('N') for s in seqs]); nt = np.mean([s.count('T') for s in seqs])
    ideal = (nn * CN + nt * CT) / 32
    print(f"{which}: nodes/ray {nn:.1f} tris/ray {nt:.1f} ideal warp-instr/ray {ideal:.1f}")
    CN, CT = 45, 55
    ideal = (nn * CN + nt * CT) / 32
    for pol, Ks in (('postpone_all', [0]), ('threshold', [8, 12, 16, 20]), ('drain', [8, 12, 16, 20, 24])):
        for K in Ks:
            for rb in (16, 22, 26):
                c, uN, uT = simulate(seqs, pol, K, rb)
                print(f"  {pol:13s} K={K:2d} refill<{rb}: {c:7.1f} warp-instr/ray (eff {100*ideal/c:4.1f}%) lanes N {uN:4.1f} T {uT:4.1f}")
    print("-- coherence experiments (threshold K=16, refill<22)")
    def key_oct(r): o, d = r; return (d[0] < 0) * 4 + (d[1] < 0) * 2 + (d[2] < 0)
    lo = tris.reshape(-1, 3).min(0); hi = tris.reshape(-1, 3).max(0)
    def key_cell(r, g=4):
        o, d = r; c = np.minimum(((o - lo) / (hi - lo + 1e-9) * g).astype(int), g - 1)
        return key_oct(r) * g**3 + c[0] * g * g + c[1] * g + c[2]
    for name, kf in (('octant', key_oct), ('octant+cell4', key_cell), ('octant+cell8', lambda r: key_cell(r, 8))):
        idx = sorted(range(len(rays)), key=lambda i: kf(rays[i]))
        c, uN, uT = simulate([seqs[i] for i in idx], 'threshold', 16, 22)
        print(f"  sorted by {name:14s}: {c:7.1f} warp-instr/ray (eff {100*ideal/c:4.1f}%) lanes N {uN:4.1f} T {uT:4.1f}")

#!/usr/bin/env python
"""blend_cull_model.py — offline model of the blend's per-warp candidate lists on the C2 scene (CPU, uses the oracle
as the source of the sorted lists).  Compares inner-loop trip counts of three warp layouts:
  A  warp = 8x4 pixels, one list per warp (what blend_culled_kernel does)
  B  warp = two 4x4 half-warps, one list per half-warp, trips = max of the two
  C  warp = 8x8 pixels, two pixels per thread, one list per warp
Termination is approximated from the oracle's n_contrib / final_T (a sub-rectangle stops once all of its
pixels have terminated).  Analysis tool only (lives under tests/ because it drives the oracle); nothing in the product imports it."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gsrast_b200 import camera, scene
from oracle import gsr_oracle

def max_power_box(a, b, c, x0, x1, y0, y1):
    mba, mbc = -b / a, -b / c
    ex = np.minimum(np.maximum(0.0, x0), x1)
    ey = np.minimum(np.maximum(0.0, y0), y1)
    dy1 = np.minimum(np.maximum(mbc * ex, y0), y1)
    dx2 = np.minimum(np.maximum(mba * ey, x0), x1)
    f1 = -0.5 * (a * ex * ex + c * dy1 * dy1) - b * ex * dy1
    f2 = -0.5 * (a * dx2 * dx2 + c * ey * ey) - b * dx2 * ey
    return np.maximum(f1, f2)

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    ntiles = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    sc, cfg = scene.make_config_scene(name)
    W, H = cfg["W"], cfg["H"]
    cam = camera.default_camera(W, H)
    ref = gsr_oracle.forward_scene(sc, cam)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    ranges = ref["ranges"].reshape(-1, 2).astype(np.int64)
    vals = ref["values"]
    m2d = ref["means2D"].reshape(-1, 2); co = ref["conic_opacity"].reshape(-1, 4)
    ncontrib = ref["n_contrib"].reshape(H, W); finalT = ref["final_T"].reshape(H, W)
    rng = np.random.default_rng(0)
    lens = ranges[:, 1] - ranges[:, 0]
    tiles = rng.choice(np.nonzero(lens > 0)[0], size=min(ntiles, int((lens > 0).sum())), replace=False)
    tot = dict(A=0, B=0, C=0, pairs=0, A_noterm=0, B_union=0)
    for t in tiles:
        tx, ty = t % gx, t // gx
        ids = vals[ranges[t, 0]:ranges[t, 1]]
        n = len(ids)
        x = m2d[ids, 0].astype(np.float64); y = m2d[ids, 1].astype(np.float64)
        a = co[ids, 0].astype(np.float64); b = co[ids, 1].astype(np.float64); c = co[ids, 2].astype(np.float64)
        o = co[ids, 3].astype(np.float64)
        ok = o >= 1 / 255.0
        thr = -np.log(np.maximum(255.0 * o, 1e-30))
        # 4x4 sub-rectangles: index (sy, sx), sy, sx in 0..3
        vis = np.zeros((4, 4, n), dtype=bool)
        stop = np.zeros((4, 4), dtype=np.int64)  # list position after which the sub-rect is finished
        for sy in range(4):
            for sx in range(4):
                X, Y = tx * 16 + 4 * sx, ty * 16 + 4 * sy
                mp = max_power_box(a, b, c, x - (X + 3), x - X, y - (Y + 3), y - Y)
                vis[sy, sx] = ok & (mp >= thr)
                ys, xs = slice(Y, min(Y + 4, H)), slice(X, min(X + 4, W))
                if ys.start >= H or xs.start >= W:
                    stop[sy, sx] = 0
                    continue
                fT = finalT[ys, xs]; nc = ncontrib[ys, xs]
                stop[sy, sx] = (nc.max() + 1) if (fT < 0.011).all() else n
        idx = np.arange(n)
        tot["pairs"] += n
        # A: warp w -> (w&1, w>>1) 8x4 = 4x4 cells (sy=w>>1, sx=2*(w&1)+{0,1})
        for w in range(8):
            sy, sx0 = w >> 1, 2 * (w & 1)
            u = vis[sy, sx0] | vis[sy, sx0 + 1]
            st = max(stop[sy, sx0], stop[sy, sx0 + 1])
            tot["A"] += int((u & (idx < st)).sum()); tot["A_noterm"] += int(u.sum())
            nl = int((vis[sy, sx0] & (idx < stop[sy, sx0])).sum()); nr = int((vis[sy, sx0 + 1] & (idx < stop[sy, sx0 + 1])).sum())
            tot["B"] += max(nl, nr)
        # C: warp -> 8x8 = cells (2*wy+{0,1}, 2*wx+{0,1})
        for wy in range(2):
            for wx in range(2):
                u = np.zeros(n, dtype=bool); st = 0
                for dy in range(2):
                    for dx in range(2):
                        u |= vis[2 * wy + dy, 2 * wx + dx]; st = max(st, stop[2 * wy + dy, 2 * wx + dx])
                tot["C"] += int((u & (idx < st)).sum())
    P = tot["pairs"]
    print("%s: %d tiles, %d staged pairs" % (name, len(tiles), P))
    print("A (8x4 per warp)       trips/pair %.3f   (no termination: %.3f)" % (tot["A"] / P, tot["A_noterm"] / P))
    print("B (2 x 4x4 half-warps) trips/pair %.3f   ratio B/A %.3f" % (tot["B"] / P, tot["B"] / tot["A"]))
    print("C (8x8, 2 px/thread)   trips/pair %.3f   ratio C/A %.3f  (x36/30 slots: %.3f)" % (tot["C"] / P, tot["C"] / tot["A"], tot["C"] / tot["A"] * 1.2))

if __name__ == "__main__":
    main()

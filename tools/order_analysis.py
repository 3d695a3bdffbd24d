"""Offline (CPU) analysis for the next k-NN step: how large is the shared staging region of a group of consecutive
queries under different internal orders?  Region = bounding box of the group's cells (1 m grid) grown by one cell;
rows = (ny)(nz) x-rows, points = map points inside.  A warp-cooperative staging needs rows <= 64 and points <= ~448."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ma-lio_b200"))
import numpy as np
from malio_b200 import synth

def spread(v, bits):
    out = np.zeros_like(v)
    for b in range(bits):
        out |= ((v >> b) & 1) << (2 * b)
    return out

def hilbert2(x, y, bits):
    d = np.zeros_like(x)
    x = x.copy(); y = y.copy()
    s = 1 << (bits - 1)
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64); ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        swap = ry == 0
        flip = swap & (rx == 1)
        x = np.where(flip, s - 1 - x, x); y = np.where(flip, s - 1 - y, y)
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        s >>= 1
    return d

def qR(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

case = synth.case_C2()
x = case.x_prop
pts = case.pts
P = pts["xyz"].astype(np.float64)
lid = pts["lidar"].astype(int)
Rs, ps = qR(np.array(x.rot)), np.array(x.pos)
W = np.zeros_like(P)
for l in range(case.n_lidar):
    m = lid == l
    RE, tE = qR(np.array(x.ext[l].q)), np.array(x.ext[l].t)
    a = P[m] @ RE.T + tE
    if l > 0:
        RC, tC = qR(np.array(case.temporal_comp[l - 1]["q"])), np.array(case.temporal_comp[l - 1]["t"])
        a = a @ RC.T + tC
    W[m] = a @ Rs.T + ps
W = W.astype(np.float32)
M = case.map_xyz
h = 1.0
o = M.min(axis=0)
mc = np.floor((M - o) / h).astype(np.int64)
dims = mc.max(axis=0) + 1
# 3-D prefix sums of the cell counts: points inside any cell box in O(1)
cnt = np.zeros(tuple(dims + 1), np.int64)
np.add.at(cnt, (mc[:, 0] + 1, mc[:, 1] + 1, mc[:, 2] + 1), 1)
cs = cnt.cumsum(0).cumsum(1).cumsum(2)
def box_count(lo, hi):   # inclusive cell boxes, arrays [G,3]
    lo = np.clip(lo, 0, dims - 1); hi = np.clip(hi, 0, dims - 1) + 1
    x0, y0, z0 = lo[:, 0], lo[:, 1], lo[:, 2]; x1, y1, z1 = hi[:, 0], hi[:, 1], hi[:, 2]
    return (cs[x1, y1, z1] - cs[x0, y1, z1] - cs[x1, y0, z1] - cs[x1, y1, z0] + cs[x0, y0, z1] + cs[x0, y1, z0] + cs[x1, y0, z0] - cs[x0, y0, z0])
qc = np.floor((W - o) / h).astype(np.int64)

def report(name, key, group):
    order = np.argsort(key, kind="stable")
    c = qc[order]
    n = (len(c) // group) * group
    g = c[:n].reshape(-1, group, 3)
    lo = g.min(axis=1) - 1; hi = g.max(axis=1) + 1
    span = hi - lo + 1
    rows = span[:, 1] * span[:, 2]
    pts_in = box_count(lo, hi)
    fit = (rows <= 64) & (pts_in <= 448)
    print(f"{name:44s} group {group:2d}: rows median {int(np.median(rows)):4d} p90 {int(np.percentile(rows, 90)):5d} | points median {int(np.median(pts_in)):5d} "
          f"p90 {int(np.percentile(pts_in, 90)):6d} | fit (rows<=64, pts<=448) {100 * fit.mean():5.1f} %  | fit (rows<=32, pts<=256) {100 * ((rows <= 32) & (pts_in <= 256)).mean():5.1f} %")

for cell in (2.0, 1.0, 0.5):
    k = np.floor(W / cell).astype(np.int64)
    kx, ky, kz = k[:, 0] - k[:, 0].min(), k[:, 1] - k[:, 1].min(), k[:, 2] - k[:, 2].min()
    bits = int(np.ceil(np.log2(max(kx.max(), ky.max()) + 1)))
    mort = (spread(ky, bits) << 1) | spread(kx, bits)
    hil = hilbert2(kx, ky, bits)
    for group in (32, 16, 8):
        report(f"cell {cell} m  z-major | Morton(x,y)", (kz << (2 * bits)) | mort, group)
        report(f"cell {cell} m  Morton(x,y) | z-minor", (mort << 8) | kz, group)
        report(f"cell {cell} m  Hilbert(x,y) | z-minor", (hil << 8) | kz, group)
    print()

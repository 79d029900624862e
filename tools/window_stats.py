"""Slot-count model of the density pass's candidate walk (CPU, numpy): how many candidate SLOTS a warp
executes under different enumeration schemes, for a jittered lattice at the reference spawn gap.

  python tools/window_stats.py [n_side] [xsub]

Used to decide between walk schemes before writing a kernel; not part of the product or the tests.
"""
import sys
import numpy as np

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
xsub = int(sys.argv[2]) if len(sys.argv) > 2 else 4
r = 0.35
gap = 0.215
rng = np.random.default_rng(1)
g = np.arange(ns) * gap + 1.0
P = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + (rng.random((ns ** 3, 3)) - 0.5) * 0.2 * gap
n = len(P)
cx = np.floor(P[:, 0] / r * xsub).astype(np.int64)
cy = np.floor(P[:, 1] / r).astype(np.int64)
cz = np.floor(P[:, 2] / r).astype(np.int64)
nxf, nyc, nzc = cx.max() + 2 * xsub, cy.max() + 2, cz.max() + 2
key = (cz * nyc + cy) * nxf + cx
order = np.argsort(key, kind="stable")
P, cx, cy, cz, key = P[order], cx[order], cy[order], cz[order], key[order]
table = np.searchsorted(key, np.arange(nzc * nyc * nxf + 1))

w = r * (1 + 1e-5)
lo = np.floor((P[:, 0] - w) / r * xsub).astype(np.int64)
hi = np.floor((P[:, 0] + w) / r * xsub).astype(np.int64)
ccx = np.floor(P[:, 0] / r).astype(np.int64)
lo = np.clip(np.maximum(lo, (ccx - 1) * xsub), 0, nxf - 1)
hi = np.clip(np.minimum(hi, (ccx + 2) * xsub - 1), 0, nxf - 1)
ylo, yhi = P[:, 1] - cy * r, (cy + 1) * r - P[:, 1]
zlo, zhi = P[:, 2] - cz * r, (cz + 1) * r - P[:, 2]
B = np.zeros((n, 9), np.int64)
E = np.zeros((n, 9), np.int64)
for rr in range(9):
    dz, dy = rr // 3 - 1, rr % 3 - 1
    ok = np.ones(n, bool)
    if dy and dz:
        a = ylo if dy < 0 else yhi
        b = zlo if dz < 0 else zhi
        ok = a * a + b * b <= w * w
    z, y = cz + dz, cy + dy
    ok &= (z >= 0) & (z < nzc) & (y >= 0) & (y < nyc)
    row = (np.clip(z, 0, nzc - 1) * nyc + np.clip(y, 0, nyc - 1)) * nxf
    B[:, rr] = np.where(ok, table[row + lo], 0)
    E[:, rr] = np.where(ok, table[row + hi + 1], 0)
L = E - B
# interior particles only (full neighbourhoods) -> warps whose particles are all interior
inter = np.all((P > g[2]) & (P < g[-3]), axis=1)
nw = n // 32
Lw = L[: nw * 32].reshape(nw, 32, 9)
Bw = B[: nw * 32].reshape(nw, 32, 9)
Ew = E[: nw * 32].reshape(nw, 32, 9)
iw = inter[: nw * 32].reshape(nw, 32).all(axis=1)
Lw, Bw, Ew = Lw[iw], Bw[iw], Ew[iw]
print("interior warps", iw.sum(), "candidates/particle", Lw.sum(2).mean())
# (a) current: per row, warp max rounded up to 4 (segments of 16 checked, groups of 4)
cur = (np.ceil(Lw.max(1) / 4) * 4).sum(1)
print("(a) per-row warp max, groups of 4   : slots/lane %.1f" % cur.mean())
# (b) flat, single candidates
print("(b) flat, 1 candidate/iter          : slots/lane %.1f" % Lw.sum(2).max(1).mean())
# (c) flat, pair-aligned (even start), 1 pair per iteration
pairs = (np.ceil(Ew / 2) - np.floor(Bw / 2)) * (Lw > 0)
print("(c) flat, aligned pairs, 1 pair/iter: slots/lane %.1f (pairs %.1f)" % (2 * pairs.sum(2).max(1).mean(), pairs.sum(2).max(1).mean()))
print("    own pairs per lane avg %.1f" % pairs.sum(2).mean())
# (d) per-row warp max of aligned pairs
print("(d) per-row warp max of aligned pairs: slots/lane %.1f" % (2 * pairs.max(1).sum(1).mean()))
# (e) flat aligned quads (4 candidates = 64 B)
quads = (np.ceil(Ew / 4) - np.floor(Bw / 4)) * (Lw > 0)
print("(e) flat, aligned quads             : slots/lane %.1f" % (4 * quads.sum(2).max(1).mean()))
# distinct 128-byte lines touched per warp-load in scheme (a) (16 B rows) ~ span of the B's
span = (Bw.max(1) - Bw.min(1))
print("row-start span within a warp (rows of 16 B): mean %.1f -> %.1f lines of 128 B" % (span.mean(), (span.mean() * 16) / 128 + 1))

"""Host model of the run-based algorithm implemented in csrc/frame.cu (test infrastructure).

It mirrors, step by step and in plain Python, what the kernels do on runs -- start masks, raster
numbering, the "walk the runs of each raster-predecessor row that intersect [a-1, b+1]" adjacency
rule of the link and pair kernels, the two-directional z-face enumeration of k_tile_pairs, the
min-root union-find numbering and the periodic-component numbering of
k_label_lut -- so the *rules* can be validated against the oracle on the CPU box, where no
kernel can run.  ``shortcuts=True`` additionally applies the three rules the kernels use to skip work:
mirror rows link run m to run m without a search (k_strip_link / k_tile_link_x), mirror rows emit no
cross-row contacts, and contacts come from voxel-on-voxel overlaps only (k_tile_pairs).
It is not a fallback: the product never imports it.
"""
import numpy as np


def runs_of(grid):
    """(runstart lin index array incl. sentinel, rowoff)."""
    X, Y, Z = grid.shape
    g = grid.reshape(X * Y, Z)
    start = np.ones_like(g, dtype=bool)
    start[:, 1:] = g[:, 1:] != g[:, :-1]
    rowcnt = start.sum(1)
    rowoff = np.concatenate([[0], np.cumsum(rowcnt)]).astype(np.int64)
    runstart = np.flatnonzero(start.reshape(-1)).astype(np.int64)
    return np.concatenate([runstart, [X * Y * Z]]), rowoff


class UF:
    def __init__(self, n):
        self.p = list(range(n))

    def find(self, v):
        while self.p[v] != v:
            self.p[v] = self.p[self.p[v]]
            v = self.p[v]
        return v

    def union(self, a, b):
        a, b = self.find(a), self.find(b)
        if a == b:
            return
        if a < b:
            a, b = b, a
        self.p[a] = b


def run_at(runstart, rowoff, Z, row, z):
    lo, hi = rowoff[row], rowoff[row + 1]
    target = row * Z + z
    return lo + int(np.searchsorted(runstart[lo:hi], target, side="right")) - 1


BACKWARD = ((-1, -1), (-1, 0), (-1, 1), (0, -1))     # raster-order predecessors among the 8 neighbour rows


def rows_mirror(runstart, rowoff, flat, Z, r1, r2):
    """Same run count, same first value, interleaving starts (csrc/frame.cu: rows_mirror)."""
    a0, a1, b0, b1 = rowoff[r1], rowoff[r1 + 1], rowoff[r2], rowoff[r2 + 1]
    n = a1 - a0
    if b1 - b0 != n or flat[r1 * Z] != flat[r2 * Z]:
        return False
    za = [int(runstart[a0 + m]) - r1 * Z for m in range(n)]
    zb = [int(runstart[b0 + m]) - r2 * Z for m in range(n)]
    return all(zb[m] < za[m + 1] and za[m] < zb[m + 1] for m in range(n - 1))


def label_runs(grid, shortcuts=False):
    X, Y, Z = grid.shape
    flat = grid.reshape(-1)
    runstart, rowoff = runs_of(grid)
    R = len(runstart) - 1
    uf = UF(R)
    for i in range(R):
        lin = int(runstart[i]); r, a = divmod(lin, Z); x, y = divmod(r, Y); v = flat[lin]
        b = int(runstart[i + 1]) - r * Z - 1
        lo, hi = max(a - 1, 0), min(b + 1, Z - 1)
        for dx, dy in BACKWARD:
            nx, ny = x + dx, y + dy
            if nx < 0 or ny < 0 or ny >= Y:
                continue
            r2 = nx * Y + ny
            if shortcuts and rows_mirror(runstart, rowoff, flat, Z, r, r2):
                uf.union(i, rowoff[r2] + (i - rowoff[r]))      # the partner of run m is run m
                continue
            j = run_at(runstart, rowoff, Z, r2, lo)
            if flat[r2 * Z + lo] != v:
                j += 1
            while j < rowoff[r2 + 1] and runstart[j] <= r2 * Z + hi:
                uf.union(i, j)
                j += 2
    lab = np.zeros(R, dtype=np.int64)
    nt = nf = 0
    rank = {}
    for i in range(R):
        if uf.find(i) == i:
            if flat[runstart[i]]:
                nt += 1; rank[i] = nt
            else:
                nf += 1; rank[i] = -nf
    for i in range(R):
        lab[i] = rank[uf.find(i)]
    return runstart, rowoff, lab, nt, nf


def expand(grid, runstart, lab):
    out = np.empty(grid.size, dtype=np.int32)
    for i in range(len(lab)):
        out[runstart[i]:runstart[i + 1]] = lab[i]
    return out.reshape(grid.shape)


def pairs(grid, runstart, rowoff, lab, periodic=True, shortcuts=False):
    X, Y, Z = grid.shape
    flat = grid.reshape(-1)
    bit = lambda row, z: flat[row * Z + z]
    contacts, bridges = set(), set()

    def emit_c(a, b):
        if a != b:
            contacts.add((min(a, b), max(a, b)))

    def emit_b(la, lb, s):
        neg = tuple(-c for c in s)
        if la < lb:
            bridges.add((la, lb) + s)
        elif la > lb:
            bridges.add((lb, la) + neg)
        else:
            bridges.add((la, la) + s); bridges.add((la, la) + neg)

    R = len(lab)
    for i in range(R):
        lin = int(runstart[i]); r, a = divmod(lin, Z); x, y = divmod(r, Y); v = flat[lin]; la = int(lab[i])
        b = int(runstart[i + 1]) - r * Z - 1
        if a > 0:
            emit_c(int(lab[i - 1]), la)
        lo, hi = max(a - 1, 0), min(b + 1, Z - 1)
        for dx, dy in BACKWARD:
            nx, ny, sx, sy = x + dx, y + dy, 0, 0
            if nx < 0: nx, sx = X - 1, -1
            if ny < 0: ny, sy = Y - 1, -1
            elif ny >= Y: ny, sy = 0, 1
            wraps = bool(sx or sy)
            if wraps and not periodic:
                continue
            r2 = nx * Y + ny
            if shortcuts and not wraps:
                # mirror rows: every cross-row contact is one of the in-row contacts; otherwise only the runs of
                # the other value that overlap [a, b] voxel on voxel (a diagonal touch is an in-row contact of r2)
                if rows_mirror(runstart, rowoff, flat, Z, r, r2):
                    continue
                j = run_at(runstart, rowoff, Z, r2, a)
                if flat[r2 * Z + a] == v:
                    j += 1
                while j < rowoff[r2 + 1] and runstart[j] <= r2 * Z + b:
                    emit_c(la, int(lab[j])); j += 2
                continue
            j = run_at(runstart, rowoff, Z, r2, lo)
            if not wraps:
                if flat[r2 * Z + lo] == v:
                    j += 1
                while j < rowoff[r2 + 1] and runstart[j] <= r2 * Z + hi:
                    emit_c(la, int(lab[j])); j += 2
            else:
                while j < rowoff[r2 + 1] and runstart[j] <= r2 * Z + hi:
                    emit_b(la, int(lab[j]), (sx, sy, 0)); j += 1
        # z faces, both directions, so that only planes x and x-1 are consulted (k_tile_pairs)
        if periodic and b == Z - 1:                       # +z step from the last run into first runs, dx in {0,-1}
            for dx in (0, -1):
                nx, sx = x + dx, 0
                if nx < 0: nx, sx = X - 1, -1
                for dy in (-1, 0, 1):
                    ny, sy = y + dy, 0
                    if ny < 0: ny, sy = Y - 1, -1
                    elif ny >= Y: ny, sy = 0, 1
                    emit_b(la, int(lab[rowoff[nx * Y + ny]]), (sx, sy, 1))
        if periodic and a == 0:                           # -z step from the first run into last runs, dx = -1
            nx, sx = x - 1, 0
            if nx < 0: nx, sx = X - 1, -1
            for dy in (-1, 0, 1):
                ny, sy = y + dy, 0
                if ny < 0: ny, sy = Y - 1, -1
                elif ny >= Y: ny, sy = 0, 1
                emit_b(la, int(lab[rowoff[nx * Y + ny + 1] - 1]), (sx, sy, -1))
    c = np.array(sorted(contacts), dtype=np.int32).reshape(-1, 2)
    b = np.array(sorted(bridges), dtype=np.int32)
    return c, b


def component_lut(bridges, nt, nf):
    """lut[label + nf] -> component id, numbering rule of SURVEY.md A.6."""
    n = nt + nf + 1
    uf = UF(n)
    for row in np.asarray(bridges).reshape(-1, 5):
        l1, l2 = int(row[0]), int(row[1])
        if l1 != l2 and (l1 < 0) == (l2 < 0):
            uf.union(l1 + nf, l2 + nf)
    lut = np.zeros(n, dtype=np.int32)
    kn = kp = 0
    rank = {}
    for j in range(n):
        if j != nf and uf.find(j) == j:
            if j < nf:
                kn += 1; rank[j] = -kn
            else:
                kp += 1; rank[j] = kp
    for j in range(n):
        if j != nf:
            lut[j] = rank[uf.find(j)]
    return lut, kp, kn

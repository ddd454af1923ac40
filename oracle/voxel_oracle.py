"""CPU oracle for the voxel hot path -- TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy (+ plain C, ``oracle.c``) restatement of what the reference computes on
the path  bin -> morph -> label -> contacts -> bridges -> components -> counts ->
voxel/atom gathers.  The product package never imports this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs do,
and only as the checker / the timed CPU baseline.

Parity pin: every function here is checked against the UNMODIFIED reference run in the
build container (``tests/test_oracle_vs_reference.py`` via ``oracle/refshim.py``), against
the golden fixtures dumped from it (``tests/golden/``) and against the known-answer table
of SURVEY.md Appendix D.  The connected-component labelling itself lives in a third-party
dependency of the reference (``scipy.ndimage.label``, ``scipy>=1.15`` per the reference's
``setup.py:20``; this image has scipy 1.18.1): ``label_3d_grid`` calls it exactly as the
reference does and ``label_3d_grid_c`` restates its result (raster-ordered flood fill).

All citations are relative to ``/root/reference/mdvcontainment/``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


# ---------------------------------------------------------------------------
# C half
# ---------------------------------------------------------------------------
def build_clib(force: bool = False) -> str:
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so, "-lm"])
    return so


def clib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_clib())
        P, I64 = ctypes.c_void_p, ctypes.c_int64
        lib.orc_bin_atoms.argtypes = [P, I64, P, P, P, P, P]
        lib.orc_bin_atoms.restype = None
        lib.orc_label.argtypes = [P, I64, I64, I64, P, P, P]
        lib.orc_label.restype = ctypes.c_int
        lib.orc_contacts.argtypes = [P, I64, I64, I64, P, I64]
        lib.orc_contacts.restype = I64
        lib.orc_bridges.argtypes = [P, I64, I64, I64, P, I64]
        lib.orc_bridges.restype = I64
        lib.orc_morph_step.argtypes = [P, P, I64, I64, I64, ctypes.c_char]
        lib.orc_morph_step.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------
# A1  binning  (atomgroup_to_voxels.py:6-30, 94-144, 168-204)
# ---------------------------------------------------------------------------
def dim2lattice(x, y, z, alpha=90, beta=90, gamma=90):
    """Box lengths/angles -> lower-triangular lattice (rows = box vectors).

    atomgroup_to_voxels.py:21-30.  The arguments arrive as np.float32 scalars
    (``atomgroup.dimensions``), so the trigonometry is evaluated in float32 and only then
    stored into a float64 matrix -- that artefact (cos 90 deg = -4.37e-8) is part of the
    contract."""
    ca, cb, cg = (np.cos(np.pi * ang / 180) for ang in (alpha, beta, gamma))
    sg = np.sin(np.pi * gamma / 180)
    zx = z * cb
    zy = z * (ca - cb * cg) / sg
    zz = np.sqrt(z ** 2 - zx ** 2 - zy ** 2)
    return np.array([x, 0, 0, y * cg, y * sg, 0, zx, zy, zz]).reshape((3, 3))


def voxel_transform(dimensions, resolution):
    """(box, nbox, T, invT) exactly as atomgroup_to_voxels.py:113-118,131,142 derive them."""
    resolution = abs(resolution)
    box = dim2lattice(*dimensions)
    nbox = (box / (10 * resolution)).round().astype(int)
    T = np.linalg.inv(box) @ nbox
    invT = np.linalg.inv(T)
    return box, nbox, T, invT


def resolution_deviation(dimensions, resolution):
    """atomgroup_to_voxels.py:118-119 (only consulted when ``max_offset == True``, :112)."""
    box, nbox, _, _ = voxel_transform(dimensions, resolution)
    unit = np.linalg.inv(nbox) @ box
    return (0.1 * (unit ** 2).sum(axis=1) ** 0.5 - abs(resolution)) / abs(resolution)


def voxelate_numpy(positions, dimensions, resolution):
    """Literal NumPy evaluation of atomgroup_to_voxels.py:131-142.

    Returns (voxels int64 [n,3], nbox, new_positions float32 [n,3])."""
    _, nbox, T, _ = voxel_transform(dimensions, resolution)
    fr = np.asarray(positions, dtype=np.float32) @ T
    vox = np.floor(fr).astype(int)
    fr -= vox
    for dim in (2, 1, 0):
        s = vox[:, dim] // nbox[dim, dim]
        vox -= s[:, None] * nbox[dim, :]
    newpos = ((vox + fr) @ np.linalg.inv(T)).astype(np.float32)
    return vox, nbox, newpos


def voxelate_fma(positions, dimensions, resolution):
    """Same as :func:`voxelate_numpy` with the float64 products as an explicit FMA chain
    (``oracle.c: orc_bin_atoms``) -- host-BLAS independent, the form the CUDA kernel uses."""
    _, nbox, T, invT = voxel_transform(dimensions, resolution)
    pos = np.ascontiguousarray(positions, dtype=np.float32)
    n = len(pos)
    vox = np.empty((n, 3), dtype=np.int64)
    newpos = np.empty((n, 3), dtype=np.float32)
    Tc = np.ascontiguousarray(T, dtype=np.float64)
    iTc = np.ascontiguousarray(invT, dtype=np.float64)
    nb = np.ascontiguousarray(nbox, dtype=np.int64)
    clib().orc_bin_atoms(_p(pos), n, _p(Tc), _p(iTc), _p(nb), _p(vox), _p(newpos))
    return vox, nbox, newpos


def occupancy(voxels, nbox):
    """atomgroup_to_voxels.py:192-197: dense bool grid of shape diag(nbox)."""
    grid = np.zeros(np.diagonal(nbox), dtype=bool)
    grid[voxels[:, 0], voxels[:, 1], voxels[:, 2]] = True
    return grid


# ---------------------------------------------------------------------------
# A3/A4 morphology  (atomgroup_to_voxels.py:32-92, 242-308)
# ---------------------------------------------------------------------------
def dilate(grid):
    """Periodic 3x3x3 binary dilation == linear_blur(v, diag(shape), 1, inplace) (:57-92, 254-257)."""
    out = np.array(grid, dtype=bool)
    for ax in range(3):
        out = out | np.roll(out, 1, axis=ax) | np.roll(out, -1, axis=ax)
    return out


def erode(grid):
    """:272-275  erosion = ~dilate(~v)."""
    return ~dilate(~np.asarray(grid, dtype=bool))


def morph(grid, morph_str="de"):
    """:278-304  left-to-right application; unknown operation -> ValueError (:303)."""
    out = np.asarray(grid, dtype=bool)
    for op in morph_str:
        if op == "d":
            out = dilate(out)
        elif op == "e":
            out = erode(out)
        else:
            raise ValueError(f"Unknown morph operation '{op}' in morph string '{morph_str}'. "
                             "Use 'd' for dilation and 'e' for erosion.")
    return out


def morph_c(grid, morph_str="de"):
    g = np.ascontiguousarray(grid, dtype=np.uint8)
    for op in morph_str:
        out = np.empty_like(g)
        if clib().orc_morph_step(_p(g), _p(out), *g.shape, op.encode()) != 0:
            raise ValueError(f"Unknown morph operation '{op}' in morph string '{morph_str}'.")
        g = out
    return g.astype(bool)


# ---------------------------------------------------------------------------
# A5 labels  (label.py:4-19)
# ---------------------------------------------------------------------------
def label_3d_grid(grid):
    """scipy.ndimage.label with the full 3x3x3 structure on the grid and on its inverse;
    False components negated; merged with np.where (label.py:7-17)."""
    from scipy import ndimage
    grid = np.asarray(grid, dtype=bool)
    st = np.ones((3, 3, 3), dtype=int)
    pos, _ = ndimage.label(grid, structure=st)
    neg, _ = ndimage.label(~grid, structure=st)
    return np.where(grid, pos, -neg)


def label_3d_grid_c(grid):
    g = np.ascontiguousarray(grid, dtype=np.uint8)
    lab = np.empty(g.shape, dtype=np.int32)
    nt, nf = ctypes.c_int32(0), ctypes.c_int32(0)
    rc = clib().orc_label(_p(g), *g.shape, _p(lab), ctypes.byref(nt), ctypes.byref(nf))
    assert rc == 0
    return lab, nt.value, nf.value


# ---------------------------------------------------------------------------
# A7 contacts  (find_label_contacts.pyx:5-77)
# ---------------------------------------------------------------------------
_BACKWARD_13 = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)
                if (dx, dy, dz) < (0, 0, 0)]


def _shifted_views(L, d):
    """Pairs (a, b) of equally shaped views with b = a displaced by d, no wrap."""
    sa, sb = [], []
    for n, k in zip(L.shape, d):
        if k == 0:
            sa.append(slice(0, n)); sb.append(slice(0, n))
        elif k < 0:
            sa.append(slice(1, n)); sb.append(slice(0, n - 1))
        else:
            sa.append(slice(0, n - 1)); sb.append(slice(1, n))
    return L[tuple(sa)], L[tuple(sb)]


def find_label_contacts(labels):
    """Sorted unique (min,max) rows of different non-zero labels that are 26-adjacent
    without wrap; int32 (K,2), (0,2) when empty (:73-77)."""
    L = np.asarray(labels)
    found = set()
    for d in _BACKWARD_13:
        a, b = _shifted_views(L, d)
        m = (a != b) & (a != 0) & (b != 0)
        if m.any():
            lo = np.minimum(a[m], b[m]); hi = np.maximum(a[m], b[m])
            found.update(map(tuple, np.unique(np.stack([lo, hi], 1), axis=0).tolist()))
    if not found:
        return np.empty((0, 2), dtype=np.int32)
    return np.array(sorted(found), dtype=np.int32)


def find_label_contacts_c(labels, cap=1 << 22):
    L = np.ascontiguousarray(labels, dtype=np.int32)
    rows = np.empty((cap, 2), dtype=np.int32)
    n = clib().orc_contacts(_p(L), *L.shape, _p(rows), cap)
    assert n <= cap
    return rows[:n].copy().reshape(-1, 2)


# ---------------------------------------------------------------------------
# A6 bridges  (find_bridges.pyx:4-89)
# ---------------------------------------------------------------------------
def find_bridges(labels):
    """Sorted unique (l1,l2,sx,sy,sz) int32 rows; empty -> shape (0,) (:89)."""
    L = np.asarray(labels)
    X, Y, Z = L.shape
    ix = np.arange(X)[:, None, None]; iy = np.arange(Y)[None, :, None]; iz = np.arange(Z)[None, None, :]
    found = set()
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                if dx == dy == dz == 0:
                    continue
                # image shift of the step a -> a+d, per axis (:52-80)
                sx = np.where(ix + dx < 0, -1, np.where(ix + dx >= X, 1, 0))
                sy = np.where(iy + dy < 0, -1, np.where(iy + dy >= Y, 1, 0))
                sz = np.where(iz + dz < 0, -1, np.where(iz + dz >= Z, 1, 0))
                crossing = np.broadcast_to((sx != 0) | (sy != 0) | (sz != 0), L.shape)
                nb = np.roll(L, (-dx, -dy, -dz), axis=(0, 1, 2))          # nb[a] = L[(a+d) mod shape]
                m = crossing & (L != 0) & (nb != 0)
                if not m.any():
                    continue
                a, b = L[m], nb[m]
                s = np.stack([np.broadcast_to(v, L.shape)[m] for v in (sx, sy, sz)], 1)
                swap = a > b                                             # (:84-87)
                l1 = np.where(swap, b, a); l2 = np.where(swap, a, b)
                s = np.where(swap[:, None], -s, s)
                rows = np.unique(np.concatenate([l1[:, None], l2[:, None], s], 1), axis=0)
                found.update(map(tuple, rows.tolist()))
    return np.array(sorted(found), dtype=np.int32)


def find_bridges_c(labels, cap=1 << 22):
    L = np.ascontiguousarray(labels, dtype=np.int32)
    rows = np.empty((cap, 5), dtype=np.int32)
    n = clib().orc_bridges(_p(L), *L.shape, _p(rows), cap)
    assert n <= cap
    if n == 0:
        return np.array([], dtype=np.int32)
    return rows[:n].copy()


# ---------------------------------------------------------------------------
# A8 components / counts  (label.py:21-38, containment_main.py:510-512)
# ---------------------------------------------------------------------------
def create_components_grid(labels, component2labels):
    """Relabel through {component: labels}; labels absent from the dict stay 0 (label.py:32-37)."""
    L = np.asarray(labels)
    lo, hi = int(L.min()), int(L.max())
    lut = np.zeros(hi - lo + 1, dtype=L.dtype)
    for comp, labs in component2labels.items():
        for lab in labs:
            if lo <= lab <= hi:
                lut[lab - lo] = comp
    return lut[L - lo]


def counts(grid):
    """dict(zip(*np.unique(grid, return_counts=True)))  (containment_main.py:510-512)."""
    return dict(zip(*np.unique(grid, return_counts=True)))


# ---------------------------------------------------------------------------
# A10 masks / positions  (containment_main.py:360-378)
# ---------------------------------------------------------------------------
def voxel_positions(components_grid, nodes):
    mask = np.isin(components_grid, nodes)
    return np.array(np.where(mask)).T


# ---------------------------------------------------------------------------
# A2/A9 voxel <-> atom mapping  (atoms_voxels_mapping.pyx:9-91)
# ---------------------------------------------------------------------------
def reference_voxel_hash(vox):
    """hash_voxel (:9-14) as the generated C evaluates it: the x product is 64-bit, the y and z
    products wrap in 32 bits before being sign-extended."""
    v = np.asarray(vox, dtype=np.int64)
    hx = v[:, 0] * 73856093
    hy = (v[:, 1] * 19349663).astype(np.int32).astype(np.int64)
    hz = (v[:, 2] * 83492791).astype(np.int32).astype(np.int64)
    return hx ^ hy ^ hz


def reference_hash_collides(vox) -> bool:
    """True if two different voxels of ``vox`` share a reference hash (then the reference's
    own mapping is wrong for them, defect B1, and parity is taken against the exact multimap)."""
    v = np.unique(np.asarray(vox, dtype=np.int64), axis=0)
    return len(np.unique(reference_voxel_hash(v))) != len(v)


def create_mapping(voxels, atom_indices):
    """The *intended* exact multimap of create_efficient_mapping_cy (:17-56): voxel tuple ->
    int32 array of positions-in-group in ascending order; identical to the reference's dict
    whenever its hash has no collision."""
    vox = np.asarray(voxels, dtype=np.int32)
    order = np.lexsort((np.arange(len(vox)), vox[:, 2], vox[:, 1], vox[:, 0]))
    sv = vox[order]
    v2a = {}
    if len(sv):
        brk = np.flatnonzero(np.any(sv[1:] != sv[:-1], axis=1)) + 1
        starts = np.concatenate([[0], brk]); ends = np.concatenate([brk, [len(sv)]])
        for s, e in zip(starts, ends):
            v2a[tuple(int(c) for c in sv[s])] = order[s:e].astype(np.int32)
    return {"atom_voxels": vox.copy(), "voxel_to_atoms": v2a,
            "atom_indices": np.asarray(atom_indices, dtype=np.int64)}


def voxels2atoms(query_voxels, mapping):
    """voxels2atomgroup_cy (:59-91): query order, then atom order; int64 atom indices."""
    out = []
    v2a = mapping["voxel_to_atoms"]
    for q in np.asarray(query_voxels, dtype=np.int32):
        hit = v2a.get((int(q[0]), int(q[1]), int(q[2])))
        if hit is not None:
            out.append(hit)
    if not out:
        return np.array([], dtype=np.int64)
    return mapping["atom_indices"][np.concatenate(out)]


def atom_components(components_grid, atom_voxels):
    """What set_betafactors (mda_containment.py:191-203) leaves in the betafactor column:
    the component id of the voxel each universe atom sits in."""
    v = np.asarray(atom_voxels)
    return np.asarray(components_grid)[v[:, 0], v[:, 1], v[:, 2]].astype(np.float64)


# ---------------------------------------------------------------------------
# Whole voxel path up to the host graph (containment_main.py:54-69)
# ---------------------------------------------------------------------------
def voxel_path(grid, use_c=True):
    labels = label_3d_grid(grid).astype(np.int32)
    if use_c:
        return labels, find_label_contacts_c(labels), find_bridges_c(labels)
    return labels, find_label_contacts(labels), find_bridges(labels)

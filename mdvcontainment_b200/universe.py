"""Minimal Universe / AtomGroup and a fixed-column ``.gro`` reader/writer.

MDAnalysis is what the reference's ``Containment`` consumes
(reference ``mda_containment.py:4,324``); it is not installable in this image, so
this module provides the small duck-typed subset of it that the voxel hot path
touches:

* ``AtomGroup``: ``.universe .ix .positions`` (float32 [n,3] in Angstrom, get/set)
  ``.dimensions`` (float32 [6], A / degrees) ``.names .resnames .tempfactors``
  ``__len__ __getitem__ __sub__ __add__``;
* ``Universe``: ``.atoms .dimensions .trajectory.frame add_TopologyAttr select_atoms``.

A real ``MDAnalysis.AtomGroup`` works with ``Containment`` just as well: only the
attributes above are used.
"""
from __future__ import annotations

import types

import numpy as np


PIN_THRESHOLD_BYTES = 4 << 20      # coordinate arrays from this size on are kept in page-locked host memory


def _host_array(shape, dtype):
    """NumPy array for per-atom data.  Large ones are backed by page-locked (pinned) host memory when a CUDA device is
    present, so that the voxel path moves them over PCIe by DMA from / to where they lie -- a pageable 122 MB
    coordinate array costs ~15 ms per host-side staging copy, three to four of them per ``Containment`` call."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if nbytes >= PIN_THRESHOLD_BYTES:
        try:
            import torch
            if torch.cuda.is_available():
                t = torch.empty(tuple(int(v) for v in shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
                return t.numpy()                    # the array keeps the tensor (and its pinned block) alive
        except Exception:                           # noqa: BLE001 - pinning is an optimisation only
            pass
    return np.empty(shape, dtype=dtype)


class Tempfactors:
    """Stand-in for ``MDAnalysis.core.topologyattrs.Tempfactors`` (values holder)."""

    attrname = "tempfactors"

    def __init__(self, values):
        self.values = np.asarray(values, dtype=np.float64)


class AtomGroup:
    """An ordered set of atom indices into a :class:`Universe`."""

    def __init__(self, universe: "Universe", ix, _sorted_unique: bool = False):
        self._u = universe
        self._ix = np.asarray(ix, dtype=np.int64).reshape(-1)
        self._range_cache = False         # decided on first use: creating a 10 M atom group should not scan it
        self._sorted_unique = _sorted_unique      # indices known to ascend strictly (selections of universe.atoms)

    @property
    def _range(self):
        """(start, stop) when the group is a contiguous index range (e.g. universe.atoms): positions are then read
        and written as slices instead of gathers."""
        if self._range_cache is False:
            ix, n = self._ix, len(self._ix)
            self._range_cache = (int(ix[0]), int(ix[0]) + n) if n and int(ix[-1]) - int(ix[0]) == n - 1 \
                and (self._sorted_unique or bool(np.all(ix[1:] - ix[:-1] == 1))) else None
        return self._range_cache

    # --- identity -------------------------------------------------------
    @property
    def universe(self):
        return self._u

    @property
    def ix(self):
        return self._ix

    @property
    def indices(self):
        return self._ix

    @property
    def atoms(self):
        return self

    @property
    def n_atoms(self):
        return len(self._ix)

    def __len__(self):
        return len(self._ix)

    # --- coordinates ----------------------------------------------------
    @property
    def positions(self):
        if self._range is not None:
            return self._u._positions[self._range[0]:self._range[1]].copy()
        return self._u._positions[self._ix]

    def _positions_view(self):
        """Writable float32 [n,3] view of this group's coordinates inside the universe's array when the group is a
        contiguous index range (no copy); None otherwise.  The voxel path copies from / to it by DMA."""
        if self._range is None:
            return None
        return self._u._positions[self._range[0]:self._range[1]]

    @positions.setter
    def positions(self, value):
        # MDAnalysis stores coordinates as float32; assignment casts.
        if self._range is not None:
            self._u._positions[self._range[0]:self._range[1]] = value
        else:
            self._u._positions[self._ix] = np.asarray(value).astype(np.float32)

    @property
    def dimensions(self):
        return self._u.dimensions

    # --- topology attributes -------------------------------------------
    @property
    def names(self):
        return self._u._names[self._ix]

    @property
    def resnames(self):
        return self._u._resnames[self._ix]

    @property
    def resids(self):
        return self._u._resids[self._ix]

    @property
    def tempfactors(self):
        if self._u._tempfactors is None:
            raise AttributeError("This Universe does not contain tempfactors information")
        if self._range is not None:
            return self._u._tempfactors[self._range[0]:self._range[1]].copy()
        return self._u._tempfactors[self._ix]

    @tempfactors.setter
    def tempfactors(self, values):
        if self._u._tempfactors is None:
            raise AttributeError("This Universe does not contain tempfactors information")
        if self._range is not None:
            self._u._tempfactors[self._range[0]:self._range[1]] = values
        else:
            self._u._tempfactors[self._ix] = values

    # --- set algebra / indexing ----------------------------------------
    def __getitem__(self, item):
        if isinstance(item, (int, np.integer)):
            return AtomGroup(self._u, self._ix[[item]])
        if isinstance(item, slice):
            return AtomGroup(self._u, self._ix[item])
        item = np.asarray(item)
        if item.size == 0:
            return AtomGroup(self._u, np.empty(0, dtype=np.int64))
        if item.dtype == bool:
            return AtomGroup(self._u, self._ix[item], _sorted_unique=self._sorted_unique)
        if self._range is not None and self._range[0] == 0:
            # indexing the whole-universe group: ix[item] == item (a gather over 10^7 indices otherwise)
            item = item.astype(np.int64, copy=False)
            n = len(self._ix)
            if item.size and (int(item.min()) < -n or int(item.max()) >= n):
                raise IndexError("atom index out of range")
            return AtomGroup(self._u, np.where(item < 0, item + n, item) if item.size and int(item.min()) < 0 else item)
        return AtomGroup(self._u, self._ix[item.astype(np.int64)])

    def __sub__(self, other):
        keep = ~np.isin(self._ix, other.ix)
        return AtomGroup(self._u, self._ix[keep])

    def __add__(self, other):
        return AtomGroup(self._u, np.concatenate([self._ix, other.ix]))

    def __eq__(self, other):
        return isinstance(other, AtomGroup) and other._u is self._u and np.array_equal(other._ix, self._ix)

    def __hash__(self):
        return id(self)

    def __repr__(self):
        return f"<AtomGroup with {len(self)} atoms>"

    def select_atoms(self, selection: str):
        return self._u._select(selection, self._ix, self._sorted_unique)


class Universe:
    """Positions (A, float32), box (float32[6]) and per-atom names."""

    def __init__(self, positions, dimensions, names=None, resnames=None, resids=None):
        src = np.asarray(positions, dtype=np.float32).reshape(-1, 3)
        self._positions = _host_array(src.shape, np.float32)
        self._positions[...] = src
        n = len(self._positions)
        dims = np.asarray(dimensions, dtype=np.float32).reshape(-1)
        if dims.size == 3:
            dims = np.concatenate([dims, np.full(3, 90.0, dtype=np.float32)])
        self.dimensions = dims.astype(np.float32)
        self._names = np.asarray(names if names is not None else ["X"] * n, dtype=object)
        self._resnames = np.asarray(resnames if resnames is not None else ["RES"] * n, dtype=object)
        self._resids = np.asarray(resids if resids is not None else np.ones(n), dtype=np.int64)
        self._tempfactors = None
        self.trajectory = types.SimpleNamespace(frame=0)
        self.atoms = AtomGroup(self, np.arange(n, dtype=np.int64), _sorted_unique=True)

    def add_TopologyAttr(self, attr, values=None):
        if isinstance(attr, str):
            if attr not in ("tempfactors", "bfactors"):
                raise ValueError(f"unsupported topology attribute {attr!r}")
            vals = np.zeros(len(self.atoms)) if values is None else values
        else:
            vals = attr.values
        self._tempfactors = np.array(vals, dtype=np.float64)

    def _adopt_tempfactors(self, values):
        """``add_TopologyAttr(Tempfactors(values))`` for a float64 array the caller hands over (no copy)."""
        values = np.asarray(values, dtype=np.float64)
        if values.shape != (len(self.atoms),):
            raise ValueError("one tempfactor per atom")
        self._tempfactors = values

    def select_atoms(self, selection: str):
        return self._select(selection, self.atoms.ix, True)

    def _select(self, selection: str, ix, sorted_unique: bool = False):
        """Tiny selection language: ``all`` | ``[not] name A B ..`` | ``[not] resname A B ..``."""
        tokens = selection.split()
        if not tokens:
            raise ValueError("empty selection")
        if tokens == ["all"]:
            return AtomGroup(self, ix, _sorted_unique=sorted_unique)
        negate = tokens[0] == "not"
        if negate:
            tokens = tokens[1:]
        if len(tokens) < 2 or tokens[0] not in ("name", "resname"):
            raise ValueError(f"unsupported selection {selection!r}")
        field = self._names if tokens[0] == "name" else self._resnames
        mask = np.isin(field[ix].astype(str), tokens[1:])
        if negate:
            mask = ~mask
        return AtomGroup(self, ix[mask], _sorted_unique=sorted_unique)


def is_atomgroup(obj) -> bool:
    """True for this module's AtomGroup and for a real MDAnalysis AtomGroup."""
    if isinstance(obj, AtomGroup):
        return True
    try:  # pragma: no cover - MDAnalysis is absent from the build image
        import MDAnalysis as mda
        return isinstance(obj, mda.core.groups.AtomGroup)
    except Exception:
        return False


# ---------------------------------------------------------------------------
# .gro I/O (fixed columns; nm on disk, Angstrom in memory like MDAnalysis)
# ---------------------------------------------------------------------------
def read_gro(path) -> Universe:
    with open(path) as fh:
        lines = fh.read().splitlines()
    n = int(lines[1])
    pos = np.empty((n, 3), dtype=np.float64)
    names, resnames, resids = [], [], []
    for i in range(n):
        ln = lines[2 + i]
        resids.append(int(ln[0:5]))
        resnames.append(ln[5:10].strip())
        names.append(ln[10:15].strip())
        pos[i, 0] = float(ln[20:28])
        pos[i, 1] = float(ln[28:36])
        pos[i, 2] = float(ln[36:44])
    box = [float(v) for v in lines[2 + n].split()]
    dims = gro_box_to_dimensions(box)
    return Universe((pos * 10.0).astype(np.float32), dims, names, resnames, resids)


def gro_box_to_dimensions(box):
    """gro box line (nm; 3 or 9 numbers) -> [lx, ly, lz, alpha, beta, gamma] (A, deg) float32."""
    if len(box) == 3:
        return np.array([box[0] * 10, box[1] * 10, box[2] * 10, 90, 90, 90], dtype=np.float32)
    xx, yy, zz, xy, xz, yx, yz, zx, zy = box
    a = np.array([xx, xy, xz]) * 10
    b = np.array([yx, yy, yz]) * 10
    c = np.array([zx, zy, zz]) * 10
    la, lb, lc = (np.linalg.norm(v) for v in (a, b, c))
    alpha = np.degrees(np.arccos(np.dot(b, c) / (lb * lc)))
    beta = np.degrees(np.arccos(np.dot(a, c) / (la * lc)))
    gamma = np.degrees(np.arccos(np.dot(a, b) / (la * lb)))
    return np.array([la, lb, lc, alpha, beta, gamma], dtype=np.float32)


def write_gro(path, universe_or_group, title="mdvcontainment_b200"):
    ag = universe_or_group.atoms
    pos = ag.positions / 10.0
    dims = ag.dimensions
    with open(path, "w") as fh:
        fh.write(f"{title}\n{len(ag):5d}\n")
        for i in range(len(ag)):
            fh.write("%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n" % (
                ag.resids[i] % 100000, str(ag.resnames[i])[:5], str(ag.names[i])[:5],
                (i + 1) % 100000, pos[i, 0], pos[i, 1], pos[i, 2]))
        fh.write("%10.5f%10.5f%10.5f\n" % (dims[0] / 10, dims[1] / 10, dims[2] / 10))


def voxels_to_universe(grid, resolution_nm=1.0) -> Universe:
    """One pseudo-atom per voxel at the voxel centre, atom name = str(voxel value).

    Debug helper with the semantics of the reference's ``voxels_to_gro``
    (``voxels_to_gro.py:6-40``), producing a Universe instead of a file."""
    grid = np.asarray(grid)
    idx = np.indices(grid.shape).reshape(3, -1).T
    pos = (idx + 0.5) * (10.0 * resolution_nm)
    names = [str(v) for v in grid.reshape(-1)]
    dims = np.array([s * 10.0 * resolution_nm for s in grid.shape] + [90, 90, 90], dtype=np.float32)
    return Universe(pos.astype(np.float32), dims, names)


def voxels_to_gro(path, arr, scale=1.0, place_in_center=True, universe=None):
    """Debug dump of a voxel grid as a ``.gro`` with one pseudo-atom per voxel, atom name = str(value) -- the
    behaviour of the reference's ``voxels_to_gro`` (``voxels_to_gro.py:6-40``): voxel edge = ``scale`` nm, or the
    universe's box divided by the grid shape when a universe is given; atoms in the voxel centre when
    ``place_in_center``; box = shape * edge, truncated to whole Angstrom like the reference's int32 cast (``:32``)."""
    arr = np.asarray(arr)
    edge = np.full(3, scale * 10.0) if universe is None else np.asarray(universe.dimensions[:3], dtype=np.float64) / arr.shape
    idx = np.indices(arr.shape).reshape(3, -1).T.astype(np.float64)
    pos = idx * edge + (edge / 2 if place_in_center else 0.0)
    dims = np.array([*(np.array(arr.shape) * edge), 90, 90, 90]).astype(np.int32).astype(np.float32)
    n = len(pos)
    write_gro(path, Universe(pos.astype(np.float32), dims, [str(v) for v in arr.reshape(-1)], ["UNK"] * n))


# ---------------------------------------------------------------------------
# .pdb I/O (fixed columns; Angstrom on disk and in memory).  Only what the voxel path needs: CRYST1 box,
# ATOM/HETATM names, residue names/ids, coordinates and the temperature factor column that
# ``Containment.set_betafactors`` fills (mda_containment.py:179-220).
# ---------------------------------------------------------------------------
def read_pdb(path) -> Universe:
    pos, names, resnames, resids, temps = [], [], [], [], []
    dims = None
    with open(path) as fh:
        for ln in fh:
            rec = ln[:6]
            if rec == "CRYST1":
                dims = np.array([float(ln[6:15]), float(ln[15:24]), float(ln[24:33]),
                                 float(ln[33:40]), float(ln[40:47]), float(ln[47:54])], dtype=np.float32)
            elif rec in ("ATOM  ", "HETATM"):
                names.append(ln[12:16].strip())
                resnames.append(ln[17:21].strip())
                rid = ln[22:26].strip()
                resids.append(int(rid) if rid.lstrip("-").isdigit() else 0)
                pos.append((float(ln[30:38]), float(ln[38:46]), float(ln[46:54])))
                tf = ln[60:66].strip() if len(ln) >= 66 else ""
                temps.append(float(tf) if tf else 0.0)
            elif rec.startswith("ENDMDL"):
                break                                   # first model only
    if dims is None:
        raise ValueError(f"{path}: no CRYST1 record (a periodic box is required)")
    u = Universe(np.asarray(pos, dtype=np.float32).reshape(-1, 3), dims, names, resnames, resids)
    u.add_TopologyAttr("tempfactors", np.asarray(temps, dtype=np.float64))
    return u


def write_pdb(path, universe_or_group, title="mdvcontainment_b200"):
    ag = universe_or_group.atoms
    pos, dims = ag.positions, ag.dimensions
    try:
        temps = ag.tempfactors
    except AttributeError:
        temps = np.zeros(len(ag))
    with open(path, "w") as fh:
        fh.write(f"TITLE     {title}\n")
        fh.write("CRYST1%9.3f%9.3f%9.3f%7.2f%7.2f%7.2f P 1           1\n" % tuple(float(v) for v in dims))
        for i in range(len(ag)):
            name = str(ag.names[i])[:4]
            name = name if len(name) == 4 else " " + name.ljust(3)
            fh.write("ATOM  %5d %4s %-4s%1s%4d    %8.3f%8.3f%8.3f%6.2f%6.2f\n" % (
                (i + 1) % 100000, name, str(ag.resnames[i])[:4], "X", int(ag.resids[i]) % 10000,
                pos[i, 0], pos[i, 1], pos[i, 2], 1.0, float(temps[i])))
        fh.write("END\n")


def load(path) -> Universe:
    """``Universe`` from a ``.gro`` or ``.pdb`` file (the stand-in for ``MDAnalysis.Universe(path)``)."""
    p = str(path).lower()
    if p.endswith(".gro"):
        return read_gro(path)
    if p.endswith(".pdb"):
        return read_pdb(path)
    raise ValueError(f"unsupported structure format: {path}")

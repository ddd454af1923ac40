"""Seam functions of the reference's ``atomgroup_to_voxels.py`` on the CUDA path.

Same names, arguments and error behaviour as the reference functions they replace (citations
relative to ``/root/reference/mdvcontainment/``); NumPy in / NumPy out at this level, device
tensors and the C-ABI underneath.  The 3x3 host-side algebra (lattice, nbox, transforms) uses the
reference's NumPy expressions verbatim in meaning -- the float32 trigonometry artefact is part of
the contract.
"""
from __future__ import annotations

import numpy as np
import torch

from . import device as dev


# ---------------------------------------------------------------------------------------------
# host-side 3x3 algebra  (atomgroup_to_voxels.py:6-30, 112-131, 142)
# ---------------------------------------------------------------------------------------------
def dim2lattice(x, y, z, alpha=90, beta=90, gamma=90):
    """Lengths/angles -> lower-triangular lattice matrix (rows are the box vectors)."""
    cos_a, cos_b, cos_g = (np.cos(np.pi * ang / 180) for ang in (alpha, beta, gamma))
    sin_g = np.sin(np.pi * gamma / 180)
    zx = z * cos_b
    zy = z * (cos_a - cos_b * cos_g) / sin_g
    zz = np.sqrt(z ** 2 - zx ** 2 - zy ** 2)
    return np.array([x, 0, 0, y * cos_g, y * sin_g, 0, zx, zy, zz]).reshape((3, 3))


def voxel_transform(dimensions, resolution):
    """(box, nbox, T, invT): T maps Angstrom positions to fractional voxel coordinates."""
    resolution = abs(resolution)
    box = dim2lattice(*dimensions)
    nbox = (box / (10 * resolution)).round().astype(int)      # 10: nm -> Angstrom
    T = np.linalg.inv(box) @ nbox
    return box, nbox, T, np.linalg.inv(T)


def _check_resolution(atomgroup, resolution, max_offset, box, nbox):
    # reference quirk kept on purpose (:112): the check only fires when max_offset == True (i.e. 1)
    if not (max_offset == True):  # noqa: E712
        return
    resolution = abs(resolution)
    unit = np.linalg.inv(nbox) @ box
    deviation = (0.1 * (unit ** 2).sum(axis=1) ** 0.5 - resolution) / resolution
    if (np.abs(deviation) > max_offset).any():
        raise ValueError(
            'A scaling artifact has occurred of more than {}% '
            'deviation from the target resolution in frame {} was '
            'detected. You could consider increasing the '
            'resolution.'.format(max_offset, atomgroup.universe.trajectory.frame))


# ---------------------------------------------------------------------------------------------
# device-level pieces used by Containment (no host round trips for the grid)
# ---------------------------------------------------------------------------------------------
class DeviceMapping:
    """Exact voxel -> atoms multimap held on the device (replaces the dict of
    ``create_efficient_mapping_cy``, atoms_voxels_mapping.pyx:17-56; the reference's hash collisions,
    SURVEY.md Appendix B1, are not reproduced).  Dict-style access builds the reference's dict lazily."""

    def __init__(self, voxels_dev: torch.Tensor, atom_indices: np.ndarray, shape):
        self.voxels_dev = voxels_dev                      # int32 [A,3]
        self.atom_indices = np.asarray(atom_indices, dtype=np.int64)
        self.shape = tuple(int(s) for s in shape)
        self._csr = None
        self._atom_ix_dev = None
        self._dict = None

    @property
    def csr(self) -> dev.AtomCSR:
        if self._csr is None:
            self._csr = dev.AtomCSR(self.voxels_dev, self.shape)
        return self._csr

    @property
    def atom_ix_dev(self):
        if self._atom_ix_dev is None:
            self._atom_ix_dev = torch.from_numpy(self.atom_indices).to(self.voxels_dev.device)
        return self._atom_ix_dev

    # --- compatibility with the reference's mapping dict ---------------------------------------
    def _as_dict(self):
        if self._dict is None:
            vox = dev.to_numpy(self.voxels_dev)
            order = dev.to_numpy(self.csr.order)
            keys = dev.to_numpy(self.csr.keys)
            v2a = {}
            if len(order):
                brk = np.flatnonzero(keys[1:] != keys[:-1]) + 1
                starts = np.concatenate([[0], brk]); ends = np.concatenate([brk, [len(keys)]])
                for s, e in zip(starts, ends):
                    v2a[tuple(int(c) for c in vox[order[s]])] = order[s:e].astype(np.int32)
            self._dict = {"atom_voxels": vox, "voxel_to_atoms": v2a, "atom_indices": self.atom_indices}
        return self._dict

    def __getitem__(self, key):
        return self._as_dict()[key]

    def keys(self):
        return self._as_dict().keys()

    def __contains__(self, key):
        return key in self._as_dict()


def _pinned_view(atomgroup, name):
    """(numpy view, torch view) of the group's coordinates / tempfactors when they can be moved by DMA in place: the
    group exposes a contiguous writable view (this package's AtomGroup on an index range) in page-locked memory."""
    getter = getattr(atomgroup, name, None)
    view = getter() if getter is not None else None
    if view is None or not view.flags.c_contiguous or not view.flags.writeable or view.size == 0:
        return None
    t = torch.from_numpy(view)
    return t if t.is_pinned() else None


def positions_to_device(atomgroup) -> torch.Tensor:
    """float32 [n,3] CUDA tensor of the group's coordinates (stream-ordered DMA from pinned memory where possible)."""
    t = _pinned_view(atomgroup, "_positions_view")
    if t is not None:
        return t.to("cuda", non_blocking=True)
    return dev.to_device(np.ascontiguousarray(atomgroup.positions, dtype=np.float32))


def positions_from_device(atomgroup, newpos: torch.Tensor):
    """``atomgroup.positions = newpos`` (atomgroup_to_voxels.py:142) without host-side staging copies where possible."""
    t = _pinned_view(atomgroup, "_positions_view")
    if t is not None:
        t.copy_(newpos, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller may read the coordinates as soon as we return
    else:
        atomgroup.positions = dev.to_numpy(newpos)


def voxelize_on_device(atomgroup, resolution, max_offset=0.05, want_mapping=True, positions_dev=None):
    """One ``create_voxels`` call with every array kept on the device.

    Returns (BitGrid, voxels int32 [n,3] device tensor | None, rewritten positions device tensor).
    Writes the wrapped positions back to the AtomGroup, as the reference does (:142)."""
    dims = atomgroup.dimensions
    box, nbox, T, invT = voxel_transform(dims, resolution)
    _check_resolution(atomgroup, resolution, max_offset, box, nbox)
    shape = tuple(int(v) for v in np.diagonal(nbox))
    if positions_dev is None:
        positions_dev = positions_to_device(atomgroup)
    bits = dev.BitGrid(shape, words=torch.empty((shape[0], shape[1], dev.pitch_words(shape[2])), dtype=torch.int32,
                                                device=positions_dev.device))
    vox, newpos = dev.bin_grid(positions_dev, T, invT, nbox, bits, want_voxels=want_mapping, want_positions=True)
    positions_from_device(atomgroup, newpos)
    return bits, vox, newpos


# ---------------------------------------------------------------------------------------------
# seam functions (reference signatures)
# ---------------------------------------------------------------------------------------------
def create_voxels(atomgroup, resolution, max_offset=0.05, return_mapping=True):
    """atomgroup_to_voxels.py:168-204 -> (bool ndarray [X,Y,Z], mapping | None)."""
    bits, vox, _ = voxelize_on_device(atomgroup, resolution, max_offset, want_mapping=return_mapping)
    explicit = dev.to_numpy(bits.to_bool())
    if return_mapping:
        return explicit, DeviceMapping(vox, atomgroup.ix, bits.shape)
    return explicit, None


def _check_morph(morph_str):
    for operation in morph_str:
        if operation not in 'de':
            raise ValueError(f"Unknown morph operation '{operation}' in morph string '{morph_str}'. "
                             "Use 'd' for dilation and 'e' for erosion.")


def morph_bits(bits: dev.BitGrid, morph_str='de') -> dev.BitGrid:
    _check_morph(morph_str)
    return dev.morph(bits, morph_str)


def morph_voxels(voxels, morph_str='de'):
    """atomgroup_to_voxels.py:278-304: 'd' periodic 3x3x3 dilation, 'e' erosion, left to right."""
    _check_morph(morph_str)
    voxels = np.asarray(voxels)
    if not morph_str:
        return voxels
    return dev.to_numpy(dev.morph(dev.BitGrid.from_bool(voxels), morph_str).to_bool())


def dilate_voxels(voxels):
    return morph_voxels(voxels, 'd')


def erode_voxels(voxels):
    return morph_voxels(voxels, 'e')


def close_voxels(voxels):
    return morph_voxels(voxels, 'de')


def voxels2atomgroup(voxels, mapping, atomgroup):
    """atomgroup_to_voxels.py:206-240: atoms of the given voxels, query order then atom order."""
    q = np.asarray(voxels, dtype=np.int32).reshape(-1, 3)
    if len(q) == 0:
        return atomgroup.universe.atoms[[]]
    if isinstance(mapping, DeviceMapping):             # device gather over the voxel-sorted atom list
        idx = mapping.csr.gather_voxels(q, mapping.atom_ix_dev)
        return atomgroup.universe.atoms[idx] if len(idx) else atomgroup.universe.atoms[[]]
    m = mapping
    v2a, atom_indices = m['voxel_to_atoms'], m['atom_indices']
    hits = [v2a[k] for k in map(tuple, q.tolist()) if k in v2a]
    if not hits:
        return atomgroup.universe.atoms[[]]
    return atomgroup.universe.atoms[atom_indices[np.concatenate(hits)]]

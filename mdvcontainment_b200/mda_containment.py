"""``Containment`` / ``ContainmentView``: the MDAnalysis-facing API on the CUDA voxel path.

Drop-in for the reference's ``mda_containment.py`` (citations relative to
``/root/reference/mdvcontainment/``): same constructor signature, validation, prints, properties and
methods.  Binning, closing/morphing, labelling, bridges, contacts, relabelling, counts and the
voxel->atom gathers run on the device; only the 3x3 box algebra and the graph logic stay on the host.
"""
from __future__ import annotations

import numpy as np
import torch

from . import device as dev
from .atomgroup_to_voxels import DeviceMapping, morph_bits, voxelize_on_device, voxels2atomgroup
from .containment_main import VoxelContainment
from .host_graph import format_dag_structure
from .universe import Tempfactors, is_atomgroup


class ContainmentBase:
    """Shared by Containment and ContainmentView (mda_containment.py:11-265).  Subclasses set
    ``_base`` (the data owner) and ``voxel_containment``."""

    def __str__(self):
        return format_dag_structure(self.voxel_containment.containment_graph,
                                    self.voxel_containment.component_ranks, self.voxel_volumes, unit='nm^3')

    @property
    def nodes(self):
        return self.voxel_containment.nodes

    @property
    def voxel_volume(self):
        """Volume of one voxel in nm^3 (mda_containment.py:36-41)."""
        total_volume = np.prod(self.universe.dimensions[:3])
        total_voxels = np.prod(self._base._grid_shape)
        return (total_volume / total_voxels) / 1000

    @property
    def voxel_volumes(self):
        if not self._base.voxel_containment._voxel_counts:
            return False
        unit = self._base.voxel_volume
        return {key: value * unit for key, value in self.voxel_containment.voxel_counts.items()}

    @property
    def atomgroup(self):
        return self._base._atomgroup

    @property
    def universe(self):
        return self._base._universe

    @property
    def negative_atomgroup(self):
        return self._base._negative_atomgroup

    @property
    def resolution(self):
        return self._base._resolution

    @property
    def closing(self):
        return self._base._closing

    @property
    def morph(self):
        return self._base._morph

    @property
    def boolean_grid(self):
        """NumPy bool array [X,Y,Z] (unpacked from the device bit grid on first access)."""
        b = self._base
        if b._boolean_grid is None:
            b._boolean_grid = dev.to_numpy(b._bits.to_bool())
        return b._boolean_grid

    @property
    def boolean_grid_device(self):
        return self._base._bits

    @property
    def voxel2atom(self):
        return self._base._voxel2atom

    # --- atoms -----------------------------------------------------------------------------------
    _NO_MAPPING_MSG = (
        "Voxel to atomgroup transformations are not possible when using the 'no_mapping' flag.\n"
        "no_mapping is only useful to speed up generating the voxel level containment,\n"
        "for it does not create breadcrumbs to work its way back to the atomgroup.")

    def get_atomgroup_from_voxel_positions(self, voxels):
        if self._base._no_mapping:
            raise ValueError(self._NO_MAPPING_MSG)
        return voxels2atomgroup(voxels, self.voxel2atom, self.atomgroup)

    def _validate_nodes(self, nodes):
        return self.voxel_containment._validate_nodes(nodes)

    def get_atomgroup_from_nodes(self, nodes, containment=False):
        """Atoms inside the given nodes (mda_containment.py:139-177): voxel C-order, then atom order.
        One ordered device selection over the voxel-sorted atom list instead of mask -> positions ->
        per-voxel dictionary lookups."""
        if self._base._no_mapping:
            raise ValueError(self._NO_MAPPING_MSG)
        nodes = self._validate_nodes(nodes)
        if not nodes:
            return self.universe.atoms[[]]
        if containment:
            nodes = self.voxel_containment.get_downstream_nodes(nodes)
        vc = self.voxel_containment
        lo, n = vc._member_span()
        mapping: DeviceMapping = self.voxel2atom
        # component of every voxel-sorted atom straight from the run tables (no dense components grid)
        idx = mapping.csr.select(self._base._csr_components(), vc._resolve_nodes_to_original(nodes), lo, n,
                                 mapping.atom_ix_dev, per_position=True)
        if len(idx) == 0:
            return self.universe.atoms[[]]
        return self.universe.atoms[idx]

    def set_betafactors(self):
        """Component id of every universe atom into the tempfactors (mda_containment.py:179-220): one
        gather ``components[voxel_of_atom]`` instead of a full-grid pass per node."""
        if self._base._no_mapping:
            raise ValueError("set_betafactors requires no_mapping='False'.")
        vc = self.voxel_containment
        mapping: DeviceMapping = self.voxel2atom
        fr = self._base.voxel_containment._frame
        comp_dev = fr.plan.atom_values(mapping.voxels_dev, fr.value_table)      # run-table lookup per atom
        comp = dev.to_numpy(comp_dev)
        betafactors = np.zeros(len(self.universe.atoms))
        is_view = isinstance(self, ContainmentView)
        if is_view:
            node_of = {}
            for node in vc.nodes:
                for orig in vc._resolve_nodes_to_original([node]):
                    node_of[int(orig)] = node
            values = np.array([node_of.get(int(c), 0) for c in np.unique(comp)], dtype=np.float64)
            comp = values[np.searchsorted(np.unique(comp), comp)]
        if len(comp) == len(betafactors) and _is_arange(mapping.atom_indices):
            betafactors = comp if comp.dtype == np.float64 else comp.astype(np.float64)
        else:
            betafactors[mapping.atom_indices] = comp
        try:
            self.universe.atoms.tempfactors = betafactors
            if is_view:
                print('NOTE: tempfactors already set in the universe, and will be overwritten with the VIEW component ids.')
            else:
                print('NOTE: tempfactors already set in the universe, and will be overwritten with the component ids.')
        except AttributeError:
            adopt = getattr(self.universe, "_adopt_tempfactors", None)
            if adopt is not None:
                adopt(betafactors)                    # this package's Universe: takes the fresh array as it is
            else:
                self.universe.add_TopologyAttr(_tempfactors_attr(betafactors))
            if is_view:
                print('Writing VIEW component ids in the tempfactors of universe.')
            else:
                print('Writing component ids in the tempfactors of universe.')

    def node_view(self, keep_nodes=None, min_size=0):
        if keep_nodes is None:
            keep_nodes = self.voxel_containment.nodes
        else:
            keep_nodes = self.voxel_containment._validate_nodes(keep_nodes)
        if min_size > 0:
            keep_nodes = self.voxel_containment.filter_nodes_on_size(min_size)
        return ContainmentView(self._base, keep_nodes)


def _is_arange(ix) -> bool:
    """``universe.atoms.ix`` is sorted and unique, so first == 0 and last == n - 1 make it arange(n) (checking every
    element of a 10 M atom index array would cost more than the gather it saves)."""
    ix = np.asarray(ix)
    n = len(ix)
    return n == 0 or (int(ix[0]) == 0 and int(ix[-1]) == n - 1)


def _tempfactors_attr(values):
    try:  # pragma: no cover - real MDAnalysis is absent from the build image
        import MDAnalysis as mda
        return mda.core.topologyattrs.Tempfactors(values)
    except Exception:
        return Tempfactors(values)


class Containment(ContainmentBase):
    """Voxelise an AtomGroup and build its containment hierarchy (mda_containment.py:268-400)."""

    @property
    def _negative_atomgroup(self):
        # the reference computes this set difference eagerly in the constructor (mda_containment.py:361);
        # same value, evaluated lazily because it costs an np.isin over the whole universe
        if self._negative_cache is None:
            self._negative_cache = self._universe.atoms - self._atomgroup
        return self._negative_cache

    def __init__(self, atomgroup, resolution, closing=False, slab=False, morph="", max_offset=0.05,
                 verbose=False, write_structures=False, no_mapping=False, return_counts=True,
                 betafactors=True):
        assert is_atomgroup(atomgroup), "Atomgroup must be an MDAnalysis AtomGroup."
        assert type(resolution) in [float, int], "Resolution must be a float or int."
        assert type(closing) == bool, "Closing must be a boolean."
        assert type(morph) == str, "Morph must be a string."
        assert set(morph).issubset({'d', 'e'}), "Morph strings can only contain 'd' and 'e' characters."
        assert type(slab) == bool, "Slab must be a boolean."
        assert type(max_offset) in [float, int], "Max_offset must be a float or int."
        assert type(verbose) == bool, "Verbose must be a boolean."
        assert type(write_structures) == bool, "Write_structures must be a boolean (write to mask.gro)."
        assert type(no_mapping) == bool, "No_mapping must be a boolean."
        assert type(return_counts) == bool, "Return_counts must be a boolean."
        assert type(betafactors) == bool, "Betafactors must be a boolean."

        if closing:
            print("WARNING: The 'closing' parameter is deprecated and will be removed in future versions. Please use the 'morph' parameter with value 'de' instead.")
            if morph != "":
                print("WARNING: Both 'closing' and 'morph' parameters are set. The 'closing' operation will be applied before the 'morph' operations.")
        if no_mapping and betafactors:
            raise ValueError("betafactors='True' requires no_mapping='False'.")

        self._atomgroup = atomgroup
        self._resolution = resolution
        self._closing = closing
        self._morph = morph
        self._slab = slab
        self._max_offset = max_offset
        self._verbose = verbose
        self._write_structures = write_structures
        self._no_mapping = no_mapping
        self._return_counts = return_counts
        self._betafactors = betafactors
        self._universe = atomgroup.universe
        self._negative_cache = None          # universe.atoms - atomgroup, built on first access
        self._csr_comp = None                # component of every entry of the voxel-sorted atom list

        self._bits, self._voxel2atom = self._voxelize_atomgroup()
        self._grid_shape = self._bits.shape
        self._boolean_grid = None
        self.voxel_containment = VoxelContainment(
            self._bits, verbose=self._verbose, write_structures=self._write_structures,
            slab=self._slab, counts=self._return_counts)
        if self._betafactors:
            self.set_betafactors()

    @property
    def _base(self):
        """This object owns the data (a property instead of ``self._base = self``: no reference cycle, so the device
        buffers are released as soon as the object is dropped)."""
        return self

    def _csr_components(self):
        if self._csr_comp is None:
            fr = self.voxel_containment._frame
            self._csr_comp = fr.plan.values_at(self._voxel2atom.csr.keys, fr.value_table)
        return self._csr_comp

    def _voxelize_atomgroup(self):
        """mda_containment.py:378-400.  With a mapping, the whole universe is binned first (its
        voxels feed the mapping and its wrapped positions are written back), then the selection is
        binned from the *rewritten* positions -- the second pass is not idempotent."""
        if not self._no_mapping:
            atoms = self._universe.atoms
            _, vox_all, newpos_all = voxelize_on_device(atoms, self._resolution, self._max_offset, want_mapping=True)
            sel = torch.from_numpy(np.ascontiguousarray(self._atomgroup.ix)).to(newpos_all.device)
            # rows of universe.atoms are ordered by index, so ix addresses the rewritten positions
            pos_sel = newpos_all.index_select(0, _positions_in(atoms.ix, sel))
            bits, _, _ = voxelize_on_device(self._atomgroup, self._resolution, self._max_offset,
                                            want_mapping=False, positions_dev=pos_sel)
            voxel2atom = DeviceMapping(vox_all, atoms.ix, bits.shape)
        else:
            bits, _, _ = voxelize_on_device(self._atomgroup, self._resolution, self._max_offset, want_mapping=False)
            voxel2atom = None
        if self._closing:
            bits = morph_bits(bits, 'de')
        if self._morph:
            bits = morph_bits(bits, self._morph)
        return bits, voxel2atom


def _positions_in(universe_ix, sel_ix_dev):
    """Row of each selected atom inside ``universe.atoms`` (identity when ix == arange)."""
    uix = np.asarray(universe_ix)
    if len(uix) and uix[0] == 0 and uix[-1] == len(uix) - 1:
        return sel_ix_dev
    lookup = torch.from_numpy(np.argsort(uix)).to(sel_ix_dev.device)
    sorted_ix = torch.from_numpy(np.sort(uix)).to(sel_ix_dev.device)
    return lookup[torch.searchsorted(sorted_ix, sel_ix_dev)]


class ContainmentView(ContainmentBase):
    """Merged-node view on a Containment (mda_containment.py:431-531)."""

    def __init__(self, base_containment, keep_nodes):
        self._base = base_containment._base
        self.voxel_containment = self._base.voxel_containment.node_view(keep_nodes)

    def get_original_nodes(self, view_node):
        return self.voxel_containment.get_original_nodes(view_node)

    def get_view_node(self, original_node):
        return self.voxel_containment.get_view_node(original_node)

    def node_view(self, keep_nodes=None, min_size=0):
        if keep_nodes is None:
            keep_nodes = self.voxel_containment.nodes
        else:
            keep_nodes = self.voxel_containment._validate_nodes(keep_nodes)
        if min_size > 0:
            keep_nodes = list(set(keep_nodes) & set(self.voxel_containment.filter_nodes_on_size(min_size)))
        if not keep_nodes:
            raise ValueError("No valid nodes to keep in sub-view")
        originals = []
        for view_node in keep_nodes:
            originals.extend(self.get_original_nodes(view_node))
        return ContainmentView(self._base, list(set(originals)))

    def translate_from_original(self, original_nodes):
        view_nodes = [self.get_view_node(n) for n in original_nodes]
        return list(dict.fromkeys(n for n in view_nodes if n is not None))

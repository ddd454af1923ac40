"""``VoxelContainment`` / ``VoxelContainmentView`` on the CUDA voxel path.

Drop-in for the reference's ``containment_main.py`` (citations relative to
``/root/reference/mdvcontainment/``): same constructor, properties and query methods; the array
work (labelling, contacts, bridges, relabel, counts, masks, positions) runs on the device through the
C-ABI, the graph work stays on the host (``host_graph.py``).  Grids live on the device and are
copied to NumPy only when a property that promises an ndarray is read.
"""
from __future__ import annotations

import time

import networkx as nx
import numpy as np
import torch

from . import device as dev
from . import host_graph as hg
from .host_graph import format_dag_structure


class FrameResult:
    """Device-side state and small host-side results of one labelled frame."""

    __slots__ = ("bits", "plan", "header", "contacts", "bridges", "unique_labels", "label_volumes",
                 "components_dev", "labels_dev", "lut", "timings", "value_table")

    def labels(self) -> torch.Tensor:
        if self.labels_dev is None:
            self.labels_dev = torch.empty(self.bits.shape, dtype=torch.int32, device=self.bits.words.device)
            self.plan.expand(self.bits, self.labels_dev, None)
        return self.labels_dev

    def components(self) -> torch.Tensor:
        """Dense int32 components grid, written (one ``k_expand`` pass from the run tables) when first asked for:
        the DAG, ranks, counts and the per-atom queries do not need it (lazy materialisation)."""
        if self.components_dev is None:
            if self.value_table == 0:                   # slab=True: labels are the components
                self.components_dev = self.labels()
            else:
                self.components_dev = torch.empty(self.bits.shape, dtype=torch.int32, device=self.bits.words.device)
                self.plan.expand(self.bits, None, self.components_dev)
        return self.components_dev


def label_frame(bits: dev.BitGrid, periodic=True, plan: dev.FramePlan | None = None) -> FrameResult:
    """Device stages of calc_containment_graph up to the host graph (containment_main.py:54-69):
    labels (kept as per-run tables), contacts, bridges, label volumes."""
    t0 = time.perf_counter()
    fr = FrameResult()
    fr.bits = bits
    fr.plan, fr.header = dev.label_with_retry(bits, plan, periodic=periodic)
    h = fr.header
    fr.contacts = fr.plan.contacts().reshape(-1, 2)
    fr.bridges = fr.plan.bridges() if periodic else None
    fr.label_volumes = fr.plan.label_volumes()
    # np.unique(nonp_labeled_grid) (:55): labels are contiguous by construction, none is zero
    fr.unique_labels = np.concatenate([np.arange(-int(h.n_false), 0, dtype=np.int32),
                                       np.arange(1, int(h.n_true) + 1, dtype=np.int32)])
    fr.components_dev = None
    fr.labels_dev = None
    fr.lut = None
    fr.value_table = 1              # which run table holds the components: 1 = runcomp, 0 = labels (slab=True)
    fr.timings = {"device_label_pairs_s": time.perf_counter() - t0}
    return fr


def apply_component_lut(fr: FrameResult, lut, also_labels=False):
    """create_components_grid (label.py:21-38) as one LUT pass fused into the dense write.  ``lut``: int32
    [n_false + n_true + 1], index = label + n_false (``PeriodicSolution.lut``), or the reference's
    ``component2labels`` dict."""
    h = fr.header
    nf = int(h.n_false)
    if isinstance(lut, dict):
        table = np.zeros(int(h.n_true) + nf + 1, dtype=np.int32)
        for comp, labs in lut.items():
            table[np.asarray(labs, dtype=np.int64) + nf] = comp
        lut = table
    fr.lut = lut
    fr.plan.set_lut(torch.from_numpy(lut).to(fr.bits.words.device))     # per-run component table; no dense pass yet
    fr.components_dev = None
    if also_labels:
        fr.labels()


def _outer_face_labels(labels_dev: torch.Tensor):
    """Labels present on the outer faces of the two largest dimensions (rank_logic.py:212-236)."""
    shape = labels_dev.shape
    relevant = np.argsort(shape)[1:]
    faces = []
    for ax in relevant:
        faces.append(labels_dev.select(int(ax), 0).reshape(-1))
        faces.append(labels_dev.select(int(ax), shape[int(ax)] - 1).reshape(-1))
    return torch.unique(torch.cat(faces)).cpu().numpy()


def calc_containment_graph(boolean_grid, verbose=False, write_structures=False, draw_graphs=False, slab=False,
                           _want_frame=False):
    """containment_main.py:18-134.  ``boolean_grid`` may be a NumPy bool array or a device ``BitGrid``.

    Returns (containment_graph, component_contact_graph, components_grid, component_ranks, contact_graph)
    with ``components_grid`` as a NumPy int32 array (a device-resident variant is used by
    ``VoxelContainment`` through ``_want_frame``)."""
    if isinstance(boolean_grid, dev.BitGrid):
        bits = boolean_grid
    else:
        g = np.asarray(boolean_grid)
        if g.dtype != np.bool_:
            raise TypeError("VoxelContainment expects a boolean grid")
        bits = dev.BitGrid.from_bool(g)
    if verbose:
        n_true = bits.count()
        counts_boolean_grid = {}
        if bits.n_voxels - n_true:
            counts_boolean_grid[np.False_] = np.int64(bits.n_voxels - n_true)
        if n_true:
            counts_boolean_grid[np.True_] = np.int64(n_true)
        print(f'Value prevalence in boolean_grid: {counts_boolean_grid}')

    fr = label_frame(bits, periodic=not slab)
    nonp_unique_labels = fr.unique_labels
    if verbose:
        print(f'Unique labels in labeled_grid: {nonp_unique_labels}')
    if write_structures:
        _write_grid('input.gro', dev.to_numpy(bits.to_bool()))
        _write_grid('nonp_labels.gro', dev.to_numpy(fr.labels()))

    t0 = time.perf_counter()
    if not slab:
        sol = hg.solve_periodic(nonp_unique_labels, fr.contacts, fr.bridges)
        if verbose:
            c2l = sol["component2labels"]
            labels2component = {labs: comp for comp, labs in c2l.items()}
            print(f'component2labels {c2l}')
            print(f'labels2component {labels2component}')
            print(f'label2component {sol["label2component"]}')
        apply_component_lut(fr, sol.lut(int(fr.header.n_false), int(fr.header.n_true)))
        component_ranks = sol["component_ranks"]
        if verbose:
            print(f'All rank of components: {component_ranks}')
            print(f'Complement ranks: {sol["complement_ranks"]}')
            print(f'Containment status: {sol["is_contained"]}')
        # np.unique(components_grid) (:132): every component owns at least one voxel -> the DAG's node array
        containment_graph = sol.containment_dag()                  # arrays; networkx form on demand
        component_contact_graph = None                              # sol["component_contact_graph"], on demand
    else:
        labels_dev = fr.labels()
        fr.components_dev = labels_dev
        fr.value_table = 0
        fr.lut = np.arange(-int(fr.header.n_false), int(fr.header.n_true) + 1, dtype=np.int32)
        sol = hg.solve_slab(nonp_unique_labels, fr.contacts, _outer_face_labels(labels_dev))
        if write_structures:
            _write_grid('mask.gro', _slab_mask(bits.shape))
        component_ranks, is_contained = sol["component_ranks"], sol["is_contained"]
        component_contact_graph = nx.Graph(sol["label_graph"].to_networkx())
        containment_graph = hg.containment_graph(is_contained, nonp_unique_labels, component_contact_graph)
    if write_structures:
        _write_grid('components.gro', dev.to_numpy(fr.components()))
    if draw_graphs:
        print('=== NON PERIODIC LABEL CONTACT GRAPH ===')
        _draw_graph(sol["label_graph"].to_networkx())
        if not slab:
            print('=== COMPONENT CONTACTS GRAPH ===')
            _draw_graph(sol["component_contact_graph"])
    fr.timings["host_graph_s"] = time.perf_counter() - t0
    if _want_frame:
        return containment_graph, (component_contact_graph if slab else sol), fr, component_ranks, sol["label_graph"]
    if not slab:
        containment_graph, component_contact_graph = containment_graph.to_networkx(), sol["component_contact_graph"]
    return (containment_graph, component_contact_graph, dev.to_numpy(fr.components()), component_ranks,
            sol["label_graph"].to_networkx())


def _slab_mask(shape):
    relevant = np.argsort(shape)[1:]
    mask = np.zeros(shape, dtype=bool)
    for ax in relevant:
        idx = [slice(None)] * 3
        idx[ax] = 0; mask[tuple(idx)] = True
        idx[ax] = -1; mask[tuple(idx)] = True
    return mask


def _write_grid(path, arr):
    from .universe import voxels_to_universe, write_gro
    write_gro(path, voxels_to_universe(arr))


def _draw_graph(graph):
    import matplotlib.pyplot as plt
    nx.draw(graph, pos=nx.kamada_kawai_layout(graph))
    nx.draw_networkx_labels(graph, nx.kamada_kawai_layout(graph))
    plt.show()


# =================================================================================================
class VoxelContainmentBase:
    """Functionality shared by VoxelContainment and VoxelContainmentView (containment_main.py:137-450).

    Subclasses provide ``_base`` (the data owner) and ``containment_graph``."""

    def __str__(self):
        return format_dag_structure(self._format_graph(), self.component_ranks, self.voxel_counts, unit='nvoxels')

    def _format_graph(self):
        return self.containment_graph

    # --- data owned by the base --------------------------------------------------------------
    @property
    def grid(self):
        return self._base._get_grid()

    @property
    def slab(self):
        return self._base._slab

    @property
    def component_contact_graph(self):
        ccg = self._base._component_contact_graph
        if isinstance(ccg, hg.PeriodicSolution):                  # built on first use
            ccg = self._base._component_contact_graph = ccg["component_contact_graph"]
        return ccg

    @property
    def components_grid(self):
        return self._base._get_components_grid()

    @property
    def components_grid_device(self):
        """int32 CUDA tensor [X,Y,Z] (no host copy)."""
        return self._base._frame.components()

    @property
    def nonp_label_contact_graph(self):
        return self._base._get_label_graph()

    @property
    def component_ranks(self):
        return self._base._component_ranks

    @property
    def voxel_counts(self):
        if not self._base.voxel_counts:
            return False
        counts = {}
        for node in self.nodes:
            counts[node] = sum(self._base.voxel_counts.get(orig, 0) for orig in self._resolve_nodes_to_original([node]))
        return counts

    @property
    def nodes(self):
        return list(self.containment_graph.nodes)

    @property
    def root_nodes(self):
        return [n for n, d in self.containment_graph.in_degree() if d == 0]

    @property
    def leaf_nodes(self):
        return [n for n, d in self.containment_graph.out_degree() if d == 0]

    # --- node bookkeeping ----------------------------------------------------------------------
    def _validate_nodes(self, nodes):
        valid_nodes = [n for n in nodes if n in self.containment_graph]
        invalid = set(nodes) - set(valid_nodes)
        if invalid:
            print(f"Warning: Nodes {invalid} not found in current graph and will be ignored")
        return valid_nodes

    def _resolve_nodes_to_original(self, nodes):
        return nodes

    def _get_nodes(self):
        return self.nodes

    def _get_roots(self):
        return self.root_nodes

    def _get_leaves(self):
        return self.leaf_nodes

    def get_counts(self, grid=None):
        """Voxels per component; with ``grid`` given, np.unique(grid, return_counts=True) of that grid
        (containment_main.py:264-284) computed by the device histogram."""
        if grid is None:
            return self.voxel_counts if self.voxel_counts else {}
        g = np.asarray(grid) if not isinstance(grid, torch.Tensor) else grid
        lo = int(g.min())
        hi = int(g.max())
        cnt = dev.grid_count(g, lo, hi - lo + 1)
        dtype = np.int32 if isinstance(grid, torch.Tensor) else np.asarray(grid).dtype.type
        return {dtype(lo + k): np.int64(c) for k, c in enumerate(cnt) if c}

    def _get_inverted_containment_graph(self):
        if not getattr(self, '_inv_containment_graph', None):
            self._inv_containment_graph = self.containment_graph.reverse()
        return self._inv_containment_graph

    def get_child_nodes(self, start_nodes):
        out = []
        for n in start_nodes:
            if n in self.containment_graph:
                out += list(self.containment_graph.neighbors(n))
        return sorted(set(out))

    def get_downstream_nodes(self, start_nodes):
        out = []
        for n in start_nodes:
            if n in self.containment_graph:
                out += list(nx.dfs_postorder_nodes(self.containment_graph, n))[::-1]
        return sorted(set(out))

    def get_upstream_nodes(self, start_nodes):
        rev = self._get_inverted_containment_graph()
        out = []
        for n in start_nodes:
            if n in rev:
                out += list(nx.dfs_postorder_nodes(rev, n))[::-1]
        return sorted(set(out))

    def get_parent_nodes(self, start_nodes):
        rev = self._get_inverted_containment_graph()
        out = []
        for n in start_nodes:
            if n in rev:
                out += list(rev.neighbors(n))
        return sorted(set(out))

    def get_total_voxel_count(self, nodes):
        if not self.voxel_counts:
            return 0
        return sum(self.voxel_counts.get(n, 0) for n in nodes)

    # --- array queries (device) ------------------------------------------------------------------
    def _member_span(self):
        lut = self._base._frame.lut
        lo = int(lut.min())
        return lo, int(lut.max()) - lo + 1

    def get_voxel_mask(self, nodes):
        """np.isin(components_grid, nodes) (containment_main.py:360-368) -> NumPy bool array."""
        lo, n = self._member_span()
        _, mask = dev.grid_select(self._base._frame.components(), self._resolve_nodes_to_original(nodes), lo, n,
                                  want_mask=True, want_positions=False)
        return dev.to_numpy(mask)

    def get_voxel_positions(self, nodes):
        """np.where of the mask as an (M,3) int64 array in C-order (containment_main.py:370-378)."""
        lo, n = self._member_span()
        pos, _ = dev.grid_select(self._base._frame.components(), self._resolve_nodes_to_original(nodes), lo, n)
        return pos

    def format_containment(self, nodes=False):
        g = self.containment_graph if nodes is False else self.containment_graph.subgraph(nodes)
        return format_dag_structure(g, self.component_ranks, self.voxel_counts)

    def draw(self, nodes=False):
        import matplotlib.pyplot as plt
        nx.draw_networkx(self.containment_graph if nodes is False else self.containment_graph.subgraph(nodes))
        plt.show()

    def filter_nodes_on_size(self, min_size):
        keep = []
        for node in self.nodes:
            total = sum(self.voxel_counts[n] for n in self.get_downstream_nodes([node]))
            if total >= min_size:
                keep.append(int(node))
        return keep

    def node_view(self, keep_nodes=None, min_size=0):
        if keep_nodes is None:
            keep_nodes = self.nodes
        if min_size > 0:
            keep_nodes = self.filter_nodes_on_size(min_size)
        return VoxelContainmentView(self._base, keep_nodes)


class VoxelContainment(VoxelContainmentBase):
    """Owner of the containment DAG of one boolean grid (containment_main.py:453-524).

    ``grid`` is a NumPy bool array [X,Y,Z] or a device ``BitGrid``."""

    def __init__(self, grid, verbose=False, write_structures=False, draw_graphs=False, counts=True, slab=False):
        self._grid = grid
        self._verbose = verbose
        self._write_structures = write_structures
        self._draw_graphs = draw_graphs
        self._slab = slab
        (self._containment_graph, self._component_contact_graph, self._frame, self._component_ranks,
         self._label_graph) = self._calc_containment()
        self._components_grid = None
        self._nonp_label_contact_graph = None
        self._inv_containment_graph = False
        self._voxel_counts = self._compute_counts() if counts else False

    @property
    def _base(self):
        """The data owner is this object itself.  A property rather than the reference's ``self._base = self``
        (containment_main.py:478): the attribute would be a reference cycle, and a cycle keeps gigabytes of device
        buffers alive until the garbage collector happens to run."""
        return self

    def _calc_containment(self):
        return calc_containment_graph(self._grid, verbose=self._verbose, write_structures=self._write_structures,
                                      draw_graphs=self._draw_graphs, slab=self._slab, _want_frame=True)

    def _compute_counts(self, grid=None):
        """dict(zip(*np.unique(components_grid, return_counts=True))) (containment_main.py:510-512)
        without a pass over the grid: label volumes summed per component."""
        if grid is not None:
            return self.get_counts(grid)
        fr = self._frame
        ids, counts = hg.counts_by_component(fr.lut, fr.label_volumes)
        return {np.int32(k): np.int64(v) for k, v in zip(ids.tolist(), counts.tolist())}

    def _get_grid(self):
        if isinstance(self._grid, dev.BitGrid):
            return dev.to_numpy(self._grid.to_bool())
        return self._grid

    def _get_components_grid(self):
        if self._components_grid is None:
            self._components_grid = dev.to_numpy(self._frame.components())
        return self._components_grid

    def _get_label_graph(self):
        if self._nonp_label_contact_graph is None:
            self._nonp_label_contact_graph = self._label_graph.to_networkx()
        return self._nonp_label_contact_graph

    @property
    def nonp_labeled_grid_device(self):
        """Non-periodic int32 label grid as a CUDA tensor (expanded on first use)."""
        return self._frame.labels()

    @property
    def timings(self):
        return self._frame.timings

    @property
    def containment_graph(self):
        if isinstance(self._containment_graph, hg.CompactDAG):    # networkx form on first use
            self._dag = self._containment_graph
            self._containment_graph = self._dag.to_networkx()
        return self._containment_graph

    def _format_graph(self):
        return self._containment_graph                            # CompactDAG or networkx: both print

    @property
    def voxel_counts(self):
        return self._voxel_counts


class VoxelContainmentView(VoxelContainmentBase):
    """Merged-node view on a VoxelContainment (containment_main.py:527-780): nodes outside
    ``keep_nodes`` are folded into their nearest kept ancestor; data stays with the base."""

    def __init__(self, base_containment, keep_nodes):
        self._base = base_containment._base
        self._keep_nodes = set(keep_nodes)
        self._node_map = self._build_node_map()
        self._reverse_node_map = {node: [] for node in self._keep_nodes}
        for orig, view_node in self._node_map.items():
            if view_node is not None:
                self._reverse_node_map[view_node].append(orig)
        self._view_graph = self._project(self._base.containment_graph, nx.DiGraph(), reduce=True)
        self._view_component_contact_graph = self._project(self._base.component_contact_graph, nx.Graph())
        self._view_nonp_label_contact_graph = self._project(self._base.nonp_label_contact_graph, nx.Graph())
        self._inv_containment_graph = None

    @property
    def containment_graph(self):
        return self._view_graph

    @property
    def component_contact_graph(self):
        return self._view_component_contact_graph

    @property
    def nonp_label_contact_graph(self):
        return self._view_nonp_label_contact_graph

    def _project(self, source, target, reduce=False):
        target.add_nodes_from(self._keep_nodes)
        for u, v in source.edges():
            mu, mv = self._node_map.get(u), self._node_map.get(v)
            if mu and mv and mu != mv:
                target.add_edge(mu, mv)
        return nx.transitive_reduction(target) if reduce else target

    def _build_node_map(self):
        node_map = {node: node for node in self._keep_nodes}
        try:
            order = list(nx.topological_sort(self._base.containment_graph))
        except Exception:
            order = self._base.nodes
        for node in order:
            if node in self._keep_nodes:
                continue
            mapped = None
            for parent in self._base.get_parent_nodes([node]):
                if parent in self._keep_nodes:
                    mapped = parent
                    break
                if parent in node_map:
                    mapped = node_map[parent]
                    if mapped:
                        break
            node_map[node] = mapped
        return node_map

    def _resolve_nodes_to_original(self, nodes):
        out = []
        for view_node in nodes:
            out.extend(self._reverse_node_map.get(view_node, []))
        return sorted(set(out))

    def get_original_nodes(self, view_node):
        return self._reverse_node_map.get(view_node, [])

    def get_view_node(self, original_node):
        return self._node_map.get(original_node)

    def node_view(self, keep_nodes=None, min_size=0):
        if keep_nodes is None:
            keep_nodes = self.nodes
        keep_nodes = self._validate_nodes(keep_nodes)
        if min_size > 0:
            keep_nodes = list(set(keep_nodes) & set(self.filter_nodes_on_size(min_size)))
        if not keep_nodes:
            raise ValueError("No valid nodes to keep in sub-view")
        originals = []
        for view_node in keep_nodes:
            originals.extend(self.get_original_nodes(view_node))
        return VoxelContainmentView(self._base, list(set(originals)))

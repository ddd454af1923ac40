"""Slab decomposition of one oversized periodic grid over several GPUs (second mode).

The grid is cut along x (the slowest axis: slabs are contiguous blocks of Y*Z planes).  Every rank labels
its own slab with the single-GPU kernels (``mdvc_frame_label`` / ``mdvc_frame_pairs`` with y/z-only
periodicity), exchanges one boundary plane of labels with its x-neighbour (``torch.distributed`` all-gather
over NCCL: 16 MiB per face at 2048^2), finds the label pairs touching across each slab boundary
(``mdvc_plane_pairs``), and all ranks solve the small global problem redundantly on the host:

* labels of different slabs joined through a boundary become one global label, numbered like the reference
  (scipy): by the first voxel in C order = by (slab, local label) of the group's first member;
* contacts / bridges = the slabs' own rows mapped to global labels + the boundary rows; the step across the
  periodic x face (last slab -> first slab) only ever yields bridges, exactly like the single-GPU path;
* label volumes are summed over slabs.

The unchanged host graph stage then runs on the global labels, and each rank writes its part of the dense
label / component grids with one ``k_expand`` pass.  There is no reference counterpart (the reference is a
single process and cannot even represent >= 2^31 voxels); parity is established against
the single-GPU path / the oracle on grids both can handle.

``merge_slabs`` (pure NumPy) holds all of the cross-slab logic so that it is testable without a GPU.
"""
from __future__ import annotations

import ctypes

import numpy as np


def slab_extents(X: int, n_slabs: int):
    """[(x0, x1)] contiguous, sizes differing by at most one plane."""
    if n_slabs < 1 or n_slabs > X:
        raise ValueError(f"cannot cut {X} planes into {n_slabs} slabs")
    q, r = divmod(X, n_slabs)
    out, x = [], 0
    for g in range(n_slabs):
        n = q + (1 if g < r else 0)
        out.append((x, x + n))
        x += n
    return out


# ------------------------------------------------------------------------------------------------
# host-side global solve
# ------------------------------------------------------------------------------------------------
def _normalise_bridges(l1, l2, s):
    """Rows (l1, l2, s) -> the reference's normal form: (min, max, +-s); both signs for l1 == l2."""
    l1 = np.asarray(l1, dtype=np.int64); l2 = np.asarray(l2, dtype=np.int64); s = np.asarray(s, dtype=np.int64).reshape(-1, 3)
    swap = l1 > l2
    a = np.where(swap, l2, l1); b = np.where(swap, l1, l2)
    ss = np.where(swap[:, None], -s, s)
    rows = np.concatenate([a[:, None], b[:, None], ss], axis=1)
    same = a == b
    if same.any():
        rows = np.concatenate([rows, np.concatenate([a[same, None], b[same, None], -ss[same]], axis=1)], axis=0)
    return rows


def merge_slabs(n_true, n_false, boundary_rows, local_contacts, local_bridges, local_volumes):
    """Global labels, contacts, bridges and volumes from per-slab results: the compiled host routine
    ``mdvc_host_merge_slabs`` (``csrc/hostgraph.cpp``; one pass over the slabs' rows, radix-sorted packed keys).  Same
    arguments and result as ``merge_slabs_numpy`` below, which documents the rules and is kept as their NumPy statement
    (``tests/test_slab_merge.py`` holds the two against each other and against the whole-grid oracle)."""
    import ctypes
    from . import _lib
    G = len(n_true)
    nt = np.asarray([int(v) for v in n_true], dtype=np.int64)
    nf = np.asarray([int(v) for v in n_false], dtype=np.int64)

    def cat(parts, cols):
        arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1, cols)) for a in parts]
        off = np.zeros(G + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(a) for a in arrs])
        return (np.concatenate(arrs) if arrs else np.empty((0, cols), np.int32)), off

    brows, boff = cat(boundary_rows, 5)
    crows, coff = cat(local_contacts, 2)
    grows, goff = cat(local_bridges, 5)
    loff = np.zeros(G + 1, dtype=np.int64)
    loff[1:] = np.cumsum(nt + nf + 1)
    vols = np.concatenate([np.ascontiguousarray(v, dtype=np.int64).reshape(-1) for v in local_volumes])
    if len(vols) != loff[-1]:
        raise ValueError("local_volumes must hold n_false + n_true + 1 entries per slab")
    lut = np.empty(int(loff[-1]), dtype=np.int32)
    contacts = np.empty((len(crows) + len(brows), 2), dtype=np.int32)
    bridges = np.empty((2 * (len(grows) + len(brows)), 5), dtype=np.int32)
    volumes = np.empty(int((nt + nf).sum()) + 1, dtype=np.int64)
    out = [ctypes.c_longlong(0) for _ in range(4)]                # NT, NF, contacts, bridges

    def ptr(a):
        return a.ctypes.data if a.size else None

    rc = _lib.load().mdvc_host_merge_slabs(G, ptr(nt), ptr(nf), ptr(brows), ptr(boff), ptr(crows), ptr(coff), ptr(grows),
                                           ptr(goff), ptr(vols), ptr(loff), ptr(lut), ctypes.addressof(out[0]),
                                           ctypes.addressof(out[1]), ptr(contacts), ctypes.addressof(out[2]), ptr(bridges),
                                           ctypes.addressof(out[3]), ptr(volumes))
    _lib.check(rc, "mdvc_host_merge_slabs")
    NT, NF, n_c, n_b = (int(o.value) for o in out)
    return {"luts": [lut[loff[g]:loff[g + 1]] for g in range(G)], "n_true": NT, "n_false": NF,
            "contacts": contacts[:n_c], "bridges": bridges[:n_b] if n_b else np.array([], dtype=np.int32),
            "volumes": volumes[:NF + NT + 1]}


def merge_slabs_numpy(n_true, n_false, boundary_rows, local_contacts, local_bridges, local_volumes):
    """Global labels, contacts, bridges and volumes from per-slab results.

    n_true[g], n_false[g]     label counts of slab g (local labels 1..n_true, -1..-n_false)
    boundary_rows[g]          int32 [K,5] rows (la, lb, sx, sy, sz) of mdvc_plane_pairs between the last plane of
                              slab g (la) and the first plane of slab (g+1) % G (lb)
    local_contacts[g]         int32 [K,2]; local_bridges[g] int32 [K,5] (y/z faces only); local_volumes[g]
                              int64 [n_false+n_true+1] (index = label + n_false)

    Returns dict(luts=[int32 arrays, index = local label + n_false[g] -> global label], n_true, n_false,
    contacts int32 [K,2], bridges int32 [K,5] (shape (0,) when empty), volumes int64)."""
    from . import host_graph as hg
    G = len(n_true)
    nt = np.asarray([int(v) for v in n_true], dtype=np.int64)
    nf = np.asarray([int(v) for v in n_false], dtype=np.int64)
    # one node per (slab, local label), positives first, each polarity in (slab, |label|) order: the smallest node
    # index of a merged group is then the group's first voxel in C order, which is the root the union-find keeps
    base_pos = np.concatenate([[0], np.cumsum(nt)])
    base_neg = base_pos[-1] + np.concatenate([[0], np.cumsum(nf)])
    n_nodes = int(base_neg[-1])

    def node(g, lab):
        lab = np.asarray(lab, dtype=np.int64)
        return np.where(lab > 0, base_pos[g] + lab - 1, base_neg[g] - lab - 1)

    eu, ev = [], []
    for g in range(G):
        rows = np.asarray(boundary_rows[g]).reshape(-1, 5)
        h = (g + 1) % G
        # same value on both sides of an in-box step: one label
        join = ~rows[:, 2:].any(axis=1) & ((rows[:, 0] > 0) == (rows[:, 1] > 0))
        if join.any():
            eu.append(node(g, rows[join, 0])); ev.append(node(h, rows[join, 1]))
    if eu:
        root, _, _ = hg.shift_graph_ranks(n_nodes, np.concatenate(eu), np.concatenate(ev))
        root = root.astype(np.int64)
    else:
        root = np.arange(n_nodes, dtype=np.int64)
    is_root = root == np.arange(n_nodes)
    NT = int(is_root[:base_pos[-1]].sum())
    NF = int(is_root[base_pos[-1]:].sum())
    number = np.empty(n_nodes, dtype=np.int64)                   # global |label| of every node
    number[:base_pos[-1]] = np.cumsum(is_root[:base_pos[-1]])
    number[base_pos[-1]:] = np.cumsum(is_root[base_pos[-1]:])
    number = number[root]                                        # merged-away labels take their representative's number
    luts = []
    for g in range(G):
        lut = np.zeros(int(nf[g] + nt[g]) + 1, dtype=np.int32)
        lut[int(nf[g]) + 1:] = number[base_pos[g]:base_pos[g + 1]]
        lut[:int(nf[g])] = -number[base_neg[g]:base_neg[g + 1]][::-1]      # index 0 is local label -nf
        luts.append(lut)

    contacts, bridges = [], []
    volumes = np.zeros(NF + NT + 1, dtype=np.int64)
    for g in range(G):
        lut, nf = luts[g], int(n_false[g])
        c = np.asarray(local_contacts[g]).reshape(-1, 2).astype(np.int64)
        if len(c):
            m = lut[c + nf].astype(np.int64)
            contacts.append(np.stack([m.min(1), m.max(1)], 1))
        b = np.asarray(local_bridges[g]).reshape(-1, 5).astype(np.int64)
        if len(b):
            bridges.append(_normalise_bridges(lut[b[:, 0] + nf], lut[b[:, 1] + nf], b[:, 2:]))
        volumes += hg.sum_by_index(lut.astype(np.int64) + NF, local_volumes[g], NF + NT + 1)
        rows = np.asarray(boundary_rows[g]).reshape(-1, 5).astype(np.int64)
        if len(rows):
            h = (g + 1) % G
            ga = lut[rows[:, 0] + nf].astype(np.int64)
            gb = luts[h][rows[:, 1] + int(n_false[h])].astype(np.int64)
            inbox = ~rows[:, 2:].any(axis=1)
            opp = inbox & ((ga > 0) != (gb > 0))
            if opp.any():
                contacts.append(np.stack([np.minimum(ga[opp], gb[opp]), np.maximum(ga[opp], gb[opp])], 1))
            if (~inbox).any():
                bridges.append(_normalise_bridges(ga[~inbox], gb[~inbox], rows[~inbox][:, 2:]))
    volumes[NF] = 0
    # sorted unique rows through one packed int64 key per row (lexicographic order of the key == of the row);
    # np.unique(axis=0) sorts rows as structured records and takes most of a second per million rows
    R = np.int64(NF + NT + 1)
    if contacts:
        c = np.concatenate(contacts)
        key = np.unique((c[:, 0] + NF) * R + (c[:, 1] + NF))
        contacts = np.stack([key // R - NF, key % R - NF], 1).astype(np.int32)
    else:
        contacts = np.empty((0, 2), np.int32)
    if bridges and int(R) * int(R) * 27 < 2 ** 63:
        b = np.concatenate(bridges)
        code = (b[:, 2] + 1) * 9 + (b[:, 3] + 1) * 3 + (b[:, 4] + 1)
        key = np.unique(((b[:, 0] + NF) * R + (b[:, 1] + NF)) * 27 + code)
        code, pair = key % 27, key // 27
        bridges = np.stack([pair // R - NF, pair % R - NF, code // 9 - 1, code // 3 % 3 - 1, code % 3 - 1], 1).astype(np.int32)
    elif bridges:
        bridges = np.unique(np.concatenate(bridges), axis=0).astype(np.int32)
    else:
        bridges = np.array([], dtype=np.int32)
    return {"luts": luts, "n_true": NT, "n_false": NF, "contacts": contacts, "bridges": bridges, "volumes": volumes}


# ------------------------------------------------------------------------------------------------
# device side
# ------------------------------------------------------------------------------------------------
class LocalGather:
    """All slabs live in this process (one GPU emulating the ranks, or tests): gathering is the identity."""

    def __init__(self, n_slabs):
        self.n_slabs = n_slabs
        self.rank0 = True

    def owned(self):
        return list(range(self.n_slabs))

    def gather_tensors(self, local):            # local: {slab index: tensor}
        return [local[g] for g in range(self.n_slabs)]

    def gather_objects(self, local):
        return [local[g] for g in range(self.n_slabs)]


class DistGather:
    """One slab per rank of the default torch.distributed group (NCCL for tensors over NVLink)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.n_slabs = dist.get_world_size()
        self.rank = dist.get_rank()
        self.rank0 = self.rank == 0

    def owned(self):
        return [self.rank]

    def gather_tensors(self, local):
        import torch
        t = local[self.rank].contiguous()
        out = [torch.empty_like(t) for _ in range(self.n_slabs)]
        self.dist.all_gather(out, t)
        return out

    def gather_objects(self, local):
        out = [None] * self.n_slabs
        self.dist.all_gather_object(out, local[self.rank])
        return out


class SlabResult:
    """Like pipeline.FrameOutput: arrays first (``dag``, ``solution``, ``count_ids`` / ``count_values``), the
    reference's networkx / dict forms on first access."""

    def __init__(self):
        self.dag = self.solution = self.count_ids = self.count_values = None
        self.contacts = self.bridges = None
        self.n_true = self.n_false = 0
        self.labels = self.components = self.extents = self.merge = self.plans = None
        self.timings = {}
        self._graph = self._ranks = self._counts = None

    @property
    def containment_graph(self):
        if self._graph is None:
            self._graph = self.dag.to_networkx()
        return self._graph

    @property
    def component_ranks(self):
        if self._ranks is None:
            self._ranks = self.solution["component_ranks"]
        return self._ranks

    @property
    def voxel_counts(self):
        if self._counts is None:
            self._counts = {np.int32(k): np.int64(v) for k, v in zip(self.count_ids.tolist(), self.count_values.tolist())}
        return self._counts

    def __str__(self):
        from . import host_graph as hg
        return hg.format_dag_structure(self.dag, self.component_ranks, self.voxel_counts)


def _plane_pairs(dev, plane_a, plane_b, sx_boundary, n_slots=1 << 16):
    import torch
    from . import _lib
    lib = _lib.load()
    Y, Z = plane_a.shape
    while True:
        slots = torch.empty(n_slots, dtype=torch.int64, device=plane_a.device)
        rows = torch.empty((n_slots, 5), dtype=torch.int32, device=plane_a.device)
        meta = torch.zeros(2, dtype=torch.int32, device=plane_a.device)
        rc = lib.mdvc_plane_pairs(plane_a.data_ptr(), plane_b.data_ptr(), Y, Z, int(sx_boundary), 1, slots.data_ptr(), n_slots,
                                  rows.data_ptr(), n_slots, meta[0:].data_ptr(), meta[1:].data_ptr(),
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "mdvc_plane_pairs")
        n, flags = (int(v) for v in meta.cpu())
        if flags & _lib.OVERFLOW_LABELS:
            raise _lib.MdvcError("label magnitude exceeds 2^28")
        if flags == 0:
            return rows[:n].cpu().numpy()
        n_slots *= 8


def analyse_slabs(slab_bits: dict, shape, gather=None, morph_ops: str = "", want_labels=True, plans: dict | None = None):
    """Containment analysis of a periodic grid given as x-slabs of bit grids.

    slab_bits   {slab index: BitGrid of that slab's own planes} for the slabs this process owns
    shape       (X, Y, Z) of the whole grid
    gather      LocalGather(G) (default, all slabs here) or DistGather() (one slab per rank)
    morph_ops   'd'/'e' string applied to the whole periodic grid first (halo planes exchanged)
    plans       {slab index: FramePlan} of an earlier call on slabs of the same shape (``result.plans``): workspaces
                and the capacities they grew to are reused instead of being rediscovered

    Returns a SlabResult; ``labels`` / ``components`` are {slab index: int32 CUDA tensor of the slab}."""
    import time
    import torch
    from . import _lib, device as dev, host_graph as hg
    t_start = time.perf_counter()
    X, Y, Z = (int(v) for v in shape)
    gather = gather or LocalGather(len(slab_bits))
    G = gather.n_slabs
    ext = slab_extents(X, G)
    own = gather.owned()

    # --- morphology with halo planes from the x-neighbours (periodic ring) -----------------------
    if morph_ops:
        m = len(morph_ops)
        if any(x1 - x0 < m for x0, x1 in ext):
            raise ValueError("slabs thinner than the morph halo")
        lo = gather.gather_tensors({g: slab_bits[g].words[:m].contiguous() for g in own})      # first m planes of every slab
        hi = gather.gather_tensors({g: slab_bits[g].words[-m:].contiguous() for g in own})     # last m planes
        new_bits = {}
        for g in own:
            left, right = hi[(g - 1) % G], lo[(g + 1) % G]
            ext_words = torch.cat([left, slab_bits[g].words, right], dim=0).contiguous()
            ext_grid = dev.BitGrid((ext_words.shape[0], Y, Z), ext_words)
            # the sweep wraps x inside this extended array; garbage from that wrap advances one plane per
            # operation, so after m operations exactly the m halo planes on each side are unusable
            closed = dev.morph(ext_grid, morph_ops)
            new_bits[g] = dev.BitGrid((ext[g][1] - ext[g][0], Y, Z), closed.words[m:-m].contiguous())
        slab_bits = new_bits

    # --- per-slab labelling, y/z-periodic pairs, boundary planes -------------------------------
    plans_in, plans = plans, {}
    headers, first_planes, last_planes = {}, {}, {}
    local = {}
    for g in own:
        bits = slab_bits[g]
        plan = plans_in.get(g) if plans_in else None
        if plan is None or plan.shape != tuple(bits.shape):
            plan = dev.FramePlan(bits.shape, device=bits.words.device)
        for _ in range(6):
            plan.label(bits)
            plan.pairs(bits, periodic=2)
            h = plan.read_header()
            if h.flags == 0:
                break
            plan = plan.grown(h)
        else:
            raise _lib.MdvcError("slab workspace still overflowing")
        plans[g], headers[g] = plan, h
        nx = bits.shape[0]
        fp = torch.empty((Y, Z), dtype=torch.int32, device=bits.words.device)
        lp = torch.empty((Y, Z), dtype=torch.int32, device=bits.words.device)
        lib = _lib.load()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for x_b, out in ((0, fp), (nx - 1, lp)):
            rc = lib.mdvc_frame_expand_planes(bits.words.data_ptr(), nx, Y, Z, bits.pitch, plan.ws.data_ptr(), plan.nbytes,
                                              ctypes.byref(plan.caps), x_b, x_b + 1, out.data_ptr(), None, st)
            _lib.check(rc, "mdvc_frame_expand_planes")
        first_planes[g], last_planes[g] = fp, lp
        local[g] = {"n_true": int(h.n_true), "n_false": int(h.n_false), "contacts": plan.contacts(),
                    "bridges": plan.bridges(), "volumes": plan.label_volumes()}

    # --- boundary pairs: last plane of slab g against the first plane of slab g+1 (ring) --------
    t_local = time.perf_counter()
    firsts = gather.gather_tensors(first_planes)
    for g in own:
        local[g]["boundary"] = _plane_pairs(dev, last_planes[g], firsts[(g + 1) % G], 1 if g == G - 1 else 0)
    everything = gather.gather_objects(local)

    # --- global solve, identical on every rank ----------------------------------------------------
    t_exchange = time.perf_counter()
    merged = merge_slabs([e["n_true"] for e in everything], [e["n_false"] for e in everything],
                         [e["boundary"] for e in everything], [e["contacts"] for e in everything],
                         [e["bridges"] for e in everything], [e["volumes"] for e in everything])
    NT, NF = merged["n_true"], merged["n_false"]
    uniq = np.concatenate([np.arange(-NF, 0, dtype=np.int32), np.arange(1, NT + 1, dtype=np.int32)])
    t_merge = time.perf_counter()
    sol = hg.solve_periodic(uniq, merged["contacts"], merged["bridges"])
    comp_lut = sol.lut(NF, NT)

    res = SlabResult()
    res.count_ids, res.count_values = hg.counts_by_component(comp_lut, merged["volumes"])
    # grids beyond the reference's reach (it raises at >= 2^31 voxels and gives up orienting more than 100 000
    # components): no guard here
    res.dag, res.solution = sol.containment_dag(max_iterations=None), sol
    res.contacts, res.bridges, res.n_true, res.n_false = merged["contacts"], merged["bridges"], NT, NF
    res.extents, res.merge, res.plans = ext, merged, plans
    t_graph = time.perf_counter()

    # --- dense grids of the owned slabs: one expand pass each ----------------------------------------
    res.labels, res.components = {}, {}
    for g in own:
        bits, plan = slab_bits[g], plans[g]
        lut_g = merged["luts"][g]                                   # local label -> global label
        comp_g = comp_lut[lut_g.astype(np.int64) + NF].astype(np.int32)
        comp_g[int(headers[g].n_false)] = 0
        plan.set_lut(torch.from_numpy(comp_g).to(bits.words.device))
        lut_dev = torch.from_numpy(lut_g).to(bits.words.device)
        nx = bits.shape[0]
        rc = plan.lib.mdvc_frame_relabel_runs(nx, Y, Z, plan.ws.data_ptr(), plan.nbytes, ctypes.byref(plan.caps),
                                              lut_dev.data_ptr(), int(lut_dev.numel()),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "mdvc_frame_relabel_runs")
        labels = torch.empty(bits.shape, dtype=torch.int32, device=bits.words.device) if want_labels else None
        comps = torch.empty(bits.shape, dtype=torch.int32, device=bits.words.device)
        plan.expand(bits, labels, comps)
        res.labels[g], res.components[g] = labels, comps
    res.timings = {"morph_label_pairs_readback_s": t_local - t_start, "exchange_boundary_pairs_s": t_exchange - t_local,
                   "merge_slabs_s": t_merge - t_exchange, "host_graph_s": t_graph - t_merge,
                   "enqueue_dense_write_s": time.perf_counter() - t_graph}
    return res

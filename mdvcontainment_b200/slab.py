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
    """Global labels, contacts, bridges and volumes from per-slab results.

    n_true[g], n_false[g]     label counts of slab g (local labels 1..n_true, -1..-n_false)
    boundary_rows[g]          int32 [K,5] rows (la, lb, sx, sy, sz) of mdvc_plane_pairs between the last plane of
                              slab g (la) and the first plane of slab (g+1) % G (lb)
    local_contacts[g]         int32 [K,2]; local_bridges[g] int32 [K,5] (y/z faces only); local_volumes[g]
                              int64 [n_false+n_true+1] (index = label + n_false)

    Returns dict(luts=[int32 arrays, index = local label + n_false[g] -> global label], n_true, n_false,
    contacts int32 [K,2], bridges int32 [K,5] (shape (0,) when empty), volumes int64)."""
    G = len(n_true)
    parent = {}

    def find(k):
        root = k
        while parent.setdefault(root, root) != root:
            root = parent[root]
        while parent[k] != root:
            parent[k], k = root, parent[k]
        return root

    for g in range(G):
        rows = np.asarray(boundary_rows[g]).reshape(-1, 5)
        h = (g + 1) % G
        inbox = ~rows[:, 2:].any(axis=1)
        for la, lb in rows[inbox][:, :2].tolist():
            if (la > 0) == (lb > 0):                       # same value on both sides of an in-box step: one label
                ra, rb = find((g, la)), find((h, lb))
                if ra != rb:
                    # the smaller (slab, |label|) stays root = the group's first voxel in C order
                    if (ra[0], abs(ra[1])) > (rb[0], abs(rb[1])):
                        ra, rb = rb, ra
                    parent[rb] = ra
    nonrep = [[[], []] for _ in range(G)]                  # [g][0]: positive locals merged away, [g][1]: |negative|
    rep_of = {}
    for k in list(parent):
        r = find(k)
        if r != k:
            nonrep[k[0]][0 if k[1] > 0 else 1].append(abs(k[1]))
            rep_of[k] = r

    luts, base_pos, base_neg = [], 0, 0
    pos_tab, neg_tab = [], []                               # per slab: global number of local label k (k = 1..n)
    for g in range(G):
        tabs = []
        for sign_idx, (n, base) in enumerate(((int(n_true[g]), base_pos), (int(n_false[g]), base_neg))):
            gone = np.zeros(n + 1, dtype=np.int64)
            if nonrep[g][sign_idx]:
                gone[np.asarray(nonrep[g][sign_idx], dtype=np.int64)] = 1
            k = np.arange(n + 1, dtype=np.int64)
            tab = base + k - np.cumsum(gone)                # valid for representatives; fixed below for the others
            tabs.append((tab, gone))
        pos_tab.append(tabs[0]); neg_tab.append(tabs[1])
        base_pos += int(n_true[g]) - len(nonrep[g][0])
        base_neg += int(n_false[g]) - len(nonrep[g][1])
    for (g, l), (rg, rl) in rep_of.items():                 # merged-away labels take their representative's number
        if l > 0:
            pos_tab[g][0][l] = pos_tab[rg][0][rl]
        else:
            neg_tab[g][0][-l] = neg_tab[rg][0][-rl]
    for g in range(G):
        nt, nf = int(n_true[g]), int(n_false[g])
        lut = np.zeros(nf + nt + 1, dtype=np.int32)
        lut[nf + 1:] = pos_tab[g][0][1:]
        lut[:nf] = -neg_tab[g][0][1:][::-1]                 # index 0 is local label -nf
        luts.append(lut)
    NT, NF = base_pos, base_neg

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
        np.add.at(volumes, lut.astype(np.int64) + NF, np.asarray(local_volumes[g], dtype=np.int64))
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
    contacts = np.unique(np.concatenate(contacts), axis=0).astype(np.int32) if contacts else np.empty((0, 2), np.int32)
    bridges = np.unique(np.concatenate(bridges), axis=0).astype(np.int32) if bridges else np.array([], dtype=np.int32)
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
    __slots__ = ("containment_graph", "component_ranks", "voxel_counts", "contacts", "bridges", "n_true", "n_false",
                 "labels", "components", "extents", "merge")

    def __str__(self):
        from . import host_graph as hg
        return hg.format_dag_structure(self.containment_graph, self.component_ranks, self.voxel_counts)


def _plane_pairs(dev, plane_a, plane_b, sx_boundary, n_slots=1 << 12):
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


def analyse_slabs(slab_bits: dict, shape, gather=None, morph_ops: str = "", want_labels=True):
    """Containment analysis of a periodic grid given as x-slabs of bit grids.

    slab_bits   {slab index: BitGrid of that slab's own planes} for the slabs this process owns
    shape       (X, Y, Z) of the whole grid
    gather      LocalGather(G) (default, all slabs here) or DistGather() (one slab per rank)
    morph_ops   'd'/'e' string applied to the whole periodic grid first (halo planes exchanged)

    Returns a SlabResult; ``labels`` / ``components`` are {slab index: int32 CUDA tensor of the slab}."""
    import torch
    from . import _lib, device as dev, host_graph as hg
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
    plans, headers, first_planes, last_planes = {}, {}, {}, {}
    local = {}
    for g in own:
        bits = slab_bits[g]
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
    firsts = gather.gather_tensors(first_planes)
    for g in own:
        local[g]["boundary"] = _plane_pairs(dev, last_planes[g], firsts[(g + 1) % G], 1 if g == G - 1 else 0)
    everything = gather.gather_objects(local)

    # --- global solve, identical on every rank ----------------------------------------------------
    merged = merge_slabs([e["n_true"] for e in everything], [e["n_false"] for e in everything],
                         [e["boundary"] for e in everything], [e["contacts"] for e in everything],
                         [e["bridges"] for e in everything], [e["volumes"] for e in everything])
    NT, NF = merged["n_true"], merged["n_false"]
    uniq = np.concatenate([np.arange(-NF, 0, dtype=np.int32), np.arange(1, NT + 1, dtype=np.int32)])
    sol = hg.solve_periodic(uniq, merged["contacts"], merged["bridges"])
    c2l = sol["component2labels"]
    comp_lut = np.zeros(NF + NT + 1, dtype=np.int32)
    for comp, labs in c2l.items():
        comp_lut[np.asarray(labs, dtype=np.int64) + NF] = comp
    unique_components = np.array(sorted(int(c) for c in c2l), dtype=np.int32)

    res = SlabResult()
    lo_c = int(comp_lut.min())
    acc = np.zeros(int(comp_lut.max()) - lo_c + 1, dtype=np.int64)
    np.add.at(acc, comp_lut.astype(np.int64) - lo_c, merged["volumes"])
    res.voxel_counts = {np.int32(lo_c + k): np.int64(v) for k, v in enumerate(acc) if v}
    res.containment_graph = hg.containment_graph(sol["is_contained"], unique_components, sol["component_contact_graph"])
    res.component_ranks = sol["component_ranks"]
    res.contacts, res.bridges, res.n_true, res.n_false = merged["contacts"], merged["bridges"], NT, NF
    res.extents, res.merge = ext, merged

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
    return res

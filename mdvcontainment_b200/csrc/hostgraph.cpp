// hostgraph.cpp -- host-side (no GPU) core of the label-graph stage between the device kernels and the containment
// DAG: connected pieces of a label graph whose edges carry periodic image shifts, and the rank (0..3) of the lattice
// spanned by the shift sums of each piece's closed walks.
//
// Reference behaviour restated (file:line relative to /root/reference/mdvcontainment/):
//   pieces            graph_logic.py:101-127  (nx.connected_components of the same-sign label subgraphs)
//                     rank_logic.py:153-160   (connected pieces of a component's complement)
//   rank of a piece   rank_logic.py:5-110     (Eulerian circuit cut into cycles, SVD rank of their summed shifts)
//
// The reference's cycle sums span the same lattice as the holonomies  phi(u) + shift(u->v) - phi(v)  of the edges
// with respect to a spanning-tree potential phi (every closed walk is an integer combination of them and the circuit
// runs through every edge), so the rank is computed from those, exactly, in integers: a union-find whose nodes carry
// their offset to the parent yields phi, every edge that closes a cycle contributes one holonomy vector to the
// piece's fraction-free echelon basis (at most three vectors).  O((V + E) alpha) instead of the reference's
// O(V * E) networkx walks (22.8 s at 48 k labels).
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "../../include/mdvc.h"

namespace {

typedef long long i64;
typedef __int128 i128;

struct Vec3 { i64 c[3]; };

static i64 gcd64(i64 a, i64 b) {
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    while (b) { i64 t = a % b; a = b; b = t; }
    return a;
}

struct Basis {
    int n;              // vectors held (0..3)
    i64 v[3][3];
    int piv[3];         // pivot column of every vector
    void add(Vec3 h) {
        if (n == 3) return;
        i64 w[3] = {h.c[0], h.c[1], h.c[2]};
        for (int k = 0; k < n; ++k) {
            const int p = piv[k];
            if (w[p]) {
                const i128 f = v[k][p], g = w[p];
                i128 t[3];
                for (int j = 0; j < 3; ++j) t[j] = f * (i128)w[j] - g * (i128)v[k][j];
                // keep the entries small: divide by their gcd (the rank does not change)
                i128 d = 0;
                for (int j = 0; j < 3; ++j) {
                    i128 a = t[j] < 0 ? -t[j] : t[j], b = d;
                    while (b) { i128 r = a % b; a = b; b = r; }
                    d = a;
                }
                if (d == 0) return;
                for (int j = 0; j < 3; ++j) w[j] = (i64)(t[j] / d);
            }
        }
        int p = -1;
        for (int j = 0; j < 3; ++j) if (w[j]) { p = j; break; }
        if (p < 0) return;
        const i64 d = gcd64(gcd64(w[0], w[1]), w[2]);
        for (int j = 0; j < 3; ++j) v[n][j] = w[j] / d;
        piv[n] = p;
        ++n;
    }
    void merge(const Basis &o) {
        for (int k = 0; k < o.n && n < 3; ++k) { Vec3 h = {{o.v[k][0], o.v[k][1], o.v[k][2]}}; add(h); }
    }
};

struct ShiftForest {
    std::vector<int32_t> parent;
    std::vector<Vec3> off;               // phi(node) - phi(parent)
    std::vector<int32_t> basis_of;       // per root: index into bases, -1 = no cycle with a non-zero sum yet
    std::vector<Basis> bases;

    explicit ShiftForest(i64 n) : parent((size_t)n), off((size_t)n), basis_of((size_t)n, -1) {
        for (i64 i = 0; i < n; ++i) { parent[(size_t)i] = (int32_t)i; off[(size_t)i] = Vec3{{0, 0, 0}}; }
    }
    // root of v and phi(v) - phi(root); compresses the path
    int32_t find(int32_t v, Vec3 &pot) {
        int32_t r = v;
        Vec3 acc = {{0, 0, 0}};
        while (parent[r] != r) { for (int j = 0; j < 3; ++j) acc.c[j] += off[r].c[j]; r = parent[r]; }
        pot = acc;
        // second pass: hang every node of the path directly under the root
        int32_t cur = v;
        Vec3 rest = acc;
        while (parent[cur] != r && cur != r) {
            const int32_t next = parent[cur];
            const Vec3 o = off[cur];
            parent[cur] = r;
            off[cur] = rest;
            for (int j = 0; j < 3; ++j) rest.c[j] -= o.c[j];
            cur = next;
        }
        return r;
    }
    // edge u -> v with image shift s:  phi(v) = phi(u) + s
    void edge(int32_t u, int32_t v, const int32_t *s) {
        Vec3 pu, pv;
        int32_t ru = find(u, pu), rv = find(v, pv);
        Vec3 h;
        for (int j = 0; j < 3; ++j) h.c[j] = pu.c[j] + s[j] - pv.c[j];
        if (ru == rv) {
            if (h.c[0] | h.c[1] | h.c[2]) {
                if (basis_of[ru] < 0) { basis_of[ru] = (int32_t)bases.size(); bases.push_back(Basis{0, {}, {}}); }
                bases[(size_t)basis_of[ru]].add(h);
            }
            return;
        }
        // the smaller index stays root (deterministic representative = the piece's first node)
        if (ru < rv) {                      // rv under ru: phi(rv) - phi(ru) = h
            parent[rv] = ru; off[rv] = h;
        } else {                            // ru under rv: phi(ru) - phi(rv) = -h
            parent[ru] = rv; for (int j = 0; j < 3; ++j) off[ru].c[j] = -h.c[j];
            const int32_t t = ru; ru = rv; rv = t;
        }
        if (basis_of[rv] >= 0) {
            if (basis_of[ru] < 0) basis_of[ru] = basis_of[rv];
            else bases[(size_t)basis_of[ru]].merge(bases[(size_t)basis_of[rv]]);
            basis_of[rv] = -1;
        }
    }
};

}  // namespace

extern "C" int mdvc_host_shift_graph_ranks(long long n_nodes, const int32_t *edge_u_host, const int32_t *edge_v_host,
                                           const int32_t *edge_shift_host, long long n_edges, const uint8_t *skip_node_host,
                                           int32_t *root_out_host, int8_t *rank_out_host, int32_t *max_rank_out_host) {
    if (n_nodes < 0 || n_edges < 0 || n_nodes >= (1ll << 31)) return MDVC_ERR_BADARG;
    if (n_edges > 0 && (!edge_u_host || !edge_v_host)) return MDVC_ERR_BADARG;
    ShiftForest F(n_nodes);
    static const int32_t zero[3] = {0, 0, 0};
    for (long long e = 0; e < n_edges; ++e) {
        const int32_t u = edge_u_host[e], v = edge_v_host[e];
        if (u < 0 || v < 0 || u >= n_nodes || v >= n_nodes) return MDVC_ERR_BADARG;
        if (skip_node_host && (skip_node_host[u] || skip_node_host[v])) continue;
        F.edge(u, v, edge_shift_host ? edge_shift_host + 3 * e : zero);
    }
    int best = 0;
    for (long long i = 0; i < n_nodes; ++i) {
        if (skip_node_host && skip_node_host[i]) {
            if (root_out_host) root_out_host[i] = -1;
            if (rank_out_host) rank_out_host[i] = -1;
            continue;
        }
        Vec3 pot;
        const int32_t r = F.find((int32_t)i, pot);
        const int rk = F.basis_of[r] < 0 ? 0 : F.bases[(size_t)F.basis_of[r]].n;
        if (root_out_host) root_out_host[i] = r;
        if (rank_out_host) rank_out_host[i] = (int8_t)rk;
        if (rk > best) best = rk;
    }
    if (max_rank_out_host) *max_rank_out_host = best;
    return MDVC_OK;
}

// ------------------------------------------------------------------------------------------
// component-level bookkeeping of create_component_contact_graph (graph_logic.py:164-198) for large label graphs
// ------------------------------------------------------------------------------------------
// The reference walks the components in order, each component's labels in ascending order, each label's successors in
// the insertion order of the label graph's edges, and collects the neighbouring components.  This returns that
// sequence with every (component, neighbour) pair kept at its first occurrence -- O(nodes + edges) with two counting
// sorts and a stamp array, where NumPy needs a stable sort and a unique over all half-edges.  Nodes must be numbered by
// ascending label value (node index order == label order inside a component).
extern "C" int mdvc_host_component_neighbours(long long n_nodes, const int32_t *comp_of_node_host, int32_t n_comps,
                                              const int32_t *src_host, const int32_t *dst_host, long long n_edges,
                                              int32_t *out_cs_host, int32_t *out_cd_host, long long *n_out_host) {
    if (n_nodes < 0 || n_edges < 0 || n_comps < 0 || !n_out_host || (n_nodes > 0 && !comp_of_node_host) ||
        (n_edges > 0 && (!src_host || !dst_host || !out_cs_host || !out_cd_host)))
        return MDVC_ERR_BADARG;
    *n_out_host = 0;
    for (long long i = 0; i < n_nodes; ++i)
        if (comp_of_node_host[i] < 0 || comp_of_node_host[i] >= n_comps) return MDVC_ERR_BADARG;
    // half-edges by source node, insertion order kept (stable counting sort)
    std::vector<long long> estart((size_t)n_nodes + 1, 0);
    for (long long e = 0; e < n_edges; ++e) {
        if (src_host[e] < 0 || src_host[e] >= n_nodes || dst_host[e] < 0 || dst_host[e] >= n_nodes) return MDVC_ERR_BADARG;
        ++estart[(size_t)src_host[e] + 1];
    }
    for (long long i = 0; i < n_nodes; ++i) estart[(size_t)i + 1] += estart[(size_t)i];
    std::vector<int32_t> edst((size_t)n_edges);
    {
        std::vector<long long> fill(estart.begin(), estart.end() - 1);
        for (long long e = 0; e < n_edges; ++e) edst[(size_t)fill[(size_t)src_host[e]]++] = dst_host[e];
    }
    // nodes by component, ascending node index inside a component
    std::vector<long long> nstart((size_t)n_comps + 1, 0);
    for (long long i = 0; i < n_nodes; ++i) ++nstart[(size_t)comp_of_node_host[i] + 1];
    for (int32_t c = 0; c < n_comps; ++c) nstart[(size_t)c + 1] += nstart[(size_t)c];
    std::vector<int32_t> nodes((size_t)n_nodes);
    {
        std::vector<long long> fill(nstart.begin(), nstart.end() - 1);
        for (long long i = 0; i < n_nodes; ++i) nodes[(size_t)fill[(size_t)comp_of_node_host[i]]++] = (int32_t)i;
    }
    std::vector<int32_t> stamp((size_t)n_comps, -1);
    long long n_out = 0;
    for (int32_t c = 0; c < n_comps; ++c) {
        for (long long k = nstart[(size_t)c]; k < nstart[(size_t)c + 1]; ++k) {
            const int32_t node = nodes[(size_t)k];
            for (long long e = estart[(size_t)node]; e < estart[(size_t)node + 1]; ++e) {
                const int32_t other = comp_of_node_host[edst[(size_t)e]];
                if (other == c || stamp[(size_t)other] == c) continue;
                stamp[(size_t)other] = c;
                out_cs_host[n_out] = c;
                out_cd_host[n_out] = other;
                ++n_out;
            }
        }
    }
    *n_out_host = n_out;
    return MDVC_OK;
}

// Adjacency lists of an undirected graph in networkx's order -- a neighbour appears where the first edge joining the two
// nodes was added, from either side (nx.Graph.add_edge) -- from the edge sequence (u, v)[n].  offsets int64[n_comps + 1],
// neighbours int32[2 n] (only offsets[n_comps] entries are used).
extern "C" int mdvc_host_adjacency(int32_t n_comps, const int32_t *eu_host, const int32_t *ev_host, long long n,
                                   long long *offsets_out_host, int32_t *nbrs_out_host) {
    if (n_comps < 0 || n < 0 || !offsets_out_host || (n > 0 && (!eu_host || !ev_host || !nbrs_out_host))) return MDVC_ERR_BADARG;
    size_t cap = 64;
    while (cap < (size_t)n * 2 + 2) cap <<= 1;
    const unsigned long long EMPTY = ~0ull;
    std::vector<unsigned long long> table(cap, EMPTY);
    std::vector<uint8_t> fresh((size_t)n, 0);
    for (int32_t c = 0; c <= n_comps; ++c) offsets_out_host[c] = 0;
    for (long long k = 0; k < n; ++k) {
        const int32_t u = eu_host[k], v = ev_host[k];
        if (u < 0 || v < 0 || u >= n_comps || v >= n_comps) return MDVC_ERR_BADARG;
        const unsigned long long a = (unsigned long long)(u < v ? u : v), b = (unsigned long long)(u < v ? v : u);
        const unsigned long long key = (a << 32) | b;
        unsigned long long h = key;                              // splitmix64 finaliser: both key halves reach every bit
        h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 27; h *= 0x94d049bb133111ebull; h ^= h >> 31;
        size_t slot = (size_t)h & (cap - 1);
        while (table[slot] != EMPTY && table[slot] != key) slot = (slot + 1) & (cap - 1);
        if (table[slot] == key) continue;
        table[slot] = key;
        fresh[(size_t)k] = 1;
        ++offsets_out_host[(size_t)u + 1];
        if (u != v) ++offsets_out_host[(size_t)v + 1];
    }
    for (int32_t c = 0; c < n_comps; ++c) offsets_out_host[(size_t)c + 1] += offsets_out_host[(size_t)c];
    std::vector<long long> fill(offsets_out_host, offsets_out_host + n_comps);
    for (long long k = 0; k < n; ++k) {
        if (!fresh[(size_t)k]) continue;
        const int32_t u = eu_host[k], v = ev_host[k];
        nbrs_out_host[fill[(size_t)u]++] = v;
        if (u != v) nbrs_out_host[fill[(size_t)v]++] = u;
    }
    return MDVC_OK;
}

// The same adjacency lists for an edge sequence the caller knows to be free of repeats (no pair twice, in either
// direction): two counting passes, no hash table.  PeriodicSolution.contact_adjacency has such a sequence: its
// (component, neighbour) list is grouped by component in ascending order and symmetric, so the first occurrence of
// every undirected edge is the one with component < neighbour.
extern "C" int mdvc_host_adjacency_unique(int32_t n_comps, const int32_t *eu_host, const int32_t *ev_host, long long n,
                                          long long *offsets_out_host, int32_t *nbrs_out_host) {
    if (n_comps < 0 || n < 0 || !offsets_out_host || (n > 0 && (!eu_host || !ev_host || !nbrs_out_host))) return MDVC_ERR_BADARG;
    for (int32_t c = 0; c <= n_comps; ++c) offsets_out_host[c] = 0;
    for (long long k = 0; k < n; ++k) {
        const int32_t u = eu_host[k], v = ev_host[k];
        if (u < 0 || v < 0 || u >= n_comps || v >= n_comps) return MDVC_ERR_BADARG;
        ++offsets_out_host[(size_t)u + 1];
        if (u != v) ++offsets_out_host[(size_t)v + 1];
    }
    for (int32_t c = 0; c < n_comps; ++c) offsets_out_host[(size_t)c + 1] += offsets_out_host[(size_t)c];
    std::vector<long long> fill(offsets_out_host, offsets_out_host + n_comps);
    for (long long k = 0; k < n; ++k) {
        const int32_t u = eu_host[k], v = ev_host[k];
        nbrs_out_host[fill[(size_t)u]++] = v;
        if (u != v) nbrs_out_host[fill[(size_t)v]++] = u;
    }
    return MDVC_OK;
}

// ------------------------------------------------------------------------------------------
// slab decomposition: the global problem from the slabs' results (mdvcontainment_b200/slab.py: merge_slabs)
// ------------------------------------------------------------------------------------------
// No reference counterpart (the reference is one process on one grid); the rules it has to
// reproduce are the reference's: global labels numbered by first voxel in C order per polarity (label.py:10-17 through
// scipy), contacts as sorted unique (min, max) rows (find_label_contacts.pyx:41-77), bridges as sorted unique
// (min, max, +-shift) rows with both signs for a label bridged to itself (find_bridges.pyx:84-87).
// One pass each over the slabs' rows instead of ~40 NumPy passes over 10^7-row arrays: union-find over one node per
// (slab, local label) -- positives first, each polarity in (slab, |label|) order, so that the smallest node of a merged
// group is the group's first voxel in C order --, numbering by a running count of the roots, rows mapped through the
// per-slab tables, packed into one 64-bit key per row (key order == row order), LSD radix sort, unique.
namespace {

static int32_t uf_find(std::vector<int32_t> &p, int32_t v) {
    int32_t r = v;
    while (p[(size_t)r] != r) r = p[(size_t)r];
    while (p[(size_t)v] != r) { const int32_t n = p[(size_t)v]; p[(size_t)v] = r; v = n; }
    return r;
}

// ascending LSD radix sort, 11 bits per pass, only over the bits any key uses
static void radix_sort_u64(std::vector<unsigned long long> &a) {
    const size_t n = a.size();
    if (n < 2) return;
    unsigned long long all = 0;
    for (size_t i = 0; i < n; ++i) all |= a[i];
    std::vector<unsigned long long> b(n);
    std::vector<size_t> cnt(2048);
    for (int shift = 0; shift < 64 && (all >> shift) != 0; shift += 11) {
        std::fill(cnt.begin(), cnt.end(), (size_t)0);
        for (size_t i = 0; i < n; ++i) ++cnt[(size_t)((a[i] >> shift) & 2047u)];
        size_t run = 0;
        for (size_t d = 0; d < 2048; ++d) { const size_t c = cnt[d]; cnt[d] = run; run += c; }
        for (size_t i = 0; i < n; ++i) b[cnt[(size_t)((a[i] >> shift) & 2047u)]++] = a[i];
        a.swap(b);
    }
}

static size_t unique_sorted(std::vector<unsigned long long> &a) {
    size_t m = 0;
    for (size_t i = 0; i < a.size(); ++i)
        if (m == 0 || a[i] != a[m - 1]) a[m++] = a[i];
    a.resize(m);
    return m;
}

}  // namespace

extern "C" int mdvc_host_merge_slabs(int n_slabs, const long long *n_true_host, const long long *n_false_host,
                                     const int32_t *boundary_host, const long long *boundary_off_host,
                                     const int32_t *contacts_host, const long long *contacts_off_host,
                                     const int32_t *bridges_host, const long long *bridges_off_host,
                                     const long long *volumes_host, const long long *lut_off_host,
                                     int32_t *lut_out_host, long long *n_true_out_host, long long *n_false_out_host,
                                     int32_t *contacts_out_host, long long *n_contacts_out_host,
                                     int32_t *bridges_out_host, long long *n_bridges_out_host,
                                     long long *volumes_out_host) {
    const int G = n_slabs;
    if (G < 1 || !n_true_host || !n_false_host || !boundary_off_host || !contacts_off_host || !bridges_off_host ||
        !lut_off_host || !lut_out_host || !n_true_out_host || !n_false_out_host || !n_contacts_out_host ||
        !n_bridges_out_host || !volumes_out_host)
        return MDVC_ERR_BADARG;
    std::vector<long long> base_pos((size_t)G + 1, 0), base_neg((size_t)G + 1, 0);
    for (int g = 0; g < G; ++g) {
        if (n_true_host[g] < 0 || n_false_host[g] < 0) return MDVC_ERR_BADARG;
        if (lut_off_host[g + 1] - lut_off_host[g] != n_true_host[g] + n_false_host[g] + 1) return MDVC_ERR_BADARG;
        base_pos[(size_t)g + 1] = base_pos[(size_t)g] + n_true_host[g];
    }
    const long long total_pos = base_pos[(size_t)G];
    base_neg[0] = total_pos;
    for (int g = 0; g < G; ++g) base_neg[(size_t)g + 1] = base_neg[(size_t)g] + n_false_host[g];
    const long long n_nodes = base_neg[(size_t)G];
    if (n_nodes >= (1ll << 31)) return MDVC_ERR_BADARG;
    // local label of slab g -> node, -1 for labels outside the slab's range (bad input)
    auto node = [&](int g, int32_t lab) -> long long {
        if (lab > 0) return lab <= n_true_host[g] ? base_pos[(size_t)g] + lab - 1 : -1;
        if (lab < 0) return -(long long)lab <= n_false_host[g] ? base_neg[(size_t)g] - lab - 1 : -1;
        return -1;
    };
    auto has_shift = [](const int32_t *r) { return (r[2] | r[3] | r[4]) != 0; };

    // --- labels joined through an in-box step across a slab boundary: one global label -----------------------------
    std::vector<int32_t> parent((size_t)n_nodes);
    for (long long i = 0; i < n_nodes; ++i) parent[(size_t)i] = (int32_t)i;
    for (int g = 0; g < G; ++g) {
        const int h = (g + 1) % G;
        for (long long k = boundary_off_host[g]; k < boundary_off_host[g + 1]; ++k) {
            const int32_t *r = boundary_host + 5 * k;
            if (has_shift(r) || (r[0] > 0) != (r[1] > 0)) continue;
            const long long a = node(g, r[0]), b = node(h, r[1]);
            if (a < 0 || b < 0) return MDVC_ERR_BADARG;
            int32_t ra = uf_find(parent, (int32_t)a), rb = uf_find(parent, (int32_t)b);
            if (ra == rb) continue;
            if (ra > rb) { const int32_t t = ra; ra = rb; rb = t; }
            parent[(size_t)rb] = ra;                             // the smaller index stays the root
        }
    }
    // --- numbering: the k-th root of a polarity (in node order) is label +-k -----------------------------------------
    std::vector<int32_t> number((size_t)n_nodes);
    long long NT = 0, NF = 0;
    for (long long i = 0; i < total_pos; ++i)
        if (parent[(size_t)i] == i) number[(size_t)i] = (int32_t)++NT;
    for (long long i = total_pos; i < n_nodes; ++i)
        if (parent[(size_t)i] == i) number[(size_t)i] = (int32_t)++NF;
    for (long long i = 0; i < n_nodes; ++i) {
        const int32_t r = uf_find(parent, (int32_t)i);
        number[(size_t)i] = number[(size_t)r];                   // roots precede their members: already final
    }
    *n_true_out_host = NT;
    *n_false_out_host = NF;
    for (int g = 0; g < G; ++g) {
        int32_t *lut = lut_out_host + lut_off_host[g];
        const long long nf = n_false_host[g], nt = n_true_host[g];
        for (long long k = 1; k <= nf; ++k) lut[nf - k] = -number[(size_t)(base_neg[(size_t)g] + k - 1)];
        lut[nf] = 0;
        for (long long k = 1; k <= nt; ++k) lut[nf + k] = number[(size_t)(base_pos[(size_t)g] + k - 1)];
    }
    // --- volumes ---------------------------------------------------------------------------------------------------------
    for (long long i = 0; i < NF + NT + 1; ++i) volumes_out_host[i] = 0;
    if (volumes_host)
        for (int g = 0; g < G; ++g) {
            const int32_t *lut = lut_out_host + lut_off_host[g];
            const long long *vol = volumes_host + lut_off_host[g];
            const long long n = lut_off_host[g + 1] - lut_off_host[g];
            for (long long i = 0; i < n; ++i) volumes_out_host[(long long)lut[i] + NF] += vol[i];
        }
    volumes_out_host[NF] = 0;

    // --- contacts and bridges as packed keys ----------------------------------------------------------------------------
    const unsigned long long R = (unsigned long long)(NF + NT + 1);
    if (R >= (1ull << 29)) return MDVC_ERR_BADARG;             // (R * R * 27 must fit 64 bits)
    std::vector<unsigned long long> ckeys, bkeys;
    ckeys.reserve((size_t)(contacts_off_host[G] + boundary_off_host[G]));
    bkeys.reserve((size_t)(2 * (bridges_off_host[G] + boundary_off_host[G])));
    auto add_contact = [&](long long a, long long b) {
        if (a > b) { const long long t = a; a = b; b = t; }
        ckeys.push_back((unsigned long long)(a + NF) * R + (unsigned long long)(b + NF));
    };
    auto add_bridge = [&](long long a, long long b, int sx, int sy, int sz) {
        if (a > b) { const long long t = a; a = b; b = t; sx = -sx; sy = -sy; sz = -sz; }
        const unsigned long long pair = (unsigned long long)(a + NF) * R + (unsigned long long)(b + NF);
        bkeys.push_back(pair * 27ull + (unsigned long long)((sx + 1) * 9 + (sy + 1) * 3 + (sz + 1)));
        if (a == b) bkeys.push_back(pair * 27ull + (unsigned long long)((1 - sx) * 9 + (1 - sy) * 3 + (1 - sz)));
    };
    for (int g = 0; g < G; ++g) {
        const int h = (g + 1) % G;
        const int32_t *lut = lut_out_host + lut_off_host[g], *lut_h = lut_out_host + lut_off_host[h];
        const long long nf = n_false_host[g], nl = lut_off_host[g + 1] - lut_off_host[g];
        const long long nf_h = n_false_host[h], nl_h = lut_off_host[h + 1] - lut_off_host[h];
        for (long long k = contacts_off_host[g]; k < contacts_off_host[g + 1]; ++k) {
            const long long ia = contacts_host[2 * k] + nf, ib = contacts_host[2 * k + 1] + nf;
            if (ia < 0 || ib < 0 || ia >= nl || ib >= nl) return MDVC_ERR_BADARG;
            add_contact(lut[ia], lut[ib]);
        }
        for (long long k = bridges_off_host[g]; k < bridges_off_host[g + 1]; ++k) {
            const int32_t *r = bridges_host + 5 * k;
            const long long ia = r[0] + nf, ib = r[1] + nf;
            if (ia < 0 || ib < 0 || ia >= nl || ib >= nl || r[2] < -1 || r[2] > 1 || r[3] < -1 || r[3] > 1 || r[4] < -1 || r[4] > 1)
                return MDVC_ERR_BADARG;
            add_bridge(lut[ia], lut[ib], r[2], r[3], r[4]);
        }
        for (long long k = boundary_off_host[g]; k < boundary_off_host[g + 1]; ++k) {
            const int32_t *r = boundary_host + 5 * k;
            const long long ia = r[0] + nf, ib = r[1] + nf_h;
            if (ia < 0 || ib < 0 || ia >= nl || ib >= nl_h || r[2] < -1 || r[2] > 1 || r[3] < -1 || r[3] > 1 || r[4] < -1 || r[4] > 1)
                return MDVC_ERR_BADARG;
            const long long ga = lut[ia], gb = lut_h[ib];
            if (has_shift(r)) add_bridge(ga, gb, r[2], r[3], r[4]);
            else if ((ga > 0) != (gb > 0)) add_contact(ga, gb);
        }
    }
    radix_sort_u64(ckeys);
    radix_sort_u64(bkeys);
    const size_t nc = unique_sorted(ckeys), nb = unique_sorted(bkeys);
    if ((nc && !contacts_out_host) || (nb && !bridges_out_host)) return MDVC_ERR_BADARG;
    for (size_t i = 0; i < nc; ++i) {
        contacts_out_host[2 * i] = (int32_t)((long long)(ckeys[i] / R) - NF);
        contacts_out_host[2 * i + 1] = (int32_t)((long long)(ckeys[i] % R) - NF);
    }
    for (size_t i = 0; i < nb; ++i) {
        const unsigned long long pair = bkeys[i] / 27ull;
        const int code = (int)(bkeys[i] % 27ull);
        int32_t *o = bridges_out_host + 5 * i;
        o[0] = (int32_t)((long long)(pair / R) - NF);
        o[1] = (int32_t)((long long)(pair % R) - NF);
        o[2] = code / 9 - 1; o[3] = code / 3 % 3 - 1; o[4] = code % 3 - 1;
    }
    *n_contacts_out_host = (long long)nc;
    *n_bridges_out_host = (long long)nb;
    return MDVC_OK;
}

/* mdvc.h -- C-ABI of the B200-native voxel hot path of MDVContainment.
 *
 * The reference (BartBruininks/mdvcontainment v1.1.1) has no FFI of its own: its hot path is
 * reached through Python import-level seam functions.  Every entry point
 * below names the reference function(s) it replaces (file:line relative to the reference
 * root).  INTEGRATION.md shows the ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in
 *    `_host`; the caller owns every buffer, the library allocates nothing persistent;
 *  - every function takes a `cudaStream_t` (passed as void*), enqueues its kernels on it and
 *    returns without synchronising unless documented otherwise;
 *  - return value: MDVC_OK (0), a negative MDVC_ERR_* for bad arguments / CUDA failures, or a
 *    positive MDVC_OVERFLOW_* bit set when a caller-provided capacity was too small (retry
 *    with the sizes reported in the header block); nothing is ever silently truncated;
 *  - grids are C-order [X][Y][Z], Z fastest (the reference's NumPy layout).  The occupancy grid
 *    is bit-packed: row (x,y) is `pitch` uint32 words, bit k of word w is voxel z = 32*w + k;
 *    `pitch` = mdvc_row_pitch_words(Z) (a multiple of 4 words = 16 B); bits with z >= Z and the
 *    pad words are zero on input and on output of every function.
 */
#ifndef MDVC_H
#define MDVC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDVC_OK 0
#define MDVC_ERR_BADARG (-1)
#define MDVC_ERR_CUDA (-2)
#define MDVC_ERR_WORKSPACE (-3)
/* capacity overflow bits (also found in mdvc_frame_header.flags) */
#define MDVC_OVERFLOW_RUNS 1
#define MDVC_OVERFLOW_CONTACTS 2
#define MDVC_OVERFLOW_BRIDGES 4
#define MDVC_OVERFLOW_LABELS 8
#define MDVC_OVERFLOW_ROWS 16

int mdvc_version(void);
/* last CUDA error string seen by the calling thread's most recent failing call */
const char *mdvc_last_error(void);
int mdvc_row_pitch_words(int Z);
/* number of CUDA kernels this library has launched so far in the calling process (for benchmarks) */
unsigned long long mdvc_launch_count(void);

/* Page-locked HOST staging memory for the coordinate copies (N1: the trajectory driver's buffers; the reference
 * reads atomgroup.positions from pageable NumPy memory, mda_containment.py:385-388).  write_combined != 0: memory the
 * CPU only writes sequentially and the GPU only reads by DMA.  The caller frees with mdvc_host_free. */
int mdvc_host_alloc(size_t bytes, int write_combined, void **ptr_out_host);
int mdvc_host_free(void *ptr_host);

/* ---------------------------------------------------------------------------------------
 * T1  bool grid <-> bit grid        (reference type: dense bool ndarray,
 *     atomgroup_to_voxels.py:195-197; containment_main.py:460)
 * ------------------------------------------------------------------------------------- */
int mdvc_pack_bool(const uint8_t *grid, int X, int Y, int Z, uint32_t *bits, int pitch, void *stream);
int mdvc_unpack_bool(const uint32_t *bits, int X, int Y, int Z, int pitch, uint8_t *grid, void *stream);
/* number of True voxels (device popcount); *count_dev is a device int64 */
int mdvc_popcount(const uint32_t *bits, int X, int Y, int Z, int pitch, unsigned long long *count_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * A1  binning        replaces _voxelate_atomgroup + the scatter of create_voxels
 *     (atomgroup_to_voxels.py:131-142, 192-197).
 *  pos[n][3] float32 (A); T = inv(box) @ nbox, invT = inv(T) (row-major float64[9], computed on
 *  the host with the reference's NumPy expressions); nbox int64[9] row-major.
 *  Optional outputs (NULL to skip): occ_bits (atomicOr into the bit grid, NOT cleared here),
 *  vox_out[n][3] int32 wrapped voxel coordinates, pos_out[n][3] float32 rewritten positions
 *  (may alias pos).  T, invT, nbox are HOST pointers (copied by value into the launch).
 * ------------------------------------------------------------------------------------- */
int mdvc_bin_atoms(const float *pos, long long n, const double *T_host, const double *invT_host,
                   const long long *nbox_host, uint32_t *occ_bits, int pitch,
                   int32_t *vox_out, float *pos_out, void *stream);

/* A1 into a grid that this call also CLEARS (the usual case: one occupancy grid per frame).  Same arguments as
 * mdvc_bin_atoms plus a scratch buffer of mdvc_bin_scratch_bytes(n) bytes (16-byte aligned device memory; NULL or
 * too small = clear + mdvc_bin_atoms).  For grids of more than 64 MiB of bit words the beads' keys are first sorted
 * into one list per 64 KiB slab of the grid; every slab is then built in shared memory and written once: the
 * coordinates are read once and the grid is written once, instead of a memset plus one DRAM read-modify-write per
 * bead.  Smaller grids stay in the L2 and are cleared and filled with plain atomics.  Result identical to
 * clear + mdvc_bin_atoms (the OR of the same bits). */
size_t mdvc_bin_scratch_bytes(long long n);
int mdvc_bin_grid(const float *pos, long long n, const double *T_host, const double *invT_host,
                  const long long *nbox_host, uint32_t *occ_bits, int pitch, int32_t *vox_out, float *pos_out,
                  void *scratch, size_t scratch_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * A3/A4  morphology  replaces dilate_voxels / erode_voxels / morph_voxels / close_voxels
 *     (atomgroup_to_voxels.py:242-308): periodic 3x3x3 dilation 'd' / erosion 'e', applied left
 *     to right.  bits_in is not modified; result in bits_out; tmp is a scratch bit grid of the
 *     same size (needed when strlen(ops) > 1; may be NULL otherwise).  Unknown op character ->
 *     MDVC_ERR_BADARG (the host layer raises the reference's ValueError, :303).
 * ------------------------------------------------------------------------------------- */
int mdvc_morph(const uint32_t *bits_in, uint32_t *bits_out, uint32_t *tmp, int X, int Y, int Z,
               int pitch, const char *ops, void *stream);
/* mdvc_morph that also counts the z-runs of every row of the RESULT into row_counts[X*Y] when its last launch is a
 * fused pair of operations (atomgroup_to_voxels.py:278-304 followed by the first pass of label.py:10-13): the rows
 * are in registers there, which saves the labelling one pass over the grid.  *counted (host) = 1 if row_counts was
 * filled, 0 if the shape or morph string did not allow it (then run mdvc_frame_label as usual). */
int mdvc_morph_counted(const uint32_t *bits_in, uint32_t *bits_out, uint32_t *tmp, int X, int Y, int Z, int pitch,
                       const char *ops, uint32_t *row_counts, int *counted, void *stream);


/* ---------------------------------------------------------------------------------------
 * Frame pipeline on the bit grid (the fast path):  A5 labels (label.py:4-19), A7 contacts
 * (find_label_contacts.pyx:5-77), A6 bridges (find_bridges.pyx:4-89), the periodic-component
 * numbering of rank_logic.py:112-134 / graph_logic.py:101-162, A8 relabel +
 * counts (label.py:21-38, containment_main.py:510-512).
 *
 * Work happens on maximal z-runs of equal voxels (both polarities at once); only
 * mdvc_frame_expand touches the dense int32 grids, exactly once.
 * ------------------------------------------------------------------------------------- */
typedef struct mdvc_frame_caps {
    uint32_t max_runs;       /* capacity of the per-run tables                         */
    uint32_t max_labels;     /* capacity of the per-label tables (n_true + n_false + 1)  */
    uint32_t contact_slots;  /* hash-set slots for contact pairs  (power of two)       */
    uint32_t bridge_slots;   /* hash-set slots for bridge records (power of two)       */
} mdvc_frame_caps;

/* 64 x uint32, first words of the workspace; fetched with mdvc_frame_read_header */
typedef struct mdvc_frame_header {
    uint32_t total_runs;     /* runs found (may exceed max_runs -> MDVC_OVERFLOW_RUNS)  */
    uint32_t n_true;         /* number of positive labels                               */
    uint32_t n_false;        /* number of negative labels                               */
    uint32_t flags;          /* MDVC_OVERFLOW_* bits                                    */
    uint32_t n_contacts;     /* unique contact rows                                     */
    uint32_t n_bridges;      /* unique bridge rows                                      */
    uint32_t n_comp_pos;     /* periodic components of True voxels                      */
    uint32_t n_comp_neg;     /* periodic components of False voxels                     */
    uint32_t reserved[56];
} mdvc_frame_header;

size_t mdvc_frame_workspace_bytes(int X, int Y, int Z, const mdvc_frame_caps *caps);

/* Stage 1: runs -> union-find -> canonical label per run (+ label volumes).
 * label numbering = rank of the component's first voxel in C-order, per polarity (scipy's). */
int mdvc_frame_label(const uint32_t *bits, int X, int Y, int Z, int pitch,
                     void *workspace, size_t workspace_bytes, const mdvc_frame_caps *caps, void *stream);
/* mdvc_frame_label whose stage 1a (z-runs per row, label.py:10-13 restated on runs) was already done by the kernel
 * that produced the grid: rows_counted != 0 means mdvc_frame_row_counts_ptr() holds the count of every row. */
int mdvc_frame_label_counted(const uint32_t *bits, int X, int Y, int Z, int pitch, void *workspace, size_t workspace_bytes,
                             const mdvc_frame_caps *caps, int rows_counted, void *stream);
/* Device array (uint32[X*Y + 2]) the label stage reads its per-row run counts from. */
uint32_t *mdvc_frame_row_counts_ptr(void *workspace, int X, int Y, int Z, const mdvc_frame_caps *caps);

/* Stage 2: unique contact pairs and bridge records, sorted like the reference's rows.
 * periodic = 0 skips bridges (the reference's slab=True path, containment_main.py:67-69); 1 = periodic in x, y
 * and z; 2 = periodic in y and z only (one slab of a slab-decomposed grid: the x faces are the caller's,
 * see mdvc_plane_pairs). */
int mdvc_frame_pairs(const uint32_t *bits, int X, int Y, int Z, int pitch,
                     void *workspace, size_t workspace_bytes, const mdvc_frame_caps *caps,
                     int periodic, void *stream);
/* Stage 3: periodic components on the device: union labels joined by same-sign bridges, number
 * them with the reference's rule, build label->component LUT, per-component voxel counts.
 * periodic = 0: every label is its own component (component id = label id). */
int mdvc_frame_components(int X, int Y, int Z, void *workspace, size_t workspace_bytes,
                          const mdvc_frame_caps *caps, int periodic, void *stream);
/* Stage 3': alternatively install a host-computed LUT (lut_host[n_false + n_true + 1], index =
 * label + n_false; the shape of create_components_grid's dict, label.py:21-38). */
int mdvc_frame_set_lut(int X, int Y, int Z, void *workspace, size_t workspace_bytes,
                       const mdvc_frame_caps *caps, const int32_t *lut_dev, uint32_t n_entries, void *stream);
/* Stage 4: write the dense int32 grids, each voxel exactly once.  Either output may be NULL. */
int mdvc_frame_expand(const uint32_t *bits, int X, int Y, int Z, int pitch,
                      void *workspace, size_t workspace_bytes, const mdvc_frame_caps *caps,
                      int32_t *labels_out, int32_t *components_out, void *stream);

/* Same for the planes [x_begin, x_end) only; outputs hold (x_end - x_begin) * Y * Z elements. */
int mdvc_frame_expand_planes(const uint32_t *bits, int X, int Y, int Z, int pitch,
                             void *workspace, size_t workspace_bytes, const mdvc_frame_caps *caps,
                             int x_begin, int x_end, int32_t *labels_out, int32_t *components_out, void *stream);
/* Slab decomposition: replace every run's label by lut[label + n_false] (local -> global numbering).  Call
 * after mdvc_frame_set_lut (whose LUT is indexed by the local labels) and before the final expand. */
int mdvc_frame_relabel_runs(int X, int Y, int Z, void *workspace, size_t workspace_bytes,
                            const mdvc_frame_caps *caps, const int32_t *lut_dev, uint32_t n_entries, void *stream);

/* N2: per-atom / per-voxel-key values straight from the run tables, no dense grid needed
 * (replaces the full-grid passes of set_betafactors / get_atomgroup_from_nodes, mda_containment.py:139-220).
 * which = 0: non-periodic label, 1: component (after mdvc_frame_set_lut / mdvc_frame_components).
 * mdvc_frame_atom_values: vox int32[n][3] (wrapped voxel of every atom) -> out double[n] (the tempfactor dtype);
 * atoms outside the grid get 0.  mdvc_frame_values_at: keys uint64[n] linear voxel indices (e.g. the sorted keys of
 * mdvc_atom_sort_by_voxel) -> out int32[n]. */
int mdvc_frame_atom_values(const int32_t *vox, long long n, int X, int Y, int Z, void *workspace, size_t workspace_bytes,
                           const mdvc_frame_caps *caps, int which, double *out, void *stream);
int mdvc_frame_values_at(const unsigned long long *keys, long long n, int X, int Y, int Z, void *workspace,
                         size_t workspace_bytes, const mdvc_frame_caps *caps, int which, int32_t *out, void *stream);

/* Synchronises the stream and copies the header to the host. Returns flags as MDVC_OVERFLOW_*. */
int mdvc_frame_read_header(const void *workspace, mdvc_frame_header *header_host, void *stream);
/* Device pointers into the workspace (valid after the producing stage; sizes from the header):
 * contacts int32[n_contacts][2], bridges int32[n_bridges][5], label volumes int64[n_false+n_true+1]
 * (index = label + n_false), LUT int32[same], component counts int64[n_comp_neg+n_comp_pos+1]
 * (index = component + n_comp_neg). */
const int32_t *mdvc_frame_contacts_ptr(const void *workspace, int X, int Y, int Z, const mdvc_frame_caps *caps);
const int32_t *mdvc_frame_bridges_ptr(const void *workspace, int X, int Y, int Z, const mdvc_frame_caps *caps);
const long long *mdvc_frame_label_volumes_ptr(const void *workspace, int X, int Y, int Z, const mdvc_frame_caps *caps);
const int32_t *mdvc_frame_lut_ptr(const void *workspace, int X, int Y, int Z, const mdvc_frame_caps *caps);
const long long *mdvc_frame_component_counts_ptr(const void *workspace, int X, int Y, int Z, const mdvc_frame_caps *caps);
/* run tables, for tests: which = 0 row offsets (uint32[X*Y+1]), 1 run starts (uint32[runs+1], linear voxel index),
 * 2 union-find parents (uint32[runs]), 3 label per run (int32[runs]), 4 component per run (int32[runs]) */
const void *mdvc_frame_debug_ptr(const void *workspace, int X, int Y, int Z, const mdvc_frame_caps *caps, int which);

/* ---------------------------------------------------------------------------------------
 * Seam-level functions on dense int32 grids (one-to-one with the reference's functions, for
 * user-supplied label grids; also the independent cross-check of the run-based path).
 * ------------------------------------------------------------------------------------- */
/* find_label_contacts (find_label_contacts.pyx:5-77).  set_slots: power-of-two scratch of uint64;
 * rows_out int32[rows_cap][2]; *n_rows_dev (device uint32) receives the unique count. */
int mdvc_grid_contacts(const int32_t *labels, int X, int Y, int Z, unsigned long long *set_slots,
                       uint32_t n_slots, int32_t *rows_out, uint32_t rows_cap, uint32_t *n_rows_dev,
                       uint32_t *flags_dev, void *stream);
/* find_bridges (find_bridges.pyx:4-89); rows_out int32[rows_cap][5]. */
int mdvc_grid_bridges(const int32_t *labels, int X, int Y, int Z, unsigned long long *set_slots,
                      uint32_t n_slots, int32_t *rows_out, uint32_t rows_cap, uint32_t *n_rows_dev,
                      uint32_t *flags_dev, void *stream);
/* Slab boundaries (no reference counterpart: the reference is single-process).  Every step from plane_a (last
 * plane of a slab, int32 [Y][Z]) to the 9 neighbours in plane_b (first plane of the next slab), y/z periodic when
 * periodic_yz: rows (la, lb, sx_boundary, sy, sz), sorted unique, NOT normalised (la, lb are numbered by
 * different slabs).  sx_boundary = 0 inside the box, +1 for the step through the x-max face. */
int mdvc_plane_pairs(const int32_t *plane_a, const int32_t *plane_b, int Y, int Z, int sx_boundary, int periodic_yz,
                     unsigned long long *set_slots, uint32_t n_slots, int32_t *rows_out, uint32_t rows_cap,
                     uint32_t *n_rows_dev, uint32_t *flags_dev, void *stream);
/* create_components_grid (label.py:21-38): out = lut[label - lut_lo] (0 where outside the LUT). */
int mdvc_grid_relabel(const int32_t *labels, long long n, const int32_t *lut, int lut_lo, int lut_n,
                      int32_t *out, void *stream);
/* np.unique(grid, return_counts=True) for values in [lo, lo+n_bins) (containment_main.py:510-512):
 * counts[v - lo] += 1.  counts is NOT cleared here. */
int mdvc_grid_count(const int32_t *grid, long long n, int lo, int n_bins, unsigned long long *counts,
                    void *stream);
/* np.isin(components_grid, nodes) + np.where (containment_main.py:360-378): member[v - lo] != 0
 * selects value v.  Writes the C-ordered voxel coordinates int64[M][3] to pos_out (capacity
 * pos_cap rows), M to *n_out_dev (device uint64).  block_counts: scratch of
 * mdvc_grid_select_scratch_words(n) uint32.  mask_out (uint8[n], optional) gets the boolean mask. */
size_t mdvc_grid_select_scratch_words(long long n);
int mdvc_grid_select(const int32_t *grid, int X, int Y, int Z, const uint8_t *member, int lo, int n_member,
                     uint32_t *scratch, long long *pos_out, unsigned long long pos_cap,
                     unsigned long long *n_out_dev, uint8_t *mask_out, void *stream);

/* ---------------------------------------------------------------------------------------
 * A2/A9  voxel <-> atom mapping   replaces create_efficient_mapping_cy / voxels2atomgroup_cy
 *     (atoms_voxels_mapping.pyx:17-91) and the per-node loop of set_betafactors
 *     (mda_containment.py:191-203).  Exact multimap keyed by the voxel's linear index (the
 *     reference's hash collisions, SURVEY.md Appendix B1, are deliberately not reproduced).
 * ------------------------------------------------------------------------------------- */
/* component id of the voxel each atom sits in: out[i] = grid[vox[i]] (float64, the betafactor) */
int mdvc_atom_components(const int32_t *vox, long long n, const int32_t *grid, int X, int Y, int Z,
                         double *out, void *stream);
/* CSR build: atoms stably sorted by voxel linear index.  order_out uint32[n] = positions-in-group
 * in (voxel C-order, atom order); keys_out uint64[n] = sorted linear indices.
 * scratch: mdvc_atom_sort_scratch_bytes(n) bytes. */
size_t mdvc_atom_sort_scratch_bytes(long long n);
int mdvc_atom_sort_by_voxel(const int32_t *vox, long long n, int X, int Y, int Z, uint32_t *order_out,
                            unsigned long long *keys_out, void *scratch, void *stream);
/* gather: atoms (in CSR order) whose voxel value is selected by member[] -> atom_ix[order] int64.
 * out capacity n. *n_out_dev device uint64. scratch as for mdvc_grid_select on n items.
 * keys == NULL: `grid` holds one value per CSR position (int32[n], e.g. from mdvc_frame_values_at) instead of a
 * dense grid indexed by keys. */
int mdvc_atom_select(const uint32_t *order, const unsigned long long *keys, long long n,
                     const int32_t *grid, const uint8_t *member, int lo, int n_member,
                     const long long *atom_ix, uint32_t *scratch, long long *out,
                     unsigned long long *n_out_dev, void *stream);

/* voxels2atomgroup_cy (atoms_voxels_mapping.pyx:59-91) for an explicit voxel list query[m][3] (int32): the atoms of
 * each query voxel, query order then atom order, mapped through atom_ix (int64, optional).  order/keys are the CSR
 * arrays of mdvc_atom_sort_by_voxel.  Call with out = NULL to obtain the count in *n_out_dev, then with a buffer of
 * that capacity.  scratch: mdvc_atom_gather_scratch_words(m) uint32. */
size_t mdvc_atom_gather_scratch_words(long long m);
int mdvc_atom_gather_voxels(const int32_t *query, long long m, int X, int Y, int Z, const uint32_t *order,
                            const unsigned long long *keys, long long n, const long long *atom_ix, uint32_t *scratch,
                            long long *out, unsigned long long out_cap, unsigned long long *n_out_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * N3  host graph stage (HOST function: plain host pointers, no stream, no GPU).  The label-level core of
 *     get_subgraphs / get_rank / get_ranks (graph_logic.py:101-127, rank_logic.py:5-110, 136-181): connected pieces
 *     of a graph whose directed edges u -> v carry an image shift s (phi(v) = phi(u) + s; contacts have s = 0,
 *     a bridge row (l1, l2, s) is the edge l1 -> l2) and the rank 0..3 of the lattice spanned by the shift sums of
 *     each piece's closed walks -- what the reference obtains from an Eulerian circuit + SVD.
 *  edge_u/edge_v int32[n_edges] node indices in [0, n_nodes); edge_shift int32[n_edges][3] (NULL = all zero);
 *  skip_node uint8[n_nodes] (optional): nodes -- and every edge touching them -- left out (complement of a component);
 *  outputs (each optional): root_out[n_nodes] = smallest node index of the node's piece (-1 for skipped nodes),
 *  rank_out[n_nodes] = rank of the node's piece (-1 skipped), *max_rank_out = largest rank over all pieces.
 * ------------------------------------------------------------------------------------- */
int mdvc_host_shift_graph_ranks(long long n_nodes, const int32_t *edge_u_host, const int32_t *edge_v_host,
                                const int32_t *edge_shift_host, long long n_edges, const uint8_t *skip_node_host,
                                int32_t *root_out_host, int8_t *rank_out_host, int32_t *max_rank_out_host);

/* create_component_contact_graph (graph_logic.py:164-198) for large label graphs, HOST functions.
 * mdvc_host_component_neighbours: the sequence in which the reference meets neighbouring components -- components in
 * order, their nodes in ascending index (= label) order, each node's successors in the insertion order of the
 * half-edges (src, dst)[n_edges] -- every (component, neighbour) pair at its first occurrence only.  Outputs hold up to
 * n_edges pairs.  mdvc_host_adjacency: adjacency lists in networkx's order (a neighbour appears where the first edge
 * joining two nodes was added, from either side) from an edge sequence; offsets int64[n_comps + 1], neighbours int32[2 n]. */
int mdvc_host_component_neighbours(long long n_nodes, const int32_t *comp_of_node_host, int32_t n_comps,
                                   const int32_t *src_host, const int32_t *dst_host, long long n_edges,
                                   int32_t *out_cs_host, int32_t *out_cd_host, long long *n_out_host);
int mdvc_host_adjacency(int32_t n_comps, const int32_t *eu_host, const int32_t *ev_host, long long n,
                        long long *offsets_out_host, int32_t *nbrs_out_host);
/* the same for an edge sequence without repeats (no pair twice, in either direction): no de-duplication pass */
int mdvc_host_adjacency_unique(int32_t n_comps, const int32_t *eu_host, const int32_t *ev_host, long long n,
                               long long *offsets_out_host, int32_t *nbrs_out_host);

/* Slab decomposition (no reference counterpart: the reference is one process on one grid): the global labels,
 * contacts, bridges and volumes from the per-slab results of n_slabs x-slabs, HOST function.  Slab g has local labels
 * 1..n_true[g] and -1..-n_false[g]; boundary rows (la, lb, sx, sy, sz) of mdvc_plane_pairs relate the last plane of
 * slab g (la) to the first plane of slab (g+1) % n_slabs (lb); contacts int32[.][2], bridges int32[.][5] (y/z faces
 * only) and volumes int64 (index = label + n_false[g]) are the slabs' own results.  All per-slab arrays are
 * concatenated, *_off int64[n_slabs + 1] are row offsets (lut_off: element offsets of the volume / lut tables,
 * lut_off[g+1] - lut_off[g] == n_false[g] + n_true[g] + 1).  Labels joined through an in-box step across a boundary
 * become one global label, numbered per polarity by first voxel in C order like scipy.ndimage.label (label.py:10-17);
 * contacts / bridges come out as the reference's sorted unique rows (find_label_contacts.pyx:41-77,
 * find_bridges.pyx:84-87).  Outputs: lut_out (local label + n_false[g] -> global label, same offsets), the global
 * label counts, contacts_out (capacity: all contact + boundary rows), bridges_out (capacity: twice all bridge +
 * boundary rows), volumes_out (capacity: sum of all label counts + 1; index = global label + n_false_out). */
int mdvc_host_merge_slabs(int n_slabs, const long long *n_true_host, const long long *n_false_host,
                          const int32_t *boundary_host, const long long *boundary_off_host,
                          const int32_t *contacts_host, const long long *contacts_off_host,
                          const int32_t *bridges_host, const long long *bridges_off_host,
                          const long long *volumes_host, const long long *lut_off_host,
                          int32_t *lut_out_host, long long *n_true_out_host, long long *n_false_out_host,
                          int32_t *contacts_out_host, long long *n_contacts_out_host,
                          int32_t *bridges_out_host, long long *n_bridges_out_host,
                          long long *volumes_out_host);

#ifdef __cplusplus
}
#endif
#endif /* MDVC_H */

// frame.cu -- run-based connected components, contacts, bridges, periodic components and the
// single dense write pass ("expand") on the bit-packed occupancy grid.
//
// Reference behaviour restated (file:line relative to /root/reference/mdvcontainment/):
//   labels      label.py:4-19 (scipy.ndimage.label, 26-connectivity, both polarities, no wrap;
//               numbering = C-order of each component's first voxel)
//   contacts    find_label_contacts.pyx:5-77
//   bridges     find_bridges.pyx:4-89
//   components  graph_logic.py:101-162 + rank_logic.py:112-134 (numbering rule SURVEY.md A.6)
//   relabel     label.py:21-38, counts containment_main.py:510-512
//
// Data model.  A *run* is a maximal stretch of equal voxels along z inside one row (x,y).  Runs are
// numbered in raster order (row-major, then z), so "smallest run id" == "first voxel in C-order":
// a union-find that hooks larger roots under smaller ones yields the reference's canonical
// numbering without ever touching a per-voxel parent array.  All graph work (union-find, pair
// detection, periodic merge) is done on runs; the dense int32 grids are written exactly once by
// k_expand, which is the only HBM-heavy kernel of the path.
#include "common.cuh"
#include "scan.cuh"

#include <limits.h>
#include <string.h>

// ------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------
struct Hdr {                       // overlays mdvc_frame_header (64 x uint32)
    uint32_t total_runs, n_true, n_false, flags, n_contacts, n_bridges, n_comp_pos, n_comp_neg;
    uint32_t run_eff;              // runs actually stored (0 when the capacity overflowed)
    uint32_t n_labels;             // n_false + n_true + 1, 0 when the label capacity overflowed
    uint32_t n_strip_roots;        // entries of the strip-root list (interior nodes of the trees)
    uint32_t n_pair_rows;          // entries of the pair stage's row list (k_row_pairs -> k_task_pairs)
    unsigned long long tot_roots;  // hi32 = False roots, lo32 = True roots
    unsigned long long tot_comps;  // hi32 = negative components, lo32 = positive components
    uint32_t pad1[48];
};
static_assert(sizeof(Hdr) == 256, "header must be 64 words");
static_assert(sizeof(mdvc_frame_header) == 256, "public header must be 64 words");

struct Dims {
    int X, Y, Z, pitch, nw;
    long long rows;
    unsigned long long N;
};

struct Layout {
    size_t hdr, rowoff, rowpol, rowflag, pairrows, tile32, tile64, runstart, parent, lab, runcomp, scan64, vol, lparent,
        lut, lscan, ccount, cset, bset, ckeys, bkeys, crows, brows, total;
};

static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static bool caps_ok(const mdvc_frame_caps *c) {
    auto pow2 = [](uint32_t v) { return v >= 64 && (v & (v - 1)) == 0; };
    return c && c->max_runs >= 1 && c->max_labels >= 3 && pow2(c->contact_slots) && pow2(c->bridge_slots);
}

static Layout make_layout(long long rows, const mdvc_frame_caps *c) {
    Layout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes); return at; };
    size_t M = c->max_runs, NL = c->max_labels;
    size_t scan_cap = M > (size_t)rows + 1 ? M : (size_t)rows + 1;
    if (NL > scan_cap) scan_cap = NL;
    L.hdr = take(sizeof(Hdr));
    L.rowoff = take(((size_t)rows + 2) * 4);
    L.rowpol = take((size_t)rows + 4);                 // value of the first voxel of every row (1 byte each)
    L.rowflag = take((size_t)rows + 4);                // bit q: the row does not mirror its q-th raster-order predecessor row
    L.pairrows = take(((size_t)rows + 1) * 4);         // rows whose run lists the pair stage has to compare
    L.tile32 = take(mdvc_scan_scratch_elems((uint32_t)scan_cap) * 4);
    L.tile64 = take(mdvc_scan_scratch_elems((uint32_t)scan_cap) * 8);
    L.runstart = take((M + 1) * 4);
    L.parent = take(M * 4);
    L.lab = take(M * 4);
    L.runcomp = take(M * 4);
    L.scan64 = take(M * 8);
    L.vol = take(NL * 8);
    L.lparent = take(NL * 4);
    L.lut = take(NL * 4);
    L.lscan = take(NL * 8);
    L.ccount = take(NL * 8);
    L.cset = take((size_t)c->contact_slots * 8);
    L.bset = take((size_t)c->bridge_slots * 8);
    L.ckeys = take((size_t)c->contact_slots * 8);
    L.bkeys = take((size_t)c->bridge_slots * 8);
    L.crows = take((size_t)c->contact_slots * 8);
    L.brows = take((size_t)c->bridge_slots * 20);
    L.total = o;
    return L;
}

static bool make_dims(int X, int Y, int Z, int pitch, Dims &D) {
    if (X <= 0 || Y <= 0 || Z <= 0 || pitch < mdvc_words(Z)) return false;
    unsigned long long N = (unsigned long long)X * Y * Z;
    if (N >= (1ull << 31)) return false;
    if (Z >= (1 << 24)) return false;      // the link / pair kernels pack a run's z-start into 24 bits (ZMASK)
    D.X = X; D.Y = Y; D.Z = Z; D.pitch = pitch; D.nw = mdvc_words(Z);
    D.rows = (long long)X * Y; D.N = N;
    return true;
}

template <typename T>
static inline T *at(void *ws, size_t off) { return reinterpret_cast<T *>(reinterpret_cast<char *>(ws) + off); }
template <typename T>
static inline const T *at(const void *ws, size_t off) { return reinterpret_cast<const T *>(reinterpret_cast<const char *>(ws) + off); }

extern "C" size_t mdvc_frame_workspace_bytes(int X, int Y, int Z, const mdvc_frame_caps *caps) {
    if (!caps_ok(caps) || X <= 0 || Y <= 0 || Z <= 0) return 0;
    return make_layout((long long)X * Y, caps).total;
}

// ------------------------------------------------------------------------------------------
// row-group helpers: G lanes (power of two) cooperate on one row, lanes walk the row's words
// ------------------------------------------------------------------------------------------
template <int G>
struct RowWalker {
    uint32_t carry_msb;   // bit Z-1.. : MSB of the last word of the previous chunk
    uint32_t run_carry;   // starts seen in previous chunks
    __device__ __forceinline__ void begin() { carry_msb = 0; run_carry = 0; }

    // B: this lane's (masked) word w.  Returns the start mask S (bit k set <=> a run starts at z=32w+k)
    __device__ __forceinline__ uint32_t starts(uint32_t B, int w, int Z, int gl) {
        uint32_t prev = __shfl_up_sync(MDVC_FULL, B, 1, G) >> 31;
        if (gl == 0) prev = carry_msb;
        carry_msb = __shfl_sync(MDVC_FULL, B, G - 1, G) >> 31;
        uint32_t S = B ^ ((B << 1) | prev);
        if (w == 0) S |= 1u;
        return S & mdvc_valid_mask(w, Z);
    }
    // exclusive count of starts before this lane's word (whole row so far); advances the carry.
    // Smooth grids have at most one start per word: then two ballots replace the shuffle scan.
    __device__ __forceinline__ uint32_t exclusive(uint32_t S, int gl) {
        uint32_t cnt = __popc(S);
        const int gshift = ((threadIdx.x & 31) / G) * G;
        const uint32_t gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
        if (__ballot_sync(MDVC_FULL, cnt > 1) == 0) {
            uint32_t has = (__ballot_sync(MDVC_FULL, cnt != 0) >> gshift) & gmask;
            uint32_t ex = run_carry + __popc(has & ((1u << gl) - 1u));
            run_carry += __popc(has);
            return ex;
        }
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            uint32_t t = __shfl_up_sync(MDVC_FULL, inc, o, G);
            if (gl >= o) inc += t;
        }
        uint32_t ex = run_carry + inc - cnt;
        run_carry += __shfl_sync(MDVC_FULL, inc, G - 1, G);
        return ex;
    }
    // number of starts in the whole row (single-chunk rows)
    __device__ __forceinline__ uint32_t row_total(uint32_t S) {
        uint32_t cnt = __popc(S);
        if (G == 32) return __reduce_add_sync(MDVC_FULL, cnt);
        uint32_t total = cnt;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) total += __shfl_xor_sync(MDVC_FULL, total, o, G);
        return total;
    }
    // start mask of a row that fits one chunk (no carry between chunks)
    __device__ __forceinline__ uint32_t starts_single(uint32_t B, int w, int Z, int gl) {
        uint32_t prev = __shfl_up_sync(MDVC_FULL, B, 1, G) >> 31;
        if (gl == 0) prev = 0;
        uint32_t S = B ^ ((B << 1) | prev);
        if (w == 0) S |= 1u;
        return S & mdvc_valid_mask(w, Z);
    }
};

static int pick_group(int nw) {
    int g = 4;
    while (g < 32 && g < nw) g <<= 1;
    return g;
}

#define MDVC_DISPATCH_G(g, CALL)            \
    switch (g) {                            \
        case 4: { constexpr int G = 4; CALL; } break;   \
        case 8: { constexpr int G = 8; CALL; } break;   \
        case 16: { constexpr int G = 16; CALL; } break; \
        default: { constexpr int G = 32; CALL; } break; \
    }

// ------------------------------------------------------------------------------------------
// stage 1a: runs per row
// ------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256)
k_count_runs(const uint32_t *__restrict__ bits, Dims D, uint32_t *__restrict__ runcount) {
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), gi = lane / G;
    constexpr int GPW = 32 / G;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    int nchunks = (D.nw + G - 1) / G;
    if (nchunks == 1) {
        // rows of at most G words: four rows per group and iteration keep four loads in flight
        constexpr int U = 4;
        const uint32_t vm = mdvc_valid_mask(gl, D.Z);
        for (long long rb = warp * (GPW * U); rb < D.rows; rb += nwarps * (GPW * U)) {
            uint32_t B[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                long long row = rb + u * GPW + gi;
                B[u] = (row < D.rows && gl < D.nw) ? (bits[row * D.pitch + gl] & vm) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                long long row = rb + u * GPW + gi;
                RowWalker<G> rw; rw.begin();
                uint32_t total = rw.row_total(rw.starts_single(B[u], gl, D.Z, gl));
                if (row < D.rows && gl == 0) runcount[row] = total;
            }
        }
        return;
    }
    for (long long rb = warp * GPW; rb < D.rows; rb += nwarps * GPW) {
        long long row = rb + gi;
        bool rv = row < D.rows;
        const uint32_t *rp = bits + (rv ? row : 0) * D.pitch;
        RowWalker<G> rw; rw.begin();
        uint32_t total = 0;
        for (int c = 0; c < nchunks; ++c) {
            int w = c * G + gl;
            uint32_t B = (rv && w < D.nw) ? (rp[w] & mdvc_valid_mask(w, D.Z)) : 0u;
            uint32_t S = rw.starts(B, w, D.Z, gl);
            total += __popc(S);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) total += __shfl_down_sync(MDVC_FULL, total, o, G);
        if (rv && gl == 0) runcount[row] = total;
    }
}

// after the row scan: publish the totals, raise the overflow flag, sentinel runstart[R] = N
__global__ void k_after_row_scan(Hdr *h, const uint32_t *rowoff, long long rows, uint32_t max_runs,
                                 uint32_t *runstart, unsigned long long N) {
    uint32_t R = rowoff[rows];
    h->total_runs = R;
    if (R > max_runs) { h->flags |= MDVC_OVERFLOW_RUNS; h->run_eff = 0; }
    else { h->run_eff = R; runstart[R] = (uint32_t)N; }
}

// stage 1b: run start positions (linear voxel index) + union-find init
template <int G>
__global__ void __launch_bounds__(256)
k_fill_runs(const uint32_t *__restrict__ bits, Dims D, const Hdr *__restrict__ h, const uint32_t *__restrict__ rowoff,
            uint32_t *__restrict__ runstart, uint32_t *__restrict__ parent, uint8_t *__restrict__ rowpol) {
    if (h->run_eff == 0) return;
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), gi = lane / G;
    constexpr int GPW = 32 / G;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    int nchunks = (D.nw + G - 1) / G;
    if (nchunks == 1) {
        constexpr int U = 4;
        const uint32_t vm = mdvc_valid_mask(gl, D.Z);
        for (long long rb = warp * (GPW * U); rb < D.rows; rb += nwarps * (GPW * U)) {
            uint32_t B[U], base[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                long long row = rb + u * GPW + gi;
                bool rv = row < D.rows;
                B[u] = (rv && gl < D.nw) ? (bits[row * D.pitch + gl] & vm) : 0u;
                base[u] = rv ? rowoff[row] : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                long long row = rb + u * GPW + gi;
                RowWalker<G> rw; rw.begin();
                uint32_t S = rw.starts_single(B[u], gl, D.Z, gl);
                if (row < D.rows && gl == 0) rowpol[row] = (uint8_t)(B[u] & 1u);
                uint32_t id = base[u] + rw.exclusive(S, gl);
                uint32_t lin0 = (uint32_t)(row * D.Z) + (uint32_t)(gl << 5);
                if (row >= D.rows) S = 0;
                while (S) {
                    int k = __ffs(S) - 1;
                    S &= S - 1;
                    runstart[id] = lin0 + k;
                    parent[id] = id;
                    ++id;
                }
            }
        }
        return;
    }
    for (long long rb = warp * GPW; rb < D.rows; rb += nwarps * GPW) {
        long long row = rb + gi;
        bool rv = row < D.rows;
        const uint32_t *rp = bits + (rv ? row : 0) * D.pitch;
        uint32_t base = rv ? rowoff[row] : 0u;
        RowWalker<G> rw; rw.begin();
        for (int c = 0; c < nchunks; ++c) {
            int w = c * G + gl;
            uint32_t B = (rv && w < D.nw) ? (rp[w] & mdvc_valid_mask(w, D.Z)) : 0u;
            uint32_t S = rw.starts(B, w, D.Z, gl);
            if (rv && w == 0) rowpol[row] = (uint8_t)(B & 1u);
            uint32_t ex = rw.exclusive(S, gl);
            uint32_t id = base + ex;
            uint32_t lin0 = (uint32_t)(row * D.Z) + (uint32_t)(w << 5);
            if (!rv) S = 0;
            while (S) {
                int k = __ffs(S) - 1;
                S &= S - 1;
                runstart[id] = lin0 + k;
                parent[id] = id;
                ++id;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 1a/1b for rows of at most 128 words (Z <= 4096): LQ lanes per row, four words (one 16-byte load) per
// lane.  Run starts inside the quad come from funnel shifts, only the quad's first word needs a neighbour lane;
// about a third of the instructions per word of the word-per-lane kernels above.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 quad_valid_mask(int ql, int nq, int Z) {
    if (ql >= nq) return make_uint4(0u, 0u, 0u, 0u);
    return make_uint4(mdvc_valid_mask(4 * ql, Z), mdvc_valid_mask(4 * ql + 1, Z), mdvc_valid_mask(4 * ql + 2, Z),
                      mdvc_valid_mask(4 * ql + 3, Z));
}

template <int LQ>
__global__ void __launch_bounds__(256)
k_count_runs_q(const uint32_t *__restrict__ bits, Dims D, uint32_t *__restrict__ runcount) {
    constexpr int RPW = 32 / LQ, U = 4;
    const int lane = threadIdx.x & 31, ql = lane & (LQ - 1), gi = lane / LQ;
    const int nq = D.pitch >> 2;
    const uint4 vm = quad_valid_mask(ql, nq, D.Z);
    const uint4 *b4 = reinterpret_cast<const uint4 *>(bits);
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long rb = warp * (RPW * U); rb < D.rows; rb += nwarps * (RPW * U)) {
        uint4 B[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            long long row = rb + u * RPW + gi;
            B[u] = (row < D.rows && ql < nq) ? __ldg(b4 + row * nq + ql) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            long long row = rb + u * RPW + gi;
            QuadStarts q = mdvc_quad_starts(B[u], vm, ql, LQ);
            uint32_t total = __popc(q.s0) + __popc(q.s1) + __popc(q.s2) + __popc(q.s3);
#pragma unroll
            for (int o = LQ / 2; o > 0; o >>= 1) total += __shfl_xor_sync(MDVC_FULL, total, o, LQ);
            if (row < D.rows && ql == 0) runcount[row] = total;
        }
    }
}

__device__ __forceinline__ uint32_t emit_starts(uint32_t S, uint32_t id, uint32_t lin0, uint32_t *__restrict__ runstart,
                                                uint32_t *__restrict__ parent) {
    while (S) {
        int k = __ffs(S) - 1;
        S &= S - 1;
        runstart[id] = lin0 + k;
        parent[id] = id;
        ++id;
    }
    return id;
}

template <int LQ>
__global__ void __launch_bounds__(256)
k_fill_runs_q(const uint32_t *__restrict__ bits, Dims D, const Hdr *__restrict__ h, const uint32_t *__restrict__ rowoff,
              uint32_t *__restrict__ runstart, uint32_t *__restrict__ parent, uint8_t *__restrict__ rowpol) {
    if (h->run_eff == 0) return;
    constexpr int RPW = 32 / LQ, U = 4;
    const int lane = threadIdx.x & 31, ql = lane & (LQ - 1), gi = lane / LQ;
    const int nq = D.pitch >> 2;
    const uint4 vm = quad_valid_mask(ql, nq, D.Z);
    const uint4 *b4 = reinterpret_cast<const uint4 *>(bits);
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long rb = warp * (RPW * U); rb < D.rows; rb += nwarps * (RPW * U)) {
        uint4 B[U];
        uint32_t base[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            long long row = rb + u * RPW + gi;
            bool rv = row < D.rows;
            B[u] = (rv && ql < nq) ? __ldg(b4 + row * nq + ql) : make_uint4(0u, 0u, 0u, 0u);
            base[u] = rv ? rowoff[row] : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            long long row = rb + u * RPW + gi;
            QuadStarts q = mdvc_quad_starts(B[u], vm, ql, LQ);
            const uint32_t cnt = __popc(q.s0) + __popc(q.s1) + __popc(q.s2) + __popc(q.s3);
            uint32_t inc = cnt;
#pragma unroll
            for (int o = 1; o < LQ; o <<= 1) {
                uint32_t t = __shfl_up_sync(MDVC_FULL, inc, o, LQ);
                if (ql >= o) inc += t;
            }
            if (row >= D.rows) continue;
            if (ql == 0) rowpol[row] = (uint8_t)(B[u].x & 1u);
            uint32_t id = base[u] + inc - cnt;
            const uint32_t lin0 = (uint32_t)(row * D.Z) + (uint32_t)(ql << 7);
            id = emit_starts(q.s0, id, lin0, runstart, parent);
            id = emit_starts(q.s1, id, lin0 + 32, runstart, parent);
            id = emit_starts(q.s2, id, lin0 + 64, runstart, parent);
            emit_starts(q.s3, id, lin0 + 96, runstart, parent);
        }
    }
}

#define MDVC_DISPATCH_LQ(lq, CALL)                        \
    switch (lq) {                                         \
        case 1: { constexpr int LQ = 1; CALL; } break;    \
        case 2: { constexpr int LQ = 2; CALL; } break;    \
        case 4: { constexpr int LQ = 4; CALL; } break;    \
        case 8: { constexpr int LQ = 8; CALL; } break;    \
        case 16: { constexpr int LQ = 16; CALL; } break;  \
        default: { constexpr int LQ = 32; CALL; } break;  \
    }

// ------------------------------------------------------------------------------------------
// run lookups
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bit_at(const uint32_t *__restrict__ bits, const Dims &D, uint32_t row, int z) {
    return (bits[(size_t)row * D.pitch + (z >> 5)] >> (z & 31)) & 1u;
}

// id of the run of `row` that contains voxel z (largest start <= row*Z + z)
__device__ __forceinline__ uint32_t run_at(const uint32_t *__restrict__ runstart, const uint32_t *__restrict__ rowoff,
                                           const Dims &D, uint32_t row, int z) {
    uint32_t lo = rowoff[row], hi = rowoff[row + 1];
    uint32_t target = row * (uint32_t)D.Z + (uint32_t)z;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (runstart[mid] <= target) lo = mid; else hi = mid;
    }
    return lo;
}

// per-CTA cache of 64-bit keys in shared memory (dedup in front of global structures)
template <int SLOTS>
struct BlockSetT {
    unsigned long long slot[SLOTS];
    __device__ void clear() {
        for (int i = threadIdx.x; i < SLOTS; i += blockDim.x) slot[i] = MDVC_KEY_EMPTY;
    }
    // 0: the key was already known to this block, 1: stored now, 2: no room for it (16 probes)
    __device__ int add_state(unsigned long long key) {
        uint32_t h = (((uint32_t)key ^ (uint32_t)(key >> 31)) * 2654435761u >> 16) & (SLOTS - 1);
        for (int probe = 0; probe < 16; ++probe) {
            unsigned long long cur = slot[h];
            if (cur == key) return 0;
            if (cur == MDVC_KEY_EMPTY) {
                unsigned long long old = atomicCAS(&slot[h], MDVC_KEY_EMPTY, key);
                if (old == MDVC_KEY_EMPTY) return 1;
                if (old == key) return 0;
            }
            h = (h + 1) & (SLOTS - 1);
        }
        return 2;
    }
    // true if the key was not yet known to this block (or the block cache is full)
    __device__ bool add(unsigned long long key) { return add_state(key) != 0; }
    // Every key stored here goes into a global set, one key per thread (all threads of the CTA, after a barrier).
    // Pair kernels park their keys in the cache and flush once: an insertion is an L2 round trip, and the thread
    // that walks a wrapped row would otherwise pay one per new key, back to back.
    __device__ void flush(unsigned long long *set, uint32_t n_slots, uint32_t *flags, uint32_t overflow_bit) {
        for (int i = threadIdx.x; i < SLOTS; i += blockDim.x) {
            const unsigned long long k = slot[i];
            if (k != MDVC_KEY_EMPTY && !mdvc_set_insert(set, n_slots, k)) atomicOr(flags, overflow_bit);
        }
    }
};
typedef BlockSetT<256> BlockSet;

// ------------------------------------------------------------------------------------------
// stage 1c: link runs of equal value that are 26-adjacent (no wrap)
// Two runs [a,b] (row r) and [c,d] (neighbour row r') touch iff c <= b+1 and d >= a-1.  Every
// unordered pair of neighbouring rows is visited once, from the later row towards its four
// raster-order predecessors (x-1,y-1) (x-1,y) (x-1,y+1) (x,y-1); the thread of run A walks the runs
// of the other row that intersect [a-1, b+1] (every second one has A's value).  The row pairs are
// processed in hierarchical rounds (k_link_round) rather than all at once.
// ------------------------------------------------------------------------------------------
// Interleaved climb (Rem's scheme): always step the cursor with the larger id; the two paths meet at
// their lowest common ancestor, so runs that are already connected never travel to the (hot) root of
// a huge component.  A root is hooked under the smaller cursor with atomicCAS.  Afterwards both start
// nodes are pointed at the meeting node (valid: it is an ancestor of both, and only non-roots are
// written with plain stores).
__device__ __forceinline__ void rem_union(uint32_t *parent, uint32_t a, uint32_t b) {
    uint32_t x = a, y = b;
    while (x != y) {
        if (x > y) {
            uint32_t px = __ldcg(&parent[x]);
            if (px == x) {
                uint32_t old = atomicCAS(&parent[x], x, y);
                if (old == x) { x = y; break; }
                px = old;
            }
            x = px;
        } else {
            uint32_t py = __ldcg(&parent[y]);
            if (py == y) {
                uint32_t old = atomicCAS(&parent[y], y, x);
                if (old == y) { y = x; break; }
                py = old;
            }
            y = py;
        }
    }
    if (a != x && __ldcg(&parent[a]) != x) parent[a] = x;
    if (b != x && __ldcg(&parent[b]) != x) parent[b] = x;
}

// One linking round.  phase 0 joins rows (x,y-1) and (x,y), phase 1 joins plane x-1 and plane x
// (offsets (-1,-1) (-1,0) (-1,+1)).  A round with stride s handles the borders at coordinates c with
// c % s == 0 and c % (s * LINK_RADIX) != 0, i.e. it merges aligned blocks of s rows / planes in
// groups of LINK_RADIX, which bounds the tree depth a round can add and keeps the contended roots of
// a round in disjoint groups.
// Most adjacent run pairs are redundant (same two trees as their neighbours): each thread climbs
// read-only to the current roots, and a per-CTA key cache lets only the first (root,root) pair of a
// CTA go on to the atomic union.
#ifndef LINK_RADIX
#define LINK_RADIX 32
#endif
__device__ __forceinline__ uint32_t climb_ro(const uint32_t *parent, uint32_t v) {
    uint32_t up;
    while ((up = __ldcg(&parent[v])) != v) v = up;
    return v;
}

__global__ void __launch_bounds__(256)
k_link_round(const uint32_t *__restrict__ bits, Dims D, const Hdr *__restrict__ h, const uint32_t *__restrict__ rowoff,
             const uint32_t *__restrict__ runstart, uint32_t *parent, int phase, int s, int radix, int force_lpr,
             uint8_t *__restrict__ rowflag) {
    __shared__ BlockSet seen;
    seen.clear();
    __syncthreads();
    if (h->run_eff == 0) return;
    // few runs per row: one thread per border row; many: 8 lanes share a row
    const int lpr = force_lpr ? force_lpr : (h->total_runs > 4ull * (unsigned long long)D.rows ? 8 : 1);
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = (int)(gtid % lpr);
    const long long t = gtid / lpr, nt = ((long long)gridDim.x * blockDim.x) / lpr;
    const int extent = phase == 0 ? D.Y : D.X;
    const int cnt = (extent - 1) / s;                       // multiples of s in [s, extent)
    const long long n_items = phase == 0 ? (long long)D.X * cnt : (long long)cnt * D.Y;
    for (long long item = t; item < n_items; item += nt) {
        int x, y, m;
        if (phase == 0) { x = (int)(item / cnt); m = (int)(item % cnt) + 1; y = m * s; }
        else { m = (int)(item / D.Y) + 1; x = m * s; y = (int)(item % D.Y); }
        if (m % radix == 0) continue;                       // border of a later round
        const uint32_t r = (uint32_t)x * D.Y + y;
        const uint32_t i1 = rowoff[r + 1];
        if (phase == 0 && rowflag && sub == 0) {
            // a strip's first row against the row before it (the previous strip's last): same test as rows_mirror,
            // on the global run table
            const uint32_t a0 = rowoff[r], b0 = rowoff[r - 1], n = i1 - a0;
            const uint32_t za = r * (uint32_t)D.Z, zb = (r - 1) * (uint32_t)D.Z;
            bool mir = (a0 - b0) == n && bit_at(bits, D, r, 0) == bit_at(bits, D, r - 1, 0);
            for (uint32_t k = 0; mir && k + 1 < n; ++k)
                mir = runstart[b0 + k] - zb < runstart[a0 + k + 1] - za && runstart[a0 + k] - za < runstart[b0 + k + 1] - zb;
            rowflag[r] = (uint8_t)!mir;
        }
        for (uint32_t i = rowoff[r] + sub; i < i1; i += lpr) {
            uint32_t lin = runstart[i];
            int a = (int)(lin - r * (uint32_t)D.Z);
            int b = (int)(runstart[i + 1] - r * (uint32_t)D.Z) - 1;
            uint32_t v = bit_at(bits, D, r, a);
            int lo = a > 0 ? a - 1 : 0, hi = b + 1 < D.Z ? b + 1 : D.Z - 1;
            int dy1 = phase == 0 ? -1 : 1;
            int nx = phase == 0 ? x : x - 1;
            uint32_t ra = climb_ro(parent, i);
            for (int dy = -1; dy <= dy1; ++dy) {
                int ny = y + dy;
                if (ny < 0 || ny >= D.Y) continue;
                uint32_t r2 = (uint32_t)nx * D.Y + ny;
                uint32_t j = run_at(runstart, rowoff, D, r2, lo);
                if (bit_at(bits, D, r2, lo) != v) ++j;         // runs alternate: the next one has value v
                uint32_t end = rowoff[r2 + 1], limit = r2 * (uint32_t)D.Z + (uint32_t)hi;
                for (; j < end && runstart[j] <= limit; j += 2) {
                    uint32_t rb = climb_ro(parent, j);
                    if (ra == rb) continue;
                    unsigned long long key = ra < rb ? ((unsigned long long)ra << 32) | rb : ((unsigned long long)rb << 32) | ra;
                    if (!seen.add(key)) continue;              // another thread of this CTA joins these two trees
                    rem_union(parent, ra, rb);
                    ra = ra < rb ? ra : rb;                    // still a node of the merged tree; good enough as a start
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 1c', tile kernels: run lists and union-find of a strip in shared memory
// ------------------------------------------------------------------------------------------
// A strip = STRIP_TY consecutive rows of one plane.  Its runs are contiguous in run-id space, so the
// whole strip is loaded with coalesced reads, linked with a shared-memory union-find (no global
// round trips) and written back flattened.  Strips whose run count exceeds the shared capacity
// (noise-like input) fall back to the global-memory walk.  Strip roots are appended to a list: they
// are the only runs that can become interior tree nodes later, so flattening that short list
// (k_flatten_list) after each merging round keeps every run within two hops of its root.
#ifndef STRIP_TY
#define STRIP_TY 64
#endif
#ifndef STRIP_CAP
#define STRIP_CAP 2048
#endif
#ifndef TILE_THREADS
#define TILE_THREADS 128
#endif
#define TILE_MIN_CTAS (1024 / TILE_THREADS)
#define ZMASK 0xffffffu            // packed run entry: (local row << 24) | z   (Z < 2^24, STRIP_TY + 2 < 256)

__device__ __forceinline__ void smem_union(volatile uint32_t *par, uint32_t a, uint32_t b) {
    uint32_t x = a, y = b;
    while (x != y) {
        if (x > y) {
            uint32_t px = par[x];
            if (px == x) {
                uint32_t old = atomicCAS((uint32_t *)&par[x], x, y);
                if (old == x) break;
                px = old;
            }
            x = px;
        } else {
            uint32_t py = par[y];
            if (py == y) {
                uint32_t old = atomicCAS((uint32_t *)&par[y], y, x);
                if (old == y) break;
                py = old;
            }
            y = py;
        }
    }
}

// last index in [n0, n1) whose z-start is <= z (n0 always qualifies: every row starts at z = 0)
__device__ __forceinline__ uint32_t tile_run_at(const uint32_t *tz, uint32_t n0, uint32_t n1, uint32_t z) {
    uint32_t lo = n0, hi = n1;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if ((tz[mid] & ZMASK) <= z) lo = mid; else hi = mid;
    }
    return lo;
}

// global-memory walk of one run against one neighbour row (fallback path and strip borders)
__device__ __forceinline__ void link_run_vs_row(const uint32_t *__restrict__ bits, const Dims &D,
                                                const uint32_t *__restrict__ rowoff, const uint32_t *__restrict__ runstart,
                                                uint32_t *parent, BlockSet &seen, uint32_t i, uint32_t &ra, uint32_t v,
                                                int lo, int hi, uint32_t r2) {
    uint32_t j = run_at(runstart, rowoff, D, r2, lo);
    if (bit_at(bits, D, r2, lo) != v) ++j;
    uint32_t end = rowoff[r2 + 1], limit = r2 * (uint32_t)D.Z + (uint32_t)hi;
    for (; j < end && runstart[j] <= limit; j += 2) {
        uint32_t rb = climb_ro(parent, j);
        if (ra == rb) continue;
        unsigned long long key = ra < rb ? ((unsigned long long)ra << 32) | rb : ((unsigned long long)rb << 32) | ra;
        if (!seen.add(key)) continue;
        rem_union(parent, ra, rb);
        ra = ra < rb ? ra : rb;
    }
}

// runs [a0, a1) of one row against runs [b0, b1) of a neighbour row (packed entries, z in the low 24 bits):
// same count, same first value, and every run starts before the next run of the other row
__device__ __forceinline__ bool rows_mirror(const uint32_t *tz, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, bool same_first) {
    const uint32_t n = a1 - a0;
    bool mirror = same_first && (b1 - b0) == n;
    for (uint32_t m = 0; mirror && m + 1 < n; ++m)
        mirror = (tz[b0 + m] & ZMASK) < (tz[a0 + m + 1] & ZMASK) && (tz[a0 + m] & ZMASK) < (tz[b0 + m + 1] & ZMASK);
    return mirror;
}

__global__ void __launch_bounds__(TILE_THREADS, TILE_MIN_CTAS)
k_strip_link(const uint32_t *__restrict__ bits, Dims D, Hdr *h, const uint32_t *__restrict__ rowoff,
             const uint8_t *__restrict__ rowpol, const uint32_t *__restrict__ runstart, uint32_t *parent, uint32_t *rootlist,
             uint8_t *__restrict__ rowflag) {
    __shared__ uint32_t srow[STRIP_TY + 1];
    __shared__ uint32_t spol[STRIP_TY];
    __shared__ uint8_t smir[STRIP_TY], stop[STRIP_TY];
    __shared__ uint32_t sz[STRIP_CAP];
    __shared__ uint32_t spar[STRIP_CAP];
    __shared__ uint32_t s_nroots, s_base;
    __shared__ BlockSet seen;
    if (h->run_eff == 0) return;
    const int spp = (D.Y + STRIP_TY - 1) / STRIP_TY;
    const long long n_strips = (long long)D.X * spp;
    for (long long sid = blockIdx.x; sid < n_strips; sid += gridDim.x) {
        const int x = (int)(sid / spp), y0 = (int)(sid % spp) * STRIP_TY;
        const int ny = min(STRIP_TY, D.Y - y0);
        const uint32_t r0 = (uint32_t)x * D.Y + y0;
        __syncthreads();                                   // previous strip fully consumed
        for (int t = threadIdx.x; t <= ny; t += blockDim.x) {
            srow[t] = rowoff[r0 + t];
            if (t < ny) spol[t] = rowpol[r0 + t];
        }
        __syncthreads();
        const uint32_t gbase = srow[0], nloc = srow[ny] - gbase;
        if (nloc <= STRIP_CAP) {
            for (uint32_t k = threadIdx.x; k < nloc; k += blockDim.x) {
                uint32_t off = runstart[gbase + k] - r0 * (uint32_t)D.Z;
                uint32_t lr = off / (uint32_t)D.Z;
                sz[k] = (lr << 24) | (off - lr * (uint32_t)D.Z);
                spar[k] = k;
            }
            __syncthreads();
            // rows that mirror the previous row (same run count, same first value, interleaving starts): run m
            // touches run m of that row and no other run of its value, so the link needs no search
            // rowflag bit 0 = "does not mirror row y-1": the pair kernel only walks such rows (this store initialises
            // the byte; the strip's first row is judged by k_link_round, the other plane's rows by k_tile_link_x)
            for (int t = threadIdx.x; t < ny; t += blockDim.x) {
                smir[t] = t > 0 && rows_mirror(sz, srow[t] - gbase, srow[t + 1] - gbase, srow[t - 1] - gbase,
                                               srow[t] - gbase, spol[t] == spol[t - 1]);
                rowflag[r0 + t] = (uint8_t)(t > 0 && !smir[t]);
            }
            __syncthreads();
            // a stack of consecutively mirrored rows links run m of every row straight to run m of the stack's top
            // row (the correspondence m -> m is transitive), which keeps the trees one level deep
            if (threadIdx.x < 32) {                          // top[t] = last row <= t that is not mirrored: max-scan by one warp
                constexpr int RPL = STRIP_TY / 32;           // consecutive rows per lane
                static_assert(STRIP_TY % 32 == 0 && STRIP_TY <= 256, "rows per lane, uint8 row indices");
                uint32_t loc[RPL], run = 0;
#pragma unroll
                for (int q = 0; q < RPL; ++q) {
                    const int t = RPL * threadIdx.x + q;
                    if (t < ny && !smir[t]) run = t;
                    loc[q] = run;
                }
                uint32_t inc = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(MDVC_FULL, inc, o);
                    if ((int)threadIdx.x >= o) inc = max(inc, up);
                }
                uint32_t ex = __shfl_up_sync(MDVC_FULL, inc, 1);
                if (threadIdx.x == 0) ex = 0;
#pragma unroll
                for (int q = 0; q < RPL; ++q) {
                    const int t = RPL * threadIdx.x + q;
                    if (t < ny) stop[t] = (uint8_t)max(ex, loc[q]);
                }
            }
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < nloc; k += blockDim.x) {
                uint32_t e = sz[k], lr = e >> 24;
                if (lr == 0) continue;
                uint32_t rb0 = srow[lr] - gbase, re0 = srow[lr + 1] - gbase;      // own row [rb0, re0)
                uint32_t n0 = srow[lr - 1] - gbase, n1 = rb0;                     // previous row [n0, n1)
                if (smir[lr]) { smem_union(spar, k, (srow[stop[lr]] - gbase) + (k - rb0)); continue; }
                int a = (int)(e & ZMASK);
                int b = (k + 1 < re0) ? (int)(sz[k + 1] & ZMASK) - 1 : D.Z - 1;
                uint32_t pol = spol[lr] ^ ((k - rb0) & 1u);
                uint32_t lo = a > 0 ? a - 1 : 0, hi = b + 1 < D.Z ? b + 1 : D.Z - 1;
                uint32_t j = tile_run_at(sz, n0, n1, lo);
                if ((spol[lr - 1] ^ ((j - n0) & 1u)) != pol) ++j;
                for (; j < n1 && (sz[j] & ZMASK) <= hi; j += 2) smem_union(spar, k, j);
            }
            __syncthreads();
            if (threadIdx.x == 0) s_nroots = 0;
            // Every union hung a run under one of the previous row, so the trees are chains as deep as the strip
            // has rows (ncu: the plain climb below was 47 % of this kernel's instructions).  Pointer jumping
            // halves the depth per round; entries only ever move to an ancestor, so the rounds need no
            // double buffering.
            for (int round = 0; round < 7; ++round) {
                bool moved = false;
                for (uint32_t k = threadIdx.x; k < nloc; k += blockDim.x) {
                    const uint32_t p = spar[k], gp = spar[p];
                    if (gp != p) { spar[k] = gp; moved = true; }
                }
                if (!__syncthreads_or(moved)) break;
            }
            for (uint32_t k = threadIdx.x; k < nloc; k += blockDim.x) {
                uint32_t root = k, up;
                while ((up = spar[root]) != root) root = up;
                parent[gbase + k] = gbase + root;
                if (root == k) sz[atomicAdd(&s_nroots, 1u)] = gbase + k;     // sz is free again: collect the roots
            }
            __syncthreads();
            if (threadIdx.x == 0) s_base = atomicAdd(&h->n_strip_roots, s_nroots);
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < s_nroots; k += blockDim.x) rootlist[s_base + k] = sz[k];
        } else {
            // too many runs for shared memory: link the rows of this strip through global memory
            seen.clear();
            for (int t = threadIdx.x; t < ny; t += blockDim.x) rowflag[r0 + t] = (uint8_t)(t > 0);
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < nloc; k += blockDim.x) {
                uint32_t i = gbase + k;
                uint32_t lin = runstart[i];
                uint32_t r = lin / (uint32_t)D.Z;
                if (r == r0) continue;
                int a = (int)(lin - r * (uint32_t)D.Z);
                int b = (int)(runstart[i + 1] - r * (uint32_t)D.Z) - 1;
                uint32_t v = bit_at(bits, D, r, a);
                int lo = a > 0 ? a - 1 : 0, hi = b + 1 < D.Z ? b + 1 : D.Z - 1;
                uint32_t ra = climb_ro(parent, i);
                link_run_vs_row(bits, D, rowoff, runstart, parent, seen, i, ra, v, lo, hi, r - 1);
            }
            // every run of such a strip may end up an interior node: list them all
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < nloc; k += blockDim.x) rootlist[atomicAdd(&h->n_strip_roots, 1u)] = gbase + k;
        }
    }
}

// flatten the listed nodes (read-only climb, each entry written by its own thread)
__global__ void __launch_bounds__(256)
k_flatten_list(const Hdr *__restrict__ h, const uint32_t *__restrict__ list, uint32_t *parent) {
    uint32_t n = h->n_strip_roots, stride = gridDim.x * blockDim.x;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        uint32_t g = list[e];
        uint32_t p = __ldcg(&parent[g]);
        if (p == g) continue;
        uint32_t root = p, up;
        while ((up = __ldcg(&parent[root])) != root) root = up;
        if (root != p) parent[g] = root;
    }
}

// x-phase tile: TY rows of plane x against rows y0-1 .. y0+TY of plane x-1, run lists and current roots
// in shared memory; only new (root,root) pairs of the CTA reach the global union.
__global__ void __launch_bounds__(TILE_THREADS, TILE_MIN_CTAS)
k_tile_link_x(const uint32_t *__restrict__ bits, Dims D, const Hdr *__restrict__ h, const uint32_t *__restrict__ rowoff,
              const uint8_t *__restrict__ rowpol, const uint32_t *__restrict__ runstart, uint32_t *parent, int s, int radix,
              uint8_t *__restrict__ rowflag) {
    __shared__ uint32_t orow[STRIP_TY + 1], brow[STRIP_TY + 3];
    __shared__ uint32_t opol[STRIP_TY], bpol[STRIP_TY + 2];
    __shared__ uint32_t tz[STRIP_CAP], troot[STRIP_CAP];       // own runs first, then the other plane's
    __shared__ uint8_t smir[3 * STRIP_TY];
    __shared__ BlockSet seen;
    if (h->run_eff == 0) return;
    const int spp = (D.Y + STRIP_TY - 1) / STRIP_TY;
    const int cnt = (D.X - 1) / s;
    const long long n_tiles = (long long)cnt * spp;
    seen.clear();
    for (long long tid = blockIdx.x; tid < n_tiles; tid += gridDim.x) {
        const int m = (int)(tid / spp) + 1, y0 = (int)(tid % spp) * STRIP_TY;
        if (m % radix == 0) continue;
        const int x = m * s;
        const int ny = min(STRIP_TY, D.Y - y0);
        const int yA = max(y0 - 1, 0), yB = min(y0 + ny, D.Y - 1);           // rows of plane x-1 in reach
        const int nb = yB - yA + 1;
        const uint32_t r0 = (uint32_t)x * D.Y + y0, q0 = (uint32_t)(x - 1) * D.Y + yA;
        __syncthreads();
        for (int t = threadIdx.x; t <= ny; t += blockDim.x) {
            orow[t] = rowoff[r0 + t];
            if (t < ny) opol[t] = rowpol[r0 + t];
        }
        for (int t = threadIdx.x; t <= nb; t += blockDim.x) {
            brow[t] = rowoff[q0 + t];
            if (t < nb) bpol[t] = rowpol[q0 + t];
        }
        __syncthreads();
        const uint32_t go = orow[0], no = orow[ny] - go, gb = brow[0], nbr = brow[nb] - gb;
        if (no + nbr <= STRIP_CAP) {
            // current representative of every run: two hops without verification (after k_flatten_list the
            // trees are at most two deep; any node of the right tree is a valid union argument anyway)
            for (uint32_t k0 = threadIdx.x; k0 < no + nbr; k0 += 4 * blockDim.x) {
                uint32_t g[4], p1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t k = k0 + u * blockDim.x;
                    g[u] = k < no ? go + k : gb + (k - no);
                    p1[u] = k < no + nbr ? __ldcg(&parent[g[u]]) : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint32_t k = k0 + u * blockDim.x;
                    if (k >= no + nbr) continue;
                    uint32_t base_row = k < no ? r0 : q0;
                    uint32_t off = runstart[g[u]] - base_row * (uint32_t)D.Z;
                    uint32_t lr = off / (uint32_t)D.Z;
                    tz[k] = (lr << 24) | (off - lr * (uint32_t)D.Z);
                    troot[k] = __ldcg(&parent[p1[u]]);
                }
            }
            __syncthreads();
            // (own row, dy) pairs whose rows mirror each other: the link partner of run m is run m, no search
            for (int t = threadIdx.x; t < 3 * ny; t += blockDim.x) {
                const int lr = t / 3, dy = t - 3 * lr - 1, yy = y0 + lr + dy;
                bool mir = false;
                if (yy >= 0 && yy < D.Y) {
                    const int bl = yy - yA;
                    mir = rows_mirror(tz, orow[lr] - go, orow[lr + 1] - go, no + (brow[bl] - gb), no + (brow[bl + 1] - gb),
                                      opol[lr] == bpol[bl]);
                }
                smir[t] = mir;
            }
            __syncthreads();
            for (int lr = threadIdx.x; lr < ny; lr += blockDim.x) {          // rowflag bits 1..3: (-1,-1) (-1,0) (-1,+1)
                uint32_t f = 0;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y0 + lr + dy;
                    if (yy >= 0 && yy < D.Y && !smir[3 * lr + dy + 1]) f |= 2u << (dy + 1);
                }
                if (f) rowflag[r0 + lr] |= (uint8_t)f;
            }
            for (uint32_t k = threadIdx.x; k < no; k += blockDim.x) {
                uint32_t e = tz[k], lr = e >> 24;
                uint32_t rb0 = orow[lr] - go, re0 = orow[lr + 1] - go;
                int a = (int)(e & ZMASK);
                int b = (k + 1 < re0) ? (int)(tz[k + 1] & ZMASK) - 1 : D.Z - 1;
                uint32_t pol = opol[lr] ^ ((k - rb0) & 1u);
                uint32_t lo = a > 0 ? a - 1 : 0, hi = b + 1 < D.Z ? b + 1 : D.Z - 1;
                uint32_t ra = troot[k];
                for (int dy = -1; dy <= 1; ++dy) {
                    int yy = y0 + (int)lr + dy;
                    if (yy < 0 || yy >= D.Y) continue;
                    int bl = yy - yA;                                          // row slot in the other plane
                    uint32_t n0 = no + (brow[bl] - gb), n1 = no + (brow[bl + 1] - gb);
                    uint32_t j, jend;
                    if (smir[3 * lr + dy + 1]) { j = n0 + (k - rb0); jend = j + 1; }
                    else {
                        j = tile_run_at(tz, n0, n1, lo);
                        if ((bpol[bl] ^ ((j - n0) & 1u)) != pol) ++j;
                        jend = n1;
                    }
                    for (; j < jend && (tz[j] & ZMASK) <= hi; j += 2) {
                        uint32_t rb = troot[j];
                        if (ra == rb) continue;
                        unsigned long long key = ra < rb ? ((unsigned long long)ra << 32) | rb : ((unsigned long long)rb << 32) | ra;
                        if (!seen.add(key)) continue;
                        rem_union(parent, ra, rb);
                        ra = ra < rb ? ra : rb;
                    }
                }
            }
        } else {
            for (int lr = threadIdx.x; lr < ny; lr += blockDim.x) rowflag[r0 + lr] |= (uint8_t)0xe;
            for (uint32_t k = threadIdx.x; k < no; k += blockDim.x) {
                uint32_t i = go + k;
                uint32_t lin = runstart[i];
                uint32_t r = lin / (uint32_t)D.Z;
                int y = (int)(r - (uint32_t)x * D.Y);
                int a = (int)(lin - r * (uint32_t)D.Z);
                int b = (int)(runstart[i + 1] - r * (uint32_t)D.Z) - 1;
                uint32_t v = bit_at(bits, D, r, a);
                int lo = a > 0 ? a - 1 : 0, hi = b + 1 < D.Z ? b + 1 : D.Z - 1;
                uint32_t ra = climb_ro(parent, i);
                for (int dy = -1; dy <= 1; ++dy) {
                    int yy = y + dy;
                    if (yy < 0 || yy >= D.Y) continue;
                    link_run_vs_row(bits, D, rowoff, runstart, parent, seen, i, ra, v, lo, hi, (uint32_t)(x - 1) * D.Y + yy);
                }
            }
        }
    }
}

// stage 1d: flatten + root numbering.  Labels follow scipy's order: the k-th True (False) root in run order is
// label k+1 (-(k+1)).  One CTA per tile of ROOT_TILE consecutive runs, eight consecutive runs per thread:
//   - read-only climb to the root and flatten (parent[i] = root);
//   - rank of every root inside its tile (block scan of packed 16+16-bit counts) -> localrank[root]
//     (bits 0-11: True roots before it, bits 16-27: False roots before it, bit 15: the root's own value, which
//     is the value of every run of its tree -- k_assign_labels needs no look-up in the bit grid);
//   - roots of the tile -> tile_sums[tile] (lo32 True, hi32 False).
// k_root_tile_offsets turns the tile sums into exclusive offsets, and k_assign_labels adds the two: no dense
// scan of one flag per run (was three more launches over 8 B per run).
__device__ __forceinline__ uint32_t ld_ca_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

#define ROOT_TILE 2048
#define ROOT_ITEMS 8
__global__ void __launch_bounds__(256)
k_flatten_flags(const uint32_t *__restrict__ bits, Dims D, const Hdr *__restrict__ h,
                const uint32_t *__restrict__ runstart, uint32_t *parent, uint32_t *__restrict__ localrank,
                unsigned long long *__restrict__ tile_sums) {
    __shared__ uint32_t sm[33];
    const uint32_t R = h->run_eff;
    const uint32_t base = blockIdx.x * ROOT_TILE;
    if (base >= R) return;
    const uint32_t i0 = base + threadIdx.x * ROOT_ITEMS;
    // read-only climb: in this kernel parent[i] is written by thread i alone, so after the kernel every entry
    // holds a root (a pointer-jumping find here could overwrite a neighbour's finished entry with a non-root
    // ancestor).  The eight climbs advance in lockstep, one round of eight independent loads at a time.
    // The loads may hit L1 (unlike in the union kernels): the only stores of this kernel replace an entry by its
    // root and roots do not change here, so a stale line still holds an ancestor and the climb still ends at the
    // root.  All runs of a component converge on a handful of root entries; served from L2 alone those few
    // sectors serialise the whole grid.
    uint32_t cur[ROOT_ITEMS], f[ROOT_ITEMS];
#pragma unroll
    for (int k = 0; k < ROOT_ITEMS; ++k) cur[k] = min(i0 + k, R - 1);
    uint32_t live = (1u << ROOT_ITEMS) - 1u;
    do {
        uint32_t up[ROOT_ITEMS];
#pragma unroll
        for (int k = 0; k < ROOT_ITEMS; ++k) up[k] = (live >> k) & 1u ? ld_ca_u32(parent + cur[k]) : cur[k];
#pragma unroll
        for (int k = 0; k < ROOT_ITEMS; ++k) {
            if (up[k] == cur[k]) live &= ~(1u << k);
            cur[k] = up[k];
        }
    } while (live);
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < ROOT_ITEMS; ++k) {
        const uint32_t i = i0 + k;
        f[k] = 0;
        if (i < R) {
            if (cur[k] == i) {
                uint32_t lin = runstart[i];
                uint32_t r = lin / (uint32_t)D.Z;
                f[k] = bit_at(bits, D, r, (int)(lin - r * (uint32_t)D.Z)) ? 1u : (1u << 16);
            } else {
                parent[i] = cur[k];
            }
        }
        mine += f[k];
    }
    uint32_t total;
    uint32_t off = mdvc_block_exscan<uint32_t>(mine, sm, &total);
#pragma unroll
    for (int k = 0; k < ROOT_ITEMS; ++k) {
        if (f[k]) localrank[i0 + k] = off | ((f[k] & 1u) << 15);      // bit 15: the tree is a True component
        off += f[k];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = (unsigned long long)(total & 0xffffu) | ((unsigned long long)(total >> 16) << 32);
}

// single block: exclusive scan of the root tile sums of the runs in use; total -> h->tot_roots
__global__ void __launch_bounds__(1024)
k_root_tile_offsets(Hdr *h, unsigned long long *tile_sums) {
    __shared__ unsigned long long sm[33];
    const uint32_t n_tiles = (h->run_eff + ROOT_TILE - 1) / ROOT_TILE;
    const uint32_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(threadIdx.x * per, n_tiles), hi = min(lo + per, n_tiles);
    unsigned long long acc = 0;
    for (uint32_t i = lo; i < hi; ++i) acc += tile_sums[i];
    unsigned long long total;
    unsigned long long off = mdvc_block_exscan<unsigned long long>(acc, sm, &total);
    for (uint32_t i = lo; i < hi; ++i) { unsigned long long v = tile_sums[i]; tile_sums[i] = off; off += v; }
    if (threadIdx.x == 0) h->tot_roots = total;
}

// publishes the label totals and zeroes the volume table of exactly the labels in use (instead of a memset of
// the whole max_labels table, 50 MB at 1024^3, every frame)
__global__ void __launch_bounds__(256)
k_after_root_scan(Hdr *h, uint32_t max_labels, unsigned long long *__restrict__ vol) {
    const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
    if (h->run_eff == 0) { if (leader) h->n_labels = 0; return; }
    uint32_t nt = (uint32_t)(h->tot_roots & 0xffffffffu), nf = (uint32_t)(h->tot_roots >> 32);
    unsigned long long nl = (unsigned long long)nt + nf + 1;
    const bool bad = nl > max_labels || nt >= MDVC_LABEL_BIAS || nf >= MDVC_LABEL_BIAS;
    if (leader) {
        h->n_true = nt; h->n_false = nf;
        if (bad) { h->flags |= MDVC_OVERFLOW_LABELS; h->n_labels = 0; }
        else h->n_labels = (uint32_t)nl;
    }
    if (bad) return;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nl;
         i += (unsigned long long)gridDim.x * blockDim.x)
        vol[i] = 0ull;
}

// stage 1e: canonical label per run + label volumes
// Volumes: consecutive runs share very few labels, so each CTA first accumulates into a small
// shared-memory table (label -> voxels) and flushes it once; labels that do not fit go straight to
// global atomics (the many-label regime, where contention is low anyway).
#define VOL_SLOTS 256
__global__ void __launch_bounds__(256)
k_assign_labels(const uint32_t *__restrict__ bits, Dims D, const Hdr *__restrict__ h,
                const uint32_t *__restrict__ runstart, const uint32_t *__restrict__ parent,
                const uint32_t *__restrict__ localrank, const unsigned long long *__restrict__ tile_off,
                int32_t *__restrict__ lab, unsigned long long *vol) {
    __shared__ int s_key[VOL_SLOTS];
    __shared__ unsigned long long s_val[VOL_SLOTS];
    for (int k = threadIdx.x; k < VOL_SLOTS; k += blockDim.x) { s_key[k] = 0; s_val[k] = 0; }
    __syncthreads();
    uint32_t R = h->run_eff;
    if (h->n_labels == 0) return;
    uint32_t nf = h->n_false;
    uint32_t stride = gridDim.x * blockDim.x;
    uint32_t rounds = (R + stride - 1) / stride;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t it = 0; it < rounds; ++it, i += stride) {
        bool live = i < R;
        int label = 0;
        uint32_t len = 0;
        if (live) {
        const uint32_t lin = runstart[i];
        const uint32_t root = parent[i];
        const unsigned long long e = tile_off[root / ROOT_TILE];
        const uint32_t lr = localrank[root];
        label = (lr & 0x8000u) ? (int)((uint32_t)(e & 0xffffffffu) + (lr & 0xfffu) + 1u)
                               : -(int)((uint32_t)(e >> 32) + ((lr >> 16) & 0xfffu) + 1u);
        lab[i] = label;
        len = runstart[i + 1] - lin;
        }
        // one shared-memory update per distinct label of the warp
        unsigned peers = __match_any_sync(MDVC_FULL, label);
        uint32_t sum = __reduce_add_sync(peers, len);
        if (!live || (threadIdx.x & 31) != (__ffs(peers) - 1)) continue;
        // shared table: key 0 = empty (labels are never 0)
        uint32_t slot = ((uint32_t)label * 2654435761u) >> 24;
        bool done = false;
        for (int probe = 0; probe < 8 && !done; ++probe) {
            int cur = s_key[slot];
            if (cur == 0) cur = atomicCAS(&s_key[slot], 0, label);
            if (cur == 0 || cur == label) { atomicAdd(&s_val[slot], (unsigned long long)sum); done = true; }
            slot = (slot + 1) & (VOL_SLOTS - 1);
        }
        if (!done) atomicAdd(&vol[(long long)label + nf], (unsigned long long)sum);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < VOL_SLOTS; k += blockDim.x)
        if (s_key[k] != 0 && s_val[k]) atomicAdd(&vol[(long long)s_key[k] + nf], s_val[k]);
}

// ------------------------------------------------------------------------------------------
// stage 2: contacts and bridges on runs
// ------------------------------------------------------------------------------------------
struct PairSink {
    BlockSet *cache;
    unsigned long long *slots;
    uint32_t n_slots;
    uint32_t *flags;
    uint32_t overflow_bit;
    unsigned long long last, last2;       // the two most recent keys of this thread
    // fast path inline, slow path (block cache + global set) out of line: the pair kernels call this from
    // ~20 sites and inlining the probing loops there blew the kernel up to 44k instructions
    __device__ __forceinline__ void put(unsigned long long key) {
        if (key == last || key == last2) return;
        put_slow(key);
    }
    __device__ __noinline__ void put_slow(unsigned long long key) {
        last2 = last;
        last = key;
        if (cache->add_state(key) != 2) return;                       // known, or parked until flush()
        if (*(volatile uint32_t *)flags & overflow_bit) return;       // already overflowed: this attempt is void anyway
        if (!mdvc_set_insert(slots, n_slots, key)) atomicOr(flags, overflow_bit);
    }
    // all threads of the CTA, once, after the last put
    __device__ void flush() {
        __syncthreads();
        cache->flush(slots, n_slots, flags, overflow_bit);
    }
};

__device__ __forceinline__ void emit_contact(PairSink &s, int a, int b) {
    if (a == b) return;
    s.put(a < b ? mdvc_pack_key(a, b, 0) : mdvc_pack_key(b, a, 0));
}
// bridge seen from a voxel labelled la stepping by image shift (sx,sy,sz) onto a voxel labelled lb;
// the reverse step is emitted too when the labels are equal (find_bridges.pyx:84-87)
__device__ __forceinline__ void emit_bridge(PairSink &s, int la, int lb, int sx, int sy, int sz) {
    if (la < lb) s.put(mdvc_pack_key(la, lb, mdvc_shift_code(sx, sy, sz)));
    else if (la > lb) s.put(mdvc_pack_key(lb, la, mdvc_shift_code(-sx, -sy, -sz)));
    else {
        s.put(mdvc_pack_key(la, la, mdvc_shift_code(sx, sy, sz)));
        s.put(mdvc_pack_key(la, la, mdvc_shift_code(-sx, -sy, -sz)));
    }
}

// Stage 2b, the row pairs that do need their run lists compared.  Relation q of a row: 0 = (0,-1) in the own plane,
// 1..3 = (-1, dy = q-2) in the plane before; every unordered pair of neighbouring rows is one relation of its later
// row.  In-box relations are skipped when the two rows mirror each other: same number of runs, same first value, and
// starts that interleave (every run starts before the next run of the other row: s'_m < s_{m+1} and s_m < s'_{m+1}).
// Then run k overlaps run k of the neighbour voxel on voxel (same value, adjacent rows: one 26-connected set, hence the
// same label) and otherwise only reaches its runs k-1 and k+1, whose labels are this row's own neighbours' labels, so
// every cross-row contact is already among the in-row contacts (k_row_pairs).  The link kernels ran that test on the
// same row pairs and left the outcome in rowflag (bit q set = not mirrored): 1-2 % of the rows of a smooth grid.
// Relations through a periodic x or y face (bridges) are always walked.
__device__ __forceinline__ uint32_t pair_relations(const Dims &D, int periodic, const uint8_t *__restrict__ rowflag,
                                                   uint32_t row, int x, int y) {
    uint32_t need = rowflag[row];                                  // in-box relations that are not mirrored
    if (periodic) {
        if (y == 0) need |= 1u;
        if (x != 0 || periodic == 1) {                             // periodic == 2: the x faces belong to the caller
            if (x == 0 || y == 0) need |= 2u;
            if (x == 0) need |= 4u;
            if (x == 0 || y == D.Y - 1) need |= 8u;
        }
    }
    return need;
}

// Stage 2a, one thread per row: everything the pair stage finds without comparing run lists --
//   * contacts between consecutive runs of a row (they always differ in value);
//   * bridges through the z faces: the last run steps +z into the first runs of the six rows with dx in {0,-1}, the
//     first run steps -z into the last runs of the three rows with dx = -1 (the reverse of the dx = +1, +z steps), so
//     only rows at or before this one in raster order are read;
//   * the list of rows whose run lists have to be compared (k_task_pairs).
// The eleven labels of a row's z-face steps (first runs of rows y-1..y+1 of both planes, last runs of the other
// plane's, the row's own first and last) almost always repeat from one row to the next, and a row whose labels equal
// the previous row's -- neither of them on an x or y face, where the shift vectors differ -- has nothing new to say.
// Every lane therefore loads the four labels of its own row only (first / last run here and in the plane before);
// the neighbours' come by shuffle, and "equal to the previous row" is three ballots.  A warp emits for 29 consecutive
// rows: lanes 0 and 1 hold the two rows before them, lane 31 the row after (comparison only).  Rows on a face take
// the slow path with explicit, wrapped loads.
// periodic: 0 none, 1 all axes, 2 y and z only (slab interior: the x faces belong to the caller).
#define RP_ROWS 29
#define RP_LIST 1024
#ifndef RP_MIN_CTAS
#define RP_MIN_CTAS 5
#endif
__global__ void __launch_bounds__(256, RP_MIN_CTAS)
k_row_pairs(Dims D, Hdr *h, const uint32_t *__restrict__ rowoff, const uint8_t *__restrict__ rowflag,
            const int32_t *__restrict__ lab, uint32_t *__restrict__ pairrows, unsigned long long *cset, uint32_t c_slots,
            unsigned long long *bset, uint32_t b_slots, int periodic) {
    __shared__ BlockSet c_cache, b_cache;
    __shared__ uint32_t s_list[RP_LIST], s_listed, s_valid, s_base;
    c_cache.clear(); b_cache.clear();
    if (threadIdx.x == 0) { s_listed = 0; s_valid = RP_LIST; }
    __syncthreads();
    if (h->n_labels == 0) return;
    PairSink cs{&c_cache, cset, c_slots, &h->flags, MDVC_OVERFLOW_CONTACTS, MDVC_KEY_EMPTY, MDVC_KEY_EMPTY};
    PairSink bs{&b_cache, bset, b_slots, &h->flags, MDVC_OVERFLOW_BRIDGES, MDVC_KEY_EMPTY, MDVC_KEY_EMPTY};
    const int lane = threadIdx.x & 31;
    const uint32_t n_rows = (uint32_t)D.rows;                      // < 2^31
    const uint32_t n_warps = (n_rows + RP_ROWS - 1) / RP_ROWS, warp_stride = (gridDim.x * blockDim.x) >> 5;
    int c_lo = 0, c_hi = 0, c_lo2 = 0, c_hi2 = 0;                  // the two contact pairs this thread emitted last
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_warps; w += warp_stride) {
        const uint32_t row = w * RP_ROWS + (uint32_t)lane - 2u;    // lanes 0, 1 of warp 0: wrapped around, not live
        const bool live = row < n_rows, mine = live && lane >= 2 && lane <= RP_ROWS + 1;
        const int x = live ? (int)(row / (uint32_t)D.Y) : 0, y = live ? (int)(row - (uint32_t)x * (uint32_t)D.Y) : 0;
        const int xo = x == 0 ? D.X - 1 : x - 1, sxo = x == 0 ? -1 : 0;
        const bool other_ok = x != 0 || periodic == 1;
        const uint32_t own0 = (uint32_t)x * D.Y, oth0 = (uint32_t)xo * D.Y;
        const uint32_t b = live ? rowoff[row] : 0u, e = live ? rowoff[row + 1] : 1u;
        {   // rows with relations to walk -> k_task_pairs' list: gathered per CTA (thousands of warps adding to one
            // global counter queue up behind each other), straight to the global list only when the buffer is full
            const uint32_t need = mine ? pair_relations(D, periodic, rowflag, row, x, y) : 0u;
            const unsigned m = __ballot_sync(MDVC_FULL, need != 0);
            if (m) {
                const uint32_t cnt = (uint32_t)__popc(m), rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
                uint32_t base = 0;
                if (lane == 0) {
                    base = atomicAdd(&s_listed, cnt);
                    if (base + cnt > RP_LIST) {                  // buffer full from `base` on
                        atomicMin(&s_valid, base);
                        base = 0x80000000u | atomicAdd(&h->n_pair_rows, cnt);
                    }
                }
                base = __shfl_sync(MDVC_FULL, base, 0);
                if (need) {
                    if (base & 0x80000000u) pairrows[(base & 0x7fffffffu) + rank] = row;
                    else s_list[base + rank] = row;
                }
            }
        }
        const int first = live ? lab[b] : 0;
        if (periodic) {
            const bool face = y == 0 || y == D.Y - 1 || x == 0;
            const bool ld_o = live && other_ok;
            const uint32_t ob = ld_o ? rowoff[oth0 + y] : 0u, oe = ld_o ? rowoff[oth0 + y + 1] : 1u;
            const int last = live ? lab[e - 1] : 0, ofirst = ld_o ? lab[ob] : 0, olast = ld_o ? lab[oe - 1] : 0;
            // the rows before (m) and after (p) this one, valid when this row is not on a face
            const int f_m = __shfl_up_sync(MDVC_FULL, first, 1), f_p = __shfl_down_sync(MDVC_FULL, first, 1);
            const int of_m = __shfl_up_sync(MDVC_FULL, ofirst, 1), of_p = __shfl_down_sync(MDVC_FULL, ofirst, 1);
            const int ol_m = __shfl_up_sync(MDVC_FULL, olast, 1), ol_p = __shfl_down_sync(MDVC_FULL, olast, 1);
            const int l_m = __shfl_up_sync(MDVC_FULL, last, 1);
            const unsigned eq3 = __ballot_sync(MDVC_FULL, first == f_m && ofirst == of_m && olast == ol_m);
            const unsigned eqw = __ballot_sync(MDVC_FULL, last == l_m);
            const unsigned inner = __ballot_sync(MDVC_FULL, live && !face);
            // equal to the previous row <=> rows y-2..y+1 agree on the three neighbour labels, rows y-1, y on the last
            const unsigned span = lane ? 7u << (lane - 1) : 0u;    // bits lane-1, lane, lane+1 (used by lanes 2..30)
            const bool same = mine && (inner >> lane & 1u) && (inner >> (lane - 1) & 1u) && (eq3 & span) == span &&
                              (eqw >> lane & 1u);
            if (mine && !same) {
                int cur[9], sy3[3] = {0, 0, 0};
                if (!face) {
                    cur[0] = f_m; cur[1] = first; cur[2] = f_p;
                    cur[3] = of_m; cur[4] = ofirst; cur[5] = of_p;
                    cur[6] = ol_m; cur[7] = olast; cur[8] = ol_p;
                } else {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        int yy = y + d - 1;
                        if (yy < 0) { yy = D.Y - 1; sy3[d] = -1; } else if (yy >= D.Y) { yy = 0; sy3[d] = 1; }
                        cur[d] = lab[rowoff[own0 + yy]];
                        cur[3 + d] = other_ok ? lab[rowoff[oth0 + yy]] : 0;
                        cur[6 + d] = other_ok ? lab[rowoff[oth0 + yy + 1] - 1u] : 0;
                    }
                }
                // q 0..2: +z into the first runs of the own plane's rows y-1, y, y+1; 3..5: of the other plane's;
                // 6..8: -z into the last runs of the other plane's
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    if (q >= 3 && !other_ok) break;
                    emit_bridge(bs, q < 6 ? last : first, cur[q], q < 3 ? 0 : sxo, sy3[q % 3], q < 6 ? 1 : -1);
                }
            }
        }
        if (mine) {
            int before = first;
            for (uint32_t k0 = b + 1; k0 < e; k0 += 4) {           // four labels in flight (the loop is load latency)
                int now[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) now[u] = k0 + u < e ? lab[k0 + u] : 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (k0 + u >= e) break;
                    const int lo = min(before, now[u]), hi = max(before, now[u]);
                    before = now[u];
                    if ((lo == c_lo && hi == c_hi) || (lo == c_lo2 && hi == c_hi2) || lo == hi) continue;
                    c_lo2 = c_lo; c_hi2 = c_hi; c_lo = lo; c_hi = hi;
                    cs.put(mdvc_pack_key(lo, hi, 0));
                }
            }
        }
    }
    __syncthreads();
    const uint32_t n_listed = min(s_listed, s_valid);              // reservations that did not fit went to the global list
    if (threadIdx.x == 0 && n_listed) s_base = atomicAdd(&h->n_pair_rows, n_listed);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_listed; i += blockDim.x) pairrows[s_base + i] = s_list[i];
    cs.flush();
    bs.flush();
}

// One task = one listed row; PR_LANES lanes share it, each takes every PR_LANES-th run A = [a,b] and walks the runs of
// the neighbour row that A reaches: contacts for in-box neighbours (runs of the other value that overlap [a,b] itself
// -- a run that merely touches diagonally sits next to a run of A's value that overlaps it, same label as A, so that
// pair is one of the neighbour row's own in-row contacts), bridges for wrapped neighbours (every run within one voxel
// of [a,b], any value, with the image shift).  The reverse steps give the same normalised rows (emit_bridge).
#define PR_LANES 8
__global__ void __launch_bounds__(256)
k_task_pairs(Dims D, Hdr *h, const uint32_t *__restrict__ rowoff, const uint8_t *__restrict__ rowpol,
             const uint8_t *__restrict__ rowflag, const uint32_t *__restrict__ tasks, const uint32_t *__restrict__ runstart,
             const int32_t *__restrict__ lab, unsigned long long *cset, uint32_t c_slots, unsigned long long *bset,
             uint32_t b_slots, int periodic) {
    __shared__ BlockSet c_cache, b_cache;
    if (h->n_labels == 0) return;
    const uint32_t n_tasks = h->n_pair_rows;
    if (blockIdx.x * (blockDim.x / PR_LANES) >= n_tasks) return;
    c_cache.clear(); b_cache.clear();
    __syncthreads();
    PairSink cs{&c_cache, cset, c_slots, &h->flags, MDVC_OVERFLOW_CONTACTS, MDVC_KEY_EMPTY, MDVC_KEY_EMPTY};
    PairSink bs{&b_cache, bset, b_slots, &h->flags, MDVC_OVERFLOW_BRIDGES, MDVC_KEY_EMPTY, MDVC_KEY_EMPTY};
    const uint32_t sub = threadIdx.x % PR_LANES;
    const uint32_t stride = gridDim.x * (blockDim.x / PR_LANES);
    for (uint32_t t = blockIdx.x * (blockDim.x / PR_LANES) + threadIdx.x / PR_LANES; t < n_tasks; t += stride) {
        const uint32_t r = tasks[t];
        const int x = (int)(r / (uint32_t)D.Y), y = (int)(r - (uint32_t)x * (uint32_t)D.Y);
        const uint32_t need = pair_relations(D, periodic, rowflag, r, x, y);
        const uint32_t rb = rowoff[r], re = rowoff[r + 1], zr = r * (uint32_t)D.Z;
        const uint32_t pol0 = rowpol[r];
        // the (up to four) neighbour rows: run range, first value, image shift
        uint32_t nb[4], ne[4], npol[4], nz[4];
        int nsx[4], nsy[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int nx = q == 0 ? x : x - 1, ny = q == 0 ? y - 1 : y + q - 2;
            nsx[q] = 0; nsy[q] = 0;
            if (nx < 0) { nx = D.X - 1; nsx[q] = -1; }
            if (ny < 0) { ny = D.Y - 1; nsy[q] = -1; } else if (ny >= D.Y) { ny = 0; nsy[q] = 1; }
            const uint32_t r2 = (uint32_t)nx * D.Y + ny;
            const bool on = (need >> q) & 1u;
            nb[q] = on ? rowoff[r2] : 0u; ne[q] = on ? rowoff[r2 + 1] : 0u;
            npol[q] = on ? rowpol[r2] : 0u; nz[q] = r2 * (uint32_t)D.Z;
        }
        for (uint32_t i = rb + sub; i < re; i += PR_LANES) {
            const uint32_t a = runstart[i] - zr, b = (i + 1 < re ? runstart[i + 1] - zr : (uint32_t)D.Z) - 1u;
            const uint32_t pol = pol0 ^ ((i - rb) & 1u);
            const int la = lab[i];
            const uint32_t lo = a > 0 ? a - 1 : 0, hi = b + 1 < (uint32_t)D.Z ? b + 1 : (uint32_t)D.Z - 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (nb[q] == ne[q]) continue;
                const bool wrapped = (nsx[q] | nsy[q]) != 0;
                const uint32_t from = (wrapped ? lo : a) + nz[q];
                uint32_t j = nb[q], top = ne[q];                   // last run of the row that starts at or before `from`
                while (top - j > 1) {
                    const uint32_t mid = (j + top) >> 1;
                    if (runstart[mid] <= from) j = mid; else top = mid;
                }
                if (!wrapped) {
                    if ((npol[q] ^ ((j - nb[q]) & 1u)) == pol) ++j;
                    for (; j < ne[q] && runstart[j] <= b + nz[q]; j += 2) emit_contact(cs, la, lab[j]);
                } else {
                    for (; j < ne[q] && runstart[j] <= hi + nz[q]; ++j) emit_bridge(bs, la, lab[j], nsx[q], nsy[q], 0);
                }
            }
        }
    }
    cs.flush();
    bs.flush();
}

// compact a hash set into a dense key list
__global__ void __launch_bounds__(256)
k_compact_set(const unsigned long long *__restrict__ slots, uint32_t n_slots, unsigned long long *__restrict__ keys,
              uint32_t *count) {
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        unsigned long long k = slots[i];
        if (k != MDVC_KEY_EMPTY) keys[atomicAdd(count, 1u)] = k;
    }
}

// single-block bitonic sort of keys[0..*count): shared memory when it fits, global memory otherwise
#define SORT_SMEM_KEYS 4096
__global__ void __launch_bounds__(1024)
k_sort_keys(unsigned long long *keys, const uint32_t *count, uint32_t cap) {
    __shared__ unsigned long long sk[SORT_SMEM_KEYS];
    uint32_t n = *count;
    if (n > cap) n = cap;
    if (n < 2) return;
    uint32_t P = 1;
    while (P < n) P <<= 1;
    if (P > cap) return;   // cannot pad (cap is a power of two >= n, so this does not happen)
    bool in_smem = P <= SORT_SMEM_KEYS;
    unsigned long long *a = in_smem ? sk : keys;
    if (in_smem) {
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) sk[i] = i < n ? keys[i] : MDVC_KEY_EMPTY;
    } else {
        for (uint32_t i = n + threadIdx.x; i < P; i += blockDim.x) keys[i] = MDVC_KEY_EMPTY;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long u = a[i], w = a[ixj];
                    bool up = (i & k) == 0;
                    if ((u > w) == up) { a[i] = w; a[ixj] = u; }
                }
            }
            __syncthreads();
        }
    }
    if (in_smem)
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) keys[i] = sk[i];
}

// Multi-block bitonic sort for key sets that do not fit one CTA's shared memory (many-label grids: millions of
// distinct contacts).  The capacity is a power of two; the live count is on the device, so every kernel works on
// P = next power of two >= count and tiles beyond P leave at once.  Compare-exchange distances below BSORT_TILE run
// in shared memory (one launch per merge size k), the larger ones take one launch per (k, j).
#define BSORT_TILE 4096
__device__ __forceinline__ uint32_t bsort_padded(const uint32_t *count, uint32_t cap) {
    uint32_t n = min(*count, cap), P = 2;
    while (P < n) P <<= 1;
    return P;
}
__global__ void __launch_bounds__(256)
k_bsort_pad(unsigned long long *keys, const uint32_t *count, uint32_t cap) {
    const uint32_t n = min(*count, cap), P = bsort_padded(count, cap);
    for (uint32_t i = n + blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) keys[i] = MDVC_KEY_EMPTY;
}
// k_lo == 0: sort every tile completely (k = 2 .. BSORT_TILE); otherwise the passes j = BSORT_TILE/2 .. 1 of merge size k_lo
__global__ void __launch_bounds__(1024)
k_bsort_local(unsigned long long *keys, const uint32_t *count, uint32_t cap, uint32_t k_lo) {
    __shared__ unsigned long long sk[BSORT_TILE];
    const uint32_t P = bsort_padded(count, cap);
    const uint32_t base = blockIdx.x * BSORT_TILE;
    if (base >= P || (k_lo && k_lo > P)) return;
    const uint32_t m = min((uint32_t)BSORT_TILE, P - base);           // P < BSORT_TILE: one partial tile (power of two)
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) sk[i] = keys[base + i];
    __syncthreads();
    for (uint32_t k = k_lo ? k_lo : 2; k <= (k_lo ? k_lo : m); k <<= 1) {
        for (uint32_t j = min(k >> 1, m >> 1); j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long u = sk[i], w = sk[ixj];
                    const bool up = ((base + i) & k) == 0;
                    if ((u > w) == up) { sk[i] = w; sk[ixj] = u; }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) keys[base + i] = sk[i];
}
__global__ void __launch_bounds__(256)
k_bsort_global(unsigned long long *keys, const uint32_t *count, uint32_t cap, uint32_t k, uint32_t j) {
    const uint32_t P = bsort_padded(count, cap);
    if (k > P) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const uint32_t ixj = i ^ j;
        if (ixj > i) {
            const unsigned long long u = keys[i], w = keys[ixj];
            const bool up = (i & k) == 0;
            if ((u > w) == up) { keys[i] = w; keys[ixj] = u; }
        }
    }
}

// ascending sort of keys[0 .. *count): one CTA while the capacity is small, the multi-block network otherwise
static int sort_keys(unsigned long long *keys, const uint32_t *count, uint32_t cap, cudaStream_t st) {
    if (cap <= 4 * BSORT_TILE) {
        k_sort_keys<<<1, 1024, 0, st>>>(keys, count, cap);
        MDVC_LAUNCH_CHECK();
        return MDVC_OK;
    }
    const unsigned tiles = cap / BSORT_TILE;
    k_bsort_pad<<<mdvc_grid(cap, 256, 4), 256, 0, st>>>(keys, count, cap);
    MDVC_LAUNCH_CHECK();
    k_bsort_local<<<tiles, 1024, 0, st>>>(keys, count, cap, 0);
    MDVC_LAUNCH_CHECK();
    for (uint32_t k = 2 * BSORT_TILE; k != 0 && k <= cap; k <<= 1) {
        for (uint32_t j = k >> 1; j >= BSORT_TILE; j >>= 1) {
            k_bsort_global<<<mdvc_grid(cap, 256, 8), 256, 0, st>>>(keys, count, cap, k, j);
            MDVC_LAUNCH_CHECK();
        }
        k_bsort_local<<<tiles, 1024, 0, st>>>(keys, count, cap, k);
        MDVC_LAUNCH_CHECK();
    }
    return MDVC_OK;
}

__global__ void __launch_bounds__(256)
k_decode_rows(const unsigned long long *__restrict__ keys, const uint32_t *count, uint32_t cap, int with_shift,
              int32_t *__restrict__ rows) {
    uint32_t n = min(*count, cap);
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int l1, l2, code;
        mdvc_unpack_key(keys[i], l1, l2, code);
        if (with_shift) {
            int32_t *o = rows + (size_t)i * 5;
            o[0] = l1; o[1] = l2; o[2] = code / 9 - 1; o[3] = (code / 3) % 3 - 1; o[4] = code % 3 - 1;
        } else {
            rows[2 * (size_t)i] = l1; rows[2 * (size_t)i + 1] = l2;
        }
    }
}

// Small sets (the default capacity): compaction, sort and decoding of one set by one CTA -- block 0 the contacts,
// block 1 the bridges -- instead of three launches per set.
__global__ void __launch_bounds__(1024)
k_finish_sets(const unsigned long long *__restrict__ cset, unsigned long long *ckeys, int32_t *__restrict__ crows,
              uint32_t *n_contacts, uint32_t c_cap, const unsigned long long *__restrict__ bset, unsigned long long *bkeys,
              int32_t *__restrict__ brows, uint32_t *n_bridges, uint32_t b_cap) {
    __shared__ unsigned long long sk[SORT_SMEM_KEYS];
    __shared__ uint32_t s_n;
    const bool bridges = blockIdx.x == 1;
    const unsigned long long *set = bridges ? bset : cset;
    unsigned long long *keys = bridges ? bkeys : ckeys;
    int32_t *rows = bridges ? brows : crows;
    const uint32_t cap = bridges ? b_cap : c_cap;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) {
        const unsigned long long k = set[i];
        if (k == MDVC_KEY_EMPTY) continue;
        const uint32_t at = atomicAdd(&s_n, 1u);
        keys[at] = k;
        if (at < SORT_SMEM_KEYS) sk[at] = k;
    }
    __syncthreads();
    const uint32_t n = s_n;
    if (threadIdx.x == 0) *(bridges ? n_bridges : n_contacts) = n;
    uint32_t P = 1;
    while (P < n) P <<= 1;
    const bool in_smem = P <= SORT_SMEM_KEYS;
    unsigned long long *a = in_smem ? sk : keys;
    for (uint32_t i = n + threadIdx.x; i < P; i += blockDim.x) a[i] = MDVC_KEY_EMPTY;       // P <= cap (a power of two)
    __syncthreads();
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long u = a[i], w = a[ixj];
                    if ((u > w) == ((i & k) == 0)) { a[i] = w; a[ixj] = u; }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long key = a[i];
        if (in_smem) keys[i] = key;
        int l1, l2, code;
        mdvc_unpack_key(key, l1, l2, code);
        if (bridges) {
            int32_t *o = rows + (size_t)i * 5;
            o[0] = l1; o[1] = l2; o[2] = code / 9 - 1; o[3] = (code / 3) % 3 - 1; o[4] = code % 3 - 1;
        } else {
            rows[2 * (size_t)i] = l1; rows[2 * (size_t)i + 1] = l2;
        }
    }
}

// ------------------------------------------------------------------------------------------
// stage 3: periodic components on the device
// ------------------------------------------------------------------------------------------
__global__ void k_label_uf_init(const Hdr *__restrict__ h, uint32_t *lparent) {
    uint32_t n = h->n_labels, stride = gridDim.x * blockDim.x;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) lparent[j] = j;
}

__global__ void k_label_uf_link(const Hdr *__restrict__ h, const unsigned long long *__restrict__ bkeys, uint32_t cap,
                                uint32_t *lparent) {
    if (h->n_labels == 0) return;
    uint32_t n = min(h->n_bridges, cap), nf = h->n_false, stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int l1, l2, code;
        mdvc_unpack_key(bkeys[i], l1, l2, code);
        if (l1 != l2 && ((l1 < 0) == (l2 < 0))) mdvc_uf_union(lparent, (uint32_t)(l1 + (int)nf), (uint32_t)(l2 + (int)nf));
    }
}

__global__ void k_label_flags(const Hdr *__restrict__ h, uint32_t *lparent, unsigned long long *__restrict__ flags) {
    uint32_t n = h->n_labels, nf = h->n_false, stride = gridDim.x * blockDim.x;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        uint32_t root = j, up;                       // read-only climb, see k_flatten_flags
        while ((up = __ldcg(&lparent[root])) != root) root = up;
        unsigned long long f = 0;
        if (root == j) f = j < nf ? (1ull << 32) : (j > nf ? 1ull : 0ull);
        else lparent[j] = root;
        flags[j] = f;
    }
}

__global__ void k_after_comp_scan(Hdr *h) {
    h->n_comp_pos = (uint32_t)(h->tot_comps & 0xffffffffu);
    h->n_comp_neg = (uint32_t)(h->tot_comps >> 32);
}

// lut[j] = component of label (j - n_false): negative components ordered by their most negative
// label, positive ones by their smallest label
__global__ void k_label_lut(const Hdr *__restrict__ h, const uint32_t *__restrict__ lparent,
                            const unsigned long long *__restrict__ rank, int32_t *__restrict__ lut,
                            const unsigned long long *__restrict__ vol, unsigned long long *ccount, int periodic) {
    uint32_t n = h->n_labels, nf = h->n_false, stride = gridDim.x * blockDim.x;
    uint32_t ncn = periodic ? h->n_comp_neg : nf;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        int comp;
        if (!periodic) comp = (int)j - (int)nf;
        else {
            unsigned long long e = rank[lparent[j]];
            comp = j < nf ? -(int)((uint32_t)(e >> 32) + 1u) : (j > nf ? (int)((uint32_t)(e & 0xffffffffu) + 1u) : 0);
        }
        lut[j] = comp;
        if (j != nf) atomicAdd(&ccount[(long long)comp + ncn], vol[j]);
    }
}

__global__ void k_run_components(const Hdr *__restrict__ h, const int32_t *__restrict__ lab, const int32_t *__restrict__ lut,
                                 uint32_t lut_n, int32_t *__restrict__ runcomp) {
    if (h->n_labels == 0) return;
    uint32_t R = h->run_eff, nf = h->n_false, stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += stride) {
        uint32_t j = (uint32_t)(lab[i] + (int)nf);
        runcomp[i] = j < lut_n ? lut[j] : 0;
    }
}

// ------------------------------------------------------------------------------------------
// stage 4: expand -- the one pass over the dense int32 grids
// Each row group recomputes the start mask of its row from the bits (L2-resident), turns it into
// run ordinals by popcount, looks the per-run value up (L1-resident: consecutive voxels share runs)
// and streams the row out with 16-byte stores.
// ------------------------------------------------------------------------------------------
// 256-bit streaming store (sm_100: STG.E.EF.256)
__device__ __forceinline__ void st_v8_cs(int32_t *p, int a, int b, int c, int d, int e, int f, int g, int h) {
#ifndef EXPAND_ST
#define EXPAND_ST "st.global.cs.v8.b32"
#endif
    asm volatile(EXPAND_ST " [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f),
                 "r"(g), "r"(h)
                 : "memory");
}

#ifndef EXPAND_LAYOUT
#define EXPAND_LAYOUT 0             // 1 = every lane expands its own word: 1.89 ms instead of 1.31 (warp stores must stay contiguous)
#endif
#ifdef EXPAND_MIN_CTAS
#define EXPAND_BOUNDS __launch_bounds__(256, EXPAND_MIN_CTAS)
#else
#define EXPAND_BOUNDS __launch_bounds__(256)
#endif
template <int G, int VEC>
__global__ void EXPAND_BOUNDS
k_expand(const uint32_t *__restrict__ bits, Dims D, const Hdr *__restrict__ h, const uint32_t *__restrict__ rowoff,
         const int32_t *__restrict__ lab, const int32_t *__restrict__ runcomp,
         int32_t *__restrict__ out_lab, int32_t *__restrict__ out_comp, long long row_begin, long long row_end) {
    if (h->n_labels == 0) return;
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), gi = lane / G;
    constexpr int GPW = 32 / G;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    int nchunks = (D.nw + G - 1) / G;
    for (long long rb = row_begin + warp * GPW; rb < row_end; rb += nwarps * GPW) {
        long long row = rb + gi;
        bool rv = row < row_end;
        const uint32_t *rp = bits + (rv ? row : 0) * D.pitch;
        uint32_t base = rv ? rowoff[row] : 0u;
        const int32_t *lrow = lab + base;
        const int32_t *crow = runcomp + base;
        size_t obase = (size_t)(rv ? row : 0) * D.Z;
        RowWalker<G> rw; rw.begin();
        for (int c = 0; c < nchunks; ++c) {
            int w = c * G + gl;
            uint32_t B = (rv && w < D.nw) ? (rp[w] & mdvc_valid_mask(w, D.Z)) : 0u;
            uint32_t S = rw.starts(B, w, D.Z, gl);
            uint32_t ex = rw.exclusive(S, gl);
            if (VEC == 8) {
                // eight voxels (a quarter word) per lane and iteration, written with one 256-bit streaming store
                // (STG.E.EF.256: a warp store covers 1 KiB of full sectors).  Most quarter words contain no run
                // start, so all eight voxels take one looked-up value.  Needs Z % 8 == 0 and 32-byte aligned outputs.
#pragma unroll
                for (int it = 0; it < 4; ++it) {
#if EXPAND_LAYOUT == 1
                    // variant: every lane expands its own word (no shuffles; a warp store covers 32 separate sectors)
                    const uint32_t Sw = S, Ew = ex;
                    const int k0 = it << 3;
                    const int z0 = (w << 5) + k0;
#else
                    const int q = it * G + gl;                 // quarter-word index inside the chunk
                    const int src = q >> 2;
                    const uint32_t Sw = __shfl_sync(MDVC_FULL, S, src, G);
                    const uint32_t Ew = __shfl_sync(MDVC_FULL, ex, src, G);
                    const int k0 = (q & 3) << 3;
                    const int z0 = ((c * G + src) << 5) + k0;
#endif
                    if (rv && z0 < D.Z) {
                        const uint32_t o0 = Ew + __popc(Sw & ((2u << k0) - 1u)) - 1u;
                        const uint32_t byte = (Sw >> k0) & 0xffu;
                        const size_t at = obase + z0;
                        if ((byte >> 1) == 0) {
                            if (out_lab) {
                                const int v = __ldg(lrow + o0);
                                st_v8_cs(out_lab + at, v, v, v, v, v, v, v, v);
                            }
                            if (out_comp) {
                                const int v = __ldg(crow + o0);
                                st_v8_cs(out_comp + at, v, v, v, v, v, v, v, v);
                            }
                        } else {
                            uint32_t o[8];
                            o[0] = o0;
#pragma unroll
                            for (int j = 1; j < 8; ++j) o[j] = o[j - 1] + ((byte >> j) & 1u);
                            if (out_lab) {
                                st_v8_cs(out_lab + at, __ldg(lrow + o[0]), __ldg(lrow + o[1]), __ldg(lrow + o[2]), __ldg(lrow + o[3]),
                                         __ldg(lrow + o[4]), __ldg(lrow + o[5]), __ldg(lrow + o[6]), __ldg(lrow + o[7]));
                            }
                            if (out_comp) {
                                st_v8_cs(out_comp + at, __ldg(crow + o[0]), __ldg(crow + o[1]), __ldg(crow + o[2]), __ldg(crow + o[3]),
                                         __ldg(crow + o[4]), __ldg(crow + o[5]), __ldg(crow + o[6]), __ldg(crow + o[7]));
                            }
                        }
                    }
                }
            } else if (VEC == 4) {
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    int q = it * G + gl;                       // quad (4 voxels) inside the chunk
                    int src = q >> 3;
                    uint32_t Sw = __shfl_sync(MDVC_FULL, S, src, G);
                    uint32_t Ew = __shfl_sync(MDVC_FULL, ex, src, G);
                    int k0 = (q & 7) << 2;
                    int z0 = ((c * G + src) << 5) + k0;
                    if (rv && z0 < D.Z) {
                        uint32_t o0 = Ew + __popc(Sw & ((2u << k0) - 1u)) - 1u;
                        uint32_t nib = (Sw >> k0) & 0xfu;
                        uint32_t o1 = o0 + ((nib >> 1) & 1u), o2 = o1 + ((nib >> 2) & 1u), o3 = o2 + ((nib >> 3) & 1u);
                        if (out_lab) {
                            int4 val;
                            val.x = __ldg(lrow + o0);
                            if (nib >> 1) { val.y = __ldg(lrow + o1); val.z = __ldg(lrow + o2); val.w = __ldg(lrow + o3); }
                            else { val.y = val.x; val.z = val.x; val.w = val.x; }
                            __stcs(reinterpret_cast<int4 *>(out_lab + obase + z0), val);
                        }
                        if (out_comp) {
                            int4 val;
                            val.x = __ldg(crow + o0);
                            if (nib >> 1) { val.y = __ldg(crow + o1); val.z = __ldg(crow + o2); val.w = __ldg(crow + o3); }
                            else { val.y = val.x; val.z = val.x; val.w = val.x; }
                            __stcs(reinterpret_cast<int4 *>(out_comp + obase + z0), val);
                        }
                    }
                }
            } else {
#pragma unroll 4
                for (int it = 0; it < 32; ++it) {
                    int vix = it * G + gl;
                    int src = vix >> 5, k = vix & 31;
                    uint32_t Sw = __shfl_sync(MDVC_FULL, S, src, G);
                    uint32_t Ew = __shfl_sync(MDVC_FULL, ex, src, G);
                    int z = ((c * G + src) << 5) + k;
                    if (rv && z < D.Z) {
                        uint32_t o = Ew + __popc(Sw & ((2u << k) - 1u)) - 1u;
                        if (out_lab) __stcs(out_lab + obase + z, __ldg(lrow + o));
                        if (out_comp) __stcs(out_comp + obase + z, __ldg(crow + o));
                    }
                }
            }
        }
    }
}

// The dense write for rows that fit one warp (17..32 words, Z % 8 == 0, 32-byte aligned outputs): one warp per row as
// in k_expand<32, 8>, but the row's dependent chain -- row offset -> run values -> stores -- is taken off the critical
// path: the NEXT row's offset and bit word are requested before the current row is expanded, its first 32 run values
// (label, component) as soon as the offset has arrived, and the value of an octet is a warp shuffle from those
// registers instead of a load that has to wait for the ordinal.  Octets that contain a run start (a few per row) and
// rows with more than 32 runs take the look-ups of the generic kernel.  A plain fill of the same 8 GiB runs at
// 7.5 TB/s on the B200, the generic kernel at 6.6: what separates them is latency per row, not bytes.
#ifndef EXPAND_PREFETCH
#define EXPAND_PREFETCH 1
#endif
template <bool HAS_LAB, bool HAS_COMP>
__global__ void __launch_bounds__(256)
k_expand_row32(const uint32_t *__restrict__ bits, Dims D, const Hdr *__restrict__ h, const uint32_t *__restrict__ rowoff,
               const int32_t *__restrict__ lab, const int32_t *__restrict__ runcomp,
               int32_t *__restrict__ out_lab, int32_t *__restrict__ out_comp, long long row_begin, long long row_end) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    long long row = row_begin + warp;
    if (row >= row_end) return;
    const uint32_t vmask = mdvc_valid_mask(lane, D.Z);
    // a CTA lives for three or four rows per warp: the loads that do not depend on each other go out together (the
    // "labels overflowed" flag is only looked at afterwards instead of stalling every warp for one round trip first)
    const uint32_t n_labels = h->n_labels;
    uint32_t base = rowoff[row], nrun = rowoff[row + 1] - base;
    uint32_t B = lane < D.nw ? (bits[(size_t)row * D.pitch + lane] & vmask) : 0u;
    if (n_labels == 0) return;
    int lv = (HAS_LAB && (uint32_t)lane < nrun) ? __ldg(lab + base + lane) : 0;
    int cv = (HAS_COMP && (uint32_t)lane < nrun) ? __ldg(runcomp + base + lane) : 0;
    for (;;) {
        const long long next = row + nwarps;
        const bool more = next < row_end;
        uint32_t base_n = 0, end_n = 0, B_n = 0;
        if (more) {                                               // requested now, needed one row later
            base_n = rowoff[next];
            end_n = rowoff[next + 1];
            B_n = lane < D.nw ? (bits[(size_t)next * D.pitch + lane] & vmask) : 0u;
        }
        // start mask and ordinal of the first run of every word (single chunk: no carries)
        uint32_t prev = __shfl_up_sync(MDVC_FULL, B, 1) >> 31;
        if (lane == 0) prev = 0;
        uint32_t S = B ^ ((B << 1) | prev);
        if (lane == 0) S |= 1u;
        S &= vmask;
        const uint32_t cnt = __popc(S);
        uint32_t ex;
        if (__ballot_sync(MDVC_FULL, cnt > 1) == 0) {
            const uint32_t has = __ballot_sync(MDVC_FULL, cnt != 0);
            ex = __popc(has & ((1u << lane) - 1u));
        } else {
            uint32_t inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(MDVC_FULL, inc, o);
                if (lane >= o) inc += t;
            }
            ex = inc - cnt;
        }
        int lv_n = 0, cv_n = 0;
        if (more) {
            const uint32_t nrun_n = end_n - base_n;
            if (HAS_LAB && (uint32_t)lane < nrun_n) lv_n = __ldg(lab + base_n + lane);
            if (HAS_COMP && (uint32_t)lane < nrun_n) cv_n = __ldg(runcomp + base_n + lane);
        }
        const bool in_regs = nrun <= 32u;                         // warp-uniform
        const int32_t *lrow = lab + base, *crow = runcomp + base;
        const size_t obase = (size_t)row * D.Z;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int q = it * 32 + lane;                         // quarter-word index inside the row
            const int src = q >> 2;
            const uint32_t Sw = __shfl_sync(MDVC_FULL, S, src);
            const uint32_t Ew = __shfl_sync(MDVC_FULL, ex, src);
            const int k0 = (q & 3) << 3;
            const int z0 = (src << 5) + k0;
            const uint32_t o0 = Ew + __popc(Sw & ((2u << k0) - 1u)) - 1u;
            const uint32_t byte = (Sw >> k0) & 0xffu;
            int vl = 0, vc = 0;
            if (in_regs) {                                        // every lane takes part in the shuffles
                if (HAS_LAB) vl = __shfl_sync(MDVC_FULL, lv, o0 & 31u);
                if (HAS_COMP) vc = __shfl_sync(MDVC_FULL, cv, o0 & 31u);
            }
            if (z0 < D.Z) {
                const size_t at = obase + z0;
                if ((byte >> 1) == 0) {
                    if (!in_regs) {
                        if (HAS_LAB) vl = __ldg(lrow + o0);
                        if (HAS_COMP) vc = __ldg(crow + o0);
                    }
                    if (HAS_LAB) st_v8_cs(out_lab + at, vl, vl, vl, vl, vl, vl, vl, vl);
                    if (HAS_COMP) st_v8_cs(out_comp + at, vc, vc, vc, vc, vc, vc, vc, vc);
                } else {
                    uint32_t o[8];
                    o[0] = o0;
#pragma unroll
                    for (int j = 1; j < 8; ++j) o[j] = o[j - 1] + ((byte >> j) & 1u);
                    if (HAS_LAB)
                        st_v8_cs(out_lab + at, __ldg(lrow + o[0]), __ldg(lrow + o[1]), __ldg(lrow + o[2]), __ldg(lrow + o[3]),
                                 __ldg(lrow + o[4]), __ldg(lrow + o[5]), __ldg(lrow + o[6]), __ldg(lrow + o[7]));
                    if (HAS_COMP)
                        st_v8_cs(out_comp + at, __ldg(crow + o[0]), __ldg(crow + o[1]), __ldg(crow + o[2]), __ldg(crow + o[3]),
                                 __ldg(crow + o[4]), __ldg(crow + o[5]), __ldg(crow + o[6]), __ldg(crow + o[7]));
                }
            }
        }
        if (!more) break;
        row = next; base = base_n; nrun = end_n - base_n; B = B_n; lv = lv_n; cv = cv_n;
    }
}

// ------------------------------------------------------------------------------------------
// host entry points
// ------------------------------------------------------------------------------------------
static int check_ws(void *ws, size_t ws_bytes, const mdvc_frame_caps *caps, long long rows, Layout &L) {
    if (!ws || !caps_ok(caps)) return MDVC_ERR_BADARG;
    L = make_layout(rows, caps);
    if (ws_bytes < L.total) return MDVC_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return MDVC_ERR_BADARG;
    return MDVC_OK;
}

static int frame_label(const uint32_t *bits, int X, int Y, int Z, int pitch, void *ws, size_t ws_bytes,
                       const mdvc_frame_caps *caps, int rows_counted, void *stream);

extern "C" int mdvc_frame_label(const uint32_t *bits, int X, int Y, int Z, int pitch, void *ws, size_t ws_bytes,
                                const mdvc_frame_caps *caps, void *stream) {
    return frame_label(bits, X, Y, Z, pitch, ws, ws_bytes, caps, 0, stream);
}

extern "C" int mdvc_frame_label_counted(const uint32_t *bits, int X, int Y, int Z, int pitch, void *ws, size_t ws_bytes,
                                        const mdvc_frame_caps *caps, int rows_counted, void *stream) {
    return frame_label(bits, X, Y, Z, pitch, ws, ws_bytes, caps, rows_counted, stream);
}

extern "C" uint32_t *mdvc_frame_row_counts_ptr(void *ws, int X, int Y, int Z, const mdvc_frame_caps *caps) {
    if (!ws || !caps_ok(caps) || X <= 0 || Y <= 0 || Z <= 0) return nullptr;
    Layout L = make_layout((long long)X * Y, caps);
    return at<uint32_t>(ws, L.rowoff);
}

static int frame_label(const uint32_t *bits, int X, int Y, int Z, int pitch, void *ws, size_t ws_bytes,
                       const mdvc_frame_caps *caps, int rows_counted, void *stream) {
    Dims D; Layout L;
    if (!bits || !make_dims(X, Y, Z, pitch, D)) return MDVC_ERR_BADARG;
    int rc = check_ws(ws, ws_bytes, caps, D.rows, L);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    Hdr *h = at<Hdr>(ws, L.hdr);
    uint32_t *rowoff = at<uint32_t>(ws, L.rowoff);
    uint8_t *rowpol = at<uint8_t>(ws, L.rowpol), *rowflag = at<uint8_t>(ws, L.rowflag);
    uint32_t *runstart = at<uint32_t>(ws, L.runstart);
    uint32_t *parent = at<uint32_t>(ws, L.parent);
    int32_t *lab = at<int32_t>(ws, L.lab);
    unsigned long long *scan64 = at<unsigned long long>(ws, L.scan64);
    unsigned long long *vol = at<unsigned long long>(ws, L.vol);

    MDVC_CUDA_CHECK(cudaMemsetAsync(h, 0, sizeof(Hdr), st));
    int g = pick_group(D.nw);
    long long row_threads = D.rows * g;
    unsigned row_grid = mdvc_grid(row_threads / 4, 256, 8);             // four rows per group and iteration
    // rows of at most 128 words on 16-byte aligned grids: the quad-per-lane kernels
    int lq = 1;
    while (lq < D.pitch / 4) lq <<= 1;
    const bool quad = lq <= 32 && D.pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(bits) & 15) == 0;
    if (quad) row_grid = mdvc_grid(D.rows * lq / 4, 256, 8);
    if (rows_counted) {
        // stage 1a was done by the kernel that produced the grid (mdvc_morph_counted)
    } else if (quad) {
        MDVC_DISPATCH_LQ(lq, (k_count_runs_q<LQ><<<row_grid, 256, 0, st>>>(bits, D, rowoff)));
        MDVC_LAUNCH_CHECK();
    } else {
        MDVC_DISPATCH_G(g, (k_count_runs<G><<<row_grid, 256, 0, st>>>(bits, D, rowoff)));
        MDVC_LAUNCH_CHECK();
    }
    rc = mdvc_exclusive_scan<uint32_t>(rowoff, rowoff, nullptr, (uint32_t)D.rows, at<uint32_t>(ws, L.tile32),
                                       rowoff + D.rows, nullptr, st);
    if (rc) return rc;
    k_after_row_scan<<<1, 1, 0, st>>>(h, rowoff, D.rows, caps->max_runs, runstart, D.N);
    MDVC_LAUNCH_CHECK();
    if (quad) {
        MDVC_DISPATCH_LQ(lq, (k_fill_runs_q<LQ><<<row_grid, 256, 0, st>>>(bits, D, h, rowoff, runstart, parent, rowpol)));
    } else {
        MDVC_DISPATCH_G(g, (k_fill_runs<G><<<row_grid, 256, 0, st>>>(bits, D, h, rowoff, runstart, parent, rowpol)));
    }
    MDVC_LAUNCH_CHECK();
    unsigned run_grid = mdvc_grid(caps->max_runs, 256, 8);
    // linking: strips in shared memory -> strip borders -> plane borders in radix-LINK_RADIX rounds;
    // the strip-root list is flattened after every merging round
    uint32_t *rootlist = reinterpret_cast<uint32_t *>(scan64);
    {
        int spp = (D.Y + STRIP_TY - 1) / STRIP_TY;
        long long n_strips = (long long)D.X * spp;
        unsigned sgrid = (unsigned)(n_strips < (long long)mdvc_num_sms() * 32 ? n_strips : (long long)mdvc_num_sms() * 32);
        unsigned lgrid = mdvc_grid(caps->max_runs, 256, 4);
        k_strip_link<<<sgrid, TILE_THREADS, 0, st>>>(bits, D, h, rowoff, rowpol, runstart, parent, rootlist, rowflag);
        MDVC_LAUNCH_CHECK();
        if (D.Y > STRIP_TY) {
            long long items = (long long)D.X * ((D.Y - 1) / STRIP_TY);
            k_link_round<<<mdvc_grid(items * 8, 256, 8), 256, 0, st>>>(bits, D, h, rowoff, runstart, parent, 0, STRIP_TY, 1 << 30, 8, rowflag);
            MDVC_LAUNCH_CHECK();
            k_flatten_list<<<lgrid, 256, 0, st>>>(h, rootlist, parent);
            MDVC_LAUNCH_CHECK();
        }
        for (long long s = 1; s < D.X; s *= LINK_RADIX) {
            long long tiles = (long long)((D.X - 1) / s) * spp;
            unsigned tgrid = (unsigned)(tiles < (long long)mdvc_num_sms() * 32 ? tiles : (long long)mdvc_num_sms() * 32);
            k_tile_link_x<<<tgrid, TILE_THREADS, 0, st>>>(bits, D, h, rowoff, rowpol, runstart, parent, (int)s, LINK_RADIX, rowflag);
            MDVC_LAUNCH_CHECK();
            k_flatten_list<<<lgrid, 256, 0, st>>>(h, rootlist, parent);
            MDVC_LAUNCH_CHECK();
        }
    }
    uint32_t *localrank = reinterpret_cast<uint32_t *>(scan64);
    unsigned long long *root_tiles = at<unsigned long long>(ws, L.tile64);
    k_flatten_flags<<<(caps->max_runs + ROOT_TILE - 1) / ROOT_TILE, 256, 0, st>>>(bits, D, h, runstart, parent, localrank, root_tiles);
    MDVC_LAUNCH_CHECK();
    k_root_tile_offsets<<<1, 1024, 0, st>>>(h, root_tiles);
    MDVC_LAUNCH_CHECK();
    k_after_root_scan<<<mdvc_num_sms(), 256, 0, st>>>(h, caps->max_labels, vol);
    MDVC_LAUNCH_CHECK();
    k_assign_labels<<<run_grid, 256, 0, st>>>(bits, D, h, runstart, parent, localrank, root_tiles, lab, vol);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

extern "C" int mdvc_frame_pairs(const uint32_t *bits, int X, int Y, int Z, int pitch, void *ws, size_t ws_bytes,
                                const mdvc_frame_caps *caps, int periodic, void *stream) {
    Dims D; Layout L;
    if (!bits || !make_dims(X, Y, Z, pitch, D)) return MDVC_ERR_BADARG;
    int rc = check_ws(ws, ws_bytes, caps, D.rows, L);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    Hdr *h = at<Hdr>(ws, L.hdr);
    unsigned long long *cset = at<unsigned long long>(ws, L.cset), *bset = at<unsigned long long>(ws, L.bset);
    unsigned long long *ckeys = at<unsigned long long>(ws, L.ckeys), *bkeys = at<unsigned long long>(ws, L.bkeys);
    MDVC_CUDA_CHECK(cudaMemsetAsync(cset, 0xff, (size_t)caps->contact_slots * 8, st));
    MDVC_CUDA_CHECK(cudaMemsetAsync(bset, 0xff, (size_t)caps->bridge_slots * 8, st));
    MDVC_CUDA_CHECK(cudaMemsetAsync(&h->n_contacts, 0, 2 * sizeof(uint32_t), st));
    MDVC_CUDA_CHECK(cudaMemsetAsync(&h->n_pair_rows, 0, sizeof(uint32_t), st));
    const int32_t *lab = at<int32_t>(ws, L.lab);
    const uint32_t *rowoff = at<uint32_t>(ws, L.rowoff);
    const uint8_t *rowflag = at<uint8_t>(ws, L.rowflag);
    uint32_t *pairrows = at<uint32_t>(ws, L.pairrows);
    k_row_pairs<<<mdvc_grid(D.rows + D.rows / RP_ROWS * 3 + 32, 256, RP_MIN_CTAS), 256, 0, st>>>(D, h, rowoff, rowflag, lab, pairrows, cset,
                                                                           caps->contact_slots, bset, caps->bridge_slots, periodic);
    MDVC_LAUNCH_CHECK();
    // the list is short on smooth grids (CTAs past its end leave at once) and holds every row on noise
    k_task_pairs<<<mdvc_grid(D.rows * PR_LANES, 256, 8), 256, 0, st>>>(D, h, rowoff, at<uint8_t>(ws, L.rowpol), rowflag, pairrows,
                                                                       at<uint32_t>(ws, L.runstart), lab, cset, caps->contact_slots,
                                                                       bset, caps->bridge_slots, periodic);
    MDVC_LAUNCH_CHECK();
    if (caps->contact_slots <= 4 * BSORT_TILE && caps->bridge_slots <= 4 * BSORT_TILE) {
        k_finish_sets<<<periodic ? 2 : 1, 1024, 0, st>>>(cset, ckeys, at<int32_t>(ws, L.crows), &h->n_contacts, caps->contact_slots,
                                                         bset, bkeys, at<int32_t>(ws, L.brows), &h->n_bridges, caps->bridge_slots);
        MDVC_LAUNCH_CHECK();
        return MDVC_OK;
    }
    k_compact_set<<<mdvc_grid(caps->contact_slots, 256, 4), 256, 0, st>>>(cset, caps->contact_slots, ckeys, &h->n_contacts);
    MDVC_LAUNCH_CHECK();
    rc = sort_keys(ckeys, &h->n_contacts, caps->contact_slots, st);
    if (rc) return rc;
    k_decode_rows<<<mdvc_grid(caps->contact_slots, 256, 4), 256, 0, st>>>(ckeys, &h->n_contacts, caps->contact_slots, 0,
                                                                          at<int32_t>(ws, L.crows));
    MDVC_LAUNCH_CHECK();
    if (periodic) {
        k_compact_set<<<mdvc_grid(caps->bridge_slots, 256, 4), 256, 0, st>>>(bset, caps->bridge_slots, bkeys, &h->n_bridges);
        MDVC_LAUNCH_CHECK();
        rc = sort_keys(bkeys, &h->n_bridges, caps->bridge_slots, st);
        if (rc) return rc;
        k_decode_rows<<<mdvc_grid(caps->bridge_slots, 256, 4), 256, 0, st>>>(bkeys, &h->n_bridges, caps->bridge_slots, 1,
                                                                             at<int32_t>(ws, L.brows));
        MDVC_LAUNCH_CHECK();
    }
    return MDVC_OK;
}

extern "C" int mdvc_frame_components(int X, int Y, int Z, void *ws, size_t ws_bytes, const mdvc_frame_caps *caps,
                                          int periodic, void *stream) {
    Dims D; Layout L;
    if (!make_dims(X, Y, Z, mdvc_words(Z), D)) return MDVC_ERR_BADARG;
    int rc = check_ws(ws, ws_bytes, caps, D.rows, L);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    Hdr *h = at<Hdr>(ws, L.hdr);
    uint32_t *lparent = at<uint32_t>(ws, L.lparent);
    int32_t *lut = at<int32_t>(ws, L.lut);
    unsigned long long *lscan = at<unsigned long long>(ws, L.lscan);
    unsigned long long *ccount = at<unsigned long long>(ws, L.ccount);
    MDVC_CUDA_CHECK(cudaMemsetAsync(ccount, 0, (size_t)caps->max_labels * 8, st));
    unsigned lgrid = mdvc_grid(caps->max_labels, 256, 4);
    if (periodic) {
        k_label_uf_init<<<lgrid, 256, 0, st>>>(h, lparent);
        MDVC_LAUNCH_CHECK();
        k_label_uf_link<<<mdvc_grid(caps->bridge_slots, 256, 4), 256, 0, st>>>(h, at<unsigned long long>(ws, L.bkeys),
                                                                                caps->bridge_slots, lparent);
        MDVC_LAUNCH_CHECK();
        k_label_flags<<<lgrid, 256, 0, st>>>(h, lparent, lscan);
        MDVC_LAUNCH_CHECK();
        rc = mdvc_exclusive_scan<unsigned long long>(lscan, lscan, &h->n_labels, caps->max_labels,
                                                     at<unsigned long long>(ws, L.tile64), &h->tot_comps, nullptr, st);
        if (rc) return rc;
        k_after_comp_scan<<<1, 1, 0, st>>>(h);
        MDVC_LAUNCH_CHECK();
    } else {
        MDVC_CUDA_CHECK(cudaMemcpyAsync(&h->n_comp_pos, &h->n_true, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        MDVC_CUDA_CHECK(cudaMemcpyAsync(&h->n_comp_neg, &h->n_false, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
    k_label_lut<<<lgrid, 256, 0, st>>>(h, lparent, lscan, lut, at<unsigned long long>(ws, L.vol), ccount, periodic);
    MDVC_LAUNCH_CHECK();
    k_run_components<<<mdvc_grid(caps->max_runs, 256, 8), 256, 0, st>>>(h, at<int32_t>(ws, L.lab), lut, caps->max_labels,
                                                                        at<int32_t>(ws, L.runcomp));
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

extern "C" int mdvc_frame_set_lut(int X, int Y, int Z, void *ws, size_t ws_bytes, const mdvc_frame_caps *caps,
                                       const int32_t *lut_dev, uint32_t n_entries, void *stream) {
    Dims D; Layout L;
    if (!lut_dev || !make_dims(X, Y, Z, mdvc_words(Z), D)) return MDVC_ERR_BADARG;
    int rc = check_ws(ws, ws_bytes, caps, D.rows, L);
    if (rc) return rc;
    if (n_entries > caps->max_labels) return MDVC_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    MDVC_CUDA_CHECK(cudaMemcpyAsync(at<int32_t>(ws, L.lut), lut_dev, (size_t)n_entries * 4, cudaMemcpyDeviceToDevice, st));
    k_run_components<<<mdvc_grid(caps->max_runs, 256, 8), 256, 0, st>>>(at<Hdr>(ws, L.hdr), at<int32_t>(ws, L.lab),
                                                                        at<int32_t>(ws, L.lut), n_entries,
                                                                        at<int32_t>(ws, L.runcomp));
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

static int expand_rows(const uint32_t *bits, int X, int Y, int Z, int pitch, void *ws, size_t ws_bytes,
                       const mdvc_frame_caps *caps, int32_t *labels_out, int32_t *components_out,
                       long long row_begin, long long row_end, void *stream) {
    Dims D; Layout L;
    if (!bits || !make_dims(X, Y, Z, pitch, D)) return MDVC_ERR_BADARG;
    int rc = check_ws(ws, ws_bytes, caps, D.rows, L);
    if (rc) return rc;
    if (row_begin < 0 || row_end > D.rows || row_begin > row_end) return MDVC_ERR_BADARG;
    if ((!labels_out && !components_out) || row_begin == row_end) return MDVC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int g = pick_group(D.nw);
    uintptr_t both = reinterpret_cast<uintptr_t>(labels_out) | reinterpret_cast<uintptr_t>(components_out);
    int vec = (Z % 8 == 0 && (both & 31) == 0) ? 8 : ((Z % 4 == 0 && (both & 15) == 0) ? 4 : 0);
    // Not a persistent grid: one short-lived CTA per 8 row groups keeps the concurrently written rows close
    // together in memory (measured on B200 at 1024^3, labels + components: 1.45 ms with 8 CTAs/SM grid-stride,
    // 1.27 ms = 6.77 TB/s with up to 256 CTAs/SM).
    unsigned grid = mdvc_grid((row_end - row_begin) * g, 256, 256);
    const Hdr *h = at<Hdr>(ws, L.hdr);
    const uint32_t *rowoff = at<uint32_t>(ws, L.rowoff);
    const int32_t *lab = at<int32_t>(ws, L.lab), *rc_ = at<int32_t>(ws, L.runcomp);
    // the kernel indexes its outputs by absolute row: shift the bases so that row_begin lands on element 0
    int32_t *lo = labels_out ? labels_out - row_begin * Z : nullptr;
    int32_t *co = components_out ? components_out - row_begin * Z : nullptr;
    if (EXPAND_PREFETCH && vec == 8 && g == 32 && D.nw <= 32) {
        if (lo && co) k_expand_row32<true, true><<<grid, 256, 0, st>>>(bits, D, h, rowoff, lab, rc_, lo, co, row_begin, row_end);
        else if (lo) k_expand_row32<true, false><<<grid, 256, 0, st>>>(bits, D, h, rowoff, lab, rc_, lo, co, row_begin, row_end);
        else k_expand_row32<false, true><<<grid, 256, 0, st>>>(bits, D, h, rowoff, lab, rc_, lo, co, row_begin, row_end);
    }
    else if (vec == 8) { MDVC_DISPATCH_G(g, (k_expand<G, 8><<<grid, 256, 0, st>>>(bits, D, h, rowoff, lab, rc_, lo, co, row_begin, row_end))); }
    else if (vec == 4) { MDVC_DISPATCH_G(g, (k_expand<G, 4><<<grid, 256, 0, st>>>(bits, D, h, rowoff, lab, rc_, lo, co, row_begin, row_end))); }
    else { MDVC_DISPATCH_G(g, (k_expand<G, 0><<<grid, 256, 0, st>>>(bits, D, h, rowoff, lab, rc_, lo, co, row_begin, row_end))); }
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

extern "C" int mdvc_frame_expand(const uint32_t *bits, int X, int Y, int Z, int pitch, void *ws, size_t ws_bytes,
                                 const mdvc_frame_caps *caps, int32_t *labels_out, int32_t *components_out, void *stream) {
    return expand_rows(bits, X, Y, Z, pitch, ws, ws_bytes, caps, labels_out, components_out, 0, (long long)X * Y, stream);
}

// planes [x_begin, x_end) only; the outputs hold (x_end - x_begin) * Y * Z elements
extern "C" int mdvc_frame_expand_planes(const uint32_t *bits, int X, int Y, int Z, int pitch, void *ws, size_t ws_bytes,
                                        const mdvc_frame_caps *caps, int x_begin, int x_end, int32_t *labels_out,
                                        int32_t *components_out, void *stream) {
    if (x_begin < 0 || x_end > X || x_begin > x_end) return MDVC_ERR_BADARG;
    return expand_rows(bits, X, Y, Z, pitch, ws, ws_bytes, caps, labels_out, components_out, (long long)x_begin * Y,
                       (long long)x_end * Y, stream);
}

// lab[i] = lut[lab[i] + n_false] for every run (slab decomposition: local label -> global label).  Call after
// mdvc_frame_set_lut, which indexes its LUT by the *local* labels.
__global__ void k_relabel_runs(const Hdr *__restrict__ h, int32_t *lab, const int32_t *__restrict__ lut, uint32_t lut_n) {
    if (h->n_labels == 0) return;
    uint32_t R = h->run_eff, nf = h->n_false, stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += stride) {
        uint32_t j = (uint32_t)(lab[i] + (int)nf);
        lab[i] = j < lut_n ? lut[j] : 0;
    }
}

extern "C" int mdvc_frame_relabel_runs(int X, int Y, int Z, void *ws, size_t ws_bytes, const mdvc_frame_caps *caps,
                                       const int32_t *lut_dev, uint32_t n_entries, void *stream) {
    Dims D; Layout L;
    if (!lut_dev || !make_dims(X, Y, Z, mdvc_words(Z), D)) return MDVC_ERR_BADARG;
    int rc = check_ws(ws, ws_bytes, caps, D.rows, L);
    if (rc) return rc;
    k_relabel_runs<<<mdvc_grid(caps->max_runs, 256, 8), 256, 0, (cudaStream_t)stream>>>(at<Hdr>(ws, L.hdr), at<int32_t>(ws, L.lab),
                                                                                         lut_dev, n_entries);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

// ------------------------------------------------------------------------------------------
// N2: per-voxel values straight from the run tables (no dense grid)
// ------------------------------------------------------------------------------------------
// value (label or component) of the run that contains linear voxel index `lin` of row `row`
__device__ __forceinline__ int32_t run_value_at(const uint32_t *__restrict__ rowoff, const uint32_t *__restrict__ runstart,
                                                const int32_t *__restrict__ table, uint32_t row, uint32_t lin) {
    uint32_t lo = rowoff[row], hi = rowoff[row + 1];
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (runstart[mid] <= lin) lo = mid; else hi = mid;
    }
    return table[lo];
}

__global__ void __launch_bounds__(256)
k_atom_run_values(const int32_t *__restrict__ vox, long long n, Dims D, const Hdr *__restrict__ h,
                  const uint32_t *__restrict__ rowoff, const uint32_t *__restrict__ runstart,
                  const int32_t *__restrict__ table, double *__restrict__ out) {
    const bool ok = h->run_eff != 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int x = vox[3 * i], y = vox[3 * i + 1], z = vox[3 * i + 2];
        int32_t v = 0;
        if (ok && (unsigned)x < (unsigned)D.X && (unsigned)y < (unsigned)D.Y && (unsigned)z < (unsigned)D.Z) {
            const uint32_t row = (uint32_t)x * (uint32_t)D.Y + (uint32_t)y;
            v = run_value_at(rowoff, runstart, table, row, row * (uint32_t)D.Z + (uint32_t)z);
        }
        out[i] = (double)v;
    }
}

__global__ void __launch_bounds__(256)
k_key_run_values(const unsigned long long *__restrict__ keys, long long n, Dims D, const Hdr *__restrict__ h,
                 const uint32_t *__restrict__ rowoff, const uint32_t *__restrict__ runstart,
                 const int32_t *__restrict__ table, int32_t *__restrict__ out) {
    const bool ok = h->run_eff != 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = keys[i];
        int32_t v = 0;
        if (ok && key < D.N) v = run_value_at(rowoff, runstart, table, (uint32_t)key / (uint32_t)D.Z, (uint32_t)key);
        out[i] = v;
    }
}

static int run_value_args(int X, int Y, int Z, void *ws, size_t ws_bytes, const mdvc_frame_caps *caps, int which, Dims &D,
                          Layout &L, const int32_t *&table) {
    if (which != 0 && which != 1) return MDVC_ERR_BADARG;
    if (!make_dims(X, Y, Z, mdvc_words(Z), D)) return MDVC_ERR_BADARG;
    int rc = check_ws(ws, ws_bytes, caps, D.rows, L);
    if (rc) return rc;
    table = at<int32_t>(ws, which ? L.runcomp : L.lab);
    return MDVC_OK;
}

extern "C" int mdvc_frame_atom_values(const int32_t *vox, long long n, int X, int Y, int Z, void *ws, size_t ws_bytes,
                                      const mdvc_frame_caps *caps, int which, double *out, void *stream) {
    if (n < 0 || (n > 0 && (!vox || !out))) return MDVC_ERR_BADARG;
    Dims D; Layout L; const int32_t *table;
    int rc = run_value_args(X, Y, Z, ws, ws_bytes, caps, which, D, L, table);
    if (rc || n == 0) return rc;
    k_atom_run_values<<<mdvc_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(vox, n, D, at<Hdr>(ws, L.hdr), at<uint32_t>(ws, L.rowoff),
                                                                              at<uint32_t>(ws, L.runstart), table, out);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

extern "C" int mdvc_frame_values_at(const unsigned long long *keys, long long n, int X, int Y, int Z, void *ws, size_t ws_bytes,
                                    const mdvc_frame_caps *caps, int which, int32_t *out, void *stream) {
    if (n < 0 || (n > 0 && (!keys || !out))) return MDVC_ERR_BADARG;
    Dims D; Layout L; const int32_t *table;
    int rc = run_value_args(X, Y, Z, ws, ws_bytes, caps, which, D, L, table);
    if (rc || n == 0) return rc;
    k_key_run_values<<<mdvc_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(keys, n, D, at<Hdr>(ws, L.hdr), at<uint32_t>(ws, L.rowoff),
                                                                             at<uint32_t>(ws, L.runstart), table, out);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

extern "C" int mdvc_frame_read_header(const void *ws, mdvc_frame_header *header_host, void *stream) {
    if (!ws || !header_host) return MDVC_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    MDVC_CUDA_CHECK(cudaMemcpyAsync(header_host, ws, sizeof(mdvc_frame_header), cudaMemcpyDeviceToHost, st));
    MDVC_CUDA_CHECK(cudaStreamSynchronize(st));
    memset(header_host->reserved, 0, sizeof(header_host->reserved));
    return (int)header_host->flags;
}

#define PTR_GETTER(name, type, field)                                                                         \
    extern "C" const type *name(const void *ws, int X, int Y, int Z, const mdvc_frame_caps *caps) {           \
        if (!ws || !caps_ok(caps) || X <= 0 || Y <= 0 || Z <= 0) return nullptr;                              \
        Layout L = make_layout((long long)X * Y, caps);                                                       \
        return at<type>(ws, L.field);                                                                         \
    }
PTR_GETTER(mdvc_frame_contacts_ptr, int32_t, crows)
PTR_GETTER(mdvc_frame_bridges_ptr, int32_t, brows)
PTR_GETTER(mdvc_frame_label_volumes_ptr, long long, vol)
PTR_GETTER(mdvc_frame_lut_ptr, int32_t, lut)
PTR_GETTER(mdvc_frame_component_counts_ptr, long long, ccount)

// test/debug access to the run tables: 0 rowoff(u32) 1 runstart(u32) 2 parent(u32) 3 lab(i32) 4 runcomp(i32)
extern "C" const void *mdvc_frame_debug_ptr(const void *ws, int X, int Y, int Z, const mdvc_frame_caps *caps, int which) {
    if (!ws || !caps_ok(caps) || X <= 0 || Y <= 0 || Z <= 0) return nullptr;
    Layout L = make_layout((long long)X * Y, caps);
    const size_t offs[5] = {L.rowoff, L.runstart, L.parent, L.lab, L.runcomp};
    if (which < 0 || which > 4) return nullptr;
    return at<char>(ws, offs[which]);
}

// gridops.cu -- seam-level kernels on dense int32 grids and the voxel<->atom gathers.
//
// Reference behaviour restated (file:line relative to /root/reference/mdvcontainment/):
//   contacts on a label grid   find_label_contacts.pyx:5-77
//   bridges on a label grid    find_bridges.pyx:4-89
//   relabel through a LUT      label.py:21-38
//   counts                     containment_main.py:510-512  (np.unique(..., return_counts=True))
//   mask + positions           containment_main.py:360-378  (np.isin + np.where)
//   atom gathers               atoms_voxels_mapping.pyx:17-91, mda_containment.py:191-203
#include "common.cuh"
#include "scan.cuh"

// ------------------------------------------------------------------------------------------
// shared pieces: per-block key cache in front of the global hash set, sort + decode
// ------------------------------------------------------------------------------------------
#define GBLOCKSET_SLOTS 512
struct GBlockSet {
    unsigned long long slot[GBLOCKSET_SLOTS];
    __device__ void clear() {
        for (int i = threadIdx.x; i < GBLOCKSET_SLOTS; i += blockDim.x) slot[i] = MDVC_KEY_EMPTY;
    }
    __device__ bool add(unsigned long long key) {
        uint32_t h = (uint32_t)mdvc_mix64(key) & (GBLOCKSET_SLOTS - 1);
        for (int probe = 0; probe < 16; ++probe) {
            unsigned long long cur = slot[h];
            if (cur == key) return false;
            if (cur == MDVC_KEY_EMPTY) {
                unsigned long long old = atomicCAS(&slot[h], MDVC_KEY_EMPTY, key);
                if (old == MDVC_KEY_EMPTY) return true;
                if (old == key) return false;
            }
            h = (h + 1) & (GBLOCKSET_SLOTS - 1);
        }
        return true;
    }
};

struct GSink {
    GBlockSet *cache;
    unsigned long long *slots;
    uint32_t n_slots;
    uint32_t *flags;
    uint32_t bit;
    unsigned long long last;
    __device__ __forceinline__ void put(unsigned long long key) {
        if (key == last) return;
        last = key;
        if (!cache->add(key)) return;
        if (!mdvc_set_insert(slots, n_slots, key)) atomicOr(flags, bit);
    }
};

__device__ __forceinline__ bool label_key_ok(int l) { return l > -MDVC_LABEL_BIAS && l < MDVC_LABEL_BIAS; }

// every voxel against its 13 lexicographically-previous neighbours, no wrap
__global__ void __launch_bounds__(256)
k_grid_contacts(const int32_t *__restrict__ L, int X, int Y, int Z, unsigned long long *slots, uint32_t n_slots,
                uint32_t *flags) {
    __shared__ GBlockSet cache;
    cache.clear();
    __syncthreads();
    GSink sink{&cache, slots, n_slots, flags, MDVC_OVERFLOW_CONTACTS, MDVC_KEY_EMPTY};
    long long N = (long long)X * Y * Z, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        int a = L[i];
        if (a == 0) continue;
        int z = (int)(i % Z);
        long long r = i / Z;
        int y = (int)(r % Y), x = (int)(r / Y);
#pragma unroll
        for (int s = 0; s < 13; ++s) {
            // (dx,dy,dz) enumerates the 13 offsets with (dx,dy,dz) < (0,0,0) lexicographically
            int dx = s < 9 ? -1 : 0;
            int dy = s < 9 ? s / 3 - 1 : (s < 12 ? -1 : 0);
            int dz = s < 12 ? s % 3 - 1 : -1;
            int nx = x + dx, ny = y + dy, nz = z + dz;
            if (nx < 0 || ny < 0 || ny >= Y || nz < 0 || nz >= Z) continue;
            int b = L[((long long)nx * Y + ny) * Z + nz];
            if (b == 0 || b == a) continue;
            int lo = a < b ? a : b, hi = a < b ? b : a;
            if (!label_key_ok(lo) || !label_key_ok(hi)) { atomicOr(flags, MDVC_OVERFLOW_LABELS); continue; }
            sink.put(mdvc_pack_key(lo, hi, 0));
        }
    }
}

// face voxels only.  Face voxel enumeration: [0, 2XY) z-faces, then 2XZ y-faces, then 2YZ x-faces
// (edge/corner voxels are visited more than once; the set dedups).
__global__ void __launch_bounds__(256)
k_grid_bridges(const int32_t *__restrict__ L, int X, int Y, int Z, unsigned long long *slots, uint32_t n_slots,
               uint32_t *flags) {
    __shared__ GBlockSet cache;
    cache.clear();
    __syncthreads();
    GSink sink{&cache, slots, n_slots, flags, MDVC_OVERFLOW_BRIDGES, MDVC_KEY_EMPTY};
    long long nz_f = 2ll * X * Y, ny_f = 2ll * X * Z, nx_f = 2ll * Y * Z;
    long long total = nz_f + ny_f + nx_f, stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        int x, y, z;
        if (t < nz_f) { long long q = t >> 1; z = (t & 1) ? Z - 1 : 0; y = (int)(q % Y); x = (int)(q / Y); }
        else if (t < nz_f + ny_f) { long long u = t - nz_f, q = u >> 1; y = (u & 1) ? Y - 1 : 0; z = (int)(q % Z); x = (int)(q / Z); }
        else { long long u = t - nz_f - ny_f, q = u >> 1; x = (u & 1) ? X - 1 : 0; z = (int)(q % Z); y = (int)(q / Z); }
        int a = L[((long long)x * Y + y) * Z + z];
        if (a == 0) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            int nx = x + dx, sx = 0;
            if (nx < 0) { nx = X - 1; sx = -1; } else if (nx >= X) { nx = 0; sx = 1; }
            for (int dy = -1; dy <= 1; ++dy) {
                int ny = y + dy, sy = 0;
                if (ny < 0) { ny = Y - 1; sy = -1; } else if (ny >= Y) { ny = 0; sy = 1; }
                for (int dz = -1; dz <= 1; ++dz) {
                    if (!(dx | dy | dz)) continue;
                    int nz = z + dz, sz = 0;
                    if (nz < 0) { nz = Z - 1; sz = -1; } else if (nz >= Z) { nz = 0; sz = 1; }
                    if (!(sx | sy | sz)) continue;
                    int b = L[((long long)nx * Y + ny) * Z + nz];
                    if (b == 0) continue;
                    if (!label_key_ok(a) || !label_key_ok(b)) { atomicOr(flags, MDVC_OVERFLOW_LABELS); continue; }
                    if (a <= b) sink.put(mdvc_pack_key(a, b, mdvc_shift_code(sx, sy, sz)));
                    else sink.put(mdvc_pack_key(b, a, mdvc_shift_code(-sx, -sy, -sz)));
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_g_compact(const unsigned long long *__restrict__ slots, uint32_t n_slots, unsigned long long *__restrict__ keys,
            uint32_t *count) {
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        unsigned long long k = slots[i];
        if (k != MDVC_KEY_EMPTY) keys[atomicAdd(count, 1u)] = k;
    }
}

// in-place variant used here: the set's own slot array is sorted (EMPTY = max key sorts last)
#define GSORT_SMEM_KEYS 4096
__global__ void __launch_bounds__(1024)
k_g_sort_slots(unsigned long long *a_glob, uint32_t P) {
    __shared__ unsigned long long sk[GSORT_SMEM_KEYS];
    bool in_smem = P <= GSORT_SMEM_KEYS;
    unsigned long long *a = in_smem ? sk : a_glob;
    if (in_smem) for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) sk[i] = a_glob[i];
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
    if (in_smem) for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) a_glob[i] = sk[i];
}

__global__ void __launch_bounds__(256)
k_g_decode(const unsigned long long *__restrict__ sorted_slots, uint32_t n_slots, int with_shift, int32_t *__restrict__ rows,
           uint32_t rows_cap, uint32_t *n_rows, uint32_t *flags, uint32_t overflow_bit) {
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        unsigned long long k = sorted_slots[i];
        if (k == MDVC_KEY_EMPTY) continue;
        bool last = (i + 1 == n_slots) || sorted_slots[i + 1] == MDVC_KEY_EMPTY;
        if (last) *n_rows = i + 1;
        if (i >= rows_cap) { if (last) atomicOr(flags, overflow_bit); continue; }
        int l1, l2, code;
        mdvc_unpack_key(k, l1, l2, code);
        if (with_shift) {
            int32_t *o = rows + (size_t)i * 5;
            o[0] = l1; o[1] = l2; o[2] = code / 9 - 1; o[3] = (code / 3) % 3 - 1; o[4] = code % 3 - 1;
        } else { rows[2 * (size_t)i] = l1; rows[2 * (size_t)i + 1] = l2; }
    }
}

static int grid_pairs(bool bridges, const int32_t *labels, int X, int Y, int Z, unsigned long long *set_slots,
                      uint32_t n_slots, int32_t *rows_out, uint32_t rows_cap, uint32_t *n_rows_dev, uint32_t *flags_dev,
                      cudaStream_t st) {
    if (!labels || !set_slots || !rows_out || !n_rows_dev || !flags_dev || X <= 0 || Y <= 0 || Z <= 0) return MDVC_ERR_BADARG;
    if (n_slots < 64 || (n_slots & (n_slots - 1))) return MDVC_ERR_BADARG;
    long long N = (long long)X * Y * Z;
    if (N >= (1ll << 31)) return MDVC_ERR_BADARG;
    MDVC_CUDA_CHECK(cudaMemsetAsync(set_slots, 0xff, (size_t)n_slots * 8, st));
    MDVC_CUDA_CHECK(cudaMemsetAsync(n_rows_dev, 0, sizeof(uint32_t), st));
    if (bridges) {
        long long faces = 2ll * X * Y + 2ll * X * Z + 2ll * Y * Z;
        k_grid_bridges<<<mdvc_grid(faces, 256, 8), 256, 0, st>>>(labels, X, Y, Z, set_slots, n_slots, flags_dev);
    } else {
        k_grid_contacts<<<mdvc_grid(N, 256, 8), 256, 0, st>>>(labels, X, Y, Z, set_slots, n_slots, flags_dev);
    }
    MDVC_LAUNCH_CHECK();
    k_g_sort_slots<<<1, 1024, 0, st>>>(set_slots, n_slots);
    MDVC_LAUNCH_CHECK();
    k_g_decode<<<mdvc_grid(n_slots, 256, 4), 256, 0, st>>>(set_slots, n_slots, bridges ? 1 : 0, rows_out, rows_cap, n_rows_dev,
                                                           flags_dev, bridges ? MDVC_OVERFLOW_BRIDGES : MDVC_OVERFLOW_CONTACTS);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

extern "C" int mdvc_grid_contacts(const int32_t *labels, int X, int Y, int Z, unsigned long long *set_slots, uint32_t n_slots,
                                  int32_t *rows_out, uint32_t rows_cap, uint32_t *n_rows_dev, uint32_t *flags_dev, void *stream) {
    return grid_pairs(false, labels, X, Y, Z, set_slots, n_slots, rows_out, rows_cap, n_rows_dev, flags_dev, (cudaStream_t)stream);
}
extern "C" int mdvc_grid_bridges(const int32_t *labels, int X, int Y, int Z, unsigned long long *set_slots, uint32_t n_slots,
                                 int32_t *rows_out, uint32_t rows_cap, uint32_t *n_rows_dev, uint32_t *flags_dev, void *stream) {
    return grid_pairs(true, labels, X, Y, Z, set_slots, n_slots, rows_out, rows_cap, n_rows_dev, flags_dev, (cudaStream_t)stream);
}

// Slab boundaries: every step from a voxel of plane A (labels la, x = last plane of one slab) to its 9
// neighbours in plane B (labels lb, x + 1 = first plane of the next slab), y and z periodic.  Rows are
// (la, lb, sx, sy, sz) with sx = sx_boundary (0 inside the box, +1 when the step leaves through the x-max
// face), NOT normalised: la and lb are numbered by different slabs, the caller maps them to global labels
// first.  sorted unique, int32 [K][5].
__global__ void __launch_bounds__(256)
k_plane_pairs(const int32_t *__restrict__ A, const int32_t *__restrict__ B, int Y, int Z, int sx_boundary, int periodic_yz,
              unsigned long long *slots, uint32_t n_slots, uint32_t *flags) {
    __shared__ GBlockSet cache;
    cache.clear();
    __syncthreads();
    GSink sink{&cache, slots, n_slots, flags, MDVC_OVERFLOW_BRIDGES, MDVC_KEY_EMPTY};
    long long n = (long long)Y * Z, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int la = A[i];
        if (la == 0) continue;
        int z = (int)(i % Z), y = (int)(i / Z);
        for (int dy = -1; dy <= 1; ++dy) {
            int ny = y + dy, sy = 0;
            if (ny < 0) { ny = Y - 1; sy = -1; } else if (ny >= Y) { ny = 0; sy = 1; }
            if (sy && !periodic_yz) continue;
            for (int dz = -1; dz <= 1; ++dz) {
                int nz = z + dz, sz = 0;
                if (nz < 0) { nz = Z - 1; sz = -1; } else if (nz >= Z) { nz = 0; sz = 1; }
                if (sz && !periodic_yz) continue;
                int lb = B[(long long)ny * Z + nz];
                if (lb == 0) continue;
                if (!label_key_ok(la) || !label_key_ok(lb)) { atomicOr(flags, MDVC_OVERFLOW_LABELS); continue; }
                sink.put(mdvc_pack_key(la, lb, mdvc_shift_code(sx_boundary, sy, sz)));
            }
        }
    }
}

extern "C" int mdvc_plane_pairs(const int32_t *plane_a, const int32_t *plane_b, int Y, int Z, int sx_boundary, int periodic_yz,
                                unsigned long long *set_slots, uint32_t n_slots, int32_t *rows_out, uint32_t rows_cap,
                                uint32_t *n_rows_dev, uint32_t *flags_dev, void *stream) {
    if (!plane_a || !plane_b || !set_slots || !rows_out || !n_rows_dev || !flags_dev || Y <= 0 || Z <= 0) return MDVC_ERR_BADARG;
    if (n_slots < 64 || (n_slots & (n_slots - 1)) || sx_boundary < -1 || sx_boundary > 1) return MDVC_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    MDVC_CUDA_CHECK(cudaMemsetAsync(set_slots, 0xff, (size_t)n_slots * 8, st));
    MDVC_CUDA_CHECK(cudaMemsetAsync(n_rows_dev, 0, sizeof(uint32_t), st));
    k_plane_pairs<<<mdvc_grid((long long)Y * Z, 256, 8), 256, 0, st>>>(plane_a, plane_b, Y, Z, sx_boundary, periodic_yz, set_slots,
                                                                      n_slots, flags_dev);
    MDVC_LAUNCH_CHECK();
    k_g_sort_slots<<<1, 1024, 0, st>>>(set_slots, n_slots);
    MDVC_LAUNCH_CHECK();
    k_g_decode<<<mdvc_grid(n_slots, 256, 4), 256, 0, st>>>(set_slots, n_slots, 1, rows_out, rows_cap, n_rows_dev, flags_dev,
                                                           MDVC_OVERFLOW_BRIDGES);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

// ------------------------------------------------------------------------------------------
// relabel through a LUT, histogram
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_grid_relabel(const int32_t *__restrict__ in, long long n, const int32_t *__restrict__ lut, int lo, int lut_n,
               int32_t *__restrict__ out) {
    long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n4 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 ? n / 4 : 0;
    for (long long q = i; q < n4; q += stride) {
        int4 v = __ldcs(reinterpret_cast<const int4 *>(in) + q), r;
        uint32_t a = (uint32_t)(v.x - lo), b = (uint32_t)(v.y - lo), c = (uint32_t)(v.z - lo), d = (uint32_t)(v.w - lo);
        r.x = a < (uint32_t)lut_n ? __ldg(lut + a) : 0; r.y = b < (uint32_t)lut_n ? __ldg(lut + b) : 0;
        r.z = c < (uint32_t)lut_n ? __ldg(lut + c) : 0; r.w = d < (uint32_t)lut_n ? __ldg(lut + d) : 0;
        __stcs(reinterpret_cast<int4 *>(out) + q, r);
    }
    for (long long q = n4 * 4 + i; q < n; q += stride) {
        uint32_t a = (uint32_t)(in[q] - lo);
        out[q] = a < (uint32_t)lut_n ? __ldg(lut + a) : 0;
    }
}

extern "C" int mdvc_grid_relabel(const int32_t *labels, long long n, const int32_t *lut, int lut_lo, int lut_n,
                                 int32_t *out, void *stream) {
    if (!labels || !lut || !out || n < 0 || lut_n <= 0) return MDVC_ERR_BADARG;
    if (n == 0) return MDVC_OK;
    k_grid_relabel<<<mdvc_grid(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(labels, n, lut, lut_lo, lut_n, out);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

// histogram with a shared-memory privatised copy when the bins fit
#define HIST_SMEM_BINS 4096
__global__ void __launch_bounds__(256)
k_grid_count(const int32_t *__restrict__ g, long long n, int lo, int n_bins, unsigned long long *counts) {
    __shared__ unsigned int sh[HIST_SMEM_BINS];
    bool priv = n_bins <= HIST_SMEM_BINS;
    if (priv) for (int i = threadIdx.x; i < n_bins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    long long stride = (long long)gridDim.x * blockDim.x;
    // a block never adds more than 2^32 voxels into one shared bin: flush every 2^24 rounds is not needed for n < 2^31
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t b = (uint32_t)(__ldcs(g + i) - lo);
        if (b >= (uint32_t)n_bins) continue;
        if (priv) atomicAdd(&sh[b], 1u); else atomicAdd(&counts[b], 1ull);
    }
    __syncthreads();
    if (priv)
        for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
            if (sh[i]) atomicAdd(&counts[i], (unsigned long long)sh[i]);
}

extern "C" int mdvc_grid_count(const int32_t *grid, long long n, int lo, int n_bins, unsigned long long *counts, void *stream) {
    if (!grid || !counts || n < 0 || n_bins <= 0) return MDVC_ERR_BADARG;
    if (n == 0) return MDVC_OK;
    k_grid_count<<<mdvc_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(grid, n, lo, n_bins, counts);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

// ------------------------------------------------------------------------------------------
// ordered selection (np.isin + np.where): tile counts -> scan -> scatter
// ------------------------------------------------------------------------------------------
#define SEL_TILE 2048   // items per block (256 threads x 8, blocked arrangement keeps C-order)

extern "C" size_t mdvc_grid_select_scratch_words(long long n) {
    size_t tiles = (size_t)((n + SEL_TILE - 1) / SEL_TILE) + 1;
    return tiles + mdvc_scan_scratch_elems((uint32_t)tiles) + 8;
}

__device__ __forceinline__ bool is_member(int v, const uint8_t *__restrict__ member, int lo, int n_member) {
    uint32_t b = (uint32_t)(v - lo);
    return b < (uint32_t)n_member && member[b] != 0;
}

__global__ void __launch_bounds__(256)
k_sel_count(const int32_t *__restrict__ g, long long n, const uint8_t *__restrict__ member, int lo, int n_member,
            uint32_t *__restrict__ tile_counts, uint8_t *__restrict__ mask_out) {
    __shared__ uint32_t sm[33];
    long long base = (long long)blockIdx.x * SEL_TILE + (long long)threadIdx.x * 8;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        long long i = base + k;
        if (i < n) {
            bool m = is_member(g[i], member, lo, n_member);
            c += m;
            if (mask_out) mask_out[i] = m;
        }
    }
    c = mdvc_block_reduce_sum<uint32_t>(c, sm);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(256)
k_sel_scatter(const int32_t *__restrict__ g, long long n, int Y, int Z, const uint8_t *__restrict__ member, int lo,
              int n_member, const uint32_t *__restrict__ tile_offsets, long long *__restrict__ pos_out,
              unsigned long long pos_cap) {
    __shared__ uint32_t sm[33];
    long long base = (long long)blockIdx.x * SEL_TILE + (long long)threadIdx.x * 8;
    bool m[8];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        long long i = base + k;
        m[k] = i < n && is_member(g[i], member, lo, n_member);
        c += m[k];
    }
    uint32_t total;
    unsigned long long off = (unsigned long long)tile_offsets[blockIdx.x] + mdvc_block_exscan<uint32_t>(c, sm, &total);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (!m[k]) continue;
        if (off < pos_cap) {
            long long i = base + k;
            long long z = i % Z, r = i / Z;
            pos_out[3 * off] = r / Y; pos_out[3 * off + 1] = r % Y; pos_out[3 * off + 2] = z;
        }
        ++off;
    }
}

__global__ void k_widen_count(const uint32_t *total32, unsigned long long *out) { *out = *total32; }

extern "C" int mdvc_grid_select(const int32_t *grid, int X, int Y, int Z, const uint8_t *member, int lo, int n_member,
                                uint32_t *scratch, long long *pos_out, unsigned long long pos_cap,
                                unsigned long long *n_out_dev, uint8_t *mask_out, void *stream) {
    if (!grid || !member || !scratch || !n_out_dev || X <= 0 || Y <= 0 || Z <= 0 || n_member <= 0) return MDVC_ERR_BADARG;
    long long n = (long long)X * Y * Z;
    if (n >= (1ll << 31)) return MDVC_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t tiles = (uint32_t)((n + SEL_TILE - 1) / SEL_TILE);
    uint32_t *tile_counts = scratch, *tile_scratch = scratch + tiles + 1, *total = tile_scratch + mdvc_scan_scratch_elems(tiles + 1);
    k_sel_count<<<tiles, 256, 0, st>>>(grid, n, member, lo, n_member, tile_counts, mask_out);
    MDVC_LAUNCH_CHECK();
    int rc = mdvc_exclusive_scan<uint32_t>(tile_counts, tile_counts, nullptr, tiles, tile_scratch, total, nullptr, st);
    if (rc) return rc;
    k_widen_count<<<1, 1, 0, st>>>(total, n_out_dev);
    MDVC_LAUNCH_CHECK();
    if (pos_out && pos_cap) {
        k_sel_scatter<<<tiles, 256, 0, st>>>(grid, n, Y, Z, member, lo, n_member, tile_counts, pos_out, pos_cap);
        MDVC_LAUNCH_CHECK();
    }
    return MDVC_OK;
}

// ------------------------------------------------------------------------------------------
// atoms
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_atom_components(const int32_t *__restrict__ vox, long long n, const int32_t *__restrict__ grid, int X, int Y, int Z,
                  double *__restrict__ out) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int x = vox[3 * i], y = vox[3 * i + 1], z = vox[3 * i + 2];
        bool ok = x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z;
        out[i] = ok ? (double)grid[((long long)x * Y + y) * Z + z] : 0.0;
    }
}

extern "C" int mdvc_atom_components(const int32_t *vox, long long n, const int32_t *grid, int X, int Y, int Z,
                                    double *out, void *stream) {
    if (n < 0 || (n > 0 && (!vox || !out)) || !grid || X <= 0 || Y <= 0 || Z <= 0) return MDVC_ERR_BADARG;
    if (n == 0) return MDVC_OK;
    k_atom_components<<<mdvc_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(vox, n, grid, X, Y, Z, out);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

// --- stable LSD radix sort of (voxel linear index, atom position) -----------------------------
// 8-bit digits; keys are < 2^31 so four passes.  Stability per pass comes from the blocked
// arrangement: a tile's items are ranked in index order and tiles are offset in tile order.
#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)

extern "C" size_t mdvc_atom_sort_scratch_bytes(long long n) {
    size_t tiles = (size_t)((n + RS_TILE - 1) / RS_TILE) + 1;
    size_t hist = tiles * 256;                                   // digit-major tile histograms
    return (size_t)n * 8 + (size_t)n * 4 + hist * 4 + mdvc_scan_scratch_elems((uint32_t)hist) * 4 + 256;
}

__global__ void __launch_bounds__(256)
k_atom_keys(const int32_t *__restrict__ vox, long long n, int Y, int Z, unsigned long long *__restrict__ keys,
            uint32_t *__restrict__ order) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        keys[i] = (unsigned long long)(((long long)vox[3 * i] * Y + vox[3 * i + 1]) * Z + vox[3 * i + 2]);
        order[i] = (uint32_t)i;
    }
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const unsigned long long *__restrict__ keys, long long n, int shift, uint32_t tiles, uint32_t *__restrict__ hist) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    long long base = (long long)blockIdx.x * RS_TILE;
    for (int k = 0; k < RS_ITEMS; ++k) {
        long long i = base + (long long)k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&sh[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * tiles + blockIdx.x] = sh[threadIdx.x];   // digit-major
}

// Stable scatter of one 8-bit digit.  Items keep their order li = k * RS_THREADS + thread inside the tile, i.e. the 64
// "rows" (k, warp) in ascending order and the lanes of a row in ascending order.  The rank of an item among the equal
// digits of its tile = (items with that digit in earlier rows) + (earlier lanes of its row with that digit): the second
// term comes from one __match_any_sync per item, the first from a per-row digit count in shared memory that 256
// threads (one per digit) turn into a prefix over the 64 rows.  The first version let every digit's thread walk the
// whole tile (2048 serial shared-memory reads per thread): 1.3 ms per pass for 20 M atoms, 5 % of the DRAM roofline.
__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const unsigned long long *__restrict__ keys_in, const uint32_t *__restrict__ val_in, long long n, int shift,
             uint32_t tiles, const uint32_t *__restrict__ hist_scanned, unsigned long long *__restrict__ keys_out,
             uint32_t *__restrict__ val_out) {
    constexpr int ROWS = RS_ITEMS * (RS_THREADS / 32);
    __shared__ uint32_t digit_base[256];
    __shared__ unsigned short cnt[ROWS][256];                    // items of digit d in row r, then: in the rows before r
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long base = (long long)blockIdx.x * RS_TILE;
    digit_base[tid] = hist_scanned[(size_t)tid * tiles + blockIdx.x];
    {
        uint32_t *z = reinterpret_cast<uint32_t *>(&cnt[0][0]);
        for (int q = tid; q < ROWS * 256 / 2; q += RS_THREADS) z[q] = 0u;
    }
    __syncthreads();
    unsigned long long key[RS_ITEMS];
    uint32_t digit[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const long long i = base + (long long)k * RS_THREADS + tid;
        const bool valid = i < n;
        key[k] = valid ? keys_in[i] : 0ull;
        const uint32_t d = valid ? (uint32_t)((key[k] >> shift) & 255u) : 256u;   // the invalid lanes only match each other
        const unsigned peers = __match_any_sync(MDVC_FULL, d);
        rank[k] = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank[k] == 0) cnt[k * (RS_THREADS / 32) + warp][d] = (unsigned short)__popc(peers);
        digit[k] = d;
    }
    __syncthreads();
    {
        uint32_t run = 0;
#pragma unroll 8
        for (int r = 0; r < ROWS; ++r) {
            const uint32_t c = cnt[r][tid];
            cnt[r][tid] = (unsigned short)run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const long long i = base + (long long)k * RS_THREADS + tid;
        if (i < n) {
            const uint32_t o = digit_base[digit[k]] + cnt[k * (RS_THREADS / 32) + warp][digit[k]] + rank[k];
            keys_out[o] = key[k];
            val_out[o] = val_in[i];
        }
    }
}

extern "C" int mdvc_atom_sort_by_voxel(const int32_t *vox, long long n, int X, int Y, int Z, uint32_t *order_out,
                                       unsigned long long *keys_out, void *scratch, void *stream) {
    if (n < 0 || n >= (1ll << 31) || X <= 0 || Y <= 0 || Z <= 0 || (n > 0 && (!vox || !order_out || !keys_out || !scratch)))
        return MDVC_ERR_BADARG;
    if ((long long)X * Y * Z >= (1ll << 31)) return MDVC_ERR_BADARG;
    if (n == 0) return MDVC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t tiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    size_t hist_n = (size_t)tiles * 256;
    char *p = reinterpret_cast<char *>(scratch);
    unsigned long long *keys_tmp = reinterpret_cast<unsigned long long *>(p); p += (size_t)n * 8;
    uint32_t *val_tmp = reinterpret_cast<uint32_t *>(p); p += (size_t)n * 4;
    p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
    uint32_t *hist = reinterpret_cast<uint32_t *>(p); p += hist_n * 4;
    uint32_t *tile_scratch = reinterpret_cast<uint32_t *>(p);
    k_atom_keys<<<mdvc_grid(n, 256, 8), 256, 0, st>>>(vox, n, Y, Z, keys_out, order_out);
    MDVC_LAUNCH_CHECK();
    unsigned long long maxkey = (unsigned long long)X * Y * Z - 1;
    int passes = 0;
    while (passes < 8 && (maxkey >> (8 * passes)) != 0) ++passes;
    if (passes == 0) passes = 1;
    if (passes & 1) ++passes;                       // even number of passes: result lands in the output arrays
    unsigned long long *kin = keys_out, *kout = keys_tmp;
    uint32_t *vin = order_out, *vout = val_tmp;
    for (int pass = 0; pass < passes; ++pass) {
        int shift = 8 * pass;
        k_rs_hist<<<tiles, RS_THREADS, 0, st>>>(kin, n, shift, tiles, hist);
        MDVC_LAUNCH_CHECK();
        int rc = mdvc_exclusive_scan<uint32_t>(hist, hist, nullptr, (uint32_t)hist_n, tile_scratch, nullptr, nullptr, st);
        if (rc) return rc;
        k_rs_scatter<<<tiles, RS_THREADS, 0, st>>>(kin, vin, n, shift, tiles, hist, kout, vout);
        MDVC_LAUNCH_CHECK();
        unsigned long long *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    return MDVC_OK;
}

// selection over the CSR order: keep atom order[i] when the value of its voxel is a member
__global__ void __launch_bounds__(256)
k_asel_count(const unsigned long long *__restrict__ keys, long long n, const int32_t *__restrict__ grid,
             const uint8_t *__restrict__ member, int lo, int n_member, uint32_t *__restrict__ tile_counts) {
    __shared__ uint32_t sm[33];
    long long base = (long long)blockIdx.x * SEL_TILE + (long long)threadIdx.x * 8;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        long long i = base + k;
        if (i < n) c += is_member(keys ? grid[keys[i]] : grid[i], member, lo, n_member);
    }
    c = mdvc_block_reduce_sum<uint32_t>(c, sm);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(256)
k_asel_scatter(const uint32_t *__restrict__ order, const unsigned long long *__restrict__ keys, long long n,
               const int32_t *__restrict__ grid, const uint8_t *__restrict__ member, int lo, int n_member,
               const long long *__restrict__ atom_ix, const uint32_t *__restrict__ tile_offsets, long long *__restrict__ out) {
    __shared__ uint32_t sm[33];
    long long base = (long long)blockIdx.x * SEL_TILE + (long long)threadIdx.x * 8;
    bool m[8];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        long long i = base + k;
        m[k] = i < n && is_member(keys ? grid[keys[i]] : grid[i], member, lo, n_member);
        c += m[k];
    }
    uint32_t total;
    unsigned long long off = (unsigned long long)tile_offsets[blockIdx.x] + mdvc_block_exscan<uint32_t>(c, sm, &total);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (m[k]) { uint32_t a = order[base + k]; out[off++] = atom_ix ? atom_ix[a] : (long long)a; }
}

extern "C" int mdvc_atom_select(const uint32_t *order, const unsigned long long *keys, long long n, const int32_t *grid,
                                const uint8_t *member, int lo, int n_member, const long long *atom_ix, uint32_t *scratch,
                                long long *out, unsigned long long *n_out_dev, void *stream) {
    if (n < 0 || n >= (1ll << 31) || !grid || !member || !scratch || !n_out_dev || n_member <= 0 ||
        (n > 0 && (!order || !out)))                       // keys == NULL: grid holds one value per CSR position
        return MDVC_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { MDVC_CUDA_CHECK(cudaMemsetAsync(n_out_dev, 0, 8, st)); return MDVC_OK; }
    uint32_t tiles = (uint32_t)((n + SEL_TILE - 1) / SEL_TILE);
    uint32_t *tile_counts = scratch, *tile_scratch = scratch + tiles + 1, *total = tile_scratch + mdvc_scan_scratch_elems(tiles + 1);
    k_asel_count<<<tiles, 256, 0, st>>>(keys, n, grid, member, lo, n_member, tile_counts);
    MDVC_LAUNCH_CHECK();
    int rc = mdvc_exclusive_scan<uint32_t>(tile_counts, tile_counts, nullptr, tiles, tile_scratch, total, nullptr, st);
    if (rc) return rc;
    k_widen_count<<<1, 1, 0, st>>>(total, n_out_dev);
    MDVC_LAUNCH_CHECK();
    k_asel_scatter<<<tiles, 256, 0, st>>>(order, keys, n, grid, member, lo, n_member, atom_ix, tile_counts, out);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

// gather for an explicit voxel list (voxels2atomgroup_cy, atoms_voxels_mapping.pyx:59-91): for every query voxel,
// in the given order, the atoms of that voxel in atom order.  Voxels without atoms (or outside the grid) add nothing.
__global__ void __launch_bounds__(256)
k_vox_ranges(const int32_t *__restrict__ q, long long m, int X, int Y, int Z, const unsigned long long *__restrict__ keys,
             long long n, uint32_t *__restrict__ start, uint32_t *__restrict__ cnt) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        int x = q[3 * i], y = q[3 * i + 1], z = q[3 * i + 2];
        uint32_t lo = 0, c = 0;
        if (x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z) {
            unsigned long long key = (unsigned long long)(((long long)x * Y + y) * Z + z);
            long long a = 0, b = n;                       // lower bound
            while (a < b) { long long mid = (a + b) >> 1; if (keys[mid] < key) a = mid + 1; else b = mid; }
            long long first = a;
            b = n;                                        // upper bound
            while (a < b) { long long mid = (a + b) >> 1; if (keys[mid] <= key) a = mid + 1; else b = mid; }
            lo = (uint32_t)first; c = (uint32_t)(a - first);
        }
        start[i] = lo; cnt[i] = c;
    }
}

__global__ void __launch_bounds__(256)
k_vox_gather(long long m, const uint32_t *__restrict__ start, const uint32_t *__restrict__ cnt_raw,
             const uint32_t *__restrict__ off, const uint32_t *__restrict__ order, const long long *__restrict__ atom_ix,
             long long *__restrict__ out, unsigned long long out_cap) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        uint32_t s0 = start[i], c = cnt_raw[i];
        unsigned long long o = off[i];
        for (uint32_t k = 0; k < c && o + k < out_cap; ++k) {
            uint32_t a = order[s0 + k];
            out[o + k] = atom_ix ? atom_ix[a] : (long long)a;
        }
    }
}

extern "C" size_t mdvc_atom_gather_scratch_words(long long m) {
    return (size_t)m * 3 + mdvc_scan_scratch_elems((uint32_t)(m > 0 ? m : 1)) + 8;
}

extern "C" int mdvc_atom_gather_voxels(const int32_t *query, long long m, int X, int Y, int Z, const uint32_t *order,
                                       const unsigned long long *keys, long long n, const long long *atom_ix,
                                       uint32_t *scratch, long long *out, unsigned long long out_cap,
                                       unsigned long long *n_out_dev, void *stream) {
    if (m < 0 || n < 0 || m >= (1ll << 31) || n >= (1ll << 31) || !scratch || !n_out_dev || X <= 0 || Y <= 0 || Z <= 0 ||
        (m > 0 && !query) || (n > 0 && (!order || !keys)))
        return MDVC_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0) { MDVC_CUDA_CHECK(cudaMemsetAsync(n_out_dev, 0, 8, st)); return MDVC_OK; }
    uint32_t *start = scratch, *cnt = scratch + m, *off = scratch + 2 * m, *tile_scratch = scratch + 3 * m;
    uint32_t *total = tile_scratch + mdvc_scan_scratch_elems((uint32_t)m);
    k_vox_ranges<<<mdvc_grid(m, 256, 8), 256, 0, st>>>(query, m, X, Y, Z, keys, n, start, cnt);
    MDVC_LAUNCH_CHECK();
    int rc = mdvc_exclusive_scan<uint32_t>(cnt, off, nullptr, (uint32_t)m, tile_scratch, total, nullptr, st);
    if (rc) return rc;
    k_widen_count<<<1, 1, 0, st>>>(total, n_out_dev);
    MDVC_LAUNCH_CHECK();
    if (out && out_cap) {
        k_vox_gather<<<mdvc_grid(m, 256, 8), 256, 0, st>>>(m, start, cnt, off, order, atom_ix, out, out_cap);
        MDVC_LAUNCH_CHECK();
    }
    return MDVC_OK;
}

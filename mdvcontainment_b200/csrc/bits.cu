// bits.cu -- bit-packed occupancy grid: pack/unpack, atom binning, periodic morphology.
//
// Reference behaviour restated (file:line relative to /root/reference/mdvcontainment/):
//   binning      atomgroup_to_voxels.py:131-142, 192-197
//   morphology   atomgroup_to_voxels.py:32-92, 242-308
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

// ------------------------------------------------------------------------------------------
// error state
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void mdvc_set_error(const char *where, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
}
void mdvc_set_error_msg(const char *msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }

unsigned long long g_mdvc_launches = 0;
extern "C" unsigned long long mdvc_launch_count(void) { return g_mdvc_launches; }
extern "C" int mdvc_version(void) { return 100; }
extern "C" const char *mdvc_last_error(void) { return g_err; }
extern "C" int mdvc_row_pitch_words(int Z) { return ((mdvc_words(Z) + 3) / 4) * 4; }

// ------------------------------------------------------------------------------------------
// host staging memory
// ------------------------------------------------------------------------------------------
// Page-locked host memory for coordinate staging buffers.  write_combined: uncached on the CPU side -- right for
// buffers the host only ever fills front to back and the GPU reads by DMA (no cache snooping on the PCIe reads);
// never read such a buffer from the CPU.
extern "C" int mdvc_host_alloc(size_t bytes, int write_combined, void **ptr_out_host) {
    if (!ptr_out_host || bytes == 0) return MDVC_ERR_BADARG;
    *ptr_out_host = nullptr;
    MDVC_CUDA_CHECK(cudaHostAlloc(ptr_out_host, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
    return MDVC_OK;
}
extern "C" int mdvc_host_free(void *ptr_host) {
    if (!ptr_host) return MDVC_OK;
    MDVC_CUDA_CHECK(cudaFreeHost(ptr_host));
    return MDVC_OK;
}

// ------------------------------------------------------------------------------------------
// T1  pack / unpack
// ------------------------------------------------------------------------------------------
// One warp packs 32 voxels per ballot; lanes walk z so byte loads are contiguous.
__global__ void __launch_bounds__(256)
k_pack_bool(const uint8_t *__restrict__ grid, long long rows, int Z, int pitch, uint32_t *__restrict__ bits) {
    int lane = threadIdx.x & 31;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    long long items = rows * pitch;                     // one item = one output word
    for (long long it = warp; it < items; it += nwarps) {
        long long row = it / pitch;
        int w = (int)(it - row * pitch);
        int z = (w << 5) + lane;
        bool on = (z < Z) && grid[row * (long long)Z + z] != 0;
        uint32_t word = __ballot_sync(MDVC_FULL, on);
        if (lane == 0) bits[it] = word;
    }
}

__global__ void __launch_bounds__(256)
k_unpack_bool(const uint32_t *__restrict__ bits, long long rows, int Z, int pitch, uint8_t *__restrict__ grid) {
    long long n = rows * (long long)Z;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        long long row = i / Z;
        int z = (int)(i - row * Z);
        grid[i] = (bits[row * pitch + (z >> 5)] >> (z & 31)) & 1u;
    }
}

__global__ void __launch_bounds__(256)
k_popcount(const uint32_t *__restrict__ bits, long long rows, int Z, int pitch, unsigned long long *out) {
    long long items = rows * pitch;
    long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += stride) {
        int w = (int)(i % pitch);
        acc += __popc(bits[i] & mdvc_valid_mask(w, Z));
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(MDVC_FULL, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

static bool bad_dims(int X, int Y, int Z, int pitch) {
    return X <= 0 || Y <= 0 || Z <= 0 || pitch < mdvc_words(Z) ||
           (long long)X * Y * Z >= (1ll << 31);
}

extern "C" int mdvc_pack_bool(const uint8_t *grid, int X, int Y, int Z, uint32_t *bits, int pitch, void *stream) {
    if (!grid || !bits || bad_dims(X, Y, Z, pitch)) return MDVC_ERR_BADARG;
    long long rows = (long long)X * Y;
    k_pack_bool<<<mdvc_grid(rows * pitch * 32, 256), 256, 0, (cudaStream_t)stream>>>(grid, rows, Z, pitch, bits);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

extern "C" int mdvc_unpack_bool(const uint32_t *bits, int X, int Y, int Z, int pitch, uint8_t *grid, void *stream) {
    if (!grid || !bits || bad_dims(X, Y, Z, pitch)) return MDVC_ERR_BADARG;
    long long rows = (long long)X * Y;
    k_unpack_bool<<<mdvc_grid(rows * Z, 256), 256, 0, (cudaStream_t)stream>>>(bits, rows, Z, pitch, grid);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

extern "C" int mdvc_popcount(const uint32_t *bits, int X, int Y, int Z, int pitch, unsigned long long *count_dev, void *stream) {
    if (!bits || !count_dev || bad_dims(X, Y, Z, pitch)) return MDVC_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    MDVC_CUDA_CHECK(cudaMemsetAsync(count_dev, 0, sizeof(unsigned long long), st));
    long long rows = (long long)X * Y;
    k_popcount<<<mdvc_grid(rows * pitch, 256, 8), 256, 0, st>>>(bits, rows, Z, pitch, count_dev);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

// ------------------------------------------------------------------------------------------
// A1  binning
// ------------------------------------------------------------------------------------------
struct BinParams {
    double T[9];
    double invT[9];
    long long nbox[9];
    int nbox32[9];       // the same matrix when every entry fits (fits32), for the all-32-bit fast path
    int fits32;
    float Tf[9];         // T rounded to float32: the certified single-precision floor of bin_floor_f32
    float Teps[3];       // 5e-7 x the column sums |T0j| + |T1j| + |T2j| (float32, rounded up)
    float mmax;          // 1e5 / the largest column sum: beyond it the single-precision floor is not attempted
};

// Python floor division a // b for b > 0.  Atoms are almost always inside the box or one image away, so
// the 64-bit division (hundreds of instructions on the GPU) is kept for the rare far-away atom only.
__device__ __forceinline__ long long floordiv_ll(long long a, long long b) {
    if (a >= 0) {
        if (a < b) return 0;
        if (a < 2 * b) return 1;
    } else if (a >= -b) {
        return -1;
    }
    long long q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

// One atom: fractional voxel coordinates as the FMA chain fma(p2,T2j, fma(p1,T1j, p0*T0j)) in float64 --
// bit-identical to the reference's NumPy/OpenBLAS dgemm; __dmul_rn/__fma_rn keep the compiler
// from contracting differently -- floor, triclinic wrap (z, y, x; Python floor division), optional outputs.
// Fast path: when the floors and the box matrix fit 28 bits (any realistic box) everything after the floor is
// 32-bit integer work; the general path keeps 64-bit integers like the reference's int64 arrays.
// Returns true when the wrapped voxel (vox) lies inside the grid; newpos = the rewritten position when want_pos.
__device__ __forceinline__ bool bin_one_atom(float x, float y, float z, const BinParams &P, bool want_pos, int (&vox)[3],
                                             float (&newpos)[3]) {
    const double p0 = (double)x, p1 = (double)y, p2 = (double)z;
    double fj[3], fl[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        fj[j] = __fma_rn(p2, P.T[6 + j], __fma_rn(p1, P.T[3 + j], __dmul_rn(p0, P.T[j])));
        fl[j] = floor(fj[j]);
    }
    const double lim = 268435456.0;                      // 2^28
    if (P.fits32 && fabs(fl[0]) < lim && fabs(fl[1]) < lim && fabs(fl[2]) < lim) {
        const int u0 = __double2int_rz(fl[0]), u1 = __double2int_rz(fl[1]), u2 = __double2int_rz(fl[2]);
        int w3[3] = {u0, u1, u2};
#pragma unroll
        for (int dim = 2; dim >= 0; --dim) {
            const int nn = P.nbox32[4 * dim], a = w3[dim];
            int sh;
            if (a >= 0) sh = a < nn ? 0 : (a < 2 * nn ? 1 : a / nn);
            else sh = a >= -nn ? -1 : -((-a + nn - 1) / nn);
#pragma unroll
            for (int j = 0; j < 3; ++j) w3[j] -= sh * P.nbox32[3 * dim + j];
        }
        if (want_pos) {
            // (voxel + fractional part) back through inv(T): atomgroup_to_voxels.py:142
            const double g0 = __dadd_rn((double)w3[0], __dsub_rn(fj[0], (double)u0)),
                         g1 = __dadd_rn((double)w3[1], __dsub_rn(fj[1], (double)u1)),
                         g2 = __dadd_rn((double)w3[2], __dsub_rn(fj[2], (double)u2));
#pragma unroll
            for (int j = 0; j < 3; ++j)
                newpos[j] = (float)__fma_rn(g2, P.invT[6 + j], __fma_rn(g1, P.invT[3 + j], __dmul_rn(g0, P.invT[j])));
        }
        vox[0] = w3[0]; vox[1] = w3[1]; vox[2] = w3[2];
        return (unsigned)w3[0] < (unsigned)P.nbox32[0] && (unsigned)w3[1] < (unsigned)P.nbox32[4] &&
               (unsigned)w3[2] < (unsigned)P.nbox32[8];
    }
    long long v[3];
    double f[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { v[j] = (long long)fl[j]; f[j] = __dsub_rn(fj[j], (double)v[j]); }
#pragma unroll
    for (int dim = 2; dim >= 0; --dim) {
        long long sft = floordiv_ll(v[dim], P.nbox[4 * dim]);
#pragma unroll
        for (int j = 0; j < 3; ++j) v[j] -= sft * P.nbox[3 * dim + j];
    }
    if (want_pos) {
        const double g0 = __dadd_rn((double)v[0], f[0]), g1 = __dadd_rn((double)v[1], f[1]), g2 = __dadd_rn((double)v[2], f[2]);
#pragma unroll
        for (int j = 0; j < 3; ++j)
            newpos[j] = (float)__fma_rn(g2, P.invT[6 + j], __fma_rn(g1, P.invT[3 + j], __dmul_rn(g0, P.invT[j])));
    }
    vox[0] = (int)v[0]; vox[1] = (int)v[1]; vox[2] = (int)v[2];
    return v[0] >= 0 && v[0] < P.nbox[0] && v[1] >= 0 && v[1] < P.nbox[4] && v[2] >= 0 && v[2] < P.nbox[8];
}

// Certified single-precision floor.  The reference floors the float64 chain above; almost every bead lies nowhere near
// a voxel face, and for those the same floor follows from float32 arithmetic.  With a = |p0 T0j| + |p1 T1j| + |p2 T2j|
// the float32 chain f = fma(p2, T2j, fma(p1, T1j, p0 * T0j)) differs from the exact value by less than 4 * 2^-24 * a
// = 2.4e-7 * a (every term passes through at most four roundings: T to float32, the product, the two sums), the float64
// chain by ~1e-13 * a, so a float32 result whose distance to the nearest integer exceeds 5e-7 * a (twice the bound) has
// the reference's floor.  a is bounded from above by max|p_i| * (|T0j| + |T1j| + |T2j|), one multiplication per column
// instead of three (the looser bound leaves a few more beads undecided, which is safe).  Anything closer -- beads ON a
// face, as coordinates rounded to 0.01 A often are, huge coordinates, NaNs -- returns false and takes the float64
// path (0.3 % of the beads of the 1024^3 scene).  ~12 instructions per column instead of ~60.
__device__ __forceinline__ bool bin_floor_f32(float x, float y, float z, const BinParams &P, int (&u)[3]) {
    const float m = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z));
    bool sure = m < P.mmax;                                       // |f| < 1e5: exact f - floor(f), int conversion safe
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float f = fmaf(z, P.Tf[6 + j], fmaf(y, P.Tf[3 + j], x * P.Tf[j]));
        const float fl = floorf(f), frac = f - fl, eps = fmaf(m, P.Teps[j], 1e-30f);
        sure = sure && frac >= eps && frac <= 1.0f - eps;         // false for NaNs too
        u[j] = (int)fl;
    }
    return sure;
}

// triclinic-aware wrap of floored voxel coordinates (dims 2, 1, 0; Python floor division), 32-bit
__device__ __forceinline__ bool bin_wrap32(const BinParams &P, int (&w3)[3]) {
    // already inside the box (nearly every bead): each of the three steps below would shift by zero
    if ((unsigned)w3[0] < (unsigned)P.nbox32[0] && (unsigned)w3[1] < (unsigned)P.nbox32[4] && (unsigned)w3[2] < (unsigned)P.nbox32[8])
        return true;
#pragma unroll
    for (int dim = 2; dim >= 0; --dim) {
        const int nn = P.nbox32[4 * dim], a = w3[dim];
        int sh;
        if (a >= 0) sh = a < nn ? 0 : (a < 2 * nn ? 1 : a / nn);
        else sh = a >= -nn ? -1 : -((-a + nn - 1) / nn);
#pragma unroll
        for (int j = 0; j < 3; ++j) w3[j] -= sh * P.nbox32[3 * dim + j];
    }
    return (unsigned)w3[0] < (unsigned)P.nbox32[0] && (unsigned)w3[1] < (unsigned)P.nbox32[4] &&
           (unsigned)w3[2] < (unsigned)P.nbox32[8];
}

// Four consecutive atoms per thread.  Positions are AoS (x,y,z per atom = 12 B), so four atoms are exactly three
// 16-byte words: each thread fetches them with three streaming float4 loads (the three loads of a warp cover one
// contiguous 1536-byte span, every sector of it used), runs the four independent FMA chains -- that instruction-level
// parallelism is what hides the load latency -- and fires its REDs.  No shared memory and no barriers: the previous
// one-atom-per-thread kernel staged tiles through shared memory and spent 11.6 stall cycles per issued instruction
// at its two barriers and 12.8 on the tile loads (ncu, profiles/r1_bin_atoms_ncu_summary.txt: 31 % issue-active).
// The optional outputs leave as 16-byte stores as well.  pos_out may alias pos: a thread reads its four atoms
// completely before it writes them.
__global__ void __launch_bounds__(256, 4)
k_bin_atoms4(const float *__restrict__ pos, long long n, BinParams P, uint32_t *__restrict__ occ, int pitch,
             int32_t *__restrict__ vox_out, float *pos_out) {
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += stride) {
        const float4 *src = reinterpret_cast<const float4 *>(pos) + 3 * g;
        const float4 q0 = __ldcs(src), q1 = __ldcs(src + 1), q2 = __ldcs(src + 2);
        const float f[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
        int32_t v[12];
        float w[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int vk[3];
            float wk[3];
            const bool inside = bin_one_atom(f[3 * k], f[3 * k + 1], f[3 * k + 2], P, pos_out != nullptr, vk, wk);
#pragma unroll
            for (int j = 0; j < 3; ++j) { v[3 * k + j] = vk[j]; w[3 * k + j] = wk[j]; }
            if (occ && inside) {
                // fire-and-forget RED.OR; inside the grid => the row index fits 32 bits (fewer than 2^31 voxels)
                const uint32_t row = (uint32_t)vk[0] * (uint32_t)P.nbox[4] + (uint32_t)vk[1];
                atomicOr(&occ[(size_t)row * pitch + ((uint32_t)vk[2] >> 5)], 1u << (vk[2] & 31));
            }
        }
        if (vox_out) {
            int4 *dst = reinterpret_cast<int4 *>(vox_out) + 3 * g;
            dst[0] = make_int4(v[0], v[1], v[2], v[3]);
            dst[1] = make_int4(v[4], v[5], v[6], v[7]);
            dst[2] = make_int4(v[8], v[9], v[10], v[11]);
        }
        if (pos_out) {
            float4 *dst = reinterpret_cast<float4 *>(pos_out) + 3 * g;
            dst[0] = make_float4(w[0], w[1], w[2], w[3]);
            dst[1] = make_float4(w[4], w[5], w[6], w[7]);
            dst[2] = make_float4(w[8], w[9], w[10], w[11]);
        }
    }
}

// One atom per thread with plain loads and stores: the last n % 4 atoms, and whole arrays whose pointers are not
// 16-byte aligned (views into the middle of a coordinate array).
__global__ void __launch_bounds__(256)
k_bin_atoms1(const float *pos, long long first, long long n, BinParams P, uint32_t *__restrict__ occ, int pitch,
             int32_t *__restrict__ vox_out, float *pos_out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int vk[3];
        float wk[3] = {0.f, 0.f, 0.f};
        const bool inside = bin_one_atom(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], P, pos_out != nullptr, vk, wk);
        if (vox_out) { vox_out[3 * i] = vk[0]; vox_out[3 * i + 1] = vk[1]; vox_out[3 * i + 2] = vk[2]; }
        if (pos_out) { pos_out[3 * i] = wk[0]; pos_out[3 * i + 1] = wk[1]; pos_out[3 * i + 2] = wk[2]; }
        if (occ && inside) {
            const uint32_t row = (uint32_t)vk[0] * (uint32_t)P.nbox[4] + (uint32_t)vk[1];
            atomicOr(&occ[(size_t)row * pitch + ((uint32_t)vk[2] >> 5)], 1u << (vk[2] & 31));
        }
    }
}

static void bin_params(BinParams &P, const double *T_host, const double *invT_host, const long long *nbox_host) {
    memcpy(P.T, T_host, sizeof(P.T));
    memcpy(P.invT, invT_host, sizeof(P.invT));
    memcpy(P.nbox, nbox_host, sizeof(P.nbox));
    P.fits32 = 1;
    for (int k = 0; k < 9; ++k) {
        if (nbox_host[k] <= -(1ll << 28) || nbox_host[k] >= (1ll << 28)) P.fits32 = 0;
        P.nbox32[k] = (int)nbox_host[k];
        P.Tf[k] = (float)T_host[k];
    }
    float cmax = 0.f;
    for (int j = 0; j < 3; ++j) {
        const float col = nextafterf((fabsf(P.Tf[j]) + fabsf(P.Tf[3 + j]) + fabsf(P.Tf[6 + j])) * 1.000001f, INFINITY);
        P.Teps[j] = nextafterf(col * 5e-7f, INFINITY);
        cmax = fmaxf(cmax, col);
    }
    P.mmax = cmax > 0.f ? 1e5f / cmax : 0.f;
}

extern "C" int mdvc_bin_atoms(const float *pos, long long n, const double *T_host, const double *invT_host,
                              const long long *nbox_host, uint32_t *occ_bits, int pitch,
                              int32_t *vox_out, float *pos_out, void *stream) {
    if (n < 0 || !T_host || !invT_host || !nbox_host || (n > 0 && !pos)) return MDVC_ERR_BADARG;
    if (nbox_host[0] <= 0 || nbox_host[4] <= 0 || nbox_host[8] <= 0) return MDVC_ERR_BADARG;
    if (occ_bits && pitch < mdvc_words((int)nbox_host[8])) return MDVC_ERR_BADARG;
    if (n == 0) return MDVC_OK;
    BinParams P;
    bin_params(P, T_host, invT_host, nbox_host);
    cudaStream_t st = (cudaStream_t)stream;
    const uintptr_t ptrs = (uintptr_t)pos | (uintptr_t)vox_out | (uintptr_t)pos_out;
    long long done = 0;
    if ((ptrs & 15) == 0 && n >= 4) {
        k_bin_atoms4<<<mdvc_grid(n >> 2, 256, 32), 256, 0, st>>>(pos, n, P, occ_bits, pitch, vox_out, pos_out);
        MDVC_LAUNCH_CHECK();
        done = n & ~3ll;
    }
    if (done < n) {
        k_bin_atoms1<<<mdvc_grid(n - done, 256, 32), 256, 0, st>>>(pos, done, n, P, occ_bits, pitch, vox_out, pos_out);
        MDVC_LAUNCH_CHECK();
    }
    return MDVC_OK;
}

// the REDs of one key list (the spill list of mdvc_bin_grid): count read from device memory
__global__ void __launch_bounds__(256)
k_bin_keys(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ count_ptr, uint32_t cap, uint32_t *__restrict__ occ) {
    const uint32_t n = min(*count_ptr, cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t k = __ldcs(keys + i);
        atomicOr(&occ[k >> 5], 1u << (k & 31));
    }
}

// ------------------------------------------------------------------------------------------
// A1  binning into a fresh grid, slab by slab in shared memory (the default path of mdvc_bin_grid)
// ------------------------------------------------------------------------------------------
// The bit grid is cut into slabs of 2^slab_shift bits (64 KiB of consecutive words at 1024^3: half an x-plane), small
// enough to be built in shared memory.  k_bin_sort drops every bead's key (= its bit index in the padded grid) into the
// key list of its slab; k_bin_fill, one CTA per slab, clears the slab in shared memory, ORs the listed beads in with
// shared-memory atomics and writes the finished words once with 16-byte stores.  No memset of the grid, no atomics on
// global memory outside the few that reserve list space, DRAM traffic = coordinates read once + grid written once
// (+ 4 B/bead of keys that stay in the L2).
// k_bin_sort works on tiles of BIN_TILE beads: keys and a per-slab count in shared memory (the shared atomicAdd also
// hands each key its rank inside the tile's share of the slab), an exclusive scan of the counts, ONE global atomicAdd per
// tile and non-empty slab to reserve a contiguous stretch of the slab's list, a counting sort of the keys inside shared
// memory, and a write-out in sorted order, so that neighbouring threads store neighbouring words.  A list that is full
// (very uneven systems: every list holds 3x the average) spills into a common list that k_bin_keys ORs into the finished
// grid afterwards.
#define BIN_TILE 8192                        // beads per tile
#define BIN_THREADS 512                      // 4 groups of four beads per thread and tile
#define BIN_MAX_SLABS 2048
#define BIN_KEY_NONE 0xffffffffu             // no bit index reaches this: the padded grid has fewer than 2^32 bits

struct BinSlabs {
    uint32_t *cursor;        // [BIN_MAX_SLABS] beads offered to each slab's list (may exceed cap), [BIN_MAX_SLABS] = spill count
    uint32_t *keys;          // [n_slabs][cap], then the spill list [n]
    uint32_t cap, n_slabs;
    int slab_shift;          // slab = key >> slab_shift
    uint32_t row_bits;       // pitch * 32: key = row * row_bits + z
};

static size_t bin_sort_smem() { return (2 * BIN_TILE + 2 * BIN_MAX_SLABS) * sizeof(uint32_t) + BIN_TILE * sizeof(unsigned short); }

// FAST (occupancy only, 32-bit box): the certified float32 floor decides almost every bead; the few it cannot decide
// (on or next to a voxel face) are queued in shared memory and redone in float64, one per thread -- densely, not as a
// divergent branch that a third of all warps would have to sit through -- before the counts are scanned.
template <bool FAST>
__global__ void __launch_bounds__(BIN_THREADS, FAST ? 2 : 1)
k_bin_sort(const float *__restrict__ pos, long long n, BinParams P, BinSlabs B, int tile_groups, int32_t *__restrict__ vox_out,
           float *pos_out) {
    extern __shared__ __align__(16) uint32_t s_dyn[];
    uint32_t *s_key = s_dyn;                                      // [BIN_TILE] key of bead (k, group) at k * BIN_TILE/4 + group
    uint32_t *s_sorted = s_key + BIN_TILE;                        // [BIN_TILE] the tile's keys grouped by slab
    uint32_t *s_cnt = s_sorted + BIN_TILE;                        // [BIN_MAX_SLABS] beads per slab, then list offset - s_off
    uint32_t *s_off = s_cnt + BIN_MAX_SLABS;                      // [BIN_MAX_SLABS] exclusive scan of the counts
    unsigned short *s_slot = reinterpret_cast<unsigned short *>(s_off + BIN_MAX_SLABS);   // [BIN_TILE] rank inside (tile, slab)
    unsigned short *s_slow = reinterpret_cast<unsigned short *>(s_sorted);                // undecided beads (dead before s_sorted is written)
    __shared__ uint32_t s_nslow, s_total, s_wsum[BIN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int TG = BIN_TILE / 4;                                  // at most this many groups of four beads (three float4) per tile
    const long long n4 = n >> 2;
    const long long n_tiles = (n4 + tile_groups - 1) / tile_groups;   // tile_groups <= TG: tiles sized so that every CTA gets the same number
    const uint32_t Y = (uint32_t)P.nbox32[4];
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        reinterpret_cast<uint4 *>(s_cnt)[tid] = make_uint4(0u, 0u, 0u, 0u);          // BIN_THREADS * 4 == BIN_MAX_SLABS
        if (tid == 0) s_nslow = 0;
        __syncthreads();
        // ---- keys, counts and ranks
#pragma unroll 2
        for (int it = 0; it < TG / BIN_THREADS; ++it) {
            const int gl = it * BIN_THREADS + tid;
            const long long g = tile * tile_groups + gl;
            uint32_t key[4] = {BIN_KEY_NONE, BIN_KEY_NONE, BIN_KEY_NONE, BIN_KEY_NONE};
            if (gl < tile_groups && g < n4) {
                const float4 *src = reinterpret_cast<const float4 *>(pos) + 3 * g;
                const float4 q0 = __ldcs(src), q1 = __ldcs(src + 1), q2 = __ldcs(src + 2);
                const float f[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
                if (FAST) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        int vk[3];
                        if (bin_floor_f32(f[3 * k], f[3 * k + 1], f[3 * k + 2], P, vk)) {
                            if (bin_wrap32(P, vk)) key[k] = ((uint32_t)vk[0] * Y + (uint32_t)vk[1]) * B.row_bits + (uint32_t)vk[2];
                        } else {
                            s_slow[atomicAdd(&s_nslow, 1u)] = (unsigned short)(gl * 4 + k);
                        }
                    }
                } else {
                    int32_t v[12];
                    float w[12];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        int vk[3];
                        float wk[3];
                        const bool inside = bin_one_atom(f[3 * k], f[3 * k + 1], f[3 * k + 2], P, pos_out != nullptr, vk, wk);
#pragma unroll
                        for (int j = 0; j < 3; ++j) { v[3 * k + j] = vk[j]; w[3 * k + j] = wk[j]; }
                        if (inside) key[k] = ((uint32_t)vk[0] * Y + (uint32_t)vk[1]) * B.row_bits + (uint32_t)vk[2];
                    }
                    if (vox_out) {
                        int4 *dst = reinterpret_cast<int4 *>(vox_out) + 3 * g;
                        dst[0] = make_int4(v[0], v[1], v[2], v[3]);
                        dst[1] = make_int4(v[4], v[5], v[6], v[7]);
                        dst[2] = make_int4(v[8], v[9], v[10], v[11]);
                    }
                    if (pos_out) {
                        float4 *dst = reinterpret_cast<float4 *>(pos_out) + 3 * g;
                        dst[0] = make_float4(w[0], w[1], w[2], w[3]);
                        dst[1] = make_float4(w[4], w[5], w[6], w[7]);
                        dst[2] = make_float4(w[8], w[9], w[10], w[11]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s_key[k * TG + gl] = key[k];
                if (key[k] != BIN_KEY_NONE) s_slot[k * TG + gl] = (unsigned short)atomicAdd(&s_cnt[key[k] >> B.slab_shift], 1u);
            }
        }
        if (FAST) {
            __syncthreads();
            for (uint32_t i = tid; i < s_nslow; i += BIN_THREADS) {
                const uint32_t nat = s_slow[i];                   // bead number inside the tile
                const long long atom = tile * tile_groups * 4 + nat;
                int vk[3];
                float wk[3];
                if (bin_one_atom(pos[3 * atom], pos[3 * atom + 1], pos[3 * atom + 2], P, false, vk, wk)) {
                    const uint32_t key = ((uint32_t)vk[0] * Y + (uint32_t)vk[1]) * B.row_bits + (uint32_t)vk[2];
                    const uint32_t at = (nat & 3u) * TG + (nat >> 2);
                    s_key[at] = key;
                    s_slot[at] = (unsigned short)atomicAdd(&s_cnt[key >> B.slab_shift], 1u);
                }
            }
        }
        __syncthreads();
        // ---- exclusive scan of the counts (four slabs per thread); list space reserved per (tile, slab)
        {
            const uint4 c = reinterpret_cast<const uint4 *>(s_cnt)[tid];
            uint4 b;                                              // reserved list offsets: issued first, the round trip
            b.x = c.x ? atomicAdd(&B.cursor[4 * tid], c.x) : 0u;  // to the L2 overlaps the scan below
            b.y = c.y ? atomicAdd(&B.cursor[4 * tid + 1], c.y) : 0u;
            b.z = c.z ? atomicAdd(&B.cursor[4 * tid + 2], c.z) : 0u;
            b.w = c.w ? atomicAdd(&B.cursor[4 * tid + 3], c.w) : 0u;
            const uint32_t mine = c.x + c.y + c.z + c.w;
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(MDVC_FULL, incl, d);
                if (lane >= d) incl += up;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t ws = lane < BIN_THREADS / 32 ? s_wsum[lane] : 0u;
                uint32_t wi = ws;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t up = __shfl_up_sync(MDVC_FULL, wi, d);
                    if (lane >= d) wi += up;
                }
                if (lane < BIN_THREADS / 32) s_wsum[lane] = wi - ws;
                if (lane == BIN_THREADS / 32 - 1) s_total = wi;
            }
            __syncthreads();
            uint4 o;
            o.x = s_wsum[warp] + incl - mine;
            o.y = o.x + c.x;
            o.z = o.y + c.y;
            o.w = o.z + c.z;
            reinterpret_cast<uint4 *>(s_off)[tid] = o;
            b.x -= o.x; b.y -= o.y; b.z -= o.z; b.w -= o.w;       // reserved list offset minus the tile-local offset
            reinterpret_cast<uint4 *>(s_cnt)[tid] = b;
        }
        __syncthreads();
        // ---- counting sort inside shared memory
#pragma unroll 4
        for (int i = tid; i < BIN_TILE; i += BIN_THREADS) {
            const uint32_t k = s_key[i];
            if (k != BIN_KEY_NONE) s_sorted[s_off[k >> B.slab_shift] + s_slot[i]] = k;
        }
        __syncthreads();
        // ---- write-out: position j of the sorted tile is entry j - s_off[slab] of this tile's stretch of the slab's list
        const uint32_t total = s_total;
#pragma unroll 4
        for (uint32_t j = tid; j < total; j += BIN_THREADS) {
            const uint32_t k = s_sorted[j], slab = k >> B.slab_shift;
            const uint32_t at = s_cnt[slab] + j;
            if (at < B.cap) B.keys[(size_t)slab * B.cap + at] = k;
            else B.keys[(size_t)B.n_slabs * B.cap + atomicAdd(&B.cursor[BIN_MAX_SLABS], 1u)] = k;   // list full: spill
        }
        __syncthreads();                                          // every array is reused by the next tile
    }
}

// One CTA per slab: clear it in shared memory, OR the listed beads in, write the words once.
__global__ void __launch_bounds__(BIN_THREADS)
k_bin_fill(BinSlabs B, uint32_t *__restrict__ occ, unsigned long long grid_words) {
    extern __shared__ __align__(16) uint32_t s_bits[];
    const uint32_t slab_words = 1u << (B.slab_shift - 5), mask = (1u << B.slab_shift) - 1u;
    const int tid = threadIdx.x;
    for (uint32_t slab = blockIdx.x; slab < B.n_slabs; slab += gridDim.x) {
        for (uint32_t w = tid; w < slab_words / 4; w += BIN_THREADS) reinterpret_cast<uint4 *>(s_bits)[w] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        const uint32_t cnt = min(__ldcg(B.cursor + slab), B.cap);
        const uint32_t *__restrict__ keys = B.keys + (size_t)slab * B.cap;
        uint32_t i = tid;
        for (; i + 3 * BIN_THREADS < cnt; i += 4 * BIN_THREADS) {          // four loads in flight per thread
            const uint32_t k0 = __ldcs(keys + i), k1 = __ldcs(keys + i + BIN_THREADS), k2 = __ldcs(keys + i + 2 * BIN_THREADS),
                           k3 = __ldcs(keys + i + 3 * BIN_THREADS);
            atomicOr(&s_bits[(k0 & mask) >> 5], 1u << (k0 & 31));
            atomicOr(&s_bits[(k1 & mask) >> 5], 1u << (k1 & 31));
            atomicOr(&s_bits[(k2 & mask) >> 5], 1u << (k2 & 31));
            atomicOr(&s_bits[(k3 & mask) >> 5], 1u << (k3 & 31));
        }
        for (; i < cnt; i += BIN_THREADS) {
            const uint32_t k = __ldcs(keys + i);
            atomicOr(&s_bits[(k & mask) >> 5], 1u << (k & 31));
        }
        __syncthreads();
        const unsigned long long w0 = (unsigned long long)slab * slab_words;
        const unsigned long long left = grid_words - w0;                   // > 0: n_slabs = ceil(grid_words / slab_words)
        const uint32_t nq = (uint32_t)((left < slab_words ? left : slab_words) / 4);   // pitch is a multiple of 4 words
        uint4 *dst = reinterpret_cast<uint4 *>(occ + w0);
        for (uint32_t w = tid; w < nq; w += BIN_THREADS) dst[w] = reinterpret_cast<const uint4 *>(s_bits)[w];
        __syncthreads();
    }
}

extern "C" size_t mdvc_bin_scratch_bytes(long long n) {
    if (n < 0) return 0;
    // counters, the slab lists (3x the average each + 256) and the spill list
    return (2 * BIN_MAX_SLABS + 3 * (size_t)n + 260 * (size_t)BIN_MAX_SLABS + (size_t)n) * sizeof(uint32_t);
}

extern "C" int mdvc_bin_grid(const float *pos, long long n, const double *T_host, const double *invT_host,
                             const long long *nbox_host, uint32_t *occ_bits, int pitch, int32_t *vox_out, float *pos_out,
                             void *scratch, size_t scratch_bytes, void *stream) {
    if (n < 0 || !T_host || !invT_host || !nbox_host || !occ_bits || (n > 0 && !pos)) return MDVC_ERR_BADARG;
    const long long X = nbox_host[0], Y = nbox_host[4], Z = nbox_host[8];
    if (X <= 0 || Y <= 0 || Z <= 0 || pitch < mdvc_words((int)Z) || X * Y * Z >= (1ll << 31)) return MDVC_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t plane_words = (size_t)Y * pitch, grid_bytes = (size_t)X * plane_words * sizeof(uint32_t);
    const unsigned long long grid_bits = (unsigned long long)X * plane_words * 32ull;
    const uintptr_t ptrs = (uintptr_t)pos | (uintptr_t)vox_out | (uintptr_t)pos_out;
    // small inputs, unaligned arrays, grids whose padded bit count does not fit 32 bits or a missing / short scratch:
    // clear everything and take the plain kernels
    bool plain = n < 4 * BIN_TILE || (ptrs & 15) != 0 || grid_bits >= (1ull << 32) || !scratch ||
                 scratch_bytes < mdvc_bin_scratch_bytes(n) || ((uintptr_t)scratch & 15) != 0;
    // slab size: 64 KiB of grid words (three CTAs of k_bin_fill per SM), 128 KiB beyond 1024^3 (more than BIN_MAX_SLABS slabs)
    int slab_shift = 19;
    while (((grid_bits + (1ull << slab_shift) - 1) >> slab_shift) > BIN_MAX_SLABS) ++slab_shift;
    if (slab_shift > 20 || ((uintptr_t)occ_bits & 15) != 0 || (pitch & 3) != 0) plain = true;
    // a grid that stays in the L2 takes the scattered REDs without DRAM traffic: 512^3 (16 MiB) 30 us plain against 56 us
    // here, 768^3 (54 MiB) 54 against 91 us, 1024^3 (128 MiB) 198 against 155 us (B200, profiles/r2_bin_slabs.txt)
    if (grid_bytes <= ((size_t)64 << 20)) plain = true;
    if (plain) {
        MDVC_CUDA_CHECK(cudaMemsetAsync(occ_bits, 0, grid_bytes, st));
        return mdvc_bin_atoms(pos, n, T_host, invT_host, nbox_host, occ_bits, pitch, vox_out, pos_out, stream);
    }
    BinParams P;
    bin_params(P, T_host, invT_host, nbox_host);
    // per device and cheap; not a stream operation, so it is fine inside a stream capture
    MDVC_CUDA_CHECK(cudaFuncSetAttribute(k_bin_sort<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_sort_smem()));
    MDVC_CUDA_CHECK(cudaFuncSetAttribute(k_bin_sort<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_sort_smem()));
    MDVC_CUDA_CHECK(cudaFuncSetAttribute(k_bin_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, 1 << 17));
    BinSlabs B;
    B.cursor = reinterpret_cast<uint32_t *>(scratch);
    B.keys = B.cursor + 2 * BIN_MAX_SLABS;
    B.slab_shift = slab_shift;
    B.n_slabs = (uint32_t)((grid_bits + (1ull << slab_shift) - 1) >> slab_shift);
    B.cap = (uint32_t)(((size_t)n * 3 / B.n_slabs) & ~(size_t)3) + 256u;
    B.row_bits = (uint32_t)pitch * 32u;
    MDVC_CUDA_CHECK(cudaMemsetAsync(B.cursor, 0, (BIN_MAX_SLABS + 1) * sizeof(uint32_t), st));
    const bool fast = P.fits32 && !vox_out && !pos_out;
    // every resident CTA gets the same number of tiles: the tile size shrinks below BIN_TILE instead of leaving a last,
    // mostly idle round (10.2 M beads: 296 CTAs x 5 tiles of 6892 beads rather than 4.2 rounds of 8192)
    const long long n4 = n >> 2, slots = (long long)mdvc_num_sms() * (fast ? 2 : 1);
    const long long min_tiles = (n4 + BIN_TILE / 4 - 1) / (BIN_TILE / 4);
    const long long sort_grid = std::min(min_tiles, slots);
    const long long rounds = (min_tiles + sort_grid - 1) / sort_grid;
    const int tile_groups = (int)((n4 + rounds * sort_grid - 1) / (rounds * sort_grid));
    if (fast) k_bin_sort<true><<<(unsigned)sort_grid, BIN_THREADS, bin_sort_smem(), st>>>(pos, n, P, B, tile_groups, vox_out, pos_out);
    else k_bin_sort<false><<<(unsigned)sort_grid, BIN_THREADS, bin_sort_smem(), st>>>(pos, n, P, B, tile_groups, vox_out, pos_out);
    MDVC_LAUNCH_CHECK();
    k_bin_fill<<<B.n_slabs, BIN_THREADS, (size_t)1 << (slab_shift - 3), st>>>(B, occ_bits, (unsigned long long)X * plane_words);
    MDVC_LAUNCH_CHECK();
    // beads that found their slab's list full (very uneven systems) and the last n % 4 beads
    k_bin_keys<<<mdvc_grid(n, 256, 4), 256, 0, st>>>(B.keys + (size_t)B.n_slabs * B.cap, B.cursor + BIN_MAX_SLABS, (uint32_t)n, occ_bits);
    MDVC_LAUNCH_CHECK();
    if (n & 3) {
        k_bin_atoms1<<<1, 256, 0, st>>>(pos, n & ~3ll, n, P, occ_bits, pitch, vox_out, pos_out);
        MDVC_LAUNCH_CHECK();
    }
    return MDVC_OK;
}

// ------------------------------------------------------------------------------------------
// A3/A4  morphology (periodic 3x3x3), one output word per thread
// ------------------------------------------------------------------------------------------
// z-dilation of word w of one row, periodic in z.  invert: operate on the complement (erosion).
__device__ __forceinline__ uint32_t dilate_z_word(const uint32_t *__restrict__ row, int w, int nw, int Z, bool invert) {
    uint32_t flip = invert ? 0xffffffffu : 0u;
    int wp = (w > 0) ? w - 1 : nw - 1;
    int wn = (w < nw - 1) ? w + 1 : 0;
    uint32_t cur = (row[w] ^ flip) & mdvc_valid_mask(w, Z);
    uint32_t prv = (row[wp] ^ flip) & mdvc_valid_mask(wp, Z);
    uint32_t nxt = (row[wn] ^ flip) & mdvc_valid_mask(wn, Z);
    int last = (Z - 1) & 31;                              // bit index of z = Z-1 inside word nw-1
    uint32_t down_in = (w > 0) ? (prv >> 31) : ((prv >> last) & 1u);     // value at z-1 for bit 0
    uint32_t up_in = nxt & 1u;                                            // value at z+1 for the top bit
    int top = (w == nw - 1) ? last : 31;
    uint32_t d = cur | (cur << 1) | down_in | (cur >> 1) | (up_in << top);
    return d & mdvc_valid_mask(w, Z);
}

__global__ void __launch_bounds__(256)
k_morph_step(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int X, int Y, int Z, int pitch, int erode) {
    int nw = mdvc_words(Z);
    long long items = (long long)X * Y * pitch;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += stride) {
        int w = (int)(i % pitch);
        long long r = i / pitch;
        if (w >= nw) { out[i] = 0; continue; }
        int y = (int)(r % Y), x = (int)(r / Y);
        uint32_t acc = 0;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            int xx = x + dx; xx = xx < 0 ? X - 1 : (xx >= X ? 0 : xx);
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                int yy = y + dy; yy = yy < 0 ? Y - 1 : (yy >= Y ? 0 : yy);
                acc |= dilate_z_word(in + ((long long)xx * Y + yy) * pitch, w, nw, Z, erode != 0);
            }
        }
        out[i] = erode ? (~acc & mdvc_valid_mask(w, Z)) : acc;
    }
}

// Tiled sweep: a CTA owns a tile of TY rows (plus one halo row on each side) x all words of a row and
// walks along x.  Per plane every thread loads one word (coalesced), parks it in shared memory and
// forms the y/z-dilated word P[x] from its 3x3 word neighbourhood; a 3-deep register window over x
// finishes the 3x3x3 OR.  Each input word is read (TY+2)/TY * (TX+2)/TX times, each output written once.
#define MORPH_MAX_THREADS 512
template <bool ERODE>
__global__ void __launch_bounds__(MORPH_MAX_THREADS)
k_morph_sweep(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int X, int Y, int Z, int pitch, int nw,
              int JT, int TX) {
    __shared__ uint32_t raw[2][MORPH_MAX_THREADS];
    const int t = threadIdx.x;
    const int j = t / nw, w = t - j * nw;
    const bool active = j < JT;
    const int TY = JT - 2;
    const int y0 = blockIdx.x * TY;
    int yy = y0 - 1 + j;
    yy %= Y; if (yy < 0) yy += Y;
    const int yout = y0 + j - 1;
    const bool is_out = active && j >= 1 && j <= TY && yout < Y;
    const int xs = blockIdx.y * TX, xe = min(xs + TX, X);
    const uint32_t vmask = mdvc_valid_mask(w, Z);
    const int last = (Z - 1) & 31;
    const int wp = w > 0 ? w - 1 : nw - 1, wn = w < nw - 1 ? w + 1 : 0;
    const int top = (w == nw - 1) ? last : 31;
    auto load = [&](int x) -> uint32_t {
        int xw = x < 0 ? x + X : (x >= X ? x - X : x);
        uint32_t v = active ? in[((size_t)xw * Y + yy) * pitch + w] : 0u;
        if (ERODE) v = ~v;
        return v & vmask;
    };
    uint32_t Pm2 = 0, Pm1 = 0;
    // two planes in flight per thread (deeper prefetch measured slower: the sweep is issue-bound, not latency-bound)
    constexpr int DEPTH = 2;
    uint32_t q[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) q[d] = (xs - 1 + d <= xe) ? load(xs - 1 + d) : 0u;
    int s = 0;
    for (int x0 = xs - 1; x0 <= xe; x0 += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const int x = x0 + d;
            if (x > xe) break;
            const uint32_t v = q[d];
            q[d] = (x + DEPTH <= xe) ? load(x + DEPTH) : 0u;
            uint32_t *buf = raw[s & 1];
            if (active) buf[t] = v;
            __syncthreads();
            uint32_t P = 0;
            if (is_out) {
                const uint32_t *r0 = buf + (j - 1) * nw, *r1 = r0 + nw, *r2 = r1 + nw;
                uint32_t c = r0[w] | r1[w] | r2[w];
                uint32_t pw = r0[wp] | r1[wp] | r2[wp];
                uint32_t nx = r0[wn] | r1[wn] | r2[wn];
                uint32_t down_in = (w > 0) ? (pw >> 31) : ((pw >> last) & 1u);
                P = (c | (c << 1) | down_in | (c >> 1) | ((nx & 1u) << top)) & vmask;
                if (s >= 2) {
                    uint32_t o = Pm2 | Pm1 | P;
                    if (ERODE) o = ~o & vmask;
                    size_t at = ((size_t)(x - 1) * Y + yout) * pitch;
                    out[at + w] = o;
                    if (w == nw - 1) for (int k = nw; k < pitch; ++k) out[at + k] = 0u;
                }
            }
            Pm2 = Pm1; Pm1 = P;
            ++s;
        }
    }
}

// Same sweep for rows of 1, 2, 4, 8, 16 or 32 words (Z <= 1024, power-of-two word count): the lanes of a
// row are neighbours in a warp segment, so the z-dilation takes its two neighbour words by shuffle (the
// segment-relative source index wraps, which is exactly the periodic z boundary) and only the z-dilated word
// goes through shared memory; the y-dilation then needs two LDS instead of nine.
template <bool ERODE>
__global__ void __launch_bounds__(MORPH_MAX_THREADS)
k_morph_sweep_p2(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int X, int Y, int Z, int pitch, int nw,
                 int JT, int TX) {
    __shared__ uint32_t dzs[2][MORPH_MAX_THREADS];
    const int t = threadIdx.x;
    const int w = t & (nw - 1), j = t / nw;
    const bool active = j < JT;
    const int TY = JT - 2;
    const int y0 = blockIdx.x * TY;
    int yy = y0 - 1 + j;
    yy %= Y; if (yy < 0) yy += Y;
    const int yout = y0 + j - 1;
    const bool is_out = active && j >= 1 && j <= TY && yout < Y;
    const int xs = blockIdx.y * TX, xe = min(xs + TX, X);
    const uint32_t vmask = mdvc_valid_mask(w, Z);
    const int last = (Z - 1) & 31;
    const int top = (w == nw - 1) ? last : 31;
    const size_t plane = (size_t)Y * pitch;
    const uint32_t *src = in + (size_t)yy * pitch + w;
    uint32_t *dst = out + (size_t)yout * pitch + w;
    uint32_t Pm2 = 0, Pm1 = 0;
    int xw = xs - 1 < 0 ? X - 1 : xs - 1;
    uint32_t v = active ? src[(size_t)xw * plane] : 0u;
    int s = 0;
    for (int x = xs - 1; x <= xe; ++x, ++s) {
        int xn = x + 1 >= X ? x + 1 - X : x + 1;
        uint32_t vnext = (x < xe && active) ? src[(size_t)xn * plane] : 0u;       // prefetch the next plane
        if (ERODE) v = ~v;
        v &= vmask;
        const uint32_t pv = __shfl_sync(MDVC_FULL, v, w - 1, nw);
        const uint32_t nv = __shfl_sync(MDVC_FULL, v, w + 1, nw);
        const uint32_t down_in = (w > 0) ? (pv >> 31) : ((pv >> last) & 1u);
        const uint32_t dz = (v | (v << 1) | down_in | (v >> 1) | ((nv & 1u) << top)) & vmask;
        uint32_t *buf = dzs[s & 1];
        buf[t] = dz;
        __syncthreads();
        uint32_t P = 0;
        if (is_out) {
            P = buf[t - nw] | dz | buf[t + nw];
            if (s >= 2) {
                uint32_t o = Pm2 | Pm1 | P;
                if (ERODE) o = ~o & vmask;
                uint32_t *op = dst + (size_t)(x - 1) * plane;
                *op = o;
                if (w == nw - 1) for (int k = 1; k < pitch - nw + 1; ++k) op[k] = 0u;
            }
        }
        Pm2 = Pm1; Pm1 = P;
        v = vnext;
    }
}

// ------------------------------------------------------------------------------------------
// Fused closing ('d' then 'e' in one pass) for rows of up to 32 words, planes staged by TMA.
//
// A CTA owns JT consecutive rows (two halo rows on each side) and walks along x.  One elected thread streams
// whole plane tiles (JT rows x pitch words, contiguous in memory apart from the periodic wrap) into an S-deep
// shared-memory ring with cp.async.bulk, completion on an mbarrier per stage.
// A thread owns four consecutive words (128 voxels) of one row: the y-OR of the three raw rows comes straight
// from the ring (3 LDS.128), the z-dilation is funnel shifts inside the quad plus ONE packed edge-bit exchange
// with the two neighbour lanes (2 shuffles per 128 voxels), the x-dilation a 3-deep register window -> dilated
// plane x-1.  Its complement goes through a double-buffered exchange tile (one __syncthreads per plane) for the
// same y/z/x steps of the erosion -> closed plane x-2, stored as one 16-byte word per thread.  Every input word
// is read once (+ halo) and every output word written once; about 25 instructions per word instead of 110 for
// the word-per-thread version (ncu: that one was issue-bound at 78 % issue-active).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 or4(uint4 a, uint4 b) { return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w); }
__device__ __forceinline__ uint4 and4(uint4 a, uint4 b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }
__device__ __forceinline__ uint4 not4(uint4 a) { return make_uint4(~a.x, ~a.y, ~a.z, ~a.w); }

// Two adjacent rows (j0 = 2*jr, j0+1) per thread: the raw rows j0-1 .. j0+2 serve both (4 LDS.128 instead of 6), the
// exchange tile is read for j0-1 and j0+2 only, the edge bits of both rows travel in one pair of shuffles, and the
// per-plane overhead (mbarrier wait, barrier, pointers) is paid once per two rows.
// E1 / E2: the first / second operation is an erosion (dilation of the complement); 'de' is <false, true>.
// ONE: a single operation (E1; E2 must be false): the first half of the pipeline only -- no exchange tile, one
// z-dilation per plane -- for the odd operation of a morph string ('d', 'e', the last of 'dde'), which used to take
// the word-per-thread kernel (114 us at 1024^3 against 69 us for two fused operations).
template <int S, bool FULLZ, bool E1, bool E2, bool ONE>
__global__ void __launch_bounds__(256)
k_close_sweep2(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int X, int Y, int Z, int pitch, int nw,
               int lq_log2, int JR, int TX, uint32_t *__restrict__ row_counts) {
    extern __shared__ __align__(128) unsigned char close_smem[];
    const int LQ = 1 << lq_log2;
    const int nq = pitch >> 2;
    const int JT = 2 * JR, TY = JT - 4;
    const int tile4 = JT * nq;
    uint4 *raw = (uint4 *)close_smem;                             // [S][JT][nq]
    uint4 *cmp = raw + S * tile4;                                 // [2][JT][nq]
    uint64_t *full = (uint64_t *)(cmp + (ONE ? 0 : 2) * tile4);   // [S]  (a single operation has no exchange tile)
    const int t = threadIdx.x;
    const int jr = t >> lq_log2, l = t & (LQ - 1);
    const int j0 = 2 * jr;
    const bool lane_ok = jr < JR && l < nq;
    const bool has_up = lane_ok && jr > 0, has_dn = lane_ok && jr < JR - 1;
    const int y0 = blockIdx.x * TY;
    const int xs = blockIdx.y * TX, xe = min(xs + TX, X);
    const int n_in = xe - xs + 4;
    const int last = (Z - 1) & 31;
    const int klast = nw - 1 - 4 * l;
    const int ktail = (klast >= 0 && klast < 3) ? klast : 3;
    const int ush = (klast >= 0 && klast <= 3) ? last : 31;
    uint4 vm = make_uint4(0u, 0u, 0u, 0u);
    if (lane_ok) vm = make_uint4(mdvc_valid_mask(4 * l, Z), mdvc_valid_mask(4 * l + 1, Z), mdvc_valid_mask(4 * l + 2, Z),
                                 mdvc_valid_mask(4 * l + 3, Z));
    const int pl = l > 0 ? l - 1 : nq - 1, nl = l < nq - 1 ? l + 1 : 0;
    const size_t plane = (size_t)Y * pitch;

    auto issue = [&](int s, int stage) {
        int x = xs - 2 + s;
        x %= X; if (x < 0) x += X;
        uint64_t *bar = full + stage;
        uint32_t *dst = (uint32_t *)(raw + stage * tile4);
        const uint32_t *pln = in + (size_t)x * plane;
        mbar_expect_tx(bar, (uint32_t)tile4 * 16u);
        int yy = (y0 - 2) % Y; if (yy < 0) yy += Y;
        for (int r = 0; r < JT;) {
            int len = min(JT - r, Y - yy);
            bulk_g2s(dst + r * pitch, pln + (size_t)yy * pitch, (uint32_t)(len * pitch) * 4u, bar);
            r += len; yy = 0;
        }
    };
    auto edges = [&](uint4 v) -> uint32_t {                      // bit 0: top bit of the quad's last word, bit 1: its first bit
        uint32_t hi;
        if (FULLZ) hi = v.w >> 31;
        else {
            const uint32_t tw = ktail == 0 ? v.x : (ktail == 1 ? v.y : (ktail == 2 ? v.z : v.w));
            hi = (tw >> ush) & 1u;
        }
        return hi | ((v.x & 1u) << 1);
    };
    auto spread = [&](uint4 v, uint32_t down, uint32_t up) -> uint4 {      // v | v<<1 | v>>1 along the periodic row
        uint4 r;
        r.x = v.x | ((v.x << 1) | down) | __funnelshift_r(v.x, v.y, 1);
        r.y = v.y | __funnelshift_l(v.x, v.y, 1) | __funnelshift_r(v.y, v.z, 1);
        r.z = v.z | __funnelshift_l(v.y, v.z, 1) | __funnelshift_r(v.z, v.w, 1);
        r.w = v.w | __funnelshift_l(v.z, v.w, 1) | (v.w >> 1);
        if (FULLZ) r.w |= up << 31;
        else {
            const uint32_t ub = up << ush;
            r.x |= ktail == 0 ? ub : 0u; r.y |= ktail == 1 ? ub : 0u; r.z |= ktail == 2 ? ub : 0u; r.w |= ktail == 3 ? ub : 0u;
            r = and4(r, vm);
        }
        return r;
    };
    auto zdil2 = [&](uint4 &a, uint4 &b) {                       // periodic z-dilation of two rows, one edge exchange
        const uint32_t e = edges(a) | (edges(b) << 2);
        const uint32_t pe = __shfl_sync(MDVC_FULL, e, pl, LQ);
        const uint32_t ne = __shfl_sync(MDVC_FULL, e, nl, LQ);
        a = spread(a, pe & 1u, (ne >> 1) & 1u);
        b = spread(b, (pe >> 2) & 1u, (ne >> 3) & 1u);
    };

    if (t == 0) {
        for (int i = 0; i < S; ++i) mbar_init(full + i, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (t == 0) for (int s = 0; s < S && s < n_in; ++s) issue(s, s);

    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    uint4 P0m1 = zero, P0m2 = zero, P1m1 = zero, P1m2 = zero, Q0m1 = zero, Q0m2 = zero, Q1m1 = zero, Q1m2 = zero;
    const int yout0 = y0 + j0 - 2;
    const bool out0 = lane_ok && j0 >= 2 && j0 <= JT - 3 && yout0 < Y;
    const bool out1 = lane_ok && j0 + 1 >= 2 && j0 + 1 <= JT - 3 && yout0 + 1 < Y;
    uint4 *op = (uint4 *)(out + (size_t)xs * plane + (size_t)((out0 || out1) ? yout0 : 0) * pitch) + l;
    const size_t plane4 = plane >> 2;
    const int me = j0 * nq + l;
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 0; s < n_in; ++s) {
        mbar_wait(full + stage, phase);
        const uint4 *rw = raw + stage * tile4 + me;
        uint4 a = zero, b = zero, c = zero, d = zero;
        if (lane_ok) { b = rw[0]; c = rw[nq]; }
        if (has_up) a = rw[-nq];
        if (has_dn) d = rw[2 * nq];
        uint4 P0, P1;
        if (E1) {                                                 // dilate the complement: ~a | ~b | ~c = ~(a & b & c)
            const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
            if (!has_up) a = ones;
            if (!has_dn) d = ones;
            const uint4 bc = and4(b, c);
            P0 = and4(not4(and4(a, bc)), vm); P1 = and4(not4(and4(bc, d)), vm);
        } else {
            const uint4 bc = or4(b, c);
            P0 = or4(a, bc); P1 = or4(bc, d);
            if (!FULLZ) { P0 = and4(P0, vm); P1 = and4(P1, vm); }
        }
        zdil2(P0, P1);
        // the dilated plane x-1 of the first operation, handed to the second one complemented iff exactly one of
        // the two operations is an erosion (erosion = complement, dilate, complement)
        uint4 n0 = or4(or4(P0m2, P0m1), P0), n1 = or4(or4(P1m2, P1m1), P1);
        if (E1 != E2) { n0 = and4(not4(n0), vm); n1 = and4(not4(n1), vm); }
        P0m2 = P0m1; P0m1 = P0; P1m2 = P1m1; P1m1 = P1;
        if (ONE) {
            __syncthreads();                                      // every thread is done with this stage of the ring
            if (t == 0 && s + S < n_in) issue(s + S, stage);
            if (s >= 3 && s < n_in - 1) {                         // n0 / n1 = finished plane xs + s - 3
                if (out0) op[0] = n0;
                if (out1) op[nq] = n1;
                op += plane4;
                if (row_counts) {
                    const QuadStarts s0 = mdvc_quad_starts(n0, vm, l, LQ), s1 = mdvc_quad_starts(n1, vm, l, LQ);
                    uint32_t c0 = __popc(s0.s0) + __popc(s0.s1) + __popc(s0.s2) + __popc(s0.s3);
                    uint32_t c1 = __popc(s1.s0) + __popc(s1.s1) + __popc(s1.s2) + __popc(s1.s3);
                    uint32_t cc = c0 | (c1 << 16);
                    for (int o = LQ >> 1; o > 0; o >>= 1) cc += __shfl_xor_sync(MDVC_FULL, cc, o, LQ);
                    const size_t row0 = (size_t)(xs + s - 3) * Y + yout0;
                    if (out0 && l == 0) row_counts[row0] = cc & 0xffffu;
                    if (out1 && l == 0) row_counts[row0 + 1] = cc >> 16;
                }
            }
            if (++stage == S) { stage = 0; phase ^= 1u; }
            continue;
        }
        uint4 *ex = cmp + (s & 1) * tile4 + me;
        if (lane_ok) { ex[0] = n0; ex[nq] = n1; }
        __syncthreads();
        if (t == 0 && s + S < n_in) issue(s + S, stage);
        const uint4 nn = or4(n0, n1);
        uint4 Q0 = nn, Q1 = nn;
        if (has_up) Q0 = or4(Q0, ex[-nq]);
        if (has_dn) Q1 = or4(Q1, ex[2 * nq]);
        zdil2(Q0, Q1);
        if (s >= 4) {
            uint4 o0 = or4(or4(Q0m2, Q0m1), Q0), o1 = or4(or4(Q1m2, Q1m1), Q1);
            if (E2) { o0 = and4(not4(o0), vm); o1 = and4(not4(o1), vm); }
            if (out0) op[0] = o0;
            if (out1) op[nq] = o1;
            op += plane4;
            if (row_counts) {
                // stage 1a of the labelling for free: z-runs per finished row (every lane takes part in the shuffles)
                const QuadStarts s0 = mdvc_quad_starts(o0, vm, l, LQ), s1 = mdvc_quad_starts(o1, vm, l, LQ);
                uint32_t c0 = __popc(s0.s0) + __popc(s0.s1) + __popc(s0.s2) + __popc(s0.s3);
                uint32_t c1 = __popc(s1.s0) + __popc(s1.s1) + __popc(s1.s2) + __popc(s1.s3);
                uint32_t cc = c0 | (c1 << 16);                    // a row has at most 1024 runs here
                for (int o = LQ >> 1; o > 0; o >>= 1) cc += __shfl_xor_sync(MDVC_FULL, cc, o, LQ);
                const size_t row0 = (size_t)(xs + s - 4) * Y + yout0;
                if (out0 && l == 0) row_counts[row0] = cc & 0xffffu;
                if (out1 && l == 0) row_counts[row0 + 1] = cc >> 16;
            }
        }
        Q0m2 = Q0m1; Q0m1 = Q0; Q1m2 = Q1m1; Q1m1 = Q1;
        if (++stage == S) { stage = 0; phase ^= 1u; }
    }
}

// Two consecutive operations (e1 then e2) -- or, with `one`, the single operation e1 -- on rows of at most 32 words.
// Returns MDVC_OK with *done = 1 when the TMA-staged kernel was launched, *done = 0 when the shape does not qualify
// (the caller then runs the word-per-thread steps).
static int launch_close(const uint32_t *src, uint32_t *dst, int X, int Y, int Z, int pitch, bool e1, bool e2, bool one,
                        cudaStream_t st, int *done, uint32_t *row_counts = nullptr, int *counted = nullptr) {
    *done = 0;
    constexpr int S = 3;
    int nw = mdvc_words(Z);
    if (nw > 32 || pitch != ((nw + 3) / 4) * 4 || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return MDVC_OK;
    int nq = pitch / 4, lq_log2 = 0;
    while ((1 << lq_log2) < nq) ++lq_log2;
    int LQ = 1 << lq_log2;
    int JT = 2 * (256 / LQ);
    if (JT > Y + 4) JT = Y + 4;                                   // one tile covers all of y
    size_t row_bytes = (size_t)(S + (one ? 0 : 2)) * pitch * 4;
    int fit = (int)((48 * 1024 - S * 8) / row_bytes);
    if (JT > fit) JT = fit;
    if (JT < 5) JT = 5;
    JT += JT & 1;                                                 // an even number of rows, two per thread
    int TY = JT - 4;
    int ny = (Y + TY - 1) / TY;
    int target = mdvc_num_sms() * 4;
    int xchunks = (target + ny - 1) / ny;
    if (xchunks > (X + 7) / 8) xchunks = (X + 7) / 8;
    if (xchunks < 1) xchunks = 1;
    if (xchunks > 65535) xchunks = 65535;
    int TX = (X + xchunks - 1) / xchunks;
    xchunks = (X + TX - 1) / TX;
    size_t smem = (size_t)JT * row_bytes + S * 8;
    dim3 grid(ny, xchunks);
    int threads = (((JT / 2) * LQ + 31) / 32) * 32;
#define MDVC_CLOSE2(FZ, A, B) k_close_sweep2<S, FZ, A, B, false><<<grid, threads, smem, st>>>(src, dst, X, Y, Z, pitch, nw, lq_log2, JT / 2, TX, row_counts)
#define MDVC_CLOSE1(FZ, A) k_close_sweep2<S, FZ, A, false, true><<<grid, threads, smem, st>>>(src, dst, X, Y, Z, pitch, nw, lq_log2, JT / 2, TX, row_counts)
    const int variant = one ? 8 + (Z % 128 == 0 ? 2 : 0) + (e1 ? 1 : 0) : (Z % 128 == 0 ? 4 : 0) | (e1 ? 2 : 0) | (e2 ? 1 : 0);
    switch (variant) {
        case 8: MDVC_CLOSE1(false, false); break;
        case 9: MDVC_CLOSE1(false, true); break;
        case 10: MDVC_CLOSE1(true, false); break;
        case 11: MDVC_CLOSE1(true, true); break;
        case 0: MDVC_CLOSE2(false, false, false); break;
        case 1: MDVC_CLOSE2(false, false, true); break;
        case 2: MDVC_CLOSE2(false, true, false); break;
        case 3: MDVC_CLOSE2(false, true, true); break;
        case 4: MDVC_CLOSE2(true, false, false); break;
        case 5: MDVC_CLOSE2(true, false, true); break;
        case 6: MDVC_CLOSE2(true, true, false); break;
        default: MDVC_CLOSE2(true, true, true); break;
    }
#undef MDVC_CLOSE2
#undef MDVC_CLOSE1
    if (counted) *counted = row_counts != nullptr;
    MDVC_LAUNCH_CHECK();
    *done = 1;
    return MDVC_OK;
}

static int launch_morph(const uint32_t *src, uint32_t *dst, int X, int Y, int Z, int pitch, bool erode, cudaStream_t st) {
    int nw = mdvc_words(Z);
    if (nw * 3 > MORPH_MAX_THREADS) {      // very long rows: the simple kernel
        long long items = (long long)X * Y * pitch;
        k_morph_step<<<mdvc_grid(items, 256), 256, 0, st>>>(src, dst, X, Y, Z, pitch, erode);
        MDVC_LAUNCH_CHECK();
        return MDVC_OK;
    }
    int JT = MORPH_MAX_THREADS / nw;
    if (JT > Y + 2) JT = Y + 2;
    if (JT < 3) JT = 3;
    int TY = JT - 2;
    int ny = (Y + TY - 1) / TY;
    int target = mdvc_num_sms() * 8;
    int xchunks = (target + ny - 1) / ny;
    if (xchunks > (X + 3) / 4) xchunks = (X + 3) / 4;
    if (xchunks < 1) xchunks = 1;
    if (xchunks > 65535) xchunks = 65535;
    int TX = (X + xchunks - 1) / xchunks;
    xchunks = (X + TX - 1) / TX;
    int threads = ((JT * nw + 31) / 32) * 32;
    dim3 grid(ny, xchunks);
    if ((nw & (nw - 1)) == 0 && nw <= 32) {
        if (erode) k_morph_sweep_p2<true><<<grid, threads, 0, st>>>(src, dst, X, Y, Z, pitch, nw, JT, TX);
        else k_morph_sweep_p2<false><<<grid, threads, 0, st>>>(src, dst, X, Y, Z, pitch, nw, JT, TX);
    } else if (erode) k_morph_sweep<true><<<grid, threads, 0, st>>>(src, dst, X, Y, Z, pitch, nw, JT, TX);
    else k_morph_sweep<false><<<grid, threads, 0, st>>>(src, dst, X, Y, Z, pitch, nw, JT, TX);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

static int morph_impl(const uint32_t *bits_in, uint32_t *bits_out, uint32_t *tmp, int X, int Y, int Z,
                      int pitch, const char *ops, uint32_t *row_counts, int *counted, void *stream) {
    if (counted) *counted = 0;
    if (!bits_in || !bits_out || !ops || bad_dims(X, Y, Z, pitch) || bits_in == bits_out) return MDVC_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    size_t n_ops = strlen(ops);
    for (size_t i = 0; i < n_ops; ++i)
        if (ops[i] != 'd' && ops[i] != 'e') { mdvc_set_error_msg("unknown morph operation"); return MDVC_ERR_BADARG; }
    size_t bytes = (size_t)X * Y * pitch * sizeof(uint32_t);
    if (n_ops == 0) {
        MDVC_CUDA_CHECK(cudaMemcpyAsync(bits_out, bits_in, bytes, cudaMemcpyDeviceToDevice, st));
        return MDVC_OK;
    }
    if (n_ops > 1 && (!tmp || tmp == bits_in || tmp == bits_out)) return MDVC_ERR_BADARG;
    // ping-pong so that the last step lands in bits_out
    // two consecutive operations (the closing 'de' above all) run as one fused launch where the shape allows it
    // (decided up front from the same conditions launch_close checks, so that the ping-pong plan below holds)
    const int nw = mdvc_words(Z);
    const uintptr_t all_ptrs = (uintptr_t)bits_in | (uintptr_t)bits_out | (uintptr_t)tmp;
    const bool can_fuse = nw <= 32 && pitch == ((nw + 3) / 4) * 4 && (all_ptrs & 15) == 0;
    auto pair_at = [&](size_t i) {
        return can_fuse && i + 1 < n_ops;
    };
    size_t n_launch = 0;
    for (size_t i = 0; i < n_ops; ++i, ++n_launch)
        if (pair_at(i)) ++i;
    const uint32_t *src = bits_in;
    size_t k = 0;
    for (size_t i = 0; i < n_ops; ++i, ++k) {
        bool to_out = ((n_launch - 1 - k) % 2) == 0;
        uint32_t *dst = to_out ? bits_out : tmp;
        if (!dst) return MDVC_ERR_BADARG;
        int fused = 0;
        if (pair_at(i)) {
            const bool last = i + 2 == n_ops;                     // only the final grid's rows are worth counting
            int rc = launch_close(src, dst, X, Y, Z, pitch, ops[i] == 'e', ops[i + 1] == 'e', false, st, &fused,
                                  last ? row_counts : nullptr, last ? counted : nullptr);
            if (rc) return rc;
        }
        if (fused) { ++i; src = dst; continue; }
        if (pair_at(i)) { mdvc_set_error_msg("fused morph pair declined after planning"); return MDVC_ERR_BADARG; }
        if (can_fuse) {                                           // the odd operation on the same TMA-staged layout
            const bool last = i + 1 == n_ops;
            int rc = launch_close(src, dst, X, Y, Z, pitch, ops[i] == 'e', false, true, st, &fused,
                                  last ? row_counts : nullptr, last ? counted : nullptr);
            if (rc) return rc;
            if (fused) { src = dst; continue; }
        }
        int rc = launch_morph(src, dst, X, Y, Z, pitch, ops[i] == 'e', st);
        if (rc) return rc;
        src = dst;
    }
    return MDVC_OK;
}

extern "C" int mdvc_morph(const uint32_t *bits_in, uint32_t *bits_out, uint32_t *tmp, int X, int Y, int Z,
                          int pitch, const char *ops, void *stream) {
    return morph_impl(bits_in, bits_out, tmp, X, Y, Z, pitch, ops, nullptr, nullptr, stream);
}

extern "C" int mdvc_morph_counted(const uint32_t *bits_in, uint32_t *bits_out, uint32_t *tmp, int X, int Y, int Z,
                                  int pitch, const char *ops, uint32_t *row_counts, int *counted, void *stream) {
    if (!counted) return MDVC_ERR_BADARG;
    return morph_impl(bits_in, bits_out, tmp, X, Y, Z, pitch, ops, row_counts, counted, stream);
}

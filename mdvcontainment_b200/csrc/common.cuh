// common.cuh -- shared device/host helpers for the mdvc kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/mdvc.h"

#define MDVC_FULL 0xffffffffu
#define MDVC_SM_COUNT_B200 148

// --- error plumbing ----------------------------------------------------------------------
void mdvc_set_error(const char *where, cudaError_t e);
void mdvc_set_error_msg(const char *msg);

#define MDVC_CUDA_CHECK(expr)                                   \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) {                                \
            mdvc_set_error(#expr, _e);                          \
            return MDVC_ERR_CUDA;                               \
        }                                                       \
    } while (0)

extern unsigned long long g_mdvc_launches;     // kernels launched by this library (host-side counter)
#define MDVC_LAUNCH_CHECK()                      \
    do {                                         \
        ++g_mdvc_launches;                       \
        MDVC_CUDA_CHECK(cudaGetLastError());     \
    } while (0)

static inline int mdvc_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = MDVC_SM_COUNT_B200;
    }
    return sms;
}

// grid for a grid-stride kernel: enough blocks for `work` items, capped at `waves` blocks per SM
static inline unsigned mdvc_grid(long long work, int block, int blocks_per_sm = 16) {
    long long need = (work + block - 1) / block;
    long long cap = (long long)mdvc_num_sms() * blocks_per_sm;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

// --- bit-grid helpers --------------------------------------------------------------------
__host__ __device__ __forceinline__ int mdvc_words(int Z) { return (Z + 31) >> 5; }

// mask of valid bits of word w in a row of Z voxels (0 for words past the row end)
__host__ __device__ __forceinline__ uint32_t mdvc_valid_mask(int w, int Z) {
    int nbits = Z - (w << 5);
    if (nbits >= 32) return 0xffffffffu;
    if (nbits <= 0) return 0u;
    return (1u << nbits) - 1u;
}

// --- packed (label,label,shift) keys -------------------------------------------------------
// 29-bit order-preserving label fields + 5-bit shift code: unsigned key order == the
// reference's signed tuple order (l1, l2, sx, sy, sz).  Valid for |label| < 2^28, which holds for
// every grid with < 2^31 voxels (at most one component per polarity per 2x2x2 block).
#define MDVC_LABEL_BIAS (1 << 28)
#define MDVC_KEY_EMPTY 0xffffffffffffffffull

__host__ __device__ __forceinline__ unsigned long long mdvc_pack_key(int l1, int l2, int code) {
    return ((unsigned long long)(uint32_t)(l1 + MDVC_LABEL_BIAS) << 34) |
           ((unsigned long long)(uint32_t)(l2 + MDVC_LABEL_BIAS) << 5) | (unsigned long long)code;
}
__host__ __device__ __forceinline__ void mdvc_unpack_key(unsigned long long k, int &l1, int &l2, int &code) {
    l1 = (int)((k >> 34) & 0x1fffffffu) - MDVC_LABEL_BIAS;
    l2 = (int)((k >> 5) & 0x1fffffffu) - MDVC_LABEL_BIAS;
    code = (int)(k & 31u);
}
__host__ __device__ __forceinline__ int mdvc_shift_code(int sx, int sy, int sz) {
    return (sx + 1) * 9 + (sy + 1) * 3 + (sz + 1);
}

#ifdef __CUDACC__
// --- hash set of 64-bit keys (open addressing, linear probing) -----------------------------
__device__ __forceinline__ unsigned long long mdvc_mix64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

// Returns false when no free slot turns up within MDVC_SET_MAX_PROBES probes: the table is full or so crowded that
// it has to grow anyway (the caller raises an overflow flag and the host retries with more slots).  The bound matters:
// probing a FULL table to the end costs n_slots loads per key, and a many-label grid pushed a million keys through
// undersized tables that way (15 s for one 1024^3 noise grid before the bound, milliseconds after).
#define MDVC_SET_MAX_PROBES 128u
__device__ __forceinline__ bool mdvc_set_insert(unsigned long long *slots, uint32_t n_slots,
                                                unsigned long long key) {
    uint32_t mask = n_slots - 1;
    uint32_t h = (uint32_t)mdvc_mix64(key) & mask;
    const uint32_t limit = n_slots < MDVC_SET_MAX_PROBES ? n_slots : MDVC_SET_MAX_PROBES;
    for (uint32_t probe = 0; probe < limit; ++probe) {
        unsigned long long cur = __ldcg(&slots[h]);
        if (cur == key) return true;
        if (cur == MDVC_KEY_EMPTY) {
            unsigned long long old = atomicCAS(&slots[h], MDVC_KEY_EMPTY, key);
            if (old == MDVC_KEY_EMPTY || old == key) return true;
        }
        h = (h + 1) & mask;
    }
    return false;
}

// --- union-find with min-index roots (parents only ever decrease) ---------------------------
__device__ __forceinline__ uint32_t mdvc_uf_find(uint32_t *parent, uint32_t v) {
    uint32_t cur = __ldcg(&parent[v]);
    if (cur != v) {
        uint32_t prev = v, next;
        while (cur > (next = __ldcg(&parent[cur]))) {   // intermediate pointer jumping
            parent[prev] = next;
            prev = cur;
            cur = next;
        }
    }
    return cur;
}

__device__ __forceinline__ void mdvc_uf_union(uint32_t *parent, uint32_t a, uint32_t b) {
    a = mdvc_uf_find(parent, a);
    b = mdvc_uf_find(parent, b);
    while (a != b) {
        if (a < b) { uint32_t t = a; a = b; b = t; }     // hook the larger root under the smaller
        uint32_t old = atomicCAS(&parent[a], a, b);
        if (old == a) break;
        a = old;                                            // a was hooked meanwhile: climb and retry
    }
}
#endif  // __CUDACC__

// --- run starts of a row held as 16-byte quads (4 words per lane, `width` lanes per row) ------------------
struct QuadStarts { uint32_t s0, s1, s2, s3; };

// Start mask of the quad: bit k of s_m set <=> a z-run starts at voxel 32*(4*ql+m)+k.  B must be masked to the
// valid voxels; vm is the quad's valid mask; every lane of the warp has to call this (shuffle inside).
__device__ __forceinline__ QuadStarts mdvc_quad_starts(uint4 B, uint4 vm, int ql, int width) {
    uint32_t prev = __shfl_up_sync(MDVC_FULL, B.w, 1, width) >> 31;
    if (ql == 0) prev = 0;
    QuadStarts q;
    q.s0 = B.x ^ ((B.x << 1) | prev);
    if (ql == 0) q.s0 |= 1u;
    q.s0 &= vm.x;
    q.s1 = (B.y ^ __funnelshift_l(B.x, B.y, 1)) & vm.y;
    q.s2 = (B.z ^ __funnelshift_l(B.y, B.z, 1)) & vm.z;
    q.s3 = (B.w ^ __funnelshift_l(B.z, B.w, 1)) & vm.w;
    return q;
}

// scan.cuh -- exclusive prefix sum over uint32 / uint64 arrays whose length may live on the device.
//
// Three launches (tile sums -> scan of tile sums -> tile scan with offset).  The arrays scanned on
// the hot path are per-row run counts (X*Y entries) and per-run root flags; both are tiny next to
// the voxel grids, so the simple scheme is enough.
#pragma once
#include "common.cuh"

#define MDVC_SCAN_THREADS 256
#define MDVC_SCAN_ITEMS 8
#define MDVC_SCAN_TILE (MDVC_SCAN_THREADS * MDVC_SCAN_ITEMS)

template <typename T>
__device__ __forceinline__ T mdvc_block_reduce_sum(T v, T *smem /* >= 32 */) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(MDVC_FULL, v, o);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? smem[threadIdx.x] : T(0);
    if (wid == 0)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(MDVC_FULL, v, o);
    return v;   // valid in thread 0
}

// exclusive scan of one value per thread across the block; returns the exclusive prefix, *total = block sum
template <typename T>
__device__ __forceinline__ T mdvc_block_exscan(T v, T *smem /* >= 33 */, T *total) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(MDVC_FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        T w = (lane < nw) ? smem[lane] : T(0);
        T winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(MDVC_FULL, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;           // exclusive warp offsets
        if (lane == 31) smem[32] = winc; // block total
    }
    __syncthreads();
    T res = smem[wid] + inc - v;
    *total = smem[32];
    __syncthreads();
    return res;
}

template <typename T>
__global__ void __launch_bounds__(MDVC_SCAN_THREADS)
k_scan_tile_sums(const T *__restrict__ in, const uint32_t *n_dev, uint32_t n_host, T *__restrict__ tile_sums) {
    __shared__ T sm[33];
    uint32_t n = n_dev ? *n_dev : n_host;
    if (n_dev && n > n_host) n = n_host;                 // never past the capacity
    size_t base = (size_t)blockIdx.x * MDVC_SCAN_TILE;
    T acc = 0;
#pragma unroll
    for (int i = 0; i < MDVC_SCAN_ITEMS; ++i) {
        size_t idx = base + (size_t)i * MDVC_SCAN_THREADS + threadIdx.x;
        if (idx < n) acc += in[idx];
    }
    acc = mdvc_block_reduce_sum(acc, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = acc;
}

// single block: exclusive scan of the tile sums in place; total -> *total_out (and *total_out2)
template <typename T>
__global__ void __launch_bounds__(1024)
k_scan_tile_offsets(T *tile_sums, uint32_t n_tiles, T *total_out, T *total_out2) {
    __shared__ T sm[33];
    uint32_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
    uint32_t lo = threadIdx.x * per, hi = min(lo + per, n_tiles);
    T acc = 0;
    for (uint32_t i = lo; i < hi; ++i) acc += tile_sums[i];
    T total;
    T off = mdvc_block_exscan(acc, sm, &total);
    for (uint32_t i = lo; i < hi; ++i) { T v = tile_sums[i]; tile_sums[i] = off; off += v; }
    if (threadIdx.x == 0) {
        if (total_out) *total_out = total;
        if (total_out2) *total_out2 = total;
    }
}

template <typename T>
__global__ void __launch_bounds__(MDVC_SCAN_THREADS)
k_scan_tiles(const T *in, const uint32_t *n_dev, uint32_t n_host, const T *__restrict__ tile_offsets, T *out) {
    __shared__ T sm[33];
    uint32_t n = n_dev ? *n_dev : n_host;
    if (n_dev && n > n_host) n = n_host;
    size_t base = (size_t)blockIdx.x * MDVC_SCAN_TILE;
    if (base >= n) return;
    // blocked arrangement: thread t owns items [t*ITEMS, (t+1)*ITEMS)
    T v[MDVC_SCAN_ITEMS];
    T acc = 0;
#pragma unroll
    for (int i = 0; i < MDVC_SCAN_ITEMS; ++i) {
        size_t idx = base + (size_t)threadIdx.x * MDVC_SCAN_ITEMS + i;
        v[i] = (idx < n) ? in[idx] : T(0);
        acc += v[i];
    }
    T total;
    T off = mdvc_block_exscan(acc, sm, &total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < MDVC_SCAN_ITEMS; ++i) {
        size_t idx = base + (size_t)threadIdx.x * MDVC_SCAN_ITEMS + i;
        if (idx < n) out[idx] = off;
        off += v[i];
    }
}

// exclusive scan of in[0..n) -> out[0..n) (in place allowed); n = *n_dev clamped to cap if n_dev.
// tile_scratch must hold ceil(cap / MDVC_SCAN_TILE) elements of T.
template <typename T>
static int mdvc_exclusive_scan(const T *in, T *out, const uint32_t *n_dev, uint32_t cap, T *tile_scratch,
                               T *total_out, T *total_out2, cudaStream_t st) {
    uint32_t n_tiles = (cap + MDVC_SCAN_TILE - 1) / MDVC_SCAN_TILE;
    if (n_tiles == 0) n_tiles = 1;
    k_scan_tile_sums<T><<<n_tiles, MDVC_SCAN_THREADS, 0, st>>>(in, n_dev, cap, tile_scratch);
    MDVC_LAUNCH_CHECK();
    k_scan_tile_offsets<T><<<1, 1024, 0, st>>>(tile_scratch, n_tiles, total_out, total_out2);
    MDVC_LAUNCH_CHECK();
    k_scan_tiles<T><<<n_tiles, MDVC_SCAN_THREADS, 0, st>>>(in, n_dev, cap, tile_scratch, out);
    MDVC_LAUNCH_CHECK();
    return MDVC_OK;
}

static inline size_t mdvc_scan_scratch_elems(uint32_t cap) {
    return (size_t)(cap + MDVC_SCAN_TILE - 1) / MDVC_SCAN_TILE + 1;
}

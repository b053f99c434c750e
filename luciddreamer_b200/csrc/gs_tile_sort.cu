// Per-tile depth sort (sm_100a).
//
// Replaces the reference's global 64-bit cub::DeviceRadixSort::SortPairs over all R (tile|depth, idx) pairs
// (RAST/cuda_rasterizer/rasterizer_impl.cu:301-309, 6 onesweep passes over 12-byte pairs) and
// identifyTileRanges (rasterizer_impl.cu:116-138).  Pairs were scattered into their tile's bucket by k_emit,
// so each tile only has to order its own bucket by key = (depth_bits << 32 | gaussian_idx): identical order
// to the reference's stable sort, whose ties (same tile, bit-identical depth) resolve by ascending Gaussian
// index (emission order).  One CTA per tile: bucket -> shared memory, bitonic network for arbitrary n
// (flip/half-cleaner form, all comparisons ascending, so the virtual +inf padding never moves), sorted
// Gaussian indices -> list.  Buckets larger than the shared-memory capacity are sorted in place in global
// memory by the same network (slow path, correctness only).
#include "gs_common.cuh"

namespace {

constexpr int kSortThreads = 256;
constexpr int kSmemKeys = 4096;     // 32 KB of u64 keys

template <typename Ptr>
__device__ __forceinline__ void bitonic_sort_any_n(Ptr a, const int n, const int tid, const int nthreads) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    const int half = np2 >> 1;
    for (int k = 2; k <= np2; k <<= 1) {
        // flip step: i pairs with the mirrored element of its k-block
        for (int t = tid; t < half; t += nthreads) {
            const int blk = t / (k >> 1), r = t - blk * (k >> 1);
            const int i = blk * k + r, j = blk * k + (k - 1 - r);
            if (j < n) {
                const unsigned long long x = a[i], y = a[j];
                if (x > y) { a[i] = y; a[j] = x; }
            }
        }
        __syncthreads();
        for (int d = k >> 2; d >= 1; d >>= 1) {
            for (int t = tid; t < half; t += nthreads) {
                const int i = ((t / d) * (d << 1)) + (t % d), j = i + d;
                if (j < n) {
                    const unsigned long long x = a[i], y = a[j];
                    if (x > y) { a[i] = y; a[j] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(kSortThreads)
k_tile_sort(const uint32_t* __restrict__ tile_off, uint32_t* __restrict__ tile_cur,
            const GsDevStatus* __restrict__ status, unsigned long long* __restrict__ keys,
            uint32_t* __restrict__ list, long long capacity) {
    if ((long long)status->num_pairs > capacity) return;
    __shared__ unsigned long long s_keys[kSmemKeys];
    const int tile = blockIdx.x;
    const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
    const int n = (int)(end - beg);
    if (threadIdx.x == 0) tile_cur[tile] = 0u;          // cursors back to zero for a possible re-render
    if (n == 0) return;
    unsigned long long* g = keys + beg;
    if (n <= kSmemKeys) {
        for (int t = threadIdx.x; t < n; t += kSortThreads) s_keys[t] = g[t];
        __syncthreads();
        if (n > 1) bitonic_sort_any_n(s_keys, n, threadIdx.x, kSortThreads);
        for (int t = threadIdx.x; t < n; t += kSortThreads) list[beg + t] = (uint32_t)s_keys[t];
    } else {
        __syncthreads();
        bitonic_sort_any_n(g, n, threadIdx.x, kSortThreads);
        for (int t = threadIdx.x; t < n; t += kSortThreads) list[beg + t] = (uint32_t)g[t];
    }
}

}  // namespace

void gs_launch_tile_sort(int G, const uint32_t* tile_off, uint32_t* tile_cur, const GsDevStatus* status,
                         unsigned long long* keys, uint32_t* list, long long capacity, cudaStream_t s) {
    k_tile_sort<<<G, kSortThreads, 0, s>>>(tile_off, tile_cur, status, keys, list, capacity);
}

// Per-tile depth sort (sm_100a).
//
// Replaces the reference's global 64-bit cub::DeviceRadixSort::SortPairs over all R (tile|depth, idx) pairs
// (RAST/cuda_rasterizer/rasterizer_impl.cu:301-309, 6 onesweep passes over 12-byte pairs) and
// identifyTileRanges (rasterizer_impl.cu:116-138).  Pairs were scattered into their tile's bucket by k_emit,
// so each tile only has to order its own bucket by key = (depth_bits << 32 | gaussian_idx): identical order
// to the reference's stable sort, whose ties (same tile, bit-identical depth) resolve by ascending Gaussian
// index (emission order).
//   k_tile_sort      one WARP per tile for buckets up to 256 keys (shared memory, __syncwarp only); larger
//                    buckets are queued (mid from the front of the queue, big from the back) for
//   k_tile_sort_mid  one 128-thread CTA per queued tile, up to 2048 keys (2 x 16 KB shared memory), persistent grid
//   k_tile_sort_big  one 512-thread CTA per queued tile, up to 8192 keys (2 x 64 KB); beyond that a bitonic network
//                    runs in place in global memory (slow path, correctness only).
// Algorithm in shared memory: runs of 32 keys are sorted in registers with a warp-shuffle bitonic network, then
// runs are merged pairwise by rank: every key binary-searches its position in the sibling run (keys are unique, so
// "number of sibling keys smaller than mine" is its exact offset) and is written to own-offset + rank in a second
// buffer.  ~log2(n/32) passes of log2(L)+1 shared-memory probes per key instead of the ~log2(n)^2/2 compare-exchange
// stages of a full bitonic network.
#include "gs_common.cuh"

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kWarpKeys = 256;       // 2 x 2 KB of u64 keys per warp
constexpr int kMidThreads = 256;
constexpr int kMidKeys = 2048;       // 2 x 16 KB static shared memory
constexpr int kBigThreads = 1024;
constexpr int kBigKeys = 8192;       // 2 x 64 KB dynamic shared memory

// sorts the 32 keys held one per lane (ascending across lanes); keys are unique
__device__ __forceinline__ unsigned long long warp_sort32(unsigned long long key, const int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j >= 1; j >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, j);
            const bool up = ((lane & k) == 0);            // ascending block?
            const bool lower = ((lane & j) == 0);         // I hold the lower position of the pair
            const bool take_min = (up == lower);
            const bool other_smaller = other < key;
            key = (take_min == other_smaller) ? other : key;
        }
    }
    return key;
}

// number of elements of a[0..len) that are < key; len <= L, L a power of two
__device__ __forceinline__ int rank_in_run(const unsigned long long* a, const int len, const int L,
                                           const unsigned long long key) {
    int pos = 0;
    for (int step = L; step >= 1; step >>= 1) {
        const int probe = pos + step;
        if (probe <= len && a[probe - 1] < key) pos = probe;
    }
    return pos;
}

// Sorts n unique keys that sit in A[0..n); B is scratch of the same capacity.  Returns the buffer holding the result.
// All `nthreads` threads of the group call; sync() synchronises the group.
template <typename Sync>
__device__ __forceinline__ unsigned long long* merge_rank_sort(unsigned long long* A, unsigned long long* B, const int n,
                                                               const int tid, const int nthreads, Sync sync) {
    const int lane = tid & 31;
    // phase 1: each warp sorts runs of 32 in registers
    for (int c = (tid >> 5) * 32; c < n; c += nthreads) {
        const int i = c + lane;
        unsigned long long key = i < n ? A[i] : ~0ull;    // +inf padding sorts to the end of the run
        key = warp_sort32(key, lane);
        if (i < n) A[i] = key;
    }
    sync();
    // phase 2: pairwise merges by rank
    for (int L = 32; L < n; L <<= 1) {
        for (int i = tid; i < n; i += nthreads) {
            const int own = i & ~(L - 1);
            const int sib = own ^ L;
            const int sl = min(L, n - sib);               // <= 0 when the sibling run does not exist
            const unsigned long long key = A[i];
            // ties cannot occur (unique keys), so "< key" ranks are exact for both sides of the pair
            const int rank = sl > 0 ? rank_in_run(A + sib, sl, L, key) : 0;
            B[min(own, sib) + (i - own) + rank] = key;
        }
        sync();
        unsigned long long* t = A; A = B; B = t;
    }
    return A;
}

// ---- merge-path sort for the long buckets (mid / big kernels).  Runs of kVT keys are sorted in registers by an
// odd-even transposition network, then sorted runs are merged pairwise: every thread produces kVT consecutive outputs
// of a merged pair -- ONE binary search along its merge-path diagonal (log2 L probes), then a serial merge (one
// shared-memory read per output) -- instead of one binary search per key per pass.  ~10 instructions per key and pass.
constexpr int kVT = 8;

__device__ __forceinline__ void cswap(unsigned long long& a, unsigned long long& b) {
    const unsigned long long lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo; b = hi;
}

// Shared-memory layout: one padding word after every kVT keys (logical index i lives at i + i / 8).  Every thread works
// on kVT CONSECUTIVE keys; unpadded, the 64-byte stride between neighbouring threads would put a warp's 64-bit accesses
// on two bank pairs (16-way conflicts); with the pad the stride is 72 bytes and a half-warp covers all 32 banks.
__device__ __forceinline__ int ph(const int i) { return i + (i >> 3); }
__host__ __device__ constexpr int padded_keys(const int n) { return n + n / 8 + 8; }

// Sorts n unique keys that sit at A[ph(0..n)); B is scratch; both hold padded_keys(capacity) words.  Returns the buffer
// that holds the result (same layout).  All `nthreads` threads call; sync() synchronises them.
template <typename Sync>
__device__ __forceinline__ unsigned long long* merge_path_sort(unsigned long long* A, unsigned long long* B, const int n,
                                                               const int tid, const int nthreads, Sync sync) {
    const int nseg = (n + kVT - 1) / kVT, npad = nseg * kVT;
    for (int i = n + tid; i < npad; i += nthreads) A[ph(i)] = ~0ull;    // +inf padding sorts to the very end
    sync();
    // phase 1: runs of kVT, in registers
    for (int sg = tid; sg < nseg; sg += nthreads) {
        unsigned long long k[kVT];
        unsigned long long* run = A + ph(sg * kVT);      // kVT consecutive words (the pad comes after them)
#pragma unroll
        for (int j = 0; j < kVT; j++) k[j] = run[j];
#pragma unroll
        for (int r = 0; r < kVT; r++) {
#pragma unroll
            for (int j = r & 1; j + 1 < kVT; j += 2) cswap(k[j], k[j + 1]);
        }
#pragma unroll
        for (int j = 0; j < kVT; j++) run[j] = k[j];
    }
    sync();
    // phase 2: pairwise merges of runs of length L (the last run of a level may be shorter or missing)
    for (int L = kVT; L < npad; L <<= 1) {
        for (int sg = tid; sg < nseg; sg += nthreads) {
            const int out0 = sg * kVT;                   // first output position of this thread's segment
            const int base = out0 & ~(2 * L - 1);        // start of the pair of runs (a multiple of kVT)
            const int la = min(L, npad - base), lb = max(0, min(L, npad - base - L));
            const unsigned long long* a = A + ph(base);          // ph(base + i) == ph(base) + ph(i) for base % 8 == 0
            const unsigned long long* b = A + ph(base + L);
            const int d = out0 - base;                   // diagonal: outputs [d, d + kVT) of the merged pair
            int lo = max(0, d - lb), hi = min(d, la);    // merge path: i = #elements taken from a before diagonal d
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (a[ph(mid)] < b[ph(d - 1 - mid)]) lo = mid + 1; else hi = mid;
            }
            int i = lo, j = d - lo;
            unsigned long long ka = i < la ? a[ph(i)] : ~0ull, kb = j < lb ? b[ph(j)] : ~0ull;
            unsigned long long* out = B + ph(out0);
#pragma unroll
            for (int o = 0; o < kVT; o++) {
                const bool take_a = (j >= lb) || (i < la && ka < kb);
                out[o] = take_a ? ka : kb;
                if (take_a) { i++; ka = i < la ? a[ph(i)] : ~0ull; }
                else { j++; kb = j < lb ? b[ph(j)] : ~0ull; }
            }
        }
        sync();
        unsigned long long* t = A; A = B; B = t;
    }
    return A;
}

template <typename Ptr, typename Sync>
__device__ __forceinline__ void bitonic_sort_any_n(Ptr a, const int n, const int tid, const int nthreads, Sync sync) {
    int lg = 0;
    while ((1 << lg) < n) lg++;                           // np2 = 1 << lg
    const int half = (1 << lg) >> 1;
    for (int lk = 1; lk <= lg; lk++) {                    // k = 1 << lk; all index maths in shifts and masks
        const int k = 1 << lk, hk = k >> 1;
        for (int t = tid; t < half; t += nthreads) {      // flip: i pairs with its mirror inside the k-block
            const int r = t & (hk - 1);
            const int b0 = (t >> (lk - 1)) << lk;
            const int i = b0 + r, j = b0 + (k - 1 - r);
            if (j < n) {
                const unsigned long long x = a[i], y = a[j];
                if (x > y) { a[i] = y; a[j] = x; }
            }
        }
        sync();
        for (int ld = lk - 2; ld >= 0; ld--) {            // d = 1 << ld
            const int d = 1 << ld;
            for (int t = tid; t < half; t += nthreads) {
                const int i = ((t >> ld) << (ld + 1)) + (t & (d - 1)), j = i + d;
                if (j < n) {
                    const unsigned long long x = a[i], y = a[j];
                    if (x > y) { a[i] = y; a[j] = x; }
                }
            }
            sync();
        }
    }
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_tile_sort(int G, const uint32_t* __restrict__ tile_off, uint32_t* __restrict__ tile_cur,
            GsDevStatus* __restrict__ status, uint32_t* __restrict__ big_list,
            const unsigned long long* __restrict__ keys, uint32_t* __restrict__ list, long long capacity) {
    if ((long long)status->num_pairs > capacity) return;
    __shared__ unsigned long long s_keys[kWarpsPerCta][2][kWarpKeys];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tile = blockIdx.x * kWarpsPerCta + wid;
    if (tile >= G) return;
    const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
    const int n = (int)(end - beg);
    if (lane == 0) tile_cur[tile] = 0u;                  // cursors back to zero for a possible re-render
    if (n == 0) return;
    if (n > kWarpKeys) {                                 // queue: mid tiles from the front, big tiles from the back
        if (lane == 0) {
            if (n <= kMidKeys) big_list[atomicAdd(&status->n_mid, 1u)] = (uint32_t)tile;
            else big_list[G - 1 - atomicAdd(&status->n_big, 1u)] = (uint32_t)tile;
        }
        return;
    }
    unsigned long long* a = s_keys[wid][0];
    const unsigned long long* g = keys + beg;
    for (int t = lane; t < n; t += 32) a[t] = g[t];
    __syncwarp();
    if (n > 1) a = merge_rank_sort(a, s_keys[wid][1], n, lane, 32, [] { __syncwarp(); });
    for (int t = lane; t < n; t += 32) list[beg + t] = (uint32_t)a[t];
}

// buckets of 257..4096 keys: one 128-thread CTA per tile, persistent over the queue
__global__ void __launch_bounds__(kMidThreads)
k_tile_sort_mid(const uint32_t* __restrict__ tile_off, const GsDevStatus* __restrict__ status,
                const uint32_t* __restrict__ big_list, const unsigned long long* __restrict__ keys,
                uint32_t* __restrict__ list, long long capacity) {
    if ((long long)status->num_pairs > capacity) return;
    __shared__ unsigned long long s_mid[2][padded_keys(kMidKeys)];
    const unsigned nmid = status->n_mid;
    for (unsigned b = blockIdx.x; b < nmid; b += gridDim.x) {
        const uint32_t tile = big_list[b];
        const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
        const int n = (int)(end - beg);
        const unsigned long long* g = keys + beg;
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += kMidThreads) s_mid[0][ph(t)] = g[t];
        __syncthreads();
        const unsigned long long* r = merge_path_sort(s_mid[0], s_mid[1], n, threadIdx.x, kMidThreads, [] { __syncthreads(); });
        for (int t = threadIdx.x; t < n; t += kMidThreads) list[beg + t] = (uint32_t)r[ph(t)];
    }
}

__global__ void __launch_bounds__(kBigThreads)
k_tile_sort_big(int G, const uint32_t* __restrict__ tile_off, const GsDevStatus* __restrict__ status,
                const uint32_t* __restrict__ big_list, unsigned long long* __restrict__ keys,
                uint32_t* __restrict__ list, long long capacity) {
    if ((long long)status->num_pairs > capacity) return;
    extern __shared__ unsigned long long s_big[];
    const unsigned nbig = status->n_big;
    for (unsigned b = blockIdx.x; b < nbig; b += gridDim.x) {
        const uint32_t tile = big_list[G - 1 - b];
        const uint32_t beg = tile_off[tile], end = tile_off[tile + 1];
        const int n = (int)(end - beg);
        unsigned long long* g = keys + beg;
        __syncthreads();
        if (n <= kBigKeys) {
            for (int t = threadIdx.x; t < n; t += kBigThreads) s_big[ph(t)] = g[t];
            __syncthreads();
            const unsigned long long* r = merge_path_sort(s_big, s_big + padded_keys(kBigKeys), n, threadIdx.x, kBigThreads, [] { __syncthreads(); });
            for (int t = threadIdx.x; t < n; t += kBigThreads) list[beg + t] = (uint32_t)r[ph(t)];
        } else {
            bitonic_sort_any_n(g, n, threadIdx.x, kBigThreads, [] { __syncthreads(); });
            for (int t = threadIdx.x; t < n; t += kBigThreads) list[beg + t] = (uint32_t)g[t];
        }
    }
}

}  // namespace

// per device, once (called from gs_context_create with the device current)
void gs_tile_sort_init() {
    cudaFuncSetAttribute(k_tile_sort_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * padded_keys(kBigKeys) * 8);
}

void gs_launch_tile_sort(int G, int num_sms, const uint32_t* tile_off, uint32_t* tile_cur, GsDevStatus* status,
                         uint32_t* big_list, unsigned long long* keys, uint32_t* list, long long capacity,
                         cudaStream_t s, cudaEvent_t* prof) {
    if (prof) cudaEventRecord(prof[0], s);
    k_tile_sort<<<(G + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32, 0, s>>>(G, tile_off, tile_cur, status,
                                                                                 big_list, keys, list, capacity);
    const int gmid = G < num_sms * 6 ? G : num_sms * 6;
    k_tile_sort_mid<<<gmid, kMidThreads, 0, s>>>(tile_off, status, big_list, keys, list, capacity);
    if (prof) { cudaEventRecord(prof[1], s); cudaEventRecord(prof[2], s); }
    const int grid = G < num_sms ? G : num_sms;
    k_tile_sort_big<<<grid, kBigThreads, 2 * padded_keys(kBigKeys) * 8, s>>>(G, tile_off, status, big_list, keys, list, capacity);
    if (prof) cudaEventRecord(prof[3], s);
}

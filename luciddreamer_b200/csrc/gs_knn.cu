// Mean squared distance to the 3 nearest neighbours of every point (sm_100a) -- the "next" row SURVEY.md 8f-3,
// the B200 equivalent of the reference's second native extension simple_knn (SimpleKNN::knn, simple_knn.cu:185-221;
// binding distCUDA2, spatial.cu:15-26; call site scene/gaussian_model.py:136, once per scene to initialise scales).
//
// Same result, different search.  The reference sorts by a 30-bit Morton code, cuts the sorted order into boxes of
// 1024 points and lets EVERY THREAD walk ALL boxes serially, scanning each accepted box point by point (divergent,
// every thread re-reading 1024 scattered points per box).  Here the sorted points are gathered once into a dense
// float4 array and organised in two levels -- leaves of 32 points (one warp) under nodes of 32 leaves -- and one WARP
// answers the 32 queries of a leaf together:
//   1. all pairs inside the own leaf (shuffle-free: the leaf is staged in shared memory) give every lane real
//      candidates, so the pruning bound is tight from the start (the reference seeds it from +-3 sorted neighbours);
//   2. lanes test 32 nodes at a time against the leaf's bounding box (box-box distance vs. the warp's largest
//      3rd-best), then the 32 leaves of each surviving node, one per lane;
//   3. a surviving leaf is staged by one coalesced 512-byte load and every lane whose own point-box distance beats
//      its own 3rd-best scans the 32 broadcast points with a branch-free min/max insertion.
// Exactness: pair distances use the reference's arithmetic (dx*dx + dy*dy + dz*dz contracted to fma(dz,dz, fma(dx,dx, dy*dy)),
// which is how nvcc compiles updateKBest, simple_knn.cu:135-146 -- read off the SASS of the reference build); box distances use the same monotone operation chain on per-axis gaps
// that are never larger than the coordinate differences of any point inside, so a pruned box can only hold points
// whose computed distance is >= the bound.  The 3 smallest distances -- hence (d0 + d1 + d2) / 3 -- are therefore
// bit-identical to the reference's for every input, including duplicates (distance 0 counts, only the query itself
// is skipped) and P < 4 (missing neighbours stay FLT_MAX and the float sum overflows exactly like the reference's).
// The 30-bit key sort is cub::DeviceRadixSort, the same library call the reference makes (simple_knn.cu:210-213).
#include <cfloat>
#include <cub/device/device_radix_sort.cuh>

#include "gs_common.cuh"

namespace {

constexpr int kLeaf = 32;
constexpr int kNodeLeaves = 32;
constexpr int kNode = kLeaf * kNodeLeaves;
constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bounds[0..2] = ordered-encoded min, bounds[3..5] = ordered-encoded max (initialised by k_knn_init)
__global__ void k_knn_init(unsigned* bounds) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256)
k_knn_bounds(const int P, const float* __restrict__ pts, unsigned* bounds) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float v = pts[3 * (size_t)i + a];
            lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(~0u, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(~0u, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) { atomicMin(&bounds[a], f2ord(lo[a])); atomicMax(&bounds[3 + a], f2ord(hi[a])); }
    }
}

__device__ __forceinline__ unsigned spread10(unsigned x) {     // 10 bits -> every third bit
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

// 30-bit Morton key of every point on the 1024^3 grid spanned by the bounding box (simple_knn.cu:45-70); only the
// ORDER it induces matters (it decides which points share a leaf), never the result.
__global__ void __launch_bounds__(256)
k_knn_morton(const int P, const float* __restrict__ pts, const unsigned* __restrict__ bounds, unsigned* keys, unsigned* vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    unsigned code = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float lo = ord2f(bounds[a]), hi = ord2f(bounds[3 + a]);
        const float ext = hi - lo;
        float t = ext > 0.f ? (pts[3 * (size_t)i + a] - lo) / ext : 0.f;
        t = fminf(fmaxf(t, 0.f), 1.f);                       // also sends NaN coordinates to cell 0
        code |= spread10((unsigned)(t * 1023.f)) << a;
    }
    keys[i] = code;
    vals[i] = (unsigned)i;
}

// one 1024-thread CTA per node: gathers the node's points in sorted order (float4: x, y, z, original index), the box
// of each of its 32 leaves (warp reduction) and the node box (shared-memory reduction over the 32 leaves)
__global__ void __launch_bounds__(kNode)
k_knn_gather(const int P, const float* __restrict__ pts, const unsigned* __restrict__ order, float4* __restrict__ sorted,
             float4* __restrict__ leaf_lo, float4* __restrict__ leaf_hi, float4* __restrict__ node_lo,
             float4* __restrict__ node_hi) {
    __shared__ float s_lo[kNodeLeaves][3], s_hi[kNodeLeaves][3];
    const int i = blockIdx.x * kNode + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < P) {
        const unsigned src = order[i];
        const float x = pts[3 * (size_t)src], y = pts[3 * (size_t)src + 1], z = pts[3 * (size_t)src + 2];
        sorted[i] = make_float4(x, y, z, __uint_as_float(src));
        lo[0] = hi[0] = x; lo[1] = hi[1] = y; lo[2] = hi[2] = z;
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(~0u, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(~0u, hi[a], o));
        }
    }
    const int leaf = blockIdx.x * kNodeLeaves + w;
    if (lane == 0) {
        if ((size_t)leaf * kLeaf < (size_t)P) {
            leaf_lo[leaf] = make_float4(lo[0], lo[1], lo[2], 0.f);
            leaf_hi[leaf] = make_float4(hi[0], hi[1], hi[2], 0.f);
        }
#pragma unroll
        for (int a = 0; a < 3; a++) { s_lo[w][a] = lo[a]; s_hi[w][a] = hi[a]; }
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            lo[a] = s_lo[lane][a]; hi[a] = s_hi[lane][a];
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                lo[a] = fminf(lo[a], __shfl_xor_sync(~0u, lo[a], o));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(~0u, hi[a], o));
            }
        }
        if (lane == 0) {
            node_lo[blockIdx.x] = make_float4(lo[0], lo[1], lo[2], 0.f);
            node_hi[blockIdx.x] = make_float4(hi[0], hi[1], hi[2], 0.f);
        }
    }
}

__device__ __forceinline__ float sq3(const float dx, const float dy, const float dz) {
    // dx*dx + dy*dy + dz*dz as nvcc contracts it in the reference's updateKBest (SASS of boxMeanDist: FMUL on y, FFMA
    // with x, FFMA with z)
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, dy * dy));
}
// squared distance point -> box (distBoxPoint, simple_knn.cu:120-132): per axis the gap to the nearer face, 0 inside
__device__ __forceinline__ float box_point2(const float4 lo, const float4 hi, const float4 p) {
    const float dx = fmaxf(fmaxf(lo.x - p.x, p.x - hi.x), 0.f);
    const float dy = fmaxf(fmaxf(lo.y - p.y, p.y - hi.y), 0.f);
    const float dz = fmaxf(fmaxf(lo.z - p.z, p.z - hi.z), 0.f);
    return sq3(dx, dy, dz);
}
__device__ __forceinline__ float box_box2(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi) {
    const float dx = fmaxf(fmaxf(blo.x - ahi.x, alo.x - bhi.x), 0.f);
    const float dy = fmaxf(fmaxf(blo.y - ahi.y, alo.y - bhi.y), 0.f);
    const float dz = fmaxf(fmaxf(blo.z - ahi.z, alo.z - bhi.z), 0.f);
    return sq3(dx, dy, dz);
}
__device__ __forceinline__ void insert3(float& b0, float& b1, float& b2, const float d) {
    const float t0 = fmaxf(b0, d);  b0 = fminf(b0, d);
    const float t1 = fmaxf(b1, t0); b1 = fminf(b1, t0);
    b2 = fminf(b2, t1);
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(~0u, v, o));
    return v;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_knn_search(const int P, const int n_leaves, const int n_nodes, const float4* __restrict__ sorted,
             const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi, const float4* __restrict__ node_lo,
             const float4* __restrict__ node_hi, float* __restrict__ out) {
    __shared__ float4 s_pts[kWarpsPerCta][kLeaf];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int leaf = blockIdx.x * kWarpsPerCta + w;
    if (leaf >= n_leaves) return;                                  // whole warps leave together
    float4* sp = s_pts[w];
    const int me = leaf * kLeaf + lane;
    const bool valid = me < P;
    const float4 p = valid ? sorted[me] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 my_lo = leaf_lo[leaf], my_hi = leaf_hi[leaf];
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;

    // 1. own leaf, all pairs
    sp[lane] = p;
    __syncwarp();
    const int own_cnt = min(kLeaf, P - leaf * kLeaf);
    for (int j = 0; j < own_cnt; j++) {
        const float4 q = sp[j];
        const float d = sq3(q.x - p.x, q.y - p.y, q.z - p.z);
        if (j != lane) insert3(b0, b1, b2, d);
    }
    __syncwarp();
    // an invalid lane never searches: its bound is negative, every box distance is >= 0
    float bound = valid ? b2 : -1.f;
    float wbound = warp_max(bound);

    // 2. the node's leaves, one per lane; every surviving leaf is staged and scanned
    auto visit_node = [&](const int nd) {
        const int lf = nd * kNodeLeaves + lane;
        float4 llo = make_float4(0.f, 0.f, 0.f, 0.f), lhi = llo;
        bool lhit = false;
        if (lf < n_leaves && lf != leaf) {
            llo = leaf_lo[lf]; lhi = leaf_hi[lf];
            lhit = !(box_box2(my_lo, my_hi, llo, lhi) > wbound);
        }
        unsigned lmask = __ballot_sync(~0u, lhit);
        while (lmask) {
            const int src = __ffs(lmask) - 1;
            lmask &= lmask - 1;
            const float4 blo = make_float4(__shfl_sync(~0u, llo.x, src), __shfl_sync(~0u, llo.y, src),
                                           __shfl_sync(~0u, llo.z, src), 0.f);
            const float4 bhi = make_float4(__shfl_sync(~0u, lhi.x, src), __shfl_sync(~0u, lhi.y, src),
                                           __shfl_sync(~0u, lhi.z, src), 0.f);
            const bool want = !(box_point2(blo, bhi, p) > bound);           // bound < 0 for invalid lanes: never
            if (__ballot_sync(~0u, want) == 0) continue;
            const int base = (nd * kNodeLeaves + src) * kLeaf;
            const int cnt = min(kLeaf, P - base);
            if (lane < cnt) sp[lane] = sorted[base + lane];
            __syncwarp();
            if (want) {
                for (int j = 0; j < cnt; j++) {
                    const float4 q = sp[j];
                    insert3(b0, b1, b2, sq3(q.x - p.x, q.y - p.y, q.z - p.z));
                }
                bound = b2;
            }
            __syncwarp();
            wbound = warp_max(bound);
        }
    };
    // own node first (its leaves are the Morton neighbours: tightens the bound before anything else is tested),
    // then all other nodes, 32 box tests at a time
    const int own_node = leaf / kNodeLeaves;
    visit_node(own_node);
    for (int nb = 0; nb < n_nodes; nb += 32) {
        const int node = nb + lane;
        bool hit = false;
        if (node < n_nodes && node != own_node) hit = !(box_box2(my_lo, my_hi, node_lo[node], node_hi[node]) > wbound);
        unsigned nmask = __ballot_sync(~0u, hit);
        while (nmask) {
            const int nd = nb + __ffs(nmask) - 1;
            nmask &= nmask - 1;
            visit_node(nd);
        }
    }
    if (valid) out[__float_as_uint(p.w)] = (b0 + b1 + b2) / 3.0f;         // simple_knn.cu:182
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct KnnLayout {
    size_t bounds, keys, vals, keys_s, vals_s, sorted, leaf_lo, leaf_hi, node_lo, node_hi, cub, total;
    int n_leaves, n_nodes;
};

KnnLayout knn_layout(int P, size_t cub_bytes) {
    KnnLayout L;
    L.n_leaves = (P + kLeaf - 1) / kLeaf;
    L.n_nodes = (P + kNode - 1) / kNode;
    size_t o = 0;
    L.bounds = o; o += 256;
    L.keys = o; o += align256((size_t)P * 4);
    L.vals = o; o += align256((size_t)P * 4);
    L.keys_s = o; o += align256((size_t)P * 4);
    L.vals_s = o; o += align256((size_t)P * 4);
    L.sorted = o; o += align256((size_t)P * 16);
    L.leaf_lo = o; o += align256((size_t)L.n_leaves * 16);
    L.leaf_hi = o; o += align256((size_t)L.n_leaves * 16);
    L.node_lo = o; o += align256((size_t)L.n_nodes * 16);
    L.node_hi = o; o += align256((size_t)L.n_nodes * 16);
    L.cub = o; o += align256(cub_bytes);
    L.total = o;
    return L;
}

// allowance for cub::DeviceRadixSort::SortPairs' temporary storage (DoubleBuffer interface: histograms and tile
// look-back state only) that does not need a device to evaluate (the real requirement is queried at launch time and checked against it)
size_t cub_allowance(int P) { return ((size_t)1 << 20) + (size_t)P; }

}  // namespace

size_t gs_knn_scratch_bytes_impl(int P) { return knn_layout(P, cub_allowance(P)).total; }

int gs_launch_knn(int P, const float* points, void* scratch, float* out, cudaStream_t s) {
    if (P <= 0) return 0;
    const KnnLayout L = knn_layout(P, cub_allowance(P));
    char* base = static_cast<char*>(scratch);
    unsigned* bounds = reinterpret_cast<unsigned*>(base + L.bounds);
    unsigned* keys = reinterpret_cast<unsigned*>(base + L.keys);
    unsigned* vals = reinterpret_cast<unsigned*>(base + L.vals);
    unsigned* keys_s = reinterpret_cast<unsigned*>(base + L.keys_s);
    unsigned* vals_s = reinterpret_cast<unsigned*>(base + L.vals_s);
    float4* sorted = reinterpret_cast<float4*>(base + L.sorted);
    float4* leaf_lo = reinterpret_cast<float4*>(base + L.leaf_lo);
    float4* leaf_hi = reinterpret_cast<float4*>(base + L.leaf_hi);
    float4* node_lo = reinterpret_cast<float4*>(base + L.node_lo);
    float4* node_hi = reinterpret_cast<float4*>(base + L.node_hi);
    size_t need = 0;
    cub::DoubleBuffer<unsigned> dk(keys, keys_s), dv(vals, vals_s);
    if (cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, P, 0, 30, s) != cudaSuccess) return -2;
    if (need > cub_allowance(P)) return -1;
    k_knn_init<<<1, 32, 0, s>>>(bounds);
    const int nb = min((P + 255) / 256, 148 * 8);
    k_knn_bounds<<<nb, 256, 0, s>>>(P, points, bounds);
    k_knn_morton<<<(P + 255) / 256, 256, 0, s>>>(P, points, bounds, keys, vals);
    size_t tmp = cub_allowance(P);
    if (cub::DeviceRadixSort::SortPairs(base + L.cub, tmp, dk, dv, P, 0, 30, s) != cudaSuccess) return -2;
    k_knn_gather<<<L.n_nodes, kNode, 0, s>>>(P, points, dv.Current(), sorted, leaf_lo, leaf_hi, node_lo, node_hi);
    k_knn_search<<<(L.n_leaves + kWarpsPerCta - 1) / kWarpsPerCta, kWarpsPerCta * 32, 0, s>>>(
        P, L.n_leaves, L.n_nodes, sorted, leaf_lo, leaf_hi, node_lo, node_hi, out);
    return 0;
}

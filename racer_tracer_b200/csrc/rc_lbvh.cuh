// rc_lbvh.cuh — BVH construction on the GPU: a linear BVH over the Morton codes of the top-level
// objects' box centres (Lauterbach et al. 2009; node topology after Karras, "Maximizing parallelism in
// the construction of BVHs, octrees, and k-d trees", HPG 2012), emitted in the threaded pre-order layout
// the traversal kernels read (rt_scene.cuh: left child = index + 1, `skip` = first node after the
// subtree, leaves = one top-level object).
//
// It replaces BoundingVolumeHirearchy::new / Node::build (src/bvh_node.rs:31-82,142-170) — a recursive
// median split over cloned objects, rebuilt on the host whenever an object moves (bvh_node.rs:176-205) —
// for scenes where the build is worth doing on the device (the Random scene's ~480 spheres and anything
// larger).  Leaves keep the reference's semantics: the object's STORED Aabb is its cull volume
// (bvh_node.rs:119), boxes of inner nodes are unions (aabb.rs:95-114).
//
// Kernels (N objects, all O(N) threads):
//   lbvh_bounds      scene bounds of the box centres (block reduction + ordered-int atomics)
//   lbvh_morton      30-bit Morton code | object index -> 64-bit keys
//   lbvh_sort_*      bitonic sort of the keys (one CTA in shared memory up to 8192 keys, global steps above)
//   lbvh_topology    Karras: every inner node's range, split and children; parents
//   lbvh_fit         bottom-up: leaves climb, the second arrival at a node unions its children (f64 boxes)
//   lbvh_emit        pre-order index of every node by walking to the root; writes DevNode / DevNodeD and the
//                    object order
//   lbvh_gather_*    primitive tables permuted into depth-first leaf order
#pragma once
#include "rt_scene.cuh"

struct LbvhObject {      // one top-level object: its stored Aabb and its run of primitives
    double lo[3], hi[3];
    int first, count;
};

struct LbvhNode {        // inner node i of the Karras tree; leaves are addressed as (n_inner + k)
    int left, right;     // child ids: < n_inner inner, >= n_inner leaf (sorted position + n_inner)
    int parent;
    int first, last;     // range of sorted leaves covered
    double lo[3], hi[3];
};

RT_D unsigned lbvh_expand_bits(unsigned v) {   // 10 bits -> every third bit
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

RT_D long long lbvh_ordered(double v) {         // order-preserving map double -> signed 64-bit
    long long b = __double_as_longlong(v);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
RT_D double lbvh_unordered(long long b) { return __longlong_as_double(b >= 0 ? b : (b ^ 0x7fffffffffffffffLL)); }

// bounds[0..2] = min of centres, bounds[3..5] = max of centres (as ordered integers)
__global__ void lbvh_bounds(const LbvhObject* __restrict__ obj, int n, long long* __restrict__ bounds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double c[3] = {0, 0, 0};
    const bool live = i < n;
    if (live) for (int a = 0; a < 3; ++a) c[a] = 0.5 * (obj[i].lo[a] + obj[i].hi[a]);
    for (int a = 0; a < 3; ++a) {
        long long mn = live ? lbvh_ordered(c[a]) : 0x7fffffffffffffffLL, mx = live ? lbvh_ordered(c[a]) : (long long)0x8000000000000000ULL;
        for (int off = 16; off > 0; off >>= 1) {
            const long long omn = __shfl_xor_sync(0xffffffffu, mn, off), omx = __shfl_xor_sync(0xffffffffu, mx, off);
            mn = omn < mn ? omn : mn;
            mx = omx > mx ? omx : mx;
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(bounds + a, mn); atomicMax(bounds + 3 + a, mx); }
    }
}

__global__ void lbvh_morton(const LbvhObject* __restrict__ obj, int n, int padded, const long long* __restrict__ bounds,
                            unsigned long long* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= padded) return;
    if (i >= n) { keys[i] = ~0ull; return; }   // padding sorts to the end
    unsigned code = 0;
    for (int a = 0; a < 3; ++a) {
        const double lo = lbvh_unordered(bounds[a]), hi = lbvh_unordered(bounds[3 + a]);
        const double c = 0.5 * (obj[i].lo[a] + obj[i].hi[a]);
        double u = hi > lo ? (c - lo) / (hi - lo) : 0.0;
        u = fmin(fmax(u * 1024.0, 0.0), 1023.0);
        code |= lbvh_expand_bits((unsigned)u) << (2 - a);
    }
    keys[i] = ((unsigned long long)code << 32) | (unsigned)i;   // the index makes every key distinct
}

#define LBVH_SMEM_KEYS 8192
__global__ void __launch_bounds__(1024) lbvh_sort_smem(unsigned long long* __restrict__ keys, int padded) {
    extern __shared__ unsigned long long sk[];
    for (int i = threadIdx.x; i < padded; i += blockDim.x) sk[i] = keys[i];
    __syncthreads();
    for (int k = 2; k <= padded; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < padded; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long a = sk[i], b = sk[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { sk[i] = b; sk[p] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < padded; i += blockDim.x) keys[i] = sk[i];
}

__global__ void lbvh_sort_step(unsigned long long* __restrict__ keys, int padded, int j, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= padded) return;
    const int p = i ^ j;
    if (p > i) {
        const unsigned long long a = keys[i], b = keys[p];
        const bool up = (i & k) == 0;
        if ((a > b) == up) { keys[i] = b; keys[p] = a; }
    }
}

// length of the common prefix of sorted keys i and j (-1 outside the array): Karras' delta
RT_D int lbvh_delta(const unsigned long long* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll(keys[i] ^ keys[j]);
}

__global__ void lbvh_topology(const unsigned long long* __restrict__ keys, int n, LbvhNode* __restrict__ nodes, int* __restrict__ leaf_parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_inner = n - 1;
    if (i >= n_inner) return;
    const int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t > 0; t >>= 1)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int first = i < j ? i : j, last = i < j ? j : i;
    const int left = first == gamma ? n_inner + gamma : gamma;
    const int right = last == gamma + 1 ? n_inner + gamma + 1 : gamma + 1;
    nodes[i].left = left; nodes[i].right = right; nodes[i].first = first; nodes[i].last = last;
    if (i == 0) nodes[0].parent = -1;
    if (left < n_inner) nodes[left].parent = i; else leaf_parent[left - n_inner] = i;
    if (right < n_inner) nodes[right].parent = i; else leaf_parent[right - n_inner] = i;
}

__global__ void lbvh_fit(const unsigned long long* __restrict__ keys, const LbvhObject* __restrict__ obj, int n, LbvhNode* __restrict__ nodes,
                         const int* __restrict__ leaf_parent, int* __restrict__ arrived) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int n_inner = n - 1;
    int node = leaf_parent[k];
    while (node >= 0) {
        if (atomicAdd(arrived + node, 1) == 0) return;   // the first child to arrive waits for its sibling
        __threadfence();
        LbvhNode& nd = nodes[node];
        for (int side = 0; side < 2; ++side) {
            const int c = side ? nd.right : nd.left;
            // an inner child's box was written by another thread: read it through L2 (ld.global.cg)
            const double* lo = c >= n_inner ? obj[(unsigned)keys[c - n_inner]].lo : nodes[c].lo;
            const double* hi = c >= n_inner ? obj[(unsigned)keys[c - n_inner]].hi : nodes[c].hi;
            for (int a = 0; a < 3; ++a) {   // Aabb::from((&a, &b)), src/aabb.rs:95-114
                const double l = __ldcg(lo + a), h = __ldcg(hi + a);
                nd.lo[a] = side ? fmin(nd.lo[a], l) : l;
                nd.hi[a] = side ? fmax(nd.hi[a], h) : h;
            }
        }
        __threadfence();
        node = nd.parent;
    }
}

// the same padding as the host path of rc_api.cu (pad_lo / pad_hi): the fp32 boxes contain the f64 ones
RT_D float lbvh_pad_lo(double v, double ext) { return (float)(v - (1e-6 * ext + 1e-6 * fabs(v) + 1e-30)); }
RT_D float lbvh_pad_hi(double v, double ext) { return (float)(v + (1e-6 * ext + 1e-6 * fabs(v) + 1e-30)); }

// One thread per node of the tree (inner nodes 0..n-2, then leaves).  Pre-order index = sum along the
// path from the root: +1 for a left step, +1 + size(left sibling's subtree) for a right step, where a
// subtree over c leaves has 2c - 1 nodes.  out_order[sorted leaf k] = object index.
__global__ void lbvh_emit(const unsigned long long* __restrict__ keys, const LbvhObject* __restrict__ obj, int n,
                          const LbvhNode* __restrict__ nodes, const int* __restrict__ leaf_parent, const int* __restrict__ obj_first_new,
                          DevNode* __restrict__ out, DevNodeD* __restrict__ out_d, int* __restrict__ preorder_of) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_inner = n - 1;
    if (id >= 2 * n - 1) return;
    const bool leaf = id >= n_inner;
    int child = id, node = leaf ? (n > 1 ? leaf_parent[id - n_inner] : -1) : nodes[id].parent;
    int pre = 0;
    while (node >= 0) {
        const LbvhNode& nd = nodes[node];
        if (nd.right == child) {
            const int l = nd.left;
            const int leaves = l >= n_inner ? 1 : nodes[l].last - nodes[l].first + 1;
            pre += 2 * leaves;          // 1 + (2 leaves - 1)
        } else pre += 1;
        child = node;
        node = nd.parent;
    }
    const int my_leaves = leaf ? 1 : nodes[id].last - nodes[id].first + 1;
    const int skip = pre + 2 * my_leaves - 1;
    const double* lo;
    const double* hi;
    int leaf_field = -1;
    if (leaf) {
        const int o = (int)(unsigned)keys[id - n_inner];
        lo = obj[o].lo; hi = obj[o].hi;
        leaf_field = obj_first_new[id - n_inner] | (obj[o].count << 24);
    } else { lo = nodes[id].lo; hi = nodes[id].hi; }
    double ext = 0.0;
    for (int a = 0; a < 3; ++a) ext = fmax(ext, hi[a] - lo[a]);
    out[pre].lo = make_float4(lbvh_pad_lo(lo[0], ext), lbvh_pad_lo(lo[1], ext), lbvh_pad_lo(lo[2], ext), __int_as_float(skip));
    out[pre].hi = make_float4(lbvh_pad_hi(hi[0], ext), lbvh_pad_hi(hi[1], ext), lbvh_pad_hi(hi[2], ext), __int_as_float(leaf_field));
    for (int a = 0; a < 3; ++a) { out_d[pre].lo[a] = lo[a]; out_d[pre].hi[a] = hi[a]; }
    out_d[pre].skip = skip;
    out_d[pre].leaf = leaf_field;
    preorder_of[id] = pre;
}

// exclusive prefix sum of the sorted objects' primitive counts (single block; N objects <= a few 10^5)
__global__ void __launch_bounds__(1024) lbvh_prim_offsets(const unsigned long long* __restrict__ keys, const LbvhObject* __restrict__ obj, int n,
                                                          int* __restrict__ first_new) {
    __shared__ int carry;
    __shared__ int warp_sums[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int k = base + threadIdx.x;
        const int v = k < n ? obj[(unsigned)keys[k]].count : 0;
        int x = v;
        for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, off); if ((threadIdx.x & 31) >= off) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_sums[threadIdx.x];
            for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, off); if (threadIdx.x >= off) w += y; }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
        if (k < n) first_new[k] = carry + before + x - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += before + x;
        __syncthreads();
    }
}

// new primitive index -> old primitive index
__global__ void lbvh_prim_order(const unsigned long long* __restrict__ keys, const LbvhObject* __restrict__ obj, int n,
                                const int* __restrict__ first_new, int* __restrict__ order) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const LbvhObject& o = obj[(unsigned)keys[k]];
    for (int q = 0; q < o.count; ++q) order[first_new[k] + q] = o.first + q;
}

template <typename T>
__global__ void lbvh_gather(const T* __restrict__ src, T* __restrict__ dst, const int* __restrict__ order, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[order[i]];
}

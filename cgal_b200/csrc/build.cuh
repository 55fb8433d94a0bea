// build.cuh — LBVH construction kernels (replace AABB_tree::build / expand, AABB_tree/include/CGAL/AABB_tree.h:787-853,
// and Compute_bbox / Split_primitives, AABB_tree/include/CGAL/AABB_traits_3.h:596-654):
//   gather_mesh_kernel   indexed Surface_mesh (points + per-face vertex indices) -> 9 doubles per primitive, + the scene bounds
//   scene_bounds_kernel  bounding box of all primitives (exact: min/max only), one atomic per block and bound (after set_triangles)
//   morton_kernel        63-bit Morton key (21 bits per axis) of each primitive's box centre in the scene box, packed with the primitive index
//                        into one sort word; the digit histograms of the sort are accumulated on the way
//   [radix_sort.cuh]     onesweep sort of the words (keys only), or of (key, index) pairs on the fall-back path
//   build_local_kernel   FUSED: leaf records (TriRec + binary32 copies) in Morton order, and topology + boxes of every node whose leaf range lies
//                        inside a block of 256 leaves, in one bottom-up pass without searches (Karras' numbering: the same tree as the classic path)
//   hierarchy_kernel     Karras 2012 radix-tree topology by searches: for the nodes that span blocks (fused path) or for all nodes (classic path)
//   refit_global         boxes of the spanning nodes, one arrival flag per node
//   refit_local          the classic path's block-local refit (fall-back of a full frontier list; b200aabb_refit)
// A primitive's exact binary64 box is rounded OUTWARD to binary32 once; min/max are exact and commute with the monotone rounding, so every
// node box is bit-equal to the outward rounding of the exact union of its primitives' boxes.
#pragma once
#include "bvh.cuh"
#include "radix_sort.cuh"

namespace b2 {

// order-preserving map double -> uint64 so that atomicMin/atomicMax on integers order like the doubles
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double from_ordered_bits(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}

// block reduction of per-thread bounds + one atomic per block and bound (shared by scene_bounds_kernel and gather_mesh_kernel)
__device__ __forceinline__ void bounds_block_reduce(double* lo, double* hi, unsigned long long* __restrict__ bounds) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (int off = 16; off > 0; off >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], off));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], off));
        }
    __shared__ double s[8][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) for (int k = 0; k < 3; ++k) { s[warp][k] = lo[k]; s[warp][3 + k] = hi[k]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = s[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fmin(v, s[w][threadIdx.x]) : fmax(v, s[w][threadIdx.x]);
        if (threadIdx.x < 3) atomicMin(&bounds[threadIdx.x], ordered_bits(v));
        else atomicMax(&bounds[threadIdx.x], ordered_bits(v));
    }
}

// indexed Surface_mesh -> 9 doubles per primitive; the scene bounds come out of the same pass (the coordinates are in registers anyway),
// so that a build after b200aabb_set_mesh does not read the primitives a second time for them
__global__ void __launch_bounds__(256) gather_mesh_kernel(const double* __restrict__ verts, const uint32_t* __restrict__ faces, uint64_t nf, double* __restrict__ tris,
                                                          unsigned long long* __restrict__ bounds) {
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += stride) {
        double* out = tris + 9 * i;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double* v = verts + 3 * (uint64_t)faces[3 * i + k];
#pragma unroll
            for (int a = 0; a < 3; ++a) { const double c = v[a]; out[3 * k + a] = c; lo[a] = fmin(lo[a], c); hi[a] = fmax(hi[a], c); }
        }
    }
    bounds_block_reduce(lo, hi, bounds);
}

// bounds[0..2] = ordered_bits(min), bounds[3..5] = ordered_bits(max); initialised to ~0 / 0 by the host
__global__ void __launch_bounds__(256) scene_bounds_kernel(const double* __restrict__ tris, uint64_t n, unsigned long long* __restrict__ bounds) {
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double* t = tris + 9 * i;
#pragma unroll
        for (int k = 0; k < 9; ++k) { const double v = t[k]; lo[k % 3] = fmin(lo[k % 3], v); hi[k % 3] = fmax(hi[k % 3], v); }
    }
    bounds_block_reduce(lo, hi, bounds);
}

// bbox() of the tree (AABB_tree.h:204-210) = the exact scene bounds, decoded from the ordered-integer atomics
__global__ void root_box_kernel(const unsigned long long* __restrict__ bounds, double* __restrict__ root_box) {
    if (threadIdx.x < 6) root_box[threadIdx.x] = from_ordered_bits(bounds[threadIdx.x]);
}

B2_HD uint64_t spread21(uint64_t x) {   // insert two zero bits after each of the low 21 bits
    x &= 0x1FFFFFull;
    x = (x | (x << 32)) & 0x1F00000000FFFFull;
    x = (x | (x << 16)) & 0x1F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
B2_HD uint64_t morton63(double cx, double cy, double cz, const double* lo, const double* scale) {
    const double fx = (cx - lo[0]) * scale[0], fy = (cy - lo[1]) * scale[1], fz = (cz - lo[2]) * scale[2];
    const uint64_t qx = (uint64_t)fmin(fmax(fx, 0.0), 2097151.0);
    const uint64_t qy = (uint64_t)fmin(fmax(fy, 0.0), 2097151.0);
    const uint64_t qz = (uint64_t)fmin(fmax(fz, 0.0), 2097151.0);
    return (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
}

// quantisation frame of the Morton keys: ISOTROPIC cells (one scale, from the longest extent of the scene box), so that
// a flat scene (terrain) is not split along its thin axis at the top of the tree
B2_HD void morton_frame(double lx, double ly, double lz, double hx, double hy, double hz, double* lo, double* scale, const double* weight = nullptr) {
    lo[0] = lx; lo[1] = ly; lo[2] = lz;
    const double ext = fmax(hx - lx, fmax(hy - ly, hz - lz));
    const double sc = ext > 0.0 ? 2097152.0 / ext : 0.0;
    scale[0] = sc * (weight ? weight[0] : 1.0); scale[1] = sc * (weight ? weight[1] : 1.0); scale[2] = sc * (weight ? weight[2] : 1.0);
}

// PACKED sort word of primitive `index` (index < 2^ib): the top 64 - ib bits of its 63-bit Morton key above its index.  Sorting the words by
// their key bits alone (stable) orders equal keys by index, so one 8-byte word per primitive replaces the (key, index) pair: a third less
// traffic per pass and — 40 key bits for 16.8 M primitives — five passes instead of eight.  The words are all different, so the radix tree
// over them needs no tie-break, and it IS the tree of the (63-bit key, index) pairs whenever no two primitives agree in the kept key bits
// (hierarchy_kernel counts the runs of four or more that do; the build falls back to full keys when they are not rare).
B2_HD uint64_t pack_sort_word(uint64_t key63, uint64_t index, int ib) { return (((key63 << 1) >> ib) << ib) | index; }
B2_HD int index_bits_for(uint64_t n) { int b = 1; while (b < 32 && (1ull << b) < n) ++b; return b; }

// PACKED: keys[i] = pack_sort_word (vals unused) and the digit histograms of the sort are accumulated on the way (rs_hist_*);
// otherwise keys[i] = the 63-bit key, vals[i] = i.
template <bool PACKED>
__global__ void __launch_bounds__(256) morton_kernel(const double* __restrict__ tris, uint64_t n, const unsigned long long* __restrict__ bounds,
                                                     uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int ib, int npasses, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[PACKED ? 8 * kRsBins : 1];
    if (PACKED) rs_hist_begin(sh);
    double lo[3], scale[3];
    morton_frame(from_ordered_bits(bounds[0]), from_ordered_bits(bounds[1]), from_ordered_bits(bounds[2]),
                 from_ordered_bits(bounds[3]), from_ordered_bits(bounds[4]), from_ordered_bits(bounds[5]), lo, scale);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double* t = tris + 9 * i;
        double c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double a = t[k], b = t[3 + k], d = t[6 + k];
            c[k] = 0.5 * (fmin(a, fmin(b, d)) + fmax(a, fmax(b, d)));
        }
        const uint64_t key = morton63(c[0], c[1], c[2], lo, scale);
        if (PACKED) {
            const uint64_t w = pack_sort_word(key, i, ib);
            keys[i] = w;
            rs_hist_add(sh, w, ib, npasses);
        } else {
            keys[i] = key;
            vals[i] = (uint32_t)i;
        }
    }
    if (PACKED) rs_hist_end(sh, hist, npasses);
}

// (perm == nullptr: the sorted PACKED words carry the primitive index in their low ib bits)
__global__ void __launch_bounds__(256) emit_leaves_kernel(const double* __restrict__ tris, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ perm,
                                                          const uint64_t* __restrict__ words, int ib, uint64_t n, TriRec* __restrict__ out, TriF* __restrict__ outf) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t src = perm ? perm[k] : (uint32_t)(words[k] & ((1ull << ib) - 1ull));
    const double* t = tris + 9 * (uint64_t)src;
    TriRec r;
#pragma unroll
    for (int j = 0; j < 9; ++j) r.v[j] = t[j];
    r.id = ids ? ids[src] : src;
    r.pad[0] = src;                                              // position of the primitive in the inserted order (b200aabb_refit re-reads it there)
#pragma unroll
    for (int j = 1; j < 5; ++j) r.pad[j] = 0;
    out[k] = r;
    TriF f;
#pragma unroll
    for (int j = 0; j < 9; ++j) f.v[j] = (float)t[j];
#pragma unroll
    for (int j = 0; j < 7; ++j) f.pad[j] = 0;
    outf[k] = f;
}

// b200aabb_refit: the primitives have moved (same count, same order): rewrite the coordinates of the leaf records in place, topology kept
__global__ void __launch_bounds__(256) refresh_leaves_kernel(const double* __restrict__ tris, uint64_t n, TriRec* __restrict__ out, TriF* __restrict__ outf) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double* t = tris + 9 * (uint64_t)out[k].pad[0];
#pragma unroll
    for (int j = 0; j < 9; ++j) { const double c = t[j]; out[k].v[j] = c; outf[k].v[j] = (float)c; }
}

// length of the common prefix of the (key, index) pairs i and j; -1 outside [0, n)
B2_HD int karras_delta(const uint64_t* __restrict__ keys, int64_t n, int64_t i, int64_t j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t x = keys[i] ^ keys[j];
#if defined(__CUDA_ARCH__)
    if (x) return __clzll((long long)x);
    return 64 + __clzll((long long)((uint64_t)i ^ (uint64_t)j));
#else
    if (x) return __builtin_clzll(x);
    return 64 + __builtin_clzll((uint64_t)i ^ (uint64_t)j);
#endif
}

// topology of internal node i (Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees", 2012)
B2_HD void karras_node(const uint64_t* __restrict__ keys, int64_t n, int64_t i, uint32_t& left, uint32_t& right, uint32_t* other_end = nullptr) {
    const int d = karras_delta(keys, n, i, i + 1) - karras_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = karras_delta(keys, n, i, i - d);
    int64_t lmax = 2;
    while (karras_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int64_t l = 0;
    for (int64_t t = lmax >> 1; t >= 1; t >>= 1)
        if (karras_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int64_t j = i + l * d;
    const int dnode = karras_delta(keys, n, i, j);
    int64_t s = 0;
    for (int64_t t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (karras_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int64_t gamma = i + s * d + (d < 0 ? -1 : 0);
    const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
    if (other_end) *other_end = (uint32_t)j;
    left = lo == gamma ? (kLeafBit | (uint32_t)gamma) : (uint32_t)gamma;
    right = hi == gamma + 1 ? (kLeafBit | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
}

// ib > 0: the keys are PACKED words; *dups += positions that start a run of four words agreeing in all their key bits (see pack_sort_word)
// list != nullptr: only the nodes list[0 .. min(*list_count, list_cap)) (those build_local_kernel left: they span blocks of leaves)
__global__ void __launch_bounds__(256) hierarchy_kernel(const uint64_t* __restrict__ keys, uint64_t n, BvhNode* __restrict__ nodes, uint32_t* __restrict__ leaf_parent,
                                                        int ib, unsigned long long* __restrict__ dups, const uint32_t* __restrict__ list,
                                                        const unsigned* __restrict__ list_count, unsigned list_cap) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (list) {
        const unsigned cnt = *list_count;
        if (i >= (cnt < list_cap ? cnt : list_cap)) return;
        i = list[i];
    }
    if (ib > 0) {
        // (runs of FOUR: pairs of primitives with equal key bits — the two triangles of a terrain quad have the same box centre in x and y —
        //  are siblings in either tree; what would hurt are long runs, which the index bits split without regard to geometry)
        const bool dup = i + 3 < n && (((keys[i] ^ keys[i + 1]) | (keys[i] ^ keys[i + 2]) | (keys[i] ^ keys[i + 3])) >> ib) == 0;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, dup);
        if (m && (threadIdx.x & 31) == 0) atomicAdd(dups, (unsigned long long)__popc(m));
    }
    if (i + 1 >= n) return;
    uint32_t l, r, j;
    karras_node(keys, (int64_t)n, (int64_t)i, l, r, &j);
    BvhNode* nd = nodes + i;
    nd->child0 = l; nd->child1 = r; nd->visit = j;           // the node covers the leaves [min(i,j), max(i,j)]; refit_kernel reads it
    if (i == 0) nd->parent = kNone;
    if (l & kLeafBit) leaf_parent[l & ~kLeafBit] = (uint32_t)i; else nodes[l].parent = (uint32_t)i;
    if (r & kLeafBit) leaf_parent[r & ~kLeafBit] = (uint32_t)i; else nodes[r].parent = (uint32_t)i;
}

// Refit, bottom-up: one thread per leaf; the second thread to arrive at a node continues upwards with the union of both child boxes.
// refit_local_kernel: a block owns kRefitBlock consecutive leaves.  The nodes whose leaf range lies inside the block (node i covers
// [min(i, visit), max(i, visit)], hierarchy_kernel) are only ever reached by this block's threads and have indices inside the
// block's range: their links are staged in shared memory and their arrival counters and child boxes live there too, so a level
// costs shared-memory latency instead of a store + fence + L2 atomic + L2 load.  A thread that reaches a node spanning blocks
// parks its box in the frontier list and ends, so that no block stays resident for the slow levels.
// refit_global_kernel: one thread per frontier entry finishes the climb with the global protocol (arrival flag = atomicExch on
// `visit`).  On Morton-ordered input a block hands over a few entries; the list holds 16 per block and a thread that finds it full
// (a degenerate, caterpillar-like tree) simply finishes its climb in place.
constexpr int kRefitBlock = 256;
constexpr int kRefitFrontierPerBlock = 16;
struct __align__(16) RefitEntry { float lo[3], hi[3]; uint32_t node, from; };
static_assert(sizeof(RefitEntry) == 32, "frontier entry");

// the climb over nodes that span blocks: global arrival flag (atomicExch on `visit`, which held the range end until then)
__device__ __forceinline__ void refit_climb_global(BvhNode* nodes, uint32_t p, uint32_t me, float* lo, float* hi) {
    while (true) {
        BvhNode* nd = nodes + p;
        const int side = (nd->child0 == me) ? 0 : 1;
        float* mylo = side ? nd->lo1 : nd->lo0; float* myhi = side ? nd->hi1 : nd->hi0;
#pragma unroll
        for (int a = 0; a < 3; ++a) { mylo[a] = lo[a]; myhi[a] = hi[a]; }
        __threadfence();
        // arrival flag = bit 31 of `visit` (the low bits keep the other end of the node's leaf range): each arrival toggles it, the first one
        // finds it clear, the second one set — and leaves it clear again, so a later refit (b200aabb_refit) finds the node as the build left it
        if ((atomicXor(&nd->visit, 0x80000000u) & 0x80000000u) == 0) return;      // first to arrive: the sibling will carry on
        __threadfence();
        // the sibling box was written by another SM: read it through L2 (this SM's L1 may hold the line from the child[] read)
        const float* olo = side ? nd->lo0 : nd->lo1; const float* ohi = side ? nd->hi0 : nd->hi1;
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], __ldcg(olo + a)); hi[a] = fmaxf(hi[a], __ldcg(ohi + a)); }
        me = p;
        p = nd->parent;
        if (p == kNone) return;
    }
}

// EMIT: the kernel also WRITES the leaf records (the work of emit_leaves_kernel: gather the primitive `src`, store the 96-byte record and its
// binary32 copy) before it climbs, so that the build reads every primitive once instead of writing the records and reading them back
// (1.6 GB at 16.8 M primitives); the stores of one thread overlap the climbs of the others.
struct EmitArgs { const double* soup; const uint32_t* ids; const uint32_t* perm; const uint64_t* words; int ib; TriF* outf; };
template <bool EMIT>
__global__ void __launch_bounds__(kRefitBlock) refit_local_kernel(TriRec* __restrict__ tris, uint64_t n, BvhNode* nodes, const uint32_t* __restrict__ leaf_parent,
                                                                  RefitEntry* __restrict__ frontier, unsigned* __restrict__ frontier_count, unsigned frontier_cap, EmitArgs em) {
    __shared__ uint4 s_link[kRefitBlock];          // child[0], child[1], parent, other end of the range, of node (block base + t)
    __shared__ unsigned s_cnt[kRefitBlock];
    __shared__ float s_box[kRefitBlock][2][6];     // child boxes of the block's own nodes
    const uint64_t base = (uint64_t)blockIdx.x * kRefitBlock;
    const uint64_t k = base + threadIdx.x;
    const uint64_t last_leaf = base + kRefitBlock - 1 < n - 1 ? base + kRefitBlock - 1 : n - 1;
    s_cnt[threadIdx.x] = 0;
    if (k + 1 < n) {
        const uint2 a = *reinterpret_cast<const uint2*>(&nodes[k].child0), b = *reinterpret_cast<const uint2*>(&nodes[k].child1);   // (child0, parent), (child1, visit)
        s_link[threadIdx.x] = make_uint4(a.x, b.x, a.y, b.y);
    }
    __syncthreads();
    if (k >= n) return;
    float lo[3], hi[3];
    if (EMIT) {
        const uint32_t src = em.perm ? em.perm[k] : (uint32_t)(em.words[k] & ((1ull << em.ib) - 1ull));
        const double* t = em.soup + 9 * (uint64_t)src;
        TriRec r;
#pragma unroll
        for (int j = 0; j < 9; ++j) r.v[j] = t[j];
        r.id = em.ids ? em.ids[src] : src;
        r.pad[0] = src;                                          // position of the primitive in the inserted order (b200aabb_refit re-reads it there)
#pragma unroll
        for (int j = 1; j < 5; ++j) r.pad[j] = 0;
        tris[k] = r;
        TriF f;
#pragma unroll
        for (int j = 0; j < 9; ++j) f.v[j] = (float)r.v[j];
#pragma unroll
        for (int j = 0; j < 7; ++j) f.pad[j] = 0;
        em.outf[k] = f;
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = float_below(fmin(r.v[a], fmin(r.v[3 + a], r.v[6 + a]))); hi[a] = float_above(fmax(r.v[a], fmax(r.v[3 + a], r.v[6 + a]))); }
    } else {
        const double* v = tris[k].v;
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = float_below(fmin(v[a], fmin(v[3 + a], v[6 + a]))); hi[a] = float_above(fmax(v[a], fmax(v[3 + a], v[6 + a]))); }
    }
    if (n == 1) return;
    uint32_t me = kLeafBit | (uint32_t)k;
    uint32_t p = leaf_parent[k];
    while (true) {
        bool local = p >= base && p <= last_leaf;
        uint4 link = make_uint4(0, 0, 0, 0);
        if (local) {
            link = s_link[p - base];
            link.w &= 0x7FFFFFFFu;
            const uint32_t a = p < link.w ? p : link.w, b = p < link.w ? link.w : p;
            local = a >= base && b <= last_leaf;
        }
        if (!local) {                               // a node that spans blocks: hand over to refit_global_kernel
            const unsigned slot = atomicAdd(frontier_count, 1u);
            if (slot >= frontier_cap) { refit_climb_global(nodes, p, me, lo, hi); return; }   // (a degenerate tree: finish here, slowly)
            RefitEntry e;
#pragma unroll
            for (int a = 0; a < 3; ++a) { e.lo[a] = lo[a]; e.hi[a] = hi[a]; }
            e.node = p; e.from = me;
            frontier[slot] = e;
            return;
        }
        BvhNode* nd = nodes + p;
        const unsigned slot = p - (uint32_t)base;
        const int side = (link.x == me) ? 0 : 1;
#pragma unroll
        for (int a = 0; a < 3; ++a) { s_box[slot][side][a] = lo[a]; s_box[slot][side][3 + a] = hi[a]; }
        __threadfence_block();
        if (atomicAdd(&s_cnt[slot], 1u) == 0) return;            // first to arrive: the sibling will carry on
        __threadfence_block();
        float olo[3], ohi[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { olo[a] = ((volatile float*)s_box[slot][1 - side])[a]; ohi[a] = ((volatile float*)s_box[slot][1 - side])[3 + a]; }
        // the second arrival writes the node at once: each 32-byte half = a child's box + the links already staged in shared memory, four 128-bit stores
        {
            const float* l0 = side ? olo : lo; const float* h0 = side ? ohi : hi;     // child 0
            const float* l1 = side ? lo : olo; const float* h1 = side ? hi : ohi;     // child 1
            float4* dst = reinterpret_cast<float4*>(nd);
            dst[0] = make_float4(l0[0], l0[1], l0[2], h0[0]);
            dst[1] = make_float4(h0[1], h0[2], __uint_as_float(link.x), __uint_as_float(link.z));
            dst[2] = make_float4(l1[0], l1[1], l1[2], h1[0]);
            dst[3] = make_float4(h1[1], h1[2], __uint_as_float(link.y), __uint_as_float(link.w));
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], olo[a]); hi[a] = fmaxf(hi[a], ohi[a]); }
        me = p;
        p = link.z;
        if (p == kNone) return;
    }
}

// (an entry of build_local_kernel does not know its parent's index yet: e.node == kNone, and the parent link written since by hierarchy_kernel is read here)
__global__ void __launch_bounds__(256) refit_global_kernel(BvhNode* nodes, const uint32_t* __restrict__ leaf_parent, const RefitEntry* __restrict__ frontier,
                                                           const unsigned* __restrict__ frontier_count, unsigned cap) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned cnt = *frontier_count;
    if (i >= (cnt < cap ? cnt : cap)) return;
    const RefitEntry e = frontier[i];
    float lo[3] = {e.lo[0], e.lo[1], e.lo[2]}, hi[3] = {e.hi[0], e.hi[1], e.hi[2]};
    const uint32_t p = e.node != kNone ? e.node : ((e.from & kLeafBit) ? leaf_parent[e.from & ~kLeafBit] : nodes[e.from].parent);
    refit_climb_global(nodes, p, e.from, lo, hi);
}

// ---------------------------------------------------------------------------------------------------------------
// build_local_kernel: leaf records + radix-tree TOPOLOGY + boxes of every node whose leaf range lies inside one block of kRefitBlock leaves,
// in ONE bottom-up pass (no searches: after Apetrei, "Fast and Simple Agglomerative LBVH Construction", 2014, but with Karras' node numbering so
// that the tree is, bit for bit, the one hierarchy_kernel + refit_local_kernel build).
//   A thread starts at its leaf with the range [l, r] = [k, k] and repeatedly joins its range with the sibling's: the parent of [l, r] splits
//   at s = r (we are its left child) when the pairs r, r+1 share a longer prefix than the pairs l-1, l, else at s = l-1 (right child).  The two
//   children meet at the shared-memory slot of s; the second to arrive knows the parent's range [L, R], and with it
//     * whether THAT node is a left child itself (same rule), hence its index in Karras' numbering: R if it is a left child, L otherwise
//       (a left child [l, gamma] is node gamma, a right child [gamma+1, r] node gamma+1, the root node 0);
//     * both child references and boxes: it writes the node, tells the children their parent, and carries on.
//   The parent of a range that touches the block's border spans blocks when the border key shares the parent's prefix: decided from s alone
//   (both children must agree without knowing each other), by comparing the pair s with the pairs just outside the block.  Such a child is
//   parked in the frontier list; hierarchy_kernel then builds the few spanning nodes (listed in `rest`) with Karras' searches and
//   refit_global_kernel finishes the boxes.  A full frontier list (a degenerate tree) sets *overflow and the build takes the classic path.
// ---------------------------------------------------------------------------------------------------------------
constexpr size_t kSoupSlack = 16;       // bytes allocated past the last inserted primitive: build_local_kernel gathers a primitive with five 128-bit loads
struct __align__(16) LocalBuildShared {
    uint64_t key[kRefitBlock + 2];      // sort words / keys of the pairs base-1 .. base+kRefitBlock
    float4 arr[kRefitBlock][2][2];      // the child that arrived at the slot of split position base + t from the left [0] / right [1]: (lo.xyz, hi.x), (hi.y, hi.z, reference,
                                        // far end of its range).  16 KB: before the climb, the warps' staging area of the coalesced leaf-record stores
    unsigned cnt[kRefitBlock];          // arrivals at the slot
    short len[kRefitBlock + 1];         // len[p - (base-1)] = karras_delta(p, p+1), p = base-1 .. last_leaf; + kSpanFlag when the node splitting at p spans blocks
    unsigned char done[kRefitBlock];    // node base + t was built here
};
constexpr short kSpanFlag = 0x4000;
static_assert(offsetof(LocalBuildShared, arr) % 16 == 0, "staging area");

// prefix length of the pairs i and j (as karras_delta) read from the block's staged keys; positions outside [0, n) give -1
B2_HD int local_delta(const LocalBuildShared& S, int64_t base, int64_t n, int64_t i, int64_t j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t x = S.key[i - (base - 1)] ^ S.key[j - (base - 1)];
#if defined(__CUDA_ARCH__)
    if (x) return __clzll((long long)x);
    return 64 + __clzll((long long)((uint64_t)i ^ (uint64_t)j));
#else
    if (x) return __builtin_clzll(x);
    return 64 + __builtin_clzll((uint64_t)i ^ (uint64_t)j);
#endif
}

#if defined(__CUDA_ARCH__)
#define B2_ARRIVE(counter) atomicAdd(counter, 1u)
#define B2_FENCE_BLOCK() __threadfence_block()
#else
#define B2_ARRIVE(counter) ((*(counter))++)
#define B2_FENCE_BLOCK()
#endif

// len[] entry of split position p, with the flag that the node splitting there spans blocks (decided from p alone: both children must agree without knowing each
// other): the key just outside the block, on either side, shares the node's prefix
B2_HD short local_len_entry(const LocalBuildShared& S, int64_t base, int64_t last_leaf, int64_t n, int64_t p) {
    if (p < 0 || p + 1 >= n) return (short)(-1);
    const int plen = local_delta(S, base, n, p, p + 1);
    const bool spans = p < base || p >= last_leaf || local_delta(S, base, n, p, base - 1) >= plen || local_delta(S, base, n, p, last_leaf + 1) >= plen;
    return (short)(plen | (spans ? kSpanFlag : 0));
}
B2_HD int len_of(short entry) { return entry < 0 ? -1 : (entry & (kSpanFlag - 1)); }

// the climb of the thread of leaf k (box lo/hi) through the block's own nodes
B2_HD void build_local_climb(LocalBuildShared& S, int64_t base, int64_t last_leaf, int64_t n, int64_t k, float* lo, float* hi, BvhNode* nodes,
                             uint32_t* leaf_parent, RefitEntry* frontier, unsigned* frontier_count, unsigned frontier_cap, unsigned* overflow) {
    uint32_t me = kLeafBit | (uint32_t)k;
    int64_t l = k, r = k;
    bool left = l == 0 || (r != n - 1 && len_of(S.len[r - (base - 1)]) > len_of(S.len[l - 1 - (base - 1)]));
    while (true) {
        const int64_t s = left ? r : l - 1;                       // split position of the parent
        if (S.len[s - (base - 1)] & kSpanFlag) {                  // the parent spans blocks
#if defined(__CUDA_ARCH__)
            const unsigned slot = atomicAdd(frontier_count, 1u);
#else
            const unsigned slot = (*frontier_count)++;
#endif
            if (slot >= frontier_cap) { *overflow = 1u; return; }
            RefitEntry e;
            for (int a = 0; a < 3; ++a) { e.lo[a] = lo[a]; e.hi[a] = hi[a]; }
            e.node = kNone; e.from = me;
            frontier[slot] = e;
            return;
        }
        const unsigned slot = (unsigned)(s - base);
        const int side = left ? 0 : 1;
#if defined(__CUDA_ARCH__)
        S.arr[slot][side][0] = make_float4(lo[0], lo[1], lo[2], hi[0]);
        S.arr[slot][side][1] = make_float4(hi[1], hi[2], __uint_as_float(me), __uint_as_float((uint32_t)(left ? l : r)));
#else
        { float4 a0, a1; a0.x = lo[0]; a0.y = lo[1]; a0.z = lo[2]; a0.w = hi[0]; a1.x = hi[1]; a1.y = hi[2];
          const uint32_t end = (uint32_t)(left ? l : r); memcpy(&a1.z, &me, 4); memcpy(&a1.w, &end, 4); S.arr[slot][side][0] = a0; S.arr[slot][side][1] = a1; }
#endif
        B2_FENCE_BLOCK();
        if (B2_ARRIVE(&S.cnt[slot]) == 0) return;                 // first to arrive: the sibling will carry on
        B2_FENCE_BLOCK();
        const float4 o0 = S.arr[slot][1 - side][0], o1 = S.arr[slot][1 - side][1];       // (the fence and the atomic order these loads after the sibling's stores)
        const float olo[3] = {o0.x, o0.y, o0.z}, ohi[3] = {o0.w, o1.x, o1.y};
        uint32_t sib, sib_end;
#if defined(__CUDA_ARCH__)
        sib = __float_as_uint(o1.z); sib_end = __float_as_uint(o1.w);
#else
        memcpy(&sib, &o1.z, 4); memcpy(&sib_end, &o1.w, 4);
#endif
        const int64_t L = left ? l : (int64_t)sib_end, R = left ? (int64_t)sib_end : r;
        const uint32_t c0 = left ? me : sib, c1 = left ? sib : me;
        const bool root = L == 0 && R == n - 1;
        const bool nleft = !root && (L == 0 || (R != n - 1 && len_of(S.len[R - (base - 1)]) > len_of(S.len[L - 1 - (base - 1)])));
        const uint32_t idx = root ? 0u : (uint32_t)(nleft ? R : L);
        BvhNode* nd = nodes + idx;
        {
            const float* l0 = side ? olo : lo; const float* h0 = side ? ohi : hi;     // child 0
            const float* l1 = side ? lo : olo; const float* h1 = side ? hi : ohi;     // child 1
#if defined(__CUDA_ARCH__)
            float4* dst = reinterpret_cast<float4*>(nd);
            dst[0] = make_float4(l0[0], l0[1], l0[2], h0[0]);
            *reinterpret_cast<float2*>(&nd->hi0[1]) = make_float2(h0[1], h0[2]);      // (the parent link next to child0 is written by whoever builds the parent)
            nd->child0 = c0;
            dst[2] = make_float4(l1[0], l1[1], l1[2], h1[0]);
            dst[3] = make_float4(h1[1], h1[2], __uint_as_float(c1), __uint_as_float(idx == (uint32_t)L ? (uint32_t)R : (uint32_t)L));
#else
            for (int a = 0; a < 3; ++a) { nd->lo0[a] = l0[a]; nd->hi0[a] = h0[a]; nd->lo1[a] = l1[a]; nd->hi1[a] = h1[a]; }
            nd->child0 = c0; nd->child1 = c1; nd->visit = idx == (uint32_t)L ? (uint32_t)R : (uint32_t)L;
#endif
            if (root) nd->parent = kNone;
        }
        if (c0 & kLeafBit) leaf_parent[c0 & ~kLeafBit] = idx; else nodes[c0].parent = idx;
        if (c1 & kLeafBit) leaf_parent[c1 & ~kLeafBit] = idx; else nodes[c1].parent = idx;
        S.done[idx - (uint32_t)base] = 1;
        if (root) return;
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], olo[a]); hi[a] = fmaxf(hi[a], ohi[a]); }
        me = idx; l = L; r = R; left = nleft;
    }
}

#if defined(__CUDACC__)
// flags[0] += positions that start a run of four PACKED words with equal key bits (ib > 0; see hierarchy_kernel); flags[1] = frontier overflow
__global__ void __launch_bounds__(kRefitBlock, 6) build_local_kernel(const uint64_t* __restrict__ keys, TriRec* __restrict__ tris, uint64_t n, BvhNode* nodes, uint32_t* leaf_parent,
                                                                  RefitEntry* __restrict__ frontier, unsigned* __restrict__ frontier_count, unsigned frontier_cap,
                                                                  EmitArgs em, uint32_t* __restrict__ rest, unsigned* __restrict__ rest_count, unsigned long long* __restrict__ flags) {
    __shared__ LocalBuildShared S;
    const int64_t base = (int64_t)blockIdx.x * kRefitBlock, k = base + threadIdx.x;
    const int64_t last_leaf = base + kRefitBlock - 1 < (int64_t)n - 1 ? base + kRefitBlock - 1 : (int64_t)n - 1;
    S.cnt[threadIdx.x] = 0;
    S.done[threadIdx.x] = 0;
    if (k < (int64_t)n) S.key[threadIdx.x + 1] = keys[k];
    if (threadIdx.x == 0) { S.key[0] = base > 0 ? keys[base - 1] : 0; S.key[last_leaf - base + 2] = last_leaf + 1 < (int64_t)n ? keys[last_leaf + 1] : 0; }
    __syncthreads();
    if (k <= last_leaf) S.len[threadIdx.x + 1] = local_len_entry(S, base, last_leaf, (int64_t)n, k);
    if (threadIdx.x == 0) S.len[0] = local_len_entry(S, base, last_leaf, (int64_t)n, base - 1);
    if (em.ib > 0) {
        const bool dup = k + 3 < (int64_t)n && (((keys[k] ^ keys[k + 1]) | (keys[k] ^ keys[k + 2]) | (keys[k] ^ keys[k + 3])) >> em.ib) == 0;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, dup);
        if (m && (threadIdx.x & 31) == 0) atomicAdd(&flags[0], (unsigned long long)__popc(m));
    }
    __syncthreads();
    // Leaf records.  A thread gathers its primitive (72 bytes at 72 * src: five 128-bit loads from the 16-byte boundary below it) and the warp writes its
    // 32 records — 3 KB of TriRec, 2 KB of TriF, contiguous in Morton order — through its 2 KB staging area with fully coalesced 128-bit stores: the L1
    // data stage bounds this kernel (profiles/r2_08), and a thread storing its own 96-byte record costs 32 sectors per instruction.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t k0 = base + warp * 32;
    const int valid = (int64_t)n - k0 >= 32 ? 32 : ((int64_t)n - k0 > 0 ? (int)((int64_t)n - k0) : 0);
    double v[9];
    uint32_t src = 0, id = 0;
    if (lane < valid) {
        src = em.perm ? em.perm[k] : (uint32_t)(S.key[threadIdx.x + 1] & ((1ull << em.ib) - 1ull));
        const double* t = em.soup + 9 * (uint64_t)src;
        if ((reinterpret_cast<uintptr_t>(em.soup) & 15) == 0) {
            const double2* q = reinterpret_cast<const double2*>(t - (src & 1));      // 72 * src is 8 modulo 16 exactly when src is odd
            const double2 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3], q4 = q[4];     // (even src: q4.y lies past the primitive — kSoupSlack)
            const bool odd = src & 1;
            v[0] = odd ? q0.y : q0.x; v[1] = odd ? q1.x : q0.y; v[2] = odd ? q1.y : q1.x; v[3] = odd ? q2.x : q1.y; v[4] = odd ? q2.y : q2.x;
            v[5] = odd ? q3.x : q2.y; v[6] = odd ? q3.y : q3.x; v[7] = odd ? q4.x : q3.y; v[8] = odd ? q4.y : q4.x;
        } else {
#pragma unroll
            for (int j = 0; j < 9; ++j) v[j] = t[j];
        }
        id = em.ids ? em.ids[src] : src;
    }
    {
        float4* stage = &S.arr[0][0][0] + warp * 128;
        double2* stage2 = reinterpret_cast<double2*>(stage);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if ((lane >> 4) == half && lane < valid) {                                // 16 records of 96 bytes at a stride of 7 x 16 bytes (bank spread)
                double2* d = stage2 + (lane & 15) * 7;
                d[0] = make_double2(v[0], v[1]); d[1] = make_double2(v[2], v[3]); d[2] = make_double2(v[4], v[5]); d[3] = make_double2(v[6], v[7]);
                d[4] = make_double2(v[8], __longlong_as_double((long long)(((unsigned long long)src << 32) | id)));   // id, pad[0] = position in the inserted order
                d[5] = make_double2(0.0, 0.0);
            }
            __syncwarp();
            const int cnt = valid - 16 * half >= 16 ? 16 : (valid - 16 * half > 0 ? valid - 16 * half : 0);
            float4* dst = reinterpret_cast<float4*>(tris + k0 + 16 * half);
            for (int e = lane; e < cnt * 6; e += 32) dst[e] = stage[(e / 6) * 7 + e % 6];
            __syncwarp();
        }
        if (lane < valid) {                                                           // 32 records of 64 bytes, 16-byte parts swizzled over the banks
            const int sw = (lane >> 1) & 3;
            stage[4 * lane + (0 ^ sw)] = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
            stage[4 * lane + (1 ^ sw)] = make_float4((float)v[4], (float)v[5], (float)v[6], (float)v[7]);
            stage[4 * lane + (2 ^ sw)] = make_float4((float)v[8], 0.0f, 0.0f, 0.0f);
            stage[4 * lane + (3 ^ sw)] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        __syncwarp();
        float4* dstf = reinterpret_cast<float4*>(em.outf + k0);
        for (int e = lane; e < valid * 4; e += 32) { const int i = e >> 2, j = e & 3; dstf[e] = stage[4 * i + (j ^ ((i >> 1) & 3))]; }
    }
    __syncthreads();                                                                  // the staging area becomes the arrival records
    if (lane < valid) {
        float lo[3], hi[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = float_below(fmin(v[a], fmin(v[3 + a], v[6 + a]))); hi[a] = float_above(fmax(v[a], fmax(v[3 + a], v[6 + a]))); }
        if (n > 1) build_local_climb(S, base, last_leaf, (int64_t)n, k, lo, hi, nodes, leaf_parent, frontier, frontier_count, frontier_cap, reinterpret_cast<unsigned*>(&flags[1]));
    }
    __syncthreads();
    if (k < (int64_t)n - 1 && !S.done[threadIdx.x]) {            // node k spans blocks: hierarchy_kernel builds it from the list
        const unsigned slot = atomicAdd(rest_count, 1u);
        if (slot < frontier_cap) rest[slot] = (uint32_t)k;
    }
}
#endif

}  // namespace b2

// build.cuh — LBVH construction kernels (replace AABB_tree::build / expand, AABB_tree/include/CGAL/AABB_tree.h:787-853,
// and Compute_bbox / Split_primitives, AABB_tree/include/CGAL/AABB_traits_3.h:596-654):
//   gather_mesh_kernel   indexed Surface_mesh (points + per-face vertex indices) -> 9 doubles per primitive
//   scene_bounds_kernel  bounding box of all primitives (exact: min/max only), one atomic per block and bound
//   morton_kernel        63-bit Morton key (21 bits per axis) of each primitive's box centre in the scene box
//   [radix_sort.cuh]     onesweep sort of (key, primitive index)
//   emit_leaves_kernel   TriRec records (and their binary32 copies, TriF) in Morton order
//   hierarchy_kernel     Karras 2012 radix-tree topology, one thread per internal node
//   refit_local/global   bottom-up boxes with one arrival counter per node (shared memory inside a block of 256 leaves, global above).  A primitive's exact binary64 box is
//                        rounded OUTWARD to binary32 once; min/max are exact and commute with the monotone rounding, so
//                        every node box is bit-equal to the outward rounding of the exact union of its primitives' boxes
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
B2_HD void morton_frame(double lx, double ly, double lz, double hx, double hy, double hz, double* lo, double* scale) {
    lo[0] = lx; lo[1] = ly; lo[2] = lz;
    const double ext = fmax(hx - lx, fmax(hy - ly, hz - lz));
    const double sc = ext > 0.0 ? 2097152.0 / ext : 0.0;
    scale[0] = sc; scale[1] = sc; scale[2] = sc;
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
__global__ void __launch_bounds__(256) hierarchy_kernel(const uint64_t* __restrict__ keys, uint64_t n, BvhNode* __restrict__ nodes, uint32_t* __restrict__ leaf_parent,
                                                        int ib, unsigned long long* __restrict__ dups) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
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

__global__ void __launch_bounds__(256) refit_global_kernel(BvhNode* nodes, const RefitEntry* __restrict__ frontier, const unsigned* __restrict__ frontier_count, unsigned cap) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned cnt = *frontier_count;
    if (i >= (cnt < cap ? cnt : cap)) return;
    const RefitEntry e = frontier[i];
    float lo[3] = {e.lo[0], e.lo[1], e.lo[2]}, hi[3] = {e.hi[0], e.hi[1], e.hi[2]};
    refit_climb_global(nodes, e.node, e.from, lo, hi);
}

}  // namespace b2

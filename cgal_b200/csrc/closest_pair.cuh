// closest_pair.cuh — closest_point_and_primitive(Point_3), second-generation warp-cooperative kernel: LANE PAIRS walk the tree.
//
// Replaces the same reference path as closest_warp.cuh (AABB_tree.h:987-1055, Projection_traits AABB_traversal_traits.h:277-327,
// AABB_traits_3.h:715-733) and returns the same records bit for bit: the winner is the candidate (reference-rounded projection, or the
// primitive's reference point when that is nearer: geom.cuh::closest_candidate) with the smallest EXACT squared distance, ties ->
// smallest primitive id + FLAG_TIE.  Only the scheduling differs.
//
// Why (profiles/r2_01): closest_warp.cuh's shared work pool keeps all lanes busy, but it pays ~160 warp instructions per node step (two ballot
// pushes per child class, pops and slot-state gathers from shared memory with bank conflicts) and two L1 wavefronts per node and lane (every
// lane fetches both 32-byte sectors of its own node): the L1 data stage ran at 74 % of its one-wavefront-per-cycle peak, 595 wavefronts per
// query, and the issue slots at 51 %.  Here
//   * a node visit is shared by a PAIR of lanes: each lane loads ONE 32-byte half of the node (one child's box + reference, bvh.cuh) — the
//     pair's two loads fall into one 64-byte line, i.e. ONE wavefront per node — tests its child in binary32 (directed rounding) and
//     swaps the outcome with its partner by one shuffle;
//   * every pair (a "worker", 16 per warp) walks ITS OWN query depth-first, nearer child first, with a private stack (24 entries in shared
//     memory, laid out so that the 16 workers hit 16 different banks; deeper entries spill to a per-warp area in global memory): the next
//     node stays in a register, a push is one predicated store, there are no ballots and no slot-state gathers in the node loop;
//   * leaves still go through the shared three-stage pipeline of closest_warp.cuh (rigorous binary32 lower bound -> binary32 approximate closest
//     point with a rigorous upper bound that shrinks the ball -> the reference's binary64 projection + exact Compare_distance_3), 32 pairs
//     per step whichever queries they belong to;
//   * a warp owns 32 query SLOTS, two per worker: when a worker's traversal ends it starts on its other slot at once while the last leaf
//     pairs of the finished query drain through the pipeline (a slot completes when as many pairs have retired as were pushed);
//   * queries are handed out in coherence order in chunks of 64 per warp, so the queries a warp works on are neighbours: a new query's first
//     ball is the distance to the reference point (first vertex) of a primitive that won a recent query of the same warp — rigorous, because
//     a primitive's candidate is never farther than its reference point — which prunes the far siblings of the root path from the start.
// Depth: a Karras tree over (63-bit key, 30-bit index) pairs is at most 94 levels deep (every level lengthens the common prefix), and a
// depth-first walk stacks at most one sibling per level: kPpStackCap = 96 can not overflow (checked all the same and reported).
//
// The body is written against a small warp abstraction (WarpCtx) so that tests/hostcheck can run the SAME code on the CPU with 32
// cooperatively scheduled fibers (test infrastructure; the product library has no host execution path).
#pragma once
#include "closest_warp.cuh"

namespace b2 {

constexpr int kPpWorkers = 16;                  // lane pairs per warp
constexpr int kPpSlots = 32;                    // query slots per warp: worker w owns slots 2w and 2w+1
constexpr int kPpStackSm = 24;                  // stack entries per worker held in shared memory
constexpr int kPpStackCap = 96;                 // >= 94 = the deepest possible tree (see above)
constexpr int kPpSpillPerWarp = kPpWorkers * (kPpStackCap - kPpStackSm);
constexpr int kPpLeafCap = 160;                 // <= 128 waiting + at most 32 added by one node step
constexpr int kPpMidCap = 192;                  // < 32 waiting + at most 160 added by one pre-test-1 phase
constexpr int kPpExactCap = 160;                // pre-test 2 stops handing over when fewer than 32 places are left
#ifndef B2_PP_CHUNK
#define B2_PP_CHUNK 64
#endif
constexpr int kPpChunk = B2_PP_CHUNK;                   // consecutive queries (coherence order) a warp takes from the batch cursor at a time
#ifndef B2_PP_WARPS_PER_SM
#define B2_PP_WARPS_PER_SM 20
#endif
constexpr int kPpWarpsPerSM = B2_PP_WARPS_PER_SM;   // resident warps per SM the kernel is compiled for (register budget 65536 / 32 / this)
enum PpSlotState : uint32_t { kPpFree = 0, kPpReady = 1, kPpTraversing = 2, kPpDraining = 3 };

struct PairShared {
    uint2 stack[kPpStackSm][kPpWorkers];   // x = node index; y = lower bound of the box's squared distance from the binary32 query point
    uint32_t lpool[kPpLeafCap + 1];        // query slot << 27 | leaf slot : awaiting pre-test 1   [+1: sink of masked pushes]
    uint2 mpool[kPpMidCap];                // x = same, y = lower bound of the distance (binary32) : awaiting pre-test 2
    uint2 epool[kPpExactCap];              // same : awaiting the binary64 projection
    float4 qa[kPpSlots];                   // as closest_warp.cuh::WarpShared
    float4 qc[kPpSlots];
    float4 v0f[kPpSlots];                  // reference point (first vertex) of the slot's incumbent primitive, rounded to binary32
    float4 hint;                           // such a point of a query this warp completed recently; w != 0 when valid
    uint32_t bprim[kPpSlots];
    uint32_t qidx[kPpSlots];
    uint32_t flags[kPpSlots];
    uint32_t pushed[kPpSlots];             // leaf pairs the slot's traversal handed to the pipeline (valid once the slot is draining)
    uint32_t retired[kPpSlots];            // leaf pairs of the slot that have left the pipeline
    uint32_t state[kPpSlots];
    uint32_t w_slotbits[32], w_cur[32], w_sp[32], w_pushed[32], w_nodes[32];   // the workers' traversal state between node phases (per lane; pair-uniform but w_nodes)
    uint32_t wstate[12];                   // warp-wide scalars kept out of registers: [0..1] next position of the warp's chunk, [2] queries left in it,
                                           // [3] batch exhausted, [4] leaf pairs pre-tested, [5] binary64 projections, [6] queries that used the exact stage, [7] ties, [8] stack overflow
};

struct PairClosestArgs {
    const BvhNode* nodes;
    const TriRec* tris;
    const TriF* trisf;
    double mag;
    const uint32_t* perm;
    unsigned long long* counters;
    const double* pts;
    uint64_t n;
    double* cp;
    uint32_t* prim;
    double* sqd;
    uint8_t* flags;
    unsigned long long* cursor;
    uint2* spill;                  // gridDim.x * kPpSpillPerWarp entries
    WarpSlots64* slots64;          // gridDim.x records
    int exact_go;                  // run the exact phase when at least this many pairs wait (1 .. kPpExactCap - 32)
    uint32_t flag_tie, flag_exact;
};

// ---------------------------------------------------------------------------------------------------------------------------------
// warp abstraction: intrinsics on the device, fibers in tests/hostcheck (B2_WARP_EMU), inert in the product's host pass (never called)
// ---------------------------------------------------------------------------------------------------------------------------------
struct WarpCtx { unsigned lane; void* emu; unsigned phase; };

#if !defined(__CUDA_ARCH__) && defined(B2_WARP_EMU)
}  // namespace b2
namespace b2emu {   // provided by tests/hostcheck/warp_emu.h
void exchange(b2::WarpCtx& w, uint64_t v, uint64_t* all);
void sync(b2::WarpCtx& w);
}
namespace b2 {
#endif

B2_HD unsigned wballot(WarpCtx& w, bool p) {
#if defined(__CUDA_ARCH__)
    return __ballot_sync(0xFFFFFFFFu, p);
#elif defined(B2_WARP_EMU)
    uint64_t all[32]; b2emu::exchange(w, p ? 1 : 0, all);
    unsigned m = 0; for (int i = 0; i < 32; ++i) m |= all[i] ? 1u << i : 0u;
    return m;
#else
    return p ? 1u << w.lane : 0u;
#endif
}
B2_HD uint32_t wshfl_xor(WarpCtx& w, uint32_t v, int m) {
#if defined(__CUDA_ARCH__)
    return __shfl_xor_sync(0xFFFFFFFFu, v, m);
#elif defined(B2_WARP_EMU)
    uint64_t all[32]; b2emu::exchange(w, v, all); return (uint32_t)all[w.lane ^ (unsigned)m];
#else
    return v;
#endif
}
B2_HD float wshfl_xor_f(WarpCtx& w, float v, int m) {
#if defined(__CUDA_ARCH__)
    return __shfl_xor_sync(0xFFFFFFFFu, v, m);
#else
    uint32_t u; memcpy(&u, &v, 4); u = wshfl_xor(w, u, m); memcpy(&v, &u, 4); return v;
#endif
}
B2_HD unsigned long long wbroadcast64(WarpCtx& w, unsigned long long v, int src) {
#if defined(__CUDA_ARCH__)
    return __shfl_sync(0xFFFFFFFFu, v, src);
#elif defined(B2_WARP_EMU)
    uint64_t all[32]; b2emu::exchange(w, v, all); return all[src];
#else
    return v;
#endif
}
// mask of the lanes holding the same value (every lane takes part)
B2_HD unsigned wmatch(WarpCtx& w, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __match_any_sync(0xFFFFFFFFu, v);
#elif defined(B2_WARP_EMU)
    uint64_t all[32]; b2emu::exchange(w, v, all);
    unsigned m = 0; for (int i = 0; i < 32; ++i) m |= all[i] == v ? 1u << i : 0u;
    return m;
#else
    return 1u << w.lane;
#endif
}
// minimum of v over the lanes of `grp` (the groups partition the warp: every lane passes the mask of its own group)
B2_HD uint32_t wreduce_min(WarpCtx& w, unsigned grp, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __reduce_min_sync(grp, v);
#elif defined(B2_WARP_EMU)
    uint64_t all[32]; b2emu::exchange(w, v, all);
    uint32_t m = v; for (int i = 0; i < 32; ++i) if ((grp >> i) & 1u) m = (uint32_t)all[i] < m ? (uint32_t)all[i] : m;
    return m;
#else
    (void)grp; return v;
#endif
}
B2_HD uint32_t wreduce_add(WarpCtx& w, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __reduce_add_sync(0xFFFFFFFFu, v);
#elif defined(B2_WARP_EMU)
    uint64_t all[32]; b2emu::exchange(w, v, all);
    uint32_t s = 0; for (int i = 0; i < 32; ++i) s += (uint32_t)all[i];
    return s;
#else
    return v;
#endif
}
B2_HD void wsync(WarpCtx& w) {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#elif defined(B2_WARP_EMU)
    b2emu::sync(w);
#else
    (void)w;
#endif
}
B2_HD uint32_t pp_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
B2_HD float pp_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
B2_HD uint2 pp_ld_spill(const uint2* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
B2_HD void pp_st_spill(uint2* p, uint2 v) {
#if defined(__CUDA_ARCH__)
    __stcg(p, v);
#else
    *p = v;
#endif
}
B2_HD int pp_popc(unsigned m) {
#if defined(__CUDA_ARCH__)
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}
B2_HD int pp_ffs(unsigned m) {
#if defined(__CUDA_ARCH__)
    return __ffs(m);
#else
    return __builtin_ffs((int)m);
#endif
}
// memory: shared-memory atomics, L2-coherent accesses of the per-warp global state, global counters
B2_HD void pp_atomic_min(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicMin(p, v);
#else
    if (v < *p) *p = v;      // (fibers: one lane runs at a time)
#endif
}
B2_HD void pp_atomic_or(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
B2_HD unsigned long long pp_global_add(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
#endif
}
B2_HD void pp_global_or(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicOr(p, v);
#else
    __atomic_fetch_or(p, v, __ATOMIC_RELAXED);
#endif
}
B2_HD double pp_ldcg(const double* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *(const volatile double*)p;
#endif
}
B2_HD void pp_stcg(double* p, double v) {
#if defined(__CUDA_ARCH__)
    __stcg(p, v);
#else
    *(volatile double*)p = v;
#endif
}
// directed binary32 arithmetic (host versions: exact in binary64 where possible, then rounded to the safe side; the bounds they feed are
// pruning aids, results never depend on them)
B2_HD float pp_sub_ru(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_ru(a, b);
#else
    return float_above((double)a - (double)b);
#endif
}
B2_HD float pp_add_ru(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_ru(a, b);
#else
    return float_above((double)a + (double)b);
#endif
}
B2_HD float pp_mul_ru(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_ru(a, b);
#else
    return float_above((double)a * (double)b);
#endif
}
B2_HD float pp_fma_ru(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_ru(a, b, c);
#else
    return nextafterf(float_above((double)a * (double)b + (double)c), INFINITY);   // (the binary64 sum may have rounded down: one more ulp)
#endif
}
B2_HD float pp_fma_rd(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rd(a, b, c);
#else
    return nextafterf(float_below((double)a * (double)b + (double)c), -INFINITY);
#endif
}
B2_HD float pp_sqrt_ru(float a) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_ru(a);
#else
    return nextafterf(float_above(sqrt((double)a)), INFINITY);
#endif
}

// one 32-byte half of a node = one child: box and reference
struct HalfNode { float lo[3], hi[3]; uint32_t ref; };
B2_HD HalfNode pp_load_half(const BvhNode* __restrict__ nodes, uint32_t i, unsigned side) {
    HalfNode h;
#if defined(__CUDA_ARCH__)
    uint32_t f[8];
    ldg256(reinterpret_cast<const char*>(nodes + i) + 32u * side, f);
    h.lo[0] = __uint_as_float(f[0]); h.lo[1] = __uint_as_float(f[1]); h.lo[2] = __uint_as_float(f[2]);
    h.hi[0] = __uint_as_float(f[3]); h.hi[1] = __uint_as_float(f[4]); h.hi[2] = __uint_as_float(f[5]);
    h.ref = f[6];
#else
    const BvhNode& s = nodes[i];
    for (int k = 0; k < 3; ++k) { h.lo[k] = side ? s.lo1[k] : s.lo0[k]; h.hi[k] = side ? s.hi1[k] : s.hi0[k]; }
    h.ref = side ? s.child1 : s.child0;
#endif
    return h;
}
// LOWER bound of the squared distance from the binary32 point (qx,qy,qz) to the closed box (every operation rounds towards the smaller result)
B2_HD float pp_box_sqdist(const float* lo, const float* hi, float qx, float qy, float qz) {
    const float dx = fmaxf(fmaxf(sub_dn(lo[0], qx), sub_dn(qx, hi[0])), 0.0f);
    const float dy = fmaxf(fmaxf(sub_dn(lo[1], qy), sub_dn(qy, hi[1])), 0.0f);
    const float dz = fmaxf(fmaxf(sub_dn(lo[2], qz), sub_dn(qz, hi[2])), 0.0f);
    return pp_fma_rd(dz, dz, pp_fma_rd(dy, dy, mul_dn(dx, dx)));
}
B2_HD RelTri pp_relative(const TriF* __restrict__ trisf, uint32_t slot, const float4& q) {
    RelTri r;
#if defined(__CUDA_ARCH__)
    uint32_t f[8];
    ldg256(trisf + slot, f);
    const float cz = __ldg(&trisf[slot].v[8]);
    const float v[9] = {__uint_as_float(f[0]), __uint_as_float(f[1]), __uint_as_float(f[2]), __uint_as_float(f[3]), __uint_as_float(f[4]),
                        __uint_as_float(f[5]), __uint_as_float(f[6]), __uint_as_float(f[7]), cz};
#else
    const float* v = trisf[slot].v;
#endif
    r.ra[0] = v[0] - q.x; r.ra[1] = v[1] - q.y; r.ra[2] = v[2] - q.z;
    r.rb[0] = v[3] - q.x; r.rb[1] = v[4] - q.y; r.rb[2] = v[5] - q.z;
    r.rc[0] = v[6] - q.x; r.rc[1] = v[7] - q.y; r.rc[2] = v[8] - q.z;
    r.m = fmaxf(fmaxf(fmaxf(fabsf(r.ra[0]), fabsf(r.ra[1])), fmaxf(fabsf(r.ra[2]), fabsf(r.rb[0]))),
                fmaxf(fmaxf(fabsf(r.rb[1]), fabsf(r.rb[2])), fmaxf(fmaxf(fabsf(r.rc[0]), fabsf(r.rc[1])), fabsf(r.rc[2]))));
    return r;
}
// shrinks slot `owner`'s ball to radius rb (any lane, any time: both bounds only ever decrease)
B2_HD void pp_shrink(PairShared& S, unsigned owner, float rb) {
    pp_atomic_min(reinterpret_cast<uint32_t*>(&S.qc[owner].x), pp_f2u(rb));
    const float r = pp_add_ru(rb, S.qc[owner].w);
    pp_atomic_min(reinterpret_cast<uint32_t*>(&S.qa[owner].w), pp_f2u(pp_mul_ru(r, r)));
}
// pruning radius from an incumbent's squared distance: as closest_warp.cuh::wp_set_bounds (shrink only)
B2_HD void pp_set_bounds(PairShared& S, unsigned owner, double sqd, double delta) {
    const float s = pp_sqrt_ru(float_above(sqd));
    pp_shrink(S, owner, pp_sqrt_ru(float_above(sqd + ((2.0 * delta * (double)s + delta * delta) + sqd * 1e-13))));
}
struct PpCand { double x, y, z; int nexact; };
B2_HD PpCand pp_candidate(const Tri3& t, const Vec3& o) {
#if defined(__CUDA_ARCH__)
    const CandOut c = wp_candidate(t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, o.x, o.y, o.z);   // out of line
    return PpCand{c.x, c.y, c.z, c.nexact};
#else
    int nexact = 0;
    const Vec3 p = closest_candidate(t, o, nexact);
    return PpCand{p.x, p.y, p.z, nexact};
#endif
}
// `r` lanes each retire one pair of slot `owner`: one shared-memory update per slot (the lanes of a slot elect a leader)
B2_HD void pp_retire(PairShared& S, WarpCtx& w, bool r, unsigned owner) {
    const unsigned grp = wmatch(w, r ? owner : 64u + w.lane);
    if (r && (grp & ((1u << w.lane) - 1u)) == 0) S.retired[owner] += (uint32_t)pp_popc(grp);
    wsync(w);
}

// (B3) of closest_pair_body: the binary64 projection + exact comparison of the pairs waiting in the exact pool.  OUT OF LINE, so that its
// register needs (binary64 geometry, the exact predicates) do not weigh on the allocation of the node loop.  Returns (pairs left in the
// pool, binary64 projections evaluated).
B2_HD_NOINLINE uint2 pp_exact_phase(PairShared& S, WarpSlots64& G, const TriRec* __restrict__ tris, double mag, int exact_go, uint32_t flag_tie, uint32_t flag_exact,
                                    int ecnt, int drain, WarpCtx* wp) {
#if defined(__CUDA_ARCH__)
    WarpCtx w; w.lane = threadIdx.x & 31u; w.emu = nullptr; w.phase = 0; (void)wp;
#else
    WarpCtx& w = *wp;
#endif
    const unsigned lane = w.lane, lt = (1u << lane) - 1u;
    uint32_t evals = 0;
    {
            // -------------------------------------------------------------------------------------------- (B3) binary64 projection + exact comparison
            // (i) pairs whose lower bound has fallen outside their slot's ball retire without any binary64 work
            int kept = 0;
            for (int base = 0; base < ecnt; base += 32) {
                const int i = base + (int)lane;
                bool keep = false;
                uint2 e = make_uint2(0, 0);
                unsigned owner = 0;
                if (i < ecnt) {
                    e = S.epool[i];
                    owner = e.x >> 27;
                    keep = !(pp_u2f(e.y) > S.qc[owner].x);
                }
                wsync(w);
                const unsigned km = wballot(w, keep);
                if (keep) S.epool[kept + pp_popc(km & lt)] = e;
                kept += pp_popc(km);
                pp_retire(S, w, i < ecnt && !keep, owner);
            }
            ecnt = kept;
            if (ecnt >= exact_go || (ecnt > 0 && drain)) {
                for (int base = 0; base < ecnt; base += 32) {
                    const int i = base + (int)lane;
                    const bool has = i < ecnt;
                    bool pending = false;
                    unsigned owner = 0;
                    uint32_t id = 0, key = 0;
                    Vec3 o = {0.0, 0.0, 0.0}, cpt = {0.0, 0.0, 0.0}, v0 = {0.0, 0.0, 0.0};
                    double sqd = 0.0;
                    if (has) {
                        const uint2 e = S.epool[i];
                        owner = e.x >> 27;
                        if (!(pp_u2f(e.y) > S.qc[owner].x)) {   // (else: an earlier group of this very phase has shrunk the ball)
                            const Tri3 t = load_tri(tris, e.x & kWpRefMask, id);
                            o = {pp_ldcg(&G.o[0][owner]), pp_ldcg(&G.o[1][owner]), pp_ldcg(&G.o[2][owner])};
                            const PpCand co = pp_candidate(t, o);
                            cpt = {co.x, co.y, co.z};
                            v0 = t.a;
                            if (co.nexact) pp_atomic_or(&S.flags[owner], flag_exact);
                            const Vec3 v = cpt - o;
                            sqd = (v.x * v.x + v.y * v.y) + v.z * v.z;   // Compute_squared_distance_3(query, closest): AABB_tree.h:1024-1032
                            key = pp_f2u((float)sqd);
                            pending = true;
                        }
                    }
                    // Every waiting candidate is compared (exactly) with its slot's incumbent, all lanes at once; the ones that are farther retire.  Of
                    // the rest, ONE per slot (the apparently nearest) replaces the incumbent, the others meet the new incumbent in the next round.
                    // Each replacement is a step the sequential reference loop could have taken, and a candidate farther than an earlier incumbent
                    // is farther than every later one, so the outcome is that of the sequential loop.
                    unsigned pm = wballot(w, pending);
                    evals += (uint32_t)pp_popc(pm);
                    while (pm) {
                        int sg = -1;                                   // (no incumbent yet: take)
                        if (pending && S.bprim[owner] != kNone) {
                            const Vec3 bp = {pp_ldcg(&G.bp[0][owner]), pp_ldcg(&G.bp[1][owner]), pp_ldcg(&G.bp[2][owner])};
                            int nexact = 0;
                            sg = cmp_dist(o, cpt, bp, nexact);
                            if (nexact) pp_atomic_or(&S.flags[owner], flag_exact);
                            if (sg > 0) pending = false;
                        }
                        wsync(w);                                      // incumbents read before any is replaced
                        const unsigned grp = wmatch(w, pending ? owner : 64u + lane);
                        const uint32_t gmin = wreduce_min(w, grp, key);
                        const unsigned wm = wballot(w, pending && key == gmin) & grp;
                        if (pending && (unsigned)(pp_ffs(wm) - 1) == lane) {
                            // the replacement step of the reference (AABB_traits_3.h:724-732); same tie rule as bvh.cuh::traverse_closest
                            uint32_t fl = S.flags[owner];
                            bool take;
                            if (sg == 0) { take = id < S.bprim[owner]; fl |= flag_tie; }
                            else { take = true; fl &= ~flag_tie; }
                            S.flags[owner] = fl;
                            if (take) {
                                pp_stcg(&G.bp[0][owner], cpt.x); pp_stcg(&G.bp[1][owner], cpt.y); pp_stcg(&G.bp[2][owner], cpt.z);
                                pp_stcg(&G.bsqd[owner], sqd);
                                S.bprim[owner] = id;
                                S.v0f[owner] = make_float4((float)v0.x, (float)v0.y, (float)v0.z, 0.0f);
                                pp_set_bounds(S, owner, sqd, 4e-16 * fmax(mag, fmax3(fabs(o.x), fabs(o.y), fabs(o.z))));
                            }
                            pending = false;
                        }
                        wsync(w);
                        pm = wballot(w, pending);
                    }
                    pp_retire(S, w, has, owner);
                }
                ecnt = 0;
            }
            }
    return make_uint2((uint32_t)ecnt, evals);
}

// (B1) + (B2) of closest_pair_body: the two binary32 pre-test stages of the leaf pipeline, OUT OF LINE like the exact phase.
// counts = lcnt | mcnt << 8 | ecnt << 16 (in and out); flush = run partial groups too.  Returns (counts, leaf pairs that entered pre-test 1).
B2_HD_NOINLINE uint2 pp_pretest_phases(PairShared& S, const TriF* __restrict__ trisf, uint32_t counts, int flush, WarpCtx* wp) {
#if defined(__CUDA_ARCH__)
    WarpCtx w; w.lane = threadIdx.x & 31u; w.emu = nullptr; w.phase = 0; (void)wp;
#else
    WarpCtx& w = *wp;
#endif
    const unsigned lane = w.lane, lt = (1u << lane) - 1u;
    int lcnt = (int)(counts & 255u), mcnt = (int)((counts >> 8) & 255u), ecnt = (int)(counts >> 16);
    uint32_t leaves = 0;
    if (mcnt < 32) {
        // ------------------------------------------------------------------------------------------------ (B1) pre-test 1: lower bound along the centroid direction
        while (lcnt >= 32 || (lcnt > 0 && flush)) {
            const int take = lcnt < 32 ? lcnt : 32;
            lcnt -= take;
            const bool valid = (int)lane < take;
            bool survive = false;
            uint32_t pair = 0;
            unsigned owner = 0;
            float lb = 0.0f;
            if (valid) {
                pair = S.lpool[lcnt + (int)lane];
                owner = pair >> 27;
                const RelTri r = pp_relative(trisf, pair & kWpRefMask, S.qa[owner]);
                const float4 qc = S.qc[owner];
                const float g[3] = {(r.ra[0] + r.rb[0]) + r.rc[0], (r.ra[1] + r.rb[1]) + r.rc[1], (r.ra[2] + r.rb[2]) + r.rc[2]};
                lb = wp_lower(g, sqrtf(dot3f(g, g)), r) - qc.y;
                survive = !(lb > qc.x);
            }
            leaves += (uint32_t)take;
            const unsigned sm = wballot(w, survive);
            if (survive) S.mpool[mcnt + pp_popc(sm & lt)] = make_uint2(pair, pp_f2u(lb));
            mcnt += pp_popc(sm);
            pp_retire(S, w, valid && !survive, owner);
        }
    }
    wsync(w);
    // ---------------------------------------------------------------------------------------------------- (B2) pre-test 2: approximate closest point
    while ((mcnt >= 32 || (mcnt > 0 && flush && lcnt == 0)) && ecnt + 32 <= kPpExactCap) {
        const int take = mcnt < 32 ? mcnt : 32;
        mcnt -= take;
        const bool valid = (int)lane < take;
        bool survive = false;
        uint2 e = make_uint2(0, 0);
        unsigned owner = 0;
        if (valid) {
            e = S.mpool[mcnt + (int)lane];
            owner = e.x >> 27;
            const float4 qc = S.qc[owner];
            if (!(pp_u2f(e.y) > qc.x)) {                    // (else: the ball has shrunk since pre-test 1)
                const RelTri r = pp_relative(trisf, e.x & kWpRefMask, S.qa[owner]);
                float lb, hi;
                wp_sharp(r, qc.z, qc.y, &lb, &hi);
                if (hi < qc.x) pp_shrink(S, owner, hi);        // this primitive is certainly nearer than the slot's ball: shrink it now, in binary32
                lb = fmaxf(lb, pp_u2f(e.y));
                if (!(lb > qc.x)) { survive = true; e.y = pp_f2u(lb); }
            }
        }
        const unsigned sm = wballot(w, survive);
        if (survive) S.epool[ecnt + pp_popc(sm & lt)] = e;
        ecnt += pp_popc(sm);
        pp_retire(S, w, valid && !survive, owner);
    }
    wsync(w);
    return make_uint2((uint32_t)lcnt | ((uint32_t)mcnt << 8) | ((uint32_t)ecnt << 16), leaves);
}

// position of the (k+1)-th set bit of m (k < popcount(m))
B2_HD unsigned pp_nth_bit(unsigned m, unsigned k) {
#if defined(__CUDA_ARCH__)
    return __fns(m, 0, (int)k + 1);
#else
    for (unsigned i = 0; i < 32; ++i) if ((m >> i) & 1u) { if (k == 0) return i; --k; }
    return 32;
#endif
}

#if defined(__CUDA_ARCH__)
#define B2_WCTX_ARG(w) nullptr
#else
#define B2_WCTX_ARG(w) (&(w))
#endif

constexpr uint32_t kPpNoSlot = 0xFFFFFFFFu;

// The node phase of closest_pair_body: every worker (lane pair) walks its own query depth-first for up to 16 steps, or until the leaf pool
// is nearly full.  OUT OF LINE and free of calls, so that its loop-carried state lives in registers; the workers' state is parked in shared
// memory (PairShared::w_*) between phases.  Returns lcnt | (1 << 16 when some worker still has nodes to visit).
B2_HD_NOINLINE uint32_t pp_node_phase(PairShared& S, const BvhNode* __restrict__ nodes, uint2* myspill, int lcnt, WarpCtx* wp) {
#if defined(__CUDA_ARCH__)
    WarpCtx w; w.lane = threadIdx.x & 31u; w.emu = nullptr; w.phase = 0; (void)wp;
#else
    WarpCtx& w = *wp;
#endif
    const unsigned lane = w.lane, lt = (1u << lane) - 1u, wk = lane >> 1, side = lane & 1u;
    // worker state, identical in both lanes of a pair: slotbits = query slot << 27 (kPpNoSlot: the worker is idle)
    uint32_t slotbits = S.w_slotbits[lane], cur = S.w_cur[lane], pushed = S.w_pushed[lane], n_nodes = S.w_nodes[lane];
    int sp = (int)S.w_sp[lane];
    float qx = 0.0f, qy = 0.0f, qz = 0.0f, bound2 = 0.0f;
    if (slotbits != kPpNoSlot) { const float4 q = S.qa[slotbits >> 27]; qx = q.x; qy = q.y; qz = q.z; bound2 = q.w; }   // (the leaf stages shrink the ball between node phases)
    bool nodework;
    for (int it = 0;; ++it) {
        const bool has = cur != kNone || sp > 0;
        if (!has && slotbits != kPpNoSlot) {                      // this traversal has ended: the slot drains, the worker is free
            if (side == 0) { S.pushed[slotbits >> 27] = pushed; S.state[slotbits >> 27] = kPpDraining; }
            slotbits = kPpNoSlot;
        }
        const unsigned hm = wballot(w, has);
        nodework = hm != 0;
        if (!nodework || lcnt > kPpLeafCap - 32 || it >= 16) break;
        // pop (one attempt per step: a stale entry — its box has fallen outside the ball — costs this worker one step, no load)
        if (has && cur == kNone) {
            --sp;
            const uint2 e = sp < kPpStackSm ? S.stack[sp][wk] : pp_ld_spill(&myspill[wk * (kPpStackCap - kPpStackSm) + (sp - kPpStackSm)]);
            if (pp_u2f(e.y) <= bound2) cur = e.x;
        }
        wsync(w);                                                   // both lanes of a pair have read the popped entry before either stacks a new one there
        const bool act = cur != kNone;
        float d = INFINITY;
        uint32_t cref = 0;
        if (act) {
            const HalfNode h = pp_load_half(nodes, cur, side);
            d = pp_box_sqdist(h.lo, h.hi, qx, qy, qz);
            cref = h.ref;
            ++n_nodes;
        }
        const float od = wshfl_xor_f(w, d, 1);
        const uint32_t oref = wshfl_xor(w, cref, 1);
        const bool in = act && d <= bound2, oin = act && od <= bound2;
        const bool mleaf = in && (cref & kLeafBit) != 0;
        const bool mint = in && (cref & kLeafBit) == 0, oint = oin && (oref & kLeafBit) == 0;
        // leaf children go to the shared pipeline
        const unsigned lm = wballot(w, mleaf);
        S.lpool[mleaf ? lcnt + pp_popc(lm & lt) : kPpLeafCap] = slotbits | (cref & kWpRefMask);
        lcnt += pp_popc(lm);
        pushed += (mleaf ? 1u : 0u) + ((oin && (oref & kLeafBit) != 0) ? 1u : 0u);
        // internal children: the nearer one is next, the farther one is stacked by the lane that tested it
        if (act) {
            if (mint && oint) {
                const bool near = d < od || (d == od && side == 0);
                if (!near) {
                    const uint2 e = make_uint2(cref, pp_f2u(d));
                    if (sp < kPpStackSm) S.stack[sp][wk] = e;
                    else if (sp < kPpStackCap) pp_st_spill(&myspill[wk * (kPpStackCap - kPpStackSm) + (sp - kPpStackSm)], e);
                    else S.wstate[8] = 1u;                          // (deeper than any tree this library builds: reported, never silent)
                }
                if (sp < kPpStackCap) ++sp;
                cur = near ? cref : oref;
            } else cur = mint ? cref : (oint ? oref : kNone);
        }
        wsync(w);                                                 // a stacked entry is popped by both lanes of the pair
    }
    S.w_slotbits[lane] = slotbits; S.w_cur[lane] = cur; S.w_pushed[lane] = pushed; S.w_nodes[lane] = n_nodes; S.w_sp[lane] = (uint32_t)sp;
    wsync(w);                                                     // the node phase's pool writes are visible to every lane
    return (uint32_t)lcnt | (nodework ? 1u << 16 : 0u);
}

B2_HD void closest_pair_body(const PairClosestArgs& a, PairShared& S, WarpSlots64& G, uint2* myspill, WarpCtx& w) {
    const unsigned lane = w.lane;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned side = lane & 1u;
    int lcnt = 0, mcnt = 0, ecnt = 0;                                // warp-uniform

    S.state[lane] = kPpFree;
    S.w_slotbits[lane] = kPpNoSlot; S.w_cur[lane] = kNone; S.w_sp[lane] = 0u; S.w_pushed[lane] = 0u; S.w_nodes[lane] = 0u;
    if (lane < 12) S.wstate[lane] = 0u;
    if (lane == 0) S.hint = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    wsync(w);

    for (;;) {
        // ------------------------------------------------------------------------------------------------ housekeeping (lane = slot)
        int flush;
        {
            const uint32_t st = S.state[lane];
            const bool done = st == kPpDraining && S.retired[lane] == S.pushed[lane];
            if (done) {
                const uint64_t qi = S.qidx[lane];
                if (a.cp) { a.cp[3 * qi] = pp_ldcg(&G.bp[0][lane]); a.cp[3 * qi + 1] = pp_ldcg(&G.bp[1][lane]); a.cp[3 * qi + 2] = pp_ldcg(&G.bp[2][lane]); }
                if (a.prim) a.prim[qi] = S.bprim[lane];
                if (a.sqd) a.sqd[qi] = pp_ldcg(&G.bsqd[lane]);
                if (a.flags) a.flags[qi] = (uint8_t)S.flags[lane];
                S.state[lane] = kPpFree;
            }
            const unsigned dm = wballot(w, done);
            if (dm) {                                                 // (warp-uniform)
                const uint32_t fl = done ? S.flags[lane] : 0u;
                const unsigned tm = wballot(w, (fl & a.flag_tie) != 0), xm = wballot(w, (fl & a.flag_exact) != 0);
                if (lane == 0) { S.wstate[6] += (uint32_t)pp_popc(xm); S.wstate[7] += (uint32_t)pp_popc(tm); }
                if ((unsigned)(pp_ffs(dm) - 1) == lane) { const float4 v = S.v0f[lane]; S.hint = make_float4(v.x, v.y, v.z, 1.0f); }
                wsync(w);
            }
            const bool isfree = st == kPpFree || done;
            const unsigned fm = wballot(w, isfree);
            unsigned ready = wballot(w, st == kPpReady);
            const bool idle = S.w_slotbits[lane] == kPpNoSlot;
            const unsigned im = wballot(w, idle) & 0x55555555u;      // idle workers (bit of their even lane)
            uint32_t wleft = S.wstate[2];
            bool exhausted = S.wstate[3] != 0;
            unsigned prepared = 0;
            if (fm != 0 && (!exhausted || wleft != 0) && (pp_popc(im) > pp_popc(ready) || pp_popc(fm) >= 8)) {   // (warp-uniform)
                unsigned long long wbase = (unsigned long long)S.wstate[0] | ((unsigned long long)S.wstate[1] << 32);
                wsync(w);
                if (wleft == 0) {                                     // the warp's chunk of the batch is used up: take the next one
                    unsigned long long base = 0;
                    if (lane == 0) base = pp_global_add(a.cursor, (unsigned long long)kPpChunk);
                    base = wbroadcast64(w, base, 0);
                    if (base >= a.n) exhausted = true;
                    else { wbase = base; wleft = (uint32_t)(a.n - base < (uint64_t)kPpChunk ? a.n - base : (uint64_t)kPpChunk); }
                }
                const uint32_t rank = (uint32_t)pp_popc(fm & lt);
                const bool take = isfree && rank < wleft;
                if (take) {
                    const float ev = float_above(a.mag * 5.9604645e-8 * 1.000001);   // >= rounding of a vertex coordinate to binary32
                    const unsigned long long pos = wbase + rank;
                    const uint32_t qi = a.perm ? a.perm[pos] : (uint32_t)pos;
                    const double* p = a.pts + 3 * (uint64_t)qi;
                    const Vec3 o = {p[0], p[1], p[2]};
                    float4 q;
                    q.x = (float)o.x; q.y = (float)o.y; q.z = (float)o.z;
                    const double ex = o.x - (double)q.x, ey = o.y - (double)q.y, ez = o.z - (double)q.z;   // exact differences
                    const float perr = float_above(sqrt((ex * ex + ey * ey) + ez * ez) * 1.0000001) + 1e-37f;
                    const double delta = 4e-16 * fmax(a.mag, fmax3(fabs(o.x), fabs(o.y), fabs(o.z)));
                    const float dslack = float_above(2.0 * delta) + 1e-37f;
                    const float e2 = pp_add_ru(pp_mul_ru(1.7320509f, ev), perr);
                    // First ball: the reference point (first vertex) h of some primitive lies within |q - h| + e2 of the query, and that
                    // primitive's candidate is never farther than its reference point (geom.cuh::closest_candidate), so the winner is not either.
                    float rb = INFINITY;
                    const float4 h = S.hint;
                    if (h.w != 0.0f) {
                        const float dx = fmaxf(pp_sub_ru(q.x, h.x), pp_sub_ru(h.x, q.x)), dy = fmaxf(pp_sub_ru(q.y, h.y), pp_sub_ru(h.y, q.y)),
                                    dz = fmaxf(pp_sub_ru(q.z, h.z), pp_sub_ru(h.z, q.z));
                        const float su = pp_sqrt_ru(pp_fma_ru(dz, dz, pp_fma_ru(dy, dy, pp_mul_ru(dx, dx))));
                        rb = pp_add_ru(pp_add_ru(pp_mul_ru(su, 1.000001f), e2), dslack);
                    }
                    const float r = pp_add_ru(rb, perr);
                    q.w = pp_mul_ru(r, r);
                    pp_stcg(&G.o[0][lane], o.x); pp_stcg(&G.o[1][lane], o.y); pp_stcg(&G.o[2][lane], o.z);
                    S.qa[lane] = q;
                    S.qc[lane] = make_float4(rb, e2, dslack, perr);
                    S.bprim[lane] = kNone;
                    S.qidx[lane] = qi;
                    S.flags[lane] = 0u;
                    S.pushed[lane] = 0u;
                    S.retired[lane] = 0u;
                    S.state[lane] = kPpReady;
                }
                prepared = wballot(w, take);
                const uint32_t need = (uint32_t)pp_popc(fm);
                const uint32_t used = need < wleft ? need : wleft;
                wbase += used; wleft -= used;
                if (lane == 0) { S.wstate[0] = (uint32_t)wbase; S.wstate[1] = (uint32_t)(wbase >> 32); S.wstate[2] = wleft; S.wstate[3] = exhausted ? 1u : 0u; }
            }
            ready |= prepared;
            wsync(w);
            // termination: nothing in flight and nothing left to hand out (the pools only hold pairs of slots in flight)
            if (((~fm) | prepared) == 0 && exhausted && wleft == 0) break;
            // the k-th idle worker starts on the k-th ready slot (both lanes of a pair decide from the same masks)
            const unsigned nready = (unsigned)pp_popc(ready);
            if (idle) {
                const unsigned k = (unsigned)pp_popc(im & ((1u << (lane & ~1u)) - 1u));
                if (k < nready) {
                    const unsigned slot = pp_nth_bit(ready, k);
                    S.w_slotbits[lane] = slot << 27; S.w_cur[lane] = 0u; S.w_sp[lane] = 0u; S.w_pushed[lane] = 0u;
                    if (side == 0) S.state[slot] = kPpTraversing;
                }
            }
            const int nidle = pp_popc(im), blocked = nidle > (int)nready ? nidle - (int)nready : 0;
            // a quarter of the workers wait for slots to drain: push the waiting pairs through the pipeline now, full groups or not
            flush = blocked >= 4 ? 1 : 0;
            wsync(w);
        }
        // ------------------------------------------------------------------------------------------------ node phase, then the leaf pipeline
        {
            const uint32_t r = pp_node_phase(S, a.nodes, myspill, lcnt, B2_WCTX_ARG(w));
            lcnt = (int)(r & 0xFFFFu);
            if (!(r >> 16)) flush = 1;                            // nothing upstream can add to the pools: drain them
        }
        if (lcnt >= 32 || mcnt >= 32 || (flush && (lcnt | mcnt) != 0)) {
            const uint2 r = pp_pretest_phases(S, a.trisf, (uint32_t)lcnt | ((uint32_t)mcnt << 8) | ((uint32_t)ecnt << 16), flush, B2_WCTX_ARG(w));
            lcnt = (int)(r.x & 255u); mcnt = (int)((r.x >> 8) & 255u); ecnt = (int)(r.x >> 16);
            if (lane == 0) S.wstate[4] += r.y;
        }
        {
            const int drain = (flush && lcnt == 0 && mcnt == 0) ? 1 : 0;
            if (ecnt >= a.exact_go || (ecnt > 0 && drain)) {
                const uint2 r = pp_exact_phase(S, G, a.tris, a.mag, a.exact_go, a.flag_tie, a.flag_exact, ecnt, drain, B2_WCTX_ARG(w));
                ecnt = (int)r.x;
                if (lane == 0) S.wstate[5] += r.y;
            }
        }
        wsync(w);
    }
    {
        const uint32_t nodes = wreduce_add(w, S.w_nodes[lane]) / 2u;   // (both lanes of a pair count the visit)
        if (lane == 0) {
            pp_global_add(&a.counters[3], (unsigned long long)nodes);
            pp_global_add(&a.counters[4], (unsigned long long)S.wstate[4]);
            pp_global_add(&a.counters[6], (unsigned long long)(S.wstate[4] - S.wstate[5]));   // leaves that never reached the binary64 projection
            if (S.wstate[6]) pp_global_add(&a.counters[0], (unsigned long long)S.wstate[6]);
            if (S.wstate[7]) pp_global_add(&a.counters[1], (unsigned long long)S.wstate[7]);
            if (S.wstate[8]) pp_global_or(&a.counters[2], 1ull);
        }
    }
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(32, kPpWarpsPerSM) closest_point_pair_kernel(const PairClosestArgs a) {
    __shared__ PairShared S;
    WarpCtx w;
    w.lane = threadIdx.x; w.emu = nullptr; w.phase = 0;
    closest_pair_body(a, S, a.slots64[blockIdx.x], a.spill + (size_t)blockIdx.x * kPpSpillPerWarp, w);
}
#endif

}  // namespace b2

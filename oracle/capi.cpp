// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under cgal_b200/ may include, link or call this.
// C entry points (ctypes) over the CPU restatement in kernel.hpp / tree.hpp.  Used by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, never by the product.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include "tree.hpp"

namespace orc { thread_local Stats g_stats; }
using namespace orc;

namespace {
std::atomic<uint64_t> g_box_tests{0}, g_prim_tests{0};   // traversal work of the batched queries, summed over the worker threads
struct Handle { Mesh mesh; std::unique_ptr<Tree> tree; double build_seconds = 0, kd_seconds = 0; };

template <class F> void parallel_for(uint64_t n, int nthreads, F f) {
    if (nthreads <= 1 || n < 1024) { const Stats s0 = g_stats; f(0, n); g_box_tests += g_stats.box_tests - s0.box_tests; g_prim_tests += g_stats.prim_tests - s0.prim_tests; return; }
    std::vector<std::thread> th;
    const uint64_t chunk = (n + nthreads - 1) / nthreads;      // static blocked ranges over queries, standing in for
    for (int t = 0; t < nthreads; ++t) {                        // tbb::parallel_for (Polygon_mesh_processing/distance.h:163-169)
        uint64_t b = t * chunk, e = std::min<uint64_t>(n, b + chunk);
        if (b >= e) break;
        th.emplace_back([=] { const Stats s0 = g_stats; f(b, e); g_box_tests += g_stats.box_tests - s0.box_tests; g_prim_tests += g_stats.prim_tests - s0.prim_tests; });
    }
    for (auto& x : th) x.join();
}
inline P3 P(const double* d) { return {d[0], d[1], d[2]}; }
inline Tri T(const double* d) { return {P(d), P(d + 3), P(d + 6)}; }
inline Box B(const double* d) { return {{d[0], d[1], d[2]}, {d[3], d[4], d[5]}}; }
inline void put(double* d, const P3& p) { d[0] = p.x; d[1] = p.y; d[2] = p.z; }
}  // namespace

extern "C" {

// verts: nv x 3 doubles; faces: nf x 3 uint32 (OFF order).  Builds the median-split tree.
void* orc_create(const double* verts, uint64_t nv, const uint32_t* faces, uint64_t nf) {
    (void)nv;
    Handle* h = new Handle();
    h->mesh.tris.resize(nf);
    for (uint64_t i = 0; i < nf; ++i)
        h->mesh.tris[i] = {P(verts + 3 * (uint64_t)faces[3 * i]), P(verts + 3 * (uint64_t)faces[3 * i + 1]), P(verts + 3 * (uint64_t)faces[3 * i + 2])};
    auto t0 = std::chrono::steady_clock::now();
    h->tree.reset(new Tree(h->mesh));
    h->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return h;
}
// triangle soup: tris = n x 9 doubles
void* orc_create_soup(const double* tris, uint64_t n) {
    Handle* h = new Handle();
    h->mesh.tris.resize(n);
    for (uint64_t i = 0; i < n; ++i) h->mesh.tris[i] = T(tris + 9 * i);
    auto t0 = std::chrono::steady_clock::now();
    h->tree.reset(new Tree(h->mesh));
    h->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return h;
}
// Segment_3 primitives: segs = n x 6 doubles (source, target)
void* orc_create_segments(const double* segs, uint64_t n) {
    Handle* h = new Handle();
    h->mesh.segments = true;
    h->mesh.tris.resize(n);
    for (uint64_t i = 0; i < n; ++i) h->mesh.tris[i] = {P(segs + 6 * i), P(segs + 6 * i + 3), P(segs + 6 * i + 3)};
    auto t0 = std::chrono::steady_clock::now();
    h->tree.reset(new Tree(h->mesh));
    h->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return h;
}
void orc_destroy(void* hv) { delete (Handle*)hv; }
double orc_build_seconds(void* hv) { return ((Handle*)hv)->build_seconds; }
double orc_kd_seconds(void* hv) { return ((Handle*)hv)->kd_seconds; }
uint64_t orc_size(void* hv) { return ((Handle*)hv)->tree->size(); }
void orc_accelerate_distance_queries(void* hv) {
    Handle* h = (Handle*)hv;
    auto t0 = std::chrono::steady_clock::now();
    h->tree->accelerate_distance_queries();
    h->kd_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
void orc_do_not_accelerate_distance_queries(void* hv) { ((Handle*)hv)->tree->do_not_accelerate_distance_queries(); }
void orc_bbox(void* hv, double* out) { Box b = ((Handle*)hv)->tree->bbox(); for (int i = 0; i < 3; ++i) { out[i] = b.lo[i]; out[3 + i] = b.hi[i]; } }

// kind: 0 = Segment_3 (source,target), 1 = Ray_3 (source, second point), 2 = Line_3 (point, direction vector): q = n x 6 doubles;
//       3 = Plane_3 (a, b, c, d): q = n x 4 doubles; 4 = Triangle_3 (three vertices): q = n x 9 doubles
static inline Tree::LineQ LQ(const double* q, uint64_t i, int kind) {
    if (kind == 3) return {P(q + 4 * i), P3{q[4 * i + 3], 0, 0}, 3};
    if (kind == 4) return {P(q + 9 * i), P(q + 9 * i + 3), 4, P(q + 9 * i + 6)};
    return {P(q + 6 * i), P(q + 6 * i + 3), kind};
}
void orc_do_intersect(void* hv, const double* q, uint64_t n, int kind, uint8_t* out, int nthreads) {
    const Tree& t = *((Handle*)hv)->tree;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) out[i] = t.do_intersect(LQ(q, i, kind)) ? 1 : 0;
    });
}
void orc_count(void* hv, const double* q, uint64_t n, int kind, uint32_t* out, int nthreads) {
    const Tree& t = *((Handle*)hv)->tree;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) out[i] = t.number_of_intersected_primitives(LQ(q, i, kind));
    });
}
// any_intersected_primitive (AABB_tree.h:947-956): 0xFFFFFFFF when none
void orc_any_intersected_primitive(void* hv, const double* q, uint64_t n, int kind, uint32_t* out, int nthreads) {
    const Tree& t = *((Handle*)hv)->tree;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) out[i] = t.any_intersected_primitive(LQ(q, i, kind));
    });
}
// all_intersected_primitives (AABB_tree.h:934-945): ids of query i, in the reference's traversal order, at ids[offsets[i] .. offsets[i+1]);
// offsets = exclusive prefix sums of orc_count.  Returns -1 if a count disagrees with the offsets.
int orc_all_intersected_primitives(void* hv, const double* q, uint64_t n, int kind, const uint64_t* offsets, uint32_t* ids, int nthreads) {
    const Tree& t = *((Handle*)hv)->tree;
    std::atomic<int> bad{0};
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        std::vector<uint32_t> v;
        for (uint64_t i = b; i < e; ++i) {
            v.clear();
            t.all_intersected_primitives(LQ(q, i, kind), v);
            if (v.size() != offsets[i + 1] - offsets[i]) { bad = 1; continue; }
            std::copy(v.begin(), v.end(), ids + offsets[i]);
        }
    });
    return bad ? -1 : 0;
}
// all_intersections(Ray_3) (AABB_tree.h:958-970).  Pass 1 (ids == nullptr): counts[i] = number of intersection objects of ray i.
// Pass 2: offsets = exclusive prefix sums of the counts; entry k of ray i -> ids[offsets[i]+k], kinds[..] (1 Point_3, 2 Segment_3),
// pts[6 * (..)] (point, or the two segment endpoints), in the reference's traversal order.  first_only: any_intersection (:972-985).
int orc_line_intersections(void* hv, const double* rays, uint64_t n, int kind, int first_only, uint32_t* counts, const uint64_t* offsets, uint32_t* ids,
                           uint8_t* kinds, double* pts, int nthreads);
int orc_ray_intersections(void* hv, const double* rays, uint64_t n, int first_only, uint32_t* counts, const uint64_t* offsets, uint32_t* ids,
                          uint8_t* kinds, double* pts, int nthreads) {
    return orc_line_intersections(hv, rays, n, 1, first_only, counts, offsets, ids, kinds, pts, nthreads);
}
// the same for every query kind of LQ(): 0 Segment_3, 1 Ray_3, 2 Line_3 (point, vector), 3 Plane_3 (a, b, c, d: 4 doubles per query); object kind 3 = the
// primitive's Triangle_3 itself (Plane_3 queries; no points)
int orc_line_intersections(void* hv, const double* rays, uint64_t n, int kind, int first_only, uint32_t* counts, const uint64_t* offsets, uint32_t* ids,
                           uint8_t* kinds, double* pts, int nthreads) {
    const Tree& t = *((Handle*)hv)->tree;
    std::atomic<int> bad{0};
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        std::vector<std::pair<uint32_t, RayHit>> v;
        for (uint64_t i = b; i < e; ++i) {
            v.clear();
            if (first_only == 2) {                                   // brute force over all primitives, no tree (the tree-independent set)
                const Mesh& m = ((Handle*)hv)->mesh;
                const Tree::LineQ lq = LQ(rays, i, kind);
                for (uint32_t f = 0; f < (uint32_t)m.tris.size(); ++f) {
                    RayHit r;
                    switch (kind) {
                        case 0: r = seg_tri_intersection(m.tris[f], lq.p, lq.q); break;
                        case 2: r = line_tri_intersection(m.tris[f], lq.p, V3{lq.q.x, lq.q.y, lq.q.z}); break;
                        case 3: { const double h[4] = {lq.p.x, lq.p.y, lq.p.z, lq.q.x}; r = plane_tri_intersection(m.tris[f], h); break; }
                        default: r = ray_tri_intersection(m.tris[f], lq.p, lq.q);
                    }
                    if (r.kind) v.push_back({f, r});
                }
            } else t.query_intersections(LQ(rays, i, kind), first_only != 0, v);
            if (counts) counts[i] = (uint32_t)v.size();
            if (!ids) continue;
            if (v.size() != offsets[i + 1] - offsets[i]) { bad = 1; continue; }
            for (size_t k = 0; k < v.size(); ++k) {
                const uint64_t o = offsets[i] + k;
                ids[o] = v[k].first; kinds[o] = (uint8_t)v[k].second.kind;
                std::memset(pts + 6 * o, 0, 48);
                if (v[k].second.kind == 1 || v[k].second.kind == 2) put(pts + 6 * o, v[k].second.pt[0]);
                if (v[k].second.kind == 2) put(pts + 6 * o + 3, v[k].second.pt[1]);
            }
        }
    });
    return bad ? -1 : 0;
}
// first_intersection(Ray_3).  prim = 0xFFFFFFFF when none.  kind: 0 none, 1 Point_3, 2 Segment_3.
// pts: n x 6 (point, or the two segment endpoints).  tie[i] = 1 if another primitive produced exactly the same t.
void orc_first_intersection(void* hv, const double* rays, uint64_t n, uint32_t* prim, double* tpar, double* pts,
                            uint8_t* kind, uint8_t* tie, int nthreads) {
    const Tree& t = *((Handle*)hv)->tree;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            Tree::FirstHit h = t.first_intersection(P(rays + 6 * i), P(rays + 6 * i + 3));
            prim[i] = h.hit ? h.prim : 0xFFFFFFFFu;
            if (tpar) tpar[i] = h.hit ? h.t : 0.0;
            if (kind) kind[i] = (uint8_t)h.r.kind;
            if (tie) tie[i] = h.t_tie;
            if (pts) { std::memset(pts + 6 * i, 0, 48); if (h.hit) { put(pts + 6 * i, h.r.pt[0]); if (h.r.kind == 2) put(pts + 6 * i + 3, h.r.pt[1]); } }
        }
    });
}
// first_intersection(Ray_3, skip): skip = one primitive id per ray (0xFFFFFFFF: none) and / or a bit mask over primitive ids
// Side_of_triangle_mesh: side[i] in {-1 ON_UNBOUNDED_SIDE, 0 ON_BOUNDARY, 1 ON_BOUNDED_SIDE}; attempts (optional) = rays cast
void orc_side_of_mesh(void* hv, const double* pts, uint64_t n, int8_t* side, uint8_t* attempts, int nthreads) {
    const Tree& t = *((Handle*)hv)->tree;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) { int a = 0; side[i] = (int8_t)t.side_of_mesh(P(pts + 3 * i), &a); if (attempts) attempts[i] = (uint8_t)std::min(a, 255); }
    });
}
int orc_classify_ray_tri(const double* tri, const double* pq) { return classify_ray_tri(T(tri), P(pq), P(pq + 3)); }
void orc_first_intersection_skip(void* hv, const double* rays, uint64_t n, const uint32_t* skip_ids, const uint32_t* mask, uint32_t mask_words, uint32_t* prim, double* tpar,
                                 int nthreads) {
    const Tree& t = *((Handle*)hv)->tree;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            const uint32_t sid = skip_ids ? skip_ids[i] : 0xFFFFFFFFu;
            auto h = t.first_intersection(P(rays + 6 * i), P(rays + 6 * i + 3),
                                          [&](uint32_t f) { return f == sid || (mask && (f >> 5) < mask_words && ((mask[f >> 5] >> (f & 31u)) & 1u)); });
            prim[i] = h.prim;
            tpar[i] = h.hit ? h.t : 0.0;
        }
    });
}
// closest_point_and_primitive(Point_3) (+ squared_distance).  Tree must be non-empty (CGAL_precondition(!empty())).
int orc_closest_point(void* hv, const double* q, uint64_t n, double* cp, uint32_t* prim, double* sqd, int nthreads) {
    const Tree& t = *((Handle*)hv)->tree;
    if (t.size() == 0) return -1;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            const P3 p = P(q + 3 * i);
            auto r = t.closest_point_and_primitive(p);
            if (cp) put(cp + 3 * i, r.first);
            if (prim) prim[i] = r.second;
            if (sqd) sqd[i] = sqdist(p, r.first);
        }
    });
    return 0;
}

// ---- brute-force counterparts (AABB_tree/test/AABB_tree/AABB_test_util.h:572-676,755-786; aabb_test_ray_intersection.cpp:38-65)
void orc_naive_count(void* hv, const double* q, uint64_t n, int kind, uint32_t* out, int nthreads) {
    const Mesh& m = ((Handle*)hv)->mesh;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t c = 0; const Tree::LineQ lq = LQ(q, i, kind);
            for (const Tri& t : m.tris) c += (m.segments ? Tree::segment_hit(t, lq) : Tree::prim_hit(t, lq)) ? 1 : 0;
            out[i] = c;
        }
    });
}
// argmin over all primitives of the exact squared distance of the rounded candidate; ntie = #primitives attaining it
void orc_naive_closest(void* hv, const double* q, uint64_t n, double* cp, uint32_t* prim, uint32_t* ntie, int nthreads) {
    const Mesh& m = ((Handle*)hv)->mesh;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            const P3 p = P(q + 3 * i);
            auto proj = [&](uint32_t f) { return m.segments ? project_on_segment(m.tris[f].a, m.tris[f].b, p) : project_on_triangle(m.tris[f], p); };
            P3 best = proj(0); uint32_t arg = 0, ties = 1;
            for (uint32_t f = 1; f < m.tris.size(); ++f) {
                const P3 c = proj(f);
                const int s = compare_distance(p, c, best);
                if (s < 0) { best = c; arg = f; ties = 1; } else if (s == 0) ++ties;
            }
            if (cp) put(cp + 3 * i, best);
            if (prim) prim[i] = arg;
            if (ntie) ntie[i] = ties;
        }
    });
}
// argmin-t over all intersections, first (lowest face id) wins on equal t; ntie = #primitives attaining min t
void orc_naive_first_intersection(void* hv, const double* rays, uint64_t n, uint32_t* prim, double* tpar, uint32_t* ntie, int nthreads) {
    const Mesh& m = ((Handle*)hv)->mesh;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            const P3 s = P(rays + 6 * i), q = P(rays + 6 * i + 3);
            const V3 v = vsub(q, s); const int mi = Tree::ray_max_axis(v);
            double t = std::numeric_limits<double>::max(); uint32_t arg = 0xFFFFFFFFu, ties = 0;
            for (uint32_t f = 0; f < m.tris.size(); ++f) {
                RayHit r = ray_tri_intersection(m.tris[f], s, q);
                if (!r.kind) continue;
                const double d = Tree::hit_param(s, v, mi, r);
                if (d < t) { t = d; arg = f; ties = 1; } else if (d == t) ++ties;
            }
            prim[i] = arg; if (tpar) tpar[i] = arg == 0xFFFFFFFFu ? 0.0 : t; if (ntie) ntie[i] = ties;
        }
    });
}

// ---- per-primitive candidates, for the parity harness' tie-set checks
void orc_candidate_closest(void* hv, uint32_t face, const double* q, double* cp) { put(cp, project_on_triangle(((Handle*)hv)->mesh.tris[face], P(q))); }
int orc_candidate_ray_hit(void* hv, uint32_t face, const double* ray, double* tpar, double* pts) {
    const P3 s = P(ray), q = P(ray + 3);
    RayHit r = ray_tri_intersection(((Handle*)hv)->mesh.tris[face], s, q);
    if (r.kind) { const V3 v = vsub(q, s); *tpar = Tree::hit_param(s, v, Tree::ray_max_axis(v), r); put(pts, r.pt[0]); if (r.kind == 2) put(pts + 3, r.pt[1]); }
    return r.kind;
}

// ---- predicate / construction level (KATs and python-fractions checks)
int orc_orientation(const double* p4) { return orientation(P(p4), P(p4 + 3), P(p4 + 6), P(p4 + 9)); }
int orc_orientation_exact(const double* p4) { return orientation_exact(P(p4), P(p4 + 3), P(p4 + 6), P(p4 + 9)); }
int orc_orientation_dir(const double* p5) { return orientation_dir(P(p5), P(p5 + 3), P(p5 + 6), P(p5 + 9), P(p5 + 12)); }
int orc_coplanar_orientation3(const double* p3) { return coplanar_orientation(P(p3), P(p3 + 3), P(p3 + 6)); }
int orc_coplanar_orientation4(const double* p4) { return coplanar_orientation(P(p4), P(p4 + 3), P(p4 + 6), P(p4 + 9)); }
int orc_compare_distance(const double* p3) { return compare_distance(P(p3), P(p3 + 3), P(p3 + 6)); }
int orc_compare_distance_exact(const double* p3) { return compare_distance_exact(P(p3), P(p3 + 3), P(p3 + 6)); }
int orc_bbox_do_intersect(const double* box, const double* pq, int is_ray) { return do_intersect_bbox(P(pq), P(pq + 3), B(box), is_ray != 0); }
int orc_bbox_do_intersect_exact(const double* box, const double* pq, int is_ray) { return bbox_segment_aux<true>(P(pq), P(pq + 3), B(box), true, !is_ray).v; }
int orc_tri_do_intersect_kind(const double* tri, const double* q, int kind) { return Tree::prim_hit(T(tri), LQ(q, 0, kind)) ? 1 : 0; }
int orc_bbox_do_intersect_kind(const double* box, const double* q, int kind) { return Tree::box_hit(LQ(q, 0, kind), B(box)) ? 1 : 0; }
// Do_intersect_3(Triangle_3 query, Triangle_3 primitive)
int orc_tri_tri_do_intersect(const double* query, const double* tri) { return do_intersect_tri_tri(T(query), T(tri)) ? 1 : 0; }
// tree.do_intersect(other_tree) (AABB_traits_3.h:679-690): the traversal of `hv` with the other tree as the query — other.do_intersect(bbox) per node (here: the
// other tree's root box against the node's box, then the primitives decide), other.do_intersect(datum) per primitive
int orc_do_intersect_tree(void* hv, void* other) {
    const Handle* h = (const Handle*)hv; const Tree& o = *((Handle*)other)->tree;
    for (const Tri& t : h->mesh.tris) if (o.do_intersect(Tree::LineQ{t.a, t.b, 4, t.c})) return 1;
    return 0;
}
// Rigid_triangle_mesh_collision_detection::does_A_intersect_B (Rigid_triangle_mesh_collision_detection.h:149-161) with the traversal traits of
// Polygon_mesh_processing/internal/AABB_traversal_traits_with_transformation.h: tree B is traversed with tree A as the query (:262-306); every primitive of B,
// transformed by B's Aff_transformation_3, traverses A under A's transformation (:141-156): do_intersect(query, datum.transform(transfo)) at the primitives,
// the transformed node box at the nodes (:36-68: the box of the 8 transformed corners in interval arithmetic; here the corners in binary64, widened by a few
// ulps, against the box of the query — a necessary condition, the primitives decide).  xf: 12 doubles, m(i, j) row major; transform(Point_3) as
// Cartesian/Aff_transformation_rep_3.h:111-117.
static inline P3 xf_point(const double* m, const P3& p) {
    return {m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7], m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]};
}
static inline Tri xf_tri(const double* m, const Tri& t) { return {xf_point(m, t.a), xf_point(m, t.b), xf_point(m, t.c)}; }
namespace {
struct XfTT {
    const Tree& t; const double* m; bool found = false;
    bool go_further() const { return !found; }
    void intersection(const Tree::LineQ& q, uint32_t f) { if (do_intersect_tri_tri(Tri{q.p, q.q, q.r}, xf_tri(m, t.mesh->datum(f)))) found = true; }
    bool do_intersect(const Tree::LineQ& q, const Box& b) const {
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, mag = 0;
        for (int c = 0; c < 8; ++c) {
            const P3 w = xf_point(m, P3{c & 1 ? b.hi[0] : b.lo[0], c & 2 ? b.hi[1] : b.lo[1], c & 4 ? b.hi[2] : b.lo[2]});
            for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], w[k]); hi[k] = std::max(hi[k], w[k]); }
        }
        for (int k = 0; k < 3; ++k)
            mag = std::max(mag, (std::fabs(m[4 * k]) * std::max(std::fabs(b.lo[0]), std::fabs(b.hi[0])) + std::fabs(m[4 * k + 1]) * std::max(std::fabs(b.lo[1]), std::fabs(b.hi[1])) +
                                 std::fabs(m[4 * k + 2]) * std::max(std::fabs(b.lo[2]), std::fabs(b.hi[2]))) + std::fabs(m[4 * k + 3]));
        for (int k = 0; k < 3; ++k) {
            const double qlo = std::min(q.p[k], std::min(q.q[k], q.r[k])), qhi = std::max(q.p[k], std::max(q.q[k], q.r[k]));
            if (qlo > hi[k] + 1e-14 * mag || qhi < lo[k] - 1e-14 * mag) return false;
        }
        return true;
    }
};
}  // namespace
int orc_do_intersect_tree_transformed(void* hv, const double* xf, void* other, const double* other_xf) {
    const Handle* a = (const Handle*)hv; const Handle* b = (const Handle*)other;
    for (const Tri& t : b->mesh.tris) {
        const Tri q = xf_tri(other_xf, t);
        XfTT tt{*a->tree, xf};
        a->tree->traversal(Tree::LineQ{q.a, q.b, 4, q.c}, tt);
        if (tt.found) return 1;
    }
    return 0;
}
// the same decided by brute force over all pairs of transformed primitives
int orc_naive_do_intersect_transformed(void* hv, const double* xf, void* other, const double* other_xf) {
    const Handle* a = (const Handle*)hv; const Handle* b = (const Handle*)other;
    std::vector<Tri> ta; ta.reserve(a->mesh.tris.size());
    for (const Tri& t : a->mesh.tris) ta.push_back(xf_tri(xf, t));
    for (const Tri& t : b->mesh.tris) { const Tri q = xf_tri(other_xf, t); for (const Tri& u : ta) if (do_intersect_tri_tri(q, u)) return 1; }
    return 0;
}
int orc_tri_do_intersect(const double* tri, const double* pq, int is_ray) {
    return (is_ray ? do_intersect_ray_tri(T(tri), P(pq), P(pq + 3)) : do_intersect_seg_tri(T(tri), P(pq), P(pq + 3))) ? 1 : 0;
}
int orc_tri_query_intersection(const double* tri, const double* q, int kind, double* pts) {      // kind 2 Line_3, 3 Plane_3
    RayHit r;
    if (kind == 2) r = line_tri_intersection(T(tri), P(q), V3{q[3], q[4], q[5]});
    else r = plane_tri_intersection(T(tri), q);
    if (r.kind == 1 || r.kind == 2) { put(pts, r.pt[0]); if (r.kind == 2) put(pts + 3, r.pt[1]); }
    return r.kind;
}
int orc_tri_seg_intersection(const double* tri, const double* seg, double* pts) {
    RayHit r = seg_tri_intersection(T(tri), P(seg), P(seg + 3));
    if (r.kind) { put(pts, r.pt[0]); if (r.kind == 2) put(pts + 3, r.pt[1]); }
    return r.kind;
}
int orc_tri_ray_intersection(const double* tri, const double* ray, double* pts) {
    RayHit r = ray_tri_intersection(T(tri), P(ray), P(ray + 3));
    if (r.kind) { put(pts, r.pt[0]); if (r.kind == 2) put(pts + 3, r.pt[1]); }
    return r.kind;
}
void orc_project_on_triangle(const double* tri, const double* q, double* out) { put(out, project_on_triangle(T(tri), P(q))); }
int orc_ball_reaches_box(const double* q, const double* box, const double* bound) { return ball_reaches_box(P(q), B(box), P(bound)) ? 1 : 0; }
int orc_intersection_distance(const double* ray, const double* box, double* t) {
    auto d = intersection_distance(P(ray), vsub(P(ray + 3), P(ray)), B(box));
    if (d) *t = *d;
    return d ? 1 : 0;
}
void orc_stats(uint64_t* out) { out[0] = g_stats.orient_calls; out[1] = g_stats.orient_exact; out[2] = g_stats.cmpdist_calls; out[3] = g_stats.cmpdist_exact; out[4] = g_stats.bbox_exact; }
// child boxes / primitives tested by the batched queries since the last call (all threads)
void orc_traversal_totals(uint64_t* out) { out[0] = g_box_tests.exchange(0); out[1] = g_prim_tests.exchange(0); }
int orc_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"

// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under cgal_b200/ may include, link or call this.
//
// CPU restatement of CGAL::AABB_tree< AABB_traits_3< Epick, AABB_face_graph_triangle_primitive<Surface_mesh> > >:
// median-split build, recursive traversals, best-first first_intersection, kd-tree hint.
// Follows (paths relative to the CGAL source root):
//   AABB_tree/include/CGAL/AABB_tree.h:513-526,582-612,787-853,858-1055
//   AABB_tree/include/CGAL/AABB_tree/internal/AABB_node.h:129-167
//   AABB_tree/include/CGAL/AABB_tree/internal/AABB_traversal_traits.h:150-184,236-327
//   AABB_tree/include/CGAL/AABB_tree/internal/AABB_ray_intersection.h:49-225
//   AABB_tree/include/CGAL/AABB_traits_3.h:596-654,764-816
// Parity status: results pinned by the reference's KATs (tests/test_oracle_kats.py); tree LAYOUT is unpinned
// (no reference test fixes the std::nth_element permutation; it only matters for exact-distance ties).
#pragma once
#include <cstdint>
#include <limits>
#include <memory>
#include <queue>
#include <vector>
#include "kernel.hpp"

namespace orc {

struct Mesh {
    // face i = OFF face i, triangle = (v0,v1,v2) in file order, reference point = v0:
    // BGL/include/CGAL/boost/graph/IO/Generic_facegraph_builder.h:130-148, Euler_operations.h:842-843,
    // BGL/include/CGAL/boost/graph/property_maps.h:52-63,158-163
    std::vector<Tri> tris;
    // segments = true: the primitives are Segment_3 (AABB_segment_primitive_3.h:40-95 / AABB_halfedge_graph_segment_primitive.h), kept as (source, target,
    // target): the box of that record is the segment's (Construct_bbox_3(Segment_3)), the reference point is the source
    bool segments = false;
    const Tri& datum(uint32_t f) const { return tris[f]; }
    const P3& reference_point(uint32_t f) const { return tris[f].a; }
};

static inline Box empty_box() {   // Bbox_3(): Kernel_23/include/CGAL/Bbox_3.h:42-49
    const double inf = std::numeric_limits<double>::infinity();
    return {{inf, inf, inf}, {-inf, -inf, -inf}};
}
static inline Box tri_box(const Tri& t) {   // Construct_bbox_3(Triangle_3): Cartesian/function_objects.h:1962-1983
    Box b;
    for (int i = 0; i < 3; ++i) {
        b.lo[i] = std::min(t.a[i], std::min(t.b[i], t.c[i]));
        b.hi[i] = std::max(t.a[i], std::max(t.b[i], t.c[i]));
    }
    return b;
}
static inline void box_add(Box& b, const Box& o) {   // Bbox_3::operator+=: Bbox_3.h:180-189
    for (int i = 0; i < 3; ++i) { b.lo[i] = std::min(b.lo[i], o.lo[i]); b.hi[i] = std::max(b.hi[i], o.hi[i]); }
}

// 1-nearest-neighbour structure over the primitives' reference points: stands in for AABB_search_tree
// (AABB_tree/internal/AABB_search_tree.h:26-72 over Spatial_searching Kd_tree / Orthogonal_k_neighbor_search, k = 1).
// Any exact 1-NN reproduces it up to equidistant points.
struct KdHint {
    struct Node { int axis; double split; int left, right, begin, end; };
    std::vector<Node> nodes;
    std::vector<P3> pts;
    std::vector<uint32_t> ids;
    void build(std::vector<P3> p, std::vector<uint32_t> id) {
        pts = std::move(p); ids = std::move(id); nodes.clear();
        if (!pts.empty()) { std::vector<int> perm(pts.size()); for (size_t i = 0; i < perm.size(); ++i) perm[i] = (int)i;
            rec(perm, 0, (int)perm.size());
            std::vector<P3> np(pts.size()); std::vector<uint32_t> ni(ids.size());
            for (size_t i = 0; i < perm.size(); ++i) { np[i] = pts[perm[i]]; ni[i] = ids[perm[i]]; }
            pts.swap(np); ids.swap(ni); }
    }
    int rec(std::vector<int>& perm, int b, int e) {
        int me = (int)nodes.size(); nodes.push_back({-1, 0, -1, -1, b, e});
        if (e - b <= 8) return me;
        double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int i = b; i < e; ++i) for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], pts[perm[i]][k]); hi[k] = std::max(hi[k], pts[perm[i]][k]); }
        int ax = 0; if (hi[1] - lo[1] > hi[ax] - lo[ax]) ax = 1; if (hi[2] - lo[2] > hi[ax] - lo[ax]) ax = 2;
        if (hi[ax] == lo[ax]) return me;
        int m = (b + e) / 2;
        std::nth_element(perm.begin() + b, perm.begin() + m, perm.begin() + e, [&](int x, int y) { return pts[x][ax] < pts[y][ax]; });
        nodes[me].axis = ax; nodes[me].split = pts[perm[m]][ax];
        int l = rec(perm, b, m); int r = rec(perm, m, e);
        nodes[me].left = l; nodes[me].right = r;
        return me;
    }
    void search(int n, const P3& q, double& best, int& arg) const {
        const Node& nd = nodes[n];
        if (nd.axis < 0) {
            for (int i = nd.begin; i < nd.end; ++i) { double d = sqdist(q, pts[i]); if (d < best) { best = d; arg = i; } }
            return;
        }
        const double diff = q[nd.axis] - nd.split;
        const int first = diff < 0 ? nd.left : nd.right, second = diff < 0 ? nd.right : nd.left;
        search(first, q, best, arg);
        if (diff * diff <= best) search(second, q, best, arg);
    }
    std::pair<P3, uint32_t> closest(const P3& q) const {
        double best = std::numeric_limits<double>::infinity(); int arg = 0;
        search(0, q, best, arg);
        return {pts[arg], ids[arg]};
    }
};

struct Tree {
    struct Node { Box bbox; int32_t left, right; };   // children: node index or position in `prims` (implicit by nb_primitives)
    const Mesh* mesh = nullptr;
    std::vector<uint32_t> prims;      // m_primitives (face ids), permuted by the build
    std::vector<Node> nodes;          // m_nodes, root first
    std::vector<Box> prim_box;        // cached datum.bbox() per face (values identical to recomputing)
    std::unique_ptr<KdHint> hint;     // m_p_search_tree
    bool use_default_search_tree = true;

    size_t size() const { return prims.size(); }

    explicit Tree(const Mesh& m) : mesh(&m) {
        prims.resize(m.tris.size());
        for (size_t i = 0; i < prims.size(); ++i) prims[i] = (uint32_t)i;
        prim_box.resize(prims.size());
        for (size_t i = 0; i < prims.size(); ++i) prim_box[i] = tri_box(m.tris[i]);
        build();
    }

    // AABB_tree::custom_build / expand: AABB_tree.h:787-853
    void build() {
        nodes.clear();
        if (prims.size() > 1) { nodes.reserve(prims.size() - 1); nodes.emplace_back(); expand(0, 0, prims.size(), prims.size()); }
    }
    static int longest_axis(const Box& b) {   // AABB_traits_3.h:786-816
        const double dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
        if (dx >= dy) return dx >= dz ? 0 : 2;
        return dy >= dz ? 1 : 2;
    }
    void expand(int node, size_t first, size_t beyond, size_t range) {
        Box bb = empty_box();
        for (size_t i = first; i < beyond; ++i) box_add(bb, prim_box[prims[i]]);   // Compute_bbox: AABB_traits_3.h:637-654
        nodes[node].bbox = bb;
        // Split_primitives: AABB_traits_3.h:596-626 (same std::nth_element call, comparator on reference points :764-778)
        const int ax = longest_axis(bb);
        const Mesh* m = mesh;
        std::nth_element(prims.begin() + first, prims.begin() + first + (beyond - first) / 2, prims.begin() + beyond,
                         [m, ax](uint32_t f1, uint32_t f2) { return m->reference_point(f1)[ax] < m->reference_point(f2)[ax]; });
        switch (range) {
        case 2: nodes[node].left = (int32_t)first; nodes[node].right = (int32_t)first + 1; break;
        case 3: {
            nodes[node].left = (int32_t)first;
            int r = (int)nodes.size(); nodes.emplace_back(); nodes[node].right = r;
            expand(r, first + 1, beyond, 2);
            break;
        }
        default: {
            const size_t new_range = range / 2;
            int l = (int)nodes.size(); nodes.emplace_back();
            int r = (int)nodes.size(); nodes.emplace_back();
            nodes[node].left = l; nodes[node].right = r;
            expand(l, first, first + new_range, new_range);
            expand(r, first + new_range, beyond, range - new_range);
        }
        }
    }
    Box bbox() const {   // AABB_tree::bbox(): AABB_tree.h:204-210
        if (size() == 0) return empty_box();
        if (size() == 1) return prim_box[prims[0]];
        return nodes[0].bbox;
    }

    // accelerate_distance_queries / build_kd_tree: AABB_tree.h:858-901
    void accelerate_distance_queries() {
        use_default_search_tree = true;
        if (prims.empty()) return;
        std::vector<P3> p; std::vector<uint32_t> id;
        p.reserve(prims.size()); id.reserve(prims.size());
        for (uint32_t f : prims) { p.push_back(mesh->reference_point(f)); id.push_back(f); }
        hint.reset(new KdHint()); hint->build(std::move(p), std::move(id));
    }
    void do_not_accelerate_distance_queries() { hint.reset(); use_default_search_tree = false; }
    std::pair<P3, uint32_t> best_hint(const P3& q) const {   // AABB_tree.h:591-612 (the lazy build is done eagerly by the caller)
        if (hint) return hint->closest(q);
        return {mesh->reference_point(prims[0]), prims[0]};  // any_reference_point_and_id :582-588
    }

    // AABB_node::traversal: internal/AABB_node.h:129-167 ; AABB_tree::traversal: AABB_tree.h:513-526
    template <class Q, class TT> static bool box_test(TT& tt, const Q& q, const Box& b) { ++g_stats.box_tests; return tt.do_intersect(q, b); }
    template <class Q, class TT> static void prim_test(TT& tt, const Q& q, uint32_t f) { ++g_stats.prim_tests; tt.intersection(q, f); }
    template <class Q, class TT> void traverse_node(int n, const Q& q, TT& tt, size_t nb) const {
        const Node& nd = nodes[n];
        switch (nb) {
        case 2:
            prim_test(tt, q, prims[nd.left]);
            if (tt.go_further()) prim_test(tt, q, prims[nd.right]);
            break;
        case 3:
            prim_test(tt, q, prims[nd.left]);
            if (tt.go_further() && box_test(tt, q, nodes[nd.right].bbox)) traverse_node(nd.right, q, tt, 2);
            break;
        default:
            if (box_test(tt, q, nodes[nd.left].bbox)) {
                traverse_node(nd.left, q, tt, nb / 2);
                if (tt.go_further() && box_test(tt, q, nodes[nd.right].bbox)) traverse_node(nd.right, q, tt, nb - nb / 2);
            } else if (box_test(tt, q, nodes[nd.right].bbox)) {
                traverse_node(nd.right, q, tt, nb - nb / 2);
            }
        }
    }
    template <class Q, class TT> void traversal(const Q& q, TT& tt) const {
        switch (size()) {
        case 0: break;
        case 1: prim_test(tt, q, prims[0]); break;
        default: traverse_node(0, q, tt, size());
        }
    }

    // kind 0: Segment_3 (source, target); 1: Ray_3 (source, second point); 2: Line_3 (point p, direction vector q); 3: Plane_3 (a, b, c = p; d = q.x)
    //      4: Triangle_3 (p, q, r)
    struct LineQ { P3 p, q; int kind; P3 r = {0, 0, 0}; };
    // Segment_3 datum (tr = (source, target, target)): Do_intersect_3(Plane_3, Segment_3) (Plane_3_Segment_3_do_intersect.h:22-40: the two oriented sides) and
    // Do_intersect_3(Triangle_3, Segment_3) (Segment_3_Triangle_3_do_intersect.h:177-262 with the QUERY as the triangle)
    static bool segment_hit(const Tri& tr, const LineQ& q) {
        if (q.kind == 4) return do_intersect_seg_tri(Tri{q.p, q.q, q.r}, tr.a, tr.b);
        const double h[4] = {q.p.x, q.p.y, q.p.z, q.q.x};
        const int s = plane_oriented_side(h, tr.a), t = plane_oriented_side(h, tr.b);
        return !(s == t && t != 0);
    }
    static bool prim_hit(const Tri& tr, const LineQ& q) {
        switch (q.kind) {
            case 4: return do_intersect_tri_tri(Tri{q.p, q.q, q.r}, tr);
            case 1: return do_intersect_ray_tri(tr, q.p, q.q);
            case 2: return do_intersect_line_tri(tr, q.p, V3{q.q.x, q.q.y, q.q.z});
            case 3: { const double h[4] = {q.p.x, q.p.y, q.p.z, q.q.x}; return do_intersect_plane_tri(tr, h); }
            default: return do_intersect_seg_tri(tr, q.p, q.q);
        }
    }
    static bool box_hit(const LineQ& q, const Box& b) {
        switch (q.kind) {
            case 4: {   // Do_intersect_3(Bbox_3, Triangle_3) (Bbox_3_Triangle_3_do_intersect.h, separating axes) is NOT restated: the box of the query against
                        // the node's box is a necessary condition, the primitives decide exactly, so every result but the traversal order is the reference's
                for (int k = 0; k < 3; ++k) {
                    const double lo = std::min(q.p[k], std::min(q.q[k], q.r[k])), hi = std::max(q.p[k], std::max(q.q[k], q.r[k]));
                    if (lo > b.hi[k] || hi < b.lo[k]) return false;
                }
                return true;
            }
            case 2: return do_intersect_line_bbox(q.p, V3{q.q.x, q.q.y, q.q.z}, b);
            case 3: { const double h[4] = {q.p.x, q.p.y, q.p.z, q.q.x}; return do_intersect_plane_bbox(h, b); }
            default: return do_intersect_bbox(q.p, q.q, b, q.kind == 1);
        }
    }

    bool hit(uint32_t f, const LineQ& q) const { return mesh->segments ? segment_hit(mesh->datum(f), q) : prim_hit(mesh->datum(f), q); }
    // Do_intersect_traits :236-271 / Listing_primitive_traits + Counting_output_iterator :26-50,150-184
    struct DoIntersectTT {
        const Tree& t; bool found = false;
        bool go_further() const { return !found; }
        void intersection(const LineQ& q, uint32_t f) {
            if (t.hit(f, q)) found = true;
        }
        bool do_intersect(const LineQ& q, const Box& b) const { return box_hit(q, b); }
    };
    struct CountTT {
        const Tree& t; uint32_t count = 0; std::vector<uint32_t>* ids = nullptr;
        bool go_further() const { return true; }
        void intersection(const LineQ& q, uint32_t f) {
            if (t.hit(f, q)) { ++count; if (ids) ids->push_back(f); }
        }
        bool do_intersect(const LineQ& q, const Box& b) const { return box_hit(q, b); }
    };
    // First_primitive_traits :188-231 (any_intersected_primitive, AABB_tree.h:947-956): the first intersected primitive in traversal order
    struct FirstPrimitiveTT {
        const Tree& t; bool found = false; uint32_t id = 0xFFFFFFFFu;
        bool go_further() const { return !found; }
        void intersection(const LineQ& q, uint32_t f) {
            if (t.hit(f, q)) { found = true; id = f; }
        }
        bool do_intersect(const LineQ& q, const Box& b) const { return box_hit(q, b); }
    };
    uint32_t any_intersected_primitive(const LineQ& q) const { FirstPrimitiveTT tt{*this}; traversal(q, tt); return tt.id; }
    // Listing_primitive_traits :150-184 (all_intersected_primitives, AABB_tree.h:934-945): ids in traversal order
    void all_intersected_primitives(const LineQ& q, std::vector<uint32_t>& out) const { CountTT tt{*this}; tt.ids = &out; traversal(q, tt); }
    // Listing_intersection_traits :105-144 (all_intersections, AABB_tree.h:958-970) and First_intersection_traits :58-101
    // (any_intersection, :972-985) for Ray_3: every primitive whose Intersect_3(Ray_3, Triangle_3) is non-empty, with the object
    struct RayObjectsTT {
        const Tree& t; bool first_only; bool found = false;
        std::vector<std::pair<uint32_t, RayHit>>* out;
        bool go_further() const { return !(first_only && found); }
        void intersection(const LineQ& q, uint32_t f) {
            RayHit r;
            switch (q.kind) {
                case 0: r = seg_tri_intersection(t.mesh->datum(f), q.p, q.q); break;
                case 2: r = line_tri_intersection(t.mesh->datum(f), q.p, V3{q.q.x, q.q.y, q.q.z}); break;
                case 3: { const double h[4] = {q.p.x, q.p.y, q.p.z, q.q.x}; r = plane_tri_intersection(t.mesh->datum(f), h); break; }
                default: r = ray_tri_intersection(t.mesh->datum(f), q.p, q.q);
            }
            if (r.kind) { found = true; out->push_back({f, r}); }
        }
        bool do_intersect(const LineQ& q, const Box& b) const { return box_hit(q, b); }
    };
    // kind 1: Ray_3 (source, second point); kind 0: Segment_3 (source, target), Intersect_3(Segment_3, Triangle_3)
    void ray_intersections(const P3& s, const P3& q, bool first_only, std::vector<std::pair<uint32_t, RayHit>>& out, int kind = 1) const {
        RayObjectsTT tt{*this, first_only, false, &out};
        traversal(LineQ{s, q, kind}, tt);
    }
    void query_intersections(const LineQ& q, bool first_only, std::vector<std::pair<uint32_t, RayHit>>& out) const {   // kinds 0..3 of LineQ
        RayObjectsTT tt{*this, first_only, false, &out};
        traversal(q, tt);
    }
    // ---- Side_of_triangle_mesh (Polygon_mesh_processing/include/CGAL/Side_of_triangle_mesh.h:265-306) ----------------------------------
    // Ray_3_Triangle_3_traversal_traits<..., Tag_false> (internal/Side_of_triangle_mesh/Ray_3_Triangle_3_traversal_traits.h:30-76): any ray
    struct SideTT {
        const Tree& t; int status = 1 /* tribool: 1 true, 0 false, 2 indeterminate */; size_t count = 0; bool stop = false;
        bool go_further() const { return !stop; }
        void intersection(const LineQ& q, uint32_t f) {
            const int k = classify_ray_tri(t.mesh->datum(f), q.p, q.q);
            if (k < 0) return;
            if (k == 0) ++count;
            else if (k == 4) { status = 0; stop = true; }
            else { status = 2; stop = true; }
        }
        bool do_intersect(const LineQ& q, const Box& b) const { return do_intersect_bbox(q.p, q.q, b, true); }
    };
    // the Tag_true specialisation (:79-222): a vertical ray, box and primitive tests by comparisons and 2D orientations of the z-projection
    struct SideVerticalTT {
        const Tree& t; int status = 1; size_t count = 0; bool stop = false;
        bool go_further() const { return !stop; }
        static bool box_test(const LineQ& q, const Box& b) {                                    // :96-110
            const bool inc_z = q.q.z > q.p.z;
            if ((inc_z && q.p.z > b.hi[2]) || (!inc_z && q.p.z < b.lo[2])) return false;
            if (q.p.x > b.hi[0] || q.p.x < b.lo[0]) return false;
            return q.p.y <= b.hi[1] && q.p.y >= b.lo[1];
        }
        bool do_intersect(const LineQ& q, const Box& b) const { return box_test(q, b); }
        void intersection(const LineQ& q, uint32_t f) {                                          // :129-221
            const Tri tr = t.mesh->datum(f);
            Box tb;
            const P3* v[3] = {&tr.a, &tr.b, &tr.c};
            for (int k = 0; k < 3; ++k) { tb.lo[k] = std::min((*v[0])[k], std::min((*v[1])[k], (*v[2])[k])); tb.hi[k] = std::max((*v[0])[k], std::max((*v[1])[k], (*v[2])[k])); }
            if (!box_test(q, tb)) return;
            int i1 = 1, i2 = 2;
            const int orient_2 = orientation2(v[0]->x, v[0]->y, v[1]->x, v[1]->y, v[2]->x, v[2]->y);
            if (orient_2 == 0) {                     // the facet is vertical: its z-projection is a segment
                const bool same01 = v[0]->x == v[1]->x && v[0]->y == v[1]->y;
                const P3* other = same01 ? v[2] : v[1];
                if (orientation2(v[0]->x, v[0]->y, other->x, other->y, q.p.x, q.p.y) != 0) return;
                status = 2; stop = true;
                return;
            }
            if (orient_2 == -1) std::swap(i1, i2);
            const P3 &p0 = *v[0], &p1 = *v[i1], &p2 = *v[i2];
            const int orient_3 = orientation(p0, p1, p2, q.p);
            const double dz = q.q.z - q.p.z;         // query.to_vector().z()
            if (orient_3 != 0 && ((dz > 0) == (orient_3 == 1))) return;
            const int o01 = orientation2(p0.x, p0.y, p1.x, p1.y, q.p.x, q.p.y);
            if (o01 == 0) { status = 2; stop = true; return; }
            if (o01 == -1) return;
            const int o12 = orientation2(p1.x, p1.y, p2.x, p2.y, q.p.x, q.p.y);
            if (o12 == 0) { status = 2; stop = true; return; }
            if (o12 == -1) return;
            const int o20 = orientation2(p2.x, p2.y, p0.x, p0.y, q.p.x, q.p.y);
            if (o20 == 0) { status = 2; stop = true; return; }
            if (o20 == -1) return;
            if (orient_3 == 0) { status = 0; stop = true; }
            else ++count;
        }
    };
    // Point_inside_vertical_ray_cast::operator() (Point_inside_vertical_ray_cast.h:69-96) behind the bbox test of Side_of_triangle_mesh.h:269-277.
    // The retry directions: the reference draws them from CGAL::Random(1340818006) through Random_points_on_sphere_3 — an un-vendored generator
    // stream that is not restated; any sequence gives the same answer for a closed mesh (this one differs from the device's on purpose).
    // returns -1 ON_UNBOUNDED_SIDE, 0 ON_BOUNDARY, 1 ON_BOUNDED_SIDE; *attempts = rays cast
    int side_of_mesh(const P3& pt, int* attempts = nullptr) const {
        const Box bb = bbox();
        if (attempts) *attempts = 0;
        if (pt.x < bb.lo[0] || pt.x > bb.hi[0] || pt.y < bb.lo[1] || pt.y > bb.hi[1] || pt.z < bb.lo[2] || pt.z > bb.hi[2]) return -1;
        auto verdict = [](int status, size_t count) { return status == 0 ? 0 : ((count & 1) ? 1 : -1); };
        {
            const double dz = 2 * pt.z < bb.hi[2] + bb.lo[2] ? -1.0 : 1.0;
            SideVerticalTT tt{*this};
            traversal(LineQ{pt, P3{pt.x + 0.0, pt.y + 0.0, pt.z + dz}, 1}, tt);
            if (attempts) ++*attempts;
            if (tt.status != 2) return verdict(tt.status, tt.count);
        }
        uint64_t lcg = 1340818006ull;
        for (int k = 0; k < 1000; ++k) {
            double d[3];
            for (int a = 0; a < 3; ++a) { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; d[a] = (double)(lcg >> 11) / 9007199254740992.0 * 2.0 - 1.0; }
            if (d[0] == 0 && d[1] == 0 && d[2] == 0) continue;
            SideTT tt{*this};
            traversal(LineQ{pt, P3{pt.x + d[0], pt.y + d[1], pt.z + d[2]}, 1}, tt);
            if (attempts) ++*attempts;
            if (tt.status != 2) return verdict(tt.status, tt.count);
        }
        return -2;
    }

    bool do_intersect(const LineQ& q) const { DoIntersectTT tt{*this}; traversal(q, tt); return tt.found; }          // AABB_tree.h:903-913
    uint32_t number_of_intersected_primitives(const LineQ& q) const { CountTT tt{*this}; traversal(q, tt); return tt.count; }  // :915-932

    // Projection_traits :277-327
    struct ProjectionTT {
        const Tree& t; P3 closest; uint32_t prim;
        bool go_further() const { return true; }
        void intersection(const P3& q, uint32_t f) {
            const P3 n = t.mesh->segments ? closest_point_step_segment(q, t.mesh->datum(f).a, t.mesh->datum(f).b, closest) : closest_point_step(q, t.mesh->datum(f), closest);
            if (n != closest) { prim = f; closest = n; }
        }
        bool do_intersect(const P3& q, const Box& b) const { return ball_reaches_box(q, b, closest); }
    };
    std::pair<P3, uint32_t> closest_point_and_primitive(const P3& q, const std::pair<P3, uint32_t>& h) const {   // AABB_tree.h:1045-1055
        ProjectionTT tt{*this, h.first, h.second};
        traversal(q, tt);
        return {tt.closest, tt.prim};
    }
    std::pair<P3, uint32_t> closest_point_and_primitive(const P3& q) const { return closest_point_and_primitive(q, best_hint(q)); }
    P3 closest_point(const P3& q) const {   // AABB_tree.h:1004-1011 -> :988-1000 (hint id := m_primitives[0].id())
        return closest_point_and_primitive(q, {best_hint(q).first, prims[0]}).first;
    }
    double squared_distance(const P3& q) const { return sqdist(q, closest_point(q)); }   // AABB_tree.h:1024-1032

    // first_intersection: internal/AABB_ray_intersection.h:49-150,201-225
    struct FirstHit { bool hit = false; uint32_t prim = 0xFFFFFFFFu; RayHit r; double t = 0; bool t_tie = false; };
    struct HeapItem { int node; size_t nb; double value; bool operator>(const HeapItem& o) const { return value > o.value; } };

    static double ray_param(const P3& s, const V3& v, int max_i, const P3& x) { return (x[max_i] - s[max_i]) / v[max_i]; }   // :189-194
    static int ray_max_axis(const V3& v) { int m = 0; for (int i = 1; i < 3; ++i) if (std::fabs(v[i]) > std::fabs(v[m])) m = i; return m; }  // :171-176
    static double hit_param(const P3& s, const V3& v, int max_i, const RayHit& r) {
        if (r.kind == 2) return std::min(ray_param(s, v, max_i, r.pt[0]), ray_param(s, v, max_i, r.pt[1]));   // :178-187
        return ray_param(s, v, max_i, r.pt[0]);
    }
    FirstHit first_intersection(const P3& s, const P3& q) const { return first_intersection(s, q, [](uint32_t) { return false; }); }
    // first_intersection(query, skip): AABB_tree.h:307-328, internal/AABB_ray_intersection.h:201-225 (skip is consulted at every primitive
    // of a tree of two or more, :86,99,113; a one-primitive tree is answered directly, :214)
    template <class Skip> FirstHit first_intersection(const P3& s, const P3& q, Skip skip) const {
        FirstHit best;
        const V3 v = vsub(q, s);
        const int max_i = ray_max_axis(v);
        auto leaf = [&](uint32_t f, double& t) {
            if (skip(f)) return;
            ++g_stats.prim_tests;
            RayHit r = ray_tri_intersection(mesh->datum(f), s, q);
            if (r.kind) {
                const double d = hit_param(s, v, max_i, r);
                if (d < t) { t = d; best.hit = true; best.prim = f; best.r = r; best.t = d; best.t_tie = false; }
                else if (d == t && best.hit && best.prim != f) best.t_tie = true;
            }
        };
        switch (size()) {
        case 0: return best;
        case 1: {   // intersection_object()(query, singleton_data()) — t only reported
            RayHit r = ray_tri_intersection(mesh->datum(prims[0]), s, q);
            if (r.kind) { best.hit = true; best.prim = prims[0]; best.r = r; best.t = hit_param(s, v, max_i, r); }
            return best;
        }
        default: break;
        }
        if (!do_intersect_bbox(s, q, nodes[0].bbox, true)) return best;
        std::priority_queue<HeapItem, std::vector<HeapItem>, std::greater<HeapItem>> pq;
        double t = std::numeric_limits<double>::max();
        pq.push({0, size(), 0.0});
        while (!pq.empty() && pq.top().value < t) {
            HeapItem cur = pq.top(); pq.pop();
            const Node& nd = nodes[cur.node];
            switch (cur.nb) {
            case 2: leaf(prims[nd.left], t); leaf(prims[nd.right], t); break;
            case 3: {
                leaf(prims[nd.left], t);
                ++g_stats.box_tests;
                auto d = intersection_distance(s, v, nodes[nd.right].bbox);
                if (d) pq.push({nd.right, 2, *d});
                break;
            }
            default: {
                g_stats.box_tests += 2;
                auto d = intersection_distance(s, v, nodes[nd.left].bbox);
                if (d) pq.push({nd.left, cur.nb / 2, *d});
                d = intersection_distance(s, v, nodes[nd.right].bbox);
                if (d) pq.push({nd.right, cur.nb - cur.nb / 2, *d});
            }
            }
        }
        return best;
    }
};

}  // namespace orc

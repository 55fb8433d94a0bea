// test_aabb_tree.cpp — exercises CGAL_B200::AABB_tree (include/CGAL_B200/AABB_tree.h) the way the reference's own tests
// exercise CGAL::AABB_tree.  Needs a B200 (run by tests/test_gpu_cpp_host.py under -m gpu); compiled on CPU as a build check.
//   no arguments : the reference's known-answer tests
//                  AABB_tree/test/AABB_tree/aabb_correctness_triangle_test.cpp:51-108, aabb_test_empty_tree.cpp:38-53,
//                  aabb_test_singleton_tree.cpp:44-67, AABB_tree/examples/AABB_tree/AABB_face_graph_triangle_example.cpp:31-39,
//                  plus single-query == batched-query consistency, lazy rebuild and move semantics
//   <in> <out>   : reads a mesh and query sets, answers them through the class, writes the results (compared with the
//                  CPU oracle by the Python test)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <variant>
#include <thread>
#include <list>
#include <vector>

#include "../../include/CGAL_B200/AABB_tree.h"
#include "mini_cgal.h"

typedef mini::Epick K;
typedef K::Point_3 Point;
typedef K::Segment_3 Segment;
typedef K::Ray_3 Ray;
typedef mini::Surface_mesh<Point> Mesh;
typedef mini::AABB_face_graph_triangle_primitive<Mesh> Primitive;
typedef mini::AABB_traits_3<K, Primitive> Traits;
typedef CGAL_B200::AABB_tree<Traits> Tree;

#define CHECK(cond)                                                              \
  do {                                                                           \
    if (!(cond)) {                                                               \
      std::fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                              \
    }                                                                            \
  } while (0)

static Mesh tetrahedron(double dx = 0.0, double dy = 0.0) {   // Data/data/meshes/tetrahedron.off (translated by (dx, dy, 0))
  Mesh m;
  auto a = m.add_vertex(Point(dx, dy, 0)), b = m.add_vertex(Point(1 + dx, dy, 0)), c = m.add_vertex(Point(dx, 1 + dy, 0)), d = m.add_vertex(Point(dx, dy, 1));
  m.add_face(a, b, c); m.add_face(a, d, b); m.add_face(a, c, d); m.add_face(b, d, c);
  return m;
}

static Mesh grid(int n) {   // n x n quads, z = a gentle bump, two triangles per quad
  Mesh m;
  std::vector<mini::Vertex_index> v;
  for (int j = 0; j <= n; ++j)
    for (int i = 0; i <= n; ++i) v.push_back(m.add_vertex(Point(i, j, 0.25 * std::sin(0.7 * i) * std::cos(0.4 * j))));
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      const auto a = v[j * (n + 1) + i], b = v[j * (n + 1) + i + 1], c = v[(j + 1) * (n + 1) + i + 1], d = v[(j + 1) * (n + 1) + i];
      m.add_face(a, b, c); m.add_face(a, c, d);
    }
  return m;
}

static double u01(uint64_t& s) {   // SplitMix64
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

static int known_answers() {
  {   // aabb_correctness_triangle_test.cpp:51-108
    Mesh m = tetrahedron();
    Tree tree(faces(m).first, faces(m).second, m);
    const Point p(-0.25, 0.251, 0.255), q(0.25, 0.253, 0.256);
    const Segment pq(p, q);
    CHECK(tree.do_intersect(pq));
    CHECK(tree.number_of_intersected_primitives(pq) == 1);
    CHECK(tree.closest_point(Point(0.0, 0.0, 3.0)) == Point(0.0, 0.0, 1.0));
    CHECK(tree.size() == 4 && !tree.empty());
    const auto bb = tree.bbox();
    CHECK(bb.xmin() == 0 && bb.ymin() == 0 && bb.zmin() == 0 && bb.xmax() == 1 && bb.ymax() == 1 && bb.zmax() == 1);
    // AABB_face_graph_triangle_example.cpp:31-39 (there on make_tetrahedron((1,0,0),(0,1,0),(0,0,1),(0,0,0)): same solid)
    CHECK(tree.number_of_intersected_primitives(Segment(Point(0.2, 0.2, 0.2), Point(1.0, 1.0, 1.0))) == 1);
    const Point c = tree.closest_point(Point(2.0, 2.0, 2.0));
    CHECK(std::fabs(c.x() - 1.0 / 3) < 1e-15 && std::fabs(c.y() - 1.0 / 3) < 1e-15 && std::fabs(c.z() - 1.0 / 3) < 1e-15);
    CHECK(tree.squared_distance(Point(2.0, 2.0, 2.0)) == (c.x() - 2) * (c.x() - 2) + (c.y() - 2) * (c.y() - 2) + (c.z() - 2) * (c.z() - 2));
    CHECK(tree.accelerate_distance_queries());
    CHECK(tree.any_reference_point_and_id().first == Point(0, 0, 0) && tree.any_reference_point_and_id().second == mini::Face_index(0));   // AABB_tree.h:582-588
    CHECK(tree.best_hint(Point(0.9, 0.1, 0.0)).first == Point(1, 0, 0) && tree.best_hint(Point(0.9, 0.1, 0.0)).second == mini::Face_index(3));  // face 3 = (b, d, c) starts at b
    CHECK(tree.closest_point(Point(2.0, 2.0, 2.0), tree.best_hint(Point(2.0, 2.0, 2.0)).first) == tree.closest_point(Point(2.0, 2.0, 2.0)));
    // first_intersection: a ray along +z from below the base face hits face 0 (z = 0) first, at (0.2,0.2,0)
    const Ray up(Point(0.2, 0.2, -1.0), Point(0.2, 0.2, 0.0));
    const auto hit = tree.first_intersection(up);
    CHECK(hit && hit->second == mini::Face_index(0));
    CHECK(std::holds_alternative<Point>(hit->first) && std::get<Point>(hit->first) == Point(0.2, 0.2, 0.0));
    CHECK(tree.first_intersected_primitive(up) && *tree.first_intersected_primitive(up) == mini::Face_index(0));
    CHECK(tree.number_of_intersected_primitives(up) == 2);
    {   // intersection objects of Line_3 and Plane_3 queries
      std::vector<Tree::Intersection_and_primitive_id<mini::Line_3>::Type> lo;
      tree.all_intersections(mini::Line_3(Point(0.2, 0.2, 5.0), mini::Vector_3(0.0, 0.0, -1.0)), std::back_inserter(lo));
      CHECK(lo.size() == 2 && std::holds_alternative<Point>(lo[0].first) && std::holds_alternative<Point>(lo[1].first));
      std::vector<Tree::Intersection_and_primitive_id<mini::Plane_3>::Type> po;
      tree.all_intersections(mini::Plane_3(0.0, 0.0, 1.0, -0.5), std::back_inserter(po));          // z = 1/2 cuts the three slanted faces in a segment each
      CHECK(po.size() == 3 && std::holds_alternative<Segment>(po[0].first) && std::holds_alternative<Segment>(po[1].first) && std::holds_alternative<Segment>(po[2].first));
      bool threw = false;                                                                           // the base plane contains face 0: a Triangle_3 object, which the stand-in
      try { po.clear(); tree.all_intersections(mini::Plane_3(0.0, 0.0, 1.0, 0.0), std::back_inserter(po)); }   // traits' variant<Point_3, Segment_3> cannot hold
      catch (const CGAL_B200::Error& e) { threw = e.code() == B200AABB_ERR_INVALID_ARGUMENT; }
      CHECK(threw);
    }
    {   // multi-GPU through the C++ class: replicate the tree with the library's NCCL communicator (two GPUs when there are two, else a world of one)
      b200aabb_tree* probe = nullptr;
      const bool two = b200aabb_create(1, &probe) == B200AABB_OK;
      if (probe) b200aabb_destroy(probe);
      try {
        std::vector<int> devices = two ? std::vector<int>{0, 1} : std::vector<int>{0};
        auto comms = CGAL_B200::Communicator::all(devices);
        Tree replica(Traits(), two ? 1 : 0);
        if (two) replica.insert(faces(m).first, faces(m).second, m);               // the same primitives in the same order: they map indices to ids on the host
        std::vector<Tree*> trees = two ? std::vector<Tree*>{&tree, &replica} : std::vector<Tree*>{&tree};
        const double ms = Tree::replicate_all(trees, comms, 0);
        CHECK(ms >= 0.0);
        Tree& r = two ? replica : tree;
        CHECK(r.size() == 4 && r.do_intersect(pq) && r.number_of_intersected_primitives(pq) == 1);
        CHECK(r.closest_point(Point(0.0, 0.0, 3.0)) == Point(0.0, 0.0, 1.0));
        const auto h2 = r.first_intersection(Ray(Point(0.2, 0.2, -1.0), Point(0.2, 0.2, 0.0)));
        CHECK(h2 && h2->second == mini::Face_index(0));
      } catch (const CGAL_B200::Error& e) {
        CHECK(e.code() == B200AABB_ERR_NCCL);                                        // libnccl.so.2 not loadable on this box: the only acceptable failure
      }
    }
    {   // intersection objects of Segment_3 queries (AABB_tree.h:958-985 with Intersect_3(Segment_3, Triangle_3))
      const Segment through(Point(0.2, 0.2, -1.0), Point(0.2, 0.2, 2.0));          // crosses the base facet at z = 0 and the slanted one at z = 0.6
      std::vector<Tree::Intersection_and_primitive_id<Segment>::Type> objs;
      tree.all_intersections(through, std::back_inserter(objs));
      CHECK(objs.size() == 2);
      bool base = false;
      for (const auto& o : objs) {          // (the point is Line_3 x Plane_3 in plain doubles: (den * 0.2) / den need not be 0.2 to the last bit)
        if (!std::holds_alternative<Point>(o.first) || !(o.second == mini::Face_index(0))) continue;
        const Point x = std::get<Point>(o.first);
        base |= std::fabs(x.x() - 0.2) < 1e-15 && std::fabs(x.y() - 0.2) < 1e-15 && std::fabs(x.z()) < 1e-15;
      }
      CHECK(base);
      CHECK(tree.any_intersection(through).has_value() && !tree.any_intersection(Segment(Point(0.2, 0.2, -1.0), Point(0.2, 0.2, -0.5))).has_value());
      const Segment in_base(Point(0.1, 0.1, 0.0), Point(0.3, 0.1, 0.0));            // lies inside the base facet: a Segment_3 object
      objs.clear();
      tree.all_intersections(in_base, std::back_inserter(objs));
      bool seg = false;
      for (const auto& o : objs) seg |= std::holds_alternative<Segment>(o.first) && o.second == mini::Face_index(0);
      CHECK(seg);
    }
    {   // Side_of_triangle_mesh (Side_of_triangle_mesh.h:265-306) on the tetrahedron (0,0,0),(1,0,0),(0,1,0),(0,0,1)
      CGAL_B200::Side_of_triangle_mesh<Tree> inside(tree);
      CHECK(inside(Point(0.1, 0.1, 0.1)) == CGAL_B200::ON_BOUNDED_SIDE);
      CHECK(inside(Point(0.5, 0.5, 0.5)) == CGAL_B200::ON_UNBOUNDED_SIDE);      // inside the bbox, outside the solid
      CHECK(inside(Point(2.0, 0.1, 0.1)) == CGAL_B200::ON_UNBOUNDED_SIDE);
      CHECK(inside(Point(0.25, 0.25, 0.0)) == CGAL_B200::ON_BOUNDARY);          // on the base facet
      CHECK(inside(Point(0.0, 0.0, 0.0)) == CGAL_B200::ON_BOUNDARY);            // a vertex
      CHECK(inside(Point(0.25, 0.25, 0.25)) == CGAL_B200::ON_BOUNDED_SIDE);
      std::vector<Point> ps = {Point(0.1, 0.1, 0.1), Point(0.9, 0.9, 0.9), Point(0.5, 0.0, 0.0)};
      std::vector<CGAL_B200::Bounded_side> rs;
      inside(ps.begin(), ps.end(), std::back_inserter(rs));
      CHECK(rs.size() == 3 && rs[0] == CGAL_B200::ON_BOUNDED_SIDE && rs[1] == CGAL_B200::ON_UNBOUNDED_SIDE && rs[2] == CGAL_B200::ON_BOUNDARY);
    }
    {   // Line_3 and Plane_3 queries (the AABB_tree concept takes any query type the traits can intersect: AABB_traits_3.h:378-400)
      const mini::Line_3 zline(Point(0.2, 0.2, 5.0), mini::Vector_3(0.0, 0.0, -1.0));
      CHECK(tree.do_intersect(zline) && tree.number_of_intersected_primitives(zline) == 2);
      CHECK(!tree.do_intersect(mini::Line_3(Point(2.0, 2.0, 0.0), Point(2.0, 2.0, 1.0))));
      CHECK(tree.number_of_intersected_primitives(mini::Plane_3(0.0, 0.0, 1.0, -0.5)) == 3);     // z = 1/2 misses the base face only
      CHECK(tree.number_of_intersected_primitives(mini::Plane_3(0.0, 0.0, 1.0, 0.0)) == 4);      // the base plane touches every face
      CHECK(!tree.do_intersect(mini::Plane_3(0.0, 0.0, 1.0, -1.5)) && !tree.any_intersected_primitive(mini::Plane_3(1.0, 1.0, 1.0, -1.5)));
      std::vector<mini::Plane_3> planes = {mini::Plane_3(0.0, 0.0, 1.0, -0.5), mini::Plane_3(1.0, 0.0, 0.0, 3.0)};
      std::vector<std::size_t> pc;
      tree.number_of_intersected_primitives(planes.begin(), planes.end(), std::back_inserter(pc));
      CHECK(pc.size() == 2 && pc[0] == 3 && pc[1] == 0);
    }
    {   // Triangle_3 queries (Triangle_3_Triangle_3_do_intersect.h) and tree-vs-tree do_intersect (AABB_traits_3.h:679-690)
      const mini::Triangle_3 cut(Point(-1.0, -1.0, 0.5), Point(3.0, -1.0, 0.5), Point(-1.0, 3.0, 0.5));   // the plane z = 1/2 across the solid
      CHECK(tree.do_intersect(cut) && tree.number_of_intersected_primitives(cut) == 3);
      const mini::Triangle_3 base(Point(0, 0, 0), Point(1, 0, 0), Point(0, 1, 0));                        // face 0 itself: touches every face
      CHECK(tree.number_of_intersected_primitives(base) == 4);
      CHECK(!tree.do_intersect(mini::Triangle_3(Point(2, 2, 2), Point(3, 2, 2), Point(2, 3, 2))));
      CHECK(!tree.any_intersected_primitive(mini::Triangle_3(Point(0.1, 0.1, 0.1), Point(0.2, 0.1, 0.1), Point(0.1, 0.2, 0.1))));   // strictly inside
      std::vector<mini::Face_index> tids;
      tree.all_intersected_primitives(cut, std::back_inserter(tids));
      CHECK(tids.size() == 3 && !(tids[0] == mini::Face_index(0)) && !(tids[1] == mini::Face_index(0)) && !(tids[2] == mini::Face_index(0)));
      Mesh m2 = tetrahedron(0.25, 0.25), m3 = tetrahedron(5.0, 0.0);                                        // overlapping / far away
      Tree near_tree(faces(m2).first, faces(m2).second, m2), far_tree(faces(m3).first, faces(m3).second, m3), empty_tree;
      CHECK(tree.do_intersect(near_tree) && near_tree.do_intersect(tree) && tree.do_intersect(tree));
      CHECK(!tree.do_intersect(far_tree) && !far_tree.do_intersect(tree) && !tree.do_intersect(empty_tree));
      // both meshes under transformations (Rigid_triangle_mesh_collision_detection): the far copy moved back onto the solid, the solid moved away, a quarter turn
      const mini::Aff_transformation_3 identity, back(1, 0, 0, -4.75, 0, 1, 0, 0.25, 0, 0, 1, 0), away(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 7.0),
          turn(0, -1, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0);                                               // (x, y, z) -> (-y, x, z)
      CHECK(tree.do_intersect(far_tree, identity, back) && far_tree.do_intersect(tree, back, identity));
      CHECK(!tree.do_intersect(near_tree, away, identity) && tree.do_intersect(near_tree, away, away));
      CHECK(tree.do_intersect(tree, identity, turn));                                              // the turned copy shares the z axis edge
      CHECK(!tree.do_intersect(far_tree, turn, identity));
      // the collision-detection consumer: three solids, moved by transformations only
      CGAL_B200::Rigid_triangle_mesh_collision_detection<Tree, mini::Aff_transformation_3> cd;
      Mesh s0 = tetrahedron(), s1 = tetrahedron(0.25, 0.25), s2 = tetrahedron(5.0, 0.0);
      const std::size_t i0 = cd.add_mesh(faces(s0).first, faces(s0).second, s0), i1 = cd.add_mesh(faces(s1).first, faces(s1).second, s1),
                        i2 = cd.add_mesh(faces(s2).first, faces(s2).second, s2);
      CHECK(i0 == 0 && i1 == 1 && i2 == 2 && cd.size() == 3);
      CHECK(cd.get_all_intersections(i0) == std::vector<std::size_t>{i1} && cd.get_all_intersections(i2).empty());
      cd.set_transformation(i2, back);                                                             // onto the other two
      CHECK((cd.get_all_intersections(i0) == std::vector<std::size_t>{i1, i2}) && (cd.get_all_intersections(i2) == std::vector<std::size_t>{i0, i1}));
      cd.set_transformation(i0, away);
      CHECK(cd.get_all_intersections(i0).empty() && cd.get_all_intersections(i1) == std::vector<std::size_t>{i2});
      cd.remove_mesh(i1);
      CHECK(cd.size() == 2 && !cd.is_valid_index(i1) && cd.get_all_intersections(i2).empty());
      Mesh s3 = tetrahedron(0.25, 0.25);
      CHECK(cd.add_mesh(faces(s3).first, faces(s3).second, s3) == i1 && cd.get_all_intersections(i2) == std::vector<std::size_t>{i1});
      // inclusions: a solid scaled down to a tenth and placed inside another one touches no face of it, but lies in its volume
      CGAL_B200::Rigid_triangle_mesh_collision_detection<Tree, mini::Aff_transformation_3> in;
      Mesh big = tetrahedron(), small = tetrahedron();
      const std::size_t ib = in.add_mesh_with_points(true, std::vector<Point>{Point(0, 0, 0)}, faces(big).first, faces(big).second, big),
                        is = in.add_mesh_with_points(true, std::vector<Point>{Point(0, 0, 0)}, faces(small).first, faces(small).second, small);
      in.set_transformation(is, mini::Aff_transformation_3(0.1, 0, 0, 0.2, 0, 0.1, 0, 0.2, 0, 0, 0.1, 0.2));     // a tenth of the size, moved to (0.2, 0.2, 0.2)
      typedef std::vector<std::pair<std::size_t, bool>> Pairs;
      CHECK(in.get_all_intersections(ib).empty());
      CHECK((in.get_all_intersections_and_inclusions(ib) == Pairs{{is, true}}) && (in.get_all_intersections_and_inclusions(is) == Pairs{{ib, true}}));
      in.set_transformation(is, mini::Aff_transformation_3(0.1, 0, 0, 0.95, 0, 0.1, 0, 0.0, 0, 0, 0.1, 0.0));    // across the vertex (1, 0, 0)
      CHECK((in.get_all_intersections_and_inclusions(ib) == Pairs{{is, false}}));
      in.set_transformation(is, mini::Aff_transformation_3(0.1, 0, 0, 3.0, 0, 0.1, 0, 3.0, 0, 0, 0.1, 3.0));     // outside
      CHECK(in.get_all_intersections_and_inclusions(ib).empty());
      in.set_transformation(ib, mini::Aff_transformation_3(1, 0, 0, 2.9, 0, 1, 0, 2.9, 0, 0, 1, 2.9));           // the large one follows: inside again
      CHECK((in.get_all_intersections_and_inclusions(ib) == Pairs{{is, true}}));
    }
    {   // all_intersected_primitives / any_intersected_primitive (AABB_tree.h:253-268): the two faces the ray crosses, base face 0 among them
      std::vector<mini::Face_index> ids;
      tree.all_intersected_primitives(up, std::back_inserter(ids));
      CHECK(ids.size() == 2 && (ids[0] == mini::Face_index(0) || ids[1] == mini::Face_index(0)) && !(ids[0] == ids[1]));
      const auto any = tree.any_intersected_primitive(up);
      CHECK(any && (*any == ids[0] || *any == ids[1]));
      CHECK(!tree.any_intersected_primitive(Ray(Point(0.2, 0.2, -1.0), Point(0.2, 0.2, -2.0))));
      std::vector<mini::Face_index> none;
      tree.all_intersected_primitives(Segment(Point(5, 5, 5), Point(6, 6, 6)), std::back_inserter(none));
      CHECK(none.empty());
      // all_intersections / any_intersection (AABB_tree.h:283-296): the ray enters through the base face at (0.2,0.2,0) and leaves
      // through the slanted face x + y + z = 1 at (0.2,0.2,0.6); first_intersection is the nearer of the two
      std::vector<Tree::Intersection_and_primitive_id<Ray>::Type> all;
      tree.all_intersections(up, std::back_inserter(all));
      CHECK(all.size() == 2);
      bool base = false, slant = false;
      for (const auto& x : all) {
        CHECK(std::holds_alternative<Point>(x.first));
        const Point p = std::get<Point>(x.first);
        if (x.second == mini::Face_index(0)) base = p == Point(0.2, 0.2, 0.0);
        else slant = std::fabs(p.x() - 0.2) < 1e-15 && std::fabs(p.y() - 0.2) < 1e-15 && std::fabs(p.z() - 0.6) < 1e-15;
      }
      CHECK(base && slant);
      const auto anyx = tree.any_intersection(up);
      CHECK(anyx && (anyx->second == all[0].second || anyx->second == all[1].second));
      CHECK(!tree.any_intersection(Ray(Point(0.2, 0.2, -1.0), Point(0.2, 0.2, -2.0))));
    }
    CHECK(!tree.first_intersection(Ray(Point(0.2, 0.2, -1.0), Point(0.2, 0.2, -2.0))));
    const auto cpp = tree.closest_point_and_primitive(Point(0.2, 0.2, -1.0));
    CHECK(cpp.first == Point(0.2, 0.2, 0.0) && cpp.second == mini::Face_index(0));
    CHECK(tree.closest_point(Point(0.2, 0.2, -1.0), Point(1, 1, 1)) == cpp.first);   // hinted overload
  }
  {   // aabb_test_empty_tree.cpp:38-53
    Tree tree;
    const Segment s(Point(0.1, 0.1, -1), Point(0.1, 0.1, 1));
    CHECK(tree.empty() && tree.size() == 0);
    CHECK(!tree.do_intersect(s) && tree.number_of_intersected_primitives(s) == 0);
    CHECK(!tree.any_intersected_primitive(s));
    CHECK(!tree.first_intersection(Ray(s.source(), s.target())));
    bool threw = false;
    try { tree.closest_point(Point(0, 0, 0)); } catch (const CGAL_B200::Error& e) { threw = e.code() == B200AABB_ERR_EMPTY_TREE; }
    CHECK(threw);   // CGAL_precondition(!empty())
  }
  {   // aabb_test_singleton_tree.cpp:44-67 + insert / clear / lazy rebuild / move
    Mesh one;
    one.add_face(one.add_vertex(Point(0, 0, 0)), one.add_vertex(Point(1, 0, 0)), one.add_vertex(Point(0, 1, 0)));
    Tree tree(faces(one).first, faces(one).second, one);
    const Segment s(Point(0.1, 0.1, -1), Point(0.1, 0.1, 1));
    CHECK(tree.size() == 1 && tree.do_intersect(s) && tree.number_of_intersected_primitives(s) == 1);
    CHECK(tree.closest_point(Point(0.25, 0.25, 5)) == Point(0.25, 0.25, 0) && tree.squared_distance(Point(0.25, 0.25, 5)) == 25.0);
    Mesh m = tetrahedron();
    tree.rebuild(faces(m).first, faces(m).second, m);
    CHECK(tree.size() == 4 && tree.number_of_intersected_primitives(Ray(Point(0.2, 0.2, -1), Point(0.2, 0.2, 0))) == 2);
    Tree moved(std::move(tree));
    CHECK(moved.size() == 4 && moved.do_intersect(s));
    moved.clear();
    CHECK(moved.empty() && !moved.do_intersect(s));
    moved.insert(faces(one).first, faces(one).second, one);   // no explicit build(): the first query builds
    CHECK(moved.number_of_intersected_primitives(s) == 1);
  }
  {   // batched == one at a time
    Mesh m = grid(24);
    Tree tree(faces(m).first, faces(m).second, m);
    tree.build();
    uint64_t seed = 23;
    std::vector<Ray> rays; std::vector<Segment> segs; std::vector<Point> pts;
    for (int i = 0; i < 600; ++i) {
      const Point a(24 * u01(seed), 24 * u01(seed), 3 * u01(seed) - 1.5), b(24 * u01(seed), 24 * u01(seed), 3 * u01(seed) - 1.5);
      rays.push_back(Ray(a, b)); segs.push_back(Segment(a, b)); pts.push_back(a);
    }
    std::vector<Tree::Ray_intersection> fi(rays.size());
    std::vector<uint8_t> flags;
    tree.first_intersection(rays.begin(), rays.end(), fi.begin(), &flags);
    std::vector<bool> hit; std::vector<Tree::size_type> cnt; std::vector<Tree::Point_and_primitive_id> cp; std::vector<double> sd;
    tree.do_intersect(segs.begin(), segs.end(), std::back_inserter(hit));
    tree.number_of_intersected_primitives(segs.begin(), segs.end(), std::back_inserter(cnt));
    tree.closest_point_and_primitive(pts.begin(), pts.end(), std::back_inserter(cp));
    tree.squared_distance(pts.begin(), pts.end(), std::back_inserter(sd));
    int nhit = 0;
    for (size_t i = 0; i < rays.size(); ++i) {
      const auto one = tree.first_intersection(rays[i]);
      CHECK(one.has_value() == fi[i].has_value());
      if (one) { CHECK(one->second == fi[i]->second && std::get<Point>(one->first) == std::get<Point>(fi[i]->first)); ++nhit; }
      CHECK(tree.do_intersect(segs[i]) == hit[i] && tree.number_of_intersected_primitives(segs[i]) == cnt[i] && (cnt[i] > 0) == hit[i]);
      const auto c1 = tree.closest_point_and_primitive(pts[i]);
      CHECK(c1.first == cp[i].first && c1.second == cp[i].second && tree.squared_distance(pts[i]) == sd[i]);
    }
    CHECK(nhit > 50 && flags.size() == rays.size());
    // first_intersection(ray, skip) (AABB_tree.h:307-328): skipping the first hit's face gives a later (or no) hit; skipping nothing changes nothing
    int nskip = 0;
    for (size_t i = 0; i < rays.size() && nskip < 40; ++i) {
      if (!fi[i]) continue;
      const auto first_id = fi[i]->second;
      const auto second = tree.first_intersection(rays[i], [&](const Tree::Primitive_id& id) { return id == first_id; });
      CHECK(!second || second->second != first_id);
      const auto same = tree.first_intersection(rays[i], [](const Tree::Primitive_id&) { return false; });
      CHECK(same && same->second == first_id);
      const auto prim = tree.first_intersected_primitive(rays[i], [&](const Tree::Primitive_id& id) { return id == first_id; });
      CHECK(prim.has_value() == second.has_value());
      ++nskip;
    }
    CHECK(nskip > 10);
    // PMP::max_distance_to_mesh's inner loop: the largest of the per-point distances
    double worst = 0;
    for (double d2 : sd) worst = std::max(worst, std::sqrt(d2));
    CHECK(tree.max_distance_to_mesh(pts.begin(), pts.end()) == worst);
    // const queries from several host threads at once (the reference's callers run tbb::parallel_for over them): serialised per handle, same answers
    std::vector<std::thread> th;
    std::vector<int> ok(4, 0);
    for (int k = 0; k < 4; ++k)
      th.emplace_back([&, k] {
        int good = 1;
        for (size_t i = k; i < pts.size(); i += 4) {
          const auto c1 = tree.closest_point_and_primitive(pts[i]);
          good &= (c1.first == cp[i].first && c1.second == cp[i].second) ? 1 : 0;
          good &= (tree.number_of_intersected_primitives(segs[i]) == cnt[i]) ? 1 : 0;
        }
        ok[k] = good;
      });
    for (auto& x : th) x.join();
    CHECK(ok[0] && ok[1] && ok[2] && ok[3]);
  }
  std::printf("test_aabb_tree: all known-answer and consistency checks passed\n");
  return 0;
}

// file format (little endian): u64 nv, nf, nrays, nsegs, npts; f64 verts[nv*3]; u32 faces[nf*3]; f64 rays[nrays*6]; f64 segs[nsegs*6]; f64 pts[npts*3]
// output: per ray u32 id (0xFFFFFFFF none) + f64 point[3]; per segment u32 count + u32 do_intersect; per point u32 id + f64 cp[3] + f64 sqdist
// other primitive types of the reference over the same class: several face graphs in one tree (ids carry the graph), and a range
// of Triangle_3 (ids are iterators)
static int other_primitives() {
  {   // AABB_face_graph_triangle_primitive<..., OneFaceGraphPerTree = Tag_false>: two meshes inserted into one tree
    typedef mini::AABB_face_graph_triangle_primitive_many<Mesh> Many;
    typedef CGAL_B200::AABB_tree<mini::AABB_traits_3<K, Many>> Tree2;
    Mesh a = tetrahedron(), b = tetrahedron(5.0, 0.0);
    Tree2 tree;
    tree.insert(faces(a).first, faces(a).second, a);
    tree.insert(faces(b).first, faces(b).second, b);
    CHECK(tree.size() == 8);
    const auto bb = tree.bbox();
    CHECK(bb.xmin() == 0 && bb.xmax() == 6);
    const Ray up_a(Point(0.2, 0.2, -1.0), Point(0.2, 0.2, 0.0)), up_b(Point(5.2, 0.2, -1.0), Point(5.2, 0.2, 0.0));
    const auto ha = tree.first_intersection(up_a), hb = tree.first_intersection(up_b);
    CHECK(ha && ha->second.second == &a && ha->second.first == mini::Face_index(0));
    CHECK(hb && hb->second.second == &b && hb->second.first == mini::Face_index(0));
    const auto cp = tree.closest_point_and_primitive(Point(5.0, 0.0, 3.0));
    CHECK(cp.first == Point(5.0, 0.0, 1.0) && cp.second.second == &b);
    CHECK(tree.number_of_intersected_primitives(Segment(Point(-1.0, 0.1, 0.1), Point(7.0, 0.1, 0.1))) == 4);   // through both solids
    std::vector<Many::Id> ids;
    tree.all_intersected_primitives(up_b, std::back_inserter(ids));
    CHECK(ids.size() == 2 && ids[0].second == &b && ids[1].second == &b);
  }
  {   // AABB_triangle_primitive_3 over a std::vector<Triangle_3>
    typedef std::vector<mini::Triangle_3>::const_iterator It;
    typedef mini::AABB_triangle_primitive_3<It> Prim;
    typedef CGAL_B200::AABB_tree<mini::AABB_traits_3<K, Prim>> Tree3;
    std::vector<mini::Triangle_3> soup;
    for (int i = 0; i < 10; ++i) soup.push_back(mini::Triangle_3(Point(0, 0, i), Point(1, 0, i), Point(0, 1, i)));   // ten parallel triangles
    Tree3 tree(soup.cbegin(), soup.cend());
    CHECK(tree.size() == 10);
    const Ray down(Point(0.2, 0.2, 20.0), Point(0.2, 0.2, 19.0));
    const auto h = tree.first_intersection(down);
    CHECK(h && h->second == soup.cbegin() + 9);
    CHECK(tree.number_of_intersected_primitives(down) == 10);
    CHECK(tree.number_of_intersected_primitives(Segment(Point(0.2, 0.2, 2.5), Point(0.2, 0.2, 6.5))) == 4);
    const auto cp = tree.closest_point_and_primitive(Point(0.2, 0.2, 3.4));
    CHECK(cp.first == Point(0.2, 0.2, 3.0) && cp.second == soup.cbegin() + 3);
    CHECK(tree.do_intersect(mini::Triangle_3(Point(0.1, 0.1, 4.5), Point(0.3, 0.1, 5.5), Point(0.1, 0.3, 5.5))));      // crosses z = 5
  }
  {   // AABB_segment_primitive_3: AABB_tree/examples/AABB_tree/AABB_segment_3_example.cpp:26-58
    typedef std::list<Segment>::const_iterator It;
    typedef mini::AABB_segment_primitive_3<It> Prim;
    typedef CGAL_B200::AABB_tree<mini::AABB_traits_3<K, Prim>> Tree4;
    const Point a(1.0, 0.0, 0.0), b(0.0, 1.0, 0.0), c(0.0, 0.0, 1.0), d(0.0, 0.0, 0.0);
    std::list<Segment> segments;
    segments.push_back(Segment(a, b)); segments.push_back(Segment(a, c)); segments.push_back(Segment(c, d));
    Tree4 tree(segments.cbegin(), segments.cend());
    CHECK(tree.size() == 3);
    CHECK(tree.number_of_intersected_primitives(mini::Plane_3(0.0, 0.0, 1.0, 0.0)) == 3);         // the plane through a, b, d
    CHECK(tree.number_of_intersected_primitives(mini::Triangle_3(a, b, c)) == 3);
    CHECK(tree.number_of_intersected_primitives(mini::Plane_3(0.0, 0.0, 1.0, -0.5)) == 2);        // z = 1/2 cuts (a, c) and (c, d)
    const Point closest = tree.closest_point(Point(2.0, 2.0, 2.0));
    CHECK(closest == Point(0.5, 0.5, 0.0) || closest == Point(0.5, 0.0, 0.5));
    CHECK(tree.squared_distance(Point(2.0, 2.0, 2.0)) == 8.5);
    const auto cp = tree.closest_point_and_primitive(Point(-1.0, -1.0, 0.25));
    CHECK(cp.first == Point(0.0, 0.0, 0.25) && cp.second == std::next(segments.cbegin(), 2));
    bool refused = false;
    try { tree.do_intersect(Segment(a, c)); } catch (const CGAL_B200::Error&) { refused = true; }   // Segment_3 x Segment_3 is not offered
    CHECK(refused);
  }
  return 0;
}

static int answer_file(const char* in, const char* out) {
  FILE* f = std::fopen(in, "rb");
  CHECK(f);
  uint64_t h[5];
  CHECK(std::fread(h, 8, 5, f) == 5);
  std::vector<double> v(3 * h[0]), rays(6 * h[2]), segs(6 * h[3]), pts(3 * h[4]);
  std::vector<uint32_t> fc(3 * h[1]);
  CHECK(std::fread(v.data(), 8, v.size(), f) == v.size() && std::fread(fc.data(), 4, fc.size(), f) == fc.size());
  CHECK(std::fread(rays.data(), 8, rays.size(), f) == rays.size() && std::fread(segs.data(), 8, segs.size(), f) == segs.size() && std::fread(pts.data(), 8, pts.size(), f) == pts.size());
  std::fclose(f);
  Mesh m;
  for (uint64_t i = 0; i < h[0]; ++i) m.add_vertex(Point(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
  for (uint64_t i = 0; i < h[1]; ++i) m.add_face(mini::Vertex_index(fc[3 * i]), mini::Vertex_index(fc[3 * i + 1]), mini::Vertex_index(fc[3 * i + 2]));
  Tree tree(faces(m).first, faces(m).second, m);
  tree.accelerate_distance_queries();
  std::vector<Ray> R; std::vector<Segment> S; std::vector<Point> P;
  for (uint64_t i = 0; i < h[2]; ++i) R.push_back(Ray(Point(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), Point(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5])));
  for (uint64_t i = 0; i < h[3]; ++i) S.push_back(Segment(Point(segs[6 * i], segs[6 * i + 1], segs[6 * i + 2]), Point(segs[6 * i + 3], segs[6 * i + 4], segs[6 * i + 5])));
  for (uint64_t i = 0; i < h[4]; ++i) P.push_back(Point(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
  std::vector<Tree::Ray_intersection> fi(R.size());
  tree.first_intersection(R.begin(), R.end(), fi.begin());
  std::vector<Tree::size_type> cnt; std::vector<bool> hit; std::vector<Tree::Point_and_primitive_id> cp; std::vector<double> sd;
  tree.number_of_intersected_primitives(S.begin(), S.end(), std::back_inserter(cnt));
  tree.do_intersect(S.begin(), S.end(), std::back_inserter(hit));
  tree.closest_point_and_primitive(P.begin(), P.end(), std::back_inserter(cp));
  tree.squared_distance(P.begin(), P.end(), std::back_inserter(sd));
  FILE* o = std::fopen(out, "wb");
  CHECK(o);
  for (size_t i = 0; i < R.size(); ++i) {
    uint32_t id = 0xFFFFFFFFu; double p[3] = {0, 0, 0};
    if (fi[i]) {
      id = (uint32_t)fi[i]->second;
      const Point x = std::holds_alternative<Point>(fi[i]->first) ? std::get<Point>(fi[i]->first) : std::get<Segment>(fi[i]->first).source();
      p[0] = x.x(); p[1] = x.y(); p[2] = x.z();
    }
    std::fwrite(&id, 4, 1, o); std::fwrite(p, 8, 3, o);
  }
  for (size_t i = 0; i < S.size(); ++i) { const uint32_t c[2] = {(uint32_t)cnt[i], hit[i] ? 1u : 0u}; std::fwrite(c, 4, 2, o); }
  for (size_t i = 0; i < P.size(); ++i) {
    const uint32_t id = (uint32_t)cp[i].second; const double p[4] = {cp[i].first.x(), cp[i].first.y(), cp[i].first.z(), sd[i]};
    std::fwrite(&id, 4, 1, o); std::fwrite(p, 8, 4, o);
  }
  std::fclose(o);
  return 0;
}

int main(int argc, char** argv) {
  try {
    if (argc == 3) return answer_file(argv[1], argv[2]);
    return known_answers() || other_primitives();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "exception: %s\n", e.what());
    return 2;
  }
}

// mini_cgal.h — TEST SCAFFOLDING: the smallest set of types with the interfaces of the CGAL classes that
// CGAL_B200::AABB_tree touches, so that the class template can be compiled and exercised in an image without Boost
// (CGAL's own headers need Boost from the first include).  Same names, members and conventions as
//   Epick kernel objects      Cartesian_kernel/include/CGAL/Cartesian/{Point_3,Segment_3,Ray_3,Triangle_3}.h
//   Bbox_3                    Kernel_23/include/CGAL/Bbox_3.h
//   Surface_mesh              Surface_mesh/include/CGAL/Surface_mesh/Surface_mesh.h (vertices, faces, halfedge(f), target, next, source)
//   AABB_face_graph_triangle_primitive  AABB_tree/include/CGAL/AABB_face_graph_triangle_primitive.h:53-196
//   AABB_traits_3             AABB_tree/include/CGAL/AABB_traits_3.h:507-560
// Nothing here computes anything: geometry is done by libb200aabb.so.
#pragma once
#include <array>
#include <cstdint>
#include <utility>
#include <variant>
#include <vector>

namespace mini {

struct Point_3 {
  double c[3] = {0, 0, 0};
  Point_3() {}
  Point_3(double x, double y, double z) : c{x, y, z} {}
  double x() const { return c[0]; }
  double y() const { return c[1]; }
  double z() const { return c[2]; }
  bool operator==(const Point_3& o) const { return c[0] == o.c[0] && c[1] == o.c[1] && c[2] == o.c[2]; }
  bool operator!=(const Point_3& o) const { return !(*this == o); }
};
struct Segment_3 {
  Point_3 s, t;
  Segment_3() {}
  Segment_3(const Point_3& a, const Point_3& b) : s(a), t(b) {}
  const Point_3& source() const { return s; }
  const Point_3& target() const { return t; }
};
struct Ray_3 {   // stored as (source, second point): Cartesian/Ray_3.h:35-47
  Point_3 s, q;
  Ray_3() {}
  Ray_3(const Point_3& a, const Point_3& b) : s(a), q(b) {}
  const Point_3& source() const { return s; }
  const Point_3& second_point() const { return q; }
};
struct Vector_3 {
  double cx = 0, cy = 0, cz = 0;
  Vector_3() {}
  Vector_3(double x, double y, double z) : cx(x), cy(y), cz(z) {}
  double x() const { return cx; }
  double y() const { return cy; }
  double z() const { return cz; }
};
struct Line_3 {   // stored as (point, direction vector); Line_3(p, q) keeps the rounded difference q - p: Cartesian/Line_3.h
  Point_3 p; Vector_3 v;
  Line_3() {}
  Line_3(const Point_3& a, const Vector_3& d) : p(a), v(d) {}
  Line_3(const Point_3& a, const Point_3& b) : p(a), v(b.x() - a.x(), b.y() - a.y(), b.z() - a.z()) {}
  const Point_3& point() const { return p; }
  const Vector_3& to_vector() const { return v; }
};
struct Plane_3 {  // a x + b y + c z + d = 0: Cartesian/Plane_3.h
  double ca, cb, cc, cd;
  Plane_3() : ca(0), cb(0), cc(1), cd(0) {}
  Plane_3(double a, double b, double c, double d) : ca(a), cb(b), cc(c), cd(d) {}
  double a() const { return ca; }
  double b() const { return cb; }
  double c() const { return cc; }
  double d() const { return cd; }
};
struct Aff_transformation_3 {  // the 3 x 4 matrix of Cartesian/Aff_transformation_rep_3.h (general representation)
  double t[3][4];
  Aff_transformation_3() : t{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}} {}
  Aff_transformation_3(double m00, double m01, double m02, double m03, double m10, double m11, double m12, double m13, double m20, double m21, double m22, double m23)
      : t{{m00, m01, m02, m03}, {m10, m11, m12, m13}, {m20, m21, m22, m23}} {}
  double m(int i, int j) const { return t[i][j]; }
};
struct Triangle_3 {
  Point_3 v[3];
  Triangle_3() {}
  Triangle_3(const Point_3& a, const Point_3& b, const Point_3& c) : v{a, b, c} {}
  const Point_3& vertex(int i) const { return v[i % 3]; }
};
struct Bbox_3 {
  double b[6];
  Bbox_3(double x0, double y0, double z0, double x1, double y1, double z1) : b{x0, y0, z0, x1, y1, z1} {}
  double xmin() const { return b[0]; }
  double ymin() const { return b[1]; }
  double zmin() const { return b[2]; }
  double xmax() const { return b[3]; }
  double ymax() const { return b[4]; }
  double zmax() const { return b[5]; }
};

struct Epick {
  typedef double FT;
  typedef mini::Point_3 Point_3;
  typedef mini::Segment_3 Segment_3;
  typedef mini::Ray_3 Ray_3;
  typedef mini::Line_3 Line_3;
  typedef mini::Plane_3 Plane_3;
  typedef mini::Vector_3 Vector_3;
  typedef mini::Triangle_3 Triangle_3;
};

// index types of Surface_mesh (Surface_mesh.h:63-198)
template <int Tag>
struct SM_Index {
  uint32_t i = 0xFFFFFFFFu;
  SM_Index() {}
  explicit SM_Index(uint32_t v) : i(v) {}
  operator uint32_t() const { return i; }
  bool operator==(const SM_Index& o) const { return i == o.i; }
  bool operator!=(const SM_Index& o) const { return i != o.i; }
};
typedef SM_Index<0> Vertex_index;
typedef SM_Index<1> Halfedge_index;
typedef SM_Index<2> Face_index;

// triangle mesh with the halfedge navigation the primitive needs.  Halfedge 3f+k of face f points to vertex k of the
// face; halfedge(f) is the one pointing to vertex 0, as Euler::add_face leaves it (Euler_operations.h:842-843).
template <class P>
class Surface_mesh {
 public:
  typedef P Point;
  typedef mini::Vertex_index Vertex_index;
  typedef mini::Face_index Face_index;
  typedef mini::Halfedge_index Halfedge_index;
  struct Face_iterator {
    uint32_t i;
    Face_index operator*() const { return Face_index(i); }
    Face_iterator& operator++() { ++i; return *this; }
    bool operator!=(const Face_iterator& o) const { return i != o.i; }
    bool operator==(const Face_iterator& o) const { return i == o.i; }
  };
  Vertex_index add_vertex(const P& p) { m_points.push_back(p); return Vertex_index((uint32_t)m_points.size() - 1); }
  Face_index add_face(Vertex_index a, Vertex_index b, Vertex_index c) { m_faces.push_back({a, b, c}); return Face_index((uint32_t)m_faces.size() - 1); }
  const P& point(Vertex_index v) const { return m_points[v]; }
  Halfedge_index halfedge(Face_index f) const { return Halfedge_index(3 * (uint32_t)f); }
  Halfedge_index next(Halfedge_index h) const { return Halfedge_index(h / 3 * 3 + (h % 3 + 1) % 3); }
  Vertex_index target(Halfedge_index h) const { return m_faces[h / 3][h % 3]; }
  Vertex_index source(Halfedge_index h) const { return m_faces[h / 3][(h % 3 + 2) % 3]; }
  uint32_t number_of_faces() const { return (uint32_t)m_faces.size(); }
  uint32_t number_of_vertices() const { return (uint32_t)m_points.size(); }
  std::pair<Face_iterator, Face_iterator> face_range() const { return {Face_iterator{0}, Face_iterator{number_of_faces()}}; }

 private:
  std::vector<P> m_points;
  std::vector<std::array<Vertex_index, 3>> m_faces;
};
template <class P>
std::pair<typename Surface_mesh<P>::Face_iterator, typename Surface_mesh<P>::Face_iterator> faces(const Surface_mesh<P>& m) { return m.face_range(); }

// AABB_face_graph_triangle_primitive<FaceGraph, Default, Tag_true /*one graph per tree*/, Tag_false /*no datum cache*/>
template <class FaceGraph>
class AABB_face_graph_triangle_primitive {
 public:
  typedef typename FaceGraph::Face_index Id;
  typedef Triangle_3 Datum;
  typedef typename FaceGraph::Point Point;
  typedef const FaceGraph* Shared_data;
  template <class Iterator>
  AABB_face_graph_triangle_primitive(Iterator it, const FaceGraph&) : m_id(*it) {}
  Id id() const { return m_id; }
  // Triangle_from_face_descriptor_map, BGL/include/CGAL/boost/graph/property_maps.h:52-63
  Datum datum(Shared_data g) const {
    const auto h = g->halfedge(m_id);
    return Datum(g->point(g->target(h)), g->point(g->target(g->next(h))), g->point(g->source(h)));
  }
  Point reference_point(Shared_data g) const { return g->point(g->target(g->halfedge(m_id))); }
  static Shared_data construct_shared_data(const FaceGraph& g) { return &g; }

 private:
  Id m_id;
};

// AABB_face_graph_triangle_primitive<FaceGraph, Default, Tag_false /*SEVERAL graphs per tree*/, Tag_false>: the id carries the graph
// (AABB_tree/include/CGAL/AABB_face_graph_triangle_primitive.h:73-88: Id = std::pair<face_descriptor, const FaceGraph*>), no shared data
template <class FaceGraph>
class AABB_face_graph_triangle_primitive_many {
 public:
  typedef std::pair<typename FaceGraph::Face_index, const FaceGraph*> Id;
  typedef Triangle_3 Datum;
  typedef typename FaceGraph::Point Point;
  template <class Iterator>
  AABB_face_graph_triangle_primitive_many(Iterator it, const FaceGraph& g) : m_id(*it, &g) {}
  Id id() const { return m_id; }
  Datum datum() const {
    const FaceGraph* g = m_id.second;
    const auto h = g->halfedge(m_id.first);
    return Datum(g->point(g->target(h)), g->point(g->target(g->next(h))), g->point(g->source(h)));
  }
  Point reference_point() const { return m_id.second->point(m_id.second->target(m_id.second->halfedge(m_id.first))); }

 private:
  Id m_id;
};

// AABB_triangle_primitive_3<GeomTraits, Iterator> over a range of Triangle_3 (AABB_tree/include/CGAL/AABB_triangle_primitive_3.h:27-95): Id = the iterator
template <class Iterator>
class AABB_triangle_primitive_3 {
 public:
  typedef Iterator Id;
  typedef Triangle_3 Datum;
  typedef Point_3 Point;
  explicit AABB_triangle_primitive_3(Iterator it) : m_it(it) {}
  Id id() const { return m_it; }
  const Datum& datum() const { return *m_it; }
  Point reference_point() const { return m_it->vertex(0); }

 private:
  Iterator m_it;
};

// AABB_segment_primitive_3<GeomTraits, Iterator> over a range of Segment_3 (AABB_tree/include/CGAL/AABB_segment_primitive_3.h:27-95): Id = the iterator,
// reference point = the source
template <class Iterator>
class AABB_segment_primitive_3 {
 public:
  typedef Iterator Id;
  typedef Segment_3 Datum;
  typedef Point_3 Point;
  explicit AABB_segment_primitive_3(Iterator it) : m_it(it) {}
  Id id() const { return m_it; }
  const Datum& datum() const { return *m_it; }
  Point reference_point() const { return m_it->source(); }

 private:
  Iterator m_it;
};

// AABB_traits_base (AABB_traits_3.h:46-76): the shared data exists only for primitives that declare it
template <class Primitive, class = void>
struct AABB_traits_base {};
template <class Primitive>
struct AABB_traits_base<Primitive, std::void_t<typename Primitive::Shared_data>> {
  template <typename... T>
  void set_shared_data(T&&... t) { m_data = Primitive::construct_shared_data(std::forward<T>(t)...); }
  const typename Primitive::Shared_data& shared_data() const { return m_data; }

 private:
  typename Primitive::Shared_data m_data = nullptr;
};

template <class GeomTraits, class AABBPrimitive>
class AABB_traits_3 : public AABB_traits_base<AABBPrimitive> {
 public:
  typedef GeomTraits Geom_traits;
  typedef typename GeomTraits::FT FT;
  typedef AABBPrimitive Primitive;
  typedef typename GeomTraits::Point_3 Point;
  typedef typename GeomTraits::Point_3 Point_3;
  typedef Bbox_3 Bounding_box;
  typedef std::pair<Point, typename Primitive::Id> Point_and_primitive_id;
  template <typename Query>
  struct Intersection_and_primitive_id {
    typedef std::pair<std::variant<typename GeomTraits::Point_3, typename GeomTraits::Segment_3>, typename Primitive::Id> Type;
  };
};

}  // namespace mini

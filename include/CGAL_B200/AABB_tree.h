// CGAL_B200/AABB_tree.h — header-only C++17 class template with the interface of CGAL::AABB_tree
// (AABB_tree/include/CGAL/AABB_tree.h:55-677 of the reference) whose work is done by libb200aabb.so (sm_100a CUDA)
// through the C ABI of b200aabb.h.  This is the reference-side binding of INTEGRATION.md: a user of
//     CGAL::AABB_tree< CGAL::AABB_traits_3< Epick, CGAL::AABB_face_graph_triangle_primitive< Surface_mesh<Point_3> > > >
// replaces `CGAL::AABB_tree` by `CGAL_B200::AABB_tree` and keeps the traits, the primitive and the mesh types.
//
// What the class needs from its template argument (all provided by CGAL::AABB_traits_3, AABB_traits_3.h:507-560):
//   AABBTraits::FT, ::Point, ::Primitive (with ::Id, ::Datum, id(), datum([shared_data])), ::Bounding_box (constructible from
//   six doubles), ::Point_and_primitive_id, ::Geom_traits (with ::Segment_3, ::Ray_3), ::Intersection_and_primitive_id<Q>::Type,
//   set_shared_data(...) / shared_data() when the primitive has Shared_data (AABBPrimitiveWithSharedData).
// Geometry objects are read through the kernel-object interface (Point_3::x()/y()/z(), Segment_3::source()/target(),
// Ray_3::source()/second_point(), Triangle_3::vertex(i)) and results are built with Point(x,y,z) / Segment_3(p,q).
// The number type must be convertible to/from double without loss (Epick, Simple_cartesian<double>).
//
// Primitives are flattened exactly as the reference sees them: the triangle of primitive p is p.datum(shared_data), i.e. for
// AABB_face_graph_triangle_primitive (P[target(h)], P[target(next(h))], P[source(h)]), h = halfedge(f)
// (BGL/include/CGAL/boost/graph/property_maps.h:52-63).  The device id of a primitive is its rank in insertion order; the
// class maps it back to Primitive::id().
//
// Besides the reference's one-query-at-a-time members (each is one small GPU batch: correct, but latency-bound) the class
// has BATCHED overloads taking an iterator range of queries and an output iterator — the way to feed a GPU.
// There is no CPU fallback: without the library or without an sm_100 device every call throws CGAL_B200::Error.
#ifndef CGAL_B200_AABB_TREE_H
#define CGAL_B200_AABB_TREE_H

#include <cstdint>
#include <iterator>
#include <atomic>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include <memory>
#include "../b200aabb.h"

namespace CGAL_B200 {

// thrown for every non-zero status of the C ABI; code() is the B200AABB_ERR_* value
class Error : public std::runtime_error {
 public:
  Error(int code, const std::string& what) : std::runtime_error("b200aabb (" + std::to_string(code) + "): " + what), m_code(code) {}
  int code() const { return m_code; }

 private:
  int m_code;
};

// One NCCL communicator per GPU, for replicating a built tree (b200aabb_comm_*: the reference has no multi-device path, its callers shard queries over
// threads; here a tree is built on one GPU, replicated, and every GPU answers its own block of the batch against its replica).
//   several processes, one GPU each:  char id[128]; if (rank == 0) Communicator::unique_id(id);  /* ship id to the other ranks */
//                                     Communicator comm(device, rank, world, id);   tree.replicate(comm, 0);
//   one process, several GPUs:        auto comms = Communicator::all({0, 1, 2, 3});  AABB_tree<Traits>::replicate_all(trees, comms, 0);
class Communicator {
 public:
  static void unique_id(char id[128]) { check(b200aabb_comm_unique_id(id), nullptr); }
  Communicator(int device, int rank, int world, const char id[128]) { check(b200aabb_comm_init(device, rank, world, id, &m_comm), nullptr); }
  ~Communicator() { if (m_comm) b200aabb_comm_destroy(m_comm); }
  Communicator(const Communicator&) = delete;
  Communicator& operator=(const Communicator&) = delete;
  static std::vector<std::unique_ptr<Communicator>> all(const std::vector<int>& devices) {
    std::vector<b200aabb_comm*> raw(devices.size(), nullptr);
    check(b200aabb_comm_init_all((int)devices.size(), devices.data(), raw.data()), nullptr);
    std::vector<std::unique_ptr<Communicator>> out;
    for (b200aabb_comm* c : raw) out.emplace_back(new Communicator(c));
    return out;
  }
  b200aabb_comm* handle() const { return m_comm; }
  static void check(int rc, const b200aabb_comm* c) {
    if (rc != B200AABB_OK) throw Error(rc, c ? b200aabb_comm_last_error(c) : "b200aabb_comm: NCCL not available or the call failed");
  }

 private:
  explicit Communicator(b200aabb_comm* c) : m_comm(c) {}
  b200aabb_comm* m_comm = nullptr;
};

namespace internal {
template <class P, class = void>
struct Has_shared_data : std::false_type {};
template <class P>
struct Has_shared_data<P, std::void_t<typename P::Shared_data>> : std::true_type {};
}  // namespace internal

template <typename AABBTraits>
class AABB_tree {
  typedef AABB_tree<AABBTraits> Self;
  typedef std::vector<typename AABBTraits::Primitive> Primitives;

 public:
  // ---- types, AABB_tree.h:69-104
  typedef AABBTraits AABB_traits;
  typedef typename AABBTraits::FT FT;
  typedef typename AABBTraits::Point Point;
  typedef typename AABBTraits::Primitive Primitive;
  typedef typename Primitive::Id Primitive_id;
  typedef typename Primitives::size_type size_type;
  typedef typename AABBTraits::Bounding_box Bounding_box;
  typedef typename AABBTraits::Point_and_primitive_id Point_and_primitive_id;
  typedef typename AABBTraits::Geom_traits::Segment_3 Segment;
  typedef typename AABBTraits::Geom_traits::Ray_3 Ray;
  template <typename Query>
  struct Intersection_and_primitive_id {
    typedef typename AABBTraits::template Intersection_and_primitive_id<Query>::Type Type;
  };
  typedef std::optional<typename Intersection_and_primitive_id<Ray>::Type> Ray_intersection;   // result of first_intersection

  // ---- creation, AABB_tree.h:115-135.  `device` = CUDA ordinal of the B200 that holds the tree.
  explicit AABB_tree(const AABBTraits& traits = AABBTraits(), int device = 0) : m_traits(traits) { create(device); }
  template <typename InputIterator, typename... T>
  AABB_tree(InputIterator first, InputIterator beyond, T&&... t) {
    create(0);
    insert(first, beyond, std::forward<T>(t)...);
  }
  AABB_tree(Self&& o) noexcept : m_traits(std::move(o.m_traits)), m_primitives(std::move(o.m_primitives)), m_handle(o.m_handle), m_uploaded(o.m_uploaded.load()) {
    o.m_handle = nullptr;
  }
  Self& operator=(Self&& o) noexcept {
    if (this != &o) {
      destroy();
      m_traits = std::move(o.m_traits); m_primitives = std::move(o.m_primitives); m_handle = o.m_handle; m_uploaded = o.m_uploaded.load();
      o.m_handle = nullptr;
    }
    return *this;
  }
  AABB_tree(const Self&) = delete;
  Self& operator=(const Self&) = delete;
  ~AABB_tree() { destroy(); }

  // ---- primitives and build, AABB_tree.h:147-216,730-853
  template <typename InputIterator, typename... T>
  void insert(InputIterator first, InputIterator beyond, T&&... t) {
    set_shared_data(std::forward<T>(t)...);
    for (; first != beyond; ++first) m_primitives.push_back(Primitive(first, std::forward<T>(t)...));
    m_uploaded = false;
  }
  void insert(const Primitive& p) { m_primitives.push_back(p); m_uploaded = false; }
  template <typename InputIterator, typename... T>
  void rebuild(InputIterator first, InputIterator beyond, T&&... t) {
    clear();
    insert(first, beyond, std::forward<T>(t)...);
    build();
  }
  template <typename... T>
  void build(T&&... t) {
    set_shared_data(std::forward<T>(t)...);
    build();
  }
  // LBVH build on the device (replaces the recursive median split of AABB_tree.h:787-853)
  void build() {
    std::lock_guard<std::mutex> lock(m_mutex);
    m_uploaded = false;
    upload_and_build();
  }
  // the reference lets the caller replace Compute_bbox / Split_primitives; the device builder has a fixed LBVH split
  template <class ComputeBbox, class SplitPrimitives>
  void custom_build(const ComputeBbox&, const SplitPrimitives&) { throw Error(B200AABB_ERR_INVALID_ARGUMENT, "custom_build is not supported by the device builder"); }
  void clear() {
    std::lock_guard<std::mutex> lock(m_mutex);
    m_primitives.clear();
    check(b200aabb_clear(m_handle));
    m_uploaded = true;   // an empty tree is consistent
  }
  const AABBTraits& traits() const { return m_traits; }
  const Bounding_box bbox() const {
    ensure_built();
    double b[6];
    check(b200aabb_bbox(m_handle, b));
    return Bounding_box(b[0], b[1], b[2], b[3], b[4], b[5]);
  }
  size_type size() const { return m_primitives.size(); }
  bool empty() const { return m_primitives.empty(); }

  // accelerate_distance_queries, AABB_tree.h:442-459,896-901.  The kd-tree of reference points is replaced by the ordered
  // BVH descent (nearest child first), so there is nothing extra to build; the call builds the tree and returns true.
  bool accelerate_distance_queries() {
    ensure_built();
    check(b200aabb_accelerate_distance_queries(m_handle));
    return true;
  }
  template <typename ConstPointIterator>
  bool accelerate_distance_queries(ConstPointIterator, ConstPointIterator) { return accelerate_distance_queries(); }
  void do_not_accelerate_distance_queries() {}

  // ---- intersection queries, one query per call (AABB_tree.h:243-354,903-932).  Query = Segment or Ray.
  template <typename Query>
  bool do_intersect(const Query& query) const {
    bool r = false;
    do_intersect(&query, &query + 1, &r);
    return r;
  }
  // do_intersect(other_tree): the AABB_tree-as-query overloads of AABB_traits_3::Do_intersect (AABB_traits_3.h:679-690), one kernel on the device that
  // holds both trees; true iff a primitive of this tree intersects a primitive of `other`
  template <typename OtherTraits>
  bool do_intersect(const AABB_tree<OtherTraits>& other) const {
    ensure_built();
    int r = 0;
    check(b200aabb_do_intersect_tree(m_handle, other.handle(), &r, nullptr, nullptr));
    return r != 0;
  }
  // the same with both trees under affine transformations: Rigid_triangle_mesh_collision_detection::does_A_intersect_B
  // (Rigid_triangle_mesh_collision_detection.h:149-161) — trees built once on the untransformed meshes, Aff = CGAL::Aff_transformation_3<K> (m(i, j))
  template <typename OtherTraits, typename Aff>
  bool do_intersect(const AABB_tree<OtherTraits>& other, const Aff& transformation, const Aff& other_transformation) const {
    ensure_built();
    double a[12], b[12];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j) { a[4 * i + j] = (double)transformation.m(i, j); b[4 * i + j] = (double)other_transformation.m(i, j); }
    int r = 0;
    check(b200aabb_do_intersect_tree_transformed(m_handle, a, other.handle(), b, &r, nullptr, nullptr));
    return r != 0;
  }
  template <typename Query>
  size_type number_of_intersected_primitives(const Query& query) const {
    size_type r = 0;
    number_of_intersected_primitives(&query, &query + 1, &r);
    return r;
  }
  // all_intersected_primitives(query, out) (AABB_tree.h:253-257,934-945): out receives Primitive_id, in traversal order
  template <typename Query, typename OutputIterator>
  OutputIterator all_intersected_primitives(const Query& query, OutputIterator out) const {
    std::vector<std::vector<Primitive_id>> r;
    all_intersected_primitives(&query, &query + 1, std::back_inserter(r));
    for (const Primitive_id& id : r[0]) *out++ = id;
    return out;
  }
  // any_intersected_primitive(query) (AABB_tree.h:265-268,947-956)
  template <typename Query>
  std::optional<Primitive_id> any_intersected_primitive(const Query& query) const {
    std::optional<Primitive_id> r;
    any_intersected_primitive(&query, &query + 1, &r);
    return r;
  }
  // all_intersections(query, out) / any_intersection(query) for Ray_3 and Segment_3 queries (AABB_tree.h:283-296,958-985): out receives
  // Intersection_and_primitive_id<Query>::Type
  template <typename OutputIterator>
  OutputIterator all_intersections(const Ray& query, OutputIterator out) const {
    std::vector<std::vector<typename Intersection_and_primitive_id<Ray>::Type>> r;
    all_intersections(&query, &query + 1, std::back_inserter(r));
    for (const auto& x : r[0]) *out++ = x;
    return out;
  }
  template <typename OutputIterator>
  OutputIterator all_intersections(const Segment& query, OutputIterator out) const {
    std::vector<std::vector<typename Intersection_and_primitive_id<Segment>::Type>> r;
    all_intersections(&query, &query + 1, std::back_inserter(r));
    for (const auto& x : r[0]) *out++ = x;
    return out;
  }
  template <typename Query, typename OutputIterator, typename = std::enable_if_t<!std::is_same<Query, Ray>::value && !std::is_same<Query, Segment>::value>>
  OutputIterator all_intersections(const Query& query, OutputIterator out) const {            // Line_3 and Plane_3 queries
    std::vector<std::vector<typename Intersection_and_primitive_id<Query>::Type>> r;
    all_intersections(&query, &query + 1, std::back_inserter(r));
    for (const auto& x : r[0]) *out++ = x;
    return out;
  }
  Ray_intersection any_intersection(const Ray& query) const {
    Ray_intersection r;
    any_intersection(&query, &query + 1, &r);
    return r;
  }
  std::optional<typename Intersection_and_primitive_id<Segment>::Type> any_intersection(const Segment& query) const {
    std::optional<typename Intersection_and_primitive_id<Segment>::Type> r;
    any_intersection(&query, &query + 1, &r);
    return r;
  }
  Ray_intersection first_intersection(const Ray& query) const {
    Ray_intersection r;
    first_intersection(&query, &query + 1, &r);
    return r;
  }
  std::optional<Primitive_id> first_intersected_primitive(const Ray& query) const {
    std::optional<Primitive_id> r;
    first_intersected_primitive(&query, &query + 1, &r);
    return r;
  }
  // first_intersection(query, skip) / first_intersected_primitive(query, skip), AABB_tree.h:307-354.  A host functor cannot run inside the
  // traversal kernel, so it is evaluated ONCE PER PRIMITIVE on the host and travels as a bit mask (b200aabb_first_intersection_skip):
  // O(size()) functor calls per call of this overload — batch the rays (first_intersection_skip below) where that matters.
  template <typename SkipFunctor>
  Ray_intersection first_intersection(const Ray& query, const SkipFunctor& skip) const {
    Ray_intersection r;
    first_intersection_skip(&query, &query + 1, &r, skip);
    return r;
  }
  template <typename SkipFunctor>
  std::optional<Primitive_id> first_intersected_primitive(const Ray& query, const SkipFunctor& skip) const {
    const Ray_intersection r = first_intersection(query, skip);
    if (r) return r->second;
    return std::nullopt;
  }

  // ---- distance queries, one query per call (AABB_tree.h:363-377,464-477,987-1055)
  FT squared_distance(const Point& query) const {
    FT r;
    squared_distance(&query, &query + 1, &r);
    return r;
  }
  Point closest_point(const Point& query) const { return closest_point_and_primitive(query).first; }
  Point_and_primitive_id closest_point_and_primitive(const Point& query) const {
    std::vector<Point_and_primitive_id> r;
    closest_point_and_primitive(&query, &query + 1, std::back_inserter(r));
    return r[0];
  }
  // A hint only seeds the reference's search ball; the result is the same except on exact-distance ties
  FT squared_distance(const Point& query, const Point&) const { return squared_distance(query); }
  Point closest_point(const Point& query, const Point&) const { return closest_point(query); }
  Point_and_primitive_id closest_point_and_primitive(const Point& query, const Point_and_primitive_id&) const { return closest_point_and_primitive(query); }

  // any_reference_point_and_id() (AABB_tree.h:582-588): the reference point of the first primitive; best_hint(query) (:591-612): the reference point nearest to the
  // query — what the reference's internal kd-tree returns.  Neither is needed by the queries above (the device traversal finds its own first ball); they are kept
  // for callers of the public members and run on the host, best_hint in O(size()).
  Point_and_primitive_id any_reference_point_and_id() const {
    if (m_primitives.empty()) throw Error(B200AABB_ERR_EMPTY_TREE, "any_reference_point_and_id on an empty tree (CGAL_assertion(!empty()))");
    return Point_and_primitive_id(reference_point(m_primitives[0]), m_primitives[0].id());
  }
  Point_and_primitive_id best_hint(const Point& query) const {
    if (m_primitives.empty()) throw Error(B200AABB_ERR_EMPTY_TREE, "best_hint on an empty tree");
    size_type arg = 0;
    double best = -1.0;
    for (size_type i = 0; i < m_primitives.size(); ++i) {
      decltype(auto) r = reference_point(m_primitives[i]);
      const double dx = (double)r.x() - (double)query.x(), dy = (double)r.y() - (double)query.y(), dz = (double)r.z() - (double)query.z();
      const double d = dx * dx + dy * dy + dz * dz;
      if (best < 0.0 || d < best) { best = d; arg = i; }
    }
    return Point_and_primitive_id(reference_point(m_primitives[arg]), m_primitives[arg].id());
  }

  // ---- BATCHED overloads (new): [first, beyond) is a range of queries, results go to `out` in the same order ------------
  // out receives bool
  template <typename QueryIterator, typename OutputIterator>
  OutputIterator do_intersect(QueryIterator first, QueryIterator beyond, OutputIterator out) const {
    typedef typename std::iterator_traits<QueryIterator>::value_type Query;
    std::vector<double> q = pack_lines(first, beyond);
    const uint64_t n = q.size() / doubles_per_query<Query>();
    std::vector<uint8_t> r(n);
    ensure_built();
    check(b200aabb_do_intersect(m_handle, q.data(), n, kind_of<Query>(), B200AABB_HOST, r.data()));
    for (uint64_t i = 0; i < n; ++i) *out++ = r[i] != 0;
    return out;
  }
  // out receives size_type
  template <typename QueryIterator, typename OutputIterator>
  OutputIterator number_of_intersected_primitives(QueryIterator first, QueryIterator beyond, OutputIterator out) const {
    typedef typename std::iterator_traits<QueryIterator>::value_type Query;
    std::vector<double> q = pack_lines(first, beyond);
    const uint64_t n = q.size() / doubles_per_query<Query>();
    std::vector<uint32_t> r(n);
    ensure_built();
    check(b200aabb_number_of_intersected_primitives(m_handle, q.data(), n, kind_of<Query>(), B200AABB_HOST, r.data()));
    for (uint64_t i = 0; i < n; ++i) *out++ = (size_type)r[i];
    return out;
  }
  // out receives std::optional<Primitive_id>
  template <typename QueryIterator, typename OutputIterator>
  OutputIterator any_intersected_primitive(QueryIterator first, QueryIterator beyond, OutputIterator out) const {
    typedef typename std::iterator_traits<QueryIterator>::value_type Query;
    std::vector<double> q = pack_lines(first, beyond);
    const uint64_t n = q.size() / doubles_per_query<Query>();
    std::vector<uint32_t> r(n);
    ensure_built();
    check(b200aabb_any_intersected_primitive(m_handle, q.data(), n, kind_of<Query>(), B200AABB_HOST, r.data()));
    for (uint64_t i = 0; i < n; ++i) {
      if (r[i] == B200AABB_NO_PRIMITIVE) *out++ = std::nullopt;
      else *out++ = std::optional<Primitive_id>(m_primitives[r[i]].id());
    }
    return out;
  }
  // out receives one std::vector<Primitive_id> per query (count pass -> prefix sums -> fill pass, include/b200aabb.h)
  template <typename QueryIterator, typename OutputIterator>
  OutputIterator all_intersected_primitives(QueryIterator first, QueryIterator beyond, OutputIterator out) const {
    typedef typename std::iterator_traits<QueryIterator>::value_type Query;
    std::vector<double> q = pack_lines(first, beyond);
    const uint64_t n = q.size() / doubles_per_query<Query>();
    std::vector<uint32_t> cnt(n);
    ensure_built();
    check(b200aabb_number_of_intersected_primitives(m_handle, q.data(), n, kind_of<Query>(), B200AABB_HOST, cnt.data()));
    std::vector<uint64_t> off(n + 1, 0);
    for (uint64_t i = 0; i < n; ++i) off[i + 1] = off[i] + cnt[i];
    std::vector<uint32_t> ids(off[n] ? off[n] : 1);
    check(b200aabb_all_intersected_primitives(m_handle, q.data(), n, kind_of<Query>(), B200AABB_HOST, off.data(), ids.data()));
    for (uint64_t i = 0; i < n; ++i) {
      std::vector<Primitive_id> v;
      for (uint64_t k = off[i]; k < off[i + 1]; ++k) v.push_back(m_primitives[ids[k]].id());
      *out++ = v;
    }
    return out;
  }
  // out receives std::optional< Intersection_and_primitive_id<Query>::Type >, Query = Ray_3 or Segment_3 (the range's value type)
  template <typename RayIterator, typename OutputIterator>
  OutputIterator any_intersection(RayIterator first, RayIterator beyond, OutputIterator out) const {
    typedef typename std::iterator_traits<RayIterator>::value_type Query;     // Ray_3, Segment_3, Line_3 or Plane_3 (kind_of)
    typedef typename Intersection_and_primitive_id<Query>::Type Result;
    typedef typename Result::first_type Intersection;
    std::vector<double> q = pack_lines(first, beyond);
    const uint64_t n = q.size() / doubles_per_query<Query>();
    std::vector<uint32_t> prim(n);
    std::vector<double> p0(3 * n), p1(3 * n);
    std::vector<uint8_t> fl(n);
    ensure_built();
    check(b200aabb_any_intersection_q(m_handle, q.data(), n, kind_of<Query>(), B200AABB_HOST, prim.data(), p0.data(), p1.data(), fl.data()));
    for (uint64_t i = 0; i < n; ++i) {
      if (!(fl[i] & B200AABB_FLAG_HIT)) { *out++ = std::nullopt; continue; }
      const Point a(FT(p0[3 * i]), FT(p0[3 * i + 1]), FT(p0[3 * i + 2]));
      if (fl[i] & B200AABB_FLAG_TRIANGLE) *out++ = Result(triangle_object<Intersection>(prim[i]), m_primitives[prim[i]].id());
      else if (fl[i] & B200AABB_FLAG_SEGMENT) *out++ = Result(Intersection(Segment(a, Point(FT(p1[3 * i]), FT(p1[3 * i + 1]), FT(p1[3 * i + 2])))), m_primitives[prim[i]].id());
      else *out++ = Result(Intersection(a), m_primitives[prim[i]].id());
    }
    return out;
  }
  // out receives one std::vector< Intersection_and_primitive_id<Query>::Type > per query (count pass -> prefix sums -> fill pass)
  template <typename RayIterator, typename OutputIterator>
  OutputIterator all_intersections(RayIterator first, RayIterator beyond, OutputIterator out) const {
    typedef typename std::iterator_traits<RayIterator>::value_type Query;     // Ray_3, Segment_3, Line_3 or Plane_3 (kind_of)
    typedef typename Intersection_and_primitive_id<Query>::Type Result;
    typedef typename Result::first_type Intersection;
    std::vector<double> q = pack_lines(first, beyond);
    const uint64_t n = q.size() / doubles_per_query<Query>();
    std::vector<uint32_t> cnt(n);
    ensure_built();
    check(b200aabb_number_of_intersections_q(m_handle, q.data(), n, kind_of<Query>(), B200AABB_HOST, cnt.data()));
    std::vector<uint64_t> off(n + 1, 0);
    for (uint64_t i = 0; i < n; ++i) off[i + 1] = off[i] + cnt[i];
    const uint64_t room = off[n] ? off[n] : 1;
    std::vector<uint32_t> ids(room);
    std::vector<uint8_t> kinds(room);
    std::vector<double> pts(6 * room);
    check(b200aabb_all_intersections_q(m_handle, q.data(), n, kind_of<Query>(), B200AABB_HOST, off.data(), ids.data(), kinds.data(), pts.data()));
    for (uint64_t i = 0; i < n; ++i) {
      std::vector<Result> v;
      for (uint64_t k = off[i]; k < off[i + 1]; ++k) {
        const Point a(FT(pts[6 * k]), FT(pts[6 * k + 1]), FT(pts[6 * k + 2]));
        if (kinds[k] == 3) v.push_back(Result(triangle_object<Intersection>(ids[k]), m_primitives[ids[k]].id()));
        else if (kinds[k] == 2) v.push_back(Result(Intersection(Segment(a, Point(FT(pts[6 * k + 3]), FT(pts[6 * k + 4]), FT(pts[6 * k + 5])))), m_primitives[ids[k]].id()));
        else v.push_back(Result(Intersection(a), m_primitives[ids[k]].id()));
      }
      *out++ = v;
    }
    return out;
  }
  // out receives std::optional< Intersection_and_primitive_id<Ray>::Type >; flags (optional) receives the B200AABB_FLAG_* bits
  template <typename RayIterator, typename OutputIterator>
  OutputIterator first_intersection(RayIterator first, RayIterator beyond, OutputIterator out, std::vector<uint8_t>* flags = nullptr) const {
    return first_intersection_masked(first, beyond, out, std::vector<uint32_t>(), flags);
  }
  // the same with the reference's skip functor (AABB_tree.h:307-328), evaluated once per primitive
  template <typename RayIterator, typename OutputIterator, typename SkipFunctor>
  OutputIterator first_intersection_skip(RayIterator first, RayIterator beyond, OutputIterator out, const SkipFunctor& skip, std::vector<uint8_t>* flags = nullptr) const {
    std::vector<uint32_t> mask((m_primitives.size() + 31) / 32 + 1, 0u);
    for (std::size_t k = 0; k < m_primitives.size(); ++k)
      if (skip(m_primitives[k].id())) mask[k >> 5] |= 1u << (k & 31);
    return first_intersection_masked(first, beyond, out, mask, flags);
  }
  // PMP::max_distance_to_mesh's inner loop (Polygon_mesh_processing/distance.h:98-190): the largest distance from the points of [first, beyond) to the primitives
  template <typename PointIterator>
  double max_distance_to_mesh(PointIterator first, PointIterator beyond) const {
    std::vector<double> q = pack_points(first, beyond);
    ensure_built();
    double d = 0;
    check(b200aabb_max_distance_to_mesh(m_handle, q.data(), q.size() / 3, B200AABB_HOST, &d, nullptr));
    return d;
  }

 private:
  template <typename RayIterator, typename OutputIterator>
  OutputIterator first_intersection_masked(RayIterator first, RayIterator beyond, OutputIterator out, const std::vector<uint32_t>& mask, std::vector<uint8_t>* flags) const {
    typedef typename Intersection_and_primitive_id<Ray>::Type Result;
    typedef typename Result::first_type Intersection;
    std::vector<double> q = pack_lines(first, beyond);
    const uint64_t n = q.size() / 6;
    std::vector<uint32_t> prim(n);
    std::vector<double> p0(3 * n), p1(3 * n);
    std::vector<uint8_t> fl(n);
    ensure_built();
    check(b200aabb_first_intersection_skip(m_handle, q.data(), n, B200AABB_HOST, nullptr, mask.empty() ? nullptr : mask.data(), (uint32_t)mask.size(), prim.data(), nullptr,
                                           p0.data(), p1.data(), fl.data()));
    for (uint64_t i = 0; i < n; ++i) {
      if (!(fl[i] & B200AABB_FLAG_HIT)) { *out++ = std::nullopt; continue; }
      const Point a(FT(p0[3 * i]), FT(p0[3 * i + 1]), FT(p0[3 * i + 2]));
      if (fl[i] & B200AABB_FLAG_SEGMENT) {
        const Point b(FT(p1[3 * i]), FT(p1[3 * i + 1]), FT(p1[3 * i + 2]));
        *out++ = Result(Intersection(Segment(a, b)), m_primitives[prim[i]].id());
      } else {
        *out++ = Result(Intersection(a), m_primitives[prim[i]].id());
      }
    }
    if (flags) flags->swap(fl);
    return out;
  }

 public:
  // out receives std::optional<Primitive_id>
  template <typename RayIterator, typename OutputIterator>
  OutputIterator first_intersected_primitive(RayIterator first, RayIterator beyond, OutputIterator out) const {
    std::vector<double> q = pack_lines(first, beyond);
    const uint64_t n = q.size() / 6;
    std::vector<uint32_t> prim(n);
    ensure_built();
    check(b200aabb_first_intersection(m_handle, q.data(), n, B200AABB_HOST, prim.data(), nullptr, nullptr, nullptr, nullptr));
    for (uint64_t i = 0; i < n; ++i) {
      if (prim[i] == B200AABB_NO_PRIMITIVE) *out++ = std::nullopt;
      else *out++ = std::optional<Primitive_id>(m_primitives[prim[i]].id());
    }
    return out;
  }
  // out receives Point_and_primitive_id; flags (optional) receives the B200AABB_FLAG_* bits (exact-distance ties)
  template <typename PointIterator, typename OutputIterator>
  OutputIterator closest_point_and_primitive(PointIterator first, PointIterator beyond, OutputIterator out, std::vector<uint8_t>* flags = nullptr) const {
    std::vector<double> q = pack_points(first, beyond);
    const uint64_t n = q.size() / 3;
    std::vector<double> cp(3 * n);
    std::vector<uint32_t> prim(n);
    std::vector<uint8_t> fl(n);
    ensure_built();
    check(b200aabb_closest_point(m_handle, q.data(), n, B200AABB_HOST, cp.data(), prim.data(), nullptr, fl.data()));
    for (uint64_t i = 0; i < n; ++i) *out++ = Point_and_primitive_id(Point(FT(cp[3 * i]), FT(cp[3 * i + 1]), FT(cp[3 * i + 2])), m_primitives[prim[i]].id());
    if (flags) flags->swap(fl);
    return out;
  }
  // out receives Point
  template <typename PointIterator, typename OutputIterator>
  OutputIterator closest_point(PointIterator first, PointIterator beyond, OutputIterator out) const {
    std::vector<double> q = pack_points(first, beyond);
    const uint64_t n = q.size() / 3;
    std::vector<double> cp(3 * n);
    ensure_built();
    check(b200aabb_closest_point(m_handle, q.data(), n, B200AABB_HOST, cp.data(), nullptr, nullptr, nullptr));
    for (uint64_t i = 0; i < n; ++i) *out++ = Point(FT(cp[3 * i]), FT(cp[3 * i + 1]), FT(cp[3 * i + 2]));
    return out;
  }
  // out receives FT
  template <typename PointIterator, typename OutputIterator>
  OutputIterator squared_distance(PointIterator first, PointIterator beyond, OutputIterator out) const {
    std::vector<double> q = pack_points(first, beyond);
    const uint64_t n = q.size() / 3;
    std::vector<double> d(n);
    ensure_built();
    check(b200aabb_closest_point(m_handle, q.data(), n, B200AABB_HOST, nullptr, nullptr, d.data(), nullptr));
    for (uint64_t i = 0; i < n; ++i) *out++ = FT(d[i]);
    return out;
  }

  // ---- multi-GPU: COLLECTIVE, every rank calls it with its own tree (created on the communicator's device).  On `root` the tree is built if needed and
  // sent; on the other ranks the device tree becomes a replica.  Every rank must have inserted the same primitives in the same order (they map the
  // primitive indices of the results to ids on the host; only the root's are uploaded).  mode 0: the built tree travels; 1: the primitives travel and every
  // rank builds.  Returns the device time of the replication on this rank in milliseconds.
  double replicate(Communicator& comm, int root, int rank, int mode = 0) {
    if (rank == root) ensure_built();
    double ms = 0;
    Communicator::check(b200aabb_comm_broadcast_tree(comm.handle(), m_handle, root, mode, &ms, nullptr), comm.handle());
    m_uploaded = true;
    return ms;
  }
  // the same for one process that drives several GPUs: trees[i] lives on the device of comms[i]
  static double replicate_all(const std::vector<Self*>& trees, const std::vector<std::unique_ptr<Communicator>>& comms, int root, int mode = 0) {
    trees[root]->ensure_built();
    std::vector<b200aabb_comm*> c;
    std::vector<b200aabb_tree*> t;
    for (size_t i = 0; i < trees.size(); ++i) { c.push_back(comms[i]->handle()); t.push_back(trees[i]->m_handle); }
    double ms = 0;
    Communicator::check(b200aabb_comm_broadcast_tree_all(c.data(), t.data(), (int)trees.size(), root, mode, &ms, nullptr), c[root]);
    for (Self* x : trees) x->m_uploaded = true;
    return ms;
  }

  // ---- access to the C ABI for callers that keep queries and results in device memory (b200aabb.h, B200AABB_DEVICE)
  b200aabb_tree* handle() const {
    ensure_built();
    return m_handle;
  }
  b200aabb_stats stats() const {
    b200aabb_stats s;
    check(b200aabb_get_stats(m_handle, &s));
    return s;
  }

 private:
  template <typename... T>
  void set_shared_data(T&&... t) {
    if constexpr (internal::Has_shared_data<Primitive>::value && sizeof...(T) > 0) m_traits.set_shared_data(std::forward<T>(t)...);
  }
  decltype(auto) datum(const Primitive& p) const {   // internal::Primitive_helper::get_datum, AABB_tree/internal/Primitive_helper.h:40-60
    if constexpr (internal::Has_shared_data<Primitive>::value) return p.datum(m_traits.shared_data());
    else return p.datum();
  }
  decltype(auto) reference_point(const Primitive& p) const {   // internal::Primitive_helper::get_reference_point, AABB_tree/internal/Primitive_helper.h:62-82
    if constexpr (internal::Has_shared_data<Primitive>::value) return p.reference_point(m_traits.shared_data());
    else return p.reference_point();
  }
  void create(int device) {
    const int rc = b200aabb_create(device, &m_handle);
    if (rc != B200AABB_OK) { m_handle = nullptr; throw Error(rc, "b200aabb_create failed (no sm_100 device? there is no CPU fallback)"); }
  }
  void destroy() {
    if (m_handle) b200aabb_destroy(m_handle);
    m_handle = nullptr;
  }
  void check(int rc) const {
    if (rc != B200AABB_OK) throw Error(rc, b200aabb_last_error(m_handle));
  }
  // lazy build on the first query, like root_node() (AABB_tree.h:636-652): thread-safe, once
  void ensure_built() const {
    if (m_uploaded) return;
    std::lock_guard<std::mutex> lock(m_mutex);
    if (!m_uploaded) const_cast<Self*>(this)->upload_and_build();
  }
  // a Segment_3 datum (AABB_segment_primitive_3, AABB_halfedge_graph_segment_primitive) has source() and target(); a Triangle_3 datum has vertex(i)
  template <typename D, typename = void> struct is_segment_datum : std::false_type {};
  template <typename D> struct is_segment_datum<D, std::void_t<decltype(std::declval<const D&>().source()), decltype(std::declval<const D&>().target())>> : std::true_type {};
  void upload_and_build() {
    const uint64_t n = m_primitives.size();
    typedef std::decay_t<decltype(datum(std::declval<const Primitive&>()))> Datum;
    if constexpr (is_segment_datum<Datum>::value) {
      std::vector<double> segs(6 * n);
      for (uint64_t i = 0; i < n; ++i) {
        decltype(auto) sg = datum(m_primitives[i]);
        decltype(auto) a = sg.source();
        decltype(auto) b = sg.target();
        segs[6 * i] = (double)a.x(); segs[6 * i + 1] = (double)a.y(); segs[6 * i + 2] = (double)a.z();
        segs[6 * i + 3] = (double)b.x(); segs[6 * i + 4] = (double)b.y(); segs[6 * i + 5] = (double)b.z();
      }
      check(b200aabb_set_segments(m_handle, segs.data(), nullptr, n, B200AABB_HOST));
    } else {
      std::vector<double> soup(9 * n);
      for (uint64_t i = 0; i < n; ++i) {
        decltype(auto) tri = datum(m_primitives[i]);
        for (int k = 0; k < 3; ++k) {
          decltype(auto) v = tri.vertex(k);
          soup[9 * i + 3 * k] = (double)v.x(); soup[9 * i + 3 * k + 1] = (double)v.y(); soup[9 * i + 3 * k + 2] = (double)v.z();
        }
      }
      check(b200aabb_set_triangles(m_handle, soup.data(), nullptr, n, B200AABB_HOST));
    }
    check(b200aabb_build(m_handle));
    m_uploaded = true;
  }
  // Query types are told apart by the kernel-object interface: Segment_3 / Ray_3 are the traits' own types; a Line_3 has point() and
  // to_vector() (Cartesian/Line_3.h), a Plane_3 has a(), b(), c(), d() (Cartesian/Plane_3.h)
  template <typename Q, typename = void> struct has_line_interface : std::false_type {};
  template <typename Q> struct has_line_interface<Q, std::void_t<decltype(std::declval<const Q&>().point()), decltype(std::declval<const Q&>().to_vector())>> : std::true_type {};
  template <typename Q, typename = void> struct has_plane_interface : std::false_type {};
  template <typename Q> struct has_plane_interface<Q, std::void_t<decltype(std::declval<const Q&>().a()), decltype(std::declval<const Q&>().d())>> : std::true_type {};
  // a Triangle_3 has vertex(int) and is none of the above (Cartesian/Triangle_3.h): do_intersect / number_of_intersected_primitives / any_ and
  // all_intersected_primitives only (the id queries)
  template <typename Q, typename = void> struct has_triangle_interface : std::false_type {};
  template <typename Q> struct has_triangle_interface<Q, std::void_t<decltype(std::declval<const Q&>().vertex(0))>> : std::true_type {};
  template <typename Query>
  static constexpr b200aabb_query_kind kind_of() {
    if constexpr (std::is_same<Query, Segment>::value) return B200AABB_SEGMENT;
    else if constexpr (std::is_same<Query, Ray>::value) return B200AABB_RAY;
    else if constexpr (has_plane_interface<Query>::value) return B200AABB_PLANE;
    else if constexpr (has_line_interface<Query>::value) return B200AABB_LINE;
    else {
      static_assert(has_triangle_interface<Query>::value, "offloaded intersection queries: Segment_3, Ray_3, Line_3, Plane_3 and Triangle_3");
      return B200AABB_TRIANGLE;
    }
  }
  // the object of a Plane_3 query that contains the whole primitive: the primitive's Triangle_3, when the traits' intersection type can hold one
  template <typename Intersection>
  Intersection triangle_object(uint32_t index) const {
    typedef std::decay_t<decltype(datum(m_primitives[index]))> Triangle;
    if constexpr (std::is_constructible<Intersection, Triangle>::value) return Intersection(datum(m_primitives[index]));
    else throw Error(B200AABB_ERR_INVALID_ARGUMENT, "the intersection is the primitive's Triangle_3, which this traits' Intersection_and_primitive_id type cannot hold");
  }
  template <typename Query>
  static constexpr uint64_t doubles_per_query() { return kind_of<Query>() == B200AABB_PLANE ? 4 : (kind_of<Query>() == B200AABB_TRIANGLE ? 9 : 6); }
  static void put(std::vector<double>& v, const Point& p) { v.push_back((double)p.x()); v.push_back((double)p.y()); v.push_back((double)p.z()); }
  template <typename Query>
  static void put_line(std::vector<double>& v, const Query& q) {
    if constexpr (kind_of<Query>() == B200AABB_SEGMENT) { put(v, q.source()); put(v, q.target()); }
    else if constexpr (kind_of<Query>() == B200AABB_RAY) { put(v, q.source()); put(v, q.second_point()); }   // Cartesian/Ray_3.h:35-47
    else if constexpr (kind_of<Query>() == B200AABB_PLANE) { v.push_back((double)q.a()); v.push_back((double)q.b()); v.push_back((double)q.c()); v.push_back((double)q.d()); }
    else if constexpr (kind_of<Query>() == B200AABB_TRIANGLE) { put(v, q.vertex(0)); put(v, q.vertex(1)); put(v, q.vertex(2)); }
    else {                                                                                                    // the stored members, not two points
      put(v, q.point());
      decltype(auto) d = q.to_vector();
      v.push_back((double)d.x()); v.push_back((double)d.y()); v.push_back((double)d.z());
    }
  }
  template <typename It>
  static std::vector<double> pack_lines(It first, It beyond) {
    std::vector<double> v;
    for (; first != beyond; ++first) put_line(v, *first);
    return v;
  }
  template <typename It>
  static std::vector<double> pack_points(It first, It beyond) {
    std::vector<double> v;
    for (; first != beyond; ++first) put(v, *first);
    return v;
  }

public:
  // Side_of_triangle_mesh::operator() on the tree of a CLOSED triangle mesh (Polygon_mesh_processing/include/CGAL/Side_of_triangle_mesh.h:265-306):
  // +1 ON_BOUNDED_SIDE, 0 ON_BOUNDARY, -1 ON_UNBOUNDED_SIDE — the values of CGAL::Bounded_side; out receives int.  See CGAL_B200::Side_of_triangle_mesh.
  template <typename PointIterator, typename OutputIterator>
  OutputIterator bounded_side(PointIterator first, PointIterator beyond, OutputIterator out) const {
    std::vector<double> q = pack_points(first, beyond);
    const uint64_t n = q.size() / 3;
    std::vector<int8_t> r(n);
    ensure_built();
    check(b200aabb_side_of_mesh(m_handle, q.data(), n, B200AABB_HOST, r.data(), nullptr));
    for (uint64_t i = 0; i < n; ++i) *out++ = (int)r[i];
    return out;
  }
  int bounded_side(const Point& p) const { int r = -1; const Point* a = &p; bounded_side(a, a + 1, &r); return r; }
  // the same with the mesh under an affine transformation and the points given in the frame of the transformed mesh (does_A_contains_a_CC_of_B,
  // Rigid_triangle_mesh_collision_detection.h:131-147)
  template <typename PointIterator, typename OutputIterator, typename Aff>
  OutputIterator bounded_side(PointIterator first, PointIterator beyond, OutputIterator out, const Aff& transformation) const {
    std::vector<double> q = pack_points(first, beyond);
    const uint64_t n = q.size() / 3;
    std::vector<int8_t> r(n);
    double m[12];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j) m[4 * i + j] = (double)transformation.m(i, j);
    ensure_built();
    check(b200aabb_side_of_mesh_transformed(m_handle, m, q.data(), n, B200AABB_HOST, r.data(), nullptr));
    for (uint64_t i = 0; i < n; ++i) *out++ = (int)r[i];
    return out;
  }

private:
  AABBTraits m_traits;
  Primitives m_primitives;
  b200aabb_tree* m_handle = nullptr;
  mutable std::atomic<bool> m_uploaded{true};   // the device tree reflects m_primitives (read without the lock by const queries)
  mutable std::mutex m_mutex;
};

// The functor of Polygon_mesh_processing/include/CGAL/Side_of_triangle_mesh.h over a CGAL_B200::AABB_tree (its constructor from an AABB_tree, :226-236):
//   CGAL_B200::Side_of_triangle_mesh<Tree> inside(tree);   int s = inside(p);   inside(points.begin(), points.end(), out);
// Values as CGAL::Bounded_side: ON_UNBOUNDED_SIDE = -1, ON_BOUNDARY = 0, ON_BOUNDED_SIDE = 1.  The mesh must be closed.
enum Bounded_side { ON_UNBOUNDED_SIDE = -1, ON_BOUNDARY = 0, ON_BOUNDED_SIDE = 1 };
template <typename Tree>
class Side_of_triangle_mesh {
public:
  explicit Side_of_triangle_mesh(const Tree& tree) : m_tree(&tree) {}
  Bounded_side operator()(const typename Tree::Point& p) const { return (Bounded_side)m_tree->bounded_side(p); }
  template <typename PointIterator, typename OutputIterator>
  OutputIterator operator()(PointIterator first, PointIterator beyond, OutputIterator out) const {
    std::vector<int> r;
    m_tree->bounded_side(first, beyond, std::back_inserter(r));
    for (int v : r) *out++ = (Bounded_side)v;
    return out;
  }
private:
  const Tree* m_tree;
};

// The face-intersection part of Polygon_mesh_processing/include/CGAL/Rigid_triangle_mesh_collision_detection.h over device trees: meshes are added once
// (one tree each, built once), moved by set_transformation, and get_all_intersections(id) lists the meshes with at least one face intersecting a face of
// mesh `id` under the current transformations (:349-388; one tree-vs-tree kernel per candidate, exact on the rounded images of the faces as the reference
// documents, :45-47).  Tree = a CGAL_B200::AABB_tree over the mesh type's faces; Aff = CGAL::Aff_transformation_3<K>.
// get_all_intersections_and_inclusions (:391-464) adds the volume inclusions of closed meshes: the caller states whether a mesh is closed and gives one point per
// connected component (the reference computes both from the mesh, :125-129, 259-260, with BGL algorithms that are host code outside this path).
template <typename Tree, typename Aff>
class Rigid_triangle_mesh_collision_detection {
public:
  typedef typename Tree::Point Point_3;
  // add_mesh with what the inclusion test needs: is the mesh closed, one point per connected component (:264-316, the overload that takes them precomputed)
  template <typename... T>
  std::size_t add_mesh_with_points(bool is_closed, const std::vector<Point_3>& points_per_cc, T&&... t) {
    const std::size_t id = add_mesh(std::forward<T>(t)...);
    m_is_closed[id] = is_closed;
    m_points_per_cc[id] = points_per_cc;
    return id;
  }
  template <typename MeshIdRange>
  std::vector<std::pair<std::size_t, bool>> get_all_intersections_and_inclusions(std::size_t mesh_id, const MeshIdRange& ids) const {   // :408-446
    std::vector<std::pair<std::size_t, bool>> res;
    for (std::size_t k : ids) {
      if (k == mesh_id || !is_valid_index(k)) continue;
      if (m_trees.at(mesh_id)->do_intersect(*m_trees[k], m_transformations[mesh_id], m_transformations[k])) res.push_back(std::make_pair(k, false));
      else if ((m_is_closed[mesh_id] && contains_a_cc_of(mesh_id, k)) || (m_is_closed[k] && contains_a_cc_of(k, mesh_id))) res.push_back(std::make_pair(k, true));
    }
    return res;
  }
  std::vector<std::pair<std::size_t, bool>> get_all_intersections_and_inclusions(std::size_t mesh_id) const {   // :454-461
    std::vector<std::size_t> all;
    for (std::size_t k = 0; k < m_trees.size(); ++k) all.push_back(k);
    return get_all_intersections_and_inclusions(mesh_id, all);
  }
  // add_mesh(tm) :231-262 — the caller passes what the Tree's range constructor takes (first, beyond, mesh); returns the id of the mesh
  template <typename... T>
  std::size_t add_mesh(T&&... t) {
    std::size_t id = 0;
    while (id < m_trees.size() && m_trees[id]) ++id;                      // the smallest free id, as the reference's id pool gives (:113-123)
    if (id == m_trees.size()) { m_trees.emplace_back(); m_transformations.emplace_back(); m_is_closed.push_back(false); m_points_per_cc.emplace_back(); }
    m_trees[id].reset(new Tree(std::forward<T>(t)...));
    m_trees[id]->build();
    m_transformations[id] = Aff();
    m_is_closed[id] = false;
    m_points_per_cc[id].clear();
    return id;
  }
  void remove_mesh(std::size_t id) { m_trees.at(id).reset(); }                                   // :481-497
  void set_transformation(std::size_t id, const Aff& transformation) { m_transformations.at(id) = transformation; }   // :320-330
  bool is_valid_index(std::size_t id) const { return id < m_trees.size() && m_trees[id]; }       // :513-517
  std::size_t size() const { std::size_t n = 0; for (const auto& t : m_trees) n += t ? 1 : 0; return n; }
  template <typename MeshIdRange>
  std::vector<std::size_t> get_all_intersections(std::size_t mesh_id, const MeshIdRange& ids) const {   // :354-377
    std::vector<std::size_t> res;
    for (std::size_t k : ids) {
      if (k == mesh_id || !is_valid_index(k)) continue;
      if (m_trees.at(mesh_id)->do_intersect(*m_trees[k], m_transformations[mesh_id], m_transformations[k])) res.push_back(k);
    }
    return res;
  }
  std::vector<std::size_t> get_all_intersections(std::size_t mesh_id) const {                      // :381-388
    std::vector<std::size_t> all;
    for (std::size_t k = 0; k < m_trees.size(); ++k) all.push_back(k);
    return get_all_intersections(mesh_id, all);
  }
private:
  // does_A_contains_a_CC_of_B (:131-147; precondition: A and B do not intersect): one point per connected component of B, moved by B's transformation as
  // Aff_transformation_3::transform(Point_3) does, classified against A under A's transformation
  bool contains_a_cc_of(std::size_t a, std::size_t b) const {
    if (m_points_per_cc[b].empty()) return false;
    const Aff& tb = m_transformations[b];
    std::vector<Point_3> moved;
    for (const Point_3& q : m_points_per_cc[b])
      moved.push_back(Point_3(tb.m(0, 0) * q.x() + tb.m(0, 1) * q.y() + tb.m(0, 2) * q.z() + tb.m(0, 3), tb.m(1, 0) * q.x() + tb.m(1, 1) * q.y() + tb.m(1, 2) * q.z() + tb.m(1, 3),
                              tb.m(2, 0) * q.x() + tb.m(2, 1) * q.y() + tb.m(2, 2) * q.z() + tb.m(2, 3)));
    std::vector<int> side;
    m_trees[a]->bounded_side(moved.begin(), moved.end(), std::back_inserter(side), m_transformations[a]);
    for (int s : side) if (s == ON_BOUNDED_SIDE) return true;
    return false;
  }
  std::vector<std::unique_ptr<Tree>> m_trees;
  std::vector<Aff> m_transformations;
  std::vector<bool> m_is_closed;
  std::vector<std::vector<Point_3>> m_points_per_cc;
};

}  // namespace CGAL_B200

#endif  // CGAL_B200_AABB_TREE_H

/* b200aabb.h — C ABI of the B200-native AABB tree (libb200aabb.so).
 *
 * This is the drop-in boundary for the hot path of
 *   CGAL::AABB_tree< AABB_traits_3< Epick, AABB_face_graph_triangle_primitive< Surface_mesh<Epick::Point_3> > > >
 * Each entry point names the reference interface it replaces (paths relative to the CGAL source root).
 * Conventions: plain pointers and sizes only; every function returns an int status (0 = ok, < 0 = error, never
 * throws); batched queries take `n` records and a memory-space flag saying whether the query/result pointers are
 * host or device pointers.  Geometry is IEEE binary64; primitive ids are uint32 (Surface_mesh face indices,
 * Surface_mesh/include/CGAL/Surface_mesh/Surface_mesh.h:63,198).  "No result" is id 0xFFFFFFFF.
 *
 * The reference-side binding a maintainer would add is include/CGAL_B200/AABB_tree.h.
 */
#ifndef B200AABB_H
#define B200AABB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200aabb_tree b200aabb_tree; /* opaque; owns all device buffers */

enum {
  B200AABB_OK = 0,
  B200AABB_ERR_INVALID_ARGUMENT = -1,
  B200AABB_ERR_CUDA = -2,          /* a CUDA call failed; see b200aabb_last_error() */
  B200AABB_ERR_OUT_OF_MEMORY = -3,
  B200AABB_ERR_EMPTY_TREE = -4,    /* CGAL_precondition(!empty()) of the distance queries, AABB_tree.h:993,1008,1039 */
  B200AABB_ERR_NO_DEVICE = -5,     /* no sm_100 device: there is NO CPU fallback */
  B200AABB_ERR_NCCL = -7,          /* NCCL is not available (libnccl.so.2 could not be loaded) or an NCCL call failed */
  B200AABB_ERR_INCOMPLETE = -6     /* a traversal could not finish (stack overflow on a tree deeper than this library ever builds): the results of
                                      the batch are incomplete.  HOST batches report it from the call itself, DEVICE batches (asynchronous) from the
                                      next b200aabb_get_stats */
};

#define B200AABB_NO_PRIMITIVE 0xFFFFFFFFu   /* (a macro: the value does not fit the int of a C enumerator) */

/* where the query / result pointers of a batched call live */
typedef enum { B200AABB_HOST = 0, B200AABB_DEVICE = 1 } b200aabb_memspace;

/* Query type of the intersection queries (do_intersect, number_of_intersected_primitives, any_ / all_intersected_primitives):
 *   Segment_3 (source, target) and Ray_3 (source, second point): 6 doubles (Cartesian_kernel/include/CGAL/Cartesian/Ray_3.h:35-47, Segment_3.h);
 *   Line_3: 6 doubles = a point of the line + its DIRECTION VECTOR, the two members a Cartesian Line_3 stores (Cartesian/Line_3.h: line.point(),
 *           line.to_vector()) — Line_3(p, q) keeps the ROUNDED difference q - p, and the reference's exact predicates work on what is stored;
 *   Plane_3: 4 doubles = the coefficients (a, b, c, d) of a x + b y + c z + d = 0 (Cartesian/Plane_3.h);
 *   Triangle_3: 9 doubles = the three vertices (Do_intersect_3(Triangle_3, Triangle_3), Intersections_3/internal/Triangle_3_Triangle_3_do_intersect.h:154-450;
 *           the id queries only — the intersection OBJECTS of the *_q calls take kinds 0..3; flat triangles are outside the reference's precondition).
 * first_intersection takes Ray_3 only. */
typedef enum { B200AABB_SEGMENT = 0, B200AABB_RAY = 1, B200AABB_LINE = 2, B200AABB_PLANE = 3, B200AABB_TRIANGLE = 4 } b200aabb_query_kind;

/* per-query result flags (bit set) */
enum {
  B200AABB_FLAG_HIT = 1,           /* first_intersection: an intersection exists */
  B200AABB_FLAG_SEGMENT = 2,       /* first_intersection: the intersection object is a Segment_3 (coplanar ray), not a Point_3 */
  B200AABB_FLAG_TIE = 4,           /* >= 2 primitives attain the winning value EXACTLY (equal t / equal squared distance):
                                      the reference's id then depends on its traversal order;
                                      this library reports the smallest primitive id among them */
  B200AABB_FLAG_EXACT = 8,         /* at least one predicate of this query needed the exact (expansion) stage */
  B200AABB_FLAG_NEAR_TIE = 16      /* first_intersection only: a primitive with a LARGER rounded t lies within the ROUNDING UNCERTAINTY of the
                                      winner's (the plain-double hit construction; large for a ray grazing a triangle; independent of FLAG_TIE).  The reference compares
                                      the same rounded values while pruning boxes with an inexact test (AABB_traits_3.h:89-125;
                                      AABB_ray_intersection.h:78), so in this band its own answer depends on the order in which its tree presents
                                      the candidates; this library reports the smallest rounded value.  closest_point never sets it: there the
                                      reference compares candidates exactly and never prunes when unsure, so only exact ties are order dependent */,
  B200AABB_FLAG_TRIANGLE = 32      /* any_intersection of a Plane_3 query: the object is the primitive's Triangle_3 itself (it lies in the plane); no point is returned */
};

/* b200aabb_set_option keys */
enum {
  B200AABB_OPT_SORT_QUERIES = 1,   /* 1 (default): every batch is traversed in a coherence order (Morton order of the points /
                                      of ray origin+direction, onesweep radix sort on the device); 0: in the caller's order */
  B200AABB_OPT_CLOSEST_KERNEL = 2, /* closest_point scheduling: 1 (default) = warp-cooperative work pool (32 query slots per persistent warp, pooled
                                      (slot,node)/(slot,leaf) pairs, closest_warp.cuh); 2 = lane pairs walk the tree with private stacks
                                      (closest_pair.cuh; measured slower on the benchmark, kept as the reference point of that design);
                                      3 = two passes: a binary32-only traversal per thread that keeps the candidates inside the final ball, then the
                                      exact projection and comparison of the candidates (2.7x slower: its warps diverge; kept for the same reason);
                                      0 = one thread per query.  Same results. */
  B200AABB_OPT_HOST_PARTS = 5,     /* HOST-pointer batches are cut into about this many chunks (each 256 Ki .. 4 Mi rays / segments, .. 16 Mi points) that flow
                                      through an upload -> sort -> traverse -> download pipeline of three streams (1..64; 0 = default: 8 for ray / segment batches;
                                      1 for point batches — a point batch is sorted into ONE coherence order) */
  B200AABB_OPT_ASYNC_HOST = 6,     /* 1: batched queries with HOST pointers return as soon as their chunks are enqueued (uploads, kernels and downloads of
                                      consecutive calls then overlap across calls); the caller keeps the query / result buffers alive and unread until
                                      b200aabb_synchronize.  0 (default): every call returns with its results in place. */
  B200AABB_OPT_RAY_KERNEL = 7,     /* first_intersection, do_intersect, counts, any_intersected_primitive: 0 (default) = one thread per query; 1 = persistent threads that
                                      take the next query of the coherence order as soon as theirs is answered (same traversal, identical results; measured equal or slower) */
  B200AABB_OPT_WARP_EXACT_GO = 4   /* warp-cooperative kernels: run the binary64 phase once this many pre-test survivors wait (1..160; 0 = the kernel's default) */
};

typedef struct b200aabb_stats {
  uint64_t kernel_launches;        /* kernels of this library launched since the handle was created */
  uint64_t exact_fallbacks;        /* queries of the last batch that used the exact stage (FLAG_EXACT) */
  uint64_t ties;                   /* queries of the last batch flagged FLAG_TIE */
  uint64_t near_ties;              /* queries of the last batch flagged FLAG_NEAR_TIE */
  uint64_t nodes_visited;          /* BVH nodes loaded by the last batch (traversal-traffic model) */
  uint64_t leaves_tested;          /* primitives visited by the last batch */
  uint64_t leaves_culled;          /* of those, rejected by the rigorous binary32 pre-test before any binary64 work (closest_point) */
  double last_build_ms;            /* device time of the last b200aabb_build (CUDA events) */
  double last_query_ms;            /* device time of the traversal kernel(s) of the last batched query */
  double last_query_total_ms;      /* same, including host<->device copies of the batch when pointers were HOST */
} b200aabb_stats;

/* ---- lifetime ------------------------------------------------------------------------------------------------ */
/* AABB_tree(const Traits& = Traits()), AABB_tree/include/CGAL/AABB_tree.h:115-116.  `device` = CUDA ordinal. */
int b200aabb_create(int device, b200aabb_tree** out);
/* ~AABB_tree(), AABB_tree.h:186-189 */
int b200aabb_destroy(b200aabb_tree* t);
/* all work of this handle is enqueued on `cuda_stream` (a cudaStream_t; NULL = the legacy default stream) */
int b200aabb_set_stream(b200aabb_tree* t, void* cuda_stream);
int b200aabb_set_option(b200aabb_tree* t, int option, int64_t value);
const char* b200aabb_last_error(const b200aabb_tree* t);
int b200aabb_get_stats(const b200aabb_tree* t, b200aabb_stats* out);
/* waits for every batch in flight on the handle (B200AABB_OPT_ASYNC_HOST, DEVICE-pointer batches) and returns B200AABB_ERR_INCOMPLETE /
 * INVALID_ARGUMENT if one of them reported it; the per-batch statistics of asynchronous batches are not kept apart */
int b200aabb_synchronize(b200aabb_tree* t);

/* ---- primitives ---------------------------------------------------------------------------------------------- */
/* The mesh as the TEXT of an OFF file (HOST memory): what `CGAL::IO::read_OFF(in, surface_mesh)` + `Tree(faces(mesh).first, faces(mesh).second, mesh)` give
 * (Surface_mesh/include/CGAL/Surface_mesh/IO/OFF.h:323-345, Stream_support/include/CGAL/IO/OFF/File_scanner_OFF.h): primitive i = facet i of the file, vertex
 * coordinates = the correctly rounded binary64 values of the decimal literals (as `std::istream >> double` yields).  The host reads the header only; the body is
 * tokenised and parsed on the device (csrc/ingest.cuh).  Accepted: plain "OFF" files whose facets are all triangles and that carry no colours, normals or
 * comments after the header; anything else returns INVALID_ARGUMENT (b200aabb_last_error says why) and leaves the tree empty. */
int b200aabb_set_mesh_off(b200aabb_tree* t, const char* text, uint64_t len);
/* Segment_3 primitives: AABB_segment_primitive_3 over a range of segments (AABB_tree/include/CGAL/AABB_segment_primitive_3.h) and
 * AABB_halfedge_graph_segment_primitive over the edges of a graph (AABB_halfedge_graph_segment_primitive.h); segs = n x 6 doubles (source, target), the
 * reference point of a primitive is its source, ids as for b200aabb_set_triangles.  Such a tree answers b200aabb_closest_point (the reference's
 * Construct_projected_point_3(Segment_3, Point_3), Cartesian/function_objects.h:3266-3285, exact comparisons, ties reported), b200aabb_max_distance_to_mesh, and
 * the id queries (do_intersect / number_of_intersected_primitives / any_ / all_intersected_primitives) with Plane_3 and Triangle_3 queries — what
 * AABB_tree/examples/AABB_tree/AABB_segment_3_example.cpp asks of it; every other entry point returns INVALID_ARGUMENT on it. */
int b200aabb_set_segments(b200aabb_tree* t, const double* segs, const uint32_t* ids, uint64_t n, b200aabb_memspace space);
/* insert(faces(m).first, faces(m).second, m), AABB_tree.h:730-747, for a triangle Surface_mesh: vertex coordinates
 * (nv x 3 doubles = mesh.points().data(), Surface_mesh/include/CGAL/Surface_mesh/Properties.h:186) and, per face in
 * faces(m) order, (target(h), target(next(h)), source(h)) with h = halfedge(f)
 * (BGL/include/CGAL/boost/graph/property_maps.h:52-63) = the OFF face's (v0,v1,v2).  Primitive id of face i is
 * `face_ids[i]`, or i when face_ids is NULL.  Replaces the previous content.  Host pointers. */
int b200aabb_set_mesh(b200aabb_tree* t, const double* vertices, uint64_t nv, const uint32_t* faces, const uint32_t* face_ids, uint64_t nf);
/* same for a triangle soup: n x 9 doubles (v0,v1,v2); `space` says where `triangles`/`ids` live. */
int b200aabb_set_triangles(b200aabb_tree* t, const double* triangles, const uint32_t* ids, uint64_t n, b200aabb_memspace space);
/* clear(), AABB_tree.h:194-201 */
int b200aabb_clear(b200aabb_tree* t);
/* build() / custom_build(), AABB_tree.h:820-853: LBVH (63-bit Morton, onesweep radix sort, Karras, atomic refit). */
int b200aabb_build(b200aabb_tree* t);
/* Deforming meshes: the reference's rebuild(first, beyond) (AABB_tree.h:167,752) is clear + insert + build.  When the
 * primitives only MOVE (same number, same order, same ids) the hierarchy can be kept: these calls replace the coordinates and recompute every
 * box bottom-up over the existing topology (no sort, no hierarchy pass).  Query results are those of a fresh build — they never depend on the
 * tree (exact leaf decisions, smallest-id tie rule) — only the traversal cost grows as the primitives drift from their Morton order.
 * b200aabb_refit: triangles = n x 9 doubles in the inserted order (HOST or DEVICE); b200aabb_refit_mesh: as b200aabb_set_mesh (host pointers). */
int b200aabb_refit(b200aabb_tree* t, const double* triangles, uint64_t n, b200aabb_memspace space);
int b200aabb_refit_mesh(b200aabb_tree* t, const double* vertices, uint64_t nv, const uint32_t* faces, uint64_t nf);
/* accelerate_distance_queries(), AABB_tree.h:896-901.  The kd-tree hint is replaced by a BVH-derived bound (nearest
 * child first descent), so this only builds the tree if needed and returns OK. */
int b200aabb_accelerate_distance_queries(b200aabb_tree* t);
/* size(), AABB_tree.h:213 ; bbox(), AABB_tree.h:204-210 (xmin,ymin,zmin,xmax,ymax,zmax; empty tree: +inf.. -inf..) */
int b200aabb_size(const b200aabb_tree* t, uint64_t* n);
int b200aabb_bbox(b200aabb_tree* t, double out[6]);

/* ---- batched queries (each lazily builds, like root_node(), AABB_tree.h:636-652) ------------------------------ */
/* do_intersect(query), AABB_tree.h:903-913: out[i] = 1 iff query i intersects some primitive.  queries: n x 6 doubles (Plane_3: n x 4, Triangle_3: n x 9). */
int b200aabb_do_intersect(b200aabb_tree* t, const double* queries, uint64_t n, b200aabb_query_kind kind,
                          b200aabb_memspace space, uint8_t* out);
/* number_of_intersected_primitives(query), AABB_tree.h:915-932 */
int b200aabb_number_of_intersected_primitives(b200aabb_tree* t, const double* queries, uint64_t n, b200aabb_query_kind kind,
                                              b200aabb_memspace space, uint32_t* out);
/* any_intersected_primitive(query), AABB_tree.h:947-956 (First_primitive_traits, internal/AABB_traversal_traits.h:188-231):
 * out[i] = id of ONE intersected primitive (which one depends on the traversal order, in the reference too), or
 * B200AABB_NO_PRIMITIVE. */
int b200aabb_any_intersected_primitive(b200aabb_tree* t, const double* queries, uint64_t n, b200aabb_query_kind kind,
                                       b200aabb_memspace space, uint32_t* out);
/* all_intersected_primitives(query, out), AABB_tree.h:934-945 (Listing_primitive_traits, :150-184).  Variable-length output in two
 * steps: the caller runs b200aabb_number_of_intersected_primitives, forms offsets[n+1] = exclusive prefix sums of the counts, and
 * passes room for offsets[n] ids; the ids of query i are written to ids[offsets[i] .. offsets[i+1]) (in traversal order, which is
 * unspecified in the reference as well).  A count that disagrees with the offsets is an error (INVALID_ARGUMENT; for DEVICE
 * pointers it is reported by the next b200aabb_get_stats); nothing is ever written outside a query's own range. */
int b200aabb_all_intersected_primitives(b200aabb_tree* t, const double* queries, uint64_t n, b200aabb_query_kind kind,
                                        b200aabb_memspace space, const uint64_t* offsets, uint32_t* ids);
/* tree.do_intersect(other_tree): the AABB_tree-as-query overloads of AABB_traits_3::Do_intersect (AABB_traits_3.h:679-690: other_tree.do_intersect(bbox)
 * per node, other_tree.do_intersect(datum) per primitive; used by Polygon_mesh_processing/Rigid_triangle_mesh_collision_detection.h).  *result = 1 iff some
 * primitive of `t` intersects some primitive of `other` (exact, symmetric, independent of both trees' layouts); prim / other_prim (optional) receive the
 * ids of ONE intersecting pair (which one is unspecified), B200AABB_NO_PRIMITIVE when there is none.  Both trees must be on the same device; one kernel:
 * the primitives of the smaller tree are Triangle_3 queries against the larger one, all threads stop at the first pair found. */
int b200aabb_do_intersect_tree(b200aabb_tree* t, b200aabb_tree* other, int* result, uint32_t* prim, uint32_t* other_prim);
/* The same with each tree standing under an affine transformation — Rigid_triangle_mesh_collision_detection::does_A_intersect_B
 * (Polygon_mesh_processing/include/CGAL/Rigid_triangle_mesh_collision_detection.h:149-161) with
 * Do_intersect_traversal_traits_with_transformation / _for_two_trees (Polygon_mesh_processing/internal/AABB_traversal_traits_with_transformation.h:113-306):
 * the trees are built once on the untransformed meshes; per call every primitive is transformed as Aff_transformation_3::transform(Point_3) does
 * (Cartesian/Aff_transformation_rep_3.h:111-117, in binary64, left to right) and the exact Do_intersect_3(Triangle_3, Triangle_3) is evaluated on the
 * rounded images, as the reference states (:45-47); node boxes are transformed conservatively.  xf / other_xf: 12 doubles = m(i, j) of the
 * Aff_transformation_3, row major (t11 t12 t13 t14  t21 ... t34); NULL = identity. */
int b200aabb_do_intersect_tree_transformed(b200aabb_tree* t, const double* xf, b200aabb_tree* other, const double* other_xf, int* result, uint32_t* prim,
                                           uint32_t* other_prim);
/* all_intersections(Ray_3, out), AABB_tree.h:958-970 (Listing_intersection_traits, internal/AABB_traversal_traits.h:105-144): every
 * primitive whose Intersect_3(Ray_3, Triangle_3) is non-empty, with the intersection object.  Two steps like all_intersected_primitives:
 * b200aabb_number_of_intersections gives the counts (they can differ from number_of_intersected_primitives for a ray nearly parallel to a
 * triangle: the reference's construction evaluates its predicates on a rounded point), the caller forms offsets[n+1] and
 * passes room for offsets[n] entries: ids[k], kinds[k] (1 = Point_3, 2 = Segment_3), points[6k..6k+5] (the point, or the two endpoints). */
int b200aabb_number_of_intersections(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_memspace space, uint32_t* out);
int b200aabb_all_intersections(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_memspace space, const uint64_t* offsets,
                               uint32_t* ids, uint8_t* kinds, double* points);
/* The same three calls for every query kind: B200AABB_RAY as above; B200AABB_SEGMENT (source, target): Intersect_3(Segment_3, Triangle_3),
 * Intersections_3/include/CGAL/Intersections_3/internal/Segment_3_Triangle_3_intersection.h:400-548; B200AABB_LINE (point, direction vector):
 * Line_3_Triangle_3_intersection.h:279-400 — a construction, so the second point of the line is the ROUNDED point + vector there, unlike in do_intersect
 *   (and unlike in the reference's own node test, which is the exact predicate: where a line grazes a primitive within rounding the reference's answer
 *   depends on its tree; this library returns the tree-independent set, every primitive whose Intersect_3 is non-empty, of which the reference's is a subset);
 * B200AABB_PLANE (a, b, c, d: 4 doubles per query): Plane_3_Triangle_3_intersection.h:27-146, whose object may also be the primitive's Triangle_3 itself
 * (kinds[k] == 3 / B200AABB_FLAG_TRIANGLE: no points; the caller has the primitive). */
int b200aabb_number_of_intersections_q(b200aabb_tree* t, const double* queries, uint64_t n, b200aabb_query_kind kind, b200aabb_memspace space, uint32_t* out);
int b200aabb_all_intersections_q(b200aabb_tree* t, const double* queries, uint64_t n, b200aabb_query_kind kind, b200aabb_memspace space, const uint64_t* offsets,
                                 uint32_t* ids, uint8_t* kinds, double* points);
int b200aabb_any_intersection_q(b200aabb_tree* t, const double* queries, uint64_t n, b200aabb_query_kind kind, b200aabb_memspace space, uint32_t* prim, double* point,
                                double* point2, uint8_t* flags);
/* any_intersection(Ray_3), AABB_tree.h:972-985 (First_intersection_traits, :58-101): ONE intersection object per ray (which one depends on
 * the traversal order, in the reference too).  Outputs as first_intersection: prim (NO_PRIMITIVE = nullopt), point, point2, flags
 * (FLAG_HIT, FLAG_SEGMENT, FLAG_EXACT); any may be NULL. */
int b200aabb_any_intersection(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_memspace space, uint32_t* prim, double* point,
                              double* point2, uint8_t* flags);
/* first_intersection(Ray_3) / first_intersected_primitive, AABB_tree/internal/AABB_ray_intersection.h:201-240.
 * rays: n x 6 (source, second point).  Outputs (any may be NULL): prim[n]; t[n] = ray parameter of the hit as the
 * reference computes it (:168-198); point[n x 3] = the Point_3, or the first endpoint of the Segment_3;
 * point2[n x 3] = second endpoint when FLAG_SEGMENT; flags[n]. */
int b200aabb_first_intersection(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_memspace space,
                                uint32_t* prim, double* tparam, double* point, double* point2, uint8_t* flags);
/* first_intersection(query, skip) / first_intersected_primitive(query, skip), AABB_tree.h:307-354 (the skip functor is consulted at every
 * primitive, internal/AABB_ray_intersection.h:86,99,113): primitives the functor rejects are ignored.  A device kernel can not call the caller's
 * functor, so it is passed as data: skip_ids[n] = ONE primitive id per ray to ignore (B200AABB_NO_PRIMITIVE = none; the usual "the face
 * the ray starts on"), and / or skip_mask = one bit per primitive id (bit id & 31 of word id >> 5; mask_words words) for an arbitrary
 * predicate evaluated once per primitive.  Either may be NULL.  Pointers live in `space` like the rays. */
int b200aabb_first_intersection_skip(b200aabb_tree* t, const double* rays, uint64_t n, b200aabb_memspace space, const uint32_t* skip_ids,
                                     const uint32_t* skip_mask, uint32_t mask_words, uint32_t* prim, double* tparam, double* point, double* point2,
                                     uint8_t* flags);
/* closest_point_and_primitive / closest_point / squared_distance (Point_3), AABB_tree.h:987-1055.
 * points: n x 3.  Outputs (any may be NULL): closest[n x 3], prim[n], sqdist[n], flags[n]. */
int b200aabb_closest_point(b200aabb_tree* t, const double* points, uint64_t n, b200aabb_memspace space,
                           double* closest, uint32_t* prim, double* sqdist, uint8_t* flags);

/* ---- multi-GPU: a built tree is a set of flat device buffers that can be broadcast (NCCL) and adopted ----------
 * b200aabb_export_layout: fills ptrs/bytes (device pointers, sizes) for up to `cap` buffers, *count = how many.
 * b200aabb_import_alloc: on a receiving rank, allocates the buffers of a built tree of `n` primitives so that
 * export_layout returns same-sized buffers to receive into; b200aabb_import_finish marks the tree built. */
int b200aabb_export_layout(b200aabb_tree* t, void** ptrs, uint64_t* bytes, int cap, int* count);
int b200aabb_import_alloc(b200aabb_tree* t, uint64_t n);
int b200aabb_import_finish(b200aabb_tree* t);

/* ---- consumers of the query path ----------------------------------------------------------------
 * Polygon_mesh_processing::max_distance_to_mesh / approximate_Hausdorff_distance's inner loop
 * (Polygon_mesh_processing/include/CGAL/Polygon_mesh_processing/distance.h:98-190, max_distance_to_mesh_impl): the largest distance from the
 * sample points to the tree's primitives = sqrt of the largest squared_distance(closest_point(p), p).  The reference chains each
 * query's result as the next query's hint (:133,180); a hint only decides exact ties, never the distance, so the value is the same.
 * *max_distance = that distance; *farthest (optional) = index of the first sample attaining it.  points: n x 3, HOST or DEVICE. */
int b200aabb_max_distance_to_mesh(b200aabb_tree* t, const double* points, uint64_t n, b200aabb_memspace space, double* max_distance, uint64_t* farthest);
/* Side_of_triangle_mesh::operator()(Point_3) (Polygon_mesh_processing/include/CGAL/Side_of_triangle_mesh.h:265-306; the ray casts of
 * internal/Side_of_triangle_mesh/Point_inside_vertical_ray_cast.h:69-121 with Ray_3_Triangle_3_traversal_traits.h:30-76): for a CLOSED triangle mesh,
 * side[i] = +1 ON_BOUNDED_SIDE, 0 ON_BOUNDARY, -1 ON_UNBOUNDED_SIDE (the values of CGAL::Bounded_side) of point i.  A point outside the tree's bbox is
 * outside; otherwise one ray from the point counts the facets it crosses, with exact predicates; a ray that passes through an edge or a vertex, or lies
 * in a facet's plane, is replaced by another (first the vertical ray towards the nearer z face of the box, then a fixed sequence of directions) — the
 * answer does not depend on the rays, only attempts[i] (optional: rays cast for point i; 0 = decided by the bbox) does.  As in the reference, the
 * result is unspecified for a mesh that is not closed.  points: n x 3, HOST or DEVICE. */
int b200aabb_side_of_mesh(b200aabb_tree* t, const double* points, uint64_t n, b200aabb_memspace space, int8_t* side, uint8_t* attempts);
/* The same with the mesh under an affine transformation (xf: 12 doubles as for b200aabb_do_intersect_tree_transformed; NULL = identity) and the points given
 * in the frame of the transformed mesh: does_A_contains_a_CC_of_B of Rigid_triangle_mesh_collision_detection.h:131-147 (Point_inside_vertical_ray_cast over the
 * Transformed_tree_helper, AABB_traversal_traits_with_transformation.h:190-224).  The rounded images of the primitives are classified exactly. */
int b200aabb_side_of_mesh_transformed(b200aabb_tree* t, const double* xf, const double* points, uint64_t n, b200aabb_memspace space, int8_t* side, uint8_t* attempts);

/* ---- multi-GPU behind the boundary --------------------------------------------------------------------
 * The reference has no multi-device path: its callers shard queries over threads (tbb::parallel_for,
 * Polygon_mesh_processing/include/CGAL/Polygon_mesh_processing/distance.h:127-169).  Here the tree is built on ONE GPU and replicated with
 * NCCL over NVLink / NVSwitch; every GPU then answers its own block of the batch against its replica (no collective in the query loop).
 * NCCL is loaded at run time (libnccl.so.2); without it these calls return B200AABB_ERR_NCCL and everything else works.
 *
 * One process per GPU: rank 0 calls b200aabb_comm_unique_id and hands the 128 bytes to the other ranks by any means (MPI, a file, a
 * socket); every rank calls b200aabb_comm_init.  One process driving several GPUs: b200aabb_comm_init_all (ncclCommInitAll) and the
 * *_all variant of the broadcast.  Both warm the communicator up, so the first replication does not pay NCCL's channel set-up. */
typedef struct b200aabb_comm b200aabb_comm;
int b200aabb_comm_unique_id(char id[128]);
int b200aabb_comm_init(int device, int rank, int world, const char id[128], b200aabb_comm** out);
int b200aabb_comm_init_all(int ndev, const int* devices /* NULL = 0..ndev-1 */, b200aabb_comm** out /* [ndev] */);
int b200aabb_comm_destroy(b200aabb_comm* c);
const char* b200aabb_comm_last_error(const b200aabb_comm* c);
/* Replicates the tree of rank `root` into the handle of every other rank (collective: every rank calls it with its own handle, created
 * on the communicator's device).  mode 0: broadcast the built tree (64-byte nodes + 96-byte primitive records + root box; the receiver
 * derives the binary32 record copies itself); mode 1: broadcast the inserted primitives (72 bytes each, + ids) and build on every rank
 * (the build runs at several G primitives/s: SURVEY.md §8e asks for both to be measured).  ms / bytes (optional): device time of the
 * replication on this rank and the bytes it moved. */
int b200aabb_comm_broadcast_tree(b200aabb_comm* c, b200aabb_tree* t, int root, int mode, double* ms, uint64_t* bytes);
int b200aabb_comm_broadcast_tree_all(b200aabb_comm** comms, b200aabb_tree** trees, int ndev, int root, int mode, double* ms, uint64_t* bytes);

/* ---- self-test hook (used by tests/, not by the reference-side binding) ---------------------------------------------
 * Evaluates one DEVICE predicate / construction per record (host pointers).  which: 0 Orientation_3(p,q,r,s) [12 doubles],
 * 1 ray_direction orientation(a,b,c, a+(q-p)) exact [15], 2 Coplanar_orientation_3(p,q,r) [9], 3 Compare_distance_3(p,q,r) [9],
 * 4 Do_intersect_3(Triangle_3, Segment_3) [15: triangle, source, target], 5 Do_intersect_3(Triangle_3, Ray_3) [15],
 * 6 Intersect_3(Triangle_3, Ray_3) [15] -> result = 0 none / 1 Point_3 / 2 Segment_3, values[8i..] = t, point, point2, t uncertainty,
 * 7 Construct_projected_point_3(Triangle_3, Point_3) [12] -> values[8i..8i+2], 8 the exact stage of Orientation_3 alone [12].
 * used_exact[i] (optional) = number of predicate evaluations of the record that reached the exact stage. */
int b200aabb_debug_predicate(int device, int which, const double* records, int doubles_per_record, uint64_t n, int* result, int* used_exact, double* values);

/* library / build identification, e.g. "b200aabb 0.1 sm_100a" */
const char* b200aabb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200AABB_H */

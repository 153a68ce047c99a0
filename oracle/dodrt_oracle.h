/* TEST INFRASTRUCTURE -- CPU oracle for the dod_raytracer hot path.  NOT product code:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it, and only as the checker.  See dodrt_oracle.c for the per-function reference citations.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against the reference's own
 * translation units (oracle/_ref/libdodrt_ref.so, built from /root/reference by oracle/Makefile) in
 * tests/test_oracle_vs_ref.py, and against the committed fixtures in tests/golden/ (generated from
 * that library by tests/golden/make_golden.py) in tests/test_oracle_golden.py.  The reference has
 * no golden vectors of its own.
 */
#ifndef DODRT_ORACLE_H
#define DODRT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MISS 0xFFFFFFFFu
#define ORC_KIND_SHIFT 29
enum { ORC_KIND_TRIANGLE = 0, ORC_KIND_SPHERE = 1, ORC_KIND_PLANE = 2, ORC_KIND_CYLINDER = 3, ORC_KIND_BOX = 4 };
enum { ORC_CLS_SPHERE = 1, ORC_CLS_PLANE = 2, ORC_CLS_CYLINDER = 4, ORC_CLS_TREE = 8, ORC_CLS_BOX = 16 };
enum { ORC_RAY_ANY = 1 };

typedef struct {
    float o[3];
    float d[3];
    float clip;     /* clippingDistance, base_shape.h:13 */
    uint32_t flags; /* bit0 = returnOnAny, base_shape.h:12 */
} orc_ray;

typedef struct {
    float t;       /* record.t of the winning primitive; the ray's clip on a miss */
    uint32_t prim; /* kind << 29 | id ; ORC_MISS on a miss ; 0 for an any-hit ray that hit */
    float u, v;    /* triangle barycentrics (triangle.cpp:137), 0 otherwise */
} orc_hit;

typedef struct {
    uint32_t nodes;  /* kd nodes fetched (interior + leaf), 8 B each */
    uint32_t leaves; /* leaves visited */
    uint32_t lanes;  /* triangle lanes tested, 288 B each */
    uint32_t max_stack;
} orc_counters;

typedef struct {
    float base[3];
    float axis[3]; /* already normalised (cylinder.cpp:224) */
    float radius_sq;
    float height;
} orc_cylinder;

typedef struct {
    /* kd-tree over triangle lanes: kdtree.h:16-48 nodes (2 x u32), triangle.h:33-44 lanes (72 floats) */
    const uint32_t *nodes;
    uint32_t num_nodes;
    const float *tri_lanes;
    uint32_t num_tri_lanes;
    float bounds[6]; /* min xyz, max xyz */
    /* sphere.cpp:12-19 lanes: x[8] y[8] z[8] radiusSq[8] */
    const float *sphere_lanes;
    uint32_t num_spheres;
    /* plane.cpp:13-20 lanes: px[8] py[8] pz[8] nx[8] ny[8] nz[8] */
    const float *plane_lanes;
    uint32_t num_planes;
    const orc_cylinder *cylinders;
    uint32_t num_cylinders;
    /* EXTENSION (no reference counterpart): box lanes minx[8] miny[8] minz[8] maxx[8] maxy[8] maxz[8] */
    const float *box_lanes;
    uint32_t num_boxes;
    float epsilon; /* Config::Epsilon, config.h:9 */
} orc_scene;

/* single-primitive-class queries; *clip is in/out exactly like _Intersect::clippingDistance */
int orc_bounds_slab(const float bounds[6], const float o[3], const float inv[3], float clip, float *tmin, float *tmax);
int orc_triangles_in_range(const float *tri_lanes, uint32_t lane_start, uint32_t num_lanes, const float o[3],
                           const float d[3], float clip, orc_hit *hit);
int orc_kdtree_intersect(const orc_scene *s, const float o[3], const float d[3], int any, float *clip, orc_hit *hit,
                         orc_counters *ctr);
int orc_sphere_intersect(const orc_scene *s, const float o[3], const float d[3], int any, float clip, orc_hit *hit);
int orc_plane_intersect(const orc_scene *s, const float o[3], const float d[3], float clip, orc_hit *hit);
int orc_cylinder_intersect(const orc_scene *s, const float o[3], const float d[3], float clip, orc_hit *hit);
int orc_box_intersect(const orc_scene *s, const float o[3], const float d[3], int any, float clip, orc_hit *hit);

/* the query chain of main.cpp:312-321 (closest) / main.cpp:198-217 (any) over the enabled classes */
int orc_query(const orc_scene *s, const orc_ray *ray, uint32_t classes, orc_hit *hit, orc_counters *ctr);
void orc_intersect(const orc_scene *s, const orc_ray *rays, uint64_t n, uint32_t classes, orc_hit *hits,
                   orc_counters *ctrs /* may be NULL */, int nthreads);

/* ray construction */
void orc_ray_tables(uint32_t width, uint32_t height, float *xs, float *ys);
void orc_primary_ray(const float *xs, const float *ys, uint32_t row, uint32_t col, orc_ray *out);
void orc_primary_rays(uint32_t width, uint32_t height, orc_ray *out);
void orc_hit_point(const float o[3], const float d[3], float t, float p[3]);
void orc_shadow_ray(const float p[3], const float light[3], orc_ray *out);

/* fused frame passes used as the checker for dodrt_trace_primary / dodrt_trace_shadow */
void orc_trace_primary(const orc_scene *s, uint32_t width, uint32_t height, uint32_t classes, orc_hit *hits,
                       orc_counters *ctrs, int nthreads);
void orc_trace_shadow(const orc_scene *s, uint32_t width, uint32_t height, uint32_t classes, const orc_hit *hits,
                      const float light[3], uint8_t *visible, orc_counters *ctrs, int nthreads);

#ifdef __cplusplus
}
#endif
#endif

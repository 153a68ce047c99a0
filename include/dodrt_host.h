/* dodrt_host.h -- C ABI of the host side above the ray-query path (libdodrt_host.so, plain C++17, no CUDA).
 *
 * The reference's host side "stays" (BASELINE.json north_star): config.ini parsing, mesh loading,
 * scene registration and its own SAH kd-tree builder.  A reference maintainer keeps using those
 * and only adds the adapter of INTEGRATION.md.  This library is the same host side for everyone
 * else -- benchmarks, tests, Python -- written from scratch but reproducing the reference's results
 * bit for bit (tests/test_host_vs_ref.py compares every array with the reference's own output):
 *
 *   Triangle::create / lanes of 8, zero padded            triangle.cpp:262-292, triangle.h:33-44
 *   KDTree::buildTree  (SAH over lane boxes, quirks kept)  kdtree.cpp:66-260
 *   Triangle::reorderLanesByIndices                        triangle.cpp:349-367
 *   Sphere/Plane::create lanes, Cylinder::Cylinder         sphere.cpp:226-242, plane.cpp:204-222, cylinder.cpp:223-229
 *   generateSpheres / generatePlanes / generateCylinders   main.cpp:26-129
 *   raster tables of rayTrace                              main.cpp:275-279,342-345
 *   Config::Load                                           config.h:16-37, config_loader.h:26-71
 *
 * It produces arrays in exactly the layouts include/dodrt.h consumes; it never traces a ray.
 */
#ifndef DODRT_HOST_H
#define DODRT_HOST_H

#include <stdint.h>

#include "dodrt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dodrt_host_scene dodrt_host_scene;

/* Config, config.h:4-14 (defaults are the reference's) */
typedef struct dodrt_host_config {
    uint32_t height, width;  /* 1080, 1920 */
    float epsilon;           /* 1e-4 */
    float frustrum_max;      /* 1000 (unused by the reference too) */
    uint32_t intersect_cost; /* 80 */
    uint32_t traversal_cost; /* 80 */
    float empty_bonus;       /* 0 */
    uint32_t max_prims;      /* 8 */
} dodrt_host_config;

typedef struct dodrt_host_sizes {
    uint32_t num_triangles;   /* Triangle::m_numTriangles */
    uint32_t num_orig_lanes;  /* lanes before the re-order */
    uint32_t num_nodes;       /* KDTree::m_nodes.size() */
    uint32_t num_lanes;       /* lanes after the re-order == m_primNums.size() */
    uint32_t max_depth;       /* KDTree::m_maxDepth */
    uint32_t num_spheres, num_planes, num_cylinders, num_boxes;
} dodrt_host_sizes;

DODRT_API const char *dodrt_host_last_error(void);

DODRT_API void dodrt_host_config_defaults(dodrt_host_config *cfg);
/* `key: value` lines, whitespace stripped, unknown keys ignored, missing keys keep defaults */
DODRT_API int dodrt_host_config_load(const char *path, dodrt_host_config *cfg);

DODRT_API int dodrt_host_scene_create(const dodrt_host_config *cfg /* NULL = defaults */, dodrt_host_scene **scene);
DODRT_API void dodrt_host_scene_destroy(dodrt_host_scene *scene);

/* ---- meshes -> triangles (Mesh::Create, mesh.cpp:9-50) ----------------------------------------------
 * Indexed triangle mesh; smooth per-position normals are generated (the loader law is documented in
 * DESIGN.md) and each face is pushed through the Triangle::create equivalent in face order.
 * `transform` (optional) = uniform scale + translation applied to the positions first: {s, tx, ty, tz}. */
DODRT_API int dodrt_host_add_mesh(dodrt_host_scene *scene, const float *positions, uint32_t num_vertices,
                                  const uint32_t *indices, uint32_t num_triangles, const float transform[4]);
/* Wavefront OBJ (v / f, fan triangulation) or the "DODM" binary mesh */
DODRT_API int dodrt_host_add_mesh_file(dodrt_host_scene *scene, const char *path, const float transform[4]);
/* Deterministic stand-in for the reference's missing assets/dragon.obj (.MISSING_LARGE_BLOBS): a displaced
 * UV sphere, r(u,v) = 2.2 + 0.25 sin7u sin5v + 0.08 sin(31u+3) sin29v, u in [0,2pi], v in [0.02,pi-0.02],
 * (n+1)^2 vertices, 2 n^2 triangles (n = 660 -> 871,200).  Writes the mesh into caller arrays. */
DODRT_API int dodrt_host_standin_dragon(uint32_t n, float *positions /* (n+1)^2*3 */, uint32_t *indices /* 2n^2*3 */);
DODRT_API int dodrt_host_write_dodm(const char *path, const float *positions, uint32_t num_vertices,
                                    const uint32_t *indices, uint32_t num_triangles);

/* ---- analytic shapes --------------------------------------------------------------------------------- */
DODRT_API int dodrt_host_add_sphere(dodrt_host_scene *scene, const float pos[3], float radius, const float color[3]);
DODRT_API int dodrt_host_add_plane(dodrt_host_scene *scene, const float normal[3], const float pos[3],
                                   const float color[3]);
DODRT_API int dodrt_host_add_cylinder(dodrt_host_scene *scene, float radius, float height, const float axis[3],
                                      const float base[3]);
DODRT_API int dodrt_host_add_box(dodrt_host_scene *scene, const float lo[3], const float hi[3]);
/* the reference's scene: srand(seed); generateSpheres(16); generatePlanes(); generateCylinders() (main.cpp:364-366) */
DODRT_API int dodrt_host_add_reference_scene(dodrt_host_scene *scene, uint32_t seed, uint32_t num_spheres);
/* BASELINE.json config 4: `count` spheres and `count` boxes, centres U[-4.5,4.5]^3, radius/half-extent
 * U[0.03,0.12] from the LCG x <- 1664525x + 1013904223, u = (x>>8)/2^24, 8 draws per index */
DODRT_API int dodrt_host_add_analytic_scene(dodrt_host_scene *scene, uint32_t seed, uint32_t count);

/* ---- kd-tree ------------------------------------------------------------------------------------------ */
/* KDTree::buildTree() (kdtree.cpp:252-260): SAH build + Triangle::reorderLanesByIndices.  The build is
 * task-parallel over subtrees (env DODRT_HOST_THREADS, default = all cores) and returns the reference's tree bit for
 * bit for every thread count (see TreeBuilder in dodrt_host.cpp). */
DODRT_API int dodrt_host_build_tree(dodrt_host_scene *scene);
/* flags: DODRT_HOST_BUILD_KEEP_CREATION_ORDER skips the host-side lane re-order (triangle.cpp:349-367, a gather that
 * multiplies the lane data by the tree's duplication factor): dodrt_host_tri_lanes / _tri_normals / _tri_attributes
 * then stay in CREATION order (num_orig_lanes records) and the re-order runs on the GPU at upload
 * (dodrt_scene_set_kdtree_indexed / dodrt_scene_set_shading_indexed with dodrt_host_prim_nums). */
#define DODRT_HOST_BUILD_KEEP_CREATION_ORDER 1u
DODRT_API int dodrt_host_build_tree_ex(dodrt_host_scene *scene, uint32_t flags);

/* ---- export: pointers stay valid until the scene is modified or destroyed ---------------------------- */
DODRT_API int dodrt_host_sizes_get(const dodrt_host_scene *scene, dodrt_host_sizes *sizes);
DODRT_API const uint64_t *dodrt_host_nodes(const dodrt_host_scene *scene);
DODRT_API const float *dodrt_host_tri_lanes(const dodrt_host_scene *scene);    /* re-ordered (see build flags), 72 floats each */
DODRT_API const uint32_t *dodrt_host_prim_nums(const dodrt_host_scene *scene); /* original lane of each lane */
DODRT_API const float *dodrt_host_bounds(const dodrt_host_scene *scene);       /* 6 floats */
DODRT_API const float *dodrt_host_tri_normals(const dodrt_host_scene *scene);  /* re-ordered, 9 floats / slot */
/* shading attributes: Triangle::m_triangleAttributes after the re-order, byte for byte (triangle.h:45-51: 320 B per
 * lane = meshAttrIdx[8], AN[8] BN[8] CN[8] as vec3), and Mesh::m_meshAttributes (one colour per mesh, mesh.cpp:23) */
DODRT_API const void *dodrt_host_tri_attributes(const dodrt_host_scene *scene);
DODRT_API const float *dodrt_host_mesh_colors(const dodrt_host_scene *scene);
DODRT_API uint32_t dodrt_host_num_meshes(const dodrt_host_scene *scene);
DODRT_API const float *dodrt_host_sphere_lanes(const dodrt_host_scene *scene);
DODRT_API const float *dodrt_host_sphere_colors(const dodrt_host_scene *scene); /* 3 floats / sphere */
DODRT_API const float *dodrt_host_plane_lanes(const dodrt_host_scene *scene);
DODRT_API const float *dodrt_host_plane_colors(const dodrt_host_scene *scene);
DODRT_API const dodrt_cylinder *dodrt_host_cylinders(const dodrt_host_scene *scene);
DODRT_API const float *dodrt_host_box_lanes(const dodrt_host_scene *scene);
DODRT_API float dodrt_host_epsilon(const dodrt_host_scene *scene);

/* raster tables: xs[0] = -W/H, xs[j+1] = xs[j] + 2(W/H)/W ; ys[0] = 1, ys[i+1] = ys[i] - 2/H (fp32 accumulation) */
DODRT_API int dodrt_host_ray_tables(uint32_t width, uint32_t height, float *xs, float *ys);

#ifdef __cplusplus
}
#endif
#endif /* DODRT_HOST_H */

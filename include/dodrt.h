/* dodrt.h -- C ABI of the B200-native ray-query path (libdodrt_cuda.so).
 *
 * This is the drop-in boundary for the hot path of AVassilev98/dod_raytracer: kd-tree traversal
 * (src/accelerators/kdtree.cpp) driving ray-triangle, ray-sphere and ray-box intersection
 * (src/shapes) for primary and shadow rays.  The reference has no process, device or FFI boundary
 * (it is one C++ executable); the entry points below are the batched mirror of the C++ interface
 * the reference's render loop calls, and each one cites the reference interface it replaces.
 * Plain pointers and sizes only; no C++ or torch types.  INTEGRATION.md shows the adapter a
 * reference maintainer adds on the C++ side.
 *
 * Conventions
 *   - every function returns 0 on success, a negative DODRT_E_* code otherwise, and never throws;
 *     dodrt_last_error() returns a thread-local description of the last failure.
 *   - the reference's bool hit/miss (base_shape.h:23) is carried in dodrt_hit.prim: DODRT_MISS = miss.
 *   - arithmetic is fp32, round-to-nearest, un-fused, IEEE div/sqrt, in the reference's operation
 *     order, so t/u/v are bit-identical to the reference CPU path and prim ids are identical.
 *   - there is NO CPU fallback: without a usable CUDA device every call fails with DODRT_E_CUDA.
 *   - a scene is immutable while queries run on it (like the reference: main.cpp:364-368 then
 *     main.cpp:371-394); queries on one scene may be issued from several host threads.
 */
#ifndef DODRT_H
#define DODRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DODRT_ABI_VERSION 2

#if defined(__GNUC__)
#define DODRT_API __attribute__((visibility("default")))
#else
#define DODRT_API
#endif

/* status codes */
#define DODRT_OK 0
#define DODRT_E_INVALID (-1) /* bad argument */
#define DODRT_E_CUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define DODRT_E_LIMIT (-3)   /* scene exceeds a compiled-in limit (e.g. kd-tree depth) */
#define DODRT_E_NOMEM (-4)

/* primitive id encoding in dodrt_hit.prim: kind << 29 | id */
#define DODRT_MISS 0xFFFFFFFFu
#define DODRT_KIND_SHIFT 29
#define DODRT_KIND_TRIANGLE 0u /* id = (laneIdx * 8 + slot) in RE-ORDERED lane space, triangle.cpp:136 */
#define DODRT_KIND_SPHERE 1u   /* id = creation index, sphere.cpp:137 */
#define DODRT_KIND_PLANE 2u    /* id = creation index, plane.cpp:110 */
#define DODRT_KIND_CYLINDER 3u /* id = creation index, cylinder.cpp:180 */
#define DODRT_KIND_BOX 4u      /* extension, see dodrt_scene_set_boxes */

/* shape classes a query visits, in the reference's fixed order (main.cpp:314-321):
 * Sphere -> [Box] -> Plane -> Cylinder -> KDTree, each clipped by the running record.t */
#define DODRT_CLS_SPHERE 1u
#define DODRT_CLS_PLANE 2u
#define DODRT_CLS_CYLINDER 4u
#define DODRT_CLS_TREE 8u
#define DODRT_CLS_BOX 16u
#define DODRT_CLS_ALL 31u

#define DODRT_RAY_ANY 1u  /* _Intersect::returnOnAny, base_shape.h:12 */
#define DODRT_RAY_SKIP 2u /* inactive slot of a batch: answered with DODRT_MISS without any work */

/* One query = the reference's `_Intersect` (base_shape.h:8-15) without the HitRecord reference. */
typedef struct dodrt_ray {
    float o[3];     /* rayOrigin */
    float d[3];     /* rayDir (the reference passes normalised directions; not required) */
    float clip;     /* clippingDistance; +inf for an unclipped closest-hit query */
    uint32_t flags; /* DODRT_RAY_ANY */
} dodrt_ray;

/* The part of the reference's HitRecord (hitrecord.h:4-10) that the intersect path produces, plus
 * the primitive id the reference keeps in a local (triangle.cpp:37, sphere.cpp:29).  hitPoint is
 * o + d*t and hitNormal/color follow from (prim,u,v); they are rebuilt by the shading stage.
 * For a DODRT_RAY_ANY query only hit/miss is defined: prim is 0 on a hit, DODRT_MISS otherwise. */
typedef struct dodrt_hit {
    float t;       /* record.t; the query's clip on a miss */
    uint32_t prim; /* kind << 29 | id, or DODRT_MISS */
    float u, v;    /* triangle barycentrics (triangle.cpp:137: bary = (1-(u+v), u, v)); 0 otherwise */
} dodrt_hit;

/* Cylinder as the reference stores it after construction (cylinder.h:33-36, cylinder.cpp:223-229). */
typedef struct dodrt_cylinder {
    float base[3];
    float axis[3]; /* normalised */
    float radius_sq;
    float height;
} dodrt_cylinder;

/* Frame description for the fused passes.  Pixels are grouped into tiles of tile_w x tile_h
 * (multiples of 8 and 4); tile k of the row-major tile grid belongs to this call iff
 * k >= first_tile and (k - first_tile) % tile_stride == 0.  One GPU: first_tile 0, tile_stride 1.
 * N GPUs (image-tile split, scene replicated): rank r passes first_tile = r, tile_stride = N.
 * With compact = 0 results are indexed by pixel (row * width + col) in a full-frame buffer (pixels
 * of other ranks' tiles are left untouched); with compact = 1 they are packed in this call's own
 * order: slot = local_tile * tile_w * tile_h + in-tile index (see dodrt_frame_local_pixels and
 * dodrt_frame_pixel_map), padded slots of partial edge tiles hold DODRT_MISS / 0. */
typedef struct dodrt_frame {
    uint32_t width, height;
    uint32_t tile_w, tile_h;
    uint32_t first_tile, tile_stride;
    uint32_t classes; /* DODRT_CLS_* mask */
    uint32_t compact;
    float origin[3]; /* camera position, main.cpp:275 = (0,0,-4.9) */
} dodrt_frame;

typedef struct dodrt_scene dodrt_scene; /* opaque; owns one replica of the scene in one GPU's HBM */

/* ---- library ------------------------------------------------------------------------------- */
DODRT_API int dodrt_abi_version(void);
DODRT_API const char *dodrt_last_error(void);
DODRT_API int dodrt_device_count(int *count);

/* ---- scene registration ---------------------------------------------------------------------
 * Replaces the reference's process-global registration (Sphere::create sphere.h:27, Plane::create
 * plane.h:25, Cylinder::create cylinder.h:29, Mesh::Create mesh.h:24 -> Triangle::create
 * triangle.h:56) followed by KDTree::buildTree() (kdtree.h:12).  The host side keeps building the
 * scene exactly as the reference does; these calls copy the finished arrays, in the reference's own
 * layouts, into the GPU (the caller keeps ownership of the host memory). */
DODRT_API int dodrt_scene_create(int device, dodrt_scene **scene);
DODRT_API int dodrt_scene_destroy(dodrt_scene *scene);

/* nodes: KDTree::m_nodes (kdtree.h:16-48,63), 8 bytes each, DFS pre-order, left child = node+1.
 * tri_lanes: Triangle::m_triangleLanes AFTER Triangle::reorderLanesByIndices (triangle.h:33-44,
 * triangle.cpp:349-367), 288 bytes each: Ax[8] Ay[8] Az[8] Bx[8] .. Cz[8]; padding slots all-zero.
 * bounds: KDTree::m_bounds (kdtree.h:67), min xyz then max xyz. */
DODRT_API int dodrt_scene_set_kdtree(dodrt_scene *scene, const uint64_t *nodes, uint32_t num_nodes, const float *tri_lanes,
                           uint32_t num_tri_lanes, const float bounds[6]);
/* Same tree, but `tri_lanes` is Triangle::m_triangleLanes BEFORE the re-order (num_src_lanes lanes in creation
 * order) and `prim_nums` is KDTree::m_primNums (kdtree.h:64; num_tri_lanes entries): lane i of the tree is
 * tri_lanes[prim_nums[i]].  Replaces Triangle::reorderLanesByIndices (triangle.cpp:349-367, called from
 * KDTree::buildTree kdtree.cpp:258): the gather runs on the GPU, fused into the upload's AB/AC pre-computation, and the
 * host ships num_src_lanes instead of num_tri_lanes (= 3.2-3.8x more) lanes. */
DODRT_API int dodrt_scene_set_kdtree_indexed(dodrt_scene *scene, const uint64_t *nodes, uint32_t num_nodes,
                                             const float *tri_lanes, uint32_t num_src_lanes, const uint32_t *prim_nums,
                                             uint32_t num_tri_lanes, const float bounds[6]);
/* sphere lanes as sphere.cpp:12-19: x[8] y[8] z[8] radiusSq[8], 128 bytes per lane, ceil(n/8) lanes */
DODRT_API int dodrt_scene_set_spheres(dodrt_scene *scene, const float *sphere_lanes, uint32_t num_spheres);
/* plane lanes as plane.cpp:13-20: px[8] py[8] pz[8] nx[8] ny[8] nz[8]; epsilon = Config::Epsilon */
DODRT_API int dodrt_scene_set_planes(dodrt_scene *scene, const float *plane_lanes, uint32_t num_planes);
DODRT_API int dodrt_scene_set_cylinders(dodrt_scene *scene, const dodrt_cylinder *cylinders, uint32_t num_cylinders);
/* EXTENSION (BASELINE.json config 4; the reference's box.h is only the kd-tree bounds helper):
 * renderable axis-aligned boxes, lanes minx[8] miny[8] minz[8] maxx[8] maxy[8] maxz[8], tested with
 * the slab arithmetic of AxisAlignedBoundingBox::intersect (box.cpp:33-53); hit distance = entry
 * distance, accepted when it is > 0 and < the running record.t. */
DODRT_API int dodrt_scene_set_boxes(dodrt_scene *scene, const float *box_lanes, uint32_t num_boxes);
/* Config::Epsilon (config.h:9), used by the plane and cylinder tests; default 1e-4 */
DODRT_API int dodrt_scene_set_epsilon(dodrt_scene *scene, float epsilon);

/* Shading attributes for dodrt_render*: Triangle::m_triangleAttributes after the re-order, in the reference's own
 * layout (triangle.h:45-51: 320 bytes per lane = unsigned meshAttrIdx[8], vec3 AN[8], vec3 BN[8], vec3 CN[8]),
 * Mesh::m_meshAttributes colours (mesh.h:13-16, 3 floats per mesh), and the colours the spheres / planes were
 * created with (sphere.h:14-17, plane.h:14-17; 3 floats each, creation order).  Cylinders render black like in the
 * reference (cylinder.cpp:172-179,204). */
DODRT_API int dodrt_scene_set_shading(dodrt_scene *scene, const void *tri_attributes, uint32_t num_tri_lanes,
                                      const float *mesh_colors, uint32_t num_meshes, const float *sphere_colors,
                                      const float *plane_colors);
/* ... with Triangle::m_triangleAttributes BEFORE the re-order + KDTree::m_primNums (see dodrt_scene_set_kdtree_indexed) */
DODRT_API int dodrt_scene_set_shading_indexed(dodrt_scene *scene, const void *tri_attributes, uint32_t num_src_lanes,
                                              const uint32_t *prim_nums, uint32_t num_tri_lanes, const float *mesh_colors,
                                              uint32_t num_meshes, const float *sphere_colors, const float *plane_colors);

/* Tuning / A-B knob: which traversal kernel variant answers the queries (all variants return identical
 * results; see dod_raytracer_b200/csrc/dodrt_kernels.cu).  variant < 0 restores the default; a variant this build
 * does not hold is refused with DODRT_E_INVALID. */
DODRT_API int dodrt_scene_set_kernel_variant(dodrt_scene *scene, int variant);
/* 1 if this build of the library holds `variant` (the product build: 0, 3, 7 and "auto" < 0; the -DDODRT_EXPERIMENTS build
 * libdodrt_cuda_exp.so: all of them, plus the one-launch frame kernels and work splitting in the donation queue) */
DODRT_API int dodrt_kernel_variant_available(int variant);

/* ---- queries: host buffers (copies in and out are part of the call) --------------------------
 * dodrt_intersect is the batch form of `bool KDTree::intersect(_Intersect&) const` (kdtree.h:13) and
 * `static bool BaseShape<D>::intersect(_Intersect&)` (base_shape.h:23) chained as in
 * main.cpp:314-321 (closest) / main.cpp:198-217 (DODRT_RAY_ANY): hits[i] answers rays[i]. */
DODRT_API int dodrt_intersect(dodrt_scene *scene, const dodrt_ray *rays, uint64_t num_rays, uint32_t classes, dodrt_hit *hits);

/* dodrt_trace_primary replaces the primary-ray half of rayTrace (main.cpp:273-321 with k = 0):
 * dir = normalize((xs[col], ys[row], 1)) (main.cpp:304; xs/ys are the accumulated raster tables of
 * main.cpp:276-279,342-345, width resp. height floats), closest-hit chain from frame->origin. */
DODRT_API int dodrt_trace_primary(dodrt_scene *scene, const dodrt_frame *frame, const float *xs, const float *ys,
                        dodrt_hit *hits);
/* dodrt_trace_shadow replaces canSeeLight (main.cpp:182-219) for every pixel with a primary hit:
 * visible[i] = 1 iff hits[i] is a hit and no shape blocks hitPoint -> light; 0 otherwise. */
DODRT_API int dodrt_trace_shadow(dodrt_scene *scene, const dodrt_frame *frame, const float *xs, const float *ys,
                       const dodrt_hit *hits, const float light[3], uint8_t *visible);
/* primary + one shadow pass per light without the intermediate round trip through the host;
 * visible is [num_lights][slots]. */
DODRT_API int dodrt_trace_frame(dodrt_scene *scene, const dodrt_frame *frame, const float *xs, const float *ys,
                      const float *lights /* num_lights x 3 */, uint32_t num_lights, dodrt_hit *hits,
                      uint8_t *visible);

/* dodrt_render replaces rayTrace (main.cpp:273-347) for the whole frame: per pixel up to `depth` mirror bounces
 * (the reference hard-codes 10, main.cpp:301), each = closest-hit chain + one canSeeLight query per light + the
 * reference's shading (main.cpp:156-244), blended with weight 1/2^k; rgb is width*height*3 bytes, row-major
 * (what the reference hands to stbi_write_png, main.cpp:396).  lights: num_lights x {x, y, z, intensity}
 * (light.h:4-8; the reference's nine are main.cpp:283-292), num_lights <= 16.  frame->classes selects the shape
 * classes.  The call renders ITS tiles (first_tile / tile_stride, like the reference's threads render their row bands,
 * main.cpp:371-394): with compact = 0 rgb is the whole frame (pixels of other calls' tiles black), with compact = 1 it
 * holds slots x 3 bytes in the call's own order (dodrt_frame_pixel_map).  dodrt_multi_render renders one frame with all
 * GPUs of a dodrt_multi into ONE row-major rgb buffer. */
DODRT_API int dodrt_render(dodrt_scene *scene, const dodrt_frame *frame, const float *xs, const float *ys,
                           const float *lights, uint32_t num_lights, uint32_t depth, uint8_t *rgb);

/* ---- queries: device-resident buffers (pointers in the scene's GPU memory, asynchronous on
 * `stream`, a cudaStream_t passed as void*; NULL = the default stream).  Used when the caller keeps
 * rays/hits on the GPU (next pipeline stage, NCCL gather, benchmarks). */
DODRT_API int dodrt_intersect_device(dodrt_scene *scene, const dodrt_ray *d_rays, uint64_t num_rays, uint32_t classes,
                           dodrt_hit *d_hits, void *stream);
DODRT_API int dodrt_trace_primary_device(dodrt_scene *scene, const dodrt_frame *frame, const float *d_xs, const float *d_ys,
                               dodrt_hit *d_hits, void *stream);
DODRT_API int dodrt_trace_shadow_device(dodrt_scene *scene, const dodrt_frame *frame, const float *d_xs, const float *d_ys,
                              const dodrt_hit *d_hits, const float light[3], uint8_t *d_visible, void *stream);

/* ---- one launch per frame share, results written where they are needed ------------------------------
 * dodrt_trace_frame_device = primary rays + one canSeeLight query per light and hit pixel (rayTrace at k = 0,
 * main.cpp:297-334 without the shading) for this call's tiles, as ONE persistent kernel with two work queues: a
 * tile's shadow batches (still a separate, coherent pass over 8x4 pixel blocks) become claimable as soon as that
 * tile's primary hit records are complete, so the tail of the primary queue overlaps shadow work.  d_hits /
 * d_visible ([num_lights][slots]) are laid out as frame->compact says; lights is a HOST array of num_lights x 3.
 * `mirror` (optional): a frame buffer view for this scene's GPU (dodrt_frame_buffer_*); every result is ALSO written
 * into that row-major frame by the kernel itself -- with an image-tile split every rank passes a view of rank 0's
 * frame buffer and the frame assembles itself over NVLink while it is traced: no gather, no assembly pass.
 * Replaces the reference's own split of the frame over its threads (main.cpp:371-394) + the shared image buffer. */
typedef struct dodrt_frame_buffer dodrt_frame_buffer; /* opaque */
DODRT_API int dodrt_trace_frame_device(dodrt_scene *scene, const dodrt_frame *frame, const float *d_xs, const float *d_ys,
                                       const float *lights, uint32_t num_lights, dodrt_hit *d_hits, uint8_t *d_visible,
                                       dodrt_frame_buffer *mirror, void *stream);

/* A row-major frame -- hit records [height*width], then visibility [num_lights][height*width] -- in the HBM of
 * `owner`'s GPU that kernels running on OTHER GPUs write into directly (NVLink / NVSwitch peer stores): the analogue of
 * the one `imageData` block all of the reference's threads write (main.cpp:369,336-340).
 *   same process : dodrt_frame_buffer_attach(scene_on_other_gpu, fb, &view)   (enables peer access)
 *   other process: dodrt_frame_buffer_export(fb, &desc) -> ship the 96-byte descriptor (MPI, torch.distributed, a pipe)
 *                  -> dodrt_frame_buffer_open(scene_on_other_gpu, &desc, &view)   (CUDA IPC, same node)
 * The owner passes `fb` itself as its own mirror.  Initial contents: DODRT_MISS / 0.  Destroy views before the owner's
 * buffer; a buffer outlives neither its owner's process nor (for views) the scene it was opened for. */
typedef struct dodrt_frame_buffer_desc {
    uint32_t width, height, num_lights, device;
    uint64_t bytes;
    uint8_t ipc_handle[64];
    uint8_t reserved[8];
} dodrt_frame_buffer_desc;
DODRT_API int dodrt_frame_buffer_create(dodrt_scene *owner, uint32_t width, uint32_t height, uint32_t num_lights,
                                        dodrt_frame_buffer **fb);
DODRT_API int dodrt_frame_buffer_export(dodrt_frame_buffer *fb, dodrt_frame_buffer_desc *desc);
DODRT_API int dodrt_frame_buffer_open(dodrt_scene *user, const dodrt_frame_buffer_desc *desc, dodrt_frame_buffer **view);
DODRT_API int dodrt_frame_buffer_attach(dodrt_scene *user, dodrt_frame_buffer *owner_fb, dodrt_frame_buffer **view);
DODRT_API int dodrt_frame_buffer_pointers(dodrt_frame_buffer *fb, dodrt_hit **d_hits, uint8_t **d_visible);
DODRT_API int dodrt_frame_buffer_destroy(dodrt_frame_buffer *fb);

/* ---- several GPUs, one host process ---------------------------------------------------------------------------
 * The reference renders one frame with all the threads of its process (main.cpp:371-394: one row band per core);
 * dodrt_multi renders one frame with all the GPUs of the process: scene replicated (one populated dodrt_scene per GPU,
 * created and filled by the caller with the same arrays), image tiles dealt round-robin over the GPUs, and ONE frame
 * buffer in host memory that every GPU fills in place.  frame describes the whole frame (first_tile / tile_stride /
 * compact, and for dodrt_multi_trace_frame the tile size, are ignored); hits is [height*width], visible
 * [num_lights][height*width], both row-major.  With pinned host buffers (cudaHostAlloc / cudaHostRegister) the frame is
 * dealt out in bands of whole pixel rows and every GPU copies its finished bands straight to their place in the
 * caller's frame over its own PCIe link, the hit records while its shadow pass still runs (DODRT_ZEROCOPY=1: the
 * kernels store into the mapped host frame themselves instead -- measured slower); with pageable buffers the GPUs
 * assemble the frame in scenes[0]'s HBM over NVLink first and one copy brings it to the host.  Calls on one dodrt_multi
 * are serialised. */
typedef struct dodrt_multi dodrt_multi; /* opaque */
DODRT_API int dodrt_multi_create(dodrt_scene *const *scenes, uint32_t num_scenes, dodrt_multi **multi);
DODRT_API int dodrt_multi_trace_frame(dodrt_multi *multi, const dodrt_frame *frame, const float *xs, const float *ys,
                                      const float *lights, uint32_t num_lights, dodrt_hit *hits, uint8_t *visible);
DODRT_API int dodrt_multi_render(dodrt_multi *multi, const dodrt_frame *frame, const float *xs, const float *ys,
                                 const float *lights, uint32_t num_lights, uint32_t depth, uint8_t *rgb);
DODRT_API int dodrt_multi_destroy(dodrt_multi *multi);

/* Multi-GPU frame assembly (kept for callers that gather compact blocks themselves, e.g. with NCCL; the frame buffer
 * mirror above makes it unnecessary).  After an image-tile split over N ranks (frame->tile_stride = N, rank r traced
 * with first_tile = r, compact = 1) and a gather that places rank r's compact results at
 * d_compact_*[r * slots_per_rank ...], this writes the row-major full frame: hits_out[row*width+col] and,
 * when both visibility pointers are non-NULL, visible_out[row*width+col].  frame->first_tile is ignored. */
DODRT_API int dodrt_frame_assemble_device(dodrt_scene *scene, const dodrt_frame *frame, const dodrt_hit *d_compact_hits,
                                          const uint8_t *d_compact_visible, uint64_t slots_per_rank,
                                          dodrt_hit *d_hits_out, uint8_t *d_visible_out, void *stream);

/* ---- frame helpers (host only, no GPU work) -------------------------------------------------- */
/* number of result slots a call with this frame description writes when compact = 1 */
DODRT_API int dodrt_frame_local_pixels(const dodrt_frame *frame, uint64_t *slots);
/* pixel_of_slot[s] = row * width + col of compact slot s, or 0xFFFFFFFF for a padded slot */
DODRT_API int dodrt_frame_pixel_map(const dodrt_frame *frame, uint32_t *pixel_of_slot, uint64_t slots);

/* ---- instrumentation ------------------------------------------------------------------------- */
/* Counters of the sphere / box culling structure (measurement only; queries run a little slower while enabled):
 * reads the counters accumulated since the last call into out (may be NULL) -- [0]/[1] nodes fetched / spheres tested,
 * [2]/[3] nodes fetched / boxes tested, [4] rays that took a culling structure -- then clears them and switches the
 * counting on (enable != 0) or off.  Synchronises the device. */
DODRT_API int dodrt_scene_debug_stats(dodrt_scene *scene, int enable, uint64_t out[8]);
/* kernels launched by this library on behalf of this scene since creation */
DODRT_API int dodrt_scene_launch_count(dodrt_scene *scene, uint64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* DODRT_H */

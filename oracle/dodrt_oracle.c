/* TEST INFRASTRUCTURE -- CPU oracle for the dod_raytracer hot path.  NOT product code.
 *
 * A plain-C, one-primitive-at-a-time restatement of the reference's ray-query path
 * (AVassilev98/dod_raytracer, mounted at /root/reference).  It exists so that the CUDA path has a
 * checker that (a) travels to machines where the reference sources do not exist, (b) exposes what
 * the reference keeps in locals (primitive ids, barycentrics) and (c) counts the work the
 * reference traversal does per ray (nodes fetched, lanes tested) for the roofline's algorithmic bytes.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED against the reference's own translation units (oracle/_ref, see
 * dodrt_oracle.h).  Build: gcc -O2 -ffp-contract=off (oracle/Makefile) -- all arithmetic below is
 * fp32, round-to-nearest, un-fused, IEEE div/sqrt, denormals kept, and every expression keeps the
 * reference's association order.
 */
#include "dodrt_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define LANE 8            /* c_triangleLaneSz, triangle.h:32 ; c_sphereLaneSz, sphere.cpp:11 */
#define TRI_LANE_FLOATS 72 /* 9 x 8, triangle.h:33-44 */

/* avxDot, avx_utils.h:13-22 == glm::dot : (x1*x2 + y1*y2) + z1*z2 */
static inline float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    float px = ax * bx;
    float py = ay * by;
    float pz = az * bz;
    float acc = px + py;
    return acc + pz;
}

/* ------------------------------------------------------------------------------------------------
 * AxisAlignedBoundingBox::intersect, box.cpp:33-53.  Comparisons are the literal ones (NaN from
 * 0*inf falls through them exactly as in the reference); no fminf/fmaxf.
 * ---------------------------------------------------------------------------------------------- */
int orc_bounds_slab(const float bounds[6], const float o[3], const float inv[3], float clip, float *tminOut,
                    float *tmaxOut)
{
    float tmin = 0;
    float tmax = clip;
    for (int i = 0; i < 3; ++i) {
        float tNear = (bounds[i] - o[i]) * inv[i];
        float tFar = (bounds[3 + i] - o[i]) * inv[i];
        if (tNear > tFar) {
            float tmp = tNear;
            tNear = tFar;
            tFar = tmp;
        }
        tmin = tNear > tmin ? tNear : tmin;
        tmax = tFar < tmax ? tFar : tmax;
        if (tmin > tmax) {
            return 0;
        }
    }
    *tminOut = tmin;
    *tmaxOut = tmax;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * Triangle::intersectInRange, triangle.cpp:22-177, one triangle at a time.
 * The AVX code tests `t < maximumDistance-at-lane-start` (triangle.cpp:45,109-111) and then
 * `laneT[j] < maximumDistance` (running, triangle.cpp:133); the running value never exceeds the
 * lane-start value, so the conjunction is the single running strict `<` used here.  The per-lane
 * early `continue`s (triangle.cpp:76-79,90-93,103-106,114-117) only skip work for lanes in which
 * no triangle survives, so they do not change any per-triangle outcome.  The last-lane mask is
 * computed but never applied (triangle.cpp:28-34): padding triangles are all-zero and fail |det|>0.
 * id = (startIdx + i) * 8 + j in re-ordered lane space, triangle.cpp:136.
 * ---------------------------------------------------------------------------------------------- */
int orc_triangles_in_range(const float *tri_lanes, uint32_t lane_start, uint32_t num_lanes, const float o[3],
                           const float d[3], float clip, orc_hit *hit)
{
    float maximumDistance = clip;
    uint32_t minTriangleIndex = ORC_MISS;
    float bu = 0, bv = 0;
    for (uint32_t i = 0; i < num_lanes; i++) {
        const float *lane = tri_lanes + (size_t)(lane_start + i) * TRI_LANE_FLOATS;
        for (int j = 0; j < LANE; j++) {
            float Ax = lane[0 * LANE + j], Ay = lane[1 * LANE + j], Az = lane[2 * LANE + j];
            float Bx = lane[3 * LANE + j], By = lane[4 * LANE + j], Bz = lane[5 * LANE + j];
            float Cx = lane[6 * LANE + j], Cy = lane[7 * LANE + j], Cz = lane[8 * LANE + j];
            /* triangle.cpp:66-69 */
            float ABx = Bx - Ax, ABy = By - Ay, ABz = Bz - Az;
            float ACx = Cx - Ax, ACy = Cy - Ay, ACz = Cz - Az;
            float px = d[1] * ACz - d[2] * ACy; /* avxCross(rayDir, AC), avx_utils.h:24-33 */
            float py = d[2] * ACx - d[0] * ACz;
            float pz = d[0] * ACy - d[1] * ACx;
            float det = dot3(px, py, pz, ABx, ABy, ABz);
            if (!(fabsf(det) > 0.0f)) { /* triangle.cpp:70-73, _CMP_GT_OS: NaN rejects */
                continue;
            }
            float inv_det = 1.0f / det; /* triangle.cpp:81 */
            float tx = o[0] - Ax, ty = o[1] - Ay, tz = o[2] - Az;
            float u = dot3(tx, ty, tz, px, py, pz) * inv_det;
            if (!(u > 0.0f && u < 1.0f)) { /* triangle.cpp:85-87 */
                continue;
            }
            float qx = ty * ABz - tz * ABy; /* avxCross(tvec, AB) */
            float qy = tz * ABx - tx * ABz;
            float qz = tx * ABy - ty * ABx;
            float v = dot3(d[0], d[1], d[2], qx, qy, qz) * inv_det;
            if (!(v > 0.0f && (u + v) < 1.0f)) { /* triangle.cpp:98-100 */
                continue;
            }
            float t = dot3(ACx, ACy, ACz, qx, qy, qz) * inv_det;
            if (!(t > 0.0f && t < maximumDistance)) { /* triangle.cpp:109-111,133 */
                continue;
            }
            maximumDistance = t;
            minTriangleIndex = (lane_start + i) * LANE + (uint32_t)j;
            bu = u;
            bv = v;
        }
    }
    if (minTriangleIndex == ORC_MISS) {
        return 0;
    }
    hit->t = maximumDistance;
    hit->prim = ((uint32_t)ORC_KIND_TRIANGLE << ORC_KIND_SHIFT) | minTriangleIndex;
    hit->u = bu;
    hit->v = bv;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * KDTree::intersect, kdtree.cpp:263-361.  Node encoding kdtree.h:16-48: word0 bits[1:0] = axis or
 * 3 = leaf, bits[31:2] = rightChildIdx (interior) / numLanes (leaf); word1 = splitOffset (float) /
 * laneStartIdx.  Left child = node + 1 (kdtree.cpp:302).
 * ---------------------------------------------------------------------------------------------- */
int orc_kdtree_intersect(const orc_scene *s, const float o[3], const float d[3], int any, float *clipInOut,
                         orc_hit *hit, orc_counters *ctr)
{
    struct {
        uint32_t node;
        float tmin, tmax;
    } worklist[64];
    float clip = *clipInOut;
    float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]}; /* kdtree.cpp:271 */
    float tmin, tmax;
    if (s->num_nodes == 0) {
        return 0;
    }
    if (!orc_bounds_slab(s->bounds, o, inv, clip, &tmin, &tmax) || tmin > clip) { /* kdtree.cpp:274 */
        return 0;
    }
    int worklistPos = 0;
    uint32_t node = 0;
    int hitAny = 0;
    for (;;) {
        if (clip < tmin) { /* kdtree.cpp:286-289 */
            break;
        }
        uint32_t w0 = s->nodes[2 * node];
        uint32_t w1 = s->nodes[2 * node + 1];
        if (ctr) {
            ctr->nodes++;
        }
        if ((w0 & 3u) != 3u) {
            uint32_t axis = w0 & 3u;
            float split;
            memcpy(&split, &w1, 4);
            float tPlane = (split - o[axis]) * inv[axis]; /* kdtree.cpp:293 */
            int leftFirst = (o[axis] < split) || (o[axis] == split && d[axis] <= 0); /* kdtree.cpp:297-299 */
            uint32_t nearChild, farChild;
            if (leftFirst) {
                nearChild = node + 1;
                farChild = w0 >> 2;
            } else {
                nearChild = w0 >> 2;
                farChild = node + 1;
            }
            if (tPlane > tmax || tPlane <= 0) { /* kdtree.cpp:312 */
                node = nearChild;
            } else if (tPlane < tmin) { /* kdtree.cpp:316 */
                node = farChild;
            } else { /* kdtree.cpp:320-329 */
                worklist[worklistPos].node = farChild;
                worklist[worklistPos].tmin = tPlane;
                worklist[worklistPos].tmax = tmax;
                ++worklistPos;
                if (ctr && (uint32_t)worklistPos > ctr->max_stack) {
                    ctr->max_stack = (uint32_t)worklistPos;
                }
                node = nearChild;
                tmax = tPlane;
            }
        } else {
            uint32_t numLanes = w0 >> 2;
            uint32_t laneStart = w1;
            if (ctr) {
                ctr->leaves++;
                ctr->lanes += numLanes;
            }
            orc_hit h;
            if (orc_triangles_in_range(s->tri_lanes, laneStart, numLanes, o, d, clip, &h)) { /* kdtree.cpp:336 */
                *hit = h;
                if (any) { /* kdtree.cpp:338-341 */
                    *clipInOut = clip;
                    return 1;
                }
                hitAny = 1;
                clip = h.t; /* kdtree.cpp:343 */
            }
            if (worklistPos > 0) { /* kdtree.cpp:347-357 */
                --worklistPos;
                node = worklist[worklistPos].node;
                tmin = worklist[worklistPos].tmin;
                tmax = worklist[worklistPos].tmax;
            } else {
                break;
            }
        }
    }
    *clipInOut = clip;
    return hitAny;
}

/* ------------------------------------------------------------------------------------------------
 * Sphere::intersect_impl, sphere.cpp:26-160, one sphere at a time but lane-structured so that the
 * any-hit early-out (sphere.cpp:138-141: break after the first LANE that produced a hit) and the
 * last-lane mask (sphere.cpp:31-37,46-49) are the reference's.  record.t starts at clip (sphere.cpp:28).
 * ---------------------------------------------------------------------------------------------- */
int orc_sphere_intersect(const orc_scene *s, const float o[3], const float d[3], int any, float clip, orc_hit *hit)
{
    float recordT = clip;
    uint32_t closest = ORC_MISS;
    uint32_t numLanes = (s->num_spheres + LANE - 1) / LANE;
    for (uint32_t i = 0; i < numLanes; i++) {
        const float *lane = s->sphere_lanes + (size_t)i * 4 * LANE;
        uint32_t minIdx = 0;
        float minDist = recordT;
        for (uint32_t j = 0; j < LANE; j++) {
            if (i * LANE + j >= s->num_spheres) {
                break;
            }
            float lx = lane[0 * LANE + j] - o[0]; /* sphere.cpp:62-64 */
            float ly = lane[1 * LANE + j] - o[1];
            float lz = lane[2 * LANE + j] - o[2];
            float distSq = dot3(lx, ly, lz, lx, ly, lz);
            float radSq = lane[3 * LANE + j];
            if (!(distSq > radSq)) { /* sphere.cpp:70 : origin must be outside */
                continue;
            }
            float tca = dot3(lx, ly, lz, d[0], d[1], d[2]); /* sphere.cpp:83 */
            float tcaSq = tca * tca;
            float d2 = distSq - tcaSq;
            if (!(d2 < radSq)) { /* sphere.cpp:88 */
                continue;
            }
            float thcSq = radSq - d2;
            float thc = sqrtf(thcSq);
            float t0 = tca - thc;
            float t1 = tca + thc;
            if (!(t0 >= 0.0f && t1 >= 0.0f)) { /* sphere.cpp:103-106 */
                continue;
            }
            float tm = t0 < t1 ? t0 : t1; /* _mm256_min_ps(t0, t1): second operand unless t0 < t1 */
            if (tm < minDist) {           /* sphere.cpp:127-133 */
                minDist = tm;
                minIdx = j;
            }
        }
        if (minDist < recordT) { /* sphere.cpp:135-142 */
            recordT = minDist;
            closest = i * LANE + minIdx;
            if (any) {
                break;
            }
        }
    }
    if (closest == ORC_MISS) {
        return 0;
    }
    hit->t = recordT;
    hit->prim = ((uint32_t)ORC_KIND_SPHERE << ORC_KIND_SHIFT) | closest;
    hit->u = hit->v = 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * Plane::intersect_impl, plane.cpp:27-139.  The "last lane" mask is all-ones (plane.cpp:35 memsets
 * every slot), padding planes have a zero normal and fail |denom| > eps.  The mask used by the
 * scalar tail is taken BEFORE the clip compare (plane.cpp:92-100), which is harmless because the
 * tail re-checks `t < minT` with minT starting at clip (plane.cpp:38,107).  No any-hit early-out.
 * ---------------------------------------------------------------------------------------------- */
int orc_plane_intersect(const orc_scene *s, const float o[3], const float d[3], float clip, orc_hit *hit)
{
    float minT = clip;
    uint32_t closest = ORC_MISS;
    float eps = s->epsilon;
    uint32_t numLanes = (s->num_planes + LANE - 1) / LANE;
    for (uint32_t i = 0; i < numLanes; i++) {
        const float *lane = s->plane_lanes + (size_t)i * 6 * LANE;
        for (uint32_t j = 0; j < LANE; j++) {
            float px = lane[0 * LANE + j], py = lane[1 * LANE + j], pz = lane[2 * LANE + j];
            float nx = lane[3 * LANE + j], ny = lane[4 * LANE + j], nz = lane[5 * LANE + j];
            float denom = dot3(d[0], d[1], d[2], nx, ny, nz); /* plane.cpp:67 */
            if (!(fabsf(denom) > eps)) {                      /* plane.cpp:68-71 */
                continue;
            }
            float vx = px - o[0], vy = py - o[1], vz = pz - o[2];
            float num = dot3(vx, vy, vz, nx, ny, nz);
            float t = num / denom; /* plane.cpp:83 */
            if (!(t > eps)) {      /* plane.cpp:85 */
                continue;
            }
            if (t < minT) { /* plane.cpp:107 */
                minT = t;
                closest = i * LANE + j;
            }
        }
    }
    if (closest == ORC_MISS) {
        return 0;
    }
    hit->t = minT;
    hit->prim = ((uint32_t)ORC_KIND_PLANE << ORC_KIND_SHIFT) | closest;
    hit->u = hit->v = 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * Cylinder::intersect_non_vectorized, cylinder.cpp:155-210 with body (:76-118) and discs (:120-152).
 * cylinder.cpp:92-93 call the unqualified sqrt on a float: that is ::sqrt(double), so the two roots
 * are evaluated in double ((double)(-b) -/+ sqrt((double)disc)) / (double)(2*a) and rounded to float.
 * prim id = cylinder index (the reference does not distinguish body and discs in its result).
 * ---------------------------------------------------------------------------------------------- */
static float min_non_negative(float a, float b) /* cylinder.cpp:8-26 */
{
    if (a < 0 && b < 0) {
        return INFINITY;
    } else if (a < 0) {
        return b;
    } else if (b < 0) {
        return a;
    }
    return fminf(a, b);
}

static int cylinder_body(const orc_cylinder *c, const float o[3], const float d[3], float eps, float *tOut)
{
    float dpx = o[0] - c->base[0], dpy = o[1] - c->base[1], dpz = o[2] - c->base[2];
    float k = dot3(d[0], d[1], d[2], c->axis[0], c->axis[1], c->axis[2]);
    float vx = d[0] - k * c->axis[0], vy = d[1] - k * c->axis[1], vz = d[2] - k * c->axis[2];
    float m = dot3(dpx, dpy, dpz, c->axis[0], c->axis[1], c->axis[2]);
    float rx = dpx - m * c->axis[0], ry = dpy - m * c->axis[1], rz = dpz - m * c->axis[2];
    float a = dot3(vx, vy, vz, vx, vy, vz);
    float b = 2.0f * dot3(vx, vy, vz, rx, ry, rz);
    float cc = dot3(rx, ry, rz, rx, ry, rz) - c->radius_sq;
    float disc = (b * b) - (4 * a * cc);
    if (disc < eps) {
        return 0;
    }
    float tSub = (float)(((double)(-b) - sqrt((double)disc)) / (double)(2 * a));
    float tAdd = (float)(((double)(-b) + sqrt((double)disc)) / (double)(2 * a));
    float t = min_non_negative(tSub, tAdd);
    if (t == INFINITY) {
        return 0;
    }
    float cx = (o[0] + d[0] * t) - c->base[0];
    float cy = (o[1] + d[1] * t) - c->base[1];
    float cz = (o[2] + d[2] * t) - c->base[2];
    float f = dot3(cx, cy, cz, c->axis[0], c->axis[1], c->axis[2]);
    if (f < 0.f || f > c->height) {
        return 0;
    }
    *tOut = t;
    return 1;
}

static int cylinder_disc(const orc_cylinder *c, const float o[3], const float d[3], float eps, float offset,
                         float clip, float *tOut)
{
    float minT = clip; /* cylinder.cpp:122: the ORIGINAL clip, not the running one */
    float px = c->base[0] + c->axis[0] * offset;
    float py = c->base[1] + c->axis[1] * offset;
    float pz = c->base[2] + c->axis[2] * offset;
    float denom = dot3(d[0], d[1], d[2], c->axis[0], c->axis[1], c->axis[2]);
    if (fabsf(denom) < eps) {
        return 0;
    }
    float vx = px - o[0], vy = py - o[1], vz = pz - o[2];
    float tnum = dot3(vx, vy, vz, c->axis[0], c->axis[1], c->axis[2]);
    float t = tnum / denom;
    if (t < eps || t > minT) {
        return 0;
    }
    float hx = o[0] + d[0] * t, hy = o[1] + d[1] * t, hz = o[2] + d[2] * t;
    float wx = hx - px, wy = hy - py, wz = hz - pz;
    if (dot3(wx, wy, wz, wx, wy, wz) > c->radius_sq) {
        return 0;
    }
    *tOut = t;
    return 1;
}

int orc_cylinder_intersect(const orc_scene *s, const float o[3], const float d[3], float clip, orc_hit *hit)
{
    uint32_t minIdx = ORC_MISS;
    float tMin = clip;
    for (uint32_t i = 0; i < s->num_cylinders; i++) {
        const orc_cylinder *c = &s->cylinders[i];
        float t;
        if (cylinder_body(c, o, d, s->epsilon, &t) && t < tMin) {
            tMin = t;
            minIdx = i;
        }
        if (cylinder_disc(c, o, d, s->epsilon, 0.0f, clip, &t) && t < tMin) {
            tMin = t;
            minIdx = i;
        }
        if (cylinder_disc(c, o, d, s->epsilon, c->height, clip, &t) && t < tMin) {
            tMin = t;
            minIdx = i;
        }
    }
    if (minIdx == ORC_MISS) {
        return 0;
    }
    hit->t = tMin;
    hit->prim = ((uint32_t)ORC_KIND_CYLINDER << ORC_KIND_SHIFT) | minIdx;
    hit->u = hit->v = 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * EXTENSION -- axis-aligned boxes as a primitive class (BASELINE.json config 4).  The reference has
 * no renderable box (box.h is the kd-tree's bounds helper only), so this defines one with the
 * reference's own slab arithmetic (box.cpp:33-53): for box k in id order, run the slab test with
 * tmin0 = 0, tmax0 = clip; the hit distance is the entry distance tmin, accepted when tmin > 0
 * (origin outside, like sphere.cpp:70) and tmin < running record.t (strict, so ties keep the lower id).
 * Any-hit stops at the first lane that produced a hit, like sphere.cpp:138-141.
 * ---------------------------------------------------------------------------------------------- */
int orc_box_intersect(const orc_scene *s, const float o[3], const float d[3], int any, float clip, orc_hit *hit)
{
    float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    float recordT = clip;
    uint32_t closest = ORC_MISS;
    uint32_t numLanes = (s->num_boxes + LANE - 1) / LANE;
    for (uint32_t i = 0; i < numLanes; i++) {
        const float *lane = s->box_lanes + (size_t)i * 6 * LANE;
        int laneHit = 0;
        for (uint32_t j = 0; j < LANE; j++) {
            if (i * LANE + j >= s->num_boxes) {
                break;
            }
            float b[6];
            for (int k = 0; k < 6; k++) {
                b[k] = lane[k * LANE + j];
            }
            float tmin, tmax;
            if (!orc_bounds_slab(b, o, inv, clip, &tmin, &tmax)) {
                continue;
            }
            if (tmin > 0.0f && tmin < recordT) {
                recordT = tmin;
                closest = i * LANE + j;
                laneHit = 1;
            }
        }
        if (laneHit && any) {
            break;
        }
    }
    if (closest == ORC_MISS) {
        return 0;
    }
    hit->t = recordT;
    hit->prim = ((uint32_t)ORC_KIND_BOX << ORC_KIND_SHIFT) | closest;
    hit->u = hit->v = 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * The query chain.  Closest hit, main.cpp:312-321: Sphere -> Plane -> Cylinder -> KDTree, each stage
 * clipped by the running record.t, so a later class needs a strictly smaller t (appendix A.8).
 * Any hit, main.cpp:198-217: same order, every class sees the original clip, first hit returns.
 * The box extension class runs right after the spheres.
 * ---------------------------------------------------------------------------------------------- */
int orc_query(const orc_scene *s, const orc_ray *ray, uint32_t classes, orc_hit *hit, orc_counters *ctr)
{
    const int any = (ray->flags & ORC_RAY_ANY) != 0;
    float clip = ray->clip;
    int found = 0;
    orc_hit h;
    hit->t = ray->clip;
    hit->prim = ORC_MISS;
    hit->u = hit->v = 0;
    if (ctr) {
        memset(ctr, 0, sizeof(*ctr));
    }
    if ((classes & ORC_CLS_SPHERE) && orc_sphere_intersect(s, ray->o, ray->d, any, clip, &h)) {
        *hit = h;
        found = 1;
        if (any) goto done;
        clip = h.t;
    }
    if ((classes & ORC_CLS_BOX) && orc_box_intersect(s, ray->o, ray->d, any, clip, &h)) {
        *hit = h;
        found = 1;
        if (any) goto done;
        clip = h.t;
    }
    if ((classes & ORC_CLS_PLANE) && orc_plane_intersect(s, ray->o, ray->d, clip, &h)) {
        *hit = h;
        found = 1;
        if (any) goto done;
        clip = h.t;
    }
    if ((classes & ORC_CLS_CYLINDER) && orc_cylinder_intersect(s, ray->o, ray->d, clip, &h)) {
        *hit = h;
        found = 1;
        if (any) goto done;
        clip = h.t;
    }
    if ((classes & ORC_CLS_TREE) && orc_kdtree_intersect(s, ray->o, ray->d, any, &clip, &h, ctr)) {
        *hit = h;
        found = 1;
    }
done:
    if (found && any) {
        hit->prim = 0;
        hit->u = hit->v = 0;
    }
    return found;
}

/* ------------------------------------------------------------------------------------------------
 * Ray construction.
 * Primary rays, main.cpp:275-279,294-310,342-345 for the canonical single band (startRow = 0):
 * xs[0] = -Ratio, xs[j+1] = xs[j] + widthStep ; ys[0] = 1, ys[i+1] = ys[i] - heightStep (repeated
 * fp32 adds, NOT j*step); dir = normalize((xs[j], ys[i], 1)) = v * (1/sqrt(dot(v,v))); origin (0,0,-4.9).
 * ---------------------------------------------------------------------------------------------- */
void orc_ray_tables(uint32_t width, uint32_t height, float *xs, float *ys)
{
    float ratio = (float)width / height; /* config.h:27 */
    float widthStep = 2.0f * ratio / width;
    float heightStep = 2.0f / height;
    float x = -ratio;
    for (uint32_t j = 0; j < width; j++) {
        xs[j] = x;
        x += widthStep;
    }
    float y = 1.0f;
    for (uint32_t i = 0; i < height; i++) {
        ys[i] = y;
        y -= heightStep;
    }
}

void orc_primary_ray(const float *xs, const float *ys, uint32_t row, uint32_t col, orc_ray *out)
{
    float vx = xs[col], vy = ys[row], vz = 1.0f;
    float inv = 1.0f / sqrtf(dot3(vx, vy, vz, vx, vy, vz)); /* glm::normalize, main.cpp:304 */
    out->o[0] = 0.0f;
    out->o[1] = 0.0f;
    out->o[2] = -4.9f;
    out->d[0] = vx * inv;
    out->d[1] = vy * inv;
    out->d[2] = vz * inv;
    out->clip = INFINITY;
    out->flags = 0;
}

void orc_primary_rays(uint32_t width, uint32_t height, orc_ray *out)
{
    float *xs = (float *)malloc(sizeof(float) * width);
    float *ys = (float *)malloc(sizeof(float) * height);
    orc_ray_tables(width, height, xs, ys);
    for (uint32_t i = 0; i < height; i++) {
        for (uint32_t j = 0; j < width; j++) {
            orc_primary_ray(xs, ys, i, j, &out[(size_t)i * width + j]);
        }
    }
    free(xs);
    free(ys);
}

/* hitPoint = rayOrigin + rayDir * t, triangle.cpp:170 / sphere.cpp:156 (mul, then add) */
void orc_hit_point(const float o[3], const float d[3], float t, float p[3])
{
    for (int k = 0; k < 3; k++) {
        float m = d[k] * t;
        p[k] = o[k] + m;
    }
}

/* canSeeLight's ray, main.cpp:184-196 */
void orc_shadow_ray(const float p[3], const float light[3], orc_ray *out)
{
    float lx = light[0] - p[0], ly = light[1] - p[1], lz = light[2] - p[2];
    float dist = sqrtf(dot3(lx, ly, lz, lx, ly, lz)); /* glm::length */
    lx /= dist;
    ly /= dist;
    lz /= dist;
    out->d[0] = lx;
    out->d[1] = ly;
    out->d[2] = lz;
    out->o[0] = p[0] + lx * 0.01f;
    out->o[1] = p[1] + ly * 0.01f;
    out->o[2] = p[2] + lz * 0.01f;
    out->clip = dist;
    out->flags = ORC_RAY_ANY;
}

/* ------------------------------------------------------------------------------------------------
 * Batch drivers (contiguous bands over threads, like main.cpp:371-393).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const orc_scene *s;
    const orc_ray *rays;
    uint32_t classes;
    orc_hit *hits;
    orc_counters *ctrs;
    uint64_t lo, hi;
    /* fused passes */
    int mode; /* 0 = explicit rays, 1 = primary, 2 = shadow */
    uint32_t width, height;
    const float *xs, *ys;
    const orc_hit *primary;
    const float *light;
    uint8_t *visible;
} band_job;

static void *band_main(void *arg)
{
    band_job *job = (band_job *)arg;
    for (uint64_t i = job->lo; i < job->hi; i++) {
        orc_counters *c = job->ctrs ? &job->ctrs[i] : NULL;
        if (job->mode == 0) {
            orc_query(job->s, &job->rays[i], job->classes, &job->hits[i], c);
        } else if (job->mode == 1) {
            orc_ray r;
            orc_primary_ray(job->xs, job->ys, (uint32_t)(i / job->width), (uint32_t)(i % job->width), &r);
            orc_query(job->s, &r, job->classes, &job->hits[i], c);
        } else {
            if (c) {
                memset(c, 0, sizeof(*c));
            }
            if (job->primary[i].prim == ORC_MISS) {
                job->visible[i] = 0;
                continue;
            }
            orc_ray r, sh;
            orc_hit h;
            float p[3];
            orc_primary_ray(job->xs, job->ys, (uint32_t)(i / job->width), (uint32_t)(i % job->width), &r);
            orc_hit_point(r.o, r.d, job->primary[i].t, p);
            orc_shadow_ray(p, job->light, &sh);
            job->visible[i] = orc_query(job->s, &sh, job->classes, &h, c) ? 0 : 1;
        }
    }
    return NULL;
}

static void run_bands(band_job *proto, uint64_t n, int nthreads)
{
    if (nthreads < 1) {
        nthreads = 1;
    }
    if ((uint64_t)nthreads > n) {
        nthreads = n ? (int)n : 1;
    }
    band_job *jobs = (band_job *)malloc(sizeof(band_job) * nthreads);
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    uint64_t chunk = (n + nthreads - 1) / nthreads;
    int started = 0;
    for (int t = 0; t < nthreads; t++) {
        uint64_t lo = chunk * t;
        uint64_t hi = lo + chunk < n ? lo + chunk : n;
        if (lo >= hi) {
            break;
        }
        jobs[t] = *proto;
        jobs[t].lo = lo;
        jobs[t].hi = hi;
        if (nthreads == 1) {
            band_main(&jobs[t]);
        } else {
            pthread_create(&tids[t], NULL, band_main, &jobs[t]);
            started++;
        }
    }
    for (int t = 0; t < started; t++) {
        pthread_join(tids[t], NULL);
    }
    free(jobs);
    free(tids);
}

void orc_intersect(const orc_scene *s, const orc_ray *rays, uint64_t n, uint32_t classes, orc_hit *hits,
                   orc_counters *ctrs, int nthreads)
{
    band_job job;
    memset(&job, 0, sizeof(job));
    job.s = s;
    job.rays = rays;
    job.classes = classes;
    job.hits = hits;
    job.ctrs = ctrs;
    job.mode = 0;
    run_bands(&job, n, nthreads);
}

void orc_trace_primary(const orc_scene *s, uint32_t width, uint32_t height, uint32_t classes, orc_hit *hits,
                       orc_counters *ctrs, int nthreads)
{
    float *xs = (float *)malloc(sizeof(float) * width);
    float *ys = (float *)malloc(sizeof(float) * height);
    orc_ray_tables(width, height, xs, ys);
    band_job job;
    memset(&job, 0, sizeof(job));
    job.s = s;
    job.classes = classes;
    job.hits = hits;
    job.ctrs = ctrs;
    job.mode = 1;
    job.width = width;
    job.height = height;
    job.xs = xs;
    job.ys = ys;
    run_bands(&job, (uint64_t)width * height, nthreads);
    free(xs);
    free(ys);
}

/* visible[i] = 1 iff pixel i had a primary hit and no shape blocks its path to the light */
void orc_trace_shadow(const orc_scene *s, uint32_t width, uint32_t height, uint32_t classes, const orc_hit *hits,
                      const float light[3], uint8_t *visible, orc_counters *ctrs, int nthreads)
{
    float *xs = (float *)malloc(sizeof(float) * width);
    float *ys = (float *)malloc(sizeof(float) * height);
    orc_ray_tables(width, height, xs, ys);
    band_job job;
    memset(&job, 0, sizeof(job));
    job.s = s;
    job.classes = classes;
    job.ctrs = ctrs;
    job.mode = 2;
    job.width = width;
    job.height = height;
    job.xs = xs;
    job.ys = ys;
    job.primary = hits;
    job.light = light;
    job.visible = visible;
    run_bands(&job, (uint64_t)width * height, nthreads);
    free(xs);
    free(ys);
}

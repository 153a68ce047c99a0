// dodrt_device.cuh -- device-side scene view and intersection routines (sm_100a).
//
// Arithmetic contract: fp32, round-to-nearest, NO fused multiply-add (this
// translation unit is compiled with -fmad=false, so every a*b+c below is mul.rn then add.rn), IEEE
// div.rn / sqrt.rn (nvcc defaults -prec-div=true -prec-sqrt=true), denormals kept (-ftz=false), and
// every expression keeps the reference's association order.  Comparisons are written exactly as the
// reference writes them so that NaN/inf (axis-parallel rays: 0*inf in the slab test) take the same
// branches; never replace them with fminf/fmaxf.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/dodrt.h"

namespace dodrt {

constexpr int kLane = 8;        // c_triangleLaneSz triangle.h:32, c_sphereLaneSz sphere.cpp:11
constexpr int kMaxStack = 32;   // >= KDTree::m_maxDepth (kdtree.cpp:72); checked at upload
constexpr uint32_t kLeafFlag = 3u;

// One triangle as the traversal kernels read it: 48 bytes = 3 x LDG.128.
//   q0 = (A.x, A.y, A.z, AB.x)  q1 = (AB.y, AB.z, AC.x, AC.y)  q2 = (AC.z, 0, 0, 0)
// AB = B - A and AC = C - A are the reference's first two operations per triangle
// (triangle.cpp:66-67); doing the same sub.rn once at upload is bit-identical.
// Index of a triangle in this array == its reference id (laneIdx*8 + slot, triangle.cpp:136).
struct DeviceScene {
    const uint2 *nodes;     // kdtree.h:16-48, 8 B each
    const float4 *tris;     // 3 x float4 per triangle slot
    // The same data in the reference's own lane shape (triangle.h:33-44: SoA of 8, 288 B per lane) with
    // B, C replaced by AB, AC: Ax[8] Ay[8] Az[8] ABx[8] ABy[8] ABz[8] ACx[8] ACy[8] ACz[8] = 18 float4.
    // float4 (c*2 + h) holds component c of triangle slots 4h..4h+3.
    const float4 *lanes4;
    uint32_t num_nodes;
    uint32_t num_tri_lanes;
    float bmin[3], bmax[3]; // KDTree::m_bounds
    const float *sphere_lanes; // x[8] y[8] z[8] r2[8]
    uint32_t num_spheres;
    const float *plane_lanes;  // px py pz nx ny nz [8]
    uint32_t num_planes;
    const dodrt_cylinder *cylinders;
    uint32_t num_cylinders;
    const float *box_lanes;    // minx miny minz maxx maxy maxz [8]
    uint32_t num_boxes;
    float epsilon;
    // optional exact culling BVHs over the sphere / box lanes (dodrt_prim_bvh.cuh); nullptr = brute force
    const float4 *sphere_bvh;
    const uint32_t *sphere_bvh_ids;
    const float4 *box_bvh;
    const uint32_t *box_bvh_ids;
    // shading attributes (only read by the render kernels): Triangle::m_triangleAttributes in the reference's
    // layout (80 words per lane, triangle.h:45-51), one colour per mesh / sphere / plane
    const uint32_t *tri_attrs;
    const float *mesh_colors;
    const float *sphere_colors;
    const float *plane_colors;
    // traversal scheduling knobs (dodrt_kernels.cu): leaf phase runs when nLeaf*tune[1] >= nNode*tune[0], or after
    // tune[2] consecutive node phases
    uint32_t tune[4];
    // (-0.0f, -0.0f): the addend that turns FFMA2 into an un-fusable packed multiply (see f2_mul)
    unsigned long long negzero2;
    // kd node steps a ray may take per node phase of the voted loop: 1 = one step per vote; default unlimited = the
    // ray walks on until it stands in a non-empty leaf (or is done), the others wait -- a vote costs about one step
    uint32_t node_burst;
    // variant 7: voted iterations between two polls of the donation queue
    uint32_t donate_poll;
    // variant 4: idle lanes that send the warp back to top up its ray pool (32 = only when every ray is finished)
    uint32_t pool_refill;
    // instrumentation (dodrt_scene_debug_stats, nullptr = off): [0]/[1] culling-BVH nodes fetched / primitives tested for
    // spheres, [2]/[3] the same for boxes, [4] rays that took a culling BVH
    unsigned long long *stats;
    // variant 7: resume steps between two fork polls of a resumed any-hit ray (0 = no work splitting), and how many
    // helpers may wait for a ray at a time (further idle warps leave the kernel)
    uint32_t fork_poll;
    uint32_t helper_limit;
};

struct Hit {
    float t;
    uint32_t prim;
    float u, v;
};

// avxDot avx_utils.h:13-22 == glm::dot: (x1*x2 + y1*y2) + z1*z2
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    float px = ax * bx;
    float py = ay * by;
    float pz = az * bz;
    float acc = px + py;
    return acc + pz;
}

// AxisAlignedBoundingBox::intersect, box.cpp:33-53
__device__ __forceinline__ bool slab(const float bmin[3], const float bmax[3], const float o[3], const float inv[3],
                                     float clip, float &tminOut, float &tmaxOut)
{
    float tmin = 0.0f;
    float tmax = clip;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float tNear = (bmin[i] - o[i]) * inv[i];
        float tFar = (bmax[i] - o[i]) * inv[i];
        if (tNear > tFar) {
            float tmp = tNear;
            tNear = tFar;
            tFar = tmp;
        }
        tmin = tNear > tmin ? tNear : tmin;
        tmax = tFar < tmax ? tFar : tmax;
        if (tmin > tmax) {
            return false;
        }
    }
    tminOut = tmin;
    tmaxOut = tmax;
    return true;
}

// One triangle of Triangle::intersectInRange (triangle.cpp:66-139).  `maxDist` is the running
// maximumDistance: the reference's lane-start compare (triangle.cpp:109-111) and running compare
// (triangle.cpp:133) collapse to this one strict `<` because the running value never grows.
__device__ __forceinline__ bool triangle_test(const float4 q0, const float4 q1, const float4 q2, const float o[3],
                                              const float d[3], float maxDist, float &tOut, float &uOut, float &vOut)
{
    const float Ax = q0.x, Ay = q0.y, Az = q0.z;
    const float ABx = q0.w, ABy = q1.x, ABz = q1.y;
    const float ACx = q1.z, ACy = q1.w, ACz = q2.x;
    float px = d[1] * ACz - d[2] * ACy; // avxCross(rayDir, AC)
    float py = d[2] * ACx - d[0] * ACz;
    float pz = d[0] * ACy - d[1] * ACx;
    float det = dot3(px, py, pz, ABx, ABy, ABz);
    if (!(fabsf(det) > 0.0f)) {
        return false;
    }
    float inv_det = 1.0f / det;
    float tx = o[0] - Ax, ty = o[1] - Ay, tz = o[2] - Az;
    float u = dot3(tx, ty, tz, px, py, pz) * inv_det;
    if (!(u > 0.0f && u < 1.0f)) {
        return false;
    }
    float qx = ty * ABz - tz * ABy; // avxCross(tvec, AB)
    float qy = tz * ABx - tx * ABz;
    float qz = tx * ABy - ty * ABx;
    float v = dot3(d[0], d[1], d[2], qx, qy, qz) * inv_det;
    if (!(v > 0.0f && (u + v) < 1.0f)) {
        return false;
    }
    float t = dot3(ACx, ACy, ACz, qx, qy, qz) * inv_det;
    if (!(t > 0.0f && t < maxDist)) {
        return false;
    }
    tOut = t;
    uOut = u;
    vOut = v;
    return true;
}

// Same result as triangle_test, bit for bit, but the IEEE division (a ~12-instruction sequence) and the
// products that follow it are only executed for triangles that survive CONSERVATIVE sign / magnitude
// tests on the un-divided numerators.  With a = dot(T,p), b = dot(D,q), c = dot(AC,q):
//   u = fl(a * fl(1/det)) > 0  requires a != 0 and sign(a) == sign(det)       (fl(1/det) has det's sign)
//   u < 1                      fails for sure when |a| > 1.00001 |det|        (two roundings ~ 2^-23 << 1e-5)
//   v > 0, (u+v) < 1, t > 0    likewise with b, |a|+|b|, c
// A triangle is only rejected early when the reference's own compare is guaranteed to reject it; all
// survivors (a few %) run the reference's exact expression sequence, so accepted (t,u,v) are identical.
// NaN/inf/denormal inputs simply fall through to the exact path.
__device__ __forceinline__ bool sign_differs_or_zero(float x, float det)
{
    return (int)(__float_as_uint(x) ^ __float_as_uint(det)) < 0 || x == 0.0f;
}

__device__ __forceinline__ bool triangle_test_fast(const float4 q0, const float4 q1, const float4 q2, const float o[3],
                                                   const float d[3], float maxDist, float &tOut, float &uOut,
                                                   float &vOut)
{
    const float Ax = q0.x, Ay = q0.y, Az = q0.z;
    const float ABx = q0.w, ABy = q1.x, ABz = q1.y;
    const float ACx = q1.z, ACy = q1.w, ACz = q2.x;
    float px = d[1] * ACz - d[2] * ACy;
    float py = d[2] * ACx - d[0] * ACz;
    float pz = d[0] * ACy - d[1] * ACx;
    float det = dot3(px, py, pz, ABx, ABy, ABz);
    if (!(fabsf(det) > 0.0f)) {
        return false;
    }
    float tx = o[0] - Ax, ty = o[1] - Ay, tz = o[2] - Az;
    float a = dot3(tx, ty, tz, px, py, pz);
    const float lim = fabsf(det) * 1.00001f;
    if (sign_differs_or_zero(a, det) || fabsf(a) > lim) {
        return false;
    }
    float qx = ty * ABz - tz * ABy;
    float qy = tz * ABx - tx * ABz;
    float qz = tx * ABy - ty * ABx;
    float b = dot3(d[0], d[1], d[2], qx, qy, qz);
    if (sign_differs_or_zero(b, det) || fabsf(a) + fabsf(b) > lim) {
        return false;
    }
    float c = dot3(ACx, ACy, ACz, qx, qy, qz);
    if (sign_differs_or_zero(c, det)) {
        return false;
    }
    // exact tail: triangle.cpp:81-111 in the reference's order
    float inv_det = 1.0f / det;
    float u = a * inv_det;
    if (!(u > 0.0f && u < 1.0f)) {
        return false;
    }
    float v = b * inv_det;
    if (!(v > 0.0f && (u + v) < 1.0f)) {
        return false;
    }
    float t = c * inv_det;
    if (!(t > 0.0f && t < maxDist)) {
        return false;
    }
    tOut = t;
    uOut = u;
    vOut = v;
    return true;
}

// First stage of triangle_test_fast for one triangle, branch-free: can this triangle still be accepted
// after the det / u tests?  (Used four triangles at a time on the SoA lanes, see lane_half_test.)
__device__ __forceinline__ bool triangle_may_hit(float Ax, float Ay, float Az, float ABx, float ABy, float ABz, float ACx,
                                                 float ACy, float ACz, const float o[3], const float d[3])
{
    float px = d[1] * ACz - d[2] * ACy;
    float py = d[2] * ACx - d[0] * ACz;
    float pz = d[0] * ACy - d[1] * ACx;
    float det = dot3(px, py, pz, ABx, ABy, ABz);
    float tx = o[0] - Ax, ty = o[1] - Ay, tz = o[2] - Az;
    float a = dot3(tx, ty, tz, px, py, pz);
    const float ad = fabsf(det);
    // (ad > 0) is the reference's |det| > 0 (NaN fails); the other two are the conservative u tests
    return (ad > 0.0f) & !sign_differs_or_zero(a, det) & !(fabsf(a) > ad * 1.00001f);
}

// All conservative stages of triangle_test_fast for one triangle, branch-free (det, u, v, u + v and the sign of t on
// the un-divided numerators): a triangle that fails is rejected by the reference's own compares for sure; what passes
// is, up to the last-bit margins, a real intersection in front of the origin.  Used where ONE warp works on one ray and
// the only instruction-level parallelism is between the slots a thread tests (dodrt_donate.inl).
__device__ __forceinline__ bool triangle_may_hit_full(float Ax, float Ay, float Az, float ABx, float ABy, float ABz, float ACx,
                                                      float ACy, float ACz, const float o[3], const float d[3])
{
    float px = d[1] * ACz - d[2] * ACy;
    float py = d[2] * ACx - d[0] * ACz;
    float pz = d[0] * ACy - d[1] * ACx;
    float det = dot3(px, py, pz, ABx, ABy, ABz);
    float tx = o[0] - Ax, ty = o[1] - Ay, tz = o[2] - Az;
    float a = dot3(tx, ty, tz, px, py, pz);
    float qx = ty * ABz - tz * ABy;
    float qy = tz * ABx - tx * ABz;
    float qz = tx * ABy - ty * ABx;
    float b = dot3(d[0], d[1], d[2], qx, qy, qz);
    float c = dot3(ACx, ACy, ACz, qx, qy, qz);
    const float ad = fabsf(det), lim = ad * 1.00001f;
    // exactly the early-out predicates of triangle_test_fast, in its polarity (NaN numerators fall through to the exact test)
    return (ad > 0.0f) & !sign_differs_or_zero(a, det) & !(fabsf(a) > lim) & !sign_differs_or_zero(b, det) &
           !(fabsf(a) + fabsf(b) > lim) & !sign_differs_or_zero(c, det);
}

// Four consecutive triangle slots (4h .. 4h+3) of one SoA lane: 9 x LDG.128, a branch-free first stage
// for all four (independent dependency chains), and the full test -- in slot order, against the running
// clip -- only for the survivors.  Returns true when at least one triangle was accepted.
__device__ __forceinline__ bool lane_half_test(const float4 *__restrict__ lane, int h, uint32_t firstId, const float o[3],
                                               const float d[3], float &clip, Hit &hit)
{
    const float4 Ax = __ldg(lane + 0 + h), Ay = __ldg(lane + 2 + h), Az = __ldg(lane + 4 + h);
    const float4 Bx = __ldg(lane + 6 + h), By = __ldg(lane + 8 + h), Bz = __ldg(lane + 10 + h);
    const float4 Cx = __ldg(lane + 12 + h), Cy = __ldg(lane + 14 + h), Cz = __ldg(lane + 16 + h);
    const bool m0 = triangle_may_hit(Ax.x, Ay.x, Az.x, Bx.x, By.x, Bz.x, Cx.x, Cy.x, Cz.x, o, d);
    const bool m1 = triangle_may_hit(Ax.y, Ay.y, Az.y, Bx.y, By.y, Bz.y, Cx.y, Cy.y, Cz.y, o, d);
    const bool m2 = triangle_may_hit(Ax.z, Ay.z, Az.z, Bx.z, By.z, Bz.z, Cx.z, Cy.z, Cz.z, o, d);
    const bool m3 = triangle_may_hit(Ax.w, Ay.w, Az.w, Bx.w, By.w, Bz.w, Cx.w, Cy.w, Cz.w, o, d);
    if (!(m0 | m1 | m2 | m3)) {
        return false;
    }
    bool any = false;
    float t, u, v;
#define DODRT_SLOT(k, M, c)                                                                                        \
    if (M && triangle_test_fast(make_float4(Ax.c, Ay.c, Az.c, Bx.c), make_float4(By.c, Bz.c, Cx.c, Cy.c),          \
                                make_float4(Cz.c, 0.0f, 0.0f, 0.0f), o, d, clip, t, u, v)) {                        \
        clip = t;                                                                                                  \
        hit.t = t;                                                                                                 \
        hit.prim = (DODRT_KIND_TRIANGLE << DODRT_KIND_SHIFT) | (firstId + k);                                      \
        hit.u = u;                                                                                                 \
        hit.v = v;                                                                                                 \
        any = true;                                                                                                \
    }
    DODRT_SLOT(0, m0, x)
    DODRT_SLOT(1, m1, y)
    DODRT_SLOT(2, m2, z)
    DODRT_SLOT(3, m3, w)
#undef DODRT_SLOT
    return any;
}

// First stage of lane_half_test alone: bit k of the result = slot 4h+k may still be accepted.
__device__ __forceinline__ uint32_t lane_half_mask(const float4 *__restrict__ lane, int h, const float o[3], const float d[3])
{
    const float4 Ax = __ldg(lane + 0 + h), Ay = __ldg(lane + 2 + h), Az = __ldg(lane + 4 + h);
    const float4 Bx = __ldg(lane + 6 + h), By = __ldg(lane + 8 + h), Bz = __ldg(lane + 10 + h);
    const float4 Cx = __ldg(lane + 12 + h), Cy = __ldg(lane + 14 + h), Cz = __ldg(lane + 16 + h);
    const uint32_t m0 = triangle_may_hit(Ax.x, Ay.x, Az.x, Bx.x, By.x, Bz.x, Cx.x, Cy.x, Cz.x, o, d);
    const uint32_t m1 = triangle_may_hit(Ax.y, Ay.y, Az.y, Bx.y, By.y, Bz.y, Cx.y, Cy.y, Cz.y, o, d);
    const uint32_t m2 = triangle_may_hit(Ax.z, Ay.z, Az.z, Bx.z, By.z, Bz.z, Cx.z, Cy.z, Cz.z, o, d);
    const uint32_t m3 = triangle_may_hit(Ax.w, Ay.w, Az.w, Bx.w, By.w, Bz.w, Cx.w, Cy.w, Cz.w, o, d);
    return m0 | (m1 << 1) | (m2 << 2) | (m3 << 3);
}

// One whole lane (8 slots), variant 8: the branch-free first stage for both halves, then ONE copy of the exact
// test in a loop over the surviving slots, lowest slot first (so the running clip evolves as in the reference's
// slot loop, triangle.cpp:119-139).  A survivor's nine floats are re-read from the lane (L1 hits).  lane_half_test
// inlines the exact test once per slot -- eight copies per lane, each executed whenever ANY ray of the warp has a
// survivor in that slot (ncu: 17 % of the shadow pass at 4-7 active threads); here the loop runs max-over-rays
// survivor-count times and the hot loop is ~500 SASS instructions shorter.
__device__ __forceinline__ bool lane_test_compact(const float4 *__restrict__ lane, uint32_t firstId, const float o[3],
                                                  const float d[3], float &clip, Hit &hit)
{
    uint32_t m = lane_half_mask(lane, 0, o, d) | (lane_half_mask(lane, 1, o, d) << 4);
    bool any = false;
    while (m != 0u) {
        const uint32_t k = (uint32_t)__ffs((int)m) - 1u;
        m &= m - 1u;
        const float *base = reinterpret_cast<const float *>(lane) + k;
        const float4 q0 = make_float4(__ldg(base), __ldg(base + 8), __ldg(base + 16), __ldg(base + 24));
        const float4 q1 = make_float4(__ldg(base + 32), __ldg(base + 40), __ldg(base + 48), __ldg(base + 56));
        const float4 q2 = make_float4(__ldg(base + 64), 0.0f, 0.0f, 0.0f);
        float t, u, v;
        if (triangle_test_fast(q0, q1, q2, o, d, clip, t, u, v)) {
            clip = t;
            hit.t = t;
            hit.prim = (DODRT_KIND_TRIANGLE << DODRT_KIND_SHIFT) | (firstId + k);
            hit.u = u;
            hit.v = v;
            any = true;
        }
    }
    return any;
}

// ---- packed fp32 pairs: sm_100a FMUL2 / FADD2 / FFMA2 ------------------------------------------------------------
// One issue slot does the SAME IEEE operation on two independent fp32 values held in a 64-bit register pair
// (PTX add/sub/fma .rn.f32x2).  Each half is rounded exactly like the scalar mul.rn / add.rn, so results are
// bit-identical to the scalar code -- PROVIDED ptxas does not contract a packed multiply with the add that
// consumes it: ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false (it honours .rn only
// for the scalar forms), and it also folds fma(a, b, -0.0) back into a multiply when the -0.0 is a literal.  The
// multiply is therefore written as fma.rn.f32x2(a, b, nz) with nz = (-0.0f, -0.0f) read from a kernel parameter
// (DeviceScene::negzero2), which ptxas cannot see through: fl(a*b + (-0)) == fl(a*b) for every input (the sum of
// an exact product and -0 keeps the product's sign, also for +-0, and inf/NaN propagate alike), and an FFMA2
// cannot be fused into the FADD2 that follows.  tests/test_sass_audit.py checks the compiled SASS: every FFMA2
// of the trace kernels has the uniform -0.0 pair as its addend.
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 f2_pack(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float f2_lo(f32x2 v)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo;
}
__device__ __forceinline__ float f2_hi(f32x2 v)
{
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return hi;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b, f32x2 nz)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
    return r;
}
// dot3 for two triangles at once, same association as avxDot: (x1*x2 + y1*y2) + z1*z2
__device__ __forceinline__ f32x2 f2_dot3(f32x2 ax, f32x2 ay, f32x2 az, f32x2 bx, f32x2 by, f32x2 bz, f32x2 nz)
{
    return f2_add(f2_add(f2_mul(ax, bx, nz), f2_mul(ay, by, nz)), f2_mul(az, bz, nz));
}

// triangle_may_hit for TWO triangle slots (the halves of each operand): the reference's det and a = dot(T, pvec),
// each half bit-identical to the scalar routine above; the comparisons run on the unpacked halves.
__device__ __forceinline__ void triangle_may_hit2(f32x2 Ax, f32x2 Ay, f32x2 Az, f32x2 ABx, f32x2 ABy, f32x2 ABz, f32x2 ACx,
                                                  f32x2 ACy, f32x2 ACz, const f32x2 o2[3], const f32x2 d2[3], f32x2 nz,
                                                  bool &m0, bool &m1)
{
    const f32x2 px = f2_sub(f2_mul(d2[1], ACz, nz), f2_mul(d2[2], ACy, nz));
    const f32x2 py = f2_sub(f2_mul(d2[2], ACx, nz), f2_mul(d2[0], ACz, nz));
    const f32x2 pz = f2_sub(f2_mul(d2[0], ACy, nz), f2_mul(d2[1], ACx, nz));
    const f32x2 det = f2_dot3(px, py, pz, ABx, ABy, ABz, nz);
    const f32x2 tx = f2_sub(o2[0], Ax), ty = f2_sub(o2[1], Ay), tz = f2_sub(o2[2], Az);
    const f32x2 a = f2_dot3(tx, ty, tz, px, py, pz, nz);
    const float det0 = f2_lo(det), det1 = f2_hi(det), a0 = f2_lo(a), a1 = f2_hi(a);
    const float ad0 = fabsf(det0), ad1 = fabsf(det1);
    m0 = (ad0 > 0.0f) & !sign_differs_or_zero(a0, det0) & !(fabsf(a0) > ad0 * 1.00001f);
    m1 = (ad1 > 0.0f) & !sign_differs_or_zero(a1, det1) & !(fabsf(a1) > ad1 * 1.00001f);
}

__device__ __forceinline__ f32x2 f2_from(const float4 &q, int pair)
{
    return pair == 0 ? f2_pack(q.x, q.y) : f2_pack(q.z, q.w);
}

// lane_half_test with the first stage on packed pairs (variant 6): 44 instead of 88 fp32 issue slots per four
// triangles.  The exact second stage is the scalar one, in slot order.
__device__ __forceinline__ bool lane_half_test_packed(const float4 *__restrict__ lane, int h, uint32_t firstId,
                                                      const float o[3], const float d[3], const f32x2 o2[3],
                                                      const f32x2 d2[3], f32x2 nz, float &clip, Hit &hit)
{
    const float4 Ax = __ldg(lane + 0 + h), Ay = __ldg(lane + 2 + h), Az = __ldg(lane + 4 + h);
    const float4 Bx = __ldg(lane + 6 + h), By = __ldg(lane + 8 + h), Bz = __ldg(lane + 10 + h);
    const float4 Cx = __ldg(lane + 12 + h), Cy = __ldg(lane + 14 + h), Cz = __ldg(lane + 16 + h);
    bool m0, m1, m2, m3;
    triangle_may_hit2(f2_from(Ax, 0), f2_from(Ay, 0), f2_from(Az, 0), f2_from(Bx, 0), f2_from(By, 0), f2_from(Bz, 0),
                      f2_from(Cx, 0), f2_from(Cy, 0), f2_from(Cz, 0), o2, d2, nz, m0, m1);
    triangle_may_hit2(f2_from(Ax, 1), f2_from(Ay, 1), f2_from(Az, 1), f2_from(Bx, 1), f2_from(By, 1), f2_from(Bz, 1),
                      f2_from(Cx, 1), f2_from(Cy, 1), f2_from(Cz, 1), o2, d2, nz, m2, m3);
    if (!(m0 | m1 | m2 | m3)) {
        return false;
    }
    bool any = false;
    float t, u, v;
#define DODRT_SLOT(k, M, c)                                                                                        \
    if (M && triangle_test_fast(make_float4(Ax.c, Ay.c, Az.c, Bx.c), make_float4(By.c, Bz.c, Cx.c, Cy.c),          \
                                make_float4(Cz.c, 0.0f, 0.0f, 0.0f), o, d, clip, t, u, v)) {                        \
        clip = t;                                                                                                  \
        hit.t = t;                                                                                                 \
        hit.prim = (DODRT_KIND_TRIANGLE << DODRT_KIND_SHIFT) | (firstId + k);                                      \
        hit.u = u;                                                                                                 \
        hit.v = v;                                                                                                 \
        any = true;                                                                                                \
    }
    DODRT_SLOT(0, m0, x)
    DODRT_SLOT(1, m1, y)
    DODRT_SLOT(2, m2, z)
    DODRT_SLOT(3, m3, w)
#undef DODRT_SLOT
    return any;
}

// Sphere::intersect_impl, sphere.cpp:26-160 (lane-structured for the any-hit break, sphere.cpp:138-141).
// Four spheres at a time: the cheap part of the reference's sequence (L, distSq, tca, d2 and its two rejections,
// sphere.cpp:62-90) runs branch-free for all four -- independent dependency chains, no divergence -- and only spheres
// that pass it take the sqrt and the t0/t1 tests, in slot order.  Same operations on the same operands as the
// slot-by-slot loop, so every accepted distance has the same bits.
__device__ __forceinline__ bool sphere_query(const DeviceScene &s, const float o[3], const float d[3], bool any,
                                             float clip, Hit &hit)
{
    float recordT = clip;
    uint32_t closest = DODRT_MISS;
    const uint32_t numLanes = (s.num_spheres + kLane - 1) / kLane;
    for (uint32_t i = 0; i < numLanes; i++) {
        const float4 *lane = reinterpret_cast<const float4 *>(s.sphere_lanes + (size_t)i * 4 * kLane);
        uint32_t minIdx = 0;
        float minDist = recordT;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float4 X = __ldg(lane + 0 + h), Y = __ldg(lane + 2 + h), Z = __ldg(lane + 4 + h), R = __ldg(lane + 6 + h);
            const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w};
            const float zs[4] = {Z.x, Z.y, Z.z, Z.w}, rs[4] = {R.x, R.y, R.z, R.w};
            float tca[4], d2[4];
            bool ok[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t j = h * 4 + k;
                float lx = xs[k] - o[0];
                float ly = ys[k] - o[1];
                float lz = zs[k] - o[2];
                float distSq = dot3(lx, ly, lz, lx, ly, lz);
                tca[k] = dot3(lx, ly, lz, d[0], d[1], d[2]);
                float tcaSq = tca[k] * tca[k];
                d2[k] = distSq - tcaSq;
                // last-lane mask (sphere.cpp:31-37,46-49), origin outside (sphere.cpp:70), line within the radius (:88)
                ok[k] = (i * kLane + j < s.num_spheres) & (distSq > rs[k]) & (d2[k] < rs[k]);
            }
            if (!(ok[0] | ok[1] | ok[2] | ok[3])) {
                continue;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (!ok[k]) {
                    continue;
                }
                float thcSq = rs[k] - d2[k];
                float thc = sqrtf(thcSq);
                float t0 = tca[k] - thc;
                float t1 = tca[k] + thc;
                if (!(t0 >= 0.0f && t1 >= 0.0f)) {
                    continue;
                }
                float tm = t0 < t1 ? t0 : t1; // _mm256_min_ps operand rule
                if (tm < minDist) {
                    minDist = tm;
                    minIdx = h * 4 + k;
                }
            }
        }
        if (minDist < recordT) {
            recordT = minDist;
            closest = i * kLane + minIdx;
            if (any) {
                break;
            }
        }
    }
    if (closest == DODRT_MISS) {
        return false;
    }
    hit.t = recordT;
    hit.prim = (DODRT_KIND_SPHERE << DODRT_KIND_SHIFT) | closest;
    hit.u = hit.v = 0.0f;
    return true;
}

// Plane::intersect_impl, plane.cpp:27-139
__device__ __forceinline__ bool plane_query(const DeviceScene &s, const float o[3], const float d[3], float clip,
                                            Hit &hit)
{
    float minT = clip;
    uint32_t closest = DODRT_MISS;
    const float eps = s.epsilon;
    const uint32_t numLanes = (s.num_planes + kLane - 1) / kLane;
    for (uint32_t i = 0; i < numLanes; i++) {
        const float *lane = s.plane_lanes + (size_t)i * 6 * kLane;
#pragma unroll
        for (int j = 0; j < kLane; j++) {
            float px = __ldg(lane + 0 * kLane + j), py = __ldg(lane + 1 * kLane + j), pz = __ldg(lane + 2 * kLane + j);
            float nx = __ldg(lane + 3 * kLane + j), ny = __ldg(lane + 4 * kLane + j), nz = __ldg(lane + 5 * kLane + j);
            float denom = dot3(d[0], d[1], d[2], nx, ny, nz);
            if (!(fabsf(denom) > eps)) {
                continue;
            }
            float vx = px - o[0], vy = py - o[1], vz = pz - o[2];
            float num = dot3(vx, vy, vz, nx, ny, nz);
            float t = num / denom;
            if (!(t > eps)) {
                continue;
            }
            if (t < minT) {
                minT = t;
                closest = i * kLane + j;
            }
        }
    }
    if (closest == DODRT_MISS) {
        return false;
    }
    hit.t = minT;
    hit.prim = (DODRT_KIND_PLANE << DODRT_KIND_SHIFT) | closest;
    hit.u = hit.v = 0.0f;
    return true;
}

// minNonNegative, cylinder.cpp:8-26
__device__ __forceinline__ float min_non_negative(float a, float b)
{
    if (a < 0 && b < 0) {
        return __int_as_float(0x7f800000);
    } else if (a < 0) {
        return b;
    } else if (b < 0) {
        return a;
    }
    return fminf(a, b);
}

// Cylinder::intersect_cylinder_body, cylinder.cpp:76-118.  The two roots go through double because
// the reference calls ::sqrt(double) on a float (cylinder.cpp:92-93).
__device__ __forceinline__ bool cylinder_body(const dodrt_cylinder &c, const float o[3], const float d[3], float eps,
                                              float &tOut)
{
    float dpx = o[0] - c.base[0], dpy = o[1] - c.base[1], dpz = o[2] - c.base[2];
    float k = dot3(d[0], d[1], d[2], c.axis[0], c.axis[1], c.axis[2]);
    float vx = d[0] - k * c.axis[0], vy = d[1] - k * c.axis[1], vz = d[2] - k * c.axis[2];
    float m = dot3(dpx, dpy, dpz, c.axis[0], c.axis[1], c.axis[2]);
    float rx = dpx - m * c.axis[0], ry = dpy - m * c.axis[1], rz = dpz - m * c.axis[2];
    float a = dot3(vx, vy, vz, vx, vy, vz);
    float b = 2.0f * dot3(vx, vy, vz, rx, ry, rz);
    float cc = dot3(rx, ry, rz, rx, ry, rz) - c.radius_sq;
    float disc = (b * b) - (4.0f * a * cc);
    if (disc < eps) {
        return false;
    }
    double sq = sqrt((double)disc);
    double den = (double)(2.0f * a);
    float tSub = (float)(__ddiv_rn(__dsub_rn((double)(-b), sq), den));
    float tAdd = (float)(__ddiv_rn(__dadd_rn((double)(-b), sq), den));
    float t = min_non_negative(tSub, tAdd);
    if (t == __int_as_float(0x7f800000)) {
        return false;
    }
    float cx = (o[0] + d[0] * t) - c.base[0];
    float cy = (o[1] + d[1] * t) - c.base[1];
    float cz = (o[2] + d[2] * t) - c.base[2];
    float f = dot3(cx, cy, cz, c.axis[0], c.axis[1], c.axis[2]);
    if (f < 0.f || f > c.height) {
        return false;
    }
    tOut = t;
    return true;
}

// Cylinder::intersect_cylinder_disc, cylinder.cpp:120-152 (minT is the ORIGINAL clip, cylinder.cpp:122)
__device__ __forceinline__ bool cylinder_disc(const dodrt_cylinder &c, const float o[3], const float d[3], float eps,
                                              float offset, float clip, float &tOut)
{
    float px = c.base[0] + c.axis[0] * offset;
    float py = c.base[1] + c.axis[1] * offset;
    float pz = c.base[2] + c.axis[2] * offset;
    float denom = dot3(d[0], d[1], d[2], c.axis[0], c.axis[1], c.axis[2]);
    if (fabsf(denom) < eps) {
        return false;
    }
    float vx = px - o[0], vy = py - o[1], vz = pz - o[2];
    float tnum = dot3(vx, vy, vz, c.axis[0], c.axis[1], c.axis[2]);
    float t = tnum / denom;
    if (t < eps || t > clip) {
        return false;
    }
    float hx = o[0] + d[0] * t, hy = o[1] + d[1] * t, hz = o[2] + d[2] * t;
    float wx = hx - px, wy = hy - py, wz = hz - pz;
    if (dot3(wx, wy, wz, wx, wy, wz) > c.radius_sq) {
        return false;
    }
    tOut = t;
    return true;
}

// Cylinder::intersect_non_vectorized, cylinder.cpp:155-210
__device__ __forceinline__ bool cylinder_query(const DeviceScene &s, const float o[3], const float d[3], float clip,
                                               Hit &hit)
{
    uint32_t minIdx = DODRT_MISS;
    float tMin = clip;
    for (uint32_t i = 0; i < s.num_cylinders; i++) {
        const dodrt_cylinder c = s.cylinders[i];
        float t;
        if (cylinder_body(c, o, d, s.epsilon, t) && t < tMin) {
            tMin = t;
            minIdx = i;
        }
        if (cylinder_disc(c, o, d, s.epsilon, 0.0f, clip, t) && t < tMin) {
            tMin = t;
            minIdx = i;
        }
        if (cylinder_disc(c, o, d, s.epsilon, c.height, clip, t) && t < tMin) {
            tMin = t;
            minIdx = i;
        }
    }
    if (minIdx == DODRT_MISS) {
        return false;
    }
    hit.t = tMin;
    hit.prim = (DODRT_KIND_CYLINDER << DODRT_KIND_SHIFT) | minIdx;
    hit.u = hit.v = 0.0f;
    return true;
}

// EXTENSION: renderable boxes with the slab arithmetic of box.cpp:33-53 (see include/dodrt.h)
__device__ __forceinline__ bool box_query(const DeviceScene &s, const float o[3], const float d[3], bool any,
                                          float clip, Hit &hit)
{
    const float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    float recordT = clip;
    uint32_t closest = DODRT_MISS;
    const uint32_t numLanes = (s.num_boxes + kLane - 1) / kLane;
    for (uint32_t i = 0; i < numLanes; i++) {
        const float *lane = s.box_lanes + (size_t)i * 6 * kLane;
        bool laneHit = false;
#pragma unroll
        for (int j = 0; j < kLane; j++) {
            if (i * kLane + j >= s.num_boxes) {
                continue;
            }
            float bmin[3] = {__ldg(lane + 0 * kLane + j), __ldg(lane + 1 * kLane + j), __ldg(lane + 2 * kLane + j)};
            float bmax[3] = {__ldg(lane + 3 * kLane + j), __ldg(lane + 4 * kLane + j), __ldg(lane + 5 * kLane + j)};
            float tmin, tmax;
            if (!slab(bmin, bmax, o, inv, clip, tmin, tmax)) {
                continue;
            }
            if (tmin > 0.0f && tmin < recordT) {
                recordT = tmin;
                closest = i * kLane + j;
                laneHit = true;
            }
        }
        if (laneHit && any) {
            break;
        }
    }
    if (closest == DODRT_MISS) {
        return false;
    }
    hit.t = recordT;
    hit.prim = (DODRT_KIND_BOX << DODRT_KIND_SHIFT) | closest;
    hit.u = hit.v = 0.0f;
    return true;
}

// work item -> pixel for the frame modes (see dodrt_frame in include/dodrt.h): tiles round-robin over
// ranks, 8x4 pixel blocks inside a tile so that one warp = one block.  `order` (optional) is the order in which
// the call's local tiles are PROCESSED; `slot` is where the item's result goes in a compact buffer and does
// not depend on it.
__device__ __forceinline__ bool slot_to_pixel(const dodrt_frame &f, uint32_t tiles_x, const uint32_t *order, uint64_t item,
                                              uint32_t &col, uint32_t &row, uint64_t &slot)
{
    const uint32_t tilePixels = f.tile_w * f.tile_h;
    uint32_t localTile = (uint32_t)(item / tilePixels);
    const uint32_t in = (uint32_t)(item % tilePixels);
    if (order) {
        localTile = __ldg(order + localTile);
    }
    slot = (uint64_t)localTile * tilePixels + in;
    const uint32_t tile = f.first_tile + localTile * f.tile_stride;
    const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
    const uint32_t block = in >> 5, lane = in & 31u;
    const uint32_t bpr = f.tile_w >> 3;
    const uint32_t bx = block % bpr, by = block / bpr;
    col = tx * f.tile_w + bx * 8 + (lane & 7u);
    row = ty * f.tile_h + by * 4 + (lane >> 3);
    return col < f.width && row < f.height;
}

// Primary ray direction, main.cpp:304: glm::normalize(v) = v * (1 / sqrt(dot(v,v)))
__device__ __forceinline__ void primary_dir(float vx, float vy, float d[3])
{
    const float vz = 1.0f;
    float inv = 1.0f / sqrtf(dot3(vx, vy, vz, vx, vy, vz));
    d[0] = vx * inv;
    d[1] = vy * inv;
    d[2] = vz * inv;
}

// canSeeLight's ray, main.cpp:184-196, from hitPoint = o + d*t (triangle.cpp:170 / sphere.cpp:156)
__device__ __forceinline__ void shadow_ray(const float o[3], const float d[3], float t, const float light[3],
                                           float so[3], float sd[3], float &clip)
{
    float p[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float m = d[k] * t;
        p[k] = o[k] + m;
    }
    float lx = light[0] - p[0], ly = light[1] - p[1], lz = light[2] - p[2];
    float dist = sqrtf(dot3(lx, ly, lz, lx, ly, lz));
    lx = lx / dist;
    ly = ly / dist;
    lz = lz / dist;
    sd[0] = lx;
    sd[1] = ly;
    sd[2] = lz;
    so[0] = p[0] + lx * 0.01f;
    so[1] = p[1] + ly * 0.01f;
    so[2] = p[2] + lz * 0.01f;
    clip = dist;
}

} // namespace dodrt

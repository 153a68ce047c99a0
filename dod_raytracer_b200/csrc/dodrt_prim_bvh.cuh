// dodrt_prim_bvh.cuh -- exact culling structure over the sphere / box lanes (BASELINE.json config 4).
//
// The reference tests EVERY sphere for every ray (sphere.cpp:39: brute force over all lanes) and keeps the
// first strictly smaller distance in id order, i.e. the lexicographic minimum of (t, id) over the spheres its
// per-sphere arithmetic accepts with t < clip; an any-hit query only asks whether such a sphere exists.
// Neither answer depends on the ORDER in which candidates are met, so any structure that (a) never skips a
// sphere the reference's arithmetic could accept and (b) reduces with the (t, id) key is result-identical.
// This file is such a structure: a binary BVH over the primitives' boxes, built on the host at upload
// (median split on the longest axis, leaves of <= 8 primitives), traversed per ray -- near child first -- with a
// CONSERVATIVE test.
//
// Why the test is conservative.  For a sphere (C, r) and a ray (O, D) the reference computes, in fp32,
// L = C-O, distSq = L.L, tca = L.D, d2 = distSq - tca*tca and requires d2 < r^2.  With e = 2^-24 every product /
// sum carries a relative error <= e, so |d2 - d2*| <= 16 e |L|^2 < 1e-6 |L|^2 against the exact value d2*:
// an accepted sphere has an exact line-to-centre distance below sqrt(r^2 + 1e-6 |L|^2) <= r + 1e-3 |L|.
// A node is therefore only skipped when the ray misses its box inflated by pad = 4e-3 * (distance from O to
// the farthest point of the box) -- four times that bound, and four orders of magnitude above the rounding of
// the slab test itself.  The same pad covers the box primitives (their own slab test is accurate to a few e).
// Distance pruning for closest-hit queries uses the same pad.  tests/test_gpu_parity.py and
// tests/test_gpu_fuzz.py compare against the brute-force oracle bit for bit.
//
// CONTRACT: the bound above needs |D| = 1 (tca = L.D is only the projection length then; with |D|^2 = 1 + delta the
// reference's d2 = distSq - tca^2 is off the geometric value by delta * tca^2 and a far-away sphere can be accepted).
// dodrt_ray.d is not required to be normalised (include/dodrt.h), so the kernels use this structure only for rays
// with | |D|^2 - 1 | <= 2e-6 (every normalised fp32 direction; the pad then still covers the bound twice) and
// run the reference's brute-force loop for all others (analytic_chain, dodrt_kernels.cu).  Primitives with NaN / inf
// coordinates or radii have no box: the upload then skips the structure for that class altogether.
#pragma once
#include "dodrt_device.cuh"

#include <vector>

namespace dodrt {

constexpr uint32_t kPrimBvhLeaf = 0x80000000u;
constexpr int kPrimBvhLeafSize = 8;
constexpr uint32_t kPrimBvhMinCount = 64; // below this the reference's brute force is used as is
constexpr float kPrimBvhPad = 4e-3f;

// 32-byte node = 2 x float4: (bmin.xyz, a) (bmax.xyz, b).  interior: a = left child, b = right child;
// leaf: a = first entry in the id list, b = kPrimBvhLeaf | count.
struct PrimBvhNode {
    float bmin[3];
    uint32_t a;
    float bmax[3];
    uint32_t b;
};

// host: boxes = N x {min xyz, max xyz}
void build_prim_bvh(const float *boxes, uint32_t count, std::vector<PrimBvhNode> &nodes, std::vector<uint32_t> &ids);

#ifdef __CUDACC__
// May the ray touch anything inside this node within [0, tLimit]?  Never answers "no" for a node that holds a
// primitive the reference arithmetic could accept (see the header comment).
__device__ __forceinline__ bool prim_bvh_may_touch(const float4 lo4, const float4 hi4, const float o[3], const float d[3],
                                                   const float inv[3], float tLimit)
{
    const float lo[3] = {lo4.x, lo4.y, lo4.z}, hi[3] = {hi4.x, hi4.y, hi4.z};
    // distance from the origin to the farthest point of the box, over-estimated by the L1 norm
    float far = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        far += fmaxf(fabsf(lo[k] - o[k]), fabsf(hi[k] - o[k]));
    }
    const float pad = kPrimBvhPad * far;
    float t0 = 0.0f, t1 = tLimit + pad;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float a = lo[k] - pad - o[k], b = hi[k] + pad - o[k];
        if (d[k] == 0.0f || !(fabsf(inv[k]) < 3.0e38f)) { // parallel (or numerically parallel) to the slab
            if (a > 0.0f || b < 0.0f) {
                return false;
            }
        } else {
            const float ta = a * inv[k], tb = b * inv[k];
            t0 = fmaxf(t0, fminf(ta, tb));
            t1 = fminf(t1, fmaxf(ta, tb));
        }
    }
    return !(t0 > t1); // NaN anywhere -> keep the node
}

// Same test, also handing back key = entry distance - pad: with a smaller limit L' the node would still pass iff
// key <= L' (its slab exits do not change), which lets a stacked node be dropped without touching memory again.  The
// re-association (t0 - pad <= L' instead of t0 <= L' + pad) moves the threshold by an ulp of a bound that is padded four
// times over.
__device__ __forceinline__ bool prim_bvh_may_touch_key(const float4 lo4, const float4 hi4, const float o[3], const float d[3],
                                                       const float inv[3], float tLimit, float &key)
{
    const float lo[3] = {lo4.x, lo4.y, lo4.z}, hi[3] = {hi4.x, hi4.y, hi4.z};
    float far = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        far += fmaxf(fabsf(lo[k] - o[k]), fabsf(hi[k] - o[k]));
    }
    const float pad = kPrimBvhPad * far;
    float t0 = 0.0f, t1 = tLimit + pad;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float a = lo[k] - pad - o[k], b = hi[k] + pad - o[k];
        if (d[k] == 0.0f || !(fabsf(inv[k]) < 3.0e38f)) {
            if (a > 0.0f || b < 0.0f) {
                return false;
            }
        } else {
            const float ta = a * inv[k], tb = b * inv[k];
            t0 = fmaxf(t0, fminf(ta, tb));
            t1 = fminf(t1, fmaxf(ta, tb));
        }
    }
    key = t0 - pad; // NaN key: never dropped later (the comparison below is false)
    return !(t0 > t1);
}

// One sphere of Sphere::intersect_impl (sphere.cpp:62-106): is it a candidate, and at which distance?
__device__ __forceinline__ bool sphere_candidate(float cx, float cy, float cz, float radSq, const float o[3],
                                                 const float d[3], float &tm)
{
    float lx = cx - o[0];
    float ly = cy - o[1];
    float lz = cz - o[2];
    float distSq = dot3(lx, ly, lz, lx, ly, lz);
    if (!(distSq > radSq)) { // sphere.cpp:70: origin must be outside
        return false;
    }
    float tca = dot3(lx, ly, lz, d[0], d[1], d[2]);
    float tcaSq = tca * tca;
    float d2 = distSq - tcaSq;
    if (!(d2 < radSq)) { // sphere.cpp:88
        return false;
    }
    float thcSq = radSq - d2;
    float thc = sqrtf(thcSq);
    float t0 = tca - thc;
    float t1 = tca + thc;
    if (!(t0 >= 0.0f && t1 >= 0.0f)) { // sphere.cpp:103-106
        return false;
    }
    tm = t0 < t1 ? t0 : t1; // _mm256_min_ps operand rule
    return true;
}

// kind: DODRT_KIND_SPHERE or DODRT_KIND_BOX.  Result = lexicographic min of (t, id) over the candidates with
// t < clip (closest) or "any candidate with t < clip" (any-hit), exactly the brute-force answer.
template <uint32_t KIND>
__device__ __forceinline__ bool prim_bvh_query(const float4 *__restrict__ nodes, const uint32_t *__restrict__ ids,
                                               const float *__restrict__ lanes, const float o[3], const float d[3], bool any,
                                               float clip, Hit &hit, unsigned long long *stats = nullptr)
{
    // instrumentation (off unless dodrt_scene_debug_stats enabled it): nodes fetched / primitives tested by this ray
    struct Count {
        unsigned long long *stats;
        uint32_t nodes = 1, prims = 0;
        __device__ ~Count()
        {
            if (stats) {
                atomicAdd(stats + (KIND == DODRT_KIND_SPHERE ? 0 : 2), (unsigned long long)nodes);
                atomicAdd(stats + (KIND == DODRT_KIND_SPHERE ? 1 : 3), (unsigned long long)prims);
                atomicAdd(stats + 4, 1ull);
            }
        }
    } count{stats};
    const float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    // Near child first: both children of an interior node are tested when it is visited, the one the ray enters first is
    // descended, the other is stacked with its pruning key and dropped unread if a closer candidate turned up meanwhile.
    // The answer is the (t, id) minimum over the same candidate set whatever the order, so only the work changes.
    uint32_t stackNode[48];
    float stackKey[48];
    int sp = 0;
    float best = clip;
    uint32_t bestId = DODRT_MISS;
    float4 lo4 = __ldg(nodes), hi4 = __ldg(nodes + 1);
    bool live = prim_bvh_may_touch(lo4, hi4, o, d, inv, best);
    for (;;) {
        if (live) {
            const uint32_t a = __float_as_uint(lo4.w), b = __float_as_uint(hi4.w);
            if (b & kPrimBvhLeaf) {
                const uint32_t n = b & ~kPrimBvhLeaf;
                count.prims += n;
                for (uint32_t k = 0; k < n; k++) {
                    const uint32_t id = __ldg(ids + a + k);
                    const float *lane = lanes + (size_t)(id >> 3) * (KIND == DODRT_KIND_SPHERE ? 32 : 48);
                    const uint32_t j = id & 7u;
                    float t;
                    bool cand;
                    if (KIND == DODRT_KIND_SPHERE) {
                        cand = sphere_candidate(__ldg(lane + j), __ldg(lane + 8 + j), __ldg(lane + 16 + j), __ldg(lane + 24 + j),
                                                o, d, t);
                    } else { // box extension: slab arithmetic of box.cpp:33-53, hit distance = entry distance > 0
                        const float bmin[3] = {__ldg(lane + j), __ldg(lane + 8 + j), __ldg(lane + 16 + j)};
                        const float bmax[3] = {__ldg(lane + 24 + j), __ldg(lane + 32 + j), __ldg(lane + 40 + j)};
                        float tmax;
                        cand = slab(bmin, bmax, o, inv, clip, t, tmax) && t > 0.0f;
                    }
                    if (cand && t < clip && (bestId == DODRT_MISS || t < best || (t == best && id < bestId))) {
                        best = t;
                        bestId = id;
                        if (any) {
                            hit.t = t;
                            hit.prim = (KIND << DODRT_KIND_SHIFT) | id;
                            hit.u = hit.v = 0.0f;
                            return true;
                        }
                    }
                }
                live = false;
            } else {
                const float4 llo = __ldg(nodes + 2 * a), lhi = __ldg(nodes + 2 * a + 1);
                const float4 rlo = __ldg(nodes + 2 * b), rhi = __ldg(nodes + 2 * b + 1);
                count.nodes += 2;
                float keyL = 0.0f, keyR = 0.0f;
                const bool touchL = prim_bvh_may_touch_key(llo, lhi, o, d, inv, best, keyL);
                const bool touchR = prim_bvh_may_touch_key(rlo, rhi, o, d, inv, best, keyR);
                const bool leftFirst = !touchR || (touchL && !(keyR < keyL));
                if (touchL && touchR) {
                    stackNode[sp] = leftFirst ? b : a;
                    stackKey[sp] = leftFirst ? keyR : keyL;
                    ++sp;
                }
                live = touchL || touchR;
                lo4 = leftFirst ? llo : rlo;
                hi4 = leftFirst ? lhi : rhi;
            }
        }
        if (!live) {
            bool got = false;
            while (sp > 0 && !got) {
                --sp;
                if (!(stackKey[sp] > best)) { // still within reach of the best candidate so far
                    const uint32_t node = stackNode[sp];
                    lo4 = __ldg(nodes + 2 * node);
                    hi4 = __ldg(nodes + 2 * node + 1);
                    got = true;
                }
            }
            if (!got) {
                break;
            }
            live = true;
        }
    }
    if (bestId == DODRT_MISS) {
        return false;
    }
    hit.t = best;
    hit.prim = (KIND << DODRT_KIND_SHIFT) | bestId;
    hit.u = hit.v = 0.0f;
    return true;
}
#endif // __CUDACC__

} // namespace dodrt

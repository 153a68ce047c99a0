// dodrt_kernels.cu -- hand-written sm_100a traversal / intersection kernels.
//
// Compile with: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
// (-fmad=false is part of the numerical contract, see dodrt_device.cuh).
//
// All kernels are persistent: a warp claims 32 consecutive work items at a time from a global
// counter (one atomic per claim); in the frame modes 32 consecutive items are an 8x4 pixel block, so a
// warp's rays are coherent.  Every thread owns ONE ray, runs the analytic classes of the reference's
// query chain (Sphere -> [Box] -> Plane -> Cylinder) and then the kd-tree with a per-thread short stack.
// The tensor cores are idle by design: the path has no dense contraction.
//
// Kernel variants (TraceParams::variant, env DODRT_VARIANT, default = kDefaultVariant).  The product build holds 0 (plain
// baseline), 3 (default) and 7 (donating); the others were measured slower and are only compiled with -DDODRT_EXPERIMENTS
// (lib/libdodrt_cuda_exp.so, which the variant parity tests also run), as recorded A/B experiments:
//   0  per-thread traversal loop, exact reference-order triangle test                (round-1 baseline)
//   1  same loop, division-deferring triangle test (triangle_test_fast)
//   2  warp-voted traversal: the warp alternates between "one kd node step" and "one triangle lane
//      (8 triangles)" and always runs the phase the MAJORITY of its live rays is waiting for, so at
//      least half of the live lanes do useful work in every instruction (profiles/r01_*: the baseline
//      loop ran with 11.8 of 32 lanes active because rays at interior nodes waited for whole leaves).
//   3  variant 2 on the SoA lane layout (the reference's own 288-B lanes with AB/AC precomputed): 18 instead
//      of 24 LDG.128 per lane and a branch-free four-triangles-at-a-time first stage.
//   4  variant 3 plus warp-level regrouping: analytic classes + bounds test for 32 fresh rays at a time, the
//      survivors compacted (ballot/popc) into a per-warp pool in shared memory from which idle lanes are
//      refilled (dodrt_pool_kernel.inl).
//   5  variant 3 plus a cooperative leaf step for stragglers (<= 4 rays waiting: 8/16/32 lanes per ray).
//   6  variant 3 with the branch-free first stage on packed fp32 pairs (FMUL2/FADD2/FFMA2, dodrt_device.cuh): half
//      the fp32 issue slots for the same bits.
//   7  variant 3 plus ray donation at the tail of a pass (dodrt_donate.inl): warps that ran out of work resume,
//      32 lanes wide, the rays suspended by warps that are still traversing.
//   8  variant 3 with ONE copy of the exact triangle test, looped over the first stage's survivors (lane_test_compact).
// Every variant computes identical results (parity tests run all of them).
#include "dodrt_kernels.cuh"
#include "dodrt_prim_bvh.cuh"

#include <cstdlib>

namespace dodrt {

namespace {

constexpr float kInfinity = __builtin_huge_valf();

__device__ __forceinline__ float pick(const float v[3], uint32_t axis)
{
    return axis == 0 ? v[0] : (axis == 1 ? v[1] : v[2]);
}

template <bool FAST>
__device__ __forceinline__ bool tri_test(const float4 *__restrict__ tri, const float o[3], const float d[3], float maxDist,
                                         float &t, float &u, float &v)
{
    const float4 q0 = __ldg(tri), q1 = __ldg(tri + 1), q2 = __ldg(tri + 2);
    return FAST ? triangle_test_fast(q0, q1, q2, o, d, maxDist, t, u, v) : triangle_test(q0, q1, q2, o, d, maxDist, t, u, v);
}

// ---- variants 0/1: KDTree::intersect, kdtree.cpp:263-361, as one loop per thread -----------------------------
template <bool FAST>
__device__ __forceinline__ bool kdtree_query(const DeviceScene &s, const float o[3], const float d[3], bool any,
                                             float &clip, Hit &hit)
{
    if (s.num_nodes == 0) {
        return false;
    }
    const float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]}; // kdtree.cpp:271
    float tmin, tmax;
    if (!slab(s.bmin, s.bmax, o, inv, clip, tmin, tmax) || tmin > clip) { // kdtree.cpp:274
        return false;
    }
    uint32_t stackNode[kMaxStack];
    float stackTmin[kMaxStack];
    float stackTmax[kMaxStack];
    int sp = 0;
    uint32_t node = 0;
    bool found = false;
    for (;;) {
        if (clip < tmin) { // kdtree.cpp:286-289
            break;
        }
        const uint2 n = __ldg(&s.nodes[node]);
        if ((n.x & 3u) != kLeafFlag) {
            const uint32_t axis = n.x & 3u;
            const float split = __uint_as_float(n.y);
            const float oa = pick(o, axis);
            const float tPlane = (split - oa) * pick(inv, axis); // kdtree.cpp:293
            const bool leftFirst = (oa < split) || (oa == split && pick(d, axis) <= 0.0f); // kdtree.cpp:297-299
            const uint32_t below = node + 1, above = n.x >> 2;
            const uint32_t nearChild = leftFirst ? below : above;
            const uint32_t farChild = leftFirst ? above : below;
            if (tPlane > tmax || tPlane <= 0.0f) { // kdtree.cpp:312
                node = nearChild;
            } else if (tPlane < tmin) { // kdtree.cpp:316
                node = farChild;
            } else { // kdtree.cpp:320-329
                stackNode[sp] = farChild;
                stackTmin[sp] = tPlane;
                stackTmax[sp] = tmax;
                ++sp;
                node = nearChild;
                tmax = tPlane;
            }
        } else {
            const uint32_t numTris = (n.x >> 2) * kLane;
            const uint32_t firstTri = n.y * kLane;
            const float4 *tri = s.tris + (size_t)firstTri * 3;
            for (uint32_t k = 0; k < numTris; k++, tri += 3) {
                float t, u, v;
                if (tri_test<FAST>(tri, o, d, clip, t, u, v)) {
                    clip = t; // running maximumDistance, then kdtree.cpp:343
                    hit.t = t;
                    hit.prim = (DODRT_KIND_TRIANGLE << DODRT_KIND_SHIFT) | (firstTri + k);
                    hit.u = u;
                    hit.v = v;
                    found = true;
                    if (any) { // kdtree.cpp:338-341: only the boolean is defined for any-hit
                        return true;
                    }
                }
            }
            if (sp > 0) { // kdtree.cpp:347-357
                --sp;
                node = stackNode[sp];
                tmin = stackTmin[sp];
                tmax = stackTmax[sp];
            } else {
                break;
            }
        }
    }
    return found;
}

// ---- variant 2: the same traversal as a warp-voted state machine ---------------------------------------------
// Per-thread state: `live` (ray still traversing), a pending triangle range [triCur, triEnd) of the leaf
// it is in, the node to visit next, the parametric interval, and the stack.  The reference's order of
// events per ray is untouched (nodes front to back, triangles of a leaf in id order, strict `<`), only
// the interleaving ACROSS the rays of a warp changes.
struct TreeState {
    float inv[3];
    float tmin, tmax;
    uint32_t node;
    uint32_t triCur, triEnd;
    int sp;
    bool live;
};

__device__ __forceinline__ void tree_pop(TreeState &st, const uint32_t *stackNode, const float *stackTmin,
                                         const float *stackTmax)
{
    if (st.sp > 0) { // kdtree.cpp:347-357
        --st.sp;
        st.node = stackNode[st.sp];
        st.tmin = stackTmin[st.sp];
        st.tmax = stackTmax[st.sp];
    } else {
        st.live = false;
    }
}

// One triangle lane (8 consecutive slots, triangle.cpp:43-140) of the leaf this ray is in.
template <bool SOA, bool PACKED, bool COMPACT = false>
__device__ __forceinline__ void leaf_step(const DeviceScene &s, TreeState &st, const float o[3], const float d[3], bool any,
                                          float &clip, Hit &hit, bool &found, const uint32_t *stackNode,
                                          const float *stackTmin, const float *stackTmax)
{
    if (COMPACT) {
        const float4 *lane = s.lanes4 + (size_t)(st.triCur >> 3) * 18;
        if (lane_test_compact(lane, st.triCur, o, d, clip, hit)) found = true;
    } else if (PACKED) {
        const float4 *lane = s.lanes4 + (size_t)(st.triCur >> 3) * 18;
        const f32x2 o2[3] = {f2_pack(o[0], o[0]), f2_pack(o[1], o[1]), f2_pack(o[2], o[2])};
        const f32x2 d2[3] = {f2_pack(d[0], d[0]), f2_pack(d[1], d[1]), f2_pack(d[2], d[2])};
        if (lane_half_test_packed(lane, 0, st.triCur, o, d, o2, d2, s.negzero2, clip, hit)) found = true;
        if (!(any && found) && lane_half_test_packed(lane, 1, st.triCur + 4, o, d, o2, d2, s.negzero2, clip, hit)) found = true;
    } else if (SOA) {
        const float4 *lane = s.lanes4 + (size_t)(st.triCur >> 3) * 18;
        if (lane_half_test(lane, 0, st.triCur, o, d, clip, hit)) found = true;
        if (!(any && found) && lane_half_test(lane, 1, st.triCur + 4, o, d, clip, hit)) found = true;
    } else {
        const float4 *tri = s.tris + (size_t)st.triCur * 3;
#pragma unroll 4
        for (uint32_t k = 0; k < (uint32_t)kLane; k++, tri += 3) {
            float t, u, v;
            if (tri_test<true>(tri, o, d, clip, t, u, v)) {
                clip = t;
                hit.t = t;
                hit.prim = (DODRT_KIND_TRIANGLE << DODRT_KIND_SHIFT) | (st.triCur + k);
                hit.u = u;
                hit.v = v;
                found = true;
            }
        }
    }
    st.triCur += kLane;
    if (any && found) {
        st.live = false; // kdtree.cpp:338-341
    } else if (st.triCur == st.triEnd) {
        tree_pop(st, stackNode, stackTmin, stackTmax);
    }
}

// One kd node visit (one iteration of the reference's while loop, kdtree.cpp:284-333,347-357).
__device__ __forceinline__ void node_step(const DeviceScene &s, TreeState &st, const float o[3], const float d[3],
                                          float clip, uint32_t *stackNode, float *stackTmin, float *stackTmax)
{
    if (clip < st.tmin) { // kdtree.cpp:286-289
        st.live = false;
        return;
    }
    const uint2 n = __ldg(&s.nodes[st.node]);
    if ((n.x & 3u) != kLeafFlag) {
        const uint32_t axis = n.x & 3u;
        const float split = __uint_as_float(n.y);
        const float oa = pick(o, axis);
        const float tPlane = (split - oa) * pick(st.inv, axis); // kdtree.cpp:293
        const bool leftFirst = (oa < split) || (oa == split && pick(d, axis) <= 0.0f); // kdtree.cpp:297-299
        const uint32_t below = st.node + 1, above = n.x >> 2;
        const uint32_t nearChild = leftFirst ? below : above;
        const uint32_t farChild = leftFirst ? above : below;
        if (tPlane > st.tmax || tPlane <= 0.0f) { // kdtree.cpp:312
            st.node = nearChild;
        } else if (tPlane < st.tmin) { // kdtree.cpp:316
            st.node = farChild;
        } else { // kdtree.cpp:320-329
            stackNode[st.sp] = farChild;
            stackTmin[st.sp] = tPlane;
            stackTmax[st.sp] = st.tmax;
            ++st.sp;
            st.node = nearChild;
            st.tmax = tPlane;
        }
    } else {
        const uint32_t numTris = (n.x >> 2) * kLane;
        if (numTris == 0) {
            tree_pop(st, stackNode, stackTmin, stackTmax);
        } else {
            st.triCur = n.y * kLane;
            st.triEnd = st.triCur + numTris;
        }
    }
}

__device__ __forceinline__ void tree_enter(const DeviceScene &s, TreeState &st, bool enter, const float o[3], const float d[3],
                                           float clip)
{
    st.live = false;
    st.triCur = st.triEnd = 0;
    st.sp = 0;
    st.node = 0;
    st.tmin = st.tmax = 0.0f;
    st.inv[0] = st.inv[1] = st.inv[2] = 0.0f;
    if (enter && s.num_nodes != 0) {
        st.inv[0] = 1.0f / d[0]; // kdtree.cpp:271
        st.inv[1] = 1.0f / d[1];
        st.inv[2] = 1.0f / d[2];
        st.live = slab(s.bmin, s.bmax, o, st.inv, clip, st.tmin, st.tmax) && !(st.tmin > clip); // kdtree.cpp:274
    }
}

// Leaf phase for STRAGGLERS: at most 4 rays of the warp wait for triangles (the others are finished or at
// nodes).  One ray per lane would leave >= 28 lanes idle, and a ray grazing the mesh visits >1000 triangle lanes
// one after the other -- the critical path that kept single SMs busy long after the rest of the grid had
// drained (profiles/r01_rank_tail.txt).  Here the warp splits into groups of G = 8/16/32 lanes, group g serves
// the g-th waiting ray, and every lane tests ONE of the ray's next G triangle slots (ray by shuffle, 9 scalar
// loads from the SoA lane).  The group's answer is the lexicographic minimum of (t, slot) over the accepted
// slots, which is exactly what the reference's slot-by-slot loop with its strict `<` leaves behind
// (triangle.cpp:119-139), so results stay bit-identical.
__device__ __forceinline__ void leaf_step_coop(const DeviceScene &s, TreeState &st, unsigned leafMask, uint32_t nLeaf,
                                               bool wantLeaf, const float o[3], const float d[3], bool any, float &clip,
                                               Hit &hit, bool &found, const uint32_t *stackNode, const float *stackTmin,
                                               const float *stackTmax)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t shift = nLeaf <= 1u ? 5u : (nLeaf <= 2u ? 4u : 3u);
    const uint32_t G = 1u << shift;
    const uint32_t g = lane >> shift, sub = lane & (G - 1u);
    unsigned m = leafMask;
    for (uint32_t i = 0; i < g; i++) {
        m &= m - 1u; // drop the g lowest waiting lanes (g <= 3)
    }
    const bool groupValid = m != 0u;
    const uint32_t owner = groupValid ? (uint32_t)__ffs((int)m) - 1u : lane;
    float ro[3], rd[3];
    ro[0] = __shfl_sync(0xffffffffu, o[0], owner);
    ro[1] = __shfl_sync(0xffffffffu, o[1], owner);
    ro[2] = __shfl_sync(0xffffffffu, o[2], owner);
    rd[0] = __shfl_sync(0xffffffffu, d[0], owner);
    rd[1] = __shfl_sync(0xffffffffu, d[1], owner);
    rd[2] = __shfl_sync(0xffffffffu, d[2], owner);
    const float rclip = __shfl_sync(0xffffffffu, clip, owner);
    const uint32_t rcur = __shfl_sync(0xffffffffu, st.triCur, owner);
    const uint32_t rend = __shfl_sync(0xffffffffu, st.triEnd, owner);
    const uint32_t tri = rcur + sub;
    bool acc = false;
    float t = 0.0f, u = 0.0f, v = 0.0f;
    if (groupValid && tri < rend) {
        const float *base = reinterpret_cast<const float *>(s.lanes4) + (size_t)(tri >> 3) * 72 + (tri & 7u);
        const float4 q0 = make_float4(__ldg(base), __ldg(base + 8), __ldg(base + 16), __ldg(base + 24));
        const float4 q1 = make_float4(__ldg(base + 32), __ldg(base + 40), __ldg(base + 48), __ldg(base + 56));
        const float4 q2 = make_float4(__ldg(base + 64), 0.0f, 0.0f, 0.0f);
        acc = triangle_test_fast(q0, q1, q2, ro, rd, rclip, t, u, v);
    }
    float best = acc ? t : kInfinity; // accepted t are finite and positive, so min() is exact
    for (uint32_t off = G >> 1; off != 0u; off >>= 1) {
        best = fminf(best, __shfl_xor_sync(0xffffffffu, best, off)); // xor < G stays inside the group
    }
    const unsigned winners = __ballot_sync(0xffffffffu, acc && t == best);
    // owners pick up their group's answer: lowest winning lane = lowest slot id
    const uint32_t myGroup = __popc(leafMask & ((1u << lane) - 1u));
    const unsigned groupLanes = (G == 32u ? 0xffffffffu : ((1u << G) - 1u)) << ((myGroup << shift) & 31u);
    const unsigned mine = wantLeaf ? (winners & groupLanes) : 0u;
    const uint32_t src = mine ? (uint32_t)__ffs((int)mine) - 1u : lane;
    const float wt = __shfl_sync(0xffffffffu, t, src);
    const float wu = __shfl_sync(0xffffffffu, u, src);
    const float wv = __shfl_sync(0xffffffffu, v, src);
    if (wantLeaf) {
        if (mine) {
            clip = wt;
            hit.t = wt;
            hit.prim = (DODRT_KIND_TRIANGLE << DODRT_KIND_SHIFT) | (st.triCur + (src & (G - 1u)));
            hit.u = wu;
            hit.v = wv;
            found = true;
        }
        st.triCur = st.triCur + G < st.triEnd ? st.triCur + G : st.triEnd;
        if (any && found) {
            st.live = false; // kdtree.cpp:338-341
        } else if (st.triCur == st.triEnd) {
            tree_pop(st, stackNode, stackTmin, stackTmax);
        }
    }
}

#ifdef DODRT_TIMELINE
// debug build only: per-warp exit times and per-donation (time, rays given, live rays kept) records of the LAST launch
__device__ unsigned long long g_tlExit[8192];
__device__ unsigned long long g_tlDonation[1 << 16];
__device__ unsigned int g_tlDonations;
__device__ unsigned int g_tlPollsEmpty;
#endif

#include "dodrt_donate.inl"

template <bool SOA, bool SHARE, bool PACKED, bool DONATE, bool COMPACT = false>
__device__ __forceinline__ bool kdtree_query_voted(const DeviceScene &s, bool enter, const float o[3], const float d[3],
                                                   bool any, float &clip, Hit &hit, const TraceParams *p = nullptr,
                                                   const Finish *fin = nullptr, bool *donated = nullptr, bool allowDonate = true)
{
    TreeState st;
    tree_enter(s, st, enter, o, d, clip);
    uint32_t stackNode[kMaxStack];
    float stackTmin[kMaxStack];
    float stackTmax[kMaxStack];
    bool found = false;
    uint32_t nodeRun = 0;
    // The voted loop proper contains no atomics, volatile loads or calls; with DONATE it is left every s.donate_poll
    // iterations for the poll / suspension below and re-entered (ptxas puts a YIELD at the head of a loop that
    // contains the poll, which cost the shadow pass 17 %).
    for (;;) {
        bool finished = false;
        uint32_t budget = s.donate_poll;
        for (;;) {
            const bool wantLeaf = st.live && st.triCur < st.triEnd;
            const bool wantNode = st.live && !wantLeaf;
            const unsigned leafMask = __ballot_sync(0xffffffffu, wantLeaf);
            const unsigned nodeMask = __ballot_sync(0xffffffffu, wantNode);
            if ((leafMask | nodeMask) == 0u) {
                finished = true;
                break;
            }
            const uint32_t nLeaf = __popc(leafMask), nNode = __popc(nodeMask);
            if (nNode == 0u || (nLeaf != 0u && (nLeaf * s.tune[1] >= nNode * s.tune[0] || nodeRun >= s.tune[2]))) {
                nodeRun = 0;
                if (SHARE && nLeaf <= 4u) {
                    leaf_step_coop(s, st, leafMask, nLeaf, wantLeaf, o, d, any, clip, hit, found, stackNode, stackTmin,
                                   stackTmax);
                } else if (wantLeaf) {
                    leaf_step<SOA, PACKED, COMPACT>(s, st, o, d, any, clip, hit, found, stackNode, stackTmin, stackTmax);
                }
            } else {
                ++nodeRun;
                if (wantNode) {
                    node_step(s, st, o, d, clip, stackNode, stackTmin, stackTmax);
                    // a ray that is still between leaves may go on: the vote (two ballots, the phase rule) costs
                    // about as much as a node step
                    for (uint32_t k = 1; k < s.node_burst && st.live && !(st.triCur < st.triEnd); k++) {
                        node_step(s, st, o, d, clip, stackNode, stackTmin, stackTmax);
                    }
                }
            }
            if (DONATE && --budget == 0u) {
                break;
            }
        }
        if (!DONATE || finished) {
            break;
        }
        // do idle warps wait for rays?
        const uint32_t want = (p->donate_slots != nullptr && allowDonate) ? donate_poll(*p) : 0u;
#ifdef DODRT_TIMELINE
        if (want == 0u && (threadIdx.x & 31u) == 0u) atomicAdd(&g_tlPollsEmpty, 1u);
#endif
        if (want != 0u) {
            donate_live_rays(*p, want, st, o, d, any, clip, hit, found, fin, stackNode, stackTmin, stackTmax, *donated);
        }
    }
    return found;
}

// The analytic part of the query chain: closest hit main.cpp:312-319, any hit main.cpp:198-208.
// Returns true when the any-hit query is already decided.
__device__ __forceinline__ bool analytic_chain(const DeviceScene &s, uint32_t classes, const float o[3],
                                               const float d[3], bool any, float &clip, Hit &hit, bool &found)
{
    Hit h;
    // The culling BVHs are only proven equivalent to the reference's brute force for unit-length directions
    // (dodrt_prim_bvh.cuh: the bound on the sphere test's d2 assumes |D| = 1; with |D|^2 = 1 + delta the accepted
    // line distance grows to r + sqrt(1e-6 + |delta|) |L|, which the node padding covers for |delta| <= 4e-6).
    // dodrt_ray.d need not be normalised, so any other ray takes the reference's own loop over every lane.
    const bool unitDir = fabsf(dot3(d[0], d[1], d[2], d[0], d[1], d[2]) - 1.0f) <= 2e-6f;
    if ((classes & DODRT_CLS_SPHERE) && s.num_spheres &&
        ((s.sphere_bvh && unitDir) ? prim_bvh_query<DODRT_KIND_SPHERE>(s.sphere_bvh, s.sphere_bvh_ids, s.sphere_lanes, o, d, any, clip, h, s.stats)
                                   : sphere_query(s, o, d, any, clip, h))) {
        hit = h;
        found = true;
        if (any) return true;
        clip = h.t;
    }
    if ((classes & DODRT_CLS_BOX) && s.num_boxes &&
        ((s.box_bvh && unitDir) ? prim_bvh_query<DODRT_KIND_BOX>(s.box_bvh, s.box_bvh_ids, s.box_lanes, o, d, any, clip, h, s.stats)
                                : box_query(s, o, d, any, clip, h))) {
        hit = h;
        found = true;
        if (any) return true;
        clip = h.t;
    }
    if ((classes & DODRT_CLS_PLANE) && s.num_planes && plane_query(s, o, d, clip, h)) {
        hit = h;
        found = true;
        if (any) return true;
        clip = h.t;
    }
    if ((classes & DODRT_CLS_CYLINDER) && s.num_cylinders && cylinder_query(s, o, d, clip, h)) {
        hit = h;
        found = true;
        if (any) return true;
        clip = h.t;
    }
    return false;
}

// One full query for this thread's ray.  `valid` = the thread has a ray at all; in the voted variant every
// lane of the warp must call this (the traversal is warp-synchronous).
// `p`, `kind`, `out` (variant 7 only): how the ray's result is written, so that a ray suspended into the donation
// queue can be finished by another warp; `*donated` is set when that happened (the caller must not write).
template <int VARIANT>
__device__ __forceinline__ bool query(const DeviceScene &s, uint32_t classes, bool valid, const float o[3],
                                      const float d[3], bool any, float clip, Hit &hit, const TraceParams *p = nullptr,
                                      uint32_t kind = 0, uint64_t out = 0, bool *donated = nullptr, uint64_t mirror = kNoMirror,
                                      bool allowDonate = true)
{
    bool found = false;
    const float clip0 = clip;
    hit.t = clip;
    hit.prim = DODRT_MISS;
    hit.u = hit.v = 0.0f;
    bool decided = !valid;
    if (valid) {
        decided = analytic_chain(s, classes, o, d, any, clip, hit, found);
    }
    const bool enter = !decided && (classes & DODRT_CLS_TREE);
    Hit h;
    if (VARIANT == kDonateVariant) {
        Finish fin;
        fin.out = out;
        fin.mirror = mirror;
        fin.kind = kind;
        fin.pre[0] = kind == kFinishAnyRecord ? clip0 : hit.t;
        fin.pre[1] = __uint_as_float(hit.prim);
        fin.pre[2] = hit.u;
        fin.pre[3] = hit.v;
        h.t = clip, h.prim = DODRT_MISS, h.u = h.v = 0.0f;
        if (kdtree_query_voted<true, false, false, true>(s, enter, o, d, any, clip, h, p, &fin, donated, allowDonate)) {
            hit = h;
            found = true;
        }
    } else if (VARIANT >= 2) {
        if (kdtree_query_voted<VARIANT >= 3, VARIANT == 5, VARIANT == 6, false, VARIANT == 8>(s, enter, o, d, any, clip, h)) {
            hit = h;
            found = true;
        }
    } else if (enter && kdtree_query<VARIANT == 1>(s, o, d, any, clip, h)) {
        hit = h;
        found = true;
    }
    return found;
}

// Heavy-first tile order.  A persistent kernel ends when its slowest warp ends, and a 32-ray batch that grazes
// the mesh runs for hundreds of microseconds; started in natural (top-to-bottom) order such batches begin
// half-way through the kernel and the other SMs idle while they finish (profiles/r01_rank_tail.txt).  One thread
// per local tile samples 4x4 of the tile's rays (the primary ray, or the pixel's shadow ray built from its
// primary hit) against the kd-tree bounds; tiles with a ray entering the bounds are queued from the front,
// the others from the back.  Only the processing order changes -- never a result.
template <int MODE> __global__ void order_tiles_kernel(const TraceParams p)
{
    const uint32_t local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= p.num_local_tiles) {
        return;
    }
    const dodrt_frame &f = p.frame;
    const uint32_t tile = f.first_tile + local * f.tile_stride;
    const uint32_t tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const float o[3] = {f.origin[0], f.origin[1], f.origin[2]};
    bool heavy = false;
    for (uint32_t sy = 0; sy < 4 && !heavy; sy++) {
        for (uint32_t sx = 0; sx < 4 && !heavy; sx++) {
            const uint32_t ix = (2 * sx + 1) * f.tile_w / 8, iy = (2 * sy + 1) * f.tile_h / 8;
            const uint32_t col = tx * f.tile_w + ix, row = ty * f.tile_h + iy;
            if (col >= f.width || row >= f.height) {
                continue;
            }
            float ro[3] = {o[0], o[1], o[2]}, rd[3];
            float clip = kInfinity;
            primary_dir(__ldg(p.xs + col), __ldg(p.ys + row), rd);
            if (MODE == kModeShadow) {
                const uint32_t in = ((iy >> 2) * (f.tile_w >> 3) + (ix >> 3)) * 32u + (iy & 3u) * 8u + (ix & 7u);
                const uint64_t idx = f.compact ? (uint64_t)local * f.tile_w * f.tile_h + in : (uint64_t)row * f.width + col;
                const float4 ph = reinterpret_cast<const float4 *>(p.hits)[idx];
                if (__float_as_uint(ph.y) == DODRT_MISS) {
                    continue;
                }
                float so[3], sd[3];
                shadow_ray(o, rd, ph.x, p.light, so, sd, clip);
                ro[0] = so[0], ro[1] = so[1], ro[2] = so[2];
                rd[0] = sd[0], rd[1] = sd[1], rd[2] = sd[2];
            }
            const float inv[3] = {1.0f / rd[0], 1.0f / rd[1], 1.0f / rd[2]};
            float tmin, tmax;
            heavy = slab(p.scene.bmin, p.scene.bmax, ro, inv, clip, tmin, tmax) && !(tmin > clip);
        }
    }
    if (heavy) {
        p.tile_order[atomicAdd(p.counter + 1, 1ull)] = local;
    } else {
        p.tile_order[p.num_local_tiles - 1u - (uint32_t)atomicAdd(p.counter + 2, 1ull)] = local;
    }
}

// Resident blocks per SM the register allocation aims at.  The plain kernels fit 5 blocks (93-96 registers) on their
// own; capped to 6 blocks (80 registers, 32-112 B of spills outside the leaf loop) the shadow pass is 4 % faster
// (dragon4k 4415 -> 4550 Mrays/s), at 7 / 8 blocks the spills win (4184 / 3529).  Measured with NVVM's
// rematerialisation off (see the Makefile); with it on, the cap made things worse.  The donating kernel stays at 5:
// at 6 it is 2 % faster on a whole frame but slower on the short passes it exists for (0.754 vs 0.735 ms per rank of 8).
#ifndef DODRT_MINBLOCKS
#define DODRT_MINBLOCKS 6
#endif
template <int MODE, int VARIANT>
__global__ void __launch_bounds__(128, VARIANT >= kDonateVariant ? 5 : DODRT_MINBLOCKS) trace_kernel(const TraceParams p)
{
    const uint32_t lane = threadIdx.x & 31u;
    if (VARIANT == kDonateVariant && p.donate_slots != nullptr && lane == 0) {
        atomicAdd(p.counter + kDonateStarted, 1ull); // see donate_helper_loop
    }
#ifdef DODRT_TIMELINE
    if (lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(p.counter + 24, t); // kernel start
    }
#endif
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(p.counter, 32ull);
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= p.count) {
            break;
        }
        const uint64_t item = base + lane;
        const bool inRange = item < p.count;
        if (MODE == kModeShadowRays) {
            // canSeeLight (main.cpp:182-219) for an explicit ray batch: shadow ray from rays[i]'s hit point
            // (hits[i]) to p.light; visible[i] = 1 iff there was a hit and nothing blocks the light
            float so[3] = {0.0f, 0.0f, 0.0f}, sd[3] = {0.0f, 0.0f, 1.0f}, sclip = 0.0f;
            bool cast = false;
            uint64_t outItem = item;
            if (inRange) {
                // all lights of a bounce in one launch: item = light * rays_per_light + ray
                uint64_t ray = item;
                float light3[3] = {p.light[0], p.light[1], p.light[2]};
                uint64_t lightBase = 0;
                if (p.num_lights > 1u) {
                    const uint32_t l = (uint32_t)(item / p.rays_per_light);
                    lightBase = (uint64_t)l * p.rays_per_light;
                    ray = item - lightBase;
                    light3[0] = p.lights[l][0], light3[1] = p.lights[l][1], light3[2] = p.lights[l][2];
                }
                if (p.ray_order != nullptr) { // spatially sorted processing order; the result stays at the ray's own index
                    ray = __ldg(p.ray_order + ray);
                    outItem = lightBase + ray;
                }
                const float4 *src = reinterpret_cast<const float4 *>(p.rays + ray);
                const float4 a = src[0], b = src[1];
                const float4 ph = reinterpret_cast<const float4 *>(p.hits)[ray];
                if (!(__float_as_uint(b.w) & DODRT_RAY_SKIP) && __float_as_uint(ph.y) != DODRT_MISS) {
                    const float o[3] = {a.x, a.y, a.z}, d[3] = {a.w, b.x, b.y};
                    shadow_ray(o, d, ph.x, light3, so, sd, sclip);
                    cast = true;
                }
            }
            Hit h;
            bool donated = false;
            const bool blocked = query<VARIANT>(p.scene, p.classes, cast, so, sd, true, sclip, h, &p, kFinishVisible, outItem, &donated);
            if (inRange && !donated) {
                p.visible[outItem] = (cast && !blocked) ? 1 : 0;
            }
        } else if (MODE == kModeRays) {
            float o[3] = {0.0f, 0.0f, 0.0f}, d[3] = {0.0f, 0.0f, 1.0f};
            float clip = 0.0f;
            bool any = false, skip = false;
            // optional processing order (dodrt_render: the cells of the previous bounce's hit points = these rays' origins)
            const uint64_t idx = (inRange && p.ray_order != nullptr) ? (uint64_t)__ldg(p.ray_order + item) : item;
            if (inRange) {
                const float4 *src = reinterpret_cast<const float4 *>(p.rays + idx);
                const float4 a = __ldg(src), b = __ldg(src + 1);
                o[0] = a.x, o[1] = a.y, o[2] = a.z;
                d[0] = a.w, d[1] = b.x, d[2] = b.y;
                clip = b.z;
                any = (__float_as_uint(b.w) & DODRT_RAY_ANY) != 0u;
                skip = (__float_as_uint(b.w) & DODRT_RAY_SKIP) != 0u;
            }
            Hit h;
            bool donated = false;
            const bool found = query<VARIANT>(p.scene, p.classes, inRange && !skip, o, d, any, clip, h, &p,
                                              any ? kFinishAnyRecord : kFinishRecord, idx, &donated);
            if (inRange && !donated) {
                if (any) { // any-hit defines only hit/miss
                    h.t = clip;
                    h.prim = found ? 0u : DODRT_MISS;
                    h.u = h.v = 0.0f;
                }
                reinterpret_cast<float4 *>(p.hits)[idx] = make_float4(h.t, __uint_as_float(h.prim), h.u, h.v);
            }
        } else {
            uint32_t col = 0, row = 0;
            uint64_t slot = 0;
            const bool inside = inRange && slot_to_pixel(p.frame, p.tiles_x, p.tile_order, item, col, row, slot);
            const uint64_t out = p.frame.compact ? slot : (uint64_t)row * p.frame.width + col;
            // where the result goes in the mirror (a peer GPU's frame or pinned host memory), if there is one
            // (a mirror with the local layout also receives the padded slots of edge tiles, like the local buffer)
            const uint64_t mirrorIdx = inside ? (p.mirror_by_pixel ? (uint64_t)row * p.frame.width + col : out)
                                              : ((inRange && p.frame.compact && !p.mirror_by_pixel) ? out : kNoMirror);
            const uint64_t mirror = p.mirror_hits != nullptr ? mirrorIdx : kNoMirror;
            const float o[3] = {p.frame.origin[0], p.frame.origin[1], p.frame.origin[2]};
            float d[3] = {0.0f, 0.0f, 1.0f};
            if (inside) {
                primary_dir(__ldg(p.xs + col), __ldg(p.ys + row), d);
            }
            if (MODE == kModePrimary) {
                Hit h;
                bool donated = false;
                query<VARIANT>(p.scene, p.classes, inside, o, d, false, kInfinity, h, &p, kFinishRecord, out, &donated, mirror);
                if ((inside || (inRange && p.frame.compact)) && !donated) { // padded slots of edge tiles read as "miss"
                    const float4 r = make_float4(h.t, __uint_as_float(h.prim), h.u, h.v);
                    reinterpret_cast<float4 *>(p.hits)[out] = r;
                    if (mirror != kNoMirror) {
                        reinterpret_cast<float4 *>(p.mirror_hits)[mirror] = r;
                    }
                }
            } else {
                bool shadowed = true, cast = false;
                float so[3] = {0.0f, 0.0f, 0.0f}, sd[3] = {0.0f, 0.0f, 1.0f}, sclip = 0.0f;
                if (inside) {
                    const float4 ph = reinterpret_cast<const float4 *>(p.hits)[out];
                    if (__float_as_uint(ph.y) != DODRT_MISS) {
                        shadow_ray(o, d, ph.x, p.light, so, sd, sclip);
                        cast = true;
                    }
                }
                Hit h;
                bool donated = false;
                const uint64_t vmirror = p.mirror_visible != nullptr ? mirrorIdx : kNoMirror;
                const bool blocked = query<VARIANT>(p.scene, p.classes, cast, so, sd, true, sclip, h, &p, kFinishVisible, out,
                                                    &donated, vmirror);
                shadowed = !cast || blocked;
                if ((inside || (inRange && p.frame.compact)) && !donated) {
                    p.visible[out] = shadowed ? 0 : 1;
                    if (vmirror != kNoMirror) {
                        p.mirror_visible[vmirror] = shadowed ? 0 : 1;
                    }
                }
            }
        }
    }
#ifdef DODRT_TIMELINE
    if (lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(p.counter + 25, t); // first warp out of work
        atomicMax(p.counter + 26, t); // last warp out of its main loop
        atomicAdd(p.counter + 22, t & 0xFFFFFFFFFull); // (sum over warps, low 36 bits: the average time a warp leaves its main loop)
        atomicAdd(p.counter + 21, 1ull);
        g_tlExit[(blockIdx.x * blockDim.x + threadIdx.x) >> 5 & 8191u] = t;
    }
#endif
#ifndef DBG_NOHELPER
    if (VARIANT == kDonateVariant && p.donate_slots != nullptr) {
        donate_helper_loop(p);
    }
#endif
#ifdef DODRT_TIMELINE
    if (lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMax(p.counter + 27, t); // last helper gone
    }
#endif
}


#ifdef DODRT_EXPERIMENTS
// ---- kModeFrame: primary + shadow rays of a frame share in ONE persistent launch ----------------------------------
// Two work queues.  Queue 1 holds the primary batches (8x4 pixel blocks, tiles in heavy-first order); queue 2 holds the
// shadow batches -- still a separate, coherent any-hit pass over 8x4 blocks (main.cpp:182-219 per pixel), but a tile's
// shadow batches become claimable as soon as ITS primary records are complete instead of after the whole primary pass:
//   * the warp that finishes the last primary batch of a tile (tile_done[t] reaches batches-per-tile) classifies the tile
//     (heavy: a shadow ray enters the kd-tree bounds) and appends it to the heavy or the light ready list (publish_tile);
//   * a warp that has run out of primary batches claims shadow batches of published tiles, heavy list first
//     (claim_shadow_batch); when nothing is claimable yet it waits: the missing tiles' primary batches are in flight on
//     resident warps.
// The tail of the primary queue (its slowest batches) therefore overlaps shadow work, and the pass has ONE ramp and ONE
// tail instead of two of each with a grid-wide barrier in between -- which is what limited an 8-way split of a 4K frame
// (0.26 + 0.53 ms per rank against 0.47 ms ideal).  Per pixel nothing changes: same primary query, same shadow ray from
// the same hit record (read back from memory, like the separate pass does), same any-hit query.
// Donation (VARIANT 7): only shadow rays are ever suspended.  A helper exists once the shadow queue is exhausted, and the
// last shadow claims wait for the last tiles, whose primary batches may still run: those must not give rays away (a
// tile would be published while a record is still pending), hence allowDonate = false in queue 1.
// One work item of trace_frame_kernel: SHADOW = false: a primary batch of queue 1; true: a shadow batch of queue 2.
// Returns false when the queue is exhausted.
template <int VARIANT, bool SHADOW> __device__ __forceinline__ bool frame_batch(const TraceParams &p)
{
    const uint32_t lane = threadIdx.x & 31u;
    const dodrt_frame &f = p.frame;
    const uint32_t tilePixels = f.tile_w * f.tile_h;
    const uint32_t bpt = tilePixels >> 5; // batches per tile
    uint32_t col = 0, row = 0, localTile = 0, light = 0;
    uint64_t slot = 0;
    bool inside;
    if (!SHADOW) {
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(p.counter, 32ull);
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= p.count) {
            return false;
        }
        inside = slot_to_pixel(f, p.tiles_x, p.tile_order, base + lane, col, row, slot);
        localTile = (uint32_t)(slot / tilePixels);
    } else {
        // the batch's tile: the (seq / batches-per-tile / lights)-th that became complete; lights one after the other
        const uint32_t perTile = bpt * p.num_lights;
        const uint32_t bpr = f.tile_w >> 3; // 8x4 blocks per tile row
        unsigned long long base = 0;
        if (lane == 0) {
            base = atomicAdd(p.counter + kShadowNext, 32ull);
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= p.shadow_count) {
            return false;
        }
        const uint64_t seq = base >> 5;
        const uint32_t entry = (uint32_t)(seq / perTile);
        const uint32_t rem = (uint32_t)(seq - (uint64_t)entry * perTile);
        light = rem / bpt;
        const uint32_t block = rem - light * bpt;
        if (lane == 0) { // not published yet: the tile's primary batches are in flight on resident warps
            unsigned ns = 100;
            while ((localTile = *reinterpret_cast<const volatile uint32_t *>(p.ready_queue + entry)) == 0u) {
                __nanosleep(ns);
                ns = ns < 1600u ? ns * 2u : ns;
            }
        }
        localTile = __shfl_sync(0xffffffffu, localTile, 0) - 1u;
        __threadfence(); // acquire side of the tile's publication
        const uint32_t tile = f.first_tile + localTile * f.tile_stride;
        const uint32_t tx = tile % p.tiles_x, ty = tile / p.tiles_x;
        const uint32_t bx = block % bpr, by = block / bpr;
        col = tx * f.tile_w + bx * 8 + (lane & 7u);
        row = ty * f.tile_h + by * 4 + (lane >> 3);
        inside = col < f.width && row < f.height;
        slot = (uint64_t)localTile * tilePixels + block * 32u + lane;
    }
    const uint64_t pixel = (uint64_t)row * f.width + col;
    const uint64_t idx = f.compact ? slot : pixel;
    const uint64_t midx = p.mirror_by_pixel ? pixel : idx;
    // (a mirror with the local layout also receives the padded slots of edge tiles, like the local buffer)
    const bool mirrored = inside || (f.compact && !p.mirror_by_pixel);
    const float o[3] = {f.origin[0], f.origin[1], f.origin[2]};
    float d[3] = {0.0f, 0.0f, 1.0f};
    if (inside) {
        primary_dir(__ldg(p.xs + col), __ldg(p.ys + row), d);
    }
    Hit h;
    bool donated = false;
    if (!SHADOW) {
        const uint64_t mirror = (p.mirror_hits != nullptr && mirrored) ? midx : kNoMirror;
        // Rays of queue 1 are only given away when there is no queue 2 (a primary-only frame): a tile must not be
        // published while one of its records is still pending in the donation queue.  (With a shadow queue helpers
        // only exist once every tile has been published anyway.)
        query<VARIANT>(p.scene, p.classes, inside, o, d, false, kInfinity, h, &p, kFinishRecord, idx, &donated, mirror,
                       p.shadow_count == 0);
        if ((inside || f.compact) && !donated) { // padded slots of edge tiles read as "miss"
            const float4 r = make_float4(h.t, __uint_as_float(h.prim), h.u, h.v);
            reinterpret_cast<float4 *>(p.hits)[idx] = r;
            if (mirror != kNoMirror) {
                reinterpret_cast<float4 *>(p.mirror_hits)[mirror] = r;
            }
        }
        if (p.shadow_count != 0) {
            // every lane's record must be visible before the tile can be counted complete
            __threadfence();
            __syncwarp();
            if (lane == 0 && atomicAdd(p.tile_done + localTile, 1u) + 1u == bpt) {
                __threadfence(); // the other batches' records (released by their atomicAdd) before the publication
                const unsigned long long k = atomicAdd(p.counter + kReadyTail, 1ull);
                *reinterpret_cast<volatile uint32_t *>(p.ready_queue + k) = localTile + 1u;
            }
        }
    } else {
        const uint64_t out = (uint64_t)light * p.visible_light_stride + idx;
        const uint64_t mirror = (p.mirror_visible != nullptr && mirrored) ? (uint64_t)light * p.mirror_light_stride + midx : kNoMirror;
        bool cast = false;
        float so[3] = {0.0f, 0.0f, 0.0f}, sd[3] = {0.0f, 0.0f, 1.0f}, sclip = 0.0f;
        if (inside) {
            const float4 ph = __ldcg(reinterpret_cast<const float4 *>(p.hits) + idx); // written in this launch: L2, not L1
            if (__float_as_uint(ph.y) != DODRT_MISS) {
                const float light3[3] = {p.lights[light][0], p.lights[light][1], p.lights[light][2]};
                shadow_ray(o, d, ph.x, light3, so, sd, sclip);
                cast = true;
            }
        }
        const bool blocked = query<VARIANT>(p.scene, p.classes, cast, so, sd, true, sclip, h, &p, kFinishVisible, out, &donated, mirror);
        if ((inside || f.compact) && !donated) { // padded slots read as "not visible"
            const uint8_t v = (cast && !blocked) ? 1 : 0;
            p.visible[out] = v;
            if (mirror != kNoMirror) {
                p.mirror_visible[mirror] = v;
            }
        }
    }
    return true;
}

// BLOCK-FUSED work item (the default frame kernel): one 8x4 pixel block end to end -- its primary batch, then, from the
// hit records still in registers, its shadow batch for every light (each a separate, coherent any-hit query over the
// same 32 pixels: main.cpp:182-219).  No warp ever waits for another one: there is no readiness bookkeeping, no second
// claim and no re-read of the hit records; while one warp is in a long primary batch the others are in shadow batches,
// so the tails of the two kinds of work overlap like with the tile queues above.  hitPoint = o + d*t is computed from
// the same float the record holds, so the shadow ray has the same bits.
template <int VARIANT> __device__ __forceinline__ bool frame_block(const TraceParams &p)
{
    const uint32_t lane = threadIdx.x & 31u;
    const dodrt_frame &f = p.frame;
    unsigned long long base = 0;
    if (lane == 0) {
        base = atomicAdd(p.counter, 32ull);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= p.count) {
        return false;
    }
    uint32_t col = 0, row = 0;
    uint64_t slot = 0;
    const bool inside = slot_to_pixel(f, p.tiles_x, p.tile_order, base + lane, col, row, slot);
    const uint64_t pixel = (uint64_t)row * f.width + col;
    const uint64_t idx = f.compact ? slot : pixel;
    const uint64_t midx = p.mirror_by_pixel ? pixel : idx;
    // (a mirror with the local layout also receives the padded slots of edge tiles, like the local buffer)
    const bool mirrored = inside || (f.compact && !p.mirror_by_pixel);
    const float o[3] = {f.origin[0], f.origin[1], f.origin[2]};
    float d[3] = {0.0f, 0.0f, 1.0f};
    if (inside) {
        primary_dir(__ldg(p.xs + col), __ldg(p.ys + row), d);
    }
    Hit h;
    bool donated = false;
    {
        const uint64_t mirror = (p.mirror_hits != nullptr && mirrored) ? midx : kNoMirror;
        // a primary ray can only be given away when nothing follows it (no lights)
        query<VARIANT>(p.scene, p.classes, inside, o, d, false, kInfinity, h, &p, kFinishRecord, idx, &donated, mirror, p.num_lights == 0);
        if ((inside || f.compact) && !donated) { // padded slots of edge tiles read as "miss"
            const float4 r = make_float4(h.t, __uint_as_float(h.prim), h.u, h.v);
            reinterpret_cast<float4 *>(p.hits)[idx] = r;
            if (mirror != kNoMirror) {
                reinterpret_cast<float4 *>(p.mirror_hits)[mirror] = r;
            }
        }
    }
    const bool cast = inside && h.prim != DODRT_MISS;
    const float tHit = h.t;
    for (uint32_t light = 0; light < p.num_lights; light++) {
        const uint64_t out = (uint64_t)light * p.visible_light_stride + idx;
        const uint64_t mirror = (p.mirror_visible != nullptr && mirrored) ? (uint64_t)light * p.mirror_light_stride + midx : kNoMirror;
        float so[3] = {0.0f, 0.0f, 0.0f}, sd[3] = {0.0f, 0.0f, 1.0f}, sclip = 0.0f;
        if (cast) {
            const float light3[3] = {p.lights[light][0], p.lights[light][1], p.lights[light][2]};
            shadow_ray(o, d, tHit, light3, so, sd, sclip);
        }
        Hit hs;
        bool given = false;
        const bool blocked = query<VARIANT>(p.scene, p.classes, cast, so, sd, true, sclip, hs, &p, kFinishVisible, out, &given, mirror);
        if ((inside || f.compact) && !given) { // padded slots read as "not visible"
            const uint8_t v = (cast && !blocked) ? 1 : 0;
            p.visible[out] = v;
            if (mirror != kNoMirror) {
                p.mirror_visible[mirror] = v;
            }
        }
    }
    return true;
}

template <int VARIANT, bool TILE_QUEUES>
__global__ void __launch_bounds__(128, VARIANT >= kDonateVariant ? 5 : DODRT_MINBLOCKS) trace_frame_kernel(const TraceParams p)
{
    const uint32_t lane = threadIdx.x & 31u;
    if (VARIANT == kDonateVariant && p.donate_slots != nullptr && lane == 0) {
        atomicAdd(p.counter + kDonateStarted, 1ull); // see donate_helper_loop
    }
    // Two loops, each with its own SPECIALISED copy of the query (closest-hit with the full hit record / any-hit with
    // nothing but the boolean): one call site with a run-time any-hit flag cost the shadow queue 40 % (the any-hit
    // kernel drops all (t, u, v, id) bookkeeping and fits its registers; measured 5.24 vs 3.82 ms per dragon4k frame).
    if (TILE_QUEUES) {
        while (frame_batch<VARIANT, false>(p)) {
        }
        while (frame_batch<VARIANT, true>(p)) {
        }
    } else {
        while (frame_block<VARIANT>(p)) {
        }
    }
    if (VARIANT == kDonateVariant && p.donate_slots != nullptr) {
        donate_helper_loop(p);
    }
}
#endif // DODRT_EXPERIMENTS (frame kernels)

#ifdef DODRT_EXPERIMENTS
#include "dodrt_pool_kernel.inl"
#endif

// One thread per triangle slot: gather the 9 SoA floats of slot j of lane i and emit A, AB, AC.
// `primNums` (optional) = KDTree::m_primNums: output lane i is source lane primNums[i] -- the reference's
// Triangle::reorderLanesByIndices (triangle.cpp:349-367) fused into the upload (dodrt_scene_set_kdtree_indexed).
__global__ void repack_triangles_kernel(const float *__restrict__ lanes, const uint32_t *__restrict__ primNums,
                                        uint32_t numLanes, float4 *__restrict__ tris, float *__restrict__ lanes4)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)numLanes * kLane) {
        return;
    }
    const uint64_t srcLane = primNums ? (uint64_t)__ldg(primNums + idx / kLane) : idx / kLane;
    const float *lane = lanes + srcLane * 72;
    const uint32_t j = (uint32_t)(idx % kLane);
    const float Ax = lane[0 * kLane + j], Ay = lane[1 * kLane + j], Az = lane[2 * kLane + j];
    const float Bx = lane[3 * kLane + j], By = lane[4 * kLane + j], Bz = lane[5 * kLane + j];
    const float Cx = lane[6 * kLane + j], Cy = lane[7 * kLane + j], Cz = lane[8 * kLane + j];
    // avxVec3Sub(B, A), avxVec3Sub(C, A): triangle.cpp:66-67
    if (tris) { // only the A/B variants 0-2 read the per-triangle records
        tris[idx * 3 + 0] = make_float4(Ax, Ay, Az, Bx - Ax);
        tris[idx * 3 + 1] = make_float4(By - Ay, Bz - Az, Cx - Ax, Cy - Ay);
        tris[idx * 3 + 2] = make_float4(Cz - Az, 0.0f, 0.0f, 0.0f);
    }
    float *out = lanes4 + (idx / kLane) * 72 + j; // SoA lane: A, AB, AC components, 8 slots each
    out[0 * kLane] = Ax;
    out[1 * kLane] = Ay;
    out[2 * kLane] = Az;
    out[3 * kLane] = Bx - Ax;
    out[4 * kLane] = By - Ay;
    out[5 * kLane] = Bz - Az;
    out[6 * kLane] = Cx - Ax;
    out[7 * kLane] = Cy - Ay;
    out[8 * kLane] = Cz - Az;
}

// The per-triangle records of variants 0-2 from the SoA lanes (A, AB, AC are already there): built on demand only.
__global__ void tris_from_lanes4_kernel(const float *__restrict__ lanes4, uint32_t numLanes, float4 *__restrict__ tris)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)numLanes * kLane) {
        return;
    }
    const float *l = lanes4 + (idx / kLane) * 72 + (idx % kLane);
    tris[idx * 3 + 0] = make_float4(l[0 * kLane], l[1 * kLane], l[2 * kLane], l[3 * kLane]);
    tris[idx * 3 + 1] = make_float4(l[4 * kLane], l[5 * kLane], l[6 * kLane], l[7 * kLane]);
    tris[idx * 3 + 2] = make_float4(l[8 * kLane], 0.0f, 0.0f, 0.0f);
}

// out[lane][w] = in[primNums[lane]][w]: the lane re-order for any per-lane record of `wordsPerLane` 32-bit words
__global__ void gather_lanes_kernel(const uint32_t *__restrict__ in, const uint32_t *__restrict__ primNums, uint32_t numLanes,
                                    uint32_t wordsPerLane, uint32_t *__restrict__ out)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)numLanes * wordsPerLane) {
        return;
    }
    const uint64_t lane = idx / wordsPerLane, w = idx % wordsPerLane;
    out[idx] = in[(uint64_t)__ldg(primNums + lane) * wordsPerLane + w];
}

// Inverse of slot_to_pixel over all ranks: pixel -> (rank, slot).  Pure data movement (16+1 B per pixel).
__global__ void assemble_kernel(const dodrt_frame f, uint32_t tilesX, const float4 *__restrict__ compactHits,
                                const uint8_t *__restrict__ compactVis, uint64_t slotsPerRank,
                                float4 *__restrict__ hitsOut, uint8_t *__restrict__ visOut)
{
    const uint64_t pixel = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= (uint64_t)f.width * f.height) {
        return;
    }
    const uint32_t col = (uint32_t)(pixel % f.width), row = (uint32_t)(pixel / f.width);
    const uint32_t tx = col / f.tile_w, ty = row / f.tile_h;
    const uint32_t tile = ty * tilesX + tx;
    const uint32_t rank = tile % f.tile_stride, localTile = tile / f.tile_stride;
    const uint32_t ix = col - tx * f.tile_w, iy = row - ty * f.tile_h;
    const uint32_t block = (iy >> 2) * (f.tile_w >> 3) + (ix >> 3);
    const uint32_t in = block * 32u + (iy & 3u) * 8u + (ix & 7u);
    const uint64_t src = (uint64_t)rank * slotsPerRank + (uint64_t)localTile * f.tile_w * f.tile_h + in;
    hitsOut[pixel] = compactHits[src];
    if (visOut) {
        visOut[pixel] = compactVis[src];
    }
}

template <int MODE, int VARIANT> cudaError_t config_for(int device, LaunchConfig *cfg)
{
    int sms = 0, perSm = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
#ifdef DODRT_EXPERIMENTS
    if constexpr (MODE == kModeFrame) { // the fused kernel exists as the plain voted kernel and as the donating one
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, trace_frame_kernel<(VARIANT == kDonateVariant ? kDonateVariant : 3), false>, 128, 0);
        int perSmQ = 0;
        if (e == cudaSuccess) {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmQ, trace_frame_kernel<(VARIANT == kDonateVariant ? kDonateVariant : 3), true>, 128, 0);
        }
        if (perSmQ < perSm) perSm = perSmQ; // one launch shape (and one donation queue size) for both forms
    } else if constexpr (VARIANT == 4 && MODE != kModeShadowRays) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, trace_kernel_pool<MODE>, 128, 0);
    } else if constexpr (VARIANT == 4) { // the pool kernel has no explicit-ray shadow mode: variant 3 serves it
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, trace_kernel<MODE, 3>, 128, 0);
    } else
#endif
    {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, trace_kernel<MODE, VARIANT>, 128, 0);
    }
    if (e != cudaSuccess) return e;
    if (perSm < 1) perSm = 1;
    cfg->grid = sms * perSm; // persistent: exactly one resident wave
    cfg->block = 128;
    return cudaSuccess;
}

template <int MODE, int VARIANT> void launch_one(const TraceParams &p, const LaunchConfig &cfg, cudaStream_t stream)
{
#ifdef DODRT_EXPERIMENTS
    if constexpr (MODE == kModeFrame) {
        if (p.tile_done != nullptr) { // tile queues (A/B form, DODRT_FRAME_QUEUES=1)
            trace_frame_kernel<(VARIANT == kDonateVariant ? kDonateVariant : 3), true><<<cfg.grid, cfg.block, 0, stream>>>(p);
        } else {
            trace_frame_kernel<(VARIANT == kDonateVariant ? kDonateVariant : 3), false><<<cfg.grid, cfg.block, 0, stream>>>(p);
        }
    } else if constexpr (VARIANT == 4 && MODE != kModeShadowRays) {
        trace_kernel_pool<MODE><<<cfg.grid, cfg.block, 0, stream>>>(p);
    } else if constexpr (VARIANT == 4) {
        trace_kernel<MODE, 3><<<cfg.grid, cfg.block, 0, stream>>>(p);
    } else
#endif
    {
        trace_kernel<MODE, VARIANT><<<cfg.grid, cfg.block, 0, stream>>>(p);
    }
}

template <int VARIANT> cudaError_t config_mode(int device, TraceMode mode, LaunchConfig *cfg)
{
    switch (mode) {
    case kModeRays: return config_for<kModeRays, VARIANT>(device, cfg);
    case kModePrimary: return config_for<kModePrimary, VARIANT>(device, cfg);
    case kModeShadow: return config_for<kModeShadow, VARIANT>(device, cfg);
#ifdef DODRT_EXPERIMENTS
    case kModeFrame: return config_for<kModeFrame, VARIANT>(device, cfg);
#else
    case kModeFrame: return cudaErrorInvalidValue;
#endif
    default: return config_for<kModeShadowRays, VARIANT>(device, cfg);
    }
}

template <int VARIANT> void launch_mode(TraceMode mode, const TraceParams &p, const LaunchConfig &cfg, cudaStream_t stream)
{
    switch (mode) {
    case kModeRays: launch_one<kModeRays, VARIANT>(p, cfg, stream); break;
    case kModePrimary: launch_one<kModePrimary, VARIANT>(p, cfg, stream); break;
    case kModeShadow: launch_one<kModeShadow, VARIANT>(p, cfg, stream); break;
#ifdef DODRT_EXPERIMENTS
    case kModeFrame: launch_one<kModeFrame, VARIANT>(p, cfg, stream); break;
#else
    case kModeFrame: break; // not in this build (variant_compiled / mode_compiled are checked by the caller)
#endif
    default: launch_one<kModeShadowRays, VARIANT>(p, cfg, stream); break;
    }
}

} // namespace

int default_variant()
{
    static const int v = [] {
        const char *e = std::getenv("DODRT_VARIANT");
        if (!e) return kVariantAuto;
        const int x = std::atoi(e);
        return variant_compiled(x) ? x : kVariantAuto;
    }();
    return v;
}

// kVariantAuto: the plain voted kernel (variant 3) unless the caller says the pass is a SHARE of a larger job
// (`split`: a frame whose tiles are dealt over several GPUs, or a bounce pass of dodrt_render) and is only a few batches
// per warp -- there the tail of the pass (its slowest warp) is a large part of it and the donating kernel (variant 7)
// wins.  Measured on dragon4k (profiles/r01_donation.txt), max over ranks of primary + shadow kernel time, variant 3 vs
// 7: whole frame (87 batches per warp) 3.79 vs 3.89 ms; half (43) 2.31 vs 2.11; a quarter (22) 1.48 vs 1.18; an eighth
// (11) 1.02 vs 0.74 ms.  Whole small frames do not qualify: teapot1080 (22 batches per warp, short rays) 0.94 ms plain
// vs 0.99 ms donating -- but a small frame over a BIG tree does (`bigTree`, >= 8192 kd nodes: rays are long, the pass
// is mostly tail): dragon1080_primary 0.51 ms plain vs 0.34 ms donating.
int resolve_variant(int variant, const LaunchConfig &donateCfg, uint64_t count, bool split, bool bigTree)
{
    if (variant != kVariantAuto) {
        return variant;
    }
    const uint64_t threads = (uint64_t)donateCfg.grid * donateCfg.block;
    return ((split || bigTree) && count < threads * kDonateBelowBatches) ? kDonateVariant : kDefaultVariant;
}

cudaError_t trace_launch_config(int device, TraceMode mode, int variant, LaunchConfig *cfg)
{
    switch (variant) {
    case 0: return config_mode<0>(device, mode, cfg);
    case 3: return config_mode<3>(device, mode, cfg);
    case 7: return config_mode<7>(device, mode, cfg);
#ifdef DODRT_EXPERIMENTS
    case 1: return config_mode<1>(device, mode, cfg);
    case 2: return config_mode<2>(device, mode, cfg);
    case 4: return config_mode<4>(device, mode, cfg);
    case 5: return config_mode<5>(device, mode, cfg);
    case 6: return config_mode<6>(device, mode, cfg);
    case 8: return config_mode<8>(device, mode, cfg);
#endif
    default: return cudaErrorInvalidValue;
    }
}

size_t donation_queue_bytes(const LaunchConfig &cfg)
{
    // donors check `tail <= capacity / 2` before reserving; every warp of the grid can pass that check at the same
    // time and reserve up to 32 slots, hence twice the grid's threads
    const size_t cap = (size_t)cfg.grid * cfg.block * 2;
    const size_t readyBytes = (cap * sizeof(uint32_t) + 255) & ~(size_t)255;
    return readyBytes + cap * kDonateSlotWords * sizeof(uint32_t);
}

#ifdef DODRT_TIMELINE
void timeline_fetch(unsigned long long *exits, unsigned long long *donations, unsigned int *numDonations, unsigned int *pollsEmpty)
{
    cudaMemcpyFromSymbol(exits, g_tlExit, sizeof(unsigned long long) * 8192);
    cudaMemcpyFromSymbol(donations, g_tlDonation, sizeof(unsigned long long) * (1 << 16));
    cudaMemcpyFromSymbol(numDonations, g_tlDonations, sizeof(unsigned int));
    cudaMemcpyFromSymbol(pollsEmpty, g_tlPollsEmpty, sizeof(unsigned int));
    const unsigned int zero = 0;
    cudaMemcpyToSymbol(g_tlDonations, &zero, sizeof(zero));
    cudaMemcpyToSymbol(g_tlPollsEmpty, &zero, sizeof(zero));
    static unsigned long long zeros[8192];
    cudaMemcpyToSymbol(g_tlExit, zeros, sizeof(zeros));
}
#endif

cudaError_t launch_trace(TraceMode mode, const TraceParams &params, const LaunchConfig &cfg, cudaStream_t stream,
                         cudaMemPool_t pool, void *persistentQueue, uint32_t epoch)
{
    TraceParams p = params;
    cudaError_t e = cudaSuccess;
    if (mode != kModeFrame) { // kModeFrame: the caller zeroes counters + tile bookkeeping with one memset
        e = cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long) * kCounterWords, stream);
    }
    if (e != cudaSuccess) return e;
#ifdef DODRT_TIMELINE
    cudaMemsetAsync(p.counter + 24, 0xFF, 16, stream); // the two minima
#endif
    // variant 7: the donation queue -- one slot per thread of the grid (x2) bounds the number of rays that can ever be
    // suspended at once
    p.donate_slots = p.donate_ready = nullptr;
    p.donate_capacity = 0;
    p.donate_epoch = 1;
    void *queue = nullptr;
    static const bool donateOn = [] { const char *e = std::getenv("DODRT_DONATE"); return !e || std::atoi(e) != 0; }();
    if (donateOn && p.variant == kDonateVariant && p.scene.num_nodes != 0 && (p.classes & DODRT_CLS_TREE) &&
        (persistentQueue != nullptr || pool != nullptr)) {
        const size_t cap = (size_t)cfg.grid * cfg.block * 2;
        const size_t readyBytes = (cap * sizeof(uint32_t) + 255) & ~(size_t)255;
        void *mem = persistentQueue;
        if (mem != nullptr && epoch != 0u) {
            p.donate_epoch = epoch; // nothing to clear: words written by earlier launches carry other epochs
        } else {
            e = cudaMallocFromPoolAsync(&queue, readyBytes + cap * kDonateSlotWords * sizeof(uint32_t), pool, stream);
            if (e != cudaSuccess) return e;
            e = cudaMemsetAsync(queue, 0, readyBytes, stream);
            if (e != cudaSuccess) return e;
            mem = queue;
        }
        p.donate_ready = static_cast<uint32_t *>(mem);
        p.donate_slots = reinterpret_cast<uint32_t *>(static_cast<char *>(mem) + readyBytes);
        p.donate_capacity = (uint32_t)cap;
    }
    if (p.tile_order && (mode == kModePrimary || mode == kModeShadow || mode == kModeFrame)) {
        const unsigned blocks = (p.num_local_tiles + 127u) / 128u;
        if (mode != kModeShadow) {
            order_tiles_kernel<kModePrimary><<<blocks, 128, 0, stream>>>(p);
        } else {
            order_tiles_kernel<kModeShadow><<<blocks, 128, 0, stream>>>(p);
        }
    }
    switch (p.variant) {
    case 0: launch_mode<0>(mode, p, cfg, stream); break;
    case 3: launch_mode<3>(mode, p, cfg, stream); break;
    case 7: launch_mode<7>(mode, p, cfg, stream); break;
#ifdef DODRT_EXPERIMENTS
    case 1: launch_mode<1>(mode, p, cfg, stream); break;
    case 2: launch_mode<2>(mode, p, cfg, stream); break;
    case 4: launch_mode<4>(mode, p, cfg, stream); break;
    case 5: launch_mode<5>(mode, p, cfg, stream); break;
    case 6: launch_mode<6>(mode, p, cfg, stream); break;
    case 8: launch_mode<8>(mode, p, cfg, stream); break;
#endif
    default: return cudaErrorInvalidValue;
    }
    e = cudaGetLastError();
    if (queue) {
        cudaFreeAsync(queue, stream);
    }
    return e;
}

cudaError_t launch_assemble(const dodrt_frame &frame, uint32_t tiles_x, const dodrt_hit *compactHits,
                            const uint8_t *compactVis, uint64_t slotsPerRank, dodrt_hit *hitsOut, uint8_t *visOut,
                            cudaStream_t stream)
{
    const uint64_t n = (uint64_t)frame.width * frame.height;
    const int block = 256;
    assemble_kernel<<<(unsigned)((n + block - 1) / block), block, 0, stream>>>(
        frame, tiles_x, reinterpret_cast<const float4 *>(compactHits), compactVis, slotsPerRank,
        reinterpret_cast<float4 *>(hitsOut), visOut);
    return cudaGetLastError();
}

cudaError_t launch_gather_lanes(const uint32_t *d_in, const uint32_t *d_prim_nums, uint32_t num_lanes, uint32_t words_per_lane,
                                uint32_t *d_out, cudaStream_t stream)
{
    const uint64_t n = (uint64_t)num_lanes * words_per_lane;
    if (n == 0) return cudaSuccess;
    const int block = 256;
    gather_lanes_kernel<<<(unsigned)((n + block - 1) / block), block, 0, stream>>>(d_in, d_prim_nums, num_lanes, words_per_lane,
                                                                                    d_out);
    return cudaGetLastError();
}

cudaError_t launch_tris_from_lanes4(const float4 *d_lanes4, uint32_t num_lanes, float4 *d_tris, cudaStream_t stream)
{
    const uint64_t n = (uint64_t)num_lanes * kLane;
    if (n == 0) return cudaSuccess;
    const int block = 256;
    tris_from_lanes4_kernel<<<(unsigned)((n + block - 1) / block), block, 0, stream>>>(reinterpret_cast<const float *>(d_lanes4),
                                                                                        num_lanes, d_tris);
    return cudaGetLastError();
}

cudaError_t launch_repack_triangles(const float *d_lanes, const uint32_t *d_prim_nums, uint32_t num_lanes, float4 *d_tris,
                                    float4 *d_lanes4, cudaStream_t stream)
{
    const uint64_t n = (uint64_t)num_lanes * kLane;
    if (n == 0) return cudaSuccess;
    const int block = 256;
    const unsigned grid = (unsigned)((n + block - 1) / block);
    repack_triangles_kernel<<<grid, block, 0, stream>>>(d_lanes, d_prim_nums, num_lanes, d_tris,
                                                        reinterpret_cast<float *>(d_lanes4));
    return cudaGetLastError();
}

} // namespace dodrt

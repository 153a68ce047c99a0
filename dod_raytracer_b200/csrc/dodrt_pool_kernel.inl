// dodrt_pool_kernel.inl -- variant 4: voted traversal with warp-level regrouping of rays.
// Textually included by dodrt_kernels.cu inside namespace dodrt::{anonymous}.
//
// Why: profiles/r01_v2_* show the triangle stage -- 55 % of all issued instructions -- running with 16 of
// 32 lanes: rays that missed the kd-tree bounds, were decided by the analytic classes, or finished early
// (any-hit) leave their lanes idle until the slowest ray of the 32-ray batch is done.
//
// How: each persistent warp splits the query into two stages and keeps a small pool of rays between them
// in shared memory.
//   stage A (full warp, 32 fresh work items): build the ray, run the analytic classes, do the bounds slab
//           test; rays that do not need the kd-tree are finished here, the survivors are COMPACTED
//           (ballot + popc prefix) into the warp's pool;
//   stage B: idle lanes pull rays out of the pool and join the warp-voted traversal (node step vs triangle
//           lane, the phase rule of kdtree_query_voted); when >= DeviceScene::pool_refill lanes are idle again the warp
//           goes back to top up (default 32 = no refill in mid-flight, the best setting measured).
// Per-ray order of events is exactly the reference's (kdtree.cpp:263-361); only which lane hosts a ray, and
// when, changes -- results are bit-identical to every other variant.

constexpr int kPoolCap = 96;     // entries per warp
constexpr int kPoolWords = 13;   // odd stride: conflict-free when consecutive lanes touch consecutive entries
constexpr int kWarpsPerBlock = 4;

struct PoolRay {
    float o[3], d[3];
    float clip, tmin, tmax;
    float hitT;
    uint32_t hitPrim;
    uint32_t out;   // result slot (frame modes: < 2^32 checked by the API; ray batches: split below)
    uint32_t flags; // bit0 = any-hit, bits 1.. = high part of the result slot
};

__device__ __forceinline__ void pool_store(uint32_t *e, const PoolRay &r)
{
    e[0] = __float_as_uint(r.o[0]); e[1] = __float_as_uint(r.o[1]); e[2] = __float_as_uint(r.o[2]);
    e[3] = __float_as_uint(r.d[0]); e[4] = __float_as_uint(r.d[1]); e[5] = __float_as_uint(r.d[2]);
    e[6] = __float_as_uint(r.clip); e[7] = __float_as_uint(r.tmin); e[8] = __float_as_uint(r.tmax);
    e[9] = __float_as_uint(r.hitT); e[10] = r.hitPrim; e[11] = r.out; e[12] = r.flags;
}

__device__ __forceinline__ void pool_load(const uint32_t *e, PoolRay &r)
{
    r.o[0] = __uint_as_float(e[0]); r.o[1] = __uint_as_float(e[1]); r.o[2] = __uint_as_float(e[2]);
    r.d[0] = __uint_as_float(e[3]); r.d[1] = __uint_as_float(e[4]); r.d[2] = __uint_as_float(e[5]);
    r.clip = __uint_as_float(e[6]); r.tmin = __uint_as_float(e[7]); r.tmax = __uint_as_float(e[8]);
    r.hitT = __uint_as_float(e[9]); r.hitPrim = e[10]; r.out = e[11]; r.flags = e[12];
}

template <int MODE>
__device__ __forceinline__ void write_result(const TraceParams &p, uint64_t out, bool any, float clip0, const Hit &hit,
                                             bool blocked)
{
    if (MODE == kModeShadow) {
        p.visible[out] = blocked ? 0 : 1;
    } else if (MODE == kModeRays && any) { // any-hit defines only hit/miss
        reinterpret_cast<float4 *>(p.hits)[out] = make_float4(clip0, __uint_as_float(blocked ? 0u : DODRT_MISS), 0.0f, 0.0f);
    } else {
        reinterpret_cast<float4 *>(p.hits)[out] = make_float4(hit.t, __uint_as_float(hit.prim), hit.u, hit.v);
    }
}

// Stage A for one work item: build the ray, run the analytic classes (Sphere -> [Box] -> Plane -> Cylinder), test
// the kd-tree bounds.  A ray that does not need the tree is FINISHED here (its result is written); returns true,
// with `r` filled, when the ray has to traverse the tree.
template <int MODE>
__device__ __forceinline__ bool stage_a(const TraceParams &p, bool useTree, uint64_t item, PoolRay &r)
{
    const DeviceScene &s = p.scene;
    const bool inRange = item < p.count;
    r.o[0] = r.o[1] = r.o[2] = 0.0f;
    r.d[0] = r.d[1] = 0.0f, r.d[2] = 1.0f;
    r.clip = 0.0f, r.flags = 0u;
    r.tmin = r.tmax = 0.0f;
    bool valid = false;  // there is a ray to trace
    bool writes = false; // this lane owns a result slot
    uint64_t slot = item;
    if (MODE == kModeRays) {
        if (inRange) {
            const float4 *src = reinterpret_cast<const float4 *>(p.rays + item);
            const float4 a = __ldg(src), b = __ldg(src + 1);
            r.o[0] = a.x, r.o[1] = a.y, r.o[2] = a.z;
            r.d[0] = a.w, r.d[1] = b.x, r.d[2] = b.y;
            r.clip = b.z;
            r.flags = __float_as_uint(b.w) & DODRT_RAY_ANY;
            writes = true;
            valid = (__float_as_uint(b.w) & DODRT_RAY_SKIP) == 0u;
        }
    } else {
        uint32_t col = 0, row = 0;
        uint64_t compactSlot = 0;
        const bool inside = inRange && slot_to_pixel(p.frame, p.tiles_x, p.tile_order, item, col, row, compactSlot);
        slot = p.frame.compact ? compactSlot : (uint64_t)row * p.frame.width + col;
        writes = inside || (inRange && p.frame.compact);
        r.o[0] = p.frame.origin[0], r.o[1] = p.frame.origin[1], r.o[2] = p.frame.origin[2];
        if (inside) {
            primary_dir(__ldg(p.xs + col), __ldg(p.ys + row), r.d);
        }
        if (MODE == kModePrimary) {
            valid = inside;
            r.clip = kInfinity;
        } else {
            r.flags = DODRT_RAY_ANY;
            if (inside) {
                const float4 ph = reinterpret_cast<const float4 *>(p.hits)[slot];
                if (__float_as_uint(ph.y) != DODRT_MISS) {
                    float so[3], sd[3];
                    shadow_ray(r.o, r.d, ph.x, p.light, so, sd, r.clip);
                    r.o[0] = so[0], r.o[1] = so[1], r.o[2] = so[2];
                    r.d[0] = sd[0], r.d[1] = sd[1], r.d[2] = sd[2];
                    valid = true;
                }
            }
        }
    }
    const bool rayAny = (r.flags & DODRT_RAY_ANY) != 0u;
    const float rayClip0 = r.clip;
    Hit h;
    h.t = r.clip, h.prim = DODRT_MISS, h.u = h.v = 0.0f;
    bool aFound = false, decided = !valid;
    if (valid) {
        decided = analytic_chain(s, p.classes, r.o, r.d, rayAny, r.clip, h, aFound);
    }
    bool survive = false;
    if (!decided && useTree) {
        const float inv[3] = {1.0f / r.d[0], 1.0f / r.d[1], 1.0f / r.d[2]};
        survive = slab(s.bmin, s.bmax, r.o, inv, r.clip, r.tmin, r.tmax) && !(r.tmin > r.clip);
    }
    if (writes && !survive) { // finished without the kd-tree
        const bool blocked = (MODE == kModeShadow) ? (!valid || aFound) : aFound;
        write_result<MODE>(p, slot, rayAny, rayClip0, h, blocked);
    }
    r.hitT = h.t;
    r.hitPrim = h.prim;
    r.out = (uint32_t)slot;
    r.flags |= (uint32_t)(slot >> 32) << 1;
    return survive;
}

template <int MODE> __global__ void __launch_bounds__(32 * kWarpsPerBlock) trace_kernel_pool(const TraceParams p)
{
    __shared__ uint32_t poolMem[kWarpsPerBlock][kPoolCap * kPoolWords];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t ltMask = (1u << lane) - 1u;
    uint32_t *pool = poolMem[threadIdx.x >> 5];
    const DeviceScene &s = p.scene;
    const bool useTree = (p.classes & DODRT_CLS_TREE) && s.num_nodes != 0;

    // the ray this lane is hosting in stage B
    float o[3] = {0.0f, 0.0f, 0.0f}, d[3] = {0.0f, 0.0f, 1.0f};
    float clip = 0.0f, clip0 = 0.0f;
    Hit hit;
    hit.t = 0.0f, hit.prim = DODRT_MISS, hit.u = hit.v = 0.0f;
    bool found = false, any = false, hosting = false;
    uint64_t out = 0;
    TreeState st;
    tree_enter(s, st, false, o, d, 0.0f);
    uint32_t stackNode[kMaxStack];
    float stackTmin[kMaxStack];
    float stackTmax[kMaxStack];

    int poolCount = 0; // warp-uniform
    bool exhausted = false;
    for (;;) {
        // ---- retire finished rays ----------------------------------------------------------------------
        if (hosting && !st.live) {
            write_result<MODE>(p, out, any, clip0, hit, found);
            hosting = false;
        }
        const int nFree = 32 - __popc(__ballot_sync(0xffffffffu, st.live));

        // ---- stage A: top up the pool ------------------------------------------------------------------------
        while (!exhausted && poolCount < nFree && poolCount <= kPoolCap - 32) {
            unsigned long long base = 0;
            if (lane == 0) {
                base = atomicAdd(p.counter, 32ull);
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= p.count) {
                exhausted = true;
                break;
            }
            PoolRay r;
            const bool survive = stage_a<MODE>(p, useTree, base + lane, r);
            const unsigned surviveMask = __ballot_sync(0xffffffffu, survive);
            if (survive) {
                pool_store(pool + (poolCount + __popc(surviveMask & ltMask)) * kPoolWords, r);
            }
            poolCount += __popc(surviveMask);
        }
        __syncwarp();

        // ---- hand pooled rays to idle lanes --------------------------------------------------------------------
        {
            const unsigned freeMask = ~__ballot_sync(0xffffffffu, st.live);
            const int myRank = __popc(freeMask & ltMask);
            if (!st.live && myRank < poolCount) {
                PoolRay r;
                pool_load(pool + (poolCount - 1 - myRank) * kPoolWords, r);
                o[0] = r.o[0], o[1] = r.o[1], o[2] = r.o[2];
                d[0] = r.d[0], d[1] = r.d[1], d[2] = r.d[2];
                clip = r.clip;
                any = (r.flags & DODRT_RAY_ANY) != 0u;
                clip0 = r.clip; // only reported for any-hit rays, whose clip never changes before the hit
                hit.t = r.hitT, hit.prim = r.hitPrim, hit.u = hit.v = 0.0f;
                out = (uint64_t)r.out | ((uint64_t)(r.flags >> 1) << 32);
                found = false;
                hosting = true;
                st.inv[0] = 1.0f / d[0]; // kdtree.cpp:271 (same IEEE divisions as in stage A)
                st.inv[1] = 1.0f / d[1];
                st.inv[2] = 1.0f / d[2];
                st.tmin = r.tmin, st.tmax = r.tmax;
                st.node = 0, st.sp = 0, st.triCur = st.triEnd = 0;
                st.live = true;
            }
            const int taken = min(poolCount, __popc(freeMask));
            poolCount -= taken;
        }
        __syncwarp();

        // ---- stage B: voted traversal until enough lanes are idle to make a top-up worthwhile -----------------------
        bool anyLive = false;
        uint32_t nodeRun = 0;
        for (;;) {
            const bool wantLeaf = st.live && st.triCur < st.triEnd;
            const bool wantNode = st.live && !wantLeaf;
            const unsigned leafMask = __ballot_sync(0xffffffffu, wantLeaf);
            const unsigned nodeMask = __ballot_sync(0xffffffffu, wantNode);
            const int nLive = __popc(leafMask | nodeMask);
            anyLive = nLive != 0;
            if (!anyLive) {
                break;
            }
            if (32 - nLive >= (int)s.pool_refill && (poolCount > 0 || !exhausted)) {
                break;
            }
            // same phase rule and node bursts as kdtree_query_voted
            const uint32_t nLeaf = __popc(leafMask), nNode = __popc(nodeMask);
            if (nNode == 0u || (nLeaf != 0u && (nLeaf * s.tune[1] >= nNode * s.tune[0] || nodeRun >= s.tune[2]))) {
                nodeRun = 0;
                if (wantLeaf) {
                    leaf_step<true, false>(s, st, o, d, any, clip, hit, found, stackNode, stackTmin, stackTmax);
                }
            } else {
                ++nodeRun;
                if (wantNode) {
                    node_step(s, st, o, d, clip, stackNode, stackTmin, stackTmax);
                    for (uint32_t k = 1; k < s.node_burst && st.live && !(st.triCur < st.triEnd); k++) {
                        node_step(s, st, o, d, clip, stackNode, stackTmin, stackTmax);
                    }
                }
            }
        }
        if (!anyLive && exhausted && poolCount == 0) {
            if (hosting) { // rays that finished in the last round
                write_result<MODE>(p, out, any, clip0, hit, found);
            }
            break;
        }
    }
}

// dodrt_kernels.cu -- hand-written sm_100a traversal / intersection kernels.
//
// Compile with: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
// (-fmad=false is part of the numerical contract, see dodrt_device.cuh).
//
// Kernel shape (v1): persistent warps.  Each warp claims 32 consecutive work items at a time from a
// global counter (one atomic per warp per claim), every thread runs ONE ray through the reference's
// query chain Sphere -> [Box] -> Plane -> Cylinder -> KDTree with a per-thread short stack, and the
// warp claims again when all of its rays are done.  In the frame modes 32 consecutive work items
// are an 8x4 pixel block, so a warp's rays are coherent and its node / triangle fetches coalesce
// into broadcasts.  The tensor cores are idle by design: the path has no dense contraction.
#include "dodrt_kernels.cuh"

namespace dodrt {

namespace {

__device__ __forceinline__ float pick(const float v[3], uint32_t axis)
{
    return axis == 0 ? v[0] : (axis == 1 ? v[1] : v[2]);
}

// KDTree::intersect, kdtree.cpp:263-361, with Triangle::intersectInRange (triangle.cpp:22-177)
// inlined for the leaves.  `clip` is _Intersect::clippingDistance (in/out, kdtree.cpp:343).
template <bool ANY>
__device__ __forceinline__ bool kdtree_query(const DeviceScene &s, const float o[3], const float d[3], float &clip,
                                             Hit &hit)
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
                const float4 q0 = __ldg(tri), q1 = __ldg(tri + 1), q2 = __ldg(tri + 2);
                float t, u, v;
                if (triangle_test(q0, q1, q2, o, d, clip, t, u, v)) {
                    clip = t; // running maximumDistance, then kdtree.cpp:343
                    hit.t = t;
                    hit.prim = (DODRT_KIND_TRIANGLE << DODRT_KIND_SHIFT) | (firstTri + k);
                    hit.u = u;
                    hit.v = v;
                    found = true;
                    if (ANY) { // kdtree.cpp:338-341: only the boolean is defined for any-hit
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

// The query chain: closest hit main.cpp:312-321, any hit main.cpp:198-217.
template <bool ANY>
__device__ __forceinline__ bool query_chain(const DeviceScene &s, uint32_t classes, const float o[3], const float d[3],
                                            float clip, Hit &hit)
{
    bool found = false;
    Hit h;
    hit.t = clip;
    hit.prim = DODRT_MISS;
    hit.u = hit.v = 0.0f;
    if ((classes & DODRT_CLS_SPHERE) && s.num_spheres && sphere_query(s, o, d, ANY, clip, h)) {
        hit = h;
        found = true;
        if (ANY) return true;
        clip = h.t;
    }
    if ((classes & DODRT_CLS_BOX) && s.num_boxes && box_query(s, o, d, ANY, clip, h)) {
        hit = h;
        found = true;
        if (ANY) return true;
        clip = h.t;
    }
    if ((classes & DODRT_CLS_PLANE) && s.num_planes && plane_query(s, o, d, clip, h)) {
        hit = h;
        found = true;
        if (ANY) return true;
        clip = h.t;
    }
    if ((classes & DODRT_CLS_CYLINDER) && s.num_cylinders && cylinder_query(s, o, d, clip, h)) {
        hit = h;
        found = true;
        if (ANY) return true;
        clip = h.t;
    }
    if ((classes & DODRT_CLS_TREE) && kdtree_query<ANY>(s, o, d, clip, h)) {
        hit = h;
        found = true;
    }
    return found;
}

// slot -> pixel for the frame modes (see dodrt_frame in include/dodrt.h): tiles round-robin over
// ranks, 8x4 pixel blocks inside a tile so that one warp = one block.
__device__ __forceinline__ bool slot_to_pixel(const dodrt_frame &f, uint32_t tiles_x, uint64_t slot, uint32_t &col,
                                              uint32_t &row)
{
    const uint32_t tilePixels = f.tile_w * f.tile_h;
    const uint32_t localTile = (uint32_t)(slot / tilePixels);
    const uint32_t in = (uint32_t)(slot % tilePixels);
    const uint32_t tile = f.first_tile + localTile * f.tile_stride;
    const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
    const uint32_t block = in >> 5, lane = in & 31u;
    const uint32_t bpr = f.tile_w >> 3;
    const uint32_t bx = block % bpr, by = block / bpr;
    col = tx * f.tile_w + bx * 8 + (lane & 7u);
    row = ty * f.tile_h + by * 4 + (lane >> 3);
    return col < f.width && row < f.height;
}

template <int MODE> __global__ void __launch_bounds__(128) trace_kernel(const TraceParams p)
{
    const uint32_t lane = threadIdx.x & 31u;
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
        if (item >= p.count) {
            continue;
        }
        if (MODE == kModeRays) {
            const float4 *src = reinterpret_cast<const float4 *>(p.rays + item);
            const float4 a = __ldg(src), b = __ldg(src + 1);
            const float o[3] = {a.x, a.y, a.z};
            const float d[3] = {a.w, b.x, b.y};
            const float clip = b.z;
            const uint32_t flags = __float_as_uint(b.w);
            Hit h;
            bool found;
            if (flags & DODRT_RAY_ANY) {
                found = query_chain<true>(p.scene, p.classes, o, d, clip, h);
                h.t = clip; // any-hit defines only hit/miss
                h.prim = found ? 0u : DODRT_MISS;
                h.u = h.v = 0.0f;
            } else {
                found = query_chain<false>(p.scene, p.classes, o, d, clip, h);
            }
            reinterpret_cast<float4 *>(p.hits)[item] = make_float4(h.t, __uint_as_float(h.prim), h.u, h.v);
        } else {
            uint32_t col, row;
            const bool inside = slot_to_pixel(p.frame, p.tiles_x, item, col, row);
            const uint64_t out = p.frame.compact ? item : (uint64_t)row * p.frame.width + col;
            if (!inside) {
                if (p.frame.compact) {
                    if (MODE == kModePrimary) {
                        reinterpret_cast<float4 *>(p.hits)[out] =
                            make_float4(__int_as_float(0x7f800000), __uint_as_float(DODRT_MISS), 0.0f, 0.0f);
                    } else {
                        p.visible[out] = 0;
                    }
                }
                continue;
            }
            const float o[3] = {p.frame.origin[0], p.frame.origin[1], p.frame.origin[2]};
            float d[3];
            primary_dir(__ldg(p.xs + col), __ldg(p.ys + row), d);
            if (MODE == kModePrimary) {
                Hit h;
                query_chain<false>(p.scene, p.classes, o, d, __int_as_float(0x7f800000), h);
                reinterpret_cast<float4 *>(p.hits)[out] = make_float4(h.t, __uint_as_float(h.prim), h.u, h.v);
            } else {
                const float4 ph = reinterpret_cast<const float4 *>(p.hits)[out];
                uint8_t vis = 0;
                if (__float_as_uint(ph.y) != DODRT_MISS) {
                    float so[3], sd[3], sclip;
                    shadow_ray(o, d, ph.x, p.light, so, sd, sclip);
                    Hit h;
                    vis = query_chain<true>(p.scene, p.classes, so, sd, sclip, h) ? 0 : 1;
                }
                p.visible[out] = vis;
            }
        }
    }
}

// One thread per triangle slot: gather the 9 SoA floats of slot j of lane i and emit A, AB, AC.
__global__ void repack_triangles_kernel(const float *__restrict__ lanes, uint32_t numLanes, float4 *__restrict__ tris)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)numLanes * kLane) {
        return;
    }
    const float *lane = lanes + (idx / kLane) * 72;
    const uint32_t j = (uint32_t)(idx % kLane);
    const float Ax = lane[0 * kLane + j], Ay = lane[1 * kLane + j], Az = lane[2 * kLane + j];
    const float Bx = lane[3 * kLane + j], By = lane[4 * kLane + j], Bz = lane[5 * kLane + j];
    const float Cx = lane[6 * kLane + j], Cy = lane[7 * kLane + j], Cz = lane[8 * kLane + j];
    // avxVec3Sub(B, A), avxVec3Sub(C, A): triangle.cpp:66-67
    tris[idx * 3 + 0] = make_float4(Ax, Ay, Az, Bx - Ax);
    tris[idx * 3 + 1] = make_float4(By - Ay, Bz - Az, Cx - Ax, Cy - Ay);
    tris[idx * 3 + 2] = make_float4(Cz - Az, 0.0f, 0.0f, 0.0f);
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

template <int MODE> cudaError_t config_for(int device, LaunchConfig *cfg)
{
    int sms = 0, perSm = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, trace_kernel<MODE>, 128, 0);
    if (e != cudaSuccess) return e;
    if (perSm < 1) perSm = 1;
    cfg->grid = sms * perSm; // persistent: exactly one resident wave
    cfg->block = 128;
    return cudaSuccess;
}

} // namespace

cudaError_t trace_launch_config(int device, TraceMode mode, LaunchConfig *cfg)
{
    switch (mode) {
    case kModeRays: return config_for<kModeRays>(device, cfg);
    case kModePrimary: return config_for<kModePrimary>(device, cfg);
    default: return config_for<kModeShadow>(device, cfg);
    }
}

cudaError_t launch_trace(TraceMode mode, const TraceParams &p, const LaunchConfig &cfg, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    switch (mode) {
    case kModeRays: trace_kernel<kModeRays><<<cfg.grid, cfg.block, 0, stream>>>(p); break;
    case kModePrimary: trace_kernel<kModePrimary><<<cfg.grid, cfg.block, 0, stream>>>(p); break;
    default: trace_kernel<kModeShadow><<<cfg.grid, cfg.block, 0, stream>>>(p); break;
    }
    return cudaGetLastError();
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

cudaError_t launch_repack_triangles(const float *d_lanes, uint32_t num_lanes, float4 *d_tris, cudaStream_t stream)
{
    const uint64_t n = (uint64_t)num_lanes * kLane;
    if (n == 0) return cudaSuccess;
    const int block = 256;
    const unsigned grid = (unsigned)((n + block - 1) / block);
    repack_triangles_kernel<<<grid, block, 0, stream>>>(d_lanes, num_lanes, d_tris);
    return cudaGetLastError();
}

} // namespace dodrt

// dodrt_render_kernels.cu -- SURVEY.md 8(f) rows f-2 / f-3: the reference's shading and bounce loop
// (rayTrace, main.cpp:273-347; getLightingFactor / shadeDiffuse / shadeSpecular / toOutputChannelType,
// main.cpp:156-244) as a wavefront around the ray-query kernels:
//
// The image is walked in the frame's tile / 8x4-block order (slot_to_pixel), so 32 consecutive rays of the first bounce are
// one coherent pixel block, and a call renders only ITS tiles (first_tile / tile_stride): the reference's split of the
// frame over its threads (main.cpp:371-394) becomes a split over GPUs (dodrt_multi_render).
//
//   render_init      per pixel: primary ray (main.cpp:304-310), finalColor = 0, pixel alive
//   for k in 0..depth-1:
//       trace        closest-hit chain for every live pixel's current ray     (launch_trace kModeRays)
//       shadow x L   canSeeLight from the hit point to light l                (launch_trace kModeShadowRays)
//       render_shade rebuild the HitRecord from (prim,t,u,v), lighting, blend with weight 1/2^k, reflect
//   render_finish    clamp(finalColor*255) -> 8-bit RGB
//
// Everything that feeds the NEXT ray (hit point, normal, reflection, epsilon offset) is fp32 in the reference's
// operation order, un-fused (-fmad=false), so bounce rays are bit-identical; the specular term goes through a
// double pow like the reference's std::pow(float,int) and only influences the colour.
#include "dodrt_kernels.cuh"

namespace dodrt {

namespace {

constexpr float kInf = __builtin_huge_valf();

__device__ __forceinline__ void normalize3(const float v[3], float out[3]) // glm::normalize = v * (1/sqrt(dot(v,v)))
{
    const float inv = 1.0f / sqrtf(dot3(v[0], v[1], v[2], v[0], v[1], v[2]));
    out[0] = v[0] * inv;
    out[1] = v[1] * inv;
    out[2] = v[2] * inv;
}

// glm::reflect(I, N) = I - N * dot(N, I) * 2
__device__ __forceinline__ void reflect3(const float I[3], const float N[3], float out[3])
{
    const float dn = dot3(N[0], N[1], N[2], I[0], I[1], I[2]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        out[k] = I[k] - (N[k] * dn) * 2.0f;
    }
}

// Which part of a cylinder produced hit distance t, and its normal: Cylinder::intersect_non_vectorized
// (cylinder.cpp:155-210) tests body, base disc, top disc in that order with a strict `<`, so the winner is
// the first part that attains the smallest distance.
__device__ __forceinline__ void cylinder_normal(const dodrt_cylinder &c, float eps, const float o[3], const float d[3],
                                                const float P[3], float N[3])
{
    float tBody = kInf, tA = kInf, tB = kInf, t;
    if (cylinder_body(c, o, d, eps, t)) tBody = t;
    if (cylinder_disc(c, o, d, eps, 0.0f, kInf, t)) tA = t;
    if (cylinder_disc(c, o, d, eps, c.height, kInf, t)) tB = t;
    int part = 0;
    float best = tBody;
    if (tA < best) {
        best = tA;
        part = 1;
    }
    if (tB < best) {
        part = 2;
    }
    if (part == 0) { // cylinder.cpp:113-116
        const float w[3] = {P[0] - c.base[0], P[1] - c.base[1], P[2] - c.base[2]};
        const float minX = dot3(w[0], w[1], w[2], c.axis[0], c.axis[1], c.axis[2]);
        const float r[3] = {w[0] - c.axis[0] * minX, w[1] - c.axis[1] * minX, w[2] - c.axis[2] * minX};
        normalize3(r, N);
    } else { // cylinder.cpp:150
        const bool flip = dot3(d[0], d[1], d[2], c.axis[0], c.axis[1], c.axis[2]) > 0.0f;
        N[0] = flip ? -c.axis[0] : c.axis[0];
        N[1] = flip ? -c.axis[1] : c.axis[1];
        N[2] = flip ? -c.axis[2] : c.axis[2];
    }
}

// The reference's HitRecord (hitrecord.h:4-10) rebuilt from what the query kernels return.
__device__ __forceinline__ void rebuild_record(const DeviceScene &s, const float o[3], const float d[3], float t,
                                               uint32_t prim, float u, float v, float P[3], float N[3], float C[3])
{
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float m = d[k] * t; // hitPoint = rayOrigin + rayDir * t, triangle.cpp:170 / sphere.cpp:156 / plane.cpp:129
        P[k] = o[k] + m;
    }
    const uint32_t kind = prim >> DODRT_KIND_SHIFT, id = prim & ((1u << DODRT_KIND_SHIFT) - 1u);
    C[0] = C[1] = C[2] = 0.0f;
    N[0] = N[1] = N[2] = 0.0f;
    if (kind == DODRT_KIND_TRIANGLE) { // triangle.cpp:147-176
        const uint32_t *a = s.tri_attrs + (size_t)(id >> 3) * 80;
        const uint32_t slot = id & 7u;
        const float *AN = reinterpret_cast<const float *>(a + 8) + slot * 3;
        const float *BN = reinterpret_cast<const float *>(a + 32) + slot * 3;
        const float *CN = reinterpret_cast<const float *>(a + 56) + slot * 3;
        const float b0 = 1.0f - (u + v), b1 = u, b2 = v; // triangle.cpp:137
#pragma unroll
        for (int k = 0; k < 3; k++) {
            N[k] = (AN[k] * b0 + BN[k] * b1) + CN[k] * b2; // mat3(AN,BN,CN) * bary, triangle.cpp:171-174
        }
        const float *mc = s.mesh_colors + (size_t)a[slot] * 3;
        C[0] = mc[0], C[1] = mc[1], C[2] = mc[2];
    } else if (kind == DODRT_KIND_SPHERE) { // sphere.cpp:153-157
        const float *lane = s.sphere_lanes + (size_t)(id >> 3) * 32;
        const float w[3] = {P[0] - lane[id & 7u], P[1] - lane[8 + (id & 7u)], P[2] - lane[16 + (id & 7u)]};
        normalize3(w, N);
        C[0] = s.sphere_colors[id * 3], C[1] = s.sphere_colors[id * 3 + 1], C[2] = s.sphere_colors[id * 3 + 2];
    } else if (kind == DODRT_KIND_PLANE) { // plane.cpp:123-130
        const float *lane = s.plane_lanes + (size_t)(id >> 3) * 48;
        N[0] = lane[24 + (id & 7u)], N[1] = lane[32 + (id & 7u)], N[2] = lane[40 + (id & 7u)];
        C[0] = s.plane_colors[id * 3], C[1] = s.plane_colors[id * 3 + 1], C[2] = s.plane_colors[id * 3 + 2];
    } else if (kind == DODRT_KIND_CYLINDER) { // colour stays (0,0,0): cylinder.cpp:172-179,204
        cylinder_normal(s.cylinders[id], s.epsilon, o, d, P, N);
    } else { // box extension: no reference shading; flat grey facing the ray
        N[0] = -d[0], N[1] = -d[1], N[2] = -d[2];
        C[0] = C[1] = C[2] = 0.5f;
    }
}

__global__ void render_init_kernel(const RenderParams p)
{
    const uint64_t pix = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; // slot of the call's share
    if (pix >= p.slots) {
        return;
    }
    uint32_t col, row;
    uint64_t slot;
    const bool inside = slot_to_pixel(p.frame, p.tiles_x, nullptr, pix, col, row, slot);
    float d[3] = {0.0f, 0.0f, 1.0f};
    if (inside) {
        primary_dir(__ldg(p.xs + col), __ldg(p.ys + row), d);
    }
    float4 *ray = reinterpret_cast<float4 *>(p.rays + pix);
    ray[0] = make_float4(p.frame.origin[0], p.frame.origin[1], p.frame.origin[2], d[0]);
    ray[1] = make_float4(d[1], d[2], kInf, __uint_as_float(inside ? 0u : DODRT_RAY_SKIP));
    p.accum[pix] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// One bounce of main.cpp:322-333 for every live pixel.
__global__ void render_shade_kernel(const RenderParams p, uint32_t bounce)
{
    const uint64_t pix = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.slots) {
        return;
    }
    float4 *ray = reinterpret_cast<float4 *>(p.rays + pix);
    const float4 r0 = ray[0], r1 = ray[1];
    if (__float_as_uint(r1.w) & DODRT_RAY_SKIP) {
        return;
    }
    const float4 h = reinterpret_cast<const float4 *>(p.hits)[pix];
    const uint32_t prim = __float_as_uint(h.y);
    if (prim == DODRT_MISS) { // main.cpp:322-325: break
        ray[1] = make_float4(r1.x, r1.y, r1.z, __uint_as_float(DODRT_RAY_SKIP));
        return;
    }
    const float o[3] = {r0.x, r0.y, r0.z}, d[3] = {r0.w, r1.x, r1.y};
    float P[3], N[3], C[3];
    rebuild_record(p.scene, o, d, h.x, prim, h.z, h.w, P, N, C);

    // getLightingFactor, main.cpp:221-244; rayDir is the un-normalised RASTER direction of the pixel (main.cpp:328)
    uint32_t col, row;
    uint64_t slot;
    slot_to_pixel(p.frame, p.tiles_x, nullptr, pix, col, row, slot); // a live ray belongs to a pixel inside the frame
    const float raster[3] = {__ldg(p.xs + col), __ldg(p.ys + row), 1.0f};
    float lighting = 0.2f; // shadeAmbientFactor
    for (uint32_t l = 0; l < p.num_lights; l++) {
        if (!p.visible[(uint64_t)l * p.slots + pix]) {
            continue;
        }
        const float L[3] = {p.lights[l][0] - P[0], p.lights[l][1] - P[1], p.lights[l][2] - P[2]};
        const float distanceFactor = p.lights[l][3] / dot3(L[0], L[1], L[2], L[0], L[1], L[2]);
        float lightDir[3];
        normalize3(L, lightDir);
        const float nd = dot3(N[0], N[1], N[2], lightDir[0], lightDir[1], lightDir[2]);
        const float diffuse = 0.0f < nd ? nd : 0.0f; // std::max(0.0f, x)
        float refl[3];
        reflect3(lightDir, N, refl);
        const float sd = dot3(refl[0], refl[1], refl[2], raster[0], raster[1], raster[2]);
        const float sbase = 0.0f < sd ? sd : 0.0f;
        const float specular = (float)pow((double)sbase, 7.0); // glm::pow(float,int) -> std::pow in double
        float single = 0.0f;
        single += diffuse;
        single += specular;
        single *= distanceFactor;
        lighting += single;
    }
    const float weight = (float)(1.0 / pow(2.0, (double)bounce)); // 1.0f / pow(2.0f, k), exact
    float4 acc = p.accum[pix];
    acc.x = ((1.0f - weight) * acc.x) + (weight * (C[0] * lighting));
    acc.y = ((1.0f - weight) * acc.y) + (weight * (C[1] * lighting));
    acc.z = ((1.0f - weight) * acc.z) + (weight * (C[2] * lighting));
    p.accum[pix] = acc;

    float nd2[3];
    reflect3(d, N, nd2); // main.cpp:332
    ray[0] = make_float4(P[0] + nd2[0] * p.scene.epsilon, P[1] + nd2[1] * p.scene.epsilon, P[2] + nd2[2] * p.scene.epsilon,
                         nd2[0]); // main.cpp:333
    ray[1] = make_float4(nd2[1], nd2[2], kInf, __uint_as_float(0u));
}

__global__ void render_finish_kernel(const RenderParams p)
{
    const uint64_t pix = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= p.slots) {
        return;
    }
    uint32_t col, row;
    uint64_t slot;
    const bool inside = slot_to_pixel(p.frame, p.tiles_x, nullptr, pix, col, row, slot);
    if (!inside && p.rgb_by_pixel) {
        return; // a padded slot of an edge tile has no pixel
    }
    const uint64_t out = p.rgb_by_pixel ? (uint64_t)row * p.frame.width + col : pix;
    const float4 acc = p.accum[pix];
    const float c[3] = {acc.x, acc.y, acc.z};
#pragma unroll
    for (int k = 0; k < 3; k++) { // toOutputChannelType, main.cpp:168-171: clamp(in*255, 0, 255) then truncation
        float x = c[k] * 255.0f;
        x = (x < 0.0f) ? 0.0f : x;
        x = (255.0f < x) ? 255.0f : x;
        p.rgb[out * 3 + k] = (uint8_t)x;
    }
}

// ---- spatial order of a bounce's hit points (counting sort over Morton cells of the room) ----------------------------
// After the first bounce the rays of neighbouring pixels have scattered: a warp's 32 shadow rays start all over the scene
// and share no kd nodes or lanes (bounces 6-10 of the dragon frame: 1.2 Grays/s against 4 for the first).  Shadow rays of
// one light that START close together stay together, so the hit points are binned into (1 << bits)^3 Morton cells of the scene
// box (the six planes of the reference's room span [-5, 5]^3; anything outside is clamped) and the shadow passes of the
// bounce walk the rays cell by cell.  Only the processing order changes: every ray's result lands at its own index.
__device__ __forceinline__ uint32_t spread10(uint32_t v) // up to 10 bits -> every third bit
{
    v &= 1023u;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// cell of slot `pix`'s hit point: Morton code of a (1 << bits)^3 grid over [-5, 5]^3, or the last cell for a ray that casts
// no shadow ray (dead or missed)
__device__ __forceinline__ uint32_t render_cell(const RenderParams &p, uint64_t pix, uint32_t bits)
{
    const uint32_t cells = 1u << (3u * bits);
    const float4 *ray = reinterpret_cast<const float4 *>(p.rays + pix);
    const float4 r0 = ray[0], r1 = ray[1];
    const float4 h = reinterpret_cast<const float4 *>(p.hits)[pix];
    if ((__float_as_uint(r1.w) & DODRT_RAY_SKIP) || __float_as_uint(h.y) == DODRT_MISS) {
        return cells - 1u;
    }
    const float P[3] = {r0.x + r0.w * h.x, r0.y + r1.x * h.x, r0.z + r1.y * h.x};
    const float scale = (float)(1u << bits) * 0.1f, top = (float)((1u << bits) - 1u);
    uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float g = (P[k] + 5.0f) * scale; // [-5, 5] -> [0, 2^bits)
        c[k] = g > 0.0f ? (g < top ? (uint32_t)g : (uint32_t)top) : 0u;
    }
    return spread10(c[0]) | (spread10(c[1]) << 1) | (spread10(c[2]) << 2);
}

__global__ void render_bin_count_kernel(const RenderParams p, uint32_t bits, uint32_t *bins)
{
    const uint64_t pix = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix < p.slots) {
        atomicAdd(bins + render_cell(p, pix, bits), 1u);
    }
}

// exclusive prefix sum of the `cells` counts into bins[cells ...] (one block, 1024 threads x cells / 1024 cells each)
__global__ void render_bin_scan_kernel(uint32_t *bins, uint32_t cells)
{
    __shared__ uint32_t partial[1024];
    const uint32_t t = threadIdx.x, per = cells / 1024u;
    uint32_t sum = 0;
    for (uint32_t k = 0; k < per; k++) {
        sum += bins[t * per + k];
    }
    partial[t] = sum;
    __syncthreads();
    for (uint32_t off = 1; off < 1024u; off <<= 1) { // Hillis-Steele inclusive scan
        const uint32_t v = t >= off ? partial[t - off] : 0u;
        __syncthreads();
        partial[t] += v;
        __syncthreads();
    }
    uint32_t run = partial[t] - sum;
    for (uint32_t k = 0; k < per; k++) {
        const uint32_t c = bins[t * per + k];
        bins[cells + t * per + k] = run;
        run += c;
    }
}

__global__ void render_bin_scatter_kernel(const RenderParams p, uint32_t bits, uint32_t *bins, uint32_t *order)
{
    const uint64_t pix = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix < p.slots) {
        order[atomicAdd(bins + (1u << (3u * bits)) + render_cell(p, pix, bits), 1u)] = (uint32_t)pix;
    }
}

} // namespace

cudaError_t launch_render_sort(const RenderParams &p, uint32_t bits, uint32_t *bins, uint32_t *order, cudaStream_t stream)
{
    const uint64_t n = p.slots;
    const uint32_t cells = 1u << (3u * bits);
    cudaError_t e = cudaMemsetAsync(bins, 0, sizeof(uint32_t) * cells, stream);
    if (e != cudaSuccess) return e;
    render_bin_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, bits, bins);
    render_bin_scan_kernel<<<1, 1024, 0, stream>>>(bins, cells);
    render_bin_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, bits, bins, order);
    return cudaGetLastError();
}

cudaError_t launch_render_init(const RenderParams &p, cudaStream_t stream)
{
    const uint64_t n = p.slots;
    render_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_render_shade(const RenderParams &p, uint32_t bounce, cudaStream_t stream)
{
    const uint64_t n = p.slots;
    render_shade_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(p, bounce);
    return cudaGetLastError();
}

cudaError_t launch_render_finish(const RenderParams &p, cudaStream_t stream)
{
    const uint64_t n = p.slots;
    render_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p);
    return cudaGetLastError();
}

} // namespace dodrt

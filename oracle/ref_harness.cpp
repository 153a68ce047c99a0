// TEST INFRASTRUCTURE -- not product code.  Nothing under dod_raytracer_b200/ may link or load this.
//
// C-callable harness around the UNMODIFIED reference translation units
// (/root/reference/src/{main,shapes/*,accelerators/kdtree}.cpp), compiled where they lie by
// oracle/Makefile into oracle/_ref/libdodrt_ref.so against the glm/assimp stand-ins in oracle/shim/.
// It replaces only main(): the reference's main.cpp is compiled with -Dmain=dodrt_reference_main so
// its scene generators (main.cpp:26-146) and rayTrace (main.cpp:273-347) stay callable.
//
// What it gives the tests:
//   * scene construction through the reference's own create()/Mesh::Create/KDTree::buildTree,
//   * export of the built tree (m_nodes, m_primNums, m_bounds) and the re-ordered triangle lanes
//     (private members, reached with `#define private public` in this TU only),
//   * batched ray queries that call the reference's own intersect functions in the order of
//     main.cpp:314-321 (closest hit) and main.cpp:198-217 (any hit).
// The reference never exposes a primitive id (triangle.cpp:37,136 / sphere.cpp:29,137 keep it in
// a local).  The harness recovers it WITHOUT touching the reference code: in "probe" mode the
// per-triangle attributes are rewritten so that the record's colour carries the id
// (triangle.cpp:166-170 copies Mesh::m_meshAttributes[meshAttrIdx].color) and the normals are the
// identity matrix, so hitNormal = mat3(e0,e1,e2)*bary = (1-(u+v), u, v) exactly (triangle.cpp:171-174).
// Spheres/planes are identified by their (unique) colours; sphere lane data lives in an anonymous
// namespace (sphere.cpp:11-24), so the harness keeps a shadow copy of everything it creates.
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <span>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "assimp/Importer.hpp"
#include "glm/glm.hpp"

#define private public
#define protected public
#include "base_shape.h"
#include "box.h"
#include "config.h"
#include "cylinder.h"
#include "hitrecord.h"
#include "kdtree.h"
#include "light.h"
#include "mesh.h"
#include "plane.h"
#include "sphere.h"
#include "triangle.h"
#undef private
#undef protected

// external-linkage functions of the reference's main.cpp (compiled with -Dmain=dodrt_reference_main)
struct RayTraceData {
    uint8_t *imageData;
    unsigned startRow;
    unsigned endRow;
    const KDTree *tree;
};
void generateSpheres(std::vector<unsigned> &sphereIds, unsigned numSpheres);
void generatePlanes(std::vector<unsigned> &planeIds);
void generateCylinders(std::vector<unsigned> &cylinderIds);
void rayTrace(const RayTraceData data);

namespace {

enum : uint32_t { CLS_SPHERE = 1, CLS_PLANE = 2, CLS_CYLINDER = 4, CLS_TREE = 8 };
enum : uint32_t { KIND_TRIANGLE = 0, KIND_SPHERE = 1, KIND_PLANE = 2, KIND_CYLINDER = 3, KIND_SHIFT = 29 };
constexpr uint32_t MISS = 0xFFFFFFFFu;

struct RefRay {
    float o[3];
    float d[3];
    float clip;
    uint32_t flags; // bit0 = returnOnAny
};
struct RefHit {
    float t;
    uint32_t prim; // kind<<29 | id, or MISS
    float u, v;
};
struct RefRecord { // the reference's HitRecord, flattened (hitrecord.h:4-10) + hit flag
    float t;
    float color[3];
    float normal[3];
    float point[3];
    uint32_t hit;
};

struct ShadowSphere {
    float pos[3];
    float radius;
    float color[3]; // colour stored inside the reference (the lookup key)
};

const KDTree *g_tree = nullptr;
std::vector<ShadowSphere> g_spheres;
std::vector<std::array<float, 3>> g_planeColors;
std::vector<Triangle::Attributes> g_realTriAttrs;
std::vector<Mesh::Attributes> g_realMeshAttrs;
bool g_probe = false;

struct ColorKey {
    uint32_t r, g, b;
    bool operator==(const ColorKey &o) const { return r == o.r && g == o.g && b == o.b; }
};
struct ColorKeyHash {
    size_t operator()(const ColorKey &k) const { return (size_t(k.r) * 0x9E3779B1u) ^ (size_t(k.g) << 17) ^ (size_t(k.b) * 0x85EBCA77u); }
};
std::unordered_map<ColorKey, uint32_t, ColorKeyHash> g_colorToPrim;

uint32_t fbits(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
ColorKey keyOf(const glm::vec3 &c) { return ColorKey{fbits(c.x), fbits(c.y), fbits(c.z)}; }

void registerSphere(const glm::vec3 &pos, float radius, const glm::vec3 &color)
{
    ShadowSphere s{{pos.x, pos.y, pos.z}, radius, {color.x, color.y, color.z}};
    uint32_t id = static_cast<uint32_t>(g_spheres.size());
    g_spheres.push_back(s);
    g_colorToPrim[keyOf(color)] = (KIND_SPHERE << KIND_SHIFT) | id;
}

void setProbe(bool on)
{
    if (on == g_probe) {
        return;
    }
    if (on) {
        g_realTriAttrs = Triangle::m_triangleAttributes;
        g_realMeshAttrs = Mesh::m_meshAttributes;
        size_t numLanes = Triangle::m_triangleLanes.size();
        Mesh::m_meshAttributes.assign(numLanes * 8, Mesh::Attributes{});
        for (size_t lane = 0; lane < numLanes; lane++) {
            Triangle::Attributes &a = Triangle::m_triangleAttributes[lane];
            for (unsigned j = 0; j < 8; j++) {
                uint32_t id = static_cast<uint32_t>(lane * 8 + j);
                a.meshAttrIdx[j] = id;
                a.AN[j] = glm::vec3(1.0f, 0.0f, 0.0f);
                a.BN[j] = glm::vec3(0.0f, 1.0f, 0.0f);
                a.CN[j] = glm::vec3(0.0f, 0.0f, 1.0f);
                // exact small integers in fp32; z = -1 marks "triangle"
                Mesh::m_meshAttributes[id].color = glm::vec3(float(id & 0xFFFFu), float(id >> 16), -1.0f);
            }
        }
    } else {
        Triangle::m_triangleAttributes = g_realTriAttrs;
        Mesh::m_meshAttributes = g_realMeshAttrs;
    }
    g_probe = on;
}

// closest-hit chain, main.cpp:312-321 ; any-hit chain, main.cpp:198-217
inline bool runChain(_Intersect &in, uint32_t classes)
{
    const bool any = in.returnOnAny;
    bool hit = false;
    if (classes & CLS_SPHERE) {
        hit |= Sphere::intersect(in);
        if (any && hit) return true;
        if (!any) in.clippingDistance = in.record.t;
    }
    if (classes & CLS_PLANE) {
        hit |= Plane::intersect(in);
        if (any && hit) return true;
        if (!any) in.clippingDistance = in.record.t;
    }
    if (classes & CLS_CYLINDER) {
        hit |= Cylinder::intersect(in);
        if (any && hit) return true;
        if (!any) in.clippingDistance = in.record.t;
    }
    if ((classes & CLS_TREE) && g_tree) {
        hit |= g_tree->intersect(in);
    }
    return hit;
}

template <typename F> void parallelFor(uint64_t n, int nthreads, F &&body)
{
    if (nthreads <= 1 || n < 2) {
        body(uint64_t(0), n);
        return;
    }
    uint64_t chunk = (n + nthreads - 1) / nthreads; // contiguous bands, main.cpp:371-393
    std::vector<std::thread> threads;
    for (int t = 0; t < nthreads; t++) {
        uint64_t lo = std::min<uint64_t>(n, chunk * t), hi = std::min<uint64_t>(n, lo + chunk);
        if (lo >= hi) break;
        threads.emplace_back([=, &body] { body(lo, hi); });
    }
    for (auto &th : threads) th.join();
}

} // namespace

extern "C" {

void ref_set_config(unsigned width, unsigned height)
{
    Config::Width = width;
    Config::Height = height;
    Config::Ratio = (float)width / height; // config.h:27
}

void ref_seed(unsigned seed) { srand(seed); }

// main.cpp:364 with a fixed seed instead of time(NULL) (main.cpp:351).  The reference draws
// r,g,b,x,y,z per sphere (main.cpp:30-37); the draw sequence is replayed to build the shadow copy.
void ref_add_reference_spheres(unsigned seed, unsigned count)
{
    std::vector<unsigned> ids;
    srand(seed);
    generateSpheres(ids, count);
    srand(seed);
    for (unsigned i = 0; i < count; i++) {
        float r = ((float)rand() / RAND_MAX);
        float g = ((float)rand() / RAND_MAX);
        float b = ((float)rand() / RAND_MAX);
        float x = ((float)rand() / RAND_MAX) * 10.0f - 5.0f;
        float y = ((float)rand() / RAND_MAX) * 10.0f - 5.0f;
        float z = ((float)rand() / RAND_MAX) * 10.0f - 5.0f;
        registerSphere(glm::vec3(x, y, z), 1.0f, glm::vec3(r, g, b));
    }
}

// An explicit sphere through Sphere::create (sphere.cpp:226-242).  The colour stored inside the
// reference is an id tag (x = id & 0xFFFF, y = id >> 16, z = -2); shading colours are not needed
// for hit parity.
void ref_add_sphere(const float pos[3], float radius)
{
    uint32_t id = static_cast<uint32_t>(g_spheres.size());
    glm::vec3 tag(float(id & 0xFFFFu), float(id >> 16), -2.0f);
    Sphere::_Create c{.position = glm::vec3(pos[0], pos[1], pos[2]), .radius = radius, .attributes = {tag}};
    Sphere::create(c);
    registerSphere(c.position, radius, tag);
}

void ref_add_reference_planes()
{
    std::vector<unsigned> ids;
    generatePlanes(ids); // main.cpp:52-109
    static const float colors[6][3] = {{0.195f, 0.410f, 0.610f}, {0.493, 0.265, 0.590}, {0.276, 0.600, 0.411},
                                       {0.292, 0.680, 0.674},   {0.720, 0.288, 0.389}, {0.680, 0.224, 0.224}};
    for (int i = 0; i < 6; i++) {
        uint32_t id = static_cast<uint32_t>(g_planeColors.size());
        g_planeColors.push_back({colors[i][0], colors[i][1], colors[i][2]});
        g_colorToPrim[keyOf(glm::vec3(colors[i][0], colors[i][1], colors[i][2]))] = (KIND_PLANE << KIND_SHIFT) | id;
    }
}

void ref_add_reference_cylinder()
{
    std::vector<unsigned> ids;
    generateCylinders(ids); // main.cpp:111-129 (draws 3 rand() for a colour that is never used)
}

// Mesh::Create (mesh.cpp:9-50) through the stand-in importer; returns the number of triangles added.
int ref_add_mesh(const char *path)
{
    unsigned before = Triangle::m_numTriangles;
    std::string p(path);
    Mesh::_Create c{.loadPath = p};
    Mesh::Create(c);
    return static_cast<int>(Triangle::m_numTriangles - before);
}

void ref_build_tree()
{
    setProbe(false);
    g_tree = new KDTree(KDTree::buildTree()); // kdtree.cpp:252-260 (also re-orders the lanes)
}

void ref_tree_sizes(uint32_t *numNodes, uint32_t *numLanes, uint32_t *numPrimNums, uint32_t *maxDepth, uint32_t *numTriangles)
{
    *numNodes = g_tree ? static_cast<uint32_t>(g_tree->m_nodes.size()) : 0;
    *numLanes = static_cast<uint32_t>(Triangle::m_triangleLanes.size());
    *numPrimNums = g_tree ? static_cast<uint32_t>(g_tree->m_primNums.size()) : 0;
    *maxDepth = g_tree ? g_tree->m_maxDepth : 0;
    *numTriangles = Triangle::m_numTriangles;
}

// nodes: 8 B each (kdtree.h:16-48); lanes: 288 B each (triangle.h:33-44), in re-ordered order;
// primNums: original lane number of every re-ordered lane (kdtree.cpp:51-55); bounds: min xyz, max xyz.
void ref_tree_export(uint64_t *nodes, float *lanes, uint32_t *primNums, float bounds[6])
{
    static_assert(sizeof(KDTree::Node) == 8, "node layout");
    static_assert(sizeof(Triangle::TriangleLane) == 288, "lane layout");
    if (g_tree) {
        std::memcpy(nodes, g_tree->m_nodes.data(), g_tree->m_nodes.size() * 8);
        std::memcpy(primNums, g_tree->m_primNums.data(), g_tree->m_primNums.size() * 4);
        for (int i = 0; i < 3; i++) {
            bounds[i] = g_tree->m_bounds.minCorner[i];
            bounds[3 + i] = g_tree->m_bounds.maxCorner[i];
        }
    }
    std::memcpy(lanes, Triangle::m_triangleLanes.data(), Triangle::m_triangleLanes.size() * 288);
}

// per re-ordered triangle slot: AN xyz, BN xyz, CN xyz (triangle.h:45-51), 9 floats
void ref_normals_export(float *normals)
{
    const std::vector<Triangle::Attributes> &attrs = g_probe ? g_realTriAttrs : Triangle::m_triangleAttributes;
    for (size_t lane = 0; lane < attrs.size(); lane++) {
        for (unsigned j = 0; j < 8; j++) {
            float *o = normals + (lane * 8 + j) * 9;
            const glm::vec3 *n[3] = {&attrs[lane].AN[j], &attrs[lane].BN[j], &attrs[lane].CN[j]};
            for (int k = 0; k < 3; k++) {
                o[k * 3 + 0] = n[k]->x;
                o[k * 3 + 1] = n[k]->y;
                o[k * 3 + 2] = n[k]->z;
            }
        }
    }
}

// Triangle::m_triangleAttributes (re-ordered) byte for byte, 320 B per lane, and the mesh colours
void ref_attrs_export(void *attrs, float *meshColors)
{
    const std::vector<Triangle::Attributes> &a = g_probe ? g_realTriAttrs : Triangle::m_triangleAttributes;
    const std::vector<Mesh::Attributes> &m = g_probe ? g_realMeshAttrs : Mesh::m_meshAttributes;
    static_assert(sizeof(Triangle::Attributes) == 320, "attribute layout");
    std::memcpy(attrs, a.data(), a.size() * sizeof(Triangle::Attributes));
    for (size_t i = 0; i < m.size(); i++) {
        meshColors[i * 3 + 0] = m[i].color.x;
        meshColors[i * 3 + 1] = m[i].color.y;
        meshColors[i * 3 + 2] = m[i].color.z;
    }
}
uint32_t ref_num_meshes() { return static_cast<uint32_t>((g_probe ? g_realMeshAttrs : Mesh::m_meshAttributes).size()); }

uint32_t ref_num_spheres() { return static_cast<uint32_t>(g_spheres.size()); }
// x, y, z, radius, r, g, b per sphere (shadow copy of what went through Sphere::create)
void ref_spheres_export(float *out)
{
    for (size_t i = 0; i < g_spheres.size(); i++) {
        std::memcpy(out + i * 7, &g_spheres[i], 7 * sizeof(float));
    }
}

// Primary rays exactly as main.cpp:275-279,294-310,342-345 for a single band starting at row 0:
// rayDir.x accumulates += widthStep per column, rayDir.y -= heightStep per row, glm::normalize per pixel.
void ref_primary_rays(RefRay *out)
{
    glm::vec3 rayDir = {-Config::Ratio, 1.0f, 1};
    float widthStep = 2.0f * Config::Ratio / Config::Width;
    float heightStep = 2.0f / Config::Height;
    uint64_t k = 0;
    for (unsigned i = 0; i < Config::Height; i++) {
        for (unsigned j = 0; j < Config::Width; j++) {
            glm::vec3 rayNorm = glm::normalize(rayDir);
            RefRay &r = out[k++];
            r.o[0] = 0;
            r.o[1] = 0;
            r.o[2] = -4.9;
            r.d[0] = rayNorm.x;
            r.d[1] = rayNorm.y;
            r.d[2] = rayNorm.z;
            r.clip = std::numeric_limits<float>::infinity();
            r.flags = 0;
            rayDir.x += widthStep;
        }
        rayDir.x = -Config::Ratio;
        rayDir.y -= heightStep;
    }
}

// Shadow rays exactly as canSeeLight builds them (main.cpp:184-196) from hit points.
void ref_shadow_rays(const float *points, uint64_t n, const float light[3], RefRay *out)
{
    glm::vec3 lightPos(light[0], light[1], light[2]);
    for (uint64_t i = 0; i < n; i++) {
        glm::vec3 hitPoint(points[i * 3], points[i * 3 + 1], points[i * 3 + 2]);
        glm::vec3 lightDir = lightPos - hitPoint;
        float lightDistance = glm::length(lightDir);
        lightDir /= lightDistance;
        glm::vec3 origin = hitPoint + lightDir * 0.01f;
        RefRay &r = out[i];
        r.o[0] = origin.x; r.o[1] = origin.y; r.o[2] = origin.z;
        r.d[0] = lightDir.x; r.d[1] = lightDir.y; r.d[2] = lightDir.z;
        r.clip = lightDistance;
        r.flags = 1;
    }
}

// Batched query with ids (probe mode).  `classes` selects Sphere|Plane|Cylinder|Tree; bit0 of each
// ray's flags is returnOnAny.  For any-hit rays prim is 0 (some hit) or MISS: the reference does not
// define which primitive an any-hit query reports.
void ref_intersect(const RefRay *rays, uint64_t n, uint32_t classes, RefHit *out, int nthreads)
{
    setProbe(true);
    parallelFor(n, nthreads, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            HitRecord hr;
            hr.t = rays[i].clip;
            hr.color = glm::vec3(0);
            hr.hitNormal = glm::vec3(0);
            hr.hitPoint = glm::vec3(0);
            _Intersect in{.rayDir = glm::vec3(rays[i].d[0], rays[i].d[1], rays[i].d[2]),
                          .rayOrigin = glm::vec3(rays[i].o[0], rays[i].o[1], rays[i].o[2]),
                          .returnOnAny = (rays[i].flags & 1u) != 0,
                          .clippingDistance = rays[i].clip,
                          .record = hr};
            bool hit = runChain(in, classes);
            RefHit &h = out[i];
            if (!hit) {
                h.t = rays[i].clip;
                h.prim = MISS;
                h.u = h.v = 0;
                continue;
            }
            h.t = hr.t;
            h.u = h.v = 0;
            if (in.returnOnAny) {
                h.prim = 0;
                continue;
            }
            if (hr.color.z == -1.0f) {
                uint32_t id = uint32_t(hr.color.x) | (uint32_t(hr.color.y) << 16);
                h.prim = (KIND_TRIANGLE << KIND_SHIFT) | id;
                h.u = hr.hitNormal.y;
                h.v = hr.hitNormal.z;
            } else {
                auto it = g_colorToPrim.find(keyOf(hr.color));
                if (it != g_colorToPrim.end()) {
                    h.prim = it->second;
                } else {
                    h.prim = (KIND_CYLINDER << KIND_SHIFT); // cylinder.cpp:172-179 leaves colour 0
                }
            }
        }
    });
}

// Batched query returning the reference's real HitRecord (real normals / colours).
void ref_intersect_records(const RefRay *rays, uint64_t n, uint32_t classes, RefRecord *out, int nthreads)
{
    setProbe(false);
    parallelFor(n, nthreads, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            HitRecord hr;
            hr.t = rays[i].clip;
            hr.color = glm::vec3(0);
            hr.hitNormal = glm::vec3(0);
            hr.hitPoint = glm::vec3(0);
            _Intersect in{.rayDir = glm::vec3(rays[i].d[0], rays[i].d[1], rays[i].d[2]),
                          .rayOrigin = glm::vec3(rays[i].o[0], rays[i].o[1], rays[i].o[2]),
                          .returnOnAny = (rays[i].flags & 1u) != 0,
                          .clippingDistance = rays[i].clip,
                          .record = hr};
            bool hit = runChain(in, classes);
            RefRecord &r = out[i];
            r.hit = hit ? 1u : 0u;
            r.t = hr.t;
            for (int k = 0; k < 3; k++) {
                r.color[k] = hr.color[k];
                r.normal[k] = hr.hitNormal[k];
                r.point[k] = hr.hitPoint[k];
            }
        }
    });
}

// The reference's own per-band renderer (main.cpp:273-347): 9 lights, 10 bounces, all shape classes.
// One call = one band, so a canonical image is rows [0,H) in a single call.
void ref_render_rows(uint8_t *image, unsigned startRow, unsigned endRow)
{
    setProbe(false);
    if (!g_tree) {
        ref_build_tree(); // mesh.cpp:17-21: a missing mesh still gets an (empty) tree, main.cpp:368
    }
    RayTraceData data{image, startRow, endRow, g_tree};
    rayTrace(data);
}

// The reference's threading (main.cpp:371-394): ceil(H/n) rows per band, one thread per band.
void ref_render_bands(uint8_t *image, int nthreads)
{
    setProbe(false);
    unsigned H = Config::Height;
    unsigned rows = (H + nthreads - 1) / nthreads;
    std::vector<std::thread> threads;
    for (unsigned start = 0; start < H; start += rows) {
        unsigned end = std::min(H, start + rows);
        threads.emplace_back([=] { ref_render_rows(image, start, end); });
    }
    for (auto &t : threads) t.join();
}

// One frame the way the reference's rayTrace does it (main.cpp:297-334) reduced to bounce k = 0 and one
// light: per pixel the closest-hit chain, then canSeeLight's any-hit chain from the hit point.  Rows are
// split into contiguous bands over threads exactly like main.cpp:371-393.  This is the CPU baseline the
// benchmark times; ids are not recovered here (real attributes stay in place).
void ref_trace_frame(uint32_t classes, const float light[3], float *tOut, uint8_t *visible, int nthreads)
{
    setProbe(false);
    const unsigned W = Config::Width, H = Config::Height;
    std::vector<float> xs(W), ys(H);
    {
        glm::vec3 rayDir = {-Config::Ratio, 1.0f, 1};
        float widthStep = 2.0f * Config::Ratio / Config::Width;
        float heightStep = 2.0f / Config::Height;
        for (unsigned j = 0; j < W; j++) { xs[j] = rayDir.x; rayDir.x += widthStep; }
        for (unsigned i = 0; i < H; i++) { ys[i] = rayDir.y; rayDir.y -= heightStep; }
    }
    const glm::vec3 lightPos(light[0], light[1], light[2]);
    parallelFor(H, nthreads, [&](uint64_t rowLo, uint64_t rowHi) {
        for (uint64_t i = rowLo; i < rowHi; i++) {
            for (unsigned j = 0; j < W; j++) {
                HitRecord hr;
                hr.t = std::numeric_limits<float>::infinity();
                _Intersect in{.rayDir = glm::normalize(glm::vec3(xs[j], ys[i], 1.0f)),
                              .rayOrigin = {0, 0, -4.9},
                              .record = hr};
                bool hit = runChain(in, classes);
                const uint64_t k = i * W + j;
                tOut[k] = hit ? hr.t : std::numeric_limits<float>::infinity();
                uint8_t vis = 0;
                if (hit) {
                    glm::vec3 lightDir = lightPos - hr.hitPoint;
                    float lightDistance = glm::length(lightDir);
                    lightDir /= lightDistance;
                    HitRecord shr;
                    shr.t = lightDistance;
                    _Intersect sin{.rayDir = lightDir,
                                   .rayOrigin = hr.hitPoint + lightDir * 0.01f,
                                   .returnOnAny = true,
                                   .clippingDistance = lightDistance,
                                   .record = shr};
                    vis = runChain(sin, classes) ? 0 : 1;
                }
                visible[k] = vis;
            }
        }
    });
}

int ref_hardware_threads() { return static_cast<int>(std::thread::hardware_concurrency()); }

} // extern "C"

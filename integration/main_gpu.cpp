// main_gpu.cpp -- the reference's program with its hot path on the GPU: a replacement for main() of
// /root/reference/src/main.cpp:349-397 that a maintainer of AVassilev98/dod_raytracer could ship next to it.
//
// Everything on the HOST side is the reference's own, unmodified code (compiled from /root/reference/src by
// integration/Makefile exactly like oracle/Makefile does): Config::Load (config.h:16-37), the scene registration through
// Sphere::create / Plane::create / Cylinder::create with the values of generateSpheres / generatePlanes /
// generateCylinders (main.cpp:26-129), Mesh::Create (mesh.cpp:9-50), KDTree::buildTree (kdtree.cpp:252-260) and
// stbi_write_png (main.cpp:396).  The per-pixel loop rayTrace (main.cpp:273-347) is replaced by ONE call into
// libdodrt_cuda.so through the adapter of dodrt_adapter.hpp: dodrt_render = up to 10 mirror bounces x (closest-hit chain
// + 9 canSeeLight queries + shading), 8-bit RGB back.  `--cpu-raw FILE` also runs the reference's own rayTrace (one row
// band = the canonical raster tables) so that a test can compare the two images.
//
//   dod_raytracer_gpu [--config config.ini] [--mesh assets/dragon.obj] [--seed N] [--out output.png]
//                     [--raw gpu.rgb] [--cpu-raw cpu.rgb] [--device D] [--depth 10] [--gpus N]
//
// Differences from the reference's main(), all of them needed for a reproducible run: the seed is an argument (the
// reference uses time(NULL), main.cpp:351), and so are the mesh and config paths (hard-coded next to the executable,
// main.cpp:353-357,135).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "dodrt_adapter.hpp"
#include "dodrt.hpp"
#include "hitrecord.h"
#include "light.h"

// the reference's own functions with external linkage (main.cpp is compiled with -Dmain=dodrt_reference_main)
struct RayTraceData {
    uint8_t *imageData;
    unsigned startRow;
    unsigned endRow;
    const KDTree *tree;
};
void rayTrace(const RayTraceData data);
extern "C" int stbi_write_png(char const *filename, int w, int h, int comp, const void *data, int stride_in_bytes);

using dodrt_integration::check;
using dodrt_integration::DodrtScene;
using dodrt_integration::SceneMirror;

static float frand() { return (float)rand() / RAND_MAX; }

// generateSpheres (main.cpp:26-50): r, g, b, then x, y, z per sphere, radius 1
static void registerSpheres(SceneMirror &m, unsigned count)
{
    for (unsigned i = 0; i < count; i++) {
        const float r = frand(), g = frand(), b = frand();
        const float x = frand() * 10.0f - 5.0f, y = frand() * 10.0f - 5.0f, z = frand() * 10.0f - 5.0f;
        Sphere::_Create c{.position = glm::vec3(x, y, z), .radius = 1.0f, .attributes = {glm::vec3(r, g, b)}};
        m.addSphere(c);
    }
}

// generatePlanes (main.cpp:52-109): the six walls of the room
static void registerPlanes(SceneMirror &m)
{
    const Plane::_Create planes[6] = {
        {.normal = {0.0f, 0.0f, -1.0f}, .position = {0.0f, 0.0f, 5.0f}, .attributes = {.color = {0.195f, 0.410f, 0.610f}}},
        {.normal = {0.0f, 0.0f, 1.0f}, .position = {0.0f, 0.0f, -5.0f}, .attributes = {.color = {0.493, 0.265, 0.590}}},
        {.normal = {0.0f, -1.0f, 0.0f}, .position = {0.0f, 5.0f, 0.0f}, .attributes = {.color = {0.276, 0.600, 0.411}}},
        {.normal = {0.0f, 1.0f, 0.0f}, .position = {0.0f, -5.0f, 0.0f}, .attributes = {.color = {0.292, 0.680, 0.674}}},
        {.normal = {1.0f, 0.0f, 0.0f}, .position = {-5.0f, 0.0f, 0.0f}, .attributes = {.color = {0.720, 0.288, 0.389}}},
        {.normal = {-1.0f, 0.0f, 0.0f}, .position = {5.0f, 0.0f, 0.0f}, .attributes = {.color = {0.680, 0.224, 0.224}}},
    };
    for (const Plane::_Create &c : planes) m.addPlane(c);
}

// generateCylinders (main.cpp:111-129): three rand() draws for a colour the reference never shows (cylinder.cpp:172-179)
static void registerCylinder(SceneMirror &m)
{
    Cylinder::_Create c = {.radius = 1.5f, .height = 4.0f, .axis = {2.2, 5, 2}, .basePosition = {-2, 0, 2},
                           .attributes = {.color = {frand(), frand(), frand()}}};
    m.addCylinder(c);
}

int main(int argc, char **argv)
{
    std::string config, mesh, out = "output.png", raw, cpuRaw;
    unsigned seed = 1, depth = 10;
    int device = 0, gpus = 1;
    for (int i = 1; i < argc; i++) {
        auto arg = [&](const char *name) { return std::strcmp(argv[i], name) == 0 && i + 1 < argc; };
        if (arg("--config")) config = argv[++i];
        else if (arg("--mesh")) mesh = argv[++i];
        else if (arg("--out")) out = argv[++i];
        else if (arg("--raw")) raw = argv[++i];
        else if (arg("--cpu-raw")) cpuRaw = argv[++i];
        else if (arg("--seed")) seed = (unsigned)std::atoi(argv[++i]);
        else if (arg("--depth")) depth = (unsigned)std::atoi(argv[++i]);
        else if (arg("--device")) device = std::atoi(argv[++i]);
        else if (arg("--gpus")) gpus = std::atoi(argv[++i]);
        else {
            std::fprintf(stderr, "unknown argument %s\n", argv[i]);
            return 2;
        }
    }
    srand(seed); // main.cpp:351 (time(NULL) there)
    if (!config.empty()) Config::Load(config); // main.cpp:353-357

    SceneMirror mirror;
    registerSpheres(mirror, 16); // main.cpp:364
    registerPlanes(mirror);      // main.cpp:365
    registerCylinder(mirror);    // main.cpp:366
    if (!mesh.empty()) {         // generateMeshes, main.cpp:131-146
        Mesh::_Create c = {.loadPath = mesh};
        Mesh::Create(c);
    }
    const KDTree tree = KDTree::buildTree(); // main.cpp:368
    const unsigned W = Config::Width, H = Config::Height;
    std::vector<uint8_t> image((size_t)W * H * 3, 0); // main.cpp:369

    // the raster tables rayTrace accumulates (main.cpp:276-279,342-345), canonical single band
    std::vector<float> xs(W), ys(H);
    const float widthStep = 2.0f * Config::Ratio / Config::Width, heightStep = 2.0f / Config::Height;
    float x = -Config::Ratio, y = 1.0f;
    for (unsigned j = 0; j < W; j++, x += widthStep) xs[j] = x;
    for (unsigned i = 0; i < H; i++, y -= heightStep) ys[i] = y;
    const float lights[9][4] = {{0.0f, 0.0f, -2.0f, 3.0f},   {4.0f, 4.3f, 3.3f, 1.0f},    {-4.f, -2.95f, 3.95f, 1.0f},
                                {3.95f, -4.2f, 3.3f, 1.0f},  {-2.9f, 4.2f, 3.8f, 1.0f},   {3.95f, 2.8f, -4.3f, 1.0f},
                                {-3.0f, -3.8f, -3.3f, 1.0f}, {4.2f, -4.2f, -3.4f, 1.0f},  {-2.9f, 4.4f, -3.5f, 1.0f}}; // main.cpp:283-292
    dodrt_frame frame{};
    frame.width = W, frame.height = H, frame.tile_w = 32, frame.tile_h = 32, frame.first_tile = 0, frame.tile_stride = 1;
    frame.classes = DODRT_CLS_SPHERE | DODRT_CLS_PLANE | DODRT_CLS_CYLINDER | DODRT_CLS_TREE;
    frame.origin[0] = 0.0f, frame.origin[1] = 0.0f, frame.origin[2] = (float)-4.9; // main.cpp:275

    const auto t0 = std::chrono::steady_clock::now();
    if (gpus <= 1) {
        DodrtScene gpu(tree, mirror, device);
        check(dodrt_render(gpu.h, &frame, xs.data(), ys.data(), &lights[0][0], 9, depth, image.data()));
    } else { // the frame split over the GPUs of this process like the reference splits it over its threads (main.cpp:371-394)
        std::vector<std::unique_ptr<DodrtScene>> replicas;
        std::vector<dodrt_scene *> handles;
        for (int d = 0; d < gpus; d++) {
            replicas.emplace_back(new DodrtScene(tree, mirror, d));
            handles.push_back(replicas.back()->h);
        }
        dodrt_multi *multi = nullptr;
        check(dodrt_multi_create(handles.data(), (uint32_t)handles.size(), &multi));
        check(dodrt_multi_render(multi, &frame, xs.data(), ys.data(), &lights[0][0], 9, depth, image.data()));
        dodrt_multi_destroy(multi);
    }
    const double gpuMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::printf("dod_raytracer_gpu: %ux%u, %zu triangle lanes, %zu kd nodes, rendered on %d GPU(s) in %.1f ms (incl. upload)\n", W, H,
                Triangle::m_triangleLanes.size(), tree.m_nodes.size(), gpus < 1 ? 1 : gpus, gpuMs);
    if (!stbi_write_png(out.c_str(), (int)W, (int)H, 3, image.data(), (int)W * 3)) { // main.cpp:396
        std::fprintf(stderr, "cannot write %s\n", out.c_str());
        return 1;
    }
    auto dump = [](const std::string &path, const std::vector<uint8_t> &img) {
        FILE *f = std::fopen(path.c_str(), "wb");
        if (!f || std::fwrite(img.data(), 1, img.size(), f) != img.size()) {
            std::fprintf(stderr, "cannot write %s\n", path.c_str());
            std::exit(1);
        }
        std::fclose(f);
    };
    if (!raw.empty()) dump(raw, image);
    if (!cpuRaw.empty()) { // the reference's own per-pixel loop, one band
        std::vector<uint8_t> cpu((size_t)W * H * 3, 0);
        const auto c0 = std::chrono::steady_clock::now();
        rayTrace(RayTraceData{cpu.data(), 0, H, &tree});
        std::printf("reference rayTrace on one host thread: %.1f ms\n",
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count());
        dump(cpuRaw, cpu);
        // 1-ray interface parity (kdtree.h:13 / base_shape.h:17-28) on a few pixels through dodrt.hpp
        DodrtScene gpu(tree, mirror, device, false);
        unsigned bad = 0, tested = 0;
        for (unsigned k = 0; k < 64; k++) {
            const unsigned col = (k * 37u + 11u) % W, row = (k * 53u + 7u) % H;
            const glm::vec3 dir = glm::normalize(glm::vec3(xs[col], ys[row], 1.0f));
            HitRecord a, b;
            _Intersect ia{.rayDir = dir, .rayOrigin = glm::vec3(0, 0, -4.9), .record = a};
            _Intersect ib{.rayDir = dir, .rayOrigin = glm::vec3(0, 0, -4.9), .record = b};
            const bool ha = tree.intersect(ia);
            const bool hb = dodrt::intersect(gpu.h, DODRT_CLS_TREE, ib);
            tested++;
            if (ha != hb || (ha && (std::memcmp(&a.t, &b.t, 4) != 0 || std::memcmp(&a.hitPoint, &b.hitPoint, 12) != 0 ||
                                    std::memcmp(&ia.clippingDistance, &ib.clippingDistance, 4) != 0))) {
                bad++;
            }
        }
        std::printf("one-ray KDTree::intersect vs dodrt::intersect: %u of %u differ\n", bad, tested);
        if (bad) return 3;
    }
    return 0;
}

// dodrt_adapter.hpp -- the adapter translation unit a maintainer of AVassilev98/dod_raytracer adds (INTEGRATION.md
// section 2), in a form that compiles against the UNMODIFIED reference sources: after the scene has been registered and
// KDTree::buildTree() has run (main.cpp:364-368), flatten everything the GPU path needs into the C ABI of dodrt.h.
//
//   * KDTree::m_nodes / m_bounds (kdtree.h:63-68) and Triangle::m_triangleLanes / m_triangleAttributes
//     (triangle.h:59-61) are private statics / members: this TU reaches them with `#define private public` around the
//     reference's headers (a maintainer would add a friend declaration or two accessors instead);
//   * sphere / plane / cylinder storage lives in anonymous namespaces (sphere.cpp:11-24, plane.cpp:10-25,
//     cylinder.cpp:29-32) and cannot be reached from outside without editing those files (INTEGRATION.md shows the
//     two-line accessors).  SceneMirror therefore records every shape AS IT IS CREATED, through the reference's own
//     create() calls, in the reference's own lane layouts (sphere.cpp:226-242, plane.cpp:204-222,
//     cylinder.cpp:211-229).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define private public
#define protected public
#include "cylinder.h"
#include "kdtree.h"
#include "mesh.h"
#include "plane.h"
#include "sphere.h"
#include "triangle.h"
#undef private
#undef protected

#include "config.h"
#include "dodrt.h"

namespace dodrt_integration {

inline void check(int rc)
{
    if (rc != DODRT_OK) {
        std::fprintf(stderr, "dodrt: %s\n", dodrt_last_error());
        std::abort();
    }
}

// What Sphere::create / Plane::create / Cylinder::create store, mirrored in the reference's lane layouts.
struct SceneMirror {
    std::vector<float> sphereLanes; // x[8] y[8] z[8] radiusSq[8] per lane, sphere.cpp:12-19
    std::vector<float> sphereColors;
    uint32_t numSpheres = 0;
    std::vector<float> planeLanes; // px[8] py[8] pz[8] nx[8] ny[8] nz[8] per lane, plane.cpp:13-20
    std::vector<float> planeColors;
    uint32_t numPlanes = 0;
    std::vector<dodrt_cylinder> cylinders;

    static void append(std::vector<float> &lanes, uint32_t index, int floatsPerLane, const float *vals, int nvals)
    {
        const uint32_t lane = index / 8, slot = index % 8;
        if (lanes.size() < (size_t)(lane + 1) * floatsPerLane) lanes.resize((size_t)(lane + 1) * floatsPerLane, 0.0f);
        for (int k = 0; k < nvals; k++) lanes[(size_t)lane * floatsPerLane + k * 8 + slot] = vals[k];
    }
    unsigned addSphere(const Sphere::_Create &c) // Sphere::create, sphere.cpp:226-242
    {
        const float v[4] = {c.position.x, c.position.y, c.position.z, c.radius * c.radius};
        append(sphereLanes, numSpheres++, 32, v, 4);
        sphereColors.insert(sphereColors.end(), {c.attributes.color.x, c.attributes.color.y, c.attributes.color.z});
        return Sphere::create(c);
    }
    unsigned addPlane(const Plane::_Create &c) // Plane::create, plane.cpp:204-222
    {
        const float v[6] = {c.position.x, c.position.y, c.position.z, c.normal.x, c.normal.y, c.normal.z};
        append(planeLanes, numPlanes++, 48, v, 6);
        planeColors.insert(planeColors.end(), {c.attributes.color.x, c.attributes.color.y, c.attributes.color.z});
        return Plane::create(c);
    }
    unsigned addCylinder(const Cylinder::_Create &c) // Cylinder::create -> Cylinder::Cylinder, cylinder.cpp:211-229
    {
        const glm::vec3 axis = glm::normalize(c.axis);
        dodrt_cylinder d{{c.basePosition.x, c.basePosition.y, c.basePosition.z}, {axis.x, axis.y, axis.z}, c.radius * c.radius, c.height};
        cylinders.push_back(d);
        return Cylinder::create(c);
    }
};

// One replica of the finished scene in one GPU.  Owns the dodrt_scene handle.
struct DodrtScene {
    dodrt_scene *h = nullptr;

    DodrtScene(const KDTree &tree, const SceneMirror &m, int device = 0, bool shading = true)
    {
        check(dodrt_scene_create(device, &h));
        static_assert(sizeof(KDTree::Node) == 8, "kdtree.h:16-48");
        static_assert(sizeof(Triangle::TriangleLane) == 288, "triangle.h:33-44");
        static_assert(sizeof(Triangle::Attributes) == 320, "triangle.h:45-51");
        const float bounds[6] = {tree.m_bounds.minCorner.x, tree.m_bounds.minCorner.y, tree.m_bounds.minCorner.z,
                                 tree.m_bounds.maxCorner.x, tree.m_bounds.maxCorner.y, tree.m_bounds.maxCorner.z};
        // lanes AFTER Triangle::reorderLanesByIndices (kdtree.cpp:258): a leaf's lanes are contiguous
        check(dodrt_scene_set_kdtree(h, reinterpret_cast<const uint64_t *>(tree.m_nodes.data()), (uint32_t)tree.m_nodes.size(),
                                     reinterpret_cast<const float *>(Triangle::m_triangleLanes.data()),
                                     (uint32_t)Triangle::m_triangleLanes.size(), bounds));
        check(dodrt_scene_set_spheres(h, m.sphereLanes.data(), m.numSpheres));
        check(dodrt_scene_set_planes(h, m.planeLanes.data(), m.numPlanes));
        check(dodrt_scene_set_cylinders(h, m.cylinders.data(), (uint32_t)m.cylinders.size()));
        check(dodrt_scene_set_epsilon(h, Config::Epsilon));
        if (shading) {
            std::vector<float> meshColors;
            for (const auto &a : Mesh::m_meshAttributes) meshColors.insert(meshColors.end(), {a.color.x, a.color.y, a.color.z});
            check(dodrt_scene_set_shading(h, Triangle::m_triangleAttributes.data(), (uint32_t)Triangle::m_triangleAttributes.size(),
                                          meshColors.data(), (uint32_t)Mesh::m_meshAttributes.size(), m.sphereColors.data(),
                                          m.planeColors.data()));
        }
    }
    ~DodrtScene() { dodrt_scene_destroy(h); }
    DodrtScene(const DodrtScene &) = delete;
    DodrtScene &operator=(const DodrtScene &) = delete;
};

} // namespace dodrt_integration

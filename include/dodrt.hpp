// dodrt.hpp -- header-only C++ convenience over the C ABI of dodrt.h.
//
// 1. dodrt::Scene: RAII handle (dodrt_scene_create / _destroy), errors as exceptions.
// 2. dodrt::intersect(): the ONE-RAY form of the reference's interface,
//        bool KDTree::intersect(_Intersect &) const                (src/accelerators/kdtree.h:13)
//        static bool BaseShape<D>::intersect(_Intersect &)         (src/shapes/base_shape.h:17-28)
//    for API parity and tests: the same in/out contract -- returns hit/miss, writes record.t and record.hitPoint
//    (hitPoint = rayOrigin + rayDir * t, triangle.cpp:170), and for a closest-hit query lowers clippingDistance to the
//    hit distance like kdtree.cpp:343 does.  It works on any struct shaped like `_Intersect` (base_shape.h:8-15):
//    members rayDir, rayOrigin (indexable 0..2), returnOnAny, clippingDistance and record with t and hitPoint --
//    in particular on the reference's own type.  record.color / record.hitNormal need the shading attributes and are
//    produced by dodrt_render; the primitive id the reference keeps in a local (triangle.cpp:37) is handed back through
//    `hit`.  One ray per call means one kernel launch per call: this is for parity checks, not for rendering -- the
//    render loop calls the batch entry points.
#ifndef DODRT_HPP
#define DODRT_HPP

#include <stdexcept>
#include <string>

#include "dodrt.h"

namespace dodrt {

inline void check(int rc)
{
    if (rc != DODRT_OK) {
        throw std::runtime_error(std::string("dodrt error ") + std::to_string(rc) + ": " + dodrt_last_error());
    }
}

class Scene {
public:
    explicit Scene(int device = 0) { check(dodrt_scene_create(device, &h_)); }
    ~Scene() { dodrt_scene_destroy(h_); }
    Scene(const Scene &) = delete;
    Scene &operator=(const Scene &) = delete;
    dodrt_scene *get() const { return h_; }
    operator dodrt_scene *() const { return h_; }

private:
    dodrt_scene *h_ = nullptr;
};

// `classes`: DODRT_CLS_TREE for KDTree::intersect, DODRT_CLS_SPHERE / _PLANE / _CYLINDER for the shape classes, or any
// union for the chain of main.cpp:314-321 (closest) / main.cpp:198-217 (returnOnAny).
template <typename IntersectT> bool intersect(dodrt_scene *scene, uint32_t classes, IntersectT &in, dodrt_hit *hit = nullptr)
{
    dodrt_ray ray;
    for (int k = 0; k < 3; k++) {
        ray.o[k] = in.rayOrigin[k];
        ray.d[k] = in.rayDir[k];
    }
    ray.clip = in.clippingDistance;
    ray.flags = in.returnOnAny ? DODRT_RAY_ANY : 0u;
    dodrt_hit h;
    check(dodrt_intersect(scene, &ray, 1, classes, &h));
    if (hit) {
        *hit = h;
    }
    if (h.prim == DODRT_MISS) {
        return false;
    }
    if (!in.returnOnAny) { // an any-hit query defines only the boolean (kdtree.cpp:338-341)
        in.record.t = h.t;
        for (int k = 0; k < 3; k++) {
            const float m = in.rayDir[k] * h.t;
            in.record.hitPoint[k] = in.rayOrigin[k] + m;
        }
        in.clippingDistance = h.t; // kdtree.cpp:343
    }
    return true;
}

} // namespace dodrt

#endif // DODRT_HPP

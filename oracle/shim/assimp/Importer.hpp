// TEST INFRASTRUCTURE -- minimal stand-in for Assimp::Importer (OBJ only).
// Surface used by the reference: ReadFile (/root/reference/src/shapes/mesh.cpp:11-14) and a
// default-constructible, copyable Importer (/root/reference/src/shapes/mesh.h:30).
// Loader semantics (documented in oracle/README.md): one mesh; faces in file order;
// polygons fan-triangulated; positions parsed with strtof; per-position smooth normals =
// normalised sum of the unit face normals of every face touching a bit-identical position.
#pragma once
#include "scene.h"
#include <memory>
namespace Assimp {
class Importer {
  public:
    Importer();
    ~Importer();
    Importer(const Importer &);
    Importer &operator=(const Importer &);
    const aiScene *ReadFile(const char *path, unsigned flags);

  private:
    struct Store;
    std::shared_ptr<Store> m_store;
};
} // namespace Assimp

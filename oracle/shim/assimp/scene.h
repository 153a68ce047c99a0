// TEST INFRASTRUCTURE -- minimal stand-in for assimp's aiScene.
// Surface used by the reference: /root/reference/src/shapes/mesh.cpp:17,27-29.
#pragma once
#include "mesh.h"
struct aiScene {
    unsigned mNumMeshes = 0;
    aiMesh **mMeshes = nullptr;
    bool HasMeshes() const { return mMeshes != nullptr && mNumMeshes > 0; }
};

// TEST INFRASTRUCTURE -- minimal stand-in for assimp's aiMesh / aiFace.
// Surface used by the reference: /root/reference/src/shapes/mesh.cpp:28-48.
#pragma once
#include "vector3.h"
struct aiFace {
    unsigned mNumIndices = 0;
    unsigned *mIndices = nullptr;
};
struct aiMesh {
    unsigned mNumVertices = 0;
    unsigned mNumFaces = 0;
    aiVector3D *mVertices = nullptr;
    aiVector3D *mNormals = nullptr;
    aiFace *mFaces = nullptr;
    bool HasFaces() const { return mFaces != nullptr && mNumFaces > 0; }
};

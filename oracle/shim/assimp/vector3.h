// TEST INFRASTRUCTURE -- minimal stand-in for assimp's aiVector3D (assimp is an un-vendored,
// unpinned dependency of the reference: /root/reference/CMakeLists.txt:46-63).
// Surface used by the reference: operator[] (/root/reference/src/utils/utils.h:55-58).
#pragma once
struct aiVector3D {
    float x, y, z;
    aiVector3D() : x(0), y(0), z(0) {}
    aiVector3D(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float operator[](unsigned i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float &operator[](unsigned i) { return i == 0 ? x : (i == 1 ? y : z); }
};

// TEST INFRASTRUCTURE -- the three post-process flags the reference passes
// (/root/reference/src/shapes/mesh.cpp:11-14).  The stand-in importer always triangulates,
// indexes by OBJ position and generates smooth normals, so the values only need to exist.
#pragma once
enum aiPostProcessSteps {
    aiProcess_JoinIdenticalVertices = 0x2,
    aiProcess_Triangulate = 0x8,
    aiProcess_GenSmoothNormals = 0x40,
};

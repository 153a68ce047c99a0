// TEST INFRASTRUCTURE -- not product code.
//
// Stand-in for the part of assimp the reference uses (Assimp::Importer::ReadFile with
// Triangulate | JoinIdenticalVertices | GenSmoothNormals, /root/reference/src/shapes/mesh.cpp:11-14).
// assimp is an unpinned FetchContent dependency (/root/reference/CMakeLists.txt:46-63) that is not
// in the reference tree, so parity is anchored on this loader: the oracle and the product's host
// loader consume the same positions / face order / normals law, and loader ulp differences
// against a true assimp build therefore cancel ("parity unpinned" at this boundary).
//
// Two inputs are understood, chosen by content:
//   * Wavefront OBJ text: `v x y z` and `f i j k ...` (i, i/t, i/t/n, i//n; negative = relative),
//     polygons fan-triangulated (0,k,k+1), everything else ignored.
//   * "DODM" binary mesh (written by tests for large synthetic meshes): char magic[4]="DODM",
//     u32 nVerts, u32 nTris, f32 positions[nVerts*3], u32 indices[nTris*3].
//
// Normals law (all fp32, contraction off): per triangle e1=v1-v0, e2=v2-v0,
// n=(e1.y*e2.z-e1.z*e2.y, e1.z*e2.x-e1.x*e2.z, e1.x*e2.y-e1.y*e2.x), len=sqrtf((n.x*n.x+n.y*n.y)+n.z*n.z),
// if len>0 n/=len (true division) and n is added, in face order then corner order, to the
// accumulator of each corner's position group (groups = bit-identical positions); finally each
// accumulator is divided by its own length if that is > 0.
#include "assimp/Importer.hpp"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace Assimp {

struct Importer::Store {
    std::vector<aiVector3D> vertices;
    std::vector<aiVector3D> normals;
    std::vector<unsigned> indices;
    std::vector<aiFace> faces;
    aiMesh mesh;
    aiMesh *meshPtr = nullptr;
    aiScene scene;
};

Importer::Importer() = default;
Importer::~Importer() = default;
Importer::Importer(const Importer &) = default;
Importer &Importer::operator=(const Importer &) = default;

namespace {

struct PosKey {
    uint32_t a, b, c;
    bool operator==(const PosKey &o) const { return a == o.a && b == o.b && c == o.c; }
};
struct PosKeyHash {
    size_t operator()(const PosKey &k) const
    {
        uint64_t h = 1469598103934665603ull;
        for (uint32_t w : {k.a, k.b, k.c}) {
            h ^= w;
            h *= 1099511628211ull;
        }
        return static_cast<size_t>(h);
    }
};

uint32_t bitsOf(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

void smoothNormals(const std::vector<aiVector3D> &pos, const std::vector<unsigned> &idx, std::vector<aiVector3D> &out)
{
    std::unordered_map<PosKey, unsigned, PosKeyHash> groupOf;
    std::vector<unsigned> group(pos.size());
    unsigned numGroups = 0;
    for (size_t i = 0; i < pos.size(); i++) {
        // +0.0 and -0.0 are the same position
        float px = pos[i].x + 0.0f, py = pos[i].y + 0.0f, pz = pos[i].z + 0.0f;
        PosKey key{bitsOf(px), bitsOf(py), bitsOf(pz)};
        auto it = groupOf.find(key);
        if (it == groupOf.end()) {
            it = groupOf.emplace(key, numGroups++).first;
        }
        group[i] = it->second;
    }
    std::vector<aiVector3D> acc(numGroups);
    for (size_t f = 0; f + 2 < idx.size(); f += 3) {
        const aiVector3D &v0 = pos[idx[f]], &v1 = pos[idx[f + 1]], &v2 = pos[idx[f + 2]];
        float e1x = v1.x - v0.x, e1y = v1.y - v0.y, e1z = v1.z - v0.z;
        float e2x = v2.x - v0.x, e2y = v2.y - v0.y, e2z = v2.z - v0.z;
        float nx = e1y * e2z - e1z * e2y;
        float ny = e1z * e2x - e1x * e2z;
        float nz = e1x * e2y - e1y * e2x;
        float len = sqrtf((nx * nx + ny * ny) + nz * nz);
        if (!(len > 0.0f)) {
            continue;
        }
        nx /= len;
        ny /= len;
        nz /= len;
        for (int c = 0; c < 3; c++) {
            aiVector3D &a = acc[group[idx[f + c]]];
            a.x += nx;
            a.y += ny;
            a.z += nz;
        }
    }
    for (aiVector3D &a : acc) {
        float len = sqrtf((a.x * a.x + a.y * a.y) + a.z * a.z);
        if (len > 0.0f) {
            a.x /= len;
            a.y /= len;
            a.z /= len;
        }
    }
    out.resize(pos.size());
    for (size_t i = 0; i < pos.size(); i++) {
        out[i] = acc[group[i]];
    }
}

bool parseBinary(const std::vector<char> &buf, std::vector<aiVector3D> &pos, std::vector<unsigned> &idx)
{
    if (buf.size() < 12 || std::memcmp(buf.data(), "DODM", 4) != 0) {
        return false;
    }
    uint32_t nv, nt;
    std::memcpy(&nv, buf.data() + 4, 4);
    std::memcpy(&nt, buf.data() + 8, 4);
    size_t need = 12 + size_t(nv) * 12 + size_t(nt) * 12;
    if (buf.size() < need) {
        return false;
    }
    pos.resize(nv);
    std::memcpy(static_cast<void *>(pos.data()), buf.data() + 12, size_t(nv) * 12);
    idx.resize(size_t(nt) * 3);
    std::memcpy(idx.data(), buf.data() + 12 + size_t(nv) * 12, size_t(nt) * 12);
    for (unsigned i : idx) {
        if (i >= nv) {
            return false;
        }
    }
    return true;
}

void parseObj(const std::vector<char> &buf, std::vector<aiVector3D> &pos, std::vector<unsigned> &idx)
{
    const char *p = buf.data();
    const char *end = p + buf.size();
    std::vector<long> corner;
    while (p < end) {
        const char *eol = static_cast<const char *>(std::memchr(p, '\n', end - p));
        if (!eol) {
            eol = end;
        }
        std::string line(p, eol);
        p = eol + 1;
        const char *s = line.c_str();
        while (*s == ' ' || *s == '\t') {
            s++;
        }
        if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
            char *q = nullptr;
            float x = strtof(s + 2, &q);
            float y = strtof(q, &q);
            float z = strtof(q, &q);
            pos.emplace_back(x, y, z);
        } else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
            corner.clear();
            const char *c = s + 2;
            while (*c) {
                while (*c == ' ' || *c == '\t' || *c == '\r') {
                    c++;
                }
                if (!*c) {
                    break;
                }
                char *q = nullptr;
                long v = strtol(c, &q, 10);
                if (q == c) {
                    break;
                }
                corner.push_back(v);
                c = q;
                while (*c && *c != ' ' && *c != '\t' && *c != '\r') {
                    c++; // skip /vt/vn
                }
            }
            for (size_t k = 1; k + 1 < corner.size(); k++) {
                long tri[3] = {corner[0], corner[k], corner[k + 1]};
                for (long v : tri) {
                    long zero = v > 0 ? v - 1 : static_cast<long>(pos.size()) + v;
                    idx.push_back(static_cast<unsigned>(zero));
                }
            }
        }
    }
}

} // namespace

const aiScene *Importer::ReadFile(const char *path, unsigned)
{
    FILE *f = std::fopen(path, "rb");
    if (!f) {
        return nullptr;
    }
    std::vector<char> buf;
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.resize(sz > 0 ? sz : 0);
    size_t got = sz > 0 ? std::fread(buf.data(), 1, sz, f) : 0;
    std::fclose(f);
    buf.resize(got);

    auto store = std::make_shared<Store>();
    if (!parseBinary(buf, store->vertices, store->indices)) {
        store->vertices.clear();
        store->indices.clear();
        parseObj(buf, store->vertices, store->indices);
    }
    for (unsigned i : store->indices) {
        if (i >= store->vertices.size()) {
            return nullptr;
        }
    }
    if (store->indices.empty()) {
        return nullptr;
    }
    smoothNormals(store->vertices, store->indices, store->normals);

    size_t numFaces = store->indices.size() / 3;
    store->faces.resize(numFaces);
    for (size_t i = 0; i < numFaces; i++) {
        store->faces[i].mNumIndices = 3;
        store->faces[i].mIndices = &store->indices[i * 3];
    }
    store->mesh.mNumVertices = static_cast<unsigned>(store->vertices.size());
    store->mesh.mNumFaces = static_cast<unsigned>(numFaces);
    store->mesh.mVertices = store->vertices.data();
    store->mesh.mNormals = store->normals.data();
    store->mesh.mFaces = store->faces.data();
    store->meshPtr = &store->mesh;
    store->scene.mNumMeshes = 1;
    store->scene.mMeshes = &store->meshPtr;
    m_store = store;
    return &m_store->scene;
}

} // namespace Assimp

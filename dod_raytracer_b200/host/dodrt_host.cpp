// dodrt_host.cpp -- host side above the ray-query path (see include/dodrt_host.h).
//
// Written from scratch; reproduces the reference's host-side RESULTS bit for bit so that the GPU
// path and the reference CPU path see the same scene: same lanes, same kd-tree (including the
// builder's quirks), same analytic shape arrays, same raster tables.
// Build: g++ -O2 -ffp-contract=off (no FMA contraction anywhere, like the reference's default build).
#include "../../include/dodrt_host.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <new>
#include <string>
#include <mutex>
#include <exception>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

namespace {

// Worker threads whose exceptions do not end the process: the first one is kept and re-thrown by join() on the calling
// thread (from where the C ABI's catch turns it into a status code), and the destructor joins whatever still runs, so
// an exception on the calling thread never meets a joinable std::thread.
class ThreadGroup {
public:
    template <typename F> void run(F body)
    {
        threads_.emplace_back([this, body]() mutable {
            try {
                body();
            } catch (...) {
                std::lock_guard<std::mutex> lock(mutex_);
                if (!error_) error_ = std::current_exception();
            }
        });
    }
    void join()
    {
        joinAll();
        if (error_) {
            std::exception_ptr e = error_;
            error_ = nullptr;
            std::rethrow_exception(e);
        }
    }
    ~ThreadGroup() { joinAll(); }

private:
    void joinAll()
    {
        for (std::thread &t : threads_) {
            if (t.joinable()) t.join();
        }
        threads_.clear();
    }
    std::vector<std::thread> threads_;
    std::mutex mutex_;
    std::exception_ptr error_;
};


thread_local std::string g_error;

int fail(const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return DODRT_E_INVALID;
}

constexpr uint32_t kLane = 8;
constexpr float kInf = std::numeric_limits<float>::infinity();

struct Vec3 {
    float x, y, z;
    float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

// glm::dot order: (x*x' + y*y') + z*z'
inline float dot(const Vec3 &a, const Vec3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// glm::min / glm::max: (y < x) ? y : x  /  (x < y) ? y : x   (box.cpp:12-16, utils.h:60-90)
inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }

struct Box {
    Vec3 lo, hi;
    void unite(const Box &b) // AxisAlignedBoundingBox::Union, box.cpp:8-19
    {
        for (int i = 0; i < 3; i++) lo[i] = gmin(lo[i], b.lo[i]);
        for (int i = 0; i < 3; i++) hi[i] = gmax(hi[i], b.hi[i]);
    }
    float surfaceArea() const // box.cpp:27-31
    {
        Vec3 v{hi.x - lo.x, hi.y - lo.y, hi.z - lo.z};
        return ((2 * v.x * v.y) + (2 * v.x * v.z) + (2 * v.y * v.z));
    }
    unsigned maximumExtent() const // box.cpp:21-25 + getMaxElementIndex utils.h:108-124 (seeded with FLT_MIN)
    {
        Vec3 v{hi.x - lo.x, hi.y - lo.y, hi.z - lo.z};
        float maxElem = std::numeric_limits<float>::min();
        unsigned maxIndex = std::numeric_limits<unsigned>::max();
        for (int i = 0; i < 3; i++) {
            if (v[i] > maxElem) {
                maxElem = v[i];
                maxIndex = i;
            }
        }
        return maxIndex;
    }
};

struct TriLane { // triangle.h:33-44
    float v[9][kLane];
};
struct NormalLane { // AN, BN, CN per slot (flat export for tools / tests)
    float n[kLane][9];
};
struct AttrLane { // Triangle::Attributes, triangle.h:45-51, byte for byte (320 B)
    uint32_t meshAttrIdx[kLane];
    float AN[kLane][3];
    float BN[kLane][3];
    float CN[kLane][3];
};

// glibc's rand()/srand() (TYPE_3 additive feedback generator), restated so that the reference's
// srand(seed) scene (main.cpp:26-50) is reproducible without touching libc's global state.
class GlibcRand {
  public:
    explicit GlibcRand(uint32_t seed)
    {
        int32_t r[34];
        r[0] = seed == 0 ? 1 : (int32_t)seed;
        for (int i = 1; i < 31; i++) {
            int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
            int64_t word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            r[i] = (int32_t)word;
        }
        for (int i = 31; i < 34; i++) r[i] = r[i - 31];
        for (int i = 0; i < 34; i++) state_[i] = (uint32_t)r[i];
        pos_ = 34;
        for (int i = 34; i < 344; i++) next();
    }
    int operator()() { return (int)(next() >> 1); }

  private:
    uint32_t next()
    {
        uint32_t v = state_[(pos_ - 31) % 34] + state_[(pos_ - 3) % 34];
        state_[pos_ % 34] = v;
        pos_++;
        return v;
    }
    uint32_t state_[34];
    uint64_t pos_;
};

constexpr float kRandMax = 2147483647; // (float)rand() / RAND_MAX: int promoted to float

} // namespace

struct dodrt_host_scene {
    dodrt_host_config cfg;
    // triangles
    uint32_t numTriangles = 0;
    std::vector<TriLane> lanes;       // before build: creation order; after build: re-ordered
    std::vector<NormalLane> normals;  // same order as lanes
    std::vector<AttrLane> attrs;      // same order as lanes; the reference's own attribute layout
    std::vector<float> meshColors;    // Mesh::m_meshAttributes, mesh.cpp:23 (3 floats per mesh)
    // kd-tree
    std::vector<uint64_t> nodes;
    std::vector<uint32_t> primNums;
    uint32_t maxDepth = 0;
    uint32_t numOrigLanes = 0;
    Box bounds{{0, 0, 0}, {0, 0, 0}};
    float boundsOut[6] = {0, 0, 0, 0, 0, 0};
    bool built = false;
    double buildSeconds = 0.0, reorderSeconds = 0.0;
    bool creationOrder = false; // lanes / normals / attrs were NOT re-ordered (DODRT_HOST_BUILD_KEEP_CREATION_ORDER)
    // analytic shapes
    std::vector<float> sphereLanes, sphereColors;
    uint32_t numSpheres = 0;
    std::vector<float> planeLanes, planeColors;
    uint32_t numPlanes = 0;
    std::vector<dodrt_cylinder> cylinders;
    std::vector<float> boxLanes;
    uint32_t numBoxes = 0;
};

namespace {

// ---- Triangle::create, triangle.cpp:262-292 ------------------------------------------------------------
void pushTriangle(dodrt_host_scene *s, const Vec3 p[3], const Vec3 n[3])
{
    const uint32_t slot = s->numTriangles % kLane;
    if (slot == 0) {
        s->lanes.emplace_back();
        s->normals.emplace_back();
        s->attrs.emplace_back();
        std::memset(&s->lanes.back(), 0, sizeof(TriLane));
        std::memset(&s->normals.back(), 0, sizeof(NormalLane));
        std::memset(&s->attrs.back(), 0, sizeof(AttrLane)); // emptyTriangleAttributes = {0}, triangle.cpp:265
    }
    TriLane &lane = s->lanes.back();
    NormalLane &nl = s->normals.back();
    AttrLane &al = s->attrs.back();
    al.meshAttrIdx[slot] = (uint32_t)(s->meshColors.size() / 3) - 1u; // Mesh::m_meshAttributes.size() - 1, triangle.cpp:286
    for (int k = 0; k < 3; k++) {
        al.AN[slot][k] = n[0][k];
        al.BN[slot][k] = n[1][k];
        al.CN[slot][k] = n[2][k];
    }
    for (int c = 0; c < 3; c++) {
        lane.v[c * 3 + 0][slot] = p[c].x;
        lane.v[c * 3 + 1][slot] = p[c].y;
        lane.v[c * 3 + 2][slot] = p[c].z;
        nl.n[slot][c * 3 + 0] = n[c].x;
        nl.n[slot][c * 3 + 1] = n[c].y;
        nl.n[slot][c * 3 + 2] = n[c].z;
    }
    s->numTriangles++;
}

// ---- mesh loading ---------------------------------------------------------------------------------------
struct PosKey {
    uint32_t a, b, c;
    bool operator==(const PosKey &o) const { return a == o.a && b == o.b && c == o.c; }
};
struct PosKeyHash {
    size_t operator()(const PosKey &k) const
    {
        uint64_t h = (uint64_t)k.a * 0x9E3779B97F4A7C15ull;
        h ^= (uint64_t)k.b * 0xC2B2AE3D27D4EB4Full + (h >> 29);
        h ^= (uint64_t)k.c * 0x165667B19E3779F9ull + (h << 7);
        return (size_t)h;
    }
};
inline uint32_t bitsOf(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

// Loader law for normals (DESIGN.md "mesh loader"): unit face normals (cross(v1-v0, v2-v0) / its length,
// skipped when the length is not > 0) are summed, face by face and corner by corner, per group of
// bit-identical positions (+0 == -0), and every sum is divided by its own length when that is > 0.
void smoothNormals(const std::vector<Vec3> &pos, const uint32_t *idx, uint32_t numTris, std::vector<Vec3> &out)
{
    std::unordered_map<PosKey, uint32_t, PosKeyHash> groupOf;
    groupOf.reserve(pos.size() * 2);
    std::vector<uint32_t> group(pos.size());
    uint32_t numGroups = 0;
    for (size_t i = 0; i < pos.size(); i++) {
        PosKey key{bitsOf(pos[i].x + 0.0f), bitsOf(pos[i].y + 0.0f), bitsOf(pos[i].z + 0.0f)};
        auto it = groupOf.find(key);
        if (it == groupOf.end()) it = groupOf.emplace(key, numGroups++).first;
        group[i] = it->second;
    }
    std::vector<Vec3> acc(numGroups, Vec3{0, 0, 0});
    for (uint32_t f = 0; f < numTris; f++) {
        const Vec3 &v0 = pos[idx[f * 3]], &v1 = pos[idx[f * 3 + 1]], &v2 = pos[idx[f * 3 + 2]];
        Vec3 e1{v1.x - v0.x, v1.y - v0.y, v1.z - v0.z}, e2{v2.x - v0.x, v2.y - v0.y, v2.z - v0.z};
        Vec3 n{e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x};
        float len = sqrtf(dot(n, n));
        if (!(len > 0.0f)) continue;
        n.x /= len;
        n.y /= len;
        n.z /= len;
        for (int c = 0; c < 3; c++) {
            Vec3 &a = acc[group[idx[f * 3 + c]]];
            a.x += n.x;
            a.y += n.y;
            a.z += n.z;
        }
    }
    for (Vec3 &a : acc) {
        float len = sqrtf(dot(a, a));
        if (len > 0.0f) {
            a.x /= len;
            a.y /= len;
            a.z /= len;
        }
    }
    out.resize(pos.size());
    for (size_t i = 0; i < pos.size(); i++) out[i] = acc[group[i]];
}

int addMesh(dodrt_host_scene *s, std::vector<Vec3> &pos, const uint32_t *idx, uint32_t numTris, const float transform[4])
{
    if (s->built) return fail("scene already built: the reference can only add meshes before KDTree::buildTree()");
    for (uint32_t i = 0; i < numTris * 3; i++) {
        if (idx[i] >= pos.size()) return fail("mesh index %u out of range (%zu vertices)", idx[i], pos.size());
    }
    if (transform) {
        const float sc = transform[0];
        for (Vec3 &p : pos) {
            p.x = p.x * sc + transform[1];
            p.y = p.y * sc + transform[2];
            p.z = p.z * sc + transform[3];
        }
    }
    std::vector<Vec3> nrm;
    const auto t0 = std::chrono::steady_clock::now();
    smoothNormals(pos, idx, numTris, nrm);
    const auto t1 = std::chrono::steady_clock::now();
    const float meshColor[3] = {(float)0.1, (float)0.8, (float)0.3}; // mesh.cpp:23
    s->meshColors.insert(s->meshColors.end(), meshColor, meshColor + 3);
    // grow geometrically: an exact reserve per mesh re-copies every earlier mesh (quadratic for many instances)
    const size_t need = s->lanes.size() + numTris / kLane + 1;
    if (need > s->lanes.capacity()) {
        const size_t cap = std::max(need, s->lanes.capacity() * 2);
        s->lanes.reserve(cap);
        s->normals.reserve(cap);
        s->attrs.reserve(cap);
    }
    for (uint32_t f = 0; f < numTris; f++) {
        Vec3 p[3] = {pos[idx[f * 3]], pos[idx[f * 3 + 1]], pos[idx[f * 3 + 2]]};
        Vec3 n[3] = {nrm[idx[f * 3]], nrm[idx[f * 3 + 1]], nrm[idx[f * 3 + 2]]};
        pushTriangle(s, p, n);
    }
    if (std::getenv("DODRT_HOST_VERBOSE")) {
        std::fprintf(stderr, "dodrt_host: add_mesh %u triangles: normals %.3f s, lanes %.3f s\n", numTris,
                     std::chrono::duration<double>(t1 - t0).count(),
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
    }
    return DODRT_OK;
}

bool readFile(const char *path, std::vector<char> &buf)
{
    FILE *f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.resize(sz > 0 ? (size_t)sz : 0);
    size_t got = sz > 0 ? std::fread(buf.data(), 1, (size_t)sz, f) : 0;
    std::fclose(f);
    buf.resize(got);
    return true;
}

unsigned hostThreads();

// One chunk of an OBJ text (whole lines).  Positive `f` indices are absolute; negative ones count back from the number
// of `v` lines seen so far IN THE FILE, which a chunk does not know -- they are stored as (negative value, number of
// v lines seen so far in this chunk) and resolved once the chunks' vertex counts are known.
struct ObjChunk {
    std::vector<Vec3> pos;
    static constexpr uint32_t kAbsolute = 0xFFFFFFFFu;
    std::vector<long> corners;       // 3 per triangle: absolute vertex index, or (<= 0) relative, see `localSeen`
    std::vector<uint32_t> localSeen; // per corner: kAbsolute, or the v lines of this chunk before the face
};

void parseObjChunk(const char *p, const char *end, ObjChunk &out)
{
    std::vector<long> corners;
    std::vector<uint32_t> seen;
    std::string line;
    while (p < end) {
        const char *eol = static_cast<const char *>(std::memchr(p, '\n', (size_t)(end - p)));
        if (!eol) eol = end;
        line.assign(p, eol);
        p = eol + 1;
        const char *c = line.c_str();
        while (*c == ' ' || *c == '\t') c++;
        if (c[0] == 'v' && (c[1] == ' ' || c[1] == '\t')) {
            char *q = nullptr;
            Vec3 v;
            v.x = strtof(c + 2, &q);
            v.y = strtof(q, &q);
            v.z = strtof(q, &q);
            out.pos.push_back(v);
        } else if (c[0] == 'f' && (c[1] == ' ' || c[1] == '\t')) {
            corners.clear();
            c += 2;
            for (;;) {
                while (*c == ' ' || *c == '\t' || *c == '\r') c++;
                if (!*c) break;
                char *q = nullptr;
                long v = strtol(c, &q, 10);
                if (q == c) break;
                corners.push_back(v);
                c = q;
                while (*c && *c != ' ' && *c != '\t' && *c != '\r') c++; // skip /vt/vn
            }
            for (size_t k = 1; k + 1 < corners.size(); k++) { // fan triangulation
                const long tri[3] = {corners[0], corners[k], corners[k + 1]};
                for (long v : tri) { // v > 0: 1-based absolute; v <= 0: relative to the vertices seen so far
                    out.corners.push_back(v > 0 ? v - 1 : v);
                    out.localSeen.push_back(v > 0 ? ObjChunk::kAbsolute : (uint32_t)out.pos.size());
                }
            }
        }
    }
}

// `v x y z` (strtof, the loader law of DESIGN.md) and `f a b c ...` lines; everything else is ignored.  The text is
// cut at line ends into one chunk per host thread, parsed concurrently and concatenated in file order -- the same
// values as a sequential pass.
void parseObj(const std::vector<char> &buf, std::vector<Vec3> &pos, std::vector<uint32_t> &idx)
{
    const char *base = buf.data(), *end = base + buf.size();
    size_t numChunks = std::max<size_t>(1, std::min<size_t>(hostThreads(), buf.size() / (1u << 20)));
    std::vector<const char *> cut(numChunks + 1, end);
    cut[0] = base;
    for (size_t k = 1; k < numChunks; k++) {
        const char *p = base + buf.size() * k / numChunks;
        if (p < cut[k - 1]) p = cut[k - 1];
        const char *eol = static_cast<const char *>(std::memchr(p, '\n', (size_t)(end - p)));
        cut[k] = eol ? eol + 1 : end;
    }
    std::vector<ObjChunk> chunks(numChunks);
    {
        ThreadGroup pool; // joins on every path; a worker's exception is re-thrown here, on the caller's thread
        for (size_t k = 1; k < numChunks; k++) pool.run([&, k] { parseObjChunk(cut[k], cut[k + 1], chunks[k]); });
        parseObjChunk(cut[0], cut[1], chunks[0]);
        pool.join();
    }
    size_t seenBefore = 0;
    for (const ObjChunk &c : chunks) {
        for (size_t i = 0; i < c.corners.size(); i++) {
            const long v = c.corners[i];
            idx.push_back((uint32_t)(c.localSeen[i] == ObjChunk::kAbsolute ? v : (long)(seenBefore + c.localSeen[i]) + v));
        }
        pos.insert(pos.end(), c.pos.begin(), c.pos.end());
        seenBefore += c.pos.size();
    }
}

// ---- KDTree::buildTree, kdtree.cpp:66-260 ---------------------------------------------------------------
struct Edge { // KDTree::AxisOffsetInEdge, kdtree.cpp:12-29
    float offset;
    uint32_t lane;
    bool isEnd;
};

// The recursion of kdtree.cpp:95-250 is a pure function of (depth, badRefines, node bounds, lane list): it emits the
// subtree's nodes in DFS pre-order and appends its leaves' lane numbers.  Only two things tie a subtree to its
// position in the whole tree -- the absolute right-child indices (Node::initInteriorNode) and the absolute
// m_primNums offsets of its leaves (Node::initLeafNode) -- so subtrees can be built CONCURRENTLY into private
// buffers with buffer-relative indices and spliced in afterwards, re-based by the splice position.  The result is
// the reference's tree bit for bit (tests/test_host_vs_ref.py, tests/test_host_parallel_build.py), because every
// decision inside a subtree (edge sort with libstdc++'s introsort on the same input sequence, the SAH sweep with its
// unsigned-cost quirk, the partition order) is executed by the same code on the same data.
// Also: the three edge lists are built and sorted LAZILY, one axis at a time, in the order the sweep visits them
// (kdtree.cpp:140-200 leaves the loop as soon as an axis found a cost below the leaf cost); the reference sorts all
// three up front (kdtree.cpp:113-137), which does not influence any value the sweep reads.
class TreeBuilder {
  public:
    TreeBuilder(dodrt_host_scene *s, unsigned threads) : s_(s), maxWorkers_(threads > 1 ? threads - 1 : 0) {}

    void run()
    {
        const uint32_t numLanes = (uint32_t)s_->lanes.size();
        // kdtree.cpp:72 -- float arithmetic throughout: std::log2(float), std::round(float)
        s_->maxDepth = (uint32_t)std::round(std::log2(8.0f + (1.3f * numLanes)));
        Box world{{kInf, kInf, kInf}, {-kInf, -kInf, -kInf}};
        std::vector<uint32_t> laneNumbers;
        laneBoxes_.reserve(numLanes);
        laneNumbers.reserve(numLanes);
        for (uint32_t i = 0; i < s_->numTriangles; i += kLane) { // kdtree.cpp:84-90
            const uint32_t len = std::min(kLane, s_->numTriangles - i);
            laneBoxes_.push_back(laneBox(i / kLane, len));
            world.unite(laneBoxes_.back());
            laneNumbers.push_back(i / kLane);
        }
        s_->bounds = world;
        SubTree root;
        build(root, s_->maxDepth, 0, world, laneNumbers);
        s_->nodes.swap(root.nodes);
        s_->primNums.swap(root.prims);
        for (int a = 0; a < 3; a++) std::vector<Edge>().swap(scratch()[a]); // this thread's copy; workers' died with them
    }

  private:
    struct SubTree { // nodes with indices relative to this buffer
        std::vector<uint64_t> nodes;
        std::vector<uint32_t> prims;
    };
    static constexpr size_t kTaskMinLanes = 4096; // smaller subtrees are not worth a thread

    Box laneBox(uint32_t lane, uint32_t len) const // Triangle::getBoundingBox, triangle.cpp:294-339
    {
        Box box{{kInf, kInf, kInf}, {-kInf, -kInf, -kInf}};
        const TriLane &l = s_->lanes[lane];
        for (uint32_t j = 0; j < len; j++) {
            Box tri{{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()},
                    {std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(),
                     std::numeric_limits<float>::lowest()}};
            for (int c = 0; c < 3; c++) {
                for (int a = 0; a < 3; a++) {
                    tri.lo[a] = gmin(tri.lo[a], l.v[c * 3 + a][j]);
                    tri.hi[a] = gmax(tri.hi[a], l.v[c * 3 + a][j]);
                }
            }
            box.unite(tri);
        }
        return box;
    }

    static void makeLeaf(SubTree &out, const std::vector<uint32_t> &laneNums) // Node::initLeafNode, kdtree.cpp:42-56
    {
        const uint32_t w0 = 3u | ((uint32_t)laneNums.size() << 2);
        const uint32_t w1 = (uint32_t)out.prims.size();
        out.nodes.push_back((uint64_t)w0 | ((uint64_t)w1 << 32));
        out.prims.insert(out.prims.end(), laneNums.begin(), laneNums.end());
    }

    // append `src` to `dst`, re-basing right-child indices and m_primNums offsets by the splice position
    static void splice(SubTree &dst, const SubTree &src)
    {
        const uint64_t nodeBase = dst.nodes.size(), primBase = dst.prims.size();
        dst.nodes.reserve(dst.nodes.size() + src.nodes.size());
        for (uint64_t n : src.nodes) {
            if ((n & 3u) == 3u) {
                n += primBase << 32; // leaf: word1 = first entry in m_primNums
            } else {
                n += nodeBase << 2; // interior: word0[31:2] = right child index
            }
            dst.nodes.push_back(n);
        }
        dst.prims.insert(dst.prims.end(), src.prims.begin(), src.prims.end());
    }

    static std::vector<Edge> *scratch()
    {
        thread_local std::vector<Edge> edges[3];
        return edges;
    }

    bool acquireWorker()
    {
        unsigned cur = workers_.load(std::memory_order_relaxed);
        while (cur < maxWorkers_) {
            if (workers_.compare_exchange_weak(cur, cur + 1, std::memory_order_acq_rel)) return true;
        }
        return false;
    }

    void build(SubTree &out, unsigned depth, unsigned badRefines, const Box &nodeBounds, std::vector<uint32_t> &laneNums)
    {
        const dodrt_host_config &cfg = s_->cfg;
        if (depth == 0 || laneNums.size() <= cfg.max_prims) { // kdtree.cpp:106
            makeLeaf(out, laneNums);
            return;
        }
        // per-thread scratch: the edge lists are dead before the recursion, so one set per thread serves every node
        std::vector<Edge> *edges = scratch();
        for (int a = 0; a < 3; a++) edges[a].clear();
        auto sortedAxis = [&](unsigned a) -> const std::vector<Edge> & {
            std::vector<Edge> &e = edges[a];
            if (e.empty()) {
                e.reserve(laneNums.size() * 2);
                for (uint32_t lane : laneNums) { // kdtree.cpp:118-127: Start then End, per lane
                    const Box &b = laneBoxes_[lane];
                    e.push_back(Edge{b.lo[a], lane, false});
                    e.push_back(Edge{b.hi[a], lane, true});
                }
                // kdtree.cpp:128-137: offset-only comparator (ties: libstdc++ introsort order)
                std::sort(e.begin(), e.end(), [](const Edge &x, const Edge &y) { return x.offset < y.offset; });
            }
            return e;
        };

        // SAH sweep, kdtree.cpp:140-200.  bestSplitCost is an UNSIGNED that receives float costs
        // (truncation on every assignment) -- a reference quirk the tree shape depends on.
        unsigned bestSplitIdx = UINT32_MAX;
        unsigned bestSplitCost = UINT32_MAX;
        const float originalSplitCost = cfg.intersect_cost * laneNums.size();
        unsigned splitAxis = 0;
        const unsigned maxAxis = nodeBounds.maximumExtent();
        const float invTotalSurfaceArea = 1.0f / nodeBounds.surfaceArea();
        for (unsigned i = 0; i < 3; i++) {
            const unsigned axis = (maxAxis + i) % 3;
            unsigned numLeft = 0;
            unsigned numRight = (unsigned)laneNums.size();
            const std::vector<Edge> &ax = sortedAxis(axis);
            for (unsigned j = 0; j < ax.size(); j++) {
                const Edge &e = ax[j];
                if (e.isEnd) numRight--;
                if (e.offset >= nodeBounds.lo[axis] && e.offset <= nodeBounds.hi[axis]) {
                    Box left = nodeBounds, right = nodeBounds;
                    left.hi[axis] = e.offset;
                    right.lo[axis] = e.offset;
                    const float pLeft = left.surfaceArea() * invTotalSurfaceArea;
                    const float pRight = right.surfaceArea() * invTotalSurfaceArea;
                    const float emptyBonus = (!numRight || !numRight) ? cfg.empty_bonus : 0.0f; // kdtree.cpp:175 (sic)
                    const float cost = cfg.traversal_cost + cfg.intersect_cost * (1 - emptyBonus) * (pLeft * numLeft + pRight * numRight);
                    if (cost < bestSplitCost) {
                        bestSplitCost = cost; // float -> unsigned truncation, kdtree.cpp:181
                        splitAxis = axis;
                        bestSplitIdx = j;
                    }
                }
                if (!e.isEnd) numLeft++;
            }
            if (bestSplitCost < originalSplitCost) break; // kdtree.cpp:196
        }
        if (bestSplitCost > originalSplitCost) badRefines++; // kdtree.cpp:202

        const size_t nodeIdx = out.nodes.size();
        if (bestSplitIdx == UINT32_MAX || badRefines == 3 ||
            (bestSplitCost > 4 * originalSplitCost && laneNums.size() < 16)) { // kdtree.cpp:208-214
            makeLeaf(out, laneNums);
            return;
        }
        out.nodes.push_back(0);

        const float splitOffset = edges[splitAxis][bestSplitIdx].offset;
        Box leftBounds = nodeBounds, rightBounds = nodeBounds;
        leftBounds.hi[splitAxis] = splitOffset;
        rightBounds.lo[splitAxis] = splitOffset;
        std::vector<uint32_t> leftLanes, rightLanes; // kdtree.cpp:224-243
        {
            const std::vector<Edge> &ax = edges[splitAxis];
            leftLanes.reserve(laneNums.size());
            rightLanes.reserve(laneNums.size());
            for (unsigned i = 0; i < bestSplitIdx; i++) {
                if (!ax[i].isEnd) leftLanes.push_back(ax[i].lane);
            }
            for (size_t i = (size_t)bestSplitIdx + 1; i < ax.size(); i++) {
                if (ax[i].isEnd) rightLanes.push_back(ax[i].lane);
            }
        }
        const bool fork = laneNums.size() >= kTaskMinLanes && acquireWorker();
        std::vector<uint32_t>().swap(laneNums);                          // the caller's copy is dead too

        uint32_t w1;
        std::memcpy(&w1, &splitOffset, 4);
        if (fork) {
            // left subtree on another thread, right subtree here, both into private buffers; splice in DFS order
            SubTree leftSub, rightSub;
            {
                ThreadGroup worker; // joins on every path; the worker's exception (bad_alloc) is re-thrown by join()
                struct Release {
                    std::atomic<unsigned> &n;
                    ~Release() { n.fetch_sub(1, std::memory_order_acq_rel); }
                };
                worker.run([&] {
                    Release release{workers_};
                    build(leftSub, depth - 1, badRefines, leftBounds, leftLanes);
                });
                build(rightSub, depth - 1, badRefines, rightBounds, rightLanes);
                worker.join();
            }
            splice(out, leftSub);
            const uint32_t w0 = splitAxis | ((uint32_t)out.nodes.size() << 2);
            out.nodes[nodeIdx] = (uint64_t)w0 | ((uint64_t)w1 << 32);
            splice(out, rightSub);
            return;
        }
        build(out, depth - 1, badRefines, leftBounds, leftLanes);
        // Node::initInteriorNode, kdtree.cpp:58-64: flags = axis, right child = next node index
        const uint32_t w0 = splitAxis | ((uint32_t)out.nodes.size() << 2);
        out.nodes[nodeIdx] = (uint64_t)w0 | ((uint64_t)w1 << 32);
        build(out, depth - 1, badRefines, rightBounds, rightLanes);
    }

    dodrt_host_scene *s_;
    std::vector<Box> laneBoxes_;
    const unsigned maxWorkers_;
    std::atomic<unsigned> workers_{0};
};

// Triangle::reorderLanesByIndices (triangle.cpp:349-367): out[i] = in[primNums[i]], split over `threads` threads
template <typename T>
void gatherLanes(const std::vector<T> &in, const std::vector<uint32_t> &primNums, std::vector<T> &out, unsigned threads)
{
    static_assert(std::is_trivially_copyable<T>::value, "lane records are plain data");
    out.resize(primNums.size());
    const size_t n = primNums.size();
    threads = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, n / 65536 + 1));
    ThreadGroup pool;
    for (unsigned t = 0; t < threads; t++) {
        const size_t lo = n * t / threads, hi = n * (t + 1) / threads;
        auto work = [&in, &primNums, &out, lo, hi] {
            for (size_t i = lo; i < hi; i++) out[i] = in[primNums[i]];
        };
        if (t + 1 == threads) {
            work();
        } else {
            pool.run(work);
        }
    }
    pool.join();
}

unsigned hostThreads()
{
    if (const char *e = std::getenv("DODRT_HOST_THREADS")) {
        const int v = std::atoi(e);
        if (v >= 1) return (unsigned)v;
    }
    const unsigned hw = std::thread::hardware_concurrency();
    return hw ? hw : 1;
}

void appendLane(std::vector<float> &lanes, uint32_t index, uint32_t floatsPerLane, const float *values, uint32_t numValues)
{
    const uint32_t slot = index % kLane;
    if (slot == 0) lanes.resize(lanes.size() + floatsPerLane, 0.0f);
    float *lane = lanes.data() + (size_t)(index / kLane) * floatsPerLane;
    for (uint32_t k = 0; k < numValues; k++) lane[k * kLane + slot] = values[k];
}

} // namespace

// The ABI never throws: every int-returning entry point is a function-try-block ending in DODRT_HOST_CATCH
// (std::bad_alloc from the vectors, std::system_error from std::thread in the task-parallel builder, ...).
static int failException()
{
    try {
        throw;
    } catch (const std::bad_alloc &) {
        fail("out of host memory");
        return DODRT_E_NOMEM;
    } catch (const std::exception &e) {
        return fail("unexpected C++ exception: %s", e.what());
    } catch (...) {
        return fail("unexpected C++ exception");
    }
}
#define DODRT_HOST_CATCH                                                                                       \
    catch (...) { return failException(); }

extern "C" {

const char *dodrt_host_last_error(void) { return g_error.c_str(); }

void dodrt_host_config_defaults(dodrt_host_config *cfg)
{
    cfg->height = 1080;
    cfg->width = 1920;
    cfg->epsilon = 0.0001f;
    cfg->frustrum_max = 1000.0f;
    cfg->intersect_cost = 80;
    cfg->traversal_cost = 80;
    cfg->empty_bonus = 0.0f;
    cfg->max_prims = 8;
}

int dodrt_host_config_load(const char *path, dodrt_host_config *cfg)
try {
    if (!path || !cfg) return fail("NULL argument");
    dodrt_host_config_defaults(cfg);
    std::ifstream in(path);
    if (!in.is_open()) return fail("cannot open %s", path); // Config::Load returns false, config.h:19-23
    std::unordered_map<std::string, std::string> kv;
    for (std::string line; std::getline(in, line);) {
        const size_t colon = line.find(':');
        std::string key = line.substr(0, colon), value = colon == std::string::npos ? "" : line.substr(colon + 1);
        auto strip = [](std::string &t) { t.erase(std::remove_if(t.begin(), t.end(), [](unsigned char c) { return std::isspace(c); }), t.end()); };
        strip(key);
        strip(value);
        kv.insert({key, value}); // first occurrence wins, like unordered_map::insert in config_loader.h:52
    }
    try {
        auto u = [&](const char *k, uint32_t &dst) { auto it = kv.find(k); if (it != kv.end()) dst = (uint32_t)std::stoi(it->second); };
        auto f = [&](const char *k, float &dst) { auto it = kv.find(k); if (it != kv.end()) dst = std::stof(it->second); };
        u("Height", cfg->height);
        u("Width", cfg->width);
        f("Epsilon", cfg->epsilon);
        f("FrustrumMax", cfg->frustrum_max);
        u("IntersectCost", cfg->intersect_cost);
        u("TraversalCost", cfg->traversal_cost);
        f("EmptyBonus", cfg->empty_bonus);
        u("MaxPrims", cfg->max_prims);
    } catch (const std::exception &e) {
        return fail("bad value in %s: %s", path, e.what());
    }
    return DODRT_OK;
}
DODRT_HOST_CATCH

int dodrt_host_scene_create(const dodrt_host_config *cfg, dodrt_host_scene **scene)
try {
    if (!scene) return fail("scene is NULL");
    dodrt_host_scene *s = new (std::nothrow) dodrt_host_scene();
    if (!s) return DODRT_E_NOMEM;
    if (cfg) s->cfg = *cfg;
    else dodrt_host_config_defaults(&s->cfg);
    *scene = s;
    return DODRT_OK;
}
DODRT_HOST_CATCH

void dodrt_host_scene_destroy(dodrt_host_scene *scene) { delete scene; }

int dodrt_host_add_mesh(dodrt_host_scene *s, const float *positions, uint32_t numVertices, const uint32_t *indices,
                        uint32_t numTriangles, const float transform[4])
try {
    if (!s || (!positions && numVertices) || (!indices && numTriangles)) return fail("NULL argument");
    std::vector<Vec3> pos(numVertices);
    for (uint32_t i = 0; i < numVertices; i++) pos[i] = Vec3{positions[i * 3], positions[i * 3 + 1], positions[i * 3 + 2]};
    return addMesh(s, pos, indices, numTriangles, transform);
}
DODRT_HOST_CATCH

int dodrt_host_add_mesh_file(dodrt_host_scene *s, const char *path, const float transform[4])
try {
    if (!s || !path) return fail("NULL argument");
    std::vector<char> buf;
    if (!readFile(path, buf)) return fail("cannot read %s", path); // mesh.cpp:17-21 prints and returns
    std::vector<Vec3> pos;
    std::vector<uint32_t> idx;
    if (buf.size() >= 12 && std::memcmp(buf.data(), "DODM", 4) == 0) {
        uint32_t nv, nt;
        std::memcpy(&nv, buf.data() + 4, 4);
        std::memcpy(&nt, buf.data() + 8, 4);
        if (buf.size() < 12 + (size_t)nv * 12 + (size_t)nt * 12) return fail("%s: truncated DODM mesh", path);
        pos.resize(nv);
        std::memcpy(static_cast<void *>(pos.data()), buf.data() + 12, (size_t)nv * 12);
        idx.resize((size_t)nt * 3);
        std::memcpy(idx.data(), buf.data() + 12 + (size_t)nv * 12, (size_t)nt * 12);
    } else {
        parseObj(buf, pos, idx);
    }
    if (idx.empty()) return fail("%s: no faces", path);
    return addMesh(s, pos, idx.data(), (uint32_t)(idx.size() / 3), transform);
}
DODRT_HOST_CATCH

int dodrt_host_standin_dragon(uint32_t n, float *positions, uint32_t *indices)
try {
    if (n < 3 || !positions || !indices) return fail("bad argument");
    const double pi = 3.14159265358979323846;
    const uint32_t stride = n + 1;
    for (uint32_t j = 0; j <= n; j++) {
        const double v = 0.02 + (pi - 0.04) * (double)j / (double)n;
        for (uint32_t i = 0; i <= n; i++) {
            const double u = 2.0 * pi * (double)i / (double)n;
            const double r = 2.2 + 0.25 * std::sin(7.0 * u) * std::sin(5.0 * v) + 0.08 * std::sin(31.0 * u + 3.0) * std::sin(29.0 * v);
            float *p = positions + ((size_t)j * stride + i) * 3;
            p[0] = (float)(r * std::sin(v) * std::cos(u));
            p[1] = (float)(r * std::cos(v));
            p[2] = (float)(r * std::sin(v) * std::sin(u));
        }
    }
    uint32_t *t = indices;
    for (uint32_t j = 0; j < n; j++) {
        for (uint32_t i = 0; i < n; i++) { // quad a b / d c split into (a,b,c), (c,d,a)
            const uint32_t a = j * stride + i, b = a + 1, c = a + stride + 1, d = a + stride;
            *t++ = a; *t++ = b; *t++ = c;
            *t++ = c; *t++ = d; *t++ = a;
        }
    }
    return DODRT_OK;
}
DODRT_HOST_CATCH

int dodrt_host_write_dodm(const char *path, const float *positions, uint32_t numVertices, const uint32_t *indices,
                          uint32_t numTriangles)
try {
    FILE *f = std::fopen(path, "wb");
    if (!f) return fail("cannot write %s", path);
    const uint32_t hdr[2] = {numVertices, numTriangles};
    bool ok = std::fwrite("DODM", 1, 4, f) == 4 && std::fwrite(hdr, 4, 2, f) == 2 &&
              std::fwrite(positions, 12, numVertices, f) == numVertices &&
              std::fwrite(indices, 12, numTriangles, f) == numTriangles;
    std::fclose(f);
    return ok ? DODRT_OK : fail("short write to %s", path);
}
DODRT_HOST_CATCH

int dodrt_host_add_sphere(dodrt_host_scene *s, const float pos[3], float radius, const float color[3])
try {
    if (!s || !pos) return fail("NULL argument");
    const float vals[4] = {pos[0], pos[1], pos[2], radius * radius}; // sphere.cpp:234-238
    appendLane(s->sphereLanes, s->numSpheres, 4 * kLane, vals, 4);
    for (int k = 0; k < 3; k++) s->sphereColors.push_back(color ? color[k] : 0.0f);
    s->numSpheres++;
    return DODRT_OK;
}
DODRT_HOST_CATCH

int dodrt_host_add_plane(dodrt_host_scene *s, const float normal[3], const float pos[3], const float color[3])
try {
    if (!s || !normal || !pos) return fail("NULL argument");
    const float vals[6] = {pos[0], pos[1], pos[2], normal[0], normal[1], normal[2]}; // plane.cpp:212-217
    appendLane(s->planeLanes, s->numPlanes, 6 * kLane, vals, 6);
    for (int k = 0; k < 3; k++) s->planeColors.push_back(color ? color[k] : 0.0f);
    s->numPlanes++;
    return DODRT_OK;
}
DODRT_HOST_CATCH

int dodrt_host_add_cylinder(dodrt_host_scene *s, float radius, float height, const float axis[3], const float base[3])
try {
    if (!s || !axis || !base) return fail("NULL argument");
    dodrt_cylinder c; // Cylinder::Cylinder, cylinder.cpp:223-229: axis = glm::normalize(axis) = v * (1/sqrt(dot))
    const Vec3 a{axis[0], axis[1], axis[2]};
    const float inv = 1.0f / sqrtf(dot(a, a));
    c.axis[0] = a.x * inv;
    c.axis[1] = a.y * inv;
    c.axis[2] = a.z * inv;
    for (int k = 0; k < 3; k++) c.base[k] = base[k];
    c.radius_sq = radius * radius;
    c.height = height;
    s->cylinders.push_back(c);
    return DODRT_OK;
}
DODRT_HOST_CATCH

int dodrt_host_add_box(dodrt_host_scene *s, const float lo[3], const float hi[3])
try {
    if (!s || !lo || !hi) return fail("NULL argument");
    const float vals[6] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
    appendLane(s->boxLanes, s->numBoxes, 6 * kLane, vals, 6);
    s->numBoxes++;
    return DODRT_OK;
}
DODRT_HOST_CATCH

int dodrt_host_add_reference_scene(dodrt_host_scene *s, uint32_t seed, uint32_t numSpheres)
try {
    if (!s) return fail("NULL argument");
    GlibcRand rnd(seed); // srand(seed) in place of srand(time(NULL)), main.cpp:351
    for (uint32_t i = 0; i < numSpheres; i++) { // generateSpheres, main.cpp:26-50: r g b x y z
        float color[3], pos[3];
        for (int k = 0; k < 3; k++) color[k] = ((float)rnd() / kRandMax);
        for (int k = 0; k < 3; k++) pos[k] = ((float)rnd() / kRandMax) * 10.0f - 5.0f;
        dodrt_host_add_sphere(s, pos, 1.0f, color);
    }
    // generatePlanes, main.cpp:52-109: {normal, position, colour}; double literals narrow to float
    static const float planes[6][9] = {
        {0.0f, 0.0f, -1.0f, 0.0f, 0.0f, 5.0f, 0.195f, 0.410f, 0.610f},
        {0.0f, 0.0f, 1.0f, 0.0f, 0.0f, -5.0f, (float)0.493, (float)0.265, (float)0.590},
        {0.0f, -1.0f, 0.0f, 0.0f, 5.0f, 0.0f, (float)0.276, (float)0.600, (float)0.411},
        {0.0f, 1.0f, 0.0f, 0.0f, -5.0f, 0.0f, (float)0.292, (float)0.680, (float)0.674},
        {1.0f, 0.0f, 0.0f, -5.0f, 0.0f, 0.0f, (float)0.720, (float)0.288, (float)0.389},
        {-1.0f, 0.0f, 0.0f, 5.0f, 0.0f, 0.0f, (float)0.680, (float)0.224, (float)0.224},
    };
    for (const float *p : planes) dodrt_host_add_plane(s, p, p + 3, p + 6);
    // generateCylinders, main.cpp:111-129 (its three rand() draws colour a cylinder that renders black)
    const float axis[3] = {(float)2.2, 5, 2}, base[3] = {-2, 0, 2};
    for (int k = 0; k < 3; k++) (void)rnd();
    return dodrt_host_add_cylinder(s, 1.5f, 4.0f, axis, base);
}
DODRT_HOST_CATCH

int dodrt_host_add_analytic_scene(dodrt_host_scene *s, uint32_t seed, uint32_t count)
try {
    if (!s) return fail("NULL argument");
    uint32_t x = seed;
    auto draw = [&x]() {
        x = 1664525u * x + 1013904223u;
        return (float)(x >> 8) / 16777216.0f;
    };
    for (uint32_t i = 0; i < count; i++) {
        float u[8];
        for (float &v : u) v = draw();
        const float sc[3] = {u[0] * 9.0f - 4.5f, u[1] * 9.0f - 4.5f, u[2] * 9.0f - 4.5f};
        const float sr = u[3] * 0.09f + 0.03f;
        const float bc[3] = {u[4] * 9.0f - 4.5f, u[5] * 9.0f - 4.5f, u[6] * 9.0f - 4.5f};
        const float bh = u[7] * 0.09f + 0.03f;
        const float lo[3] = {bc[0] - bh, bc[1] - bh, bc[2] - bh}, hi[3] = {bc[0] + bh, bc[1] + bh, bc[2] + bh};
        dodrt_host_add_sphere(s, sc, sr, nullptr);
        dodrt_host_add_box(s, lo, hi);
    }
    return DODRT_OK;
}
DODRT_HOST_CATCH

int dodrt_host_build_tree_ex(dodrt_host_scene *s, uint32_t flags)
try {
    if (!s) return fail("NULL argument");
    if (s->built) return fail("tree already built");
    if (flags & ~(uint32_t)DODRT_HOST_BUILD_KEEP_CREATION_ORDER) return fail("unknown build flags 0x%x", flags);
    const unsigned threads = hostThreads();
    s->numOrigLanes = (uint32_t)s->lanes.size();
    const auto t0 = std::chrono::steady_clock::now();
    TreeBuilder tb(s, threads);
    tb.run();
    const auto t1 = std::chrono::steady_clock::now();
    if (!(flags & DODRT_HOST_BUILD_KEEP_CREATION_ORDER)) {
        // Triangle::reorderLanesByIndices, triangle.cpp:349-367
        std::vector<TriLane> lanes;
        std::vector<NormalLane> normals;
        std::vector<AttrLane> attrs;
        gatherLanes(s->lanes, s->primNums, lanes, threads);
        s->lanes.swap(lanes);
        std::vector<TriLane>().swap(lanes);
        gatherLanes(s->normals, s->primNums, normals, threads);
        s->normals.swap(normals);
        std::vector<NormalLane>().swap(normals);
        gatherLanes(s->attrs, s->primNums, attrs, threads);
        s->attrs.swap(attrs);
    }
    s->creationOrder = (flags & DODRT_HOST_BUILD_KEEP_CREATION_ORDER) != 0;
    s->buildSeconds = std::chrono::duration<double>(t1 - t0).count();
    s->reorderSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    if (std::getenv("DODRT_HOST_VERBOSE")) {
        std::fprintf(stderr, "dodrt_host: kd build %.3f s, lane re-order %.3f s, %u threads, %zu nodes, %zu lanes\n",
                     s->buildSeconds, s->reorderSeconds, threads, s->nodes.size(), s->primNums.size());
    }
    for (int i = 0; i < 3; i++) {
        s->boundsOut[i] = s->bounds.lo[i];
        s->boundsOut[3 + i] = s->bounds.hi[i];
    }
    s->built = true;
    return DODRT_OK;
}
DODRT_HOST_CATCH

int dodrt_host_build_tree(dodrt_host_scene *s) { return dodrt_host_build_tree_ex(s, 0); }

int dodrt_host_sizes_get(const dodrt_host_scene *s, dodrt_host_sizes *z)
try {
    if (!s || !z) return fail("NULL argument");
    z->num_triangles = s->numTriangles;
    z->num_orig_lanes = s->built ? s->numOrigLanes : (uint32_t)s->lanes.size();
    z->num_nodes = (uint32_t)s->nodes.size();
    z->num_lanes = s->built ? (uint32_t)s->primNums.size() : (uint32_t)s->lanes.size();
    z->max_depth = s->maxDepth;
    z->num_spheres = s->numSpheres;
    z->num_planes = s->numPlanes;
    z->num_cylinders = (uint32_t)s->cylinders.size();
    z->num_boxes = s->numBoxes;
    return DODRT_OK;
}
DODRT_HOST_CATCH

const uint64_t *dodrt_host_nodes(const dodrt_host_scene *s) { return s->nodes.data(); }
const float *dodrt_host_tri_lanes(const dodrt_host_scene *s) { return reinterpret_cast<const float *>(s->lanes.data()); }
const uint32_t *dodrt_host_prim_nums(const dodrt_host_scene *s) { return s->primNums.data(); }
const float *dodrt_host_bounds(const dodrt_host_scene *s) { return s->boundsOut; }
const float *dodrt_host_tri_normals(const dodrt_host_scene *s) { return reinterpret_cast<const float *>(s->normals.data()); }
const void *dodrt_host_tri_attributes(const dodrt_host_scene *s) { return s->attrs.data(); }
const float *dodrt_host_mesh_colors(const dodrt_host_scene *s) { return s->meshColors.data(); }
uint32_t dodrt_host_num_meshes(const dodrt_host_scene *s) { return (uint32_t)(s->meshColors.size() / 3); }
const float *dodrt_host_sphere_lanes(const dodrt_host_scene *s) { return s->sphereLanes.data(); }
const float *dodrt_host_sphere_colors(const dodrt_host_scene *s) { return s->sphereColors.data(); }
const float *dodrt_host_plane_lanes(const dodrt_host_scene *s) { return s->planeLanes.data(); }
const float *dodrt_host_plane_colors(const dodrt_host_scene *s) { return s->planeColors.data(); }
const dodrt_cylinder *dodrt_host_cylinders(const dodrt_host_scene *s) { return s->cylinders.data(); }
const float *dodrt_host_box_lanes(const dodrt_host_scene *s) { return s->boxLanes.data(); }
float dodrt_host_epsilon(const dodrt_host_scene *s) { return s->cfg.epsilon; }

int dodrt_host_ray_tables(uint32_t width, uint32_t height, float *xs, float *ys)
try {
    if (!width || !height || !xs || !ys) return fail("bad argument");
    const float ratio = (float)width / height; // Config::Ratio, config.h:27
    const float widthStep = 2.0f * ratio / width; // main.cpp:278-279
    const float heightStep = 2.0f / height;
    float x = -ratio; // main.cpp:276
    for (uint32_t j = 0; j < width; j++) {
        xs[j] = x;
        x += widthStep; // main.cpp:342
    }
    float y = 1.0f;
    for (uint32_t i = 0; i < height; i++) {
        ys[i] = y;
        y -= heightStep; // main.cpp:345 (canonical single band: startRow = 0, main.cpp:295)
    }
    return DODRT_OK;
}
DODRT_HOST_CATCH

} // extern "C"

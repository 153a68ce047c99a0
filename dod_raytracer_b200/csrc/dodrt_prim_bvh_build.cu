// dodrt_prim_bvh_build.cu -- host-side builder of the sphere / box culling BVH (see dodrt_prim_bvh.cuh).
// Any tree shape gives the same query results (the traversal is conservative and reduces with (t, id)); the shape
// only decides how much is culled: median split of the primitive centres along the longest axis, leaves <= 8.
#include "dodrt_prim_bvh.cuh"

#include <algorithm>
#include <cstring>

namespace dodrt {

namespace {

struct Builder {
    const float *boxes;
    std::vector<PrimBvhNode> &nodes;
    std::vector<uint32_t> &ids;

    uint32_t build(uint32_t begin, uint32_t end)
    {
        const uint32_t index = (uint32_t)nodes.size();
        nodes.emplace_back();
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        for (uint32_t i = begin; i < end; i++) {
            const float *b = boxes + (size_t)ids[i] * 6;
            for (int k = 0; k < 3; k++) {
                lo[k] = std::min(lo[k], b[k]);
                hi[k] = std::max(hi[k], b[3 + k]);
            }
        }
        PrimBvhNode n;
        std::memcpy(n.bmin, lo, sizeof(lo));
        std::memcpy(n.bmax, hi, sizeof(hi));
        if (end - begin <= (uint32_t)kPrimBvhLeafSize) {
            std::sort(ids.begin() + begin, ids.begin() + end);
            n.a = begin;
            n.b = kPrimBvhLeaf | (end - begin);
            nodes[index] = n;
            return index;
        }
        int axis = 0;
        if (hi[1] - lo[1] > hi[axis] - lo[axis]) axis = 1;
        if (hi[2] - lo[2] > hi[axis] - lo[axis]) axis = 2;
        const uint32_t mid = begin + (end - begin) / 2;
        std::nth_element(ids.begin() + begin, ids.begin() + mid, ids.begin() + end, [&](uint32_t x, uint32_t y) {
            const float cx = boxes[(size_t)x * 6 + axis] + boxes[(size_t)x * 6 + 3 + axis];
            const float cy = boxes[(size_t)y * 6 + axis] + boxes[(size_t)y * 6 + 3 + axis];
            return cx < cy || (cx == cy && x < y);
        });
        n.a = build(begin, mid);
        n.b = build(mid, end);
        nodes[index] = n;
        return index;
    }
};

} // namespace

void build_prim_bvh(const float *boxes, uint32_t count, std::vector<PrimBvhNode> &nodes, std::vector<uint32_t> &ids)
{
    nodes.clear();
    ids.resize(count);
    for (uint32_t i = 0; i < count; i++) ids[i] = i;
    if (count == 0) return;
    nodes.reserve(2 * (count / kPrimBvhLeafSize + 1));
    Builder b{boxes, nodes, ids};
    b.build(0, count);
}

} // namespace dodrt

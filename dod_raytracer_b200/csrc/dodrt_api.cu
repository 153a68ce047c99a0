// dodrt_api.cu -- the extern "C" boundary declared in include/dodrt.h.
//
// Host-side responsibilities only: argument validation, scene upload (one replica per GPU, kept
// resident in HBM for the lifetime of the handle), staging of host buffers, kernel launches and
// error translation.  All ray/shape arithmetic lives in dodrt_device.cuh / dodrt_kernels.cu.
// There is deliberately no CPU implementation behind any entry point.
#include "dodrt_kernels.cuh"
#include "dodrt_prim_bvh.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace dodrt;

namespace {

thread_local std::string g_lastError;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_lastError = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            return fail(DODRT_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,     \
                        __LINE__);                                                                           \
        }                                                                                                    \
    } while (0)

// The ABI promises "never throws": every extern "C" body is a function-try-block that ends in DODRT_CATCH.
int failException()
{
    try {
        throw;
    } catch (const std::bad_alloc &) {
        return fail(DODRT_E_NOMEM, "out of host memory");
    } catch (const std::exception &e) {
        return fail(DODRT_E_INVALID, "unexpected C++ exception: %s", e.what());
    } catch (...) {
        return fail(DODRT_E_INVALID, "unexpected C++ exception");
    }
}
#define DODRT_CATCH                                                                                           \
    catch (...) { return failException(); }

constexpr int kCounterSlots = 256;

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(device) == cudaSuccess) {
            ok = true;
        }
    }
    ~DeviceGuard()
    {
        if (prev >= 0) {
            cudaSetDevice(prev);
        }
    }
};

template <typename T> void freeDevice(T *&p)
{
    if (p) {
        cudaFree(const_cast<void *>(static_cast<const void *>(p)));
        p = nullptr;
    }
}

} // namespace

struct dodrt_scene {
    int device = 0;
    DeviceScene dev{};
    uint2 *d_nodes = nullptr;
    float4 *d_tris = nullptr;
    float4 *d_lanes4 = nullptr;
    float *d_spheres = nullptr;
    float *d_planes = nullptr;
    dodrt_cylinder *d_cylinders = nullptr;
    float *d_boxes = nullptr;
    float4 *d_sphereBvh = nullptr, *d_boxBvh = nullptr;
    uint32_t *d_sphereBvhIds = nullptr, *d_boxBvhIds = nullptr;
    uint32_t *d_triAttrs = nullptr;
    float *d_meshColors = nullptr, *d_sphereColors = nullptr, *d_planeColors = nullptr;
    unsigned long long *d_counters = nullptr;
    unsigned long long *d_stats = nullptr; // dodrt_scene_debug_stats
    std::atomic<uint32_t> nextCounter{0};
    std::atomic<uint64_t> launches{0};
    LaunchConfig cfg[kNumVariants][kNumModes]{};
    int variant = kVariantAuto; // kVariantAuto or an explicit variant (dodrt_scene_set_kernel_variant / DODRT_VARIANT)
    uint32_t treeDepth = 0;
    std::mutex mutex; // guards scene mutation and the lazily created staging stream
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr; // D2H of finished bands while the next band is traced
    // Frame shares traced in CHUNKS on concurrent streams (traceFrameOnDevice): chunk c's primary / shadow passes run on
    // chunkStream[c-1] (chunk 0 on the caller's stream), fenced with events from a recycled ring
    static constexpr int kMaxChunks = 4;
    cudaStream_t chunkStream[kMaxChunks - 1] = {nullptr, nullptr, nullptr};
    static constexpr int kChunkEvents = 64;
    cudaEvent_t chunkEvent[kChunkEvents] = {};
    std::atomic<uint32_t> nextChunkEvent{0};
    // staging memory of the host-buffer entry points: a private pool that keeps freed blocks cached
    // (release threshold = max), so a per-frame call does not pay for physical allocation every time
    cudaMemPool_t pool = nullptr;
    // Persistent donation queues of the donating kernel (variant 7): a small ring, each zero-filled once and then
    // re-used with a fresh epoch per launch, so a pass costs neither an allocation nor a memset of the ready words.
    // Launches on one stream share a queue (they cannot overlap); a queue moves to another stream only when the launch
    // that used it has finished (its event); with more than kDonateQueues streams in flight a launch falls back to a
    // per-launch pool allocation.
    static constexpr int kDonateQueues = 4;
    void *donateQueue[kDonateQueues] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t donateDone[kDonateQueues] = {nullptr, nullptr, nullptr, nullptr};
    bool donateBusy[kDonateQueues] = {false, false, false, false};
    cudaStream_t donateOwner[kDonateQueues] = {nullptr, nullptr, nullptr, nullptr}; // stream of the launch that used it last
    std::atomic<uint32_t> donateEpoch{0};
    // Host-buffer entry points (dodrt_trace_frame & co): they run on the scene's own stream one call at a time
    // (hostMutex), so their staging memory and fence events are PERSISTENT -- grown on demand, never freed before the
    // scene is: a per-frame call pays for no allocation, no event creation and no pool bookkeeping.
    std::mutex hostMutex;
    float *stTables = nullptr;
    size_t stTablesBytes = 0;
    dodrt_hit *stHits = nullptr;
    size_t stHitsBytes = 0;
    uint8_t *stVis = nullptr;
    size_t stVisBytes = 0;
    std::vector<cudaEvent_t> stEvents;
    size_t stEventsUsed = 0;
};

// dodrt_frame_buffer: a row-major frame in one GPU's HBM that kernels on other GPUs write into (include/dodrt.h)
struct dodrt_frame_buffer {
    int device = 0;      // GPU whose kernels use this view (owner: where the memory lives)
    int ownerDevice = 0;
    bool owner = false;  // allocated here (cudaFree) ...
    bool ipc = false;    // ... or opened from another process (cudaIpcCloseMemHandle); neither = same-process peer view
    uint32_t width = 0, height = 0, numLights = 0;
    size_t bytes = 0;
    char *base = nullptr;
    dodrt_hit *hits() const { return reinterpret_cast<dodrt_hit *>(base); }
    uint8_t *visible() const { return reinterpret_cast<uint8_t *>(base + (size_t)width * height * sizeof(dodrt_hit)); }
};

namespace {

// depth of the DFS pre-order tree (left child = node+1, right child index in word0 >> 2), and
// validation of every index the kernels will dereference
int validateTree(const uint64_t *nodes, uint32_t numNodes, uint32_t numLanes, uint32_t *depthOut)
{
    if (numNodes == 0) {
        *depthOut = 0;
        return DODRT_OK;
    }
    struct Item {
        uint32_t node, depth;
    };
    std::vector<Item> stack;
    stack.push_back({0, 1});
    uint32_t maxDepth = 0;
    uint64_t visited = 0;
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        if (it.node >= numNodes) {
            return fail(DODRT_E_INVALID, "kd-tree child index %u out of range (%u nodes)", it.node, numNodes);
        }
        if (++visited > numNodes) {
            return fail(DODRT_E_INVALID, "kd-tree is not a tree (cycle or shared node)");
        }
        maxDepth = it.depth > maxDepth ? it.depth : maxDepth;
        const uint32_t w0 = (uint32_t)(nodes[it.node] & 0xFFFFFFFFu);
        const uint32_t w1 = (uint32_t)(nodes[it.node] >> 32);
        if ((w0 & 3u) == kLeafFlag) {
            const uint64_t n = w0 >> 2;
            if (n && (uint64_t)w1 + n > numLanes) {
                return fail(DODRT_E_INVALID, "leaf %u references lanes [%u,%llu) beyond %u", it.node, w1,
                            (unsigned long long)(w1 + n), numLanes);
            }
        } else {
            stack.push_back({w0 >> 2, it.depth + 1});
            stack.push_back({it.node + 1, it.depth + 1});
        }
    }
    *depthOut = maxDepth;
    return DODRT_OK;
}

constexpr uint32_t kMaxTileSide = 4096; // keeps tile_w * tile_h and (width + tile_w - 1) inside 32 bits
int checkFrame(const dodrt_frame *f)
{
    if (!f) return fail(DODRT_E_INVALID, "frame is NULL");
    if (f->width == 0 || f->height == 0) return fail(DODRT_E_INVALID, "empty frame %ux%u", f->width, f->height);
    if (f->tile_w == 0 || f->tile_h == 0 || (f->tile_w % 8) || (f->tile_h % 4)) {
        return fail(DODRT_E_INVALID, "tile %ux%u must be a non-zero multiple of 8x4", f->tile_w, f->tile_h);
    }
    if (f->tile_w > kMaxTileSide || f->tile_h > kMaxTileSide) {
        return fail(DODRT_E_INVALID, "tile %ux%u exceeds the limit of %u per side", f->tile_w, f->tile_h, kMaxTileSide);
    }
    if (f->tile_stride == 0) return fail(DODRT_E_INVALID, "tile_stride must be >= 1");
    if ((uint64_t)f->width * f->height >= 0xFFFFFFFFull) return fail(DODRT_E_INVALID, "frame too large");
    return DODRT_OK;
}

void frameTiles(const dodrt_frame *f, uint32_t *tilesX, uint32_t *localTiles)
{
    const uint32_t tx = (uint32_t)(((uint64_t)f->width + f->tile_w - 1) / f->tile_w);
    const uint32_t ty = (uint32_t)(((uint64_t)f->height + f->tile_h - 1) / f->tile_h);
    const uint64_t total = (uint64_t)tx * ty;
    *tilesX = tx;
    *localTiles = f->first_tile < total ? (uint32_t)((total - f->first_tile + f->tile_stride - 1) / f->tile_stride) : 0;
}

uint64_t frameSlots(const dodrt_frame *f)
{
    uint32_t tx, lt;
    frameTiles(f, &tx, &lt);
    return (uint64_t)lt * f->tile_w * f->tile_h;
}

unsigned long long *nextCounter(dodrt_scene *s)
{
    return s->d_counters + (size_t)(s->nextCounter.fetch_add(1) % kCounterSlots) * kCounterWords;
}

int ensurePool(dodrt_scene *s)
{
    std::lock_guard<std::mutex> lock(s->mutex);
    if (!s->pool) {
        DeviceGuard guard(s->device);
        if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = s->device;
        CUDA_TRY(cudaMemPoolCreate(&s->pool, &props));
        uint64_t keep = UINT64_MAX;
        CUDA_TRY(cudaMemPoolSetAttribute(s->pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    return DODRT_OK;
}

int ensureStream(dodrt_scene *s)
{
    int rc = ensurePool(s);
    if (rc != DODRT_OK) return rc;
    std::lock_guard<std::mutex> lock(s->mutex);
    // the streams must belong to the scene's device, whatever the calling thread's current device is
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    if (!s->stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    }
    if (!s->copyStream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s->copyStream, cudaStreamNonBlocking));
    }
    return DODRT_OK;
}

// launch_trace with the scene's donation machinery: a persistent queue when one is free, else the pool.
// The scene mutex is held from the choice of a queue until the launch has been enqueued AND its completion event
// recorded: donateDone[q] therefore always describes the LAST launch that used queue q, and another host thread can
// never see "busy" together with an event that is unrecorded or still holds an older, finished record (which would
// hand one queue to two concurrent launches).  Launches are asynchronous, so the lock is held for microseconds.
cudaError_t launchTraceOn(dodrt_scene *s, TraceMode mode, const TraceParams &p, cudaStream_t stream)
{
    const LaunchConfig &cfg = s->cfg[p.variant][mode];
    if (p.variant != kDonateVariant || s->dev.num_nodes == 0 || !(p.classes & DODRT_CLS_TREE)) {
        return launch_trace(mode, p, cfg, stream, nullptr);
    }
    std::unique_lock<std::mutex> lock(s->mutex);
    int slot = -1;
    // launches on one stream run one after the other, so they can share a queue without waiting for anything
    for (int q = 0; q < dodrt_scene::kDonateQueues && slot < 0; q++) {
        if (s->donateQueue[q] && s->donateBusy[q] && s->donateOwner[q] == stream) slot = q;
    }
    for (int q = 0; q < dodrt_scene::kDonateQueues && slot < 0; q++) {
        if (s->donateBusy[q] && cudaEventQuery(s->donateDone[q]) == cudaSuccess) s->donateBusy[q] = false;
        if (s->donateBusy[q]) continue; // in flight on another stream
        if (!s->donateQueue[q]) {
            size_t bytes = 0; // a queue serves every mode: size it for the largest grid of the donating kernels
            for (int m = 0; m < kNumModes; m++) bytes = std::max(bytes, donation_queue_bytes(s->cfg[kDonateVariant][m]));
            if (cudaMalloc(&s->donateQueue[q], bytes) != cudaSuccess ||
                cudaMemset(s->donateQueue[q], 0, bytes) != cudaSuccess ||
                cudaEventCreateWithFlags(&s->donateDone[q], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                if (s->donateQueue[q]) cudaFree(s->donateQueue[q]);
                s->donateQueue[q] = nullptr;
                break; // out of memory: fall back to the pool path below
            }
        }
        slot = q;
    }
    if (slot < 0) {
        lock.unlock();
        return launch_trace(mode, p, cfg, stream, s->pool);
    }
    uint32_t epoch = s->donateEpoch.fetch_add(1) + 1u;
    if (epoch == 0u) epoch = s->donateEpoch.fetch_add(1) + 1u; // 0 marks "never written"
    cudaError_t e = launch_trace(mode, p, cfg, stream, nullptr, s->donateQueue[slot], epoch);
    const cudaError_t er = cudaEventRecord(s->donateDone[slot], stream);
    // busy only with a recorded event behind it; if the record failed nobody may wait on the stale event
    s->donateBusy[slot] = er == cudaSuccess;
    s->donateOwner[slot] = stream;
    return e != cudaSuccess ? e : er;
}

// Second destination of a frame pass's results, written by the kernels themselves (TraceParams::mirror_*): a peer GPU's
// frame buffer (by pixel) or the caller's pinned host buffers (same layout as the device results).
struct Mirror {
    dodrt_hit *hits = nullptr;
    uint8_t *visible = nullptr; // of the light the pass handles (two-launch path) / of light 0 (fused path)
    uint64_t lightStride = 0;
    uint32_t byPixel = 0;
};

// Traces local tiles [tileBegin, tileBegin + tileCount) of the frame (tileCount is clamped); d_hits / d_visible
// always point at slot 0 of the call's result buffers.
int launchFrame(dodrt_scene *s, TraceMode mode, const dodrt_frame *frame, const float *d_xs, const float *d_ys,
                dodrt_hit *d_hits, const float *light, uint8_t *d_visible, cudaStream_t stream, uint32_t tileBegin = 0,
                uint32_t tileCount = 0xFFFFFFFFu, const Mirror *mirror = nullptr)
{
    TraceParams p{};
    p.scene = s->dev;
    p.classes = frame->classes;
    p.frame = *frame;
    if (mirror) {
        p.mirror_by_pixel = mirror->byPixel;
        p.mirror_light_stride = mirror->lightStride;
    }
    uint32_t localTiles;
    frameTiles(frame, &p.tiles_x, &localTiles);
    if (tileBegin >= localTiles) {
        return DODRT_OK;
    }
    const uint32_t tiles = tileCount < localTiles - tileBegin ? tileCount : localTiles - tileBegin;
    const uint64_t tilePixels = (uint64_t)frame->tile_w * frame->tile_h;
    const uint64_t slotBase = frame->compact ? tileBegin * tilePixels : 0; // full-frame results are indexed by pixel
    p.frame.first_tile = frame->first_tile + tileBegin * frame->tile_stride;
    p.count = tiles * tilePixels;
    p.hits = d_hits + slotBase;
    p.xs = d_xs;
    p.ys = d_ys;
    p.visible = d_visible ? d_visible + slotBase : nullptr;
    if (mirror) { // by pixel: absolute indices; else the mirror has the layout of the local buffers
        const uint64_t mbase = mirror->byPixel ? 0 : slotBase;
        p.mirror_hits = (mode == kModePrimary && mirror->hits) ? mirror->hits + mbase : nullptr;
        p.mirror_visible = (mode == kModeShadow && mirror->visible) ? mirror->visible + mbase : nullptr;
    }
    if (light) {
        p.light[0] = light[0];
        p.light[1] = light[1];
        p.light[2] = light[2];
    }
    p.counter = nextCounter(s);
    p.variant = resolve_variant(s->variant, s->cfg[kDonateVariant][mode], p.count, frame->tile_stride > 1, s->dev.num_nodes >= kBigTreeNodes);
    if (p.count == 0) {
        return DODRT_OK;
    }
    // heavy-first tile order (see order_tiles_kernel): worth it when the kd-tree is in play and there are
    // enough tiles to reorder
    p.num_local_tiles = tiles;
    p.tile_order = nullptr;
    static const bool orderTiles = [] { const char *e = std::getenv("DODRT_TILE_ORDER"); return !e || std::atoi(e) != 0; }();
    // (not for the donating kernel: its tail is spread over the idle warps anyway -- measured 0.750 vs 0.760 ms per rank of
    // 8 -- and the order kernel plus its allocation are ~10 us of a 0.25 ms pass)
    if (orderTiles && p.variant != kDonateVariant && (frame->classes & DODRT_CLS_TREE) && s->dev.num_nodes != 0 && tiles >= 64) {
        int rc = ensurePool(s);
        if (rc != DODRT_OK) return rc;
        CUDA_TRY(cudaMallocFromPoolAsync(&p.tile_order, sizeof(uint32_t) * tiles, s->pool, stream));
    }
    if (p.variant == kDonateVariant) {
        int rc = ensurePool(s);
        if (rc != DODRT_OK) return rc;
    }
    cudaError_t le = launchTraceOn(s, mode, p, stream);
#ifdef DODRT_TIMELINE
    { // debug build: where does a pass spend its time?  (start, first warp out of work, last warp busy, last helper gone)
        unsigned long long tl[8];
        cudaStreamSynchronize(stream);
        cudaMemcpy(tl, p.counter + 24, 64, cudaMemcpyDeviceToHost);
        unsigned long long served = 0;
        cudaMemcpy(&served, p.counter + kDonateServed, 8, cudaMemcpyDeviceToHost);
        if (served) {
            std::fprintf(stderr, "   resumed rays %llu: node steps %.1f per ray at %.0f cycles each, leaf steps (128 slots) %.1f per ray at %.0f cycles each\n",
                         served, (double)tl[6] / served, tl[6] ? (double)tl[4] / tl[6] : 0.0,
                         (double)tl[7] / served, tl[7] ? (double)tl[5] / tl[7] : 0.0);
        }
        unsigned long long ex[2] = {0, 0}; // [21] warps, [22] sum of their main-loop exit times
        cudaMemcpy(ex, p.counter + 21, 16, cudaMemcpyDeviceToHost);
        std::fprintf(stderr, "timeline mode %d variant %d count %llu: first-idle %+.1f us, mean main-loop exit %+.1f us (%llu warps), last-busy %+.1f us, end %+.1f us; resumed rays took %.1f us per warp\n",
                     (int)mode, p.variant, (unsigned long long)p.count, (tl[1] - tl[0]) * 1e-3,
                     ex[0] ? ((double)ex[1] / ex[0] - (double)(tl[0] & 0xFFFFFFFFFull)) * 1e-3 : 0.0, ex[0], (tl[2] - tl[0]) * 1e-3, (tl[3] - tl[0]) * 1e-3,
                     ex[0] ? (double)(tl[4] + tl[5]) / ex[0] / 1.92e3 : 0.0);
        { // histograms over 16-us bins since the kernel started: warps leaving their main loop; donations (events, rays given, rays kept)
            static unsigned long long exits[8192], dons[1 << 16];
            unsigned int nd = 0, pollsEmpty = 0;
            timeline_fetch(exits, dons, &nd, &pollsEmpty);
            unsigned hExit[48] = {0}, hDon[48] = {0}, hGiven[48] = {0}, hKept[48] = {0};
            for (int i = 0; i < 8192; i++) {
                if (exits[i] >= tl[0]) hExit[std::min<unsigned long long>((exits[i] - tl[0]) / 16000ull, 47ull)]++;
            }
            const unsigned long long t0m = tl[0] & ((1ull << 52) - 1ull);
            for (unsigned i = 0; i < std::min(nd, 1u << 16); i++) {
                const unsigned long long t = dons[i] >> 12;
                const unsigned long long b = t >= t0m ? std::min<unsigned long long>((t - t0m) / 16000ull, 47ull) : 47ull;
                hDon[b]++, hGiven[b] += (unsigned)((dons[i] >> 6) & 63u), hKept[b] += (unsigned)(dons[i] & 63u);
            }
            std::fprintf(stderr, "   %u donations, %u polls found no waiting helper\n   bin(16us): exits | donations given kept\n", nd, pollsEmpty);
            for (int b = 0; b < 48; b++) {
                if (hExit[b] | hDon[b]) std::fprintf(stderr, "   %3d: %5u | %5u %5u %5u\n", b, hExit[b], hDon[b], hGiven[b], hKept[b]);
            }
        }
    }
#endif
    if (p.tile_order) {
        cudaFreeAsync(p.tile_order, stream);
        s->launches.fetch_add(1);
    }
    if (le != cudaSuccess) return fail(DODRT_E_CUDA, "kernel launch failed: %s", cudaGetErrorString(le));
    s->launches.fetch_add(1);
    return DODRT_OK;
}


// Primary + shadow queues of the frame share in ONE launch (trace_frame_kernel).  d_visible: [num_lights][visStride].
// Returns DODRT_OK with *done = false when the selected kernel variant has no fused form (A/B variants): the caller
// then runs the separate passes.
int launchFrameFused(dodrt_scene *s, const dodrt_frame *frame, const float *d_xs, const float *d_ys, const float *lights,
                     uint32_t numLights, dodrt_hit *d_hits, uint8_t *d_visible, uint64_t visStride, const Mirror *mirror,
                     cudaStream_t stream, bool *done)
{
    *done = false;
    // A/B knob, read per call: 1 = the one-launch frame kernels.  Default: separate primary / shadow passes, which are
    // faster at every share size (profiles/r02_frame_kernel_ab.txt)
    const char *fusedEnv = std::getenv("DODRT_FUSED");
    const bool fusedOn = mode_compiled(kModeFrame) && fusedEnv && std::atoi(fusedEnv) != 0;
    if (!fusedOn || numLights > (uint32_t)kMaxLights) return DODRT_OK;
    TraceParams p{};
    p.scene = s->dev;
    p.classes = frame->classes;
    p.frame = *frame;
    uint32_t tiles;
    frameTiles(frame, &p.tiles_x, &tiles);
    const uint64_t tilePixels = (uint64_t)frame->tile_w * frame->tile_h;
    p.count = tiles * tilePixels;
    p.shadow_count = p.count * numLights;
    if (p.count == 0) {
        *done = true;
        return DODRT_OK;
    }
    // donation pays when the whole job is only a few batches per warp (see resolve_variant)
    p.variant = resolve_variant(s->variant, s->cfg[kDonateVariant][kModeFrame], p.count + p.shadow_count, frame->tile_stride > 1,
                                s->dev.num_nodes >= kBigTreeNodes);
    if (p.variant != kDefaultVariant && p.variant != kDonateVariant) return DODRT_OK;
    p.hits = d_hits;
    p.visible = d_visible;
    p.visible_light_stride = visStride;
    p.xs = d_xs;
    p.ys = d_ys;
    p.num_lights = numLights;
    for (uint32_t l = 0; l < numLights; l++) {
        for (int k = 0; k < 3; k++) p.lights[l][k] = lights[3 * l + k];
    }
    if (mirror) {
        p.mirror_hits = mirror->hits;
        p.mirror_visible = numLights ? mirror->visible : nullptr;
        p.mirror_light_stride = mirror->lightStride;
        p.mirror_by_pixel = mirror->byPixel;
    }
    int rc = ensurePool(s);
    if (rc != DODRT_OK) return rc;
    // one stream-ordered block: counters | tile_order | (tile queues only: tile_done | ready_queue); one memset
    static const bool orderTiles = [] { const char *e = std::getenv("DODRT_TILE_ORDER"); return !e || std::atoi(e) != 0; }();
    const bool order = orderTiles && (frame->classes & DODRT_CLS_TREE) && s->dev.num_nodes != 0 && tiles >= 64;
    const char *qEnv = std::getenv("DODRT_FRAME_QUEUES"); // A/B: 1 = per-tile ready queues instead of block-fused batches
    const bool queues = qEnv && std::atoi(qEnv) != 0 && numLights != 0;
    const size_t counterBytes = sizeof(unsigned long long) * kCounterWords;
    const size_t tileBytes = sizeof(uint32_t) * (size_t)tiles;
    char *block = nullptr;
    CUDA_TRY(cudaMallocFromPoolAsync(&block, counterBytes + tileBytes * 3, s->pool, stream));
    cudaError_t e = cudaMemsetAsync(block, 0, counterBytes + (queues ? 3 : 0) * tileBytes, stream);
    p.counter = reinterpret_cast<unsigned long long *>(block);
    p.tile_order = order ? reinterpret_cast<uint32_t *>(block + counterBytes) : nullptr;
    p.tile_done = queues ? reinterpret_cast<uint32_t *>(block + counterBytes) + tiles : nullptr;
    p.ready_queue = queues ? p.tile_done + tiles : nullptr;
    p.num_local_tiles = tiles;
    if (e == cudaSuccess) e = launchTraceOn(s, kModeFrame, p, stream);
    cudaFreeAsync(block, stream);
    if (e != cudaSuccess) return fail(DODRT_E_CUDA, "fused frame launch failed: %s", cudaGetErrorString(e));
    s->launches.fetch_add(order ? 2 : 1);
    *done = true;
    return DODRT_OK;
}

// The whole frame share: fused launch when the variant has one, else primary pass + one shadow pass per light.
int ensureChunkStreams(dodrt_scene *s)
{
    std::lock_guard<std::mutex> lock(s->mutex);
    if (s->chunkEvent[0]) return DODRT_OK;
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    for (cudaStream_t &cs : s->chunkStream) CUDA_TRY(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    for (int i = dodrt_scene::kChunkEvents - 1; i >= 0; i--) { // [0] last: it doubles as the "all created" flag
        CUDA_TRY(cudaEventCreateWithFlags(&s->chunkEvent[i], cudaEventDisableTiming));
    }
    return DODRT_OK;
}

cudaEvent_t nextChunkEvent(dodrt_scene *s) { return s->chunkEvent[s->nextChunkEvent.fetch_add(1) % dodrt_scene::kChunkEvents]; }

// In how many chunks is this frame share traced?  A/B knob (DODRT_FRAME_CHUNKS, default 1).  The idea: a short share (a rank
// of a split frame) ends each of its passes with a tail in which a few warps finish long rays while most SMs idle (about
// 160 us per pass on a 1-of-8 share of a 4K frame); cut into halves whose passes run on two streams, the tail of one half's
// pass could overlap the main phase of the other's next pass -- idle warps leave the persistent kernel (helper limit), their
// SM slots go to the other stream's kernel, and no kernel ever waits for another one.  MEASURED (tests/tools/chunk_sweep.sh,
// 1-of-8 share of dragon4k): 1 chunk 0.714 ms, 2 chunks 0.855, 4 chunks 1.063 (N=1: 3.82 -> 4.27): every kernel is a
// persistent grid sized for the whole GPU, so two of them in flight halve each other's resident warps and mix closest-hit and
// any-hit programs on an SM, which costs more than the overlapped tails give (same finding as the one-launch frame kernels,
// profiles/r02_frame_kernel_ab.txt).
uint32_t frameChunks(const dodrt_frame *frame, uint32_t tiles)
{
    const char *e = std::getenv("DODRT_FRAME_CHUNKS"); // read per call (A/B)
    uint32_t chunks = e ? (uint32_t)std::max(1, std::atoi(e)) : 1u;
    (void)frame;
    chunks = std::min<uint32_t>(chunks, dodrt_scene::kMaxChunks);
    while (chunks > 1 && tiles / chunks < 128) chunks--; // a chunk is still a few batches per warp
    return chunks;
}

// The whole frame share: primary pass + one shadow pass per light (experiments build, DODRT_FUSED=1: one launch).
int traceFrameOnDevice(dodrt_scene *s, const dodrt_frame *frame, const float *d_xs, const float *d_ys, const float *lights,
                       uint32_t numLights, dodrt_hit *d_hits, uint8_t *d_visible, uint64_t visStride, const Mirror *mirror,
                       cudaStream_t stream)
{
    bool done = false;
    int rc = launchFrameFused(s, frame, d_xs, d_ys, lights, numLights, d_hits, d_visible, visStride, mirror, stream, &done);
    if (rc != DODRT_OK || done) return rc;
    uint32_t tilesX, tiles;
    frameTiles(frame, &tilesX, &tiles);
    const uint32_t chunks = frameChunks(frame, tiles);
    if (chunks > 1) {
        rc = ensureChunkStreams(s);
        if (rc != DODRT_OK) return rc;
        cudaEvent_t fork = nextChunkEvent(s);
        CUDA_TRY(cudaEventRecord(fork, stream));
        for (uint32_t c = 1; c < chunks; c++) CUDA_TRY(cudaStreamWaitEvent(s->chunkStream[c - 1], fork, 0));
    }
    for (uint32_t c = 0; c < chunks && rc == DODRT_OK; c++) {
        cudaStream_t st = c == 0 ? stream : s->chunkStream[c - 1];
        const uint32_t begin = (uint32_t)((uint64_t)tiles * c / chunks), end = (uint32_t)((uint64_t)tiles * (c + 1) / chunks);
        rc = launchFrame(s, kModePrimary, frame, d_xs, d_ys, d_hits, nullptr, nullptr, st, begin, end - begin, mirror);
        for (uint32_t l = 0; l < numLights && rc == DODRT_OK; l++) {
            Mirror m;
            if (mirror) {
                m = *mirror;
                m.visible = mirror->visible ? mirror->visible + (uint64_t)l * mirror->lightStride : nullptr;
            }
            rc = launchFrame(s, kModeShadow, frame, d_xs, d_ys, d_hits, lights + 3 * l, d_visible + (uint64_t)l * visStride, st, begin,
                             end - begin, mirror ? &m : nullptr);
        }
    }
    for (uint32_t c = 1; c < chunks; c++) { // join: the caller's stream continues when every chunk is done
        cudaEvent_t joined = nextChunkEvent(s);
        CUDA_TRY(cudaEventRecord(joined, s->chunkStream[c - 1]));
        CUDA_TRY(cudaStreamWaitEvent(stream, joined, 0));
    }
    return rc;
}

// Is [ptr, ptr + bytes) pinned host memory the scene's GPU can write to?  Returns its device-side address then.
void *mappedHostPointer(const void *ptr, size_t bytes)
{
    if (!ptr || !bytes) return nullptr;
    cudaPointerAttributes a{}, b{};
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess ||
        cudaPointerGetAttributes(&b, static_cast<const char *>(ptr) + bytes - 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a.type != cudaMemoryTypeHost || b.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return a.devicePointer;
}

// grow-only staging of the host-buffer entry points (dodrt_scene::hostMutex held)
template <typename T> cudaError_t growStaging(T *&ptr, size_t &have, size_t need)
{
    if (need <= have) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    have = 0;
    const size_t bytes = need + need / 8; // head room: a slightly larger frame does not re-allocate
    cudaError_t e = cudaMalloc(&ptr, bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

// `to` waits for what `from` has queued so far; events are pre-created and recycled call after call
// The next event of the scene's per-call pool (dodrt_scene::stEventsUsed is reset at the start of a host-buffer call).
cudaError_t stagingEvent(dodrt_scene *s, cudaEvent_t *ev)
{
    if (s->stEventsUsed == s->stEvents.size()) {
        cudaEvent_t fresh = nullptr;
        cudaError_t e = cudaEventCreateWithFlags(&fresh, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        s->stEvents.push_back(fresh);
    }
    *ev = s->stEvents[s->stEventsUsed++];
    return cudaSuccess;
}

cudaError_t stagingFence(dodrt_scene *s, cudaStream_t from, cudaStream_t to)
{
    cudaEvent_t ev = nullptr;
    cudaError_t e = stagingEvent(s, &ev);
    if (e == cudaSuccess) e = cudaEventRecord(ev, from);
    return e == cudaSuccess ? cudaStreamWaitEvent(to, ev, 0) : e;
}

} // namespace

extern "C" {

int dodrt_abi_version(void) { return DODRT_ABI_VERSION; }

int dodrt_kernel_variant_available(int variant) { return (variant < 0 || variant_compiled(variant)) ? 1 : 0; }

const char *dodrt_last_error(void) { return g_lastError.c_str(); }

int dodrt_device_count(int *count)
try {
    if (!count) return fail(DODRT_E_INVALID, "count is NULL");
    *count = 0;
    CUDA_TRY(cudaGetDeviceCount(count));
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_scene_create(int device, dodrt_scene **scene)
try {
    if (!scene) return fail(DODRT_E_INVALID, "scene is NULL");
    *scene = nullptr;
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(DODRT_E_INVALID, "device %d out of range (%d devices)", device, count);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", device);
    dodrt_scene *s = new (std::nothrow) dodrt_scene();
    if (!s) return fail(DODRT_E_NOMEM, "out of host memory");
    s->device = device;
    s->dev.epsilon = 0.0001f; // Config::Epsilon default, config.h:9
    s->dev.negzero2 = 0x8000000080000000ull;
    {
        const char *t = std::getenv("DODRT_TUNE"); // "num,den,maxNodeRun" (exploration knob)
        unsigned a = 3, b = 1, c = 0xFFFFFFFFu; // measured best on dragon4k (profiles/r01_vote_rule_sweep.txt)
        if (t) std::sscanf(t, "%u,%u,%u", &a, &b, &c);
        s->dev.tune[0] = a, s->dev.tune[1] = b, s->dev.tune[2] = c;
        // test knob for variant 7: suspend rays at every poll, helpers or not (exercises the resume path everywhere)
        const char *burst = std::getenv("DODRT_NODE_BURST");
        // measured on dragon4k (profiles/r01_node_burst.txt): 1 -> 3779, 2 -> 4003, 4 -> 4115, 8 -> 4203, 64 -> 4286 Mrays/s
        s->dev.node_burst = burst ? (uint32_t)std::max(1, std::atoi(burst)) : 0xFFFFFFFFu;
        const char *poll = std::getenv("DODRT_DONATE_POLL");
        s->dev.donate_poll = poll ? (uint32_t)std::max(1, std::atoi(poll)) : 32u;
        const char *refill = std::getenv("DODRT_POOL_REFILL");
        s->dev.pool_refill = refill ? (uint32_t)std::min(32, std::max(1, std::atoi(refill))) : 32u;
        const char *fork = std::getenv("DODRT_FORK_POLL");
        // measured on a 1-of-8 share of dragon4k (profiles/r02_donation_fork.txt): work splitting of resumed any-hit rays
        // loses (0.71 -> 0.74-0.79 ms: the forks take the waiting helpers away from the donors), so it is off; 512 waiting
        // helpers serve the donors as well as all ~2900 idle warps do, and poll six times less
        s->dev.fork_poll = fork ? (uint32_t)std::max(0, std::atoi(fork)) : 0u;
        const char *limit = std::getenv("DODRT_HELPER_LIMIT");
        s->dev.helper_limit = limit ? (uint32_t)std::max(1, std::atoi(limit)) : 512u;
        const char *always = std::getenv("DODRT_DONATE_ALWAYS");
        s->dev.tune[3] = (always && std::atoi(always) != 0) ? 1u : 0u;
    }
    cudaError_t e = cudaMalloc(&s->d_counters, sizeof(unsigned long long) * kCounterSlots * kCounterWords);
    s->variant = default_variant();
    for (int v = 0; v < kNumVariants; v++) {
        for (int m = 0; m < kNumModes && e == cudaSuccess; m++) {
            if (variant_compiled(v) && mode_compiled(m)) e = trace_launch_config(device, (TraceMode)m, v, &s->cfg[v][m]);
        }
    }
    if (e != cudaSuccess) {
        cudaFree(s->d_counters);
        delete s;
        return fail(DODRT_E_CUDA, "scene_create: %s", cudaGetErrorString(e));
    }
    *scene = s;
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_scene_destroy(dodrt_scene *s)
try {
    if (!s) return DODRT_OK;
    DeviceGuard guard(s->device);
    cudaDeviceSynchronize();
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->copyStream) cudaStreamDestroy(s->copyStream);
    for (cudaStream_t cs : s->chunkStream) {
        if (cs) cudaStreamDestroy(cs);
    }
    for (cudaEvent_t ev : s->chunkEvent) {
        if (ev) cudaEventDestroy(ev);
    }
    if (s->pool) cudaMemPoolDestroy(s->pool);
    for (cudaEvent_t ev : s->stEvents) cudaEventDestroy(ev);
    if (s->stTables) cudaFree(s->stTables);
    if (s->stHits) cudaFree(s->stHits);
    if (s->stVis) cudaFree(s->stVis);
    for (int q = 0; q < dodrt_scene::kDonateQueues; q++) {
        if (s->donateDone[q]) cudaEventDestroy(s->donateDone[q]);
        if (s->donateQueue[q]) cudaFree(s->donateQueue[q]);
    }
    freeDevice(s->d_nodes);
    freeDevice(s->d_tris);
    freeDevice(s->d_lanes4);
    freeDevice(s->d_spheres);
    freeDevice(s->d_planes);
    freeDevice(s->d_cylinders);
    freeDevice(s->d_boxes);
    freeDevice(s->d_sphereBvh);
    freeDevice(s->d_sphereBvhIds);
    freeDevice(s->d_boxBvh);
    freeDevice(s->d_boxBvhIds);
    freeDevice(s->d_triAttrs);
    freeDevice(s->d_meshColors);
    freeDevice(s->d_sphereColors);
    freeDevice(s->d_planeColors);
    freeDevice(s->d_counters);
    freeDevice(s->d_stats);
    delete s;
    return DODRT_OK;
}
DODRT_CATCH

// Shared body of dodrt_scene_set_kdtree (prim_nums == nullptr: `tri_lanes` holds num_tri_lanes re-ordered lanes) and
// dodrt_scene_set_kdtree_indexed (`tri_lanes` holds num_src_lanes lanes in creation order and lane i of the tree is
// tri_lanes[prim_nums[i]]: Triangle::reorderLanesByIndices, triangle.cpp:349-367, done by the repack kernel's gather).
static int ensureTris(dodrt_scene *s);
static int setKdtree(dodrt_scene *s, const uint64_t *nodes, uint32_t num_nodes, const float *tri_lanes, uint32_t num_src_lanes,
                     const uint32_t *prim_nums, uint32_t num_tri_lanes, const float bounds[6])
{
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    if ((num_nodes && !nodes) || (num_tri_lanes && !tri_lanes) || !bounds) {
        return fail(DODRT_E_INVALID, "NULL array with non-zero count");
    }
    if (prim_nums) {
        for (uint32_t i = 0; i < num_tri_lanes; i++) {
            if (prim_nums[i] >= num_src_lanes) {
                return fail(DODRT_E_INVALID, "prim_nums[%u] = %u out of range (%u lanes)", i, prim_nums[i], num_src_lanes);
            }
        }
    }
    if ((uint64_t)num_tri_lanes * kLane >= (1ull << DODRT_KIND_SHIFT)) {
        return fail(DODRT_E_LIMIT, "%u lanes exceed the %u-bit triangle id space", num_tri_lanes, DODRT_KIND_SHIFT);
    }
    uint32_t depth = 0;
    int rc = validateTree(nodes, num_nodes, num_tri_lanes, &depth);
    if (rc != DODRT_OK) return rc;
    if (depth > (uint32_t)kMaxStack) {
        return fail(DODRT_E_LIMIT, "kd-tree depth %u exceeds the traversal stack (%d)", depth, kMaxStack);
    }
    std::lock_guard<std::mutex> lock(s->mutex);
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    freeDevice(s->d_nodes);
    freeDevice(s->d_tris);
    freeDevice(s->d_lanes4);
    s->dev.nodes = nullptr;
    s->dev.tris = nullptr;
    s->dev.lanes4 = nullptr;
    s->dev.num_nodes = 0;
    s->dev.num_tri_lanes = 0;
    if (num_nodes) {
        CUDA_TRY(cudaMalloc(&s->d_nodes, (size_t)num_nodes * sizeof(uint2)));
        CUDA_TRY(cudaMemcpy(s->d_nodes, nodes, (size_t)num_nodes * sizeof(uint2), cudaMemcpyHostToDevice));
    }
    if (num_tri_lanes) {
        float *d_lanes = nullptr;
        uint32_t *d_prim = nullptr;
        const size_t laneBytes = (size_t)num_tri_lanes * 288;
        const size_t srcBytes = (size_t)(prim_nums ? num_src_lanes : num_tri_lanes) * 288;
        CUDA_TRY(cudaMalloc(&d_lanes, srcBytes));
        cudaError_t e = cudaMemcpy(d_lanes, tri_lanes, srcBytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && prim_nums) e = cudaMalloc(&d_prim, (size_t)num_tri_lanes * sizeof(uint32_t));
        if (e == cudaSuccess && prim_nums) {
            e = cudaMemcpy(d_prim, prim_nums, (size_t)num_tri_lanes * sizeof(uint32_t), cudaMemcpyHostToDevice);
        }
        if (e == cudaSuccess) e = cudaMalloc(&s->d_lanes4, laneBytes);
        if (e == cudaSuccess) e = launch_repack_triangles(d_lanes, d_prim, num_tri_lanes, s->d_tris, s->d_lanes4, nullptr);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaFree(d_lanes);
        cudaFree(d_prim);
        if (e != cudaSuccess) return fail(DODRT_E_CUDA, "triangle upload: %s", cudaGetErrorString(e));
        s->launches.fetch_add(1);
    }
    s->dev.nodes = s->d_nodes;
    s->dev.tris = nullptr; // per-triangle records: built on demand by ensureTris (variants 0-2 only)
    s->dev.lanes4 = s->d_lanes4;
    s->dev.num_nodes = num_nodes;
    s->dev.num_tri_lanes = num_tri_lanes;
    for (int i = 0; i < 3; i++) {
        s->dev.bmin[i] = bounds[i];
        s->dev.bmax[i] = bounds[3 + i];
    }
    s->treeDepth = depth;
    if (s->variant >= 0 && s->variant < 3) return ensureTris(s); // DODRT_VARIANT / an earlier set_kernel_variant
    return DODRT_OK;
}

int dodrt_scene_set_kdtree(dodrt_scene *s, const uint64_t *nodes, uint32_t num_nodes, const float *tri_lanes,
                           uint32_t num_tri_lanes, const float bounds[6])
try {
    return setKdtree(s, nodes, num_nodes, tri_lanes, num_tri_lanes, nullptr, num_tri_lanes, bounds);
}
DODRT_CATCH

int dodrt_scene_set_kdtree_indexed(dodrt_scene *s, const uint64_t *nodes, uint32_t num_nodes, const float *tri_lanes,
                                   uint32_t num_src_lanes, const uint32_t *prim_nums, uint32_t num_tri_lanes,
                                   const float bounds[6])
try {
    if (num_tri_lanes && !prim_nums) return fail(DODRT_E_INVALID, "prim_nums is NULL");
    return setKdtree(s, nodes, num_nodes, tri_lanes, num_src_lanes, prim_nums, num_tri_lanes, bounds);
}
DODRT_CATCH

static int uploadLanes(dodrt_scene *s, float **slot, const float **view, uint32_t *countField, const float *lanes,
                       uint32_t count, uint32_t floatsPerLane)
{
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    if (count && !lanes) return fail(DODRT_E_INVALID, "NULL lanes with non-zero count");
    std::lock_guard<std::mutex> lock(s->mutex);
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    freeDevice(*slot);
    *view = nullptr;
    *countField = 0;
    if (count) {
        const size_t bytes = (size_t)((count + kLane - 1) / kLane) * floatsPerLane * sizeof(float);
        CUDA_TRY(cudaMalloc(slot, bytes));
        CUDA_TRY(cudaMemcpy(*slot, lanes, bytes, cudaMemcpyHostToDevice));
    }
    *view = *slot;
    *countField = count;
    return DODRT_OK;
}

// Builds and uploads the culling BVH over `count` primitives whose boxes are given as count x {min xyz, max xyz}.
static int uploadPrimBvh(dodrt_scene *s, const std::vector<float> &boxes, uint32_t count, float4 **d_nodes, uint32_t **d_ids,
                         const float4 **nodesView, const uint32_t **idsView)
{
    std::lock_guard<std::mutex> lock(s->mutex);
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    freeDevice(*d_nodes);
    freeDevice(*d_ids);
    *nodesView = nullptr;
    *idsView = nullptr;
    static const bool enabled = [] { const char *e = std::getenv("DODRT_PRIM_BVH"); return !e || std::atoi(e) != 0; }();
    if (!enabled || count < kPrimBvhMinCount) return DODRT_OK;
    for (float v : boxes) { // NaN / inf primitives cannot be boxed (and would break nth_element's ordering): brute force then
        if (!std::isfinite(v)) return DODRT_OK;
    }
    std::vector<PrimBvhNode> nodes;
    std::vector<uint32_t> ids;
    build_prim_bvh(boxes.data(), count, nodes, ids);
    CUDA_TRY(cudaMalloc(d_nodes, nodes.size() * sizeof(PrimBvhNode)));
    CUDA_TRY(cudaMemcpy(*d_nodes, nodes.data(), nodes.size() * sizeof(PrimBvhNode), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(d_ids, ids.size() * sizeof(uint32_t)));
    CUDA_TRY(cudaMemcpy(*d_ids, ids.data(), ids.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    *nodesView = *d_nodes;
    *idsView = *d_ids;
    return DODRT_OK;
}

int dodrt_scene_set_spheres(dodrt_scene *s, const float *sphere_lanes, uint32_t num_spheres)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    int rc = uploadLanes(s, &s->d_spheres, &s->dev.sphere_lanes, &s->dev.num_spheres, sphere_lanes, num_spheres, 4 * kLane);
    if (rc != DODRT_OK) return rc;
    std::vector<float> boxes((size_t)num_spheres * 6);
    for (uint32_t i = 0; i < num_spheres; i++) { // box of sphere i: centre +- radius (radius from radiusSq, rounded up)
        const float *lane = sphere_lanes + (size_t)(i / kLane) * 4 * kLane;
        const uint32_t j = i % kLane;
        const float r = std::sqrt(lane[3 * kLane + j]) * 1.000001f;
        for (int k = 0; k < 3; k++) {
            boxes[(size_t)i * 6 + k] = lane[k * kLane + j] - r;
            boxes[(size_t)i * 6 + 3 + k] = lane[k * kLane + j] + r;
        }
    }
    return uploadPrimBvh(s, boxes, num_spheres, &s->d_sphereBvh, &s->d_sphereBvhIds, &s->dev.sphere_bvh, &s->dev.sphere_bvh_ids);
}
DODRT_CATCH

int dodrt_scene_set_planes(dodrt_scene *s, const float *plane_lanes, uint32_t num_planes)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    return uploadLanes(s, &s->d_planes, &s->dev.plane_lanes, &s->dev.num_planes, plane_lanes, num_planes, 6 * kLane);
}
DODRT_CATCH

int dodrt_scene_set_boxes(dodrt_scene *s, const float *box_lanes, uint32_t num_boxes)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    int rc = uploadLanes(s, &s->d_boxes, &s->dev.box_lanes, &s->dev.num_boxes, box_lanes, num_boxes, 6 * kLane);
    if (rc != DODRT_OK) return rc;
    std::vector<float> boxes((size_t)num_boxes * 6);
    for (uint32_t i = 0; i < num_boxes; i++) {
        const float *lane = box_lanes + (size_t)(i / kLane) * 6 * kLane;
        const uint32_t j = i % kLane;
        for (int k = 0; k < 3; k++) { // tolerate inverted boxes (min > max): the slab test swaps per axis anyway
            const float a = lane[k * kLane + j], b = lane[(3 + k) * kLane + j];
            boxes[(size_t)i * 6 + k] = a < b ? a : b;
            boxes[(size_t)i * 6 + 3 + k] = a < b ? b : a;
        }
    }
    return uploadPrimBvh(s, boxes, num_boxes, &s->d_boxBvh, &s->d_boxBvhIds, &s->dev.box_bvh, &s->dev.box_bvh_ids);
}
DODRT_CATCH

int dodrt_scene_set_cylinders(dodrt_scene *s, const dodrt_cylinder *cylinders, uint32_t num_cylinders)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    if (num_cylinders && !cylinders) return fail(DODRT_E_INVALID, "NULL cylinders with non-zero count");
    std::lock_guard<std::mutex> lock(s->mutex);
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    freeDevice(s->d_cylinders);
    s->dev.cylinders = nullptr;
    s->dev.num_cylinders = 0;
    if (num_cylinders) {
        CUDA_TRY(cudaMalloc(&s->d_cylinders, sizeof(dodrt_cylinder) * num_cylinders));
        CUDA_TRY(cudaMemcpy(s->d_cylinders, cylinders, sizeof(dodrt_cylinder) * num_cylinders, cudaMemcpyHostToDevice));
    }
    s->dev.cylinders = s->d_cylinders;
    s->dev.num_cylinders = num_cylinders;
    return DODRT_OK;
}
DODRT_CATCH

static int setShading(dodrt_scene *s, const void *tri_attributes, uint32_t num_src_lanes, const uint32_t *prim_nums,
                      uint32_t num_tri_lanes, const float *mesh_colors, uint32_t num_meshes, const float *sphere_colors,
                      const float *plane_colors)
{
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    if (num_tri_lanes != s->dev.num_tri_lanes) {
        return fail(DODRT_E_INVALID, "attributes for %u lanes but the kd-tree has %u", num_tri_lanes, s->dev.num_tri_lanes);
    }
    if (num_tri_lanes && (!tri_attributes || !mesh_colors || !num_meshes)) return fail(DODRT_E_INVALID, "NULL attributes");
    if ((s->dev.num_spheres && !sphere_colors) || (s->dev.num_planes && !plane_colors)) {
        return fail(DODRT_E_INVALID, "NULL sphere/plane colours");
    }
    if (prim_nums) {
        for (uint32_t i = 0; i < num_tri_lanes; i++) {
            if (prim_nums[i] >= num_src_lanes) {
                return fail(DODRT_E_INVALID, "prim_nums[%u] = %u out of range (%u lanes)", i, prim_nums[i], num_src_lanes);
            }
        }
    }
    if (num_tri_lanes) { // every meshAttrIdx must address a mesh colour
        const uint32_t *a = static_cast<const uint32_t *>(tri_attributes);
        for (uint32_t lane = 0; lane < num_src_lanes; lane++) {
            for (int j = 0; j < kLane; j++) {
                if (a[(size_t)lane * 80 + j] >= num_meshes) {
                    return fail(DODRT_E_INVALID, "meshAttrIdx %u of lane %u out of range (%u meshes)",
                                a[(size_t)lane * 80 + j], lane, num_meshes);
                }
            }
        }
    }
    std::lock_guard<std::mutex> lock(s->mutex);
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    freeDevice(s->d_triAttrs);
    freeDevice(s->d_meshColors);
    freeDevice(s->d_sphereColors);
    freeDevice(s->d_planeColors);
    auto upload = [&](auto **dst, const void *src, size_t bytes) -> cudaError_t {
        if (!bytes) return cudaSuccess;
        cudaError_t e = cudaMalloc(dst, bytes);
        return e == cudaSuccess ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : e;
    };
    if (prim_nums && num_tri_lanes) { // Triangle::reorderLanesByIndices (triangle.cpp:358-364) for the attributes, on the GPU
        uint32_t *d_src = nullptr, *d_prim = nullptr;
        cudaError_t e = upload(&d_src, tri_attributes, (size_t)num_src_lanes * 320);
        if (e == cudaSuccess) e = upload(&d_prim, prim_nums, (size_t)num_tri_lanes * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&s->d_triAttrs, (size_t)num_tri_lanes * 320);
        if (e == cudaSuccess) e = launch_gather_lanes(d_src, d_prim, num_tri_lanes, 80, s->d_triAttrs, nullptr);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaFree(d_src);
        cudaFree(d_prim);
        if (e != cudaSuccess) return fail(DODRT_E_CUDA, "attribute upload: %s", cudaGetErrorString(e));
        s->launches.fetch_add(1);
    } else {
        CUDA_TRY(upload(&s->d_triAttrs, tri_attributes, (size_t)num_tri_lanes * 320));
    }
    CUDA_TRY(upload(&s->d_meshColors, mesh_colors, (size_t)num_meshes * 12));
    CUDA_TRY(upload(&s->d_sphereColors, sphere_colors, (size_t)s->dev.num_spheres * 12));
    CUDA_TRY(upload(&s->d_planeColors, plane_colors, (size_t)s->dev.num_planes * 12));
    s->dev.tri_attrs = s->d_triAttrs;
    s->dev.mesh_colors = s->d_meshColors;
    s->dev.sphere_colors = s->d_sphereColors;
    s->dev.plane_colors = s->d_planeColors;
    return DODRT_OK;
}

int dodrt_scene_set_shading(dodrt_scene *s, const void *tri_attributes, uint32_t num_tri_lanes, const float *mesh_colors,
                            uint32_t num_meshes, const float *sphere_colors, const float *plane_colors)
try {
    return setShading(s, tri_attributes, num_tri_lanes, nullptr, num_tri_lanes, mesh_colors, num_meshes, sphere_colors,
                      plane_colors);
}
DODRT_CATCH

int dodrt_scene_set_shading_indexed(dodrt_scene *s, const void *tri_attributes, uint32_t num_src_lanes,
                                    const uint32_t *prim_nums, uint32_t num_tri_lanes, const float *mesh_colors,
                                    uint32_t num_meshes, const float *sphere_colors, const float *plane_colors)
try {
    if (num_tri_lanes && !prim_nums) return fail(DODRT_E_INVALID, "prim_nums is NULL");
    return setShading(s, tri_attributes, num_src_lanes, prim_nums, num_tri_lanes, mesh_colors, num_meshes, sphere_colors,
                      plane_colors);
}
DODRT_CATCH

// Variants 0-2 read one 48-B record per triangle slot (133 MB for the 871k-triangle mesh, 2 GB for config 5); the
// default kernels never do, so the array only exists once such a variant has been asked for.
static int ensureTris(dodrt_scene *s)
{
    if (s->d_tris || s->dev.num_tri_lanes == 0) return DODRT_OK;
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    CUDA_TRY(cudaMalloc(&s->d_tris, (size_t)s->dev.num_tri_lanes * kLane * 3 * sizeof(float4)));
    cudaError_t e = launch_tris_from_lanes4(s->d_lanes4, s->dev.num_tri_lanes, s->d_tris, nullptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(DODRT_E_CUDA, "triangle records: %s", cudaGetErrorString(e));
    s->dev.tris = s->d_tris;
    s->launches.fetch_add(1);
    return DODRT_OK;
}

int dodrt_scene_set_kernel_variant(dodrt_scene *s, int variant)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    if (variant >= kNumVariants) return fail(DODRT_E_INVALID, "kernel variant %d out of range [0,%d)", variant, kNumVariants);
    if (variant >= 0 && !variant_compiled(variant)) {
        return fail(DODRT_E_INVALID, "kernel variant %d is an experiment: only in the -DDODRT_EXPERIMENTS build (libdodrt_cuda_exp.so)", variant);
    }
    std::lock_guard<std::mutex> lock(s->mutex);
    s->variant = variant < 0 ? default_variant() : variant;
    if (s->variant >= 0 && s->variant < 3) return ensureTris(s);
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_scene_set_epsilon(dodrt_scene *s, float epsilon)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    std::lock_guard<std::mutex> lock(s->mutex);
    s->dev.epsilon = epsilon;
    return DODRT_OK;
}
DODRT_CATCH

// ---- device-resident entry points ---------------------------------------------------------------

int dodrt_intersect_device(dodrt_scene *s, const dodrt_ray *d_rays, uint64_t num_rays, uint32_t classes,
                           dodrt_hit *d_hits, void *stream)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    if (num_rays == 0) return DODRT_OK;
    if (!d_rays || !d_hits) return fail(DODRT_E_INVALID, "NULL ray/hit buffer");
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    TraceParams p{};
    p.scene = s->dev;
    p.classes = classes;
    p.rays = d_rays;
    p.count = num_rays;
    p.hits = d_hits;
    p.counter = nextCounter(s);
    p.variant = resolve_variant(s->variant, s->cfg[kDonateVariant][kModeRays], p.count, false, s->dev.num_nodes >= kBigTreeNodes);
    p.tile_order = nullptr;
    p.num_local_tiles = 0;
    if (p.variant == kDonateVariant) {
        int rc = ensurePool(s);
        if (rc != DODRT_OK) return rc;
    }
    CUDA_TRY(launchTraceOn(s, kModeRays, p, static_cast<cudaStream_t>(stream)));
    s->launches.fetch_add(1);
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_trace_primary_device(dodrt_scene *s, const dodrt_frame *frame, const float *d_xs, const float *d_ys,
                               dodrt_hit *d_hits, void *stream)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    int rc = checkFrame(frame);
    if (rc != DODRT_OK) return rc;
    if (!d_xs || !d_ys || !d_hits) return fail(DODRT_E_INVALID, "NULL table/hit buffer");
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    return launchFrame(s, kModePrimary, frame, d_xs, d_ys, d_hits, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}
DODRT_CATCH

int dodrt_trace_shadow_device(dodrt_scene *s, const dodrt_frame *frame, const float *d_xs, const float *d_ys,
                              const dodrt_hit *d_hits, const float light[3], uint8_t *d_visible, void *stream)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    int rc = checkFrame(frame);
    if (rc != DODRT_OK) return rc;
    if (!d_xs || !d_ys || !d_hits || !light || !d_visible) return fail(DODRT_E_INVALID, "NULL argument");
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    return launchFrame(s, kModeShadow, frame, d_xs, d_ys, const_cast<dodrt_hit *>(d_hits), light, d_visible,
                       static_cast<cudaStream_t>(stream));
}
DODRT_CATCH

int dodrt_frame_assemble_device(dodrt_scene *s, const dodrt_frame *frame, const dodrt_hit *d_compact_hits,
                                const uint8_t *d_compact_visible, uint64_t slots_per_rank, dodrt_hit *d_hits_out,
                                uint8_t *d_visible_out, void *stream)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    int rc = checkFrame(frame);
    if (rc != DODRT_OK) return rc;
    if (!d_compact_hits || !d_hits_out) return fail(DODRT_E_INVALID, "NULL hit buffer");
    if ((d_visible_out != nullptr) != (d_compact_visible != nullptr)) {
        return fail(DODRT_E_INVALID, "visibility input and output must both be given or both be NULL");
    }
    dodrt_frame rank0 = *frame;
    rank0.first_tile = 0;
    if (slots_per_rank < frameSlots(&rank0)) {
        return fail(DODRT_E_INVALID, "slots_per_rank %llu is smaller than rank 0's %llu slots",
                    (unsigned long long)slots_per_rank, (unsigned long long)frameSlots(&rank0));
    }
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    uint32_t tilesX, localTiles;
    frameTiles(frame, &tilesX, &localTiles);
    CUDA_TRY(launch_assemble(*frame, tilesX, d_compact_hits, d_compact_visible, slots_per_rank, d_hits_out,
                             d_visible_out, static_cast<cudaStream_t>(stream)));
    s->launches.fetch_add(1);
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_trace_frame_device(dodrt_scene *s, const dodrt_frame *frame, const float *d_xs, const float *d_ys, const float *lights,
                             uint32_t num_lights, dodrt_hit *d_hits, uint8_t *d_visible, dodrt_frame_buffer *mirror, void *stream)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    int rc = checkFrame(frame);
    if (rc != DODRT_OK) return rc;
    if (!d_xs || !d_ys || !d_hits) return fail(DODRT_E_INVALID, "NULL table/hit buffer");
    if (num_lights && (!lights || !d_visible)) return fail(DODRT_E_INVALID, "NULL lights/visible buffer");
    if (num_lights > (uint32_t)kMaxLights) return fail(DODRT_E_LIMIT, "%u lights exceed the limit of %d", num_lights, kMaxLights);
    Mirror m;
    if (mirror) {
        if (mirror->device != s->device) {
            return fail(DODRT_E_INVALID, "frame buffer view belongs to device %d, the scene to device %d (open / attach it for this scene)",
                        mirror->device, s->device);
        }
        if (mirror->width != frame->width || mirror->height != frame->height || mirror->numLights < num_lights) {
            return fail(DODRT_E_INVALID, "frame buffer is %ux%u with %u lights, the frame %ux%u with %u", mirror->width,
                        mirror->height, mirror->numLights, frame->width, frame->height, num_lights);
        }
        m.hits = mirror->hits();
        m.visible = mirror->visible();
        m.lightStride = (uint64_t)mirror->width * mirror->height;
        m.byPixel = 1;
    }
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    const uint64_t slots = frame->compact ? frameSlots(frame) : (uint64_t)frame->width * frame->height;
    return traceFrameOnDevice(s, frame, d_xs, d_ys, lights, num_lights, d_hits, d_visible, slots, mirror ? &m : nullptr,
                              static_cast<cudaStream_t>(stream));
}
DODRT_CATCH

// ---- frame buffers other GPUs write into -------------------------------------------------------------

int dodrt_frame_buffer_create(dodrt_scene *owner, uint32_t width, uint32_t height, uint32_t num_lights, dodrt_frame_buffer **fb)
try {
    if (!owner || !fb) return fail(DODRT_E_INVALID, "NULL argument");
    *fb = nullptr;
    if (!width || !height || (uint64_t)width * height >= 0xFFFFFFFFull) return fail(DODRT_E_INVALID, "bad frame size %ux%u", width, height);
    if (num_lights > (uint32_t)kMaxLights) return fail(DODRT_E_LIMIT, "%u lights exceed the limit of %d", num_lights, kMaxLights);
    DeviceGuard guard(owner->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", owner->device);
    std::unique_ptr<dodrt_frame_buffer> f(new dodrt_frame_buffer());
    f->device = f->ownerDevice = owner->device;
    f->owner = true;
    f->width = width, f->height = height, f->numLights = num_lights;
    f->bytes = (size_t)width * height * (sizeof(dodrt_hit) + num_lights);
    // plain cudaMalloc (not a pool): the allocation must be exportable with cudaIpcGetMemHandle
    CUDA_TRY(cudaMalloc(&f->base, f->bytes));
    cudaError_t e = cudaMemset(f->base, 0xFF, (size_t)width * height * sizeof(dodrt_hit)); // "miss" until written
    if (e == cudaSuccess && num_lights) e = cudaMemset(f->visible(), 0, (size_t)width * height * num_lights);
    if (e != cudaSuccess) {
        cudaFree(f->base);
        return fail(DODRT_E_CUDA, "frame buffer: %s", cudaGetErrorString(e));
    }
    *fb = f.release();
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_frame_buffer_export(dodrt_frame_buffer *fb, dodrt_frame_buffer_desc *desc)
try {
    if (!fb || !desc) return fail(DODRT_E_INVALID, "NULL argument");
    if (!fb->owner) return fail(DODRT_E_INVALID, "only the owner of a frame buffer can export it");
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(desc->ipc_handle), "ipc handle does not fit the descriptor");
    std::memset(desc, 0, sizeof(*desc));
    DeviceGuard guard(fb->ownerDevice);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", fb->ownerDevice);
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, fb->base));
    std::memcpy(desc->ipc_handle, &h, sizeof(h));
    desc->width = fb->width, desc->height = fb->height, desc->num_lights = fb->numLights;
    desc->device = (uint32_t)fb->ownerDevice;
    desc->bytes = fb->bytes;
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_frame_buffer_open(dodrt_scene *user, const dodrt_frame_buffer_desc *desc, dodrt_frame_buffer **fb)
try {
    if (!user || !desc || !fb) return fail(DODRT_E_INVALID, "NULL argument");
    *fb = nullptr;
    if ((size_t)desc->width * desc->height * (sizeof(dodrt_hit) + desc->num_lights) != desc->bytes || !desc->bytes) {
        return fail(DODRT_E_INVALID, "inconsistent frame buffer descriptor");
    }
    DeviceGuard guard(user->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", user->device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, desc->ipc_handle, sizeof(h));
    void *base = nullptr;
    // maps the owner's allocation into this process; peer access from the current device is enabled on demand
    CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    dodrt_frame_buffer *f = new (std::nothrow) dodrt_frame_buffer();
    if (!f) {
        cudaIpcCloseMemHandle(base);
        return fail(DODRT_E_NOMEM, "out of host memory");
    }
    f->device = user->device;
    f->ownerDevice = (int)desc->device;
    f->ipc = true;
    f->width = desc->width, f->height = desc->height, f->numLights = desc->num_lights;
    f->bytes = desc->bytes;
    f->base = static_cast<char *>(base);
    *fb = f;
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_frame_buffer_attach(dodrt_scene *user, dodrt_frame_buffer *owner_fb, dodrt_frame_buffer **fb)
try {
    if (!user || !owner_fb || !fb) return fail(DODRT_E_INVALID, "NULL argument");
    *fb = nullptr;
    if (!owner_fb->owner) return fail(DODRT_E_INVALID, "attach needs the frame buffer its owner created");
    if (user->device != owner_fb->ownerDevice) {
        int can = 0;
        CUDA_TRY(cudaDeviceCanAccessPeer(&can, user->device, owner_fb->ownerDevice));
        if (!can) return fail(DODRT_E_CUDA, "device %d cannot address the memory of device %d", user->device, owner_fb->ownerDevice);
        DeviceGuard guard(user->device);
        if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", user->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(owner_fb->ownerDevice, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) {
            cudaGetLastError();
        } else if (e != cudaSuccess) {
            return fail(DODRT_E_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", user->device, owner_fb->ownerDevice, cudaGetErrorString(e));
        }
    }
    dodrt_frame_buffer *f = new (std::nothrow) dodrt_frame_buffer(*owner_fb);
    if (!f) return fail(DODRT_E_NOMEM, "out of host memory");
    f->device = user->device;
    f->owner = false;
    f->ipc = false;
    *fb = f;
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_frame_buffer_pointers(dodrt_frame_buffer *fb, dodrt_hit **d_hits, uint8_t **d_visible)
try {
    if (!fb) return fail(DODRT_E_INVALID, "frame buffer is NULL");
    if (d_hits) *d_hits = fb->hits();
    if (d_visible) *d_visible = fb->numLights ? fb->visible() : nullptr;
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_frame_buffer_destroy(dodrt_frame_buffer *fb)
try {
    if (!fb) return DODRT_OK;
    if (fb->owner || fb->ipc) {
        DeviceGuard guard(fb->owner ? fb->ownerDevice : fb->device);
        cudaDeviceSynchronize();
        if (fb->owner) {
            cudaFree(fb->base);
        } else {
            cudaIpcCloseMemHandle(fb->base);
        }
    }
    delete fb;
    return DODRT_OK;
}
DODRT_CATCH

// ---- host-buffer entry points -------------------------------------------------------------------

int dodrt_intersect(dodrt_scene *s, const dodrt_ray *rays, uint64_t num_rays, uint32_t classes, dodrt_hit *hits)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    if (num_rays == 0) return DODRT_OK;
    if (!rays || !hits) return fail(DODRT_E_INVALID, "NULL ray/hit buffer");
    int rc = ensureStream(s);
    if (rc != DODRT_OK) return rc;
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    dodrt_ray *d_rays = nullptr;
    dodrt_hit *d_hits = nullptr;
    cudaStream_t st = s->stream;
    CUDA_TRY(cudaMallocFromPoolAsync(&d_rays, num_rays * sizeof(dodrt_ray), s->pool, st));
    cudaError_t e = cudaMallocFromPoolAsync(&d_hits, num_rays * sizeof(dodrt_hit), s->pool, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rays, rays, num_rays * sizeof(dodrt_ray), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        rc = dodrt_intersect_device(s, d_rays, num_rays, classes, d_hits, st);
        if (rc == DODRT_OK) {
            e = cudaMemcpyAsync(hits, d_hits, num_rays * sizeof(dodrt_hit), cudaMemcpyDeviceToHost, st);
        }
    }
    if (d_rays) cudaFreeAsync(d_rays, st);
    if (d_hits) cudaFreeAsync(d_hits, st);
    cudaError_t es = cudaStreamSynchronize(st);
    if (rc != DODRT_OK) return rc;
    if (e == cudaSuccess) e = es;
    if (e != cudaSuccess) return fail(DODRT_E_CUDA, "dodrt_intersect: %s", cudaGetErrorString(e));
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_trace_frame(dodrt_scene *s, const dodrt_frame *frame, const float *xs, const float *ys, const float *lights,
                      uint32_t num_lights, dodrt_hit *hits, uint8_t *visible)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    int rc = checkFrame(frame);
    if (rc != DODRT_OK) return rc;
    if (!xs || !ys || !hits) return fail(DODRT_E_INVALID, "NULL table/hit buffer");
    if (num_lights && (!lights || !visible)) return fail(DODRT_E_INVALID, "NULL lights/visible buffer");
    rc = ensureStream(s);
    if (rc != DODRT_OK) return rc;
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    const uint64_t slots = frame->compact ? frameSlots(frame) : (uint64_t)frame->width * frame->height;
    if (slots == 0) return DODRT_OK;
    std::lock_guard<std::mutex> hostLock(s->hostMutex); // one host-buffer call at a time per scene: persistent staging
    cudaStream_t st = s->stream;
    s->stEventsUsed = 0;
    const size_t tableFloats = (size_t)frame->width + frame->height;
    cudaError_t e = growStaging(s->stTables, s->stTablesBytes, tableFloats * sizeof(float));
    if (e == cudaSuccess) e = growStaging(s->stHits, s->stHitsBytes, slots * sizeof(dodrt_hit));
    if (e == cudaSuccess) e = growStaging(s->stVis, s->stVisBytes, slots * (size_t)(num_lights ? num_lights : 1));
    if (e != cudaSuccess) return fail(DODRT_E_CUDA, "dodrt_trace_frame staging: %s", cudaGetErrorString(e));
    float *d_tables = s->stTables;
    dodrt_hit *d_hits = s->stHits;
    uint8_t *d_vis = s->stVis;
    e = cudaMemcpyAsync(d_tables, xs, frame->width * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        e = cudaMemcpyAsync(d_tables + frame->width, ys, frame->height * sizeof(float), cudaMemcpyHostToDevice, st);
    }
    // ---- zero-copy.  Opt-in (DODRT_ZEROCOPY=1, read per call): the kernels store every result into the caller's PINNED buffers themselves
    // (mirror = the host buffers) instead of staging + DMA copies.  Measured on this pool's B200s it loses at every share
    // size -- SM stores over PCIe reach about half the DMA rate and stall the SMs (dragon4k: 5.85 vs 3.99 ms whole frame,
    // 0.96 vs 0.77 ms for a 1-of-8 share; tests/tools/e2e_probe.py) -- so the default is the staged path below.
    const char *zcEnv = std::getenv("DODRT_ZEROCOPY");
    const bool zcOn = zcEnv && std::atoi(zcEnv) != 0;
    void *mHits = zcOn ? mappedHostPointer(hits, slots * sizeof(dodrt_hit)) : nullptr;
    void *mVis = (zcOn && num_lights) ? mappedHostPointer(visible, slots * (size_t)num_lights) : nullptr;
    const bool zeroCopy = mHits && (num_lights == 0 || mVis) && (frame->compact || frame->tile_stride == 1);
    if (e == cudaSuccess && zeroCopy) {
        Mirror m;
        m.hits = static_cast<dodrt_hit *>(mHits);
        m.visible = static_cast<uint8_t *>(mVis);
        m.lightStride = slots;
        m.byPixel = 0;
        rc = traceFrameOnDevice(s, frame, d_tables, d_tables + frame->width, lights, num_lights, d_hits, d_vis, slots, &m, st);
        cudaError_t es = cudaStreamSynchronize(st);
        if (rc != DODRT_OK) return rc;
        if (es != cudaSuccess) return fail(DODRT_E_CUDA, "dodrt_trace_frame: %s", cudaGetErrorString(es));
        return DODRT_OK;
    }
    if (e == cudaSuccess && !frame->compact) {
        // pixels of tiles that belong to other ranks must read as "miss" / "not visible"
        e = cudaMemsetAsync(d_hits, 0xFF, slots * sizeof(dodrt_hit), st);
        if (e == cudaSuccess && num_lights) e = cudaMemsetAsync(d_vis, 0, slots * (size_t)num_lights, st);
    }
    // ---- staged (pageable host buffers).  Copies run on a second stream so that they hide behind kernels:
    //  * with shadow passes, the primary hit records (16 of the 17 B per pixel) go to the host WHILE the shadow
    //    kernels -- which only read them -- run; each light's visibility bytes follow their own pass;
    //  * a primary-only frame is traced in `bands` bands (whole tile rows for full-frame results, runs of local
    //    tiles for compact results, so a band's results are contiguous) and band k is copied while band k+1 is
    //    traced.  More bands cost ~0.2 ms each in launch/tail overhead (profiles/r01_e2e_bands.txt), hence 2.
    if (e == cudaSuccess) {
        uint32_t tilesX, localTiles;
        frameTiles(frame, &tilesX, &localTiles);
        const uint64_t tilePixels = (uint64_t)frame->tile_w * frame->tile_h;
        const bool bandable = num_lights == 0 && (frame->compact || frame->tile_stride == 1) && slots >= (1u << 19);
        uint32_t bandTiles = localTiles;
        if (bandable) {
            static const uint32_t bands = [] {
                const char *b = std::getenv("DODRT_BANDS");
                const int v = b ? std::atoi(b) : 2;
                return (uint32_t)(v > 0 ? v : 2);
            }();
            bandTiles = (localTiles + bands - 1) / bands;
            if (!frame->compact) bandTiles = ((bandTiles + tilesX - 1) / tilesX) * tilesX; // whole tile rows
        }
        for (uint32_t begin = 0; begin < localTiles && rc == DODRT_OK && e == cudaSuccess; begin += bandTiles) {
            const uint32_t count = bandTiles < localTiles - begin ? bandTiles : localTiles - begin;
            rc = launchFrame(s, kModePrimary, frame, d_tables, d_tables + frame->width, d_hits, nullptr, nullptr, st, begin,
                             count);
            if (rc != DODRT_OK) break;
            uint64_t lo = 0, hi = slots; // result range of this band
            if (bandable && frame->compact) {
                lo = begin * tilePixels;
                hi = (begin + (uint64_t)count) * tilePixels;
            } else if (bandable) {
                lo = (uint64_t)(begin / tilesX) * frame->tile_h * frame->width;
                hi = (uint64_t)((begin + count) / tilesX) * frame->tile_h * frame->width;
                if (begin + count >= localTiles || hi > slots) hi = slots;
            }
            e = stagingFence(s, st, s->copyStream);
            if (e == cudaSuccess) {
                e = cudaMemcpyAsync(hits + lo, d_hits + lo, (hi - lo) * sizeof(dodrt_hit), cudaMemcpyDeviceToHost, s->copyStream);
            }
        }
        for (uint32_t l = 0; l < num_lights && rc == DODRT_OK && e == cudaSuccess; l++) {
            rc = launchFrame(s, kModeShadow, frame, d_tables, d_tables + frame->width, d_hits, lights + 3 * l,
                             d_vis + slots * l, st);
            if (rc != DODRT_OK) break;
            e = stagingFence(s, st, s->copyStream);
            if (e == cudaSuccess) {
                e = cudaMemcpyAsync(visible + slots * l, d_vis + slots * l, slots, cudaMemcpyDeviceToHost, s->copyStream);
            }
        }
    }
    // the staging buffers are re-used by the next call: both streams must have drained
    cudaError_t ec = cudaStreamSynchronize(s->copyStream);
    cudaError_t es = cudaStreamSynchronize(st);
    if (rc != DODRT_OK) return rc;
    if (e == cudaSuccess) e = ec;
    if (e == cudaSuccess) e = es;
    if (e != cudaSuccess) return fail(DODRT_E_CUDA, "dodrt_trace_frame: %s", cudaGetErrorString(e));
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_trace_primary(dodrt_scene *s, const dodrt_frame *frame, const float *xs, const float *ys, dodrt_hit *hits)
try {
    return dodrt_trace_frame(s, frame, xs, ys, nullptr, 0, hits, nullptr);
}
DODRT_CATCH

int dodrt_trace_shadow(dodrt_scene *s, const dodrt_frame *frame, const float *xs, const float *ys,
                       const dodrt_hit *hits, const float light[3], uint8_t *visible)
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    int rc = checkFrame(frame);
    if (rc != DODRT_OK) return rc;
    if (!xs || !ys || !hits || !light || !visible) return fail(DODRT_E_INVALID, "NULL argument");
    rc = ensureStream(s);
    if (rc != DODRT_OK) return rc;
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    cudaStream_t st = s->stream;
    const uint64_t slots = frame->compact ? frameSlots(frame) : (uint64_t)frame->width * frame->height;
    if (slots == 0) return DODRT_OK;
    std::lock_guard<std::mutex> hostLock(s->hostMutex); // persistent staging, like dodrt_trace_frame
    cudaError_t e = growStaging(s->stTables, s->stTablesBytes, ((size_t)frame->width + frame->height) * sizeof(float));
    if (e == cudaSuccess) e = growStaging(s->stHits, s->stHitsBytes, slots * sizeof(dodrt_hit));
    if (e == cudaSuccess) e = growStaging(s->stVis, s->stVisBytes, slots);
    if (e != cudaSuccess) return fail(DODRT_E_CUDA, "dodrt_trace_shadow staging: %s", cudaGetErrorString(e));
    float *d_tables = s->stTables;
    dodrt_hit *d_hits = s->stHits;
    uint8_t *d_vis = s->stVis;
    e = cudaMemcpyAsync(d_tables, xs, frame->width * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        e = cudaMemcpyAsync(d_tables + frame->width, ys, frame->height * sizeof(float), cudaMemcpyHostToDevice, st);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_hits, hits, slots * sizeof(dodrt_hit), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_vis, 0, slots, st);
    if (e == cudaSuccess) {
        rc = launchFrame(s, kModeShadow, frame, d_tables, d_tables + frame->width, d_hits, light, d_vis, st);
        if (rc == DODRT_OK) e = cudaMemcpyAsync(visible, d_vis, slots, cudaMemcpyDeviceToHost, st);
    }
    cudaError_t es = cudaStreamSynchronize(st);
    if (rc != DODRT_OK) return rc;
    if (e == cudaSuccess) e = es;
    if (e != cudaSuccess) return fail(DODRT_E_CUDA, "dodrt_trace_shadow: %s", cudaGetErrorString(e));
    return DODRT_OK;
}
DODRT_CATCH

// ---- several GPUs, one host process ------------------------------------------------------------------------------

struct dodrt_multi {
    std::vector<dodrt_scene *> scenes;
    std::mutex mutex;
    dodrt_frame_buffer *frameBuffer = nullptr;          // pageable host buffers: the frame assembles in scenes[0]'s HBM
    std::vector<dodrt_frame_buffer *> views;            // per scene; views[0] == frameBuffer
};

static void multiDropFrameBuffer(dodrt_multi *m)
{
    for (size_t i = 1; i < m->views.size(); i++) dodrt_frame_buffer_destroy(m->views[i]);
    m->views.clear();
    dodrt_frame_buffer_destroy(m->frameBuffer);
    m->frameBuffer = nullptr;
}

int dodrt_multi_create(dodrt_scene *const *scenes, uint32_t num_scenes, dodrt_multi **multi)
try {
    if (!scenes || !multi || num_scenes == 0) return fail(DODRT_E_INVALID, "NULL or empty scene list");
    *multi = nullptr;
    for (uint32_t i = 0; i < num_scenes; i++) {
        if (!scenes[i]) return fail(DODRT_E_INVALID, "scenes[%u] is NULL", i);
        for (uint32_t j = 0; j < i; j++) {
            if (scenes[j]->device == scenes[i]->device) return fail(DODRT_E_INVALID, "scenes[%u] and scenes[%u] share device %d", j, i, scenes[i]->device);
        }
        int rc = ensureStream(scenes[i]);
        if (rc != DODRT_OK) return rc;
    }
    std::unique_ptr<dodrt_multi> m(new dodrt_multi());
    m->scenes.assign(scenes, scenes + num_scenes);
    *multi = m.release();
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_multi_destroy(dodrt_multi *m)
try {
    if (!m) return DODRT_OK;
    multiDropFrameBuffer(m);
    delete m;
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_multi_trace_frame(dodrt_multi *m, const dodrt_frame *frame, const float *xs, const float *ys, const float *lights,
                            uint32_t num_lights, dodrt_hit *hits, uint8_t *visible)
try {
    if (!m) return fail(DODRT_E_INVALID, "multi is NULL");
    int rc = checkFrame(frame);
    if (rc != DODRT_OK) return rc;
    if (!xs || !ys || !hits) return fail(DODRT_E_INVALID, "NULL table/hit buffer");
    if (num_lights && (!lights || !visible)) return fail(DODRT_E_INVALID, "NULL lights/visible buffer");
    if (num_lights > (uint32_t)kMaxLights) return fail(DODRT_E_LIMIT, "%u lights exceed the limit of %d", num_lights, kMaxLights);
    std::lock_guard<std::mutex> lock(m->mutex);
    const uint32_t n = (uint32_t)m->scenes.size();
    const uint64_t pixels = (uint64_t)frame->width * frame->height;
    // where do the results go?  Pinned host buffers: straight from every GPU's kernel.  Otherwise via scenes[0]'s HBM.
    void *mHits = nullptr, *mVis = nullptr;
    {
        DeviceGuard guard(m->scenes[0]->device);
        mHits = mappedHostPointer(hits, pixels * sizeof(dodrt_hit));
        mVis = num_lights ? mappedHostPointer(visible, pixels * num_lights) : nullptr;
    }
    const bool pinned = mHits && (num_lights == 0 || mVis);
    // Pinned host frame, default: STAGED BANDS.  The frame is cut into bands of whole pixel rows (tiles as wide as the
    // frame, or half / a third of it beyond kMaxTileSide), dealt round-robin; every GPU traces its bands into its own
    // full-frame staging buffer and DMAs each band to the same place of the caller's frame -- contiguous rows, so a band
    // is one copy -- with the hit records on their way while the shadow pass runs.  DODRT_ZEROCOPY=1 selects the round-2
    // first form instead (the kernels store into the mapped host frame themselves), which measured slower on this pool:
    // dragon4k 5.8 / 3.5 / 3.2 / 3.6 ms on 1 / 2 / 4 / 8 GPUs (tests/tools/multi_bench.py).
    const char *zcEnv = std::getenv("DODRT_ZEROCOPY");
    const bool direct = pinned && zcEnv && std::atoi(zcEnv) != 0;
    if (pinned && !direct) {
        const uint32_t W = frame->width, H = frame->height;
        const uint32_t tilesX = (W + kMaxTileSide - 1) / kMaxTileSide;
        const uint32_t tileW = (((W + tilesX - 1) / tilesX) + 7u) & ~7u;
        uint32_t bandH = (H / (16u * n)) & ~3u; // ~16 bands per GPU: the shares differ by one band at most
        bandH = bandH < 4u ? 4u : (bandH > 64u ? 64u : bandH);
        const uint32_t bandsY = (H + bandH - 1) / bandH;
        const uint64_t totalTiles = (uint64_t)tilesX * bandsY;
        std::vector<cudaEvent_t> primaryDone(n, nullptr);
        std::vector<std::vector<cudaEvent_t>> shadowDone(n);
        std::vector<std::unique_lock<std::mutex>> hostLocks; // released when the call leaves, whichever way
        hostLocks.reserve(n);
        for (uint32_t i = 0; i < n; i++) shadowDone[i].reserve(num_lights);
        cudaError_t e = cudaSuccess;
        uint32_t locked = 0;
        // phase 1: every GPU gets its tables and all its launches (asynchronous; the GPUs trace at the same time)
        for (uint32_t i = 0; i < n && rc == DODRT_OK && e == cudaSuccess; i++) {
            dodrt_scene *s = m->scenes[i];
            DeviceGuard guard(s->device);
            if (!guard.ok) {
                rc = fail(DODRT_E_CUDA, "cannot select device %d", s->device);
                break;
            }
            hostLocks.emplace_back(s->hostMutex); // persistent staging, like dodrt_trace_frame
            locked = i + 1;
            s->stEventsUsed = 0;
            if (i >= totalTiles) continue;
            dodrt_frame f = *frame;
            f.tile_w = tileW, f.tile_h = bandH, f.first_tile = i, f.tile_stride = n, f.compact = 0;
            e = growStaging(s->stTables, s->stTablesBytes, ((size_t)W + H) * sizeof(float));
            if (e == cudaSuccess) e = growStaging(s->stHits, s->stHitsBytes, pixels * sizeof(dodrt_hit));
            if (e == cudaSuccess) e = growStaging(s->stVis, s->stVisBytes, pixels * (size_t)(num_lights ? num_lights : 1));
            if (e == cudaSuccess) e = cudaMemcpyAsync(s->stTables, xs, W * sizeof(float), cudaMemcpyHostToDevice, s->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(s->stTables + W, ys, H * sizeof(float), cudaMemcpyHostToDevice, s->stream);
            if (e != cudaSuccess) break;
            rc = launchFrame(s, kModePrimary, &f, s->stTables, s->stTables + W, s->stHits, nullptr, nullptr, s->stream);
            if (rc != DODRT_OK) break;
            e = stagingEvent(s, &primaryDone[i]);
            if (e == cudaSuccess) e = cudaEventRecord(primaryDone[i], s->stream);
            for (uint32_t l = 0; l < num_lights && rc == DODRT_OK && e == cudaSuccess; l++) {
                rc = launchFrame(s, kModeShadow, &f, s->stTables, s->stTables + W, s->stHits, lights + 3 * l, s->stVis + pixels * l,
                                 s->stream);
                if (rc != DODRT_OK) break;
                cudaEvent_t ev = nullptr;
                e = stagingEvent(s, &ev);
                if (e == cudaSuccess) e = cudaEventRecord(ev, s->stream);
                if (e == cudaSuccess) shadowDone[i].push_back(ev);
            }
        }
        // phase 2: the copies, band by band, on every GPU's copy stream behind the pass that produces them.  (One host
        // thread per GPU instead of the two phases was measured too: 1.87 vs 1.78 ms on 8 GPUs -- the frame is bound by
        // the D2H traffic into one host frame, ~90 GB/s in aggregate, not by the ~40 CUDA calls per GPU.)
        auto copyBands = [&](dodrt_scene *s, uint32_t i, const void *src, void *dst, size_t elem) -> cudaError_t {
            cudaError_t ce = cudaSuccess;
            for (uint64_t k = i; k < totalTiles && ce == cudaSuccess; k += n) {
                const uint32_t tx = (uint32_t)(k % tilesX), ty = (uint32_t)(k / tilesX);
                const uint32_t row0 = ty * bandH, rows = bandH < H - row0 ? bandH : H - row0;
                const uint32_t col0 = tx * tileW, cols = tileW < W - col0 ? tileW : W - col0;
                const size_t off = ((size_t)row0 * W + col0) * elem;
                if (cols == W) {
                    ce = cudaMemcpyAsync(static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, (size_t)rows * W * elem,
                                         cudaMemcpyDeviceToHost, s->copyStream);
                } else {
                    ce = cudaMemcpy2DAsync(static_cast<char *>(dst) + off, (size_t)W * elem, static_cast<const char *>(src) + off,
                                           (size_t)W * elem, (size_t)cols * elem, rows, cudaMemcpyDeviceToHost, s->copyStream);
                }
            }
            return ce;
        };
        for (uint32_t i = 0; i < locked && rc == DODRT_OK && e == cudaSuccess; i++) {
            dodrt_scene *s = m->scenes[i];
            if (!primaryDone[i]) continue;
            DeviceGuard guard(s->device);
            e = cudaStreamWaitEvent(s->copyStream, primaryDone[i], 0);
            if (e == cudaSuccess) e = copyBands(s, i, s->stHits, hits, sizeof(dodrt_hit));
            for (uint32_t l = 0; l < (uint32_t)shadowDone[i].size() && e == cudaSuccess; l++) {
                e = cudaStreamWaitEvent(s->copyStream, shadowDone[i][l], 0);
                if (e == cudaSuccess) e = copyBands(s, i, s->stVis + pixels * l, visible + pixels * l, 1);
            }
        }
        for (uint32_t i = 0; i < locked; i++) { // the staging buffers are re-used by the next call: both streams must have drained
            dodrt_scene *s = m->scenes[i];
            DeviceGuard guard(s->device);
            cudaError_t ec = cudaStreamSynchronize(s->copyStream);
            cudaError_t es = cudaStreamSynchronize(s->stream);
            if (e == cudaSuccess) e = ec;
            if (e == cudaSuccess) e = es;
        }
        if (rc != DODRT_OK) return rc;
        if (e != cudaSuccess) return fail(DODRT_E_CUDA, "dodrt_multi_trace_frame: %s", cudaGetErrorString(e));
        return DODRT_OK;
    }
    if (!direct) {
        dodrt_frame_buffer *fb = m->frameBuffer;
        if (!fb || fb->width != frame->width || fb->height != frame->height || fb->numLights < num_lights) {
            multiDropFrameBuffer(m);
            rc = dodrt_frame_buffer_create(m->scenes[0], frame->width, frame->height, num_lights, &m->frameBuffer);
            if (rc != DODRT_OK) return rc;
            m->views.assign(1, m->frameBuffer);
            for (uint32_t i = 1; i < n; i++) {
                dodrt_frame_buffer *v = nullptr;
                rc = dodrt_frame_buffer_attach(m->scenes[i], m->frameBuffer, &v);
                if (rc != DODRT_OK) {
                    multiDropFrameBuffer(m);
                    return rc;
                }
                m->views.push_back(v);
            }
        }
    }
    cudaError_t e = cudaSuccess;
    uint32_t launched = 0;
    std::vector<std::unique_lock<std::mutex>> hostLocks; // released when the call leaves, whichever way
    hostLocks.reserve(n);
    for (uint32_t i = 0; i < n && rc == DODRT_OK && e == cudaSuccess; i++) { // asynchronous: all GPUs trace at the same time
        dodrt_scene *s = m->scenes[i];
        DeviceGuard guard(s->device);
        if (!guard.ok) {
            rc = fail(DODRT_E_CUDA, "cannot select device %d", s->device);
            break;
        }
        hostLocks.emplace_back(s->hostMutex);
        launched = i + 1;
        dodrt_frame f = *frame;
        f.first_tile = i;
        f.tile_stride = n;
        f.compact = 1;
        const uint64_t slots = frameSlots(&f);
        if (slots == 0) continue;
        const size_t tableFloats = (size_t)frame->width + frame->height;
        e = growStaging(s->stTables, s->stTablesBytes, tableFloats * sizeof(float));
        if (e == cudaSuccess) e = growStaging(s->stHits, s->stHitsBytes, slots * sizeof(dodrt_hit));
        if (e == cudaSuccess) e = growStaging(s->stVis, s->stVisBytes, slots * (size_t)(num_lights ? num_lights : 1));
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->stTables, xs, frame->width * sizeof(float), cudaMemcpyHostToDevice, s->stream);
        if (e == cudaSuccess) {
            e = cudaMemcpyAsync(s->stTables + frame->width, ys, frame->height * sizeof(float), cudaMemcpyHostToDevice, s->stream);
        }
        if (e != cudaSuccess) break;
        Mirror mir;
        mir.byPixel = 1;
        mir.lightStride = pixels;
        if (direct) {
            mir.hits = static_cast<dodrt_hit *>(mHits);
            mir.visible = static_cast<uint8_t *>(mVis);
        } else {
            mir.hits = m->views[i]->hits();
            mir.visible = m->views[i]->visible();
        }
        rc = traceFrameOnDevice(s, &f, s->stTables, s->stTables + frame->width, lights, num_lights, s->stHits, s->stVis, slots, &mir,
                                s->stream);
    }
    for (uint32_t i = 0; i < launched; i++) {
        dodrt_scene *s = m->scenes[i];
        DeviceGuard guard(s->device);
        cudaError_t es = cudaStreamSynchronize(s->stream);
        if (e == cudaSuccess) e = es;
    }
    if (rc != DODRT_OK) return rc;
    if (e == cudaSuccess && !direct) {
        DeviceGuard guard(m->scenes[0]->device);
        e = cudaMemcpy(hits, m->frameBuffer->hits(), pixels * sizeof(dodrt_hit), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && num_lights) e = cudaMemcpy(visible, m->frameBuffer->visible(), pixels * num_lights, cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) return fail(DODRT_E_CUDA, "dodrt_multi_trace_frame: %s", cudaGetErrorString(e));
    return DODRT_OK;
}
DODRT_CATCH

// ---- shading + bounce loop ----------------------------------------------------------------

// rayTrace (main.cpp:273-347) for the call's tiles, asynchronous on the scene's stream.  `d_rgb`: where the finish kernel
// stores the 8-bit pixels (by pixel: any memory this GPU can write -- local, a peer GPU's, pinned host; by slot: local).
// The working arrays come from the scene's pool and are released on the stream.
static int renderOnDevice(dodrt_scene *s, const dodrt_frame *frame, const float *xs, const float *ys, const float *lights,
                          uint32_t num_lights, uint32_t depth, uint8_t *d_rgb, bool rgbByPixel)
{
    cudaStream_t st = s->stream;
    RenderParams rp{};
    rp.scene = s->dev;
    rp.frame = *frame;
    uint32_t tiles;
    frameTiles(frame, &rp.tiles_x, &tiles);
    const uint64_t n = (uint64_t)tiles * frame->tile_w * frame->tile_h;
    rp.slots = n;
    if (n == 0) return DODRT_OK;
    rp.num_lights = num_lights;
    for (uint32_t l = 0; l < num_lights; l++) {
        for (int k = 0; k < 4; k++) rp.lights[l][k] = lights[l * 4 + k];
    }
    float *d_tables = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync(&d_tables, ((size_t)frame->width + frame->height) * sizeof(float), s->pool, st);
    if (e == cudaSuccess) e = cudaMallocFromPoolAsync(&rp.rays, n * sizeof(dodrt_ray), s->pool, st);
    if (e == cudaSuccess) e = cudaMallocFromPoolAsync(&rp.hits, n * sizeof(dodrt_hit), s->pool, st);
    if (e == cudaSuccess) e = cudaMallocFromPoolAsync(&rp.visible, n * (num_lights ? num_lights : 1), s->pool, st);
    if (e == cudaSuccess) e = cudaMallocFromPoolAsync(&rp.accum, n * sizeof(float4), s->pool, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tables, xs, frame->width * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        e = cudaMemcpyAsync(d_tables + frame->width, ys, frame->height * sizeof(float), cudaMemcpyHostToDevice, st);
    }
    // scratch of the per-bounce spatial sort (DODRT_RENDER_SORT=0 switches it off; rays < 2^32)
    const char *sortEnv = std::getenv("DODRT_RENDER_SORT");
    const bool sortBounces = (!sortEnv || std::atoi(sortEnv) != 0) && num_lights != 0 && depth > 1 && n < 0xFFFFFFFFull;
    uint32_t *bins = nullptr, *order = nullptr;
    const char *cellEnv = std::getenv("DODRT_RENDER_CELL_BITS");
    const uint32_t cellBits = cellEnv ? (uint32_t)std::min(7, std::max(4, std::atoi(cellEnv))) : 5u;
    const char *sortFromEnv = std::getenv("DODRT_RENDER_SORT_FROM");
    const uint32_t sortFrom = sortFromEnv ? (uint32_t)std::max(0, std::atoi(sortFromEnv)) : 1u;
    if (sortBounces && e == cudaSuccess) e = cudaMallocFromPoolAsync(&bins, sizeof(uint32_t) * (2u << (3u * cellBits)), s->pool, st);
    if (sortBounces && e == cudaSuccess) e = cudaMallocFromPoolAsync(&order, sizeof(uint32_t) * n, s->pool, st);
    rp.xs = d_tables;
    rp.ys = d_tables ? d_tables + frame->width : nullptr;
    rp.rgb = d_rgb;
    rp.rgb_by_pixel = rgbByPixel ? 1u : 0u;
    if (e == cudaSuccess) e = launch_render_init(rp, st);
    if (e == cudaSuccess) s->launches.fetch_add(1);
    for (uint32_t k = 0; k < depth && e == cudaSuccess; k++) { // main.cpp:312
        TraceParams p{};
        p.scene = s->dev;
        p.classes = frame->classes;
        p.rays = rp.rays;
        p.count = n;
        p.hits = rp.hits;
        // bounce passes: incoherent rays, long tails (dragon as-is frame 182 -> 166 ms with donation)
        p.variant = resolve_variant(s->variant, s->cfg[kDonateVariant][kModeRays], p.count, true, s->dev.num_nodes >= kBigTreeNodes);
        // bounce k >= 2 starts where bounce k-1 hit: the cell order of those hit points is the order of these origins
        if (k >= sortFrom + 1 && sortBounces && order) p.ray_order = order;
        p.counter = nextCounter(s);
        e = launchTraceOn(s, kModeRays, p, st); // closest-hit chain, main.cpp:314-321
        if (e == cudaSuccess) s->launches.fetch_add(1);
        if (num_lights && e == cudaSuccess) { // canSeeLight for every light (main.cpp:226): one launch, light-major
            // from the second bounce on the hit points of neighbouring pixels are scattered: walk them cell by cell
            if (k >= sortFrom && sortBounces && order) {
                e = launch_render_sort(rp, cellBits, bins, order, st);
                if (e == cudaSuccess) s->launches.fetch_add(3);
                p.ray_order = order;
            }
            p.counter = nextCounter(s);
            p.visible = rp.visible;
            p.num_lights = num_lights;
            p.rays_per_light = n;
            p.count = n * num_lights;
            for (uint32_t l = 0; l < num_lights; l++) {
                for (int c = 0; c < 3; c++) p.lights[l][c] = rp.lights[l][c];
            }
            for (int c = 0; c < 3; c++) p.light[c] = rp.lights[0][c];
            p.variant = resolve_variant(s->variant, s->cfg[kDonateVariant][kModeShadowRays], p.count, true, s->dev.num_nodes >= kBigTreeNodes);
            e = launchTraceOn(s, kModeShadowRays, p, st);
            if (e == cudaSuccess) s->launches.fetch_add(1);
        }
        if (e == cudaSuccess) e = launch_render_shade(rp, k, st);
        if (e == cudaSuccess) s->launches.fetch_add(1);
    }
    if (e == cudaSuccess) e = launch_render_finish(rp, st);
    if (e == cudaSuccess) s->launches.fetch_add(1);
    if (d_tables) cudaFreeAsync(d_tables, st);
    if (rp.rays) cudaFreeAsync(rp.rays, st);
    if (rp.hits) cudaFreeAsync(rp.hits, st);
    if (rp.visible) cudaFreeAsync(rp.visible, st);
    if (rp.accum) cudaFreeAsync(rp.accum, st);
    if (bins) cudaFreeAsync(bins, st);
    if (order) cudaFreeAsync(order, st);
    if (e != cudaSuccess) return fail(DODRT_E_CUDA, "dodrt_render: %s", cudaGetErrorString(e));
    return DODRT_OK;
}

static int checkRender(dodrt_scene *s, const dodrt_frame *frame, const float *xs, const float *ys, const float *lights,
                       uint32_t num_lights, const uint8_t *rgb)
{
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    int rc = checkFrame(frame);
    if (rc != DODRT_OK) return rc;
    if (!xs || !ys || !rgb || (num_lights && !lights)) return fail(DODRT_E_INVALID, "NULL argument");
    if (num_lights > (uint32_t)kMaxLights) return fail(DODRT_E_LIMIT, "%u lights exceed the limit of %d", num_lights, kMaxLights);
    if ((s->dev.num_tri_lanes && !s->dev.tri_attrs) || (s->dev.num_spheres && !s->dev.sphere_colors) ||
        (s->dev.num_planes && !s->dev.plane_colors)) {
        return fail(DODRT_E_INVALID, "dodrt_scene_set_shading has not been called for this scene");
    }
    return DODRT_OK;
}

int dodrt_render(dodrt_scene *s, const dodrt_frame *frame, const float *xs, const float *ys, const float *lights,
                 uint32_t num_lights, uint32_t depth, uint8_t *rgb)
try {
    int rc = checkRender(s, frame, xs, ys, lights, num_lights, rgb);
    if (rc != DODRT_OK) return rc;
    rc = ensureStream(s);
    if (rc != DODRT_OK) return rc;
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    std::lock_guard<std::mutex> hostLock(s->hostMutex);
    cudaStream_t st = s->stream;
    const bool byPixel = !frame->compact;
    const uint64_t outPixels = byPixel ? (uint64_t)frame->width * frame->height : frameSlots(frame);
    if (outPixels == 0) return DODRT_OK;
    uint8_t *d_rgb = nullptr;
    CUDA_TRY(cudaMallocFromPoolAsync(&d_rgb, outPixels * 3, s->pool, st));
    cudaError_t e = cudaSuccess;
    if (byPixel && frame->tile_stride > 1) e = cudaMemsetAsync(d_rgb, 0, outPixels * 3, st); // other ranks' pixels read as black
    if (e == cudaSuccess) rc = renderOnDevice(s, frame, xs, ys, lights, num_lights, depth, d_rgb, byPixel);
    if (e == cudaSuccess && rc == DODRT_OK) e = cudaMemcpyAsync(rgb, d_rgb, outPixels * 3, cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(d_rgb, st);
    cudaError_t es = cudaStreamSynchronize(st);
    if (rc != DODRT_OK) return rc;
    if (e == cudaSuccess) e = es;
    if (e != cudaSuccess) return fail(DODRT_E_CUDA, "dodrt_render: %s", cudaGetErrorString(e));
    return DODRT_OK;
}
DODRT_CATCH

// The reference renders the frame with all the threads of the process (one row band each, main.cpp:371-394); this renders
// it with all the GPUs of the process: tiles dealt round-robin, every GPU's finish kernel stores its pixels straight into
// the ONE rgb frame -- the caller's pinned host buffer, or (pageable host memory) a frame in scenes[0]'s HBM over peer
// stores, copied out once.
int dodrt_multi_render(dodrt_multi *m, const dodrt_frame *frame, const float *xs, const float *ys, const float *lights,
                       uint32_t num_lights, uint32_t depth, uint8_t *rgb)
try {
    if (!m) return fail(DODRT_E_INVALID, "multi is NULL");
    std::lock_guard<std::mutex> lock(m->mutex);
    const uint32_t n = (uint32_t)m->scenes.size();
    for (uint32_t i = 0; i < n; i++) {
        int rc = checkRender(m->scenes[i], frame, xs, ys, lights, num_lights, rgb);
        if (rc != DODRT_OK) return rc;
    }
    const uint64_t bytes = (uint64_t)frame->width * frame->height * 3;
    uint8_t *dest = nullptr, *d_frame = nullptr;
    {
        DeviceGuard guard(m->scenes[0]->device);
        dest = static_cast<uint8_t *>(mappedHostPointer(rgb, bytes));
        if (!dest) { // pageable: assemble in scenes[0]'s HBM (peer access as for the frame buffers)
            CUDA_TRY(cudaMalloc(&d_frame, bytes));
            dest = d_frame;
        }
    }
    int rc = DODRT_OK;
    uint32_t launched = 0;
    std::vector<std::unique_lock<std::mutex>> hostLocks; // released when the call leaves, whichever way
    hostLocks.reserve(n);
    for (uint32_t i = 0; i < n && rc == DODRT_OK; i++) {
        dodrt_scene *s = m->scenes[i];
        DeviceGuard guard(s->device);
        if (!guard.ok) {
            rc = fail(DODRT_E_CUDA, "cannot select device %d", s->device);
            break;
        }
        if (d_frame && s->device != m->scenes[0]->device) {
            cudaError_t pe = cudaDeviceEnablePeerAccess(m->scenes[0]->device, 0);
            if (pe == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
            } else if (pe != cudaSuccess) {
                rc = fail(DODRT_E_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", s->device, m->scenes[0]->device, cudaGetErrorString(pe));
                break;
            }
        }
        hostLocks.emplace_back(s->hostMutex);
        launched = i + 1;
        dodrt_frame f = *frame;
        f.first_tile = i;
        f.tile_stride = n;
        f.compact = 0;
        rc = renderOnDevice(s, &f, xs, ys, lights, num_lights, depth, dest, true);
    }
    cudaError_t e = cudaSuccess;
    for (uint32_t i = 0; i < launched; i++) {
        dodrt_scene *s = m->scenes[i];
        DeviceGuard guard(s->device);
        cudaError_t es = cudaStreamSynchronize(s->stream);
        if (e == cudaSuccess) e = es;
    }
    if (d_frame) {
        DeviceGuard guard(m->scenes[0]->device);
        if (rc == DODRT_OK && e == cudaSuccess) e = cudaMemcpy(rgb, d_frame, bytes, cudaMemcpyDeviceToHost);
        cudaFree(d_frame);
    }
    if (rc != DODRT_OK) return rc;
    if (e != cudaSuccess) return fail(DODRT_E_CUDA, "dodrt_multi_render: %s", cudaGetErrorString(e));
    return DODRT_OK;
}
DODRT_CATCH

// ---- frame helpers ------------------------------------------------------------------------------

int dodrt_frame_local_pixels(const dodrt_frame *frame, uint64_t *slots)
try {
    int rc = checkFrame(frame);
    if (rc != DODRT_OK) return rc;
    if (!slots) return fail(DODRT_E_INVALID, "slots is NULL");
    *slots = frameSlots(frame);
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_frame_pixel_map(const dodrt_frame *f, uint32_t *pixel_of_slot, uint64_t slots)
try {
    int rc = checkFrame(f);
    if (rc != DODRT_OK) return rc;
    if (!pixel_of_slot && slots) return fail(DODRT_E_INVALID, "pixel_of_slot is NULL");
    if (slots != frameSlots(f)) return fail(DODRT_E_INVALID, "slots does not match the frame description");
    uint32_t tilesX, localTiles;
    frameTiles(f, &tilesX, &localTiles);
    const uint32_t tilePixels = f->tile_w * f->tile_h;
    const uint32_t bpr = f->tile_w >> 3;
    for (uint64_t slot = 0; slot < slots; slot++) { // same mapping as slot_to_pixel in dodrt_kernels.cu
        const uint32_t localTile = (uint32_t)(slot / tilePixels), in = (uint32_t)(slot % tilePixels);
        const uint32_t tile = f->first_tile + localTile * f->tile_stride;
        const uint32_t tx = tile % tilesX, ty = tile / tilesX;
        const uint32_t block = in >> 5, lane = in & 31u;
        const uint32_t col = tx * f->tile_w + (block % bpr) * 8 + (lane & 7u);
        const uint32_t row = ty * f->tile_h + (block / bpr) * 4 + (lane >> 3);
        pixel_of_slot[slot] = (col < f->width && row < f->height) ? row * f->width + col : 0xFFFFFFFFu;
    }
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_scene_debug_stats(dodrt_scene *s, int enable, uint64_t out[8])
try {
    if (!s) return fail(DODRT_E_INVALID, "scene is NULL");
    std::lock_guard<std::mutex> lock(s->mutex);
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(DODRT_E_CUDA, "cannot select device %d", s->device);
    CUDA_TRY(cudaDeviceSynchronize());
    if (out) {
        std::memset(out, 0, 8 * sizeof(uint64_t));
        if (s->d_stats) CUDA_TRY(cudaMemcpy(out, s->d_stats, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    }
    if (enable && !s->d_stats) CUDA_TRY(cudaMalloc(&s->d_stats, 8 * sizeof(unsigned long long)));
    if (s->d_stats) CUDA_TRY(cudaMemset(s->d_stats, 0, 8 * sizeof(unsigned long long)));
    s->dev.stats = enable ? s->d_stats : nullptr;
    return DODRT_OK;
}
DODRT_CATCH

int dodrt_scene_launch_count(dodrt_scene *s, uint64_t *launches)
try {
    if (!s || !launches) return fail(DODRT_E_INVALID, "NULL argument");
    *launches = s->launches.load();
    return DODRT_OK;
}
DODRT_CATCH

} // extern "C"

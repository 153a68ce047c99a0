// dodrt_kernels.cuh -- launch interface between the C ABI (dodrt_api.cu) and the kernels.
#pragma once
#include "dodrt_device.cuh"

namespace dodrt {

enum TraceMode : int {
    kModeRays = 0,    // explicit dodrt_ray batch            (dodrt_intersect*)
    kModePrimary = 1, // in-kernel primary ray generation     (dodrt_trace_primary*)
    kModeShadow = 2,  // in-kernel shadow ray generation      (dodrt_trace_shadow*)
    kModeShadowRays = 3, // shadow rays from an explicit ray batch + its hits (bounce loop of dodrt_render*)
    kModeFrame = 4,   // primary + shadow batches of a frame share in ONE launch (trace_frame_kernel, dodrt_trace_frame_device)
};
constexpr int kNumModes = 5;

constexpr int kNumVariants = 9;    // see the header comment of dodrt_kernels.cu
constexpr int kDefaultVariant = 3;
constexpr int kVariantAuto = -1;   // pick per launch, see resolve_variant
int default_variant();             // kVariantAuto unless env DODRT_VARIANT names a (compiled) variant
#ifdef DODRT_EXPERIMENTS
constexpr bool kExperiments = true;
#else
constexpr bool kExperiments = false;
#endif
// the product build instantiates variants 0, 3, 7 and no frame kernels; -DDODRT_EXPERIMENTS adds the rest
constexpr bool variant_compiled(int v) { return v == 0 || v == 3 || v == 7 || (kExperiments && v >= 0 && v < kNumVariants); }
constexpr bool mode_compiled(int m) { return m != 4 /* kModeFrame */ || kExperiments; }

struct TraceParams {
    DeviceScene scene;
    uint32_t classes;
    int variant;
    // kModeRays
    const dodrt_ray *rays;
    // all modes: number of work items (rays, or result slots of the frame incl. edge padding)
    uint64_t count;
    // kModeRays/kModePrimary: output ; kModeShadow: input
    dodrt_hit *hits;
    // frame modes
    dodrt_frame frame;
    uint32_t tiles_x;
    const float *xs, *ys;
    float light[3];
    uint8_t *visible;
    // dynamic work distribution: kCounterWords counters per launch, zeroed on the stream before the launch:
    // [0] next work item, [1] heavy tiles placed so far, [2] light tiles placed so far
    unsigned long long *counter;
    // frame modes: optional processing order of the call's local tiles (heavy-first, see order_tiles_kernel);
    // results still land in the natural slots.  nullptr = natural order.
    uint32_t *tile_order;
    uint32_t num_local_tiles;
    // variant 7 (ray donation, dodrt_donate.inl): per-launch queue of suspended rays, filled by warps that are still
    // working when others have run out of work; nullptr = donation off
    uint32_t *donate_slots;   // donate_capacity x kDonateSlotWords words
    uint32_t *donate_ready;   // donate_capacity words; slot i is filled when donate_ready[i] == donate_epoch
    uint32_t donate_epoch;    // non-zero, unique per launch on a persistent queue (stale words of earlier launches never match)
    uint32_t donate_capacity;
    // Frame modes: optional MIRROR of the results, written by the kernels themselves next to the local copy while they
    // run -- a peer GPU's frame over NVLink (image-tile split: every rank fills rank 0's row-major frame directly, no
    // gather, no assembly pass) or pinned host memory (the caller's buffers: no D2H copy after the kernel).  Indexed by
    // pixel (row * width + col) when mirror_by_pixel, else like the local buffers.  Visibility of light l at
    // mirror_visible[l * mirror_light_stride + index].
    dodrt_hit *mirror_hits;
    uint8_t *mirror_visible;
    uint64_t mirror_light_stride;
    uint32_t mirror_by_pixel;
    // kModeFrame: the lights of the shadow queue, visible[l * visible_light_stride + slot]; per-tile bookkeeping that
    // makes a tile's shadow batches claimable once its primary records are complete (all zeroed before the launch):
    // tile_done[t] = primary batches of local tile t finished; ready_queue[k] = 1 + the k-th tile that became complete.
    // nullptr = the block-fused form of the frame kernel (no queues at all)
    // kModeShadowRays with num_lights > 1: ONE launch for all lights of a bounce (item = light * rays_per_light + ray,
    // which is also its index in `visible`)
    uint32_t num_lights;
    float lights[16][3];
    uint64_t rays_per_light;
    // kModeShadowRays, optional: process the rays in this order (ray = ray_order[position]); results stay indexed by ray.
    // dodrt_render sorts the hit points of a bounce into spatial cells so that a warp's 32 shadow rays start close together
    const uint32_t *ray_order;
    uint64_t visible_light_stride;
    uint64_t shadow_count; // count * num_lights
    uint32_t *tile_done;
    uint32_t *ready_queue;
};

// counters (one 256-B block per launch): [0] next work item, [1]/[2] heavy/light tiles placed (order_tiles_kernel);
// kModeFrame with tile queues: [3] next shadow item, [4] tiles published in ready_queue;
// on their own 128-B line, away from the work counter every warp hammers: [16] warps that left the main loop,
// [17] donation tickets taken by helpers, [18] donation slots reserved by donors and forking helpers, [19] warps that
// entered the kernel, [20] donation slots completely served; [8] set once the queue is quiescent (dodrt_donate.inl)
constexpr int kCounterWords = 32;
constexpr int kShadowNext = 3, kReadyTail = 4;
constexpr int kDonateFinished = 16, kDonateHead = 17, kDonateTail = 18, kDonateStarted = 19, kDonateServed = 20;
constexpr int kDonateQuiet = 8; // on the work counter's line, which nobody writes any more once helpers exist
constexpr int kDonateVariant = 7;
constexpr int kDonateSlotWords = 80; // 24 header words + 16 stack entries x 3 + mirror index (2) + 6 spare = 320 B
constexpr int kDonateMirrorWord = 72;
constexpr int kDonateMaxStack = 16;
constexpr uint64_t kDonateBelowBatches = 64; // auto: donate when a pass has fewer 32-ray batches per warp than this

struct LaunchConfig {
    int grid;
    int block;
};

// Occupancy-derived persistent launch shape for the given device (cached by the caller).
cudaError_t trace_launch_config(int device, TraceMode mode, int variant, LaunchConfig *cfg);
int resolve_variant(int variant, const LaunchConfig &donateCfg, uint64_t count, bool split, bool bigTree);
constexpr uint32_t kBigTreeNodes = 8192;
// Donation queue of variant 7: either `queue` (persistent memory of donation_queue_bytes(cfg), zero-filled once when it was
// allocated, used with a fresh non-zero `epoch` per launch -- nothing to clear), or a per-launch allocation from `pool`
// (stream-ordered, ready words zeroed on the stream); neither = no donation.
size_t donation_queue_bytes(const LaunchConfig &cfg);
#ifdef DODRT_TIMELINE
void timeline_fetch(unsigned long long *exits, unsigned long long *donations, unsigned int *numDonations, unsigned int *pollsEmpty);
#endif
cudaError_t launch_trace(TraceMode mode, const TraceParams &p, const LaunchConfig &cfg, cudaStream_t stream,
                         cudaMemPool_t pool = nullptr, void *queue = nullptr, uint32_t epoch = 0);

// ---- shading / bounce loop (dodrt_render_kernels.cu): rayTrace, main.cpp:273-347 ---------------------------------
constexpr int kMaxLights = 16;
struct RenderParams {
    DeviceScene scene;
    // the call's share of the image: result slots of `frame` (tiles first_tile, first_tile + tile_stride, ..., 8x4 pixel
    // blocks inside a tile; padded slots of edge tiles carry DODRT_RAY_SKIP rays).  All per-pixel arrays are indexed by slot.
    dodrt_frame frame;
    uint32_t tiles_x;
    uint64_t slots;
    const float *xs, *ys;
    dodrt_ray *rays;     // current ray of every slot
    dodrt_hit *hits;     // its closest hit
    uint8_t *visible;    // [num_lights][slots] canSeeLight results of this bounce
    float4 *accum;       // finalColor
    // output, 3 bytes per pixel: indexed by pixel (row * width + col; local, a peer GPU's or pinned host memory -- the
    // kernel stores straight into it) when rgb_by_pixel, else by slot (padded slots zero)
    uint8_t *rgb;
    uint32_t rgb_by_pixel;
    uint32_t num_lights;
    float lights[kMaxLights][4]; // position xyz, intensity (light.h:4-8)
};
cudaError_t launch_render_init(const RenderParams &p, cudaStream_t stream);
cudaError_t launch_render_shade(const RenderParams &p, uint32_t bounce, cudaStream_t stream);
cudaError_t launch_render_finish(const RenderParams &p, cudaStream_t stream);
// Spatial order of a bounce's hit points for its shadow passes: order[k] = k-th ray in cell order (counting sort over the
// (1 << bits)^3 Morton cells of the room, 4 <= bits <= 7; `bins` = 2 << (3 * bits) words of scratch)
cudaError_t launch_render_sort(const RenderParams &p, uint32_t bits, uint32_t *bins, uint32_t *order, cudaStream_t stream);

// Multi-GPU: gathered per-rank compact results -> row-major frame (one thread per pixel).
cudaError_t launch_assemble(const dodrt_frame &frame, uint32_t tiles_x, const dodrt_hit *compactHits,
                            const uint8_t *compactVis, uint64_t slotsPerRank, dodrt_hit *hitsOut, uint8_t *visOut,
                            cudaStream_t stream);

// Upload helper: reference lanes (288 B, SoA of 8) -> per-triangle 48-B records with AB/AC.
// `d_prim_nums` (optional): output lane i = source lane d_prim_nums[i] (Triangle::reorderLanesByIndices fused in).
cudaError_t launch_repack_triangles(const float *d_lanes, const uint32_t *d_prim_nums, uint32_t num_lanes, float4 *d_tris,
                                    float4 *d_lanes4, cudaStream_t stream);
// per-triangle 48-B records (variants 0-2 only) from the SoA lanes; `d_tris` of launch_repack_triangles may be nullptr
cudaError_t launch_tris_from_lanes4(const float4 *d_lanes4, uint32_t num_lanes, float4 *d_tris, cudaStream_t stream);
// generic per-lane gather (shading attributes): out[i] = in[d_prim_nums[i]], `words_per_lane` 32-bit words each
cudaError_t launch_gather_lanes(const uint32_t *d_in, const uint32_t *d_prim_nums, uint32_t num_lanes, uint32_t words_per_lane,
                                uint32_t *d_out, cudaStream_t stream);

} // namespace dodrt

// dodrt_donate.inl -- variant 7: variant 3 plus RAY DONATION at the tail of a pass.
// Textually included by dodrt_kernels.cu inside namespace dodrt::{anonymous}.
//
// Why (profiles/r01_rank_tail.txt): a persistent pass ends with its slowest warp, and a 32-ray batch grazing the
// mesh runs 0.2-0.5 ms on a lone warp -- as long as a whole 8-GPU share of the frame.  The SMs that have run out of
// work idle meanwhile (rank 0 of 8: 44 % busy in the primary pass).
//
// How: a warp that cannot claim another batch does not exit; it becomes a HELPER and waits on a global queue.
// A warp still traversing polls (one volatile load every DeviceScene::donate_poll voted iterations) whether helpers exist; if
// they outnumber the queued rays it SUSPENDS its live rays -- ray, running clip, best hit, node / leaf cursor and
// the short stack -- into the queue and goes on to become a helper itself.  A helper resumes ONE ray with the whole
// warp: kd node steps are warp-uniform, a leaf is tested 128 triangle slots at a time (four per lane) and reduced with the
// lexicographic (t, slot) minimum, which is what the reference's slot-by-slot loop with its strict `<` leaves behind
// (triangle.cpp:119-139) -- the same reduction as leaf_step_coop, bit-identical to every other variant.
// A suspended ray resumes exactly where it stopped: no node is visited twice, the order of events per ray is the
// reference's (kdtree.cpp:263-361).  Helpers wait only for warps that have ENTERED the kernel (counted at entry), never
// for the grid size: if the SMs are shared with another kernel -- e.g. a second donating launch on another stream -- some
// blocks may not be resident yet, and they could never start while everybody spins.  A block that starts late finds
// the work counter exhausted and has nothing to donate, so the pass ends when every warp that entered has left its
// main loop and the queue is drained (tests/test_gpu_parity.py::test_concurrent_donating_launches).
//
// WORK SPLITTING of a resumed any-hit ray (round 2, profiles/r02_donation_fork.txt).  What was left of a short pass was a
// chain-length bound: one resumed ray walks ~500 dependent node / leaf steps at L2 latency (~0.2 us each) whoever executes
// it, while thousands of helpers wait.  For an any-hit query the answer is an OR over the leaves the no-hit traversal
// visits (kdtree.cpp:338-341), so the sub-trees on the ray's short stack are INDEPENDENT jobs: a helper that
// sees other helpers waiting (head > tail) FORKS -- it pushes its oldest stack entries as new queue slots (same ray, node
// = the entry, empty stack) and keeps the rest.  Pieces fork again when their own stacks grow.  The result needs no join:
// the first fork writes the optimistic answer ("visible" / "miss") once, before the children are published, and from
// then on a piece only ever writes "blocked" / "hit"; a piece that sees the answer already decided stops.
// Termination with forking helpers: a waiting helper may leave only in a QUIESCENT state -- every warp has left its main
// loop, and every reserved slot has been completely served (kDonateServed == tail: nobody is resuming, so nobody can
// fork) -- and then only if the final tail does not cover its ticket.  The state is detected by whoever causes it (the
// last warp to leave its main loop / the helper that serves the last slot: donate_check_quiescent) and announced in ONE
// word on a line nobody writes any more; the thousands of waiting helpers poll that word and their own ready word, not
// the queue's counters -- with every helper reading finished / started / served / tail each round, the counter line
// saturated and the donors' own polls and reservations took microseconds (1-of-8 share: 0.71 -> 1.04 ms).
enum FinishKind : uint32_t { kFinishRecord = 0, kFinishAnyRecord = 1, kFinishVisible = 2 };
constexpr uint32_t kFlagAny = 1u, kFlagFound = 2u, kFlagForked = 4u; // slot word 7, bits 0-2; FinishKind in bits 3-4

// What remains to be done with the kd-tree's answer for one ray (main.cpp:320-325 / 209-217 as the trace kernel
// applies them): where the result goes and what it is when the tree finds nothing.
constexpr uint64_t kNoMirror = ~0ull;
struct Finish {
    uint64_t out;
    uint64_t mirror; // index into the mirror buffers (TraceParams::mirror_*), kNoMirror = none
    uint32_t kind;
    float pre[4]; // kFinishRecord: the record before the tree (analytic hit or miss); kFinishAnyRecord: pre[0] = the ray's clip
};

__device__ __forceinline__ void finish_write(const TraceParams &p, uint32_t kind, uint64_t out, uint64_t mirror, bool found,
                                             const Hit &hit, const float pre[4])
{
    if (kind == kFinishVisible) {
        p.visible[out] = found ? 0 : 1;
        if (mirror != kNoMirror) {
            p.mirror_visible[mirror] = found ? 0 : 1;
        }
    } else if (kind == kFinishAnyRecord) {
        reinterpret_cast<float4 *>(p.hits)[out] = make_float4(pre[0], __uint_as_float(found ? 0u : DODRT_MISS), 0.0f, 0.0f);
    } else {
        const float4 r = found ? make_float4(hit.t, __uint_as_float(hit.prim), hit.u, hit.v) : make_float4(pre[0], pre[1], pre[2], pre[3]);
        reinterpret_cast<float4 *>(p.hits)[out] = r;
        if (mirror != kNoMirror) {
            reinterpret_cast<float4 *>(p.mirror_hits)[mirror] = r;
        }
    }
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// Queue protocol (counters kDonateFinished / kDonateHead / kDonateTail, on their own 128-B line):
//   * a helper takes a TICKET h = atomicAdd(head, 1) and then waits on ready[h] alone -- a word nobody else polls;
//   * donors reserve slots with atomicAdd(tail, n), fill them, fence, and set ready[slot];
//   * head - tail (when positive) is the number of helpers waiting with a ticket and no ray: donors give at most
//     that many rays, so a ray is only ever suspended when a whole warp is idle and waiting for it;
//   * a ticket holder whose slot was never filled leaves when every warp has left its main loop (no donor is left)
//     and tail <= its ticket.
// Donor side, called by every lane of the warp: should this warp suspend rays now, and how many at most?
__device__ __noinline__ uint32_t donate_poll_impl(const unsigned long long *counter, uint32_t capacity, uint32_t always)
{
    uint32_t want = 0;
    if ((threadIdx.x & 31u) == 0u) {
        const unsigned long long head = ld_volatile_u64(counter + kDonateHead);
        const unsigned long long tail = ld_volatile_u64(counter + kDonateTail);
        if (tail <= capacity / 2u) {
            if (always != 0u) {
                want = 32u;
            } else if (head > tail) {
                want = (uint32_t)(head - tail < 32ull ? head - tail : 32ull);
            }
        }
    }
    return __shfl_sync(0xffffffffu, want, 0);
}
__device__ __forceinline__ uint32_t donate_poll(const TraceParams &p) { return donate_poll_impl(p.counter, p.donate_capacity, p.scene.tune[3]); }

// Everything a helper needs to resume one ray.  The suspension is a cold path: the donor packs this struct (a few
// dozen stores to local memory, executed only when a ray is actually given away) and a NON-INLINED routine copies it
// into the queue, so the register allocation of the hot voted loop does not have to keep all of it live together with
// the temporaries of the copy (inlined, the suspension raised the kernels from 96 to 114 registers = 5 -> 4 blocks
// per SM).
struct SuspendedRay {
    float4 w[6]; // w[3] and the `out` half of w[5] are filled by donate_store from the Finish record
    const uint32_t *stackNode;
    const float *stackTmin, *stackTmax;
    int sp;
};

// `fin` lives in the caller's local memory (its address escapes to this non-inlined routine), so the finisher
// record -- never read by the traversal itself -- does not occupy registers across the voted loop.
__device__ __noinline__ void donate_store(uint32_t *slots, uint32_t *ready, uint32_t epoch, uint32_t slot,
                                          const SuspendedRay &r, const Finish *fin)
{
    float4 *w = reinterpret_cast<float4 *>(slots + (size_t)slot * kDonateSlotWords);
    w[0] = r.w[0];
    w[1] = make_float4(r.w[1].x, r.w[1].y, r.w[1].z, __uint_as_float(__float_as_uint(r.w[1].w) | (fin->kind << 3)));
    w[2] = r.w[2];
    w[3] = make_float4(fin->pre[0], fin->pre[1], fin->pre[2], fin->pre[3]);
    w[4] = r.w[4];
    w[5] = make_float4(r.w[5].x, r.w[5].y, __uint_as_float((uint32_t)fin->out), __uint_as_float((uint32_t)(fin->out >> 32)));
    uint32_t *stk = slots + (size_t)slot * kDonateSlotWords + 24;
    stk[kDonateMirrorWord - 24] = (uint32_t)fin->mirror;
    stk[kDonateMirrorWord - 24 + 1] = (uint32_t)(fin->mirror >> 32);
    for (int i = 0; i < r.sp; i++) {
        stk[3 * i + 0] = r.stackNode[i];
        stk[3 * i + 1] = __float_as_uint(r.stackTmin[i]);
        stk[3 * i + 2] = __float_as_uint(r.stackTmax[i]);
    }
    __threadfence();
    *reinterpret_cast<volatile uint32_t *>(ready + slot) = epoch;
}

// Suspends up to `want` live rays of the warp into the queue (lanes in ascending order).  The lanes concerned have
// st.live == false and donated == true afterwards.
__device__ __forceinline__ void donate_live_rays(const TraceParams &p, uint32_t want, TreeState &st, const float o[3],
                                                 const float d[3], bool any, float clip, const Hit &hit, bool found,
                                                 const Finish *fin, const uint32_t *stackNode, const float *stackTmin,
                                                 const float *stackTmax, bool &donated)
{
#ifdef DBG_NODONATE
    return;
#endif
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned candidates = __ballot_sync(0xffffffffu, st.live && st.sp <= kDonateMaxStack);
    const uint32_t rank = __popc(candidates & ((1u << lane) - 1u));
    const bool give = ((candidates >> lane) & 1u) != 0u && rank < want;
    const uint32_t n = min((uint32_t)__popc(candidates), want);
    if (n == 0u) {
        return;
    }
    unsigned long long base = 0;
    if (lane == 0) {
        base = atomicAdd(p.counter + kDonateTail, (unsigned long long)n);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
#ifdef DODRT_TIMELINE
    if (lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_tlDonation[atomicAdd(&g_tlDonations, 1u) & 0xFFFFu] = (t << 12) | ((unsigned long long)n << 6) | (unsigned long long)(__popc(candidates) - n);
    }
#endif
    if (give) {
        SuspendedRay r;
        const uint32_t flags = (any ? kFlagAny : 0u) | (found ? kFlagFound : 0u);
        r.w[0] = make_float4(o[0], o[1], o[2], d[0]);
        r.w[1] = make_float4(d[1], d[2], clip, __uint_as_float(flags));
        r.w[2] = make_float4(hit.t, __uint_as_float(hit.prim), hit.u, hit.v);
        r.w[4] = make_float4(st.tmin, st.tmax, __uint_as_float(st.node), __uint_as_float(st.triCur));
        r.w[5] = make_float4(__uint_as_float(st.triEnd), __uint_as_float((uint32_t)st.sp), 0.0f, 0.0f);
        r.stackNode = stackNode;
        r.stackTmin = stackTmin;
        r.stackTmax = stackTmax;
        r.sp = st.sp;
        donate_store(p.donate_slots, p.donate_ready, p.donate_epoch, (uint32_t)base + rank, r, fin);
        st.live = false;
        donated = true;
    }
}

// Called by one lane right after it has made ITS contribution visible (finished++ or served++): if that was the event
// that made the queue quiescent, announce it.  finished <= started and served <= tail at all times and all four only
// grow; the counters share one L2 line (one order of events), and served is read before tail, so equality means: when
// tail was read, every warp that had entered the kernel had left its main loop and nobody was resuming a ray.  Warps
// that enter later find the work counter exhausted and run through here themselves.
__device__ __forceinline__ void donate_check_quiescent(const TraceParams &p)
{
    __threadfence();
    const unsigned long long done = ld_volatile_u64(p.counter + kDonateFinished);
    if (done != ld_volatile_u64(p.counter + kDonateStarted)) {
        return;
    }
    const unsigned long long served = ld_volatile_u64(p.counter + kDonateServed);
    __threadfence();
    if (served == ld_volatile_u64(p.counter + kDonateTail)) {
        *reinterpret_cast<volatile unsigned long long *>(p.counter + kDonateQuiet) = 1ull;
    }
}

// The answer of an any-hit ray as far as it is known when its first piece forks: "nothing blocks" / "miss".
__device__ __forceinline__ void write_optimistic(const TraceParams &p, uint32_t kind, uint64_t out, uint64_t mirror, const float pre[4])
{
    if (kind == kFinishVisible) {
        p.visible[out] = 1;
        if (mirror != kNoMirror) p.mirror_visible[mirror] = 1;
    } else { // kFinishAnyRecord
        const float4 r = make_float4(pre[0], __uint_as_float(DODRT_MISS), 0.0f, 0.0f);
        reinterpret_cast<float4 *>(p.hits)[out] = r;
        if (mirror != kNoMirror) reinterpret_cast<float4 *>(p.mirror_hits)[mirror] = r;
    }
}

// Has another piece of this (forked, any-hit) ray already found a hit?
__device__ __forceinline__ bool answer_decided(const TraceParams &p, uint32_t kind, uint64_t out)
{
    if (kind == kFinishVisible) {
        return *reinterpret_cast<const volatile uint8_t *>(p.visible + out) == 0;
    }
    return *reinterpret_cast<const volatile uint32_t *>(reinterpret_cast<const uint32_t *>(p.hits + out) + 1) != DODRT_MISS;
}

// Helper side: the whole warp resumes the ray (or the piece of an any-hit ray) in `slot` and writes its result.
__device__ __forceinline__ float pick4(const float4 &q, int j) { return j == 0 ? q.x : j == 1 ? q.y : j == 2 ? q.z : q.w; }

__device__ __forceinline__ void resume_ray(const TraceParams &p, uint32_t slot)
{
    const DeviceScene &s = p.scene;
    const uint32_t lane = threadIdx.x & 31u;
    const float4 *w = reinterpret_cast<const float4 *>(p.donate_slots + (size_t)slot * kDonateSlotWords);
    const float4 w0 = __ldcg(w), w1 = __ldcg(w + 1), w2 = __ldcg(w + 2), w3 = __ldcg(w + 3), w4 = __ldcg(w + 4),
                 w5 = __ldcg(w + 5);
    const float o[3] = {w0.x, w0.y, w0.z}, d[3] = {w0.w, w1.x, w1.y};
    float clip = w1.z;
    const uint32_t flags = __float_as_uint(w1.w);
    const bool any = (flags & kFlagAny) != 0u;
    bool found = (flags & kFlagFound) != 0u;
    bool forked = (flags & kFlagForked) != 0u;
    const uint32_t kind = flags >> 3;
    Hit hit;
    hit.t = w2.x, hit.prim = __float_as_uint(w2.y), hit.u = w2.z, hit.v = w2.w;
    const float pre[4] = {w3.x, w3.y, w3.z, w3.w};
    TreeState st;
    st.inv[0] = 1.0f / d[0]; // kdtree.cpp:271, the same division the donor did
    st.inv[1] = 1.0f / d[1];
    st.inv[2] = 1.0f / d[2];
    st.tmin = w4.x, st.tmax = w4.y, st.node = __float_as_uint(w4.z), st.triCur = __float_as_uint(w4.w);
    st.triEnd = __float_as_uint(w5.x);
    st.sp = (int)__float_as_uint(w5.y);
    st.live = true;
    const uint64_t out = (uint64_t)__float_as_uint(w5.z) | ((uint64_t)__float_as_uint(w5.w) << 32);
    uint32_t stackNode[kMaxStack];
    float stackTmin[kMaxStack];
    float stackTmax[kMaxStack];
    const uint32_t *stk = p.donate_slots + (size_t)slot * kDonateSlotWords + 24;
    const uint64_t mirror = (uint64_t)__ldcg(stk + kDonateMirrorWord - 24) | ((uint64_t)__ldcg(stk + kDonateMirrorWord - 24 + 1) << 32);
    for (int i = 0; i < st.sp; i++) {
        stackNode[i] = __ldcg(stk + 3 * i);
        stackTmin[i] = __uint_as_float(__ldcg(stk + 3 * i + 1));
        stackTmax[i] = __uint_as_float(__ldcg(stk + 3 * i + 2));
    }
    uint32_t sincePoll = 0;
#ifdef DODRT_TIMELINE
    unsigned long long tlNode = 0, tlLeaf = 0, tlNodeSteps = 0, tlLeafSteps = 0;
#endif
    while (st.live) { // warp-uniform: every lane holds the same state
#ifdef DODRT_TIMELINE
        const long long tl0 = clock64();
        const bool tlIsLeaf = st.triCur < st.triEnd;
#endif
        if (st.triCur < st.triEnd) {
            // 128 triangle slots per step: lane L tests slots triCur + 4L .. 4L + 3, which are four adjacent columns of ONE
            // SoA triangle lane (triangle.h:33-44: 9 rows of 8 floats), so nine 16-byte loads fetch all four triangles
            // and a leaf of up to 16 triangle lanes is one step (round 1 tested 32 slots per step with nine 4-byte loads
            // per lane: 8.7 steps per resumed ray against 4.0 now; profiles/r02_donation_fork.txt, section 6).
            const uint32_t first = st.triCur + 4u * lane; // triCur, triEnd are multiples of 8: all four slots are in or out
            const bool in = first < st.triEnd;
            float4 r[9];
#pragma unroll
            for (int k = 0; k < 9; k++) {
                r[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
            if (in) {
                const float4 *base = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(s.lanes4) +
                                                                      (size_t)(first >> 3) * 72 + (first & 7u));
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    r[k] = __ldg(base + 2 * k);
                }
            }
            // First stage for the four slots at once: every conservative compare of triangle_test_fast, branch-free, four
            // independent dependency chains (a warp that works on one ray has no other source of instruction-level
            // parallelism).  The exact test, with its division, runs only for what is almost certainly a hit, lowest
            // slot first, and strict `<` keeps the lower slot on a tie (triangle.cpp:133).  All slots are compared
            // with the clip at the START of the step, then reduced with the lexicographic (t, slot) minimum: what the
            // reference's running clip leaves behind.
            uint32_t m = 0;
            if (in) {
                m = (uint32_t)triangle_may_hit_full(r[0].x, r[1].x, r[2].x, r[3].x, r[4].x, r[5].x, r[6].x, r[7].x, r[8].x, o, d) |
                    ((uint32_t)triangle_may_hit_full(r[0].y, r[1].y, r[2].y, r[3].y, r[4].y, r[5].y, r[6].y, r[7].y, r[8].y, o, d) << 1) |
                    ((uint32_t)triangle_may_hit_full(r[0].z, r[1].z, r[2].z, r[3].z, r[4].z, r[5].z, r[6].z, r[7].z, r[8].z, o, d) << 2) |
                    ((uint32_t)triangle_may_hit_full(r[0].w, r[1].w, r[2].w, r[3].w, r[4].w, r[5].w, r[6].w, r[7].w, r[8].w, o, d) << 3);
            }
            float t = 0.0f, u = 0.0f, v = 0.0f;
            uint32_t slotId = first;
            bool acc = false;
            while (m != 0u) {
                const int j = __ffs((int)m) - 1;
                m &= m - 1u;
                const float4 q0 = make_float4(pick4(r[0], j), pick4(r[1], j), pick4(r[2], j), pick4(r[3], j));
                const float4 q1 = make_float4(pick4(r[4], j), pick4(r[5], j), pick4(r[6], j), pick4(r[7], j));
                const float4 q2 = make_float4(pick4(r[8], j), 0.0f, 0.0f, 0.0f);
                float tj = 0.0f, uj = 0.0f, vj = 0.0f;
                if (triangle_test_fast(q0, q1, q2, o, d, clip, tj, uj, vj) && (!acc || tj < t)) {
                    t = tj, u = uj, v = vj;
                    slotId = first + (uint32_t)j;
                    acc = true;
                }
            }
            const unsigned accMask = __ballot_sync(0xffffffffu, acc);
            if (accMask != 0u) {
                float best = acc ? t : kInfinity; // accepted t are finite and positive: min() is exact
#pragma unroll
                for (uint32_t off = 16; off != 0u; off >>= 1) {
                    best = fminf(best, __shfl_xor_sync(0xffffffffu, best, off));
                }
                // lowest slot id among equal t = what the reference's slot-by-slot loop leaves behind
                const uint32_t winSlot = __reduce_min_sync(0xffffffffu, (acc && t == best) ? slotId : 0xFFFFFFFFu);
                const unsigned winners = __ballot_sync(0xffffffffu, acc && t == best && slotId == winSlot);
                const uint32_t src = (uint32_t)__ffs((int)winners) - 1u;
                clip = best;
                hit.t = best;
                hit.prim = (DODRT_KIND_TRIANGLE << DODRT_KIND_SHIFT) | winSlot;
                hit.u = __shfl_sync(0xffffffffu, u, src);
                hit.v = __shfl_sync(0xffffffffu, v, src);
                found = true;
            }
            st.triCur = st.triCur + 128u < st.triEnd ? st.triCur + 128u : st.triEnd;
            if (any && found) {
                st.live = false; // kdtree.cpp:338-341
            } else if (st.triCur == st.triEnd) {
                tree_pop(st, stackNode, stackTmin, stackTmax);
            }
        } else {
            node_step(s, st, o, d, clip, stackNode, stackTmin, stackTmax);
        }
#ifdef DODRT_TIMELINE
        if (tlIsLeaf) {
            tlLeaf += (unsigned long long)(clock64() - tl0);
            tlLeafSteps++;
        } else {
            tlNode += (unsigned long long)(clock64() - tl0);
            tlNodeSteps++;
        }
#endif
#ifdef DODRT_EXPERIMENTS
        // ---- work splitting (any-hit rays only): give the oldest stack entries to helpers that wait without a ray
        if (any && st.live && s.fork_poll != 0u && ++sincePoll >= s.fork_poll) {
            sincePoll = 0;
            uint32_t m = 0;
            bool decided = false;
            if (lane == 0) {
                if (forked) {
                    decided = answer_decided(p, kind, out);
                }
                if (!decided && st.sp > 0) {
                    const unsigned long long head = ld_volatile_u64(p.counter + kDonateHead);
                    const unsigned long long tail = ld_volatile_u64(p.counter + kDonateTail);
                    // (same bound as the donors: every warp adds at most 32 slots per check, capacity = 2 x threads)
                    if (tail <= p.donate_capacity / 2u) {
                        if (p.scene.tune[3] != 0u) {
                            m = (uint32_t)st.sp; // DODRT_DONATE_ALWAYS (tests): fork everything, helpers waiting or not
                        } else if (head > tail) {
                            const unsigned long long waiting = head - tail;
                            m = (uint32_t)(waiting < (unsigned long long)st.sp ? waiting : (unsigned long long)st.sp);
                        }
                    }
                }
            }
            decided = __shfl_sync(0xffffffffu, decided ? 1 : 0, 0) != 0;
            m = __shfl_sync(0xffffffffu, m, 0);
            if (decided) {
                st.live = false; // a sibling piece found a hit: the OR is true whatever is left here
                break;
            }
            if (m != 0u) {
                if (!forked) { // first fork of this ray: the optimistic answer goes out before any child can answer
                    if (lane == 0) {
                        write_optimistic(p, kind, out, mirror, pre);
                        __threadfence();
                    }
                    forked = true;
                }
                unsigned long long base = 0;
                if (lane == 0) {
                    if (mirror != kNoMirror) {
                        __threadfence_system(); // the mirror lives on a peer GPU / in host memory
                    }
                    base = atomicAdd(p.counter + kDonateTail, (unsigned long long)m);
                }
                base = __shfl_sync(0xffffffffu, base, 0);
                __syncwarp(); // lane 0's optimistic answer is ordered before any child's publication
                if (lane < m) { // child `lane` = stack entry `lane` (the oldest entries: the largest sub-trees)
                    const uint32_t child = (uint32_t)base + lane;
                    float4 *cw = reinterpret_cast<float4 *>(p.donate_slots + (size_t)child * kDonateSlotWords);
                    cw[0] = w0;
                    cw[1] = make_float4(d[1], d[2], clip, __uint_as_float(kFlagAny | kFlagForked | (kind << 3)));
                    cw[2] = make_float4(clip, __uint_as_float(DODRT_MISS), 0.0f, 0.0f);
                    cw[3] = w3;
                    cw[4] = make_float4(stackTmin[lane], stackTmax[lane], __uint_as_float(stackNode[lane]), __uint_as_float(0u));
                    cw[5] = make_float4(__uint_as_float(0u), __uint_as_float(0u), w5.z, w5.w);
                    uint32_t *cs = p.donate_slots + (size_t)child * kDonateSlotWords + 24;
                    cs[kDonateMirrorWord - 24] = (uint32_t)mirror;
                    cs[kDonateMirrorWord - 24 + 1] = (uint32_t)(mirror >> 32);
                    __threadfence();
                    *reinterpret_cast<volatile uint32_t *>(p.donate_ready + child) = p.donate_epoch;
                }
                for (int i = 0; i + (int)m < st.sp; i++) { // keep the younger entries
                    stackNode[i] = stackNode[i + m];
                    stackTmin[i] = stackTmin[i + m];
                    stackTmax[i] = stackTmax[i + m];
                }
                st.sp -= (int)m;
            }
        }
#endif // DODRT_EXPERIMENTS (work splitting)
    }
    (void)sincePoll;
#ifdef DODRT_TIMELINE
    if (lane == 0) { // debug build: where does a resumed ray spend its cycles?  [28] node cycles [29] leaf cycles [30] node steps [31] leaf steps
        atomicAdd(p.counter + 28, tlNode);
        atomicAdd(p.counter + 29, tlLeaf);
        atomicAdd(p.counter + 30, tlNodeSteps);
        atomicAdd(p.counter + 31, tlLeafSteps);
    }
#endif
    if (lane == 0) {
        if (forked) { // pieces of a forked ray only ever turn the optimistic answer into "blocked" / "hit"
            if (found) {
                finish_write(p, kind, out, mirror, true, hit, pre);
            }
        } else {
            finish_write(p, kind, out, mirror, found, hit, pre);
        }
        __threadfence();
        atomicAdd(p.counter + kDonateServed, 1ull); // this slot is completely served (see donate_helper_loop)
        donate_check_quiescent(p);
    }
}

// Back-off of a waiting helper between two looks at its ready word (nanosleep, doubling).  1-of-8 share of dragon4k, rank 0 /
// rank 1, ms: cap 400 ns 0.701 / 0.748, 800 ns 0.692 / 0.741, 3200 ns (round 1) 0.702 / 0.745, 12800 ns 0.734 / 0.775.
constexpr unsigned kHelperSleepMinNs = 100u, kHelperSleepCapNs = 800u;

// Runs after the main loop of a warp: count this warp as finished, then serve the queue -- take a ticket, wait for that
// slot to be filled, resume the ray -- until every warp has left its main loop and no slot at or beyond the ticket
// was reserved.
__device__ __forceinline__ void donate_helper_loop(const TraceParams &p)
{
    const uint32_t lane = threadIdx.x & 31u;
    if (lane == 0) {
        __threadfence(); // this warp's donations (if any) are published before it counts as finished
        atomicAdd(p.counter + kDonateFinished, 1ull);
        donate_check_quiescent(p);
    }
    for (;;) {
        uint32_t ticket = 0;
        if (lane == 0) {
            // enough helpers wait already: this warp would only add polling traffic
            const unsigned long long head = ld_volatile_u64(p.counter + kDonateHead);
            const unsigned long long tail = ld_volatile_u64(p.counter + kDonateTail);
            ticket = (head > tail && head - tail >= p.scene.helper_limit) ? 0xFFFFFFFFu
                                                                         : (uint32_t)atomicAdd(p.counter + kDonateHead, 1ull);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket >= p.donate_capacity) {
            break; // over the helper limit, or (cannot happen while donors respect the capacity bound) past the queue
        }
        uint32_t ready = 0;
        if (lane == 0) {
            unsigned ns = kHelperSleepMinNs;
            for (;;) {
                if (*reinterpret_cast<const volatile uint32_t *>(p.donate_ready + ticket) == p.donate_epoch) {
                    ready = 1;
                    break;
                }
                // Quiescent (see donate_check_quiescent): no warp is in its main loop -- counted when they ENTER the
                // kernel, not taken from the grid size: blocks that are not resident yet (the SMs may be shared with
                // another kernel, e.g. a second donating launch on another stream) must not be waited for -- and no
                // helper is resuming a ray, so no slot will ever be reserved again: tail is final.
                if (ld_volatile_u64(p.counter + kDonateQuiet) != 0ull) {
                    __threadfence();
                    if (ld_volatile_u64(p.counter + kDonateTail) <= ticket) {
                        break;
                    }
                    // tail covers the ticket: the slot was reserved and is being filled
                }
                __nanosleep(ns);
                ns = ns < kHelperSleepCapNs ? ns * 2u : ns;
            }
        }
        ready = __shfl_sync(0xffffffffu, ready, 0);
        if (!ready) {
            break;
        }
        __threadfence();
        resume_ray(p, ticket);
    }
}

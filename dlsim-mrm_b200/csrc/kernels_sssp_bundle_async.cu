// K1, origin-bundle search WITHOUT round barriers inside a label window (experiment: DTL_BUNDLE_ASYNC=1).
// Same labels, same node-major layout, same queue bits and far pile as sssp_bundle_kernel (kernels_sssp_bundle.cu); what
// changes is the near queue: a ring buffer in shared memory that the groups of a CTA consume while other groups still
// produce into it.  A round-synchronous pass has every warp in the same phase (an issue burst, then everybody waits for
// the same round trips, then a barrier); here the phases of different warps overlap.
//   * producers reserve a slot with one atomicAdd on the tail, write the row, fence, then publish node+1 into the slot;
//   * every group owns one ticket (atomicAdd on the head) and polls its slot; the ticket is kept until its entry shows up;
//   * the high bits of s_pt count entries that are queued or being expanded (its low bits are the ring tail); when it reaches zero the window is drained, all warps
//     leave the loop, and the CTA re-buckets the far pile exactly like the round-synchronous kernel;
//   * an idle warp that has polled for about a second without the window draining gives the bundle up (watchdog);
//   * the ring has far more slots (>= 1 024) than the CTA has ticket holders (128 groups), so a ticket can never meet the
//     slot of an older lap; a ring that wraps onto an unconsumed slot, or a far pile that overflows, aborts the bundle (its trees are recomputed
//     by the global-queue kernel, like any other overflow).
// Correctness does not depend on the order of expansions (see kernels_sssp_bundle.cu): the protocol "clear the node's queue
// bits, then read its label row" / "atomicMin, then set the bit and push" is unchanged.
#include <cstdio>
#include <cstdlib>

#include "sssp_bundle_common.cuh"

namespace dtl {

#define PT_SHIFT 17
#define PT_ONE_PENDING (1u << PT_SHIFT)
#define PT_TAIL_MASK (PT_ONE_PENDING - 1u)
#ifndef BND_ASYNC_SLEEP
#define BND_ASYNC_SLEEP 40 /* ns an idle warp sleeps between two polls of its ring slots */
#endif

// The label -> predecessor pass of the finished bundles (fuse_pred): bundles are taken in the order they finished, tiles of a
// bundle from a counter.  Run by the CTAs of the search kernel that have nothing (left) to search and by the CTAs of
// bundle_pred_helper_kernel; s_bundle / s_head are two shared ints of the calling CTA.
template <int BLOCK, int B>
__device__ __forceinline__ void bundle_pred_loop(const BundleLaunch& a, int* s_bundle, int* s_head)
{
    constexpr int TILE = BLOCK / (B / 2) * 8;
    const int tid = threadIdx.x;
    const int N = a.s.N;
    const int n_tiles = (N + TILE - 1) / TILE;
    int* const done = a.work + 1;
    int* const tile_counter = a.work + 1 + a.n_bundles;
    for (int i = 0; i < a.n_bundles; ++i) {
        __syncthreads();
        if (tid == 0) {
            int b;
            while ((b = *reinterpret_cast<volatile int*>(&done[i])) < 0) __nanosleep(400);
            *s_bundle = b;
        }
        __syncthreads();
        const int b = *s_bundle;
        __threadfence();
        for (;;) {
            __syncthreads();
            if (tid == 0) *s_head = atomicAdd(&tile_counter[b], 1);
            __syncthreads();
            const int tile = *s_head;
            if (tile >= n_tiles) break;
            bundle_pred_nodes<B, BLOCK, true>(a, b, tile * TILE, min(N, (tile + 1) * TILE), tile == 0);
        }
    }
}

// Helper CTAs for the label -> predecessor pass of a RUNNING search kernel (launched on another stream when the search kernel was
// given only as many CTAs as it has bundles, so that the SMs without a bundle can first serve that stream's other work -- the
// link-volume scatter).  A helper only joins in when every CTA of the search kernel is resident (work[1 + 2 * n_bundles] counts
// them): it then cannot keep a search CTA from being placed, and the search cannot wait for it.
template <int BLOCK, int B>
__global__ void __launch_bounds__(BLOCK, 1) bundle_pred_helper_kernel(BundleLaunch a)
{
    __shared__ int s_bundle, s_head, s_go;
    if (threadIdx.x == 0) {
        const volatile int* started = a.work + 1 + 2 * a.n_bundles;
        int go = 0;
        for (int spin = 0; spin < 256 && !(go = *started >= a.grid); ++spin) __nanosleep(200);
        s_go = go;
    }
    __syncthreads();
    if (!s_go) return;
    bundle_pred_loop<BLOCK, B>(a, &s_bundle, &s_head);
}

// EPL = out-edges relaxed per lane and pass (4, 2 or 1).  A ring entry (one node) is served by a group of
// G = (B / 2) * (4 / EPL) lanes: lane (es, op) of the group relaxes edges es*EPL .. es*EPL+EPL-1 of the node for origins 2*op and
// 2*op+1.  Only EPL = 4 is instantiated: on average a warp pass serves 0.9 nodes at the bench configuration (24.5 M passes for
// 21.7 M expansions), which suggested spreading a node over more lanes to shorten the serial instruction stream of a pass, but
// the ready nodes come in bursts and the number of GROUPS a CTA has is what counts then -- measured per 2 000 trees with 1 024
// threads: EPL = 4 (128 groups) 7.3 ms, EPL = 2 (64 groups) 7.7 ms, EPL = 1 (32 groups, one node per warp) 12.2 ms.
// BOUNDED: destination-bounded search (BundleLaunch::dest_ptr) -- after every window opening the smallest FRESH key of anything still
// queued is a lower bound on every label that can still change; when all destinations with demand lie below it the search stops.
template <int BLOCK, int B, int EPL, bool TAGS, int CPS, bool BOUNDED>
__global__ void __launch_bounds__(BLOCK, CPS) sssp_bundle_async_kernel(BundleLaunch a)
{
    constexpr int LPT = 2;
    constexpr int OPL = B / LPT;      // lanes that share an edge: one per pair of origins
    constexpr int ESG = BND_U / EPL;  // edge subgroups of a group
    constexpr int G = OPL * ESG;      // lanes per ring entry
    constexpr int OWN = OPL / EPL;    // lanes of an edge subgroup that end up holding the same edge's key
    static_assert(BND_U == 4 && (OPL == 4 || OPL == 8) && (EPL == 1 || EPL == 2 || EPL == 4) && G <= 32 && OWN >= 1,
                  "a pass relaxes four out-edges per group");
    extern __shared__ int s_dyn[];
    __shared__ int s_bundle, s_head, s_n_far, s_n_far2, s_overflow;
    // {entries queued or in flight (bits 17..31), ring tail of this window (bits 0..16)}: ONE shared-memory atomic per push --
    // same-address shared-memory atomics are what this kernel's throughput hangs on
    __shared__ unsigned s_pt;
    __shared__ unsigned s_red[32];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int g = tid & (G - 1);        // lane of the group: g == 0 pops the entry
    const int op = g & (OPL - 1);       // origin pair of this lane
    const int es = g / OPL;             // edge subgroup of this lane
    const int gfirst = lane & ~(G - 1);
    const int N = a.s.N;
    const int RING = a.cap_near; // power of two
    const int RMASK = RING - 1;
    const int CAPF = a.cap_far;
    const int n_words = (N + 15) >> 4;
    unsigned* const s_state = reinterpret_cast<unsigned*>(s_dyn);            // [n_words]
    int2* const s_r = reinterpret_cast<int2*>(s_dyn + ((n_words + 1) & ~1)); // [RING] forward-star row of the entry's node
    volatile int* const s_q = reinterpret_cast<int*>(s_r + RING);            // [RING] node + 1, 0 = empty slot
    int2* q_far = a.far + (size_t)blockIdx.x * 2 * CAPF;
    int2* q_far2 = q_far + CAPF;
    unsigned long long c_relax = 0;
    unsigned c_pops = 0, c_rounds = 0;
    if (a.fuse_pred && tid == 0) { // resident (see bundle_pred_helper_kernel); the last one to arrive tells whoever waits for it
        const int r = atomicAdd(&a.work[1 + 2 * a.n_bundles], 1);
        if (a.started_word && r == (int)gridDim.x - 1) {
            __threadfence();
            *reinterpret_cast<volatile unsigned*>(a.started_word) = a.started_tag;
        }
    }
#ifdef SSSP_PROFILE
    if (tid == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMin(&a.s.stats->stage[4], now);
    }
#endif

    // publish one entry into the ring (any thread)
    auto push_near = [&](int node, int2 row) {
        const unsigned pt = atomicAdd(&s_pt, PT_ONE_PENDING + 1u);
        const int queued = (int)(pt >> PT_SHIFT);
        const int tail = (int)(pt & PT_TAIL_MASK);
        const int slot = tail & RMASK;
        // the ring wrapped onto an unconsumed entry, the test hook's limit, or the tail field is about to run out of bits
        if (s_q[slot] != 0 || queued >= a.near_limit || tail >= (int)PT_TAIL_MASK - 4096) {
            s_overflow = 1;
            atomicSub(&s_pt, PT_ONE_PENDING);
            return;
        }
        s_r[slot] = row;
        __threadfence_block();
        s_q[slot] = node + 1;
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) s_bundle = atomicAdd(a.bundle_counter, 1);
        __syncthreads();
        const int bundle = s_bundle;
        if (bundle >= a.n_bundles) break;
#ifdef SSSP_PROFILE
        const long long pf_start = clock64(); // per-bundle search time: stage[1] = max, stage[2] = sum, stage[3] = bundles
        long long pf_t = pf_start, pf_fill = 0, pf_min = 0, pf_dist = 0, pf_drain = 0, pf_nfar = 0; // thread 0's cycles per phase
#define APF(x_) do { const long long n_ = clock64(); x_ += n_ - pf_t; pf_t = n_; } while (0)
#else
#define APF(x_) do { } while (0)
#endif
        int zone[LPT];
        unsigned has_task = 0;
#pragma unroll
        for (int t = 0; t < LPT; ++t) {
            const int task = __ldg(&a.bundle_tasks[bundle * B + op * LPT + t]);
            zone[t] = TAGS && task >= 0 ? __ldg(&a.s.task_zone[task]) : -1;
            has_task |= (unsigned)(task >= 0) << t;
        }
        unsigned long long* const bl = a.bl + (size_t)bundle * N * B;
        unsigned long long* const blg = bl + op * LPT;

        {
            ulonglong2* v = reinterpret_cast<ulonglong2*>(bl);
            const size_t nv = (size_t)N * B / 2;
            const ulonglong2 x = make_ulonglong2(INF_BITS, INF_BITS);
            for (size_t i = tid; i < nv; i += BLOCK) v[i] = x;
        }
        for (int i = tid; i < n_words; i += BLOCK) s_state[i] = 0u;
        for (int i = tid; i < RING; i += BLOCK) s_q[i] = 0;
        if (tid == 0) {
            s_head = 0; s_pt = 0u; s_n_far = 0; s_n_far2 = 0; s_overflow = 0;
            if (a.bundle_bound) a.bundle_bound[bundle] = BND_DEAD_KEY; // unless the search stops early, every label is final
        }
        __syncthreads();
        if (tid < B) { // seed: every origin relaxes its own (unhidden) row once with label 0; the heads wait in the far pile
            const int task = __ldg(&a.bundle_tasks[bundle * B + tid]);
            if (task >= 0) {
                const int origin = __ldg(&a.s.task_origin[task]);
                const int oz = __ldg(&a.s.task_zone[task]);
                bl[(size_t)origin * B + tid] = 0ULL;
                a.s.tie_nodes[task] = 0;
                const int r1 = __ldg(&a.s.row_ptr[origin + 1]);
                for (int e = __ldg(&a.s.row_ptr[origin]); e < r1; ++e) {
                    const Edge ed = ld_edge(&a.s.edges[e]);
                    int to = ed.to;
                    if (to < 0) {
                        to &= 0x7fffffff;
                        if (__ldg(&a.s.link_tag[ed.link]) != oz) continue;
                    }
                    if (to == origin) continue;
                    const double c = 0.0 + ed.cost;
                    const unsigned long long cb = (unsigned long long)__double_as_longlong(c);
                    const unsigned long long old = atomicMin(&bl[(size_t)to * B + tid], cb);
                    ++c_relax;
                    if (old > cb) {
                        const unsigned o = atomicOr(&s_state[ST_WORD(to)], 2u << ST_SHIFT(to));
                        if (!((o >> ST_SHIFT(to)) & 3u)) {
                            const int pos = atomicAdd(&s_n_far, 1);
                            if (pos < CAPF) q_far[pos] = make_int2(to, (int)key_of(cb));
                            else s_overflow = 1;
                        }
                    }
                }
            }
        }
        __syncthreads();
        unsigned thr = 0u;
        bool failed = false;
        APF(pf_fill);

        for (;;) {
            // ---- open the next label window from the far pile (block-wide, like the round-synchronous kernel) ----
            const int n_far = min(s_n_far, CAPF);
            if (n_far == 0) break;
            unsigned mm = BND_DEAD_KEY;
            for (int i = tid; i < n_far; i += BLOCK) {
                const int2 q = q_far[i];
                if ((s_state[ST_WORD(q.x)] >> ST_SHIFT(q.x)) & 2u) mm = min(mm, (unsigned)q.y);
            }
            mm = block_min_u32<BLOCK>(mm, s_red);
#ifdef SSSP_PROFILE
            pf_nfar += n_far;
#endif
            APF(pf_min);
            if (mm == BND_DEAD_KEY) break;
            thr = key_of((unsigned long long)__double_as_longlong(__hiloint2double((int)mm, 0) + a.s.delta));
            if (thr <= mm) thr = mm + 1;
            if (tid == 0) { s_head = 0; s_pt = 0u; } // the ring is empty here: every published entry was consumed and finished
            __syncthreads();
            unsigned lb = BND_DEAD_KEY; // BOUNDED: smallest fresh key of the live entries
            // one far entry per OPL lanes (the lanes of a group's first edge subgroup suffice: a label row is OPL 16-byte loads)
            for (int base = 0; base < n_far; base += BLOCK / OPL) {
                if (base + (tid & ~31) / OPL >= n_far) continue;
                const int i = base + tid / OPL;
                const bool have = i < n_far;
                const int v = have ? q_far[i].x : 0;
                const int w = ST_WORD(v), sh = ST_SHIFT(v);
                const int ofirst = lane & ~(OPL - 1);
                unsigned o = 0;
                if (have && op == 0) o = atomicAnd(&s_state[w], ~(2u << sh)) >> sh;
                o = __shfl_sync(0xffffffffu, o, ofirst);
                const bool live = (o & 2u) != 0;
                unsigned k = BND_DEAD_KEY;
                if (live) {
                    const ulonglong2 lab = __ldcg(reinterpret_cast<const ulonglong2*>(blg + (size_t)v * B));
                    k = min(key_of(lab.x), key_of(lab.y));
                }
                const unsigned key = group_min_u32<OPL>(k);
                if (BOUNDED && live) lb = min(lb, key);
                if (live && op == 0) {
                    if (key < thr) {
                        const unsigned o2 = atomicOr(&s_state[w], 1u << sh);
                        if (!((o2 >> sh) & 1u)) push_near(v, __ldg(&a.s.rows[v]));
                    } else {
                        atomicOr(&s_state[w], 2u << sh);
                        const int pos = atomicAdd(&s_n_far2, 1);
                        q_far2[pos] = make_int2(v, (int)key);
                    }
                }
            }
            __syncthreads();
            if (tid == 0) { s_n_far = s_n_far2; s_n_far2 = 0; }
            int2* t2 = q_far; q_far = q_far2; q_far2 = t2;
            ++c_rounds;
            __syncthreads();
            APF(pf_dist);
            if (s_overflow) { failed = true; break; }
            if (BOUNDED && a.dest_ptr) {
                lb = block_min_u32<BLOCK>(lb, s_red);
                unsigned far_dest = 0; // largest key of a destination label (an unreached destination keeps the search going)
                const int d1 = __ldg(&a.dest_ptr[bundle + 1]);
                for (int i = __ldg(&a.dest_ptr[bundle]) + tid; i < d1; i += BLOCK)
                    far_dest = max(far_dest, key_of(__ldcg(&bl[__ldg(&a.dest_idx[i])])));
                far_dest = ~block_min_u32<BLOCK>(~far_dest, s_red);
                if (far_dest < lb) { // nothing queued can reach below lb: every destination label is final
                    if (tid == 0) a.bundle_bound[bundle] = lb;
                    break;
                }
            }

            // ---- drain the window: every group consumes ring entries until nothing is queued or in flight ----
            int ticket = -1; // this group's claim on a ring position (leader's register, broadcast when needed)
            unsigned idle_polls = 0;
            for (;;) {
                if (g == 0 && ticket < 0) ticket = atomicAdd(&s_head, 1);
                int node1 = 0;
                bool reserved = false; // a producer holds this group's slot and is about to publish into it
                if (g == 0) {
                    node1 = s_q[ticket & RMASK];
                    reserved = node1 == 0 && ticket < (int)(*reinterpret_cast<volatile unsigned*>(&s_pt) & PT_TAIL_MASK);
                }
                const unsigned ready = __ballot_sync(0xffffffffu, node1 != 0);
                const unsigned soon = __ballot_sync(0xffffffffu, reserved);
                if (ready == 0u || soon != 0u) {
                    // nothing for this warp yet: leave when the window is drained or the bundle is aborted, otherwise poll again.
                    // With some entries ready and others about to be published the warp polls once more, so that one pass of
                    // the instruction stream below serves all of them.
                    const int pend = (int)(*reinterpret_cast<volatile unsigned*>(&s_pt) >> PT_SHIFT);
                    const int ovf = *reinterpret_cast<volatile int*>(&s_overflow);
                    if (ovf || (ready == 0u && pend == 0)) break;
                    if (soon == 0u) __nanosleep(BND_ASYNC_SLEEP);
                    // watchdog: a window drains in micro- to milliseconds; a warp that has polled for about a second gives the
                    // bundle up (its trees are recomputed by the global-queue kernel) instead of spinning on a lost wake-up
                    if (++idle_polls > (1u << 24)) s_overflow = 1;
                    continue;
                }
                idle_polls = 0;
                const bool have = __shfl_sync(0xffffffffu, node1, gfirst) != 0;
                int u = 0;
                int2 row = make_int2(0, 0);
                unsigned was = 0;
                if (have && g == 0) {
                    u = node1 - 1;
                    const volatile int* rr = reinterpret_cast<const volatile int*>(&s_r[ticket & RMASK]);
                    row.x = rr[0]; row.y = rr[1];
                    // the slot is free again (no fence: shared-memory accesses of a thread are performed in order, and the slot is
                    // not handed out again before the tail has gone round the whole ring)
                    s_q[ticket & RMASK] = 0;
                    ticket = -1;
                    was = atomicAnd(&s_state[ST_WORD(u)], ~(3u << ST_SHIFT(u))) >> ST_SHIFT(u);
                }
                u = __shfl_sync(0xffffffffu, u, gfirst);
                int e0 = __shfl_sync(0xffffffffu, row.x, gfirst) + es * EPL; // first edge of this lane's subgroup
                const int re = __shfl_sync(0xffffffffu, row.y, gfirst);
                // edge records are loaded unconditionally (slots past the end of the row repeat its last edge and are masked out
                // of the atomics): no predicated register moves in the instruction stream
                const int elast = max(re - 1, 0);
                double cost[EPL];
                int to[EPL];
#pragma unroll
                for (int j = 0; j < EPL; ++j) {
                    const int4 v = __ldg(reinterpret_cast<const int4*>(&a.s.edges[min(e0 + j, elast)]));
                    cost[j] = __hiloint2double(v.y, v.x); to[j] = v.z;
                }
                was = __shfl_sync(0xffffffffu, was, gfirst) & 3u;
                double lu[LPT];
                bool act[LPT];
                {
                    ulonglong2 lab = make_ulonglong2(INF_BITS, INF_BITS);
                    if (was) lab = __ldcg(reinterpret_cast<const ulonglong2*>(blg + (size_t)u * B));
                    lu[0] = __longlong_as_double((long long)lab.x);
                    lu[1] = __longlong_as_double((long long)lab.y);
                    act[0] = lab.x < INF_BITS && (has_task & 1u);
                    act[1] = lab.y < INF_BITS && (has_task & 2u);
                }
                c_pops += (unsigned)(have && g == 0);
                if (es == 0) c_relax += ((unsigned)act[0] + (unsigned)act[1]) * (unsigned)max(re - e0, 0);
                // One pass relaxes up to BND_U out-edges of the group's node, EPL of them per lane, for the lane's two origins: all
                // atomics are issued first (an atomic that is not issued "returns" its own candidate, i.e. no improvement); when
                // they have returned, the lanes of an edge subgroup decide per edge whether the head improved for any origin and
                // what its key is now, and one lane per edge queues the head.
                for (;;) {
                    unsigned long long old[EPL][LPT];
#pragma unroll
                    for (int j = 0; j < EPL; ++j) {
                        bool ok[LPT];
#pragma unroll
                        for (int t = 0; t < LPT; ++t) ok[t] = act[t] && e0 + j < re;
                        if (TAGS) {
                            if (to[j] < 0) { // routing.h:779-786
                                const int tag = __ldg(&a.s.link_tag[__ldg(&a.s.edges[min(e0 + j, elast)].link)]);
#pragma unroll
                                for (int t = 0; t < LPT; ++t) ok[t] = ok[t] && tag == zone[t];
                            }
                            to[j] &= 0x7fffffff;
                        }
                        unsigned long long* const p = blg + (size_t)to[j] * B;
#pragma unroll
                        for (int t = 0; t < LPT; ++t)
                            old[j][t] = atom_min_if(p + t, (unsigned long long)__double_as_longlong(lu[t] + cost[j]), ok[t]); // routing.h:804,816
                    }
                    // the edge this lane queues: its head and the head's forward-star row (requesting the row before the atomics
                    // instead -- so that it arrives in their shadow -- measured no difference: 7.869 against 7.873 ms)
                    const int je = op / OWN; // index among this lane's EPL edges
                    int myto = to[0];
#pragma unroll
                    for (int j = 1; j < EPL; ++j) myto = je == j ? to[j] : myto;
                    const bool owner = (op % OWN) == 0 && e0 + je < re;
                    int2 prow = make_int2(0, 0);
                    if (owner) prow = __ldg(&a.s.rows[myto]);
                    unsigned impm = 0;
                    unsigned nk[EPL];
#pragma unroll
                    for (int j = 0; j < EPL; ++j) {
                        const unsigned long long cb0 = (unsigned long long)__double_as_longlong(lu[0] + cost[j]);
                        const unsigned long long cb1 = (unsigned long long)__double_as_longlong(lu[1] + cost[j]);
                        impm |= (unsigned)(old[j][0] > cb0 || old[j][1] > cb1) << j;
                        nk[j] = min(min(key_of(old[j][0]), key_of(cb0)), min(key_of(old[j][1]), key_of(cb1)));
                    }
#pragma unroll
                    for (int o = OPL / 2; o > 0; o >>= 1) impm |= __shfl_xor_sync(0xffffffffu, impm, o);
                    if (__any_sync(0xffffffffu, impm != 0u)) {
                        const unsigned key = subgroup_min_edges<OPL, EPL>(nk, op);
                        if (owner && ((impm >> je) & 1u)) {
                            const int w = ST_WORD(myto), sh = ST_SHIFT(myto);
                            if (key < thr) {
                                const unsigned o = atomicOr(&s_state[w], 1u << sh);
                                if (!((o >> sh) & 1u)) push_near(myto, prow);
                            } else {
                                const unsigned o = atomicOr(&s_state[w], 2u << sh);
                                if (!((o >> sh) & 3u)) {
                                    const int pos = atomicAdd(&s_n_far, 1);
                                    if (pos < CAPF) q_far[pos] = make_int2(myto, (int)key);
                                    else s_overflow = 1;
                                }
                            }
                        }
                    }
                    e0 += BND_U;
                    if (!__any_sync(0xffffffffu, e0 < re)) break; // rows of more than BND_U edges take another pass
#pragma unroll
                    for (int j = 0; j < EPL; ++j) {
                        const int4 v = __ldg(reinterpret_cast<const int4*>(&a.s.edges[min(e0 + j, elast)]));
                        cost[j] = __hiloint2double(v.y, v.x); to[j] = v.z;
                    }
                }
                __syncwarp();
                if (have && g == 0) atomicSub(&s_pt, PT_ONE_PENDING); // after every push of this expansion
            }
            __syncthreads();
            APF(pf_drain);
            if (s_overflow) { failed = true; break; }
        }
#ifdef SSSP_PROFILE
        if (tid == 0) {
            atomicAdd(&a.s.stats->prof[1], (unsigned long long)pf_fill);
            atomicAdd(&a.s.stats->prof[2], (unsigned long long)pf_min);
            atomicAdd(&a.s.stats->prof[3], (unsigned long long)pf_dist);
            atomicAdd(&a.s.stats->prof[4], (unsigned long long)pf_drain);
            atomicAdd(&a.s.stats->prof[5], (unsigned long long)pf_nfar);
            atomicAdd(&a.s.stats->prof[6], (unsigned long long)c_rounds);
            const unsigned long long dt = (unsigned long long)(clock64() - pf_start);
            atomicMax(&a.s.stats->stage[1], dt);
            atomicAdd(&a.s.stats->stage[2], dt);
            atomicAdd(&a.s.stats->stage[3], 1ULL);
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            atomicMax(&a.s.stats->stage[5], now); // the last search of the launch ended
            atomicMin(&a.s.stats->stage[7], now); // the first search ended
#ifdef BUNDLE_DUMP
            printf("bundle %d cycles %llu windows %u origin0 %d\n", bundle, dt, c_rounds, a.s.task_origin[a.bundle_tasks[bundle * B]]);
#endif
        }
#endif
        if (failed) {
            __syncthreads();
            if (tid < B) {
                const int task = __ldg(&a.bundle_tasks[bundle * B + tid]);
                if (task >= 0) a.s.tie_nodes[task] = -1;
            }
            if (tid == 0) atomicAdd(&a.s.stats->overflow, 1ULL);
        }
        if (a.fuse_pred) { // the bundle's labels are final: hand it to the label -> predecessor pass below (any CTA)
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicExch(&a.work[1 + atomicAdd(&a.work[0], 1)], bundle);
        }
    }
    // ---- label -> predecessor pass of the finished bundles, by every CTA that has nothing left to search: the CTAs beyond
    // the number of bundles from the start, the others as their bundle finishes (bundles differ by some 10 % in search time).
    if (a.fuse_pred) bundle_pred_loop<BLOCK, B>(a, &s_bundle, &s_head);
#ifdef SSSP_PROFILE
    if (tid == 0) { // stage[4] = first CTA started, [7] = first search ended, [5] = last search ended, [6] = last CTA left (ns, globaltimer)
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMax(&a.s.stats->stage[6], now);
    }
#endif
    for (int o = 16; o > 0; o >>= 1) {
        c_relax += __shfl_xor_sync(0xffffffffu, c_relax, o);
        c_pops += __shfl_xor_sync(0xffffffffu, c_pops, o);
    }
    if (lane == 0) {
        atomicAdd(&a.s.stats->relaxations, c_relax);
        atomicAdd(&a.s.stats->pops, (unsigned long long)c_pops);
    }
    if (tid == 0) atomicAdd(&a.s.stats->rounds, (unsigned long long)c_rounds);
}

size_t bundle_async_smem_bytes(int N, int ring)
{
    const size_t words = (((size_t)(N + 15) >> 4) + 1) & ~(size_t)1;
    return words * sizeof(unsigned) + (size_t)ring * (sizeof(int2) + sizeof(int));
}

template <int BLOCK, int B, int EPL, int CPS, bool BOUNDED>
static void async_launch_bd(const BundleLaunch& a, size_t smem, cudaStream_t st)
{
    // a.tagged: some node outside the hidden centroid rows has a zone-tagged out-link (routing.h:779-786 then has to be
    // evaluated inside the edge loop); ordinary networks only tag the connectors that leave a centroid
    if (a.tagged) {
        cudaFuncSetAttribute(sssp_bundle_async_kernel<BLOCK, B, EPL, true, CPS, BOUNDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(sssp_bundle_async_kernel<BLOCK, B, EPL, true, CPS, BOUNDED>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        sssp_bundle_async_kernel<BLOCK, B, EPL, true, CPS, BOUNDED><<<a.grid, BLOCK, smem, st>>>(a);
    } else {
        cudaFuncSetAttribute(sssp_bundle_async_kernel<BLOCK, B, EPL, false, CPS, BOUNDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(sssp_bundle_async_kernel<BLOCK, B, EPL, false, CPS, BOUNDED>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        static bool said = false;
        if (!said && getenv("DTL_BUNDLE_DEBUG")) {
            said = true;
            int occ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sssp_bundle_async_kernel<BLOCK, B, EPL, false, CPS, BOUNDED>, BLOCK, smem);
            fprintf(stderr, "async bundle search: %d threads x %d CTAs/SM asked, %d edges per lane, occupancy %d CTAs/SM, %zu B dynamic smem, grid %d, %d bundles of %d\n",
                    BLOCK, CPS, EPL, occ, smem, a.grid, a.n_bundles, B);
        }
        sssp_bundle_async_kernel<BLOCK, B, EPL, false, CPS, BOUNDED><<<a.grid, BLOCK, smem, st>>>(a);
    }
}

template <int BLOCK, int B, int EPL, int CPS>
static void async_launch(const BundleLaunch& a, size_t smem, cudaStream_t st)
{
    async_launch_bd<BLOCK, B, EPL, CPS, false>(a, smem, st);
}

template <int B, int EPL>
static void async_launch_b(const BundleLaunch& a, size_t smem, cudaStream_t st)
{
    if (a.dest_ptr && a.block == 768 && a.ctas_per_sm == 1) { // the destination-bounded search exists for the default CTA shape only
        async_launch_bd<768, B, EPL, 1, true>(a, smem, st);
        return;
    }
    // CTA size x resident CTAs per SM: one big CTA per SM, or two smaller ones (each with its own bundle, ring and queue bits)
    switch (a.block * 10 + a.ctas_per_sm) {
    case 5122: async_launch<512, B, EPL, 2>(a, smem, st); break;
    case 5121: async_launch<512, B, EPL, 1>(a, smem, st); break;
    case 7681: async_launch<768, B, EPL, 1>(a, smem, st); break;
    case 8961: async_launch<896, B, EPL, 1>(a, smem, st); break;
    case 8321: async_launch<832, B, EPL, 1>(a, smem, st); break;
    default: async_launch<1024, B, EPL, 1>(a, smem, st); break;
    }
}

template <int B>
static void helper_launch_b(const BundleLaunch& a, int grid, cudaStream_t st)
{
    switch (a.block) {
    case 512: bundle_pred_helper_kernel<512, B><<<grid, 512, 0, st>>>(a); break;
    case 768: bundle_pred_helper_kernel<768, B><<<grid, 768, 0, st>>>(a); break;
    case 832: bundle_pred_helper_kernel<832, B><<<grid, 832, 0, st>>>(a); break;
    case 896: bundle_pred_helper_kernel<896, B><<<grid, 896, 0, st>>>(a); break;
    default: bundle_pred_helper_kernel<1024, B><<<grid, 1024, 0, st>>>(a); break;
    }
}

// `grid` helper CTAs for the fused label -> predecessor pass of the search kernel launched with `a` (same CTA size: the tiles are
// shared); a.work must have been reset before either kernel starts
void launch_bundle_pred_helpers(const BundleLaunch& a, int grid, cudaStream_t st)
{
    if (!a.fuse_pred || a.n_bundles == 0 || grid <= 0) return;
    if (a.B == 8) helper_launch_b<8>(a, grid, st);
    else helper_launch_b<16>(a, grid, st);
}

void launch_sssp_bundle_async_search(const BundleLaunch& a, cudaStream_t st)
{
    if (a.n_bundles == 0) return;
    const size_t smem = bundle_async_smem_bytes(a.s.N, a.cap_near);
    if (a.B == 8) async_launch_b<8, 4>(a, smem, st);
    else async_launch_b<16, 4>(a, smem, st);
}

} // namespace dtl

// K1, origin-bundle variant: B neighbouring origins share ONE label-correcting search (reference routing.h:633-939 per
// origin).  The labels of a bundle are stored node-major, bl[node][B] (one 8*B-byte row per node: 128 B for B = 16), and a
// group of B lanes owns one queue entry = one NODE: lane b relaxes the node's out-edges for origin b.  Every label access
// of a relaxation is then one coalesced row instead of B scattered 8-byte accesses, and the queue, row-pointer and edge
// loads are shared by the B origins.
//
//   * Schedule.  A node's key is the minimum of its B labels; nodes with a key inside the current label window are
//     expanded round by round (all their "dirty" lanes at once, also lanes whose label lies beyond the window), the others
//     wait in a far pile (near/far label correcting as in kernels_sssp.cu, but with one shared clock per bundle).  Lanes
//     that met at a node travel together from there on, so away from the origins a node is expanded ~1.3-1.5 times for
//     ALL B origins (simulated on the config-5 grid; 1.15 times per origin in the single-origin kernels).
//   * Exactness.  A label only ever takes values "label of an upstream node + link cost" (left-to-right FP64 sums of real
//     paths) and only decreases (64-bit atomicMin on the bit pattern).  The search stops when no lane of any node is
//     dirty, i.e. when no edge can improve a label: the Bellman fixed point, which is unique because IEEE addition is
//     monotone.  Labels are therefore bit-identical to the reference's under any schedule.
//   * Two queue bits per node in SHARED memory: "has an entry in the near queue of this or the next round" and "has an
//     entry in the far pile".  Whoever sets the first of them pushes the entry (at most one live entry per node); the
//     expander clears both BEFORE it requests the node's label row, so a concurrent improvement is either seen by this
//     expansion or finds the bits clear and queues the node again.  An expansion relaxes every lane that has reached the
//     node.  The dependent chain of a round is two L2 round trips: label row (edge records in its shadow: queue entries
//     carry the forward-star row) -> atomicMin; everything else is shared memory.
//   * Predecessors are not tracked in flight: bundle_finish_kernel derives them from the final labels (the reference's
//     predecessor link is the unique tight in-edge; two tight upstream nodes = an order-dependent tie, marked for the
//     serial replay exactly like in the other kernels) while it transposes the labels into the per-tree layout.
#include <cstdlib>

#include "sssp_bundle_common.cuh"

namespace dtl {

// B origins per bundle, LPT per thread: a group of G = B/LPT threads owns one queue entry (a node), thread g holds the labels
// of origins g*LPT .. g*LPT+LPT-1 (16-byte loads).  The instruction stream of a pass (queue pop, edge records, tag checks, row
// prefetch, push logic) is shared by 32/G entries, so more origins per thread means fewer instructions per expanded node.
template <int BLOCK, int B, int LPT>
__global__ void __launch_bounds__(BLOCK, BLOCK <= 512 && LPT <= 2 ? 2 : 1) sssp_bundle_kernel(BundleLaunch a)
{
    constexpr int G = B / LPT;
    constexpr int PR = (BND_U + G - 1) / G; // forward-star rows of the heads prefetched per thread (edge j by thread j % G)
    extern __shared__ int s_dyn[];
    __shared__ int s_bundle, s_cnt[3], s_n_far, s_n_far2, s_overflow;
    __shared__ unsigned s_red[32];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int g = tid & (G - 1);        // this thread holds origins g*LPT .. g*LPT+LPT-1 of its bundle
    const int gfirst = lane & ~(G - 1); // first warp lane of this thread's group
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << gfirst;
    const int N = a.s.N;
    const int CAP = a.cap_near;
    const int CAPF = a.cap_far;
    const int n_words = (N + 15) >> 4;
    unsigned* const s_state = reinterpret_cast<unsigned*>(s_dyn);            // [n_words]
    int2* const s_r = reinterpret_cast<int2*>(s_dyn + ((n_words + 1) & ~1)); // [2][CAP] forward-star row of the entry's node
    int* const s_q = reinterpret_cast<int*>(s_r + 2 * CAP);                  // [2][CAP] node
    int2* q_far = a.far + (size_t)blockIdx.x * 2 * CAPF; // {node, key}
    int2* q_far2 = q_far + CAPF;
    unsigned long long c_relax = 0;
    unsigned c_pops = 0, c_rounds = 0;
#ifdef SSSP_PROFILE
    // thread 0's cycles: [1] init + seed [2] expansion [3] round barrier [4] re-bucketing; [6] passes of warp 0 [7] entries read
    long long pf_t = clock64(), pf[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#define PF(k_) do { const long long n_ = clock64(); pf[k_] += n_ - pf_t; pf_t = n_; } while (0)
#else
#define PF(k_) do { } while (0)
#endif

    for (;;) {
        __syncthreads();
        if (tid == 0) s_bundle = atomicAdd(a.bundle_counter, 1);
        __syncthreads();
        const int bundle = s_bundle;
        if (bundle >= a.n_bundles) break;
        int zone[LPT];
        unsigned has_task = 0; // bit t: origin slot g*LPT+t holds a tree of this batch
#pragma unroll
        for (int t = 0; t < LPT; ++t) {
            const int task = __ldg(&a.bundle_tasks[bundle * B + g * LPT + t]);
            zone[t] = task >= 0 ? __ldg(&a.s.task_zone[task]) : -1;
            has_task |= (unsigned)(task >= 0) << t;
        }
        unsigned long long* const bl = a.bl + (size_t)bundle * N * B;
        unsigned long long* const blg = bl + g * LPT; // this thread's origins inside a node row

        { // routing.h:677-689 for B trees at once
            ulonglong2* v = reinterpret_cast<ulonglong2*>(bl);
            const size_t nv = (size_t)N * B / 2;
            const ulonglong2 x = make_ulonglong2(INF_BITS, INF_BITS);
            for (size_t i = tid; i < nv; i += BLOCK) v[i] = x;
        }
        for (int i = tid; i < n_words; i += BLOCK) s_state[i] = 0u;
        if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0; s_n_far = 0; s_n_far2 = 0; s_overflow = 0; }
        __syncthreads();
        // seed: every origin relaxes its own (unhidden) row once with label 0; the heads wait in the far pile
        if (tid < B) {
            const int task = __ldg(&a.bundle_tasks[bundle * B + tid]);
            if (task >= 0) {
                const int origin = __ldg(&a.s.task_origin[task]);
                const int zone = __ldg(&a.s.task_zone[task]);
                bl[(size_t)origin * B + tid] = 0ULL; // +0.0
                a.s.tie_nodes[task] = 0;
                const int r1 = __ldg(&a.s.row_ptr[origin + 1]);
                for (int e = __ldg(&a.s.row_ptr[origin]); e < r1; ++e) {
                    const Edge ed = ld_edge(&a.s.edges[e]);
                    int to = ed.to;
                    if (to < 0) { // routing.h:779-786
                        to &= 0x7fffffff;
                        if (__ldg(&a.s.link_tag[ed.link]) != zone) continue;
                    }
                    if (to == origin) continue;
                    const double c = 0.0 + ed.cost; // routing.h:804
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
        PF(1);
        int n_near = 0;
        int buf = 0; // near-queue buffer read this round
        int ci = 0;  // s_cnt[ci] counted this round's entries, pushes count into s_cnt[ci+1], s_cnt[ci+2] is cleared
        unsigned thr = 0u;
        bool failed = false;

        for (;;) {
            while (n_near > 0) {
                const int ci1 = ci == 2 ? 0 : ci + 1;
                const int ci2 = ci1 == 2 ? 0 : ci1 + 1;
                const int in0 = buf * CAP, out0 = (buf ^ 1) * CAP;
                // Expansion: a group of G threads owns one queue entry (a node); thread g relaxes the node's out-edges for origins
                // g and g + G.  Chain of a pass: clear the node's queue bits (shared memory) -> label row (the edge records are
                // loaded in its shadow: the entry carries the row) -> atomicMin round trip (the heads' rows are loaded in its
                // shadow) -> queue bits of the heads (shared memory) -> push.
                for (int base = 0; base < n_near; base += BLOCK / G) {
                    if (base + (tid & ~31) / G >= n_near) continue; // no entry for this warp
                    const int i = base + tid / G;
                    const bool have = i < n_near;
                    int u = 0;
                    int2 row = make_int2(0, 0);
                    unsigned was = 0;
                    if (have) {
                        u = s_q[in0 + i];
                        row = s_r[in0 + i];
                        // the label row is requested only after the queue bits are cleared: an improvement of this node that
                        // the row load below does not see finds the bits clear and queues the node again
                        if (g == 0) was = atomicAnd(&s_state[ST_WORD(u)], ~(3u << ST_SHIFT(u))) >> ST_SHIFT(u);
                    }
                    int e0 = row.x;
                    const int re = row.y;
                    double cost[BND_U];
                    int to[BND_U];
#pragma unroll
                    for (int j = 0; j < BND_U; ++j) {
                        cost[j] = 0.0; to[j] = 0;
                        if (e0 + j < re) {
                            const int4 v = __ldg(reinterpret_cast<const int4*>(&a.s.edges[e0 + j]));
                            cost[j] = __hiloint2double(v.y, v.x); to[j] = v.z;
                        }
                    }
                    was = __shfl_sync(0xffffffffu, was, gfirst) & 3u;
                    double lu[LPT];
                    unsigned act = 0; // bit t: origin t of this thread has reached the node
                    {
                        ulonglong2 lab[LPT / 2];
#pragma unroll
                        for (int t = 0; t < LPT / 2; ++t) lab[t] = make_ulonglong2(INF_BITS, INF_BITS);
                        if (was) {
#pragma unroll
                            for (int t = 0; t < LPT / 2; ++t) lab[t] = __ldcg(reinterpret_cast<const ulonglong2*>(blg + (size_t)u * B) + t);
                        }
#pragma unroll
                        for (int t = 0; t < LPT / 2; ++t) {
                            lu[2 * t] = __longlong_as_double((long long)lab[t].x);
                            lu[2 * t + 1] = __longlong_as_double((long long)lab[t].y);
                            act |= (unsigned)(lab[t].x < INF_BITS) << (2 * t) | (unsigned)(lab[t].y < INF_BITS) << (2 * t + 1);
                        }
                        act &= has_task;
                    }
                    if (have) {
                        if (g == 0) ++c_pops;
                        c_relax += (unsigned)__popc(act) * (unsigned)(re - e0);
                    }
#ifdef SSSP_PROFILE
                    pf[6] += 1;
                    if (lu[0] + lu[LPT - 1] == -1.0) pf[6] += 1; // consume the label row
                    PF(0);
#endif
                    while (__any_sync(0xffffffffu, e0 < re)) {
                        unsigned long long old[BND_U][LPT];
                        int2 prow[PR];
                        unsigned okm = 0; // bit j*LPT+t: origin t of this thread may use edge j
#pragma unroll
                        for (int j = 0; j < BND_U; ++j) {
                            const bool valid = e0 + j < re;
                            unsigned ok = valid ? act : 0u;
                            if (to[j] < 0) { // outgoing connector of a zone: only that zone's own tree may use it (routing.h:779-786)
                                to[j] &= 0x7fffffff;
                                const int tag = __ldg(&a.s.link_tag[__ldg(&a.s.edges[e0 + j].link)]);
#pragma unroll
                                for (int t = 0; t < LPT; ++t)
                                    if (tag != zone[t]) ok &= ~(1u << t);
                            }
                            unsigned long long* const p = blg + (size_t)to[j] * B;
#pragma unroll
                            for (int t = 0; t < LPT; ++t) {
                                old[j][t] = 0ULL;
#if BND_PRECHECK
                                if (valid) old[j][t] = __ldcg(p + t);
#else
                                // routing.h:804,816
                                if ((ok >> t) & 1u) old[j][t] = atomicMin(p + t, (unsigned long long)__double_as_longlong(lu[t] + cost[j]));
#endif
                            }
                            okm |= ok << (j * LPT);
                            if (valid && g == (j % G)) prow[j / G] = __ldg(&a.s.rows[to[j]]); // in the shadow of the atomics
                        }
#if BND_PRECHECK
#pragma unroll
                        for (int j = 0; j < BND_U; ++j) { // routing.h:804,816: the atomic only where the label improves
                            unsigned long long* const p = blg + (size_t)to[j] * B;
#pragma unroll
                            for (int t = 0; t < LPT; ++t) {
                                const unsigned long long cb = (unsigned long long)__double_as_longlong(lu[t] + cost[j]);
                                if (((okm >> (j * LPT + t)) & 1u) && cb < old[j][t]) old[j][t] = atomicMin(p + t, cb);
                            }
                        }
#endif
#ifdef SSSP_PROFILE
                        if (old[0][0] + old[BND_U - 1][LPT - 1] == 1ULL) pf[6] += 1; // consume the atomics
                        PF(5);
#endif
#pragma unroll
                        for (int j = 0; j < BND_U; ++j) {
                            bool imp = false;
                            unsigned k = BND_DEAD_KEY;
#pragma unroll
                            for (int t = 0; t < LPT; ++t) {
                                const unsigned long long cb = (unsigned long long)__double_as_longlong(lu[t] + cost[j]);
                                if ((okm >> (j * LPT + t)) & 1u) {
                                    imp = imp || old[j][t] > cb;
                                    k = min(k, min(key_of(old[j][t]), key_of(cb)));
                                }
                            }
                            const unsigned bal = __ballot_sync(0xffffffffu, imp);
                            if (bal == 0u) continue;
                            const unsigned key = group_min_u32<G>(k);
                            if ((bal & gmask) && g == (j % G)) { // the thread that fetched the head's row pushes it
                                const int w = ST_WORD(to[j]), sh = ST_SHIFT(to[j]);
                                if (key < thr) {
                                    const unsigned o = atomicOr(&s_state[w], 1u << sh);
                                    if (!((o >> sh) & 1u)) {
                                        const int pos = atomicAdd(&s_cnt[ci1], 1);
                                        if (pos < CAP) { s_q[out0 + pos] = to[j]; s_r[out0 + pos] = prow[j / G]; }
                                        else s_overflow = 1;
                                    }
                                } else {
                                    const unsigned o = atomicOr(&s_state[w], 2u << sh);
                                    if (!((o >> sh) & 3u)) {
                                        const int pos = atomicAdd(&s_n_far, 1);
                                        if (pos < CAPF) q_far[pos] = make_int2(to[j], (int)key);
                                        else s_overflow = 1;
                                    }
                                }
                            }
                        }
                        e0 += BND_U;
#pragma unroll
                        for (int j = 0; j < BND_U; ++j)
                            if (e0 + j < re) {
                                const int4 v = __ldg(reinterpret_cast<const int4*>(&a.s.edges[e0 + j]));
                                cost[j] = __hiloint2double(v.y, v.x); to[j] = v.z;
                            }
                    }
                }
#ifdef SSSP_PROFILE
                pf[7] += n_near;
#endif
                PF(2);
                if (tid == 0) s_cnt[ci2] = 0; // last read before the previous barrier, pushed into from the round after next
                failed = __syncthreads_or(s_overflow); // the one barrier of a round
                PF(3);
                ci = ci1;
                buf ^= 1;
                n_near = min(s_cnt[ci], CAP);
                ++c_rounds;
                if (failed) break;
            }
            if (failed) break;
            // near queue drained: open the next label window from the far pile
            const int n_far = min(s_n_far, CAPF);
            if (n_far == 0) break;
            unsigned mm = BND_DEAD_KEY;
            for (int i = tid; i < n_far; i += BLOCK) {
                const int2 q = q_far[i];
                if ((s_state[ST_WORD(q.x)] >> ST_SHIFT(q.x)) & 2u) mm = min(mm, (unsigned)q.y);
            }
            mm = block_min_u32<BLOCK>(mm, s_red);
            if (mm == BND_DEAD_KEY) break; // every far entry is stale: no node waits for an expansion
            thr = key_of((unsigned long long)__double_as_longlong(__hiloint2double((int)mm, 0) + a.s.delta));
            if (thr <= mm) thr = mm + 1; // window narrower than the key resolution
            for (int base = 0; base < n_far; base += BLOCK / G) {
                if (base + (tid & ~31) / G >= n_far) continue;
                const int i = base + tid / G;
                const bool have = i < n_far;
                const int v = have ? q_far[i].x : 0;
                const int w = ST_WORD(v), sh = ST_SHIFT(v);
                unsigned o = 0;
                if (have && g == 0) o = atomicAnd(&s_state[w], ~(2u << sh)) >> sh; // a stale duplicate finds the bit already consumed
                o = __shfl_sync(0xffffffffu, o, gfirst);
                const bool live = (o & 2u) != 0;
                unsigned k = BND_DEAD_KEY;
                if (live) {
#pragma unroll
                    for (int t = 0; t < LPT / 2; ++t) {
                        const ulonglong2 lab = __ldcg(reinterpret_cast<const ulonglong2*>(blg + (size_t)v * B) + t);
                        k = min(k, min(key_of(lab.x), key_of(lab.y)));
                    }
                }
                const unsigned key = group_min_u32<G>(k);
                if (live && g == 0) {
                    if (key < thr) {
                        const unsigned o2 = atomicOr(&s_state[w], 1u << sh);
                        if (!((o2 >> sh) & 1u)) {
                            const int pos = atomicAdd(&s_cnt[ci], 1);
                            if (pos < CAP) { s_q[buf * CAP + pos] = v; s_r[buf * CAP + pos] = __ldg(&a.s.rows[v]); }
                            else s_overflow = 1;
                        }
                    } else {
                        atomicOr(&s_state[w], 2u << sh);
                        const int pos = atomicAdd(&s_n_far2, 1);
                        q_far2[pos] = make_int2(v, (int)key); // pos < n_far <= CAPF
                    }
                }
            }
            failed = __syncthreads_or(s_overflow);
            n_near = min(s_cnt[ci], CAP);
            __syncthreads();
            if (tid == 0) { s_n_far = s_n_far2; s_n_far2 = 0; }
            int2* t2 = q_far; q_far = q_far2; q_far2 = t2;
            ++c_rounds;
            __syncthreads();
            PF(4);
            if (failed) break;
        }
        PF(4);
        if (failed) {
            // a queue was too small: the trees of this bundle are recomputed by the global-queue kernel with worst-case
            // capacities (host: run_sssp_batch)
            __syncthreads();
            if (tid < B) {
                const int task = __ldg(&a.bundle_tasks[bundle * B + tid]);
                if (task >= 0) a.s.tie_nodes[task] = -1;
            }
            if (tid == 0) atomicAdd(&a.s.stats->overflow, 1ULL);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        c_relax += __shfl_xor_sync(0xffffffffu, c_relax, o);
        c_pops += __shfl_xor_sync(0xffffffffu, c_pops, o);
    }
    if (lane == 0) {
        atomicAdd(&a.s.stats->relaxations, c_relax);
        atomicAdd(&a.s.stats->pops, (unsigned long long)c_pops);
    }
    if (tid == 0) {
        atomicAdd(&a.s.stats->rounds, (unsigned long long)c_rounds);
#ifdef SSSP_PROFILE
        for (int j = 1; j < 8; ++j) atomicAdd(&a.s.stats->prof[j], (unsigned long long)pf[j]);
        atomicAdd(&a.s.stats->stage[0], (unsigned long long)pf[0]);
#endif
    }
}

// Final labels of a bundle -> per-tree labels[task][N] and predecessor links.  One block per tile of TN nodes and bundle:
//   1. the reverse-star records of the tile (upstream node, cost, link, zone-tag flag) are staged in shared memory by all
//      threads at once (two dependent round trips for the whole tile instead of three per node);
//   2. one group of B/2 threads per node (two origins per thread): the label rows of the node and of its upstream nodes are
//      coalesced 16-byte loads, issued together; the tight in-edge is the predecessor (two tight upstream nodes = a tie, left to the serial replay);
//   3. the tile is transposed through shared memory so that every tree receives TN consecutive labels / predecessors.
// SLIM (the caller does not keep per-tree arrays: the assignment loop, which only back-traces the trees): step 3 is dropped and
// the predecessors stay node-major next to the labels -- bp[node][B] records {predecessor link, predecessor node}, one
// coalesced row per node -- for the node-major back-trace (kernels_columns.cu: backtrace_nm_kernel).  Order-dependent ties
// and overflowed bundles raise a.batch_flag instead (the host then redoes the batch through the per-tree path).
#define FIN_CAP 384 /* reverse-star records staged per tile (a tile with more falls back to global loads for the rest) */
template <int B, int TN, bool SLIM>
__global__ void __launch_bounds__(256, 4) bundle_finish_kernel(BundleLaunch a)
{
    __shared__ unsigned long long s_lab[SLIM ? 1 : B][TN + 1];
    __shared__ int s_pred[SLIM ? 1 : B][TN + 1];
    __shared__ int s_task[B];
    __shared__ int s_ptr[TN + 1];
    __shared__ double s_cost[FIN_CAP];
    __shared__ int2 s_fl[FIN_CAP]; // {upstream node, link | bit 31 = zone-tagged connector}, in reverse-star (= forward scan) order
    constexpr int G = B / 2; // threads per node: thread g holds origins 2g and 2g+1 (16-byte loads)
    const int tid = threadIdx.x;
    const int g = tid & (G - 1);
    const int N = a.s.N;
    const int bundle = blockIdx.y;
    const int v0 = blockIdx.x * TN;
    const int nt = min(TN, N - v0);
    int task[2], origin[2], zone[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        task[t] = __ldg(&a.bundle_tasks[bundle * B + 2 * g + t]);
        if (task[t] >= 0 && a.s.tie_nodes[task[t]] < 0) { // the search of this bundle overflowed: the host re-runs its trees
            task[t] = -1;
            if (SLIM && tid < G && blockIdx.x == 0) atomicOr(a.batch_flag, 2);
        }
        origin[t] = task[t] >= 0 ? __ldg(&a.s.task_origin[task[t]]) : -1;
        zone[t] = task[t] >= 0 ? __ldg(&a.s.task_zone[task[t]]) : -1;
    }
    if (tid < G) { s_task[2 * tid] = task[0]; s_task[2 * tid + 1] = task[1]; }
    const ulonglong2* const bl2 = reinterpret_cast<const ulonglong2*>(a.bl + (size_t)bundle * N * B) + g;
    if (tid <= nt) s_ptr[tid] = __ldg(&a.s.in_ptr[v0 + tid]);
    __syncthreads();
    const int p0 = s_ptr[0];
    const int n_stage = min(s_ptr[nt] - p0, FIN_CAP);
    for (int i = tid; i < n_stage; i += 256) {
        const int2 ie = __ldg(&a.s.in_edges[p0 + i]);
        const Edge ed = ld_edge(&a.s.edges[ie.y]);
        s_cost[i] = ed.cost;
        s_fl[i] = make_int2(ie.x, ed.link | (ed.to < 0 ? 0x80000000 : 0));
    }
    __syncthreads();
    unsigned n_ties0 = 0, n_ties1 = 0;
    for (int j = tid / G; j < nt; j += 256 / G) {
        const int v = v0 + j;
        const ulonglong2 lvb = __ldg(bl2 + (size_t)v * G);
        const bool need0 = task[0] >= 0 && lvb.x < INF_BITS && v != origin[0];
        const bool need1 = task[1] >= 0 && lvb.y < INF_BITS && v != origin[1];
        const double lv0 = __longlong_as_double((long long)lvb.x), lv1 = __longlong_as_double((long long)lvb.y);
        int first0 = -1, first1 = -1, link0 = DTL_NO_PRED, link1 = DTL_NO_PRED;
        bool multi0 = false, multi1 = false;
        const int j1 = s_ptr[j + 1] - p0;
        for (int jj = s_ptr[j] - p0; jj < j1; jj += 4) { // four upstream rows in flight per thread
            int2 fl[4];
            double cost[4];
            ulonglong2 lub[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                fl[q] = make_int2(-1, 0); cost[q] = 0.0;
                lub[q] = make_ulonglong2(INF_BITS, INF_BITS);
                if (jj + q < j1) {
                    if (jj + q < FIN_CAP) {
                        fl[q] = s_fl[jj + q]; cost[q] = s_cost[jj + q];
                    } else { // tile with more reverse-star records than the staging area
                        const int2 ie = __ldg(&a.s.in_edges[p0 + jj + q]);
                        const Edge ed = ld_edge(&a.s.edges[ie.y]);
                        fl[q] = make_int2(ie.x, ed.link | (ed.to < 0 ? 0x80000000 : 0)); cost[q] = ed.cost;
                    }
                    lub[q] = __ldg(bl2 + (size_t)fl[q].x * G);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (fl[q].x < 0) continue;
                bool ok0 = need0 && lub[q].x < INF_BITS, ok1 = need1 && lub[q].y < INF_BITS;
                if (fl[q].y < 0) { // routing.h:779-786
                    const int tag = __ldg(&a.s.link_tag[fl[q].y & 0x7fffffff]);
                    ok0 = ok0 && tag == zone[0];
                    ok1 = ok1 && tag == zone[1];
                }
                // the reverse star lists a node's in-edges in forward scan order: of parallel links from one upstream node the
                // first one met is the one a serial scan keeps
                if (ok0 && __longlong_as_double((long long)lub[q].x) + cost[q] == lv0) {
                    if (first0 < 0) { first0 = fl[q].x; link0 = fl[q].y & 0x7fffffff; }
                    else if (fl[q].x != first0) multi0 = true;
                }
                if (ok1 && __longlong_as_double((long long)lub[q].y) + cost[q] == lv1) {
                    if (first1 < 0) { first1 = fl[q].x; link1 = fl[q].y & 0x7fffffff; }
                    else if (fl[q].x != first1) multi1 = true;
                }
            }
        }
        if (multi0) { link0 = TIE_MARK; ++n_ties0; }
        if (multi1) { link1 = TIE_MARK; ++n_ties1; }
        if (SLIM) {
            reinterpret_cast<int4*>(a.bp + ((size_t)bundle * N + v) * B)[g] = make_int4(link0, first0, link1, first1);
        } else {
            s_lab[2 * g][j] = lvb.x; s_lab[2 * g + 1][j] = lvb.y;
            s_pred[2 * g][j] = link0; s_pred[2 * g + 1][j] = link1;
        }
    }
    if (SLIM) {
        if (n_ties0 | n_ties1) atomicOr(a.batch_flag, 1);
        return;
    }
    if (n_ties0) atomicAdd(&a.s.tie_nodes[task[0]], (int)n_ties0);
    if (n_ties1) atomicAdd(&a.s.tie_nodes[task[1]], (int)n_ties1);
    __syncthreads();
    for (int idx = tid; idx < B * TN; idx += 256) {
        const int b = idx / TN, j = idx - b * TN;
        const int t = s_task[b];
        if (t >= 0 && j < nt) {
            a.s.labels[(size_t)t * N + v0 + j] = s_lab[b][j];
            a.s.pred_link[(size_t)t * N + v0 + j] = s_pred[b][j];
        }
    }
}

#ifndef PRED_MIN_BLOCKS
#define PRED_MIN_BLOCKS 4
#endif
template <int B>
__global__ void __launch_bounds__(256, PRED_MIN_BLOCKS) bundle_pred_kernel(BundleLaunch a)
{
    constexpr int per_block = 256 / (B / 2) * 8;
    bundle_pred_nodes<B, 256, false>(a, blockIdx.y, blockIdx.x * per_block, min(a.s.N, (blockIdx.x + 1) * per_block), blockIdx.x == 0);
}

template <int B>
static void pred_launch(const BundleLaunch& a, cudaStream_t st)
{
    constexpr int per_block = 256 / (B / 2) * 8;
    const dim3 grid((a.s.N + per_block - 1) / per_block, a.n_bundles);
    bundle_pred_kernel<B><<<grid, 256, 0, st>>>(a);
}

size_t bundle_smem_bytes(int N, int cap_near)
{
    const size_t words = (((size_t)(N + 15) >> 4) + 1) & ~(size_t)1;
    return words * sizeof(unsigned) + (size_t)2 * cap_near * (sizeof(int2) + sizeof(int));
}

template <int BLOCK, int B, int LPT>
static void bundle_launch(const BundleLaunch& a, cudaStream_t st)
{
    const size_t smem = bundle_smem_bytes(a.s.N, a.cap_near);
    cudaFuncSetAttribute(sssp_bundle_kernel<BLOCK, B, LPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sssp_bundle_kernel<BLOCK, B, LPT><<<a.grid, BLOCK, smem, st>>>(a);
}

template <int B>
static void finish_launch(const BundleLaunch& a, cudaStream_t st)
{
    constexpr int TN = 64;
    const dim3 grid((a.s.N + TN - 1) / TN, a.n_bundles);
    static const bool old_slim = getenv("DTL_SLIM_FINISH_OLD") != nullptr; // experiment hook: the tile-staged pass writing bp
    if (a.slim && !old_slim) pred_launch<B>(a, st);
    else if (a.slim) bundle_finish_kernel<B, TN, true><<<grid, 256, 0, st>>>(a);
    else bundle_finish_kernel<B, TN, false><<<grid, 256, 0, st>>>(a);
}

template <int B, int LPT>
static void bundle_launch_block(const BundleLaunch& a, cudaStream_t st)
{
    if (a.block >= 1024) bundle_launch<1024, B, LPT>(a, st);
    else if (a.block >= 768) bundle_launch<768, B, LPT>(a, st);
    else bundle_launch<512, B, LPT>(a, st);
}

void launch_sssp_bundle_finish(const BundleLaunch& a, cudaStream_t st)
{
    if (a.n_bundles == 0) return;
    switch (a.B) {
    case 2: finish_launch<2>(a, st); break;
    case 4: finish_launch<4>(a, st); break;
    case 8: finish_launch<8>(a, st); break;
    default: finish_launch<16>(a, st); break;
    }
}

void launch_sssp_bundle(const BundleLaunch& a, cudaStream_t st)
{
    if (a.n_bundles == 0) return;
    if (a.async && a.slim && a.fuse_pred) { // the search kernel also runs the label -> predecessor pass (CTAs without a bundle)
        launch_sssp_bundle_async_search(a, st);
        return;
    }
    if (a.async) launch_sssp_bundle_async_search(a, st); // experiment: no round barriers inside a label window
    else
    switch (a.B) {
    case 4: bundle_launch_block<4, 2>(a, st); break;
    case 8: a.lpt >= 4 ? bundle_launch_block<8, 4>(a, st) : bundle_launch_block<8, 2>(a, st); break;
    default: a.lpt >= 4 ? bundle_launch_block<16, 4>(a, st) : bundle_launch_block<16, 2>(a, st); break;
    }
    launch_sssp_bundle_finish(a, st);
}

} // namespace dtl

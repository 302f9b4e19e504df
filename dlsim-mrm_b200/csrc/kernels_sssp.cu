// K1: batched many-source shortest-path trees (replaces NetworkForSP::optimal_label_correcting,
// reference routing.h:633-882, for every (agent type, period, origin zone) task of an iteration).
//
// Design (B200): one persistent CTA per origin at a time, CTAs pull origins from a device counter so
// that all 148 SMs stay busy whatever the per-origin frontier sizes are.  Per origin the CTA runs a
// near/far (delta-window) label-correcting search:
//   * labels are FP64 bit patterns in HBM/L2 (one 8-byte word per node and origin); a relaxation is
//     one 16-byte edge load, one 8-byte label load and -- only when it improves -- one 64-bit
//     atomicMin (non-negative doubles order like unsigned integers), so any interleaving converges to
//     the same Bellman fixed point the reference's serial deque reaches: min over paths of the
//     left-to-right FP64 sum.  Labels are therefore bit-exact whatever the schedule.
//   * the frontier is a queue of (node, label-at-push, link) entries compacted with warp-aggregated
//     shared-memory counters; an entry is live iff its label still equals the node's label, so each
//     label value of a node is expanded exactly once and the entry's link IS the tight predecessor.
//   * entries beyond the current label window go to a far pile that is re-bucketed when the near
//     queue drains (window = min live far label + delta).
// Predecessors: the reference's predecessor is the unique tight in-edge unless two different upstream
// nodes tie exactly in FP64; every such tie is observed in-flight as an equality c == label[w]
//  and recorded as a candidate, verified against the final labels over the reverse
// star, and origins with a real tie are re-run by sssp_replay_kernel -- a faithful serial emulation of
// the reference deque (push-back for new nodes, push-front for re-entered ones).
//
// Roofline: HBM/L2 bound, 20 B per counted edge.  Compiled with -fmad=false.
#include <cstdlib>

#include "dtl_internal.h"

namespace dtl {

#define INF_BITS 0x430c6bf526340000ULL /* bit pattern of 1.0e15 */
#define TIE_MARK (-7777)
#ifndef SSSP_UNR
#define SSSP_UNR 4 /* edges expanded together per queue entry (memory-level parallelism) */
#endif
#ifndef SSSP_ROWINFO
#define SSSP_ROWINFO 0 /* 1: queue entries carry the node's forward-star row (loaded by the pusher, off the critical path) */
#endif
#ifndef SSSP_EDGEPAR
#define SSSP_EDGEPAR 0 /* 1: edge-parallel expansion (one edge per thread) instead of one queue entry per lane */
#endif
#ifndef SSSP_NOPRE
#define SSSP_NOPRE 0 /* 1: no label pre-check load, every relaxation is one atomicMin */
#endif

__device__ __forceinline__ double ld_label(const unsigned long long* p)
{
    return __longlong_as_double(__ldcg(reinterpret_cast<const long long*>(p)));
}
__device__ __forceinline__ Edge ld_edge(const Edge* p)
{
    const int4 v = __ldg(reinterpret_cast<const int4*>(p));
    Edge e;
    e.cost = __hiloint2double(v.y, v.x);
    e.to = v.z;
    e.link = v.w;
    return e;
}
__device__ __forceinline__ QEntry ld_entry(const QEntry* p)
{
    const int4 v = *reinterpret_cast<const int4*>(p);
    QEntry q;
    q.lab = __hiloint2double(v.y, v.x);
    q.node = v.z;
    q.link = v.w;
    return q;
}
__device__ __forceinline__ void st_entry(QEntry* p, double lab, int node, int link)
{
    int4 v;
    v.x = __double2loint(lab); v.y = __double2hiint(lab); v.z = node; v.w = link;
    *reinterpret_cast<int4*>(p) = v;
}

// warp-aggregated slot allocation on a shared-memory counter (works under divergence)
__device__ __forceinline__ int agg_inc(int* ctr)
{
    const unsigned act = __activemask();
    const unsigned m = __match_any_sync(act, (unsigned long long)(uintptr_t)ctr);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(ctr, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1));
}

template <int BLOCK>
__device__ __forceinline__ double block_min(double v, double* s_red)
{
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double w = threadIdx.x < BLOCK / 32 ? s_red[threadIdx.x] : DTL_MAX_LABEL_COST;
        for (int o = 16; o > 0; o >>= 1) w = fmin(w, __shfl_xor_sync(0xffffffffu, w, o));
        if (threadIdx.x == 0) s_red[0] = w;
    }
    __syncthreads();
    const double r = s_red[0];
    __syncthreads();
    return r;
}

template <int BLOCK, int UNR>
__global__ void __launch_bounds__(BLOCK) sssp_kernel(SsspLaunch a)
{
    __shared__ int s_task, s_n_next, s_n_far, s_n_far2, s_n_tie, s_overflow;
    __shared__ double s_red[32];
#if SSSP_EDGEPAR
    __shared__ double s_lab[BLOCK];
    __shared__ int s_rb[BLOCK], s_off[BLOCK + 1], s_wsum[BLOCK / 32];
#endif
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int N = a.N;
    QEntry* q_near = a.ws + (size_t)blockIdx.x * a.ws_stride;
    QEntry* q_next = q_near + a.cap_near;
    QEntry* q_far = q_next + a.cap_near;
    QEntry* q_far2 = q_far + a.cap_far;
#if SSSP_ROWINFO
    // side arrays {row begin, row end} of every queued node: the expansion need not wait for row_ptr
    int2* r_near = reinterpret_cast<int2*>(q_far2 + a.cap_far);
    int2* r_next = r_near + a.cap_near;
    int2* r_far = r_next + a.cap_near;
    int2* r_far2 = r_far + a.cap_far;
#endif
    int* tie_cand = a.tie_cand + (size_t)blockIdx.x * a.tie_cap;
    unsigned long long c_relax = 0, c_pops = 0, c_rounds = 0;
#ifdef SSSP_PROFILE
    long long pf_t = clock64(), pf_init = 0, pf_near = 0, pf_book = 0, pf_far = 0, pf_tie = 0;
#define PF(bucket) do { const long long n_ = clock64(); bucket += n_ - pf_t; pf_t = n_; } while (0)
    long long ps_t = 0, ps[6] = { 0, 0, 0, 0, 0, 0 };
    // stage clock: the asm consumes `dep`, so it issues only once that value has arrived
#define PS(k, dep) do { if (tid == 0) { long long n_; unsigned long long d_ = (unsigned long long)(dep); \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(n_) : "l"(d_) : "memory"); ps[k] += n_ - ps_t; ps_t = n_; } } while (0)
#else
#define PF(bucket) do { } while (0)
#define PS(k, dep) do { } while (0)
#endif

    for (;;) {
        __syncthreads();
        if (tid == 0) s_task = atomicAdd(a.task_counter, 1);
        __syncthreads();
        if (s_task >= a.n_run) break;
        const int task = a.task_list ? a.task_list[s_task] : s_task;
        const int origin = a.task_origin[task];
        const int zone = a.task_zone[task];
        unsigned long long* lab = a.labels + (size_t)task * N;
        int* pred = a.pred_link + (size_t)task * N;

        for (int i = tid; i < N; i += BLOCK) { // routing.h:677-689
            lab[i] = INF_BITS;
            pred[i] = DTL_NO_PRED;
        }
        if (tid == 0) {
            s_n_next = 0; s_n_far = 0; s_n_far2 = 0; s_n_tie = 0; s_overflow = 0;
            a.tie_nodes[task] = 0;
        }
        __syncthreads();
        if (a.row_ptr[origin + 1] - a.row_ptr[origin] == 0) continue; // routing.h:692: origin has no out-link
        if (tid == 0) {
            lab[origin] = 0ULL; // +0.0
            st_entry(&q_near[0], 0.0, origin, DTL_NO_PRED);
#if SSSP_ROWINFO
            r_near[0] = make_int2(a.row_ptr[origin], a.row_ptr[origin + 1]);
#endif
        }
        __syncthreads();
        PF(pf_init);
        int n_near = 1;
        double thr = a.delta;

        for (;;) {
            while (n_near > 0) {
#if SSSP_EDGEPAR
                // Edge-parallel expansion: phase A validates a chunk of BLOCK queue entries (one per thread) and scans their
                // degrees; phase B gives every thread ONE edge of the chunk, so all lanes relax and pushes are ballots.
                for (int base = 0; base < n_near; base += BLOCK) {
                    const int i = base + tid;
                    int deg = 0;
                    if (i < n_near) {
                        const QEntry q = ld_entry(&q_near[i]);
                        const double cur = ld_label(&lab[q.node]);
                        if (!(cur < q.lab)) { // else: a better label was pushed later, that entry expands the node
                            const int rb = __ldg(&a.row_ptr[q.node]), re = __ldg(&a.row_ptr[q.node + 1]);
                            pred[q.node] = q.link; // the entry that carries the current label names the tight link
                            ++c_pops;
                            deg = re - rb;
                            s_lab[tid] = q.lab;
                            s_rb[tid] = rb;
                        }
                    }
                    int incl = deg;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    if (lane == 31) s_wsum[tid >> 5] = incl;
                    __syncthreads();
                    int wbase = 0;
#pragma unroll
                    for (int w = 0; w < BLOCK / 32; ++w)
                        if (w < (tid >> 5)) wbase += s_wsum[w];
                    s_off[tid] = wbase + incl - deg;
                    if (tid == BLOCK - 1) s_off[BLOCK] = wbase + incl;
                    __syncthreads();
                    const int total = s_off[BLOCK];
                    for (int eb = 0; eb < total; eb += BLOCK) {
                        const int e = eb + tid;
                        bool ok = e < total;
                        int lo = 0;
                        if (ok) { // owner entry: last j with s_off[j] <= e
                            int hi = BLOCK;
#pragma unroll
                            for (int step = 0; step < 10; ++step) {
                                const int mid = (lo + hi) >> 1;
                                if (hi - lo > 1) { if (s_off[mid] <= e) lo = mid; else hi = mid; }
                            }
                        }
                        Edge ed;
                        ed.cost = 0.0; ed.to = 0; ed.link = 0;
                        double qlab = 0.0;
                        if (ok) {
                            ed = ld_edge(&a.edges[s_rb[lo] + (e - s_off[lo])]);
                            qlab = s_lab[lo];
                        }
                        int to = ed.to;
                        if (ok && to < 0) { // outgoing connector of a zone: only the origin zone may leave its centroid
                            to &= 0x7fffffff;
                            if (__ldg(&a.link_tag[ed.link]) != zone) ok = false; // routing.h:779-786
                        }
                        bool is_near = false, is_far = false, is_tie = false;
                        const double c = qlab + ed.cost; // routing.h:804
                        if (ok) {
                            ++c_relax;
                            const double lw = ld_label(&lab[to]);
                            if (c < lw) {             // routing.h:816
                                const unsigned long long cb = (unsigned long long)__double_as_longlong(c);
                                const unsigned long long old = atomicMin(&lab[to], cb);
                                if (old > cb) { is_near = c < thr; is_far = !is_near; }
                                else if (old == cb && to != origin) is_tie = true;
                            } else if (c == lw && to != origin) is_tie = true;
                        }
                        const unsigned nm = __ballot_sync(0xffffffffu, is_near), fm = __ballot_sync(0xffffffffu, is_far);
                        int bn = 0, bf = 0;
                        if (lane == 0) {
                            if (nm) bn = atomicAdd(&s_n_next, __popc(nm));
                            if (fm) bf = atomicAdd(&s_n_far, __popc(fm));
                        }
                        bn = __shfl_sync(0xffffffffu, bn, 0);
                        bf = __shfl_sync(0xffffffffu, bf, 0);
                        const unsigned lt = (1u << lane) - 1;
                        if (is_near) {
                            const int pos = bn + __popc(nm & lt);
                            if (pos < a.cap_near) st_entry(&q_next[pos], c, to, ed.link);
                            else s_overflow = 1;
                        } else if (is_far) {
                            const int pos = bf + __popc(fm & lt);
                            if (pos < a.cap_far) st_entry(&q_far[pos], c, to, ed.link);
                            else s_overflow = 1;
                        }
                        if (is_tie) { // exact FP64 ties are rare: slow path
                            const int pos = atomicAdd(&s_n_tie, 1);
                            if (pos < a.tie_cap) tie_cand[pos] = to;
                        }
                    }
                    __syncthreads(); // the chunk's shared arrays are reused
                }
#else
                // Warp-convergent expansion: every lane owns one queue entry, UNR of its edges are relaxed per pass
                // (edge loads, label loads and atomics of a pass are issued back to back), and the pushes of a pass
                // are placed with ONE warp scan + one shared-memory atomic per queue instead of per-push atomics.
                for (int base = 0; base < n_near; base += BLOCK) {
                    const int i = base + tid;
                    bool active = i < n_near;
                    QEntry q;
                    q.lab = 0.0; q.node = 0; q.link = DTL_NO_PRED;
                    int rb = 0, re = 0;
#ifdef SSSP_PROFILE
                    if (tid == 0) ps_t = clock64();
#endif
                    if (active) {
                        q = ld_entry(&q_near[i]);
#if SSSP_ROWINFO
                        const int2 rr = r_near[i];
                        rb = rr.x; re = rr.y;
#endif
                        PS(0, q.node);
                        const double cur = ld_label(&lab[q.node]);
#if !SSSP_ROWINFO
                        rb = __ldg(&a.row_ptr[q.node]);
                        re = __ldg(&a.row_ptr[q.node + 1]);
#endif
                        PS(1, __double_as_longlong(cur) + rb + re);
                        if (cur < q.lab) active = false; // a better label was pushed later: that entry expands the node
                    }
                    if (active) {
                        pred[q.node] = q.link;           // the entry that carries the current label names the tight link
                        ++c_pops;
                    } else re = rb;
                    for (int e0 = rb; __any_sync(0xffffffffu, e0 < re); e0 += UNR) {
                        Edge ed[UNR];
                        bool ok[UNR];
                        int to[UNR];
                        double c[UNR], lw[UNR];
                        unsigned long long old[UNR];
#pragma unroll
                        for (int j = 0; j < UNR; ++j) {
                            ok[j] = e0 + j < re;
                            if (ok[j]) ed[j] = ld_edge(&a.edges[e0 + j]);
                        }
                        PS(2, (ok[0] ? ed[0].to : 0) + (ok[UNR - 1] ? ed[UNR - 1].to : 0));
#pragma unroll
                        for (int j = 0; j < UNR; ++j) {
                            to[j] = 0; c[j] = 0.0; lw[j] = 0.0;
                            if (!ok[j]) continue;
                            to[j] = ed[j].to;
                            if (to[j] < 0) { // outgoing connector of a zone: only the origin zone may leave its centroid
                                to[j] &= 0x7fffffff;
                                if (__ldg(&a.link_tag[ed[j].link]) != zone) ok[j] = false; // routing.h:779-786
                            }
                        }
#pragma unroll
                        for (int j = 0; j < UNR; ++j) {
                            if (!ok[j]) continue;
                            ++c_relax;
                            c[j] = q.lab + ed[j].cost; // routing.h:804
#if !SSSP_NOPRE
                            lw[j] = ld_label(&lab[to[j]]);
#else
                            lw[j] = DTL_MAX_LABEL_COST * 2; // no pre-check: the atomic below decides
#endif
                        }
                        PS(3, __double_as_longlong(lw[0]) + __double_as_longlong(lw[UNR - 1]));
#pragma unroll
                        for (int j = 0; j < UNR; ++j) {
                            old[j] = 0ULL;
#if defined(SSSP_EXP_RED)
                            if (ok[j] && c[j] < lw[j]) { // experiment: fire-and-forget min, push decided by the pre-check
                                atomicMin(&lab[to[j]], (unsigned long long)__double_as_longlong(c[j]));
                                old[j] = ~0ULL;
                            }
#elif defined(SSSP_EXP_PLAIN)
                            if (ok[j] && c[j] < lw[j]) { // experiment: racy plain store
                                lab[to[j]] = (unsigned long long)__double_as_longlong(c[j]);
                                old[j] = ~0ULL;
                            }
#else
                            if (ok[j] && c[j] < lw[j])  // routing.h:816
                                old[j] = atomicMin(&lab[to[j]], (unsigned long long)__double_as_longlong(c[j]));
#endif
                        }
#if SSSP_ROWINFO
                        int2 prow[UNR];
#pragma unroll
                        for (int j = 0; j < UNR; ++j) {
                            prow[j] = make_int2(0, 0);
                            if (ok[j] && c[j] < lw[j]) { prow[j].x = __ldg(&a.row_ptr[to[j]]); prow[j].y = __ldg(&a.row_ptr[to[j] + 1]); }
                        }
#endif
                        PS(4, old[0] + old[UNR - 1]);
                        // classify: bit j of near_m / far_m / tie_m
                        unsigned near_m = 0, far_m = 0, tie_m = 0;
#pragma unroll
                        for (int j = 0; j < UNR; ++j) {
                            if (!ok[j]) continue;
                            const unsigned long long cb = (unsigned long long)__double_as_longlong(c[j]);
                            if (c[j] < lw[j]) {
                                if (old[j] > cb) { if (c[j] < thr) near_m |= 1u << j; else far_m |= 1u << j; }
                                else if (old[j] == cb && to[j] != origin) tie_m |= 1u << j;
                            } else if (c[j] == lw[j] && to[j] != origin) tie_m |= 1u << j;
#if SSSP_ROWINFO && SSSP_NOPRE
                            if (!((near_m | far_m) & (1u << j))) prow[j] = make_int2(0, 0);
#endif
                        }
                        // one inclusive scan carries both counts (near in the low half, far in the high half)
                        const int mine = __popc(near_m) | (__popc(far_m) << 16);
                        int incl = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int v = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += v;
                        }
                        const int total = __shfl_sync(0xffffffffu, incl, 31);
                        int base_near = 0, base_far = 0;
                        if (lane == 31) {
                            if (total & 0xffff) base_near = atomicAdd(&s_n_next, total & 0xffff);
                            if (total >> 16) base_far = atomicAdd(&s_n_far, total >> 16);
                        }
                        base_near = __shfl_sync(0xffffffffu, base_near, 31);
                        base_far = __shfl_sync(0xffffffffu, base_far, 31);
                        int pos_near = base_near + ((incl - mine) & 0xffff);
                        int pos_far = base_far + ((incl - mine) >> 16);
#pragma unroll
                        for (int j = 0; j < UNR; ++j) {
                            if (near_m & (1u << j)) {
                                if (pos_near < a.cap_near) {
                                    st_entry(&q_next[pos_near], c[j], to[j], ed[j].link);
#if SSSP_ROWINFO
                                    r_next[pos_near] = prow[j];
#endif
                                } else s_overflow = 1;
                                ++pos_near;
                            } else if (far_m & (1u << j)) {
                                if (pos_far < a.cap_far) {
                                    st_entry(&q_far[pos_far], c[j], to[j], ed[j].link);
#if SSSP_ROWINFO
                                    r_far[pos_far] = prow[j];
#endif
                                } else s_overflow = 1;
                                ++pos_far;
                            }
                        }
                        if (tie_m) { // exact FP64 ties are rare: slow path
#pragma unroll
                            for (int j = 0; j < UNR; ++j)
                                if (tie_m & (1u << j)) {
                                    const int pos = atomicAdd(&s_n_tie, 1);
                                    if (pos < a.tie_cap) tie_cand[pos] = to[j];
                                }
                        }
                        PS(5, 0);
                    }
                }
#endif
#ifdef SSSP_PROFILE
                const long long pf_wait0 = clock64();
#endif
                __syncthreads();
#ifdef SSSP_PROFILE
                pf_tie += clock64() - pf_wait0; // reuse: barrier wait of thread 0 after its own work
#endif
                PF(pf_near);
                n_near = min(s_n_next, a.cap_near);
                const int n_far_now = s_n_far;
                __syncthreads();
                if (tid == 0) s_n_next = 0;
                QEntry* t = q_near; q_near = q_next; q_next = t;
#if SSSP_ROWINFO
                { int2* tr = r_near; r_near = r_next; r_next = tr; }
#endif
                ++c_rounds;
                if (n_far_now > a.far_trigger) {
                    // far pile close to capacity: drop dead entries (at most one live entry per node survives)
                    const int nf = min(n_far_now, a.cap_far);
                    for (int i = tid; i < nf; i += BLOCK) {
                        const QEntry q = ld_entry(&q_far[i]);
                        if (ld_label(&lab[q.node]) == q.lab) {
                            const int pos = agg_inc(&s_n_far2);
                            st_entry(&q_far2[pos], q.lab, q.node, q.link);
#if SSSP_ROWINFO
                            r_far2[pos] = r_far[i];
#endif
                        }
                    }
                    __syncthreads();
                    if (tid == 0) { s_n_far = s_n_far2; s_n_far2 = 0; }
                    QEntry* t2 = q_far; q_far = q_far2; q_far2 = t2;
#if SSSP_ROWINFO
                    { int2* tr = r_far; r_far = r_far2; r_far2 = tr; }
#endif
                }
                __syncthreads();
                PF(pf_book);
            }
            // near queue drained: open the next label window from the far pile
            const int n_far = min(s_n_far, a.cap_far);
            if (n_far == 0) break;
            double m = DTL_MAX_LABEL_COST;
            for (int i = tid; i < n_far; i += BLOCK) {
                const QEntry q = ld_entry(&q_far[i]);
                if (ld_label(&lab[q.node]) == q.lab) m = fmin(m, q.lab);
            }
            m = block_min<BLOCK>(m, s_red);
            if (!(m < DTL_MAX_LABEL_COST)) break; // every far entry is dead
            thr = m + a.delta;
            for (int i = tid; i < n_far; i += BLOCK) {
                const QEntry q = ld_entry(&q_far[i]);
                if (ld_label(&lab[q.node]) == q.lab) {
                    if (q.lab < thr) {
                        const int pos = agg_inc(&s_n_next);
                        if (pos < a.cap_near) {
                            st_entry(&q_near[pos], q.lab, q.node, q.link);
#if SSSP_ROWINFO
                            r_near[pos] = r_far[i];
#endif
                        } else s_overflow = 1;
                    } else {
                        const int pos = agg_inc(&s_n_far2);
                        st_entry(&q_far2[pos], q.lab, q.node, q.link);
#if SSSP_ROWINFO
                        r_far2[pos] = r_far[i];
#endif
                    }
                }
            }
            __syncthreads();
            n_near = min(s_n_next, a.cap_near);
            __syncthreads();
            if (tid == 0) { s_n_far = s_n_far2; s_n_far2 = 0; s_n_next = 0; }
            QEntry* t2 = q_far; q_far = q_far2; q_far2 = t2;
#if SSSP_ROWINFO
            { int2* tr = r_far; r_far = r_far2; r_far2 = tr; }
#endif
            ++c_rounds;
            __syncthreads();
            PF(pf_far);
        }
        PF(pf_far);

        // ---- verify tie candidates against the final labels (reverse star) ----
        __syncthreads();
        const int n_tie = s_n_tie;
        if (n_tie > 0) {
            const bool all_nodes = n_tie > a.tie_cap;
            const int n_check = all_nodes ? N : n_tie;
            for (int i = tid; i < n_check; i += BLOCK) {
                const int w = all_nodes ? i : tie_cand[i];
                if (w == origin) continue;
                const double lw = ld_label(&lab[w]);
                if (!(lw < DTL_MAX_LABEL_COST)) continue;
                int first_u = -1, best_e = 0x7fffffff;
                bool multi = false;
                for (int j = a.in_ptr[w]; j < a.in_ptr[w + 1]; ++j) {
                    const int2 ie = a.in_edges[j];
                    const double lu = ld_label(&lab[ie.x]);
                    if (!(lu < DTL_MAX_LABEL_COST)) continue;
                    const Edge ed = ld_edge(&a.edges[ie.y]);
                    if (ed.to < 0 && a.link_tag[ed.link] != zone) continue;
                    if (lu + ed.cost == lw) {
                        if (first_u < 0) first_u = ie.x;
                        if (ie.x != first_u) multi = true;
                        else if (ie.y < best_e) best_e = ie.y;
                    }
                }
                if (multi) {
                    if (atomicExch(&pred[w], TIE_MARK) != TIE_MARK) atomicAdd(&a.tie_nodes[task], 1);
                } else if (first_u >= 0) {
                    // parallel links from one upstream node: the reference keeps the first in scan order
                    if (pred[w] != TIE_MARK) pred[w] = a.edges[best_e].link;
                }
            }
        }
        __syncthreads();
        PF(pf_tie);
        if (tid == 0) {
            atomicAdd(&a.stats->tie_candidates, (unsigned long long)n_tie);
            if (s_overflow) { // a queue was too small: the tree is incomplete, the host re-runs it with worst-case queues
                atomicAdd(&a.stats->overflow, 1ULL);
                a.tie_nodes[task] = -1;
            }
        }
    }
    // flush per-thread counters
    for (int o = 16; o > 0; o >>= 1) {
        c_relax += __shfl_xor_sync(0xffffffffu, c_relax, o);
        c_pops += __shfl_xor_sync(0xffffffffu, c_pops, o);
    }
    if ((tid & 31) == 0) {
        atomicAdd(&a.stats->relaxations, c_relax);
        atomicAdd(&a.stats->pops, c_pops);
    }
    if (tid == 0) atomicAdd(&a.stats->rounds, c_rounds);
#ifdef SSSP_PROFILE
    if (tid == 0) {
        atomicAdd(&a.stats->prof[0], (unsigned long long)pf_init); atomicAdd(&a.stats->prof[1], (unsigned long long)pf_near);
        atomicAdd(&a.stats->prof[2], (unsigned long long)pf_book); atomicAdd(&a.stats->prof[3], (unsigned long long)pf_far);
        atomicAdd(&a.stats->prof[4], (unsigned long long)pf_tie);
        for (int k = 0; k < 6; ++k) atomicAdd(&a.stats->stage[k], (unsigned long long)ps[k]);
    }
#endif
}

void launch_sssp(const SsspLaunch& a, cudaStream_t st)
{
    if (a.n_run == 0) return;
    static bool carveout_set = false;
    if (!carveout_set) { // the kernel needs < 5 KB of shared memory: give the rest of the 228 KB to L1 (edges, row pointers)
        carveout_set = true;
        const int pct = getenv("DTL_SSSP_CARVEOUT") ? atoi(getenv("DTL_SSSP_CARVEOUT")) : 0;
        cudaFuncSetAttribute(sssp_kernel<128, SSSP_UNR>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(sssp_kernel<256, SSSP_UNR>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(sssp_kernel<512, SSSP_UNR>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        cudaFuncSetAttribute(sssp_kernel<1024, SSSP_UNR>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    switch (a.block) {
    case 128: sssp_kernel<128, SSSP_UNR><<<a.grid, 128, 0, st>>>(a); break;
    case 512: sssp_kernel<512, SSSP_UNR><<<a.grid, 512, 0, st>>>(a); break;
    case 1024: sssp_kernel<1024, SSSP_UNR><<<a.grid, 1024, 0, st>>>(a); break;
    default: sssp_kernel<256, SSSP_UNR><<<a.grid, 256, 0, st>>>(a); break;
    }
}

void sssp_occupancy(int block, int* max_ctas_per_sm, int* regs)
{
    int n = 0;
    cudaFuncAttributes fa{};
    switch (block) {
    case 128: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sssp_kernel<128, SSSP_UNR>, 128, 0); cudaFuncGetAttributes(&fa, sssp_kernel<128, SSSP_UNR>); break;
    case 512: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sssp_kernel<512, SSSP_UNR>, 512, 0); cudaFuncGetAttributes(&fa, sssp_kernel<512, SSSP_UNR>); break;
    case 1024: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sssp_kernel<1024, SSSP_UNR>, 1024, 0); cudaFuncGetAttributes(&fa, sssp_kernel<1024, SSSP_UNR>); break;
    default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sssp_kernel<256, SSSP_UNR>, 256, 0); cudaFuncGetAttributes(&fa, sssp_kernel<256, SSSP_UNR>); break;
    }
    if (max_ctas_per_sm) *max_ctas_per_sm = n;
    if (regs) *regs = fa.numRegs;
}

// Serial emulation of the reference deque label-correcting scan (routing.h:706-882) for origins whose
// tree contains an exact FP64 tie between different upstream nodes: the winner then depends on the
// whole scan history, so the history is replayed.  One thread walks the deque; the block only helps
// with initialisation.  replay_tasks lists the batch-local task indices.
__global__ void __launch_bounds__(128) sssp_replay_kernel(SsspLaunch a, const int* replay_tasks, int* status_ws,
                                                          int* next_ws)
{
    const int task = replay_tasks[blockIdx.x];
    const int N = a.N;
    const int origin = a.task_origin[task];
    const int zone = a.task_zone[task];
    double* label = reinterpret_cast<double*>(a.labels + (size_t)task * N);
    int* pred = a.pred_link + (size_t)task * N;
    int* status = status_ws + (size_t)blockIdx.x * N;
    int* next = next_ws + (size_t)blockIdx.x * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        status[i] = 0; label[i] = DTL_MAX_LABEL_COST; pred[i] = DTL_NO_PRED; next[i] = -1;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (a.row_ptr[origin + 1] - a.row_ptr[origin] == 0) return;
    label[origin] = 0.0;
    int front = origin, tail = origin;
    next[origin] = -1;
    unsigned long long pops = 0;
    while (front != -1) {
        const int u = front;
        front = next[u];
        next[u] = -1;
        status[u] = 2;
        ++pops;
        const double lu = label[u];
        for (int e = a.row_ptr[u]; e < a.row_ptr[u + 1]; ++e) {
            const Edge ed = a.edges[e];
            int v = ed.to;
            if (v < 0) {
                v &= 0x7fffffff;
                if (a.link_tag[ed.link] != zone) continue;
            }
            const double c = lu + ed.cost;
            if (c < label[v]) {
                label[v] = c;
                pred[v] = ed.link;
                if (status[v] == 0) { // push back
                    if (front == -1) { front = v; tail = v; next[v] = -1; }
                    else { next[tail] = v; next[v] = -1; tail = v; }
                    status[v] = 1;
                }
                if (status[v] == 2) { // push front
                    if (front == -1) { next[v] = -1; front = v; tail = v; }
                    else { next[v] = front; front = v; }
                    status[v] = 1;
                }
            }
        }
    }
    atomicAdd(&a.stats->replayed, 1ULL);
    atomicAdd(&a.stats->pops, pops);
}

void launch_sssp_replay(const SsspLaunch& a, const int* replay_tasks, int n_replay, int* status_ws, int* next_ws,
                        cudaStream_t st)
{
    if (n_replay == 0) return;
    sssp_replay_kernel<<<n_replay, 128, 0, st>>>(a, replay_tasks, status_ws, next_ws);
}

// Graph500-style edge count: forward-star edges of reachable nodes that pass the connector
// filter, each once per tree.
__global__ void count_edges_kernel(SsspLaunch a)
{
    const int N = a.N;
    unsigned long long cnt = 0;
    const size_t total = (size_t)a.n_tasks * N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int task = (int)(idx / N);
        const int u = (int)(idx - (size_t)task * N);
        if (!(__longlong_as_double((long long)a.labels[idx]) < DTL_MAX_LABEL_COST)) continue;
        const int zone = a.task_zone[task];
        for (int e = a.row_ptr[u]; e < a.row_ptr[u + 1]; ++e) {
            const Edge ed = ld_edge(&a.edges[e]);
            if (ed.to < 0 && a.link_tag[ed.link] != zone) continue;
            ++cnt;
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&a.stats->counted_edges, cnt);
}

void launch_count_edges(const SsspLaunch& a, cudaStream_t st)
{
    if (a.n_tasks == 0) return;
    count_edges_kernel<<<148 * 8, 256, 0, st>>>(a);
}

} // namespace dtl

// K1: batched many-source shortest-path trees (replaces NetworkForSP::optimal_label_correcting,
// reference routing.h:633-882, for every (agent type, period, origin zone) task of an iteration).
//
// Design (B200): one persistent CTA per origin at a time, CTAs pull origins from a device counter so
// that all 148 SMs stay busy whatever the per-origin frontier sizes are.  Per origin the CTA runs a
// near/far (delta-window) label-correcting search:
//   * labels are FP64 bit patterns in HBM/L2 (one 8-byte word per node and origin); a relaxation is
//     one 64-bit atomicMin (non-negative doubles order like unsigned integers), so any interleaving
//     converges to the same Bellman fixed point the reference's serial deque reaches: min over paths of
//     the left-to-right FP64 sum.  Labels are therefore bit-exact whatever the schedule.
//   * the frontier of a round lives in SHARED MEMORY as EDGE ITEMS: when a node's label improves, the
//     improving lane reserves one item per out-edge of that node, writes {label, node, parent} and
//     starts an asynchronous global->shared copy (cp.async) of the 16-byte edge record into the item.
//     The next round gives every lane ONE item: all lanes of a warp relax (no per-node loops, no idle
//     lanes), and the only long-latency step of a round is the atomicMin round trip -- the edge fetch
//     overlaps the rest of the pushing round.  Nodes with more than SSSP_DMAX out-edges get one
//     "range" item that a single lane walks; the origin's own star is read by the whole CTA.
//   * next to the items a round carries one HEAD entry {label-at-push, node, link} per pushed node: the
//     head is live iff its label still equals the node's label, so exactly one head per final label
//     value exists and its link IS the tight predecessor link.  Items of superseded heads are relaxed
//     anyway (their label is a real path cost, so this is harmless and avoids a dependent load).
//   * the relaxation back to the node the item came from is skipped when c > label-at-push: it can
//     never improve (c > label-at-push >= the parent's label when it pushed >= its label now).
//   * nodes beyond the current label window go to a far pile in global memory that is re-bucketed
//     when the round runs dry (window = min live far label + delta); pushes that do not fit the
//     shared-memory queues spill to the same pile.
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
#define RANGE_MARK (-5) /* link field of a range item: its edge slot holds {row begin, row end} instead of an edge */
#ifndef SSSP_DMAX
#define SSSP_DMAX 16 /* nodes with more out-edges than this are expanded by one lane walking the row */
#endif
#ifndef SSSP_MINB
#define SSSP_MINB 4 /* register cap: resident CTAs per SM wanted at 256 threads (scaled with the block size) */
#endif
#define SSSP_MIN_CTAS(BLOCK) ((SSSP_MINB * 256 / (BLOCK)) > 0 ? (SSSP_MINB * 256 / (BLOCK)) : 1)

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
__device__ __forceinline__ int4 pack_entry(double lab, int node, int link)
{
    int4 v;
    v.x = __double2loint(lab); v.y = __double2hiint(lab); v.z = node; v.w = link;
    return v;
}
__device__ __forceinline__ void st_entry(QEntry* p, double lab, int node, int link)
{
    *reinterpret_cast<int4*>(p) = pack_entry(lab, node, link);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
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

// routing.h:677-689 for one tree: label = 1e15, predecessor = -9999 (16-byte stores where the alignment allows)
template <int BLOCK>
__device__ __forceinline__ void init_tree(unsigned long long* lab, int* pred, int N)
{
    const int tid = threadIdx.x;
    {
        const int head = (int)(((uintptr_t)lab >> 3) & 1); // elements before the first 16-byte boundary
        const int nv = (N - head) >> 1;
        ulonglong2* v = reinterpret_cast<ulonglong2*>(lab + head);
        const ulonglong2 x = make_ulonglong2(INF_BITS, INF_BITS);
        for (int i = tid; i < nv; i += BLOCK) v[i] = x;
        if (tid == 0) {
            if (head) lab[0] = INF_BITS;
            for (int i = head + 2 * nv; i < N; ++i) lab[i] = INF_BITS;
        }
    }
    {
        const int head = min(N, (int)((4 - (((uintptr_t)pred >> 2) & 3)) & 3));
        const int nv = (N - head) >> 2;
        int4* v = reinterpret_cast<int4*>(pred + head);
        const int4 x = make_int4(DTL_NO_PRED, DTL_NO_PRED, DTL_NO_PRED, DTL_NO_PRED);
        for (int i = tid; i < nv; i += BLOCK) v[i] = x;
        if (tid == 0) {
            for (int i = 0; i < head; ++i) pred[i] = DTL_NO_PRED;
            for (int i = head + 4 * nv; i < N; ++i) pred[i] = DTL_NO_PRED;
        }
    }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, SSSP_MIN_CTAS(BLOCK)) sssp_kernel(SsspLaunch a)
{
    extern __shared__ int4 s_dyn[];
    __shared__ int s_task, s_icnt[3], s_hcnt[3], s_n_far, s_n_far2, s_n_tie, s_overflow, s_spill;
    __shared__ double s_red[32];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int N = a.N;
    const int ICAP = a.cap_items, HCAP = a.cap_heads;
    // shared-memory round queues, two buffers each (this round / next round)
    int4* const s_ie = s_dyn;              // [2*ICAP] item: the edge record {cost lo, cost hi, to, link} (cp.async target)
    int4* const s_iq = s_ie + 2 * ICAP;    // [2*ICAP] item: {label lo, label hi, node, parent}; node < 0 marks an unused slot
    int4* const s_hd = s_iq + 2 * ICAP;    // [2*HCAP] head: {label lo, label hi, node, link};  node < 0 marks an unused slot
    QEntry* q_far = a.ws + (size_t)blockIdx.x * a.ws_stride;
    QEntry* q_far2 = q_far + a.cap_far;
    int* tie_cand = a.tie_cand + (size_t)blockIdx.x * a.tie_cap;
    unsigned c_relax = 0, c_pops = 0, c_rounds = 0;
    if (tid == 0) s_spill = 0;
#ifdef SSSP_PROFILE
    // thread 0's cycles per phase: [1] init [2] expansion [3] round barrier [4] re-bucketing [5] tie check; [6] re-bucketings [7] items read
    long long pf_t = clock64(), pf[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#define PF(k_) do { const long long n_ = clock64(); pf[k_] += n_ - pf_t; pf_t = n_; } while (0)
    // stage clocks of thread 0 inside the expansion; the asm consumes `dep`, so it issues only once that value has arrived:
    // [0] item read [2] filter/add [3] atomic + row [4] push [5] heads + cp.async wait; [6] passes
    long long ps_t = 0, ps[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#define PS(k_, dep) do { if (tid == 0) { long long n_; unsigned long long d_ = (unsigned long long)(dep); \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(n_) : "l"(d_) : "memory"); ps[k_] += n_ - ps_t; ps_t = n_; } } while (0)
#else
#define PF(k_) do { } while (0)
#define PS(k_, dep) do { } while (0)
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

        init_tree<BLOCK>(lab, pred, N);
        const int2 orow = make_int2(__ldg(&a.row_ptr[origin]), __ldg(&a.row_ptr[origin + 1])); // a.rows hides connector-only rows
        if (tid == 0) {
            s_icnt[0] = 0; s_icnt[1] = 0; s_icnt[2] = 0;
            s_hcnt[0] = 0; s_hcnt[1] = 0; s_hcnt[2] = 0;
            s_n_far = 0; s_n_far2 = 0; s_n_tie = 0; s_overflow = 0;
            a.tie_nodes[task] = 0;
        }
        __syncthreads();
        if (orow.y - orow.x == 0) continue; // routing.h:692: origin has no out-link
        if (tid == 0) {
            lab[origin] = 0ULL; // +0.0
            s_hd[0] = pack_entry(0.0, origin, DTL_NO_PRED);
        }
        __syncthreads();
        PF(1);
        // round 0: the items are the origin's own forward star, read straight from global memory by the whole CTA
        int n_items = orow.y - orow.x, n_heads = 1;
        bool from_origin = true;
        int buf = 0; // queue buffer read this round
        int ci = 0;  // s_?cnt[ci] counted this round's entries, pushes count into [ci+1], [ci+2] is cleared
        double thr = a.delta;

        for (;;) {
            while (n_items > 0 || n_heads > 0) {
                const int ci1 = ci == 2 ? 0 : ci + 1;
                const int ci2 = ci1 == 2 ? 0 : ci1 + 1;
                const int in_i = buf * ICAP, out_i = (buf ^ 1) * ICAP, in_h = buf * HCAP, out_h = (buf ^ 1) * HCAP;
                // heads: the liveness load is issued now and consumed after the items of the round
                int4 hq = make_int4(0, 0, -1, 0);
                unsigned long long hcur = 0ULL;
                if (tid < n_heads) hq = s_hd[in_h + tid];
                if (hq.z >= 0) hcur = __ldcg(&lab[hq.z]);
                for (int base = 0; base < n_items; base += BLOCK) {
                    if (base + (tid & ~31) >= n_items) continue; // no item for this warp
                    const int i = base + tid;
                    double qlab = 0.0;
                    int node = -1, parent = -1, e0 = 0, re = 0; // e0 < re: this lane has an edge to relax
                    Edge ed;
                    ed.cost = 0.0; ed.to = 0; ed.link = 0;
#ifdef SSSP_PROFILE
                    if (tid == 0) ps_t = clock64();
#endif
                    if (i < n_items) {
                        if (from_origin) {
                            node = origin; re = 1;
                            ed = ld_edge(&a.edges[orow.x + i]);
                        } else {
                            const int4 qv = s_iq[in_i + i];
                            const int4 ev = s_ie[in_i + i];
                            qlab = __hiloint2double(qv.y, qv.x); node = qv.z; parent = qv.w;
                            if (node >= 0) {
                                if (ev.w == RANGE_MARK) { // high-degree node: this lane walks the row
                                    e0 = ev.x; re = ev.y;
                                    ed = ld_edge(&a.edges[e0]);
                                } else {
                                    ed.cost = __hiloint2double(ev.y, ev.x); ed.to = ev.z; ed.link = ev.w;
                                    re = 1;
                                }
                            }
                        }
                    }
                    PS(0, node + ed.to);
                    while (__any_sync(0xffffffffu, e0 < re)) {
                        bool ok = e0 < re;
                        int to = ed.to;
                        if (ok && to < 0) { // outgoing connector of a zone: only the origin zone may leave its centroid
                            to &= 0x7fffffff;
                            if (__ldg(&a.link_tag[ed.link]) != zone) ok = false; // routing.h:779-786
                        }
                        const double c = qlab + ed.cost; // routing.h:804
                        if (to == parent && c > qlab) ok = false; // back to where the item came from: cannot improve
                        const bool near_c = c < thr;
                        PS(2, ok + near_c + to);
                        // the one long trip of a round: the atomic, and in its shadow the forward-star row of the head node
                        const unsigned long long cb = (unsigned long long)__double_as_longlong(c);
                        unsigned long long old = 0ULL;
                        int2 prow = make_int2(0, 0);
                        if (ok) { // routing.h:816
                            ++c_relax;
                            old = atomicMin(&lab[to], cb);
                            if (near_c) prow = __ldg(&a.rows[to]);
                        }
                        const bool imp = ok && old > cb;
                        const bool is_near = imp && near_c, is_far = imp && !near_c;
                        PS(3, imp + prow.x);
                        const unsigned nm = __ballot_sync(0xffffffffu, is_near), fm = __ballot_sync(0xffffffffu, is_far);
                        if (nm | fm) {
                            // a pushed node takes one head slot and one item slot per out-edge (one range item if it has many)
                            const bool range = prow.y - prow.x > SSSP_DMAX;
                            const int d = is_near ? (range ? 1 : prow.y - prow.x) : 0;
                            int incl = d;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                                if (lane >= o) incl += v;
                            }
                            int ib = 0, hb = 0, fb = 0;
                            if (lane == 31) {
                                if (incl) ib = atomicAdd(&s_icnt[ci1], incl);
                                if (nm) hb = atomicAdd(&s_hcnt[ci1], __popc(nm));
                                if (fm) fb = atomicAdd(&s_n_far, __popc(fm));
                            }
                            ib = __shfl_sync(0xffffffffu, ib, 31);
                            hb = __shfl_sync(0xffffffffu, hb, 31);
                            fb = __shfl_sync(0xffffffffu, fb, 31);
                            const unsigned lt = (1u << lane) - 1;
                            if (is_near) {
                                const int hp = hb + __popc(nm & lt);
                                const int ip = ib + incl - d;
                                if (hp < HCAP && ip + d <= ICAP) {
                                    s_hd[out_h + hp] = pack_entry(c, to, ed.link);
                                    const int4 qv = make_int4(__double2loint(c), __double2hiint(c), to, node);
                                    if (range) {
                                        s_iq[out_i + ip] = qv;
                                        s_ie[out_i + ip] = make_int4(prow.x, prow.y, 0, RANGE_MARK);
                                    } else {
                                        for (int t = 0; t < d; ++t) {
                                            s_iq[out_i + ip + t] = qv;
                                            cp_async16(&s_ie[out_i + ip + t], &a.edges[prow.x + t]);
                                        }
                                    }
                                } else { // the round queues are full: blank the reserved slots, the node waits in the far pile
                                    if (hp < HCAP) s_hd[out_h + hp] = make_int4(0, 0, -1, 0);
                                    for (int t = 0; t < d && ip + t < ICAP; ++t) s_iq[out_i + ip + t] = make_int4(0, 0, -1, -1);
                                    const int pf_ = atomicAdd(&s_n_far, 1);
                                    if (pf_ < a.cap_far) st_entry(&q_far[pf_], c, to, ed.link);
                                    else s_overflow = 1;
                                    atomicAdd(&s_spill, 1);
                                }
                            } else if (is_far) {
                                const int pos = fb + __popc(fm & lt);
                                if (pos < a.cap_far) st_entry(&q_far[pos], c, to, ed.link);
                                else s_overflow = 1;
                            }
                        }
                        if (ok && old == cb && to != origin) { // exact FP64 ties are rare: slow path
                            const int pos = atomicAdd(&s_n_tie, 1);
                            if (pos < a.tie_cap) tie_cand[pos] = to;
                        }
                        PS(4, 0);
                        ++e0;
                        if (e0 < re) ed = ld_edge(&a.edges[e0]); // range items only
#ifdef SSSP_PROFILE
                        ps[6] += 1;
#endif
                    }
                }
#ifdef SSSP_PROFILE
                if (tid == 0) ps_t = clock64();
#endif
                // the head that carries the node's current label names the tight link (a superseded head leaves it to its successor)
                if (hq.z >= 0 && !(hcur < (((unsigned long long)(unsigned)hq.y << 32) | (unsigned)hq.x))) {
                    pred[hq.z] = hq.w;
                    ++c_pops;
                }
                for (int h = tid + BLOCK; h < n_heads; h += BLOCK) {
                    const int4 q = s_hd[in_h + h];
                    if (q.z >= 0 && !(__ldcg(&lab[q.z]) < (((unsigned long long)(unsigned)q.y << 32) | (unsigned)q.x))) {
                        pred[q.z] = q.w;
                        ++c_pops;
                    }
                }
                cp_async_wait_all(); // the edge records of the items this thread pushed have landed
                PS(5, 0);
#ifdef SSSP_PROFILE
                pf[7] += n_items;
#endif
                PF(2);
                if (tid == 0) { s_icnt[ci2] = 0; s_hcnt[ci2] = 0; } // last read before the previous barrier, pushed into from the round after next
                // the one barrier of a round; thread 0's view of the far pile may miss this round's pushes of slower warps
                // (accounted for in the worst-case capacity), but every thread gets the same decision
                const int compact = __syncthreads_or(tid == 0 && s_n_far > a.far_trigger);
                PF(3);
                ci = ci1;
                buf ^= 1;
                from_origin = false;
                n_items = min(s_icnt[ci], ICAP);
                n_heads = min(s_hcnt[ci], HCAP);
                ++c_rounds;
                if (compact) {
                    // far pile close to capacity: drop dead entries (at most one live entry per node survives)
                    __syncthreads(); // every thread has read its counts; nobody pushes any more
                    const int nf = min(s_n_far, a.cap_far);
                    __syncthreads();
                    for (int i = tid; i < nf; i += BLOCK) {
                        const QEntry q = ld_entry(&q_far[i]);
                        if (ld_label(&lab[q.node]) == q.lab) {
                            const int pos = agg_inc(&s_n_far2);
                            st_entry(&q_far2[pos], q.lab, q.node, q.link);
                        }
                    }
                    __syncthreads();
                    if (tid == 0) { s_n_far = s_n_far2; s_n_far2 = 0; }
                    QEntry* t2 = q_far; q_far = q_far2; q_far2 = t2;
                    __syncthreads();
                }
            }
            // the round ran dry: open the next label window from the far pile
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
                    bool placed = false;
                    if (q.lab < thr) {
                        const int2 row = __ldg(&a.rows[q.node]);
                        const bool range = row.y - row.x > SSSP_DMAX;
                        const int d = range ? 1 : row.y - row.x;
                        const int hp = atomicAdd(&s_hcnt[ci], 1);
                        const int ip = d ? atomicAdd(&s_icnt[ci], d) : 0;
                        if (hp < HCAP && ip + d <= ICAP) {
                            placed = true;
                            s_hd[buf * HCAP + hp] = pack_entry(q.lab, q.node, q.link);
                            const int4 qv = make_int4(__double2loint(q.lab), __double2hiint(q.lab), q.node,
                                                      q.link >= 0 ? __ldg(&a.link_from[q.link]) : -1);
                            if (range) {
                                s_iq[buf * ICAP + ip] = qv;
                                s_ie[buf * ICAP + ip] = make_int4(row.x, row.y, 0, RANGE_MARK);
                            } else {
                                for (int t = 0; t < d; ++t) {
                                    s_iq[buf * ICAP + ip + t] = qv;
                                    cp_async16(&s_ie[buf * ICAP + ip + t], &a.edges[row.x + t]);
                                }
                            }
                        } else {
                            if (hp < HCAP) s_hd[buf * HCAP + hp] = make_int4(0, 0, -1, 0);
                            for (int t = 0; t < d && ip + t < ICAP; ++t) s_iq[buf * ICAP + ip + t] = make_int4(0, 0, -1, -1);
                        }
                    }
                    if (!placed) { // beyond the window, or the round queues are full: stays in the pile
                        const int pf_ = agg_inc(&s_n_far2);
                        st_entry(&q_far2[pf_], q.lab, q.node, q.link);
                    }
                }
            }
            cp_async_wait_all();
            __syncthreads();
            n_items = min(s_icnt[ci], ICAP);
            n_heads = min(s_hcnt[ci], HCAP);
            __syncthreads();
            if (tid == 0) { s_n_far = s_n_far2; s_n_far2 = 0; }
            QEntry* t2 = q_far; q_far = q_far2; q_far2 = t2;
            ++c_rounds;
            __syncthreads();
#ifdef SSSP_PROFILE
            pf[6] += 1;
#endif
            PF(4);
        }
        PF(4);

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
        PF(5);
        if (tid == 0) {
            atomicAdd(&a.stats->tie_candidates, (unsigned long long)n_tie);
            if (s_overflow) { // the far pile was too small: the tree is incomplete, the host re-runs it with worst-case capacity
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
        atomicAdd(&a.stats->relaxations, (unsigned long long)c_relax);
        atomicAdd(&a.stats->pops, (unsigned long long)c_pops);
    }
    __syncthreads();
    if (tid == 0) {
        atomicAdd(&a.stats->rounds, (unsigned long long)c_rounds);
        if (s_spill) atomicAdd(&a.stats->prof[0], (unsigned long long)s_spill);
#ifdef SSSP_PROFILE
        for (int j = 1; j < 8; ++j) atomicAdd(&a.stats->prof[j], (unsigned long long)pf[j]);
        for (int j = 0; j < 8; ++j) atomicAdd(&a.stats->stage[j], (unsigned long long)ps[j]);
#endif
    }
}

static size_t sssp_smem_bytes(int cap_items, int cap_heads) { return ((size_t)4 * cap_items + (size_t)2 * cap_heads) * sizeof(int4); }

template <int BLOCK>
static cudaError_t sssp_prepare(size_t smem)
{
    return cudaFuncSetAttribute(sssp_kernel<BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

void launch_sssp(const SsspLaunch& a, cudaStream_t st)
{
    if (a.n_run == 0) return;
    const size_t smem = sssp_smem_bytes(a.cap_items, a.cap_heads);
    switch (a.block) {
    case 128: sssp_prepare<128>(smem); sssp_kernel<128><<<a.grid, 128, smem, st>>>(a); break;
    case 512: sssp_prepare<512>(smem); sssp_kernel<512><<<a.grid, 512, smem, st>>>(a); break;
    case 1024: sssp_prepare<1024>(smem); sssp_kernel<1024><<<a.grid, 1024, smem, st>>>(a); break;
    default: sssp_prepare<256>(smem); sssp_kernel<256><<<a.grid, 256, smem, st>>>(a); break;
    }
}

void sssp_occupancy(int block, int cap_items, int cap_heads, int* max_ctas_per_sm, int* regs)
{
    int n = 0;
    cudaFuncAttributes fa{};
    const size_t smem = sssp_smem_bytes(cap_items, cap_heads);
    switch (block) {
    case 128: sssp_prepare<128>(smem); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sssp_kernel<128>, 128, smem); cudaFuncGetAttributes(&fa, sssp_kernel<128>); break;
    case 512: sssp_prepare<512>(smem); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sssp_kernel<512>, 512, smem); cudaFuncGetAttributes(&fa, sssp_kernel<512>); break;
    case 1024: sssp_prepare<1024>(smem); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sssp_kernel<1024>, 1024, smem); cudaFuncGetAttributes(&fa, sssp_kernel<1024>); break;
    default: sssp_prepare<256>(smem); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sssp_kernel<256>, 256, smem); cudaFuncGetAttributes(&fa, sssp_kernel<256>); break;
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

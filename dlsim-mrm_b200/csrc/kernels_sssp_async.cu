// K1, asynchronous-window variant for FEW trees per SM (an origin shard on one of several GPUs): the same near/far
// label-correcting search as kernels_sssp.cu / kernels_sssp_smq.cu (same labels, same predecessor rule, same tie
// detection), but inside a label window there are NO rounds and NO CTA barriers.
//   * the near queue of a window is one linear array of entries in shared memory (28 B: label-at-push, node, link,
//     forward-star row, parent node) with four counters: `tail` (slots reserved), `commit` (slots written, in order),
//     `head` (slots claimed) and `pending` (entries written and not yet finished);
//   * every warp loops: claim up to 32/G committed entries (CAS on `head`), relax them (G lanes per entry, one edge per
//     lane; chain: entry from shared memory -> edge load -> atomicMin), push improved nodes -- reserve slots with one
//     atomicAdd on `tail` per warp, write, then commit in reservation order -- and retire its entries from `pending`;
//   * the window is drained when `pending` reaches zero (a producer is always itself a pending entry, so the count can
//     not touch zero while pushes are still possible); only then the CTA meets at a barrier and re-buckets the far pile.
// The critical path of a tree is therefore (hops of the deepest branch) x (one relaxation chain), not
// (rounds) x (slowest warp of the round + barrier).  Entries beyond the window, and entries that do not fit the array,
// go to the far pile in global memory; a tree whose far pile overflows is re-run by the global-queue kernel with the
// worst-case capacities.
#include "sssp_common.cuh"

namespace dtl {

#ifndef ASY_G
#define ASY_G 4 /* lanes per queue entry (power of two <= 32) */
#endif
#define ASY_SPIN_LIMIT (1 << 22) /* iterations of a spin wait before the tree is handed to the global-queue kernel */
#define ASY_MIN_CTAS(BLOCK) (1024 / (BLOCK)) /* 64 registers per thread: 1024 threads per SM */

__device__ __forceinline__ int ld_vol(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_vol(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }

template <int BLOCK, int G>
__global__ void __launch_bounds__(BLOCK, ASY_MIN_CTAS(BLOCK)) sssp_async_kernel(SsspLaunch a)
{
    extern __shared__ int4 s_dyn[];
    __shared__ int s_task, s_head, s_tail, s_commit, s_pending, s_stuck, s_n_far, s_n_far2, s_n_tie, s_overflow, s_spill;
    __shared__ double s_red[32];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int k = tid & (G - 1);   // edge slot of this lane inside its group of G lanes (one group per queue entry)
    const int grp = lane / G;      // entry of this lane inside the warp's claim
    const int N = a.N;
    const int R = a.cap_smem;      // entries of one window in shared memory
    int4* const s_q = s_dyn;                               // {label lo, label hi, node, link}
    int2* const s_r = reinterpret_cast<int2*>(s_q + R);    // {row begin, row end} of the node's forward star
    int* const s_p = reinterpret_cast<int*>(s_r + R);      // the node the entry was pushed from
    QEntry* q_far = a.ws + (size_t)blockIdx.x * a.ws_stride;
    QEntry* q_far2 = q_far + a.cap_far;
    int* tie_cand = a.tie_cand + (size_t)blockIdx.x * a.tie_cap;
    unsigned c_relax = 0, c_pops = 0, c_rounds = 0;
    if (tid == 0) s_spill = 0;

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
            s_n_far = 0; s_n_far2 = 0; s_n_tie = 0; s_overflow = 0; s_stuck = 0;
            a.tie_nodes[task] = 0;
        }
        __syncthreads();
        if (orow.y - orow.x == 0) continue; // routing.h:692: origin has no out-link
        if (tid == 0) {
            lab[origin] = 0ULL; // +0.0
            s_q[0] = pack_entry(0.0, origin, DTL_NO_PRED);
            s_r[0] = orow;
            s_p[0] = -1;
            s_head = 0; s_tail = 1; s_commit = 1; s_pending = 1;
        }
        __syncthreads();
        double thr = a.delta;

        for (;;) { // label windows
            // ---------------- asynchronous phase: every warp claims, relaxes and pushes until the window is drained
            for (;;) {
                int h = 0, n = 0;
                if (lane == 0) {
                    for (int spin = 0;; ++spin) {
                        h = ld_vol(&s_head);
                        const int c = ld_vol(&s_commit);
                        if (c > h) {
                            n = min(32 / G, c - h);
                            if (atomicCAS(&s_head, h, h + n) == h) break;
                            n = 0;
                            continue;
                        }
                        if (ld_vol(&s_pending) <= 0 || ld_vol(&s_stuck)) break; // nothing committed, nothing in flight: the window is drained
                        if (spin > ASY_SPIN_LIMIT) { st_vol(&s_stuck, 1); break; } // never seen; turns a protocol bug into a re-run, not a hang
                        __nanosleep(40);
                    }
                }
                h = __shfl_sync(0xffffffffu, h, 0);
                n = __shfl_sync(0xffffffffu, n, 0);
                if (n == 0) break;
                __threadfence_block(); // the claimed entries were written before their commit became visible
                double qlab = 0.0;
                int node = 0, parent = -1, e0 = 0, re = 0;
                if (grp < n) {
                    const int4 qv = s_q[h + grp];
                    const int2 rr = s_r[h + grp];
                    parent = s_p[h + grp];
                    qlab = __hiloint2double(qv.y, qv.x); node = qv.z;
                    e0 = rr.x + k; re = rr.y;
                }
                Edge ed;
                ed.cost = 0.0; ed.to = 0; ed.link = 0;
                if (e0 < re) ed = ld_edge(&a.edges[e0]);
                while (__any_sync(0xffffffffu, e0 < re)) {
                    bool ok = e0 < re;
                    int to = ed.to;
                    if (ok && to < 0) { // outgoing connector of a zone: only the origin zone may leave its centroid
                        to &= 0x7fffffff;
                        if (__ldg(&a.link_tag[ed.link]) != zone) ok = false; // routing.h:779-786
                    }
                    const double c = qlab + ed.cost; // routing.h:804
                    if (to == parent && c > qlab) ok = false; // back to where the entry came from: cannot improve
                    const bool near_c = c < thr;
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
                    const unsigned nm = __ballot_sync(0xffffffffu, is_near), fm = __ballot_sync(0xffffffffu, is_far);
                    if (nm | fm) {
                        const int cnt = __popc(nm);
                        int bn = 0, bf = 0;
                        if (lane == 0) {
                            if (cnt) {
                                bn = atomicAdd(&s_tail, cnt);
                                const int valid = max(0, min(cnt, R - bn));
                                if (valid) atomicAdd(&s_pending, valid); // before the commit: a written entry is always counted
                            }
                            if (fm) bf = atomicAdd(&s_n_far, __popc(fm));
                        }
                        bn = __shfl_sync(0xffffffffu, bn, 0);
                        bf = __shfl_sync(0xffffffffu, bf, 0);
                        const unsigned lt = (1u << lane) - 1;
                        if (is_near) {
                            const int pos = bn + __popc(nm & lt);
                            if (pos < R) {
                                s_q[pos] = pack_entry(c, to, ed.link);
                                s_r[pos] = prow;
                                s_p[pos] = node;
                            } else { // the window array is full: the entry waits in the far pile for the next re-bucketing
                                const int pf_ = atomicAdd(&s_n_far, 1);
                                if (pf_ < a.cap_far) st_entry(&q_far[pf_], c, to, ed.link);
                                else s_overflow = 1;
                                atomicAdd(&s_spill, 1);
                            }
                        } else if (is_far) {
                            const int pos = bf + __popc(fm & lt);
                            if (pos < a.cap_far) st_entry(&q_far[pos], c, to, ed.link);
                            else s_overflow = 1;
                        }
                        if (cnt && bn < R) { // commit in reservation order: consumers only see completely written entries
                            __threadfence_block();
                            __syncwarp();
                            if (lane == 0) {
                                int spin = 0;
                                while (ld_vol(&s_commit) != bn && !ld_vol(&s_stuck))
                                    if (++spin > ASY_SPIN_LIMIT) st_vol(&s_stuck, 1);
                                st_vol(&s_commit, min(bn + cnt, R));
                            }
                        }
                    }
                    if (ok && old == cb && to != origin) { // exact FP64 ties are rare: slow path
                        const int pos = atomicAdd(&s_n_tie, 1);
                        if (pos < a.tie_cap) tie_cand[pos] = to;
                    }
                    e0 += G;
                    if (e0 < re) ed = ld_edge(&a.edges[e0]);
                }
                __syncwarp();
                if (lane == 0) atomicSub(&s_pending, n); // after the pushes of these entries were counted
            }
            __syncthreads();
            ++c_rounds;
            // Predecessors of the window, with the labels at rest: the entry that carries a node's current label names the
            // tight link.  (Inside the window this would race: a superseded entry's late store could overwrite its successor's.)
            {
                const int n_done = min(s_tail, R);
                for (int i = tid; i < n_done; i += BLOCK) {
                    const int4 q = s_q[i];
                    if (__ldcg(&lab[q.z]) == (((unsigned long long)(unsigned)q.y << 32) | (unsigned)q.x)) {
                        pred[q.z] = q.w;
                        ++c_pops;
                    }
                }
            }
            if (s_stuck) s_overflow = 1; // the host re-runs the tree with the global-queue kernel
            __syncthreads();
            if (s_stuck) break;
            // ---------------- the window is drained: open the next one from the far pile
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
            if (tid == 0) { s_tail = 0; s_n_far2 = 0; }
            __syncthreads();
            for (int i = tid; i < n_far; i += BLOCK) {
                const QEntry q = ld_entry(&q_far[i]);
                if (ld_label(&lab[q.node]) == q.lab) {
                    int pos = R;
                    if (q.lab < thr) pos = agg_inc(&s_tail);
                    if (pos < R) {
                        s_q[pos] = pack_entry(q.lab, q.node, q.link);
                        s_r[pos] = __ldg(&a.rows[q.node]);
                        s_p[pos] = q.link >= 0 ? __ldg(&a.link_from[q.link]) : -1;
                    } else { // beyond the window, or the array is full: stays in the pile
                        const int pf_ = agg_inc(&s_n_far2);
                        st_entry(&q_far2[pf_], q.lab, q.node, q.link);
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                const int nn = min(s_tail, R);
                s_head = 0; s_tail = nn; s_commit = nn; s_pending = nn;
                s_n_far = s_n_far2;
            }
            QEntry* t2 = q_far; q_far = q_far2; q_far2 = t2;
            __syncthreads();
        }

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
        atomicAdd(&a.stats->rounds, (unsigned long long)c_rounds); // label windows, for this kernel
        if (s_spill) atomicAdd(&a.stats->prof[0], (unsigned long long)s_spill); // entries that did not fit the window array
    }
}

static size_t asy_smem_bytes(int cap_smem) { return (size_t)cap_smem * (sizeof(int4) + sizeof(int2) + sizeof(int)); }

template <int BLOCK>
static void asy_launch(const SsspLaunch& a, cudaStream_t st)
{
    const size_t smem = asy_smem_bytes(a.cap_smem);
    cudaFuncSetAttribute(sssp_async_kernel<BLOCK, ASY_G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sssp_async_kernel<BLOCK, ASY_G><<<a.grid, BLOCK, smem, st>>>(a);
}

void launch_sssp_async(const SsspLaunch& a, cudaStream_t st)
{
    if (a.n_run == 0) return;
    switch (a.block) {
    case 256: asy_launch<256>(a, st); break;
    case 512: asy_launch<512>(a, st); break;
    default: asy_launch<1024>(a, st); break;
    }
}

} // namespace dtl

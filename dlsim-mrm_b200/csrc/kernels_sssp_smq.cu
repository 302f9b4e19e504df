// K1, shared-memory-queue variant: the same near/far label-correcting search as kernels_sssp.cu (same labels, same
// predecessor rule, same tie detection), tuned for FEW trees per SM -- the regime of an origin shard on one of 8 GPUs,
// where the time of a launch is the latency of one tree, not the throughput of the SM.
//   * the near queues of the current and of the next round live in shared memory: entries {label-at-push, node, link,
//     forward-star row, parent node};
//   * G lanes own one queue entry and relax U out-edges each per pass (G = U = 2: a second pass only for nodes with more
//     than four out-edges), so a round of a big CTA is a single pass with every warp busy;
//   * the dependent chain of a round is two round trips: edge load -> atomicMin.  The row comes with the entry (fetched by
//     the pusher in the shadow of its atomic), the liveness of the entry is not waited for (relaxing with a superseded
//     label is harmless: it is a real path cost), only the predecessor store needs it;
//   * entries beyond the label window, and entries that do not fit the shared-memory queue, go to the far pile in global
//     memory, exactly like in the global-queue kernel.
// Measured on B200 (config-5 grid): 148 trees on 148 SMs in 2.6 ms (1024 threads per tree) against 5.9 ms for the
// global-queue kernel; at 7 trees per SM the global-queue kernel wins (DESIGN.md "K1").
#include "sssp_common.cuh"

namespace dtl {

#ifndef SMQ_G
#define SMQ_G 2 /* lanes per queue entry (power of two <= 32) */
#endif
#ifndef SMQ_U
#define SMQ_U 2 /* edges per lane and pass: G * U out-edges of a node are relaxed in one pass */
#endif
#ifndef SMQ_PREFETCH
#define SMQ_PREFETCH 0 /* 1: the pusher prefetches the pushed node's edge row into L1 (no measurable effect on B200) */
#endif
#ifndef SMQ_BALLOT
#define SMQ_BALLOT 0 /* 1: queue positions from ballots instead of a five-step shuffle scan (measured: no difference, 3.37 vs 3.36 ms per 250-tree shard) */
#endif
#ifndef SMQ_PUSH_PREFETCH
#define SMQ_PUSH_PREFETCH 0 /* 1: the pusher prefetches the pushed node's edge records into L1 (measured slower: 3.50 vs 3.36 ms per 250-tree shard) */
#endif
#define SMQ_MIN_CTAS(BLOCK) (1024 / (BLOCK)) /* 64 registers per thread: 1024 threads per SM */

// BOUNDED: destination-bounded search (SsspLaunch::bd_cell_ptr) -- when a label window is opened the smallest live label of the far
// pile bounds every label that can still change from below; the search stops when all destinations with demand lie below it.
template <int BLOCK, int G, int U, bool BOUNDED>
__global__ void __launch_bounds__(BLOCK, SMQ_MIN_CTAS(BLOCK)) sssp_smq_kernel(SsspLaunch a)
{
    extern __shared__ int4 s_dyn[];
    __shared__ int s_task, s_cnt[3], s_n_far, s_n_far2, s_n_tie, s_overflow, s_spill;
    __shared__ double s_red[32];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int k = tid & (G - 1); // edge slot of this lane inside its group of G lanes (one group per queue entry)
    const int N = a.N;
    const int CAP = a.cap_smem;
    // shared-memory near queues: two buffers (this round / next round) of CAP entries, structure of arrays
    int4* const s_q = s_dyn;                                   // {label lo, label hi, node, link}
    int2* const s_r = reinterpret_cast<int2*>(s_q + 2 * CAP);  // {row begin, row end} of the node's forward star
    int* const s_p = reinterpret_cast<int*>(s_r + 2 * CAP);    // the node the entry was pushed from
    QEntry* q_far = a.ws + (size_t)blockIdx.x * a.ws_stride;
    QEntry* q_far2 = q_far + a.cap_far;
    int* tie_cand = a.tie_cand + (size_t)blockIdx.x * a.tie_cap;
    unsigned c_relax = 0, c_pops = 0, c_rounds = 0;
    if (tid == 0) s_spill = 0;
#ifdef SSSP_PROFILE
    // thread 0's cycles: [1] init [2] expansion [3] round barrier [4] re-bucketing [5] tie check; [6] passes of warp 0 [7] entries read
    long long pf_t = clock64(), pf[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#define PF(k_) do { const long long n_ = clock64(); pf[k_] += n_ - pf_t; pf_t = n_; } while (0)
#else
#define PF(k_) do { } while (0)
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
            s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0;
            s_n_far = 0; s_n_far2 = 0; s_n_tie = 0; s_overflow = 0;
            a.tie_nodes[task] = 0;
        }
        __syncthreads();
        if (orow.y - orow.x == 0) continue; // routing.h:692: origin has no out-link
        if (tid == 0) {
            lab[origin] = 0ULL; // +0.0
            s_q[0] = pack_entry(0.0, origin, DTL_NO_PRED);
            s_r[0] = orow;
            s_p[0] = -1;
        }
        __syncthreads();
        PF(1);
        int n_near = 1;
        double bound = 2.0 * DTL_MAX_LABEL_COST; // BOUNDED: labels at or beyond it are not final when the search stopped early
        int buf = 0; // near-queue buffer read this round
        int ci = 0;  // s_cnt[ci] counted this round's entries, pushes count into s_cnt[ci+1], s_cnt[ci+2] is cleared
        double thr = a.delta;

        for (;;) {
            while (n_near > 0) {
                const int ci1 = ci == 2 ? 0 : ci + 1;
                const int ci2 = ci1 == 2 ? 0 : ci1 + 1;
                const int in0 = buf * CAP, out0 = (buf ^ 1) * CAP; // first slot of the queue read / written this round
                // Expansion: a group of G lanes owns one queue entry, every lane U of its edges (a second pass only for nodes
                // with more than G * U out-edges).  The chain of a pass is: entry from shared memory -> edge (L1: prefetched by the
                // pusher) -> atomicMin round trip -> ballot-placed push.  The liveness of the entry is NOT waited for: relaxing
                // with a superseded label is harmless (it is a real path cost), only the predecessor store needs it.
                for (int base = 0; base < n_near; base += BLOCK / G) {
                    if (base + (tid & ~31) / G >= n_near) continue; // no entry for this warp
                    const int i = base + tid / G;
                    double qlab = 0.0;
                    int node = 0, qlink = DTL_NO_PRED, parent = -1, e0 = 0, re = 0;
                    if (i < n_near) {
                        const int4 qv = s_q[in0 + i];
                        const int2 rr = s_r[in0 + i];
                        parent = s_p[in0 + i];
                        qlab = __hiloint2double(qv.y, qv.x); node = qv.z; qlink = qv.w;
                        e0 = rr.x + k * U; re = rr.y;
                    }
                    unsigned long long cur = 0ULL;
                    const bool head = k == 0 && i < n_near;
                    if (head) cur = __ldcg(&lab[node]); // consumed after the atomics have returned
                    Edge ed[U];
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        ed[j].cost = 0.0; ed[j].to = 0; ed[j].link = 0;
                        if (e0 + j < re) ed[j] = ld_edge(&a.edges[e0 + j]);
                    }
                    bool first = true;
                    while (__any_sync(0xffffffffu, e0 < re)) {
#ifdef SSSP_PROFILE
                        pf[6] += 1;
#endif
                        unsigned ok_m = 0, near_c = 0; // bit j: edge j is relaxed / its head falls into the near window
                        double c[U];
                        int to[U];
#pragma unroll
                        for (int j = 0; j < U; ++j) {
                            to[j] = ed[j].to;
                            c[j] = qlab + ed[j].cost; // routing.h:804
                            bool ok = e0 + j < re;
                            if (ok && to[j] < 0) { // outgoing connector of a zone: only the origin zone may leave its centroid
                                to[j] &= 0x7fffffff;
                                if (__ldg(&a.link_tag[ed[j].link]) != zone) ok = false; // routing.h:779-786
                            }
                            if (to[j] == parent && c[j] > qlab) ok = false; // back to where the entry came from: cannot improve
                            if (ok) {
                                ok_m |= 1u << j;
                                if (c[j] < thr) near_c |= 1u << j;
                            }
                        }
                        c_relax += __popc(ok_m);
                        // the one long trip: the atomics, and in their shadow the forward-star rows of the heads that may enter the near queue
                        unsigned long long old[U];
                        int2 prow[U];
#pragma unroll
                        for (int j = 0; j < U; ++j) {
                            old[j] = 0ULL;
                            prow[j] = make_int2(0, 0);
                            if (ok_m & (1u << j)) { // routing.h:816
                                old[j] = atomicMin(&lab[to[j]], (unsigned long long)__double_as_longlong(c[j]));
                                if (near_c & (1u << j)) prow[j] = __ldg(&a.rows[to[j]]);
                            }
                        }
                        unsigned near_m = 0, far_m = 0, tie_m = 0;
#pragma unroll
                        for (int j = 0; j < U; ++j) {
                            if (!(ok_m & (1u << j))) continue;
                            const unsigned long long cb = (unsigned long long)__double_as_longlong(c[j]);
                            if (old[j] > cb) { if (near_c & (1u << j)) near_m |= 1u << j; else far_m |= 1u << j; }
                            else if (old[j] == cb && to[j] != origin) tie_m |= 1u << j;
                        }
#if SMQ_BALLOT
                        // positions from ballots (independent of each other, unlike the steps of a shuffle scan): the entries of a
                        // lane follow those of the lanes below it
                        const unsigned lt = (1u << lane) - 1u;
                        int pre_near = 0, pre_far = 0, tot_near = 0, tot_far = 0;
#pragma unroll
                        for (int j = 0; j < U; ++j) {
                            const unsigned bn_ = __ballot_sync(0xffffffffu, (near_m >> j) & 1u);
                            const unsigned bf_ = __ballot_sync(0xffffffffu, (far_m >> j) & 1u);
                            pre_near += __popc(bn_ & lt); tot_near += __popc(bn_);
                            pre_far += __popc(bf_ & lt); tot_far += __popc(bf_);
                        }
                        const int total = tot_near | (tot_far << 16);
                        const int mine = 0, incl = pre_near | (pre_far << 16);
                        if (total) {
                            int bn = 0, bf = 0;
                            if (lane == 31) {
                                if (tot_near) bn = atomicAdd(&s_cnt[ci1], tot_near);
                                if (tot_far) bf = atomicAdd(&s_n_far, tot_far);
                            }
                            bn = __shfl_sync(0xffffffffu, bn, 31);
                            bf = __shfl_sync(0xffffffffu, bf, 31);
#else
                        // one inclusive scan carries both counts (near in the low half, far in the high half)
                        const int mine = __popc(near_m) | (__popc(far_m) << 16);
                        int incl = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int v = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += v;
                        }
                        const int total = __shfl_sync(0xffffffffu, incl, 31);
                        if (total) {
                            int bn = 0, bf = 0;
                            if (lane == 31) {
                                if (total & 0xffff) bn = atomicAdd(&s_cnt[ci1], total & 0xffff);
                                if (total >> 16) bf = atomicAdd(&s_n_far, total >> 16);
                            }
                            bn = __shfl_sync(0xffffffffu, bn, 31);
                            bf = __shfl_sync(0xffffffffu, bf, 31);
#endif
                            int pos_near = bn + ((incl - mine) & 0xffff);
                            int pos_far = bf + ((incl - mine) >> 16);
#pragma unroll
                            for (int j = 0; j < U; ++j) {
                                if (near_m & (1u << j)) {
                                    if (pos_near < CAP) {
                                        s_q[out0 + pos_near] = pack_entry(c[j], to[j], ed[j].link);
                                        s_r[out0 + pos_near] = prow[j];
                                        s_p[out0 + pos_near] = node;
#if SMQ_PUSH_PREFETCH
                                        // the head's edge records are on their way to L1 while this round finishes and the CTA meets
                                        // at the barrier: the next round's expansion starts from L1 instead of an L2 round trip
                                        if (prow[j].y > prow[j].x) {
                                            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.edges + prow[j].x));
                                            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.edges + prow[j].y - 1));
                                        }
#endif
                                    } else { // shared-memory queue full: the entry waits in the far pile for the next re-bucketing
                                        const int pf_ = atomicAdd(&s_n_far, 1);
                                        if (pf_ < a.cap_far) st_entry(&q_far[pf_], c[j], to[j], ed[j].link);
                                        else s_overflow = 1;
                                        atomicAdd(&s_spill, 1);
                                    }
                                    ++pos_near;
                                } else if (far_m & (1u << j)) {
                                    if (pos_far < a.cap_far) st_entry(&q_far[pos_far], c[j], to[j], ed[j].link);
                                    else s_overflow = 1;
                                    ++pos_far;
                                }
                            }
                        }
                        if (tie_m) { // exact FP64 ties are rare: slow path
#pragma unroll
                            for (int j = 0; j < U; ++j)
                                if (tie_m & (1u << j)) {
                                    const int pos = atomicAdd(&s_n_tie, 1);
                                    if (pos < a.tie_cap) tie_cand[pos] = to[j];
                                }
                        }
                        if (first) {
                            first = false;
                            // the entry that carries the current label names the tight link (a superseded entry leaves it to its successor)
                            if (head && !(cur < (unsigned long long)__double_as_longlong(qlab))) {
                                pred[node] = qlink;
                                ++c_pops;
                            }
                        }
                        e0 += G * U;
#pragma unroll
                        for (int j = 0; j < U; ++j)
                            if (e0 + j < re) ed[j] = ld_edge(&a.edges[e0 + j]);
                    }
                    if (first && head && !(cur < (unsigned long long)__double_as_longlong(qlab))) { // entry without usable out-edges
                        pred[node] = qlink;
                        ++c_pops;
                    }
                }
#ifdef SSSP_PROFILE
                pf[7] += n_near;
#endif
                PF(2);
                if (tid == 0) s_cnt[ci2] = 0; // last read before the previous barrier, pushed into from the round after next
                // the one barrier of a round; thread 0's view of the far pile may miss this round's pushes of slower warps
                // (accounted for in the worst-case capacity), but every thread gets the same decision
                const int compact = __syncthreads_or(tid == 0 && s_n_far > a.far_trigger);
                PF(3);
                ci = ci1;
                buf ^= 1;
                n_near = min(s_cnt[ci], CAP);
                ++c_rounds;
                if (compact) {
                    // far pile close to capacity: drop dead entries (at most one live entry per node survives)
                    __syncthreads(); // every thread has read its n_near; nobody pushes any more
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
            if (BOUNDED && a.bd_cell_ptr) {
                double far_dest = 0.0; // an unreached destination (1e15) keeps the search going
                const int c1 = __ldg(&a.bd_cell_ptr[task + 1]);
                for (int i = __ldg(&a.bd_cell_ptr[task]) + tid; i < c1; i += BLOCK)
                    far_dest = fmax(far_dest, ld_label(&lab[__ldg(&a.bd_dest[__ldg(&a.bd_cells[i])])]));
                far_dest = -block_min<BLOCK>(-far_dest, s_red);
                if (far_dest < m) { bound = m; break; } // nothing queued can get below m: every destination label is final
            }
            thr = m + a.delta;
            for (int i = tid; i < n_far; i += BLOCK) {
                const QEntry q = ld_entry(&q_far[i]);
                if (ld_label(&lab[q.node]) == q.lab) {
                    int pos = CAP;
                    if (q.lab < thr) pos = agg_inc(&s_cnt[ci]);
                    if (pos < CAP) {
                        const int2 row = __ldg(&a.rows[q.node]);
                        s_q[buf * CAP + pos] = pack_entry(q.lab, q.node, q.link);
                        s_r[buf * CAP + pos] = row;
                        s_p[buf * CAP + pos] = q.link >= 0 ? __ldg(&a.link_from[q.link]) : -1;
#if SMQ_PREFETCH
                        if (row.y > row.x) {
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.edges + row.x));
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.edges + row.y - 1));
                        }
#endif
                    } else { // beyond the window, or the shared-memory queue is full: stays in the pile
                        const int pf_ = agg_inc(&s_n_far2);
                        st_entry(&q_far2[pf_], q.lab, q.node, q.link);
                    }
                }
            }
            __syncthreads();
            n_near = min(s_cnt[ci], CAP);
            __syncthreads();
            if (tid == 0) { s_n_far = s_n_far2; s_n_far2 = 0; }
            QEntry* t2 = q_far; q_far = q_far2; q_far2 = t2;
            ++c_rounds;
            __syncthreads();
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
                if (!(lw < DTL_MAX_LABEL_COST) || !(lw < bound)) continue;
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
        atomicAdd(&a.stats->rounds, (unsigned long long)c_rounds);
        if (s_spill) atomicAdd(&a.stats->prof[0], (unsigned long long)s_spill); // entries that did not fit the shared-memory queue
#ifdef SSSP_PROFILE
        for (int j = 1; j < 8; ++j) atomicAdd(&a.stats->prof[j], (unsigned long long)pf[j]);
#endif
    }
}

static size_t smq_smem_bytes(int cap_smem) { return (size_t)2 * cap_smem * (sizeof(int4) + sizeof(int2) + sizeof(int)); }

template <int BLOCK>
static void smq_launch(const SsspLaunch& a, cudaStream_t st)
{
    const size_t smem = smq_smem_bytes(a.cap_smem);
    if (a.bd_cell_ptr) {
        cudaFuncSetAttribute(sssp_smq_kernel<BLOCK, SMQ_G, SMQ_U, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sssp_smq_kernel<BLOCK, SMQ_G, SMQ_U, true><<<a.grid, BLOCK, smem, st>>>(a);
        return;
    }
    cudaFuncSetAttribute(sssp_smq_kernel<BLOCK, SMQ_G, SMQ_U, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sssp_smq_kernel<BLOCK, SMQ_G, SMQ_U, false><<<a.grid, BLOCK, smem, st>>>(a);
}

void launch_sssp_smq(const SsspLaunch& a, cudaStream_t st)
{
    if (a.n_run == 0) return;
    switch (a.block) {
    case 256: smq_launch<256>(a, st); break;
    case 512: smq_launch<512>(a, st); break;
    default: smq_launch<1024>(a, st); break;
    }
}

} // namespace dtl

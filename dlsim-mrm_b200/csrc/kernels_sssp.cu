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
//   * the relaxation back to the node the entry came from is skipped when c > label-at-push: it can
//     never improve (c > label-at-push >= the parent's label when it pushed >= its label now).
//   * entries beyond the current label window go to a far pile that is re-bucketed when the near
//     queue drains (window = min live far label + delta).
//   * one CTA barrier per round (three rotating queue counters).
// The kernel is bound by the SM's L1TEX pipe (one scattered 8/16-byte access per cycle and SM) and by
// dependent L2/HBM round trips, not by HBM bytes: see DESIGN.md "K1" for the measured variants
// (shared-memory queues, lanes-per-entry and edge-item mappings were all slower at full occupancy).
// Predecessors: the reference's predecessor is the unique tight in-edge unless two different upstream
// nodes tie exactly in FP64; every such tie is observed in-flight as an equality c == label[w]
//  and recorded as a candidate, verified against the final labels over the reverse
// star, and origins with a real tie are re-run by sssp_replay_kernel -- a faithful serial emulation of
// the reference deque (push-back for new nodes, push-front for re-entered ones).
//
// Roofline: HBM/L2 bound, 20 B per counted edge.  Compiled with -fmad=false.
#include <cstdlib>

#include "sssp_common.cuh"

namespace dtl {

#ifndef SSSP_UNR
#define SSSP_UNR 4 /* edges expanded together per queue entry (memory-level parallelism) */
#endif
#ifndef SSSP_NOPRE
#define SSSP_NOPRE 0 /* 1: no label pre-check load, every relaxation is one atomicMin */
#endif
#ifndef SSSP_MINB
#define SSSP_MINB 7 /* register cap: resident CTAs per SM wanted at 128 threads (scaled with the block size) */
#endif
#define SSSP_MIN_CTAS(BLOCK) ((SSSP_MINB * 128 / (BLOCK)) > 0 ? (SSSP_MINB * 128 / (BLOCK)) : 1)

template <int BLOCK, int UNR>
__global__ void __launch_bounds__(BLOCK, SSSP_MIN_CTAS(BLOCK)) sssp_kernel(SsspLaunch a)
{
    __shared__ int s_task, s_cnt[3], s_n_far, s_n_far2, s_n_tie, s_overflow, s_rounds;
    __shared__ double s_red[32];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int N = a.N;
    QEntry* q_near = a.ws + (size_t)blockIdx.x * a.ws_stride;
    QEntry* q_next = q_near + a.cap_near;
    QEntry* q_far = q_next + a.cap_near;
    QEntry* q_far2 = q_far + a.cap_far;
    unsigned c_relax = 0, c_pops = 0;
    if (threadIdx.x == 0) s_rounds = 0;

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
        // a.rows hides the rows of connector-only nodes (centroids): only the origin of the tree expands its own
        const int2 orow = make_int2(__ldg(&a.row_ptr[origin]), __ldg(&a.row_ptr[origin + 1]));
        if (tid == 0) {
            s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0;
            s_n_far = 0; s_n_far2 = 0; s_n_tie = 0; s_overflow = 0;
            a.tie_nodes[task] = 0;
        }
        __syncthreads();
        if (orow.y - orow.x == 0) continue; // routing.h:692: origin has no out-link
        if (tid == 0) {
            lab[origin] = 0ULL; // +0.0
            st_entry(&q_near[0], 0.0, origin, DTL_NO_PRED);
        }
        __syncthreads();
        int n_near = 1;
        int ci = 0; // s_cnt[ci] counted this round's entries, pushes count into s_cnt[ci+1], s_cnt[ci+2] is cleared
        double thr = a.delta;

        for (;;) {
            while (n_near > 0) {
                // Warp-convergent expansion: every lane owns one queue entry, UNR of its edges are relaxed per pass
                // (edge loads, label loads and atomics of a pass are issued back to back), and the pushes of a pass
                // are placed with ONE warp scan + one shared-memory atomic per queue instead of per-push atomics.
                const int ci1 = ci == 2 ? 0 : ci + 1;
                const int ci2 = ci1 == 2 ? 0 : ci1 + 1;
                for (int base = 0; base < n_near; base += BLOCK) {
                    if (base + (tid & ~31) >= n_near) continue; // no entry for this warp
                    const int i = base + tid;
                    bool active = i < n_near;
                    QEntry q;
                    q.lab = 0.0; q.node = 0; q.link = DTL_NO_PRED;
                    int rb = 0, re = 0, parent = -1;
                    if (active) {
                        q = ld_entry(&q_near[i]);
                        const double cur = ld_label(&lab[q.node]);
                        if (q.link >= 0) { // the link the entry came over names the parent and the row of this node in one load
                            const int4 li = __ldg(&a.link_info[q.link]);
                            parent = li.x; rb = li.y; re = li.z;
                        } else { rb = __ldg(&a.row_ptr[q.node]); re = __ldg(&a.row_ptr[q.node + 1]); } // the origin's own entry
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
                            c[j] = q.lab + ed[j].cost; // routing.h:804
                            if (to[j] == parent && c[j] > q.lab) { ok[j] = false; continue; } // back to the parent: cannot improve
                            ++c_relax;
#if !SSSP_NOPRE
                            lw[j] = ld_label(&lab[to[j]]);
#else
                            lw[j] = DTL_MAX_LABEL_COST * 2; // no pre-check: the atomic below decides
#endif
                        }
#pragma unroll
                        for (int j = 0; j < UNR; ++j) {
                            old[j] = 0ULL;
                            if (ok[j] && c[j] < lw[j])  // routing.h:816
                                old[j] = atomicMin(&lab[to[j]], (unsigned long long)__double_as_longlong(c[j]));
                        }
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
                            if (total & 0xffff) base_near = atomicAdd(&s_cnt[ci1], total & 0xffff);
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
                                } else s_overflow = 1;
                                ++pos_near;
                            } else if (far_m & (1u << j)) {
                                if (pos_far < a.cap_far) {
                                    st_entry(&q_far[pos_far], c[j], to[j], ed[j].link);
                                } else s_overflow = 1;
                                ++pos_far;
                            }
                        }
                        if (tie_m) { // exact FP64 ties are rare: slow path
#pragma unroll
                            for (int j = 0; j < UNR; ++j)
                                if (tie_m & (1u << j)) {
                                    const int pos = atomicAdd(&s_n_tie, 1);
                                    if (pos < a.tie_cap) a.tie_cand[(size_t)blockIdx.x * a.tie_cap + pos] = to[j];
                                }
                        }
                    }
                }
                if (tid == 0) s_cnt[ci2] = 0; // last read before the previous barrier, pushed into from the round after next
                // the one barrier of a round; thread 0's view of the far pile may miss this round's pushes of slower warps
                // (accounted for in the worst-case capacity), but every thread gets the same decision
                const int compact = __syncthreads_or(tid == 0 && s_n_far > a.far_trigger);
                ci = ci1;
                n_near = min(s_cnt[ci], a.cap_near);
                QEntry* t = q_near; q_near = q_next; q_next = t;
                if (tid == 0) ++s_rounds;
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
            thr = m + a.delta;
            for (int i = tid; i < n_far; i += BLOCK) {
                const QEntry q = ld_entry(&q_far[i]);
                if (ld_label(&lab[q.node]) == q.lab) {
                    if (q.lab < thr) {
                        const int pos = agg_inc(&s_cnt[ci]);
                        if (pos < a.cap_near) {
                            st_entry(&q_near[pos], q.lab, q.node, q.link);
                        } else s_overflow = 1;
                    } else {
                        const int pos = agg_inc(&s_n_far2);
                        st_entry(&q_far2[pos], q.lab, q.node, q.link);
                    }
                }
            }
            __syncthreads();
            n_near = min(s_cnt[ci], a.cap_near);
            __syncthreads();
            if (tid == 0) { s_n_far = s_n_far2; s_n_far2 = 0; }
            QEntry* t2 = q_far; q_far = q_far2; q_far2 = t2;
            if (tid == 0) ++s_rounds;
            __syncthreads();
        }

        // ---- verify tie candidates against the final labels (reverse star) ----
        __syncthreads();
        const int n_tie = s_n_tie;
        if (n_tie > 0) {
            const bool all_nodes = n_tie > a.tie_cap;
            const int n_check = all_nodes ? N : n_tie;
            for (int i = tid; i < n_check; i += BLOCK) {
                const int w = all_nodes ? i : a.tie_cand[(size_t)blockIdx.x * a.tie_cap + i];
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
        atomicAdd(&a.stats->relaxations, (unsigned long long)c_relax);
        atomicAdd(&a.stats->pops, (unsigned long long)c_pops);
    }
    __syncthreads();
    if (tid == 0) atomicAdd(&a.stats->rounds, (unsigned long long)s_rounds);
}

void launch_sssp(const SsspLaunch& a, cudaStream_t st)
{
    if (a.n_run == 0) return;
    switch (a.block) {
    case 64: sssp_kernel<64, SSSP_UNR><<<a.grid, 64, 0, st>>>(a); break;
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
    case 64: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, sssp_kernel<64, SSSP_UNR>, 64, 0); cudaFuncGetAttributes(&fa, sssp_kernel<64, SSSP_UNR>); break;
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

// the same count on the node-major labels of an origin-bundle batch (the trees were never written per tree): one thread per
// (bundle, node) walks the node's row once for all B origins of the bundle
__global__ void count_edges_nm_kernel(BundleLaunch a)
{
    const int N = a.s.N, B = a.B;
    unsigned long long cnt = 0;
    const size_t total = (size_t)a.n_bundles * N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int bundle = (int)(idx / N);
        const int u = (int)(idx - (size_t)bundle * N);
        const unsigned long long* row = a.bl + idx * B;
        unsigned reached = 0;
        for (int o = 0; o < B; ++o)
            if (a.bundle_tasks[bundle * B + o] >= 0 && __longlong_as_double((long long)row[o]) < DTL_MAX_LABEL_COST) reached |= 1u << o;
        if (!reached) continue;
        const int n_reached = __popc(reached);
        for (int e = a.s.row_ptr[u]; e < a.s.row_ptr[u + 1]; ++e) {
            const Edge ed = ld_edge(&a.s.edges[e]);
            if (ed.to >= 0) { cnt += n_reached; continue; }
            const int tag = a.s.link_tag[ed.link];                     // routing.h:779-786: only the origin zone may use it
            for (int o = 0; o < B; ++o)
                if (((reached >> o) & 1u) && a.s.task_zone[a.bundle_tasks[bundle * B + o]] == tag) ++cnt;
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&a.s.stats->counted_edges, cnt);
}

void launch_count_edges_nm(const BundleLaunch& a, cudaStream_t st)
{
    if (a.n_bundles == 0) return;
    count_edges_nm_kernel<<<148 * 8, 256, 0, st>>>(a);
}

} // namespace dtl

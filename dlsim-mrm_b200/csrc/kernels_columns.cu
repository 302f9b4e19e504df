// K2a  tree -> column builder        (NetworkForSP::backtrace_shortest_path_tree, reference routing.h:413-629)
// K3   column -> link volume scatter (g_reset_and_update_link_volume_based_on_columns, assignment.h:53-195)
// K2b  column-pool gradient projection (g_update_gradient_cost_and_assigned_flow_in_column_pool,
//      assignment.h:565-928)
//
// The column pool is a CSR of columns: fixed slots per OD cell (count, node_sum, n_links, offset, volume)
// and one int32 arena of path links in origin->destination order.  All three kernels are HBM/L2 bound
// gather/scatter work; algorithmic bytes per unit are listed in DESIGN.md.  Compiled with -fmad=false.
#include <algorithm>
#include <cstdlib>

#include "dtl_internal.h"

namespace dtl {

__device__ __forceinline__ double mxd(double a, double b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------------
// K2a: one thread per OD cell of the batch.  Walks the predecessor chain destination -> origin twice:
// first to form the reference's column key node_sum = sum(node + pred_link) (int32 wrap-around,
// routing.h:536) and the path length, then -- only when the key is new for the cell -- to store the
// links reversed (DTA.h:551-559).  Adds od_volume/(k+1) to the column (routing.h:435,502,619) and
// writes the cell's share of the least system travel time (routing.h:517-520).
__global__ void __launch_bounds__(128) backtrace_kernel(ColumnLaunch a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_batch_cells) return;
    const int cell = a.batch_cells[t];
    const int task = a.od_task[cell];
    a.least_contrib[cell] = 0.0;
    if (a.task_skip_backtrace[task]) return; // routing.h:428
    const int lt = task - a.task_first;      // tree index inside the batch
    const int N = a.N, L = a.L;
    const unsigned long long* lab = a.labels + (size_t)lt * N;
    const int* pred = a.pred_link + (size_t)lt * N;
    const double od_volume = a.od_vol[cell];
    if (!(od_volume > 0.000001)) return;     // routing.h:505
    const int dest = a.od_dest_node[cell];
    const float kp = (float)1 / (float)(a.k + 1); // routing.h:435
    const double volume = od_volume * kp;    // routing.h:502

    int pl = pred[dest];
    if (pl >= 0) a.least_contrib[cell] = od_volume * __longlong_as_double((long long)lab[dest]); // :517-520
    unsigned node_sum = 0;
    int n_nodes = 0, n_links = 0;
    int cur = dest;
    while (cur >= 0 && cur < N) {            // routing.h:522-572
        ++n_nodes;
        pl = pred[cur];
        node_sum += (unsigned)cur + (unsigned)pl;
        if (pl >= 0 && pl < L) { ++n_links; cur = a.link_from[pl]; }
        else cur = -1;
    }
    atomicAdd(a.path_nodes, (unsigned long long)n_nodes);
    if (n_nodes < 2) return;                 // routing.h:584
    const int key = (int)node_sum;
    const ColumnPool& P = a.pool;
    const size_t base = (size_t)cell * P.cap;
    int cnt = P.count[cell];
    int slot = -1;
    for (int j = 0; j < cnt; ++j)
        if (P.node_sum[base + j] == key) { slot = j; break; }
    if (slot < 0) {                          // routing.h:595-617 insert a new column
        if (cnt >= P.cap) { atomicOr(a.error_flag, 2); return; }
        const long long need = (n_links + 3) & ~3; // keep every column 16-byte aligned
        const long long off = (long long)atomicAdd(P.arena_used, (unsigned long long)need);
        if (off + need > P.arena_cap) { atomicOr(a.error_flag, 1); return; }
        slot = cnt;
        P.node_sum[base + slot] = key;
        if (P.seq) P.seq[base + slot] = slot;
        P.n_links[base + slot] = n_links;
        P.n_nodes[base + slot] = n_nodes;
        P.offset[base + slot] = off;
        P.volume[base + slot] = 0.0;
        P.cost[base + slot] = 0.0;
        P.time[base + slot] = 0.0;
        int* out = P.arena + off;
        int w = n_links;
        cur = dest;
        while (cur >= 0 && cur < N) {
            pl = pred[cur];
            if (pl >= 0 && pl < L) { out[--w] = pl; cur = a.link_from[pl]; }
            else cur = -1;
        }
        for (int q = n_links; q < need; ++q) out[q] = -1;
        P.count[cell] = cnt + 1;
    }
    P.volume[base + slot] += volume;         // routing.h:619
}

// K2a on the node-major trees an origin-bundle batch leaves behind (bundle_finish_kernel<.., SLIM>): the record of a node row
// holds the predecessor link AND the predecessor node, so a hop of the walk is ONE dependent 8-byte load (two in the
// per-tree layout: pred_link, then link_from).  Same arithmetic and the same column pool updates as backtrace_kernel.
// The walk is done ONCE: its links go to a per-thread stack (local memory: the lanes of a warp advance in lockstep, so a push
// is one coalesced line) and a new column is copied from there, reversed, in 16-byte stores; only paths longer than the stack
// are walked a second time.  The cells of an origin come sorted by the length of their last path (sort_cells_by_path_length).
#define BT_STACK 128
__global__ void __launch_bounds__(128) backtrace_nm_kernel(ColumnLaunch a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_batch_cells) return;
    if (*a.batch_flag) return;               // ties or an overflowed bundle: the host redoes the batch through the per-tree path
    const int cell = a.batch_cells[t];
    const int task = a.od_task[cell];
    a.least_contrib[cell] = 0.0;
    if (a.task_skip_backtrace[task]) return; // routing.h:428
    const int N = a.N, L = a.L, B = a.B;
    const int ts = a.task_slot[task - a.task_first];
    const size_t row0 = (size_t)(ts / B) * N * B + (ts % B); // + node * B
    const int2* pred = a.bp + row0;
    const double od_volume = a.od_vol[cell];
    if (!(od_volume > 0.000001)) return;     // routing.h:505
    const int dest = a.od_dest_node[cell];
    const float kp = (float)1 / (float)(a.k + 1); // routing.h:435
    const double volume = od_volume * kp;    // routing.h:502

    int2 rec = __ldg(&pred[(size_t)dest * B]);
    if (rec.x >= 0) a.least_contrib[cell] = od_volume * __longlong_as_double((long long)__ldg(&a.bl[row0 + (size_t)dest * B])); // :517-520
    int stk[BT_STACK];
    unsigned node_sum = 0;
    int n_nodes = 0, n_links = 0;
    int cur = dest;
    for (;;) {                               // routing.h:522-572
        ++n_nodes;
        node_sum += (unsigned)cur + (unsigned)rec.x;
        if (rec.x < 0 || rec.x >= L) break;
        if (n_links < BT_STACK) stk[n_links] = rec.x;
        ++n_links;
        cur = rec.y;
        if (cur < 0 || cur >= N) break;
        rec = __ldg(&pred[(size_t)cur * B]);
    }
    atomicAdd(a.path_nodes, (unsigned long long)n_nodes);
    if (a.path_len) a.path_len[cell] = n_nodes;
    if (n_nodes < 2) return;                 // routing.h:584
    const int key = (int)node_sum;
    const ColumnPool& P = a.pool;
    const size_t base = (size_t)cell * P.cap;
    int cnt = P.count[cell];
    int slot = -1;
    for (int j = 0; j < cnt; ++j)
        if (P.node_sum[base + j] == key) { slot = j; break; }
    if (slot < 0) {                          // routing.h:595-617 insert a new column
        if (cnt >= P.cap) { atomicOr(a.error_flag, 2); return; }
        const long long need = (n_links + 3) & ~3; // keep every column 16-byte aligned
        const long long off = (long long)atomicAdd(P.arena_used, (unsigned long long)need);
        if (off + need > P.arena_cap) { atomicOr(a.error_flag, 1); return; }
        slot = cnt;
        P.node_sum[base + slot] = key;
        if (P.seq) P.seq[base + slot] = slot;
        P.n_links[base + slot] = n_links;
        P.n_nodes[base + slot] = n_nodes;
        P.offset[base + slot] = off;
        P.volume[base + slot] = 0.0;
        P.cost[base + slot] = 0.0;
        P.time[base + slot] = 0.0;
        int* out = P.arena + off;
        if (n_links <= BT_STACK) {           // origin -> destination order (DTA.h:551-559) = the stack from the top
            for (int q = 0; q < (int)need; q += 4) {
                int4 v;
                v.x = stk[n_links - 1 - q];
                v.y = q + 1 < n_links ? stk[n_links - 2 - q] : -1;
                v.z = q + 2 < n_links ? stk[n_links - 3 - q] : -1;
                v.w = q + 3 < n_links ? stk[n_links - 4 - q] : -1;
                *reinterpret_cast<int4*>(out + q) = v;
            }
        } else {
            int w = n_links;
            cur = dest;
            while (w > 0) {
                rec = __ldg(&pred[(size_t)cur * B]);
                out[--w] = rec.x;
                cur = rec.y;
            }
            for (int q = n_links; q < need; ++q) out[q] = -1;
        }
        P.count[cell] = cnt + 1;
    }
    P.volume[base + slot] += volume;         // routing.h:619
}

// The OD cells of every origin, ordered by the length of the path the back-trace last found for them, longest first: the lanes
// of a warp then walk paths of similar length (a warp waits for its longest walk) and still belong to ONE tree, whose trunk
// they share through L1/L2.  (Sorting the cells of the whole batch by length was measured slower than no sorting, 0.70 against
// 0.49 ms: it scatters the cells of a tree over the launch.)  Path lengths barely change between iterations, so this runs once
// per batch.  One CTA per origin, rank by counting; origins with more cells than fit the shared arrays keep their order.
#define CELL_SORT_MAX 4096
__global__ void __launch_bounds__(256) cell_sort_in_task_kernel(const int* path_len, int* cells, const int* cell_ptr, int first)
{
    __shared__ int s_cell[CELL_SORT_MAX], s_key[CELL_SORT_MAX];
    const int c0 = cell_ptr[blockIdx.x] - first, n = cell_ptr[blockIdx.x + 1] - first - c0;
    if (n < 2 || n > CELL_SORT_MAX) return;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = cells[c0 + i];
        s_cell[i] = c;
        s_key[i] = path_len[c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int k = s_key[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const int kj = s_key[j];
            rank += (kj > k || (kj == k && j < i)) ? 1 : 0;
        }
        cells[c0 + rank] = s_cell[i];
    }
}
// cells: the batch's cell list; cell_ptr: [n_tasks + 1] offsets of the batch's tasks into the engine's cell list, `first` = the
// offset of the batch's first cell
void sort_cells_by_path_length(const int* path_len, int* cells, const int* cell_ptr, int first, int n_tasks, cudaStream_t st)
{
    if (n_tasks <= 0) return;
    cell_sort_in_task_kernel<<<n_tasks, 256, 0, st>>>(path_len, cells, cell_ptr, first);
}

__global__ void __launch_bounds__(256) induced_delay_kernel(signed char* design_flag, const double* doc, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && design_flag[i] == 0 && doc[i] > 3) design_flag[i] = -2;
}
void launch_induced_delay_flags(signed char* design_flag, const double* doc, size_t n, cudaStream_t st)
{
    if (n) induced_delay_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(design_flag, doc, n);
}

// gates of the pipelined loop (engine.cu: iterations_impl).  *go = tag when none of the n flag words of a search is set, else 0 (the
// main stream waits for tag before it rewrites the link costs).  *start = tag lets the side stream begin the back-trace: at once when
// a flag is set (the batch is redone first, nothing of the next iteration will run meanwhile) or when the next search does not
// announce itself (start_when_clean); otherwise the next search writes it when all its CTAs are resident.
__global__ void __launch_bounds__(256) set_go_kernel(const int* flags, int n, unsigned* go, unsigned* start, unsigned tag, int start_when_clean)
{
    int any = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) any |= flags[i] != 0;
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) {
        __threadfence();
        if (any || start_when_clean) *reinterpret_cast<volatile unsigned*>(start) = tag;
        *reinterpret_cast<volatile unsigned*>(go) = any ? 0u : tag;
    }
}
void launch_set_go(const int* flags, int n, unsigned* go, unsigned* start, unsigned tag, int start_when_clean, cudaStream_t st)
{
    set_go_kernel<<<1, 256, 0, st>>>(flags, n, go, start, tag, start_when_clean);
}

void launch_backtrace(const ColumnLaunch& a, cudaStream_t st)
{
    if (a.n_batch_cells == 0) return;
    if (a.bp) backtrace_nm_kernel<<<(a.n_batch_cells + 127) / 128, 128, 0, st>>>(a);
    else backtrace_kernel<<<(a.n_batch_cells + 127) / 128, 128, 0, st>>>(a);
}

// ------------------------------------------------------------------------------------------------
// K3: one warp per OD cell.  A lane reads FOUR consecutive path links of a column per step (columns are 16-byte aligned in the
// arena and padded with -1: one 16-byte load, 512 bytes per warp step) and adds the column's PCE volume to the per-link
// Q31.32 accumulators with integer atomics, so the sums do not depend on the order columns are visited in -- across warps,
// launches or GPUs.  The addend is the reference's own  (double)(float)path_volume * pce  (assignment.h:87,127,140)
// rounded once to 2^-32.  Warps are dealt to cells with a stride (a.cell_stride, coprime to the number of cells): the
// cells of one origin all run over the same first links, and neighbouring warps hammering the same accumulators serialise
// in the L2 atomic units.
#ifndef SCATTER_MIN_BLOCKS
#define SCATTER_MIN_BLOCKS 4
#endif
template <bool PERSON>
__global__ void __launch_bounds__(256, PERSON ? 3 : SCATTER_MIN_BLOCKS) scatter_kernel(ScatterLaunch a)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= a.pool.n_od) return;
    const int cell = (int)(((long long)warp * a.cell_stride) % a.pool.n_od);
    const ColumnPool& P = a.pool;
    const int cnt = P.count[cell];
    if (cnt == 0) return;
    const int at = a.od_at[cell], tau = a.od_tau[cell];
    const double* pce = a.pce + ((size_t)tau * a.AT + at) * a.L;
    const double* occ = a.occ + ((size_t)tau * a.AT + at) * a.L;
    // one PCE / occupancy for every link of this (period, agent type) is the common case: no gather then (same product)
    const double pce_u = a.pce_uni[tau * a.AT + at], occ_u = PERSON ? a.occ_uni[tau * a.AT + at] : 0.0;
    const bool pce_same = pce_u == pce_u, occ_same = occ_u == occ_u;
    unsigned long long* acc = a.pce_fix + (size_t)tau * a.L;
    unsigned long long* per = PERSON ? a.person_fix + (size_t)tau * (1 + a.AT) * a.L : nullptr;
    unsigned long long* per_at = PERSON ? per + (size_t)(1 + at) * a.L : nullptr;
    const size_t base = (size_t)cell * P.cap;
    unsigned long long refs = 0;
    // slot records of the cell's columns: lane j holds column j (cells have at most a few dozen columns)
    for (int j0 = 0; j0 < cnt; j0 += 32) {
        const int jm = min(32, cnt - j0);
        double vol_l = 0.0;
        int n_l = 0;
        long long off_l = 0;
        if (lane < jm) {
            vol_l = P.volume[base + j0 + lane];
            n_l = P.n_links[base + j0 + lane];
            off_l = P.offset[base + j0 + lane];
            if (a.decay_k >= 0 && (a.decay_cell ? a.decay_cell[cell] != 0 : (!a.decay_at || a.decay_at[at]))) // assignment.h:182-186 MSA self-reduction (:150-176 in the SA stage)
                P.volume[base + j0 + lane] = vol_l * ((double)a.decay_k / (double)(a.decay_k + 1));
        }
        // four columns per step: their link groups are requested together (a warp otherwise has ONE load in flight per column
        // and spends its time waiting for it); a lane's first group covers columns of up to 128 links, longer ones take the tail loop
        for (int j = 0; j < jm; j += 4) {
            int4 v[4];
            float f[4];
            int n[4];
            const int4* links[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int jj = (j + c) & 31;
                const double vol = __shfl_sync(0xffffffffu, vol_l, jj);
                n[c] = __shfl_sync(0xffffffffu, n_l, jj);
                const long long off = __shfl_sync(0xffffffffu, off_l, jj);
                f[c] = (float)vol;                                   // assignment.h:87,127
                if (j + c >= jm) n[c] = 0;
                refs += n[c];
                if (f[c] == 0.0f) n[c] = 0;
                links[c] = reinterpret_cast<const int4*>(P.arena + off);
                v[c] = make_int4(-1, -1, -1, -1);
                if (lane * 4 < n[c]) v[c] = __ldg(&links[c][lane]);
            }
            // first link group of the four columns: the columns of a cell are alternative paths of ONE origin-destination pair
            // and share their first links position by position, so a lane often holds the same link in several columns --
            // their addends are summed before the atomic (integer sums: the same bits)
            unsigned long long x_u[4], y_u[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                x_u[c] = (unsigned long long)__double2ll_rn((double)f[c] * pce_u * DTL_FIX_SCALE);
                y_u[c] = PERSON ? (unsigned long long)__double2ll_rn((double)f[c] * occ_u * DTL_FIX_SCALE) : 0ULL;
            }
            if (pce_same && (!PERSON || occ_same)) {
                const int ls[4][4] = { { v[0].x, v[0].y, v[0].z, v[0].w }, { v[1].x, v[1].y, v[1].z, v[1].w },
                                       { v[2].x, v[2].y, v[2].z, v[2].w }, { v[3].x, v[3].y, v[3].z, v[3].w } };
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    int l[4] = { ls[0][r], ls[1][r], ls[2][r], ls[3][r] };
                    unsigned long long x[4] = { x_u[0], x_u[1], x_u[2], x_u[3] };
                    unsigned long long y[4] = { y_u[0], y_u[1], y_u[2], y_u[3] };
#pragma unroll
                    for (int c = 3; c > 0; --c)
#pragma unroll
                        for (int c2 = 0; c2 < c; ++c2)
                            if (l[c] >= 0 && l[c] == l[c2]) { x[c2] += x[c]; y[c2] += y[c]; l[c] = -1; }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (l[c] < 0) continue;                      // padding, nothing for this lane, or merged into an earlier column
                        atomicAdd(&acc[l[c]], x[c]);
                        if (PERSON) {
                            atomicAdd(&per[l[c]], y[c]);             // assignment.h:141
                            atomicAdd(&per_at[l[c]], y[c]);          // assignment.h:142
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) v[c] = make_int4(-1, -1, -1, -1);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                for (int q = lane;;) {
                    const int ls[4] = { v[c].x, v[c].y, v[c].z, v[c].w };
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int l = ls[r];
                        if (l < 0) continue;                         // padding of the last group / nothing for this lane
                        const unsigned long long x =
                            pce_same ? x_u[c] : (unsigned long long)__double2ll_rn(f[c] * __ldg(&pce[l]) * DTL_FIX_SCALE); // assignment.h:140
                        atomicAdd(&acc[l], x);
                        if (PERSON) {
                            const unsigned long long y =
                                occ_same ? y_u[c] : (unsigned long long)__double2ll_rn(f[c] * __ldg(&occ[l]) * DTL_FIX_SCALE);
                            atomicAdd(&per[l], y);                   // assignment.h:141
                            atomicAdd(&per_at[l], y);                // assignment.h:142
                        }
                    }
                    q += 32;
                    if (q * 4 >= n[c]) break;
                    v[c] = __ldg(&links[c][q]);
                }
            }
        }
    }
    if (lane == 0) atomicAdd(a.link_refs, refs);
}

static int coprime_stride(int n)
{
    // a stride near n / golden ratio, coprime to n: consecutive warps get cells of different origins, every cell exactly once
    if (n < 64) return 1;
    int s = (int)(n * 0.6180339887498949);
    auto gcd = [](int x, int y) { while (y) { const int t = x % y; x = y; y = t; } return x; };
    while (gcd(s, n) != 1) ++s;
    return s;
}

void launch_scatter(const ScatterLaunch& a0, cudaStream_t st)
{
    if (a0.pool.n_od == 0) return;
    ScatterLaunch a = a0;
    static const bool in_order = getenv("DTL_SCATTER_IN_ORDER") != nullptr; // experiment hook: cells in origin order
    a.cell_stride = in_order ? 1 : coprime_stride(a.pool.n_od);
    const int blocks = (a.pool.n_od + 7) / 8;
    if (a.person_fix) scatter_kernel<true><<<blocks, 256, 0, st>>>(a);
    else scatter_kernel<false><<<blocks, 256, 0, st>>>(a);
}

// ------------------------------------------------------------------------------------------------
// K2b, three launches.  (1) per (period, agent type, link): {travel time, first-order gradient cost} with the reference's
// expression  tt + penalty + toll / VOT * 60  (DTA.h:1157) evaluated ONCE per link instead of once per column link, in one
// 16-byte record so that a column link costs one gather.  (2) one thread per column slot: path cost / time as
// left-to-right FP64 sums in path order exactly like the reference (DTA.h:1154-1164, assignment.h:723-736).  (3) one thread
// per OD cell: arg-min ("first strict minimum in ascending node_sum order", assignment.h:757) and the flow shifts.
__global__ void __launch_bounds__(256) pool_link_cost_kernel(PoolUpdateLaunch a)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)a.T * a.AT * a.L;
    if (i >= n) return;
    const int l = (int)(i % a.L);
    const int ta = (int)(i / a.L); // tau * AT + at
    const int tau = ta / a.AT, at = ta - tau * a.AT;
    const double t = a.tt[(size_t)tau * a.L + l];
    a.link_cost[i] = make_double2(t, t + a.penalty[(size_t)tau * a.L + l] + a.toll[i] / a.vot[at] * 60); // DTA.h:1157
}

__global__ void __launch_bounds__(128) pool_column_cost_kernel(PoolUpdateLaunch a)
{
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const ColumnPool& P = a.pool;
    if (slot >= (size_t)P.n_od * P.cap) return;
    const int cell = (int)(slot / P.cap);
    if ((int)(slot - (size_t)cell * P.cap) >= P.count[cell]) return;
    const double2* lc = a.link_cost + ((size_t)a.od_tau[cell] * a.AT + a.od_at[cell]) * a.L;
    const int n = P.n_links[slot];
    const int4* links = reinterpret_cast<const int4*>(P.arena + P.offset[slot]);
    double path_cost = 0, path_time = 0;
    const bool watch = a.sa && a.rt_info[a.od_at[cell]] == 1;          // assignment.h:699-713
    const signed char* flag = a.design_flag + (size_t)a.od_tau[cell] * a.L;
    bool impacted = false;
    for (int q = 0; q < n; q += 4) {
        const int4 v = __ldg(&links[q >> 2]);
        const int ls[4] = { v.x, v.y, v.z, v.w };
        double2 c[4];
        if (watch)
#pragma unroll
            for (int r = 0; r < 4; ++r) impacted = impacted || (q + r < n && flag[ls[r]] == -1);
#pragma unroll
        for (int r = 0; r < 4; ++r) // the four gathers of a 16-byte group of link ids are issued together
            c[r] = q + r < n ? __ldg(&lc[ls[r]]) : make_double2(0.0, 0.0);
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (q + r < n) { path_time += c[r].x; path_cost += c[r].y; }
    }
    P.cost[slot] = path_cost;
    P.time[slot] = path_time;
    if (impacted) a.od_impact[cell] = 1;
}

__global__ void __launch_bounds__(128) pool_update_kernel(PoolUpdateLaunch a)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= a.pool.n_od) return;
    const ColumnPool& P = a.pool;
    const int cnt = P.count[cell];
    double* sums = a.od_sums + (size_t)cell * 4;
    double s_gap = 0, s_cost = 0, s_time = 0, s_dem = 0;
    if (cnt > 0) {
        const size_t base = (size_t)cell * P.cap;
        const double od_volume = a.od_vol[cell];
        // pass 2: arg-min in ascending node_sum order with strict '<' (std::map order, assignment.h:686-767)
        double least_cost = 999999;
        int least = -1, least_key = 0;
        for (int j = 0; j < cnt; ++j) {
            const double c = P.cost[base + j];
            const double v = P.volume[base + j];
            const int key = P.node_sum[base + j];
            s_time += P.time[base + j] * v;                              // :746
            s_dem += v;                                                  // :747
            if (cnt == 1) s_cost += c * v;                               // :750-754
            if (c < 999999 && (least < 0 || c < least_cost || (c == least_cost && key < least_key))) {
                least_cost = c; least = j; least_key = key;
            }
        }
        if (cnt >= 2) {                                                  // :769
            double switched = 0;
            // the reference walks the std::map, i.e. ascending node_sum: keep that order for the FP sums
            int prev_key = 0;
            bool first = true;
            for (int step = 0; step < cnt; ++step) {
                int j = -1, key = 0;
                for (int c2 = 0; c2 < cnt; ++c2) { // next column in ascending key order (cnt is small)
                    const int k2 = P.node_sum[base + c2];
                    if ((first || k2 > prev_key) && (j < 0 || k2 < key)) { j = c2; key = k2; }
                }
                prev_key = key; first = false;
                if (P.seq && least >= 0 ? P.seq[base + j] == P.seq[base + least] : j == least) continue; // :775 (path_seq_no != least path_seq_no)
                const double c = P.cost[base + j];
                const double v = P.volume[base + j];
                const double diff = c - least_cost;
                const double rel = diff / mxd(0.0001, least_cost);       // :781
                s_gap += diff * v;                                       // :786
                s_cost += c * v;                                         // :787
                const double stepsz = 1.0 / (a.n + 2) * od_volume;       // :802
                double shift = stepsz * mxd(0.0000, rel);                // :808
                if (rel > 3 * 60) shift = v;                             // :811
                if (shift > v) shift = v;                                // :816
                if (a.sa && !a.od_impact[cell]) shift = 0;               // :828-846
                const double nv = mxd(0.0, v - shift);                   // :851
                P.volume[base + j] = nv;
                switched += (v - nv);                                    // :859
            }
            if (least >= 0) {                                            // :869-886
                const double nv = P.volume[base + least] + switched;
                P.volume[base + least] = nv;
                s_cost += P.cost[base + least] * nv;
            }
        }
    }
    sums[0] = s_gap; sums[1] = s_cost; sums[2] = s_time; sums[3] = s_dem;
}

void launch_pool_update(const PoolUpdateLaunch& a, cudaStream_t st)
{
    if (a.pool.n_od == 0) return;
    const size_t n_lc = (size_t)a.T * a.AT * a.L, n_slots = (size_t)a.pool.n_od * a.pool.cap;
    pool_link_cost_kernel<<<(unsigned)((n_lc + 255) / 256), 256, 0, st>>>(a);
    pool_column_cost_kernel<<<(unsigned)((n_slots + 127) / 128), 128, 0, st>>>(a);
    pool_update_kernel<<<(a.pool.n_od + 127) / 128, 128, 0, st>>>(a);
}

} // namespace dtl

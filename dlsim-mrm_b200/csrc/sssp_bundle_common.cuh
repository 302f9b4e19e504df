// Shared device helpers of the origin-bundle K1 kernels (kernels_sssp_bundle.cu: round-synchronous search and the
// label->tree pass; kernels_sssp_bundle_async.cu: the search without round barriers).
#pragma once
#include "sssp_common.cuh"

namespace dtl {

#ifndef BND_U
#define BND_U 4 /* out-edges of a node relaxed per pass */
#endif
#define BND_DEAD_KEY 0xffffffffu
#ifndef BND_PRECHECK
#define BND_PRECHECK 0 /* 1: load the head's label row first and issue atomics only for origins that improve (one more round trip, far fewer atomics) */
#endif

// per-node queue bits in shared memory, two per node: bit 0 = the node has an entry in the near queue of this or the next
// round, bit 1 = it has an entry in the far pile
#define ST_WORD(n) ((n) >> 4)
#define ST_SHIFT(n) (((n) & 15) << 1)

// scheduling key of a label: the high word of the FP64 bit pattern (non-negative doubles order like unsigned integers;
// 20 mantissa bits are plenty for deciding which label window a node belongs to -- the key never touches a label)
__device__ __forceinline__ unsigned key_of(unsigned long long bits) { return (unsigned)(bits >> 32); }

// 64-bit atomicMin on global memory that is predicated instead of branched around; when `on` is false nothing is issued and
// the candidate itself comes back, which reads as "no improvement" to the caller
__device__ __forceinline__ unsigned long long atom_min_if(unsigned long long* p, unsigned long long v, bool on)
{
    unsigned long long old = v;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q atom.global.min.u64 %0, [%1], %2;\n\t}"
                 : "+l"(old) : "l"(p), "l"(v), "r"((unsigned)on) : "memory");
    return old;
}

template <int BLOCK>
__device__ __forceinline__ unsigned block_min_u32(unsigned v, unsigned* s_red)
{
    v = __reduce_min_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned w = threadIdx.x < BLOCK / 32 ? s_red[threadIdx.x] : 0xffffffffu;
        w = __reduce_min_sync(0xffffffffu, w);
        if (threadIdx.x == 0) s_red[0] = w;
    }
    __syncthreads();
    const unsigned r = s_red[0];
    __syncthreads();
    return r;
}

template <int G>
__device__ __forceinline__ unsigned group_min_u32(unsigned v)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Minimum over the OPL lanes of an edge subgroup of EPL values per lane, transposed: every exchange step halves the values a
// lane still carries until one is left, then the remaining steps are a plain butterfly.  Lane `op` ends up with the minimum of
// value index op / (OPL / EPL) (the high bits of op pick the value), e.g. 4 shuffles instead of 12 for EPL = 4, OPL = 8.
template <int OPL, int EPL>
__device__ __forceinline__ unsigned subgroup_min_edges(const unsigned (&v)[EPL], int op)
{
    static_assert((OPL == 4 || OPL == 8) && (EPL == 1 || EPL == 2 || EPL == 4) && EPL <= OPL, "lanes per edge, edges per lane");
    unsigned w[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) w[j] = v[j];
    int n = EPL;
#pragma unroll
    for (int o = OPL / 2; o > 0; o >>= 1) {
        if (n > 1) {
            const bool hi = (op & o) != 0;
            n >>= 1;
#pragma unroll
            for (int j = 0; j < EPL / 2; ++j)
                if (j < n) w[j] = min(hi ? w[j + n] : w[j], __shfl_xor_sync(0xffffffffu, hi ? w[j] : w[j + n], o));
        } else {
            w[0] = min(w[0], __shfl_xor_sync(0xffffffffu, w[0], o));
        }
    }
    return w[0];
}

// The node-major label -> predecessor pass of the assignment loop (trees are only back-traced, no per-tree arrays): a group
// of B/2 lanes per node, two origins per lane.  The node's first four in-edges come as one 32-byte record (ForwardStar::rin4),
// their edge records and the label rows of the upstream nodes are all requested at once, and the group writes one coalesced
// row of {predecessor link, predecessor node} records.  The tight in-edge is the predecessor (label[u] + cost == label[v] in
// exact FP64; of parallel links from one node the first in scan order); two tight upstream nodes are an order-dependent tie
// and, like an overflowed bundle, raise a.batch_flag: the host then redoes the batch through bundle_finish_kernel.
// Nodes [v_begin, v_end) of one bundle by the BLOCK threads of the calling CTA (BLOCK / (B/2) nodes per step).  SAME_LAUNCH: the
// labels were written by other CTAs of the running kernel (the pass runs inside the search kernel), so they are read through
// L2 instead of the read-only path.
template <int B, int BLOCK, bool SAME_LAUNCH>
__device__ __forceinline__ void bundle_pred_nodes(const BundleLaunch& a, int bundle, int v_begin, int v_end, bool report_overflow)
{
    auto ld_row = [](const ulonglong2* p) { return SAME_LAUNCH ? __ldcg(p) : __ldg(p); };
    constexpr int G = B / 2;
    constexpr int NODES = BLOCK / G;   // nodes per step
    const int tid = threadIdx.x;
    const int g = tid & (G - 1);
    const int N = a.s.N;
    int task[2], origin[2], zone[2];
    bool bad = false;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        task[t] = __ldg(&a.bundle_tasks[bundle * B + 2 * g + t]);
        if (task[t] >= 0 && __ldcg(&a.s.tie_nodes[task[t]]) < 0) { task[t] = -1; bad = true; }
        origin[t] = task[t] >= 0 ? __ldg(&a.s.task_origin[task[t]]) : -1;
        zone[t] = task[t] >= 0 ? __ldg(&a.s.task_zone[task[t]]) : -1;
    }
    if (bad && report_overflow && tid < G) atomicOr(a.batch_flag, 2);
    const ulonglong2* const bl2 = reinterpret_cast<const ulonglong2*>(a.bl + (size_t)bundle * N * B) + g;
    int4* const out = reinterpret_cast<int4*>(a.bp + (size_t)bundle * N * B) + g;
    // destination-bounded search: labels whose key is not below the bundle's bound are not final (treated as unreached)
    const unsigned bound = a.bundle_bound ? (SAME_LAUNCH ? __ldcg(&a.bundle_bound[bundle]) : __ldg(&a.bundle_bound[bundle])) : BND_DEAD_KEY;
    bool tie = false;
    for (int v = v_begin + tid / G; v < v_end; v += NODES) {
        const ulonglong2 lvb = ld_row(bl2 + (size_t)v * G);
        const int4 r01 = __ldg(reinterpret_cast<const int4*>(a.s.rin4 + (size_t)v * 4));
        const int4 r23 = __ldg(reinterpret_cast<const int4*>(a.s.rin4 + (size_t)v * 4) + 1);
        const bool need0 = task[0] >= 0 && lvb.x < INF_BITS && v != origin[0] && key_of(lvb.x) < bound;
        const bool need1 = task[1] >= 0 && lvb.y < INF_BITS && v != origin[1] && key_of(lvb.y) < bound;
        // a search that stopped early: a node beyond the bound for both origins of this lane gets no record at all (nothing is requested,
        // nothing written -- the row keeps whatever an earlier iteration left there).  The back-trace never comes by: it walks from
        // destinations below the bound (that is when the search stopped) towards smaller labels.  The origin rows are always written.
        if (bound != BND_DEAD_KEY && !need0 && !need1 && v != origin[0] && v != origin[1]) continue;
        const int from[4] = { r01.x, r01.z, r23.x, r23.z };
        const int eidx[4] = { r01.y, r01.w, r23.y, r23.w & 0x3fffffff };
        ulonglong2 lub[4];
        int4 ed[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            lub[q] = make_ulonglong2(INF_BITS, INF_BITS);
            ed[q] = make_int4(0, 0, 0, 0);
            if (from[q] >= 0) {
                lub[q] = ld_row(bl2 + (size_t)from[q] * G);
                ed[q] = __ldg(reinterpret_cast<const int4*>(&a.s.edges[eidx[q]]));
            }
        }
        const double lv0 = __longlong_as_double((long long)lvb.x), lv1 = __longlong_as_double((long long)lvb.y);
        int first0 = -1, first1 = -1, link0 = DTL_NO_PRED, link1 = DTL_NO_PRED;
        auto consider = [&](int u, double cost, int link, bool tagged, ulonglong2 lu) {
            bool ok0 = need0 && lu.x < INF_BITS, ok1 = need1 && lu.y < INF_BITS;
            if (tagged) { // routing.h:779-786
                const int tag = __ldg(&a.s.link_tag[link]);
                ok0 = ok0 && tag == zone[0];
                ok1 = ok1 && tag == zone[1];
            }
            if (ok0 && __longlong_as_double((long long)lu.x) + cost == lv0) {
                if (first0 < 0) { first0 = u; link0 = link; }
                else if (u != first0) tie = true;
            }
            if (ok1 && __longlong_as_double((long long)lu.y) + cost == lv1) {
                if (first1 < 0) { first1 = u; link1 = link; }
                else if (u != first1) tie = true;
            }
        };
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (from[q] >= 0) consider(from[q], __hiloint2double(ed[q].y, ed[q].x), ed[q].w, ed[q].z < 0, lub[q]);
        if (r23.w & 0x40000000) { // more than four in-edges: the rest through the reverse star
            const int j1 = __ldg(&a.s.in_ptr[v + 1]);
            for (int jj = __ldg(&a.s.in_ptr[v]) + 4; jj < j1; ++jj) {
                const int2 ie = __ldg(&a.s.in_edges[jj]);
                const Edge e2 = ld_edge(&a.s.edges[ie.y]);
                consider(ie.x, e2.cost, e2.link, e2.to < 0, ld_row(bl2 + (size_t)ie.x * G));
            }
        }
        out[(size_t)v * G] = make_int4(link0, first0, link1, first1);
    }
    if (tie) atomicOr(a.batch_flag, 1);
}


} // namespace dtl

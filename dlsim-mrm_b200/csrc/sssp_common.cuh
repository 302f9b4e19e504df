// Shared device helpers of the two K1 (shortest-path tree) kernels: kernels_sssp.cu (global-memory queues, many trees
// per SM) and kernels_sssp_smq.cu (shared-memory queues, few trees per SM).
#pragma once
#include "dtl_internal.h"

namespace dtl {

#define INF_BITS 0x430c6bf526340000ULL /* bit pattern of 1.0e15 */
#define TIE_MARK (-7777)
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


} // namespace dtl

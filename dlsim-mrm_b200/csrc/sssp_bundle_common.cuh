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

} // namespace dtl

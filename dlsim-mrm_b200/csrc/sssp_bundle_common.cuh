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

// Group-wide minimum of four values per lane, transposed: after log2(G) exchange steps the lane holds the minimum, over the
// G lanes of its group, of the value with index (g >> 1) for G = 8 and index g for G = 4 -- each step halves the values a
// lane still carries (4 shuffles for G = 8 instead of the 12 of four separate butterfly reductions).
template <int G>
__device__ __forceinline__ unsigned group_min_edges(const unsigned (&v)[4], int g)
{
    static_assert(G == 4 || G == 8, "groups of 4 or 8 lanes");
    unsigned a0, a1;
    if (G == 8) {
        const bool hi = (g & 4) != 0;
        a0 = min(hi ? v[2] : v[0], __shfl_xor_sync(0xffffffffu, hi ? v[0] : v[2], 4));
        a1 = min(hi ? v[3] : v[1], __shfl_xor_sync(0xffffffffu, hi ? v[1] : v[3], 4));
        const bool h2 = (g & 2) != 0;
        unsigned b = min(h2 ? a1 : a0, __shfl_xor_sync(0xffffffffu, h2 ? a0 : a1, 2));
        return min(b, __shfl_xor_sync(0xffffffffu, b, 1));
    } else {
        const bool hi = (g & 2) != 0;
        a0 = min(hi ? v[2] : v[0], __shfl_xor_sync(0xffffffffu, hi ? v[0] : v[2], 2));
        a1 = min(hi ? v[3] : v[1], __shfl_xor_sync(0xffffffffu, hi ? v[1] : v[3], 2));
        const bool h1 = (g & 1) != 0;
        return min(h1 ? a1 : a0, __shfl_xor_sync(0xffffffffu, h1 ? a0 : a1, 1));
    }
}

} // namespace dtl

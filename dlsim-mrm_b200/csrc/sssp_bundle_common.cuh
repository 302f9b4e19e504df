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


} // namespace dtl

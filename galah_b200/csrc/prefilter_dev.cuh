// Device-side helpers shared by the K2 kernels (prefilter.cu, prefilter_join.cu).
#pragma once
#include "common.cuh"
#include "prefilter.cuh"

namespace gb200 {

// ------------------------------------------------------------------------------------------
// device: warp-cooperative exact intersection of two sorted distinct u64 lists
// ------------------------------------------------------------------------------------------

// Merge-path intersection.  All 32 lanes call with the same (A, na, B, nb); returns |A n B| on
// every lane.  Ties go to A first, so a common value is seen as "take b while the previous a
// equals it"; that test also works across lane boundaries because A[i-1] is re-read.
template <typename PtrT>
__device__ __forceinline__ uint32_t warp_intersect(PtrT A, uint32_t na, PtrT B, uint32_t nb) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t len = na + nb;
    const uint32_t per = (len + 31) >> 5;
    const uint32_t d0 = min(lane * per, len);
    const uint32_t d1 = min(d0 + per, len);
    // partition: smallest i with !(A[i] <= B[d0-1-i])
    uint32_t lo = d0 > nb ? d0 - nb : 0, hi = min(d0, na);
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (A[mid] <= B[d0 - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    uint32_t i = lo, j = d0 - lo;
    uint64_t a = i < na ? A[i] : 0, b = j < nb ? B[j] : 0;
    uint64_t a_prev = i > 0 ? A[i - 1] : 0;
    bool have_prev = i > 0;
    uint32_t common = 0;
    for (uint32_t t = d0; t < d1; t++) {
        bool take_a = (j >= nb) || (i < na && a <= b);
        if (take_a) {
            a_prev = a; have_prev = true; i++;
            if (i < na) a = A[i];
        } else {
            common += (have_prev && a_prev == b) ? 1u : 0u;
            j++;
            if (j < nb) b = B[j];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) common += __shfl_xor_sync(0xffffffffu, common, o);
    return common;
}

// #{x in X[0..nx) : x <= v}
template <typename PtrT>
__device__ __forceinline__ uint32_t upper_rank(PtrT X, uint32_t nx, uint64_t v) {
    uint32_t lo = 0, hi = nx;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (X[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Lane 0 finishes a pair: conservative test, exact `total`, append.
template <typename PtrT>
__device__ __forceinline__ void finish_pair(const KernelParams &p, uint32_t gi, uint32_t gj, PtrT A,
                                            uint32_t na, PtrT B, uint32_t nb, uint32_t common) {
    const uint32_t tmin = min(na, nb);
    if (common < p.cmin_by_tmin[tmin]) return;
    uint32_t total;
    if (na == 0 || nb == 0) {
        total = 0;  // the reference's loop body never runs: i = j = 0
    } else {
        const uint64_t amax = A[na - 1], bmax = B[nb - 1];
        if (amax <= bmax) total = na + upper_rank(B, nb, amax) - common;
        else total = nb + upper_rank(A, na, bmax) - common;
    }
    if (common < p.cmin_by_total[total]) return;
    unsigned long long slot = atomicAdd(p.n_cand, 1ull);
    if (slot < p.cand_cap) p.cand[slot] = make_uint4(gi, gj, common, total);
}


}  // namespace gb200

// Instantiates align_kernel<K, DEBUG> for K = 1..KMAX and dispatches a launch to the right one.
#include "wxb_kernel.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace wxb {

template <int K, bool DEBUG>
static cudaError_t configure(size_t smem) {
    // Every instantiation asks for the maximum shared-memory carveout: kernels that want different
    // L1/shared splits cannot be co-resident on an SM, and a plan runs one launch per
    // (warps, tokens-per-lane) class concurrently.
    static size_t configured_for = 0;
    static bool first = true;
    if (!first && smem <= configured_for) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(align_kernel<K, DEBUG>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(align_kernel<K, DEBUG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        configured_for = 227 * 1024;
    } else {
        configured_for = std::max<size_t>(configured_for, 48 * 1024);
    }
    first = false;
    return cudaSuccess;
}

template <int K>
static cudaError_t launch_one(int warps, int nb, int te, const Params &p, int first, int count, size_t smem, cudaStream_t st, bool debug) {
    cudaError_t e = debug ? configure<K, true>(smem) : configure<K, false>(smem);
    if (e != cudaSuccess) return e;
    const int threads = 32 * (warps + 1);            // + the producer warp
    if (getenv("WXB_DEBUG_OCC")) {
        int nbk = 0;
        cudaFuncAttributes fa;
        if (debug) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbk, align_kernel<K, true>, threads, smem); cudaFuncGetAttributes(&fa, align_kernel<K, true>); }
        else { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbk, align_kernel<K, false>, threads, smem); cudaFuncGetAttributes(&fa, align_kernel<K, false>); }
        fprintf(stderr, "[wxb] align_kernel<%d,%d> W=%d NB=%d TE=%d: regs=%d smem=%zu grid=%d -> %d CTAs/SM\n", K, (int)debug, warps, nb, te, fa.numRegs, smem, count, nbk);
    }
    if (debug) align_kernel<K, true><<<count, threads, smem, st>>>(p, first, nb, te);
    else       align_kernel<K, false><<<count, threads, smem, st>>>(p, first, nb, te);
    return cudaGetLastError();
}

template <int K>
static cudaError_t dispatch(int k, int warps, int nb, int te, const Params &p, int first, int count, size_t smem, cudaStream_t st, bool debug) {
    if (k == K) return launch_one<K>(warps, nb, te, p, first, count, smem, st, debug);
    if constexpr (K < KMAX) return dispatch<K + 1>(k, warps, nb, te, p, first, count, smem, st, debug);
    else return cudaErrorInvalidValue;
}

cudaError_t launch_align(int k, int warps, int nb, int te, const Params &p, int first, int count, size_t smem, cudaStream_t st, bool debug) {
    if (warps < 1 || warps > WMAX || nb < 2) return cudaErrorInvalidValue;
    return dispatch<1>(k, warps, nb, te, p, first, count, smem, st, debug);
}

}  // namespace wxb

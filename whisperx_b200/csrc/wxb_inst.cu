// Instantiates align_kernel<K, DEBUG, WILD> for K = 1..KMAX and dispatches a launch to the right one.
#include "wxb_kernel.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace wxb {

template <int K, bool DEBUG, bool WILD>
static cudaError_t configure(size_t smem) {
    // Every instantiation asks for the maximum shared-memory carveout: kernels that want different
    // L1/shared splits cannot be co-resident on an SM, and a plan runs one launch per
    // (warps, tokens-per-lane) class concurrently.
    // Function attributes are per *device*: the cache is keyed on the current device (one engine per GPU may
    // live in the same process) and guarded, since contexts may launch from different threads.
    constexpr int MAX_DEV = 64;
    static std::mutex mu;
    static size_t configured_for[MAX_DEV] = {};      // 0 = this device has not been configured yet
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool cached = dev >= 0 && dev < MAX_DEV;
    std::lock_guard<std::mutex> lock(mu);
    if (cached && configured_for[dev] != 0 && smem <= configured_for[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(align_kernel<K, DEBUG, WILD>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    size_t now = 48 * 1024;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(align_kernel<K, DEBUG, WILD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        now = 227 * 1024;
    }
    if (cached) configured_for[dev] = std::max(configured_for[dev], now);
    return cudaSuccess;
}

template <int K, bool DEBUG, bool WILD>
static cudaError_t launch_variant(int warps, int nb, int te, const Params &p, int first, int count, size_t smem, cudaStream_t st) {
    cudaError_t e = configure<K, DEBUG, WILD>(smem);
    if (e != cudaSuccess) return e;
    const int threads = 32 * (warps + 1);            // + the producer warp
    if (getenv("WXB_DEBUG_OCC")) {
        int nbk = 0;
        cudaFuncAttributes fa;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbk, align_kernel<K, DEBUG, WILD>, threads, smem);
        cudaFuncGetAttributes(&fa, align_kernel<K, DEBUG, WILD>);
        fprintf(stderr, "[wxb] align_kernel<%d,%d,%d> W=%d NB=%d TE=%d: regs=%d smem=%zu grid=%d -> %d CTAs/SM\n", K, (int)DEBUG, (int)WILD,
                warps, nb, te, fa.numRegs, smem, count, nbk);
    }
    align_kernel<K, DEBUG, WILD><<<count, threads, smem, st>>>(p, first, nb, te);
    return cudaGetLastError();
}

template <int K>
static cudaError_t launch_one(int warps, int nb, int te, const Params &p, int first, int count, size_t smem, cudaStream_t st, bool debug, bool wild) {
    if (debug) return wild ? launch_variant<K, true, true>(warps, nb, te, p, first, count, smem, st)
                           : launch_variant<K, true, false>(warps, nb, te, p, first, count, smem, st);
    return wild ? launch_variant<K, false, true>(warps, nb, te, p, first, count, smem, st)
                : launch_variant<K, false, false>(warps, nb, te, p, first, count, smem, st);
}

template <int K>
static cudaError_t dispatch(int k, int warps, int nb, int te, const Params &p, int first, int count, size_t smem, cudaStream_t st, bool debug, bool wild) {
    if (k == K) return launch_one<K>(warps, nb, te, p, first, count, smem, st, debug, wild);
    if constexpr (K < KMAX) return dispatch<K + 1>(k, warps, nb, te, p, first, count, smem, st, debug, wild);
    else return cudaErrorInvalidValue;
}

cudaError_t launch_align(int k, int warps, int nb, int te, const Params &p, int first, int count, size_t smem, cudaStream_t st, bool debug, bool wild) {
    if (warps < 1 || warps > WMAX || nb < 2) return cudaErrorInvalidValue;
    return dispatch<1>(k, warps, nb, te, p, first, count, smem, st, debug, wild);
}

}  // namespace wxb

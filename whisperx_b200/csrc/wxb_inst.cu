// Instantiates align_kernel<K, DEBUG> for K in [WXB_KLO, WXB_KHI]; built once per range so the
// ranges compile in parallel (see whisperx_b200/build.py).
#include "wxb_kernel.cuh"

#include <cstdio>
#include <cstdlib>

#ifndef WXB_KLO
#error "compile with -DWXB_KLO=<lo> -DWXB_KHI=<hi>"
#endif

namespace wxb {

template <int K, bool DEBUG>
static cudaError_t configure(size_t smem) {
    // Every instantiation asks for the same (maximum) shared-memory carveout: kernels that want
    // different L1/shared splits cannot be co-resident on an SM, and the plan runs one kernel per
    // tokens-per-lane class concurrently (measured: without this the classes execute in groups).
    static size_t configured_for = static_cast<size_t>(-1);
    if (configured_for == smem) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(align_kernel<K, DEBUG>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(align_kernel<K, DEBUG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (getenv("WXB_DEBUG_OCC")) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, align_kernel<K, DEBUG>, 32, smem);
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, align_kernel<K, DEBUG>);
        fprintf(stderr, "[wxb] align_kernel<%d,%d>: regs=%d smem=%zu -> %d CTAs/SM\n", K, (int)DEBUG, fa.numRegs, smem, nb);
    }
    configured_for = smem;
    return cudaSuccess;
}

template <int K>
static cudaError_t launch_one(const Params &p, int first, int count, size_t smem, cudaStream_t st, bool debug) {
    cudaError_t e = debug ? configure<K, true>(smem) : configure<K, false>(smem);
    if (e != cudaSuccess) return e;
    if (debug) align_kernel<K, true><<<count, 32, smem, st>>>(p, first);
    else       align_kernel<K, false><<<count, 32, smem, st>>>(p, first);
    return cudaGetLastError();
}

template <int K, int KHI>
struct Dispatch {
    static cudaError_t go(int k, const Params &p, int first, int count, size_t smem, cudaStream_t st, bool debug) {
        if (k == K) return launch_one<K>(p, first, count, smem, st, debug);
        if constexpr (K < KHI) return Dispatch<K + 1, KHI>::go(k, p, first, count, smem, st, debug);
        else return cudaErrorInvalidValue;
    }
};

#define WXB_CAT2(a, b) a##b
#define WXB_CAT(a, b) WXB_CAT2(a, b)

cudaError_t WXB_CAT(launch_range_, WXB_KLO)(int k, const Params &p, int first, int count, size_t smem,
                                           cudaStream_t st, bool debug) {
    return Dispatch<WXB_KLO, WXB_KHI>::go(k, p, first, count, smem, st, debug);
}

}  // namespace wxb

// Instantiates align_kernel<K, DEBUG> for K in [WXB_KLO, WXB_KHI]; built once per range so the
// ranges compile in parallel (see whisperx_b200/build.py).
#include "wxb_kernel.cuh"

#ifndef WXB_KLO
#error "compile with -DWXB_KLO=<lo> -DWXB_KHI=<hi>"
#endif

namespace wxb {

template <int K>
static cudaError_t launch_one(const Params &p, int first, int count, size_t smem, cudaStream_t st, bool debug) {
    if (smem > 48 * 1024) {
        cudaError_t e = debug
            ? cudaFuncSetAttribute(align_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
            : cudaFuncSetAttribute(align_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
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

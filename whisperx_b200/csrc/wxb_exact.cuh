// Exact host/device helpers shared by the kernels and the C ABI (single source => same bits).
//
//   round3()        Python round(x, 3)  — reference use: whisperx/alignment.py:262-264, 308
//   round3_np()     np.float64.__round__(3)      — alignment.py:308 (and :262-264 for numpy-typed times)
//   pairwise_sum()  numpy add.reduce order behind pandas .mean() — alignment.py:308
//   row_log_softmax torch.log_softmax(dim=-1) formula — alignment.py:233
//
// Compiled with -fmad=false: the only fused multiply-add in this file is the explicit fma() in
// round3(), which is there *because* it is exact.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define WXB_HD __host__ __device__ __forceinline__
#else
#define WXB_HD inline
#endif

namespace wxb {

// Python round(x, 3) == strtod(printf("%.3f", x)): round the *exact* binary value of x to 3
// decimals, ties to even, then return the double nearest to that decimal.
//   p = fl(1000 x), e = 1000 x - p exactly (one fma)  =>  1000 x = p + e as real numbers.
//   n = floor(p), f = p - n is exact; f and 0.5 are both multiples of ulp(p) (p < 2^52), so
//   f != 0.5 decides the side on its own (|e| <= ulp(p)/2 < |f - 0.5|); at f == 0.5 the sign of
//   e decides, and e == 0 is a true decimal tie -> even n.
//   m / 1000.0 is one correctly rounded IEEE division = the double nearest to the decimal m/1000.
// Valid for |x| < 2^42 (timestamps and scores are far below that); non-finite x passes through.
WXB_HD double round3(double x) {
    if (!(x == x) || x == INFINITY || x == -INFINITY) return x;
    const double ax = fabs(x);
    if (ax >= 4398046511104.0) return x;       // 2^42: outside the contract above, passed through unchanged
    const double p = ax * 1000.0;
    const double e = fma(ax, 1000.0, -p);
    const double n = floor(p);
    const double f = p - n;
    bool up;
    if (f > 0.5) up = true;
    else if (f < 0.5) up = false;
    else if (e > 0.0) up = true;
    else if (e < 0.0) up = false;
    else up = (fmod(n, 2.0) != 0.0);
    const double m = up ? n + 1.0 : n;
    const double r = m / 1000.0;
    return x < 0.0 ? -r : r;
}

// round(x, 3) when x is a numpy float64 scalar (np.float64.__round__ == np.round): scale, rint
// (half to even on the *scaled binary* value), unscale.  Differs from round3() in ~2% of word
// means; the reference's word score is round(Series.mean(), 3) on a np.float64 (alignment.py:308).
WXB_HD double round3_np(double x) {
    return rint(x * 1000.0) / 1000.0;
}

// numpy pairwise_sum_DOUBLE (numpy/_core/src/umath/loops_utils.h.src) over a[idx(i)], i in [0,n).
// `get(i)` returns element i.  Recursion depth is log2(n/128); words are short, so the n > 128
// branch is only there for completeness.
template <class Get>
WXB_HD double pairwise_block(Get get, long lo, long n) {
    if (n < 8) {
        double r = -0.0;
        for (long i = 0; i < n; ++i) r += get(lo + i);
        return r;
    }
    double r0 = get(lo + 0), r1 = get(lo + 1), r2 = get(lo + 2), r3 = get(lo + 3);
    double r4 = get(lo + 4), r5 = get(lo + 5), r6 = get(lo + 6), r7 = get(lo + 7);
    long i;
    for (i = 8; i < n - (n % 8); i += 8) {
        r0 += get(lo + i + 0); r1 += get(lo + i + 1); r2 += get(lo + i + 2); r3 += get(lo + i + 3);
        r4 += get(lo + i + 4); r5 += get(lo + i + 5); r6 += get(lo + i + 6); r7 += get(lo + i + 7);
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += get(lo + i);
    return res;
}

template <class Get>
WXB_HD double pairwise_sum(Get get, long n) {
    // iterative form of the n > 128 halving recursion: numpy splits n into n2 = (n/2) - (n/2)%8
    // and n - n2, recursively.  An explicit stack keeps this usable on the device.
    if (n <= 128) return 0.0 + pairwise_block(get, 0, n);
    struct Frame { long lo, n; int state; double left; };
    Frame st[48];
    int sp = 0;
    st[sp++] = {0, n, 0, 0.0};
    double ret = 0.0;
    while (sp > 0) {
        Frame &f = st[sp - 1];
        if (f.n <= 128) { ret = pairwise_block(get, f.lo, f.n); --sp; continue; }
        long n2 = f.n / 2; n2 -= n2 % 8;
        if (f.state == 0) { f.state = 1; st[sp++] = {f.lo, n2, 0, 0.0}; }
        else if (f.state == 1) { f.left = ret; f.state = 2; st[sp++] = {f.lo + n2, f.n - n2, 0, 0.0}; }
        else { ret = f.left + ret; --sp; }
    }
    return 0.0 + ret;
}

// log_softmax of one row, sequential over the vocabulary in ascending index order:
//   m = max_v x[v];  s = sum_v exp(x[v] - m);  out[v] = (x[v] - m) - logf(s)
// This order *defines* the engine's log-probs (torch's CPU kernel cannot be matched bit for bit,
// SURVEY §7.2 item 1); every kernel that needs a log-prob from logits goes through here.
// exp(d), d <= 0, is one multiply and one MUFU (ex2.approx, relative error <= 2^-22 on the scaled
// argument): the terms that matter for the sum have small |d|, where this is as accurate as expf
// at a fifth of its issue slots; the largest term is exactly 1.
__device__ __forceinline__ float exp_nonpos(float d) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d * 1.4426950408889634f));
    return y;
}
// (the per-symbol loops are left to the compiler's own unrolling: explicit factors 1 / 2 / 4 / 8 measured 17 / 10 / 4 /
// 13 % slower on the benchmark step than its default — profiles/r2_ablation.md)
template <class Load>
__device__ __forceinline__ void row_lse(Load x, int V, float &m, float &ls) {
    float mx = x(0);
    for (int v = 1; v < V; ++v) mx = fmaxf(mx, x(v));
    float s = 0.0f;
    for (int v = 0; v < V; ++v) s += exp_nonpos(x(v) - mx);
    m = mx;
    ls = logf(s);
}
// Rows wider than WIDE_V symbols are reduced by G lanes together (lane l takes symbols l, l+G, ...):
// no lane adds thousands of terms into one running sum (the plain serial sum drops every term below
// half an ulp of it: measured bias 8.5e-6 at V = 7,001).  Lane sums are compensated (Kahan) and
// combined by a fixed xor tree, so the result is deterministic.  All G lanes return the same m and ls.
// Every lane of the warp must call it (full-mask shuffles); G = 32 is one row per warp.
constexpr int WIDE_V = 64;
template <int G>
__device__ __forceinline__ void row_lse_group(const float *row, int V, int lane_in, float &m, float &ls) {
    float mx = -INFINITY;
    for (int v = lane_in; v < V; v += G) mx = fmaxf(mx, row[v]);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.0f, c = 0.0f;
    for (int v = lane_in; v < V; v += G) {
        const float y = exp_nonpos(row[v] - mx) - c;
        const float t = s + y;
        c = (t - s) - y;
        s = t;
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    m = mx;
    ls = logf(s);
}
__device__ __forceinline__ void row_lse_warp(const float *row, int V, int lane, float &m, float &ls) {
    row_lse_group<32>(row, V, lane, m, ls);
}
__device__ __forceinline__ float logp_from(float x, float m, float ls) { return (x - m) - ls; }

}  // namespace wxb

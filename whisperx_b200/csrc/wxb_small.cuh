// The small kernels around align_kernel: records for un-alignable segments, backtrack over the choice bits,
// per-token finalisation (merge_repeats + timestamps), per-word reduction, stand-alone log-softmax,
// backtrack from a caller's trellis, speaker assignment, metadata upload.  Included by wxb_align.cu only.
#pragma once
#include "../../include/wxb_align.h"
#include "wxb_kernel.cuh"

namespace wxb {

// ------------------------------------------------------------------ small kernels
struct ShortSeg { long long tok_base; long long frame_base; int N; int T; int seg; int pad; };

// Segments the planner proved un-alignable (N == 0 or N > T: column N of the trellis is -inf in
// every row, so argmax = 0 and the walk is empty — alignment.py:396-420) get the FAILED record.
__global__ void mark_failed_kernel(const ShortSeg *list, int n, int *status, int *nframes, int *c_start,
                                   int *c_end, double *c_score, double *c_start_s, double *c_end_s,
                                   double *c_score_r, int *path_tok, float *path_logp) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const ShortSeg s = list[i];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (threadIdx.x == 0) { status[s.seg] = WXB_SEG_FAILED; nframes[s.seg] = 0; }
    for (int k = threadIdx.x; k < s.N; k += blockDim.x) {
        const long long o = s.tok_base + k;
        c_start[o] = -1; c_end[o] = -1; c_score[o] = nan;
        if (c_start_s) { c_start_s[o] = nan; c_end_s[o] = nan; c_score_r[o] = nan; }
    }
    if (path_tok)
        for (int f = threadIdx.x; f < s.T; f += blockDim.x) { path_tok[s.frame_base + f] = -1; path_logp[s.frame_base + f] = 0.0f; }
}

// Plan metadata upload by the SMs instead of the copy engine: wxb_align_host has the host->device
// engine busy with the emission stream for milliseconds, and a DMA of the (small) metadata would
// queue behind it.  `src` is pinned, mapped host memory; 16-byte loads over PCIe.
__global__ void __launch_bounds__(256) upload_kernel(uint4 *dst, const uint4 *src, size_t n16) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n16; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dst[i] = src[i];
}

// ---- backtrack (:395-421) over the choice bits: one warp per segment, control flow warp-uniform ----
// The walk needs one choice word per visited (column, 32-frame word) pair: ~N + T/32 *dependent*
// loads if done naively.  Instead the warp fetches a window of 32 columns x 4 words (128 frames)
// with 4 independent loads per lane and walks inside it on register shuffles; a new window is
// fetched only when the path leaves the block: about max(N/32, T/96) dependent round trips per
// segment.
__global__ void __launch_bounds__(128) backtrack_kernel(const Params p, const int seg_first, const int count) {
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= count) return;
    const SegDesc db = p.segs[seg_first + idx];
    const int K = db.K, W = db.W, KL = db.KL;
    const int t_start = p.nframes[db.seg];
    const long long wstep = static_cast<long long>(W) * K * 32;
    auto owner = [&](int k, int &g, int &slot) { column_owner(k, db.N, W, K, KL, g, slot); };     // token k -> global lane / register slot
    int status = 0;
    if (t_start == 0) {
        status = 1;                                  // empty walk -> for/else -> None
    } else {
        int j = db.N;                                // current column (token j-1 is being located)
        int s = t_start - 1;                         // frame of the current point
        int *cs = p.c_start + db.tok_base;           // token k's entry frame goes to cs[k]
        bool more = true;
        while (more) {
            const int j0 = j;                        // window columns j0, j0-1, ..., j0-31
            const int w0 = s >> 5;                   // window words   w0, w0-1, w0-2, w0-3
            uint32_t wA = 0u, wB = 0u, wC = 0u, wD = 0u;
            const int c = j0 - lane;
            if (c >= 1) {
                int g, slot;
                owner(c - 1, g, slot);
                const uint32_t *bp = p.bits + db.bits_off + static_cast<long long>(w0) * wstep + ((g >> 5) * K + slot) * 32 + (g & 31);
                wA = __ldcg(bp);
                if (w0 >= 1) wB = __ldcg(bp - 1 * wstep);
                if (w0 >= 2) wC = __ldcg(bp - 2 * wstep);
                if (w0 >= 3) wD = __ldcg(bp - 3 * wstep);
            }
            const int jmin = max(j0 - 31, 1);        // leave the window below this column ...
            const int smin = max((w0 - 2) << 5, 0);  // ... or below this frame (the look-back needs word ws and ws-1)
            int my_entry = -1;                       // entry frame of this lane's column, once the walk has passed it
            // One iteration per token (or per 32 frames the path stayed): the choice bits of the current column for
            // frames s, s-1, ..., s-31 are cut out of two neighbouring words with one funnel shift — bit q = frame
            // s-q, whatever the word alignment of s — so the lowest set bit is the frame that entered the token.
            // Entry frames stay in the lane that owns the column and leave as one coalesced store per window.
            do {
                const int dw = w0 - (s >> 5);        // 0..2 inside the window
                const bool d1 = (dw == 1), d2 = (dw == 2);
                const uint32_t lo = d2 ? wC : (d1 ? wB : wA);      // word holding frame s
                const uint32_t hi = d2 ? wD : (d1 ? wC : wB);      // the 32 frames before it
                const uint32_t mine = __funnelshift_r(lo, hi, 31 - (s & 31));    // bit (31 - f%32) of a word is frame f
                const int src_lane = j0 - j;
                const uint32_t m = __shfl_sync(FULL, mine, src_lane);
                const bool found = (m != 0u);
                const int sc = s - (__ffs(m) - 1);   // frame at which token j-1 was entered (if found)
                if (found && lane == src_lane) my_entry = sc;
                j -= found ? 1 : 0;
                s = found ? sc - 1 : s - 32;
            } while (j >= jmin && s >= smin);
            if (my_entry >= 0) cs[c - 1] = my_entry;
            more = (j > 0) && (s >= 0);
        }
        status = (j == 0) ? 0 : 1;                   // ran out of frames before reaching token 0 -> None
    }
    if (lane == 0) {
        p.status[db.seg] = status;
        p.nframes[db.seg] = status ? 0 : t_start;
        if (p.trace) p.trace[4ll * (seg_first + idx) + 2] = global_ns();
    }
    // merge_repeats, scores and timestamps run token-parallel in finalize_kernel
}

// ---- phase 3 as its own kernel: one thread per token of every segment ----
// merge_repeats (:438-454): token k spans frames [start_k, start_{k+1}) (the last one ends at
// t_start); its score is the mean of the path probabilities: exp(logp[start, token]) for the frame
// that entered it and exp(logp[f, 0]) — literal column 0 (:409) — for the frames that stayed.
// Char timestamps (:252-264): round(frame*ratio + t1, 3), round(score, 3).
__global__ void __launch_bounds__(128) finalize_kernel(const Params p, const FinBlock *blocks) {
    const FinBlock fb = blocks[blockIdx.x];
    const SegDesc d = p.segs[fb.desc];
    const int N = d.N, T = d.T, V = p.V;
    const int status = p.status[d.seg];
    const int t_start = p.nframes[d.seg];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const bool want_sec = (p.seg_t1 != nullptr) && (p.c_start_s != nullptr);
    if (p.path_tok && fb.k0 == 0) {
        // frames off the path: everything on failure, else [0, start_0) and [t_start, T)
        const int first = status ? T : p.c_start[d.tok_base];
        for (int f = threadIdx.x; f < T; f += blockDim.x)
            if (status || f < first || f >= t_start) { p.path_tok[d.frame_base + f] = -1; p.path_logp[d.frame_base + f] = 0.0f; }
    }
    const int k = fb.k0 + threadIdx.x;
    if (k >= N) return;
    const long long o = d.tok_base + k;
    if (status) {
        p.c_start[o] = -1;
        p.c_end[o] = -1;
        p.c_score[o] = nan;
        if (want_sec) { p.c_start_s[o] = nan; p.c_end_s[o] = nan; p.c_score_r[o] = nan; }
        return;
    }
    const float *Eg = p.emis + d.emis_off;
    const int sf = p.c_start[o];
    const int ef = (k == N - 1) ? t_start : p.c_start[o + 1];
    const int t4 = __ldg(p.tok4 + o);
    float x;
    if (t4 < 0) {
        x = p.wild[d.frame_base + sf];                    // wildcard token: max over the non-blank vocabulary (a log-prob already)
    } else {
        x = __ldg(Eg + static_cast<size_t>(sf) * V + (t4 >> 2));
        if (p.is_logits) x = logp_from(x, p.rowmax[d.frame_base + sf], p.rowlse[d.frame_base + sf]);
    }
    double sum = static_cast<double>(expf(x));           // Python sum(): 0 + p_first + ...
    for (int f = sf + 1; f < ef; ++f) sum += static_cast<double>(p.p0[d.frame_base + f]);
    const double score = sum / static_cast<double>(ef - sf);
    p.c_end[o] = ef;
    p.c_score[o] = score;
    if (want_sec) {
        const double t1 = p.seg_t1[d.seg], ratio = p.seg_ratio[d.seg];
        const double ts = __dadd_rn(__dmul_rn(static_cast<double>(sf), ratio), t1);
        const double te = __dadd_rn(__dmul_rn(static_cast<double>(ef), ratio), t1);
        // numpy-typed segment times make `frame*ratio + t1` a np.float64, whose round() is numpy's
        const bool np_times = p.seg_round_np != nullptr && p.seg_round_np[d.seg] != 0;
        p.c_start_s[o] = np_times ? round3_np(ts) : round3(ts);
        p.c_end_s[o] = np_times ? round3_np(te) : round3(te);
        p.c_score_r[o] = round3(score);                 // Segment.score is always a Python float
    }
    if (p.path_tok) {
        p.path_tok[d.frame_base + sf] = k;
        p.path_logp[d.frame_base + sf] = x;
        for (int f = sf + 1; f < ef; ++f) {
            float e0 = __ldg(Eg + static_cast<size_t>(f) * V);
            if (p.is_logits) e0 = logp_from(e0, p.rowmax[d.frame_base + f], p.rowlse[d.frame_base + f]);
            p.path_tok[d.frame_base + f] = k;
            p.path_logp[d.frame_base + f] = e0;
        }
    }
}

// Word reduction, alignment.py:297-320: start = min, end = max over the word's aligned chars
// (first / last aligned char: rounded times are monotone in the frame index), score =
// round(mean(rounded char scores), 3) with pandas' zero-filled numpy pairwise sum / count; the
// mean is a np.float64, so that round() is numpy's (rint(x*1000)/1000), not Python's.
__global__ void word_reduce_kernel(long long W, const int *word_seg, const long long *seg_tok_base,
                                   const long long *wchar_off, const int *wchar_idx, const int *status,
                                   const double *c_start_s, const double *c_end_s, const double *c_score_r,
                                   double *w_start, double *w_end, double *w_score) {
    // (the caller shifts word_seg / wchar_off / w_* to the first word of the range it wants reduced)
    const long long w = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (w >= W) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const int seg = word_seg[w];
    const long long base = seg_tok_base[seg];
    const long long lo = wchar_off[w];
    const long n = static_cast<long>(wchar_off[w + 1] - lo);
    int first = -1, last = -1;
    long count = 0;
    if (status[seg] == WXB_SEG_OK) {
        for (long i = 0; i < n; ++i) {
            const int idx = wchar_idx[lo + i];
            if (idx >= 0) { if (first < 0) first = idx; last = idx; ++count; }
        }
    }
    if (count == 0) { w_start[w] = nan; w_end[w] = nan; w_score[w] = nan; return; }
    const double sum = pairwise_sum(
        [&](long i) { const int idx = wchar_idx[lo + i]; return idx >= 0 ? c_score_r[base + idx] : 0.0; }, n);
    w_start[w] = c_start_s[base + first];
    w_end[w] = c_end_s[base + last];
    w_score[w] = round3_np(sum / static_cast<double>(count));   // np.float64.__round__, not Python round
}

// Kernel (i) on its own — torch.log_softmax(emissions, dim=-1), alignment.py:233.  HBM-bound: 8*V bytes
// per row.  Three shapes (launch_log_softmax picks):
//
// log_softmax_kernel, group = 1 (V <= 64): a CTA stages `rows_per_cta` contiguous rows in shared memory
//   with coalesced (128-bit when aligned) loads, one thread per row normalises in the same
//   row_lse()/logp_from() order as the fused staging, and the tile goes back coalesced.  Rows are staged
//   with an odd stride (pad = 1 for even V) so that the 32 rows a warp works on start in 32 different
//   banks — with the plain stride an even V serialises every access (V = 32: 32-way; measured 0.56 vs
//   5.8 TB/s at V = 29).  The flat index -> staged slot map is i + i / V, the division one multiply-high.
// log_softmax_rows_kernel<ITER, WPR> (64 < V <= 8192): a warp (V <= 1024) or a whole CTA per row straight
//   from global memory, the row held in ITER registers per lane — no staging, ITER independent loads in
//   flight per lane (the staged shape keeps too few bytes in flight once a CTA's tile is tens of KB:
//   1.7-2.8 TB/s at V = 100-200 against 5.3-6.9 this way).  Lane sums are compensated, then a fixed tree.
// log_softmax_kernel, group = 32 (wider still): staged, a warp per row (row_lse_warp).
__global__ void __launch_bounds__(128) log_softmax_kernel(const float *in, float *out, long long rows, int V,
                                                          int rows_per_cta, int group, int pad, unsigned magic) {
    extern __shared__ __align__(16) float tile[];
    const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
    const int nr = static_cast<int>(min(static_cast<long long>(rows_per_cta), rows - row0));
    if (nr <= 0) return;
    const long long base = row0 * V;
    const int nel = nr * V;
    const bool vec = ((reinterpret_cast<uintptr_t>(in + base) | reinterpret_cast<uintptr_t>(out + base)) & 15) == 0;
    const int stride = V + pad;
    auto slot = [&](int i) { return pad ? i + static_cast<int>(__umulhi(static_cast<unsigned>(i), magic)) * pad : i; };
    if (vec) {
        const float4 *src = reinterpret_cast<const float4 *>(in + base);
        if (pad) {
            for (int i = threadIdx.x; i < nel / 4; i += blockDim.x) {
                const float4 v = __ldg(src + i);
                tile[slot(4 * i)] = v.x; tile[slot(4 * i + 1)] = v.y; tile[slot(4 * i + 2)] = v.z; tile[slot(4 * i + 3)] = v.w;
            }
        } else {
            float4 *t4 = reinterpret_cast<float4 *>(tile);
            for (int i = threadIdx.x; i < nel / 4; i += blockDim.x) t4[i] = __ldg(src + i);
        }
        for (int i = (nel & ~3) + threadIdx.x; i < nel; i += blockDim.x) tile[slot(i)] = __ldg(in + base + i);
    } else {
        for (int i = threadIdx.x; i < nel; i += blockDim.x) tile[slot(i)] = __ldg(in + base + i);
    }
    __syncthreads();
    if (group == 1) {
        for (int rr = threadIdx.x; rr < nr; rr += blockDim.x) {
            float *row = tile + rr * stride;
            float m, ls;
            row_lse([&](int v) { return row[v]; }, V, m, ls);
            for (int v = 0; v < V; ++v) row[v] = logp_from(row[v], m, ls);
        }
    } else {
        const int lane = threadIdx.x & 31;
        for (int rr = threadIdx.x >> 5; rr < nr; rr += blockDim.x >> 5) {
            float *row = tile + rr * stride;
            float m, ls;
            row_lse_warp(row, V, lane, m, ls);
            for (int v = lane; v < V; v += 32) row[v] = logp_from(row[v], m, ls);
        }
    }
    __syncthreads();
    if (vec) {
        float4 *dst = reinterpret_cast<float4 *>(out + base);
        if (pad) {
            for (int i = threadIdx.x; i < nel / 4; i += blockDim.x)
                dst[i] = make_float4(tile[slot(4 * i)], tile[slot(4 * i + 1)], tile[slot(4 * i + 2)], tile[slot(4 * i + 3)]);
        } else {
            const float4 *t4 = reinterpret_cast<const float4 *>(tile);
            for (int i = threadIdx.x; i < nel / 4; i += blockDim.x) dst[i] = t4[i];
        }
        for (int i = (nel & ~3) + threadIdx.x; i < nel; i += blockDim.x) out[base + i] = tile[slot(i)];
    } else {
        for (int i = threadIdx.x; i < nel; i += blockDim.x) out[base + i] = tile[slot(i)];
    }
}

template <int ITER, int WPR>      // WPR warps per row: 1 (a row per warp, 8 rows per CTA) or 8 (a row per CTA)
__global__ void __launch_bounds__(256) log_softmax_rows_kernel(const float *in, float *out, long long rows, int V) {
    static_assert(WPR == 1 || WPR == 8, "one warp or the whole CTA per row");
    __shared__ float red[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int RPC = 8 / WPR, CS = 32 * WPR;
    const int col = (warp % WPR) * 32 + lane;
    const long long step = static_cast<long long>(gridDim.x) * RPC;
    for (long long row = static_cast<long long>(blockIdx.x) * RPC + warp / WPR; row < rows; row += step) {
        const float *src = in + row * V;
        float x[ITER];
#pragma unroll
        for (int k = 0; k < ITER; ++k) x[k] = (col + CS * k < V) ? __ldg(src + col + CS * k) : -INFINITY;
        float mx = x[0];
#pragma unroll
        for (int k = 1; k < ITER; ++k) mx = fmaxf(mx, x[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if constexpr (WPR == 8) {
            if (lane == 0) red[0][warp] = mx;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < 8; ++w) mx = fmaxf(mx, red[0][w]);
        }
        float s = 0.0f, c = 0.0f;           // WPR == 1: the order of row_lse_group<32>
#pragma unroll
        for (int k = 0; k < ITER; ++k) {
            if (col + CS * k < V) {
                const float y = exp_nonpos(x[k] - mx) - c;
                const float t = s + y;
                c = (t - s) - y;
                s = t;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if constexpr (WPR == 8) {
            if (lane == 0) red[1][warp] = s;
            __syncthreads();
            s = red[1][0];
#pragma unroll
            for (int w = 1; w < 8; ++w) s += red[1][w];
        }
        const float ls = logf(s);
        float *dst = out + row * V;
#pragma unroll
        for (int k = 0; k < ITER; ++k)
            if (col + CS * k < V) dst[col + CS * k] = logp_from(x[k], mx, ls);
        if constexpr (WPR == 8) __syncthreads();      // red[] is reused by the next row
    }
}

// backtrack() from a caller-supplied trellis — the literal algorithm of alignment.py:387-421
// (used by the functional drop-in `backtrack(trellis, emission, tokens, blank_id)`).
__global__ void __launch_bounds__(256) backtrack_trellis_kernel(const float *R, const float *E, const int *tok,
                                                                int T, int N, int V, int blank, int *path_tok,
                                                                float *path_logp, int *status, int *nframes) {
    __shared__ float s_val[256];
    __shared__ int s_idx[256];
    const int W = N + 1;
    for (int t = threadIdx.x; t < T; t += blockDim.x) { path_tok[t] = -1; path_logp[t] = 0.0f; }
    // torch.argmax(trellis[:, N]): first maximal index, NaN counts as maximal
    auto better = [](float av, int at, float bv, int bt) {
        const bool an = av != av, bn = bv != bv;
        if (an != bn) return an;
        if (an) return at < bt;
        return av > bv || (av == bv && at < bt);
    };
    float bv = R[N];
    int bt = 0;
    for (int t = threadIdx.x; t <= T; t += blockDim.x) {
        const float v = R[static_cast<size_t>(t) * W + N];
        if (better(v, t, bv, bt)) { bv = v; bt = t; }
    }
    s_val[threadIdx.x] = bv; s_idx[threadIdx.x] = bt;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off && better(s_val[threadIdx.x + off], s_idx[threadIdx.x + off], s_val[threadIdx.x], s_idx[threadIdx.x])) {
            s_val[threadIdx.x] = s_val[threadIdx.x + off];
            s_idx[threadIdx.x] = s_idx[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const int t_start = s_idx[0];
    int j = N;
    bool reached = false;
    for (int t = t_start; t > 0; --t) {
        const float *e = E + static_cast<size_t>(t - 1) * V;
        const float stayed = R[static_cast<size_t>(t - 1) * W + j] + e[blank];
        const float changed = R[static_cast<size_t>(t - 1) * W + j - 1] + e[tok[j - 1]];
        const bool took = changed > stayed;
        path_tok[t - 1] = j - 1;
        path_logp[t - 1] = e[took ? tok[j - 1] : 0];
        if (took) { --j; if (j == 0) { reached = true; break; } }
    }
    *status = reached ? WXB_SEG_OK : WXB_SEG_FAILED;
    *nframes = t_start;
}

// assign_word_speakers (whisperx/diarize.py:35-67): one thread per query interval, diarization turns
// staged through shared memory in row order so that every per-speaker sum is accumulated exactly as
// pandas' groupby-sum does (Kahan compensation, row order).
constexpr int SPK_MAX = 64;
constexpr int SPK_TILE = 256;
// One launch handles speakers [spk_base, spk_base + n_spk), n_spk <= SPK_MAX.  More speakers than that are
// covered by successive launches in ascending id order that carry the best (sum, id) so far in
// best_sum / out: a later speaker replaces it only when strictly larger, which keeps the reference's
// "ties to the first label".  Ids outside the launch's range are ignored (also protects the local arrays).
__global__ void __launch_bounds__(128) assign_speakers_kernel(const double *d_start, const double *d_end, const int *d_spk,
                                                              long long n_diar, int spk_base, int n_spk, const double *q_start,
                                                              const double *q_end, long long n_query, int fill_nearest,
                                                              int *out, double *best_sum) {
    __shared__ double s_start[SPK_TILE], s_end[SPK_TILE];
    __shared__ int s_spk[SPK_TILE];
    const long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const bool live = q < n_query;
    const double qs = live ? q_start[q] : 0.0, qe = live ? q_end[q] : 0.0;
    double sum[SPK_MAX], comp[SPK_MAX];
    for (int k = 0; k < n_spk; ++k) { sum[k] = 0.0; comp[k] = 0.0; }
    unsigned long long seen = 0ull;              // speakers with at least one qualifying row
    for (long long base = 0; base < n_diar; base += SPK_TILE) {
        const int n = static_cast<int>(min(static_cast<long long>(SPK_TILE), n_diar - base));
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            s_start[i] = d_start[base + i]; s_end[i] = d_end[base + i]; s_spk[i] = d_spk[base + i];
        }
        __syncthreads();
        if (!live) continue;
        for (int i = 0; i < n; ++i) {
            const double inter = fmin(s_end[i], qe) - fmax(s_start[i], qs);
            const int k = s_spk[i] - spk_base;
            if ((fill_nearest || inter > 0.0) && k >= 0 && k < n_spk) {
                const double y = inter - comp[k];
                const double t = sum[k] + y;
                comp[k] = (t - sum[k]) - y;
                sum[k] = t;
                seen |= 1ull << k;
            }
        }
    }
    if (!live) return;
    int best = -1;
    for (int k = 0; k < n_spk; ++k)
        if ((seen >> k) & 1ull)
            if (best < 0 || sum[k] > sum[best]) best = k;      // descending sort, ties keep label order
    if (!best_sum) { out[q] = best < 0 ? -1 : spk_base + best; return; }
    const int prev = spk_base > 0 ? out[q] : -1;                // earlier launches' winner
    if (best >= 0 && (prev < 0 || sum[best] > best_sum[q])) { out[q] = spk_base + best; best_sum[q] = sum[best]; }
    else if (spk_base == 0) out[q] = -1;
}

}  // namespace wxb

// Fused forced-alignment kernel for sm_100a: one CTA per transcript segment, the token axis of
// the trellis split over the CTA's warps and swept as a wavefront.
//
// Replaces, for a whole batch of segments, the reference's per-segment Python:
//   log_softmax            whisperx/alignment.py:233      (producer warp, fused into the tile staging)
//   get_trellis            whisperx/alignment.py:359-379  (compute warps: recurrence, token axis in registers)
//   backtrack              whisperx/alignment.py:387-421  (backtrack_kernel in wxb_small.cuh: walk over packed choice bits)
//   merge_repeats          whisperx/alignment.py:438-454  (finalize_kernel: per-token start/end/mean prob)
//   char timestamps        whisperx/alignment.py:252-264  (finalize_kernel: exact decimal rounding)
//
// CTA = W compute warps + 1 producer warp.  Every segment of a batch runs with the batch's K (one kernel
// instantiation per batch: the instruction cache decides between a fast and a slow step here), W = ceil(N / 32K).
// Warps 0 .. W-2 are "full": every lane owns K consecutive token columns.  The last warp holds the
// m = N - 32K(W-1) columns that remain (1 <= m <= 32K) with KL = ceil(m / 32) <= K columns per lane: its first
// rL = m - 32(KL-1) lanes own KL columns, the others KL-1, so the last column of the transcript is the last
// column of its last occupied lane.  The recurrence loop is instantiated per column count: a transcript of 399
// tokens runs as 192 + 192 + 32 column slots (K = 6, KL = 1) instead of 3 x 192.
//
//   producer warp   streams the segment's emission rows HBM -> shared memory exactly once, TE frames
//                   per tile through a ring of NB buffers (one cp.async.bulk per tile, mbarrier
//                   completion), applies log_softmax in place (one row per lane), and builds trellis
//                   column 0 for the tile (float64 running sum of literal vocabulary column 0).
//   compute warp w  runs the recurrence over its 32 lanes x K columns for the tile; the column to the
//                   left of lane 0 comes from a per-tile boundary array in shared memory — column 0
//                   for warp 0, the last column of warp w-1 otherwise — and lane 31 writes its own
//                   last column into the boundary array of warp w+1.  Warp w therefore trails warp
//                   w-1 by one tile: a systolic wavefront; handshakes are mbarriers, once per tile.
//
// Nothing of size O(T*V) or O(T*N*4) is ever written: the trellis lives in registers, and what the
// backtrack needs is a 1-bit-per-cell stay/change matrix (T*N/8 bytes, L2 resident) plus the running
// first-argmax of the last column.
//
// Exactness: float32 add / NaN-propagating max in the reference's order, no
// FMA (-fmad=false; the recurrence has no multiply anyway), column 0 = float(double cumsum) of
// literal vocabulary column 0, +inf tail in column 0, strict `changed > stayed` (tie -> stay),
// first-argmax over the last column, literal column 0 for the probability of a "stay" frame.
#pragma once
#include "wxb_exact.cuh"

namespace wxb {

struct SegDesc {
    long long emis_off;     // float offset of the segment's first emission row
    long long frame_base;   // frame_off[s]: index base for per-frame scratch and path outputs
    long long tok_base;     // tok_off[s]:   index base for tokens and per-char outputs
    long long bits_off;     // 32-bit word offset into the choice-bit scratch
    long long trellis_off;  // float offset into the debug trellis (DEBUG only)
    int T, N, seg, K;       // K: columns per lane of the full warps
    int W, pad;             // compute warps of the segment's CTA; pad: emission chunk (host path)
    int KL, pad2;           // columns per lane of the last warp (its first lanes; the others hold KL-1)
};

struct Params {
    const float *emis;
    const int *tok4;        // [sum_N] token ids * 4 (byte offset inside an emission row)
    const SegDesc *segs;
    uint32_t *bits;         // choice words of a segment: word ((t/32)*W*K + w*K + slot)*32 + lane, bit 31-(t%32)
    float *p0;              // [sum_T] expf(logp[t, 0])
    float *wild;            // [sum_T] (wildcard batches) max over non-blank v of logp[t, v]
    float *rowmax;          // [sum_T] (logits mode) row max
    float *rowlse;          // [sum_T] (logits mode) log(sum(exp(x - max)))
    int V, blank, is_logits, pad0;
    int *status, *nframes;
    int *c_start, *c_end;
    double *c_score, *c_start_s, *c_end_s, *c_score_r;
    const double *seg_t1, *seg_ratio;
    const unsigned char *seg_round_np;   // [S] or null: 1 = numpy rounding for this segment's timestamps
    int *path_tok;
    float *path_logp, *final_col, *trellis, *logprobs;
    unsigned long long *trace;   // debug: [4 * n_desc] {t_begin, t_forward_done, t_end, smid} per CTA, or null
    unsigned long long *trace_hist;   // debug: [4][2] frames handled per scheduler by compute / producer warps, or null
};

constexpr int UNR = 4;           // frames per unrolled block of the recurrence: boundary values move as float4
constexpr uint32_t FULL = 0xffffffffu;
constexpr int KMAX = 8;          // largest tokens-per-lane class that is instantiated
constexpr int WMAX = 31;         // compute warps per CTA (+1 producer warp = 1024 threads)

// Shared-memory carve-up of one CTA (all offsets in bytes from the dynamic smem base).
struct SmemLayout {
    uint32_t tile_bytes, off_left, off_eb, off_wild, off_xd, off_fin, off_bar, total;
};
__host__ __device__ inline SmemLayout smem_layout(int W, int NB, int TE, int V, bool wild = false) {
    SmemLayout L;
    const uint32_t per_ring = static_cast<uint32_t>(NB) * TE * 4u;        // one float per frame of every ring slot
    L.tile_bytes = static_cast<uint32_t>(TE) * V * 4u;                    // TE % 4 == 0 -> multiple of 16
    L.off_left = static_cast<uint32_t>(NB) * L.tile_bytes;                // left[W+1][NB][TE]: column left of each warp's first
    L.off_eb = L.off_left + static_cast<uint32_t>(W + 1) * per_ring;      // eb[NB][TE]: logp[t, blank]
    L.off_wild = L.off_eb + per_ring;                                     // wild[NB][TE]: wildcard emission (WILD kernels)
    L.off_xd = (L.off_wild + (wild ? per_ring : 0u) + 15u) & ~15u;        // 32 doubles
    L.off_fin = L.off_xd + 256u;                                          // trellis[T, N] (one float)
    L.off_bar = L.off_fin + 16u;                                          // tma[NB] full[NB] empty[NB] edge[W][NB]
    L.total = L.off_bar + static_cast<uint32_t>(3 + W) * NB * 8u;
    return L;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float a) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32x4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// torch.maximum: NaN-propagating max in one FMNMX.NAN
__device__ __forceinline__ float fmax_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Blocks until the barrier's phase `parity` has completed.  try_wait suspends the thread in hardware until the
// phase flips or a time limit expires; with the default (short, implementation-defined) limit the retry loop
// of the waiting warps — producers ahead of their ring, compute warps behind their neighbour — was 17.6 % of all
// instructions the kernel issued (ncu source view, profiles/r2_align_kernel_lines.md).  With an explicit limit of
// 10 ms (the value CUTLASS passes) a waiting warp costs no issue slots; it still wakes as soon as the phase flips.
#ifndef WXB_WAIT_HINT_NS
#define WXB_WAIT_HINT_NS 0x989680u      // experiments: -DWXB_WAIT_HINT_NS=1u restores (nearly) the default behaviour
#endif
// (one out-of-line copy of this loop instead of ~20 inlined ones would save 4 KB of hot code, but the calls cost
// 27 % of the step: measured, profiles/r2_ablation.md)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if WXB_WAIT_HINT_NS
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(WXB_WAIT_HINT_NS) : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar) : "memory");
}

// token k (0-based) -> owning global lane (32*warp + lane) and register slot, for a segment of N tokens laid out
// as described at the top of this file
__device__ __forceinline__ void column_owner(int k, int N, int W, int K, int KL, int &g, int &slot) {
    const int full = 32 * K * (W - 1);
    if (k < full) { g = k / K; slot = k - g * K; return; }
    const int kk = k - full;
    const int rL = (N - full) - 32 * (KL - 1);          // lanes of the last warp that own KL columns
    int l;
    if (KL == 1 || kk < rL * KL) { l = kk / KL; slot = kk - l * KL; }
    else { const int k2 = kk - rL * KL; const int q = k2 / (KL - 1); l = rL + q; slot = k2 - q * (KL - 1); }
    g = 32 * (W - 1) + l;
}

// Register budget per thread: the recurrence keeps 3K values live (state, choice accumulators, token
// offsets) plus ~2K transients; a CTA of up to 1024 threads needs <= 64.  Fewer registers = more
// resident warps to hide the dependent-issue latency of the recurrence.
#ifndef WXB_REGS
#define WXB_REGS(K) ((K) <= 4 ? 40 : (K) <= 6 ? 56 : 64)
#endif

// One compute warp: the recurrence over its 32 lanes x KC columns, tile after tile behind its left neighbour.
// KC = columns per lane of this warp; MIXED = the CTA's last warp (first r lanes hold KC columns, the rest KC-1,
// and its last column is column N of the trellis); KF = columns per lane of the CTA's full warps (fixes the
// warp's first column and the layout of the choice words).
template <int KC, bool MIXED, bool DEBUG, bool WILD>
__device__ __forceinline__ void compute_warp(const Params &p, const SegDesc &d, const SmemLayout &L, const uint32_t sm_base,
                                             const int NB, const int TE, const int ntiles, const int wi, const int lane,
                                             const int W, const int KF) {
    const int T = d.T, N = d.N, V = p.V;
    const uint32_t bar_tma = sm_base + L.off_bar;
    const uint32_t bar_full = bar_tma + 8u * NB;
    const uint32_t bar_empty = bar_full + 8u * NB;
    const uint32_t bar_edge = bar_empty + 8u * NB;     // [w][b], w = 0..W-1 (row 0 unused)
    const uint32_t left_stride_w = static_cast<uint32_t>(NB) * TE * 4u;   // bytes between left[w] and left[w+1]
    // columns of this warp: [col0w, col0w + 32 KC) for a full warp; the last (MIXED) warp holds the m columns
    // that remain, KC in its first r lanes and KC-1 in the others
    const int col0w = wi * 32 * KF;
    const int m = N - col0w;
    const int r = MIXED ? m - 32 * (KC - 1) : 32;   // 1..32
    const bool wide = !MIXED || lane < r;
    const int cb = col0w + (wide ? lane * KC : r * KC + (lane - r) * (KC - 1));
    const int ncols = wide ? KC : KC - 1;
    const bool last_warp = MIXED;
    // lane of this warp whose last real column is the warp's last column (KC = 1: the last warp may
    // end before lane 31); in the last warp that column is column N of the trellis
    const int laneN = (KC == 1 && MIXED) ? (m - 1) : 31;

    uint32_t tokoff[KC];
    uint32_t wmask = 0u;            // (WILD) bit i: column i is a wildcard token
#pragma unroll
    for (int i = 0; i < KC; ++i) {
        // byte offsets (token*4) come pre-multiplied from the plan: LDS [R + UR] with a uniform row base
        const int t4 = (i < ncols) ? __ldg(p.tok4 + d.tok_base + cb + i) : 0;
        if (WILD && t4 < 0) wmask |= 1u << i;
        tokoff[i] = (t4 < 0) ? 0u : static_cast<uint32_t>(t4);
    }
    float R[KC];
    uint32_t acc[KC];
#pragma unroll
    for (int i = 0; i < KC; ++i) { R[i] = -INFINITY; acc[i] = 0u; }   // trellis[0, 1:] = -inf  (:369)
    float src = -INFINITY;          // this lane's last real column at the current row
    const uint32_t rowstep = static_cast<uint32_t>(V) * 4u;
    const uint32_t left_in0 = sm_base + L.off_left + wi * left_stride_w;      // left[wi][b][tt]
    const uint32_t wait_bar0 = (wi == 0) ? bar_full : bar_edge + 8u * NB * wi;
    const uint32_t next_bar0 = last_warp ? bar_empty : bar_edge + 8u * NB * (wi + 1);
    uint32_t *bits_w = p.bits + d.bits_off + (static_cast<long long>(wi) * KF) * 32 + lane;   // next word to flush
    const int bits_step = W * KF * 32;                                         // words per 32 frames

    int t = 0;                      // frames consumed so far == current trellis row
    int b = -1;                     // ring position and lap parity of the current tile
    uint32_t ph = 0;
    // Trellis column j is -inf in every row t < j (:369), so until frame `first column of this warp`
    // all of the warp's cells are -inf and all of its choice bits 0 (-inf > -inf is false): whole
    // tiles below that frame are passed on without running the recurrence.
    const int skip_tiles = DEBUG ? 0 : min(col0w / TE, ntiles);
    int tau = 0;
    for (; tau < skip_tiles; ++tau) {
        ++b;
        const bool wrap = (b == NB);
        b = wrap ? 0 : b;
        ph ^= wrap ? 1u : 0u;
        const int n = min(TE, T - tau * TE);
        mbar_wait(wait_bar0 + 8u * b, ph);
        if (lane < n) sts_f32(left_in0 + left_stride_w + static_cast<uint32_t>(b * TE + lane) * 4u, -INFINITY);
        t += n;
        if ((t & 31) == 0) {            // a 32-frame choice word is complete: all zeros
#pragma unroll
            for (int i = 0; i < KC; ++i) bits_w[i * 32] = 0u;
            bits_w += bits_step;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(next_bar0 + 8u * b);
    }
    for (; tau < ntiles; ++tau) {
        // branch-free ring step: a select keeps b / ph (and every address derived from them) in
        // uniform registers, an `if` does not
        ++b;
        const bool wrap = (b == NB);
        b = wrap ? 0 : b;
        ph ^= wrap ? 1u : 0u;
        const int n = min(TE, T - tau * TE);
        mbar_wait(wait_bar0 + 8u * b, ph);

        uint32_t rowaddr = sm_base + b * L.tile_bytes;
        uint32_t left_in = left_in0 + static_cast<uint32_t>(b * TE) * 4u;     // advances with the tile
        uint32_t left_out = left_in + left_stride_w;
        uint32_t eb_in = sm_base + L.off_eb + static_cast<uint32_t>(b * TE) * 4u;
        uint32_t wild_in = sm_base + L.off_wild + static_cast<uint32_t>(b * TE) * 4u;
        asm volatile("" : "+r"(left_in), "+r"(left_out), "+r"(eb_in));   // materialise once per tile, not per frame under a predicate

        // ---- the recurrence (:372-378): one frame ----
        // eb = logp[t, blank]; lv = trellis[t, first column of this warp - 1] (used by lane 0); wv = wildcard emission
        auto step = [&](float eb, float lv, float wv) {
            float left = __shfl_up_sync(FULL, src, 1);
            if (lane == 0) left = lv;
            // chg first (it reads the *old* left neighbours), then stay / sign-difference as packed
            // f32x2 adds (FADD2 on sm_100: one issue slot for two cells; IEEE round-to-nearest per lane)
            float chg[KC];
#pragma unroll
            for (int i = KC - 1; i >= 0; --i) {
                float e = lds_f32(rowaddr + tokoff[i]);
                if (WILD && ((wmask >> i) & 1u)) e = wv;
                chg[i] = ((i == 0) ? left : R[i - 1]) + e;    // trellis[t, j-1] + emission[t, tokens[j-1]]
            }
#pragma unroll
            for (int i = 0; i + 1 < KC; i += 2) {
                const float2 st = __fadd2_rn(make_float2(R[i], R[i + 1]), make_float2(eb, eb));   // trellis[t, j] + emission[t, blank]
                const float2 dd = __fadd2_rn(st, make_float2(-chg[i], -chg[i + 1]));             // sign set <=> chg > stay (strict)
                R[i] = fmax_nan(st.x, chg[i]);
                R[i + 1] = fmax_nan(st.y, chg[i + 1]);
                acc[i] = __funnelshift_l(__float_as_uint(dd.x), acc[i], 1);
                acc[i + 1] = __funnelshift_l(__float_as_uint(dd.y), acc[i + 1], 1);
            }
            if (KC & 1) {
                const float stay = R[KC - 1] + eb;
                const float dd = stay - chg[KC - 1];
                R[KC - 1] = fmax_nan(stay, chg[KC - 1]);
                acc[KC - 1] = __funnelshift_l(__float_as_uint(dd), acc[KC - 1], 1);
            }
            rowaddr += rowstep;
            ++t;
            if (KC >= 2 && MIXED) src = wide ? R[KC - 1] : R[(KC >= 2) ? KC - 2 : 0];
            else src = R[KC - 1];
            if (DEBUG) {
                float *tr = p.trellis + d.trellis_off + static_cast<long long>(t) * (N + 1);
#pragma unroll
                for (int i = 0; i < KC; ++i)
                    if (i < ncols) tr[1 + cb + i] = R[i];
                if (p.final_col && last_warp && lane == laneN) p.final_col[d.frame_base + d.seg + t] = src;
            }
        };
        auto flush_if_word_done = [&]() {
            if ((t & 31) == 0) {
#pragma unroll
                for (int i = 0; i < KC; ++i) bits_w[i * 32] = acc[i];
                bits_w += bits_step;
            }
        };

        // four frames at a time: their blank log-probs and left-boundary values arrive as one float4 each,
        // and the four rows of this warp's last column leave as one float4 (the LSU pipe is the busiest one of
        // this loop: profiles/r2_align_kernel_lines.md; a two-frame body with float2 exchanges is 2-4 % slower)
        int tt = 0;
        for (; tt + UNR <= n; tt += UNR) {
            const float4 e4 = lds_f32x4(eb_in);
            const float4 l4 = lds_f32x4(left_in);
            float4 w4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (WILD) w4 = lds_f32x4(wild_in);
            const float s0 = src; step(e4.x, l4.x, w4.x);
            const float s1 = src; step(e4.y, l4.y, w4.y);
            const float s2 = src; step(e4.z, l4.z, w4.z);
            const float s3 = src; step(e4.w, l4.w, w4.w);
            if (lane == laneN) sts_f32x4(left_out, s0, s1, s2, s3);    // rows t..t+3 of my last column -> warp wi+1
            eb_in += 16u; left_in += 16u; left_out += 16u; wild_in += 16u;
        }
        for (; tt < n; ++tt) {             // the last, partial block of a segment
            const float eb = lds_f32(eb_in), lv = lds_f32(left_in);
            const float wv = WILD ? lds_f32(wild_in) : 0.0f;
            if (lane == laneN) sts_f32(left_out, src);
            step(eb, lv, wv);
            eb_in += 4u; left_in += 4u; left_out += 4u; wild_in += 4u;
        }
        // tiles are 32, 16, 8 or 4 frames long and start at multiples of their length: a 32-frame choice word
        // can only end where a tile ends, so the check sits here and not in the frame loop (six predicated
        // instructions per four frames otherwise)
        flush_if_word_done();
        __syncwarp();
        // hand the tile on: boundary values to warp wi+1, or the buffer back to the producer
        if (lane == 0) mbar_arrive(next_bar0 + 8u * b);
    }
    if (t & 31) {   // last, partial word: left-align so that frame t' sits at bit 31-(t'%32)
        const int sh = 32 - (t & 31);
#pragma unroll
        for (int i = 0; i < KC; ++i) bits_w[i * 32] = acc[i] << sh;
    }
    if (last_warp && lane == laneN) sts_f32(sm_base + L.off_fin, src);      // trellis[T, N]
}

// WILD: the batch contains wildcard tokens (id -1, an extension over the reference — see DESIGN.md):
// their emission at frame t is max over non-blank v of logp[t, v]; the producer computes it per frame
// into its own per-tile array next to the blank log-probs.
template <int K, bool DEBUG, bool WILD>
__global__ void __maxnreg__((DEBUG || WILD) ? 64 : WXB_REGS(K)) align_kernel(const Params p, const int seg_first, const int NB, const int TE) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x & 31;
    // warp index through a shuffle: tells the compiler it is warp-uniform (uniform registers for the
    // tile / barrier addresses derived from it)
    const int wi = __shfl_sync(FULL, static_cast<int>(threadIdx.x >> 5), 0);
    const int W = (blockDim.x >> 5) - 1;            // compute warps; warp W is the producer
    const SegDesc d = p.segs[seg_first + blockIdx.x];
    const int T = d.T, N = d.N, V = p.V;
    if (p.trace && threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        p.trace[4ll * (seg_first + blockIdx.x) + 0] = global_ns();
        p.trace[4ll * (seg_first + blockIdx.x) + 3] = smid;
    }
    if (p.trace_hist && lane == 0) {       // debug: frames of work per (scheduler = hardware warp slot % 4, role)
        uint32_t hw;
        asm volatile("mov.u32 %0, %warpid;" : "=r"(hw));
        atomicAdd(p.trace_hist + (hw & 3u) * 2u + (wi == W ? 1u : 0u), static_cast<unsigned long long>(T));
    }
    const SmemLayout L = smem_layout(W, NB, TE, V, WILD);
    const uint32_t sm_base = smem_u32(sm);
    const uint32_t bar_tma = sm_base + L.off_bar;
    const uint32_t bar_full = bar_tma + 8u * NB;
    const uint32_t bar_empty = bar_full + 8u * NB;
    const uint32_t bar_edge = bar_empty + 8u * NB;     // [w][b], w = 0..W-1 (row 0 unused)
    const uint32_t left_stride_w = static_cast<uint32_t>(NB) * TE * 4u;   // bytes between left[w] and left[w+1]
    const int ntiles = (T + TE - 1) / TE;
    const float *Eg = p.emis + d.emis_off;

    if (threadIdx.x == 0) {
        for (int i = 0; i < (3 + W) * NB; ++i) mbar_init(bar_tma + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (DEBUG) {
        float *tr = p.trellis + d.trellis_off;
        const int infstart0 = T + 1 - N;
        for (int j = threadIdx.x; j <= N; j += blockDim.x)
            tr[j] = (j == 0) ? ((infstart0 <= 0) ? INFINITY : 0.0f) : -INFINITY;
        if (p.final_col && threadIdx.x == 0) p.final_col[d.frame_base + d.seg] = -INFINITY;
    }
    fence_proxy_async();
    __syncthreads();

    int t_start = 0;                // first argmax of trellis[:, N]; meaningful in the producer warp after the forward pass
    float best_val = -INFINITY;

    if (wi == W) {
        // ================================================================ producer warp
        const bool bulk_ok = ((reinterpret_cast<uintptr_t>(Eg) & 15) == 0);
        double *xd_s = reinterpret_cast<double *>(sm + L.off_xd);
        double cum = 0.0;               // float64 running sum of logp[:, 0]  (torch.cumsum, :368)
        const int infstart = T + 1 - N; // rows >= infstart of column 0 are +inf (:370)

        // ring position of the tile being issued (bi, lap parity pi) and processed (b, ph): plain
        // counters instead of tau % NB, tau / NB keep the addresses in uniform registers
        int bi = 0;
        uint32_t pi = 0, lap = 0;
        // Running first-argmax of trellis[:, N] (torch.argmax, :396).  The last compute warp leaves the
        // rows a tile started from in the boundary array of the warp that would follow it; the
        // producer folds them in, tile by tile in frame order, when it takes the ring slot back.
        float best = -INFINITY;         // trellis[0, N] = -inf
        int bestt = 0;
        int sb = 0, st0 = 0;            // ring slot and first row of the next tile to scan
        uint32_t sph = 0;
        auto scan_next = [&]() {
            mbar_wait(bar_empty + 8u * sb, sph);
            const float *fin = reinterpret_cast<const float *>(sm + L.off_left) + (W * NB + sb) * TE;
            float v = (st0 + lane < T && lane < TE) ? fin[lane] : -INFINITY;
            int vi = st0 + lane;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(FULL, v, off);
                const int oi = __shfl_xor_sync(FULL, vi, off);
                if (ov > v || (ov == v && oi < vi)) { v = ov; vi = oi; }
            }
            if (v > best) { best = v; bestt = vi; }
            st0 += TE;
            ++sb;
            const bool wraps = (sb == NB);
            sb = wraps ? 0 : sb;
            sph ^= wraps ? 1u : 0u;
            __syncwarp();
        };
        auto issue = [&](int tau) -> bool {     // start the copy of tile tau into its ring buffer
            const int b = bi;
            if (lap) scan_next();               // waits until the last compute warp is done with the slot's previous tile
            ++bi;
            const bool wrapi = (bi == NB);
            bi = wrapi ? 0 : bi;
            pi ^= wrapi ? 1u : 0u;
            lap |= wrapi ? 1u : 0u;
            const int t0 = tau * TE;
            const int n = min(TE, T - t0);
            const float *srcp = Eg + static_cast<size_t>(t0) * V;
            const uint32_t dst = sm_base + b * L.tile_bytes;
            const uint32_t bytes = static_cast<uint32_t>(n * V) * 4u;
            if (bulk_ok && (bytes & 15u) == 0u) {
                if (lane == 0) {
                    fence_proxy_async();            // earlier generic reads/writes of this buffer
                    mbar_expect_tx(bar_tma + 8u * b, bytes);
                    bulk_g2s(dst, srcp, bytes, bar_tma + 8u * b);
                }
                return true;
            }
            float *tb = reinterpret_cast<float *>(sm + b * L.tile_bytes);
            for (int i = lane; i < n * V; i += 32) tb[i] = __ldg(srcp + i);
            return false;
        };

        // (a single loop with one call site of issue() — step s issues tile s and processes tile s-1 — is 7 % slower
        // on the benchmark batch: measured A/B, profiles/r2_ablation.md)
        bool cur_bulk = (ntiles > 0) ? issue(0) : false;
        int b = -1;
        uint32_t ph = 0;
        for (int tau = 0; tau < ntiles; ++tau) {
            ++b;
            const bool wrap = (b == NB);
            b = wrap ? 0 : b;
            ph ^= wrap ? 1u : 0u;
            const int t0 = tau * TE;
            const int n = min(TE, T - t0);
            bool next_bulk = false;
            if (tau + 1 < ntiles) next_bulk = issue(tau + 1);
            if (cur_bulk) mbar_wait(bar_tma + 8u * b, ph);
            cur_bulk = next_bulk;
            __syncwarp();
            float *tb = reinterpret_cast<float *>(sm + b * L.tile_bytes);
            // ---- kernel (i), fused: log_softmax of the tile rows in place (one row per lane) ----
            if (p.is_logits) {
                if (lane < n) {         // odd V <= WIDE_V only (conflict-free); other batches are normalised by log_softmax_kernel first
                    float *row = tb + lane * V;
                    float m, ls;
                    row_lse([&](int v) { return row[v]; }, V, m, ls);
                    for (int v = 0; v < V; ++v) row[v] = logp_from(row[v], m, ls);
                    p.rowmax[d.frame_base + t0 + lane] = m;
                    p.rowlse[d.frame_base + t0 + lane] = ls;
                }
                __syncwarp();
                if (p.logprobs) {
                    float *lp = p.logprobs + (d.frame_base + t0) * static_cast<long long>(V);
                    for (int i = lane; i < n * V; i += 32) lp[i] = tb[i];
                }
            }
            // ---- column 0 of the trellis for rows t0 .. t0+n-1 (exclusive float64 prefix) ----
            float *col0 = reinterpret_cast<float *>(sm + L.off_left) + b * TE;     // left[0][b][:] = trellis[t, 0]
            float *ebs = reinterpret_cast<float *>(sm + L.off_eb) + b * TE;        // logp[t, blank]
            const float e0 = (lane < n) ? tb[lane * V] : 0.0f;
            if (lane < n) p.p0[d.frame_base + t0 + lane] = expf(e0);
            float mine_f = INFINITY;
            if (t0 < infstart) {
                xd_s[lane] = (lane < n) ? static_cast<double>(e0) : 0.0;
                __syncwarp();
                // the same sequential chain in every lane (torch.cumsum's order); prefix i is parked in shared memory
                // (every lane stores the same value) and picked up by lane i afterwards
                double *pre_s = xd_s;               // in place: addend i is read before prefix i overwrites it
#pragma unroll 2
                for (int i = 0; i < TE; i += 2) {
                    const double2 x2 = *reinterpret_cast<const double2 *>(xd_s + i);
                    pre_s[i] = cum;
                    cum += x2.x;
                    pre_s[i + 1] = cum;
                    cum += x2.y;
                }
                __syncwarp();
                if (t0 + lane < infstart) mine_f = static_cast<float>(pre_s[lane]);
                __syncwarp();
            }
            float wild_wide = -INFINITY;
            const bool wild_coop = V > WIDE_V || (V & 1) == 0;      // rows too long for one lane, or all in the same banks
            if (WILD && wild_coop) {                // the warp scans the rows one after another
                for (int f = 0; f < n; ++f) {
                    float wv = -INFINITY;
                    for (int v = lane; v < V; v += 32)
                        if (v != p.blank) wv = fmaxf(wv, tb[f * V + v]);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) wv = fmaxf(wv, __shfl_xor_sync(0xffffffffu, wv, o));
                    if (f == lane) wild_wide = wv;
                }
            }
            if (lane < n) {
                const float ebl = tb[lane * V + p.blank];
                col0[lane] = mine_f;
                ebs[lane] = ebl;
                if (WILD) {
                    float wv = -INFINITY;           // max over the non-blank vocabulary
                    if (!wild_coop) {
                        for (int v = 0; v < V; ++v)
                            if (v != p.blank) wv = fmaxf(wv, tb[lane * V + v]);
                    } else {
                        wv = wild_wide;
                    }
                    p.wild[d.frame_base + t0 + lane] = wv;
                    (reinterpret_cast<float *>(sm + L.off_wild) + b * TE)[lane] = wv;
                }
            }
            if (DEBUG) {
                float *tr = p.trellis + d.trellis_off;
                if (lane < n) tr[static_cast<long long>(t0 + lane) * (N + 1)] = mine_f;
                if (tau == ntiles - 1 && lane == 0)
                    tr[static_cast<long long>(T) * (N + 1)] = (T >= infstart) ? INFINITY : static_cast<float>(cum);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8u * b);
        }
        while (st0 < T) scan_next();    // the tiles still in flight when the last one was issued
        t_start = bestt;
        best_val = best;
    } else if (wi < W - 1) {
        compute_warp<K, false, DEBUG, WILD>(p, d, L, sm_base, NB, TE, ntiles, wi, lane, W, K);
    } else {
        // the last warp: its own column count (warp-uniform switch; every case is the same loop for another KC)
        switch (d.KL) {
            case 1: compute_warp<1, true, DEBUG, WILD>(p, d, L, sm_base, NB, TE, ntiles, wi, lane, W, K); break;
            case 2: if constexpr (K >= 2) compute_warp<2, true, DEBUG, WILD>(p, d, L, sm_base, NB, TE, ntiles, wi, lane, W, K); break;
            case 3: if constexpr (K >= 3) compute_warp<3, true, DEBUG, WILD>(p, d, L, sm_base, NB, TE, ntiles, wi, lane, W, K); break;
            case 4: if constexpr (K >= 4) compute_warp<4, true, DEBUG, WILD>(p, d, L, sm_base, NB, TE, ntiles, wi, lane, W, K); break;
            case 5: if constexpr (K >= 5) compute_warp<5, true, DEBUG, WILD>(p, d, L, sm_base, NB, TE, ntiles, wi, lane, W, K); break;
            case 6: if constexpr (K >= 6) compute_warp<6, true, DEBUG, WILD>(p, d, L, sm_base, NB, TE, ntiles, wi, lane, W, K); break;
            case 7: if constexpr (K >= 7) compute_warp<7, true, DEBUG, WILD>(p, d, L, sm_base, NB, TE, ntiles, wi, lane, W, K); break;
            default: if constexpr (K >= 8) compute_warp<8, true, DEBUG, WILD>(p, d, L, sm_base, NB, TE, ntiles, wi, lane, W, K); break;
        }
    }
    __syncthreads();      // every warp's choice words (and trellis[T, N]) are visible to the walker
    if (wi != W) return;  // the producer warp holds the arg-max and walks back
    // best_val covers rows 0..T-1; row T wins only if strictly larger (first maximal index)
    if (T > 0 && *reinterpret_cast<const float *>(sm + L.off_fin) > best_val) t_start = T;
    if (lane == 0) {
        p.nframes[d.seg] = t_start;            // arg-max of the last column; the walk starts here (backtrack_kernel)
        if (p.trace) p.trace[4ll * (seg_first + blockIdx.x) + 1] = global_ns();
    }
    // backtrack runs in backtrack_kernel (one warp per segment, no shared memory): the CTA gives its
    // registers and shared memory back as soon as the forward pass is done
}

struct FinBlock { int desc; int k0; };   // finalize_kernel work item (see wxb_align.cu)

}  // namespace wxb

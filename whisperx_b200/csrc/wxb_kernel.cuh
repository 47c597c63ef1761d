// Fused forced-alignment kernel for sm_100a: one warp (= one CTA) per transcript segment.
//
// Replaces, for a whole batch of segments, the reference's per-segment Python:
//   log_softmax            whisperx/alignment.py:233      (fused into the tile staging, optional)
//   get_trellis            whisperx/alignment.py:359-379  (phase 1: recurrence, token axis in registers)
//   backtrack              whisperx/alignment.py:387-421  (phase 2: walk over packed choice bits)
//   merge_repeats          whisperx/alignment.py:438-454  (finalize_kernel: per-token start/end/mean prob)
//   char timestamps        whisperx/alignment.py:252-264  (finalize_kernel: exact decimal rounding)
//
// Data movement: the segment's emission rows [T,V] are streamed HBM -> shared memory exactly once,
// TE frames per tile, double buffered, with one cp.async.bulk (TMA bulk copy, mbarrier completion)
// per tile when the rows are 16-byte aligned and 4-byte cp.async otherwise.  Nothing else of size
// O(T*V) or O(T*N*4) ever exists: the trellis lives in registers (K token columns per lane), and
// what backtrack needs is a 1-bit-per-cell stay/change matrix (T*N/8 bytes, L2 resident) plus
// the running first-argmax of the last column.
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
    int T, N, seg, K;
};

struct Params {
    const float *emis;
    const int *tokens;      // [sum_N] token ids
    const int *tok4;        // [sum_N] token ids * 4 (byte offset inside an emission row)
    const SegDesc *segs;
    uint32_t *bits;         // [sum ceil(T/32)*K*32] choice words: word (w,slot,lane), bit 31-(t%32)
    float *p0;              // [sum_T] expf(logp[t, 0])
    float *rowmax;          // [sum_T] (logits mode) row max
    float *rowlse;          // [sum_T] (logits mode) log(sum(exp(x - max)))
    int V, blank, TE, is_logits;
    int *status, *nframes;
    int *c_start, *c_end;
    double *c_score, *c_start_s, *c_end_s, *c_score_r;
    const double *seg_t1, *seg_ratio;
    const unsigned char *seg_round_np;   // [S] or null: 1 = numpy rounding for this segment's timestamps
    int *path_tok;
    float *path_logp, *final_col, *trellis, *logprobs;
    unsigned long long *trace;   // debug: [4 * n_desc] {t_begin, t_forward_done, t_end, smid} per CTA, or null
};

#ifndef WXB_UNROLL
#define WXB_UNROLL 2      // frames per unrolled block of the recurrence (must divide 32)
#endif
constexpr uint32_t FULL = 0xffffffffu;
constexpr int SM_TAIL_BYTES = 144 + 256 + 16;   // col0[36] floats + xd[32] doubles + 2 mbarriers

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(__cvta_generic_to_global(src)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// token k (0-based) -> owning lane / register slot, for the "r lanes hold K columns, the rest K-1" split
template <int K>
__device__ __forceinline__ void owner_of(int k, int r, int &ln, int &slot) {
    if (K == 1 || k < r * K) { ln = k / K; slot = k - ln * K; }
    else { const int kk = k - r * K; const int q = kk / (K > 1 ? K - 1 : 1); ln = r + q; slot = kk - q * (K - 1); }
}

template <int K, bool DEBUG>
__global__ void __launch_bounds__(32) align_kernel(const Params p, const int seg_first) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x;
    const SegDesc d = p.segs[seg_first + blockIdx.x];
    const int T = d.T, N = d.N, V = p.V, TE = p.TE;
    if (p.trace && lane == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        p.trace[4ll * (seg_first + blockIdx.x) + 0] = global_ns();
        p.trace[4ll * (seg_first + blockIdx.x) + 3] = smid;
    }
    const uint32_t tile_bytes = static_cast<uint32_t>(TE * V) * 4u;
    float *col0_s = reinterpret_cast<float *>(sm + 2 * tile_bytes);
    double *xd_s = reinterpret_cast<double *>(sm + 2 * tile_bytes + 144);
    const uint32_t sm_base = smem_u32(sm);
    const uint32_t bar0 = sm_base + 2 * tile_bytes + 144 + 256;

    // ---- column ownership: lanes [0,r) hold K token columns, lanes [r,32) hold K-1 ----
    const int r = N - 32 * (K - 1);                 // 1..32 by construction of the K bucket
    const bool wide = lane < r;
    const int cb = wide ? lane * K : r * K + (lane - r) * (K - 1);
    const int ncols = wide ? K : K - 1;
    const int laneN = (K == 1) ? (N - 1) : 31;      // lane whose last real column is column N

    uint32_t tokoff[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        // byte offsets (token*4) come pre-multiplied from the plan: a shift visible in PTX makes
        // ptxas emit one LEA per cell per frame instead of folding LDS [R + UR] (row base is uniform)
        tokoff[i] = (i < ncols) ? static_cast<uint32_t>(__ldg(p.tok4 + d.tok_base + cb + i)) : 0u;
    }
    float R[K];
    uint32_t acc[K];
#pragma unroll
    for (int i = 0; i < K; ++i) { R[i] = -INFINITY; acc[i] = 0u; }   // trellis[0, 1:] = -inf  (:369)
    float src = -INFINITY;          // this lane's last real column at the current row
    float best = -INFINITY;         // running max of trellis[:, N]; trellis[0, N] = -inf
    int bestt = 0;                  // first argmax (torch.argmax, :396)
    double cum = 0.0;               // float64 running sum of logp[:, 0]  (torch.cumsum, :368)
    const int infstart = T + 1 - N; // rows >= infstart of column 0 are +inf (:370)

    const float *Eg = p.emis + d.emis_off;
    const bool bulk_ok = ((reinterpret_cast<uintptr_t>(Eg) & 15) == 0);
    const int ntiles = (T + TE - 1) / TE;

    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncwarp();

    uint32_t phase_bits = 0;        // bit b = parity to wait for on buffer b
    uint32_t bulk_bits = 0;         // bit b = buffer b was filled by a bulk copy

    auto issue_tile = [&](int tau, int b) {
        const int t0 = tau * TE;
        const int n = min(TE, T - t0);
        const float *srcp = Eg + static_cast<size_t>(t0) * V;
        const uint32_t dst = sm_base + b * tile_bytes;
        const uint32_t bytes = static_cast<uint32_t>(n * V) * 4u;
        if (bulk_ok && (bytes & 15u) == 0u) {
            if (lane == 0) {
                fence_proxy_async();            // earlier generic reads/writes of this buffer
                mbar_expect_tx(bar0 + 8 * b, bytes);
                bulk_g2s(dst, srcp, bytes, bar0 + 8 * b);
            }
            bulk_bits |= (1u << b);
        } else {
            for (int i = lane; i < n * V; i += 32) cp_async4(dst + 4u * i, srcp + i);
            bulk_bits &= ~(1u << b);
        }
    };

    if (DEBUG) {
        float *tr = p.trellis + d.trellis_off;
        for (int j = lane; j <= N; j += 32)
            tr[j] = (j == 0) ? ((infstart <= 0) ? INFINITY : 0.0f) : -INFINITY;
    }
    if (DEBUG && p.final_col && lane == 0) p.final_col[d.frame_base + d.seg] = -INFINITY;

    if (ntiles > 0) issue_tile(0, 0);

    int t = 0;                      // frames consumed so far == current trellis row
    for (int tau = 0; tau < ntiles; ++tau) {
        const int b = tau & 1;
        const int t0 = tau * TE;
        const int n = min(TE, T - t0);
        // ---- wait for tile tau, prefetch tile tau+1 into the other buffer ----
        if (bulk_bits & (1u << b)) {
            mbar_wait(bar0 + 8 * b, (phase_bits >> b) & 1u);
            phase_bits ^= (1u << b);
        } else {
            cp_async_wait_all();
        }
        __syncwarp();
        if (tau + 1 < ntiles) issue_tile(tau + 1, b ^ 1);

        float *tb = reinterpret_cast<float *>(sm + b * tile_bytes);
        // ---- kernel (i), fused: log_softmax of the tile rows in place (one row per lane) ----
        if (p.is_logits) {
            if (lane < n) {
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
        // ---- column 0 of the trellis for rows t0 .. t0+n (exclusive float64 prefix) ----
        {
            const float e0 = (lane < n) ? tb[lane * V] : 0.0f;
            if (lane < n) p.p0[d.frame_base + t0 + lane] = expf(e0);
            if (t0 < infstart) {
                xd_s[lane] = static_cast<double>(e0);
                __syncwarp();
                for (int i = 0; i < n; ++i) {
                    col0_s[i] = static_cast<float>(cum);   // same value from every lane
                    cum += xd_s[i];
                }
                col0_s[n] = static_cast<float>(cum);
                __syncwarp();
            }
            for (int i = lane; i <= n; i += 32)
                if (t0 + i >= infstart) col0_s[i] = INFINITY;
            __syncwarp();
        }

        // ---- phase 1: the recurrence (:372-378), 8 frames unrolled ----
        uint32_t rowaddr = sm_base + b * tile_bytes;
        const uint32_t rowstep = static_cast<uint32_t>(V) * 4u;
        const uint32_t blankoff = static_cast<uint32_t>(p.blank) * 4u;

        auto step = [&](int tt) {
            const float eb = lds_f32(rowaddr + blankoff);
            const float c0 = col0_s[tt];
            float left = __shfl_up_sync(FULL, src, 1);
            if (lane == 0) left = c0;
            // chg first (it reads the *old* left neighbours), then stay / sign-difference as packed
            // f32x2 adds (FADD2 on sm_100: one issue slot for two cells; IEEE round-to-nearest per lane)
            float chg[K];
#pragma unroll
            for (int i = K - 1; i >= 0; --i) {
                const float e = lds_f32(rowaddr + tokoff[i]);
                chg[i] = ((i == 0) ? left : R[i - 1]) + e;    // trellis[t, j-1] + emission[t, tokens[j-1]]
            }
#pragma unroll
            for (int i = 0; i + 1 < K; i += 2) {
                const float2 st = __fadd2_rn(make_float2(R[i], R[i + 1]), make_float2(eb, eb));   // trellis[t, j] + emission[t, blank]
                const float2 dd = __fadd2_rn(st, make_float2(-chg[i], -chg[i + 1]));             // sign set <=> chg > stay (strict)
                R[i] = fmax_nan(st.x, chg[i]);
                R[i + 1] = fmax_nan(st.y, chg[i + 1]);
                acc[i] = __funnelshift_l(__float_as_uint(dd.x), acc[i], 1);
                acc[i + 1] = __funnelshift_l(__float_as_uint(dd.y), acc[i + 1], 1);
            }
            if (K & 1) {
                const float stay = R[K - 1] + eb;
                const float dd = stay - chg[K - 1];
                R[K - 1] = fmax_nan(stay, chg[K - 1]);
                acc[K - 1] = __funnelshift_l(__float_as_uint(dd), acc[K - 1], 1);
            }
            rowaddr += rowstep;
            ++t;
            if (K >= 2) src = wide ? R[K - 1] : R[(K >= 2) ? K - 2 : 0];
            else src = R[0];
            if (src > best) { best = src; bestt = t; }
            if (DEBUG) {
                float *tr = p.trellis + d.trellis_off + static_cast<long long>(t) * (N + 1);
#pragma unroll
                for (int i = 0; i < K; ++i)
                    if (i < ncols) tr[1 + cb + i] = R[i];
                if (lane == 0) tr[0] = col0_s[tt + 1];
            }
            if (DEBUG && p.final_col && lane == laneN) p.final_col[d.frame_base + d.seg + t] = src;
        };
        auto flush_if_word_done = [&]() {
            if ((t & 31) == 0) {
                uint32_t *w = p.bits + d.bits_off + (static_cast<long long>((t >> 5) - 1) * K) * 32 + lane;
#pragma unroll
                for (int i = 0; i < K; ++i) w[i * 32] = acc[i];
            }
        };

        int tt = 0;
        for (; tt + WXB_UNROLL <= n; tt += WXB_UNROLL) {
#pragma unroll
            for (int u = 0; u < WXB_UNROLL; ++u) step(tt + u);
            flush_if_word_done();
        }
        for (; tt < n; ++tt) {
            step(tt);
            flush_if_word_done();
        }
        // this buffer is re-filled through the async proxy two tiles from now: order every lane's
        // generic reads / in-place log-softmax writes before that copy
        fence_proxy_async();
        __syncwarp();
    }
    if (t & 31) {   // last, partial word: left-align so that frame t' sits at bit 31-(t'%32)
        const int sh = 32 - (t & 31);
        uint32_t *w = p.bits + d.bits_off + (static_cast<long long>(t >> 5) * K) * 32 + lane;
#pragma unroll
        for (int i = 0; i < K; ++i) w[i * 32] = acc[i] << sh;
    }
    __threadfence_block();
    __syncwarp();
    if (p.trace && lane == 0) p.trace[4ll * (seg_first + blockIdx.x) + 1] = global_ns();

    // ---- phase 2: backtrack (:395-421) over the choice bits; control flow is warp-uniform ----
    // The walk needs one choice word per visited (column, 32-frame word) pair: ~N + T/32
    // *dependent* loads if done naively (measured: 30% of all warp time).  Instead the warp
    // fetches a window of 32 columns x 4 words (128 frames) with 4 independent loads per lane
    // and walks inside it on register shuffles; a new window is fetched only when the path leaves
    // the block: about max(N/32, T/128) dependent round trips per segment.
    const int t_start = __shfl_sync(FULL, bestt, laneN);
    int status = 0;
    int *cstart = p.c_start + d.tok_base;
    if (t_start == 0) {
        status = 1;                                  // empty walk -> for/else -> None
    } else {
        int j = N;
        int s = t_start - 1;                         // frame of the current point
        bool done = false;
        while (!done) {
            const int j0 = j;                        // window columns j0, j0-1, ..., j0-31
            const int w0 = s >> 5;                   // window words   w0, w0-1, w0-2, w0-3
            uint32_t wA = 0u, wB = 0u, wC = 0u, wD = 0u;
            const int c = j0 - lane;
            if (c >= 1) {
                int ln, slot;
                owner_of<K>(c - 1, r, ln, slot);
                const uint32_t *bp = p.bits + d.bits_off + (static_cast<long long>(w0) * K + slot) * 32 + ln;
                wA = __ldcg(bp);
                if (w0 >= 1) wB = __ldcg(bp - 1 * K * 32);
                if (w0 >= 2) wC = __ldcg(bp - 2 * K * 32);
                if (w0 >= 3) wD = __ldcg(bp - 3 * K * 32);
            }
            while (true) {
                const int dw = w0 - (s >> 5);        // 0..3 inside the window
                const uint32_t mine = (dw == 0) ? wA : (dw == 1) ? wB : (dw == 2) ? wC : wD;
                const uint32_t word = __shfl_sync(FULL, mine, j0 - j);
                // bit (31 - f%32) is frame f; shift so that bit 0 = frame s, bit q = frame s-q
                const uint32_t m = word >> (31 - (s & 31));
                if (m == 0u) {
                    s = (s & ~31) - 1;               // stayed through the rest of this word
                    if (s < 0) { status = 1; done = true; break; }
                } else {
                    const int sc = s - (__ffs(m) - 1);   // frame at which token j-1 was entered
                    if (lane == 0) cstart[j - 1] = sc;
                    --j;
                    if (j == 0) { done = true; break; }
                    s = sc - 1;
                    if (s < 0) { status = 1; done = true; break; }
                }
                if (j < j0 - 31 || (s >> 5) < w0 - 3) break;     // left the window: fetch the next one
            }
        }
    }
    if (lane == 0) {
        p.status[d.seg] = status;
        p.nframes[d.seg] = status ? 0 : t_start;
        if (p.trace) p.trace[4ll * (seg_first + blockIdx.x) + 2] = global_ns();
    }
    // phase 3 (merge_repeats, scores, timestamps) runs token-parallel in finalize_kernel
}

struct FinBlock { int desc; int k0; };   // finalize_kernel work item (see wxb_align.cu)

}  // namespace wxb

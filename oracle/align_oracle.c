/*
 * TEST INFRASTRUCTURE ONLY — scalar C restatement of the reference's forced-alignment DP.
 *
 * This is the checker, never the product.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; nothing under whisperx_b200/ links or calls it.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference):
 *   wxo_forward_trellis  <- get_trellis     whisperx/alignment.py:359-379
 *   wxo_walk_back        <- backtrack       whisperx/alignment.py:387-421
 *   wxo_collapse_repeats <- merge_repeats   whisperx/alignment.py:438-454
 *   wxo_round3           <- Python round(x, 3) used at alignment.py:262-264,308
 *   wxo_pairwise_sum     <- numpy add.reduce order behind pandas .mean() at alignment.py:308
 *
 * The arithmetic the reference delegates to third-party code (torch ATen CPU kernels, pinned by
 * requirements.txt:1 only as torch>=2; container has torch 2.11.0) is restated from its
 * published behaviour and PINNED by tests/test_oracle_pinned.py against the reference itself
 * imported in the build container, and by the committed vectors in tests/golden/:
 *   - torch.cumsum(float32) on CPU accumulates in double and rounds every prefix to float;
 *   - torch.maximum propagates NaN;  torch.argmax returns the first maximal index, NaN counts
 *     as maximal;  `changed > stayed` is a strict compare (tie -> stay);
 *   - Tensor.exp() on a float32 scalar is a float32 exp, widened to double by .item().
 *     (glibc expf vs torch's vectorised exp differ by <= 1 ulp: scores are compared at 1e-6 rel.)
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no FMA contraction, SSE2 float math).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define WXO_OK 0
#define WXO_FAILED 1      /* backtrack() returned None */
#define WXO_NOMEM 2

static inline float max_nan(float a, float b) {      /* torch.maximum */
    if (a != a) return a;
    if (b != b) return b;
    return a > b ? a : b;
}

/* emission[t, token] as the trellis and the backtrack read it.  EXTENSION (not in the reference fork,
 * SURVEY Appendix C): token id -1 is a wildcard scored as the best non-blank symbol of the frame. */
static inline float tok_emis(const float *e, int V, int tok, int blank) {
    if (tok >= 0) return e[tok];
    float m = -INFINITY;
    for (int v = 0; v < V; ++v)
        if (v != blank && e[v] > m) m = e[v];
    return m;
}

/* get_trellis, alignment.py:359-379.  R is (T+1) x (N+1), row-major. */
void wxo_forward_trellis(const float *E, int T, int V, const int32_t *tok, int N, int blank,
                         float *R) {
    const int W = N + 1;
    /* :367  trellis[0,0] = 0 */
    R[0] = 0.0f;
    /* :368  trellis[1:,0] = cumsum(emission[:,0])  — literal column 0; double accumulator */
    double acc = 0.0;
    for (int t = 0; t < T; ++t) {
        acc += (double)E[(size_t)t * V];
        R[(size_t)(t + 1) * W] = (float)acc;
    }
    /* :369  trellis[0,-N:] = -inf   (Python slice: N==0 means the whole row) */
    {
        int c0 = (N == 0) ? 0 : (W - N < 0 ? 0 : W - N);
        for (int j = c0; j < W; ++j) R[j] = -INFINITY;
    }
    /* :370  trellis[-N:,0] = +inf   (clamps: N >= T+1 covers every row, incl. R[0,0]) */
    {
        int r0 = (N == 0) ? 0 : (T + 1 - N < 0 ? 0 : T + 1 - N);
        for (int t = r0; t <= T; ++t) R[(size_t)t * W] = INFINITY;
    }
    /* :372-378 */
    for (int t = 0; t < T; ++t) {
        const float *row = R + (size_t)t * W;
        float *nxt = R + (size_t)(t + 1) * W;
        const float *e = E + (size_t)t * V;
        const float eb = e[blank];
        for (int j = 1; j <= N; ++j) {
            float stay = row[j] + eb;
            float move = row[j - 1] + tok_emis(e, V, tok[j - 1], blank);
            nxt[j] = max_nan(stay, move);
        }
    }
}

/* backtrack, alignment.py:387-421.  Points are returned in ascending time order.
 * p_tok/p_time/p_logp/p_prob need room for T entries.  Returns WXO_OK or WXO_FAILED. */
int wxo_walk_back(const float *R, const float *E, int T, int V, const int32_t *tok, int N,
                  int blank, int32_t *p_tok, int32_t *p_time, float *p_logp, double *p_prob,
                  int32_t *p_len, int32_t *t_start_out) {
    const int W = N + 1;
    int j = N;
    /* :396 torch.argmax(trellis[:, j]) — first max, NaN is maximal */
    int t_start = 0;
    float best = R[j];
    for (int t = 1; t <= T; ++t) {
        float v = R[(size_t)t * W + j];
        if ((v > best) || (v != v && best == best)) { best = v; t_start = t; }
    }
    if (t_start_out) *t_start_out = t_start;
    int n = 0, reached = 0;
    for (int t = t_start; t > 0; --t) {
        if (j == 0) break;  /* unreachable for N>=1; guards N==0 like the empty loop would */
        const float *e = E + (size_t)(t - 1) * V;
        float stayed = R[(size_t)(t - 1) * W + j] + e[blank];
        float changed = R[(size_t)(t - 1) * W + j - 1] + tok_emis(e, V, tok[j - 1], blank);
        int took = changed > stayed;
        float lp = took ? tok_emis(e, V, tok[j - 1], blank) : e[0];        /* :409 literal column 0 on a stay */
        p_tok[n] = j - 1;
        p_time[n] = t - 1;
        p_logp[n] = lp;
        p_prob[n] = (double)expf(lp);
        ++n;
        if (took) {
            --j;
            if (j == 0) { reached = 1; break; }
        }
    }
    if (!reached) { *p_len = 0; return WXO_FAILED; }
    /* path[::-1] */
    for (int a = 0, b = n - 1; a < b; ++a, --b) {
        int32_t ti = p_tok[a]; p_tok[a] = p_tok[b]; p_tok[b] = ti;
        ti = p_time[a]; p_time[a] = p_time[b]; p_time[b] = ti;
        float f = p_logp[a]; p_logp[a] = p_logp[b]; p_logp[b] = f;
        double d = p_prob[a]; p_prob[a] = p_prob[b]; p_prob[b] = d;
    }
    *p_len = n;
    return WXO_OK;
}

/* merge_repeats, alignment.py:438-454: one (start,end,score) per run of equal token_index;
 * score = Python sum (double, left to right, from integer 0) / run length. Returns #runs. */
int wxo_collapse_repeats(const int32_t *p_tok, const int32_t *p_time, const double *p_prob,
                         int n, int32_t *seg_tok, int32_t *seg_start, int32_t *seg_end,
                         double *seg_score) {
    int a = 0, m = 0;
    while (a < n) {
        int b = a;
        while (b < n && p_tok[b] == p_tok[a]) ++b;
        double s = 0.0;
        for (int k = a; k < b; ++k) s += p_prob[k];
        seg_tok[m] = p_tok[a];
        seg_start[m] = p_time[a];
        seg_end[m] = p_time[b - 1] + 1;
        seg_score[m] = s / (double)(b - a);
        ++m;
        a = b;
    }
    return m;
}

/* One segment end to end (alignment.py:242-250).  Outputs have room for N entries.
 * `R` may be NULL (scratch is malloc'ed).  Returns WXO_OK / WXO_FAILED / WXO_NOMEM. */
int wxo_align_segment(const float *E, int T, int V, const int32_t *tok, int N, int blank,
                      float *R, int32_t *c_start, int32_t *c_end, double *c_score,
                      int32_t *t_start_out) {
    float *own = NULL;
    if (!R) {
        own = (float *)malloc(sizeof(float) * (size_t)(T + 1) * (size_t)(N + 1));
        if (!own) return WXO_NOMEM;
        R = own;
    }
    size_t cap = (size_t)(T > 0 ? T : 1);
    int32_t *p_tok = (int32_t *)malloc(cap * sizeof(int32_t) * 2);
    float *p_logp = (float *)malloc(cap * sizeof(float));
    double *p_prob = (double *)malloc(cap * sizeof(double));
    int32_t *s_tok = (int32_t *)malloc(cap * sizeof(int32_t));
    int rc = WXO_NOMEM;
    if (p_tok && p_logp && p_prob && s_tok) {
        int32_t *p_time = p_tok + cap;
        int32_t n = 0;
        wxo_forward_trellis(E, T, V, tok, N, blank, R);
        rc = wxo_walk_back(R, E, T, V, tok, N, blank, p_tok, p_time, p_logp, p_prob, &n,
                           t_start_out);
        if (rc == WXO_OK) wxo_collapse_repeats(p_tok, p_time, p_prob, n, s_tok, c_start, c_end, c_score);
    }
    free(p_tok); free(p_logp); free(p_prob); free(s_tok); free(own);
    return rc;
}

/* Python round(x, 3): correctly rounded decimal (half-even on the exact binary value), then the
 * nearest double to that decimal.  CPython float.__round__ -> _Py_dg_dtoa(mode 3) + strtod;
 * glibc printf("%.3f") / strtod are both exact, so this is the definition. */
double wxo_round3(double x) {
    if (!(x == x) || isinf(x)) return x;
    char buf[512];
    snprintf(buf, sizeof buf, "%.3f", x);
    return strtod(buf, NULL);
}

/* round(x, 3) on a numpy float64 scalar (np.float64.__round__ == np.round: scale, rint, unscale);
 * this is what `round(word_chars["score"].mean(), 3)` evaluates (alignment.py:308). */
double wxo_round3_np(double x) { return rint(x * 1000.0) / 1000.0; }

/* numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE),
 * as reached by pandas nanmean -> ndarray.sum() on a contiguous float64 vector. */
static double pw(const double *a, long n) {
    if (n < 8) {
        double r = -0.0;
        for (long i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        long i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    long n2 = n / 2;
    n2 -= n2 % 8;
    return pw(a, n2) + pw(a + n2, n - n2);
}

double wxo_pairwise_sum(const double *a, long n) {
    return 0.0 + pw(a, n);   /* add.reduce starts from the identity 0 */
}

/* Batch driver for CPU-baseline timing: S segments, concatenated inputs (same layout as the
 * engine's C-ABI).  Returns the number of failed segments. */
long wxo_align_batch(const float *E, const int64_t *frame_off, const int32_t *tokens,
                     const int64_t *tok_off, long S, int V, int blank, int32_t *status,
                     int32_t *c_start, int32_t *c_end, double *c_score) {
    long failed = 0;
    for (long s = 0; s < S; ++s) {
        int T = (int)(frame_off[s + 1] - frame_off[s]);
        int N = (int)(tok_off[s + 1] - tok_off[s]);
        int32_t ts = 0;
        int rc = wxo_align_segment(E + (size_t)frame_off[s] * V, T, V, tokens + tok_off[s], N, blank,
                                   NULL, c_start + tok_off[s], c_end + tok_off[s],
                                   c_score + tok_off[s], &ts);
        status[s] = rc;
        failed += (rc != WXO_OK);
    }
    return failed;
}

/*
 * wxb_align.h — C ABI of the B200 forced-alignment engine (libwxb_align.so).
 *
 * The reference (whisperX 3.1.1, pure Python) has no FFI for this path; these entry points are
 * what a binding for whisperx/alignment.py would call instead of its per-segment Python code.
 * Every entry point names the reference lines it replaces (paths relative to the reference
 * checkout).  Plain pointers and sizes only — no torch / C++ types cross this boundary.
 *
 * Batch layout ("concatenated segments"), shared by all calls:
 *   emissions  float32 [rows, V] row-major:  segment s owns rows frame_off[s] .. frame_off[s]+T_s
 *   tokens     int32   [sum_N]:              segment s owns     tok_off[s]   .. tok_off[s+1]
 * which is the batched form of the per-segment `emission[T,V]` / `tokens` list the reference
 * builds at alignment.py:209-235.
 *
 * Error convention: every function returns WXB_OK (0) or a negative WXB_E* code and never
 * throws; wxb_last_error() returns a thread-local, human-readable message for the last failure.
 * A segment that cannot be aligned is NOT an error: its status[] entry is WXB_SEG_FAILED, the
 * batched equivalent of `backtrack()` returning None (alignment.py:418-420, handled at :245-248).
 */
#ifndef WXB_ALIGN_H
#define WXB_ALIGN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WXB_ABI_VERSION 1

/* return codes */
#define WXB_OK            0
#define WXB_EINVAL       -1   /* bad argument (null pointer, V out of range, offsets not monotone ...) */
#define WXB_ECUDA        -2   /* CUDA runtime error; message carries cudaGetErrorString */
#define WXB_ENOMEM       -3   /* host or device allocation failed */
#define WXB_EUNSUPPORTED -4   /* shape outside what this build handles (see wxb_limits) */

/* per-segment status[] values */
#define WXB_SEG_OK      0
#define WXB_SEG_FAILED  1     /* backtrack() -> None: no monotone path reaches token 0 (e.g. N > T) */

/* flags for wxb_batch.flags */
#define WXB_F_LOGITS        1u   /* emissions are raw logits: fuse log_softmax(dim=-1) (alignment.py:233) */
#define WXB_F_NO_SHORTCUT   2u   /* run the DP even for N > T (tests); default marks them FAILED on the host */
#define WXB_F_WILDCARD      4u   /* EXTENSION (not in the reference fork): token id -1 is a wildcard whose emission at
                                    frame t is max over non-blank v of logp[t, v]; its path probability is exp of that */

#define WXB_F_CLOBBER       8u   /* the emission buffer is the caller's scratch: the engine may overwrite it (with
                                    WXB_F_LOGITS, batches that are normalised before the recurrence are then
                                    normalised in place instead of into a copy; run such a plan once per buffer) */

typedef struct wxb_ctx wxb_ctx;     /* per-device workspace (scratch buffers, side streams) */
typedef struct wxb_plan wxb_plan;   /* a batch description uploaded and bucketed once, runnable many times */

/* Input description.  All pointers are HOST pointers (small metadata, copied once per plan). */
typedef struct wxb_batch {
    int64_t        n_segments;   /* S */
    int32_t        vocab;        /* V: emission row width */
    int32_t        blank_id;     /* alignment.py:237-240 */
    uint32_t       flags;        /* WXB_F_* */
    uint32_t       reserved;
    const int64_t *frame_off;    /* [S+1] ascending row offsets; segment s starts at row frame_off[s] */
    const int32_t *frame_len;    /* [S] or NULL.  NULL: T_s = frame_off[s+1]-frame_off[s].  Otherwise
                                    T_s = frame_len[s] <= frame_off[s+1]-frame_off[s]; the rows in between
                                    are padding (e.g. segment starts rounded up to 4 rows so that every
                                    segment is 16-byte aligned and takes the bulk-copy path). */
    const int64_t *tok_off;      /* [S+1] ascending; N_s = tok_off[s+1]-tok_off[s] >= 0 */
    const int32_t *tokens;       /* [tok_off[S]] values in [0, V); -1 = wildcard with WXB_F_WILDCARD */
    /* Optional timestamp conversion (alignment.py:252-264): second = round(frame*ratio + t1, 3).
     * Both NULL -> the *_s / *_r / word outputs are not produced. */
    const double  *seg_t1;       /* [S] segment start time t1 (alignment.py:184) */
    const double  *seg_ratio;    /* [S] duration/T as the reference computes it (alignment.py:253) */
    const uint8_t *seg_round_np; /* [S] or NULL.  1: the segment's start/end were numpy scalars in the caller, so
                                    `frame*ratio + t1` is a np.float64 and round(.,3) is numpy's rint(x*1000)/1000;
                                    0 / NULL: Python float rounding (exact decimal, half-even).  Scores: char scores
                                    always round the Python way (:264), word scores always the numpy way (:308). */
    /* Optional word groups (alignment.py:297-320): word w of the batch owns entries
     * wchar_off[w] .. wchar_off[w+1] of wchar_idx; segment s owns words word_off[s] .. word_off[s+1].
     * wchar_idx holds, for each non-space character of the word in text order, the index of its
     * aligned character inside the segment (0 .. N_s-1), or -1 for a character the model cannot
     * align (those count as NaN rows in the reference's table). */
    int64_t        n_words;      /* W (0 -> no word reduction) */
    const int64_t *word_off;     /* [S+1] */
    const int64_t *wchar_off;    /* [W+1] */
    const int32_t *wchar_idx;    /* [wchar_off[W]] */
} wxb_batch;

/* Output buffers.  DEVICE pointers for wxb_plan_run / wxb_align_device, HOST pointers for
 * wxb_align_host.  Any pointer may be NULL to skip that output, except status/char_start/
 * char_end/char_score which are required. */
typedef struct wxb_outputs {
    int32_t *status;        /* [S]     WXB_SEG_*                                                  */
    int32_t *n_frames;      /* [S]     t_start = argmax_t trellis[t, N] (alignment.py:396); 0 on failure */
    int32_t *char_start;    /* [sum_N] first frame of token k   (Segment.start, alignment.py:448) */
    int32_t *char_end;      /* [sum_N] last frame + 1           (Segment.end,   alignment.py:449) */
    double  *char_score;    /* [sum_N] mean path probability    (Segment.score, alignment.py:444), unrounded */
    double  *char_start_s;  /* [sum_N] round(start*ratio+t1, 3) (alignment.py:262)  needs seg_t1/seg_ratio */
    double  *char_end_s;    /* [sum_N] round(end*ratio+t1, 3)   (alignment.py:263)                */
    double  *char_score_r;  /* [sum_N] round(score, 3)          (alignment.py:264)                */
    double  *word_start_s;  /* [W] min start over the word's aligned chars, NaN if none (alignment.py:306) */
    double  *word_end_s;    /* [W] max end                                              (alignment.py:307) */
    double  *word_score;    /* [W] round(mean(rounded char scores), 3), NaN if none     (alignment.py:308) */
    /* ---- optional introspection, used by the parity tests and the functional API ---- */
    int32_t *path_token;    /* [sum_T] token index on the path at frame t, -1 off the path (Point.token_index) */
    float   *path_logp;     /* [sum_T] log-prob whose exp() is Point.score (alignment.py:409), 0 off the path */
    float   *final_col;     /* [sum_T + S] trellis[0..T, N] per segment, at frame_off[s] + s */
    float   *trellis;       /* [sum (T+1)(N+1)] full get_trellis() table (alignment.py:359-379), segment s at
                               trellis_off = sum_{r<s} (T_r+1)(N_r+1) */
    float   *logprobs;      /* [sum_T, V] the log-probs the engine used (only with WXB_F_LOGITS) */
} wxb_outputs;

typedef struct wxb_limits {
    int32_t max_vocab;          /* largest V the fused emission staging handles */
    int32_t max_tokens;         /* largest N per segment (32 lanes x max warps x max K) */
    int32_t tokens_per_lane;    /* largest K (token columns held in registers per lane) */
    int32_t sm_count;           /* SMs of the device the ctx was created on */
} wxb_limits;

/* ---- lifecycle ---- */
int  wxb_abi_version(void);
const char *wxb_last_error(void);
int  wxb_create(int device, wxb_ctx **out);             /* binds to CUDA device `device` */
void wxb_destroy(wxb_ctx *ctx);
int  wxb_get_limits(const wxb_ctx *ctx, wxb_limits *out);

/* ---- plan: validate + bucket by (warps, tokens-per-lane) + cost-sort + upload descriptors ----
 * Replaces the per-segment token lookup and loop set-up of alignment.py:182-210. */
int  wxb_plan_create(wxb_ctx *ctx, const wxb_batch *batch, wxb_plan **out);
void wxb_plan_destroy(wxb_plan *plan);
/* number of kernel launches one wxb_plan_run issues (for bench.py's gpu_launches) */
int  wxb_plan_launch_count(const wxb_plan *plan);
/* sum over runnable segments of T*N (trellis cells) and of the algorithmic bytes 4TV+20N */
int  wxb_plan_stats(const wxb_plan *plan, int64_t *cells, int64_t *alg_bytes, int64_t *n_shortcut);

/* ---- the hot path ----
 * wxb_plan_run: get_trellis + backtrack + merge_repeats (+ timestamp/word reduction) for every
 * segment of the plan (alignment.py:242-250, 252-264, 297-320).  `emissions_dev` and all
 * non-NULL `out` pointers are DEVICE memory; work is enqueued on `stream` (a cudaStream_t cast
 * to void*; NULL = legacy default stream) and the call returns without synchronising.
 * `elapsed_dp_ms`, if not NULL, makes the call synchronise and report the CUDA-event time of
 * the DP kernels alone.  Calls that repeat the same (emissions, outputs) pointers are replayed as a
 * CUDA graph from the third one on (WXB_GRAPHS=0 disables).
 * All plans of a ctx share its scratch buffers and side streams: a run waits (on the device, through an event)
 * for the previous wxb_plan_run of the same ctx, whatever stream that was enqueued on.
 * Calls into one ctx must come from one host thread at a time; use one ctx per thread / per GPU for concurrency. */
int  wxb_plan_run(wxb_plan *plan, const float *emissions_dev, const wxb_outputs *out_dev,
                  void *stream, float *elapsed_dp_ms);

/* Overlapped runs (a serving loop over many batches): same work as wxb_plan_run, but `stream` is not made to wait
 * for it.  Calls alternate between the ctx's two run slots (each has its own scratch and side streams), so up to
 * two runs are in flight and the tail of one batch — its last, longest segments, when most SMs have run out of
 * work — is filled by the head of the next; call n+2 waits on the device for call n.  wxb_ctx_join makes `stream`
 * wait for every outstanding run of the ctx: results are complete once `stream` has passed that point.  The
 * caller must not hand an output buffer to a new run while an earlier run that writes it may still be in flight. */
int  wxb_plan_run_async(wxb_plan *plan, const float *emissions_dev, const wxb_outputs *out_dev, void *stream);
int  wxb_ctx_join(wxb_ctx *ctx, void *stream);

/* One-shot convenience: plan_create + plan_run + plan_destroy, device buffers. */
int  wxb_align_device(wxb_ctx *ctx, const wxb_batch *batch, const float *emissions_dev,
                      const wxb_outputs *out_dev, void *stream);

/* One-shot with HOST buffers: stages emissions H2D in chunks overlapped with the kernels and
 * copies the requested outputs back.  Blocking.  This is the call bench.py's `e2e` times. */
int  wxb_align_host(wxb_ctx *ctx, const wxb_batch *batch, const float *emissions_host,
                    const wxb_outputs *out_host);

/* Emission bytes per host->device chunk of wxb_align_host (default 64 MiB, env WXB_CHUNK_MB).  Any value >= 1
 * is legal — a chunk may be smaller than one segment.  Returns the previous value. */
int64_t wxb_set_chunk_bytes(wxb_ctx *ctx, int64_t bytes);

/* ---- kernel (i) on its own: log_softmax(dim=-1) over [rows, V] (alignment.py:233) ----
 * Same arithmetic, in the same order, as the fused staging inside wxb_plan_run. */
int  wxb_log_softmax_device(wxb_ctx *ctx, const float *logits_dev, int64_t rows, int32_t vocab,
                            float *out_dev, void *stream);

/* ---- backtrack from a caller-supplied trellis (functional drop-in for alignment.py:387-421) ----
 * trellis_dev [T+1, N+1], emission_dev [T, V] (log-probs), tokens_dev [N].  Writes
 * path_token/path_logp [T] (as in wxb_outputs) and *status (device pointers). */
int  wxb_backtrack_trellis_device(wxb_ctx *ctx, const float *trellis_dev, const float *emission_dev,
                                  const int32_t *tokens_dev, int32_t T, int32_t N, int32_t vocab,
                                  int32_t blank_id, int32_t *path_token_dev, float *path_logp_dev,
                                  int32_t *status_dev, int32_t *n_frames_dev, void *stream);

/* ---- speaker assignment (whisperx/diarize.py:35-67, assign_word_speakers) ----
 * For every query interval q (a segment or a word) the speaker whose diarization turns overlap it most:
 *   intersection_d = min(d_end, q_end) - max(d_start, q_start)          (float64, diarize.py:39,52)
 *   rows with intersection_d > 0 only, unless fill_nearest                (:42-45, :55-58)
 *   per-speaker sums in row order (pandas groupby-sum: Kahan compensated), arg-max, ties to the
 *   lowest speaker id (ids must follow the sorted label order, as groupby does)   (:47-48, :60-61)
 * out_speaker[q] = speaker id, or -1 when no row qualifies.  Any n_speakers (64 per launch; more are covered by
 * successive launches that carry the running winner).  Ids outside [0, n_speakers) are ignored by the device entry
 * point and rejected (WXB_EINVAL) by the host one.
 * _device: all pointers are device memory, work is enqueued on `stream`; _host: host pointers, blocking. */
int  wxb_assign_speakers_device(wxb_ctx *ctx, const double *diar_start, const double *diar_end,
                                const int32_t *diar_speaker, int64_t n_diar, int32_t n_speakers,
                                const double *query_start, const double *query_end, int64_t n_query,
                                int32_t fill_nearest, int32_t *out_speaker, void *stream);
int  wxb_assign_speakers_host(wxb_ctx *ctx, const double *diar_start, const double *diar_end,
                              const int32_t *diar_speaker, int64_t n_diar, int32_t n_speakers,
                              const double *query_start, const double *query_end, int64_t n_query,
                              int32_t fill_nearest, int32_t *out_speaker);

/* ---- debug aid: per-CTA timeline of the last wxb_plan_run (ctx created with env WXB_TRACE=1) ----
 * rows: [cap_rows][8] = {t_begin_ns, t_forward_done_ns, t_end_ns, smid, T, N, K, segment}; returns the
 * number of rows written or a negative error code. */
int  wxb_plan_debug_trace(wxb_plan *plan, uint64_t *rows, int64_t cap_rows);

/* ---- exact scores on the host (alignment.py:409, :444, :264, :308) ----
 * The reference's point probability is `emission[t, idx].exp().item()` on a CPU tensor (:409), i.e. whatever exp
 * the host's torch build dispatches to (MKL VML here: its float32 results differ between the AVX2 and AVX-512 code
 * paths in ~4 % of the values, profiles/r2_exp_host_dependence.md) — no fixed device formula can reproduce it on
 * every host.  A caller that needs the dict's `score` fields byte-identical to the reference run on the same
 * machine asks wxb_plan_run for path_logp (bit-exact float32), applies its own exp to it (the Python side calls
 * torch.exp on the host, the very op the reference calls) and hands the probabilities to this function, which
 * redoes the three reductions in the reference's order: per character the Python-order float64 sum / count
 * (:444), round(score, 3) (:264), and per word round(mean of rounded char scores, 3) in numpy's pairwise order
 * (:308).  All pointers are HOST memory; `batch` is the description the plan was made from (frame_off, tok_off and,
 * for word_score, the word groups).  char_score_r / word_score may be NULL.  No GPU work, no ctx. */
int  wxb_exact_scores_host(const wxb_batch *batch, const float *path_prob, const int32_t *status,
                           const int32_t *char_start, const int32_t *char_end,
                           double *char_score, double *char_score_r, double *word_score);

/* ---- host-side exact helpers (no GPU needed; the device code uses the same source) ---- */
double wxb_round3(double x);                              /* Python round(x, 3) (alignment.py:262-264) */
double wxb_round3_np(double x);                           /* np.float64.__round__(3) (alignment.py:308) */
double wxb_pairwise_sum(const double *a, int64_t n);      /* numpy add.reduce order (alignment.py:308) */
int    wxb_tokens_per_lane(int32_t n_tokens);             /* K (token columns per lane) for N tokens, <=0 if unsupported */
int    wxb_segment_warps(int32_t n_tokens);               /* W (compute warps of the segment's CTA), <=0 if unsupported */
/* Planner knob: every segment of a batch runs with K = k_target token columns per lane in its full warps (one
 * kernel instantiation per batch; the last warp of a segment holds what is left with KL <= K columns per lane).
 * By default (0) the planner picks per batch: 8 for vocabularies >= 64, else 6, or 3 when the batch is too small to
 * fill the GPU; 1..8 forces a value, -1 restores the default, 0 only queries.  Returns the previous value.
 * Process-wide default (atomic); a ctx created with WXB_K_TARGET set keeps its own.  wxb_tokens_per_lane /
 * wxb_segment_warps answer for the forced value, or for 4 when none is forced. */
int    wxb_set_k_target(int k_target);

#ifdef __cplusplus
}
#endif
#endif /* WXB_ALIGN_H */

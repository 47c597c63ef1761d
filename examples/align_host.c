/* Plain C against the C ABI (include/wxb_align.h): align one tiny segment from host buffers.
 *
 *   gcc -std=c99 -Iinclude examples/align_host.c -Lwhisperx_b200 -lwxb_align \
 *       -Wl,-rpath,$PWD/whisperx_b200 -o /tmp/align_host && /tmp/align_host
 *
 * The program hands wxb_align_host() what whisperx/alignment.py:242-264 works on for one segment — the
 * log-prob emission matrix [T, V], the transcript's token ids, the blank id and the segment's time span —
 * and prints one line per character: start frame, end frame, score, start second, end second.
 * No CUDA headers are needed on the caller's side. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "wxb_align.h"

int main(void) {
    enum { T = 8, V = 4, N = 3 };
    /* raw scores; normalised to log-probs below (the reference does log_softmax at alignment.py:233) */
    static const float raw[T][V] = {
        {2.0f, 0.1f, 0.0f, -1.0f}, {0.2f, 2.5f, 0.0f, -0.5f}, {0.1f, 2.0f, 0.3f, 0.0f}, {1.5f, 0.0f, 0.2f, 0.1f},
        {0.0f, 0.1f, 2.2f, 0.0f},  {0.3f, 0.0f, 1.9f, 0.4f},  {0.1f, 0.0f, 0.2f, 2.4f}, {1.8f, 0.0f, 0.1f, 0.6f}};
    float emis[T * V];
    for (int t = 0; t < T; ++t) {
        double m = raw[t][0], s = 0.0;
        for (int v = 1; v < V; ++v) m = raw[t][v] > m ? raw[t][v] : m;
        for (int v = 0; v < V; ++v) s += exp(raw[t][v] - m);
        for (int v = 0; v < V; ++v) emis[t * V + v] = (float)(raw[t][v] - m - log(s));
    }
    const int64_t frame_off[2] = {0, T}, tok_off[2] = {0, N};
    const int32_t tokens[N] = {1, 2, 3};
    const double t1[1] = {12.5}, ratio[1] = {(13.3 - 12.5) / T};

    wxb_batch b = {0};
    b.n_segments = 1; b.vocab = V; b.blank_id = 0;
    b.frame_off = frame_off; b.tok_off = tok_off; b.tokens = tokens;
    b.seg_t1 = t1; b.seg_ratio = ratio;

    int32_t status[1], n_frames[1], cs[N], ce[N];
    double score[N], start_s[N], end_s[N], score_r[N];
    wxb_outputs o = {0};
    o.status = status; o.n_frames = n_frames; o.char_start = cs; o.char_end = ce; o.char_score = score;
    o.char_start_s = start_s; o.char_end_s = end_s; o.char_score_r = score_r;

    wxb_ctx *ctx = NULL;
    if (wxb_create(0, &ctx) != WXB_OK) { fprintf(stderr, "wxb_create: %s\n", wxb_last_error()); return 2; }
    int rc = wxb_align_host(ctx, &b, emis, &o);
    if (rc != WXB_OK) { fprintf(stderr, "wxb_align_host: %s\n", wxb_last_error()); wxb_destroy(ctx); return 3; }
    printf("status %d t_start %d\n", status[0], n_frames[0]);
    for (int k = 0; k < N; ++k)
        printf("%d %d %d %.17g %.3f %.3f %.3f\n", tokens[k], cs[k], ce[k], score[k], start_s[k], end_s[k], score_r[k]);
    wxb_destroy(ctx);
    return 0;
}

// libwxb_align.so — host side of the C ABI declared in include/wxb_align.h: per-device workspace,
// batch planner ((warps, tokens-per-lane) classes, longest-first order, emission chunks), launch
// orchestration (side streams, CUDA graph replay, pipelined host path).  The kernels live in
// wxb_kernel.cuh (fused alignment kernel) and wxb_small.cuh (everything around it).  Reference lines
// replaced are cited at each entry point.
#include "../../include/wxb_align.h"
#include "wxb_kernel.cuh"
#include "wxb_small.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace wxb {
cudaError_t launch_align(int k, int warps, int nb, int te, const Params &p, int first, int count, size_t smem, cudaStream_t st, bool debug, bool wild);

constexpr int NSIDE = 8;
constexpr int SMEM_LIMIT = 227 * 1024;

}  // namespace wxb

using namespace wxb;

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(x)                                                                          \
    do {                                                                                     \
        cudaError_t e_ = (x);                                                                \
        if (e_ != cudaSuccess) return fail(WXB_ECUDA, "%s failed: %s", #x, cudaGetErrorString(e_)); \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

// One device allocation mirrored by one pinned host buffer: a plan's metadata is assembled in the
// mirror and goes up in a single cudaMemcpyAsync.  Arenas are pooled per context (grow-only).
struct Arena {
    unsigned char *dev = nullptr, *host = nullptr, *host_dev = nullptr;    // host_dev: device-side address of `host`
    size_t cap = 0;
    cudaEvent_t uploaded = nullptr;
    cudaError_t ensure(size_t bytes) {
        if (!uploaded) {
            cudaError_t e = cudaEventCreateWithFlags(&uploaded, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        if (bytes <= cap) return cudaSuccess;
        release_mem();
        const size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&dev), want);
        if (e != cudaSuccess) { dev = nullptr; return e; }
        e = cudaHostAlloc(reinterpret_cast<void **>(&host), want, cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer(reinterpret_cast<void **>(&host_dev), host, 0);
        if (e != cudaSuccess) { cudaFree(dev); if (host) cudaFreeHost(host); dev = host = host_dev = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release_mem() {
        if (dev) cudaFree(dev);
        if (host) cudaFreeHost(host);
        dev = host = host_dev = nullptr; cap = 0;
    }
    void release() { release_mem(); if (uploaded) cudaEventDestroy(uploaded); uploaded = nullptr; }
};

// Everything one run of a plan needs for itself: the scratch the kernels write (choice bits, exp(logp[t,0]), row
// statistics of the fused log-softmax ...), the side streams its classes run on and the events that fork / join
// them.  A ctx has two, so that two runs can be in flight at once (wxb_plan_run_async): the tail of one batch —
// its last, longest segments draining the GPU — overlaps the head of the next.
struct RunSlot {
    DevBuf bits, p0, rowmax, rowlse, wild, nframes_scratch, wide_lp;
    cudaStream_t side[NSIDE] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[NSIDE] = {};
    cudaEvent_t done = nullptr;               // the last run that used this slot has finished (scratch reusable, results complete)
    cudaStream_t done_on = nullptr;           // stream `done` was last recorded on (a later run on that stream need not wait)
    cudaError_t create() {
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < NSIDE && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
        for (int i = 0; i < NSIDE && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
        return e;
    }
    void destroy() {
        DevBuf *bufs[] = {&bits, &p0, &rowmax, &rowlse, &wild, &nframes_scratch, &wide_lp};
        for (DevBuf *b : bufs) b->release();
        for (int i = 0; i < NSIDE; ++i) {
            if (side[i]) cudaStreamDestroy(side[i]);
            if (ev_join[i]) cudaEventDestroy(ev_join[i]);
        }
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (done) cudaEventDestroy(done);
    }
};

struct wxb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t main = nullptr;              // used by wxb_align_host and for plan uploads
    cudaStream_t copy = nullptr;              // host->device emission chunks (wxb_align_host)
    cudaStream_t d2h = nullptr;               // per-chunk word reduction + device->host result copies (wxb_align_host)
    cudaEvent_t ev_d2h = nullptr;
    cudaStream_t jstream = nullptr;           // joins the side streams of an overlapped run (wxb_plan_run_async)
    cudaEvent_t ev_dbg[3] = {};               // WXB_DEBUG_TIMING: kernels joined / results on the host / per-chunk
    RunSlot slot[2];                          // [0]: wxb_plan_run / wxb_align_host; overlapped runs alternate between the two
    int next_async = 0;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    std::vector<cudaEvent_t> ev_chunk;        // one per emission chunk in flight
    std::vector<cudaEvent_t> ev_soft;         // wide-vocabulary logits: chunk normalised
    DevBuf trace, spk_best;
    int side_streams = NSIDE;                 // WXB_SIDE_STREAMS (1..NSIDE)
    int k_target = 0;                         // WXB_K_TARGET / wxb_ctx_set_k_target: 0 = process default, then per batch
    int tile_pref = 32;                       // WXB_TILE_FRAMES (experiments): largest tile length tried
    int ring_extra = 2;                       // WXB_RING_EXTRA (experiments): ring depth = W + this
    bool want_trace = false;                  // WXB_TRACE=1
    bool use_graphs = true;                   // WXB_GRAPHS=0 disables graph replay in wxb_plan_run
    size_t chunk_bytes = 64u << 20;           // WXB_CHUNK_MB: emission bytes per pipelined H2D chunk
    bool chunk_taper = true;                  // WXB_CHUNK_TAPER=0: no tapered tail (see make_chunk_bounds)
    std::vector<Arena *> arena_pool;          // metadata arenas of destroyed plans, ready for reuse
    // wxb_align_host staging
    DevBuf h_emis, h_i32[4], h_f64[7], h_path_tok, h_f32[4];
};

struct Bucket { int K; int W; int NB; int TE; size_t smem; int first; int count; long long cost; int maxT; int chunk; int fin_first; int fin_count; };

struct wxb_plan {
    wxb_ctx *ctx = nullptr;
    int64_t S = 0, W = 0, sum_T = 0, sum_N = 0;
    int V = 0, blank = 0;
    uint32_t flags = 0;
    bool has_sec = false, has_words = false, has_round_np = false, has_wild = false;
    std::vector<Bucket> buckets;              // chunk-major, longest segments first inside a chunk
    std::vector<int64_t> chunk_row;           // [n_chunks+1] emission row range of each chunk
    std::vector<int64_t> chunk_seg;           // [n_chunks+1] input-segment range whose last row lies in each chunk
    int n_short = 0;
    long long bits_words = 0, trellis_floats = 0, cells = 0, alg_bytes = 0;
    int n_fin_blocks = 0, n_desc = 0;
    std::vector<SegDesc> host_descs;          // kept for wxb_plan_debug_trace
    Arena *arena = nullptr;
    cudaStream_t last_stream = nullptr;
    bool ran = false;
    // wxb_plan_run replays a CUDA graph from the third identical call on (same emissions pointer, same outputs)
    const float *graph_emis = nullptr;
    wxb_outputs graph_out{};
    int graph_same = 0;                       // consecutive calls with this key
    cudaGraphExec_t graph_exec = nullptr;
    void *graph_scratch[7] = {};              // context scratch pointers baked into the graph (another plan may regrow them)
    bool graph_off = false;                   // capture failed once: stay eager
    // device views into the arena
    SegDesc *d_segs = nullptr; FinBlock *d_fin = nullptr; int *d_tok4 = nullptr; ShortSeg *d_short = nullptr;
    double *d_t1 = nullptr, *d_ratio = nullptr; unsigned char *d_round_np = nullptr;
    int *d_word_seg = nullptr; long long *d_seg_tok_base = nullptr, *d_wchar_off = nullptr; int *d_wchar_idx = nullptr;
};

// Warps and tokens-per-lane for a transcript of n tokens: as few compute warps as keep K <= k_target
// (more columns per lane amortise the per-frame overhead; more warps shorten the segment's latency).
// Process-wide *default* only (wxb_set_k_target, atomic): a plan uses its ctx's own k_target (WXB_K_TARGET at
// wxb_create) and falls back to this one, so contexts on different devices / threads share no mutable planner state.
static std::atomic<int> g_k_target{0};        // 0 = chosen per batch (see pick_k_target); 1..KMAX = forced

// The batch's K (columns per lane of every full warp).  Measured on B200 with the uniform-K planner
// (bench.py --only-main, WXB_K_TARGET sweep, profiles/r2_k_sweep.md): the 4,096-segment VAD-cut batch runs
// 0.50 / 0.47 / 0.49 / 0.46 / 0.45 / 0.46 / 0.50 ms for K = 2 .. 8, all-30-s batches 0.90 / 0.84 / 0.85 / 0.85 ms for
// K = 4 / 5 / 6 / 8, the V = 100 long-transcript batch 1.00 / 0.91 / 0.89 ms for K = 4 / 6 / 8: more columns per lane
// spend fewer issue slots per cell (the per-frame overhead is shared by more columns) until the 64 registers
// of K >= 7 cost residency.  A batch too small to fill the GPU gets short wavefronts (more warps per segment)
// instead: its segments' own latency is then all that matters.
static int pick_k_target(const int64_t *tok_off, const int64_t *frame_off, const int32_t *frame_len, int64_t S, int sm_count, int V) {
    if (V >= 64) return KMAX;
    const int k_big = 6, k_small = 3;
    long long warps = 0;
    for (int64_t s = 0; s < S; ++s) {
        const long long n = tok_off[s + 1] - tok_off[s];
        const long long t = frame_len ? frame_len[s] : frame_off[s + 1] - frame_off[s];
        if (n <= 0 || n > t) continue;
        warps += (n + 32 * k_big - 1) / (32 * k_big) + 1;
    }
    return warps >= 4ll * sm_count ? k_big : k_small;
}

// -> W compute warps, K columns per lane in the full warps (0 .. W-2), KL <= K in the last one (wxb_kernel.cuh).
// Every segment of a batch gets the batch's K (k_target) — one align_kernel instantiation, one recurrence loop
// body for all full warps: the instruction cache is what separates a fast batch from a slow one here (the same
// step with two or three instantiations co-resident on an SM runs 10-20 % slower).  Only a transcript too long
// for 31 warps of K columns gets a larger K of its own.
static bool shape_of(long long n, int *W, int *K, int k_target = 0, int *KL = nullptr) {
    if (n <= 0) return false;
    if (k_target <= 0) { const int g = g_k_target.load(std::memory_order_relaxed); k_target = g > 0 ? g : 4; }
    long long k = k_target;
    long long w = (n + 32 * k - 1) / (32 * k);
    if (w > WMAX) {
        k = (n + 32ll * WMAX - 1) / (32ll * WMAX);
        if (k > KMAX) return false;
        w = (n + 32 * k - 1) / (32 * k);
    }
    const long long m = n - 32 * k * (w - 1);             // 1 .. 32K columns left for the last warp
    *W = static_cast<int>(w);
    *K = static_cast<int>(k);
    if (KL) *KL = static_cast<int>((m + 31) / 32);
    return true;
}

// most token columns any lane of the segment's CTA holds
static int tokens_per_lane(long long n) {
    int W, K, KL;
    if (n <= 0) return 0;
    if (!shape_of(n, &W, &K, 0, &KL)) return -1;
    return W == 1 ? KL : K;
}

// Frames per shared-memory tile (cp.async.bulk needs 16-byte multiples, i.e. TE % 4 == 0): 32 when the
// widest CTA of the batch still fits a ring of at least 2 tiles, else 16, 8, 4.
// (tile_pref = WXB_TILE_FRAMES, ring_extra = WXB_RING_EXTRA: per-ctx experiment knobs, defaults 32 and 2)

// Tile length (frames) and ring depth of a W-warp CTA over a V-wide vocabulary.  32-frame tiles
// amortise the per-tile hand-offs best (measured: 16 costs +25 % on the V=29 benchmark), so they are
// kept unless a CTA would then be so large that fewer than ~16 warps fit on an SM (wide vocabularies
// with many warps); the ring is W+2 deep — every warp on its own tile plus one being filled — or
// shallower if shared memory says so (any depth >= 2 is deadlock-free).  cp.async.bulk needs
// 16-byte multiples, i.e. TE % 4 == 0.
static int pick_shape(int tile_pref, int ring_extra, int W, int V, bool wild, int *TE, int *NB, size_t *smem) {
    const int pref[4] = {32, 16, 8, 4};
    const int want_ctas = std::max(1, (16 + W) / (W + 1));
    int chosen = -1;
    for (int i = 0; i < 4 && chosen < 0; ++i) {
        if (pref[i] > tile_pref) continue;
        if (smem_layout(W, W + ring_extra, pref[i], V, wild).total * static_cast<uint32_t>(want_ctas) <= static_cast<uint32_t>(SMEM_LIMIT)) chosen = pref[i];
    }
    for (int i = 0; i < 4 && chosen < 0; ++i)         // nothing reaches the occupancy target: the longest tile that fits at all
        if (pref[i] <= tile_pref && smem_layout(W, 2, pref[i], V, wild).total <= static_cast<uint32_t>(SMEM_LIMIT)) chosen = pref[i];
    if (chosen < 0) return -1;
    int nb = std::max(2, W + ring_extra);
    while (nb > 2 && smem_layout(W, nb, chosen, V, wild).total > static_cast<uint32_t>(SMEM_LIMIT)) --nb;
    *TE = chosen; *NB = nb;
    *smem = smem_layout(W, nb, chosen, V, wild).total;
    return 0;
}

// ------------------------------------------------------------------ ABI: lifecycle
extern "C" int wxb_abi_version(void) { return WXB_ABI_VERSION; }
extern "C" const char *wxb_last_error(void) { return g_err.c_str(); }
extern "C" double wxb_round3(double x) { return wxb::round3(x); }
extern "C" double wxb_round3_np(double x) { return wxb::round3_np(x); }
extern "C" double wxb_pairwise_sum(const double *a, int64_t n) {
    return wxb::pairwise_sum([&](long i) { return a[i]; }, static_cast<long>(n));
}
extern "C" int wxb_tokens_per_lane(int32_t n_tokens) { return tokens_per_lane(n_tokens); }
extern "C" int wxb_segment_warps(int32_t n_tokens) {
    int W, K;
    if (n_tokens <= 0) return 0;
    return shape_of(n_tokens, &W, &K) ? W : -1;
}
extern "C" int wxb_set_k_target(int k_target) {
    const int old = g_k_target.load();
    if (k_target > 0) g_k_target.store(std::max(1, std::min(KMAX, k_target)));
    if (k_target < 0) g_k_target.store(0);       // back to the per-batch choice
    return old;
}

extern "C" int wxb_create(int device, wxb_ctx **out) {
    if (!out) return fail(WXB_EINVAL, "wxb_create: out is NULL");
    *out = nullptr;
    int count = 0;
    CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(WXB_EINVAL, "wxb_create: device %d out of range (%d visible)", device, count);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(WXB_EUNSUPPORTED, "wxb_create: device %d is sm_%d%d; this library is built for sm_100a only (no fallback)", device, prop.major, prop.minor);
    wxb_ctx *c = new (std::nothrow) wxb_ctx();
    if (!c) return fail(WXB_ENOMEM, "wxb_create: out of host memory");
    c->device = device;
    if (const char *e = getenv("WXB_SIDE_STREAMS")) c->side_streams = std::max(1, std::min(NSIDE, atoi(e)));
    if (const char *e = getenv("WXB_TRACE")) c->want_trace = atoi(e) != 0;
    if (const char *e = getenv("WXB_GRAPHS")) c->use_graphs = atoi(e) != 0;
    if (const char *e = getenv("WXB_K_TARGET")) c->k_target = std::max(1, std::min(KMAX, atoi(e)));
    if (const char *e = getenv("WXB_TILE_FRAMES")) c->tile_pref = std::max(4, std::min(32, atoi(e)));
    if (const char *e = getenv("WXB_RING_EXTRA")) c->ring_extra = std::max(1, std::min(4, atoi(e)));
    if (const char *e = getenv("WXB_CHUNK_TAPER")) c->chunk_taper = atoi(e) != 0;
    if (const char *e = getenv("WXB_CHUNK_MB")) c->chunk_bytes = static_cast<size_t>(std::max(1, atoi(e))) << 20;
    c->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&c->main, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_d2h, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->jstream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = c->slot[i].create();
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev_dbg[i]);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_t0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_t1);
    if (e != cudaSuccess) {
        wxb_destroy(c);
        return fail(WXB_ECUDA, "wxb_create: %s", cudaGetErrorString(e));
    }
    *out = c;
    return WXB_OK;
}

extern "C" void wxb_destroy(wxb_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf *bufs[] = {&c->trace, &c->spk_best, &c->h_emis, &c->h_path_tok};
    for (RunSlot &sl : c->slot) sl.destroy();
    if (c->jstream) cudaStreamDestroy(c->jstream);
    for (DevBuf *b : bufs) b->release();
    for (auto &b : c->h_i32) b.release();
    for (auto &b : c->h_f64) b.release();
    for (auto &b : c->h_f32) b.release();
    for (Arena *a : c->arena_pool) { a->release(); delete a; }
    for (cudaEvent_t ev : c->ev_chunk) cudaEventDestroy(ev);
    for (cudaEvent_t ev : c->ev_soft) cudaEventDestroy(ev);
    if (c->main) cudaStreamDestroy(c->main);
    if (c->copy) cudaStreamDestroy(c->copy);
    if (c->d2h) cudaStreamDestroy(c->d2h);
    if (c->ev_d2h) cudaEventDestroy(c->ev_d2h);
    for (int i = 0; i < 3; ++i) if (c->ev_dbg[i]) cudaEventDestroy(c->ev_dbg[i]);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    delete c;
}

extern "C" int64_t wxb_set_chunk_bytes(wxb_ctx *ctx, int64_t bytes) {
    if (!ctx) return -1;
    const int64_t old = static_cast<int64_t>(ctx->chunk_bytes);
    if (bytes >= 1) ctx->chunk_bytes = static_cast<size_t>(bytes);
    return old;
}

extern "C" int wxb_get_limits(const wxb_ctx *ctx, wxb_limits *out) {
    if (!ctx || !out) return fail(WXB_EINVAL, "wxb_get_limits: NULL argument");
    int vmax = 1;
    while (smem_layout(1, 2, 4, vmax + 1).total <= static_cast<uint32_t>(SMEM_LIMIT)) ++vmax;       // one warp, 4-frame tiles
    out->max_vocab = vmax;
    out->max_tokens = 32 * WMAX * KMAX;
    out->tokens_per_lane = KMAX;
    out->sm_count = ctx->sm_count;
    return WXB_OK;
}

// ------------------------------------------------------------------ ABI: plan
static size_t up256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

static void plan_free(wxb_plan *pl) {
    if (!pl) return;
    if (pl->graph_exec) cudaGraphExecDestroy(pl->graph_exec);
    if (pl->arena) pl->ctx->arena_pool.push_back(pl->arena);
    delete pl;
}

// Row boundaries of the host->device emission chunks of wxb_align_host: full chunks of chunk_rows rows, then a
// tapered tail (1/2, 1/4, 1/8 of a chunk).  What is left to do when the last byte has landed is the kernels and
// result copies of the last chunk's segments only, so the smaller that chunk, the closer the call ends to the
// link's own time (64 MB last chunk: ~0.45 ms of tail on the 4,096-segment benchmark batch; 8 MB: ~0.1 ms).
static std::vector<int64_t> make_chunk_bounds(int64_t rows, int64_t chunk_rows, bool taper) {
    std::vector<int64_t> b{0};
    if (rows <= 0) { b.push_back(0); return b; }
    int64_t tail[3] = {chunk_rows / 2, chunk_rows / 4, chunk_rows / 8};
    const bool use_tail = taper && chunk_rows >= 8 && rows > 2 * chunk_rows;
    const int64_t tail_total = use_tail ? tail[0] + tail[1] + tail[2] : 0;
    int64_t r = 0;
    while (rows - tail_total - r > chunk_rows) { r += chunk_rows; b.push_back(r); }
    if (!use_tail) { b.push_back(rows); return b; }
    r = rows - tail_total;
    b.push_back(r);
    for (int i = 0; i < 3; ++i) { r += tail[i]; b.push_back(r); }
    return b;
}

// bounds = nullptr: one chunk (the emissions are already resident).  Otherwise the emission matrix arrives in
// chunks with the given row boundaries (wxb_align_host starts those copies before planning) and a segment
// belongs to the chunk that holds its last row: its kernels wait for that chunk only.
static int plan_create_impl(wxb_ctx *ctx, const wxb_batch *bt, const std::vector<int64_t> *bounds, wxb_plan **out, bool copy_engine_busy = false) {
    if (!ctx || !bt || !out) return fail(WXB_EINVAL, "wxb_plan_create: NULL argument");
    *out = nullptr;
    const int64_t S = bt->n_segments;
    if (S < 0) return fail(WXB_EINVAL, "wxb_plan_create: n_segments < 0");
    if (S > 0 && (!bt->frame_off || !bt->tok_off)) return fail(WXB_EINVAL, "wxb_plan_create: frame_off/tok_off is NULL");
    if (bt->vocab < 1) return fail(WXB_EINVAL, "wxb_plan_create: vocab must be >= 1");
    if (bt->blank_id < 0 || bt->blank_id >= bt->vocab) return fail(WXB_EINVAL, "wxb_plan_create: blank_id %d outside [0,%d)", bt->blank_id, bt->vocab);
    if ((bt->seg_t1 == nullptr) != (bt->seg_ratio == nullptr)) return fail(WXB_EINVAL, "wxb_plan_create: seg_t1 and seg_ratio must be given together");
    if (bt->n_words < 0) return fail(WXB_EINVAL, "wxb_plan_create: n_words < 0");
    if (bt->n_words > 0 && (!bt->word_off || !bt->wchar_off || !bt->wchar_idx || !bt->seg_t1))
        return fail(WXB_EINVAL, "wxb_plan_create: word groups need word_off, wchar_off, wchar_idx and seg_t1/seg_ratio");

    wxb_plan *pl = new (std::nothrow) wxb_plan();
    if (!pl) return fail(WXB_ENOMEM, "wxb_plan_create: out of host memory");
    struct Guard { wxb_plan *p; ~Guard() { plan_free(p); } } guard{pl};
    pl->ctx = ctx;
    pl->S = S;
    pl->V = bt->vocab;
    pl->blank = bt->blank_id;
    pl->flags = bt->flags;
    pl->has_sec = bt->seg_t1 != nullptr;
    pl->W = bt->n_words;
    pl->has_words = bt->n_words > 0;
    pl->has_wild = (bt->flags & WXB_F_WILDCARD) != 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (S == 0) { guard.p = nullptr; *out = pl; return WXB_OK; }

    const int64_t *fo = bt->frame_off, *to = bt->tok_off;
    if (fo[0] < 0 || to[0] < 0) return fail(WXB_EINVAL, "wxb_plan_create: negative offset");
    for (int64_t s = 0; s < S; ++s) {
        if (fo[s + 1] < fo[s] || to[s + 1] < to[s]) return fail(WXB_EINVAL, "wxb_plan_create: offsets not ascending at segment %lld", (long long)s);
        if (fo[s + 1] - fo[s] > 0x3fffffff || to[s + 1] - to[s] > 0x3fffffff) return fail(WXB_EINVAL, "wxb_plan_create: segment %lld too large", (long long)s);
        if (bt->frame_len && (bt->frame_len[s] < 0 || bt->frame_len[s] > fo[s + 1] - fo[s]))
            return fail(WXB_EINVAL, "wxb_plan_create: frame_len[%lld]=%d outside [0,%lld]", (long long)s, bt->frame_len[s], (long long)(fo[s + 1] - fo[s]));
    }
    pl->sum_T = fo[S];
    pl->sum_N = to[S];
    if (pl->sum_N > 0 && !bt->tokens) return fail(WXB_EINVAL, "wxb_plan_create: tokens is NULL");
    const int64_t W = pl->W;
    if (pl->has_words) {
        if (bt->word_off[0] != 0 || bt->word_off[S] != W) return fail(WXB_EINVAL, "wxb_plan_create: word_off must run from 0 to n_words");
        for (int64_t s = 0; s < S; ++s)
            if (bt->word_off[s + 1] < bt->word_off[s]) return fail(WXB_EINVAL, "wxb_plan_create: word_off not ascending");
        for (int64_t w = 0; w < W; ++w)
            if (bt->wchar_off[w + 1] < bt->wchar_off[w]) return fail(WXB_EINVAL, "wxb_plan_create: wchar_off not ascending");
    }

    // ---- shape every segment; shortcut the provably failing ones ----
    std::vector<SegDesc> descs;
    descs.reserve(static_cast<size_t>(S));
    std::vector<int> desc_chunk;
    desc_chunk.reserve(static_cast<size_t>(S));
    std::vector<ShortSeg> shorts;
    long long trellis_off = 0;
    int wmax = 1;
    const bool no_shortcut = (bt->flags & WXB_F_NO_SHORTCUT) != 0;
    const int k_forced = ctx->k_target > 0 ? ctx->k_target : g_k_target.load(std::memory_order_relaxed);
    const int k_target = k_forced > 0 ? k_forced : pick_k_target(to, fo, bt->frame_len, S, ctx->sm_count, pl->V);
    if (bounds && bounds->size() >= 2 && bounds->front() == 0 && bounds->back() == fo[S]) pl->chunk_row = *bounds;
    else { pl->chunk_row.push_back(0); pl->chunk_row.push_back(fo[S]); }
    const int n_chunks = static_cast<int>(pl->chunk_row.size()) - 1;
    pl->chunk_seg.assign(static_cast<size_t>(n_chunks) + 1, S);
    pl->chunk_seg[0] = 0;
    int chunk_prev = 0;
    for (int64_t s = 0; s < S; ++s) {
        const int T = bt->frame_len ? bt->frame_len[s] : static_cast<int>(fo[s + 1] - fo[s]);
        const int N = static_cast<int>(to[s + 1] - to[s]);
        // chunk holding the segment's last row; last rows are non-decreasing in s, so the search only moves forward
        int chunk = chunk_prev;
        const int64_t last_row = fo[s] + std::max(T, 1) - 1;
        while (chunk + 1 < n_chunks && last_row >= pl->chunk_row[chunk + 1]) ++chunk;
        for (; chunk_prev < chunk; ++chunk_prev) pl->chunk_seg[chunk_prev + 1] = s;
        const long long this_trellis = trellis_off;
        trellis_off += static_cast<long long>(T + 1) * (N + 1);
        int K = -1, Wc = 0, KL = 0;
        const bool fits = shape_of(N, &Wc, &K, k_target, &KL);
        if (N == 0 || (N > T && (!no_shortcut || !fits))) {
            shorts.push_back({to[s], fo[s], N, T, static_cast<int>(s), 0});
            continue;
        }
        if (!fits)
            return fail(WXB_EUNSUPPORTED, "wxb_plan_create: segment %lld has %d tokens; this build aligns at most %d per segment", (long long)s, N, 32 * WMAX * KMAX);
        SegDesc d;
        d.emis_off = fo[s] * static_cast<long long>(pl->V);
        d.frame_base = fo[s];
        d.tok_base = to[s];
        d.bits_off = 0;
        d.trellis_off = this_trellis;
        d.T = T; d.N = N; d.seg = static_cast<int>(s); d.K = K; d.W = Wc; d.pad = chunk; d.KL = KL; d.pad2 = 0;
        wmax = std::max(wmax, Wc);
        descs.push_back(d);
        pl->cells += static_cast<long long>(T) * N;
        pl->alg_bytes += 4ll * T * pl->V + 20ll * N;
    }
    pl->trellis_floats = trellis_off;
    pl->n_short = static_cast<int>(shorts.size());
    {
        int te, nbq; size_t sm;
        if (pick_shape(ctx->tile_pref, ctx->ring_extra, wmax, pl->V, pl->has_wild, &te, &nbq, &sm) != 0)
            return fail(WXB_EUNSUPPORTED, "wxb_plan_create: vocab %d needs more than %d bytes of shared memory for a %d-warp segment", pl->V, SMEM_LIMIT, wmax);
    }
    // one launch per (chunk, warps, tokens-per-lane) class; longest-first inside a class: the block
    // scheduler hands out CTAs in index order (LPT)
    std::stable_sort(descs.begin(), descs.end(), [](const SegDesc &a, const SegDesc &b) {
        if (a.pad != b.pad) return a.pad < b.pad;
        if (a.W != b.W) return a.W > b.W;
        if (a.K != b.K) return a.K > b.K;
        return a.T > b.T;
    });
    long long bits = 0;
    int fin_next = 0;               // finalize work items (128 tokens each) are laid out in descriptor order
    for (size_t i = 0; i < descs.size(); ++i) {
        descs[i].bits_off = bits;
        bits += static_cast<long long>((descs[i].T + 31) / 32) * descs[i].W * descs[i].K * 32;
        if (i == 0 || descs[i].K != descs[i - 1].K || descs[i].W != descs[i - 1].W || descs[i].pad != descs[i - 1].pad) {
            Bucket b{descs[i].K, descs[i].W, 0, 0, 0, static_cast<int>(i), 0, 0, descs[i].T, descs[i].pad, fin_next, 0};
            pick_shape(ctx->tile_pref, ctx->ring_extra, b.W, pl->V, pl->has_wild, &b.TE, &b.NB, &b.smem);
            pl->buckets.push_back(b);
        }
        pl->buckets.back().count += 1;
        pl->buckets.back().cost += static_cast<long long>(descs[i].T) * (descs[i].K * (descs[i].W - 1) + descs[i].KL);
        pl->buckets.back().fin_count += (descs[i].N + 127) / 128;
        fin_next += (descs[i].N + 127) / 128;
    }
    pl->bits_words = bits;
    // launch order = placement priority: inside a chunk, classes holding the longest segments go
    // first (LPT across classes; a segment's latency grows with its frame count)
    std::stable_sort(pl->buckets.begin(), pl->buckets.end(), [](const Bucket &a, const Bucket &b) {
        if (a.chunk != b.chunk) return a.chunk < b.chunk;
        if (a.maxT != b.maxT) return a.maxT > b.maxT;
        return a.cost > b.cost;
    });
    size_t n_fin = 0;
    for (const SegDesc &d : descs) n_fin += static_cast<size_t>((d.N + 127) / 128);
    pl->n_fin_blocks = static_cast<int>(n_fin);
    pl->n_desc = static_cast<int>(descs.size());
    if (ctx->want_trace) pl->host_descs = descs;

    // ---- one arena: lay out, fill the pinned mirror, upload once ----
    const size_t sN = static_cast<size_t>(pl->sum_N), nW = static_cast<size_t>(W);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += up256(std::max<size_t>(bytes, 1)); return o; };
    const size_t o_segs = take(descs.size() * sizeof(SegDesc));
    const size_t o_fin = take(n_fin * sizeof(FinBlock));
    const size_t o_tok4 = take(sN * 4);
    const size_t o_short = take(shorts.size() * sizeof(ShortSeg));
    const size_t o_t1 = take(pl->has_sec ? static_cast<size_t>(S) * 8 : 0);
    const size_t o_ratio = take(pl->has_sec ? static_cast<size_t>(S) * 8 : 0);
    pl->has_round_np = pl->has_sec && bt->seg_round_np != nullptr;
    const size_t o_rnp = take(pl->has_round_np ? static_cast<size_t>(S) : 0);
    const size_t o_wseg = take(pl->has_words ? nW * 4 : 0);
    const size_t o_stb = take(pl->has_words ? static_cast<size_t>(S) * 8 : 0);
    const size_t o_woff = take(pl->has_words ? (nW + 1) * 8 : 0);
    const size_t n_wc = pl->has_words ? static_cast<size_t>(bt->wchar_off[W]) : 0;
    const size_t o_widx = take(n_wc * 4);

    Arena *ar = nullptr;
    if (!ctx->arena_pool.empty()) { ar = ctx->arena_pool.back(); ctx->arena_pool.pop_back(); }
    else ar = new (std::nothrow) Arena();
    if (!ar) return fail(WXB_ENOMEM, "wxb_plan_create: out of host memory");
    pl->arena = ar;
    CUDA_TRY(ar->ensure(off));
    unsigned char *h = ar->host;
    memcpy(h + o_segs, descs.data(), descs.size() * sizeof(SegDesc));
    {
        FinBlock *fin = reinterpret_cast<FinBlock *>(h + o_fin);
        size_t k = 0;
        for (size_t i = 0; i < descs.size(); ++i)
            for (int k0 = 0; k0 < descs[i].N; k0 += 128) fin[k++] = {static_cast<int>(i), k0};
    }
    {
        int *tok4 = reinterpret_cast<int *>(h + o_tok4);
        const int32_t *tk = bt->tokens;
        const uint32_t V = static_cast<uint32_t>(pl->V);
        const uint32_t lo = pl->has_wild ? 1u : 0u;          // with wildcards, id -1 is legal: check tk + 1 in [0, V + 1)
        uint32_t bad = 0;
        for (size_t i = 0; i < sN; ++i) {
            bad |= (static_cast<uint32_t>(tk[i]) + lo >= V + lo);   // other negative ids wrap to large unsigned values
            tok4[i] = tk[i] * 4;                                     // wildcard -> -4: "no row offset"
        }
        if (bad)
            for (size_t i = 0; i < sN; ++i)
                if (tk[i] < -static_cast<int>(lo) || tk[i] >= pl->V) return fail(WXB_EINVAL, "wxb_plan_create: tokens[%lld]=%d outside [%d,%d)", (long long)i, tk[i], -static_cast<int>(lo), pl->V);
    }
    if (!shorts.empty()) memcpy(h + o_short, shorts.data(), shorts.size() * sizeof(ShortSeg));
    if (pl->has_sec) {
        memcpy(h + o_t1, bt->seg_t1, static_cast<size_t>(S) * 8);
        memcpy(h + o_ratio, bt->seg_ratio, static_cast<size_t>(S) * 8);
        if (pl->has_round_np) memcpy(h + o_rnp, bt->seg_round_np, static_cast<size_t>(S));
    }
    if (pl->has_words) {
        int *word_seg = reinterpret_cast<int *>(h + o_wseg);
        long long *stb = reinterpret_cast<long long *>(h + o_stb);
        for (int64_t s = 0; s < S; ++s) {
            stb[s] = to[s];
            const int N = static_cast<int>(to[s + 1] - to[s]);
            for (int64_t w = bt->word_off[s]; w < bt->word_off[s + 1]; ++w) {
                word_seg[w] = static_cast<int>(s);
                for (int64_t i = bt->wchar_off[w]; i < bt->wchar_off[w + 1]; ++i)
                    if (bt->wchar_idx[i] < -1 || bt->wchar_idx[i] >= N) return fail(WXB_EINVAL, "wxb_plan_create: wchar_idx[%lld]=%d outside [-1,%d)", (long long)i, bt->wchar_idx[i], N);
            }
        }
        memcpy(h + o_woff, bt->wchar_off, (nW + 1) * 8);
        if (n_wc) memcpy(h + o_widx, bt->wchar_idx, n_wc * 4);
    }
    if (copy_engine_busy) {
        const size_t n16 = (off + 15) / 16;               // the arena is padded to 256 bytes
        upload_kernel<<<static_cast<unsigned>(std::min<size_t>((n16 + 255) / 256, 4 * static_cast<size_t>(ctx->sm_count))), 256, 0, ctx->main>>>(
            reinterpret_cast<uint4 *>(ar->dev), reinterpret_cast<const uint4 *>(ar->host_dev), n16);
        CUDA_TRY(cudaGetLastError());
    } else {
        CUDA_TRY(cudaMemcpyAsync(ar->dev, ar->host, off, cudaMemcpyHostToDevice, ctx->main));
    }
    CUDA_TRY(cudaEventRecord(ar->uploaded, ctx->main));
    unsigned char *dv = ar->dev;
    pl->d_segs = reinterpret_cast<SegDesc *>(dv + o_segs);
    pl->d_fin = reinterpret_cast<FinBlock *>(dv + o_fin);
    pl->d_tok4 = reinterpret_cast<int *>(dv + o_tok4);
    pl->d_short = reinterpret_cast<ShortSeg *>(dv + o_short);
    pl->d_t1 = reinterpret_cast<double *>(dv + o_t1);
    pl->d_ratio = reinterpret_cast<double *>(dv + o_ratio);
    pl->d_round_np = dv + o_rnp;
    pl->d_word_seg = reinterpret_cast<int *>(dv + o_wseg);
    pl->d_seg_tok_base = reinterpret_cast<long long *>(dv + o_stb);
    pl->d_wchar_off = reinterpret_cast<long long *>(dv + o_woff);
    pl->d_wchar_idx = reinterpret_cast<int *>(dv + o_widx);
    guard.p = nullptr;
    *out = pl;
    return WXB_OK;
}

extern "C" int wxb_plan_create(wxb_ctx *ctx, const wxb_batch *bt, wxb_plan **out) {
    return plan_create_impl(ctx, bt, nullptr, out);
}

extern "C" void wxb_plan_destroy(wxb_plan *plan) {
    if (!plan) return;
    cudaSetDevice(plan->ctx->device);
    // the arena goes back to the pool: nothing may still be reading it
    if (plan->arena && plan->arena->uploaded) cudaEventSynchronize(plan->arena->uploaded);
    if (plan->ran) {
        cudaStreamSynchronize(plan->last_stream);
        for (RunSlot &sl : plan->ctx->slot) cudaEventSynchronize(sl.done);       // overlapped runs end on the join stream
    }
    plan_free(plan);
}

extern "C" int wxb_plan_launch_count(const wxb_plan *plan) {
    if (!plan) return 0;
    int n = (plan->n_short > 0 ? 1 : 0) + (plan->has_words ? 1 : 0);
    for (const Bucket &b : plan->buckets) n += 2 + (b.fin_count > 0 ? 1 : 0);      // align_kernel + backtrack_kernel + finalize_kernel
    return n;
}

extern "C" int wxb_plan_stats(const wxb_plan *plan, int64_t *cells, int64_t *alg_bytes, int64_t *n_shortcut) {
    if (!plan) return fail(WXB_EINVAL, "wxb_plan_stats: plan is NULL");
    if (cells) *cells = plan->cells;
    if (alg_bytes) *alg_bytes = plan->alg_bytes;
    if (n_shortcut) *n_shortcut = plan->n_short;
    return WXB_OK;
}

// torch.log_softmax over `rows` rows of V floats (alignment.py:233); in == out is allowed.
static int launch_log_softmax(const float *in, float *out, int64_t rows, int V, cudaStream_t st) {
    const size_t row_bytes = static_cast<size_t>(V) * 4;
    if (row_bytes > static_cast<size_t>(SMEM_LIMIT)) return fail(WXB_EUNSUPPORTED, "log_softmax: vocab %d too large", V);
    // shapes: see log_softmax_kernel
    if (V > wxb::WIDE_V && V <= 8192) {
        const int rpc = V <= 1024 ? 8 : 1;
        const long long need = (rows + rpc - 1) / rpc;
        const unsigned blocks = static_cast<unsigned>(std::min<long long>(need, 148ll * 8 * 16));
        if (V <= 128) log_softmax_rows_kernel<4, 1><<<blocks, 256, 0, st>>>(in, out, rows, V);
        else if (V <= 256) log_softmax_rows_kernel<8, 1><<<blocks, 256, 0, st>>>(in, out, rows, V);
        else if (V <= 512) log_softmax_rows_kernel<16, 1><<<blocks, 256, 0, st>>>(in, out, rows, V);
        else if (V <= 1024) log_softmax_rows_kernel<32, 1><<<blocks, 256, 0, st>>>(in, out, rows, V);
        else if (V <= 2048) log_softmax_rows_kernel<8, 8><<<blocks, 256, 0, st>>>(in, out, rows, V);
        else if (V <= 4096) log_softmax_rows_kernel<16, 8><<<blocks, 256, 0, st>>>(in, out, rows, V);
        else log_softmax_rows_kernel<32, 8><<<blocks, 256, 0, st>>>(in, out, rows, V);
        CUDA_TRY(cudaGetLastError());
        return WXB_OK;
    }
    int rpc, group, pad;
    if (V <= wxb::WIDE_V) { group = 1; pad = (V & 1) ? 0 : 1; }
    else { group = 32; pad = 0; }
    const size_t staged = static_cast<size_t>(V + pad) * 4;
    if (group == 32) {
        rpc = static_cast<int>(std::min<size_t>(128, (40 * 1024) / staged));
        if (rpc >= 4) rpc &= ~3;
        if (rpc < 4) rpc = 4;                               // one row per warp of the CTA
        while (rpc > 1 && static_cast<size_t>(rpc) * staged > static_cast<size_t>(SMEM_LIMIT)) --rpc;
    } else {
        rpc = static_cast<int>(std::min<size_t>(128, (72 * 1024) / staged));
        rpc = std::max(rpc & ~3, 4);
    }
    const size_t smem = static_cast<size_t>(rpc) * staged;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(log_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    // i / V for i < 2^14 as one multiply-high (exact: the error i / 2^32 is far below 1 / V)
    const unsigned magic = V > 1 ? static_cast<unsigned>((0x100000000ull + static_cast<unsigned>(V) - 1) / static_cast<unsigned>(V)) : 0u;
    const long long blocks = (rows + rpc - 1) / rpc;
    log_softmax_kernel<<<static_cast<unsigned>(blocks), 128, smem, st>>>(in, out, rows, V, rpc, group, pad, magic);
    CUDA_TRY(cudaGetLastError());
    return WXB_OK;
}

// ------------------------------------------------------------------ ABI: run
// chunk_ready != nullptr: pipelined host path — chunk c of the emissions is on its way to `emis`
// on the copy stream and chunk_ready[c] fires when it has landed; the kernels of a segment wait
// for the chunk that holds its last row.
// With a sink (wxb_align_host), the results of a chunk's segments go back to the host as soon as that
// chunk's kernels are done, on their own stream: the device->host copies overlap the rest of the
// host->device emission stream (PCIe is full duplex).
struct HostSink {
    const wxb_outputs *host;        // caller's buffers
    const int64_t *tok_off, *word_off, *frame_off;
};

// slot_i: which RunSlot (scratch + side streams) the run uses.  overlapped: do not make `st` wait for the run —
// the side streams are joined on the ctx's join stream instead, where the slot's `done` event is recorded
// (wxb_plan_run_async / wxb_ctx_join); every class then goes to a side stream, also when there is only one.
static int run_impl(wxb_plan *pl, const float *emis, const cudaEvent_t *chunk_ready, const wxb_outputs *o, cudaStream_t st,
                    float *elapsed_dp_ms, const HostSink *sink = nullptr, int slot_i = 0, bool overlapped = false) {
    if (!pl || !o) return fail(WXB_EINVAL, "wxb_plan_run: NULL argument");
    if (pl->S == 0) { if (elapsed_dp_ms) *elapsed_dp_ms = 0.f; return WXB_OK; }
    if (!o->status || !o->char_start || !o->char_end || !o->char_score)
        return fail(WXB_EINVAL, "wxb_plan_run: status, char_start, char_end and char_score are required outputs");
    if (!emis && pl->sum_T > 0) return fail(WXB_EINVAL, "wxb_plan_run: emissions is NULL");
    if ((o->path_token == nullptr) != (o->path_logp == nullptr)) return fail(WXB_EINVAL, "wxb_plan_run: path_token and path_logp go together");
    const bool want_sec = o->char_start_s || o->char_end_s || o->char_score_r;
    if (want_sec && !(o->char_start_s && o->char_end_s && o->char_score_r && pl->has_sec))
        return fail(WXB_EINVAL, "wxb_plan_run: second-valued outputs need all three char_*_s/_r buffers and seg_t1/seg_ratio in the batch");
    const bool want_words = o->word_start_s || o->word_end_s || o->word_score;
    if (want_words && !(o->word_start_s && o->word_end_s && o->word_score && pl->has_words && want_sec))
        return fail(WXB_EINVAL, "wxb_plan_run: word outputs need all three word buffers, word groups in the batch and the char_*_s outputs");
    const bool logits = (pl->flags & WXB_F_LOGITS) != 0;
    if (o->logprobs && !logits) return fail(WXB_EINVAL, "wxb_plan_run: logprobs output is only produced with WXB_F_LOGITS");
    const bool debug = o->trellis || o->final_col;
    // Logits are normalised inside align_kernel (producer warp, one row per lane) only for small odd
    // vocabularies.  An even V puts the 32 rows of a tile in the same banks (V = 32: the fused C2 batch
    // takes 1.07 ms instead of 0.33), and one warp cannot normalise thousands of symbols per frame at the
    // rate the recurrence consumes them (7.1 ms fused vs 0.38 + 0.27 ms split at V = 2,304).  Those batches
    // go through log_softmax_kernel first — into the caller's logprobs buffer if there is one, in place
    // when the emissions sit in our own staging buffer (host path), else into scratch — and the
    // recurrence then runs on log-probs.  Up to V = 64 the arithmetic (and so every bit) is the same.
    const bool wide_logits = logits && (pl->V > wxb::WIDE_V || (pl->V & 1) == 0);

    wxb_ctx *c = pl->ctx;
    RunSlot &sl = c->slot[slot_i];
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(sl.bits.ensure(static_cast<size_t>(std::max<long long>(pl->bits_words, 1)) * 4));
    CUDA_TRY(sl.p0.ensure(static_cast<size_t>(std::max<int64_t>(pl->sum_T, 1)) * 4));
    if (pl->has_wild) CUDA_TRY(sl.wild.ensure(static_cast<size_t>(std::max<int64_t>(pl->sum_T, 1)) * 4));
    float *wide_dst = nullptr;
    if (wide_logits) {
        if (o->logprobs) wide_dst = o->logprobs;
        else if (chunk_ready || (pl->flags & WXB_F_CLOBBER)) wide_dst = const_cast<float *>(emis);   // c->h_emis (ours), or the caller's scratch
        else {
            CUDA_TRY(sl.wide_lp.ensure(static_cast<size_t>(std::max<int64_t>(pl->sum_T, 1)) * pl->V * 4));
            wide_dst = sl.wide_lp.as<float>();
        }
    }
    if (logits && !wide_logits) {
        CUDA_TRY(sl.rowmax.ensure(static_cast<size_t>(std::max<int64_t>(pl->sum_T, 1)) * 4));
        CUDA_TRY(sl.rowlse.ensure(static_cast<size_t>(std::max<int64_t>(pl->sum_T, 1)) * 4));
    }
    int *nframes = o->n_frames;
    if (!nframes) {
        CUDA_TRY(sl.nframes_scratch.ensure(static_cast<size_t>(pl->S) * 4));
        nframes = sl.nframes_scratch.as<int>();
    }

    Params P;
    memset(&P, 0, sizeof P);
    P.emis = emis;
    P.tok4 = pl->d_tok4;
    P.segs = pl->d_segs;
    P.bits = sl.bits.as<uint32_t>();
    P.p0 = sl.p0.as<float>();
    P.rowmax = sl.rowmax.as<float>();
    P.rowlse = sl.rowlse.as<float>();
    P.wild = sl.wild.as<float>();
    P.V = pl->V; P.blank = pl->blank; P.is_logits = (logits && !wide_logits) ? 1 : 0;
    if (wide_logits) P.emis = wide_dst;
    P.status = o->status; P.nframes = nframes;
    P.c_start = o->char_start; P.c_end = o->char_end; P.c_score = o->char_score;
    if (want_sec) {
        P.c_start_s = o->char_start_s; P.c_end_s = o->char_end_s; P.c_score_r = o->char_score_r;
        P.seg_t1 = pl->d_t1; P.seg_ratio = pl->d_ratio;
        P.seg_round_np = pl->has_round_np ? pl->d_round_np : nullptr;
    }
    P.path_tok = o->path_token; P.path_logp = o->path_logp;
    P.final_col = o->final_col; P.trellis = o->trellis; P.logprobs = wide_logits ? nullptr : o->logprobs;
    if (c->want_trace) {
        CUDA_TRY(c->trace.ensure(static_cast<size_t>(std::max(pl->n_desc, 1)) * 32 + 64));
        P.trace = c->trace.as<unsigned long long>();
        P.trace_hist = P.trace + 4ll * std::max(pl->n_desc, 1);
        CUDA_TRY(cudaMemsetAsync(P.trace_hist, 0, 64, st));
    }

    pl->last_stream = st;
    pl->ran = true;
    {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        CUDA_TRY(cudaStreamIsCapturing(st, &cap));
        // the plan's metadata upload (ctx stream); while capturing, the caller has already waited for it on the host
        if (cap == cudaStreamCaptureStatusNone) CUDA_TRY(cudaStreamWaitEvent(st, pl->arena->uploaded, 0));
    }
    if (elapsed_dp_ms) CUDA_TRY(cudaEventRecord(c->ev_t0, st));
    const int nb = static_cast<int>(pl->buckets.size());
    const cudaEvent_t *rows_ready = chunk_ready;        // what a class waits for before it reads its chunk
    if (wide_logits && pl->sum_T > 0) {
        if (!chunk_ready) {
            int rc = launch_log_softmax(emis, wide_dst, pl->sum_T, pl->V, st);
            if (rc != WXB_OK) return rc;
        } else {
            // host path: normalise each chunk on the main stream as it lands; the classes wait for that
            const int n_chunks = static_cast<int>(pl->chunk_seg.size()) - 1;
            while (static_cast<int>(c->ev_soft.size()) < n_chunks) {
                cudaEvent_t ev;
                CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                c->ev_soft.push_back(ev);
            }
            for (int k = 0; k < n_chunks; ++k) {
                const int64_t r0 = pl->chunk_row[k], r1 = pl->chunk_row[k + 1];
                CUDA_TRY(cudaStreamWaitEvent(st, chunk_ready[k], 0));
                if (r1 > r0) {
                    int rc = launch_log_softmax(emis + r0 * pl->V, wide_dst + r0 * pl->V, r1 - r0, pl->V, st);
                    if (rc != WXB_OK) return rc;
                }
                CUDA_TRY(cudaEventRecord(c->ev_soft[k], st));
            }
            rows_ready = c->ev_soft.data();
        }
    }
    // every class's backtrack_kernel and finalize_kernel follow its align_kernel on the same stream.  The block
    // scheduler places CTAs in launch order, so the small walk kernels of the early classes queue behind the later
    // align_kernel launches and run near the end of the step (timeline, profiles/r2_timeline.md); moving them to
    // highest-priority companion streams made them start as soon as their class was done but moved the end of the
    // step by nothing (0.543 vs 0.516 ms: the walks cost their issue slots wherever they run) — not kept.
    auto finalize = [&](const Bucket &b, cudaStream_t s) -> cudaError_t {
        backtrack_kernel<<<(b.count + 3) / 4, 128, 0, s>>>(P, b.first, b.count);
        if (b.fin_count > 0) finalize_kernel<<<b.fin_count, 128, 0, s>>>(P, pl->d_fin + b.fin_first);
        return cudaGetLastError();
    };
    if (pl->n_short > 0) {
        mark_failed_kernel<<<pl->n_short, 64, 0, st>>>(pl->d_short, pl->n_short, o->status, nframes,
                                                        o->char_start, o->char_end, o->char_score,
                                                        want_sec ? o->char_start_s : nullptr, o->char_end_s, o->char_score_r, o->path_token, o->path_logp);
        CUDA_TRY(cudaGetLastError());
    }
    auto reduce_words = [&](int64_t w0, int64_t w1, cudaStream_t s) -> cudaError_t {
        if (!want_words || w1 <= w0) return cudaSuccess;
        const int threads = 128;
        const long long blocks = (w1 - w0 + threads - 1) / threads;
        word_reduce_kernel<<<static_cast<unsigned>(blocks), threads, 0, s>>>(
            w1 - w0, pl->d_word_seg + w0, pl->d_seg_tok_base, pl->d_wchar_off + w0, pl->d_wchar_idx, o->status, o->char_start_s,
            o->char_end_s, o->char_score_r, o->word_start_s + w0, o->word_end_s + w0, o->word_score + w0);
        return cudaGetLastError();
    };
    if (nb == 1 && !chunk_ready && !overlapped) {
        const Bucket &b = pl->buckets[0];
        CUDA_TRY(launch_align(b.K, b.W, b.NB, b.TE, P, b.first, b.count, b.smem, st, debug, pl->has_wild));
        CUDA_TRY(finalize(b, st));
    } else if (nb >= 1) {
        // classes are independent: fork onto side streams so that every SM sees a mix of them
        const int used = std::min(nb, c->side_streams);
        CUDA_TRY(cudaEventRecord(sl.ev_fork, st));
        for (int i = 0; i < used; ++i) CUDA_TRY(cudaStreamWaitEvent(sl.side[i], sl.ev_fork, 0));
        if (sink) CUDA_TRY(cudaStreamWaitEvent(c->d2h, sl.ev_fork, 0));
        int seen_chunk[NSIDE];
        for (int i = 0; i < NSIDE; ++i) seen_chunk[i] = -1;
        auto drain_chunk = [&](int chunk) -> int {      // results of the chunk's segments -> host, on the d2h stream
            for (int i = 0; i < used; ++i) {
                CUDA_TRY(cudaEventRecord(sl.ev_join[i], sl.side[i]));
                CUDA_TRY(cudaStreamWaitEvent(c->d2h, sl.ev_join[i], 0));
            }
            const int64_t s0 = pl->chunk_seg[chunk], s1 = pl->chunk_seg[chunk + 1];
            if (s1 <= s0) return WXB_OK;
            const int64_t k0 = sink->tok_off[s0], k1 = sink->tok_off[s1];
            const int64_t w0 = pl->has_words ? sink->word_off[s0] : 0, w1 = pl->has_words ? sink->word_off[s1] : 0;
            const int64_t f0 = sink->frame_off[s0], f1 = sink->frame_off[s1];
            CUDA_TRY(reduce_words(w0, w1, c->d2h));
            const wxb_outputs *h = sink->host;
#define WXB_D2H(field, lo, hi, type)                                                                          \
            if (h->field && o->field && (hi) > (lo))                                                          \
                CUDA_TRY(cudaMemcpyAsync(h->field + (lo), o->field + (lo), static_cast<size_t>((hi) - (lo)) * sizeof(type), \
                                         cudaMemcpyDeviceToHost, c->d2h));
            WXB_D2H(status, s0, s1, int32_t) WXB_D2H(n_frames, s0, s1, int32_t)
            WXB_D2H(char_start, k0, k1, int32_t) WXB_D2H(char_end, k0, k1, int32_t) WXB_D2H(char_score, k0, k1, double)
            WXB_D2H(char_start_s, k0, k1, double) WXB_D2H(char_end_s, k0, k1, double) WXB_D2H(char_score_r, k0, k1, double)
            WXB_D2H(word_start_s, w0, w1, double) WXB_D2H(word_end_s, w0, w1, double) WXB_D2H(word_score, w0, w1, double)
            WXB_D2H(path_token, f0, f1, int32_t) WXB_D2H(path_logp, f0, f1, float)
#undef WXB_D2H
            return WXB_OK;
        };
        int cur_chunk = nb > 0 ? pl->buckets[0].chunk : 0;
        for (int i = 0; i < nb; ++i) {
            const Bucket &b = pl->buckets[i];
            const int si = i % used;
            if (sink && b.chunk != cur_chunk) {
                for (int ch = cur_chunk; ch < b.chunk; ++ch) { int rc = drain_chunk(ch); if (rc != WXB_OK) return rc; }
                cur_chunk = b.chunk;
            }
            if (rows_ready && seen_chunk[si] < b.chunk) {
                CUDA_TRY(cudaStreamWaitEvent(sl.side[si], rows_ready[b.chunk], 0));
                seen_chunk[si] = b.chunk;
            }
            CUDA_TRY(launch_align(b.K, b.W, b.NB, b.TE, P, b.first, b.count, b.smem, sl.side[si], debug, pl->has_wild));
            CUDA_TRY(finalize(b, sl.side[si]));
        }
        if (sink) {
            const int n_chunks = static_cast<int>(pl->chunk_seg.size()) - 1;
            for (int ch = cur_chunk; ch < n_chunks; ++ch) { int rc = drain_chunk(ch); if (rc != WXB_OK) return rc; }
            CUDA_TRY(cudaEventRecord(c->ev_dbg[1], c->d2h));
            CUDA_TRY(cudaEventRecord(c->ev_d2h, c->d2h));
        }
        cudaStream_t joiner = overlapped ? c->jstream : st;
        for (int i = 0; i < used; ++i) {
            CUDA_TRY(cudaEventRecord(sl.ev_join[i], sl.side[i]));
            CUDA_TRY(cudaStreamWaitEvent(joiner, sl.ev_join[i], 0));
        }
        if (sink) {
            CUDA_TRY(cudaEventRecord(c->ev_dbg[0], st));          // all kernels done
            CUDA_TRY(cudaStreamWaitEvent(st, c->ev_d2h, 0));
        }
    }
    if (elapsed_dp_ms) CUDA_TRY(cudaEventRecord(c->ev_t1, st));
    if (overlapped) {
        // the join stream has waited for every class; order it after this call's own work on `st` too (the records of
        // un-alignable segments, a plan without classes), reduce the words there and publish the slot's `done`
        if (nb == 0) { CUDA_TRY(cudaEventRecord(sl.ev_fork, st)); }
        CUDA_TRY(cudaStreamWaitEvent(c->jstream, sl.ev_fork, 0));
        CUDA_TRY(reduce_words(0, pl->W, c->jstream));
        CUDA_TRY(cudaEventRecord(sl.done, c->jstream));
        sl.done_on = c->jstream;
        return WXB_OK;
    }
    if (!sink) CUDA_TRY(reduce_words(0, pl->W, st));
    if (elapsed_dp_ms) {
        CUDA_TRY(cudaEventSynchronize(c->ev_t1));
        CUDA_TRY(cudaEventElapsedTime(elapsed_dp_ms, c->ev_t0, c->ev_t1));
    }
    return WXB_OK;
}

static int plan_run_unserialised(wxb_plan *pl, const float *emis, const wxb_outputs *o, cudaStream_t st, float *elapsed_dp_ms);

// The scratch of a RunSlot is shared by every plan that runs in it: a run first makes its stream wait (on the
// device) for the slot's previous run, whatever stream that was enqueued on, and publishes its own end in the
// slot's `done` event.  wxb_plan_run always uses slot 0 (a no-op wait when the caller keeps to one stream).
extern "C" int wxb_plan_run(wxb_plan *pl, const float *emis, const wxb_outputs *o, void *stream_, float *elapsed_dp_ms) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (!pl || !o) return fail(WXB_EINVAL, "wxb_plan_run: NULL argument");
    wxb_ctx *c = pl->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CUDA_TRY(cudaStreamIsCapturing(st, &cap));
    const bool serialise = cap == cudaStreamCaptureStatusNone && pl->S > 0;
    RunSlot &sl = c->slot[0];
    if (serialise && sl.done_on != st) CUDA_TRY(cudaStreamWaitEvent(st, sl.done, 0));
    const int rc = plan_run_unserialised(pl, emis, o, st, elapsed_dp_ms);
    if (serialise && rc == WXB_OK) {
        CUDA_TRY(cudaEventRecord(sl.done, st));
        sl.done_on = st;
    }
    return rc;
}

// Overlapped run: same work as wxb_plan_run, but `stream` is not made to wait for it.  Consecutive calls
// alternate between the ctx's two RunSlots (own scratch, own side streams), so at most two runs are in flight:
// call n+2 waits, on the device, for call n.  The tail of a batch — its last, longest segments, when most SMs
// have nothing left to do — is filled by the head of the next one.  Results are complete once wxb_ctx_join()'s
// stream has passed the join; the caller must not reuse output buffers of a run that may still be in flight.
extern "C" int wxb_plan_run_async(wxb_plan *pl, const float *emis, const wxb_outputs *o, void *stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (!pl || !o) return fail(WXB_EINVAL, "wxb_plan_run_async: NULL argument");
    if (pl->S == 0) return WXB_OK;
    wxb_ctx *c = pl->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CUDA_TRY(cudaStreamIsCapturing(st, &cap));
    if (cap != cudaStreamCaptureStatusNone) return fail(WXB_EUNSUPPORTED, "wxb_plan_run_async: the stream is being captured");
    if (o->trellis || o->final_col || c->want_trace) return fail(WXB_EUNSUPPORTED, "wxb_plan_run_async: debug outputs need wxb_plan_run");
    const int si = c->next_async;
    c->next_async ^= 1;
    RunSlot &sl = c->slot[si];
    CUDA_TRY(cudaStreamWaitEvent(st, sl.done, 0));          // the run that used this slot two calls ago
    const int rc = run_impl(pl, emis, nullptr, o, st, nullptr, nullptr, si, /*overlapped=*/true);
    if (rc == WXB_OK) pl->last_stream = c->jstream;
    return rc;
}

extern "C" int wxb_ctx_join(wxb_ctx *c, void *stream_) {
    if (!c) return fail(WXB_EINVAL, "wxb_ctx_join: ctx is NULL");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    for (RunSlot &sl : c->slot)
        if (sl.done_on != st) CUDA_TRY(cudaStreamWaitEvent(st, sl.done, 0));
    return WXB_OK;
}

// A plan that is run again and again on the same buffers (steady-state serving, bench.py) is a fixed
// DAG of ~25 launches over 8 side streams joined by events: from the third identical call on it is
// replayed as one CUDA graph, which removes the per-step event/launch latencies between steps.
static int plan_run_unserialised(wxb_plan *pl, const float *emis, const wxb_outputs *o, cudaStream_t st, float *elapsed_dp_ms) {
    wxb_ctx *c = pl->ctx;
    if (elapsed_dp_ms || !c->use_graphs || c->want_trace || pl->graph_off || pl->S == 0 || (pl->flags & WXB_F_CLOBBER))
        return run_impl(pl, emis, nullptr, o, st, elapsed_dp_ms);
    const bool same = pl->graph_emis == emis && memcmp(&pl->graph_out, o, sizeof(wxb_outputs)) == 0;
    if (!same) {
        if (pl->graph_exec) { cudaGraphExecDestroy(pl->graph_exec); pl->graph_exec = nullptr; }
        pl->graph_emis = emis;
        pl->graph_out = *o;
        pl->graph_same = 0;
    }
    ++pl->graph_same;
    if (pl->graph_same < 3) return run_impl(pl, emis, nullptr, o, st, nullptr);      // eager: allocates scratch, configures kernels
    CUDA_TRY(cudaSetDevice(c->device));
    const RunSlot &s0 = c->slot[0];
    void *scratch[7] = {s0.bits.p, s0.p0.p, s0.rowmax.p, s0.rowlse.p, s0.wild.p, s0.nframes_scratch.p, s0.wide_lp.p};
    if (pl->graph_exec && memcmp(scratch, pl->graph_scratch, sizeof scratch) != 0) {
        cudaGraphExecDestroy(pl->graph_exec);     // a larger plan regrew the shared scratch since the capture
        pl->graph_exec = nullptr;
        return run_impl(pl, emis, nullptr, o, st, nullptr);     // re-establishes the scratch; captured again next call
    }
    if (!pl->graph_exec) {
        CUDA_TRY(cudaEventSynchronize(pl->arena->uploaded));
        CUDA_TRY(cudaStreamSynchronize(c->main));
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture(c->main, cudaStreamCaptureModeThreadLocal);
        int rc = WXB_OK;
        if (e == cudaSuccess) {
            rc = run_impl(pl, emis, nullptr, o, c->main, nullptr);
            e = cudaStreamEndCapture(c->main, &graph);
        }
        if (e == cudaSuccess && rc == WXB_OK) e = cudaGraphInstantiate(&pl->graph_exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess || rc != WXB_OK) {
            cudaGetLastError();               // clear the sticky capture error; this plan stays eager
            pl->graph_exec = nullptr;
            pl->graph_off = true;
            return run_impl(pl, emis, nullptr, o, st, nullptr);
        }
        void *now[7] = {s0.bits.p, s0.p0.p, s0.rowmax.p, s0.rowlse.p, s0.wild.p, s0.nframes_scratch.p, s0.wide_lp.p};
        memcpy(pl->graph_scratch, now, sizeof now);
    }
    pl->last_stream = st;
    pl->ran = true;
    CUDA_TRY(cudaGraphLaunch(pl->graph_exec, st));
    return WXB_OK;
}

extern "C" int wxb_align_device(wxb_ctx *ctx, const wxb_batch *batch, const float *emissions_dev,
                                const wxb_outputs *out_dev, void *stream) {
    wxb_plan *pl = nullptr;
    int rc = wxb_plan_create(ctx, batch, &pl);
    if (rc != WXB_OK) return rc;
    rc = wxb_plan_run(pl, emissions_dev, out_dev, stream, nullptr);
    wxb_plan_destroy(pl);         // synchronises the stream: the plan's device arrays outlive the kernels
    return rc;
}

extern "C" int wxb_align_host(wxb_ctx *c, const wxb_batch *bt, const float *emis_host, const wxb_outputs *oh) {
    if (!c || !bt || !oh) return fail(WXB_EINVAL, "wxb_align_host: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    const bool timing = getenv("WXB_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    // 1. start moving the emissions before anything else: fixed-size row chunks on the copy stream, one
    //    event per chunk.  Planning (host) and the kernels of earlier chunks overlap the rest of the copy.
    int64_t rows = 0, chunk_rows = 0;
    if (bt->n_segments > 0 && bt->frame_off && bt->vocab >= 1) {
        rows = bt->frame_off[bt->n_segments];
        if (rows > 0 && !emis_host) return fail(WXB_EINVAL, "wxb_align_host: emissions is NULL");
        chunk_rows = std::max<int64_t>(1, static_cast<int64_t>(c->chunk_bytes / (4ull * bt->vocab)));
    }
    int n_chunks = 0;
    std::vector<int64_t> bounds;
    if (rows > 0) {
        const size_t V = static_cast<size_t>(bt->vocab);
        CUDA_TRY(c->h_emis.ensure(static_cast<size_t>(rows) * V * 4));
        bounds = make_chunk_bounds(rows, chunk_rows, c->chunk_taper);
        n_chunks = static_cast<int>(bounds.size()) - 1;
        while (static_cast<int>(c->ev_chunk.size()) < n_chunks) {
            cudaEvent_t ev;
            CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            c->ev_chunk.push_back(ev);
        }
        if (timing) CUDA_TRY(cudaEventRecord(c->ev_t0, c->copy));
        for (int k = 0; k < n_chunks; ++k) {
            const int64_t r0 = bounds[k], r1 = bounds[k + 1];
            if (r1 > r0)
                CUDA_TRY(cudaMemcpyAsync(c->h_emis.as<float>() + r0 * V, emis_host + r0 * V, static_cast<size_t>(r1 - r0) * V * 4,
                                         cudaMemcpyHostToDevice, c->copy));
            CUDA_TRY(cudaEventRecord(c->ev_chunk[k], c->copy));
        }
        if (timing) CUDA_TRY(cudaEventRecord(c->ev_t1, c->copy));
    }
    // 2. plan while the copy engine works
    const double t_copies = now();
    wxb_plan *pl = nullptr;
    int rc = plan_create_impl(c, bt, n_chunks > 0 ? &bounds : nullptr, &pl, /*copy_engine_busy=*/n_chunks > 0);
    const double t_plan = now();
    if (rc != WXB_OK) { cudaStreamSynchronize(c->copy); return rc; }
    struct Guard { wxb_plan *p; wxb_ctx *c; ~Guard() { cudaStreamSynchronize(c->copy); wxb_plan_destroy(p); } } guard{pl, c};
    if (pl->S == 0) return WXB_OK;
    const size_t S = static_cast<size_t>(pl->S), sN = static_cast<size_t>(std::max<int64_t>(pl->sum_N, 1));
    const size_t sT = static_cast<size_t>(std::max<int64_t>(pl->sum_T, 1)), W = static_cast<size_t>(std::max<int64_t>(pl->W, 1));
    const size_t V = static_cast<size_t>(pl->V);
    wxb_outputs od;
    memset(&od, 0, sizeof od);
    CUDA_TRY(c->h_i32[0].ensure(S * 4)); od.status = c->h_i32[0].as<int32_t>();
    CUDA_TRY(c->h_i32[1].ensure(S * 4)); od.n_frames = c->h_i32[1].as<int32_t>();
    CUDA_TRY(c->h_i32[2].ensure(sN * 4)); od.char_start = c->h_i32[2].as<int32_t>();
    CUDA_TRY(c->h_i32[3].ensure(sN * 4)); od.char_end = c->h_i32[3].as<int32_t>();
    CUDA_TRY(c->h_f64[0].ensure(sN * 8)); od.char_score = c->h_f64[0].as<double>();
    const bool want_sec = oh->char_start_s || oh->char_end_s || oh->char_score_r || oh->word_start_s;
    if (want_sec) {
        CUDA_TRY(c->h_f64[1].ensure(sN * 8)); od.char_start_s = c->h_f64[1].as<double>();
        CUDA_TRY(c->h_f64[2].ensure(sN * 8)); od.char_end_s = c->h_f64[2].as<double>();
        CUDA_TRY(c->h_f64[3].ensure(sN * 8)); od.char_score_r = c->h_f64[3].as<double>();
    }
    if (oh->word_start_s || oh->word_end_s || oh->word_score) {
        CUDA_TRY(c->h_f64[4].ensure(W * 8)); od.word_start_s = c->h_f64[4].as<double>();
        CUDA_TRY(c->h_f64[5].ensure(W * 8)); od.word_end_s = c->h_f64[5].as<double>();
        CUDA_TRY(c->h_f64[6].ensure(W * 8)); od.word_score = c->h_f64[6].as<double>();
    }
    if (oh->path_token) {
        CUDA_TRY(c->h_path_tok.ensure(sT * 4)); od.path_token = c->h_path_tok.as<int32_t>();
        CUDA_TRY(c->h_f32[0].ensure(sT * 4)); od.path_logp = c->h_f32[0].as<float>();
    }
    if (oh->final_col) { CUDA_TRY(c->h_f32[1].ensure((sT + S) * 4)); od.final_col = c->h_f32[1].as<float>(); }
    if (oh->trellis) { CUDA_TRY(c->h_f32[2].ensure(static_cast<size_t>(std::max<long long>(pl->trellis_floats, 1)) * 4)); od.trellis = c->h_f32[2].as<float>(); }
    if (oh->logprobs) { CUDA_TRY(c->h_f32[3].ensure(sT * V * 4)); od.logprobs = c->h_f32[3].as<float>(); }

    cudaStream_t st = c->main;
    CUDA_TRY(cudaStreamWaitEvent(st, c->slot[0].done, 0));       // an overlapped run may still be using slot 0's scratch
    // results stream back per chunk when there is anything to run (see HostSink); otherwise in bulk below
    const bool stream_back = !pl->buckets.empty();
    HostSink sink{oh, bt->tok_off, bt->word_off, bt->frame_off};
    rc = run_impl(pl, c->h_emis.as<float>(), n_chunks > 0 ? c->ev_chunk.data() : nullptr, &od, st, nullptr,
                  stream_back ? &sink : nullptr);
    if (rc != WXB_OK) return rc;
#define WXB_D2H(field, count, type)                                                                   \
    if (oh->field && (count) > 0)                                                                     \
        CUDA_TRY(cudaMemcpyAsync(oh->field, od.field, static_cast<size_t>(count) * sizeof(type), cudaMemcpyDeviceToHost, st));
    if (!stream_back) {
        WXB_D2H(status, pl->S, int32_t) WXB_D2H(n_frames, pl->S, int32_t)
        WXB_D2H(char_start, pl->sum_N, int32_t) WXB_D2H(char_end, pl->sum_N, int32_t) WXB_D2H(char_score, pl->sum_N, double)
        WXB_D2H(char_start_s, pl->sum_N, double) WXB_D2H(char_end_s, pl->sum_N, double) WXB_D2H(char_score_r, pl->sum_N, double)
        WXB_D2H(word_start_s, pl->W, double) WXB_D2H(word_end_s, pl->W, double) WXB_D2H(word_score, pl->W, double)
        WXB_D2H(path_token, pl->sum_T, int32_t) WXB_D2H(path_logp, pl->sum_T, float)
    }
    WXB_D2H(final_col, pl->sum_T + pl->S, float) WXB_D2H(trellis, pl->trellis_floats, float)
    WXB_D2H(logprobs, pl->sum_T * pl->V, float)
#undef WXB_D2H
    const double t_enq = now();
    CUDA_TRY(cudaStreamSynchronize(st));
    if (timing && rows > 0) {
        float h2d_ms = 0.f;
        cudaEventElapsedTime(&h2d_ms, c->ev_t0, c->ev_t1);
        float k_ms = 0.f, d_ms = 0.f;
        if (stream_back) { cudaEventElapsedTime(&k_ms, c->ev_t0, c->ev_dbg[0]); cudaEventElapsedTime(&d_ms, c->ev_t0, c->ev_dbg[1]); }
        fprintf(stderr, "[wxb] align_host: since the first copy started: H2D done %.3f ms (%d chunks), kernels done %.3f ms, results on host %.3f ms\n",
                h2d_ms, n_chunks, k_ms, d_ms);
    }
    if (timing)
        fprintf(stderr, "[wxb] align_host: copies enqueued %.3f ms, plan %.3f ms, launches enqueued %.3f ms, sync %.3f ms, total %.3f ms\n",
                t_copies - t_begin, t_plan - t_copies, t_enq - t_plan, now() - t_enq, now() - t_begin);
    return WXB_OK;
}

// Debug aid (not part of the alignment path): per-CTA timeline of the last wxb_plan_run when the
// ctx was created with WXB_TRACE=1.  rows[i] = {t_begin_ns, t_forward_done_ns, t_end_ns, smid, T, N, K, seg}.
extern "C" int wxb_plan_debug_trace(wxb_plan *pl, uint64_t *rows, int64_t cap_rows) {
    if (!pl || !rows) return fail(WXB_EINVAL, "wxb_plan_debug_trace: NULL argument");
    wxb_ctx *c = pl->ctx;
    if (!c->want_trace || !c->trace.p) return fail(WXB_EINVAL, "wxb_plan_debug_trace: create the ctx with WXB_TRACE=1 and run the plan first");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaDeviceSynchronize());
    if (pl->host_descs.empty()) return fail(WXB_EINVAL, "wxb_plan_debug_trace: the plan was created without WXB_TRACE=1");
    const int n = static_cast<int>(std::min<int64_t>(pl->n_desc, cap_rows));
    std::vector<unsigned long long> raw(static_cast<size_t>(std::max(n, 1)) * 4);
    CUDA_TRY(cudaMemcpy(raw.data(), c->trace.p, static_cast<size_t>(n) * 32, cudaMemcpyDeviceToHost));
    {
        unsigned long long hist[8];
        CUDA_TRY(cudaMemcpy(hist, c->trace.as<unsigned long long>() + 4ll * std::max(pl->n_desc, 1), 64, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[wxb] frames per scheduler (compute warps | producer warps):");
        for (int i = 0; i < 4; ++i) fprintf(stderr, "  s%d %llu | %llu", i, hist[2 * i], hist[2 * i + 1]);
        fprintf(stderr, "\n");
    }
    for (int i = 0; i < n; ++i) {
        const SegDesc &d = pl->host_descs[i];
        uint64_t *r = rows + 8ll * i;
        r[0] = raw[4 * i + 0]; r[1] = raw[4 * i + 1]; r[2] = raw[4 * i + 2]; r[3] = raw[4 * i + 3];
        r[4] = d.T; r[5] = d.N; r[6] = d.K; r[7] = d.seg;
    }
    return n;
}

// ------------------------------------------------------------------ ABI: speaker assignment
extern "C" int wxb_assign_speakers_device(wxb_ctx *ctx, const double *diar_start, const double *diar_end,
                                          const int32_t *diar_speaker, int64_t n_diar, int32_t n_speakers,
                                          const double *query_start, const double *query_end, int64_t n_query,
                                          int32_t fill_nearest, int32_t *out_speaker, void *stream) {
    if (!ctx) return fail(WXB_EINVAL, "wxb_assign_speakers: ctx is NULL");
    if (n_diar < 0 || n_query < 0) return fail(WXB_EINVAL, "wxb_assign_speakers: negative count");
    if (n_speakers < 0) return fail(WXB_EINVAL, "wxb_assign_speakers: n_speakers < 0");
    if (n_query == 0) return WXB_OK;
    if (!query_start || !query_end || !out_speaker || (n_diar > 0 && (!diar_start || !diar_end || !diar_speaker)))
        return fail(WXB_EINVAL, "wxb_assign_speakers: NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long blocks = (n_query + 127) / 128;
    // up to SPK_MAX speakers fit one launch; more go through successive launches that carry the running winner
    double *best_sum = nullptr;
    if (n_speakers > SPK_MAX) {
        CUDA_TRY(ctx->spk_best.ensure(static_cast<size_t>(n_query) * 8));
        best_sum = ctx->spk_best.as<double>();
    }
    for (int base = 0; base == 0 || base < n_speakers; base += SPK_MAX) {
        assign_speakers_kernel<<<static_cast<unsigned>(blocks), 128, 0, st>>>(
            diar_start, diar_end, diar_speaker, n_diar, base, std::min(SPK_MAX, std::max(n_speakers - base, 0)), query_start, query_end,
            n_query, fill_nearest, out_speaker, best_sum);
        CUDA_TRY(cudaGetLastError());
    }
    return WXB_OK;
}

extern "C" int wxb_assign_speakers_host(wxb_ctx *ctx, const double *diar_start, const double *diar_end,
                                        const int32_t *diar_speaker, int64_t n_diar, int32_t n_speakers,
                                        const double *query_start, const double *query_end, int64_t n_query,
                                        int32_t fill_nearest, int32_t *out_speaker) {
    if (!ctx) return fail(WXB_EINVAL, "wxb_assign_speakers: ctx is NULL");
    if (n_diar < 0 || n_query < 0) return fail(WXB_EINVAL, "wxb_assign_speakers: negative count");
    if (n_query == 0) return WXB_OK;
    if (!query_start || !query_end || !out_speaker || (n_diar > 0 && (!diar_start || !diar_end || !diar_speaker)))
        return fail(WXB_EINVAL, "wxb_assign_speakers: NULL argument");
    for (int64_t i = 0; i < n_diar; ++i)
        if (diar_speaker[i] < 0 || diar_speaker[i] >= n_speakers) return fail(WXB_EINVAL, "wxb_assign_speakers: diar_speaker[%lld]=%d outside [0,%d)", (long long)i, diar_speaker[i], n_speakers);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t nd = static_cast<size_t>(std::max<int64_t>(n_diar, 1)), nq = static_cast<size_t>(n_query);
    CUDA_TRY(ctx->h_f64[0].ensure(nd * 8)); CUDA_TRY(ctx->h_f64[1].ensure(nd * 8)); CUDA_TRY(ctx->h_i32[0].ensure(nd * 4));
    CUDA_TRY(ctx->h_f64[2].ensure(nq * 8)); CUDA_TRY(ctx->h_f64[3].ensure(nq * 8)); CUDA_TRY(ctx->h_i32[1].ensure(nq * 4));
    cudaStream_t st = ctx->main;
    if (n_diar > 0) {
        CUDA_TRY(cudaMemcpyAsync(ctx->h_f64[0].p, diar_start, nd * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ctx->h_f64[1].p, diar_end, nd * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ctx->h_i32[0].p, diar_speaker, nd * 4, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(ctx->h_f64[2].p, query_start, nq * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ctx->h_f64[3].p, query_end, nq * 8, cudaMemcpyHostToDevice, st));
    int rc = wxb_assign_speakers_device(ctx, ctx->h_f64[0].as<double>(), ctx->h_f64[1].as<double>(), ctx->h_i32[0].as<int32_t>(), n_diar,
                                        n_speakers, ctx->h_f64[2].as<double>(), ctx->h_f64[3].as<double>(), n_query, fill_nearest,
                                        ctx->h_i32[1].as<int32_t>(), st);
    if (rc != WXB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(out_speaker, ctx->h_i32[1].p, nq * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return WXB_OK;
}

// ------------------------------------------------------------------ ABI: exact scores on the host
extern "C" int wxb_exact_scores_host(const wxb_batch *bt, const float *path_prob, const int32_t *status,
                                     const int32_t *char_start, const int32_t *char_end,
                                     double *char_score, double *char_score_r, double *word_score) {
    if (!bt || !status || !char_start || !char_end || !char_score) return fail(WXB_EINVAL, "wxb_exact_scores_host: NULL argument");
    const int64_t S = bt->n_segments;
    if (S <= 0) return WXB_OK;
    if (!bt->frame_off || !bt->tok_off) return fail(WXB_EINVAL, "wxb_exact_scores_host: frame_off/tok_off is NULL");
    if (!path_prob && bt->frame_off[S] > 0) return fail(WXB_EINVAL, "wxb_exact_scores_host: path_prob is NULL");
    const double nan = std::nan("");
    for (int64_t s = 0; s < S; ++s) {
        const int64_t k0 = bt->tok_off[s], k1 = bt->tok_off[s + 1];
        const int64_t T = bt->frame_len ? bt->frame_len[s] : bt->frame_off[s + 1] - bt->frame_off[s];
        const float *p = path_prob + bt->frame_off[s];
        for (int64_t k = k0; k < k1; ++k) {
            if (status[s] != WXB_SEG_OK) { char_score[k] = nan; if (char_score_r) char_score_r[k] = nan; continue; }
            const int a = char_start[k], b = char_end[k];
            if (a < 0 || b <= a || b > T) return fail(WXB_EINVAL, "wxb_exact_scores_host: char %lld of segment %lld spans [%d,%d) outside its %lld frames", (long long)(k - k0), (long long)s, a, b, (long long)T);
            double sum = static_cast<double>(p[a]);                   // Python sum(): 0 + p_first + ...
            for (int f = a + 1; f < b; ++f) sum += static_cast<double>(p[f]);
            const double score = sum / static_cast<double>(b - a);
            char_score[k] = score;
            if (char_score_r) char_score_r[k] = wxb::round3(score);
        }
    }
    if (!word_score || bt->n_words <= 0) return WXB_OK;
    if (!char_score_r || !bt->word_off || !bt->wchar_off || !bt->wchar_idx) return fail(WXB_EINVAL, "wxb_exact_scores_host: word scores need char_score_r and the word groups");
    for (int64_t s = 0; s < S; ++s) {
        const int64_t base = bt->tok_off[s];
        for (int64_t w = bt->word_off[s]; w < bt->word_off[s + 1]; ++w) {
            const int64_t lo = bt->wchar_off[w];
            const long n = static_cast<long>(bt->wchar_off[w + 1] - lo);
            long count = 0;
            if (status[s] == WXB_SEG_OK)
                for (long i = 0; i < n; ++i) count += bt->wchar_idx[lo + i] >= 0;
            if (count == 0) { word_score[w] = nan; continue; }
            const double sum = wxb::pairwise_sum([&](long i) { const int idx = bt->wchar_idx[lo + i]; return idx >= 0 ? char_score_r[base + idx] : 0.0; }, n);
            word_score[w] = wxb::round3_np(sum / static_cast<double>(count));
        }
    }
    return WXB_OK;
}

// ------------------------------------------------------------------ ABI: stand-alone kernels
extern "C" int wxb_log_softmax_device(wxb_ctx *ctx, const float *logits_dev, int64_t rows, int32_t vocab,
                                      float *out_dev, void *stream) {
    if (!ctx || (rows > 0 && (!logits_dev || !out_dev))) return fail(WXB_EINVAL, "wxb_log_softmax_device: NULL argument");
    if (rows < 0 || vocab < 1) return fail(WXB_EINVAL, "wxb_log_softmax_device: bad shape");
    if (rows == 0) return WXB_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    return launch_log_softmax(logits_dev, out_dev, rows, vocab, static_cast<cudaStream_t>(stream));
}

extern "C" int wxb_backtrack_trellis_device(wxb_ctx *ctx, const float *trellis_dev, const float *emission_dev,
                                            const int32_t *tokens_dev, int32_t T, int32_t N, int32_t vocab,
                                            int32_t blank_id, int32_t *path_token_dev, float *path_logp_dev,
                                            int32_t *status_dev, int32_t *n_frames_dev, void *stream) {
    if (!ctx || !trellis_dev || !emission_dev || !tokens_dev || !path_token_dev || !path_logp_dev || !status_dev || !n_frames_dev)
        return fail(WXB_EINVAL, "wxb_backtrack_trellis_device: NULL argument");
    if (T < 0 || N < 1 || vocab < 1 || blank_id < 0 || blank_id >= vocab)
        return fail(WXB_EINVAL, "wxb_backtrack_trellis_device: bad shape (T=%d N=%d V=%d blank=%d)", T, N, vocab, blank_id);
    CUDA_TRY(cudaSetDevice(ctx->device));
    backtrack_trellis_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(trellis_dev, emission_dev, tokens_dev, T, N, vocab,
                                                                              blank_id, path_token_dev, path_logp_dev,
                                                                              status_dev, n_frames_dev);
    CUDA_TRY(cudaGetLastError());
    return WXB_OK;
}

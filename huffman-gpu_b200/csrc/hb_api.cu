/*
 * hb_api.cu -- the C ABI of libhuffb200.so (see include/huffman_b200.h).
 *
 * Mirrors the five-step sequence of the reference's runVLCTest (main_test_cu.cu:52-180):
 *   init -> histogram (hist.cu) -> codebook (huffTree.h, load_data.h:40-47)
 *        -> encode (vlc_kernel_sm64huff.cu + scan.cu + pack_kernels.cu) -> free
 * with caller-owned data buffers, explicit streams, error codes instead of exit(), and no hidden
 * globals (the reference keeps scan scratch in file-scope statics, scan.cu:63-65).
 * There is no CPU fallback anywhere in this file.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/huffman_b200.h"
#include "hb_ctx.h"
#include "hb_kernels.cuh"


namespace {

int cuda_fail(hb_ctx *ctx, cudaError_t e)
{
    if (ctx) ctx->last_cuda = (int)e;
    (void)cudaGetLastError();   // clear the sticky-less error state
    return HB_ERR_CUDA;
}

#define HB_CUDA(ctx, call)                                   \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return cuda_fail((ctx), e__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess) {
            ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
        } else {
            ok = cudaSetDevice(dev) == cudaSuccess;
            prev = -1;
        }
    }
    ~DeviceGuard()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) (void)cudaSetDevice(prev);
    }
};

// The kernel variant for a codebook; $HB_FORCE_GROUP overrides the group size (tuning aid only).
int forced_group()
{
    const char *env = getenv("HB_FORCE_GROUP");
    return env ? atoi(env) : 0;
}

hb::EncVariant choose_variant(const uint32_t len[256])
{
    hb::EncVariant v = hb::pick_variant(len);
    const int g = forced_group();
    if (g == 1 || g == 2 || g == 4 || (!v.wide && (g == 3 || g == 6 || g == 8))) {
        uint32_t max_len = 0;
        for (int s = 0; s < 256; s++)
            if (len[s] > max_len) max_len = len[s];
        v.group = g;
        v.check = g > 1 && (v.wide || (uint32_t)g * max_len > 31u);
    }
    return v;
}

// Validate the caller's tables against the parity domain and pack them for the kernel.
int pack_tables(const uint32_t cw[256], const uint32_t len[256], uint32_t packed[512],
                hb::EncVariant *variant)
{
    for (int s = 0; s < 256; s++) {
        if (len[s] > HB_MAX_CODE_LEN) return HB_ERR_CODELEN;
        if (len[s] < 32 && (cw[s] >> len[s]) != 0) return HB_ERR_CODEWORD;
    }
    const hb::EncVariant v = choose_variant(len);
    memset(packed, 0, 512 * sizeof(uint32_t));
    for (int s = 0; s < 256; s++) {
        const uint32_t l = len[s];
        const uint32_t left = l ? (cw[s] << (32u - l)) : 0u;
        if (v.wide) {
            packed[2 * s] = left;
            packed[2 * s + 1] = l;
        } else {
            packed[s] = left | l;      // l <= 24: the codeword bits sit at or above bit 8
        }
    }
    *variant = v;
    return HB_OK;
}

int set_codebook(hb_ctx *ctx, const uint32_t cw[256], const uint32_t len[256], cudaStream_t stream)
{
    const int forced = forced_group();
    if (ctx->table_valid && forced == ctx->forced && memcmp(cw, ctx->cw_cache, sizeof(ctx->cw_cache)) == 0 &&
        memcmp(len, ctx->len_cache, sizeof(ctx->len_cache)) == 0)
        return HB_OK;
    uint32_t packed[512];
    hb::EncVariant v;
    const int rc = pack_tables(cw, len, packed, &v);
    if (rc != HB_OK) return rc;
    // the pinned staging block may still be in flight from the previous upload
    HB_CUDA(ctx, cudaEventSynchronize(ctx->table_uploaded));
    memcpy(ctx->h_table, packed, sizeof(packed));
    HB_CUDA(ctx, cudaMemcpyAsync(ctx->d_table, ctx->h_table, sizeof(packed), cudaMemcpyHostToDevice,
                                 stream));
    HB_CUDA(ctx, cudaEventRecord(ctx->table_uploaded, stream));
    memcpy(ctx->cw_cache, cw, sizeof(ctx->cw_cache));
    memcpy(ctx->len_cache, len, sizeof(ctx->len_cache));
    ctx->variant = v;
    ctx->forced = forced;
    ctx->table_valid = true;
    return HB_OK;
}

uint64_t tiles_of(uint64_t n_words) { return (n_words + hb::kTileWords - 1) / hb::kTileWords; }

// where hist_kernel reports that it refused to run (mapped pinned: the device writes it, the host reads it)
unsigned long long *hist_flag(hb_ctx *ctx) { return &ctx->h_result[kMaxChunks].overflow; }

// The look-back trees after a failed or refused launch: nothing may be assumed about them any more.
void reset_trees(hb_ctx *ctx)
{
    (void)cudaDeviceSynchronize();
    for (int i = 0; i < 2; i++) {
        (void)cudaMemset(ctx->d_tree[i], 0, ctx->max_tiles * sizeof(unsigned long long));
        ctx->tree_dirty[i] = 0;
    }
    (void)cudaDeviceSynchronize();
    (void)cudaGetLastError();
}

// A context serialises its jobs: they share the device table, the result block and the two alternating trees.  A job
// given another stream than the previous one is ordered behind it with an event (same stream: stream order).
int order_after_previous_job(hb_ctx *ctx, cudaStream_t stream)
{
    if (ctx->have_last_stream && ctx->last_stream != stream)
        HB_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->job_done, 0));
    return HB_OK;
}

int job_launched(hb_ctx *ctx, cudaStream_t stream)
{
    HB_CUDA(ctx, cudaEventRecord(ctx->job_done, stream));
    ctx->last_stream = stream;
    ctx->have_last_stream = true;
    return HB_OK;
}

// Launch the encode kernel over tiles [first_tile, end_tile) of a job.  A launch with first_tile == 0 starts a new job:
// it takes the tree the previous job's kernel cleared and clears the previous job's tree in turn; that bookkeeping is
// committed only once the launch has been accepted.
int launch_tiles(hb_ctx *ctx, const uint32_t *d_in, uint64_t n_words, uint64_t first_tile,
                 uint64_t end_tile, uint32_t *d_out, uint64_t cap_words, uint64_t start_bit,
                 cudaStream_t stream, int result_slot = 0)
{
    const bool new_job = first_tile == 0;
    const int cur = new_job ? (ctx->tree_cur ^ 1) : ctx->tree_cur;
    hb::EncParams p;
    p.in = d_in;
    p.n_words = n_words;
    p.n_tiles = tiles_of(n_words);
    if (p.n_tiles > hb::kMaxJobTiles) return HB_ERR_CAPACITY;      // the tree node's tile-count field (hb_encode.cu)
    p.first_tile = first_tile;
    p.end_tile = end_tile;
    p.out = d_out;
    p.out_cap_words = cap_words;
    p.start_bit = start_bit;
    p.tree = ctx->d_tree[cur];
    p.tree_zero = ctx->d_tree[cur ^ 1];
    p.zero_count = new_job ? ctx->tree_dirty[cur ^ 1] : 0;
    p.table = ctx->d_table;
    p.result = ctx->h_result + result_slot;
    p.prof = ctx->d_prof;
    p.seam_flags = ctx->next_seam_flags;
    {
        static const char *pf = getenv("HB_L2_PREFETCH");
        p.l2_prefetch = (pf && atoi(pf) == 0) ? 0u : 1u;    // measured: +1.4 % (1 GiB, H 2.2) .. +2.5 % (H 7.9); $HB_L2_PREFETCH=0 turns it off
    }

    // one persistent CTA per SM (its shared memory holds the 64 KiB table and the staging rings)
    const uint64_t tiles = end_tile - first_tile;
    uint64_t grid = (uint64_t)ctx->sm_count;
    if (grid > tiles) grid = tiles;
    const cudaError_t e = hb::launch_encode(ctx->variant, p, (int)grid, stream);
    if (e != cudaSuccess) {
        const int rc = cuda_fail(ctx, e);
        reset_trees(ctx);                       // earlier launches of this job may have used the tree
        return rc;
    }
    if (new_job) {
        ctx->tree_cur = cur;
        ctx->tree_dirty[cur ^ 1] = 0;           // being cleared by this kernel
        ctx->tree_dirty[cur] = p.n_tiles;
    }
    ctx->launches++;
    return HB_OK;
}

}  // namespace

extern "C" {

int hb_init(hb_ctx **out, int device, uint64_t max_words)
{
    if (!out) return HB_ERR_ARG;
    *out = nullptr;
    hb_ctx *ctx = new (std::nothrow) hb_ctx();
    if (!ctx) return HB_ERR_NOMEM;
    ctx->device = device;
    memset(ctx->cw_cache, 0, sizeof(ctx->cw_cache));
    memset(ctx->len_cache, 0, sizeof(ctx->len_cache));

    DeviceGuard g(device);                    // the caller's current device is restored on return
    cudaError_t e = g.ok ? cudaSuccess : cudaGetLastError();
    if (e == cudaSuccess && !g.ok) e = cudaErrorInvalidDevice;
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess && prop.major < 10) e = cudaErrorNoKernelImageForDevice;   // sm_100a only
    if (e == cudaSuccess) e = hb::encode_configure();
    if (e == cudaSuccess) e = hb::histogram_configure();
    if (e != cudaSuccess) {
        int rc = cuda_fail(ctx, e);
        fprintf(stderr, "hb_init: CUDA error %d (%s); libhuffb200 has no CPU fallback\n", (int)e,
                cudaGetErrorString(e));
        delete ctx;
        return rc;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_words = max_words ? max_words : 1;
    if (tiles_of(ctx->max_words) > hb::kMaxJobTiles) {          // 64 GiB of input per job (hb_encode.cu, tree node layout)
        delete ctx;
        return HB_ERR_CAPACITY;
    }
    ctx->max_tiles = tiles_of(ctx->max_words) + 1;

    // the first failing call decides the status: allocation failures are HB_ERR_NOMEM, anything else HB_ERR_CUDA
    cudaError_t first = cudaSuccess;
    auto tryc = [&](cudaError_t r) {
        if (first == cudaSuccess && r != cudaSuccess) first = r;
        return first == cudaSuccess;
    };
    for (int i = 0; i < 2; i++)
        if (tryc(cudaMalloc(&ctx->d_tree[i], ctx->max_tiles * sizeof(unsigned long long))))
            tryc(cudaMemset(ctx->d_tree[i], 0, ctx->max_tiles * sizeof(unsigned long long)));
    tryc(cudaMalloc(&ctx->d_table, 512 * sizeof(uint32_t)));
    tryc(cudaMallocHost(&ctx->h_table, 512 * sizeof(uint32_t)));
    tryc(cudaHostAlloc(&ctx->h_result, (kMaxChunks + 1) * sizeof(hb::EncResult), cudaHostAllocMapped));
    tryc(cudaMalloc(&ctx->d_hist, 256 * sizeof(unsigned long long)));
    tryc(cudaMalloc(&ctx->d_thr, 256 * sizeof(uint32_t)));
    tryc(cudaMalloc(&ctx->d_symmap, 256));
    tryc(cudaEventCreateWithFlags(&ctx->table_uploaded, cudaEventDisableTiming));
    tryc(cudaEventCreateWithFlags(&ctx->job_done, cudaEventDisableTiming));
    for (int i = 0; i < kMaxChunks; i++) {
        tryc(cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming));
        tryc(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
    }
    tryc(cudaStreamCreateWithFlags(&ctx->s_main, cudaStreamNonBlocking));
    tryc(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    tryc(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    if (getenv("HB_PROFILE")) {
        // 32 global counters, then per worker index: cycles spent waiting for records, and the worker's total
        if (tryc(cudaMalloc(&ctx->d_prof, (32 + 512 + 1280 + 960) * sizeof(unsigned long long))))
            tryc(cudaMemset(ctx->d_prof, 0, (32 + 512 + 1280 + 960) * sizeof(unsigned long long)));
    }
    tryc(cudaDeviceSynchronize());
    if (first == cudaSuccess) {
        // The histogram kernel refuses to run when the shared window is not laid out as it assumes, and says so in
        // the result block (never in the data).  That is a property of driver + kernel image: probe it once, here, so
        // that the asynchronous hb_histogram_device can never hand a refused histogram to the codebook builder.
        memset(ctx->h_result, 0, (kMaxChunks + 1) * sizeof(hb::EncResult));
        tryc(cudaMemset(ctx->d_hist, 0, 256 * sizeof(unsigned long long)));
        tryc(cudaMemset(ctx->d_table, 0, 512 * sizeof(uint32_t)));
        tryc(hb::launch_histogram(ctx->d_table, 16, ctx->d_hist, ctx->sm_count, hist_flag(ctx), nullptr));
        tryc(cudaDeviceSynchronize());
        if (first == cudaSuccess && ctx->h_result[kMaxChunks].overflow) {
            fprintf(stderr, "hb_init: hist_kernel refused its shared-memory layout on this driver\n");
            hb_free(ctx);
            return HB_ERR_STATE;
        }
    }
    if (first != cudaSuccess) {
        const int rc = first == cudaErrorMemoryAllocation ? HB_ERR_NOMEM : HB_ERR_CUDA;
        fprintf(stderr, "hb_init: CUDA error %d (%s)\n", (int)first, cudaGetErrorString(first));
        (void)cudaGetLastError();
        hb_free(ctx);
        return rc;
    }
    *out = ctx;
    return HB_OK;
}

void hb_free(hb_ctx *ctx)
{
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    (void)cudaDeviceSynchronize();
    if (ctx->d_prof) {
        static const char *names[] = {"w0_wait_tile", "w0_wait_prefix", "w0_total", "(unused)", "rs_wait_agg",
                                      "rs_lookback", "rs_bits_before", "rs_total", "tiles", "lookback_polls",
                                      "w0_pass1", "w0_emit", "w0_copy"};
        static unsigned long long v[32 + 512 + 1280 + 960];
        if (cudaMemcpy(v, ctx->d_prof, sizeof(v), cudaMemcpyDeviceToHost) == cudaSuccess) {
            const double tiles = v[8] ? (double)v[8] : 1.0;
            for (int i = 0; i < 13; i++)
                fprintf(stderr, "hb_prof %-16s %14llu  per tile %10.1f\n", names[i], v[i], (double)v[i] / tiles);
            // share of its time each worker warp waited for records (the global counters above then sum all 16 workers)
            fprintf(stderr, "hb_prof per-worker wait share (%%), all CTAs:");
            for (int b = 0; b < 16; b++)
                fprintf(stderr, "%s%3.0f", (b % 37) ? " " : "\n  ", v[32 + 256 + b] ? 100.0 * (double)v[32 + b] / (double)v[32 + 256 + b] : 0.0);
            fprintf(stderr, "\n");
            if (getenv("HB_PROFILE_SKEW")) {
                static const char *what[] = {"pass 1", "scan + pass 2 (+ waits for ring space)", "copy-out", "waits for records"};
                for (int c = 0; c < 4; c++) {
                    fprintf(stderr, "hb_prof worker 0, kilocycles in %s, per CTA (last launch):", what[c]);
                    for (int b = 0; b < ctx->sm_count && b < 160; b++)
                        fprintf(stderr, "%s%4.0f", (b % 37) ? " " : "\n   ", (double)v[32 + 512 + 640 + c * 160 + b] * 1e-3);
                    fprintf(stderr, "\n");
                }
            }
            {
                // timeline of the last launch, relative to the first CTA's entry (global timer)
                static const char *stage[] = {"kernel entry", "prologue done", "first tile published", "last tile published",
                                              "worker 0 encoded its last chunk", "worker 0 copied its last chunk out"};
                unsigned long long t0 = ~0ULL;
                for (int b = 0; b < ctx->sm_count && b < 160; b++)
                    if (v[32 + 512 + 1280 + b] && v[32 + 512 + 1280 + b] < t0) t0 = v[32 + 512 + 1280 + b];
                for (int c = 0; c < 6; c++) {
                    unsigned long long lo = ~0ULL, hi = 0;
                    for (int b = 0; b < ctx->sm_count && b < 160; b++) {
                        const unsigned long long t = v[32 + 512 + 1280 + c * 160 + b];
                        if (t && t < lo) lo = t;
                        if (t > hi) hi = t;
                    }
                    if (hi)
                        fprintf(stderr, "hb_prof timeline %-36s first CTA %8.2f us, last CTA %8.2f us\n", stage[c],
                                (double)(lo - t0) * 1e-3, (double)(hi - t0) * 1e-3);
                }
            }
            // skew between CTAs: when each published its tile at 1/4, 1/2, 3/4 and the end of its sequence
            for (int c = 0; c < 4; c++) {
                unsigned long long lo = ~0ULL, hi = 0;
                for (int b = 0; b < ctx->sm_count && b < 160; b++) {
                    const unsigned long long t = v[32 + 512 + c * 160 + b];
                    if (t && t < lo) lo = t;
                    if (t > hi) hi = t;
                }
                fprintf(stderr, "hb_prof publish skew at %d/4 of the run: %.2f us between first and last CTA\n", c + 1,
                        hi ? (double)(hi - lo) * 1e-3 : 0.0);
                if (getenv("HB_PROFILE_SKEW")) {
                    if (c == 0) {
                        fprintf(stderr, "  SM id, per CTA:");
                        for (int b = 0; b < ctx->sm_count && b < 160; b++)
                            fprintf(stderr, "%s%3llu", (b % 37) ? " " : "\n   ", v[32 + 256 + 64 + b]);
                        fprintf(stderr, "\n");
                    }
                    fprintf(stderr, "  lag behind the first CTA (us), per CTA:");
                    for (int b = 0; b < ctx->sm_count && b < 160; b++)
                        fprintf(stderr, "%s%3.0f", (b % 37) ? " " : "\n   ", (double)(v[32 + 512 + c * 160 + b] - lo) * 1e-3);
                    fprintf(stderr, "\n");
                }
            }
        }
        cudaFree(ctx->d_prof);
    }
    cudaFree(ctx->d_tree[0]);
    cudaFree(ctx->d_tree[1]);
    cudaFree(ctx->d_table);
    if (ctx->h_table) cudaFreeHost(ctx->h_table);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    cudaFree(ctx->d_hist);
    cudaFree(ctx->d_thr);
    cudaFree(ctx->d_symmap);
    cudaFree(ctx->d_in_buf);
    cudaFree(ctx->d_out_buf);
    cudaFree(ctx->d_dec_lut);
    cudaFree(ctx->d_dec_trie);
    cudaFree(ctx->d_dec_error);
    if (ctx->h_dec_stage) cudaFreeHost(ctx->h_dec_stage);
    if (ctx->table_uploaded) cudaEventDestroy(ctx->table_uploaded);
    if (ctx->job_done) cudaEventDestroy(ctx->job_done);
    for (int i = 0; i < kMaxChunks; i++) {
        if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
    }
    if (ctx->s_main) cudaStreamDestroy(ctx->s_main);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    (void)cudaGetLastError();
    delete ctx;
}

int hb_histogram_device(hb_ctx *ctx, const uint32_t *d_in, uint64_t n_words, uint64_t *d_hist,
                        void *stream)
{
    if (!ctx || !d_hist || (!d_in && n_words) || ((uintptr_t)d_in & 3u)) return HB_ERR_ARG;
    DeviceGuard g(ctx->device);
    if (*(volatile unsigned long long *)hist_flag(ctx)) return HB_ERR_STATE;     // refused before (see hb_init)
    HB_CUDA(ctx, hb::launch_histogram(d_in, n_words, (unsigned long long *)d_hist, ctx->sm_count, hist_flag(ctx),
                                      (cudaStream_t)stream));
    if (n_words) ctx->launches++;
    return HB_OK;
}

int hb_histogram(hb_ctx *ctx, const uint32_t *d_in, uint64_t n_words, uint64_t hist[256], void *stream)
{
    if (!ctx || !hist) return HB_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA(ctx, cudaMemsetAsync(ctx->d_hist, 0, 256 * sizeof(unsigned long long), st));
    const int rc = hb_histogram_device(ctx, d_in, n_words, (uint64_t *)ctx->d_hist, stream);
    if (rc != HB_OK) return rc;
    HB_CUDA(ctx, cudaMemcpyAsync(hist, ctx->d_hist, 256 * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, st));
    HB_CUDA(ctx, cudaStreamSynchronize(st));
    if (*(volatile unsigned long long *)hist_flag(ctx)) return HB_ERR_STATE;     // the kernel refused its shared-memory layout
    return HB_OK;
}

int hb_encode_async(hb_ctx *ctx, const uint32_t *d_in, uint64_t n_words, const uint32_t codewords[256],
                    const uint32_t codewordlens[256], uint32_t *d_out, uint64_t out_capacity_words,
                    uint64_t start_bit, void *stream)
{
    if (!ctx || !codewords || !codewordlens || !d_out || (!d_in && n_words)) return HB_ERR_ARG;
    if (((uintptr_t)d_in & 31u) || ((uintptr_t)d_out & 3u)) return HB_ERR_ARG;
    if (n_words > ctx->max_words) return HB_ERR_CAPACITY;
    if (start_bit >> 47) return HB_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    // a job on another stream than the previous one waits for it: the table upload below must not overtake a kernel
    // that still reads the old table
    int rc = order_after_previous_job(ctx, st);
    if (rc != HB_OK) return rc;
    if ((rc = set_codebook(ctx, codewords, codewordlens, st)) != HB_OK) return rc;

    // Several encodes may be queued on the same stream: the kernel only ever SETS result->overflow and
    // the last launch's last tile writes result->bits_end; the host touches the block in
    // hb_encode_result only, after the stream has drained.
    if (n_words == 0) {
        ctx->next_seam_flags = 0;
        // cpuencode.cpp:17 -- the first output word is cleared even for an empty input
        if ((start_bit >> 5) >= out_capacity_words) return HB_ERR_CAPACITY;
        if ((start_bit & 31u) == 0)
            HB_CUDA(ctx, cudaMemsetAsync(d_out + (start_bit >> 5), 0, sizeof(uint32_t), st));
        ctx->pending = true;
        ctx->pending_empty = true;
        ctx->pending_start_bit = start_bit;
        return HB_OK;
    }
    rc = launch_tiles(ctx, d_in, n_words, 0, tiles_of(n_words), d_out, out_capacity_words, start_bit, st);
    ctx->next_seam_flags = 0;
    if (rc != HB_OK) return rc;
    if ((rc = job_launched(ctx, st)) != HB_OK) return rc;
    ctx->pending = true;
    ctx->pending_empty = false;
    ctx->pending_start_bit = start_bit;
    ctx->last_job_tiles = tiles_of(n_words);
    ctx->last_job_words = n_words;
    ctx->last_job_start_bit = start_bit;
    ctx->last_job_valid = true;
    return HB_OK;
}

int hb_encode_result(hb_ctx *ctx, uint64_t *total_bits, void *stream)
{
    if (!ctx) return HB_ERR_ARG;
    if (!ctx->pending) return HB_ERR_STATE;
    DeviceGuard g(ctx->device);
    HB_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    ctx->pending = false;
    const unsigned long long overflow = ctx->h_result->overflow;
    ctx->h_result->overflow = 0;
    if (total_bits)
        *total_bits = ctx->pending_empty ? 0 : ctx->h_result->bits_end - ctx->pending_start_bit;
    if (overflow == 2ULL) {                          // the kernel refused its shared-memory layout: it cleared nothing
        reset_trees(ctx);
        return HB_ERR_STATE;
    }
    return overflow ? HB_ERR_CAPACITY : HB_OK;
}

int hb_encode(hb_ctx *ctx, const uint32_t *d_in, uint64_t n_words, const uint32_t codewords[256],
              const uint32_t codewordlens[256], uint32_t *d_out, uint64_t out_capacity_words,
              uint64_t start_bit, uint64_t *total_bits, void *stream)
{
    const int rc = hb_encode_async(ctx, d_in, n_words, codewords, codewordlens, d_out,
                                   out_capacity_words, start_bit, stream);
    if (rc != HB_OK) return rc;
    return hb_encode_result(ctx, total_bits, stream);
}

// ---- host-buffer pipeline ------------------------------------------------------------------------
static int ensure_buffers(hb_ctx *ctx, uint64_t in_words, uint64_t out_words)
{
    if (in_words > ctx->in_buf_words) {
        cudaFree(ctx->d_in_buf);
        ctx->d_in_buf = nullptr;
        ctx->in_buf_words = 0;
        HB_CUDA(ctx, cudaMalloc(&ctx->d_in_buf, (in_words + 8) * sizeof(uint32_t)));
        ctx->in_buf_words = in_words;
    }
    if (out_words > ctx->out_buf_words) {
        cudaFree(ctx->d_out_buf);
        ctx->d_out_buf = nullptr;
        ctx->out_buf_words = 0;
        HB_CUDA(ctx, cudaMalloc(&ctx->d_out_buf, (out_words + 8) * sizeof(uint32_t)));
        ctx->out_buf_words = out_words;
    }
    return HB_OK;
}

int hb_vlc_encode_host(hb_ctx *ctx, const uint32_t *h_in, uint64_t n_words, uint32_t *h_out,
                       uint64_t out_capacity_words, const uint32_t codewords[256],
                       const uint32_t codewordlens[256], uint64_t *out_bytes, uint64_t *total_bits)
{
    if (!ctx || !h_out || !codewords || !codewordlens || (!h_in && n_words)) return HB_ERR_ARG;
    if (n_words > ctx->max_words) return HB_ERR_CAPACITY;
    if (out_capacity_words == 0) return HB_ERR_CAPACITY;
    DeviceGuard g(ctx->device);

    uint32_t max_len = 0;
    for (int s = 0; s < 256; s++) {
        if (codewordlens[s] > HB_MAX_CODE_LEN) return HB_ERR_CODELEN;
        if (codewordlens[s] > max_len) max_len = codewordlens[s];
    }
    // device output: never more than the caller can take, never more than the worst case
    uint64_t worst = (n_words * 4 * (uint64_t)max_len) / 32 + 2;
    uint64_t dev_out_words = worst < out_capacity_words ? worst : out_capacity_words;
    int rc = ensure_buffers(ctx, n_words, dev_out_words);
    if (rc != HB_OK) return rc;

    cudaStream_t st = ctx->s_main;
    if (ctx->pending) {                        // un-fetched async encodes: drain them first
        HB_CUDA(ctx, cudaDeviceSynchronize());
        ctx->pending = false;
    }
    if ((rc = order_after_previous_job(ctx, st)) != HB_OK) return rc;
    if ((rc = set_codebook(ctx, codewords, codewordlens, st)) != HB_OK) return rc;
    ctx->last_job_valid = false;               // (hb_encode_tile_index describes device-buffer jobs only)
    memset(ctx->h_result, 0, kMaxChunks * sizeof(hb::EncResult));

    uint64_t bits = 0;
    if (n_words == 0) {
        h_out[0] = 0;                          // cpuencode.cpp:17
    } else {
        // Chunked, three streams: the H2D copies run back to back on their own stream, the encode of chunk k (which
        // waits for copy k only) overlaps copy k+1, and the D2H copy of the words chunk k completed overlaps both
        // (PCIe is full duplex).  Small chunks keep the fill (first copy) and the drain (last encode + last D2H) short.
        // All chunks belong to ONE job (one look-back tree, one output stream); a later launch looks back into the
        // tree nodes and symbols of the earlier ones.
        const uint64_t total_tiles = tiles_of(n_words);
        uint64_t chunk_tiles = (16ull << 20) / hb::kTileBytes;           // 16 MiB of input per chunk ...
        static const char *env = getenv("HB_CHUNK_MIB");
        if (env && atoi(env) > 0) chunk_tiles = ((uint64_t)atoi(env) << 20) / hb::kTileBytes;
        if (chunk_tiles < 1) chunk_tiles = 1;
        if ((total_tiles + chunk_tiles - 1) / chunk_tiles > (uint64_t)kMaxChunks)    // ... at most kMaxChunks of them
            chunk_tiles = (total_tiles + kMaxChunks - 1) / kMaxChunks;
        // Every exit after the first launch goes through `fail`: launches may still be queued on s_main and copies into
        // the caller's h_out in flight on s_d2h; nothing may be left running, dirty or half-ordered behind our back.
        bool launched = false, kernel_state_lost = false;
        auto fail = [&](int status) {
            if (launched) {
                (void)cudaStreamSynchronize(ctx->s_h2d);
                (void)cudaStreamSynchronize(st);
                (void)cudaStreamSynchronize(ctx->s_d2h);
                (void)cudaGetLastError();
                (void)job_launched(ctx, st);
            }
            memset(ctx->h_result, 0, kMaxChunks * sizeof(hb::EncResult));
            if (kernel_state_lost) reset_trees(ctx);
            return status;
        };
#define HB_TRY(call)                                             \
        do {                                                     \
            cudaError_t e__ = (call);                            \
            if (e__ != cudaSuccess) return fail(cuda_fail(ctx, e__)); \
        } while (0)
        int n_chunks = 0;
        for (uint64_t t0 = 0; t0 < total_tiles; t0 += chunk_tiles, n_chunks++) {
            const uint64_t t1 = (t0 + chunk_tiles < total_tiles) ? t0 + chunk_tiles : total_tiles;
            const uint64_t w0 = t0 * hb::kTileWords;
            const uint64_t w1 = (t1 * hb::kTileWords < n_words) ? t1 * hb::kTileWords : n_words;
            HB_TRY(cudaMemcpyAsync(ctx->d_in_buf + w0, h_in + w0, (w1 - w0) * sizeof(uint32_t),
                                   cudaMemcpyHostToDevice, ctx->s_h2d));
            HB_TRY(cudaEventRecord(ctx->ev_h2d[n_chunks], ctx->s_h2d));
            HB_TRY(cudaStreamWaitEvent(st, ctx->ev_h2d[n_chunks], 0));
            rc = launch_tiles(ctx, ctx->d_in_buf, n_words, t0, t1, ctx->d_out_buf, dev_out_words, 0, st, n_chunks);
            if (rc != HB_OK) {
                // A later launch of a job that never happens leaves the earlier ones polling tree nodes nobody will
                // complete: launch_tiles has already synchronised the device and reset the trees in that case.
                return fail(rc);
            }
            launched = true;
            HB_TRY(cudaEventRecord(ctx->ev_chunk[n_chunks], st));
        }
        // as each launch retires, every output word below its end bit is final: send it home
        uint64_t done_words = 0;
        for (int c = 0; c < n_chunks; c++) {
            HB_TRY(cudaEventSynchronize(ctx->ev_chunk[c]));
            if (ctx->h_result[c].overflow == 2ULL) {
                kernel_state_lost = true;
                return fail(HB_ERR_STATE);
            }
            if (ctx->h_result[c].overflow) return fail(HB_ERR_CAPACITY);
            bits = ctx->h_result[c].bits_end;
            // the last launch also owns the final partial word and the reference's courtesy zero word
            uint64_t upto = (c + 1 == n_chunks) ? bits / 32 + 1 : bits / 32;
            if (c + 1 == n_chunks) {
                if (upto > out_capacity_words) upto = out_capacity_words;
                if (upto < (bits + 31) / 32) return fail(HB_ERR_CAPACITY);
            }
            if (upto > done_words) {
                HB_TRY(cudaMemcpyAsync(h_out + done_words, ctx->d_out_buf + done_words,
                                       (upto - done_words) * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                       ctx->s_d2h));
                done_words = upto;
            }
        }
        HB_TRY(cudaStreamSynchronize(ctx->s_d2h));
#undef HB_TRY
        if ((rc = job_launched(ctx, st)) != HB_OK) return rc;
        memset(ctx->h_result, 0, kMaxChunks * sizeof(hb::EncResult));
    }
    if (total_bits) *total_bits = bits;
    if (out_bytes) *out_bytes = (bits + 7) / 8;
    return HB_OK;
}

// ---- decoder -----------------------------------------------------------------------
int hb_encode_tile_index(hb_ctx *ctx, uint64_t total_bits, uint64_t *d_tile_bits, void *stream)
{
    if (!ctx || !d_tile_bits) return HB_ERR_ARG;
    if (!ctx->last_job_valid || ctx->pending) return HB_ERR_STATE;      // after hb_encode / hb_encode_result, before the next job
    DeviceGuard g(ctx->device);
    HB_CUDA(ctx, hb::launch_tile_index(ctx->d_tree[ctx->tree_cur], ctx->last_job_tiles, ctx->last_job_start_bit,
                                       ctx->last_job_start_bit + total_bits, (unsigned long long *)d_tile_bits,
                                       (cudaStream_t)stream));
    ctx->launches++;
    return HB_OK;
}

int hb_decode(hb_ctx *ctx, const uint32_t *d_stream, uint64_t stream_words, const uint64_t *d_tile_bits, uint64_t n_words,
              const uint32_t codewords[256], const uint32_t codewordlens[256], uint32_t *d_out, void *stream)
{
    if (!ctx || !codewords || !codewordlens || (n_words && (!d_stream || !d_tile_bits || !d_out))) return HB_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int kLut = 1 << hb::kDecLutBits;
    if (!ctx->d_dec_lut) {
        HB_CUDA(ctx, cudaMalloc(&ctx->d_dec_lut, kLut * sizeof(uint16_t)));
        HB_CUDA(ctx, cudaMalloc(&ctx->d_dec_trie, 1024 * sizeof(int16_t)));
        HB_CUDA(ctx, cudaMalloc(&ctx->d_dec_error, sizeof(unsigned long long)));
        HB_CUDA(ctx, cudaMallocHost(&ctx->h_dec_stage, kLut * sizeof(uint16_t) + 1024 * sizeof(int16_t) + 16));
    }
    // the code trie (every used code must be a leaf: a prefix code) and the table of the codes of up to kDecLutBits bits
    uint16_t *lut = (uint16_t *)ctx->h_dec_stage;
    int16_t *trie = (int16_t *)(lut + kLut);
    unsigned long long *h_err = (unsigned long long *)(trie + 1024);
    HB_CUDA(ctx, cudaStreamSynchronize(st));                  // (the staging block may still feed an earlier decode)
    memset(lut, 0, kLut * sizeof(uint16_t));
    memset(trie, 0, 1024 * sizeof(int16_t));
    int used = 1;
    for (int s = 0; s < 256; s++) {
        const uint32_t l = codewordlens[s];
        if (!l) continue;
        if (l > HB_MAX_CODE_LEN) return HB_ERR_CODELEN;
        if (codewords[s] >> l) return HB_ERR_CODEWORD;
        int cur = 0;
        for (int b = (int)l - 1; b >= 0; b--) {
            const int bit = (int)((codewords[s] >> b) & 1u);
            int16_t &slot = trie[cur * 2 + bit];
            if (b == 0) {
                if (slot != 0) return HB_ERR_CODEWORD;           // not a prefix code
                slot = (int16_t)(-(s + 1));
            } else {
                if (slot < 0) return HB_ERR_CODEWORD;
                if (slot == 0) {
                    if (used >= 512) return HB_ERR_CODEWORD;
                    slot = (int16_t)used++;
                }
                cur = slot;
            }
        }
        if (l <= (uint32_t)hb::kDecLutBits) {
            const uint32_t base = codewords[s] << (hb::kDecLutBits - l), span = 1u << (hb::kDecLutBits - l);
            for (uint32_t i = 0; i < span; i++) lut[base + i] = (uint16_t)(s | (l << 8));
        }
    }
    if (n_words == 0) return HB_OK;
    *h_err = 0;
    HB_CUDA(ctx, cudaMemcpyAsync(ctx->d_dec_lut, lut, kLut * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    HB_CUDA(ctx, cudaMemcpyAsync(ctx->d_dec_trie, trie, 1024 * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    HB_CUDA(ctx, cudaMemsetAsync(ctx->d_dec_error, 0, sizeof(unsigned long long), st));
    HB_CUDA(ctx, hb::launch_decode(d_stream, (const unsigned long long *)d_tile_bits, tiles_of(n_words), n_words * 4, stream_words,
                                   ctx->d_dec_lut, ctx->d_dec_trie, d_out, ctx->d_dec_error, st));
    ctx->launches++;
    HB_CUDA(ctx, cudaMemcpyAsync(h_err, ctx->d_dec_error, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    HB_CUDA(ctx, cudaStreamSynchronize(st));
    return *h_err ? HB_ERR_CODEWORD : HB_OK;                  // a tile did not end where the next one starts: not this stream's tables
}

static hb_ctx *g_default_ctx = nullptr;
static std::mutex g_default_mutex;

int hb_vlc_encode(unsigned int *indata, unsigned int num_elements, unsigned int *outdata,
                  unsigned int *outsize, unsigned int *codewords, unsigned int *codewordlens)
{
    if (!outdata || !outsize || !codewords || !codewordlens) return HB_ERR_ARG;
    std::lock_guard<std::mutex> lock(g_default_mutex);
    if (!g_default_ctx || g_default_ctx->max_words < num_elements) {
        const char *dev = getenv("HB_DEVICE");
        hb_ctx *fresh = nullptr;
        uint64_t want = num_elements > (1u << 20) ? num_elements : (1u << 20);
        int rc = hb_init(&fresh, dev ? atoi(dev) : 0, want);
        if (rc != HB_OK) return rc;
        hb_free(g_default_ctx);
        g_default_ctx = fresh;
    }
    // The reference signature carries no output capacity: outdata must hold floor(bits/32)+1 words
    // (cpuencode.cpp:39).  Use the codebook's worst case as the bound.
    uint32_t max_len = 0;
    for (int s = 0; s < 256; s++)
        if (codewordlens[s] > max_len) max_len = codewordlens[s];
    const uint64_t cap = ((uint64_t)num_elements * 4 * max_len) / 32 + 2;
    uint64_t bytes = 0;
    const int rc = hb_vlc_encode_host(g_default_ctx, indata, num_elements, outdata, cap, codewords,
                                      codewordlens, &bytes, nullptr);
    if (rc == HB_OK) *outsize = (unsigned int)bytes;    // cpuencode.cpp:45 (uint32 bytes)
    return rc;
}

int hb_host_alloc(void **p, uint64_t bytes)
{
    if (!p) return HB_ERR_ARG;
    const cudaError_t e = cudaMallocHost(p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        *p = nullptr;
        return HB_ERR_NOMEM;
    }
    return HB_OK;
}

void hb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int hb_stitch_seam(hb_ctx *ctx, uint32_t *d_dst, const uint32_t *d_src, uint64_t n_words, void *stream)
{
    if (!ctx || (!d_dst && n_words) || (!d_src && n_words)) return HB_ERR_ARG;
    DeviceGuard g(ctx->device);
    HB_CUDA(ctx, hb::launch_or_words(d_dst, d_src, n_words, (cudaStream_t)stream));
    if (n_words) ctx->launches++;
    return HB_OK;
}

int hb_synth_fill(hb_ctx *ctx, uint8_t *d_out, uint64_t first, uint64_t n, uint64_t seed, int mode,
                  int nbits, const uint32_t *thr, int K, const uint8_t *symmap, void *stream)
{
    if (!ctx || (!d_out && n) || !thr || K < 1 || K > 256) return HB_ERR_ARG;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    // thr/symmap are tiny pageable host arrays: the copies below are synchronous w.r.t. the host
    HB_CUDA(ctx, cudaMemcpyAsync(ctx->d_thr, thr, (size_t)K * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (symmap)
        HB_CUDA(ctx, cudaMemcpyAsync(ctx->d_symmap, symmap, (size_t)K, cudaMemcpyHostToDevice, st));
    HB_CUDA(ctx, hb::launch_synth(d_out, first, n, seed, mode, nbits, ctx->d_thr, K,
                                  symmap ? ctx->d_symmap : nullptr, st));
    if (n) ctx->launches++;
    return HB_OK;
}

uint32_t hb_tile_bytes(void) { return (uint32_t)hb::kTileBytes; }

uint64_t hb_launch_count(const hb_ctx *ctx) { return ctx ? ctx->launches : 0; }

const char *hb_encode_variant(const uint32_t codewordlens[256])
{
    if (!codewordlens) return "?";
    for (int s = 0; s < 256; s++)
        if (codewordlens[s] > HB_MAX_CODE_LEN) return "rejected";
    return hb::variant_name(choose_variant(codewordlens));
}

const char *hb_strerror(int status)
{
    switch (status) {
    case HB_OK: return "ok";
    case HB_ERR_ARG: return "bad argument (NULL, size, or alignment)";
    case HB_ERR_CAPACITY: return "output or context capacity too small";
    case HB_ERR_CODELEN: return "codeword length > 31";
    case HB_ERR_CODEWORD: return "codeword has bits above its length";
    case HB_ERR_CUDA: return "CUDA error (see hb_last_cuda_error)";
    case HB_ERR_NOMEM: return "out of memory";
    case HB_ERR_STATE: return "call sequence error";
    case HB_ERR_NCCL: return "NCCL unavailable or an NCCL call failed (see hb_comm_last_nccl_error)";
    default: return "unknown status";
    }
}

int hb_last_cuda_error(const hb_ctx *ctx) { return ctx ? ctx->last_cuda : 0; }

const char *hb_version(void) { return "huffman-b200 0.2 (sm_100a)"; }

}  // extern "C"

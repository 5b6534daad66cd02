/*
 * hb_ctx.h -- the context behind the opaque hb_ctx of include/huffman_b200.h (internal: shared by hb_api.cu and
 * hb_comm.cu, never installed).
 */
#ifndef HB_CTX_H_
#define HB_CTX_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "hb_kernels.cuh"

constexpr int kMaxChunks = 64;                // launches of one chunked host job (hb_vlc_encode_host)

struct hb_ctx {
    int device = 0;
    int sm_count = 0;
    uint64_t max_words = 0;
    uint64_t max_tiles = 0;

    unsigned long long *d_tree[2] = {nullptr, nullptr};   // look-back Fenwick trees, used by alternate jobs
    uint64_t tree_dirty[2] = {0, 0};          // entries a job left non-zero (cleared by the next job's kernel)
    int tree_cur = 0;

    uint32_t *d_table = nullptr;              // 512 words: packed[256] or wide uint2[256]
    uint32_t *h_table = nullptr;              // pinned staging for the table upload
    cudaEvent_t table_uploaded = nullptr;
    uint32_t cw_cache[256];
    uint32_t len_cache[256];
    bool table_valid = false;
    int forced = 0;                           // $HB_FORCE_GROUP the cached table was packed under
    hb::EncVariant variant = {1, false, false};

    hb::EncResult *h_result = nullptr;        // mapped pinned, kMaxChunks slots (+ 1, see below); the kernels write them directly
    // h_result[kMaxChunks].overflow: set by hist_kernel when it refuses its shared-memory layout
    cudaStream_t last_stream = nullptr;       // the stream of the previous job; a job on another stream waits for job_done
    bool have_last_stream = false;
    cudaEvent_t job_done = nullptr;
    uint64_t pending_start_bit = 0;
    bool pending = false;
    bool pending_empty = false;

    unsigned long long *d_hist = nullptr;     // 256 bins
    uint32_t *d_thr = nullptr;                // synth: thresholds
    uint8_t *d_symmap = nullptr;

    // host-buffer pipeline (hb_vlc_encode_host)
    uint32_t *d_in_buf = nullptr;
    uint64_t in_buf_words = 0;
    uint32_t *d_out_buf = nullptr;
    uint64_t out_buf_words = 0;
    cudaStream_t s_main = nullptr;
    cudaStream_t s_d2h = nullptr;
    cudaStream_t s_h2d = nullptr;
    cudaEvent_t ev_chunk[kMaxChunks] = {};            // one per launch of a chunked host job
    cudaEvent_t ev_h2d[kMaxChunks] = {};              // ... and one per input copy

    // decoder (hb_decode): device copies of the decode table and the code trie, an error counter; the job the trees
    // still describe (hb_encode_tile_index)
    uint16_t *d_dec_lut = nullptr;
    int16_t *d_dec_trie = nullptr;
    unsigned long long *d_dec_error = nullptr;
    void *h_dec_stage = nullptr;              // pinned staging for the two tables + the error word
    uint64_t last_job_tiles = 0, last_job_words = 0, last_job_start_bit = 0;
    bool last_job_valid = false;

    uint32_t next_seam_flags = 0;             // for the next hb_encode_async only (set by hb_shard_encode_direct_async)

    unsigned long long *d_prof = nullptr;     // $HB_PROFILE: kernel cycle counters, dumped by hb_free
    uint64_t launches = 0;
    int last_cuda = 0;
};

#endif

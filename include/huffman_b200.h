/*
 * huffman_b200.h -- C ABI of libhuffb200.so: the B200-native (sm_100a) Huffman variable-length
 * ENCODE hot path, a drop-in for the reference's histogram -> codebook -> encode -> scan -> pack
 * sequence (vlnguyen92/Huffman-GPU, main_test_cu.cu:85-170).
 *
 * Plain C: pointers and sizes only, no CUDA or torch types (a stream is passed as the opaque
 * cudaStream_t value, `void *`; NULL = the legacy default stream).  Every function returns an
 * hb_status (0 = ok, negative = error); nothing ever calls exit() (the reference does:
 * cutil.h:781-787, hist.cu:19-27).
 *
 * Conventions shared with the reference (cpuencode.h:4-7, cpuencode.cpp:12-46):
 *   - input  = symbols packed 4 per little-endian uint32, consumed MOST SIGNIFICANT BYTE FIRST;
 *   - tables = codewords[256] (right-aligned values) and codewordlens[256] (bits), host memory;
 *   - output = one contiguous bitstream in uint32 words, stream bit j = bit 31-(j%32) of word j/32,
 *              unused low bits of the last word zero, bit-identical to cpu_vlc_encode.
 * Parity domain: codeword lengths 0..31 and codewords[s] < 2^codewordlens[s] (cpuencode.cpp:34
 * is undefined for length 32 and does not mask straddling codewords).
 * Anything else is rejected with HB_ERR_CODELEN / HB_ERR_CODEWORD.
 *
 * There is NO CPU fallback: every data-path entry point fails with HB_ERR_CUDA when no sm_100
 * device / kernel image is available.
 */
#ifndef HUFFMAN_B200_H_
#define HUFFMAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* libhuffb200.so is built with -fvisibility=hidden: only the functions declared here are exported */
#if defined(__GNUC__)
#define HB_API __attribute__((visibility("default")))
#else
#define HB_API
#endif

#define HB_NUM_SYMBOLS 256          /* parameters.h:25 NUM_SYMBOLS */
#define HB_MAX_CODE_LEN 31          /* parity domain of cpu_vlc_encode */

typedef enum {
    HB_OK = 0,
    HB_ERR_ARG = -1,        /* NULL pointer, bad size, misaligned device pointer */
    HB_ERR_CAPACITY = -2,   /* output (or context scratch) too small */
    HB_ERR_CODELEN = -3,    /* a used codeword length is > 31 (or a histogram needs > 31 bits) */
    HB_ERR_CODEWORD = -4,   /* codewords[s] has bits set at or above codewordlens[s] */
    HB_ERR_CUDA = -5,       /* CUDA runtime error; see hb_last_cuda_error() */
    HB_ERR_NOMEM = -6,
    HB_ERR_STATE = -7,      /* call sequence error (e.g. hb_encode_result without hb_encode_async) */
    HB_ERR_NCCL = -8        /* NCCL not available in this process, or an NCCL call failed (hb_comm_last_nccl_error) */
} hb_status;

typedef struct hb_ctx hb_ctx;

/* ---- lifecycle (replaces InitCUDA, cuda_helpers.h:11-38, and the cudaMalloc block of
 *      runVLCTest, main_test_cu.cu:93-110; scan.cu:67-112 preallocBlockSums/deallocBlockSums).
 * The context owns only scratch: tile descriptors for inputs up to `max_words` uint32 words, the
 * device copy of the codebook, a pinned result block.  One context per (host thread, device);
 * calls on a context are ordered on the stream they are given; a context serialises its jobs (they share its
 * scratch), so a job given another stream than the previous one is made to wait for it.  No hidden globals.
 * max_words is limited to 2^21 encode tiles (64 GiB of input per job; HB_ERR_CAPACITY beyond). */
HB_API int hb_init(hb_ctx **ctx, int device, uint64_t max_words);
HB_API void hb_free(hb_ctx *ctx);

/* ---- byte histogram of a device buffer (replaces runHisto/histo_kernel, hist.cu:34-125, minus
 *      the file read and minus runHisto's window bug: this counts ALL 4*n_words bytes).
 * d_in: device pointer, 4-byte aligned.  hist: HOST array, 64-bit counts.  Synchronises `stream`. */
HB_API int hb_histogram(hb_ctx *ctx, const uint32_t *d_in, uint64_t n_words, uint64_t hist[256],
                 void *stream);
/* Same, but ADDS into a DEVICE array of 256 uint64 and does not synchronise (multi-GPU path: the
 * caller all-reduces d_hist with NCCL before building the codebook). */
HB_API int hb_histogram_device(hb_ctx *ctx, const uint32_t *d_in, uint64_t n_words, uint64_t *d_hist,
                        void *stream);

/* ---- host-side codebook (replaces BuildTree + GenerateCodes, huffTree.h:55-94, and the table
 *      flatten of loadData, load_data.h:40-47).  Pure function, no device work.  Tie-breaking is
 *      identical to std::priority_queue<INode*, vector<INode*>, NodeCmp>; weights are 64-bit.
 * Returns the maximum code length (>= 0) or HB_ERR_CODELEN if a code would exceed 31 bits. */
HB_API int hb_build_codebook(const uint64_t hist[256], uint32_t codewords[256],
                      uint32_t codewordlens[256]);

/* sum_s hist[s] * codewordlens[s]: the exact output size in bits, known before encoding
 * (used to derive per-shard start bits without an extra pass). */
HB_API uint64_t hb_bits_from_hist(const uint64_t hist[256], const uint32_t codewordlens[256]);

/* ---- single-pass encode (replaces vlc_encode_kernel_sm64huff + prescanArray + cudaMemset +
 *      pack2, main_test_cu.cu:142-166, and is bit-exact with cpu_vlc_encode, cpuencode.cpp:12-46).
 * d_in:   device, 32-byte aligned, n_words uint32 (4 symbols each).
 * d_out:  device, 4-byte aligned, capacity in words.  No pre-zeroing needed.  Words
 *         [start_bit/32, ceil((start_bit+bits)/32)) are written; like the reference
 *         (cpuencode.cpp:39) one extra zero word is written after a word-aligned end when it
 *         fits.  The `start_bit % 32` leading bits of the first word are written as ZERO (a shard
 *         seam is OR-ed by the caller, see hb_stitch_seam).
 * start_bit: global bit position of this stream's first bit (0 for a single-GPU encode).
 * total_bits: HOST pointer; receives the number of bits produced (excluding start_bit).
 * hb_encode synchronises `stream`; hb_encode_async does not (fetch with hb_encode_result). */
HB_API int hb_encode(hb_ctx *ctx, const uint32_t *d_in, uint64_t n_words, const uint32_t codewords[256],
              const uint32_t codewordlens[256], uint32_t *d_out, uint64_t out_capacity_words,
              uint64_t start_bit, uint64_t *total_bits, void *stream);
HB_API int hb_encode_async(hb_ctx *ctx, const uint32_t *d_in, uint64_t n_words,
                    const uint32_t codewords[256], const uint32_t codewordlens[256],
                    uint32_t *d_out, uint64_t out_capacity_words, uint64_t start_bit,
                    void *stream);
HB_API int hb_encode_result(hb_ctx *ctx, uint64_t *total_bits, void *stream);

/* ---- host-buffer entry points (the call a reference user makes today) ------------------------
 * hb_vlc_encode has EXACTLY the signature and semantics of the reference's
 *   extern "C" void cpu_vlc_encode(unsigned int *indata, unsigned int num_elements,
 *        unsigned int *outdata, unsigned int *outsize, unsigned int *codewords,
 *        unsigned int *codewordlens)                                   (cpuencode.h:4-7)
 * with HOST pointers: outsize is in BYTES = ceil(bits/8) (cpuencode.cpp:44-45) and outdata needs
 * floor(bits/32)+1 words.  It runs H2D -> GPU encode -> D2H on a lazily created per-process
 * context (device = $HB_DEVICE or 0) and returns an hb_status instead of void.
 * hb_vlc_encode_host is the same with 64-bit sizes, an explicit context and capacity, and chunked
 * copy/encode overlap; buffers from hb_host_alloc (pinned) are copied by DMA directly. */
HB_API int hb_vlc_encode(unsigned int *indata, unsigned int num_elements, unsigned int *outdata,
                  unsigned int *outsize, unsigned int *codewords, unsigned int *codewordlens);
HB_API int hb_vlc_encode_host(hb_ctx *ctx, const uint32_t *h_in, uint64_t n_words, uint32_t *h_out,
                       uint64_t out_capacity_words, const uint32_t codewords[256],
                       const uint32_t codewordlens[256], uint64_t *out_bytes,
                       uint64_t *total_bits);
HB_API int hb_host_alloc(void **p, uint64_t bytes);      /* pinned host memory */
HB_API void hb_host_free(void *p);

/* ---- multi-GPU (no reference equivalent: hist.cu:67 hard-codes device 0) ------------
 * One process (or host thread) per GPU, each with its own hb_ctx and hb_comm; every rank holds one contiguous
 * shard of the input (boundaries on multiples of hb_tile_bytes(), except the stream's ragged end).  NCCL over
 * NVLink carries exactly two tiny collectives -- the all-reduce of the 256-bin histogram and the all-gather of
 * one uint64 per rank (the shard's bit total, known before encoding) -- and every rank then encodes its shard
 * already in global bit phase.  The data path has no collective.  NCCL is bound at run time (the libnccl.so.2
 * the process already has, else the system one; $HB_NCCL_LIB overrides); without it these calls return HB_ERR_NCCL.
 *
 *   rank 0: hb_comm_unique_id(id); the caller hands the 128 bytes to every rank (any channel it has)
 *   all   : hb_comm_init(&comm, ctx, rank, n, id)          -- or hb_comm_adopt() of an existing ncclComm_t
 *   all   : hb_shard_plan_build(...)                        -- histogram -> all-reduce -> codebook -> all-gather
 *   all   : hb_shard_encode_async(...), hb_shard_encode_result(...)
 *   (opt) : hb_stitch_open(...), hb_stitch_push(...)        -- one stream on the root GPU, bit-exact with the
 *                                                              single-GPU stream and with cpu_vlc_encode
 */
#define HB_UNIQUE_ID_BYTES 128      /* sizeof(ncclUniqueId) */
#define HB_MAX_RANKS 64
typedef struct hb_comm hb_comm;

typedef struct {
    int32_t rank, n_ranks;
    int32_t max_len;                /* longest code of the (global) codebook */
    uint32_t phase;                 /* start_bit % 32: the start_bit argument of this shard's encode */
    uint64_t shard_bits;            /* bits this shard produces: sum_s hist_local[s] * codewordlens[s] */
    uint64_t start_bit;             /* global bit offset of this shard = sum of the lower ranks' shard_bits */
    uint64_t total_bits;            /* of the whole stream */
    uint64_t first_word;            /* start_bit / 32: where this shard's words go in the stitched stream */
    uint64_t local_words;           /* words the shard's encode writes: ceil((phase + shard_bits) / 32), >= 1;
                                       give the local buffer local_offset_words + local_words + 1 words */
    uint32_t local_offset_words;    /* first_word % 4: the shard's words start this far into the local buffer, so
                                       that they have the same 16-byte alignment as their place in the stream */
    uint32_t reserved;
} hb_shard_plan;

HB_API int hb_comm_unique_id(uint8_t id[HB_UNIQUE_ID_BYTES]);
HB_API int hb_comm_init(hb_comm **comm, hb_ctx *ctx, int rank, int n_ranks, const uint8_t id[HB_UNIQUE_ID_BYTES]);
/* the same on a communicator the caller already has (an ncclComm_t of the same NCCL library, passed as void *) */
HB_API int hb_comm_adopt(hb_comm **comm, hb_ctx *ctx, void *nccl_comm, int rank, int n_ranks);
HB_API void hb_comm_free(hb_comm *comm);                   /* collective when a stitch buffer is open */
HB_API int hb_comm_last_nccl_error(const hb_comm *comm);   /* ncclResult_t of the last HB_ERR_NCCL */

/* Collective.  Histogram of this rank's shard (device buffer) -> ncclAllReduce -> the identical codebook on every
 * rank (codewords / codewordlens, HOST arrays, out) -> ncclAllGather of the shard bit totals -> *plan.
 * hist_global (HOST, optional) receives the all-reduced histogram.  Synchronises `stream` (twice: the codebook is
 * built on the host, as in the reference). */
HB_API int hb_shard_plan_build(hb_comm *comm, const uint32_t *d_in, uint64_t n_words, uint32_t codewords[256],
                        uint32_t codewordlens[256], uint64_t hist_global[256], hb_shard_plan *plan,
                        void *stream);
/* every rank's offsets of the last plan (arrays of n_ranks, either may be NULL) */
HB_API int hb_comm_plan_offsets(const hb_comm *comm, uint64_t *start_bits, uint64_t *shard_bits);
/* hb_encode_async / hb_encode_result of this rank's shard in global phase: the words land at
 * d_local + plan->local_offset_words.  Not collective. */
HB_API int hb_shard_encode_async(hb_comm *comm, const uint32_t *d_in, uint64_t n_words, const uint32_t codewords[256],
                          const uint32_t codewordlens[256], uint32_t *d_local, uint64_t local_capacity_words,
                          const hb_shard_plan *plan, void *stream);
HB_API int hb_shard_encode_result(hb_comm *comm, uint64_t *shard_bits, void *stream);

/* Optional stitch: the shards gathered into one stream on GPU `root` over NVLink, without a funnel -- the root's
 * buffer is shared through CUDA IPC and every rank stores its own words straight to their final place (peer
 * stores), all ranks at once; a seam word is written by the lowest rank with bits in it, OR-ed with the head
 * words of the others.  All three calls are collective.
 * hb_stitch_open: the root allocates capacity_words (>= total_bits / 32 + 1); *d_stitched = the stream on the root
 * (its peer mapping elsewhere).  hb_stitch_push: asynchronous on `stream`; when the root's stream has drained, the
 * stream is complete (words [0, total_bits / 32], incl. the zero word after a word-aligned end). */
HB_API int hb_stitch_open(hb_comm *comm, uint64_t capacity_words, int root, uint32_t **d_stitched, void *stream);
HB_API int hb_stitch_push(hb_comm *comm, const uint32_t *d_local, const hb_shard_plan *plan, void *stream);
HB_API int hb_stitch_close(hb_comm *comm);
/* The fused form of encode + stitch (collective; needs hb_stitch_open): the shard is encoded STRAIGHT into the root's
 * stream -- the kernel's copy-out stores go to peer memory over NVLink, already in global phase; no local output buffer,
 * no second pass.  The two words a shard may share with its neighbours are OR-ed into words the root has zeroed.
 * Asynchronous on `stream`; fetch the bit count with hb_shard_encode_result.  When the root's stream has drained, the
 * stream is complete. */
HB_API int hb_shard_encode_direct_async(hb_comm *comm, const uint32_t *d_in, uint64_t n_words,
                                 const uint32_t codewords[256], const uint32_t codewordlens[256],
                                 const hb_shard_plan *plan, void *stream);

/* building blocks of the above for callers that bring their own exchange:
 * hb_shard_offsets: exclusive prefix of per-shard bit totals -> start_bit of every shard. */
HB_API int hb_shard_offsets(const uint64_t *shard_bits, int n_shards, uint64_t *start_bits,
                     uint64_t *total_bits);
/* OR `n_words` words of d_src into d_dst (both device, same GPU): merges the seam word(s) that two
 * neighbouring shards both own after a peer-to-peer gather. */
HB_API int hb_stitch_seam(hb_ctx *ctx, uint32_t *d_dst, const uint32_t *d_src, uint64_t n_words,
                   void *stream);

/* ---- decoder (no reference equivalent: the reference cannot decode) -----------------
 * A tile-parallel decoder, so that encode -> decode round trips can be checked on the device at any size.  Not part
 * of the measured hot path.
 * hb_encode_tile_index: the bit offset of every encode tile (hb_tile_bytes() of input each) of the LAST job of this
 * context, n_tiles + 1 device uint64 (the last entry = the end of the stream); call it after hb_encode /
 * hb_encode_result and before the next encode (the offsets come from the job's look-back tree).
 * hb_decode: d_stream (stream_words words readable) -> the n_words input words into d_out, bit-exact with the
 * encoder's input.  The tables must form a prefix code (HB_ERR_CODEWORD otherwise, and when the stream does not
 * decode tile by tile to exactly the indexed offsets); symbols with length 0 cannot occur.  Synchronises `stream`. */
HB_API int hb_encode_tile_index(hb_ctx *ctx, uint64_t total_bits, uint64_t *d_tile_bits, void *stream);
HB_API int hb_decode(hb_ctx *ctx, const uint32_t *d_stream, uint64_t stream_words, const uint64_t *d_tile_bits,
              uint64_t n_words, const uint32_t codewords[256], const uint32_t codewordlens[256], uint32_t *d_out,
              void *stream);

/* ---- tooling -----------------------------------------------------------------------------------
 * Deterministic synthetic input generator on the device (the reference's
 * testdatagen.h:62-67 cannot control entropy).  Byte i of the stream, i in [first, first+n):
 *   mode 0: u = splitmix64(seed + (i+1)*0x9E3779B97F4A7C15) >> 32
 *   mode 1: u = perm_nbits(i, seed), a bijection on [0, 2^nbits)
 *   sym = first k < K-1 with u < thr[k], else K-1;  byte = symmap ? symmap[sym] : sym.
 * thr / symmap are HOST arrays (K <= 256). */
HB_API int hb_synth_fill(hb_ctx *ctx, uint8_t *d_out, uint64_t first, uint64_t n, uint64_t seed, int mode,
                  int nbits, const uint32_t *thr, int K, const uint8_t *symmap, void *stream);

/* Input bytes per encode tile (the unit of the look-back and of the work distribution): shard boundaries on
 * multiples of it keep every shard's tiles full.  A property of the build, not of a context. */
HB_API uint32_t hb_tile_bytes(void);
/* Kernel launches issued through this context so far (bench.py's gpu_launches counter). */
HB_API uint64_t hb_launch_count(const hb_ctx *ctx);
/* Name of the encode kernel variant chosen for a codebook ("packed_g2", "wide_g1", ...). */
HB_API const char *hb_encode_variant(const uint32_t codewordlens[256]);
HB_API const char *hb_strerror(int status);
HB_API int hb_last_cuda_error(const hb_ctx *ctx);        /* cudaError_t of the last HB_ERR_CUDA */
HB_API const char *hb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HUFFMAN_B200_H_ */

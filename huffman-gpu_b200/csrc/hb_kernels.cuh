/*
 * hb_kernels.cuh -- internal interface between the C-ABI layer (hb_api.cu) and the sm_100a kernels.
 */
#ifndef HB_KERNELS_CUH_
#define HB_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

// ---- geometry of the single-pass encoder ---------------------------------------------------------
// One persistent CTA per SM: kEncWorkers worker warps + a publisher warp + kEncResolvers resolver warps.  A tile is the CTA's unit of
// work and of the look-back; a warp chunk (1/kEncWorkers of a tile) is a worker warp's unit.
#ifndef HB_WORKERS
#define HB_WORKERS 16                  // (8: the warp-count experiment of round 2)
#endif
constexpr int kEncWorkers = HB_WORKERS;
#ifndef HB_RESOLVERS
#define HB_RESOLVERS 2
#endif
constexpr int kEncResolvers = HB_RESOLVERS;                           // resolver warps: each takes every kEncResolvers-th tile
constexpr int kEncThreads = (kEncWorkers + 1 + kEncResolvers) * 32;
#ifndef HB_SYM_PER_THREAD
#define HB_SYM_PER_THREAD 64
#endif
constexpr int kSymPerThread = HB_SYM_PER_THREAD;                                     // symbols (bytes) per lane per chunk
constexpr int kChunkBytes = 32 * kSymPerThread;                       // one warp: 2 KiB
constexpr int kTileBytes = kEncWorkers * kChunkBytes;                 // 32 KiB
constexpr int kTileWords = kTileBytes / 4;
// A look-back tree node counts tiles in 22 bits and sums bits in 42 (hb_encode.cu): a job has at most 2^21 tiles
// (64 GiB of input; a node then counts at most 2^20 tiles and sums at most 64 GiB * 31 bits < 2^42).
constexpr unsigned long long kMaxJobTiles = 1ULL << 21;
constexpr int kTreeCountShift = 42;                                   // node = [63:42] tiles counted | [41:0] bits summed

// Written by the kernel into mapped pinned host memory (zero-copy), read by the host after sync.
struct EncResult {
    unsigned long long bits_end;      // global bit position after the last tile of the launch
    unsigned long long overflow;      // != 0: the output did not fit out_cap_words
};

struct EncParams {
    const uint32_t *in;               // base of the job's symbol buffer (32-byte aligned)
    unsigned long long n_words;       // words in the whole job
    unsigned long long n_tiles;       // tiles in the whole job
    unsigned long long first_tile;    // this launch encodes tiles [first_tile, end_tile)
    unsigned long long end_tile;
    uint32_t *out;
    unsigned long long out_cap_words;
    unsigned long long start_bit;     // global bit position of the job's first bit
    unsigned long long *tree;         // Fenwick tree over the job's tile bit counts, 1-based, n_tiles entries
    unsigned long long *tree_zero;    // the tree of the next job: entries [0, zero_count) are cleared
    unsigned long long zero_count;
    const uint32_t *table;            // packed: uint32[256]; wide: uint2[256] as uint32[512]
    EncResult *result;
    unsigned long long *prof;         // optional cycle counters ($HB_PROFILE), else nullptr
    uint32_t l2_prefetch;             // pull the tile after next into L2 (on by default; $HB_L2_PREFETCH=0)
    uint32_t seam_flags;              // kSeam*: the output is a stream shared with neighbouring shards (hb_comm.cu)
};

// seam_flags: the first output word is shared with the previous shard (OR it), so is the final partial word with the next
// (OR it), the word after a word-aligned end belongs to the next shard (do not write the courtesy zero word)
constexpr uint32_t kSeamFirst = 1u, kSeamLast = 2u, kSeamNoZeroWord = 4u;

// Encode kernel variants.
//   "packed": table entry = (cw << (32-len)) | len, needs every len <= 24;
//   "wide":   entry = {cw << (32-len), len}, len <= 31.
//   G = symbols whose codewords are chained in one 32-bit register before the word-boundary test.
//   A group must stay below 32 bits; when G * max_len > 31 that is checked at run time (CHECK) and
//   a warp that meets such a group re-encodes its chunk symbol by symbol.
struct EncVariant {
    int group;        // G in {1, 2, 3, 4, 6, 8}
    bool wide;
    bool check;
};

// Picks the variant from the code lengths alone (2^-len is the symbol probability a Huffman code
// implies): the largest G whose chance of an over-long group is negligible.
EncVariant pick_variant(const uint32_t codewordlens[256]);
const char *variant_name(const EncVariant &v);
size_t encode_smem_bytes(const EncVariant &v);
cudaError_t encode_configure();                        // opt-in smem attributes, once per process
cudaError_t launch_encode(const EncVariant &v, const EncParams &p, int grid, cudaStream_t stream);

cudaError_t histogram_configure();                     // opt-in smem attribute, once per device context
cudaError_t launch_histogram(const uint32_t *d_in, unsigned long long n_words,
                             unsigned long long *d_hist, int sm_count, unsigned long long *refused,
                             cudaStream_t stream);
cudaError_t launch_or_words(uint32_t *d_dst, const uint32_t *d_src, unsigned long long n_words,
                            cudaStream_t stream);
// decoder (hb_decode.cu): per-tile bit offsets from the look-back tree an encode left behind,
// and the tile-parallel table/trie decoder
cudaError_t launch_tile_index(const unsigned long long *d_tree, unsigned long long n_tiles, unsigned long long start_bit,
                              unsigned long long end_bit, unsigned long long *d_tile_bits, cudaStream_t stream);
cudaError_t launch_decode(const uint32_t *d_stream, const unsigned long long *d_tile_bits, unsigned long long n_tiles,
                          unsigned long long n_symbols, unsigned long long n_stream_words, const uint16_t *d_lut,
                          const int16_t *d_trie, uint32_t *d_out_words, unsigned long long *d_error, cudaStream_t stream);
constexpr int kDecLutBits = 10;                                       // primary decode table: 2^10 entries of {symbol, length}
cudaError_t launch_synth(uint8_t *d_out, unsigned long long first, unsigned long long n,
                         unsigned long long seed, int mode, int nbits, const uint32_t *d_thr, int K,
                         const uint8_t *d_symmap, cudaStream_t stream);

}  // namespace hb
#endif

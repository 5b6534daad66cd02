/*
 * hb_decode.cu -- tile-parallel Huffman DECODER for sm_100a.
 *
 * The reference has no decoder at all (its only check is CPU-encode vs GPU-encode, main_test_cu.cu:170-171); this one
 * exists so that streams can be proven decodable at full size on the device: encode -> decode must give the input back,
 * a size-independent parity property next to the comparison with cpu_vlc_encode.  It is not on the measured hot path.
 *
 * A Huffman stream is not self-synchronising in any cheap way, but the encoder already knows where every 32 KiB tile of
 * input starts in the stream: the look-back tree of an encode job holds the bit count of every tile.  tile_index_kernel
 * turns the tree into an array of per-tile bit offsets (a prefix query per tile: one node per set bit of the tile number,
 * exactly what the encoder's resolver reads); decode_kernel then gives every thread one tile: a 64-bit bit buffer, a
 * 2^10-entry table {symbol, length} for codes of up to 10 bits, a bit-serial walk of the code trie for longer ones,
 * four symbols per 32-bit store in the reference's order (first symbol = most significant byte, cpuencode.cpp:28).
 */
#include "hb_kernels.cuh"

namespace hb {
namespace {

constexpr unsigned long long kSumMask = (1ULL << kTreeCountShift) - 1ULL;
constexpr int kTileSymbols = kTileBytes;
constexpr int kDecThreads = 64;

__global__ void tile_index_kernel(const unsigned long long *__restrict__ tree, unsigned long long n_tiles,
                                  unsigned long long start_bit, unsigned long long end_bit,
                                  unsigned long long *__restrict__ tile_bits)
{
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if (t == n_tiles) {
        tile_bits[t] = end_bit;                               // (nodes at or beyond the last tile are never written)
        return;
    }
    // prefix [0, t): for every set bit b of t, the node t with the bits below b cleared (it covers 2^b tiles)
    unsigned long long sum = 0;
    for (unsigned long long rest = t; rest; rest &= rest - 1ULL) sum += tree[rest] & kSumMask;
    tile_bits[t] = start_bit + sum;
}

// lut[i] (i = the next kDecLutBits stream bits): symbol | length << 8, or 0 when the code is longer than kDecLutBits.
// trie[node][bit] = child node (> 0), or -(symbol + 1) for a leaf, or 0 for a dead prefix.
__global__ void __launch_bounds__(kDecThreads) decode_kernel(const uint32_t *__restrict__ stream,
                                                             const unsigned long long *__restrict__ tile_bits,
                                                             unsigned long long n_tiles, unsigned long long n_symbols,
                                                             unsigned long long n_stream_words,
                                                             const uint16_t *__restrict__ g_lut, const int16_t *__restrict__ g_trie,
                                                             uint32_t *__restrict__ out_words, unsigned long long *error)
{
    __shared__ uint16_t lut[1 << kDecLutBits];
    __shared__ int16_t trie[512 * 2];
    for (int i = threadIdx.x; i < (1 << kDecLutBits); i += kDecThreads) lut[i] = g_lut[i];
    for (int i = threadIdx.x; i < 1024; i += kDecThreads) trie[i] = g_trie[i];
    __syncthreads();
    const unsigned long long t = (unsigned long long)blockIdx.x * kDecThreads + threadIdx.x;
    if (t >= n_tiles) return;
    const unsigned long long first = t * (unsigned long long)kTileSymbols;
    const unsigned long long count = n_symbols - first < (unsigned long long)kTileSymbols ? n_symbols - first : kTileSymbols;
    unsigned long long bit = tile_bits[t];
    const unsigned long long end = tile_bits[t + 1];
    // 64-bit buffer, the next stream bit in its most significant position; `avail` valid bits
    const uint32_t *wp = stream + (bit >> 5);
    unsigned long long buf = ((unsigned long long)wp[0] << 32) << (bit & 31u);
    int avail = 32 - (int)(bit & 31u);
    wp++;
    uint32_t *out = out_words + first / 4;
    bool bad = false;
    for (unsigned long long s = 0; s < count; s += 4) {
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (avail <= 32) {                                 // refill: one whole word fits below the valid bits
                const uint32_t next = wp < stream + n_stream_words ? *wp : 0u;     // (never read past the stream)
                wp++;
                buf |= (unsigned long long)next << (32 - avail);
                avail += 32;
            }
            uint32_t sym, len;
            const uint32_t e = lut[(uint32_t)(buf >> (64 - kDecLutBits))];
            if (e) {
                sym = e & 0xFFu;
                len = e >> 8;
            } else {
                int node = 0;
                len = 0;
                do {
                    node = trie[node * 2 + (int)((buf >> (63 - len)) & 1ULL)];
                    len++;
                } while (node > 0 && len < 32u);
                if (node >= 0) {                               // a dead prefix, or no leaf within 31 bits
                    bad = true;
                    node = -1;
                }
                sym = (uint32_t)(-node - 1);
            }
            buf <<= len;
            avail -= (int)len;
            bit += len;
            word |= sym << (24 - 8 * j);
        }
        out[s / 4] = word;
    }
    if (bad || bit != end) atomicAdd(error, 1ULL);            // every tile must end exactly where the next one starts
}

}  // namespace

cudaError_t launch_tile_index(const unsigned long long *d_tree, unsigned long long n_tiles, unsigned long long start_bit,
                              unsigned long long end_bit, unsigned long long *d_tile_bits, cudaStream_t stream)
{
    const unsigned long long n = n_tiles + 1;
    tile_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_tree, n_tiles, start_bit, end_bit, d_tile_bits);
    return cudaGetLastError();
}

cudaError_t launch_decode(const uint32_t *d_stream, const unsigned long long *d_tile_bits, unsigned long long n_tiles,
                          unsigned long long n_symbols, unsigned long long n_stream_words, const uint16_t *d_lut,
                          const int16_t *d_trie, uint32_t *d_out_words, unsigned long long *d_error, cudaStream_t stream)
{
    if (n_tiles == 0) return cudaSuccess;
    decode_kernel<<<(unsigned)((n_tiles + kDecThreads - 1) / kDecThreads), kDecThreads, 0, stream>>>(
        d_stream, d_tile_bits, n_tiles, n_symbols, n_stream_words, d_lut, d_trie, d_out_words, d_error);
    return cudaGetLastError();
}

}  // namespace hb

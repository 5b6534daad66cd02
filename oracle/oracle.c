/*
 * oracle.c -- CPU restatement of the reference's Huffman encode hot path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Parity status: PINNED against the unmodified
 * reference (oracle/_ref/libref.so, built by oracle/Makefile) and tests/golden/.
 * Nothing here is linked into, loaded by, or called from libhuffb200.so.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * cpuencode.cpp:12-46
 * ---------------------------------------------------------------------------------- */
int orc_vlc_encode(const uint32_t *in, uint64_t n_words, uint32_t *out,
                   uint64_t *out_bytes, uint64_t *total_bits,
                   const uint32_t *codewords, const uint32_t *codewordlens)
{
    uint32_t *pt = out;          /* cpuencode.cpp:16 bitstreamPt */
    uint32_t startbit = 0;       /* cpuencode.cpp:18 */
    uint64_t bytes = 0;          /* cpuencode.cpp:19 totalBytes (64-bit here) */
    uint64_t bits = 0;
    *pt = 0;                     /* cpuencode.cpp:17 */

    for (uint64_t k = 0; k < n_words; k++) {
        uint32_t val32 = in[k];
        for (unsigned i = 0; i < 4; i++) {
            /* cpuencode.cpp:28 -- most significant byte of the word is the first symbol */
            unsigned symbol = (val32 >> (8 * (3 - i))) & 0xFFu;
            uint32_t cw32 = codewords[symbol];
            uint32_t numbits = codewordlens[symbol];
            if (numbits > 31)
                return -1;       /* cpuencode.cpp:34 would evaluate 1<<32 */
            bits += numbits;
            while (numbits > 0) {                                   /* cpuencode.cpp:32 */
                uint32_t room = 32 - startbit;
                uint32_t writebits = room < numbits ? room : numbits;
                uint32_t mask32;
                if (numbits == writebits)                           /* cpuencode.cpp:34 */
                    mask32 = (cw32 & ((1u << numbits) - 1u)) << (32 - startbit - numbits);
                else                                                /* cpuencode.cpp:35 */
                    mask32 = cw32 >> (numbits - writebits);
                *pt |= mask32;                                      /* cpuencode.cpp:36 */
                numbits -= writebits;
                startbit = (startbit + writebits) % 32;
                if (startbit == 0) {                                /* cpuencode.cpp:39 */
                    pt++;
                    *pt = 0;
                    bytes += 4;
                }
            }
        }
    }
    bytes += (startbit / 8) + ((startbit % 8 == 0) ? 0 : 1);       /* cpuencode.cpp:44 */
    if (out_bytes) *out_bytes = bytes;
    if (total_bits) *total_bits = bits;
    return 0;
}

void orc_cpu_vlc_encode(unsigned int *indata, unsigned int num_elements,
                        unsigned int *outdata, unsigned int *outsize,
                        unsigned int *codewords, unsigned int *codewordlens)
{
    uint64_t bytes = 0;
    (void)orc_vlc_encode(indata, num_elements, outdata, &bytes, NULL, codewords, codewordlens);
    *outsize = (unsigned int)bytes;                                 /* cpuencode.cpp:45 */
}

/* ------------------------------------------------------------------------------------
 * hist.cu:34-52, without runHisto's window bug
 * ---------------------------------------------------------------------------------- */
void orc_histogram(const uint8_t *data, uint64_t n, uint64_t hist[256])
{
    memset(hist, 0, 256 * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++)
        hist[data[i]]++;
}

/* ------------------------------------------------------------------------------------
 * huffTree.h:55-94 + load_data.h:40-47
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int64_t f;      /* huffTree.h:22 (int there) */
    int left;       /* huffTree.h:31 */
    int right;      /* huffTree.h:32 */
    int sym;        /* huffTree.h:45; -1 for internal nodes */
} orc_node;

/* NodeCmp, huffTree.h:50-53: comp(lhs, rhs) = lhs->f > rhs->f */
static int orc_cmp(const orc_node *nd, int lhs, int rhs) { return nd[lhs].f > nd[rhs].f; }

/* libstdc++ bits/stl_heap.h __push_heap */
static void orc_push_heap(int *heap, int hole, int top, int value, const orc_node *nd)
{
    int parent = (hole - 1) / 2;
    while (hole > top && orc_cmp(nd, heap[parent], value)) {
        heap[hole] = heap[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    heap[hole] = value;
}

/* libstdc++ bits/stl_heap.h __adjust_heap */
static void orc_adjust_heap(int *heap, int hole, int len, int value, const orc_node *nd)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (orc_cmp(nd, heap[child], heap[child - 1]))
            child--;
        heap[hole] = heap[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        heap[hole] = heap[child - 1];
        hole = child - 1;
    }
    orc_push_heap(heap, hole, top, value, nd);
}

/* priority_queue::pop = pop_heap + pop_back; returns the popped top */
static int orc_pq_pop(int *heap, int *size, const orc_node *nd)
{
    int top = heap[0];
    int n = *size;
    if (n > 1) {
        int value = heap[n - 1];          /* __pop_heap: value = *result; *result = *first */
        heap[n - 1] = heap[0];
        orc_adjust_heap(heap, 0, n - 1, value, nd);
    }
    *size = n - 1;
    return top;
}

static void orc_pq_push(int *heap, int *size, int value, const orc_node *nd)
{
    heap[*size] = value;
    (*size)++;
    orc_push_heap(heap, *size - 1, 0, value, nd);
}

int orc_build_codebook(const uint64_t hist[256], uint32_t codewords[256],
                       uint32_t codewordlens[256])
{
    orc_node nd[511];
    int heap[256];
    int size = 0, nn = 0;

    memset(codewords, 0, 256 * sizeof(uint32_t));
    memset(codewordlens, 0, 256 * sizeof(uint32_t));

    for (int i = 0; i < 256; i++) {                       /* huffTree.h:59-63 */
        if (hist[i] != 0) {
            nd[nn].f = (int64_t)hist[i];
            nd[nn].left = nd[nn].right = -1;
            nd[nn].sym = i;
            orc_pq_push(heap, &size, nn, nd);
            nn++;
        }
    }
    if (size == 0)
        return 0;                                          /* reference: top() on empty queue (UB) */
    while (size > 1) {                                     /* huffTree.h:64-74 */
        int childR = orc_pq_pop(heap, &size, nd);
        int childL = orc_pq_pop(heap, &size, nd);
        nd[nn].f = nd[childR].f + nd[childL].f;            /* huffTree.h:34 */
        nd[nn].left = childR;                              /* InternalNode(childR, childL): c0 -> left */
        nd[nn].right = childL;
        nd[nn].sym = -1;
        orc_pq_push(heap, &size, nn, nd);
        nn++;
    }

    /* GenerateCodes (huffTree.h:78-94): DFS, left appends 0, right appends 1.  The code
     * value is kept as a 64-bit integer with the root edge as MSB, which is what the
     * flatten loop of load_data.h:40-47 produces. */
    int stack_node[512];
    uint64_t stack_code[512];
    int stack_len[512];
    int sp = 0, maxlen = 0;
    stack_node[0] = heap[0]; stack_code[0] = 0; stack_len[0] = 0; sp = 1;
    while (sp > 0) {
        sp--;
        int n = stack_node[sp];
        uint64_t c = stack_code[sp];
        int l = stack_len[sp];
        if (nd[n].sym >= 0) {
            if (l > 32)
                return -1;
            codewords[nd[n].sym] = (uint32_t)c;
            codewordlens[nd[n].sym] = (uint32_t)l;
            if (l > maxlen) maxlen = l;
        } else {
            stack_node[sp] = nd[n].right; stack_code[sp] = (c << 1) | 1u; stack_len[sp] = l + 1; sp++;
            stack_node[sp] = nd[n].left;  stack_code[sp] = (c << 1);      stack_len[sp] = l + 1; sp++;
        }
    }
    return maxlen;
}

/* ------------------------------------------------------------------------------------ */
uint64_t orc_word_fnv(const uint32_t *words, uint64_t n_words)
{
    uint64_t h = 1469598103934665603ULL;
    for (uint64_t i = 0; i < n_words; i++) {
        h ^= words[i];
        h *= 1099511628211ULL;
    }
    return h;
}

/* ------------------------------------------------------------------------------------
 * decoder (no reference equivalent)
 * ---------------------------------------------------------------------------------- */
uint64_t orc_vlc_decode(const uint32_t *stream, uint64_t bit0, uint64_t n_symbols,
                        uint8_t *out_file_order, const uint32_t *codewords,
                        const uint32_t *codewordlens)
{
    /* binary trie: child[node][bit]; leaf symbol stored as -(sym+2); 0 = absent */
    int (*child)[2] = calloc(2 * 256 * 32 + 2, sizeof(*child));
    int nn = 1, zero_sym = -1;
    if (!child) return (uint64_t)-1;
    for (int s = 0; s < 256; s++) {
        uint32_t l = codewordlens[s];
        if (l == 0) { if (zero_sym < 0) zero_sym = s; continue; }
        int n = 0;
        for (uint32_t b = 0; b < l; b++) {
            int bit = (codewords[s] >> (l - 1 - b)) & 1;
            if (b == l - 1) {
                child[n][bit] = -(s + 2);
            } else {
                if (child[n][bit] <= 0) child[n][bit] = nn++;
                n = child[n][bit];
            }
        }
    }
    uint64_t pos = bit0;
    for (uint64_t i = 0; i < n_symbols; i++) {
        int n = 0, sym = -1;
        if (nn == 1 && child[0][0] == 0 && child[0][1] == 0) {
            sym = zero_sym;               /* single-symbol alphabet: length-0 code */
        } else {
            for (;;) {
                int bit = (stream[pos >> 5] >> (31 - (pos & 31))) & 1;
                pos++;
                int c = child[n][bit];
                if (c == 0) { free(child); return (uint64_t)-1; }
                if (c < 0) { sym = -c - 2; break; }
                n = c;
            }
        }
        if (sym < 0) { free(child); return (uint64_t)-1; }
        /* symbol i is byte (3 - i%4) of little-endian word i/4 (cpuencode.cpp:28) */
        out_file_order[(i & ~(uint64_t)3) + (3 - (i & 3))] = (uint8_t)sym;
    }
    free(child);
    return pos;
}

/* ------------------------------------------------------------------------------------
 * synthetic inputs
 * ---------------------------------------------------------------------------------- */
static inline uint64_t orc_mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline uint32_t orc_perm(uint64_t i, uint64_t seed, int nbits)
{
    const uint64_t mask = (nbits >= 64) ? ~0ULL : ((1ULL << nbits) - 1);
    uint64_t x = (i + seed) & mask;
    x = (x * 0x9E3779B97F4A7C15ULL) & mask;
    x ^= x >> (nbits / 2 + 1);
    x = (x * 0xBF58476D1CE4E5B9ULL) & mask;
    x ^= x >> (nbits / 2);
    x = (x * 0x94D049BB133111EBULL) & mask;
    x ^= x >> (nbits / 2 + 2);
    return (uint32_t)x;
}

void orc_synth_fill(uint8_t *out, uint64_t first, uint64_t n, uint64_t seed, int mode,
                    int nbits, const uint32_t *thr, int K, const uint8_t *symmap)
{
    for (uint64_t j = 0; j < n; j++) {
        uint64_t i = first + j;
        uint32_t u = (mode == 0)
                         ? (uint32_t)(orc_mix64(seed + (i + 1) * 0x9E3779B97F4A7C15ULL) >> 32)
                         : orc_perm(i, seed, nbits);
        /* first k in [0, K-1) with u < thr[k]; else K-1 */
        int lo = 0, hi = K - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (u < thr[mid]) hi = mid; else lo = mid + 1;
        }
        out[j] = symmap ? symmap[lo] : (uint8_t)lo;
    }
}

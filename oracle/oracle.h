/*
 * oracle.h -- CPU restatement of the reference's Huffman encode hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load liboracle.so.  The product
 * library (libhuffb200.so) never links, loads or calls anything declared here.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle.py) against
 *   (a) the unmodified reference sources compiled into oracle/_ref/libref.so by
 *       oracle/Makefile (cpuencode.cpp, huffTree.h), when /root/reference is present, and
 *   (b) the golden vectors of SURVEY.md section 8c committed under tests/golden/.
 *
 * Each function cites the reference file:line it follows (paths relative to the
 * reference checkout, vlnguyen92/Huffman-GPU).
 */
#ifndef HB_ORACLE_H_
#define HB_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cpuencode.cpp:12-46 (cpu_vlc_encode).  Serial MSB-first bit packer into uint32 words.
 * Symbols are taken MSB-byte-first from each uint32 input word (cpuencode.cpp:28).
 * Writes floor(bits/32)+1 words (the word after a word-aligned end is zeroed,
 * cpuencode.cpp:39).  *out_bytes = ceil(bits/8) as a 64-bit value (the reference
 * truncates it to uint32, cpuencode.cpp:44-45); *total_bits is the exact bit count.
 * Returns 0, or -1 if any used codeword length is > 31 (outside the parity domain:
 * cpuencode.cpp:34 evaluates 1<<32, undefined behaviour). */
int orc_vlc_encode(const uint32_t *in, uint64_t n_words, uint32_t *out,
                   uint64_t *out_bytes, uint64_t *total_bits,
                   const uint32_t *codewords, const uint32_t *codewordlens);

/* Same argument list as the reference entry point (cpuencode.h:4-7), for A/B runs
 * against oracle/_ref's cpu_vlc_encode. */
void orc_cpu_vlc_encode(unsigned int *indata, unsigned int num_elements,
                        unsigned int *outdata, unsigned int *outsize,
                        unsigned int *codewords, unsigned int *codewordlens);

/* Plain byte count over all n bytes.  This is what hist.cu:34-52 (histo_kernel)
 * computes per launch; runHisto's windowing bug (hist.cu:98-102) is NOT reproduced
 * (SURVEY.md section 8 a-2). */
void orc_histogram(const uint8_t *data, uint64_t n, uint64_t hist[256]);

/* huffTree.h:55-76 (BuildTree: std::priority_queue with NodeCmp, leaves pushed in
 * symbol order, first-popped child is `left`), huffTree.h:78-94 (GenerateCodes:
 * left=0, right=1, root edge first) and load_data.h:40-47 (flatten: first edge is the
 * MSB of a right-aligned `len`-bit value).  The heap follows libstdc++'s
 * push_heap/pop_heap exactly (bits/stl_heap.h __push_heap/__adjust_heap) so ties are
 * broken like the reference.  Weights are 64-bit (the reference's `int f`,
 * huffTree.h:22, overflows above INT_MAX).  Returns the maximum code length, or -1 if
 * a code would be longer than 32 bits.  Absent symbols get (0,0); a single present
 * symbol gets length 0; an all-zero histogram gives all-zero tables. */
int orc_build_codebook(const uint64_t hist[256], uint32_t codewords[256],
                       uint32_t codewordlens[256]);

/* SURVEY.md section 8c hash: h=1469598103934665603; for each word: h^=w; h*=1099511628211. */
uint64_t orc_word_fnv(const uint32_t *words, uint64_t n_words);

/* Bit-serial prefix decoder (no reference equivalent: SURVEY.md section 8 f-4).  Decodes
 * exactly n_symbols from the MSB-first uint32 word stream starting at bit `bit0` and
 * stores them as bytes in decode order (= file order shuffled back through the
 * MSB-byte-first rule, i.e. out[] is the original uint32-packed input viewed as bytes).
 * Returns the bit position after the last symbol, or (uint64_t)-1 on a dead prefix. */
uint64_t orc_vlc_decode(const uint32_t *stream, uint64_t bit0, uint64_t n_symbols,
                        uint8_t *out_file_order, const uint32_t *codewords,
                        const uint32_t *codewordlens);

/* Deterministic synthetic inputs.  The reference has no usable
 * generator (testdatagen.h:62-67 only draws uniform words).  Byte i of the stream is
 *   mode 0 (iid):   u = splitmix64(seed + i) >> 32;  sym = first k with u < thr[k], else K-1
 *   mode 1 (exact): u = perm_nbits(i, seed) (a bijection on [0, 2^nbits)); same search,
 *                   thr[] = exact cumulative counts, so symbol k occurs exactly
 *                   thr[k]-thr[k-1] times over the 2^nbits positions.
 * out[j] = symmap[sym] for i = first+j. */
void orc_synth_fill(uint8_t *out, uint64_t first, uint64_t n, uint64_t seed, int mode,
                    int nbits, const uint32_t *thr, int K, const uint8_t *symmap);

#ifdef __cplusplus
}
#endif
#endif

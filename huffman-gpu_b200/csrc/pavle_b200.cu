/*
 * pavle_b200.cu -- command-line driver with the call sequence and the output fields of the reference's `pavle`
 * (main_test_cu.cu:41-180 main/runVLCTest, load_data.h:8-58 loadData; `./pavle data/test1024_H2.206587175259.in`),
 * on top of the C ABI of libhuffb200.so.  SURVEY.md section 8 f-1.
 *
 *   pavle_b200 <file> [--repeats N] [--no-check] [--cpu-lib <shared library exporting cpu_vlc_encode>]
 *
 * file -> device -> byte histogram (hb_histogram) -> Huffman codebook (hb_build_codebook) -> single-pass encode
 * (hb_encode_async, mean of N launches between CUDA events, as the reference times NT = 10 launches,
 * main_test_cu.cu:117,141-156) -> "GPU Encoded to %d [B]" -> PASS!/FAIL!.
 *
 * The reference decides PASS/FAIL by comparing with its CPU encoder (compare_vectors, main_test_cu.cu:171).  This
 * driver does not carry a second encoder: it DECODES the GPU stream on the host with the same tables and compares the
 * symbols with the file (a round trip), and checks the bit count against sum(hist[s] * len[s]).  Exit status 0 = PASS.
 * There is no CPU encode path here: without an sm_100 device hb_init fails and so does the driver.
 *
 * --cpu-lib: the reference's two CPU fields ("CPU Encoding time (CPU)", "CPU Encoded to %d [B]", main_test_cu.cu:120-125)
 * and its own verdict, compare_vectors over ceil(bytes/4) words (main_test_cu.cu:126,171; comparison_helpers.h:5-16), for
 * callers who have a cpu_vlc_encode at hand (e.g. the reference's cpuencode.cpp built as a shared library).  The
 * library is dlopen()ed at run time and never linked: the product carries no CPU encoder.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "../../include/huffman_b200.h"

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                       \
            return 3;                                                                         \
        }                                                                                     \
    } while (0)
#define HB(call)                                                                              \
    do {                                                                                      \
        int rc_ = (call);                                                                     \
        if (rc_ != HB_OK) {                                                                   \
            fprintf(stderr, "%s: %s\n", #call, hb_strerror(rc_));                             \
            return 3;                                                                         \
        }                                                                                     \
    } while (0)

/* Binary trie over the codebook: node[i][bit] = child index, or -(symbol + 1) for a leaf. */
static int build_trie(const uint32_t cw[256], const uint32_t len[256], int (*node)[2], int cap)
{
    int used = 1;
    node[0][0] = node[0][1] = 0;
    for (int s = 0; s < 256; s++) {
        if (!len[s]) continue;
        int cur = 0;
        for (int b = (int)len[s] - 1; b >= 0; b--) {
            const int bit = (cw[s] >> b) & 1;
            if (b == 0) {
                if (node[cur][bit] != 0) return -1;              /* not a prefix code */
                node[cur][bit] = -(s + 1);
            } else {
                if (node[cur][bit] < 0) return -1;
                if (node[cur][bit] == 0) {
                    if (used >= cap) return -1;
                    node[used][0] = node[used][1] = 0;
                    node[cur][bit] = used++;
                }
                cur = node[cur][bit];
            }
        }
    }
    return used;
}

/* Decodes `n_sym` symbols from the MSB-first word stream and compares them with the input words (whose bytes are
 * consumed most significant first, cpuencode.cpp:28).  Returns the number of bits consumed, or -1 on a mismatch. */
static long long decode_and_compare(const uint32_t *stream, uint64_t total_bits, const uint32_t *in_words,
                                    uint64_t n_sym, int (*node)[2])
{
    uint64_t bit = 0;
    for (uint64_t i = 0; i < n_sym; i++) {
        int cur = 0;
        for (;;) {
            if (bit >= total_bits) return -1;
            const int b = (stream[bit >> 5] >> (31 - (bit & 31))) & 1;
            bit++;
            cur = node[cur][b];
            if (cur < 0) break;
            if (cur == 0) return -1;
        }
        const uint32_t want = (in_words[i >> 2] >> (24 - 8 * (i & 3))) & 0xFFu;
        if ((uint32_t)(-cur - 1) != want) return -1;
    }
    return (long long)bit;
}

int main(int argc, char **argv)
{
    const char *path = NULL, *cpu_lib = NULL;
    int repeats = 10, check = 1;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--repeats") && i + 1 < argc)
            repeats = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--cpu-lib") && i + 1 < argc)
            cpu_lib = argv[++i];
        else if (!strcmp(argv[i], "--no-check"))
            check = 0;
        else if (argv[i][0] != '-' && !path)
            path = argv[i];
        else
            path = NULL, i = argc;
    }
    if (!path || repeats < 1) {
        printf("No input file\n");                               /* load_data.h:27 */
        fprintf(stderr, "usage: pavle_b200 <file> [--repeats N] [--no-check] [--cpu-lib lib.so]   (%s)\n", hb_version());
        return 2;
    }

    /* ---- load (load_data.h:8-31): the file as uint32 words; a ragged tail of < 4 bytes is dropped as there ---- */
    FILE *f = fopen(path, "rb");
    if (!f) {
        perror(path);
        return 2;
    }
    fseek(f, 0, SEEK_END);
    const long fsize = ftell(f);
    fseek(f, 0, SEEK_SET);
    const uint64_t n_words = (uint64_t)fsize / 4;
    const uint64_t mem_size = n_words * 4;
    uint32_t *source = NULL;
    CK(cudaMallocHost(&source, mem_size ? mem_size : 4));
    if (mem_size && fread(source, 1, mem_size, f) != mem_size) {
        fprintf(stderr, "%s: short read\n", path);
        return 2;
    }
    fclose(f);
    printf("CUDA! Starting VLC Tests!\n");                       /* main_test_cu.cu:53 */

    uint32_t *d_in = NULL;
    CK(cudaMalloc(&d_in, mem_size ? mem_size : 4));
    CK(cudaMemcpy(d_in, source, mem_size, cudaMemcpyHostToDevice));

    hb_ctx *ctx = NULL;
    HB(hb_init(&ctx, 0, n_words));

    /* ---- histogram, entropy, codebook (load_data.h:33-56) ---- */
    uint64_t hist[256];
    uint32_t codewords[256], codewordlens[256];
    HB(hb_histogram(ctx, d_in, n_words, hist, NULL));
    double H = 0.0;
    for (int s = 0; s < 256; s++)
        if (hist[s]) {
            const double pr = (double)hist[s] / (double)mem_size;
            H -= pr * log2(pr);
        }
    const int max_len = hb_build_codebook(hist, codewords, codewordlens);     /* >= 0, or an error */
    if (max_len < 0) {
        fprintf(stderr, "hb_build_codebook: %s\n", hb_strerror(max_len));
        return 3;
    }
    printf("\n%s, %llu bytes, entropy %f\n\n", path, (unsigned long long)mem_size, H);
    const uint64_t bits_expected = hb_bits_from_hist(hist, codewordlens);
    printf("Parameters: num_elements: %llu, max code length: %d, tile: %u bytes, kernel: %s\n----------------------------\n",
           (unsigned long long)n_words, max_len, hb_tile_bytes(), hb_encode_variant(codewordlens));

    const uint64_t cap_words = bits_expected / 32 + 2;

    /* ---- optional: the caller's CPU encoder, timed as the reference times it (main_test_cu.cu:32-36,120-125) ---- */
    uint32_t *cref = NULL;
    unsigned int refbytesize = 0;
    if (cpu_lib) {
        typedef void (*cpu_fn)(unsigned int *, unsigned int, unsigned int *, unsigned int *, unsigned int *, unsigned int *);
        void *h = dlopen(cpu_lib, RTLD_NOW | RTLD_LOCAL);
        cpu_fn cpu = h ? (cpu_fn)dlsym(h, "cpu_vlc_encode") : NULL;
        if (!cpu) {
            fprintf(stderr, "%s: no cpu_vlc_encode (%s)\n", cpu_lib, dlerror());
            return 2;
        }
        if (n_words >> 32) {
            fprintf(stderr, "cpu_vlc_encode counts words in 32 bits (cpuencode.h:4-7): input too large for --cpu-lib\n");
            return 2;
        }
        cref = (uint32_t *)calloc(cap_words + 1, 4);
        if (!cref) return 3;
        struct timeval tv;
        gettimeofday(&tv, NULL);
        const long long t0 = tv.tv_sec * 1000000LL + tv.tv_usec;
        cpu((unsigned int *)source, (unsigned int)n_words, cref, &refbytesize, codewords, codewordlens);
        gettimeofday(&tv, NULL);
        const float msec = (float)((tv.tv_sec * 1000000LL + tv.tv_usec - t0) / 1000.0);
        printf("CPU Encoding time (CPU): %f (ms)\n", msec);
        printf("CPU Encoded to %d [B]\n", refbytesize);
    }

    /* ---- encode: warm-up, then the mean of `repeats` launches between events (main_test_cu.cu:136-156) ---- */
    uint32_t *d_out = NULL;
    CK(cudaMalloc(&d_out, cap_words * 4));
    uint64_t bits = 0;
    HB(hb_encode(ctx, d_in, n_words, codewords, codewordlens, d_out, cap_words, 0, &bits, NULL));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, NULL));
    for (int r = 0; r < repeats; r++)
        HB(hb_encode_async(ctx, d_in, n_words, codewords, codewordlens, d_out, cap_words, 0, NULL));
    CK(cudaEventRecord(e1, NULL));
    HB(hb_encode_result(ctx, &bits, NULL));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("GPU Encoding time (B200 single pass): %f (ms)\n", ms / (float)repeats);
    printf("GPU Encoded to %llu [B]\n", (unsigned long long)((bits + 7) / 8));
    if (ms > 0.f)
        printf("GPU Encoding rate: %f (GB/s of input)\n", (double)mem_size * repeats / ((double)ms * 1e-3) / 1e9);

    /* ---- PASS / FAIL (main_test_cu.cu:171 compares with the CPU encoder; here: bit count + round trip) ---- */
    int ok = bits == bits_expected;
    if (ok && check && bits) {
        uint32_t *stream = (uint32_t *)malloc(cap_words * 4);
        int(*node)[2] = (int(*)[2])malloc(sizeof(int[2]) * 512);
        if (!stream || !node) return 3;
        CK(cudaMemcpy(stream, d_out, cap_words * 4, cudaMemcpyDeviceToHost));
        if (build_trie(codewords, codewordlens, node, 512) < 0) {
            fprintf(stderr, "the codebook is not a prefix code: round trip skipped\n");
        } else {
            const long long used = decode_and_compare(stream, bits, source, mem_size, node);
            ok = used >= 0 && (uint64_t)used == bits;
            /* the padding of the last word must be zero (cpuencode.cpp:39-41) */
            if (ok && (bits & 31)) ok = (stream[bits >> 5] & (0xFFFFFFFFu >> (bits & 31))) == 0;
        }
        free(stream);
        free(node);
    }
    if (ok && cref) {
        /* compare_vectors(crefData, destData, num_ints) with num_ints = ceil(refbytesize / 4): main_test_cu.cu:126,171 */
        const uint64_t num_ints = (bits + 31) / 32;              /* (refbytesize itself wraps at 4 GiB, cpuencode.cpp:45) */
        uint32_t *stream = (uint32_t *)malloc(cap_words * 4);
        if (!stream) return 3;
        CK(cudaMemcpy(stream, d_out, cap_words * 4, cudaMemcpyDeviceToHost));
        ok = refbytesize == (unsigned int)((bits + 7) / 8);
        for (uint64_t i = 0; ok && i < num_ints; i++)
            if (stream[i] != cref[i]) {
                printf("Error at word %llu: CPU %08x GPU %08x\n", (unsigned long long)i, cref[i], stream[i]);
                ok = 0;
            }
        free(stream);
        if (ok) printf("PASS! vectors are matching!\n");        /* comparison_helpers.h:13 */
    }
    printf(ok ? "PASS!\n" : "FAIL!\n");
    free(cref);

    cudaFree(d_out);
    cudaFree(d_in);
    cudaFreeHost(source);
    hb_free(ctx);
    return ok ? 0 : 1;
}

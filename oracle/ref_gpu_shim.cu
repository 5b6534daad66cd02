/*
 * ref_gpu_shim.cu -- runs the UNMODIFIED reference GPU pipeline (vlc_encode_kernel_sm64huff -> prescanArray ->
 * cudaMemset -> pack2) as a measured comparator.  TEST / BENCH INFRASTRUCTURE ONLY: nothing here is linked into
 * libhuffb200.so.  The reference kernels are #included from the reference checkout where they lie (-I$(REF));
 * no reference source is copied into this repository.  Only the launch sequence of runVLCTest
 * (main_test_cu.cu:130-168) is restated here, because main_test_cu.cu itself pulls in load_data.h, which does not
 * compile (missing ';' at load_data.h:28) and reads its input from a file.
 *
 * Validity limits of the reference path: 4 codewords <= 64 bits, <= 8192 bits
 * per 1024-symbol block, total output < 2^32 bits, num_elements a multiple of 256*16.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "parameters.h"
#include "vlc_kernel_sm64huff.cu"
#include "scan.cu"
#include "pack_kernels.cu"

extern "C" int ref_gpu_pipeline(const unsigned int *d_source, unsigned int num_elements,
                                const unsigned int *h_codewords, const unsigned int *h_codewordlens,
                                unsigned int *d_packed, unsigned long long packed_bytes, unsigned int *total_bits,
                                float *ms_encode, float *ms_scan, float *ms_memset_pack, int repeats)
{
    const unsigned int num_block_threads = 256;                      /* main_test_cu.cu:43 */
    if (num_elements == 0 || num_elements % (num_block_threads * 16) != 0) return -1;   /* :166, load_data.h:20 */
    const unsigned int num_blocks = num_elements / num_block_threads;
    const size_t mem_size = (size_t)num_elements * sizeof(unsigned int);
    if (packed_bytes < mem_size) return -2;                          /* the reference sizes it like the input, :95 */

    unsigned int *d_dest = 0, *d_cw = 0, *d_cl = 0, *d_cindex = 0, *d_cindex2 = 0;
    if (cudaMalloc(&d_dest, mem_size) != cudaSuccess) return -3;
    cudaMalloc(&d_cw, NUM_SYMBOLS * sizeof(unsigned int));
    cudaMalloc(&d_cl, NUM_SYMBOLS * sizeof(unsigned int));
    cudaMalloc(&d_cindex, num_blocks * sizeof(unsigned int));
    cudaMalloc(&d_cindex2, num_blocks * sizeof(unsigned int));
    cudaMemcpy(d_cw, h_codewords, NUM_SYMBOLS * sizeof(unsigned int), cudaMemcpyHostToDevice);
    cudaMemcpy(d_cl, h_codewordlens, NUM_SYMBOLS * sizeof(unsigned int), cudaMemcpyHostToDevice);
    cudaMemset(d_dest, 0, mem_size);                                 /* :110 uploads a zeroed destData */

    dim3 grid_size(num_blocks, 1, 1), block_size(num_block_threads, 1, 1);
    /* CACHECWLUT is defined (parameters.h:16): main_test_cu.cu:132-134 */
    const unsigned int sm_size = 2 * NUM_SYMBOLS * sizeof(int) + block_size.x * sizeof(unsigned int);

    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    float t_enc = 0, t_scan = 0, t_pack = 0;
    preallocBlockSums(num_blocks);                                   /* :160 */
    for (int it = 0; it < repeats; it++) {
        cudaEventRecord(e0, 0);
        /* :142-146; the TESTING-only cw32/cw32len/cw32idx arguments are never dereferenced by the kernel */
        vlc_encode_kernel_sm64huff<<<grid_size, block_size, sm_size>>>((unsigned int *)d_source, d_cw, d_cl,
                                                                       (unsigned int *)0, (unsigned int *)0,
                                                                       (unsigned int *)0, d_dest, d_cindex);
        cudaEventRecord(e1, 0);
        prescanArray(d_cindex2, d_cindex, num_blocks);               /* :164 */
        cudaEventRecord(e2, 0);
        cudaMemset(d_packed, 0, mem_size);                           /* :162 */
        pack2<<<num_blocks / 16, 16>>>(d_dest, d_cindex, d_cindex2, d_packed, num_elements / num_blocks);   /* :166 */
        cudaEventRecord(e3, 0);
        cudaEventSynchronize(e3);
        float a, b, c;
        cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2); cudaEventElapsedTime(&c, e2, e3);
        t_enc += a; t_scan += b; t_pack += c;
    }
    deallocBlockSums();                                              /* :168 */
    const cudaError_t err = cudaDeviceSynchronize();
    unsigned int last_off = 0, last_bits = 0;
    cudaMemcpy(&last_off, d_cindex2 + num_blocks - 1, sizeof(unsigned int), cudaMemcpyDeviceToHost);
    cudaMemcpy(&last_bits, d_cindex + num_blocks - 1, sizeof(unsigned int), cudaMemcpyDeviceToHost);
    if (total_bits) *total_bits = last_off + last_bits;
    if (ms_encode) *ms_encode = t_enc / repeats;
    if (ms_scan) *ms_scan = t_scan / repeats;
    if (ms_memset_pack) *ms_memset_pack = t_pack / repeats;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    cudaFree(d_dest); cudaFree(d_cw); cudaFree(d_cl); cudaFree(d_cindex); cudaFree(d_cindex2);
    return err == cudaSuccess ? 0 : -(int)err - 100;
}

// tools/ubench.cu -- issue-rate microbenchmark for the integer instructions of the encode kernel's hot loops (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu && build/ubench
// Every test runs W warps per SM sub-partition (4 sub-partitions per SM, one CTA per SM), each warp executing N
// independent-chain instructions of one kind; reported: SMSP cycles per warp-instruction (1.0 = full issue rate,
// 2.0 = a half-rate pipe).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 256, kChains = 8;

template <int OP>
__device__ __forceinline__ void op(uint32_t &a, uint32_t b, uint32_t c, uint32_t &l)
{
    if (OP == 0) asm volatile("prmt.b32 %0, %0, %1, 0x6504;" : "+r"(a) : "r"(b));                      // PRMT reg,reg,imm
    if (OP == 1) asm volatile("shf.l.wrap.b32 %0, %1, %0, %1;" : "+r"(a) : "r"(b));                    // SHF.L.W 3 regs (append)
    if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));            // LOP3
    if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(a ^ b));                           // IADD3 / IMAD.IADD (ptxas picks)
    if (OP == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));                // IMAD 3 regs
    if (OP == 5) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c));              // IDP.4A
    if (OP == 6) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a) : "r"(b));                     // SHF imm shift
    if (OP == 7) asm volatile("mad.lo.u32 %0, %0, 5, %1;" : "+r"(a) : "r"(b));                         // IMAD reg,imm,reg
    if (OP == 8) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));                  // PRMT reg selector
    if (OP == 9) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(b));                            // IMAD.HI
    if (OP == 10) {                                                                                      // the pass-1 mix without the load
        uint32_t t;
        asm volatile("prmt.b32 %0, %1, %2, 0x6504;" : "=r"(t) : "r"(a), "r"(b));
        asm volatile("shf.l.wrap.b32 %0, %1, %0, %1;" : "+r"(a) : "r"(t));
        asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(l) : "r"(t), "r"(b));
    }
}

template <int OP>
__global__ void bench(unsigned long long *out, uint32_t seed)
{
    uint32_t r[kChains], l[kChains];
    for (int i = 0; i < kChains; i++) r[i] = seed * (i + 1) + threadIdx.x, l[i] = i;
    const uint32_t b = seed | 1u, c = seed ^ 0x55u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < kChains; i++) op<OP>(r[i], b, c, l[i]);
    }
    const long long t1 = clock64();
    uint32_t x = 0;
    for (int i = 0; i < kChains; i++) x ^= r[i] ^ l[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
    if (x == 0x12345678u) out[1] = x;
}

template <int OP>
void run(const char *name, int insts_per_op)
{
    unsigned long long *d, h[2];
    cudaMalloc(&d, 16);
    for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
        bench<OP><<<148, warps_per_smsp * 4 * 32>>>(d, 12345u);
        cudaDeviceSynchronize();
        bench<OP><<<148, warps_per_smsp * 4 * 32>>>(d, 12345u);
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        const double n = (double)kIters * 4 * kChains * insts_per_op * warps_per_smsp;   // warp-instructions per SMSP
        printf("%-34s %d warp(s)/SMSP: %.2f cycles per warp-instruction\n", name, warps_per_smsp, (double)h[0] / n);
    }
    cudaFree(d);
}

int main()
{
    run<0>("PRMT r,r,imm", 1);
    run<8>("PRMT r,r,r", 1);
    run<1>("SHF.L.W r,r,r (append)", 1);
    run<6>("SHF.R.W r,r,imm", 1);
    run<2>("LOP3 r,r,r", 1);
    run<3>("add.u32 r,r", 1);
    run<4>("IMAD r,r,r", 1);
    run<7>("IMAD r,imm,r", 1);
    run<9>("IMAD.HI r,r", 1);
    run<5>("IDP.4A r,r,r", 1);
    run<10>("PRMT+SHF+IDP4A (pass 1 w/o LDS)", 3);
    return 0;
}

// Pipe-rate microbenchmark for the integer ops the stepping kernels are made of (sm_100a).
// Reports warp-instructions per clock per SM for: SHF (funnel shift, ALU pipe), LOP3 (ALU), IMAD (FMA pipe),
// IMAD.WIDE.U32 (FMA pipe, 64-bit result) and ALU/FMA mixes.  nvcc -arch=sm_100a -O3 pipes.cu -o pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096, CH = 8;

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint32_t *out, uint32_t seed, long long *clk, uint32_t mulc) {
    uint32_t a[CH], b[CH];
    uint64_t w[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = seed + i * 77 + threadIdx.x; b[i] = seed * 3 + i; w[i] = a[i]; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) asm volatile("shf.l.wrap.b32 %0, %1, %2, 10;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]));
            if (MODE == 1) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]), "r"(seed));
            if (MODE == 2) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]), "r"(seed));
            if (MODE == 3) asm volatile("mad.wide.u32 %0, %1, %3, %2;" : "=l"(w[i]) : "r"(a[i]), "l"(w[i]), "r"(mulc));
            if (MODE == 4) { // 1 SHF : 1 IMAD
                asm volatile("shf.l.wrap.b32 %0, %1, %2, 10;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]));
                asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(b[i]) : "r"(b[i]), "r"(seed), "r"(seed));
            }
            if (MODE == 5) { // 1 LOP3 : 1 IMAD.WIDE
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]), "r"(seed));
                asm volatile("mad.wide.u32 %0, %1, %3, %2;" : "=l"(w[i]) : "r"(b[i]), "l"(w[i]), "r"(mulc));
            }
            if (MODE == 6) { // 2 LOP3 : 1 IMAD.WIDE
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]), "r"(seed));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(b[i]) : "r"(a[i]), "r"(b[i]), "r"(seed));
                asm volatile("mad.wide.u32 %0, %1, %3, %2;" : "=l"(w[i]) : "r"(b[i]), "l"(w[i]), "r"(mulc));
            }
            if (MODE == 7) { // umulhi (IMAD.HI)
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]));
            }
            if (MODE == 8) { // popc
                asm volatile("popc.b32 %0, %1;" : "=r"(a[i]) : "r"(a[i]));
            }
            if (MODE == 9) { // 1 LOP3 : 1 IMAD.HI
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]), "r"(seed));
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(b[i]) : "r"(b[i]), "r"(seed));
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int MODE>
void run(const char *name, int per_iter) {
    uint32_t *out; long long *clk, h;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 8);
    k<MODE><<<148, 512>>>(out, 12345u, clk, 1024u);
    k<MODE><<<148, 512>>>(out, 12345u, clk, 1024u);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    double winst = (double)ITER * CH * per_iter * 16; // warp-instructions per SM
    printf("%-28s %6.3f warp-inst/clk/SM  (%5.3f per SMSP)\n", name, winst / h, winst / h / 4);
    cudaFree(out); cudaFree(clk);
}

int main() {
    run<0>("SHF", 1); run<1>("LOP3", 1); run<2>("IMAD", 1); run<3>("IMAD.WIDE.U32", 1);
    run<4>("SHF + IMAD", 2); run<5>("LOP3 + IMAD.WIDE", 2); run<6>("2 LOP3 + IMAD.WIDE", 3);
    run<7>("IMAD.HI", 1); run<8>("POPC", 1); run<9>("LOP3 + IMAD.HI", 2);
    return 0;
}

// Issue-rate microbenchmark for the integer forms the coder kernels use (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void k(uint32_t *out, uint32_t a, uint32_t b, long long *cyc)
{
    uint32_t r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = threadIdx.x * 8 + j + a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int jj = 0; jj < 64; jj++) {
            const int j = jj & 7;
            if (OP == 0) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(a));
            if (OP == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(a));
            if (OP == 2) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(a));
            if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[j]) : "r"(b), "r"(a));
            if (OP == 4) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(a));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[j]) : "r"(b), "r"(a)); }
            if (OP == 5) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(a));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[j]) : "r"(b), "r"(a));
                           asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(a));
                           asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[j]) : "r"(b), "r"(a)); }
            if (OP == 7) asm volatile("vabsdiff.u32.u32.u32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(a));
            if (OP == 8) asm volatile("min.u32 %0, %0, %1;" : "+r"(r[j]) : "r"(b));
            if (OP == 9) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(a));
            if (OP == 10) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[j]) : "r"(b));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= r[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char *name, int per)
{
    uint32_t *out; long long *cyc;
    int blocks = 148, threads = 1024;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    k<OP><<<blocks, threads>>>(out, 3, 5, cyc); cudaDeviceSynchronize();
    k<OP><<<blocks, threads>>>(out, 3, 5, cyc); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < blocks; i++) c += h[i]; c /= blocks;
    double winst = (double)N * 64 * per * (threads / 32);
    printf("%-28s %.2f warp-inst/clk/SM  (%.2f clk per warp-inst per SMSP)\n", name, winst / c, c / (winst / 4));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0>("mad.hi.u32 (IMAD.HI)", 1);
    run<1>("mad.lo.u32 (IMAD)", 1);
    run<2>("shf.r.clamp (SHF)", 1);
    run<3>("lop3 (LOP3)", 1);
    run<10>("add (IADD3/VIADD)", 1);
    run<7>("vabsdiff", 1);
    run<8>("min.u32", 1);
    run<9>("prmt", 1);
    run<4>("IMAD + LOP3 alternating", 2);
    run<5>("IMAD.HI + LOP3 + SHF + LOP3", 4);
    return 0;
}

// tools/ubench.cu - issue-rate micro-benchmarks for the integer/video instructions K1 leans on
// (SURVEY.md section 7 "ALU vs HBM": VABSDIFF4 issue rate on B200 is unknown - measure first).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096
__device__ __forceinline__ uint32_t sad4(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d; asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
template <int OP>
__global__ void k(uint32_t *out, uint32_t seed) {
    uint32_t a0 = threadIdx.x * 0x01010101u + seed, a1 = a0 ^ 0x55aa55aau, a2 = a0 + 0x01020304u, a3 = ~a0;
    uint32_t b = seed * 0x9E3779B9u + threadIdx.x, c0 = 0, c1 = 1, c2 = 2, c3 = 3;
    __shared__ uint32_t sm[1024];
    sm[threadIdx.x & 1023] = a0;
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (OP == 0) { c0 = sad4(a0, c1, c0); c1 = sad4(a1, c2, c1); c2 = sad4(a2, c3, c2); c3 = sad4(a3, c0, c3); }
            if (OP == 1) { c0 = __vavgu4(c0, a0); c1 = __vavgu4(c1, a1); c2 = __vavgu4(c2, a2); c3 = __vavgu4(c3, a3); }
            if (OP == 2) { c0 = __funnelshift_r(c0, a0, b); c1 = __funnelshift_r(c1, a1, b); c2 = __funnelshift_r(c2, a2, b); c3 = __funnelshift_r(c3, a3, b); }
            if (OP == 3) { c0 = __byte_perm(c0, a0, b); c1 = __byte_perm(c1, a1, b); c2 = __byte_perm(c2, a2, b); c3 = __byte_perm(c3, a3, b); }
            if (OP == 4) { c0 = c0 * a0 + b; c1 = c1 * a1 + b; c2 = c2 * a2 + b; c3 = c3 * a3 + b; }
            if (OP == 5) { c0 += sm[(c0 + a0) & 1023]; c1 += sm[(c1 + a1) & 1023]; c2 += sm[(c2 + a2) & 1023]; c3 += sm[(c3 + a3) & 1023]; }
            if (OP == 6) { c0 = (c0 + a0) ^ b; c1 = (c1 + a1) ^ b; c2 = (c2 + a2) ^ b; c3 = (c3 + a3) ^ b; }
            if (OP == 8) { c0 = __mulhi((int)c0, 1 << 20) + a0; c1 = __mulhi((int)c1, 1 << 20) + a1; c2 = __mulhi((int)c2, 1 << 20) + a2; c3 = __mulhi((int)c3, 1 << 20) + a3; }
            if (OP == 9) { c0 = sad4(a0, c1, c0); c1 = c1 * a1 + b; c2 = sad4(a2, c3, c2); c3 = c3 * a3 + b; }
            if (OP == 10) { c0 = ((int)c0 >> 12) + a0; c1 = ((int)c1 >> 12) + a1; c2 = ((int)c2 >> 12) + a2; c3 = ((int)c3 >> 12) + a3; }
            if (OP == 7) { c0 = __vsub2(c0, a0); c1 = __vsub2(c1, a1); c2 = __vsub2(c2, a2); c3 = __vsub2(c3, a3); }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3;
}

template <int OP> void run(const char *name, int ops_per_inner, uint32_t *d) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4, threads = 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(d, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<OP><<<blocks, threads>>>(d, 2); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ops = (double)blocks * threads * ITERS * 8.0 * 4.0 * ops_per_inner;
    printf("%-28s %8.3f ms  %8.2f Gop/s(thread-ops)  %6.2f thread-ops/clk/SM @%d MHz nominal\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000);
}

int main() {
    uint32_t *d; cudaMalloc(&d, 148 * 4 * 512 * 4 * 2);
    run<0>("vsadu4+acc (VABSDIFF4.ACC)", 1, d);
    run<1>("vavgu4 (emulated)", 1, d);
    run<2>("funnelshift_r (SHF)", 1, d);
    run<3>("byte_perm (PRMT)", 1, d);
    run<4>("imad", 1, d);
    run<5>("lds.32 + iadd", 1, d);
    run<6>("iadd+xor (LOP3/IADD3)", 1, d);
    run<7>("vsub2 (emulated)", 1, d);
    run<8>("mulhi(x,2^20)+a (IMAD.HI)", 1, d);
    run<9>("half VABSDIFF4 + half IMAD", 1, d);
    run<10>("(x>>12)+a (SHF+IADD)", 1, d);
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

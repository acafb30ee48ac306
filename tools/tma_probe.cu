// tools/tma_probe.cu - isolates which TMA box shapes / coordinates the B200 accepts for the u8
// tensors K1 uses.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, uint32_t bytes, uint8_t *out) {
    __shared__ __align__(128) uint8_t buf[4096];
    __shared__ unsigned long long bar;
    const uint32_t b = smem_u32(&bar), d = smem_u32(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(d), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(d), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(b) : "memory");
    }
    __syncwarp();
    asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(b), "r"(0) : "memory");
    for (uint32_t i = threadIdx.x; i < bytes; i += 32) out[i] = buf[i];
}

int main(int argc, char **argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1; int ci = -1;
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int W = 64, H = 64, F = 5;
    std::vector<uint8_t> h((size_t)W * H * 3 * F);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d, *out; cudaMalloc(&d, h.size()); cudaMalloc(&out, 4096); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    const cuuint32_t es[4] = {1, 1, 1, 1};
    struct Case { const char *name; int rank; cuuint64_t dims[4]; cuuint64_t strides[3]; cuuint32_t box[4]; int c[4]; };
    Case cases[] = {
        {"4d box16x16 in-bounds", 4, {64, 64, 3, 5}, {64, 4096, 12288}, {16, 16, 1, 1}, {16, 16, 1, 2}},
        {"3d box48x30 at (-16,-7)", 3, {64, 64, 15, 0}, {64, 4096, 0}, {48, 30, 1, 0}, {-16, -7, 0, 0}},
        {"3d box48x22 at (0,-3)", 3, {64, 64, 15, 0}, {64, 4096, 0}, {48, 22, 1, 0}, {0, -3, 2, 0}},
        {"3d box48x30 at (32,41)", 3, {64, 64, 15, 0}, {64, 4096, 0}, {48, 30, 1, 0}, {32, 41, 1, 0}},
        {"4d box32x16 at (-16,-4)", 4, {32, 32, 2, 1}, {32, 1024, 6144}, {32, 16, 1, 1}, {-16, -4, 1, 0}},
        {"4d box32x16 at (16,20)", 4, {32, 32, 2, 1}, {32, 1024, 6144}, {32, 16, 1, 1}, {16, 20, 0, 0}},
        {"3d box32x30 at (8,9) unaligned", 3, {64, 64, 15, 0}, {64, 4096, 0}, {32, 30, 1, 0}, {8, 9, 0, 0}},
    };
    for (auto &c : cases) {
        if (++ci, only >= 0 && ci != only) continue;
        CUtensorMap tm;
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, c.rank, d, c.dims, c.strides, c.box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        uint32_t bytes = c.box[0] * c.box[1];
        cudaMemset(out, 0xEE, 4096);
        if (r == CUDA_SUCCESS) {
            if (c.rank == 3) probe<3><<<1, 32>>>(tm, c.c[0], c.c[1], c.c[2], c.c[3], bytes, out);
            else probe<4><<<1, 32>>>(tm, c.c[0], c.c[1], c.c[2], c.c[3], bytes, out);
        }
        cudaError_t e = cudaDeviceSynchronize();
        uint8_t ho[64]; cudaMemcpy(ho, out, 64, cudaMemcpyDeviceToHost);
        printf("%-32s encode=%d run=%s first bytes %02x %02x %02x %02x .. %02x\n", c.name, (int)r, cudaGetErrorString(e), ho[0], ho[1], ho[8], ho[9], ho[40]);
        if (e != cudaSuccess) { printf("sticky error; stopping\n"); return 1; }
    }
    return 0;
}

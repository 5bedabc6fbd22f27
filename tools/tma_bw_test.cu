// Microbenchmark: L2 -> shared-memory bandwidth of TMA tile loads when every SM streams 128x64 bf16 tiles (16 KB, 128B
// swizzle) from an L2-resident matrix, `depth` loads in flight per SM.  Gives the ceiling for operand feeding.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I spec_b200/csrc -o tools/tma_bw_test.bin tools/tma_bw_test.cu
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace sb;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int DEPTH>
__global__ void __launch_bounds__(32) bw_kernel(const __grid_constant__ CUtensorMap map, int iters, int row_tiles, int k_tiles) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar = sbase + DEPTH * 16384;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DEPTH; ++s) mbar_init(bar + s * 8, 1);
        mbar_fence_init();
        uint32_t tile = blockIdx.x;
        for (int i = 0; i < iters + DEPTH; ++i) {
            const int s = i % DEPTH;
            if (i >= DEPTH) mbar_wait(bar + s * 8, ((i / DEPTH) - 1) & 1);
            if (i < iters) {
                mbar_arrive_expect_tx(bar + s * 8, 16384);
                const int rt = tile % row_tiles, kt = (tile / row_tiles) % k_tiles;
                tma_load_2d(sbase + s * 16384, &map, bar + s * 8, kt * 64, rt * 128);
                tile += gridDim.x;
            }
        }
    }
}

int main() {
    void* q = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr);
    auto enc = reinterpret_cast<EncodeTiledFn>(q);
    for (long long mb : {32LL, 512LL}) {                 // 32 MB: L2-resident; 512 MB: streams from HBM
        const int K = 1024;
        const long long rows = mb * 1024 * 1024 / (K * 2);
        void* d; cudaMalloc(&d, rows * K * 2); cudaMemset(d, 0, rows * K * 2);
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
        enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        auto run = [&](auto kern, int depth) {
            const int smem = depth * 16384 + 256 + 1024;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            const int iters = 4000;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                kern<<<148, 32, smem>>>(map, iters, (int)(rows / 128), K / 64);
                cudaEventRecord(e1);
                cudaDeviceSynchronize();
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("buffer %4lld MB  depth %2d x 16 KB per SM : %7.2f TB/s into smem (148 SMs)\n", mb, depth, 148.0 * iters * 16384 / (ms * 1e-3) / 1e12);
        };
        run(bw_kernel<2>, 2); run(bw_kernel<4>, 4); run(bw_kernel<8>, 8); run(bw_kernel<12>, 12);
        cudaFree(d);
    }
    return 0;
}

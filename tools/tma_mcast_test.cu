// Microbenchmark for round 2 (profiles/round2_plan.md item 1): what rate do MULTICAST TMA tile loads sustain?
// A cluster of CL CTAs streams 128x64 bf16 tiles (16 KB, 128B swizzle) out of an L2-resident matrix; for every tile each CTA
// fetches 1/CL of it (128/CL rows) and multicasts its slice to all CL CTAs, so every CTA ends up with the whole tile in its own
// shared memory while L2 is read only once per cluster.  Reported: bytes landed in shared memory per second (chip-wide) and the
// L2 read rate that implies (landed / CL).  Compare with tools/tma_bw_test.cu (no multicast: 12.1 TB/s landed = 12.1 TB/s of L2).
// Stage reuse is cluster-wide: a stage may be overwritten only after EVERY CTA of the cluster has seen it full, so every CTA
// arrives on every CTA's empty barrier (count CL) -- the same protocol a multicast conv kernel needs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I spec_b200/csrc -o tools/tma_mcast_test.bin tools/tma_mcast_test.cu
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace sb;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CL, int DEPTH>
__global__ void __launch_bounds__(32) mcast_kernel(const __grid_constant__ CUtensorMap map, int iters, int row_tiles, int k_tiles) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar_full = sbase + DEPTH * 16384;
    const uint32_t bar_empty = bar_full + DEPTH * 8;
    const uint32_t rank = cluster_ctarank();
    constexpr int SLICE_ROWS = 128 / CL, SLICE_BYTES = SLICE_ROWS * 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DEPTH; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, CL); }
        mbar_fence_init();
    }
    cluster_sync_all();                                           // every CTA's barriers exist before anyone multicasts into them
    if (threadIdx.x == 0) {
        uint32_t tile = cluster_id_x();
        const uint32_t step = cluster_nclusters_x();
        for (int i = 0; i < iters + DEPTH; ++i) {
            const int s = i % DEPTH;
            if (i >= DEPTH) {
                // consume: the tile of iteration i-DEPTH is complete in OUR smem; tell every CTA of the cluster
                mbar_wait(bar_full + s * 8, ((i / DEPTH) - 1) & 1);
                for (uint32_t r = 0; r < CL; ++r) mbar_arrive_cluster(mapa_u32(bar_empty + s * 8, r));
            }
            if (i < iters) {
                if (i >= DEPTH) mbar_wait(bar_empty + s * 8, ((i / DEPTH) - 1) & 1);   // all CL CTAs are done with stage s
                mbar_arrive_expect_tx(bar_full + s * 8, 16384);                      // CL slices of 16384/CL bytes will land here
                const int rt = tile % row_tiles, kt = (tile / row_tiles) % k_tiles;
                tma_load_2d_mcast(sbase + s * 16384 + rank * SLICE_BYTES, &map, bar_full + s * 8, kt * 64, rt * 128 + rank * SLICE_ROWS,
                                  static_cast<uint16_t>((1u << CL) - 1));
                tile += step;
            }
        }
    }
    __syncwarp();                                                 // reconverge before the .aligned cluster barrier
    cluster_sync_all();                                           // nobody exits while a peer may still signal its barriers
}

// Two-warp variant (the shape of the real kernels): warp 0 = producer (wait empty, arm, issue its slice), warp 1 = consumer
// (wait full, release the stage in every CTA of the cluster).  Separates the issue rate of the producer from the handshake.
template <int CL, int DEPTH>
__global__ void __launch_bounds__(64) mcast2_kernel(const __grid_constant__ CUtensorMap map, int iters, int row_tiles, int k_tiles) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar_full = sbase + DEPTH * 16384;
    const uint32_t bar_empty = bar_full + DEPTH * 8;
    const uint32_t rank = cluster_ctarank();
    constexpr int SLICE_ROWS = 128 / CL, SLICE_BYTES = SLICE_ROWS * 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DEPTH; ++s) { mbar_init(bar_full + s * 8, 1); mbar_init(bar_empty + s * 8, CL); }
        mbar_fence_init();
    }
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x == 0) {                                        // producer
        uint32_t tile = cluster_id_x();
        const uint32_t step = cluster_nclusters_x();
        for (int i = 0; i < iters; ++i) {
            const int s = i % DEPTH;
            mbar_wait(bar_empty + s * 8, ((i / DEPTH) & 1) ^ 1);
            mbar_arrive_expect_tx(bar_full + s * 8, 16384);
            const int rt = tile % row_tiles, kt = (tile / row_tiles) % k_tiles;
            tma_load_2d_mcast(sbase + s * 16384 + rank * SLICE_BYTES, &map, bar_full + s * 8, kt * 64, rt * 128 + rank * SLICE_ROWS,
                              static_cast<uint16_t>((1u << CL) - 1));
            tile += step;
        }
    } else if (threadIdx.x == 32) {                                // consumer
        for (int i = 0; i < iters; ++i) {
            const int s = i % DEPTH;
            mbar_wait(bar_full + s * 8, (i / DEPTH) & 1);
            for (uint32_t r = 0; r < CL; ++r) mbar_arrive_cluster(mapa_u32(bar_empty + s * 8, r));
        }
    }
    __syncwarp();
    __syncthreads();
    cluster_sync_all();
}

template <int CL, int DEPTH>
static void run(const CUtensorMap& map, long long rows, int K, long long mb) {
    auto kern = mcast_kernel<CL, DEPTH>;
    const int smem = DEPTH * 16384 + 256 + 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int iters = 4000;
    const int nsm = 148 / CL * CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nsm); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaError_t err = cudaSuccess;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        err = cudaLaunchKernelEx(&cfg, kern, map, iters, (int)(rows / 128), K / 64);
        cudaEventRecord(e1);
        if (err == cudaSuccess) err = cudaDeviceSynchronize();
        if (err != cudaSuccess) break;
    }
    if (err != cudaSuccess) { printf("cluster %d depth %2d: %s\n", CL, DEPTH, cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double landed = (double)nsm * iters * 16384 / (ms * 1e-3) / 1e12;
    printf("buffer %4lld MB  cluster %d  depth %2d x 16 KB per SM : %6.2f TB/s landed in smem, %6.2f TB/s read from L2 (%d SMs)\n",
           mb, CL, DEPTH, landed, landed / CL, nsm);
}

template <int CL, int DEPTH>
static void run2(const CUtensorMap& map, long long rows, int K, long long mb) {
    auto kern = mcast2_kernel<CL, DEPTH>;
    const int smem = DEPTH * 16384 + 256 + 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int iters = 4000;
    const int nsm = 148 / CL * CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nsm); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaError_t err = cudaSuccess;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        err = cudaLaunchKernelEx(&cfg, kern, map, iters, (int)(rows / 128), K / 64);
        cudaEventRecord(e1);
        if (err == cudaSuccess) err = cudaDeviceSynchronize();
        if (err != cudaSuccess) break;
    }
    if (err != cudaSuccess) { printf("cluster %d depth %2d: %s\n", CL, DEPTH, cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double landed = (double)nsm * iters * 16384 / (ms * 1e-3) / 1e12;
    printf("[2 warps] buffer %4lld MB  cluster %d  depth %2d x 16 KB per SM : %6.2f TB/s landed in smem, %6.2f TB/s read from L2 (%d SMs)\n",
           mb, CL, DEPTH, landed, landed / CL, nsm);
}

int main() {
    void* q = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr);
    auto enc = reinterpret_cast<EncodeTiledFn>(q);
    for (long long mb : {32LL, 512LL}) {                 // 32 MB: L2-resident; 512 MB: streams from HBM
        const int K = 1024;
        const long long rows = mb * 1024 * 1024 / (K * 2);
        void* d; cudaMalloc(&d, rows * K * 2); cudaMemset(d, 0, rows * K * 2);
        auto mk = [&](int box_rows) {
            CUtensorMap map;
            cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
            enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            return map;
        };
        CUtensorMap m1 = mk(128), m2 = mk(64), m4 = mk(32);
        run<1, 4>(m1, rows, K, mb); run<1, 8>(m1, rows, K, mb);
        run<2, 4>(m2, rows, K, mb); run<2, 8>(m2, rows, K, mb);
        run<4, 4>(m4, rows, K, mb); run<4, 8>(m4, rows, K, mb);
        run2<1, 8>(m1, rows, K, mb); run2<2, 8>(m2, rows, K, mb); run2<4, 8>(m4, rows, K, mb); run2<2, 12>(m2, rows, K, mb);
        cudaFree(d);
    }
    return 0;
}

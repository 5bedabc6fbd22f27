// Probe for "max-pool inside the stem epilogue": throughput of 16-byte bf16x8 MAX reductions (REDG.E.MAX.BF16x8) into a pooled
// NHWC tensor [N][56][56][64] (103 MB at N=256) when every pooled pixel receives ~5 partial maxima of 128 B each (eight 16-byte
// reductions), as the 8x16-pixel stem tiles would emit them; compared with a plain 16-byte-store kernel over the same bytes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/red_bw_test.bin tools/red_bw_test.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void red_kernel(uint32_t* pooled, long long npix, int partials, int use_red) {
    // one warp handles 18 pixels per "tile" (lanes 0..17 active), 8 chunk instructions each, like one epilogue warp
    const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long items = npix * partials / 18;
    for (long long it = warp_global; it < items; it += nwarps) {
        if (lane < 18) {
            // pixel: 9 consecutive pooled pixels of one row + 9 of the next row (56 apart), shifted per partial so that pixels overlap
            const long long base = (it * 9 / partials * 1) % (npix - 80);
            const long long pix = base + (lane < 9 ? lane : 56 + lane - 9);
            uint32_t* p = pooled + pix * 32;
            const uint32_t v = 0x3f803f80u + static_cast<uint32_t>(it & 7);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (use_red)
                    asm volatile("red.global.max.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p + c * 4), "r"(v), "r"(v), "r"(v), "r"(v) : "memory");
                else
                    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p + c * 4), "r"(v), "r"(v), "r"(v), "r"(v) : "memory");
            }
        }
    }
}

int main() {
    const long long npix = 256LL * 56 * 56;
    uint32_t* d;
    cudaMalloc(&d, npix * 128);
    cudaMemset(d, 0, npix * 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int use_red = 1; use_red >= 0; --use_red)
        for (int partials : {2, 4, 6}) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                red_kernel<<<148 * 4, 128>>>(d, npix, partials, use_red);
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double bytes = static_cast<double>(npix) * partials * 128;
            printf("%s partials/pixel=%d: %.3f ms, %.0f GB/s of operand bytes (%.1f M 16-byte ops)\n", use_red ? "REDG.MAX.BF16x8" : "STG.128        ",
                   partials, best, bytes / best / 1e6, bytes / 16 / 1e6);
        }
    float ms;
    cudaEventRecord(e0); cudaMemsetAsync(d, 0, npix * 128); cudaEventRecord(e1); cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemset of the pooled tensor (103 MB): %.3f ms\n", ms);
    return 0;
}

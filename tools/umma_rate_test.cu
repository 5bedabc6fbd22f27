// Microbenchmark: steady-state cycles per tcgen05.mma when operands are already in shared memory (no TMA traffic):
// isolates the tensor pipe + smem operand fetch rate for N = 64/128/256, cta_group 1 and 2, one CTA (pair) per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I spec_b200/csrc -o tools/umma_rate_test.bin tools/umma_rate_test.cu
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace sb;

template <int N, int CTAS>
__global__ void __launch_bounds__(128) rate_kernel(long long* cycles_out, int iters, int nbuf) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (sbase - smem_u32(raw));
    // nbuf (A,B) stage pairs so consecutive MMAs read different smem like a real pipeline
    const uint32_t a_base = sbase, b_base = sbase + nbuf * 16384;
    const uint32_t bar = b_base + nbuf * (N / CTAS) * 128;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(gen + (bar - sbase) + 8);
    const int t = threadIdx.x;
    for (uint32_t i = t; i < (bar - sbase) / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u;   // bf16 ~0.0078
    if (t == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (t < 32) {
        if (CTAS == 2) tmem_alloc_2cta(smem_u32(tptr), N < 32 ? 32 : N);
        else { tmem_alloc(smem_u32(tptr), N < 32 ? 32 : N); tmem_relinquish(); }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tacc = *tptr;
    const bool issuer = t == 0 && (CTAS == 1 || cluster_ctarank() == 0);
    long long t0 = 0, t1 = 0;
    if (issuer) {
        constexpr uint32_t idesc = umma_idesc_f16(1, 128 * CTAS, N);
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t s = i % nbuf;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t ad = umma_desc_sw128(a_base + s * 16384 + k * 32), bd = umma_desc_sw128(b_base + s * (N / CTAS) * 128 + k * 32);
                if (CTAS == 2) umma_f16_2cta(tacc, ad, bd, idesc, 1); else umma_f16(tacc, ad, bd, idesc, 1);
            }
        }
        if (CTAS == 2) umma_commit_2cta(bar); else umma_commit(bar);
    }
    mbar_wait(bar, 0);
    if (issuer) { t1 = clock64(); if (blockIdx.x == 0) cycles_out[0] = t1 - t0; }
    tc_fence_before();
    if (CTAS == 2) cluster_sync_all(); else __syncthreads();
    if (t < 32) { if (CTAS == 2) tmem_dealloc_2cta(tacc, N < 32 ? 32 : N); else tmem_dealloc(tacc, N < 32 ? 32 : N); }
}

template <int N, int CTAS>
void run(int nbuf) {
    long long* d; cudaMalloc(&d, 8);
    const int iters = 2000;
    const int smem = nbuf * 16384 + nbuf * (N / CTAS) * 128 + 64 + 1024;
    cudaFuncSetAttribute(rate_kernel<N, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, rate_kernel<N, CTAS>, d, iters, nbuf);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d CTAS=%d: %s\n", N, CTAS, cudaGetErrorString(e)); return; }
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double per_mma = (double)cyc / (iters * 4.0);
    const double macs = 128.0 * N * 16;      // per SM per MMA
    const double tflops = 148.0 * iters * 4.0 * macs * 2 / (ms * 1e-3) / 1e12;
    printf("N=%3d cta_group::%d nbuf=%d: %7.1f cycles/MMA  (%6.0f MAC/cycle/SM)   chip %7.1f TFLOP/s (event-timed, all 148 SMs)\n", N, CTAS, nbuf, per_mma,
           macs / per_mma, tflops);
}

int main() {
    run<64, 1>(4); run<128, 1>(4); run<256, 1>(3); run<128, 2>(4); run<256, 2>(4);
    run<256, 1>(1); run<256, 2>(1);
    return 0;
}

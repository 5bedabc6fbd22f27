// Microbenchmark (round 2): is the 87-cycle floor of N <= 128 tcgen05.mma (tools/umma_rate_test.cu) a DEPENDENT-ACCUMULATE latency?
// Same loop, but consecutive MMAs rotate over NACC independent TMEM accumulators (column offsets j*N).  If cycles/MMA drops
// towards max(M,128)*N/256 (32 cycles at N=64, 64 at N=128) with NACC >= 2, kernels whose tiles have N <= 128 should interleave
// the MMAs of two tiles / two K halves instead of issuing one accumulation chain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I spec_b200/csrc -o tools/umma_rate2_test.bin tools/umma_rate2_test.cu
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace sb;

template <int N>
__global__ void __launch_bounds__(128) rate_kernel(long long* cycles_out, int iters, int nbuf, int nacc, int same_a) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (sbase - smem_u32(raw));
    const uint32_t a_base = sbase, b_base = sbase + nbuf * 16384;
    const uint32_t bar = b_base + nbuf * N * 128;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(gen + (bar - sbase) + 8);
    const int t = threadIdx.x;
    for (uint32_t i = t; i < (bar - sbase) / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u;
    if (t == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (t < 32) { tmem_alloc(smem_u32(tptr), 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = *tptr;
    long long t0 = 0, t1 = 0;
    if (t == 0) {
        constexpr uint32_t idesc = umma_idesc_f16(1, 128, N);
        t0 = clock64();
        uint32_t j = 0;
        for (int i = 0; i < iters; ++i) {
            const uint32_t s = i % nbuf;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t ad = umma_desc_sw128(a_base + (same_a ? 0 : s * 16384) + k * 32), bd = umma_desc_sw128(b_base + s * N * 128 + k * 32);
                umma_f16(tacc + j * N, ad, bd, idesc, 1);
                j = (j + 1 == static_cast<uint32_t>(nacc)) ? 0 : j + 1;
            }
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    if (t == 0) { t1 = clock64(); if (blockIdx.x == 0) cycles_out[0] = t1 - t0; }
    tc_fence_before();
    __syncthreads();
    if (t < 32) tmem_dealloc(tacc, 512);
}

template <int N>
void run(int nbuf, int nacc, int same_a) {
    long long* d; cudaMalloc(&d, 8);
    const int iters = 2000;
    const int smem = nbuf * 16384 + nbuf * N * 128 + 64 + 1024;
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        rate_kernel<N><<<148, 128, smem>>>(d, iters, nbuf, nacc, same_a);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return; }
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double per_mma = (double)cyc / (iters * 4.0);
    const double macs = 128.0 * N * 16;
    const double tflops = 148.0 * iters * 4.0 * macs * 2 / (ms * 1e-3) / 1e12;
    printf("N=%3d nacc=%d nbuf=%d same_a=%d: %7.1f cycles/MMA (%6.0f MAC/cycle/SM)  chip %7.1f TFLOP/s\n", N, nacc, nbuf, same_a, per_mma, macs / per_mma, tflops);
}

int main() {
    for (int nacc : {1, 2, 4, 8}) run<64>(4, nacc, 0);
    for (int nacc : {1, 2, 4}) run<128>(4, nacc, 0);
    for (int nacc : {1, 2}) run<256>(3, nacc, 0);
    run<64>(4, 2, 1); run<64>(1, 2, 0); run<128>(1, 2, 0);
    return 0;
}

// Experiment: minimal cta_group::2 UMMA (M=256 across a CTA pair, N=128, K=64) without TMA, to pin down the
// conventions before building the 2-CTA conv kernel: TMEM alloc in both CTAs, one MMA issued by the leader,
// multicast commit to both CTAs' mbarriers, each CTA holds A rows [128r,128r+128) and B rows [64r,64r+64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I spec_b200/csrc -o tools/umma_2cta_test.bin tools/umma_2cta_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "common.cuh"
using namespace sb;

constexpr int NN = 128;     // full N
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
k2cta(const __nv_bfloat16* A, const __nv_bfloat16* B, float* C) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (sbase - smem_u32(raw));
    const uint32_t a_base = sbase, b_base = sbase + 16384, bar = sbase + 16384 + 8192;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(gen + 16384 + 8192 + 8);
    const int t = threadIdx.x;
    const uint32_t rank = cluster_rank();
    for (int i = t; i < 128 * 8; i += 128) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(gen + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + (rank * 128 + r) * 64 + j * 8);
    }
    for (int i = t; i < 64 * 8; i += 128) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(gen + 16384 + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + (rank * 64 + r) * 64 + j * 8);
    }
    if (t == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tacc = *tptr;
    if (rank == 0 && t == 0) {
        constexpr uint32_t idesc = umma_idesc_f16(1, 256, NN);
        for (int k = 0; k < 4; ++k) {
            const uint64_t ad = umma_desc_sw128(a_base + k * 32), bd = umma_desc_sw128(b_base + k * 32);
            const uint32_t acc = k != 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tacc), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const int warp = t >> 5;
    for (int c = 0; c < NN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tacc + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        for (int e = 0; e < 32; ++e) C[(rank * 128 + t) * NN + c * 32 + e] = __uint_as_float(v[e]);
    }
    tc_fence_before();
    cluster_sync_all();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tacc), "r"(128) : "memory");
}

int main() {
    std::vector<__nv_bfloat16> hA(256 * 64), hB(NN * 64);
    std::vector<float> fA(256 * 64), fB(NN * 64);
    srand(2);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dC;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, 256 * NN * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0, 256 * NN * 4);
    const int smem = 16384 + 8192 + 64 + 1024;
    cudaFuncSetAttribute(k2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k2cta<<<2, 128, smem>>>(dA, dB, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hC(256 * NN);
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    // hypotheses for the column order of the result
    double err_plain = 0, err_swap = 0;
    for (int m = 0; m < 256; ++m)
        for (int n = 0; n < NN; ++n) {
            double ref = 0, ref_sw = 0;
            const int nsw = (n + NN / 2) % NN;
            for (int k = 0; k < 64; ++k) { ref += (double)fA[m * 64 + k] * fB[n * 64 + k]; ref_sw += (double)fA[m * 64 + k] * fB[nsw * 64 + k]; }
            err_plain = fmax(err_plain, fabs(ref - hC[m * NN + n]));
            err_swap = fmax(err_swap, fabs(ref_sw - hC[m * NN + n]));
        }
    printf("2-CTA UMMA M=256 N=%d: max|err| plain column order = %.3e, halves swapped = %.3e  -> %s\n", NN, err_plain, err_swap,
           err_plain < 1e-3 ? "PLAIN OK" : (err_swap < 1e-3 ? "SWAPPED OK" : "WRONG"));
    for (int m : {0, 127, 128, 255}) printf("  C[%d][0..3] = %f %f %f %f\n", m, hC[m * NN], hC[m * NN + 1], hC[m * NN + 2], hC[m * NN + 3]);
    return 0;
}

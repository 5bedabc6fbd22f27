// Experiment: can a UMMA smem descriptor address a K-major SWIZZLE_128B operand whose first row is NOT at a
// 1024-byte (8-row) boundary?  If yes, a 3x3 conv can keep ONE haloed input tile in smem and feed the nine taps
// as row-shifted views of it (no per-tap re-load).  Tests start offsets r0 = 0..17 rows with
//   variant 0: base_offset field = 0            variant 1: base_offset = (start_addr >> 7) & 7
// against a CPU reference.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I spec_b200/csrc -o /tmp/umma_shift tools/umma_shift_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "common.cuh"
using namespace sb;

constexpr int ROWS = 160;          // rows of the big A buffer (20 KB)
__global__ void __launch_bounds__(128) shift_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* C, int r0, int variant) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (sbase - smem_u32(raw));
    const uint32_t a_base = sbase, b_base = sbase + ROWS * 128, bar = b_base + 64 * 128;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(gen + ROWS * 128 + 64 * 128 + 8);
    const int t = threadIdx.x;
    // swizzled fill: row r, 16-byte chunk j -> r*128 + ((j ^ (r & 7)) << 4)   (what TMA SWIZZLE_128B writes)
    for (int i = t; i < ROWS * 8; i += 128) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(gen + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + j * 8);
    }
    for (int i = t; i < 64 * 8; i += 128) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(gen + ROWS * 128 + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + j * 8);
    }
    if (t == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (t < 32) { tmem_alloc(smem_u32(tptr), 64); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tacc = *tptr;
    if (t == 0) {
        constexpr uint32_t idesc = umma_idesc_f16(1, 128, 64);
        for (int k = 0; k < 4; ++k) {
            const uint32_t a_addr = a_base + r0 * 128 + k * 32;
            uint64_t ad = umma_desc_sw128(a_addr);
            if (variant == 1) ad |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
            umma_f16(tacc, ad, umma_desc_sw128(b_base + k * 32), idesc, k != 0);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const int warp = t >> 5;
    for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tacc + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
        tmem_ld_wait();
        for (int e = 0; e < 32; ++e) C[t * 64 + c * 32 + e] = __uint_as_float(v[e]);
    }
    tc_fence_before();
    __syncthreads();
    if (t < 32) tmem_dealloc(tacc, 64);
}

int main() {
    std::vector<__nv_bfloat16> hA(ROWS * 64), hB(64 * 64);
    std::vector<float> fA(ROWS * 64), fB(64 * 64);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dC;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, 128 * 64 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    const int smem = ROWS * 128 + 64 * 128 + 64 + 1024;
    cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> hC(128 * 64);
    for (int variant = 0; variant < 2; ++variant)
        for (int r0 : {0, 1, 2, 3, 7, 8, 9, 16, 17, 18, 31}) {
            cudaMemset(dC, 0, 128 * 64 * 4);
            shift_kernel<<<1, 128, smem>>>(dA, dB, dC, r0, variant);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant %d r0 %d: CUDA error %s\n", variant, r0, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 64; ++k) ref += (double)fA[(r0 + m) * 64 + k] * fB[n * 64 + k];
                    maxerr = fmax(maxerr, fabs(ref - hC[m * 64 + n]));
                }
            printf("variant %d (base_offset %s)  r0 = %2d : max |err| = %.3e  %s\n", variant, variant ? "set" : "0", r0, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
        }
    return 0;
}

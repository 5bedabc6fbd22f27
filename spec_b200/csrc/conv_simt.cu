// fp32 SIMT kernels: the "parity mode" convolution (same NHWC implicit-GEMM view as conv_tc.cu but
// plain FFMA with fp32 storage, used to meet the north-star fp32 tolerances against the CPU oracle)
// and the fp32 linear layer used by the CamCalib and HMR-head tails
// (reference ops: cuBLAS addmm at /root/reference/camcalib/model.py:77-79 and the HMRHead GEMMs,
// SURVEY.md section 2.2).
#include "common.cuh"
#include "internal.h"

namespace sb {

// ------------------------------------------------------------------ conv, fp32
// Tile 64 pixels x 64 channels x 16 k; 256 threads, 4x4 outputs each.
constexpr int SC_BM = 64, SC_BN = 64, SC_BK = 16;

__global__ void __launch_bounds__(256)
conv_f32_kernel(const ConvParams p, const float* __restrict__ wt /* [K][Cout] */)
{
    __shared__ __align__(16) float As[SC_BK][SC_BM + 4];
    __shared__ __align__(16) float Bs[SC_BK][SC_BN];
    const float* __restrict__ in = static_cast<const float*>(p.in);
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * SC_BM;
    const int n0 = blockIdx.y * SC_BN;

    // A-load role: row a_r, k-quad a_k
    const int a_r = tid >> 2, a_k = (tid & 3) * 4;
    const long long r = static_cast<long long>(m0) + a_r;
    const bool row_ok = r < p.M;
    int n = 0, oh = 0, ow = 0;
    if (row_ok) {
        const int hw = p.Ho * p.Wo;
        n = static_cast<int>(r / hw);
        const int rem = static_cast<int>(r - static_cast<long long>(n) * hw);
        oh = rem / p.Wo;
        ow = rem - oh * p.Wo;
    }
    const int ih0 = oh * p.stride - p.pad, iw0 = ow * p.stride - p.pad;
    const float* base = in + static_cast<size_t>(n) * p.H * p.W * p.Cin;
    // B-load role
    const int b_k = tid >> 4, b_c = (tid & 15) * 4;
    // compute role
    const int ty = tid >> 4, tx = tid & 15;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += SC_BK) {
        {   // A tile (Cin % 4 == 0, so a k-quad never straddles a tap)
            const int kidx = k0 + a_k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_ok && kidx < p.K) {
                const int tap = kidx / p.Cin;
                const int c = kidx - tap * p.Cin;
                const int khi = tap / p.kw, kwi = tap - khi * p.kw;
                const int ih = ih0 + khi, iw = iw0 + kwi;
                if (static_cast<unsigned>(ih) < static_cast<unsigned>(p.H) && static_cast<unsigned>(iw) < static_cast<unsigned>(p.W))
                    v = *reinterpret_cast<const float4*>(base + (static_cast<size_t>(ih) * p.W + iw) * p.Cin + c);
            }
            As[a_k + 0][a_r] = v.x; As[a_k + 1][a_r] = v.y; As[a_k + 2][a_r] = v.z; As[a_k + 3][a_r] = v.w;
        }
        {   // B tile
            const int kidx = k0 + b_k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kidx < p.K && n0 + b_c < p.Cout)
                v = *reinterpret_cast<const float4*>(wt + static_cast<size_t>(kidx) * p.Cout + n0 + b_c);
            *reinterpret_cast<float4*>(&Bs[b_k][b_c]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SC_BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    float* __restrict__ out = static_cast<float*>(p.out);
    const float* __restrict__ res = static_cast<const float*>(p.res);
    const int col = n0 + tx * 4;
    if (col < p.Cout) {
        const float4 bq = *reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long row = static_cast<long long>(m0) + ty * 4 + i;
            if (row < p.M) {
                float4 o = make_float4(acc[i][0] + bq.x, acc[i][1] + bq.y, acc[i][2] + bq.z, acc[i][3] + bq.w);
                if (res != nullptr) {
                    const float4 rv = *reinterpret_cast<const float4*>(res + static_cast<size_t>(row) * p.res_ld + col);
                    o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
                }
                if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                *reinterpret_cast<float4*>(out + static_cast<size_t>(row) * p.out_ld + p.out_coff + col) = o;
            }
        }
    }
}

bool conv_f32_launch(const ConvParams& p, const ConvWeights& w, cudaStream_t s) {
    if (w.w_f32 == nullptr) { set_error("conv_f32: weights not packed"); return false; }
    if ((p.Cin % 4) != 0 || (p.Cout % 4) != 0 || (p.out_ld % 4) != 0 || (p.out_coff % 4) != 0) {
        set_error("conv_f32: channel counts must be multiples of 4");
        return false;
    }
    dim3 grid((p.M + SC_BM - 1) / SC_BM, (p.Cout + SC_BN - 1) / SC_BN);
    conv_f32_kernel<<<grid, 256, 0, s>>>(p, w.w_f32);
    return check_cuda(cudaGetLastError(), "conv_f32 launch");
}

// ------------------------------------------------------------------ linear, fp32
// out[m, n] = sum_k A[m,k] * W[n,k] + bias[n] + add[m,n];  tile 32 x 64 x 32, 256 threads, 2x4 each.
constexpr int LN_BM = 32, LN_BN = 64, LN_BK = 32;

__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                  const float* __restrict__ bias, const float* add, int add_ld, float* out, int out_ld,
                  int M, int N, int K, int k_chunk, size_t split_stride)
{
    // split-K: slice blockIdx.z handles k in [z*k_chunk, min(K, (z+1)*k_chunk)) and writes its partial sums to
    // out + z*split_stride (bias / addend only in slice 0); the consumer adds the slices in a fixed order.
    const int kz0 = blockIdx.z * k_chunk;
    A += kz0; W += kz0;
    K = min(K - kz0, k_chunk);
    out += blockIdx.z * split_stride;
    if (blockIdx.z != 0) { bias = nullptr; add = nullptr; }
    __shared__ __align__(16) float As[LN_BK][LN_BM + 4];
    __shared__ __align__(16) float Ws[LN_BK][LN_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * LN_BM, n0 = blockIdx.x * LN_BN;
    const int a_r = tid >> 3, a_k = (tid & 7) * 4;          // 32 rows x 8 k-quads
    const int ty = tid >> 4, tx = tid & 15;                 // rows ty*2.., cols tx*4..
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};

    // register double-buffering: the global loads of k-block i+1 are in flight while k-block i is multiplied (the grid of the
    // CamCalib GEMM is only 96 CTAs, so nothing else hides the load latency)
    auto load_tiles = [&](int k0, float4& va, float4 (&vw)[2]) {
        va = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + a_r < M && k0 + a_k < K)
            va = *reinterpret_cast<const float4*>(A + static_cast<size_t>(m0 + a_r) * lda + k0 + a_k);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int w_r = (tid >> 3) + h * 32;
            vw[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + w_r < N && k0 + a_k < K)
                vw[h] = *reinterpret_cast<const float4*>(W + static_cast<size_t>(n0 + w_r) * ldw + k0 + a_k);
        }
    };
    float4 va, vw[2];
    load_tiles(0, va, vw);
    for (int k0 = 0; k0 < K; k0 += LN_BK) {
        As[a_k + 0][a_r] = va.x; As[a_k + 1][a_r] = va.y; As[a_k + 2][a_r] = va.z; As[a_k + 3][a_r] = va.w;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int w_r = (tid >> 3) + h * 32;
            Ws[a_k + 0][w_r] = vw[h].x; Ws[a_k + 1][w_r] = vw[h].y; Ws[a_k + 2][w_r] = vw[h].z; Ws[a_k + 3][w_r] = vw[h].w;
        }
        __syncthreads();
        if (k0 + LN_BK < K) load_tiles(k0 + LN_BK, va, vw);
#pragma unroll
        for (int k = 0; k < LN_BK; ++k) {
            const float2 a = *reinterpret_cast<const float2*>(&As[k][ty * 2]);
            const float4 b = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
            acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
            acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
            acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty * 2 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nn = n0 + tx * 4 + j;
            if (nn >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[nn];
            if (add) v += add[static_cast<size_t>(m) * add_ld + nn];
            out[static_cast<size_t>(m) * out_ld + nn] = v;
        }
    }
}

bool linear_f32_launch(const float* A, int lda, const float* W, int ldw, const float* bias, const float* add,
                       int add_ld, float* out, int out_ld, int M, int N, int K, cudaStream_t s, int ksplit,
                       size_t split_stride, int* ksplit_used) {
    if ((K % 4) != 0 || (lda % 4) != 0 || (ldw % 4) != 0 ||
        (reinterpret_cast<uintptr_t>(A) & 15) != 0 || (reinterpret_cast<uintptr_t>(W) & 15) != 0) {
        set_error("linear_f32: K, lda, ldw must be multiples of 4 and pointers 16-byte aligned");
        return false;
    }
    if (ksplit < 1) ksplit = 1;
    int k_chunk = (K + ksplit - 1) / ksplit;
    k_chunk = (k_chunk + 3) / 4 * 4;
    ksplit = (K + k_chunk - 1) / k_chunk;                    // rounding k_chunk up to a multiple of 4 can leave fewer slices than asked for
    if (ksplit_used) *ksplit_used = ksplit;
    dim3 grid((N + LN_BN - 1) / LN_BN, (M + LN_BM - 1) / LN_BM, ksplit);
    linear_f32_kernel<<<grid, 256, 0, s>>>(A, lda, W, ldw, bias, add, add_ld, out, out_ld, M, N, K, k_chunk, split_stride);
    return check_cuda(cudaGetLastError(), "linear_f32 launch");
}

}  // namespace sb

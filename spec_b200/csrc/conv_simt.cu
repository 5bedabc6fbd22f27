// fp32 SIMT kernels: the "parity mode" convolution (same NHWC implicit-GEMM view as conv_tc.cu but
// plain FFMA with fp32 storage, used to meet the north-star fp32 tolerances against the CPU oracle)
// and the fp32 linear layer used by the CamCalib and HMR-head tails
// (reference ops: cuBLAS addmm at /root/reference/camcalib/model.py:77-79 and the HMRHead GEMMs,
// SURVEY.md section 2.2).
#include <algorithm>

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
// out[m, n] = sum_k A[m,k] * W[n,k] + bias[n] + add[m,n].  Tile BM x 64 x 16 (BM = 64 or 32), 256 threads, each thread the
// rows ty + 16 i and the columns tx + 16 j (i < BM/16, j < 4).  Both operands stay k-contiguous in shared memory exactly as
// they are in global memory (row pitch 20 words: conflict-free for 16-byte reads at a lane stride of one row), filled by a
// 3-stage cp.async pipeline, and are read four k at a time: 8 LDS.128 per 64 FFMA at BM = 64.  (The round-1 kernel transposed
// through registers with 4-way-conflicting scalar stores and issued 2 LDS per 8 FFMA; the CamCalib GEMM, 256 x 768 x 2048 on
// 96 CTAs, took 95 us = 8.5 TFLOP/s.)
//
// Split-K, two flavours:
//   external  (split_stride != 0): slice z writes its partial sums to out + z*split_stride, bias/addend in slice 0 only; the
//             consumer adds the slices in a fixed order (the HMR head's iteration kernel does);
//   internal  (red != nullptr): slices write partials to red->partial; the LAST slice of a tile to arrive (device-scope counter
//             after a fence) sums all slices in the fixed order z = 0..ks-1, adds bias/addend and writes out -- deterministic
//             whichever slice is last -- and re-arms the counter for the next launch.
constexpr int LN_BN = 64, LN_BK = 16, LN_LD = LN_BK + 4, LN_STAGES = 3;

template <int BM>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                  const float* __restrict__ bias, const float* add, int add_ld, float* out, int out_ld,
                  int M, int N, int K, int k_chunk, size_t split_stride, float* red_partial, unsigned* red_counters)
{
    constexpr int RM = BM / 16;
    __shared__ __align__(16) float As[LN_STAGES][BM][LN_LD];
    __shared__ __align__(16) float Ws[LN_STAGES][LN_BN][LN_LD];
    __shared__ bool s_last;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * LN_BN;
    const int kz0 = blockIdx.z * k_chunk;
    A += kz0; W += kz0;
    const int Kz = min(K - kz0, k_chunk);
    const int nkb = (Kz + LN_BK - 1) / LN_BK;

    const int l_r = tid >> 2, l_k = (tid & 3) * 4;          // loader: row, k offset of this thread's 16-byte chunk
    auto issue = [&](int stage, int k0) {
        if (BM == 64 || tid < BM * 4) {
            const bool ok = (m0 + l_r < M) && (k0 + l_k < Kz);
            const float* src = ok ? A + static_cast<size_t>(m0 + l_r) * lda + k0 + l_k : A;
            cp_async16(smem_u32(&As[stage][l_r][l_k]), src, ok);
        }
        {
            const bool ok = (n0 + l_r < N) && (k0 + l_k < Kz);
            const float* src = ok ? W + static_cast<size_t>(n0 + l_r) * ldw + k0 + l_k : W;
            cp_async16(smem_u32(&Ws[stage][l_r][l_k]), src, ok);
        }
    };

    float acc[RM][4];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int st = 0; st < LN_STAGES - 1; ++st) {
        if (st < nkb) issue(st, st * LN_BK);
        cp_async_commit();
    }
    for (int kb = 0; kb < nkb; ++kb) {
        cp_async_wait<LN_STAGES - 2>();                        // k-block kb has landed (this thread's part) ...
        __syncthreads();                                       // ... everyone's part; and everyone is done with k-block kb-1
        if (kb + LN_STAGES - 1 < nkb) issue((kb + LN_STAGES - 1) % LN_STAGES, (kb + LN_STAGES - 1) * LN_BK);
        cp_async_commit();
        const int st = kb % LN_STAGES;
#pragma unroll
        for (int k4 = 0; k4 < LN_BK; k4 += 4) {
            float4 a[RM], b[4];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = *reinterpret_cast<const float4*>(&As[st][ty + 16 * i][k4]);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(&Ws[st][tx + 16 * j][k4]);
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]); acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]); acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                }
        }
    }
    cp_async_wait<0>();

    const int ks = gridDim.z, z = blockIdx.z;
    if (red_partial != nullptr && ks > 1) {
        float* mine = red_partial + static_cast<size_t>(z) * M * N;
#pragma unroll
        for (int i = 0; i < RM; ++i) {
            const int m = m0 + ty + 16 * i;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int nn = n0 + tx + 16 * j;
                if (nn < N) __stcg(mine + static_cast<size_t>(m) * N + nn, acc[i][j]);
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            unsigned* c = red_counters + blockIdx.y * gridDim.x + blockIdx.x;
            const unsigned arrived = atomicAdd(c, 1u);
            s_last = (arrived == static_cast<unsigned>(ks - 1));
            if (s_last) *c = 0;                                // all ks slices of this tile have arrived: re-arm for the next launch
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = m0 + ty + 16 * i, nn = n0 + tx + 16 * j;
                float v = 0.f;
                if (m < M && nn < N)
                    for (int q = 0; q < ks; ++q) v += __ldcg(red_partial + (static_cast<size_t>(q) * M + m) * N + nn);
                acc[i][j] = v;
            }
    } else {
        out += z * split_stride;
        if (z != 0) { bias = nullptr; add = nullptr; }
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int m = m0 + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nn = n0 + tx + 16 * j;
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
                       size_t split_stride, int* ksplit_used, const LinearRedWs* red) {
    if ((K % 4) != 0 || (lda % 4) != 0 || (ldw % 4) != 0 ||
        (reinterpret_cast<uintptr_t>(A) & 15) != 0 || (reinterpret_cast<uintptr_t>(W) & 15) != 0) {
        set_error("linear_f32: K, lda, ldw must be multiples of 4 and pointers 16-byte aligned");
        return false;
    }
    if (ksplit < 1) ksplit = 1;
    const int nt = (N + LN_BN - 1) / LN_BN;
    const bool internal = red != nullptr && red->partial != nullptr && red->counters != nullptr && split_stride == 0;
    if (internal) {
        // The slice count depends on K ONLY -- never on M: an image's logits must not change with the size or composition of the
        // batch it is in (tests/test_gpu_parity.py::test_batch_composition_invariance_at_full_size holds that bit for bit), and
        // the order of the k-sum is what the split fixes.  K/256 slices, at most 8: 2048 -> 8 slices of 16 k-blocks.
        ksplit = std::max(1, std::min(8, K / 256));
        if (ksplit > 1 && (static_cast<size_t>(ksplit) * M * N > red->partial_floats || nt * ((M + 31) / 32) > red->n_counters)) {
            set_error("linear_f32: split-K scratch too small for this batch (falling back would change the summation order)");
            return false;
        }
    }
    int k_chunk = (K + ksplit - 1) / ksplit;
    k_chunk = (k_chunk + LN_BK - 1) / LN_BK * LN_BK;         // whole k-blocks per slice
    ksplit = (K + k_chunk - 1) / k_chunk;                    // rounding k_chunk up can leave fewer slices than asked for
    if (ksplit_used) *ksplit_used = ksplit;
    float* rp = (internal && ksplit > 1) ? red->partial : nullptr;
    unsigned* rc = (internal && ksplit > 1) ? red->counters : nullptr;
    const bool big = M >= 64 && nt * ((M + 63) / 64) * ksplit >= 148;
    if (big) {
        dim3 grid(nt, (M + 63) / 64, ksplit);
        linear_f32_kernel<64><<<grid, 256, 0, s>>>(A, lda, W, ldw, bias, add, add_ld, out, out_ld, M, N, K, k_chunk, split_stride, rp, rc);
    } else {
        dim3 grid(nt, (M + 31) / 32, ksplit);
        linear_f32_kernel<32><<<grid, 256, 0, s>>>(A, lda, W, ldw, bias, add, add_ld, out, out_ld, M, N, K, k_chunk, split_stride, rp, rc);
    }
    return check_cuda(cudaGetLastError(), "linear_f32 launch");
}

}  // namespace sb

// Eval-side metrics on the device: H36M joint regression from the predicted mesh,
// pelvis centring, MPJPE, Procrustes-aligned MPJPE and per-vertex error -- what
// /root/reference/spec/trainer.py:272-316 and /root/reference/spec/utils/compute_error.py:33-86 do after copying the
// 21 MB of vertices per batch to the host (numpy SVD per sample).  Only three floats per image come back.
#include "common.cuh"
#include "internal.h"
#include "tail.h"

namespace sb {

// ------------------------------------------------------------------ 17 H36M joints = J_regressor_h36m (17 x 6890) . verts
// grid (B, nsets), block 256.  set 0 = predicted mesh, set 1 = ground-truth mesh (optional).  JT is [6890][20].
__global__ void __launch_bounds__(256)
h36m_joints_kernel(const float* __restrict__ v0, long long ld0, const float* __restrict__ v1, long long ld1,
                   const float* __restrict__ JT, float* __restrict__ out /*[nsets][B][17][3]*/, int B)
{
    __shared__ float red[8][51];
    const int b = blockIdx.x, set = blockIdx.y;
    const float* verts = (set == 0 ? v0 + b * ld0 : v1 + b * ld1);
    float acc[51];
#pragma unroll
    for (int i = 0; i < 51; ++i) acc[i] = 0.f;
    for (int v = threadIdx.x; v < SMPL_NV; v += 256) {
        const float x = verts[v * 3 + 0], y = verts[v * 3 + 1], z = verts[v * 3 + 2];
        const float4* jr = reinterpret_cast<const float4*>(JT + static_cast<size_t>(v) * 20);
        float w[20];
#pragma unroll
        for (int q = 0; q < 5; ++q) { const float4 t = jr[q]; w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w; }
#pragma unroll
        for (int j = 0; j < 17; ++j) {
            acc[j * 3 + 0] = fmaf(w[j], x, acc[j * 3 + 0]);
            acc[j * 3 + 1] = fmaf(w[j], y, acc[j * 3 + 1]);
            acc[j * 3 + 2] = fmaf(w[j], z, acc[j * 3 + 2]);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 51; ++i) {
        const float s = warp_sum(acc[i]);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 51) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        out[(static_cast<size_t>(set) * B + b) * 51 + threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------ 3x3 helpers (one thread per image)
__device__ inline void jacobi_eig3(float A[3][3], float V[3][3], float lam[3]) {
    // cyclic Jacobi on a symmetric 3x3; V accumulates the rotations (det +1)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.f : 0.f;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const float off = fabsf(A[0][1]) + fabsf(A[0][2]) + fabsf(A[1][2]);
        if (off < 1e-20f) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
            const float apq = A[p][q];
            if (fabsf(apq) < 1e-30f) continue;
            const float theta = (A[q][q] - A[p][p]) / (2.f * apq);
            const float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
            const float c = rsqrtf(t * t + 1.f), s = t * c;
#pragma unroll
            for (int k = 0; k < 3; ++k) {           // A <- A J
                const float akp = A[k][p], akq = A[k][q];
                A[k][p] = c * akp - s * akq;
                A[k][q] = s * akp + c * akq;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {           // A <- J^T A
                const float apk = A[p][k], aqk = A[q][k];
                A[p][k] = c * apk - s * aqk;
                A[q][k] = s * apk + c * aqk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float vkp = V[k][p], vkq = V[k][q];
                V[k][p] = c * vkp - s * vkq;
                V[k][q] = s * vkp + c * vkq;
            }
        }
    }
    lam[0] = A[0][0]; lam[1] = A[1][1]; lam[2] = A[2][2];
}

// grid ceil(B/64), block 64: thread = image.
// pred17 / gt17: [B][17][3] regressed joints; gt14: [B][14][3] given keypoints (used when non-null).
__global__ void __launch_bounds__(64)
eval_metrics_kernel(const float* __restrict__ pred17, const float* __restrict__ gt17, const float* __restrict__ gt14,
                    const int* __restrict__ map14, float* __restrict__ mpjpe, float* __restrict__ pampjpe,
                    float* __restrict__ pred14_out, float* __restrict__ pelvis_out /*[2][B][3]*/, int B)
{
    const int b = blockIdx.x * 64 + threadIdx.x;
    if (b >= B) return;
    float P[14][3], G[14][3];
    const float* pj = pred17 + static_cast<size_t>(b) * 51;
    const float ppx = pj[0], ppy = pj[1], ppz = pj[2];                  // pred pelvis = joint 0 (trainer.py:277)
    float gpx = 0.f, gpy = 0.f, gpz = 0.f;
    if (gt14 == nullptr) { const float* gj = gt17 + static_cast<size_t>(b) * 51; gpx = gj[0]; gpy = gj[1]; gpz = gj[2]; }
    for (int j = 0; j < 14; ++j) {
        const int m = map14[j];
        P[j][0] = pj[m * 3 + 0] - ppx; P[j][1] = pj[m * 3 + 1] - ppy; P[j][2] = pj[m * 3 + 2] - ppz;
        if (gt14 != nullptr) {
            G[j][0] = gt14[(static_cast<size_t>(b) * 14 + j) * 3 + 0]; G[j][1] = gt14[(static_cast<size_t>(b) * 14 + j) * 3 + 1];
            G[j][2] = gt14[(static_cast<size_t>(b) * 14 + j) * 3 + 2];
        } else {
            const float* gj = gt17 + static_cast<size_t>(b) * 51;
            G[j][0] = gj[m * 3 + 0] - gpx; G[j][1] = gj[m * 3 + 1] - gpy; G[j][2] = gj[m * 3 + 2] - gpz;
        }
        if (pred14_out) {
            float* o = pred14_out + (static_cast<size_t>(b) * 14 + j) * 3;
            o[0] = P[j][0]; o[1] = P[j][1]; o[2] = P[j][2];
        }
    }
    if (pelvis_out) {
        pelvis_out[b * 3 + 0] = ppx; pelvis_out[b * 3 + 1] = ppy; pelvis_out[b * 3 + 2] = ppz;
        pelvis_out[(static_cast<size_t>(B) + b) * 3 + 0] = gpx; pelvis_out[(static_cast<size_t>(B) + b) * 3 + 1] = gpy;
        pelvis_out[(static_cast<size_t>(B) + b) * 3 + 2] = gpz;
    }
    // MPJPE
    float e = 0.f;
    for (int j = 0; j < 14; ++j) {
        const float dx = P[j][0] - G[j][0], dy = P[j][1] - G[j][1], dz = P[j][2] - G[j][2];
        e += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    mpjpe[b] = e / 14.f;
    // Procrustes (pare/SPIN compute_similarity_transform): align P (S1) to G (S2)
    float mu1[3] = {0.f, 0.f, 0.f}, mu2[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < 14; ++j)
        for (int c = 0; c < 3; ++c) { mu1[c] += P[j][c]; mu2[c] += G[j][c]; }
    for (int c = 0; c < 3; ++c) { mu1[c] /= 14.f; mu2[c] /= 14.f; }
    float K[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    float var1 = 0.f;
    for (int j = 0; j < 14; ++j) {
        float x1[3], x2[3];
        for (int c = 0; c < 3; ++c) { x1[c] = P[j][c] - mu1[c]; x2[c] = G[j][c] - mu2[c]; var1 += x1[c] * x1[c]; }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) K[r][c] = fmaf(x1[r], x2[c], K[r][c]);          // K = X1 X2^T
    }
    // K = U S V^T : eigen-decomposition of K^T K = V S^2 V^T
    float KtK[3][3], V[3][3], lam[3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) KtK[r][c] = K[0][r] * K[0][c] + K[1][r] * K[1][c] + K[2][r] * K[2][c];
    jacobi_eig3(KtK, V, lam);
    int o0 = 0, o1 = 1, o2 = 2;                                   // sort descending
    if (lam[o0] < lam[o1]) { int t = o0; o0 = o1; o1 = t; }
    if (lam[o0] < lam[o2]) { int t = o0; o0 = o2; o2 = t; }
    if (lam[o1] < lam[o2]) { int t = o1; o1 = o2; o2 = t; }
    float v[3][3];                                                // columns v1,v2,v3 (as rows of v[])
    for (int k = 0; k < 3; ++k) { v[0][k] = V[k][o0]; v[1][k] = V[k][o1]; v[2][k] = V[k][o2]; }
    {   // keep det(V) = +1 after the permutation
        const float det = v[0][0] * (v[1][1] * v[2][2] - v[1][2] * v[2][1]) - v[0][1] * (v[1][0] * v[2][2] - v[1][2] * v[2][0]) +
                          v[0][2] * (v[1][0] * v[2][1] - v[1][1] * v[2][0]);
        if (det < 0.f) { v[2][0] = -v[2][0]; v[2][1] = -v[2][1]; v[2][2] = -v[2][2]; }
    }
    const float s1 = sqrtf(fmaxf(lam[o0], 0.f)), s2 = sqrtf(fmaxf(lam[o1], 0.f)), s3 = sqrtf(fmaxf(lam[o2], 0.f));
    float u1[3], u2[3], u3[3];
    for (int r = 0; r < 3; ++r) {
        u1[r] = (K[r][0] * v[0][0] + K[r][1] * v[0][1] + K[r][2] * v[0][2]) / fmaxf(s1, 1e-30f);
        u2[r] = (K[r][0] * v[1][0] + K[r][1] * v[1][1] + K[r][2] * v[1][2]) / fmaxf(s2, 1e-30f);
    }
    {   // re-orthonormalise u2 against u1 (fp32), u3 = u1 x u2
        const float n1 = rsqrtf(fmaxf(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2], 1e-30f));
        for (int r = 0; r < 3; ++r) u1[r] *= n1;
        const float d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
        for (int r = 0; r < 3; ++r) u2[r] -= d * u1[r];
        const float n2 = rsqrtf(fmaxf(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2], 1e-30f));
        for (int r = 0; r < 3; ++r) u2[r] *= n2;
        u3[0] = u1[1] * u2[2] - u1[2] * u2[1]; u3[1] = u1[2] * u2[0] - u1[0] * u2[2]; u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    }
    // R = V Z U^T with Z = diag(1,1,sign det(U V^T)) == v1 u1^T + v2 u2^T + v3 (u1 x u2)^T ; trace(R K) = s1 + s2 + d s3
    float R[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[r][c] = v[0][r] * u1[c] + v[1][r] * u2[c] + v[2][r] * u3[c];
    const float detK = K[0][0] * (K[1][1] * K[2][2] - K[1][2] * K[2][1]) - K[0][1] * (K[1][0] * K[2][2] - K[1][2] * K[2][0]) +
                       K[0][2] * (K[1][0] * K[2][1] - K[1][1] * K[2][0]);
    const float dsg = detK < 0.f ? -1.f : 1.f;
    const float scale = (s1 + s2 + dsg * s3) / fmaxf(var1, 1e-30f);
    float tr[3];
    for (int r = 0; r < 3; ++r) tr[r] = mu2[r] - scale * (R[r][0] * mu1[0] + R[r][1] * mu1[1] + R[r][2] * mu1[2]);
    float re = 0.f;
    for (int j = 0; j < 14; ++j) {
        float d2 = 0.f;
        for (int r = 0; r < 3; ++r) {
            const float h = scale * (R[r][0] * P[j][0] + R[r][1] * P[j][1] + R[r][2] * P[j][2]) + tr[r] - G[j][r];
            d2 += h * h;
        }
        re += sqrtf(d2);
    }
    pampjpe[b] = re / 14.f;
}

// ------------------------------------------------------------------ per-vertex error (compute_error_verts): mean_v |p_v - g_v|
__global__ void __launch_bounds__(256)
v2v_kernel(const float* __restrict__ pv, long long ldp, const float* __restrict__ gv, long long ldg,
           const float* __restrict__ pelvis /*[2][B][3] or null*/, float* __restrict__ out, int B)
{
    __shared__ float red[8];
    const int b = blockIdx.x;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (pelvis) {       // (p - pelvis_p) - (g - pelvis_g)
        ox = pelvis[b * 3 + 0] - pelvis[(static_cast<size_t>(B) + b) * 3 + 0];
        oy = pelvis[b * 3 + 1] - pelvis[(static_cast<size_t>(B) + b) * 3 + 1];
        oz = pelvis[b * 3 + 2] - pelvis[(static_cast<size_t>(B) + b) * 3 + 2];
    }
    const float* p = pv + b * ldp;
    const float* g = gv + b * ldg;
    float acc = 0.f;
    for (int v = threadIdx.x; v < SMPL_NV; v += 256) {
        const float dx = p[v * 3] - g[v * 3] - ox, dy = p[v * 3 + 1] - g[v * 3 + 1] - oy, dz = p[v * 3 + 2] - g[v * 3 + 2] - oz;
        acc += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        out[b] = s / static_cast<float>(SMPL_NV);
    }
}

bool eval_launch(const float* JT, const int* map14, int B, const float* pred_verts, long long ld_pred, const float* gt_kp14,
                 const float* gt_verts, long long ld_gt, int center_v2v, float* ws /*[2][B][51] + [2][B][3]*/, float* mpjpe,
                 float* pampjpe, float* v2v, float* pred_kp14, cudaStream_t s) {
    float* j17 = ws;
    float* pelvis = ws + static_cast<size_t>(2) * B * 51;
    const bool regress_gt = (gt_kp14 == nullptr);
    if (regress_gt && gt_verts == nullptr) { set_error("eval: need gt keypoints or gt vertices"); return false; }
    dim3 grid(B, regress_gt ? 2 : 1);
    h36m_joints_kernel<<<grid, 256, 0, s>>>(pred_verts, ld_pred, gt_verts, ld_gt, JT, j17, B);
    if (!check_cuda(cudaGetLastError(), "h36m_joints")) return false;
    eval_metrics_kernel<<<(B + 63) / 64, 64, 0, s>>>(j17, j17 + static_cast<size_t>(B) * 51, gt_kp14, map14, mpjpe, pampjpe, pred_kp14, pelvis, B);
    if (!check_cuda(cudaGetLastError(), "eval_metrics")) return false;
    if (v2v != nullptr && gt_verts != nullptr) {
        v2v_kernel<<<B, 256, 0, s>>>(pred_verts, ld_pred, gt_verts, ld_gt, (center_v2v && regress_gt) ? pelvis : nullptr, v2v, B);
        if (!check_cuda(cudaGetLastError(), "v2v")) return false;
    }
    return true;
}

}  // namespace sb

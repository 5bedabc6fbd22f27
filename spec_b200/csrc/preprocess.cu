// Input side of the SPEC demo loop on the device: uint8 full image -> network inputs.
//
//  * person crops: what /root/reference/spec/tester.py:118-125 gets from pare's get_single_image_crop_demo
//    (VIBE gen_trans_from_patch_cv -> cv2.getAffineTransform -> cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT) ->
//    ToTensor -> Normalize), one launch for all detections of an image instead of a CPU warp + H2D per detection;
//  * CamCalib input: transforms.Resize(min_size) (Pillow antialiased bilinear) -> ToTensor -> Normalize of
//    /root/reference/camcalib/pano_dataset.py:156-162.
//
// Both are integer/byte work and are reproduced BIT-EXACTLY: OpenCV's 8-bit warp uses 10-bit fixed-point source
// coordinates (round-half-even of double products), 5-bit bilinear fractions and 15-bit weights; Pillow's 8-bit
// resample uses 22-bit fixed-point triangle-filter coefficients, a horizontal pass into a uint8 image and then a
// vertical pass.  ToTensor+Normalize is a 3x256 float32 table built on the host with the same float32 operations.
// HBM-bound byte kernels (a 1080p frame is 6 MB: L2-resident), no tensor cores involved.
#include "common.cuh"
#include "internal.h"
#include "../../include/specb200.h"
#include <cmath>
#include <map>
#include <utility>
#include <vector>

namespace sb {

constexpr int CROP_CHUNK = 32;
struct CropBatch { double inv[CROP_CHUNK][6]; };     // dst -> src matrices, passed by value (1.5 KB of kernel params)

__global__ void __launch_bounds__(256)
crop_normalize_kernel(const uint8_t* __restrict__ img, int H, int W, long long stride, int bgr, const CropBatch mats, int cs,
                      const float* __restrict__ lut, float* __restrict__ out, uint8_t* __restrict__ raw)
{
    const int n = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cs * cs) return;
    const int y = idx / cs, x = idx - y * cs;
    const double* m = mats.inv[n];
    // cv::warpAffine: adelta/bdelta per column, X0/Y0 per row, AB_BITS = 10, round_delta = AB_SCALE / INTER_TAB_SIZE / 2
    const int ad = __double2int_rn(__dmul_rn(__dmul_rn(m[0], static_cast<double>(x)), 1024.0));
    const int bd = __double2int_rn(__dmul_rn(__dmul_rn(m[3], static_cast<double>(x)), 1024.0));
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], static_cast<double>(y)), m[2]), 1024.0)) + 16;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], static_cast<double>(y)), m[5]), 1024.0)) + 16;
    const int X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));   // stored as short upstream
    const int fx = X & 31, fy = Y & 31;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    int acc[3] = {1 << 14, 1 << 14, 1 << 14};
    const bool x0ok = static_cast<unsigned>(sx) < static_cast<unsigned>(W), x1ok = static_cast<unsigned>(sx + 1) < static_cast<unsigned>(W);
    const bool y0ok = static_cast<unsigned>(sy) < static_cast<unsigned>(H), y1ok = static_cast<unsigned>(sy + 1) < static_cast<unsigned>(H);
    const uint8_t* r0 = img + static_cast<long long>(sy) * stride + static_cast<long long>(sx) * 3;
    const uint8_t* r1 = r0 + stride;
    if (y0ok && x0ok) { acc[0] += w00 * r0[0]; acc[1] += w00 * r0[1]; acc[2] += w00 * r0[2]; }
    if (y0ok && x1ok) { acc[0] += w01 * r0[3]; acc[1] += w01 * r0[4]; acc[2] += w01 * r0[5]; }
    if (y1ok && x0ok) { acc[0] += w10 * r1[0]; acc[1] += w10 * r1[1]; acc[2] += w10 * r1[2]; }
    if (y1ok && x1ok) { acc[0] += w11 * r1[3]; acc[1] += w11 * r1[4]; acc[2] += w11 * r1[5]; }
    const size_t plane = static_cast<size_t>(cs) * cs;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int v = acc[bgr ? 2 - c : c] >> 15;                 // <= 255 by construction (weights sum to 32768)
        out[(static_cast<size_t>(n) * 3 + c) * plane + idx] = lut[c * 256 + v];
        if (raw != nullptr) raw[(static_cast<size_t>(n) * plane + idx) * 3 + c] = static_cast<uint8_t>(v);
    }
}

// Pillow ImagingResampleHorizontal_8bpc / Vertical_8bpc: one thread per output pixel (3 channels).
// src is [rows][in][3] with row pitch `pitch` bytes; the resampled axis is the middle one for the horizontal pass
// (axis_stride = 3 bytes) and the first one for the vertical pass (axis_stride = pitch).
__global__ void __launch_bounds__(256)
resample_kernel(const uint8_t* __restrict__ src, long long pitch, int vertical, int out_rows, int out_cols,
                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, uint8_t* __restrict__ dst)
{
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(out_rows) * out_cols) return;
    const int r = static_cast<int>(idx / out_cols), c = static_cast<int>(idx - static_cast<long long>(r) * out_cols);
    const int o = vertical ? r : c;                                  // index along the resampled axis
    const int first = bounds[2 * o], taps = bounds[2 * o + 1];
    const int* k = kk + static_cast<size_t>(o) * ksize;
    const uint8_t* p = vertical ? src + static_cast<long long>(first) * pitch + static_cast<long long>(c) * 3
                                : src + static_cast<long long>(r) * pitch + static_cast<long long>(first) * 3;
    const long long step = vertical ? pitch : 3;
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;                    // 1 << (PRECISION_BITS - 1)
    for (int t = 0; t < taps; ++t, p += step) {
        const int w = k[t];
        s0 += p[0] * w; s1 += p[1] * w; s2 += p[2] * w;
    }
    uint8_t* q = dst + idx * 3;
    q[0] = static_cast<uint8_t>(max(0, min(255, s0 >> 22)));
    q[1] = static_cast<uint8_t>(max(0, min(255, s1 >> 22)));
    q[2] = static_cast<uint8_t>(max(0, min(255, s2 >> 22)));
}

// ToTensor + Normalize of a uint8 [rows][cols][3] image -> float32 planar [3][rows][cols] (+ optional RGB uint8 copy)
__global__ void __launch_bounds__(256)
normalize_kernel(const uint8_t* __restrict__ src, long long pitch, int rows, int cols, int bgr, const float* __restrict__ lut,
                 float* __restrict__ out, uint8_t* __restrict__ raw)
{
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long plane = static_cast<long long>(rows) * cols;
    if (idx >= plane) return;
    const int r = static_cast<int>(idx / cols), c = static_cast<int>(idx - static_cast<long long>(r) * cols);
    const uint8_t* p = src + static_cast<long long>(r) * pitch + static_cast<long long>(c) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const int v = p[bgr ? 2 - ch : ch];
        out[ch * plane + idx] = lut[ch * 256 + v];
        if (raw != nullptr) raw[idx * 3 + ch] = static_cast<uint8_t>(v);
    }
}

// ------------------------------------------------------------------------------------------------ host arithmetic
// cv2.getAffineTransform: 6x6 system, OpenCV's LU with partial pivoting (matrix_decomp.cpp LUImpl), double.
static void get_affine_transform(const float src[3][2], const float dst[3][2], double M[6]) {
    double A[6][6] = {}, B[6] = {};
    for (int i = 0; i < 3; ++i) {
        A[2 * i][0] = A[2 * i + 1][3] = src[i][0];
        A[2 * i][1] = A[2 * i + 1][4] = src[i][1];
        A[2 * i][2] = A[2 * i + 1][5] = 1.0;
        B[2 * i] = dst[i][0];
        B[2 * i + 1] = dst[i][1];
    }
    for (int i = 0; i < 6; ++i) {
        int k = i;
        for (int j = i + 1; j < 6; ++j) if (std::fabs(A[j][i]) > std::fabs(A[k][i])) k = j;
        if (k != i) {
            for (int j = i; j < 6; ++j) std::swap(A[i][j], A[k][j]);
            std::swap(B[i], B[k]);
        }
        const double d = -1.0 / A[i][i];
        for (int j = i + 1; j < 6; ++j) {
            const double alpha = A[j][i] * d;
            for (int c = i + 1; c < 6; ++c) { const double t = alpha * A[i][c]; A[j][c] += t; }
            const double t = alpha * B[i];
            B[j] += t;
        }
    }
    for (int i = 5; i >= 0; --i) {
        double s = B[i];
        for (int c = i + 1; c < 6; ++c) { const double t = A[i][c] * B[c]; s -= t; }
        B[i] = s / A[i][i];
    }
    for (int i = 0; i < 6; ++i) M[i] = B[i];
}

// VIBE gen_trans_from_patch_cv (rot = 0, inv = False): float32 point triples exactly as the numpy code rounds them.
static void patch_transform(double c_x, double c_y, double bw, double bh, int dst, double scale, double M[6]) {
    const double sw = bw * scale, sh = bh * scale;
    const float down = static_cast<float>(sh * 0.5), right = static_cast<float>(sw * 0.5);
    const float src[3][2] = {{static_cast<float>(c_x), static_cast<float>(c_y)},
                             {static_cast<float>(c_x), static_cast<float>(c_y + static_cast<double>(down))},
                             {static_cast<float>(c_x + static_cast<double>(right)), static_cast<float>(c_y)}};
    const float dc = static_cast<float>(dst * 0.5), dd = static_cast<float>(dst * 0.5);
    const float dstp[3][2] = {{dc, dc}, {dc, dc + dd}, {dc + dd, dc}};
    get_affine_transform(src, dstp, M);
}

// the dst->src matrix cv::warpAffine derives (imgwarp.cpp), same operation order
static void invert_affine(const double M[6], double inv[6]) {
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1.0 / D : 0.0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    inv[0] = A11; inv[1] = M[1] * (-D); inv[3] = M[3] * (-D); inv[4] = A22;
    inv[2] = -inv[0] * M[2] - inv[1] * M[5];
    inv[5] = -inv[3] * M[2] - inv[4] * M[5];
}

// Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc, bilinear filter, box = whole axis
static int resample_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
    const double scale = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
    bounds.assign(static_cast<size_t>(out_size) * 2, 0);
    kk.assign(static_cast<size_t>(out_size) * ksize, 0);
    std::vector<double> w(ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            w[x] = a < 1.0 ? 1.0 - a : 0.0;
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) w[x] /= ww;
            kk[static_cast<size_t>(xx) * ksize + x] = w[x] < 0 ? static_cast<int>(-0.5 + w[x] * (1 << 22)) : static_cast<int>(0.5 + w[x] * (1 << 22));
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    return ksize;
}

struct CoeffTable { int* bounds = nullptr; int* kk = nullptr; int ksize = 0; };

}  // namespace sb

using namespace sb;

struct specb200_preproc {
    float* lut = nullptr;                                       // [3][256]
    std::map<std::pair<int, int>, CoeffTable> tables;           // (in, out) -> device coefficient table
};

static bool get_table(specb200_preproc* t, int in_size, int out_size, CoeffTable& out) {
    auto it = t->tables.find({in_size, out_size});
    if (it != t->tables.end()) { out = it->second; return true; }
    std::vector<int> bounds, kk;
    CoeffTable c;
    c.ksize = resample_coeffs(in_size, out_size, bounds, kk);
    if (!check_cuda(cudaMalloc(&c.bounds, bounds.size() * sizeof(int)), "cudaMalloc") ||
        !check_cuda(cudaMalloc(&c.kk, kk.size() * sizeof(int)), "cudaMalloc") ||
        !check_cuda(cudaMemcpy(c.bounds, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice), "upload") ||
        !check_cuda(cudaMemcpy(c.kk, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice), "upload"))
        return false;
    t->tables[{in_size, out_size}] = c;
    out = c;
    return true;
}

extern "C" int specb200_preproc_create(specb200_preproc_t** out, const float* mean3, const float* std3) {
    if (!out || !mean3 || !std3) { set_error("preproc_create: bad arguments"); return 1; }
    std::vector<float> lut(3 * 256);
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < 256; ++v) {
            volatile float x = static_cast<float>(v) / 255.0f;      // ToTensor: uint8 -> float32, .div(255)
            volatile float y = x - mean3[c];                         // Normalize: sub_(mean).div_(std), float32
            lut[c * 256 + v] = y / std3[c];
        }
    specb200_preproc* t = new specb200_preproc();
    if (!check_cuda(cudaMalloc(&t->lut, lut.size() * sizeof(float)), "cudaMalloc") ||
        !check_cuda(cudaMemcpy(t->lut, lut.data(), lut.size() * sizeof(float), cudaMemcpyHostToDevice), "upload")) {
        specb200_preproc_destroy(t);
        return 1;
    }
    *out = t;
    return 0;
}

extern "C" void specb200_preproc_destroy(specb200_preproc_t* t) {
    if (!t) return;
    if (t->lut) cudaFree(t->lut);
    for (auto& kv : t->tables) { cudaFree(kv.second.bounds); cudaFree(kv.second.kk); }
    delete t;
}

extern "C" int specb200_preproc_crop_transforms(const double* boxes_host, int32_t n, double scale, int32_t crop_size,
                                                double* trans_host, double* inv_host) {
    if (!boxes_host || n < 0 || crop_size <= 0) { set_error("preproc_crop_transforms: bad arguments"); return 1; }
    for (int i = 0; i < n; ++i) {
        double M[6], inv[6];
        patch_transform(boxes_host[4 * i], boxes_host[4 * i + 1], boxes_host[4 * i + 2], boxes_host[4 * i + 3], crop_size, scale, M);
        invert_affine(M, inv);
        for (int j = 0; j < 6; ++j) {
            if (trans_host) trans_host[6 * i + j] = M[j];
            if (inv_host) inv_host[6 * i + j] = inv[j];
        }
    }
    return 0;
}

extern "C" int specb200_preproc_crop(specb200_preproc_t* t, const uint8_t* image_dev, int32_t height, int32_t width,
                                     int64_t row_stride_bytes, int32_t bgr, const double* boxes_host, int32_t n, double scale,
                                     int32_t crop_size, float* out_dev, uint8_t* raw_dev, void* stream) {
    if (!t || !image_dev || !out_dev || height <= 0 || width <= 0 || n < 0 || crop_size <= 0 || (n > 0 && !boxes_host) ||
        row_stride_bytes < static_cast<int64_t>(width) * 3) { set_error("preproc_crop: bad arguments"); return 1; }
    if (height > 32767 || width > 32767) { set_error("preproc_crop: images above 32767 px per side are not supported (cv2 stores source coordinates as int16)"); return 1; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t plane = static_cast<size_t>(crop_size) * crop_size;
    for (int base = 0; base < n; base += CROP_CHUNK) {
        const int cnt = n - base < CROP_CHUNK ? n - base : CROP_CHUNK;
        CropBatch mats;
        for (int i = 0; i < cnt; ++i) {
            double M[6];
            const double* b = boxes_host + 4 * static_cast<size_t>(base + i);
            if (!(b[2] * scale > 0) || !(b[3] * scale > 0)) { set_error("preproc_crop: bbox width/height must be positive"); return 1; }
            patch_transform(b[0], b[1], b[2], b[3], crop_size, scale, M);
            invert_affine(M, mats.inv[i]);
        }
        dim3 grid(static_cast<unsigned>((plane + 255) / 256), cnt);
        crop_normalize_kernel<<<grid, 256, 0, s>>>(image_dev, height, width, row_stride_bytes, bgr, mats, crop_size, t->lut,
                                                   out_dev + static_cast<size_t>(base) * 3 * plane,
                                                   raw_dev ? raw_dev + static_cast<size_t>(base) * 3 * plane : nullptr);
        if (!check_cuda(cudaGetLastError(), "crop_normalize launch")) return 1;
    }
    return 0;
}

extern "C" int specb200_preproc_resized_shape(int32_t height, int32_t width, int32_t min_size, int32_t* out_h, int32_t* out_w) {
    if (height <= 0 || width <= 0 || min_size <= 0 || !out_h || !out_w) { set_error("preproc_resized_shape: bad arguments"); return 1; }
    // torchvision _compute_resized_output_size: the short side becomes min_size, the long one int(min_size * long / short)
    if (width <= height) { *out_w = min_size; *out_h = static_cast<int32_t>(static_cast<double>(min_size) * height / width); }
    else { *out_h = min_size; *out_w = static_cast<int32_t>(static_cast<double>(min_size) * width / height); }
    return 0;
}

extern "C" int64_t specb200_preproc_resize_workspace_bytes(int32_t height, int32_t width, int32_t out_h, int32_t out_w) {
    if (height <= 0 || width <= 0 || out_h <= 0 || out_w <= 0) { set_error("preproc_resize_workspace_bytes: bad arguments"); return -1; }
    return (static_cast<int64_t>(height) * out_w + static_cast<int64_t>(out_h) * out_w) * 3 + 512;
}

extern "C" int specb200_preproc_resize(specb200_preproc_t* t, const uint8_t* image_dev, int32_t height, int32_t width,
                                       int64_t row_stride_bytes, int32_t bgr, int32_t out_h, int32_t out_w, void* workspace_dev,
                                       int64_t workspace_bytes, float* out_dev, uint8_t* raw_dev, void* stream) {
    if (!t || !image_dev || !out_dev || height <= 0 || width <= 0 || out_h <= 0 || out_w <= 0 ||
        row_stride_bytes < static_cast<int64_t>(width) * 3) { set_error("preproc_resize: bad arguments"); return 1; }
    if (workspace_bytes < specb200_preproc_resize_workspace_bytes(height, width, out_h, out_w) || !workspace_dev) {
        set_error("preproc_resize: workspace too small");
        return 1;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint8_t* tmp_h = static_cast<uint8_t*>(workspace_dev);
    uint8_t* tmp_v = tmp_h + (static_cast<size_t>(height) * out_w * 3 + 255) / 256 * 256;
    const uint8_t* cur = image_dev;
    long long pitch = row_stride_bytes;
    if (out_w != width) {                                         // horizontal pass over every source row
        CoeffTable c;
        if (!get_table(t, width, out_w, c)) return 1;
        const long long total = static_cast<long long>(height) * out_w;
        resample_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(cur, pitch, 0, height, out_w, c.bounds, c.kk, c.ksize, tmp_h);
        if (!check_cuda(cudaGetLastError(), "resample (horizontal) launch")) return 1;
        cur = tmp_h;
        pitch = static_cast<long long>(out_w) * 3;
    }
    if (out_h != height) {                                        // vertical pass
        CoeffTable c;
        if (!get_table(t, height, out_h, c)) return 1;
        const long long total = static_cast<long long>(out_h) * out_w;
        resample_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(cur, pitch, 1, out_h, out_w, c.bounds, c.kk, c.ksize, tmp_v);
        if (!check_cuda(cudaGetLastError(), "resample (vertical) launch")) return 1;
        cur = tmp_v;
        pitch = static_cast<long long>(out_w) * 3;
    }
    const long long plane = static_cast<long long>(out_h) * out_w;
    normalize_kernel<<<static_cast<unsigned>((plane + 255) / 256), 256, 0, s>>>(cur, pitch, out_h, out_w, bgr, t->lut, out_dev, raw_dev);
    return check_cuda(cudaGetLastError(), "normalize launch") ? 0 : 1;
}

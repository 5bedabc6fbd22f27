"""CPU oracle for the INPUT side of the SPEC demo loop — TEST INFRASTRUCTURE, NOT PRODUCT.

Restates, in numpy integer arithmetic, the two image transforms the reference applies before its networks:

* person crops: ``get_single_image_crop_demo(img, bbox, kp_2d=None, scale=1.0, crop_size=224)`` called at
  /root/reference/spec/tester.py:118-125 (imported from ``pare.utils.vibe_image_utils``, tester.py:30; un-vendored).
  Upstream (VIBE ``lib/data_utils/img_utils.py``, recalled): ``gen_trans_from_patch_cv`` builds three float32 point
  pairs (centre, centre+down, centre+right), ``cv2.getAffineTransform`` -> ``cv2.warpAffine(img, trans, (224,224),
  flags=INTER_LINEAR, borderMode=BORDER_CONSTANT)`` -> ``ToTensor`` + ``Normalize(IMAGENET mean/std)``;
* CamCalib's input: ``transforms.Compose([Resize(min_size=600), ToTensor(), Normalize(...)])`` on a PIL image,
  /root/reference/camcalib/pano_dataset.py:156-162.

PARITY STATUS: **pinned**.  The arithmetic lives in OpenCV (``cv2.warpAffine`` 8-bit path: 10-bit fixed-point source
coordinates, 5-bit bilinear fractions, 15-bit weights) and Pillow (``ImagingResample`` 8-bit path: 22-bit fixed-point
triangle-filter coefficients, horizontal then vertical pass with a uint8 intermediate).  Both libraries are installed
in the build container (cv2 4.13.0, Pillow 12.2.0, torchvision 0.26.0), so these restatements are checked
BIT-EXACTLY against the library calls themselves: tests/golden/make_preprocess_golden.py commits library-made
vectors (tests/golden/preprocess.npz), tests/test_preprocess.py compares the oracle with them (and with the live
libraries when importable).
"""
import math

import numpy as np

IMG_NORM_MEAN = (0.485, 0.456, 0.406)      # /root/reference/spec/constants.py:20
IMG_NORM_STD = (0.229, 0.224, 0.225)       # /root/reference/spec/constants.py:21


# ------------------------------------------------------------------------------------------------ affine crop
def get_affine_transform_cv(src, dst):
    """cv2.getAffineTransform(src, dst) for three float32 point pairs: the 6x6 system solved by OpenCV's own LU with
    partial pivoting (modules/core/src/matrix_decomp.cpp LUImpl), same operation order, in double -> bit-equal M."""
    m = 6
    A = [[0.0] * 6 for _ in range(6)]
    B = [0.0] * 6
    for i in range(3):
        x, y = float(src[i][0]), float(src[i][1])
        A[2 * i][0], A[2 * i][1], A[2 * i][2] = x, y, 1.0
        A[2 * i + 1][3], A[2 * i + 1][4], A[2 * i + 1][5] = x, y, 1.0
        B[2 * i], B[2 * i + 1] = float(dst[i][0]), float(dst[i][1])
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(A[j][i]) > abs(A[k][i]):
                k = j
        if k != i:
            A[i], A[k] = A[i][:i] + A[k][i:], A[k][:i] + A[i][i:]
            B[i], B[k] = B[k], B[i]
        d = -1.0 / A[i][i]
        for j in range(i + 1, m):
            alpha = A[j][i] * d
            for c in range(i + 1, m):
                A[j][c] += alpha * A[i][c]
            B[j] += alpha * B[i]
    for i in range(m - 1, -1, -1):
        s = B[i]
        for c in range(i + 1, m):
            s -= A[i][c] * B[c]
        B[i] = s / A[i][i]
    return np.array(B, dtype=np.float64).reshape(2, 3)


def gen_trans_from_patch(c_x, c_y, src_w, src_h, dst_w, dst_h, scale=1.0):
    """VIBE gen_trans_from_patch_cv with rot=0, inv=False: three float32 point pairs (centre, centre+down,
    centre+right; the direction vectors are rounded to float32 BEFORE being added to the float64 centre, as the numpy
    code upstream does) -> cv2.getAffineTransform."""
    f32 = np.float32
    sw, sh = src_w * scale, src_h * scale
    down, right = float(f32(sh * 0.5)), float(f32(sw * 0.5))
    src = [(f32(c_x), f32(c_y)), (f32(c_x), f32(float(c_y) + down)), (f32(float(c_x) + right), f32(c_y))]
    dcx, dcy = f32(dst_w * 0.5), f32(dst_h * 0.5)
    dst = [(dcx, dcy), (dcx, f32(dcy + f32(dst_h * 0.5))), (f32(dcx + f32(dst_w * 0.5)), dcy)]
    return get_affine_transform_cv(src, dst)


def invert_affine_cv(M):
    """The dst->src matrix cv2.warpAffine derives from a forward matrix (imgwarp.cpp, !WARP_INVERSE_MAP), same op order."""
    M = np.asarray(M, dtype=np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    i0, i1, i3, i4 = A11, M[0, 1] * (-D), M[1, 0] * (-D), A22
    b1 = -i0 * M[0, 2] - i1 * M[1, 2]
    b2 = -i3 * M[0, 2] - i4 * M[1, 2]
    return np.array([i0, i1, b1, i3, i4, b2], dtype=np.float64)


def warp_affine_u8(img, inv6, dst_w, dst_h):
    """cv2.warpAffine(img, M, (dst_w, dst_h), INTER_LINEAR, BORDER_CONSTANT=0) for uint8 HxWxC given the INVERSE
    matrix: AB_BITS=10 fixed-point coordinates with round_delta 16, INTER_BITS=5 fractions, weights
    (32-fx)(32-fy)*32 / 32768 (exact on this grid, so the table-normalisation branch of OpenCV never fires)."""
    H, W = img.shape[:2]
    x = np.arange(dst_w, dtype=np.float64)
    y = np.arange(dst_h, dtype=np.float64)
    ad = np.rint(inv6[0] * x * 1024.0).astype(np.int64)
    bd = np.rint(inv6[3] * x * 1024.0).astype(np.int64)
    X0 = np.rint((inv6[1] * y + inv6[2]) * 1024.0).astype(np.int64) + 16
    Y0 = np.rint((inv6[4] * y + inv6[5]) * 1024.0).astype(np.int64) + 16
    X = (X0[:, None] + ad[None, :]) >> 5
    Y = (Y0[:, None] + bd[None, :]) >> 5
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    fx, fy = X & 31, Y & 31
    acc = np.zeros((dst_h, dst_w, img.shape[2]), dtype=np.int64)
    for dy, dx, w in ((0, 0, (32 - fx) * (32 - fy)), (0, 1, fx * (32 - fy)), (1, 0, (32 - fx) * fy), (1, 1, fx * fy)):
        yy, xx = sy + dy, sx + dx
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        p = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
        acc += (w * 32 * ok)[..., None] * p
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


def normalize_lut(mean=IMG_NORM_MEAN, std=IMG_NORM_STD):
    """ToTensor (uint8 -> float32 / 255) followed by Normalize ((x - mean) / std), all in float32: [3][256]."""
    v = np.arange(256, dtype=np.float32) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32)[:, None]
    s = np.asarray(std, dtype=np.float32)[:, None]
    return ((v[None, :] - m) / s).astype(np.float32)


def to_tensor_normalize(img_u8, mean=IMG_NORM_MEAN, std=IMG_NORM_STD):
    lut = normalize_lut(mean, std)
    return np.stack([lut[c][img_u8[..., c]] for c in range(3)], axis=0)


def get_single_image_crop_demo(image, bbox, kp_2d=None, scale=1.0, crop_size=224):
    """-> (norm_img (3, cs, cs) float32, raw_img (cs, cs, 3) uint8, kp_2d).  bbox = (c_x, c_y, w, h) in pixels."""
    M = gen_trans_from_patch(bbox[0], bbox[1], bbox[2], bbox[3], crop_size, crop_size, scale)
    raw = warp_affine_u8(image, invert_affine_cv(M), crop_size, crop_size)
    if kp_2d is not None:
        kp_2d = np.array(kp_2d, dtype=np.float64, copy=True)
        pts = np.concatenate([kp_2d[:, :2], np.ones((kp_2d.shape[0], 1))], axis=1)
        kp_2d[:, :2] = pts @ M.T
    return to_tensor_normalize(raw), raw, kp_2d


# ------------------------------------------------------------------------------------------------ PIL resize
def pil_bilinear_coeffs(in_size, out_size):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR (triangle) filter, box = full
    image: per output index the first source index, the tap count and the 22-bit fixed-point weights."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = []
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            v = 1.0 - a if a < 1.0 else 0.0
            w.append(v)
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis1(a, bounds, kk):
    out = np.empty((a.shape[0], bounds.shape[0], a.shape[2]), dtype=np.uint8)
    for xx in range(bounds.shape[0]):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full((a.shape[0], a.shape[2]), 1 << 21, dtype=np.int64)
        for x in range(n):
            acc += a[:, x0 + x, :].astype(np.int64) * int(kk[xx, x])
        out[:, xx, :] = np.clip(acc >> 22, 0, 255)
    return out


def pil_resize_bilinear_u8(img, out_h, out_w):
    """PIL.Image.resize((out_w, out_h), BILINEAR) for an 8-bit HxWxC image: horizontal pass, then vertical pass."""
    H, W = img.shape[:2]
    t = img
    if out_w != W:
        t = _resample_axis1(t, *pil_bilinear_coeffs(W, out_w))
    if out_h != H:
        t = _resample_axis1(t.transpose(1, 0, 2), *pil_bilinear_coeffs(H, out_h)).transpose(1, 0, 2)
    return np.ascontiguousarray(t)


def resized_shape(h, w, min_size):
    """torchvision.transforms.Resize(int) output size: the short side becomes min_size, the long side is truncated."""
    if w <= h:
        return int(min_size * h / w), min_size
    return min_size, int(min_size * w / h)


def camcalib_transform(image, min_size=600):
    """Resize(min_size) + ToTensor + Normalize of camcalib/pano_dataset.py:156-162 -> (3, oh, ow) float32."""
    oh, ow = resized_shape(image.shape[0], image.shape[1], min_size)
    return to_tensor_normalize(pil_resize_bilinear_u8(image, oh, ow))

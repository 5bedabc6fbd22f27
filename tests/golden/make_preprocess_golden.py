"""Generates tests/golden/preprocess.npz with the LIBRARIES the reference's input path calls -- cv2.getAffineTransform
/ cv2.warpAffine (through VIBE's gen_trans_from_patch_cv, restated below from upstream since pare is not installable),
PIL.Image.resize via torchvision.transforms.Resize, ToTensor, Normalize -- so that the oracle (oracle/preprocess.py)
and the CUDA kernels are pinned bit-exactly against library output, not against each other.

    python -m tests.golden.make_preprocess_golden      (needs cv2, Pillow, torchvision: present in the build container;
                                                        made with cv2 4.13.0, Pillow 12.2.0, torchvision 0.26.0)
"""
import os

import numpy as np

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def frame(h, w, seed):
    """Seeded uint8 frame with both smooth structure and noise (so that interpolation errors are visible)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([127 + 120 * np.sin(xx / 7.0 + c) * np.cos(yy / 5.0 - c) for c in range(3)], axis=-1)
    return np.clip(base + rng.integers(-40, 41, (h, w, 3)), 0, 255).astype(np.uint8)


def vibe_trans(c_x, c_y, src_width, src_height, dst_width, dst_height, scale, rot=0):
    import cv2

    def rotate_2d(pt, rot_rad):
        x, y = pt[0], pt[1]
        sn, cs = np.sin(rot_rad), np.cos(rot_rad)
        return np.array([x * cs - y * sn, x * sn + y * cs], dtype=np.float32)

    src_w, src_h = src_width * scale, src_height * scale
    src_center = np.zeros(2)
    src_center[0], src_center[1] = c_x, c_y
    rot_rad = np.pi * rot / 180
    src_downdir = rotate_2d(np.array([0, src_h * 0.5], dtype=np.float32), rot_rad)
    src_rightdir = rotate_2d(np.array([src_w * 0.5, 0], dtype=np.float32), rot_rad)
    dst_center = np.array([dst_width * 0.5, dst_height * 0.5], dtype=np.float32)
    dst_downdir = np.array([0, dst_height * 0.5], dtype=np.float32)
    dst_rightdir = np.array([dst_width * 0.5, 0], dtype=np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    src[0, :], src[1, :], src[2, :] = src_center, src_center + src_downdir, src_center + src_rightdir
    dst = np.zeros((3, 2), dtype=np.float32)
    dst[0, :], dst[1, :], dst[2, :] = dst_center, dst_center + dst_downdir, dst_center + dst_rightdir
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def library_crop(img, bbox, scale, crop_size):
    import cv2
    import torchvision.transforms as T
    trans = vibe_trans(bbox[0], bbox[1], bbox[2], bbox[3], crop_size, crop_size, scale)
    raw = cv2.warpAffine(img.copy(), trans, (crop_size, crop_size), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    norm = T.Compose([T.ToTensor(), T.Normalize(mean=MEAN, std=STD)])(raw).numpy()
    return raw, norm, trans


def library_resize(img, min_size):
    from PIL import Image
    import torchvision.transforms as T
    pil = T.Resize(min_size)(Image.fromarray(img))
    norm = T.Compose([T.ToTensor(), T.Normalize(mean=MEAN, std=STD)])(pil).numpy()
    return np.asarray(pil).copy(), norm


CROP_FRAME = (135, 240, 11)
CROP_BOXES_64 = [  # (c_x, c_y, w, h), scale
    ((120.0, 67.0, 64.0, 64.0), 1.0),          # identity: integer centre, box == crop
    ((120.5, 67.25, 64.0, 64.0), 1.0),         # fractional shift
    ((100.3, 60.7, 100.0, 100.0), 1.0),        # downscale
    ((100.3, 60.7, 100.0, 100.0), 1.2),        # VIBE default scale
    ((30.0, 20.0, 90.0, 90.0), 1.0),           # over the top-left border
    ((230.0, 125.0, 70.5, 70.5), 1.0),         # over the bottom-right border
    ((120.0, 67.0, 400.0, 400.0), 1.0),        # frame inside the crop
    ((-80.0, -80.0, 50.0, 50.0), 1.0),         # completely outside -> zeros
    ((117.77, 64.31, 23.9, 23.9), 1.0),        # upscale
    ((64.0, 64.0, 128.0, 128.0), 1.0),         # exact 2x downscale (ties in the fixed-point rounding)
]
CROP_BOX_224 = ((128.4, 70.2, 130.0, 130.0), 1.0)
RESIZE_CASES = [((135, 240, 21), 75), ((101, 77, 22), 60), ((40, 56, 23), 64), ((60, 80, 24), 60), ((90, 90, 25), 48)]


def main():
    out = {}
    img = frame(*CROP_FRAME)
    raws, transs = [], []
    for box, scale in CROP_BOXES_64:
        raw, norm, trans = library_crop(img, box, scale, 64)
        raws.append(raw)
        transs.append(trans)
    out['crop64_raw'] = np.stack(raws)
    out['crop64_trans'] = np.stack(transs)
    raw, norm, trans = library_crop(img, CROP_BOX_224[0], CROP_BOX_224[1], 224)
    out['crop224_raw'], out['crop224_norm'], out['crop224_trans'] = raw, norm, trans
    for i, (fr, ms) in enumerate(RESIZE_CASES):
        raw, norm = library_resize(frame(*fr), ms)
        out[f'resize{i}_raw'] = raw
        if i == 0:
            out['resize0_norm'] = norm
    import cv2, PIL, torchvision
    out['versions'] = np.array([cv2.__version__, PIL.__version__, torchvision.__version__])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'preprocess.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()

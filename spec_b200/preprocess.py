"""Input side of the demo loop on the GPU: uint8 frame -> network inputs.

Replaces, per frame, the CPU work of /root/reference/spec/tester.py:105-125 (``cv2.warpAffine`` crop + ``ToTensor`` +
``Normalize`` + one H2D copy PER DETECTION, through pare's ``get_single_image_crop_demo``) with one kernel launch over
all detections of a frame that is already on the device, and CamCalib's ``Resize(600) + ToTensor + Normalize``
(/root/reference/camcalib/pano_dataset.py:156-162) with three byte kernels.  Results are bit-identical to
cv2 / Pillow / torchvision (8-bit fixed-point paths restated in spec_b200/csrc/preprocess.cu); the frame must be a CUDA
uint8 tensor -- there is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

IMG_NORM_MEAN = (0.485, 0.456, 0.406)      # /root/reference/spec/constants.py:20
IMG_NORM_STD = (0.229, 0.224, 0.225)       # /root/reference/spec/constants.py:21


def _check_frame(image):
    _lib.require_device(image)
    if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3:
        raise ValueError(f'expected a uint8 (H, W, 3) frame, got {image.dtype} {tuple(image.shape)}')
    if image.stride(2) != 1 or image.stride(1) != 3:
        image = image.contiguous()
    return image


class Preprocessor:
    """Owns the device tables (normalisation LUT, cached Pillow coefficient tables); one per process is enough."""

    def __init__(self, mean=IMG_NORM_MEAN, std=IMG_NORM_STD):
        self.mean = np.asarray(mean, dtype=np.float32)
        self.std = np.asarray(std, dtype=np.float32)
        self._handles = {}
        self._ws = {}

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.lib().specb200_preproc_destroy(h)
        except Exception:
            pass

    def _handle(self, device):
        h = self._handles.get(device)
        if h is None:
            h = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(_lib.lib().specb200_preproc_create(C.byref(h), self.mean.ctypes.data, self.std.ctypes.data))
            self._handles[device] = h
        return h

    # -------------------------------------------------------------------------------------------- person crops
    @torch.no_grad()
    def crop(self, image, bboxes, scale=1.0, crop_size=224, bgr=False, return_raw=False):
        """image: CUDA uint8 (H, W, 3); bboxes: (N, 4) array-like of (c_x, c_y, w, h) pixels (host values, as the
        detector hands them over at tester.py:101,116) -> float32 (N, 3, crop, crop) [, uint8 (N, crop, crop, 3)]."""
        image = _check_frame(image)
        boxes = np.ascontiguousarray(np.asarray(bboxes, dtype=np.float64).reshape(-1, 4))
        n = boxes.shape[0]
        out = torch.empty(n, 3, crop_size, crop_size, dtype=torch.float32, device=image.device)
        raw = torch.empty(n, crop_size, crop_size, 3, dtype=torch.uint8, device=image.device) if return_raw else None
        if n == 0:                                           # a frame without detections (tester.py:102-103)
            return (out, raw) if return_raw else out
        with torch.cuda.device(image.device):
            _lib.check(_lib.lib().specb200_preproc_crop(
                self._handle(image.device), image.data_ptr(), image.shape[0], image.shape[1], image.stride(0), int(bool(bgr)),
                boxes.ctypes.data, n, float(scale), crop_size, out.data_ptr(), raw.data_ptr() if raw is not None else None,
                torch.cuda.current_stream().cuda_stream))
        return (out, raw) if return_raw else out

    # -------------------------------------------------------------------------------------------- CamCalib input
    @torch.no_grad()
    def resize(self, image, min_size=600, bgr=False, return_raw=False):
        """-> float32 (1, 3, oh, ow) with the short side resized to ``min_size`` (antialiased bilinear, Pillow rule)."""
        image = _check_frame(image)
        H, W = image.shape[:2]
        oh, ow = resized_shape(H, W, min_size)
        l = _lib.lib()
        need = l.specb200_preproc_resize_workspace_bytes(H, W, oh, ow)
        key = image.device
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=image.device)
            self._ws[key] = ws
        out = torch.empty(1, 3, oh, ow, dtype=torch.float32, device=image.device)
        raw = torch.empty(oh, ow, 3, dtype=torch.uint8, device=image.device) if return_raw else None
        with torch.cuda.device(image.device):
            _lib.check(l.specb200_preproc_resize(
                self._handle(image.device), image.data_ptr(), H, W, image.stride(0), int(bool(bgr)), oh, ow, ws.data_ptr(),
                ws.numel(), out.data_ptr(), raw.data_ptr() if raw is not None else None,
                torch.cuda.current_stream().cuda_stream))
        return (out, raw) if return_raw else out


def resized_shape(height, width, min_size):
    """torchvision ``Resize(int)`` output size (host arithmetic inside the library)."""
    oh, ow = C.c_int32(), C.c_int32()
    _lib.check(_lib.lib().specb200_preproc_resized_shape(int(height), int(width), int(min_size), C.byref(oh), C.byref(ow)))
    return oh.value, ow.value


def crop_transforms(bboxes, scale=1.0, crop_size=224):
    """The forward 2x3 ``trans`` matrices of the crops (what ``get_single_image_crop_demo`` maps kp_2d with) and
    the dst->src matrices cv2.warpAffine derives from them: two float64 (N, 2, 3) arrays.  Host only."""
    boxes = np.ascontiguousarray(np.asarray(bboxes, dtype=np.float64).reshape(-1, 4))
    n = boxes.shape[0]
    trans = np.empty((n, 6), dtype=np.float64)
    inv = np.empty((n, 6), dtype=np.float64)
    _lib.check(_lib.lib().specb200_preproc_crop_transforms(boxes.ctypes.data, n, float(scale), int(crop_size),
                                                          trans.ctypes.data, inv.ctypes.data))
    return trans.reshape(n, 2, 3), inv.reshape(n, 2, 3)


_default = None


def default_preprocessor():
    global _default
    if _default is None:
        _default = Preprocessor()
    return _default


def get_single_image_crop_demo(image, bbox, kp_2d=None, scale=1.2, crop_size=224):
    """Same name, argument order and return triple as pare.utils.vibe_image_utils.get_single_image_crop_demo
    (tester.py:30,118-125; called there with scale=1.0) for a frame that lives on the GPU:
    -> (norm_img (3, cs, cs) float32 CUDA, raw_img (cs, cs, 3) uint8 CUDA, kp_2d)."""
    norm, raw = default_preprocessor().crop(image, [bbox], scale=scale, crop_size=crop_size, return_raw=True)
    if kp_2d is not None:
        trans, _ = crop_transforms([bbox], scale, crop_size)
        kp_2d = np.array(kp_2d, dtype=np.float64, copy=True)
        pts = np.concatenate([kp_2d[:, :2], np.ones((kp_2d.shape[0], 1))], axis=1)
        kp_2d[:, :2] = pts @ trans[0].T
    return norm[0], raw[0], kp_2d


def camcalib_transform(image, min_size=600):
    """``data_transform`` of camcalib/pano_dataset.py:156-162 for a CUDA uint8 frame -> (3, oh, ow) float32."""
    return default_preprocessor().resize(image, min_size=min_size)[0]

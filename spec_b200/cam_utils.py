"""CamCalib decode and the CamCalib -> SPEC glue, on the GPU.

Mirrors /root/reference/camcalib/cam_utils.py:110-145 (``convert_preds_to_angles`` -- the
``softargmax_l2`` / ``softargmax_biased_l2`` branch used by scripts/camcalib_demo.py:227),
/root/reference/scripts/camcalib_demo.py:127-129 (``f_pix``) and
/root/reference/spec/utils/cam_params.py:24-50 (``read_cam_params``: R = euler(pitch,0,roll), K with
K[2,2] left 0) -- one fused kernel instead of softmax/arange/sum + a pkl round trip.
"""
import torch

from . import _lib

VFOV_RANGE = (0.2617, 2.1)      # cam_utils.py:55
PITCH_RANGE = (-0.6, 0.6)       # cam_utils.py:39
ROLL_RANGE = (-0.6, 0.6)        # cam_utils.py:133


def _as_f32(x, batch, device):
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    x = x.to(device=device, dtype=torch.float32)
    if x.dim() == 0:
        x = x.expand(batch)
    return x.contiguous()


def decode_logits(logits, img_h=None, img_w=None):
    """logits: (B, 3*D) fp32 [vfov|pitch|roll].  Returns angles (B,3) and, when img_h/img_w are given,
    cam_rotmat (B,3,3), cam_intrinsics (B,3,3), f_pix (B,)."""
    _lib.require_device(logits)
    B, D3 = logits.shape
    D = D3 // 3
    dev = logits.device
    angles = torch.empty(B, 3, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    if img_h is None:
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().specb200_camcalib_decode(logits.data_ptr(), logits.stride(0), D, B, 0, 0,
                                                           angles.data_ptr(), 0, 0, 0, stream))
        return angles
    ih, iw = _as_f32(img_h, B, dev), _as_f32(img_w, B, dev)
    R = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
    K = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
    f = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().specb200_camcalib_decode(logits.data_ptr(), logits.stride(0), D, B, ih.data_ptr(),
                                                       iw.data_ptr(), angles.data_ptr(), R.data_ptr(), K.data_ptr(),
                                                       f.data_ptr(), stream))
    return angles, R, K, f


@torch.no_grad()
def convert_preds_to_angles(pred_vfov, pred_pitch, pred_roll, loss_type='softargmax_l2', return_type='torch',
                            legacy=False):
    """Same signature as the reference function.  Only the soft-argmax branch is on the hot path; the
    legacy argmax-bins branch (cam_utils.py:123-126) is out of scope."""
    if loss_type not in ('softargmax_l2', 'softargmax_biased_l2') or legacy:
        raise NotImplementedError('only the soft-argmax decode is part of the SPEC hot path')
    logits = torch.cat([pred_vfov, pred_pitch, pred_roll], 1).float().contiguous()
    ang = decode_logits(logits)
    out = ang[:, 0], ang[:, 1], ang[:, 2]
    if return_type == 'np':
        return tuple(o.cpu().numpy() for o in out)
    return out


# ---- CamCalib <-> SPEC wire format: the reference hands camera parameters from camcalib_demo.py to
# spec_demo.py through one pickle per image; the fused in-process path does not need it but can still produce it.
def save_camcalib_pkl(output_path, image_names, angles, f_pix):
    """Writes ``<output_path>/camcalib/<basename>.pkl`` = {'vfov','f_pix','pitch','roll'} exactly as
    /root/reference/scripts/camcalib_demo.py:135-140,174 does (numpy scalars), from the tensors returned by
    ``CameraRegressorNetwork.predict_camera``."""
    import os
    import joblib
    import numpy as np
    os.makedirs(os.path.join(output_path, 'camcalib'), exist_ok=True)
    a = angles.detach().float().cpu().numpy()
    f = f_pix.detach().float().cpu().numpy()
    out = []
    for i, name in enumerate(image_names):
        p = os.path.join(output_path, 'camcalib', os.path.basename(name) + '.pkl')
        joblib.dump({'vfov': np.float32(a[i, 0]), 'f_pix': np.float32(f[i]), 'pitch': np.float32(a[i, 1]),
                     'roll': np.float32(a[i, 2])}, p)
        out.append(p)
    return out


def load_camcalib_pkl(output_path, img_fname, orig_shape):
    """Host mirror of /root/reference/spec/utils/cam_params.py:24-50 (``read_cam_params``) for files written above:
    returns (cam_rotmat, cam_int, vfov, pitch, roll, focal_length) as CPU tensors / floats.  The rotation is built with
    the same quaternion formula as the device kernel (tail.cu::euler_to_rotmat)."""
    import math
    import os
    import joblib
    d = joblib.load(os.path.join(output_path, 'camcalib', os.path.basename(img_fname) + '.pkl'))
    pitch, roll, vfov, f = float(d['pitch']), float(d['roll']), float(d['vfov']), float(d['f_pix'])
    hx, hz = pitch * 0.5, roll * 0.5
    qw, qx, qy, qz = math.cos(hx) * math.cos(hz), math.sin(hx) * math.cos(hz), -math.sin(hx) * math.sin(hz), math.cos(hx) * math.sin(hz)
    n = math.sqrt(qw * qw + qx * qx + qy * qy + qz * qz)
    qw, qx, qy, qz = qw / n, qx / n, qy / n, qz / n
    R = torch.tensor([[qw * qw + qx * qx - qy * qy - qz * qz, 2 * qx * qy - 2 * qw * qz, 2 * qw * qy + 2 * qx * qz],
                      [2 * qw * qz + 2 * qx * qy, qw * qw - qx * qx + qy * qy - qz * qz, 2 * qy * qz - 2 * qw * qx],
                      [2 * qx * qz - 2 * qw * qy, 2 * qw * qx + 2 * qy * qz, qw * qw - qx * qx - qy * qy + qz * qz]], dtype=torch.float32)
    K = torch.zeros(3, 3)
    K[0, 0] = K[1, 1] = f
    K[0, 2], K[1, 2] = orig_shape[1] / 2, orig_shape[0] / 2
    return R, K, vfov, pitch, roll, f

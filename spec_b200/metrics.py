"""Eval-side metrics on the GPU.

Replaces what /root/reference/spec/trainer.py:272-316 and /root/reference/spec/utils/compute_error.py:33-86 do per
validation batch -- ``J_regressor_h36m @ vertices`` -> 14 LSP joints (``H36M_TO_J14``) -> pelvis centring -> MPJPE,
Procrustes-aligned MPJPE (``reconstruction_error``: numpy SVD per sample on the host) and per-vertex error
(``compute_error_verts``) -- without copying the 21 MB of vertices per batch to the host.  Units are the inputs' (the
reference multiplies by 1000 afterwards).
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .constants import H36M_TO_J14


class EvalMetrics(nn.Module):
    """``J_regressor`` is the (17, 6890) H36M regressor the reference loads from ``data/J_regressor_h36m.npy`` and
    registers as the buffer ``J_regressor`` (spec/trainer.py:96-99)."""

    def __init__(self, J_regressor_h36m, joint_mapper=H36M_TO_J14):
        super().__init__()
        J = torch.as_tensor(np.asarray(J_regressor_h36m)).float()
        assert J.shape == (17, 6890)
        self.register_buffer('J_regressor', J)
        self.joint_mapper = list(joint_mapper)
        self._handle = None
        self._device = None
        self._ws = None

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().specb200_eval_destroy(self._handle)
        except Exception:
            pass

    def _ensure(self, device):
        if self._handle is not None and self._device == device:
            return
        _lib.require_device()
        J = self.J_regressor.detach().float().contiguous().cpu()
        m = torch.tensor(self.joint_mapper, dtype=torch.int32)
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib().specb200_eval_create(C.byref(h), J.data_ptr(), m.data_ptr()))
        self._handle, self._device = h, device

    @torch.no_grad()
    def forward(self, pred_vertices, gt_keypoints_3d=None, gt_vertices=None, center_v2v=False):
        """pred_vertices (B,6890,3) [may be a strided view of the packed record].  Give either ``gt_keypoints_3d``
        (B,14,3) as the trainer does, or ``gt_vertices`` (B,6890,3) as compute_error.py does (or both: keypoints for
        the joint errors, vertices for v2v).  Returns dict of (B,) tensors + ``pred_keypoints_3d`` (B,14,3)."""
        _lib.require_device(pred_vertices)
        dev = pred_vertices.device
        B = pred_vertices.shape[0]
        if pred_vertices.stride(2) != 1 or pred_vertices.stride(1) != 3 or pred_vertices.dtype != torch.float32:
            pred_vertices = pred_vertices.float().contiguous()
        if gt_keypoints_3d is None and gt_vertices is None:
            raise ValueError('need gt_keypoints_3d or gt_vertices')
        if center_v2v and gt_keypoints_3d is not None:
            # the centred vertex error (compute_error.py) subtracts the pelvis regressed from BOTH meshes; with ground-truth
            # keypoints supplied the kernel takes the joint errors from them and never regresses the GT pelvis
            raise ValueError('center_v2v=True needs the joints regressed from gt_vertices: do not pass gt_keypoints_3d with it')
        self._ensure(dev)
        L = _lib.lib()
        kp = gt_keypoints_3d.to(dev, torch.float32).contiguous() if gt_keypoints_3d is not None else None
        gv = gt_vertices.to(dev, torch.float32).contiguous() if gt_vertices is not None else None
        n = L.specb200_eval_workspace_bytes(self._handle, B)
        if self._ws is None or self._ws.numel() < n or self._ws.device != dev:
            self._ws = torch.empty(n, dtype=torch.uint8, device=dev)
        out = {k: torch.empty(B, dtype=torch.float32, device=dev) for k in ('mpjpe', 'pa_mpjpe')}
        v2v = torch.empty(B, dtype=torch.float32, device=dev) if gv is not None else None
        pk = torch.empty(B, 14, 3, dtype=torch.float32, device=dev)
        ptr = lambda t: t.data_ptr() if t is not None else 0
        with torch.cuda.device(dev):
            _lib.check(L.specb200_eval_forward(self._handle, B, pred_vertices.data_ptr(), pred_vertices.stride(0), ptr(kp), ptr(gv),
                                               gv.stride(0) if gv is not None else 0, int(bool(center_v2v)), self._ws.data_ptr(),
                                               self._ws.numel(), out['mpjpe'].data_ptr(), out['pa_mpjpe'].data_ptr(), ptr(v2v),
                                               pk.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        if v2v is not None:
            out['v2v'] = v2v
        out['pred_keypoints_3d'] = pk
        return out

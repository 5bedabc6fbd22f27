"""``CameraRegressorNetwork`` -- drop-in for /root/reference/camcalib/model.py:24-81.

Same constructor arguments, same ``forward(images) -> [vfov, pitch, roll]`` (three (B,256) logit
tensors), same state_dict keys (``backbone.*``, ``fc_vfov.*``, ``fc_pitch.*``, ``fc_roll.*``; the
multi-layer variant keeps ``fc_*.{i}.*``).  Compute: libspecb200 trunk -> global-average pool ->
one fused fp32 GEMM for the three heads.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from . import backbone as _bb
from .cam_utils import decode_logits


class CameraRegressorNetwork(nn.Module):
    def __init__(self, backbone='resnet50', num_fc_layers=1, num_fc_channels=1024, num_out_channels=256):
        super().__init__()
        if backbone not in _bb._RESNETS:
            raise ValueError(f'unsupported CamCalib backbone {backbone}')
        self.backbone = getattr(_bb, backbone)(pretrained=True)
        self.num_out_channels = num_out_channels
        out_channels = _bb.get_backbone_info(backbone)['n_output_channels']
        assert num_fc_layers > 0, 'Number of FC layers should be more than 0'
        if num_fc_layers == 1:
            self.fc_vfov = nn.Linear(out_channels, num_out_channels)
            self.fc_pitch = nn.Linear(out_channels, num_out_channels)
            self.fc_roll = nn.Linear(out_channels, num_out_channels)
            for fc in (self.fc_vfov, self.fc_pitch, self.fc_roll):       # model.py:45-52
                nn.init.normal_(fc.weight, mean=0, std=0.01)
                nn.init.constant_(fc.bias, 0)
        else:
            self.fc_vfov = self._get_fc_layers(num_fc_layers, num_fc_channels, out_channels)
            self.fc_pitch = self._get_fc_layers(num_fc_layers, num_fc_channels, out_channels)
            self.fc_roll = self._get_fc_layers(num_fc_layers, num_fc_channels, out_channels)
        self._handle = None
        self._dirty = True
        self._watch = None
        self._ws = None
        self.register_load_state_dict_post_hook(lambda m, k: m._mark_dirty())

    def _get_fc_layers(self, num_layers, num_channels, inp_channels):     # model.py:54-70 (no activations)
        mods = []
        for i in range(num_layers):
            if i == 0:
                mods.append(nn.Linear(inp_channels, num_channels))
            elif i == num_layers - 1:
                mods.append(nn.Linear(num_channels, self.num_out_channels))
            else:
                mods.append(nn.Linear(num_channels, num_channels))
        return nn.Sequential(*mods)

    def _mark_dirty(self):
        self._dirty = True

    def _apply(self, fn, *a, **k):
        self._dirty = True
        return super()._apply(fn, *a, **k)

    def _release(self):
        if self._handle is not None:
            _lib.lib().specb200_camtail_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure(self, device):
        heads = (self.fc_vfov, self.fc_pitch, self.fc_roll)
        if self._handle is not None and not self._dirty and self._device == device and not self._watch.changed():
            return
        _lib.require_device()
        self._release()
        L = _lib.lib()
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.specb200_camtail_create(C.byref(h), self.backbone.n_output_channels, self.num_out_channels))
            self._handle = h
            for which, fc in enumerate((self.fc_vfov, self.fc_pitch, self.fc_roll)):
                for lin in ([fc] if isinstance(fc, nn.Linear) else list(fc)):
                    w = lin.weight.detach().float().contiguous().cpu()
                    b = lin.bias.detach().float().contiguous().cpu()
                    _lib.check(L.specb200_camtail_add_linear(h, which, w.data_ptr(), b.data_ptr(), w.shape[0], w.shape[1]))
            _lib.check(L.specb200_camtail_finalize(h))
        self._device = device
        self._dirty = False
        self._watch = _lib.VersionWatch(*heads)

    def logits(self, images):
        """(B, 3*num_out) fp32 logits [vfov|pitch|roll]."""
        _lib.require_device(images)
        _lib.refuse_training(self)
        dev = images.device
        self._ensure(dev)
        B = images.shape[0]
        Cn = self.backbone.n_output_channels
        pooled = torch.empty(B, Cn, dtype=torch.float32, device=dev)
        self.backbone.run(images, pooled=pooled, pooled_ld=Cn)
        out = torch.empty(B, 3 * self.num_out_channels, dtype=torch.float32, device=dev)
        L = _lib.lib()
        n = L.specb200_camtail_workspace_bytes(self._handle, B)
        if self._ws is None or self._ws.numel() < n or self._ws.device != dev:
            self._ws = torch.empty(n, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.specb200_camtail_forward(self._handle, pooled.data_ptr(), Cn, B, self._ws.data_ptr(),
                                                  self._ws.numel(), out.data_ptr(),
                                                  torch.cuda.current_stream(dev).cuda_stream))
        return out

    def forward(self, images):
        lg = self.logits(images).view(images.shape[0], 3, self.num_out_channels)
        return [lg[:, 0], lg[:, 1], lg[:, 2]]

    @torch.no_grad()
    def predict_camera(self, images, img_h, img_w):
        """In-process CamCalib -> SPEC glue (replaces the subprocess + pkl of spec/tester.py:86-88 and
        spec/utils/cam_params.py:24-50): returns angles (B,3), cam_rotmat, cam_intrinsics, f_pix."""
        return decode_logits(self.logits(images), img_h, img_w)

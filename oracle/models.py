"""The two L2 wrappers of the reference, restated on top of the oracle internals.

``CameraRegressorNetwork`` follows /root/reference/camcalib/model.py:24-81 line by line;
``HMR`` follows /root/reference/spec/models/hmr.py:28-122 (the estimate_var / uncertainty
options of the ctor are accepted and ignored: they are never enabled by spec/tester.py:53-59 or
spec/trainer.py:50-56).  ``spec_full_forward`` is the in-process CamCalib -> glue -> HMR pipeline
that BASELINE.json's "SPEC full forward" configs time.
"""
import torch
import torch.nn as nn

from . import resnet as _resnet
from . import hrnet as _hrnet
from .head import HMRHead, SMPLCamHead, SMPLHead
from .geometry import convert_preds_to_angles, cam_params_from_angles

_BACKBONES = {'resnet18': _resnet.resnet18, 'resnet34': _resnet.resnet34, 'resnet50': _resnet.resnet50,
              'resnet101': _resnet.resnet101}


def get_backbone_info(backbone):
    """pare.models.backbone.utils.get_backbone_info ([UPSTREAM-RECALLED] A.1)."""
    return {'resnet18': 512, 'resnet34': 512, 'resnet50': 2048, 'resnet101': 2048,
            'hrnet_w32': 480, 'hrnet_w48': 720}[backbone]


class CameraRegressorNetwork(nn.Module):
    def __init__(self, backbone='resnet50', num_fc_layers=1, num_fc_channels=1024, num_out_channels=256):
        super().__init__()
        self.backbone = _BACKBONES[backbone]()
        self.num_out_channels = num_out_channels
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        out_channels = get_backbone_info(backbone)
        assert num_fc_layers > 0
        if num_fc_layers == 1:
            self.fc_vfov = nn.Linear(out_channels, num_out_channels)
            self.fc_pitch = nn.Linear(out_channels, num_out_channels)
            self.fc_roll = nn.Linear(out_channels, num_out_channels)
            for fc in (self.fc_vfov, self.fc_pitch, self.fc_roll):
                nn.init.normal_(fc.weight, mean=0, std=0.01)
                nn.init.constant_(fc.bias, 0)
        else:
            self.fc_vfov = self._get_fc_layers(num_fc_layers, num_fc_channels, out_channels)
            self.fc_pitch = self._get_fc_layers(num_fc_layers, num_fc_channels, out_channels)
            self.fc_roll = self._get_fc_layers(num_fc_layers, num_fc_channels, out_channels)

    def _get_fc_layers(self, num_layers, num_channels, inp_channels):
        mods = []
        for i in range(num_layers):
            if i == 0:
                mods.append(nn.Linear(inp_channels, num_channels))
            elif i == num_layers - 1:
                mods.append(nn.Linear(num_channels, self.num_out_channels))
            else:
                mods.append(nn.Linear(num_channels, num_channels))
        return nn.Sequential(*mods)

    def forward(self, images):
        x = self.backbone(images)
        x = torch.flatten(self.avgpool(x), 1)
        return [self.fc_vfov(x), self.fc_pitch(x), self.fc_roll(x)]


class HMR(nn.Module):
    def __init__(self, backbone='resnet50', focal_length=5000., img_res=224, pretrained=None, use_cam=False,
                 p=0.0, estimate_var=False, use_separate_var_branch=False, uncertainty_activation='',
                 use_cam_feats=False, smpl_data=None, mean_params=None):
        super().__init__()
        if backbone.startswith('hrnet'):
            backbone, use_conv = backbone.split('-')
            ctor = {'hrnet_w32': _hrnet.hrnet_w32, 'hrnet_w48': _hrnet.hrnet_w48}[backbone]
            self.backbone = ctor(pretrained=True, downsample=True, use_conv=(use_conv == 'conv'))
        else:
            self.backbone = _BACKBONES[backbone]()
        self.use_cam_feats = use_cam_feats
        self.head = HMRHead(num_input_features=get_backbone_info(backbone), use_cam_feats=use_cam_feats,
                            mean_params=mean_params)
        self.use_cam = use_cam
        if use_cam:
            self.smpl = SMPLCamHead(smpl_data, img_res=img_res)
        else:
            self.smpl = SMPLHead(smpl_data, focal_length=focal_length, img_res=img_res)

    def forward(self, images, cam_rotmat=None, cam_intrinsics=None, bbox_scale=None, bbox_center=None,
                img_w=None, img_h=None):
        features = self.backbone(images)
        if self.use_cam_feats:
            cam_vfov = 2 * torch.atan(img_h / (2 * cam_intrinsics[:, 0, 0]))
            hmr_output = self.head(features, cam_rotmat=cam_rotmat, cam_vfov=cam_vfov)
        else:
            hmr_output = self.head(features)
        if self.use_cam:
            smpl_output = self.smpl(rotmat=hmr_output['pred_pose'], shape=hmr_output['pred_shape'],
                                    cam=hmr_output['pred_cam'], cam_rotmat=cam_rotmat,
                                    cam_intrinsics=cam_intrinsics, bbox_scale=bbox_scale,
                                    bbox_center=bbox_center, img_w=img_w, img_h=img_h,
                                    normalize_joints2d=False)
        else:
            smpl_output = self.smpl(rotmat=hmr_output['pred_pose'], shape=hmr_output['pred_shape'],
                                    cam=hmr_output['pred_cam'], normalize_joints2d=True)
        smpl_output.update(hmr_output)
        return smpl_output


@torch.no_grad()
def spec_full_forward(camcalib, hmr, images, bbox_scale, bbox_center, img_w, img_h):
    """CamCalib(images) -> angles -> (R,K) -> HMR(images,R,K,...) (BASELINE.json configs 2,3,5)."""
    logits = camcalib(images)
    vfov, pitch, roll = convert_preds_to_angles(*logits)
    R, K, f_pix = cam_params_from_angles(vfov, pitch, roll, img_h, img_w)
    out = hmr(images, R, K, bbox_scale, bbox_center, img_w, img_h)
    out.update({'cam_vfov': vfov, 'cam_pitch': pitch, 'cam_roll': roll, 'cam_rotmat': R, 'cam_intrinsics': K})
    return out

"""Precision-matched CPU oracle for the 16-bit tensor-core modes (TEST INFRASTRUCTURE).

The bf16/fp16 CUDA path stores activations and BN-folded weights in 16 bits and accumulates in fp32.
Against the fp32 oracle that differs by ~1e-2 relative, far above the north-star
fp32 tolerances, so the 16-bit modes are checked against THIS emulation, which rounds at exactly the
points the kernels round (image, folded weights, every conv epilogue output) and accumulates in fp32
(products of 16-bit values are exact in fp32; only the summation order differs).
Walks the oracle's own ResNet modules (oracle/resnet.py) -- independent of the product's op program.
"""
import torch
import torch.nn.functional as F

from .resnet import Bottleneck, BasicBlock


def _rnd(x, dtype):
    return x.to(dtype).float()


def fold_bn(conv, bn):
    scale = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
    w = (conv.weight.double() * scale.view(-1, 1, 1, 1)).float()
    b = (bn.bias.double() - bn.running_mean.double() * scale).float()
    return w, b


def conv_bn_act(x, conv, bn, dtype, relu, res=None):
    w, b = fold_bn(conv, bn)
    y = F.conv2d(x, _rnd(w, dtype), None, conv.stride, conv.padding) + b.view(1, -1, 1, 1)
    if res is not None:
        y = y + res
    if relu:
        y = torch.relu(y)
    return _rnd(y, dtype)


@torch.no_grad()
def resnet_trunk_lowp(trunk, images, dtype=torch.bfloat16):
    x = _rnd(images, dtype)
    x = conv_bn_act(x, trunk.conv1, trunk.bn1, dtype, True)
    x = F.max_pool2d(x, 3, 2, 1)
    for layer in (trunk.layer1, trunk.layer2, trunk.layer3, trunk.layer4):
        for blk in layer:
            idt = x if blk.downsample is None else conv_bn_act(x, blk.downsample[0], blk.downsample[1], dtype, False)
            if isinstance(blk, Bottleneck):
                t = conv_bn_act(x, blk.conv1, blk.bn1, dtype, True)
                t = conv_bn_act(t, blk.conv2, blk.bn2, dtype, True)
                x = conv_bn_act(t, blk.conv3, blk.bn3, dtype, True, res=idt)
            else:
                assert isinstance(blk, BasicBlock)
                t = conv_bn_act(x, blk.conv1, blk.bn1, dtype, True)
                x = conv_bn_act(t, blk.conv2, blk.bn2, dtype, True, res=idt)
    return x


@torch.no_grad()
def camcalib_lowp(model, images, dtype=torch.bfloat16):
    f = resnet_trunk_lowp(model.backbone, images, dtype)
    x = f.mean((2, 3))
    return [model.fc_vfov(x), model.fc_pitch(x), model.fc_roll(x)]


@torch.no_grad()
def hmr_lowp(model, images, cam_rotmat, cam_intrinsics, bbox_scale, bbox_center, img_w, img_h, dtype=torch.bfloat16):
    """HMR.forward with the ResNet trunk in emulated 16-bit; head / SMPL stay fp32 exactly as on the GPU."""
    feats = resnet_trunk_lowp(model.backbone, images, dtype)
    cam_vfov = 2 * torch.atan(img_h / (2 * cam_intrinsics[:, 0, 0]))
    out = model.head(feats, cam_rotmat=cam_rotmat, cam_vfov=cam_vfov)
    so = model.smpl(rotmat=out['pred_pose'], shape=out['pred_shape'], cam=out['pred_cam'], cam_rotmat=cam_rotmat,
                    cam_intrinsics=cam_intrinsics, bbox_scale=bbox_scale, bbox_center=bbox_center, img_w=img_w,
                    img_h=img_h, normalize_joints2d=False)
    so.update(out)
    return so

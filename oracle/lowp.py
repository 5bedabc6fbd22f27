"""Precision-matched CPU oracle for the 16-bit tensor-core modes (TEST INFRASTRUCTURE).

The bf16/fp16 CUDA path stores activations and BN-folded weights in 16 bits and accumulates in fp32.
Against the fp32 oracle that differs by ~1e-2 relative, far above the north-star
fp32 tolerances, so the 16-bit modes are checked against THIS emulation, which rounds at exactly the
points the kernels round (image, folded weights, every conv epilogue output) and accumulates in fp32
(products of 16-bit values are exact in fp32; only the summation order differs).
Walks the oracle's own ResNet / HRNet modules (oracle/resnet.py, oracle/hrnet.py) -- independent of the product's op program.
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


TRACE = None      # tests may set this to a dict: conv module -> emulated output (layer-wise comparison with the CUDA path)


def conv_bn_act(x, conv, bn, dtype, relu, res=None, keep_fp32=False):
    w, b = fold_bn(conv, bn)
    y = F.conv2d(x, _rnd(w, dtype), None, conv.stride, conv.padding) + b.view(1, -1, 1, 1)
    if res is not None:
        y = y + res
    if relu:
        y = torch.relu(y)
    if not keep_fp32:
        y = _rnd(y, dtype)
    if TRACE is not None:
        TRACE[conv] = y
    return y


def _fused_downsample(blk):
    """The 64 -> 64 -> 256 stride-1 bottleneck WITH a downsample conv (ResNet-50 / HRNet layer1.0) runs as one kernel
    (spec_b200/csrc/conv_bneck.cu) in which the downsample conv and conv3 share ONE fp32 accumulator: the identity branch is
    never rounded to 16 bits on its own.  (Everything else rounds after every conv, fused or not.)"""
    return (blk.downsample is not None and isinstance(blk, Bottleneck) and blk.conv1.in_channels == 64 and blk.conv1.out_channels == 64
            and blk.conv3.out_channels == 256 and blk.conv2.stride == (1, 1) and blk.downsample[0].stride == (1, 1))


@torch.no_grad()
def resnet_trunk_lowp(trunk, images, dtype=torch.bfloat16):
    x = _rnd(images, dtype)
    x = conv_bn_act(x, trunk.conv1, trunk.bn1, dtype, True)
    x = F.max_pool2d(x, 3, 2, 1)
    for layer in (trunk.layer1, trunk.layer2, trunk.layer3, trunk.layer4):
        for blk in layer:
            if isinstance(blk, Bottleneck):
                x = _bottleneck_lowp(blk, x, dtype)
            else:
                assert isinstance(blk, BasicBlock)
                x = _basic_block_lowp(blk, x, dtype)
    return x


def _basic_block_lowp(blk, x, dtype):
    idt = x if blk.downsample is None else conv_bn_act(x, blk.downsample[0], blk.downsample[1], dtype, False)
    t = conv_bn_act(x, blk.conv1, blk.bn1, dtype, True)
    return conv_bn_act(t, blk.conv2, blk.bn2, dtype, True, res=idt)


def _bottleneck_lowp(blk, x, dtype):
    idt = x if blk.downsample is None else conv_bn_act(x, blk.downsample[0], blk.downsample[1], dtype, False,
                                                       keep_fp32=_fused_downsample(blk) and dtype != torch.float32)
    t = conv_bn_act(x, blk.conv1, blk.bn1, dtype, True)
    t = conv_bn_act(t, blk.conv2, blk.bn2, dtype, True)
    return conv_bn_act(t, blk.conv3, blk.bn3, dtype, True, res=idt)


def _hr_module_lowp(m, xs, dtype):
    """oracle/hrnet.py::HRModule.forward with a 16-bit rounding after every op the CUDA path materialises: each conv
    epilogue (the last conv of a down-path adds the running sum before it rounds), each nearest-upsample-add, ReLU fused
    into the last accumulation -- the same left-to-right order over j as the reference's ``y = y + t``."""
    xs = list(xs)
    for i in range(m.num_branches):
        for blk in m.branches[i]:
            xs[i] = _basic_block_lowp(blk, xs[i], dtype)
    if m.num_branches == 1:
        return xs
    out = []
    nb = m.num_branches
    for i in range(len(m.fuse_layers)):
        y = None
        for j in range(nb):
            last = j == nb - 1
            if j < i:
                t = xs[j]
                chain = m.fuse_layers[i][j]
                for k, step in enumerate(chain):
                    final = k == len(chain) - 1
                    t = conv_bn_act(t, step[0], step[1], dtype, relu=not final, res=y if final else None)
                y = t
            elif j == i:
                if y is None:
                    y = xs[i]
                else:
                    y = xs[i] + y
                    y = _rnd(torch.relu(y) if last else y, dtype)
            else:
                fl = m.fuse_layers[i][j]
                t = conv_bn_act(xs[j], fl[0], fl[1], dtype, False)
                y = y + F.interpolate(t, scale_factor=2 ** (j - i), mode='nearest')
                y = _rnd(torch.relu(y) if last else y, dtype)
        out.append(y)
    return out


@torch.no_grad()
def hrnet_trunk_lowp(trunk, images, dtype=torch.bfloat16):
    """oracle/hrnet.py::HRNetTrunk.forward in emulated 16-bit storage (see resnet_trunk_lowp)."""
    cba = lambda x, seq: conv_bn_act(x, seq[0], seq[1], dtype, True)
    x = _rnd(images, dtype)
    x = conv_bn_act(x, trunk.conv1, trunk.bn1, dtype, True)
    x = conv_bn_act(x, trunk.conv2, trunk.bn2, dtype, True)
    for blk in trunk.layer1:
        x = _bottleneck_lowp(blk, x, dtype)

    def transition(tr, ys, n_prev):
        res = []
        for i in range(len(tr)):
            if i < n_prev and tr[i] is None:
                res.append(ys[i])
            elif i < n_prev:
                res.append(cba(ys[i], tr[i]))
            else:
                t = ys[n_prev - 1]
                for step in tr[i]:
                    t = cba(t, step)
                res.append(t)
        return res
    xs = [cba(x, t) for t in trunk.transition1[:1]] + [cba(x, trunk.transition1[1][0])]
    xs = _hr_module_lowp(trunk.stage2[0], xs, dtype)
    xs = transition(trunk.transition2, xs, 2)
    for m in trunk.stage3:
        xs = _hr_module_lowp(m, xs, dtype)
    xs = transition(trunk.transition3, xs, 3)
    for m in trunk.stage4:
        xs = _hr_module_lowp(m, xs, dtype)
    if trunk.use_conv:
        outs = []
        for i in range(3):
            t = xs[i]
            seq = trunk.downsample_layers[i]
            for k in range(0, len(seq), 3):
                t = conv_bn_act(t, seq[k], seq[k + 1], dtype, True)
            outs.append(t)
        outs.append(xs[3])
    else:
        h, w = xs[3].shape[2:]
        outs = [_rnd(F.interpolate(xs[i], size=(h, w), mode='bilinear', align_corners=True), dtype) for i in range(3)] + [xs[3]]
    return torch.cat(outs, 1)


def trunk_lowp(trunk, images, dtype=torch.bfloat16):
    from .hrnet import HRNetTrunk
    return hrnet_trunk_lowp(trunk, images, dtype) if isinstance(trunk, HRNetTrunk) else resnet_trunk_lowp(trunk, images, dtype)


@torch.no_grad()
def camcalib_lowp(model, images, dtype=torch.bfloat16):
    f = trunk_lowp(model.backbone, images, dtype)
    x = f.mean((2, 3))
    return [model.fc_vfov(x), model.fc_pitch(x), model.fc_roll(x)]


@torch.no_grad()
def hmr_lowp(model, images, cam_rotmat, cam_intrinsics, bbox_scale, bbox_center, img_w, img_h, dtype=torch.bfloat16):
    """HMR.forward with the trunk in emulated 16-bit; head / SMPL stay fp32 exactly as on the GPU."""
    feats = trunk_lowp(model.backbone, images, dtype)
    cam_vfov = 2 * torch.atan(img_h / (2 * cam_intrinsics[:, 0, 0]))
    out = model.head(feats, cam_rotmat=cam_rotmat, cam_vfov=cam_vfov)
    so = model.smpl(rotmat=out['pred_pose'], shape=out['pred_shape'], cam=out['pred_cam'], cam_rotmat=cam_rotmat,
                    cam_intrinsics=cam_intrinsics, bbox_scale=bbox_scale, bbox_center=bbox_center, img_w=img_w,
                    img_h=img_h, normalize_joints2d=False)
    so.update(out)
    return so

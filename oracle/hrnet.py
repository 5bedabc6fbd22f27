"""HRNet-W32 / W48 pose trunk for the CPU oracle.

Restates ``pare.models.backbone.hrnet.hrnet_w32/w48(pretrained, downsample, use_conv)``
([UPSTREAM-RECALLED] SURVEY.md A.2; call site /root/reference/spec/models/hmr.py:44-51).
PARITY UNPINNED: upstream is absent, so this module *defines* the architecture the CUDA path is
checked against; the open points (tail conv bias / BN / ReLU placement) are listed in DESIGN.md.

Structure: stem 2x(3x3 s2 conv+BN+ReLU) -> layer1 (4 Bottlenecks, 64->256) -> transition1 ->
stage2 (1 module, 2 branches) -> transition2 -> stage3 (4 modules, 3 branches) -> transition3 ->
stage4 (3 modules, 4 branches, multi-scale output) -> tail (``downsample=True``):
  use_conv=True : branch i<3 through (3-i) x [3x3 s2 conv C->C + BN + ReLU], concat with branch 3
  use_conv=False: bilinear (align_corners=True) resize to branch-3 size, concat.
Output (B, sum(C), H/32, W/32): 480 channels for w32, 720 for w48.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .resnet import BasicBlock, Bottleneck


class HRModule(nn.Module):
    def __init__(self, num_branches, num_blocks, channels, multi_scale_output=True):
        super().__init__()
        self.num_branches = num_branches
        self.channels = channels
        self.multi_scale_output = multi_scale_output
        self.branches = nn.ModuleList([
            nn.Sequential(*[BasicBlock(channels[i], channels[i]) for _ in range(num_blocks)])
            for i in range(num_branches)])
        self.fuse_layers = self._make_fuse_layers()
        self.relu = nn.ReLU(inplace=True)

    def _make_fuse_layers(self):
        if self.num_branches == 1:
            return None
        C = self.channels
        fuse = []
        for i in range(self.num_branches if self.multi_scale_output else 1):
            row = []
            for j in range(self.num_branches):
                if j > i:
                    row.append(nn.Sequential(
                        nn.Conv2d(C[j], C[i], 1, 1, 0, bias=False), nn.BatchNorm2d(C[i]),
                        nn.Upsample(scale_factor=2 ** (j - i), mode='nearest')))
                elif j == i:
                    row.append(None)
                else:
                    convs = []
                    for k in range(i - j):
                        if k == i - j - 1:
                            convs.append(nn.Sequential(
                                nn.Conv2d(C[j], C[i], 3, 2, 1, bias=False), nn.BatchNorm2d(C[i])))
                        else:
                            convs.append(nn.Sequential(
                                nn.Conv2d(C[j], C[j], 3, 2, 1, bias=False), nn.BatchNorm2d(C[j]),
                                nn.ReLU(inplace=True)))
                    row.append(nn.Sequential(*convs))
            fuse.append(nn.ModuleList(row))
        return nn.ModuleList(fuse)

    def forward(self, xs):
        xs = [self.branches[i](xs[i]) for i in range(self.num_branches)]
        if self.num_branches == 1:
            return xs
        out = []
        for i in range(len(self.fuse_layers)):
            y = None
            for j in range(self.num_branches):
                t = xs[j] if j == i else self.fuse_layers[i][j](xs[j])
                y = t if y is None else y + t
            out.append(self.relu(y))
        return out


class HRNetTrunk(nn.Module):
    def __init__(self, width=32, downsample=True, use_conv=True):
        super().__init__()
        self.width, self.downsample, self.use_conv = width, downsample, use_conv
        C = [width, width * 2, width * 4, width * 8]
        self.conv1 = nn.Conv2d(3, 64, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.conv2 = nn.Conv2d(64, 64, 3, 2, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        ds = nn.Sequential(nn.Conv2d(64, 256, 1, bias=False), nn.BatchNorm2d(256))
        self.layer1 = nn.Sequential(Bottleneck(64, 64, 1, ds), Bottleneck(256, 64), Bottleneck(256, 64),
                                    Bottleneck(256, 64))
        self.transition1 = self._make_transition([256], C[:2])
        self.stage2 = nn.Sequential(HRModule(2, 4, C[:2]))
        self.transition2 = self._make_transition(C[:2], C[:3])
        self.stage3 = nn.Sequential(*[HRModule(3, 4, C[:3]) for _ in range(4)])
        self.transition3 = self._make_transition(C[:3], C[:4])
        self.stage4 = nn.Sequential(*[HRModule(4, 4, C[:4], True) for _ in range(3)])
        self.n_output_channels = sum(C)
        if downsample and use_conv:
            mods = []
            for i in range(3):
                seq = []
                for _ in range(3 - i):
                    seq += [nn.Conv2d(C[i], C[i], 3, 2, 1, bias=False), nn.BatchNorm2d(C[i]),
                            nn.ReLU(inplace=True)]
                mods.append(nn.Sequential(*seq))
            self.downsample_layers = nn.ModuleList(mods)

    @staticmethod
    def _make_transition(pre, cur):
        layers = []
        for i in range(len(cur)):
            if i < len(pre):
                if cur[i] != pre[i]:
                    layers.append(nn.Sequential(nn.Conv2d(pre[i], cur[i], 3, 1, 1, bias=False),
                                                nn.BatchNorm2d(cur[i]), nn.ReLU(inplace=True)))
                else:
                    layers.append(None)
            else:
                convs = []
                for j in range(i + 1 - len(pre)):
                    inc = pre[-1]
                    outc = cur[i] if j == i - len(pre) else inc
                    convs.append(nn.Sequential(nn.Conv2d(inc, outc, 3, 2, 1, bias=False),
                                               nn.BatchNorm2d(outc), nn.ReLU(inplace=True)))
                layers.append(nn.Sequential(*convs))
        return nn.ModuleList(layers)

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        x = self.relu(self.bn2(self.conv2(x)))
        x = self.layer1(x)
        xs = [t(x) for t in self.transition1]
        ys = self.stage2[0](xs)
        xs = [ys[i] if (i < 2 and self.transition2[i] is None) else self.transition2[i](ys[min(i, 1)])
              for i in range(3)]
        for m in self.stage3:
            xs = m(xs)
        ys = xs
        xs = [ys[i] if (i < 3 and self.transition3[i] is None) else self.transition3[i](ys[min(i, 2)])
              for i in range(4)]
        for m in self.stage4:
            xs = m(xs)
        if not self.downsample:
            return xs[0]
        if self.use_conv:
            outs = [self.downsample_layers[i](xs[i]) for i in range(3)] + [xs[3]]
        else:
            h, w = xs[3].shape[2:]
            outs = [F.interpolate(xs[i], size=(h, w), mode='bilinear', align_corners=True)
                    for i in range(3)] + [xs[3]]
        return torch.cat(outs, 1)


def hrnet_w32(pretrained=False, downsample=True, use_conv=True):
    return HRNetTrunk(32, downsample, use_conv)


def hrnet_w48(pretrained=False, downsample=True, use_conv=True):
    return HRNetTrunk(48, downsample, use_conv)

"""Test infrastructure: a CPU (torch) interpreter of the op program that spec_b200.backbone compiles for libspecb200.
Running the program with plain torch ops and comparing with the oracle trunk validates the HOST logic (architecture
compilation, buffer reuse, BN folding, concat offsets, fuse order) without a GPU."""
import torch
import torch.nn.functional as F

from spec_b200._lib import OP_CONV, OP_MAXPOOL, OP_UPADD, OP_BILINEAR, OP_COPY


@torch.no_grad()
def run_program(trunk, images):
    P = trunk._program
    mods = dict(trunk.named_modules())
    buf = {0: images}
    for o in P.ops:
        x = buf[o['src']]
        if o['type'] == OP_CONV:
            cp, bp, cout, cin, k = P.convs[o['wslot']]
            conv, bn = mods[cp], mods[bp]
            scale = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
            w = (conv.weight.double() * scale.view(-1, 1, 1, 1)).float()
            b = (bn.bias.double() - bn.running_mean.double() * scale).float()
            y = F.conv2d(x, w, b, o['stride'], o['pad'])
            if o['src2'] >= 0:
                y = y + buf[o['src2']]
            if o['relu']:
                y = torch.relu(y)
            if o['dst_coff'] == 0 and cout == P.buf_ch[o['dst']]:
                buf[o['dst']] = y
            else:
                if o['dst'] not in buf or buf[o['dst']].shape[1] != P.buf_ch[o['dst']] or buf[o['dst']].shape[2:] != y.shape[2:]:
                    buf[o['dst']] = torch.zeros(y.shape[0], P.buf_ch[o['dst']], *y.shape[2:])
                buf[o['dst']][:, o['dst_coff']:o['dst_coff'] + cout] = y
        elif o['type'] == OP_MAXPOOL:
            buf[o['dst']] = F.max_pool2d(x, 3, 2, 1)
        elif o['type'] == OP_UPADD:
            up = x if o['shift'] == 0 else F.interpolate(x, scale_factor=2 ** o['shift'], mode='nearest')
            y = buf[o['dst']] + up
            buf[o['dst']] = torch.relu(y) if o['relu'] else y
        elif o['type'] == OP_BILINEAR:
            ref = buf[o['src2']]
            y = F.interpolate(x, size=ref.shape[2:], mode='bilinear', align_corners=True)
            if o['dst'] not in buf or buf[o['dst']].shape[1] != P.buf_ch[o['dst']] or buf[o['dst']].shape[2:] != y.shape[2:]:
                buf[o['dst']] = torch.zeros(y.shape[0], P.buf_ch[o['dst']], *y.shape[2:])
            buf[o['dst']][:, o['dst_coff']:o['dst_coff'] + y.shape[1]] = y
        elif o['type'] == OP_COPY:
            if o['dst_coff'] == 0 and x.shape[1] == P.buf_ch[o['dst']]:
                buf[o['dst']] = x.clone()
            else:
                if o['dst'] not in buf or buf[o['dst']].shape[1] != P.buf_ch[o['dst']] or buf[o['dst']].shape[2:] != x.shape[2:]:
                    buf[o['dst']] = torch.zeros(x.shape[0], P.buf_ch[o['dst']], *x.shape[2:])
                buf[o['dst']][:, o['dst_coff']:o['dst_coff'] + x.shape[1]] = x
        else:
            raise ValueError(o['type'])
    return buf[trunk._out_buf]

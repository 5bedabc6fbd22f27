"""Backbone trunks (ResNet-18/34/50/101, HRNet-W32/W48) of the B200-native SPEC hot path.

Host-side mirror of ``pare.models.backbone.{resnet,hrnet}`` as the reference uses them
(/root/reference/camcalib/model.py:33,73; /root/reference/spec/models/hmr.py:44-53,92): a module
that owns parameters under the reference's state_dict names (torchvision / HRNet naming, OIHW conv
weights, BatchNorm running statistics) and whose ``forward(images)`` returns the final NCHW feature map.
No torch operator computes anything here: construction *compiles* the architecture into a flat op
program (conv+foldedBN+ReLU+residual, max-pool, upsample-add, bilinear, copy) that libspecb200 runs
with hand-written sm_100a kernels over NHWC activations.  BatchNorm is folded into the conv weights
when the weights are packed (eval-mode semantics; the hot path is inference only).
"""
import ctypes as C
import os
import warnings

import torch
import torch.nn as nn

from . import _lib
from ._lib import Op, OP_CONV, OP_MAXPOOL, OP_UPADD, OP_BILINEAR, OP_COPY

_BN_EPS = 1e-5


class _Node(nn.Module):
    """Pure parameter container (never called)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError('parameter container: computation happens in libspecb200')


def _set(root, path, module):
    parts = path.split('.')
    cur = root
    for p in parts[:-1]:
        if p not in cur._modules:
            cur.add_module(p, _Node())
        cur = cur._modules[p]
    cur.add_module(parts[-1], module)


class _Program:
    """Collects ops, activation buffers (with reuse) and the conv->parameter map."""

    def __init__(self):
        self.ops = []
        self.buf_ch = [0]            # buffer 0 = NHWC image; its channel padding depends on precision
        self.free = {}
        self.convs = []              # wslot -> (conv path, bn path, cout, cin, k)

    def new(self, ch):
        lst = self.free.get(ch)
        if lst:
            return lst.pop()
        self.buf_ch.append(ch)
        return len(self.buf_ch) - 1

    def release(self, *ids):
        for i in ids:
            if i is not None and i != 0:
                self.free.setdefault(self.buf_ch[i], []).append(i)

    def conv(self, root, src, cin, cout, k, stride, pad, conv_path, bn_path, relu, res=None, dst=None, coff=0):
        conv = nn.Conv2d(cin, cout, k, stride, pad, bias=False)
        nn.init.kaiming_normal_(conv.weight, mode='fan_out', nonlinearity='relu')
        _set(root, conv_path, conv)
        _set(root, bn_path, nn.BatchNorm2d(cout, eps=_BN_EPS))
        if dst is None:
            dst = self.new(cout)
        wslot = len(self.convs)
        self.convs.append((conv_path, bn_path, cout, cin, k))
        # 32 -> 32 channel 3x3/1 convs (HRNet's high-resolution branch) run on the pixel-pair view of the same memory
        pair = int(cin == 32 and cout == 32 and k == 3 and stride == 1 and pad == 1 and coff == 0 and src != 0)
        self.ops.append(dict(type=OP_CONV, src=src, src2=-1 if res is None else res, dst=dst, cin=cin, cout=cout,
                             kh=k, kw=k, stride=stride, pad=pad, relu=int(relu), dst_coff=coff, shift=0, wslot=wslot,
                             pair=pair))
        return dst

    def op(self, type, src, dst, src2=-1, relu=0, coff=0, shift=0):
        self.ops.append(dict(type=type, src=src, src2=src2, dst=dst, cin=0, cout=0, kh=0, kw=0, stride=0, pad=0,
                             relu=int(relu), dst_coff=coff, shift=shift, wslot=-1, pair=0))
        return dst


# ------------------------------------------------------------------------------------------ ResNet
def _basic_block(P, root, x, pre, inpl, planes, stride, ds):
    idt = P.conv(root, x, inpl, planes, 1, stride, 0, pre + 'downsample.0', pre + 'downsample.1', False) if ds else x
    t = P.conv(root, x, inpl, planes, 3, stride, 1, pre + 'conv1', pre + 'bn1', True)
    out = P.conv(root, t, planes, planes, 3, 1, 1, pre + 'conv2', pre + 'bn2', True, res=idt)
    P.release(t, x, idt if ds else None)
    return out


def _bottleneck(P, root, x, pre, inpl, planes, stride, ds):
    idt = P.conv(root, x, inpl, planes * 4, 1, stride, 0, pre + 'downsample.0', pre + 'downsample.1', False) if ds else x
    t1 = P.conv(root, x, inpl, planes, 1, 1, 0, pre + 'conv1', pre + 'bn1', True)
    t2 = P.conv(root, t1, planes, planes, 3, stride, 1, pre + 'conv2', pre + 'bn2', True)
    out = P.conv(root, t2, planes, planes * 4, 1, 1, 0, pre + 'conv3', pre + 'bn3', True, res=idt)
    P.release(t1, t2, x, idt if ds else None)
    return out


def _build_resnet(root, P, block, layers):
    exp = 4 if block is _bottleneck else 1
    x = P.conv(root, 0, 3, 64, 7, 2, 3, 'conv1', 'bn1', True)         # op cin becomes 4 (NHWC4 image) at pack time
    y = P.op(OP_MAXPOOL, x, P.new(64))
    P.release(x)
    x = y
    inpl = 64
    for li, (planes, n, stride) in enumerate(zip((64, 128, 256, 512), layers, (1, 2, 2, 2)), start=1):
        for bi in range(n):
            s = stride if bi == 0 else 1
            ds = bi == 0 and (s != 1 or inpl != planes * exp)
            x = block(P, root, x, f'layer{li}.{bi}.', inpl, planes, s, ds)
            inpl = planes * exp
    return x, inpl


# ------------------------------------------------------------------------------------------ HRNet
def _hr_module(P, root, xs, pre, C, multi_scale_output=True):
    nb = len(xs)
    for i in range(nb):
        for blk in range(4):
            xs[i] = _basic_block(P, root, xs[i], f'{pre}branches.{i}.{blk}.', C[i], C[i], 1, False)
    if nb == 1:
        return xs
    outs = []
    for i in range(nb if multi_scale_output else 1):
        acc = None
        for j in range(nb):
            last = j == nb - 1
            fp = f'{pre}fuse_layers.{i}.{j}.'
            if j < i:
                t = xs[j]
                for k in range(i - j):
                    final = k == i - j - 1
                    cout = C[i] if final else C[j]
                    t2 = P.conv(root, t, C[j], cout, 3, 2, 1, f'{fp}{k}.0', f'{fp}{k}.1', relu=not final,
                                res=acc if final else None)
                    if t != xs[j]:
                        P.release(t)
                    t = t2
                if acc is not None:
                    P.release(acc)
                acc = t
            elif j == i:
                if acc is None:
                    acc = P.op(OP_COPY, xs[i], P.new(C[i]))
                else:
                    P.op(OP_UPADD, xs[i], acc, relu=last, shift=0)
            else:
                t = P.conv(root, xs[j], C[j], C[i], 1, 1, 0, f'{fp}0', f'{fp}1', relu=False)
                P.op(OP_UPADD, t, acc, relu=last, shift=j - i)
                P.release(t)
        outs.append(acc)
    P.release(*xs)
    return outs


def _build_hrnet(root, P, width, use_conv):
    C = [width, width * 2, width * 4, width * 8]
    x = P.conv(root, 0, 3, 64, 3, 2, 1, 'conv1', 'bn1', True)
    y = P.conv(root, x, 64, 64, 3, 2, 1, 'conv2', 'bn2', True)
    P.release(x)
    x = y
    inpl = 64
    for bi in range(4):
        x = _bottleneck(P, root, x, f'layer1.{bi}.', inpl, 64, 1, bi == 0)
        inpl = 256
    # transition1
    xs = [P.conv(root, x, 256, C[0], 3, 1, 1, 'transition1.0.0', 'transition1.0.1', True),
          P.conv(root, x, 256, C[1], 3, 2, 1, 'transition1.1.0.0', 'transition1.1.0.1', True)]
    P.release(x)
    xs = _hr_module(P, root, xs, 'stage2.0.', C[:2])
    xs.append(P.conv(root, xs[1], C[1], C[2], 3, 2, 1, 'transition2.2.0.0', 'transition2.2.0.1', True))
    for m in range(4):
        xs = _hr_module(P, root, xs, f'stage3.{m}.', C[:3])
    xs.append(P.conv(root, xs[2], C[2], C[3], 3, 2, 1, 'transition3.3.0.0', 'transition3.3.0.1', True))
    for m in range(3):
        xs = _hr_module(P, root, xs, f'stage4.{m}.', C[:4])
    # tail: bring branches 0..2 to the branch-3 resolution and concatenate
    cat = P.new(sum(C))
    coff = 0
    if use_conv:
        P.op(OP_COPY, xs[3], cat, coff=sum(C[:3]))       # first writer defines the concat buffer's spatial size
        for i in range(3):
            t = xs[i]
            n = 3 - i
            for k in range(n):
                final = k == n - 1
                t2 = P.conv(root, t, C[i], C[i], 3, 2, 1, f'downsample_layers.{i}.{3 * k}', f'downsample_layers.{i}.{3 * k + 1}',
                            True, dst=cat if final else None, coff=coff if final else 0)
                if t != xs[i]:
                    P.release(t)
                t = t2
            coff += C[i]
    else:
        P.op(OP_COPY, xs[3], cat, coff=sum(C[:3]))
        for i in range(3):
            P.op(OP_BILINEAR, xs[i], cat, src2=xs[3], coff=coff)
            coff += C[i]
    return cat, sum(C)


_RESNETS = {'resnet18': (_basic_block, [2, 2, 2, 2]), 'resnet34': (_basic_block, [3, 4, 6, 3]),
            'resnet50': (_bottleneck, [3, 4, 6, 3]), 'resnet101': (_bottleneck, [3, 4, 23, 3])}
_N_OUT = {'resnet18': 512, 'resnet34': 512, 'resnet50': 2048, 'resnet101': 2048, 'hrnet_w32': 480, 'hrnet_w48': 720}


def get_backbone_info(backbone):
    """pare.models.backbone.utils.get_backbone_info (call sites model.py:37, hmr.py:58)."""
    return {'n_output_channels': _N_OUT[backbone]}


def default_precision():
    p = os.environ.get('SPECB200_PRECISION', 'bf16').lower()
    if p not in _lib.PREC:
        raise ValueError(f'SPECB200_PRECISION must be one of {list(_lib.PREC)}')
    return p


class Trunk(nn.Module):
    """A backbone: reference-named parameters + the compiled op program + the libspecb200 handle."""

    def __init__(self, arch, use_conv=True, pretrained=False, precision=None, builder=None):
        super().__init__()
        self.arch = arch
        P = _Program()
        if builder is not None:           # custom op program (unit tests of single kernels)
            out_buf, nch = builder(self, P)
        elif arch in _RESNETS:
            block, layers = _RESNETS[arch]
            out_buf, nch = _build_resnet(self, P, block, layers)
        elif arch in ('hrnet_w32', 'hrnet_w48'):
            out_buf, nch = _build_hrnet(self, P, 32 if arch == 'hrnet_w32' else 48, use_conv)
        else:
            raise ValueError(f'unknown backbone {arch}')
        self.n_output_channels = nch
        self._program = P
        self._out_buf = out_buf
        self.precision = precision or default_precision()
        self.chunk = int(os.environ.get('SPECB200_CHUNK', '0'))
        self._handle = None
        self._handle_key = None
        self._dirty = True
        self._ws = {}
        self.register_load_state_dict_post_hook(lambda m, k: m.mark_dirty())
        if pretrained:
            warnings.warn('pretrained ImageNet weights cannot be downloaded offline; backbone keeps its seeded '
                          'random initialisation until a state_dict is loaded', stacklevel=3)

    # ---- bookkeeping
    def mark_dirty(self):
        self._dirty = True

    def _apply(self, fn, *a, **k):
        self._dirty = True
        return super()._apply(fn, *a, **k)

    def set_precision(self, precision):
        if precision not in _lib.PREC:
            raise ValueError(precision)
        if precision != self.precision:
            self.precision = precision
            self._dirty = True
        return self

    def _release(self):
        if self._handle is not None:
            _lib.lib().specb200_trunk_destroy(self._handle)
            self._handle = None
        self._ws = {}

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def conv_flops_per_image(self, h=224, w=224):
        """2*MACs of all convolutions at an h x w input (the roofline numerator)."""
        shapes = {0: (h, w)}
        macs = 0
        for o in self._program.ops:
            sh = shapes[o['src']]
            if o['type'] == OP_CONV:
                ho = (sh[0] + 2 * o['pad'] - o['kh']) // o['stride'] + 1
                wo = (sh[1] + 2 * o['pad'] - o['kw']) // o['stride'] + 1
                cin = 3 if o['src'] == 0 else o['cin']
                macs += ho * wo * o['cout'] * cin * o['kh'] * o['kw']
                shapes[o['dst']] = (ho, wo)
            elif o['type'] == OP_MAXPOOL:
                shapes[o['dst']] = ((sh[0] - 1) // 2 + 1, (sh[1] - 1) // 2 + 1)
            elif o['type'] == OP_BILINEAR:
                shapes[o['dst']] = shapes[o['src2']]
            elif o['type'] == OP_COPY:
                shapes[o['dst']] = sh
        return 2 * macs

    # ---- engine
    def _weights_changed(self):
        w = getattr(self, '_watch', None)
        return w is None or w.changed()

    def _ensure(self, device):
        key = (device, self.precision)
        L = _lib.lib()
        if self._handle is not None and key == self._handle_key and not self._dirty and not self._weights_changed():
            return
        _lib.require_device()
        self._release()
        prec = _lib.PREC[self.precision]
        cpad = 4                      # image stored NHWC4 (RGB + zero channel)
        P = self._program
        buf_ch = list(P.buf_ch)
        buf_ch[0] = cpad
        ops = (Op * len(P.ops))()
        for i, o in enumerate(P.ops):
            d = dict(o)
            if d['type'] == OP_CONV and d['src'] == 0:
                d['cin'] = cpad
            for k, v in d.items():
                setattr(ops[i], k, v)
        h = C.c_void_p()
        bc = (C.c_int32 * len(buf_ch))(*buf_ch)
        with torch.cuda.device(device):
            _lib.check(L.specb200_trunk_create(C.byref(h), ops, len(P.ops), bc, len(buf_ch), len(P.convs), self._out_buf, prec))
            self._handle = h
            self._handle_key = key
            _lib.check(L.specb200_trunk_set_chunk(h, self.chunk))
            # BatchNorm is folded ON THE HOST in fp64 (one D2H copy of the whole state_dict, then CPU arithmetic): no torch
            # kernels are launched on the device for it -- the only device work of a forward is libspecb200's own kernels
            host = {k: v.detach().cpu() for k, v in self.state_dict().items()}
            for slot, (cp, bp, cout, cin, k) in enumerate(P.convs):
                w = host[cp + '.weight'].double()
                scale = host[bp + '.weight'].double() / torch.sqrt(host[bp + '.running_var'].double() + _BN_EPS)
                wf = (w * scale.view(-1, 1, 1, 1)).float().contiguous()
                bf = (host[bp + '.bias'].double() - host[bp + '.running_mean'].double() * scale).float().contiguous()
                _lib.check(L.specb200_trunk_set_conv(h, slot, wf.data_ptr(), bf.data_ptr(), cout, wf.shape[1], k, k))
        self._dirty = False
        self._watch = _lib.VersionWatch(self)

    def out_shape(self, h, w):
        c, ho, wo = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(_lib.lib().specb200_trunk_out_shape(self._handle, h, w, C.byref(c), C.byref(ho), C.byref(wo)))
        return c.value, ho.value, wo.value

    def _workspace(self, batch, h, w, device):
        key = (batch, h, w)
        ws = self._ws.get(key)
        if ws is None:
            n = _lib.lib().specb200_trunk_workspace_bytes(self._handle, batch, h, w)
            if n < 0:
                _lib.check(1)
            ws = torch.empty(n, dtype=torch.uint8, device=device)
            self._ws = {key: ws}              # keep only the most recent shape
        return ws

    def run(self, images, pooled=None, pooled_ld=0, want_features=False):
        """Enqueue the trunk on the current stream.  ``pooled``: fp32 tensor (or raw pointer) receiving the
        global-average-pooled feature per image with row stride ``pooled_ld`` floats."""
        _lib.require_device(images)
        _lib.refuse_training(self)
        if images.dim() != 4 or images.shape[1] != 3 or not images.is_floating_point():
            raise ValueError('images must be a floating-point (B,3,H,W) tensor')
        if images.shape[0] == 0:
            raise ValueError('empty batch')
        images = images.float().contiguous()
        B, _, H, W = images.shape
        self._ensure(images.device)
        ws = self._workspace(B, H, W, images.device)
        feat = None
        if want_features:
            c, ho, wo = self.out_shape(H, W)
            feat = torch.empty(B, c, ho, wo, dtype=torch.float32, device=images.device)
        pptr = pooled.data_ptr() if torch.is_tensor(pooled) else (pooled or 0)
        stream = torch.cuda.current_stream(images.device).cuda_stream
        with torch.cuda.device(images.device):
            _lib.check(_lib.lib().specb200_trunk_forward(
                self._handle, images.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(), pptr, pooled_ld,
                feat.data_ptr() if feat is not None else 0, stream))
        return feat

    def activation_after(self, images, op_index):
        """Diagnostic: the fp32 NCHW activation in the destination buffer of program op ``op_index`` right after it ran
        (``specb200_trunk_forward_until``) -- for conv ops that is conv + folded BN (+ residual) (+ ReLU) as stored."""
        _lib.require_device(images)
        images = images.float().contiguous()
        B, _, H, W = images.shape
        self._ensure(images.device)
        ws = self._workspace(B, H, W, images.device)
        c, ho, wo = C.c_int32(), C.c_int32(), C.c_int32()
        L = _lib.lib()
        stream = torch.cuda.current_stream(images.device).cuda_stream
        with torch.cuda.device(images.device):
            _lib.check(L.specb200_trunk_forward_until(self._handle, images.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(), op_index,
                                                      C.byref(c), C.byref(ho), C.byref(wo), None, stream))
            out = torch.empty(B, c.value, ho.value, wo.value, dtype=torch.float32, device=images.device)
            _lib.check(L.specb200_trunk_forward_until(self._handle, images.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(), op_index,
                                                      C.byref(c), C.byref(ho), C.byref(wo), out.data_ptr(), stream))
        return out

    def last_launches(self):
        return int(_lib.lib().specb200_trunk_last_launches(self._handle)) if self._handle else 0

    def profile_ops(self, images):
        """Per-op device milliseconds (CUDA events on the current stream; synchronises).  Returns a list of dicts
        with op geometry, algorithmic FLOPs / bytes and measured time -- the live per-layer roofline table."""
        import numpy as np
        _lib.require_device(images)
        images = images.contiguous()
        B, _, H, W = images.shape
        self._ensure(images.device)
        ws = self._workspace(B, H, W, images.device)
        pooled = torch.empty(B, self.n_output_channels, dtype=torch.float32, device=images.device)
        n = len(self._program.ops)
        ms = np.zeros(n + 2, dtype=np.float32)
        with torch.cuda.device(images.device):
            _lib.check(_lib.lib().specb200_trunk_profile(
                self._handle, images.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(), pooled.data_ptr(),
                self.n_output_channels, ms.ctypes.data, torch.cuda.current_stream(images.device).cuda_stream))
        es = 4 if self.precision == 'fp32' else 2
        shapes = {0: (H, W)}
        shapes_in = {}
        rows = [dict(name='images_to_nhwc', type=0, ms=float(ms[0]), flops=0, bytes=B * H * W * (12 + es * 4))]
        for i, o in enumerate(self._program.ops):
            sh = shapes[o['src']]
            shapes_in[i] = sh
            r = dict(type=o['type'], ms=float(ms[i + 1]), flops=0)
            if o['type'] == OP_CONV:
                ho = (sh[0] + 2 * o['pad'] - o['kh']) // o['stride'] + 1
                wo = (sh[1] + 2 * o['pad'] - o['kw']) // o['stride'] + 1
                cin = 3 if o['src'] == 0 else o['cin']
                cin_s = 4 if o['src'] == 0 else o['cin']
                r.update(name=self._program.convs[o['wslot']][0], cin=cin, cout=o['cout'], k=o['kh'], stride=o['stride'],
                         hin=sh[0], hout=ho, flops=2 * B * ho * wo * o['cout'] * cin * o['kh'] * o['kw'],
                         bytes=es * (B * sh[0] * sh[1] * cin_s + B * ho * wo * o['cout'] * (2 if o['src2'] >= 0 else 1)
                                     + o['cout'] * cin_s * o['kh'] * o['kw']))
                shapes[o['dst']] = (ho, wo)
            elif o['type'] == OP_MAXPOOL:
                shapes[o['dst']] = ((sh[0] - 1) // 2 + 1, (sh[1] - 1) // 2 + 1)
                c = self._program.buf_ch[o['src']]
                r.update(name='maxpool', bytes=es * B * c * (sh[0] * sh[1] + shapes[o['dst']][0] * shapes[o['dst']][1]))
            elif o['type'] == OP_BILINEAR:
                shapes[o['dst']] = shapes[o['src2']]
                r.update(name='bilinear', bytes=0)
            elif o['type'] == OP_COPY:
                shapes[o['dst']] = sh
                r.update(name='copy', bytes=2 * es * B * sh[0] * sh[1] * self._program.buf_ch[o['src']])
            else:
                d = shapes[o['dst']]
                r.update(name='upadd', bytes=es * B * self._program.buf_ch[o['dst']] * (2 * d[0] * d[1] + sh[0] * sh[1]))
            rows.append(r)
        # fused bottleneck groups run as ONE launch: the time is on the group's first op; its algorithmic bytes are the block's
        # input + output (+ weights), the 64-channel intermediates never reach HBM
        L = _lib.lib()
        groups = {}
        for i in range(n):
            g = int(L.specb200_trunk_fused_group_first_op(self._handle, i)) if self.precision != 'fp32' else -1
            if g >= 0:
                groups.setdefault(g, []).append(i)
        for g, members in groups.items():
            rs = [rows[1 + i] for i in members]
            first, last = rs[0], rs[-1]
            w_bytes = sum(es * r['cout'] * r['cin'] * r['k'] * r['k'] for r in rs)
            hw = B * shapes_in[members[0]][0] * shapes_in[members[0]][1]
            first_cin = rows[1 + members[0]]['cin']
            total = es * hw * (first_cin + last['cout']) + w_bytes
            for r in rs:
                r['bytes'] = 0
                r['fused_into'] = first['name']
            first['bytes'] = total
            first['name'] = first['name'] + ' [fused block x%d]' % len(members)
        rows.append(dict(name='avgpool', type=0, ms=float(ms[n + 1]), flops=0, bytes=0))
        return rows

    def forward(self, images):
        """``self.backbone(images)`` of the reference: the final NCHW fp32 feature map."""
        return self.run(images, want_features=True)

    def pooled_features(self, images):
        B = images.shape[0]
        out = torch.empty(B, self.n_output_channels, dtype=torch.float32, device=images.device)
        self.run(images, pooled=out, pooled_ld=self.n_output_channels)
        return out


def resnet18(pretrained=False, **kw):
    return Trunk('resnet18', pretrained=pretrained, **kw)


def resnet34(pretrained=False, **kw):
    return Trunk('resnet34', pretrained=pretrained, **kw)


def resnet50(pretrained=False, **kw):
    return Trunk('resnet50', pretrained=pretrained, **kw)


def resnet101(pretrained=False, **kw):
    return Trunk('resnet101', pretrained=pretrained, **kw)


def hrnet_w32(pretrained=False, downsample=True, use_conv=True, **kw):
    if not downsample:
        raise NotImplementedError('the SPEC hot path always uses downsample=True (hmr.py:47-51)')
    return Trunk('hrnet_w32', use_conv=use_conv, pretrained=pretrained, **kw)


def hrnet_w48(pretrained=False, downsample=True, use_conv=True, **kw):
    if not downsample:
        raise NotImplementedError('the SPEC hot path always uses downsample=True (hmr.py:47-51)')
    return Trunk('hrnet_w48', use_conv=use_conv, pretrained=pretrained, **kw)

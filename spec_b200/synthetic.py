"""Seeded synthetic stand-ins for the assets the reference downloads (scripts/prepare_data.sh:4-6):
the licensed SMPL model, ``smpl_mean_params.npz``, ``J_regressor_extra.npy`` and checkpoints.
Shapes and statistics follow SURVEY.md section 8(d); used by bench.py, the tests and as the loud
fallback of the module constructors when the real files are absent."""
import numpy as np
import torch

from .constants import SMPL_PARENTS


def synthetic_smpl_data(seed=0):
    g = np.random.RandomState(seed)
    V = 6890
    f32 = np.float32

    def simplex(rows, cols, sharp):
        a = g.rand(rows, cols).astype(np.float64) ** sharp
        return (a / a.sum(1, keepdims=True)).astype(f32)

    lw = g.randn(V, 24) * 2.0
    lw = np.exp(lw - lw.max(1, keepdims=True))
    return {
        'v_template': (g.randn(V, 3) * 0.3).astype(f32),
        'shapedirs': (g.randn(V, 3, 10) * 0.01).astype(f32),
        'posedirs': (g.randn(207, V * 3) * 0.001).astype(f32),
        'J_regressor': simplex(24, V, 8.0),
        'lbs_weights': (lw / lw.sum(1, keepdims=True)).astype(f32),
        'J_regressor_extra': simplex(9, V, 8.0),
        'parents': np.asarray(SMPL_PARENTS, dtype=np.int64),
    }


def synthetic_mean_params(seed=0):
    g = np.random.RandomState(seed + 1)
    pose = np.tile(np.array([1., 0, 0, 1, 0, 0], dtype=np.float32), 24) + (g.randn(144) * 0.05).astype(np.float32)
    return {'pose': pose.astype(np.float32), 'shape': np.zeros(10, np.float32), 'cam': np.array([0.9, 0., 0.], np.float32)}


@torch.no_grad()
def randomize_module_(module, seed=0):
    """Non-trivial seeded weights: Kaiming convs, BN gamma~U(.5,1) beta~N(0,.1) mean~N(0,.1) var~U(.5,1.5)
    (so BN folding is really exercised), Linear layers keep their constructor initialisation re-drawn."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.Conv2d):
            fan_out = m.weight.shape[0] * m.weight.shape[2] * m.weight.shape[3]
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan_out) ** 0.5)
        elif isinstance(m, torch.nn.BatchNorm2d):
            n = m.weight.shape[0]
            m.weight.copy_(torch.rand(n, generator=g) * 0.5 + 0.5)
            m.bias.copy_(torch.randn(n, generator=g) * 0.1)
            m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(n, generator=g) + 0.5)
    return module


def synthetic_batch(batch, seed=0, img_res=224, img_h=1080., img_w=1920., device='cpu'):
    """Images + detection boxes as SPEC sees them: N(0,1) crops, bbox_scale~U(.5,4),
    bbox_center~U(.2,.8)*(w,h), AGORA-sized full image."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, img_res, img_res, generator=g)
    scale = torch.rand(batch, generator=g) * 3.5 + 0.5
    center = (torch.rand(batch, 2, generator=g) * 0.6 + 0.2) * torch.tensor([img_w, img_h])
    out = {'images': images, 'bbox_scale': scale, 'bbox_center': center,
           'img_w': torch.full((batch,), float(img_w)), 'img_h': torch.full((batch,), float(img_h))}
    return {k: v.to(device) for k, v in out.items()}


def synthetic_camera(batch, seed=0, img_h=1080., img_w=1920.):
    """Pre-computed CamCalib parameters as the eval loop reads them from npz (cam_dataset.py:617-653):
    pitch, roll ~ U(-.6,.6), vfov ~ U(.2617, 2.1)."""
    g = torch.Generator().manual_seed(seed + 7)
    pitch = torch.rand(batch, generator=g) * 1.2 - 0.6
    roll = torch.rand(batch, generator=g) * 1.2 - 0.6
    vfov = torch.rand(batch, generator=g) * (2.1 - 0.2617) + 0.2617
    return vfov, pitch, roll


def synthetic_camera_matrices(batch, seed=0, img_h=1080., img_w=1920., device='cpu'):
    """Dataset-supplied camera of the eval loop (BASELINE configs[3]; spec/trainer.py:235-236 ``batch['cam_rotmat']``,
    ``batch['cam_int']``): R = Rx(pitch) Rz(roll), K = [[f,0,w/2],[0,f,h/2],[0,0,0]] with f = h/2/tan(vfov/2)
    (K[2,2] stays 0 as spec/utils/cam_params.py:39-46 leaves it)."""
    vfov, pitch, roll = synthetic_camera(batch, seed, img_h, img_w)
    cp, sp, cr, sr = torch.cos(pitch), torch.sin(pitch), torch.cos(roll), torch.sin(roll)
    z, o = torch.zeros_like(cp), torch.ones_like(cp)
    Rx = torch.stack([o, z, z, z, cp, -sp, z, sp, cp], 1).view(-1, 3, 3)
    Rz = torch.stack([cr, -sr, z, sr, cr, z, z, z, o], 1).view(-1, 3, 3)
    f = img_h / 2. / torch.tan(vfov / 2.)
    K = torch.zeros(batch, 3, 3)
    K[:, 0, 0] = f
    K[:, 1, 1] = f
    K[:, 0, 2] = img_w / 2.
    K[:, 1, 2] = img_h / 2.
    return (Rx @ Rz).contiguous().to(device), K.to(device)

"""Generates tests/golden/spec_resnet50_b2.npz from the CPU oracle (fp32, seeded synthetic weights and
inputs: the reference ships no golden vectors and its hot-path arithmetic -- pare / smplx -- is not
installable offline, so these are oracle-made; see oracle/__init__.py "parity unpinned").

    python -m tests.golden.make_golden
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def amplify_decoders_(hmr, gain=30.0):
    """xavier(gain=0.01) decoders make the regressed deltas ~1e-3; scale them so the head GEMMs matter."""
    with torch.no_grad():
        for m in (hmr.head.decpose, hmr.head.decshape, hmr.head.deccam):
            m.weight.mul_(gain)
    return hmr


def build_models(seed=0):
    from oracle import models as om
    from spec_b200.synthetic import synthetic_smpl_data, synthetic_mean_params, randomize_module_
    torch.manual_seed(seed)
    hmr = om.HMR('resnet50', use_cam=True, use_cam_feats=True, smpl_data=synthetic_smpl_data(0),
                 mean_params=synthetic_mean_params(0)).eval()
    randomize_module_(hmr.backbone, seed)
    amplify_decoders_(hmr)
    torch.manual_seed(seed + 1)
    cc = om.CameraRegressorNetwork('resnet50').eval()
    randomize_module_(cc.backbone, seed + 1)
    with torch.no_grad():
        for fc in (cc.fc_vfov, cc.fc_pitch, cc.fc_roll):
            fc.weight.mul_(8.0)
    return cc, hmr


def build_case(batch=2, seed=0):
    from oracle.models import spec_full_forward
    from spec_b200.synthetic import synthetic_batch
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    cc, hmr = build_models(seed)
    b = synthetic_batch(batch, seed)
    out = spec_full_forward(cc, hmr, b['images'], b['bbox_scale'], b['bbox_center'], b['img_w'], b['img_h'])
    res = {k: out[k].numpy() for k in ('cam_vfov', 'cam_pitch', 'cam_roll', 'cam_rotmat', 'cam_intrinsics', 'pred_cam',
                                       'pred_shape', 'pred_pose_6d', 'pred_pose', 'pred_cam_t', 'smpl_joints3d',
                                       'smpl_joints2d')}
    v = out['smpl_vertices'].numpy()
    res['smpl_vertices_first64'] = v[:, :64]
    res['smpl_vertices_mean'] = v.mean(1)
    res['smpl_vertices_abs_sum'] = np.abs(v.astype(np.float64)).sum((1, 2))
    return res


if __name__ == '__main__':
    out = build_case()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'spec_resnet50_b2.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items()})

"""HMRHead and SMPL / SMPLCamHead for the CPU oracle.

[UPSTREAM-RECALLED] restatement of ``pare.models.head.HMRHead``,
``pare.models.head.SMPLCamHead`` + ``pare.models.SMPL`` + ``smplx.lbs`` 0.1.28 (A.4-A.6).
Reference call sites: /root/reference/spec/models/hmr.py:57-69 (ctors), :94-113 (forward);
the buffer name ``init_pose`` is evidenced at /root/reference/scripts/spec_eval.py:57.
"""
import torch
import torch.nn as nn

from .constants import JOINT_MAP_49, SMPL_VERTEX_IDS_21, SMPL_PARENTS
from .geometry import (rot6d_to_rotmat, rotmat_to_rot6d, convert_pare_to_full_img_cam,
                       perspective_projection)


class HMRHead(nn.Module):
    def __init__(self, num_input_features, use_cam_feats=False, mean_params=None):
        super().__init__()
        npose = 24 * 6
        self.npose = npose
        self.use_cam_feats = use_cam_feats
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        n_cam = 7 if use_cam_feats else 0          # rot6d(6) + vfov(1)
        self.fc1 = nn.Linear(num_input_features + npose + 13 + n_cam, 1024)
        self.drop1 = nn.Dropout()
        self.fc2 = nn.Linear(1024, 1024)
        self.drop2 = nn.Dropout()
        self.decpose = nn.Linear(1024, npose)
        self.decshape = nn.Linear(1024, 10)
        self.deccam = nn.Linear(1024, 3)
        nn.init.xavier_uniform_(self.decpose.weight, gain=0.01)
        nn.init.xavier_uniform_(self.decshape.weight, gain=0.01)
        nn.init.xavier_uniform_(self.deccam.weight, gain=0.01)
        if mean_params is None:
            mean_params = {'pose': torch.tensor([1., 0, 0, 1, 0, 0]).repeat(24),
                           'shape': torch.zeros(10), 'cam': torch.tensor([0.9, 0., 0.])}
        self.register_buffer('init_pose', torch.as_tensor(mean_params['pose']).float().reshape(1, npose))
        self.register_buffer('init_shape', torch.as_tensor(mean_params['shape']).float().reshape(1, 10))
        self.register_buffer('init_cam', torch.as_tensor(mean_params['cam']).float().reshape(1, 3))

    def forward(self, features, cam_rotmat=None, cam_vfov=None, n_iter=3):
        B = features.shape[0]
        xf = self.avgpool(features).reshape(B, -1)
        pred_pose = self.init_pose.expand(B, -1)
        pred_shape = self.init_shape.expand(B, -1)
        pred_cam = self.init_cam.expand(B, -1)
        for _ in range(n_iter):
            if self.use_cam_feats:
                xc = torch.cat([xf, pred_pose, pred_shape, pred_cam,
                                rotmat_to_rot6d(cam_rotmat), cam_vfov.unsqueeze(-1)], 1)
            else:
                xc = torch.cat([xf, pred_pose, pred_shape, pred_cam], 1)
            xc = self.drop1(self.fc1(xc))
            xc = self.drop2(self.fc2(xc))
            pred_pose = self.decpose(xc) + pred_pose
            pred_shape = self.decshape(xc) + pred_shape
            pred_cam = self.deccam(xc) + pred_cam
        pred_rotmat = rot6d_to_rotmat(pred_pose).view(B, 24, 3, 3)
        return {'pred_pose': pred_rotmat, 'pred_cam': pred_cam, 'pred_shape': pred_shape,
                'pred_pose_6d': pred_pose}


def lbs(betas, rotmats, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    """smplx.lbs.lbs with pose2rot=False (A.4).  rotmats (B,24,3,3).  Returns verts (B,V,3), J_posed (B,24,3)."""
    B = betas.shape[0]
    v_shaped = v_template.unsqueeze(0) + torch.einsum('bl,mkl->bmk', betas, shapedirs)
    J = torch.einsum('bik,ji->bjk', v_shaped, J_regressor)
    ident = torch.eye(3, dtype=betas.dtype)
    pose_feature = (rotmats[:, 1:] - ident).reshape(B, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, posedirs).view(B, -1, 3)
    # batch_rigid_transform
    par = torch.as_tensor(parents, dtype=torch.long)
    rel_J = J.clone()
    rel_J[:, 1:] = rel_J[:, 1:] - J[:, par[1:]]
    T = torch.zeros(B, 24, 4, 4, dtype=betas.dtype)
    T[:, :, :3, :3] = rotmats
    T[:, :, :3, 3] = rel_J
    T[:, :, 3, 3] = 1
    chain = [T[:, 0]]
    for i in range(1, 24):
        chain.append(torch.matmul(chain[int(par[i])], T[:, i]))
    G = torch.stack(chain, 1)
    J_posed = G[:, :, :3, 3]
    J_h = torch.cat([J, torch.zeros(B, 24, 1, dtype=betas.dtype)], 2).unsqueeze(-1)     # (B,24,4,1)
    A = G - torch.nn.functional.pad(torch.matmul(G, J_h), [3, 0])
    Tv = torch.matmul(lbs_weights.unsqueeze(0).expand(B, -1, -1), A.view(B, 24, 16)).view(B, -1, 4, 4)
    v_h = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=betas.dtype)], 2)
    verts = torch.matmul(Tv, v_h.unsqueeze(-1))[:, :, :3, 0]
    return verts, J_posed


class SMPL49(nn.Module):
    """``pare.models.SMPL`` restated: smplx.SMPL LBS + 21 selected vertices + 9 extra-regressed
    joints -> gather by JOINT_MAP_49.  Buffers are named as in smplx so state_dicts line up."""

    def __init__(self, smpl_data):
        super().__init__()
        f = lambda k: torch.as_tensor(smpl_data[k]).float()
        self.register_buffer('v_template', f('v_template'))          # (6890,3)
        self.register_buffer('shapedirs', f('shapedirs'))            # (6890,3,10)
        self.register_buffer('posedirs', f('posedirs'))              # (207,20670)
        self.register_buffer('J_regressor', f('J_regressor'))        # (24,6890)
        self.register_buffer('lbs_weights', f('lbs_weights'))        # (6890,24)
        self.register_buffer('J_regressor_extra', f('J_regressor_extra'))   # (9,6890)
        self.register_buffer('parents', torch.as_tensor(SMPL_PARENTS, dtype=torch.long))
        self.register_buffer('joint_map', torch.as_tensor(JOINT_MAP_49, dtype=torch.long))
        self.register_buffer('vertex_ids', torch.as_tensor(SMPL_VERTEX_IDS_21, dtype=torch.long))

    def forward(self, rotmat, shape):
        verts, J_posed = lbs(shape, rotmat, self.v_template, self.shapedirs, self.posedirs,
                             self.J_regressor, self.parents.tolist(), self.lbs_weights)
        extra_v = verts[:, self.vertex_ids]
        extra_j = torch.einsum('jv,bvk->bjk', self.J_regressor_extra, verts)
        joints54 = torch.cat([J_posed, extra_v, extra_j], 1)
        return verts, joints54[:, self.joint_map]


class SMPLCamHead(nn.Module):
    def __init__(self, smpl_data, img_res=224):
        super().__init__()
        self.smpl = SMPL49(smpl_data)
        self.img_res = img_res

    def forward(self, rotmat, shape, cam, cam_rotmat, cam_intrinsics, bbox_scale, bbox_center, img_w, img_h,
                normalize_joints2d=False):
        verts, joints3d = self.smpl(rotmat, shape)
        cam_t = convert_pare_to_full_img_cam(
            pare_cam=cam, bbox_height=bbox_scale * 200., bbox_center=bbox_center, img_w=img_w, img_h=img_h,
            focal_length=cam_intrinsics[:, 0, 0], crop_res=self.img_res)
        joints2d = perspective_projection(joints3d, rotation=cam_rotmat, translation=cam_t,
                                          cam_intrinsics=cam_intrinsics)
        if normalize_joints2d:
            joints2d = joints2d / (self.img_res / 2.)
        return {'smpl_vertices': verts, 'smpl_joints3d': joints3d, 'smpl_joints2d': joints2d,
                'pred_cam_t': cam_t}


class SMPLHead(nn.Module):
    """``use_cam=False`` branch (/root/reference/spec/models/hmr.py:70-74,114-121): fixed focal
    length, weak-perspective camera -> translation, identity rotation, crop-centred projection."""

    def __init__(self, smpl_data, focal_length=5000., img_res=224):
        super().__init__()
        self.smpl = SMPL49(smpl_data)
        self.focal_length, self.img_res = focal_length, img_res

    def forward(self, rotmat, shape, cam, normalize_joints2d=False):
        verts, joints3d = self.smpl(rotmat, shape)
        B = rotmat.shape[0]
        cam_t = torch.stack([cam[:, 1], cam[:, 2], 2 * self.focal_length / (self.img_res * cam[:, 0] + 1e-9)], -1)
        K = torch.zeros(B, 3, 3)
        K[:, 0, 0] = self.focal_length
        K[:, 1, 1] = self.focal_length
        K[:, 2, 2] = 1.
        joints2d = perspective_projection(joints3d, rotation=torch.eye(3).unsqueeze(0).expand(B, -1, -1),
                                          translation=cam_t, cam_intrinsics=K)
        if normalize_joints2d:
            joints2d = joints2d / (self.img_res / 2.)
        return {'smpl_vertices': verts, 'smpl_joints3d': joints3d, 'smpl_joints2d': joints2d,
                'pred_cam_t': cam_t}

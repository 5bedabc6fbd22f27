"""CPU oracle of the eval-side metrics (TEST INFRASTRUCTURE).

``reconstruction_error`` / ``compute_similarity_transform`` / ``compute_error_verts`` live in ``pare.utils.eval_utils``
([UPSTREAM-RECALLED]: SPIN's utils/pose_utils.py; call sites /root/reference/spec/trainer.py:291-316,
/root/reference/spec/utils/compute_error.py:33-86) -- parity unpinned; the wrappers around them are followed line by line.
"""
import numpy as np
import torch

from .constants import H36M_TO_J14


def compute_similarity_transform(S1, S2):
    """Procrustes: similarity transform (s, R, t) of S1 (N,3) closest to S2 (N,3) in the least-squares sense."""
    S1, S2 = S1.T, S2.T
    mu1, mu2 = S1.mean(axis=1, keepdims=True), S2.mean(axis=1, keepdims=True)
    X1, X2 = S1 - mu1, S2 - mu2
    var1 = np.sum(X1 ** 2)
    K = X1.dot(X2.T)
    U, s, Vh = np.linalg.svd(K)
    V = Vh.T
    Z = np.eye(U.shape[0])
    Z[-1, -1] *= np.sign(np.linalg.det(U.dot(V.T)))
    R = V.dot(Z.dot(U.T))
    scale = np.trace(R.dot(K)) / var1
    t = mu2 - scale * (R.dot(mu1))
    return (scale * R.dot(S1) + t).T


def reconstruction_error(S1, S2):
    S1_hat = np.stack([compute_similarity_transform(a, b) for a, b in zip(S1, S2)])
    return np.sqrt(((S1_hat - S2) ** 2).sum(axis=-1)).mean(axis=-1)


def compute_error_verts(pred_verts, target_verts):
    return np.sqrt(((pred_verts - target_verts) ** 2).sum(axis=2)).mean(axis=1)


def trainer_metrics(pred_vertices, gt_keypoints_3d, J_regressor, gt_vertices=None):
    """spec/trainer.py:272-316 (the 14-joint part): returns mpjpe, pa_mpjpe, v2v, pred_keypoints_3d."""
    Jb = J_regressor[None].expand(pred_vertices.shape[0], -1, -1)
    pred = torch.matmul(Jb, pred_vertices)
    pelvis = pred[:, [0], :].clone()
    pred = pred[:, H36M_TO_J14, :] - pelvis
    mpjpe = torch.sqrt(((pred - gt_keypoints_3d) ** 2).sum(dim=-1)).mean(dim=-1).numpy()
    pa = reconstruction_error(pred.numpy(), gt_keypoints_3d.numpy())
    v2v = compute_error_verts(pred_vertices.numpy(), gt_vertices.numpy()) if gt_vertices is not None else None
    return mpjpe, pa, v2v, pred


def eval_single(pred_vertices, gt_vertices, J_regressor):
    """spec/utils/compute_error.py:49-86."""
    Jb = J_regressor[None].expand(pred_vertices.shape[0], -1, -1)
    pj = torch.matmul(Jb, pred_vertices); pp = pj[:, [0], :].clone(); pj = pj[:, H36M_TO_J14, :] - pp
    gj = torch.matmul(Jb, gt_vertices); gp = gj[:, [0], :].clone(); gj = gj[:, H36M_TO_J14, :] - gp
    v2v = compute_error_verts((pred_vertices - pp).numpy(), (gt_vertices - gp).numpy())
    pa = reconstruction_error(pj.numpy(), gj.numpy())
    mpjpe = torch.sqrt(((pj - gj) ** 2).sum(dim=-1)).mean(dim=-1).numpy()
    return mpjpe, pa, v2v

"""Latency of the module-level API at demo-like batch sizes (eager launches vs the graph pipeline)."""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import torch
import spec_b200 as sb
from spec_b200.synthetic import synthetic_batch, synthetic_camera, randomize_module_
from oracle import geometry as og
dev = 'cuda:0'
hmr = sb.HMR('resnet50', use_cam=True, use_cam_feats=True); randomize_module_(hmr.backbone, 0); hmr.to(dev)
cc = sb.CameraRegressorNetwork('resnet50'); randomize_module_(cc.backbone, 1); cc.to(dev)
for B in (1, 4, 16):
    b = synthetic_batch(B, 0, device=dev)
    vfov, pitch, roll = synthetic_camera(B, 0)
    R, K, _ = og.cam_params_from_angles(vfov, pitch, roll, b['img_h'].cpu(), b['img_w'].cpu())
    R, K = R.to(dev), K.to(dev)
    f = lambda: hmr(b['images'], R, K, b['bbox_scale'], b['bbox_center'], b['img_w'], b['img_h'])
    for _ in range(5): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): f()
    torch.cuda.synchronize(); t_e = (time.perf_counter() - t0) / 50
    pipe = sb.SPECPipeline(cc, hmr, use_graph=True)
    g = lambda: pipe.forward_packed(b['images'], b['bbox_scale'], b['bbox_center'], b['img_w'], b['img_h'])
    for _ in range(5): g()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): g()
    torch.cuda.synchronize(); t_g = (time.perf_counter() - t0) / 50
    print(f'B={B:3d}: HMR.forward eager {t_e*1e3:.3f} ms/call ({hmr.last_launches()} launches) | full pipeline (2 trunks) as graph {t_g*1e3:.3f} ms/call')
x = torch.randn(1, 3, 600, 800, device=dev)
for _ in range(5): cc(x)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(30): cc(x)
torch.cuda.synchronize(); print(f'CamCalib 1x600x800 eager: {(time.perf_counter()-t0)/30*1e3:.3f} ms/call')

import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import torch
import spec_b200 as sb
from spec_b200.synthetic import synthetic_batch, synthetic_camera, randomize_module_
from oracle import geometry as og
dev = 'cuda:0'
hmr = sb.HMR('resnet50', use_cam=True, use_cam_feats=True); randomize_module_(hmr.backbone, 0); hmr.to(dev)
def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def gpu_time(f, n=10):
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for B in (1, 2, 4):
    b = synthetic_batch(B, 0, device=dev)
    vfov, pitch, roll = synthetic_camera(B, 0)
    R, K, _ = og.cam_params_from_angles(vfov, pitch, roll, b['img_h'].cpu(), b['img_w'].cpu())
    R, K = R.to(dev), K.to(dev)
    t_trunk = timeit(lambda: hmr.backbone.pooled_features(b['images']))
    t_full = timeit(lambda: hmr(b['images'], R, K, b['bbox_scale'], b['bbox_center'], b['img_w'], b['img_h']))
    g_full = gpu_time(lambda: hmr(b['images'], R, K, b['bbox_scale'], b['bbox_center'], b['img_w'], b['img_h']))
    print(f'B={B}: trunk-only wall {t_trunk:.3f} ms | full forward wall {t_full:.3f} ms | full forward event-timed {g_full:.3f} ms')

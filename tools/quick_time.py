"""Throw-away first-signal timing of the full pipeline (eager + graph) on one GPU."""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import torch
import spec_b200 as sb
from spec_b200.synthetic import synthetic_batch, randomize_module_

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = 'cuda:0'
cc = sb.CameraRegressorNetwork('resnet50'); randomize_module_(cc.backbone, 1)
hmr = sb.HMR('resnet50', use_cam=True, use_cam_feats=True); randomize_module_(hmr.backbone, 0)
for m in (cc, hmr):
    m.backbone.set_precision(prec); m.backbone.chunk = chunk; m.to(dev)
b = synthetic_batch(B, 0, device=dev)
for graph in (False, True):
    pipe = sb.SPECPipeline(cc, hmr, use_graph=graph)
    args = (b['images'], b['bbox_scale'], b['bbox_center'], b['img_w'], b['img_h'])
    for _ in range(3):
        pipe.forward_packed(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        pipe.forward_packed(*args)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2 * 8.174272512e9 * B
    print(f'B={B} prec={prec} chunk={chunk} graph={graph}: {ms:.3f} ms/step  {B/ms*1e3:.0f} img/s  conv {fl/ms/1e9:.1f} TFLOP/s  launches={pipe.launches_per_step()}')
# trunk only
x = b['images']
tr = hmr.backbone
pooled = torch.empty(B, 2048, device=dev)
for _ in range(3): tr.run(x, pooled=pooled, pooled_ld=2048)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): tr.run(x, pooled=pooled, pooled_ld=2048)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f'trunk only: {ms:.3f} ms  {8.174272512e9*B/ms/1e9:.1f} TFLOP/s')

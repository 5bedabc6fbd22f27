"""Per-op live roofline table of a trunk (CUDA events through specb200_trunk_profile)."""
import sys, os, json, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import torch
import spec_b200 as sb
from spec_b200.synthetic import randomize_module_

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
arch = sys.argv[3] if len(sys.argv) > 3 else 'resnet50'
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
t = getattr(sb, arch)() if not arch.startswith('hrnet') else sb.hrnet_w32(use_conv=True)
randomize_module_(t, 0)
t.set_precision(prec); t.chunk = chunk; t.to('cuda:0')
x = torch.randn(B, 3, 224, 224, device='cuda:0')
t.profile_ops(x)
rows = t.profile_ops(x)
tot = sum(r['ms'] for r in rows)
print(f'# {arch} B={B} {prec} chunk={chunk}: sum of op times {tot:.3f} ms; conv TFLOP/s overall {sum(r["flops"] for r in rows)/tot/1e9:.1f}')
print(f'{"op":34s} {"cin":>5s} {"cout":>5s} k s {"hin":>4s} {"ms":>8s} {"TF/s":>8s} {"GB/s":>8s} {"%":>5s}')
for r in rows:
    tf = r['flops'] / r['ms'] / 1e9 if r['ms'] > 0 else 0
    gb = r.get('bytes', 0) / r['ms'] / 1e6 if r['ms'] > 0 else 0
    print(f'{r["name"]:34s} {r.get("cin",0):5d} {r.get("cout",0):5d} {r.get("k",0)} {r.get("stride",0)} {r.get("hin",0):4d} {r["ms"]:8.4f} {tf:8.1f} {gb:8.0f} {100*r["ms"]/tot:5.1f}')
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows, open(f'gpurun_out/layers_{arch}_{prec}_b{B}_c{chunk}.json', 'w'))

"""Per-tile phase timeline of the fused bottleneck kernel (conv_bneck.cu, TRACE instantiation): runs ONE eager ResNet-50 trunk
forward at B=256 with SPECB200_BNECK_TRACE set; CTA 0 of each of the three layer1 launches stamps clock64() at every phase
boundary of its first tiles.  Output: gpurun_out/bneck_trace.txt (cycles relative to each tile's MMA start)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
os.makedirs('gpurun_out', exist_ok=True)
out = 'gpurun_out/bneck_trace.txt'
if os.path.exists(out):
    os.remove(out)
os.environ['SPECB200_BNECK_TRACE'] = out
import torch
import spec_b200 as sb
from spec_b200.synthetic import randomize_module_
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
t = sb.resnet50(); randomize_module_(t, 0); t.set_precision('bf16'); t.eval(); t.to('cuda:0')
x = torch.randn(B, 3, 224, 224, device='cuda:0')
with torch.no_grad():
    t.pooled_features(x)
torch.cuda.synchronize()
print(open(out).read())

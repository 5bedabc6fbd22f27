"""Multi-GPU tests (need >= 2 GPUs on the box: ``gpurun --gpus 2``; skipped on the 1-GPU round-end box, where the
world-size-2 gloo tests of tests/test_dist_gloo.py cover the host logic): the peer-memory record gather
(specb200_allgather_outputs, copy-engine and push-kernel modes) and the NCCL baseline deliver every rank's block of exactly
the submitted step under skewed ranks and late consumers (tools/gather_check.py)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mode', ['copy', 'push', 'nccl'])
def test_record_gather_across_gpus(mode):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    n = 2 if n < 4 else (4 if n < 8 else 8)
    port = 29600 + {'copy': 1, 'push': 2, 'nccl': 3}[mode]
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
                        '--master-port', str(port), os.path.join(ROOT, 'tools', 'gather_check.py'), mode, '24', '64'],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'mismatching buffers = 0' in r.stdout

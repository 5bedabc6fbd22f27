"""Non-default kernel variants kept in the binary for A/B timing (DESIGN.md "Switches"): each runs the conv-kernel and whole-path
16-bit parity tests in a subprocess with its switch set (the switches are read once per process).  They only run when
SPECB200_RUN_EXPERIMENTAL=1 -- the round-end validation should spend its GPU time on the default path.

Round-2 status (B200): the two round-1 experiments both passed the parity subset incl. the B=256 bench-config test.
  * SPECB200_SPLIT_PRODUCER -- second TMA producer thread for the weight tiles: PROMOTED to the default; ``=0`` is the old path
                               (kept as the A/B baseline, exercised here).
  * SPECB200_MCAST_B        -- one-tile kernel in 2-CTA clusters sharing the weight tile by TMA multicast: parity-clean but slower
                               on the clock (layer2 3x3 0.081-0.093 vs 0.072-0.084 ms): REMOVED from the binary (profiles/README.md).

  * SPECB200_PDL=0          -- trunk kernels launched in plain stream order instead of with programmatic dependent launch
                               (the default since the end of round 2): the A/B baseline, exercised here.

    SPECB200_RUN_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get('SPECB200_RUN_EXPERIMENTAL') != '1', reason='set SPECB200_RUN_EXPERIMENTAL=1 to run the non-default kernel variants')
@pytest.mark.parametrize('switch,value', [('SPECB200_SPLIT_PRODUCER', '0'), ('SPECB200_NO_BNECK', '1'), ('SPECB200_PDL', '0')])
def test_non_default_variant_keeps_parity(switch, value):
    env = dict(os.environ)
    env[switch] = value
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_gpu_parity.py'), '-m', 'gpu', '-x', '-q',
                        '-k', 'conv_kernels or full_forward_lowp_parity or golden or ragged'],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]

import os
import sys
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# there is no network for the licensed SMPL files: the tests opt in to the seeded synthetic stand-ins explicitly
# (without this variable a missing asset raises FileNotFoundError, tests/test_boundary.py::test_missing_assets_raise)
os.environ.setdefault('SPECB200_SYNTHETIC_ASSETS', '1')
warnings.filterwarnings('ignore', message='.*SYNTHETIC.*')
warnings.filterwarnings('ignore', message='.*pretrained ImageNet.*')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA sm_100 (B200) device')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


@pytest.fixture(autouse=True)
def _inference_mode():
    """The product is inference-only and refuses ``training and grad-enabled`` forwards (spec_b200/_lib.py::refuse_training;
    covered by tests/test_boundary.py::test_training_mode_forward_raises): the tests run under no_grad like every caller of
    the reference's inference path (tester.py:143 ``with torch.no_grad()``)."""
    with torch.no_grad():
        yield


# ---------------------------------------------------------------- shared builders (tests only)
def make_pair(backbone='resnet50', seed=0, use_cam=True, use_cam_feats=True, amplify=True):
    """(product HMR, oracle HMR) with identical seeded non-trivial weights, via state_dict transfer."""
    import spec_b200 as sb
    from spec_b200.synthetic import synthetic_smpl_data, synthetic_mean_params, randomize_module_
    from oracle import models as om
    smpl, mean = synthetic_smpl_data(0), synthetic_mean_params(0)
    torch.manual_seed(seed)
    ref = om.HMR(backbone, use_cam=use_cam, use_cam_feats=use_cam_feats, smpl_data=smpl, mean_params=mean).eval()
    randomize_module_(ref.backbone, seed)
    from tests.golden.make_golden import amplify_decoders_
    if amplify:
        amplify_decoders_(ref)
    prod = sb.HMR(backbone, use_cam=use_cam, use_cam_feats=use_cam_feats, smpl_data=smpl, mean_params=mean).eval()
    missing = prod.load_state_dict(ref.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return prod, ref


def make_camcalib_pair(backbone='resnet50', seed=1, num_fc_layers=1):
    import spec_b200 as sb
    from spec_b200.synthetic import randomize_module_
    from oracle import models as om
    torch.manual_seed(seed)
    ref = om.CameraRegressorNetwork(backbone, num_fc_layers=num_fc_layers).eval()
    randomize_module_(ref.backbone, seed)
    # N(0, 0.01) heads on |feature| ~ 1 give nearly flat logits; scale them up so the decode is exercised
    with torch.no_grad():
        for fc in (ref.fc_vfov, ref.fc_pitch, ref.fc_roll):
            for p in fc.parameters():
                if p.dim() == 2:
                    p.mul_(8.0)
    prod = sb.CameraRegressorNetwork(backbone, num_fc_layers=num_fc_layers).eval()
    prod.load_state_dict(ref.state_dict(), strict=True)
    return prod, ref

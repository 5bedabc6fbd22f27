"""Pins the oracle's wrapper layer to OUTPUTS OF THE REFERENCE ITSELF.

tests/golden/reference_wrappers.py executes the unmodified /root/reference/camcalib/model.py, spec/models/hmr.py,
camcalib/cam_utils.py and spec/utils/cam_params.py (with ``pare`` -- un-vendored, not installable -- stubbed in sys.modules
by the oracle's restatements of its internals).  Here:

* with /root/reference mounted (build container): the live reference wrappers and oracle/models.py + oracle/geometry.py are
  run on the same seeded weights / inputs and must agree BIT FOR BIT, and the live run must reproduce the committed fixture;
* everywhere (the GPU box has no /root/reference): the oracle must reproduce tests/golden/reference_wrappers.npz, the
  committed output of that reference run, to fp32 round-off (a few ulps: the fixture may have been made on another CPU) --
  the oracle's wrapper layer is checked against the reference's own code, not against itself.  The CUDA path is compared
  with the same fixture in tests/test_gpu_parity.py::test_against_reference_wrapper_fixture.

What this does NOT pin: the internals behind the stubs (pare's backbones / HMRHead / SMPL, smplx's LBS), which remain
restated from the published algorithms (oracle/__init__.py)."""
import numpy as np
import pytest
import torch

from tests.golden import reference_wrappers as rw


@pytest.fixture(scope='module')
def fixture():
    return np.load(rw.FIXTURE)


@pytest.fixture(scope='module')
def live():
    if not rw.reference_available():
        pytest.skip('reference not mounted')
    return rw.load_reference()


def _same(name, got, want, exact=True):
    """exact: same process, same thread count, same torch CPU kernels on both sides -> bit equality.  Against the COMMITTED
    fixture (possibly made on another CPU / thread count, where the GEMM blocking and hence the fp32 summation order differ)
    the bound is a few fp32 ulps: 2e-6 + 2e-6*|ref|, three orders below the north-star tolerances."""
    assert got.shape == want.shape and got.dtype == want.dtype, (name, got.shape, want.shape, got.dtype, want.dtype)
    if got.dtype.kind in 'US':
        assert list(got) == list(want), name
    elif exact:
        assert np.array_equal(got, want), f'{name}: max abs diff {np.abs(got.astype(np.float64) - want.astype(np.float64)).max():.3e}'
    else:
        err = np.abs(got.astype(np.float64) - want.astype(np.float64))
        tol = 2e-6 + 2e-6 * np.abs(want.astype(np.float64))
        if name.endswith('joints2d') or name.endswith('cam_intrinsics'):
            tol = tol * 50                                # pixel-valued (O(1e3)) quantities
        assert (err <= tol).all(), f'{name}: max abs diff {err.max():.3e}'


@pytest.mark.parametrize('case', list(rw.CASES))
def test_oracle_reproduces_reference_fixture(fixture, case):
    """oracle == committed outputs of the reference's own wrapper code (to fp32 round-off, see _same)."""
    torch.set_num_threads(4)
    seed = 40 + list(rw.CASES).index(case)
    got = rw.run_oracle_case(rw.CASES[case], seed)
    keys = [k[len(case) + 1:] for k in fixture.files if k.startswith(case + '/')]
    assert sorted(keys) == sorted(got), (sorted(keys), sorted(got))
    for k in keys:
        _same(f'{case}/{k}', got[k], fixture[f'{case}/{k}'], exact=False)


@pytest.mark.parametrize('case', list(rw.CASES))
def test_live_reference_wrappers_equal_oracle_and_fixture(live, fixture, case, tmp_path):
    torch.set_num_threads(4)
    seed = 40 + list(rw.CASES).index(case)
    ref_out = rw.run_reference_case(live, rw.CASES[case], seed, str(tmp_path))
    ora_out = rw.run_oracle_case(rw.CASES[case], seed)
    assert sorted(ref_out) == sorted(ora_out)
    for k in ref_out:
        _same(f'live {case}/{k}', ora_out[k], ref_out[k])
        _same(f'fixture {case}/{k}', ref_out[k], fixture[f'{case}/{k}'], exact=False)


def test_reference_output_contract(fixture):
    """Key order of the reference's output dict (smpl_output first, updated with hmr_output: hmr.py:100-113) -- the product
    returns the same order (tests/test_gpu_parity.py::test_consumer_shim_contract)."""
    want = ['smpl_vertices', 'smpl_joints3d', 'smpl_joints2d', 'pred_cam_t', 'pred_pose', 'pred_cam', 'pred_shape', 'pred_pose_6d']
    for case in rw.CASES:
        assert list(fixture[f'{case}/keys']) == want
    K = fixture['spec_resnet50/cam_intrinsics']
    assert (K[:, 2, 2] == 0).all() and (K[:, 0, 0] == K[:, 1, 1]).all()        # cam_params.py:39-46 leaves K[2,2] = 0


def test_reference_bins_match_oracle_ranges(live):
    """The soft-argmax ranges the oracle hard-codes are the end points of the reference's own bin tables
    (cam_utils.py:39,55,128-133)."""
    from oracle import geometry as og
    cu = live['cam_utils']
    assert float(np.min(cu.vfov_bins)) == og.VFOV_MIN and float(np.max(cu.vfov_bins)) == og.VFOV_MAX
    assert float(np.min(cu.pitch_bins)) == og.PITCH_MIN and float(np.max(cu.pitch_bins)) == og.PITCH_MAX


def test_product_decode_formula_matches_reference_fixture(fixture):
    """Host-side check of the f_pix / K glue the product's decode kernel implements (fp32 tanf on the device): the fp32
    formula stays within the north-star camera tolerance (1e-5 relative) of the reference's float64-then-rounded value."""
    for case in ('spec_resnet50', 'spec_hrnet_w32_conv'):
        vfov = torch.from_numpy(fixture[f'{case}/cam_vfov'])
        K = torch.from_numpy(fixture[f'{case}/cam_intrinsics'])
        h = 2 * K[:, 1, 2]
        f32 = h / 2. / torch.tan(vfov / 2.)
        assert torch.allclose(f32, K[:, 0, 0], rtol=1e-5, atol=0)

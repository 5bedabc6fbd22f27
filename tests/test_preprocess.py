"""Input side: person crops and the CamCalib resize, bit-exact against cv2 / Pillow /
torchvision.  CPU part: the oracle (oracle/preprocess.py) and the library's host arithmetic against the committed
library-made vectors (tests/golden/preprocess.npz, made by tests/golden/make_preprocess_golden.py) and, when the
libraries are importable, against live calls.  GPU part: the CUDA kernels against the same vectors and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as op
from tests.golden import make_preprocess_golden as mg

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'preprocess.npz'))


# ------------------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_crops_match_cv2_golden():
    img = mg.frame(*mg.CROP_FRAME)
    for i, (box, scale) in enumerate(mg.CROP_BOXES_64):
        norm, raw, _ = op.get_single_image_crop_demo(img, box, None, scale, 64)
        assert np.array_equal(op.gen_trans_from_patch(*box, 64, 64, scale), GOLD['crop64_trans'][i]), i
        assert np.array_equal(raw, GOLD['crop64_raw'][i]), i
    assert GOLD['crop64_raw'][7].max() == 0                         # box completely outside the frame
    assert np.array_equal(GOLD['crop64_raw'][0], img[67 - 32:67 + 32, 120 - 32:120 + 32])      # identity crop
    box, scale = mg.CROP_BOX_224
    norm, raw, _ = op.get_single_image_crop_demo(img, box, None, scale, 224)
    assert np.array_equal(raw, GOLD['crop224_raw'])
    assert np.array_equal(norm, GOLD['crop224_norm'])               # ToTensor + Normalize table, float32 bit-exact


def test_oracle_resize_matches_pillow_golden():
    for i, (fr, ms) in enumerate(mg.RESIZE_CASES):
        img = mg.frame(*fr)
        oh, ow = op.resized_shape(img.shape[0], img.shape[1], ms)
        want = GOLD[f'resize{i}_raw']
        assert want.shape == (oh, ow, 3)
        assert np.array_equal(op.pil_resize_bilinear_u8(img, oh, ow), want), i
    assert np.array_equal(op.camcalib_transform(mg.frame(*mg.RESIZE_CASES[0][0]), mg.RESIZE_CASES[0][1]), GOLD['resize0_norm'])


def test_oracle_keypoints_follow_the_crop():
    box = (100.3, 60.7, 100.0, 100.0)
    kp = np.array([[100.3, 60.7, 1.0], [50.3, 10.7, 1.0], [150.3, 110.7, 0.5]])
    _, _, out = op.get_single_image_crop_demo(mg.frame(*mg.CROP_FRAME), box, kp, 1.0, 64)
    np.testing.assert_allclose(out[:, :2], [[32, 32], [0, 0], [64, 64]], atol=1e-4)
    assert np.array_equal(out[:, 2], kp[:, 2])


def test_oracle_against_live_libraries():
    cv2 = pytest.importorskip('cv2')
    pytest.importorskip('PIL')
    pytest.importorskip('torchvision')
    rng = np.random.default_rng(7)
    img = mg.frame(200, 320, 3)
    for i in range(60):
        box = (rng.uniform(-30, 350), rng.uniform(-30, 230), *(2 * [rng.uniform(10, 500)]))
        scale = float(rng.choice([1.0, 1.2]))
        cs = int(rng.choice([24, 40, 56]))
        raw, norm, trans = mg.library_crop(img, box, scale, cs)
        n2, r2, _ = op.get_single_image_crop_demo(img, box, None, scale, cs)
        assert np.array_equal(op.gen_trans_from_patch(*box, cs, cs, scale), trans)
        assert np.array_equal(r2, raw) and np.array_equal(n2, norm)
    for (h, w, ms) in [(200, 320, 90), (131, 97, 50), (48, 64, 100), (77, 77, 33)]:
        fr = mg.frame(h, w, h)
        raw, norm = mg.library_resize(fr, ms)
        assert np.array_equal(op.camcalib_transform(fr, ms), norm)


# ------------------------------------------------------------------------------------------------ CPU: host logic of the library
def test_library_host_transforms_match_cv2_golden():
    from spec_b200 import preprocess as pp
    boxes = [b for b, s in mg.CROP_BOXES_64 if s == 1.0]
    want = [GOLD['crop64_trans'][i] for i, (b, s) in enumerate(mg.CROP_BOXES_64) if s == 1.0]
    trans, inv = pp.crop_transforms(boxes, 1.0, 64)
    for i in range(len(boxes)):
        assert np.array_equal(trans[i], want[i]), i
        assert np.array_equal(inv[i].reshape(6), op.invert_affine_cv(want[i])), i
    trans, _ = pp.crop_transforms([mg.CROP_BOXES_64[3][0]], 1.2, 64)
    assert np.array_equal(trans[0], GOLD['crop64_trans'][3])
    trans, _ = pp.crop_transforms([mg.CROP_BOX_224[0]], 1.0, 224)
    assert np.array_equal(trans[0], GOLD['crop224_trans'])
    rng = np.random.default_rng(0)
    boxes = np.stack([rng.uniform(-50, 2000, 500), rng.uniform(-50, 1100, 500), rng.uniform(5, 900, 500), rng.uniform(5, 900, 500)], 1)
    trans, inv = pp.crop_transforms(boxes, 1.2, 224)
    for i in range(500):
        m = op.gen_trans_from_patch(*boxes[i], 224, 224, 1.2)
        assert np.array_equal(trans[i], m) and np.array_equal(inv[i].reshape(6), op.invert_affine_cv(m))


def test_library_resized_shape():
    from spec_b200 import preprocess as pp
    for (h, w, ms) in [(1080, 1920, 600), (1920, 1080, 600), (600, 600, 600), (333, 500, 600), (7, 1000, 5), (135, 240, 75)]:
        assert pp.resized_shape(h, w, ms) == op.resized_shape(h, w, ms)
    assert pp.resized_shape(1080, 1920, 600) == (600, 1066)


def test_preprocess_has_no_cpu_path():
    from spec_b200 import preprocess as pp
    with pytest.raises(RuntimeError, match='no CPU path'):
        pp.Preprocessor().crop(torch.zeros(8, 8, 3, dtype=torch.uint8), [(4, 4, 8, 8)])
    with pytest.raises(RuntimeError, match='no CPU path'):
        pp.camcalib_transform(torch.zeros(8, 8, 3, dtype=torch.uint8))


# ------------------------------------------------------------------------------------------------ GPU: kernels
@pytest.mark.gpu
def test_gpu_crops_match_cv2_golden():
    from spec_b200 import preprocess as pp
    img = torch.from_numpy(mg.frame(*mg.CROP_FRAME)).cuda()
    P = pp.Preprocessor()
    for scale in (1.0, 1.2):
        idx = [i for i, (b, s) in enumerate(mg.CROP_BOXES_64) if s == scale]
        norm, raw = P.crop(img, [mg.CROP_BOXES_64[i][0] for i in idx], scale=scale, crop_size=64, return_raw=True)
        assert np.array_equal(raw.cpu().numpy(), GOLD['crop64_raw'][idx])
        assert np.array_equal(norm.cpu().numpy(), np.stack([op.to_tensor_normalize(GOLD['crop64_raw'][i]) for i in idx]))
    norm, raw, _ = pp.get_single_image_crop_demo(img, mg.CROP_BOX_224[0], None, scale=1.0, crop_size=224)
    assert np.array_equal(raw.cpu().numpy(), GOLD['crop224_raw'])
    assert np.array_equal(norm.cpu().numpy(), GOLD['crop224_norm'])


@pytest.mark.gpu
def test_gpu_resize_matches_pillow_golden():
    from spec_b200 import preprocess as pp
    P = pp.Preprocessor()
    for i, (fr, ms) in enumerate(mg.RESIZE_CASES):
        img = torch.from_numpy(mg.frame(*fr)).cuda()
        out, raw = P.resize(img, min_size=ms, return_raw=True)
        assert np.array_equal(raw.cpu().numpy(), GOLD[f'resize{i}_raw']), i
        if i == 0:
            assert np.array_equal(out[0].cpu().numpy(), GOLD['resize0_norm'])
            assert np.array_equal(pp.camcalib_transform(img, ms).cpu().numpy(), GOLD['resize0_norm'])


@pytest.mark.gpu
def test_gpu_full_hd_frame_against_oracle():
    """1080p frame, 40 detections (two launches of 32), BGR input with a padded row pitch; bit-exact vs the oracle."""
    from spec_b200 import preprocess as pp
    rng = np.random.default_rng(5)
    rgb = mg.frame(1080, 1920, 9)
    padded = torch.zeros(1080, 1952, 3, dtype=torch.uint8, device='cuda')
    padded[:, :1920] = torch.from_numpy(rgb[:, :, ::-1].copy()).cuda()
    bgr_view = padded[:, :1920]                                       # row pitch 1952*3 bytes, channels B,G,R
    boxes = np.stack([rng.uniform(-50, 1970, 40), rng.uniform(-50, 1130, 40), rng.uniform(40, 900, 40)], 1)
    boxes = np.concatenate([boxes, boxes[:, 2:3]], 1)
    boxes[0] = (960.0, 540.0, 224.0, 224.0)
    P = pp.Preprocessor()
    norm, raw = P.crop(bgr_view, boxes, scale=1.0, crop_size=224, bgr=True, return_raw=True)
    raw, norm = raw.cpu().numpy(), norm.cpu().numpy()
    assert np.array_equal(raw[0], rgb[540 - 112:540 + 112, 960 - 112:960 + 112])         # identity crop == slice
    for i in range(40):
        n2, r2, _ = op.get_single_image_crop_demo(rgb, boxes[i], None, 1.0, 224)
        assert np.array_equal(raw[i], r2), i
        assert np.array_equal(norm[i], n2), i
    out, rraw = P.resize(torch.from_numpy(rgb).cuda(), min_size=600, return_raw=True)
    assert tuple(out.shape) == (1, 3, 600, 1066)
    want = op.pil_resize_bilinear_u8(rgb, 600, 1066)
    assert np.array_equal(rraw.cpu().numpy(), want)
    assert np.array_equal(out[0].cpu().numpy(), op.to_tensor_normalize(want))


@pytest.mark.gpu
def test_gpu_preprocess_properties():
    from spec_b200 import preprocess as pp
    P = pp.Preprocessor()
    lut = op.normalize_lut()
    const = torch.full((300, 400, 3), 77, dtype=torch.uint8, device='cuda')
    out = P.resize(const, min_size=120)
    for c in range(3):
        assert torch.all(out[0, c] == float(lut[c, 77]))              # a constant frame stays constant
    same = torch.from_numpy(mg.frame(96, 96, 1)).cuda()
    out, raw = P.resize(same, min_size=96, return_raw=True)           # same size: Pillow copies, so do we
    assert torch.equal(raw, same)
    norm = P.crop(const, [(200, 150, 100, 100)], crop_size=32)       # box inside a constant frame
    for c in range(3):
        assert torch.all(norm[0, c] == float(lut[c, 77]))
    assert P.crop(const, np.zeros((0, 4)), crop_size=32).shape == (0, 3, 32, 32)      # frame without detections
    with pytest.raises(RuntimeError, match='positive'):
        P.crop(const, [(10, 10, 0, 5)])
    kp = np.array([[200.0, 150.0, 1.0]])
    _, _, k2 = pp.get_single_image_crop_demo(const, (200, 150, 100, 100), kp, scale=1.0, crop_size=32)
    np.testing.assert_allclose(k2[0, :2], [16, 16], atol=1e-5)


@pytest.mark.gpu
def test_gpu_demo_loop_on_a_frame_matches_oracle():
    """SPECPipeline.run_on_frame == the demo loop of tester.py:99-167 restated with the oracle: Resize -> CamCalib ->
    (R, K) for the frame; crop every detection -> HMR.  fp32 mode, north-star tolerances."""
    import spec_b200 as sb
    from oracle import geometry as og
    from tests.conftest import make_pair, make_camcalib_pair
    hmr, hmr_ref = make_pair('resnet50', seed=0)
    cc, cc_ref = make_camcalib_pair('resnet50', seed=1)
    cc.backbone.set_precision('fp32')
    hmr.backbone.set_precision('fp32')
    pipe = sb.SPECPipeline(cc.to('cuda:0'), hmr.to('cuda:0'), use_graph=False)
    rgb = mg.frame(270, 480, 4)
    dets = np.array([[240.0, 135.0, 180.0, 180.0], [60.5, 200.25, 150.0, 150.0], [400.0, 80.0, 260.0, 260.0]])
    got = pipe.run_on_frame(torch.from_numpy(rgb[:, :, ::-1].copy()).cuda(), dets, bgr=True, camcalib_min_size=160)
    torch.cuda.synchronize()
    assert pipe.run_on_frame(torch.from_numpy(rgb).cuda(), np.zeros((0, 4))) == {}
    # oracle demo loop
    with torch.no_grad():
        full = torch.from_numpy(op.camcalib_transform(rgb, 160))[None]
        vfov, pitch, roll = og.convert_preds_to_angles(*cc_ref(full))
        R, K, _ = og.cam_params_from_angles(vfov, pitch, roll, 270, 480)
        crops = torch.from_numpy(np.stack([op.get_single_image_crop_demo(rgb, d, None, 1.0, 224)[0] for d in dets]))
        n = len(dets)
        ref = hmr_ref(crops, R.expand(n, 3, 3), K.expand(n, 3, 3), torch.tensor(dets[:, 2] / 200.0, dtype=torch.float32),
                      torch.tensor(dets[:, :2], dtype=torch.float32), torch.full((n,), 480.0), torch.full((n,), 270.0))
    assert torch.equal(got['inp_images'].cpu(), crops)                                   # byte work: bit-exact
    def close(name, a, b, atol, rtol=0.0):
        err = (a.cpu() - b).abs()
        assert bool((err <= atol + rtol * b.abs()).all()), (name, float(err.max()))
    close('cam angles', got['cam_angles'][0], torch.stack([vfov[0], pitch[0], roll[0]]), 1e-5)
    close('smpl_vertices', got['smpl_vertices'], ref['smpl_vertices'], 1e-3)
    close('smpl_joints3d', got['smpl_joints3d'], ref['smpl_joints3d'], 1e-3)
    close('pred_cam', got['pred_cam'], ref['pred_cam'], 1e-5, 1e-5)
    close('smpl_joints2d', got['smpl_joints2d'], ref['smpl_joints2d'], 0.05, 1e-4)

"""CPU oracle for the SPEC inference hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, in plain fp32 PyTorch on the CPU, the arithmetic of the
reference path CamCalib -> backbone -> HMR head -> SMPL/LBS -> projection
(SURVEY.md section 8a, Appendix A).  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
it.  The product (``spec_b200``) never does, and fails loudly without its CUDA
library.

PARITY STATUS: *parity unpinned* for everything except the ResNet trunk and the
INPUT SIDE (oracle/preprocess.py: person crops and the CamCalib resize are pinned
bit-exactly against cv2 / Pillow / torchvision, which ARE installed here).
The reference keeps the arithmetic of this path in two un-vendored third-party
packages -- ``pare`` (git+https://github.com/mkocabas/PARE.git, NO commit
pinned, /root/reference/requirements.txt:28) and ``smplx==0.1.28``
(requirements.txt:7) -- neither of which is present offline, and it ships no
tests, golden vectors or fixtures.  What IS pinned:

* the wrappers that live in /root/reference are followed line by line
  (camcalib/model.py:24-81, spec/models/hmr.py:28-122,
  camcalib/cam_utils.py:39-63,110-145, spec/utils/cam_params.py:24-50,
  spec/constants.py:20-113);
* the ResNet trunk is checked layer-for-layer against torchvision's own
  ``resnet50``/``resnet34`` (pare's trunk is a copy of torchvision's with
  avgpool/fc removed) in tests/test_oracle.py;
* domain invariants (rotation orthonormality, identity-pose LBS, optical-axis
  projection, soft-argmax of uniform logits) and the committed golden vectors
  under tests/golden/ (made by tests/golden/make_golden.py from this oracle).
"""

"""Integer tables and image statistics of the hot path (host-side mirror of
/root/reference/spec/constants.py:20-21,29-113).  JOINT_MAP_49 is the resolved gather table
``[JOINT_MAP[n] for n in JOINT_NAMES]`` (49 indices into the 54 candidate joints = 24 SMPL joints +
21 selected vertices + 9 extra-regressed joints); it is uploaded verbatim to the GPU (bit-exact)."""

IMG_NORM_MEAN = [0.485, 0.456, 0.406]
IMG_NORM_STD = [0.229, 0.224, 0.225]

JOINT_MAP_49 = [24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                8, 5, 45, 46, 4, 7, 21, 19, 17, 16, 18, 20, 47, 48, 49, 50, 51, 52, 53, 24, 26, 25, 28, 27]

# smplx VertexJointSelector vertex ids appended after the 24 SMPL joints (candidates 24..44)
SMPL_VERTEX_IDS_21 = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                      2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

H36M_TO_J17 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9]
H36M_TO_J14 = H36M_TO_J17[:14]

SMPL_NUM_VERTS = 6890
# per-image packed output record (floats): the multi-GPU all-gather unit
RECORD_LAYOUT = (('smpl_vertices', 20670, (6890, 3)), ('smpl_joints3d', 147, (49, 3)), ('smpl_joints2d', 98, (49, 2)),
                 ('pred_cam_t', 3, (3,)), ('pred_cam', 3, (3,)), ('pred_shape', 10, (10,)), ('pred_pose', 216, (24, 3, 3)),
                 ('pred_pose_6d', 144, (144,)), ('cam_angles', 3, (3,)))
RECORD_FLOATS = sum(n for _, n, _ in RECORD_LAYOUT)          # 21294 floats = 85,176 B per image

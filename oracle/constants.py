"""Joint tables of the 49-joint layout.

Restates /root/reference/spec/constants.py:29-105 (JOINT_NAMES / JOINT_MAP) as
the resolved integer gather table ``JOINT_MAP_49`` (= ``[JOINT_MAP[n] for n in
JOINT_NAMES]``), plus the smplx 0.1.28 ``VertexJointSelector`` vertex ids
([UPSTREAM-RECALLED]) and H36M selectors (constants.py:109-113).
tests/test_oracle.py re-derives JOINT_MAP_49 from the name tables below.
"""

IMG_NORM_MEAN = [0.485, 0.456, 0.406]   # constants.py:20
IMG_NORM_STD = [0.229, 0.224, 0.225]    # constants.py:21

JOINT_NAMES = [
    'OP Nose', 'OP Neck', 'OP RShoulder', 'OP RElbow', 'OP RWrist', 'OP LShoulder', 'OP LElbow',
    'OP LWrist', 'OP MidHip', 'OP RHip', 'OP RKnee', 'OP RAnkle', 'OP LHip', 'OP LKnee', 'OP LAnkle',
    'OP REye', 'OP LEye', 'OP REar', 'OP LEar', 'OP LBigToe', 'OP LSmallToe', 'OP LHeel',
    'OP RBigToe', 'OP RSmallToe', 'OP RHeel',
    'Right Ankle', 'Right Knee', 'Right Hip', 'Left Hip', 'Left Knee', 'Left Ankle', 'Right Wrist',
    'Right Elbow', 'Right Shoulder', 'Left Shoulder', 'Left Elbow', 'Left Wrist', 'Neck (LSP)',
    'Top of Head (LSP)', 'Pelvis (MPII)', 'Thorax (MPII)', 'Spine (H36M)', 'Jaw (H36M)',
    'Head (H36M)', 'Nose', 'Left Eye', 'Right Eye', 'Left Ear', 'Right Ear',
]

JOINT_MAP = {
    'OP Nose': 24, 'OP Neck': 12, 'OP RShoulder': 17, 'OP RElbow': 19, 'OP RWrist': 21,
    'OP LShoulder': 16, 'OP LElbow': 18, 'OP LWrist': 20, 'OP MidHip': 0, 'OP RHip': 2,
    'OP RKnee': 5, 'OP RAnkle': 8, 'OP LHip': 1, 'OP LKnee': 4, 'OP LAnkle': 7, 'OP REye': 25,
    'OP LEye': 26, 'OP REar': 27, 'OP LEar': 28, 'OP LBigToe': 29, 'OP LSmallToe': 30,
    'OP LHeel': 31, 'OP RBigToe': 32, 'OP RSmallToe': 33, 'OP RHeel': 34, 'Right Ankle': 8,
    'Right Knee': 5, 'Right Hip': 45, 'Left Hip': 46, 'Left Knee': 4, 'Left Ankle': 7,
    'Right Wrist': 21, 'Right Elbow': 19, 'Right Shoulder': 17, 'Left Shoulder': 16,
    'Left Elbow': 18, 'Left Wrist': 20, 'Neck (LSP)': 47, 'Top of Head (LSP)': 48,
    'Pelvis (MPII)': 49, 'Thorax (MPII)': 50, 'Spine (H36M)': 51, 'Jaw (H36M)': 52,
    'Head (H36M)': 53, 'Nose': 24, 'Left Eye': 26, 'Right Eye': 25, 'Left Ear': 28, 'Right Ear': 27,
}

JOINT_MAP_49 = [JOINT_MAP[n] for n in JOINT_NAMES]

# smplx VertexJointSelector ('smplh' ids): face(5) + feet(6) + finger tips(10)
SMPL_VERTEX_IDS_21 = [
    332, 6260, 2800, 4071, 583,                 # nose, reye, leye, rear, lear
    3216, 3226, 3387, 6617, 6624, 6787,         # LBigToe LSmallToe LHeel RBigToe RSmallToe RHeel
    2746, 2319, 2445, 2556, 2673,               # l thumb,index,middle,ring,pinky
    6191, 5782, 5905, 6016, 6133,               # r thumb,index,middle,ring,pinky
]

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]

H36M_TO_J17 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9]   # constants.py:109
H36M_TO_J14 = H36M_TO_J17[:14]

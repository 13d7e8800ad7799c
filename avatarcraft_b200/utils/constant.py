"""Constants of the hot path (values from the reference's utils/constant.py:5-42)."""
DEFAULT_GEO_THRESH = 0.05     # mesh-guided near/far radius AND squared-distance mask threshold
NSR_BOUND = 1.6               # half edge of the scene cube
GLOBAL_SEED = 42
WHITE_BKG, BLACK_BKG, NOISE_BKG, CHESSBOARD_BKG = 0, 1, 2, 3
SMPL_SCALE = 0.9
CANONICAL_CAMERA_DIST_TRAIN = 2.0 * SMPL_SCALE
CANONICAL_CAMERA_DIST_VAL = 1.6 * SMPL_SCALE
CAN_HEAD_CAMERA_DIST = 0.5 * SMPL_SCALE
CAN_HEAD_OFFSET = 0.47 * SMPL_SCALE
CANONICAL_ZOOM_FACTOR = 1000 / 1280

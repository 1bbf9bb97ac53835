"""Configuration surface of the network builder.

Mirrors the constant NAMES and default VALUES of the reference's ``yolo/config.py`` (lines cited per
group) so that code written against ``import yolo.config as cfg`` keeps working; like the reference,
the values are plain module attributes that callers may overwrite before constructing ``YOLONet``
(e.g. ``cfg.BATCH_SIZE = 1``, calculate_test_map.py:354).
"""
import os

import numpy as np

# --- paths (config.py:12-16); only used by the reference's drivers ---------------------------
MODEL_PATH = os.environ.get('DISYOLO_MODEL_PATH', os.getcwd())
DATASET = os.path.join(MODEL_PATH, 'data')
OUTPUT_DIR = os.path.join(MODEL_PATH, 'output')
WEIGHTS_FILE = os.path.join(MODEL_PATH, 'pretrained_weights', 'yolov3_3class_coco.ckpt')

# --- device (config.py:18) -------------------------------------------------------------------
GPU = '0'

# --- classes and anchors from dimension clustering at 576 px (config.py:21-22) ----------------
CLASSES = ['crack', 'spall', 'rebar']
ANCHORS = np.array([[31, 23], [62, 58], [143, 91],
                    [213, 186], [61, 337], [194, 432],
                    [474, 248], [551, 93], [478, 454]], dtype=np.float32)

# --- augmentation switches (config.py:25-26); host data pipeline, unused by the hot path ------
FLIPPED = True
BLUR_NOISE_LIGHT = True

# --- schedule (config.py:31-35) ----------------------------------------------------------------
MAX_ITER = 10000
SUMMARY_ITER = 50
SAVE_ITER = 500

# --- network (config.py:38-46) -----------------------------------------------------------------
ALPHA = 0.1                        # leaky-ReLU slope
BATCH_SIZE = 2                     # baked into the reference graph; here: the net's max batch
IMAGE_SIZE = 576
K_MAP = 3                          # k x k position-sensitive score maps
BASE_GRID = int(IMAGE_SIZE / 32)   # cells of the stride-32 map

# --- loss scales (config.py:49-57) -------------------------------------------------------------
OBJECT_SCALE = 2.0
NOOBJECT_SCALE = 1.0
CLASS_SCALE = 1.0
COORD_SCALE = 1.0
MASK_SCALE = 5.0
SCORE_SCALE = 2.0
IGNORE_THRESH = 0.5

# --- detection thresholds (config.py:60-72) ----------------------------------------------------
OBJ_THRESHOLD = 0.25
IOU_THRESHOLD = 0.3
TEST_SIZE = 576
MAX_BOX_PER_IMAGE = 20
MAX_DETECTION = 30

"""Behavioural constants of the hot path (oracle side; test infrastructure).

Values restated from the reference's ``src/config.py`` and tracker core:
  YOLO_INPUT_SHAPE        src/config.py:16
  YOLO_CONF_THRESHOLD     src/config.py:17
  YOLO_NMS_THRESHOLD      src/config.py:18   (stored by the detector, never used by it)
  DEEPSORT_*              src/config.py:23-29
  REID_INPUT_SHAPE        src/config.py:32
  CLASSES                 src/config.py:36-48
  CLASSES_TO_TRACK        src/config.py:53
  INFTY_COST              src/tracker/core/linear_assignment.py:9
  CHI2INV95[4]            src/tracker/core/kalman_filter.py:16
"""

YOLO_INPUT_SHAPE = (640, 640)
YOLO_CONF_THRESHOLD = 0.3
YOLO_NMS_THRESHOLD = 0.5
YOLO_TOPK = 100  # upstream end-to-end export default; unpinned by the reference

DEEPSORT_MAX_DIST = 0.2
DEEPSORT_MIN_CONFIDENCE = 0.3
DEEPSORT_MAX_IOU_DISTANCE = 0.7
DEEPSORT_MAX_AGE = 70
DEEPSORT_N_INIT = 3
DEEPSORT_NN_BUDGET = 100

REID_INPUT_SHAPE = (128, 64)
REID_FEATURE_DIM = 512

CLASSES = (
    'person', 'bicycle', 'car', 'motorcycle', 'airplane', 'bus', 'train', 'truck', 'boat',
    'traffic light', 'fire hydrant', 'stop sign', 'parking meter', 'bench', 'bird', 'cat',
    'dog', 'horse', 'sheep', 'cow', 'elephant', 'bear', 'zebra', 'giraffe', 'backpack',
    'umbrella', 'handbag', 'tie', 'suitcase', 'frisbee', 'skis', 'snowboard', 'sports ball',
    'kite', 'baseball bat', 'baseball glove', 'skateboard', 'surfboard', 'tennis racket',
    'bottle', 'wine glass', 'cup', 'fork', 'knife', 'spoon', 'bowl', 'banana', 'apple',
    'sandwich', 'orange', 'broccoli', 'carrot', 'hot dog', 'pizza', 'donut', 'cake', 'chair',
    'couch', 'potted plant', 'bed', 'dining table', 'toilet', 'tv', 'laptop', 'mouse',
    'remote', 'keyboard', 'cell phone', 'microwave', 'oven', 'toaster', 'sink',
    'refrigerator', 'book', 'clock', 'vase', 'scissors', 'teddy bear', 'hair drier',
    'toothbrush'
)
CLASSES_TO_TRACK = {'person', 'car', 'bus', 'truck', 'motorcycle'}
TRACKED_CLASS_IDS = tuple(i for i, n in enumerate(CLASSES) if n in CLASSES_TO_TRACK)  # (0, 2, 3, 5, 7)

INFTY_COST = 1e5
CHI2INV95_4 = 9.487729036781154

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
LETTERBOX_PAD_VALUE = 114

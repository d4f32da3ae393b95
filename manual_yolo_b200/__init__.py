"""manual_yolo_b200 -- B200-native (sm_100a) detection post-processing path of kanaksharma67/manual-yolo.

letterbox -> YOLOv8 Detect-head decode + confidence filter -> sort -> class-aware NMS -> ROI crop/resize,
hand-written CUDA behind the C ABI of ``include/b200yolo.h`` (``libb200yolo.so``).  See DESIGN.md.
"""

from . import classifier, geometry, handoff, tracking  # noqa: F401
from .classifier import RankClassifier, load_rank_classifier  # noqa: F401
from .api import (Candidates, DenseChain, Detections, Workspace, classify_preprocess, crop_resize_rois, decode_and_filter, filter_decoded,  # noqa: F401
                  gather_slice_detections, greedy_nmm, iou_cost_matrix, preprocess_slices,
                  letterbox, nms_candidates, nms_sorted, non_max_suppression, postprocess_dense, postprocess_small,
                  preprocess,
                  rois_from_detections, scale_boxes, scale_params_tensor,
                  select_rois, sort_candidates, stage_head_classes_h2d, stage_rows_h2d)
from .pipeline import BatchStream, HostRunner, Pipeline, PipelineResult, SlicedPipeline  # noqa: F401

__version__ = "0.1.0"

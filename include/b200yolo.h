/*
 * b200yolo.h -- C ABI of libb200yolo.so: the B200 (sm_100a) detection post-processing path.
 *
 * The reference (kanaksharma67/manual-yolo) has no FFI of its own: its scripts reach this path
 * through the Ultralytics Python API (detect.py:20-21,541; yolo.py:354,361; pipe.py:147,179;
 * classifier detect.py:121).  The entry points below are what a binding for that seam needs:
 * one per replaceable upstream function (ultralytics==8.3.176, requirements.txt:95).  The Python
 * host side (manual_yolo_b200/api.py) binds them with ctypes; INTEGRATION.md shows the
 * reference-side stub.
 *
 * Conventions
 *  - every pointer is a DEVICE-ACCESSIBLE pointer owned by the caller (e.g. torch tensor.data_ptr());
 *    `frames` of the K5 entry points may also be PINNED HOST memory (cudaHostAlloc / torch
 *    pin_memory(): same pointer on the device under unified addressing) -- the crops are then
 *    read over PCIe without staging the frames (the `levels` of b200yolo_postprocess_small
 *    likewise); b200yolo_stage_rows_h2d / b200yolo_copy2d_h2d are the only calls that take a host
 *    source by contract;
 *    the library never allocates, frees or keeps caller memory, and has no mutable global state;
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it;
 *  - return 0 = OK, < 0 = argument error detected on the host before any launch,
 *    > 0 = cudaError_t observed after launch.  b200yolo_strerror() names both kinds;
 *  - no exceptions, no abort, no CPU fallback.
 */
#ifndef B200YOLO_H_
#define B200YOLO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200YOLO_VERSION 100 /* 0.1.0 */

enum {
  B200YOLO_OK = 0,
  B200YOLO_ERR_NULL = -1,        /* required pointer is NULL */
  B200YOLO_ERR_SHAPE = -2,       /* non-positive / inconsistent dimension */
  B200YOLO_ERR_ALIGN = -3,       /* pointer or pitch not aligned as documented */
  B200YOLO_ERR_UNSUPPORTED = -4, /* outside the implemented envelope (limits below) */
  B200YOLO_ERR_WORKSPACE = -5,   /* workspace too small: see b200yolo_workspace_bytes */
  B200YOLO_ERR_RANGE = -6        /* threshold outside [0,1] etc. (UL asserts the same) */
};

#define B200YOLO_REG_MAX 16      /* DFL bins per box side (ultralytics nn/modules/head.py Detect.reg_max) */
#define B200YOLO_MAX_LEVELS 3
#define B200YOLO_MAX_CLASSES 4096
#define B200YOLO_MAX_SORT 65536  /* candidates per image the sort/NMS kernels accept (cap) */
#define B200YOLO_MAX_ANCHORS 65536 /* anchors per image: the sort key carries the anchor index in 16 bits */

int b200yolo_version(void);
const char* b200yolo_strerror(int code);

/* ---- K1: letterbox (+ normalise) ------------------------------------------------------------
 * Replaces ultralytics/data/augment.py::LetterBox.__call__ (cv2.resize INTER_LINEAR +
 * cv2.copyMakeBorder value=114) and, for the f32 variant, the rest of
 * ultralytics/engine/predictor.py::BasePredictor.preprocess (BGR->RGB, HWC->CHW, float, /255),
 * entered from detect.py:541 / yolo.py:361 / pipe.py:179.
 * src: B frames of H x W x 3 uint8 (BGR, interleaved), row pitch / frame stride in bytes.
 * The resized image is new_w x new_h placed at (left, top) inside outW x outH; the host computes
 * that geometry (manual_yolo_b200/geometry.py, same arithmetic as LetterBox).  Arithmetic is
 * cv2's 8-bit fixed point (11-bit coefficients), bit-exact for down- and up-scales.
 *   _u8_to_f32: dst = (B,3,outH,outW) float32, planes R,G,B when swap_rb != 0; value/255 (true division)
 *   _u8       : dst = (B,outH,outW,3) uint8, channel order preserved (plain LetterBox drop-in) */
int b200yolo_letterbox_u8_to_f32(const uint8_t* src, int B, int H, int W, int64_t src_pitch,
                                 int64_t src_batch_stride, float* dst, int outH, int outW, int new_w,
                                 int new_h, int top, int left, int pad_value, int swap_rb, void* stream);
/* half=True form (ultralytics predict(half=True): `im.half(); im /= 255` on the device): dst = (B,3,outH,outW)
 * float16; each value is fl16(fl32(v / 255)), as torch's fp16 division (computed in fp32) rounds it.  Halves the
 * bytes K1 writes.  The reference runs with half=False (runs/rank_classifier/args.yaml:42): optional. */
int b200yolo_letterbox_u8_to_f16(const uint8_t* src, int B, int H, int W, int64_t src_pitch,
                                 int64_t src_batch_stride, void* dst, int outH, int outW, int new_w, int new_h,
                                 int top, int left, int pad_value, int swap_rb, void* stream);
int b200yolo_letterbox_u8(const uint8_t* src, int B, int H, int W, int64_t src_pitch,
                          int64_t src_batch_stride, uint8_t* dst, int outH, int outW, int new_w,
                          int new_h, int top, int left, int pad_value, void* stream);

/* K1 in slice mode -- the front end of SAHI-style sliced prediction (sahi.predict.get_sliced_prediction as
 * called at pipe.py:183-194: 640x640 windows, overlap ratio 0.2; window geometry = sahi.slicing.get_slice_bboxes,
 * restated in manual_yolo_b200/geometry.py::slice_boxes).  Batch item f * n_slices + s is the slice_h x slice_w
 * window of frame f whose top-left corner is (slice_xy[2s], slice_xy[2s+1]) (HOST int array); each item is
 * letterboxed exactly like a frame of that size.  dst = (n_frames * n_slices, 3, outH, outW) float32. */
int b200yolo_letterbox_slices_u8_to_f32(const uint8_t* frames, int n_frames, int frame_h, int frame_w,
                                        int64_t pitch, int64_t frame_stride, const int* slice_xy, int n_slices,
                                        int slice_h, int slice_w, float* dst, int outH, int outW, int new_w,
                                        int new_h, int top, int left, int pad_value, int swap_rb, void* stream);

/* ---- K2: Detect-head decode + confidence filter + compaction ---------------------------------
 * Replaces ultralytics/nn/modules/head.py::Detect._inference (DFL softmax expectation,
 * dist2bbox(xywh) * stride, class sigmoid) fused with the head of
 * ultralytics/utils/ops.py::non_max_suppression (amax > conf, xywh2xyxy, cls.max, conf filter,
 * optional `classes` filter).  The head may be the concatenated (B, 64+nc, A) tensor or the three
 * per-level (B, 64+nc, Hi, Wi) tensors: each level is described by a b200yolo_level. */
typedef struct b200yolo_level {
  const float* ptr;        /* element (b, ch, i) lives at ptr[b*batch_stride + ch*chan_stride + i] */
  int64_t batch_stride;    /* in elements */
  int64_t chan_stride;     /* in elements */
  int h, w;                /* feature-map size; anchors are row-major, centre (x+0.5, y+0.5) */
  float stride;            /* 8 / 16 / 32 */
} b200yolo_level;

/* cand: (B, cap, 6) float32 rows [x1,y1,x2,y2,score,class] in letterboxed pixels, slot order
 * arbitrary; cand_anchor: (B, cap) int32 anchor index of each slot; cand_count: (B) int32, MUST be
 * zeroed by the caller; may exceed cap on overflow (slots beyond cap are dropped, the count is not
 * clamped so the host can detect it).  class_mask: optional nc-bit allow-list (uint32 words). */
int b200yolo_decode_filter(const b200yolo_level* levels, int n_levels, int B, int nc, float conf_thres,
                           const uint32_t* class_mask, float* cand, int* cand_anchor, int* cand_count,
                           int cap, void* stream);

/* First half of b200yolo_decode_filter only: class reduction + confidence filter + compaction.  The rows
 * of the survivors carry score and class (columns 4,5); their boxes are left to
 * b200yolo_postprocess_small, which decodes them inside the fused per-image kernel. */
int b200yolo_class_filter(const b200yolo_level* levels, int n_levels, int B, int nc, float conf_thres,
                          const uint32_t* class_mask, float* cand, int* cand_anchor, int* cand_count,
                          int cap, void* stream);

/* Same outputs from an already decoded UL-format prediction (B, 4+nc(+extra), A): rows 0-3 xywh,
 * rows 4.. class scores.  This is the non_max_suppression(prediction, ...) drop-in entry. */
int b200yolo_filter_decoded(const float* pred, int B, int channels, int nc, int A, float conf_thres,
                            const uint32_t* class_mask, float* cand, int* cand_anchor, int* cand_count,
                            int cap, void* stream);

/* ---- K3: per-image sort (score descending, anchor ascending on ties) + max_nms cap ----------
 * Replaces the stable descending sort inside torchvision.ops.nms and UL's `n > max_nms` cap.
 * order: (B, cap) int32, order[b][r] = slot of rank r for r < min(count, cap, max_nms).
 * workspace: b200yolo_workspace_bytes(B, cap) bytes, 16-byte aligned; pass the SAME workspace to b200yolo_nms.
 * It starts with a header (per image: how many entries of order[] are in order, and a fallback flag).  For
 * cap > 2048 only the best 2048 entries of an image are ordered (radix select + bitonic sort): greedy NMS stops at
 * max_det keeps long before it needs more; if it ever does, b200yolo_nms re-sorts that image completely and
 * re-runs it, so results are exact in every case.  workspace may be NULL for cap <= 12288: everything is then
 * fully sorted (no header, no fallback). */
int b200yolo_sort_topk(const float* cand, const int* cand_anchor, const int* cand_count, int B, int cap,
                       int max_nms, int* order, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K4: class-aware greedy NMS with max_det cap ---------------------------------------------
 * Replaces `boxes = x[:, :4] + cls * max_wh; i = torchvision.ops.nms(boxes, scores, iou); i[:max_det]`
 * of ops.non_max_suppression: fp32 class-offset boxes, IoU in fp32, comparison against the DOUBLE
 * threshold, NaN never suppresses.  out: (B, max_det, 6) rows in kept (score-descending) order;
 * out_anchor: (B, max_det) anchor indices (UL return_idxs); out_count: (B).
 * scale: optional (B, 5) float32 {gain, pad_x, pad_y, w0, h0}: when non-NULL the boxes are also
 * mapped to source pixels as ops.scale_boxes + clip_boxes does ((x - pad) / gain, clamp).
 * roi_cnt: optional (B) int32: number of kept detections of image b whose class is set in the
 * roi_nc-bit allow-list roi_class_mask (the *_rank classes) -- consumed by
 * b200yolo_roi_from_detections. */
int b200yolo_nms(const float* cand, const int* cand_anchor, const int* cand_count, const int* order, int B,
                 int cap, int max_nms, double iou_thres, float max_wh, int agnostic, int max_det,
                 const float* scale, float* out, int* out_anchor, int* out_count,
                 const uint32_t* roi_class_mask, int roi_nc, int* roi_cnt, void* workspace,
                 size_t workspace_bytes, void* stream);

/* ---- K2b+K3+K4 fused for the sparse regime (cap <= 1024 candidates per image) -------------------
 * One launch, one CTA per image, all in shared memory: DFL box decode of the survivors (levels != NULL:
 * rows come from b200yolo_class_filter; levels == NULL: rows already hold boxes), the K3 sort, the K4
 * NMS with max_det, optional scale_boxes, ROI counts.  Same outputs and bit-exact same results as
 * b200yolo_sort_topk + b200yolo_nms.  An image whose cand_count exceeds cap is processed on its first
 * cap slots only: the caller must detect it -- cand_seen[b] > cap -- and treat the image as failed (the
 * Python host layer raises; SURVEY.md section 8(b): "raises rather than truncating silently").
 * cand_seen: optional (B) int32.  When non-NULL the kernel copies the unclamped cand_count[b] into it and then
 * RESETS cand_count[b] to 0 (cand_count is written in that case): the compaction counters are re-armed for the
 * next step without a memset launch, and cand_seen is what the host reads back. */
int b200yolo_postprocess_small(const b200yolo_level* levels, int n_levels, float* cand, const int* cand_anchor,
                               const int* cand_count, int B, int cap, int max_nms, double iou_thres,
                               float max_wh, int agnostic, int max_det, const float* scale, float* out,
                               int* out_anchor, int* out_count, const uint32_t* roi_class_mask, int roi_nc,
                               int* roi_cnt, int* cand_seen, void* stream);

/* ---- K2b+K3+K4 chained for the dense regime (cap > 1024; e.g. the conf = 0.001 evaluation setting) ----------
 * Input: the survivors of b200yolo_class_filter (scores, classes, anchors; boxes not decoded yet).  One host call
 * enqueues: select + sort of the best 2048 entries per image; DFL box decode of the best 512 of them (one NMS window,
 * visited in slot order so that the head tensor is read sector by sector); windowed NMS, which decodes the boxes of
 * any further window itself when it loads it; then, for images whose NMS ran out of ordered entries before max_det
 * keeps (flag in the workspace header), the full sort and the NMS again.  Same outputs, bit for bit, as
 * b200yolo_decode_filter + b200yolo_sort_topk + b200yolo_nms; the work is proportional to what the NMS consumes
 * (about 500 of 8400 candidates per image at conf = 0.001, max_det = 300).
 * order: (B, cap) int32 scratch; workspace: b200yolo_workspace_bytes(B, cap) bytes (required). */
int b200yolo_postprocess_dense(const b200yolo_level* levels, int n_levels, float* cand, const int* cand_anchor,
                               const int* cand_count, int B, int cap, int max_nms, double iou_thres, float max_wh,
                               int agnostic, int max_det, const float* scale, float* out, int* out_anchor,
                               int* out_count, const uint32_t* roi_class_mask, int roi_nc, int* roi_cnt, int* order,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---- a10: ops.scale_boxes + clip_boxes on a flat (n,>=4) xyxy array (in place) --------------- */
int b200yolo_scale_boxes(float* boxes, int n, int row_stride, float gain, float pad_x, float pad_y,
                         float w0, float h0, void* stream);

/* ---- K5: ROI crop + resize to size x size (rank-classifier batch) ----------------------------
 * Replaces detect.py:100-113 safe_crop (after the consumers' int() truncation, detect.py:581) and
 * ultralytics ClassificationPredictor.preprocess with the checkpoint's transforms
 * (Resize(size, bilinear, antialias) -> CenterCrop(size) -> ToTensor -> Normalize(0,1)), i.e. Pillow's
 * two-pass 8-bit fixed-point resample, for every ROI of a batch in one launch (detect.py:121 calls
 * the predictor once per crop).
 * frames: (B,H,W,3) uint8 BGR; boxes: (N,4) float32 xyxy source pixels; batch_idx: (N) int32;
 * roi_count: optional device int32 -- when non-NULL only the first min(*roi_count, N) ROIs are done.
 * dst: (N,3,size,size) float32 RGB in [0,1] (size must be 64); valid: (N) int32: 1 = ok, 2 = ok and
 * produced by the general split launch (resample scale > 3 or > 198 referenced columns: far beyond a rank
 * card), 0 = safe_crop returns None, -1 = ROI
 * beyond the envelope (short side > 31*size); the dst rows of invalid ROIs are zero-filled. */
int b200yolo_roi_crop_resize(const uint8_t* frames, int B, int H, int W, int64_t pitch,
                             int64_t batch_stride, const float* boxes, const int* batch_idx,
                             const int* roi_count, int N, int pad, int size, float* dst, int* valid,
                             void* stream);

/* Pipeline form of K5: ROI g (CTA g of roi_cap) is the g-th detection, image-major and in kept order,
 * whose class is in class_mask; it is located on the device from det (B,max_det,6), det_count (B) and
 * the per-image counts roi_cnt (B) written by b200yolo_nms -- no selection launch, no host round trip
 * (the reference loops over detections on the host, detect.py:580-588).  dst: (roi_cap,3,size,size);
 * roi_batch / roi_det / valid: (roi_cap); roi_total: (1) = sum roi_cnt, NOT clamped: a value above roi_cap tells the
 * caller that detections beyond the first roi_cap were not cropped (consumers use min(roi_total, roi_cap) rows). */
int b200yolo_roi_from_detections(const uint8_t* frames, int B, int H, int W, int64_t pitch,
                                 int64_t batch_stride, const float* det, const int* det_count,
                                 const int* roi_cnt, int max_det, const uint32_t* class_mask, int nc,
                                 int pad, int size, float* dst, int* roi_batch, int* roi_det, int* valid,
                                 int* roi_total, int roi_cap, void* stream);

/* Gather the detections of selected classes (the *_rank ids) into a dense ROI list, image-major,
 * detection order preserved: feeds K5 without a host round trip (detect.py:580-588 loop).
 * det: (B,max_det,6) from b200yolo_nms (source pixels); roi_boxes: (roi_cap,4); roi_batch:
 * (roi_cap); roi_det: (roi_cap) detection row index; roi_count: (1) = number of selected detections, NOT clamped
 * (rows beyond roi_cap are dropped; the count tells). */
int b200yolo_select_rois(const float* det, const int* det_count, int B, int max_det,
                         const uint32_t* class_mask, int nc, float* roi_boxes, int* roi_batch,
                         int* roi_det, int* roi_count, int roi_cap, void* stream);

/* Merge input of sliced prediction: det (n_frames * n_slices, max_det, 6) / det_count from b200yolo_nms hold each
 * slice's kept detections in slice pixels; per frame they are concatenated (slice-major, rank order), shifted by
 * the slice origin (SAHI shift_amount) and written as candidates (cand (n_frames, cap, 6), cand_anchor =
 * slice * max_det + rank, cand_count not clamped).  full_det / full_count: optional (n_frames, max_det, 6) / (n_frames)
 * detections of the FULL frame in frame pixels -- SAHI's perform_standard_pred=True, its default -- appended after the
 * slices (cand_anchor = n_slices * max_det + rank).  The candidates feed b200yolo_greedy_nmm (SAHI's default merge)
 * or one more b200yolo_sort_topk + b200yolo_nms. */
int b200yolo_gather_slice_detections(const float* det, const int* det_count, int n_frames, int n_slices,
                                     int max_det, const int* slice_xy, const float* full_det, const int* full_count,
                                     float* cand, int* cand_anchor, int* cand_count, int cap, void* stream);

/* SAHI's default merge of the gathered predictions (sahi.predict.get_sliced_prediction as called at pipe.py:186-193:
 * postprocess_type "GREEDYNMM", match_metric "IOS" = 1 ("IOU" = 0), match_threshold 0.5, class-aware): greedy
 * non-maximum MERGING -- a matched box is merged into the kept one (union box, max score) instead of being dropped.
 * cand / cand_src / cand_count: as written by b200yolo_gather_slice_detections (cap <= 8192).  out: (n_frames,
 * max_det, 6) merged rows in descending score of the kept boxes; out_src: (n_frames, max_det) provenance of each kept
 * box; out_count: (n_frames).  roi_class_mask / roi_nc / roi_cnt: as for b200yolo_nms. */
int b200yolo_greedy_nmm(const float* cand, const int* cand_src, const int* cand_count, int n_frames, int cap,
                        int match_metric, double match_threshold, int agnostic, int max_det, float* out, int* out_src,
                        int* out_count, const uint32_t* roi_class_mask, int roi_nc, int* roi_cnt, void* stream);

/* ---- N2: tracker association costs -------------------------------------------------------------------------
 * The reference hands each frame's detections to supervision ByteTrack (detect.py:557); its association step is
 * cost = 1 - box_iou_batch(tracks, detections), optionally fused with the detection scores
 * (1 - (1 - cost) * score), then a linear assignment on the host.  tracks: (B, T, 4) float32 xyxy (16-byte
 * aligned) + track_count (B); det / det_count: the padded output of b200yolo_nms.  cost: (B, T, max_det)
 * float32; entries outside (track_count[b], det_count[b]) are set to pad_cost. */
int b200yolo_iou_cost_matrix(const float* tracks, const int* track_count, const float* det, const int* det_count,
                             int B, int T, int max_det, int fuse_score, float pad_cost, float* cost, void* stream);

/* ---- N2: ByteTrack state arithmetic (supervision ByteTrack as the reference uses it, detect.py:22,557) --------------
 * Batched over tracks, on state arrays that stay on the device: mean (capacity, 8) float64 [x, y, a, h, vx, vy, va, vh],
 * cov (capacity, 8, 8) float64.  slots: (n) int32 rows of those arrays.
 *   predict : STrack.multi_predict -- zero_vh[t] = 0: constant-velocity Kalman prediction; 1: vh zeroed first (tracks not
 *             in the Tracked state); 2: no prediction (unconfirmed tracks / reading the current boxes).  tlbr (n,4)
 *             float32 (optional) receives each track's box afterwards, ready to be the `tracks` argument of
 *             b200yolo_iou_cost_matrix;
 *   update  : KalmanFilter.update of track slots[t] with the detection box boxes[box_idx[t] * box_stride .. +4) (xyxy
 *             float32, e.g. rows of the NMS output with box_stride = 6);
 *   initiate: KalmanFilter.initiate of slots[t] from such a box (new tracks). */
int b200yolo_kalman_predict(double* mean, double* cov, const int* slots, const int* zero_vh, int n, float* tlbr,
                            void* stream);
int b200yolo_kalman_update(double* mean, double* cov, const int* slots, const float* boxes, int box_stride,
                           const int* box_idx, int n, void* stream);
int b200yolo_kalman_initiate(double* mean, double* cov, const int* slots, const float* boxes, int box_stride,
                             const int* box_idx, int n, void* stream);

size_t b200yolo_workspace_bytes(int B, int cap);

/* ---- host -> device staging of the source rows K1 references ------------------------------------
 * The one entry point that takes a HOST pointer (pinned memory, or the copy is not asynchronous).
 * Replaces the `.to(device)` of ultralytics/engine/predictor.py::BasePredictor.preprocess for the
 * frames of a batch (detect.py:541 passes one host numpy frame per call).  Copies rows
 * row0, row0 + row_step, ... (n_rows of them) of each of the B host frames (H x W x 3 uint8, row
 * pitch / frame stride in bytes) into dev_rows (B, n_rows, W, 3).  manual_yolo_b200/geometry.py::
 * referenced_rows() gives (row0, row_step, n_rows): when the vertical scale is an odd integer k the
 * bilinear weights are (2048, 0) and K1 reads one row in k, so the staged image is passed to
 * b200yolo_letterbox_* with H = new_h (vertical identity) and gives bit-identical output. */
int b200yolo_stage_rows_h2d(const uint8_t* host_frames, int B, int H, int W, int64_t pitch,
                            int64_t batch_stride, int row0, int row_step, int n_rows, uint8_t* dev_rows,
                            int64_t dev_pitch, int64_t dev_batch_stride, void* stream);

/* Generic strided host -> device copy (one 2-D DMA, pinned source): `height` runs of `width` bytes.
 * Stages only the class channels of a host Detect-head tensor (one run per frame) when the DFL
 * channels of the few survivors are read zero-copy by b200yolo_postprocess_small. */
int b200yolo_copy2d_h2d(void* dev_dst, int64_t dpitch, const void* host_src, int64_t spitch, int64_t width,
                        int64_t height, void* stream);

/* ---- device self-test of the fast fp32 forms inside the DFL decode (test hook, no reference analogue) ----
 * The decode must be bit-identical to torch's CPU softmax (Sleef expf_u10, true division).  Two of its
 * inner forms are cheaper restatements; this entry proves them on the device:
 *   mode 0: the exponent-add scaling of exp(d), d <= 0, against the two-step Sleef form for EVERY float in
 *           [-80, 0] (n ignored);
 *   mode 1: quotient-from-reciprocal (one rcp + Markstein correction) against IEEE division over n
 *           pseudo-random (numerator, denominator) pairs of the softmax domain.
 * *mismatches (device uint64, caller-zeroed) receives the number of disagreeing inputs: must stay 0. */
int b200yolo_selftest_math(int mode, uint64_t n, uint64_t* mismatches, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200YOLO_H_ */

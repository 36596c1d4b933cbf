/* hd_b200.h -- C ABI of libhd_b200.so: the B200 (sm_100a) post-CNN box pipeline of
 * HeltonDetection (YOLOv5 decode/filter, IoU, NMS, RPN proposals, RoIAlign/RoIPool, WBF).
 *
 * Drop-in boundary.  The reference (README.md:2 "based on PyTorch") reaches this path through
 * Python functions over torch ATen ops and the torchvision ops
 *     torchvision::nms(Tensor dets, Tensor scores, float iou_threshold) -> Tensor
 *     torchvision::roi_align(Tensor input, Tensor rois, float spatial_scale, SymInt pooled_height,
 *                            SymInt pooled_width, int sampling_ratio, bool aligned) -> Tensor
 *     torchvision::roi_pool(Tensor input, Tensor rois, float spatial_scale, SymInt pooled_height,
 *                           SymInt pooled_width) -> (Tensor, Tensor)
 * (torchvision/ops/boxes.py:48, roi_align.py:243-260, roi_pool.py:15-53).  The dev-branch call
 * sites are not in the mount (README.md:6); each entry point below cites the reference feature
 * (README.md:N) and the executable torchvision interface it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter is documented "host";
 *   - plain C types only; `stream` is a cudaStream_t passed as void*;
 *   - no allocation inside: the caller passes `workspace` of at least hd_<op>_workspace_size() bytes
 *     (256-byte aligned);
 *   - kernels are enqueued on `stream`; no entry point synchronises the device;
 *   - return 0 on success, a negative HD_ERR_* otherwise; hd_last_error() (thread-local) has the text.
 *     No C++ exception crosses the ABI.
 */
#ifndef HD_B200_H
#define HD_B200_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define HD_API __attribute__((visibility("default")))
#else
#define HD_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define HD_OK 0
#define HD_ERR_INVALID (-1)   /* bad argument */
#define HD_ERR_CUDA (-2)      /* CUDA runtime error at launch */
#define HD_ERR_WORKSPACE (-3) /* workspace too small */

#define HD_MAX_LEVELS 8
#define HD_MAX_ANCHORS 8

/* flags */
#define HD_FLAG_CONF_GE 1     /* keep conf >= thr (bubbliiiing) instead of conf > thr (ultralytics) */
#define HD_FLAG_DENSE_READ 2  /* stream every head element (no objectness-tile skip) */

/* class handling of the batched NMS */
#define HD_NMS_AGNOSTIC 0
#define HD_NMS_CLASS_EXACT 1  /* suppress only equal class ids (= torchvision _batched_nms_vanilla) */
#define HD_NMS_CLASS_OFFSET 2 /* nms(boxes + cls*offset_scale), fp32 add (= ultralytics max_wh trick) */

HD_API int hd_version(void);
HD_API const char* hd_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * YOLOv5 head (README.md:9; lineage `decode_box` / `non_max_suppression`)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float* data;  /* device, [B, A*(5+nc), H, W] contiguous (the head's native NCHW) */
    int32_t H, W;
    float stride;                         /* 8 / 16 / 32 */
    float anchor_wh[2 * HD_MAX_ANCHORS];  /* (w,h) px of the A anchors of this level */
} hd_yolo_level; /* host struct */

/* Replaces decode_box: heads -> pred [B, total_anchors, 5+nc] = (cx,cy,w,h,obj,cls..) px,
 * flat index level_offset + (a*H+i)*W + j. */
HD_API int hd_yolo_decode(const hd_yolo_level* levels /*host*/, int n_levels, int B, int A, int nc, float* pred, void* stream);

/* Fused decode + sigmoid + confidence filter + compaction straight from the raw heads
 * (replaces decode_box followed by the filter half of non_max_suppression).
 * Candidates of image b are written, in unspecified order, to slot s < cand_count[b] of
 *   cand_box[b*cap+s] (xyxy px), cand_score (obj*cls), cand_cls (best class), cand_anchor (flat anchor index).
 * cap >= total anchors guarantees no overflow; on overflow the count saturates at cap.
 * cand_count is zeroed by the call. */
HD_API int hd_yolo_decode_filter(const hd_yolo_level* levels /*host*/, int n_levels, int B, int A, int nc, double conf_thres,
                          int flags, float* cand_box, float* cand_score, int32_t* cand_cls, int32_t* cand_anchor,
                          int32_t* cand_count, int cap, void* stream);

/* Same filter on an already decoded prediction [B, N, 5+nc] (drop-in for non_max_suppression(prediction,..)). */
HD_API int hd_yolo_filter_pred(const float* pred, int B, int N, int nc, double conf_thres, int flags, float* cand_box,
                        float* cand_score, int32_t* cand_cls, int32_t* cand_anchor, int32_t* cand_count, int cap,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batched per-image sort + greedy NMS (torchvision.ops.nms / batched_nms, boxes.py:20-120; A.4).
 * For every image b: take its cnt=counts[b] (or n_fixed if counts==NULL) candidates at stride `cap`,
 * order them by (score desc, tiebreak asc) [NaN score first], keep the first max_nms (<=0: all), run
 * greedy NMS (suppress iff IoU > iou_thres, compared in double like the CPU kernel) and write the
 * first max_det keeps, in score order:
 *   out_det[b, r, 0:6] = (x1,y1,x2,y2,score,cls)   (nullable)
 *   out_idx[b, r]      = tiebreak id of the kept box (slot index if tiebreak==NULL)   (nullable)
 *   out_count[b]       = number of keeps written.
 * tiebreak ids must be unique per image.  cls may be NULL for HD_NMS_AGNOSTIC.
 * ------------------------------------------------------------------------------------------- */
HD_API size_t hd_sort_nms_workspace_size(int B, int cap);
HD_API int hd_sort_nms_batched(const float* boxes, const float* scores, const int32_t* cls, const int32_t* tiebreak,
                        const int32_t* counts, int n_fixed, int B, int cap, double iou_thres, int class_mode,
                        float offset_scale, int max_nms, int max_det, float* out_det, int64_t* out_idx,
                        int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);

/* box_iou (boxes.py:308-370): iou[N,M] = inter / (area1 + area2 - inter), fp32, no eps. */
HD_API int hd_box_iou(const float* boxes1, int64_t N, const float* boxes2, int64_t M, float* iou, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HD_B200_H */

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
#define HD_FLAG_IN_F16 4      /* hd_yolo_decode_filter / hd_yolo_postprocess*: levels[].data points to IEEE half heads */
#define HD_FLAG_IN_BF16 8     /* ... to bfloat16 heads.  Elements are widened to fp32 on load (exact); all arithmetic is fp32 */
#define HD_FLAG_MULTI_LABEL 256 /* ultralytics multi_label: every (anchor, class) with obj*cls > thr is a candidate; index = anchor*nc + class.
                                 * Candidates of an image beyond the buffer capacity (= anchors per image) are dropped. NCHW heads only. */
#define HD_FLAG_IN_NHWC 128   /* ... to fp32 heads stored [B, H, W, A*(5+nc)] (torch channels_last of the NCHW head): the layout the
                               * final 1x1 conv produces; same candidates as the NCHW path */

/* class handling of the batched NMS */
#define HD_NMS_AGNOSTIC 0
#define HD_NMS_CLASS_EXACT 1  /* suppress only equal class ids (= torchvision _batched_nms_vanilla) */
#define HD_NMS_CLASS_OFFSET 2 /* nms(boxes + cls*offset_scale), fp32 add (= ultralytics max_wh trick) */

HD_API int hd_version(void);
/* developer aid: SM clock stamps of the phases of block 0 of the last sort_nms_kernel (which=0) or
 * rpn_select_nms_kernel (which=1); host array of 16; synchronises the device */
HD_API int hd_debug_phases(int which, long long* out16 /*host*/);
HD_API const char* hd_last_error(void);
/* kernels launched (or captured into a CUDA graph) by this library since it was loaded, over all threads and devices;
 * callers take differences (bench.py reports the launches of one step as `gpu_launches`) */
HD_API unsigned long long hd_debug_launch_count(void);
/* developer aid: cycle counters of the streamed RoIAlign kernel (enable != 0 allocates them on the current device; out8 != NULL
 * synchronises, copies the 8 counters {tables, waits, row loop, output, items, producer waits, -, -} to the host and clears them) */
HD_API int hd_debug_roi_profile(int enable, unsigned long long* out8 /*host, nullable*/);

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

/* One-call post-process: hd_yolo_decode_filter followed by hd_sort_nms_batched on an internal candidate buffer
 * carved from `workspace` (three launches: 1 KB memset, decode kernel, small-image NMS kernel; a fourth,
 * radix-sort kernel only does work for images with more than 512 candidates).  Outputs as hd_sort_nms_batched;
 * out_idx holds flat anchor indices.  `data` pointers may also be pinned, mapped HOST memory (UVA zero-copy):
 * with the default objectness skip only the 32-byte sectors of possible survivors cross PCIe. */
HD_API size_t hd_yolo_postprocess_workspace_size(int B, int total_anchors);
HD_API int hd_yolo_postprocess(const hd_yolo_level* levels /*host*/, int n_levels, int B, int A, int nc, double conf_thres,
                               double iou_thres, int flags, int class_mode, float offset_scale, int max_nms, int max_det,
                               float* out_det, int64_t* out_idx, int32_t* out_count, void* workspace, size_t workspace_bytes,
                               void* stream);

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
/* Images with more than 512 candidates are sorted and suppressed by one CTA each, or -- when the batch is small enough to
 * leave most SMs idle (B <= 33 on a B200) -- by a thread-block cluster of 4 or 8 CTAs per image; bit-identical outputs.
 * One-CTA images with more than 2048 candidates are ordered by a bucket sort on the score word (bitonic network as the fallback
 * for degenerate score distributions); the 1024-thread kernel is used while one of the last 64 calls on the same device
 * saw such an image (one word per device), a 256-thread variant (co-resident with other kernels) otherwise -- same bits either way.
 * hd_nms_set_mode: 0 = automatic (default), 1 = one CTA per image only, 2 = clusters whenever the batch allows, 3 = always the
 * 256-thread variant, 4 = always the 1024-thread variant, 5 = as 1 with the bitonic network for every size.  Returns the
 * previous mode (developer / test aid). */
HD_API int hd_nms_set_mode(int mode);
HD_API int hd_sort_nms_batched(const float* boxes, const float* scores, const int32_t* cls, const int32_t* tiebreak,
                        const int32_t* counts, int n_fixed, int B, int cap, double iou_thres, int class_mode,
                        float offset_scale, int max_nms, int max_det, float* out_det, int64_t* out_idx,
                        int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);

/* Replicated outputs (multi-GPU): besides out_det/out_count, every kept row and every count is also stored to
 * `n` more buffers -- typically the slices of the gather buffers of the peer GPUs, mapped into this process
 * (CUDA IPC / symmetric memory) -- so the all-gather of the padded detections happens in the NMS
 * kernel's epilogue as posted NVLink stores instead of a separate pack + NCCL collective.  det[i] addresses a
 * [B, max_det, 6] block, count[i] a [B] block, at the place this rank's images occupy in replica i. */
#define HD_MAX_REPLICAS 16
typedef struct {
    int32_t n;
    float* det[HD_MAX_REPLICAS];
    int32_t* count[HD_MAX_REPLICAS];
} hd_replicas; /* host struct */
HD_API int hd_sort_nms_batched_replicated(const float* boxes, const float* scores, const int32_t* cls, const int32_t* tiebreak,
                                          const int32_t* counts, int n_fixed, int B, int cap, double iou_thres, int class_mode,
                                          float offset_scale, int max_nms, int max_det, float* out_det, int64_t* out_idx,
                                          int32_t* out_count, const hd_replicas* replicas /*host, nullable*/, void* workspace,
                                          size_t workspace_bytes, void* stream);
HD_API int hd_yolo_postprocess_replicated(const hd_yolo_level* levels /*host*/, int n_levels, int B, int A, int nc,
                                          double conf_thres, double iou_thres, int flags, int class_mode, float offset_scale,
                                          int max_nms, int max_det, float* out_det, int64_t* out_idx, int32_t* out_count,
                                          const hd_replicas* replicas /*host, nullable*/, void* workspace,
                                          size_t workspace_bytes, void* stream);

/* box_iou (boxes.py:308-370): iou[N,M] = inter / (area1 + area2 - inter), fp32, no eps. */
HD_API int hd_box_iou(const float* boxes1, int64_t N, const float* boxes2, int64_t M, float* iou, void* stream);

/* ---------------------------------------------------------------------------------------------
 * RoIAlign / RoIPool / FPN level assignment (README.md:65,73-78; torchvision roi_align.py:204-260,
 * roi_pool.py:15-53, poolers.py:73-84,147-227).
 * rois [K,5] = (batch_idx, x1,y1,x2,y2) image px; out [K,C,PH,PW] (NCHW, like torchvision).
 * Multi-level: RoI k is pooled from levels[level_ids[k]] and written to out[k] (original RoI order,
 * as MultiScaleRoIAlign); level_ids may be NULL when n_levels == 1 (the README "P2" single-level heads).
 * layout: memory layout of every levels[l].data -- HD_LAYOUT_NCHW [B,C,H,W] or HD_LAYOUT_NHWC [B,H,W,C]
 * (torch channels_last).  NHWC is the fast path (channel-contiguous gathers, TMA tile store).
 * ------------------------------------------------------------------------------------------- */
#define HD_LAYOUT_NCHW 0
#define HD_LAYOUT_NHWC 1
typedef struct {
    const float* data; /* device */
    int32_t H, W;
    float spatial_scale; /* 1/stride */
} hd_roi_level;        /* host struct */

HD_API int hd_roi_align(const hd_roi_level* levels /*host*/, int n_levels, int layout, int C, const float* rois,
                        const int32_t* level_ids, int64_t K, int pooled_h, int pooled_w, int sampling_ratio, int aligned,
                        float* out, void* stream);

/* The same operation with the batch size and a caller workspace (hd_roi_align_workspace_size bytes; the answer depends only on
 * the shapes).  By default this runs the per-RoI kernels of hd_roi_align (the fastest measured on B200).  Two bucketed variants for
 * many RoIs on NHWC features (C % 32 == 0, output <= 7x7, 1 <= sampling_ratio <= 2) are opt-in through hd_roi_set_mode -- 2: the
 * STREAMED kernel (csrc/roi_strip.cu: RoIs bucketed by (image, level, x-strip) and sorted by first feature row on the device, rows
 * cross L2->SM once through a TMA-fed shared-memory ring), 3: the row-walk kernel; both return the same bits and both measured
 * slower.  Replaces: torchvision MultiScaleRoIAlign / roi_align call site (poolers.py:147-227,
 * roi_align.py:204-260; README.md:65). */
HD_API size_t hd_roi_align_workspace_size(const hd_roi_level* levels /*host*/, int n_levels, int C, int batch, int64_t K, int pooled_h,
                                          int pooled_w, int sampling_ratio);
HD_API int hd_roi_align_ws(const hd_roi_level* levels /*host*/, int n_levels, int layout, int C, int batch, const float* rois,
                           const int32_t* level_ids, int64_t K, int pooled_h, int pooled_w, int sampling_ratio, int aligned, float* out,
                           void* workspace, size_t workspace_bytes, void* stream);
/* developer experiment: channel-sliced gather kernel, `slices` CTAs per RoI, every CTA walking `chunk` consecutive
 * entries of the device list `list[0 .. *list_count)` (RoI indices the caller sorted by position).  sampling_ratio 1..2, NHWC. */
HD_API int hd_debug_roi_align_sliced(const hd_roi_level* levels /*host*/, int n_levels, int C, const float* rois, const int32_t* level_ids,
                                     int64_t K, int pooled_h, int pooled_w, int sampling_ratio, int aligned, float* out,
                                     const int32_t* list, const int32_t* list_count, int slices, int chunk, void* stream);
/* argmax (int32 [K,C,PH,PW], index h*W+w inside the channel plane, -1 for an empty bin) may be NULL. */
HD_API int hd_roi_pool(const hd_roi_level* levels /*host*/, int n_levels, int layout, int C, const float* rois,
                       const int32_t* level_ids, int64_t K, int pooled_h, int pooled_w, float* out, int32_t* argmax,
                       void* stream);
/* RoIAlign backward (torchvision _roi_align_backward): grad_out [K,C,PH,PW] is scattered into the gradient
 * buffers levels[l].data (written through, although the struct declares them const; same shapes/layout as the forward
 * inputs), which the caller has zero-filled: grad_in[b,c,y,x] += grad_out[k,c,ph,pw] * w_corner / count per sample.
 * NHWC with C % 4 == 0 uses 128-bit reductions; NCHW / other C use scalar atomics (single level only).  The order of the
 * floating-point additions is not deterministic. */
HD_API int hd_roi_align_backward(const float* grad_out, const float* rois, const int32_t* level_ids, int64_t K,
                                 const hd_roi_level* levels /*host*/, int n_levels, int layout, int C, int pooled_h, int pooled_w,
                                 int sampling_ratio, int aligned, void* stream);
/* RoIPool backward (torchvision _roi_pool_backward): grad_in[b,c,argmax[k,c,ph,pw]] += grad_out[k,c,ph,pw]; grad_in ([B,C,H,W] in the
 * given layout) must be zero-filled by the caller; argmax as written by hd_roi_pool. */
HD_API int hd_roi_pool_backward(const float* grad_out, const int32_t* argmax, const float* rois, int64_t K, float* grad_in, int layout, int C,
                                int H, int W, int pooled_h, int pooled_w, void* stream);
/* RoIAlign kernel choice (process-wide; developer / test aid): 0 = automatic (the gather kernels), 1 = gather kernels only,
 * 2 = experimental staged-row (TMA ring) kernel whenever eligible (NHWC, C % 4 == 0, sampling_ratio 1 or 2, output <= 8x8;
 * bit-identical results, currently slower).  Returns the previous mode. */
HD_API int hd_roi_set_mode(int mode);
/* [B,C,H,W] -> [B,H,W,C] layout pass used in front of the NHWC kernels. */
HD_API int hd_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, void* stream);
/* Level of each RoI.  rois: rows of roi_stride floats with the xyxy box at box_offset.
 * style 0 (torchvision LevelMapper): floor(canonical_level + log2(sqrt(area)/canonical_scale) + eps) clamped to
 *   [k_min,k_max], minus k_min;  style 1 (mmdet): floor(log2(sqrt(area)/finest_scale + eps)) clamped to [0,k_max-k_min],
 *   finest_scale = canonical_scale / 2^(canonical_level-k_min).  Either output may be NULL. */
HD_API int hd_roi_level_map(const float* rois, int roi_stride, int box_offset, int64_t K, int style, int k_min, int k_max,
                            float canonical_scale, float canonical_level, float eps, int32_t* levels32, int64_t* levels64,
                            void* stream);

/* ---------------------------------------------------------------------------------------------
 * RoI-head output post-process -- the FasterRCNN final stage (README.md:8).  Replaces
 * torchvision RoIHeads.postprocess_detections (models/detection/roi_heads.py:668-723) with BoxCoder.decode_single
 * (_utils.py:183-224), and the lineage DecodeBox (bubbliiiing frcnn) as a flag variant.
 *   cls_logits [B*R, n_class] (class 0 = background), box_deltas [B*R, n_class*4], rois [B*R, 5] = (b,x1,y1,x2,y2)
 *   with the R rows of image b contiguous (the hd_rpn_proposals layout); roi_count[b] (nullable) = valid rows.
 * Per (roi, class c >= 1): score = softmax(logits)[c]; box = decode(roi, deltas[c] / weights)  [weights 10,10,5,5]
 * -> clip to the image -> keep score > score_thresh and w,h >= min_size (pass -INFINITY to skip the size test)
 * -> candidate (box, score, label c, id = r*(n_class-1) + c-1).  hd_roi_head_postprocess then runs the class-aware
 * NMS (suppress only equal labels, = batched_nms) and writes the first max_det keeps by score:
 *   out_det [B, max_det, 6] = (x1,y1,x2,y2,score,label), out_idx [B, max_det] = candidate id, out_count [B].
 * ------------------------------------------------------------------------------------------- */
#define HD_ROIHEAD_MUL_STD 16      /* weights are multipliers (std 0.1,0.1,0.2,0.2: lineage) instead of divisors (10,10,5,5) */
#define HD_ROIHEAD_CLAMP_DWH 32    /* clamp dw,dh to clamp_dwh before exp (torchvision bbox_xform_clip = log(1000/16)) */
#define HD_ROIHEAD_LABEL_MINUS1 64 /* labels c-1 (lineage) instead of c (torchvision) */
HD_API int hd_roi_head_decode_filter(const float* cls_logits, const float* box_deltas, const float* rois, const int32_t* roi_count,
                                     int B, int R, int n_class, const float* weights /*host[4]*/, int flags, float clamp_dwh,
                                     float img_h, float img_w, double score_thresh, float min_size, float* cand_box,
                                     float* cand_score, int32_t* cand_cls, int32_t* cand_id, int32_t* cand_count, int cap,
                                     void* stream);
HD_API size_t hd_roi_head_postprocess_workspace_size(int B, int R, int n_class);
HD_API int hd_roi_head_postprocess(const float* cls_logits, const float* box_deltas, const float* rois, const int32_t* roi_count,
                                   int B, int R, int n_class, const float* weights /*host[4]*/, int flags, float clamp_dwh,
                                   float img_h, float img_w, double score_thresh, float min_size, double nms_iou, int max_det,
                                   float* out_det, int64_t* out_idx, int32_t* out_count, void* workspace, size_t workspace_bytes,
                                   void* stream);

/* Output formats: letterbox inverse + clip of padded detections [B, max_det, 6] (ultralytics
 * scale_coords / clip_coords) and optionally the COCO result-json box (x, y, w, h).  meta [B,5] (device) =
 * (pad_x, pad_y, gain, w0, h0) per image; rows >= count[b] (count nullable) are zeroed.  out may alias det. */
#define HD_BOX_XYWH 1
HD_API int hd_scale_detections(const float* det, const int32_t* count, int B, int max_det, const float* meta, int flags, float* out,
                               void* stream);

/* ---------------------------------------------------------------------------------------------
 * IoU-based label assignment (training-side user of box_iou): torchvision det_utils.Matcher
 * (models/detection/_utils.py:318-400) over box_iou (boxes.py:308-370) without materialising the [G,N] matrix.
 *   gt_boxes [B, Gmax, 4] with gt_count[b] (nullable: all Gmax) valid rows; pred_boxes [N,4] shared by all images
 *   (anchors; pred_per_image = 0) or [B,N,4] (proposals; pred_per_image = 1).
 *   matches [B,N] int64: index of the best GT (first on ties), -1 if its IoU < low, -2 if low <= IoU < high; with
 *   allow_low_quality every prediction that attains some GT's best IoU keeps its arg-max match.
 *   matched_iou [B,N] (nullable): the best IoU.  Boxes are expected proper (x2 >= x1, y2 >= y1).
 * ------------------------------------------------------------------------------------------- */
/* Regression targets once the matching is known: torchvision BoxCoder.encode_single (_utils.py:75-110) / lineage bbox2loc.
 * targets[b,i] = weights * encode(gt[b, max(matches[b,i],0)], pred[(b,)i]); matches NULL = elementwise (Gmax == N).  weights: host[4]. */
HD_API int hd_box_encode(const float* gt_boxes, int B, int Gmax, const int64_t* matches, const float* pred_boxes, int pred_per_image, int N,
                         const float* weights /*host[4]*/, float* targets, void* stream);
HD_API size_t hd_match_workspace_size(int B, int Gmax, int N);
HD_API int hd_match(const float* gt_boxes, const int32_t* gt_count, int B, int Gmax, const float* pred_boxes, int pred_per_image, int N,
                    double high_threshold, double low_threshold, int allow_low_quality, int64_t* matches, float* matched_iou,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * RPN proposal creation (README.md:8,63-65; lineage ProposalCreator / loc2bbox, SURVEY.md A.3;
 * cross-check torchvision models/detection/rpn.py:231-297, _utils.py:183-224).
 * Per level: objectness [B, A, H, W] (sigmoid) or [B, 2A, H, W] (HD_RPN_SOFTMAX, channel a*2+{bg,fg}),
 * deltas [B, 4A, H, W] (channel a*4+k), anchors generated on the fly as anchor_base[a] + (j,i,j,i)*stride.
 * Flat proposal index inside an image = level_offset + (i*W+j)*A + a.
 * Pipeline per image: decode -> clip to [0,img_w]x[0,img_h] -> drop w or h < min_size -> top n_pre by
 * (score desc, index asc) -> greedy NMS(nms_iou) -> first n_post.  The lineage's random re-sampling pad
 * is not reproduced: short results are zero-padded and out_count[b] gives the number of real rows.
 *   out_rois   [B, n_post, 5] = (b, x1,y1,x2,y2)  -- directly the roi_align input
 *   out_scores [B, n_post] (nullable), out_idx [B, n_post] flat proposal index or -1 (nullable)
 * ------------------------------------------------------------------------------------------- */
#define HD_RPN_SOFTMAX 1   /* 2-channel softmax objectness instead of 1-channel sigmoid */
#define HD_RPN_CLAMP_DWH 2 /* clamp dw,dh to clamp_dwh before exp (torchvision bbox_xform_clip) */
#define HD_RPN_KEY_LOGIT 4 /* sigmoid mode: keys[] order the raw logits instead of the probabilities (torchvision's per-level top-k) */
#define HD_RPN_EXACT_MATH 8 /* sigmoid / softmax / exp(dw), exp(dh) are evaluated in fp64 and rounded once to fp32, i.e. the correctly
                             * rounded fp32 value (up to ~1e-8 of the inputs): scores and boxes then do not depend on which libm /
                             * SIMD exp produced them, so the whole decode -> top-k -> NMS chain is bit-reproducible against a CPU
                             * that does the same (oracle `exact_math=True`).  Default (flag clear): fp32 expf, as the reference. */
typedef struct {
    const float* objectness; /* device */
    const float* deltas;     /* device */
    int32_t H, W;
    float stride;
    float anchor_base[4 * HD_MAX_ANCHORS]; /* (x1,y1,x2,y2) of the A base anchors, centred on cell (0,0) */
} hd_rpn_level;                            /* host struct */

HD_API int hd_rpn_num_anchors(const hd_rpn_level* levels /*host*/, int n_levels, int A);
/* stage 1 only: boxes [B,N,4], scores [B,N], keys [B,N] (sortable score bits, 0 = dropped by min_size) */
HD_API int hd_rpn_decode(const hd_rpn_level* levels /*host*/, int n_levels, int B, int A, int flags, float img_h, float img_w,
                         float min_size, float clamp_dwh, float* boxes, float* scores, uint32_t* keys, void* stream);
/* stage 2 only: top-k + sort + NMS on the stage-1 arrays */
HD_API size_t hd_rpn_select_nms_workspace_size(int B, int N, int n_pre);
HD_API int hd_rpn_select_nms(const float* boxes, const float* scores, const uint32_t* keys, int B, int N, int n_pre, int n_post,
                             double nms_iou, float* out_rois, float* out_scores, int64_t* out_idx, int32_t* out_count,
                             void* workspace, size_t workspace_bytes, void* stream);
/* the same on a level slice (or any sub-range) of wider per-image arrays: image b's N elements start at b * image_stride */
HD_API int hd_rpn_select_nms_strided(const float* boxes, const float* scores, const uint32_t* keys, int B, int N, int64_t image_stride,
                                     int n_pre, int n_post, double nms_iou, float* out_rois, float* out_scores, int64_t* out_idx,
                                     int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);
/* Stage 2 runs as one thread-block CLUSTER (8 SMs) per image when n_pre <= 16384, else as one CTA per image; both
 * give bit-identical outputs.  hd_rpn_set_mode: 0 = automatic (default), 1 = force one CTA per image,
 * 2 = cluster whenever eligible.  Returns the previous mode (process-wide; developer / test aid). */
HD_API int hd_rpn_set_mode(int mode);
/* number of clusters (of the given size) of the stage-2 kernel the current device holds at once (developer aid; -1 on error) */
HD_API int hd_rpn_cluster_capacity(int cluster_size /* 1, 2, 4 or 8 */);
/* force the cluster size of stage 2 (1, 2, 4, 8; 0 = chosen per launch from batch size and device capacity); returns the previous value */
HD_API int hd_rpn_set_cluster_size(int cluster_size);
/* both stages */
HD_API size_t hd_rpn_proposals_workspace_size(int B, int N, int n_pre);
HD_API int hd_rpn_proposals(const hd_rpn_level* levels /*host*/, int n_levels, int B, int A, int flags, float img_h, float img_w,
                            float min_size, float clamp_dwh, int n_pre, int n_post, double nms_iou, float* out_rois,
                            float* out_scores, int64_t* out_idx, int32_t* out_count, void* workspace, size_t workspace_bytes,
                            void* stream);
/* Per-level selection (torchvision RegionProposalNetwork.filter_proposals, rpn.py:231-297), composed on the device:
 * hd_rpn_decode(HD_RPN_KEY_LOGIT | HD_RPN_CLAMP_DWH) -> one hd_rpn_select_nms per level slice with nms_iou = 2 (= stable top-k + sort;
 * the levels are independent: issue them on different streams) -> hd_rpn_merge_levels -> hd_sort_nms_batched(HD_NMS_CLASS_EXACT, cls =
 * level) -> hd_rpn_finish_levels.
 *   hd_rpn_merge_levels: sel[l] (HOST array of L device pointers) = [B, k[l]] int64 indices inside level l's slice (-1 = empty) as
 *     written by hd_rpn_select_nms(out_idx); level_off[l] = first flat anchor of level l.  Gathers the boxes / probabilities, drops boxes
 *     with a side < min_size or a probability < score_thresh, and compacts the survivors in (level, score) order:
 *     cand_box [B,K,4], cand_score [B,K], cand_lvl [B,K], cand_anchor [B,K] (flat anchor index), cand_count [B], K = sum k[l].
 *   hd_rpn_finish_levels: det / slot / count of the NMS call -> rois [B*n_post,5] = (b,x1,y1,x2,y2), scores [B,n_post] (nullable),
 *     idx [B,n_post] flat anchor index (nullable); rows beyond the count are zero, index -1. */
HD_API int hd_rpn_merge_levels(const float* boxes, const float* scores, const int64_t* const* sel /*host*/, const int32_t* k /*host*/,
                               const int32_t* level_off /*host*/, int n_levels, int B, int N, float min_size, float score_thresh,
                               float* cand_box, float* cand_score, int32_t* cand_lvl, int32_t* cand_anchor, int32_t* cand_count, void* stream);
HD_API int hd_rpn_finish_levels(const float* det, const int64_t* slot, const int32_t* count, const int32_t* cand_anchor, int B, int K,
                                int n_post, float* rois, float* scores, int64_t* idx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * TTA map-back + Weighted Boxes Fusion (README.md:19; ensemble-boxes weighted_boxes_fusion).
 * WBF inputs are padded per (image, view): boxes [B,V,M,4] xyxy normalised to [0,1], scores [B,V,M],
 * labels [B,V,M] (float holding integers in [0,num_labels)), counts [B,V] valid rows per view.
 * weights: HOST array of V doubles (NULL = all 1).  Outputs, sorted by fused score desc, padded to V*M rows:
 *   out_boxes [B,V*M,4] f32, out_scores [B,V*M] f64 (the reference returns float64), out_labels [B,V*M] f32,
 *   out_count [B]; rows of image b at and beyond out_count[b] are not written.
 * hd_tta_map_back writes view v of the WBF inputs from that view's detections det [B,max_det,6] (view px):
 *   un-flip x' = view_w - x (corners swapped), / scale, / (img_w, img_h).
 * ------------------------------------------------------------------------------------------- */
#define HD_WBF_AVG 0
#define HD_WBF_MAX 1
#define HD_WBF_BOX_AND_MODEL_AVG 2       /* score * n / sum(w of the boxes) * sum(w of the distinct models in the cluster) / sum(weights) */
#define HD_WBF_ABSENT_MODEL_AWARE_AVG 3  /* score * n / (sum(w of the boxes) + sum(w of the models absent from the cluster)) */
/* OR-ed into conf_type: the 'avg' rescale uses min(n, sum(weights)) (ensemble-boxes <= 1.0.4) instead of min(n, len(weights))
 * (ensemble-boxes >= 1.0.5, the default).  The two agree for unit weights. */
#define HD_WBF_RESCALE_SUM_WEIGHTS 256
HD_API size_t hd_wbf_workspace_size(int B, int V, int M, int num_labels);
HD_API int hd_wbf(const float* boxes, const float* scores, const float* labels, const int32_t* counts, int B, int V, int M,
                  int num_labels, const double* weights /*host, nullable*/, double iou_thr, double skip_box_thr, int conf_type,
                  int allows_overflow, float* out_boxes, double* out_scores, float* out_labels, int32_t* out_count,
                  void* workspace, size_t workspace_bytes, void* stream);
HD_API int hd_tta_map_back(const float* det, const int32_t* count, int B, int max_det, float scale, int hflip, float view_w,
                           float img_w, float img_h, float* boxes, float* scores, float* labels, int32_t* counts, int V, int v,
                           int M, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HD_B200_H */

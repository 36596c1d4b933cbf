// RoI-head output post-process (the FasterRCNN final stage, README.md:8): softmax over the class
// logits, per-class box decode (std 0.1,0.1,0.2,0.2 = weights 10,10,5,5), clip, score / min-size filter, compaction;
// the class-aware NMS + top-k then runs in hd_sort_nms_batched.  Restates torchvision RoIHeads.postprocess_detections
// (models/detection/roi_heads.py:668-723) with BoxCoder.decode_single (_utils.py:183-224); the lineage DecodeBox
// (bubbliiiing frcnn utils_bbox.py) is the HD_ROIHEAD_MUL_STD | HD_ROIHEAD_LABEL_MINUS1 variant.
#include "hd_common.cuh"

struct RoiHeadParams {
    const float* logits;   // [B*R, n_class]
    const float* deltas;   // [B*R, n_class*4]
    const float* rois;     // [B*R, 5]
    const int* roi_count;  // [B] or NULL
    int B, R, n_class, flags;
    float wx, wy, ww, wh, clamp_dwh, img_h, img_w, score_thr, min_size;
    float4* cand_box; float* cand_score; int* cand_cls; int* cand_id; int* cand_count;
    int cap;
};

// one warp per RoI: lanes stride over the classes (coalesced rows of the two head tensors)
__global__ void __launch_bounds__(256) roi_head_decode_filter_kernel(const __grid_constant__ RoiHeadParams p) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (long long)p.B * p.R) return;
    const int b = (int)(row / p.R), r = (int)(row - (long long)b * p.R);
    if (p.roi_count && r >= p.roi_count[b]) return;
    const float* __restrict__ lg = p.logits + row * p.n_class;
    // softmax exactly as ATen's vectorised CPU kernel states it: max, sum of exp(x - max), exp(x - max) / sum
    float m = -INFINITY;
    for (int c = lane; c < p.n_class; c += 32) m = fmaxf(m, lg[c]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(HD_FULL, m, d));
    float s = 0.0f;
    for (int c = lane; c < p.n_class; c += 32) s += expf(__fsub_rn(lg[c], m));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(HD_FULL, s, d);
    const float* roi = p.rois + row * 5;
    const float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
    const float wa = __fsub_rn(x2, x1), ha = __fsub_rn(y2, y1);
    const float cxa = __fadd_rn(x1, __fmul_rn(0.5f, wa)), cya = __fadd_rn(y1, __fmul_rn(0.5f, ha));
    const float* __restrict__ dl = p.deltas + row * p.n_class * 4;
    const bool mul = (p.flags & HD_ROIHEAD_MUL_STD) != 0;
    const int nfg = p.n_class - 1;
    for (int c0 = 1; c0 < p.n_class; c0 += 32) {
        const int c = c0 + lane;
        bool keep = false;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        float score = 0.0f;
        if (c < p.n_class) {
            score = __fdiv_rn(expf(__fsub_rn(lg[c], m)), s);
            keep = (p.flags & HD_FLAG_CONF_GE) ? (score >= p.score_thr) : (score > p.score_thr);
            if (keep) {
                const float4 d = *reinterpret_cast<const float4*>(dl + 4 * c);
                float dx, dy, dw, dh;
                if (mul) { dx = __fmul_rn(d.x, p.wx); dy = __fmul_rn(d.y, p.wy); dw = __fmul_rn(d.z, p.ww); dh = __fmul_rn(d.w, p.wh); }
                else { dx = __fdiv_rn(d.x, p.wx); dy = __fdiv_rn(d.y, p.wy); dw = __fdiv_rn(d.z, p.ww); dh = __fdiv_rn(d.w, p.wh); }
                if (p.flags & HD_ROIHEAD_CLAMP_DWH) { dw = fminf(dw, p.clamp_dwh); dh = fminf(dh, p.clamp_dwh); }
                const float cx = __fadd_rn(__fmul_rn(dx, wa), cxa), cy = __fadd_rn(__fmul_rn(dy, ha), cya);
                const float w = __fmul_rn(expf(dw), wa), h = __fmul_rn(expf(dh), ha);
                const float hw = __fmul_rn(0.5f, w), hh = __fmul_rn(0.5f, h);
                box.x = __fsub_rn(cx, hw); box.y = __fsub_rn(cy, hh); box.z = __fadd_rn(cx, hw); box.w = __fadd_rn(cy, hh);
                // clip_boxes_to_image
                box.x = fminf(fmaxf(box.x, 0.0f), p.img_w); box.z = fminf(fmaxf(box.z, 0.0f), p.img_w);
                box.y = fminf(fmaxf(box.y, 0.0f), p.img_h); box.w = fminf(fmaxf(box.w, 0.0f), p.img_h);
                // remove_small_boxes: keep (w >= min_size) & (h >= min_size)
                if (p.min_size > -INFINITY) keep = (__fsub_rn(box.z, box.x) >= p.min_size) && (__fsub_rn(box.w, box.y) >= p.min_size);
            }
        }
        const unsigned mk = __ballot_sync(HD_FULL, keep);
        if (mk) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&p.cand_count[b], __popc(mk));
            base = __shfl_sync(HD_FULL, base, 0);
            const int slot = base + __popc(mk & hd_lanemask_lt());
            if (keep && slot < p.cap) {
                const size_t o = (size_t)b * p.cap + slot;
                p.cand_box[o] = box; p.cand_score[o] = score;
                p.cand_cls[o] = (p.flags & HD_ROIHEAD_LABEL_MINUS1) ? c - 1 : c;
                p.cand_id[o] = r * nfg + (c - 1);   // flat index of (roi, class) in the reference's reshape(-1)
            }
        }
    }
}

// count saturates at cap when more candidates than slots were found
__global__ void roi_head_clamp_count_kernel(int* cnt, int B, int cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B && cnt[i] > cap) cnt[i] = cap;
}

static int roi_head_fill(RoiHeadParams& p, const float* logits, const float* deltas, const float* rois, const int32_t* roi_count, int B, int R,
                         int n_class, const float* weights, int flags, float clamp_dwh, float img_h, float img_w, double score_thresh,
                         float min_size, float* cand_box, float* cand_score, int32_t* cand_cls, int32_t* cand_id, int32_t* cand_count,
                         int cap) {
    HD_CHECK_ARG(B >= 0 && R >= 0 && n_class >= 2, "bad shape B=%d R=%d n_class=%d (n_class counts the background class 0)", B, R, n_class);
    HD_CHECK_ARG(weights != nullptr, "weights is NULL");
    HD_CHECK_ARG(cap >= 0, "cap must be >= 0");
    p.logits = logits; p.deltas = deltas; p.rois = rois; p.roi_count = roi_count; p.B = B; p.R = R; p.n_class = n_class; p.flags = flags;
    p.wx = weights[0]; p.wy = weights[1]; p.ww = weights[2]; p.wh = weights[3];
    p.clamp_dwh = clamp_dwh; p.img_h = img_h; p.img_w = img_w;
    p.score_thr = (float)score_thresh;   // torch compares the fp32 scores with the scalar cast to fp32 (scores > score_thresh), as yolo.cu does
    p.min_size = min_size;
    p.cand_box = (float4*)cand_box; p.cand_score = cand_score; p.cand_cls = cand_cls; p.cand_id = cand_id; p.cand_count = cand_count; p.cap = cap;
    return HD_OK;
}

extern "C" HD_API int hd_roi_head_decode_filter(const float* cls_logits, const float* box_deltas, const float* rois, const int32_t* roi_count,
                                                int B, int R, int n_class, const float* weights, int flags, float clamp_dwh, float img_h,
                                                float img_w, double score_thresh, float min_size, float* cand_box, float* cand_score,
                                                int32_t* cand_cls, int32_t* cand_id, int32_t* cand_count, int cap, void* stream) {
    RoiHeadParams p;
    int rc = roi_head_fill(p, cls_logits, box_deltas, rois, roi_count, B, R, n_class, weights, flags, clamp_dwh, img_h, img_w, score_thresh, min_size,
                           cand_box, cand_score, cand_cls, cand_id, cand_count, cap);
    if (rc) return rc;
    if (B == 0) return HD_OK;
    HD_CHECK_ARG(cand_count != nullptr, "cand_count is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    HD_CUDA_CALL(cudaMemsetAsync(cand_count, 0, sizeof(int) * (size_t)B, st));
    if (R == 0) return HD_OK;
    HD_CHECK_ARG(cls_logits && box_deltas && rois && cand_box && cand_score && cand_cls && cand_id, "null pointer");
    HD_CHECK_ARG(((uintptr_t)box_deltas & 15) == 0, "box_deltas must be 16-byte aligned");
    const long long rows = (long long)B * R, blocks = (rows + 7) / 8;
    HD_CHECK_ARG(blocks < (1ll << 31), "grid too large");
    roi_head_decode_filter_kernel<<<(unsigned)blocks, 256, 0, st>>>(p);
    HD_CUDA_LAUNCH_CHECK("roi_head_decode_filter_kernel");
    if ((long long)cap < (long long)R * (n_class - 1)) {   // only a caller-chosen smaller cap can overflow
        roi_head_clamp_count_kernel<<<(B + 255) / 256, 256, 0, st>>>(cand_count, B, cap);
        HD_CUDA_LAUNCH_CHECK("roi_head_clamp_count_kernel");
    }
    return HD_OK;
}

static size_t roi_head_cand_bytes(int B, int cap, size_t* offs) {
    size_t n = (size_t)B * cap, o = 0;
    offs[0] = o; o = hd_align_up(o + n * 16, 256);
    offs[1] = o; o = hd_align_up(o + n * 4, 256);
    offs[2] = o; o = hd_align_up(o + n * 4, 256);
    offs[3] = o; o = hd_align_up(o + n * 4, 256);
    offs[4] = o; o = hd_align_up(o + (size_t)B * 4, 256);
    return o;
}

extern "C" HD_API size_t hd_roi_head_postprocess_workspace_size(int B, int R, int n_class) {
    if (B < 0 || R < 0 || n_class < 2) return 0;
    size_t offs[5];
    const long long capl = (long long)R * (n_class - 1);
    if (capl >= (1ll << 31)) return 0;
    return roi_head_cand_bytes(B, (int)capl, offs) + hd_sort_nms_workspace_size(B, (int)capl) + 512;
}

extern "C" HD_API int hd_roi_head_postprocess(const float* cls_logits, const float* box_deltas, const float* rois, const int32_t* roi_count, int B,
                                              int R, int n_class, const float* weights, int flags, float clamp_dwh, float img_h, float img_w,
                                              double score_thresh, float min_size, double nms_iou, int max_det, float* out_det,
                                              int64_t* out_idx, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
    HD_CHECK_ARG(B >= 0 && R >= 0 && n_class >= 2, "bad shape B=%d R=%d n_class=%d", B, R, n_class);
    if (B == 0) return HD_OK;
    const long long capl = (long long)R * (n_class - 1);
    HD_CHECK_ARG(capl < (1ll << 31), "R * (n_class - 1) too large");
    const int cap = (int)capl;
    size_t offs[5];
    const size_t cb = roi_head_cand_bytes(B, cap, offs);
    uintptr_t w0 = hd_align_up((uintptr_t)workspace, 256);
    if (!workspace || w0 + cb + hd_sort_nms_workspace_size(B, cap) > (uintptr_t)workspace + workspace_bytes)
        HD_FAIL(HD_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", hd_roi_head_postprocess_workspace_size(B, R, n_class), workspace_bytes);
    float* cbox = (float*)(w0 + offs[0]); float* csc = (float*)(w0 + offs[1]);
    int32_t* ccls = (int32_t*)(w0 + offs[2]); int32_t* cid = (int32_t*)(w0 + offs[3]); int32_t* ccnt = (int32_t*)(w0 + offs[4]);
    int rc = hd_roi_head_decode_filter(cls_logits, box_deltas, rois, roi_count, B, R, n_class, weights, flags, clamp_dwh, img_h, img_w, score_thresh,
                                       min_size, cbox, csc, ccls, cid, ccnt, cap, stream);
    if (rc) return rc;
    void* nws = (void*)(w0 + cb);
    return hd_sort_nms_batched(cbox, csc, ccls, cid, ccnt, 0, B, cap, nms_iou, HD_NMS_CLASS_EXACT, 0.0f, 0, max_det, out_det, out_idx, out_count, nws,
                               (size_t)((uintptr_t)workspace + workspace_bytes - (uintptr_t)nws), stream);
}

// ------------------------------------------------------------------------------------------------ output formats (8f-2)
// Letterbox inverse + clip of padded detections (ultralytics scale_coords / clip_coords): per image
//   x = clamp((x - pad_x) / gain, 0, w0),  y = clamp((y - pad_y) / gain, 0, h0)      (fp32, IEEE sub / div)
// and optionally the COCO result-json box (x, y, w, h).  Rows >= count[b] are written as zeros.
__global__ void __launch_bounds__(256) scale_detections_kernel(const float* __restrict__ det, const int* __restrict__ count, int B, int max_det,
                                                               const float* __restrict__ meta, int flags, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * max_det) return;
    const int b = (int)(i / max_det), r = (int)(i - (long long)b * max_det);
    float* o = out + i * 6;
    if (count && r >= count[b]) { o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.0f; return; }
    const float* d = det + i * 6;
    const float px = meta[b * 5], py = meta[b * 5 + 1], gain = meta[b * 5 + 2], w0 = meta[b * 5 + 3], h0 = meta[b * 5 + 4];
    float x1 = __fdiv_rn(__fsub_rn(d[0], px), gain), y1 = __fdiv_rn(__fsub_rn(d[1], py), gain);
    float x2 = __fdiv_rn(__fsub_rn(d[2], px), gain), y2 = __fdiv_rn(__fsub_rn(d[3], py), gain);
    x1 = fminf(fmaxf(x1, 0.0f), w0); x2 = fminf(fmaxf(x2, 0.0f), w0);
    y1 = fminf(fmaxf(y1, 0.0f), h0); y2 = fminf(fmaxf(y2, 0.0f), h0);
    o[0] = x1; o[1] = y1;
    if (flags & HD_BOX_XYWH) { o[2] = __fsub_rn(x2, x1); o[3] = __fsub_rn(y2, y1); }
    else { o[2] = x2; o[3] = y2; }
    o[4] = d[4]; o[5] = d[5];
}

extern "C" HD_API int hd_scale_detections(const float* det, const int32_t* count, int B, int max_det, const float* meta, int flags, float* out,
                                          void* stream) {
    HD_CHECK_ARG(B >= 0 && max_det >= 0, "bad shape B=%d max_det=%d", B, max_det);
    if (B == 0 || max_det == 0) return HD_OK;
    HD_CHECK_ARG(det && meta && out, "null pointer");
    const long long n = (long long)B * max_det, blocks = (n + 255) / 256;
    HD_CHECK_ARG(blocks < (1ll << 31), "grid too large");
    scale_detections_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(det, count, B, max_det, meta, flags, out);
    HD_CUDA_LAUNCH_CHECK("scale_detections_kernel");
    return HD_OK;
}

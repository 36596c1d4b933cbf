// Per-level RPN selection glue (torchvision RegionProposalNetwork.filter_proposals, models/detection/rpn.py:231-297): after the
// per-level top-k (hd_rpn_select_nms with an IoU threshold nothing exceeds), (1) gather the selected boxes of all levels of an
// image, drop the small / low-score ones and compact the rest in (level, score) order -- the order batched_nms sees them in --
// and (2) after the class(=level)-aware NMS, turn the kept slots into roi rows and flat anchor indices.  Each is one launch;
// they replace ~15 eager gather / argsort / where launches of the round-1 host composition.
#include "hd_common.cuh"

struct RpnMergeParams {
    const float4* boxes;    // [B, N]
    const float* scores;    // [B, N] probabilities
    const long long* sel[HD_MAX_LEVELS];   // per level [B, k_l]: index inside the level's slice, -1 = empty
    int k[HD_MAX_LEVELS], level_off[HD_MAX_LEVELS];
    int n_levels, B, N, K;  // K = sum k_l
    float min_size, score_thresh;
    float4* cand_box; float* cand_score; int* cand_lvl; int* cand_anchor; int* cand_count;   // [B, K] / [B]
};

__global__ void __launch_bounds__(1024) rpn_merge_levels_kernel(const __grid_constant__ RpnMergeParams p) {
    __shared__ int wsum[32];
    __shared__ int carry;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < p.K; base += 1024) {
        const int j = base + tid;
        bool valid = false;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        float sc = 0.f;
        int lvl = 0, anchor = -1;
        if (j < p.K) {
            int r = j;
#pragma unroll
            for (int l = 0; l < HD_MAX_LEVELS; ++l)
                if (l < p.n_levels - 1 && r >= p.k[l] && lvl == l) { r -= p.k[l]; lvl = l + 1; }
            const long long s = p.sel[lvl][(size_t)b * p.k[lvl] + r];
            if (s >= 0) {
                anchor = p.level_off[lvl] + (int)s;
                bx = p.boxes[(size_t)b * p.N + anchor];
                sc = p.scores[(size_t)b * p.N + anchor];
                // remove_small_boxes (ws >= min_size & hs >= min_size) and scores >= score_thresh, as torchvision (rpn.py:281-287)
                valid = (__fsub_rn(bx.z, bx.x) >= p.min_size) && (__fsub_rn(bx.w, bx.y) >= p.min_size) && (sc >= p.score_thresh);
            }
        }
        const unsigned m = __ballot_sync(HD_FULL, valid);
        if (lane == 0) wsum[wid] = __popc(m);
        __syncthreads();
        int pre = carry;
        for (int w = 0; w < wid; ++w) pre += wsum[w];
        if (valid) {
            const size_t o = (size_t)b * p.K + pre + __popc(m & hd_lanemask_lt());
            p.cand_box[o] = bx; p.cand_score[o] = sc; p.cand_lvl[o] = lvl; p.cand_anchor[o] = anchor;
        }
        __syncthreads();
        if (tid == 1023) carry = pre + __popc(m);
        __syncthreads();
    }
    if (tid == 0) p.cand_count[b] = carry;
}

__global__ void __launch_bounds__(256) rpn_finish_levels_kernel(const float* __restrict__ det, const long long* __restrict__ slot,
                                                                const int* __restrict__ count, const int* __restrict__ cand_anchor, int B, int K,
                                                                int n_post, float* __restrict__ rois, float* __restrict__ scores,
                                                                long long* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * n_post) return;
    const int b = i / n_post, j = i - b * n_post;
    float* r = rois + (size_t)i * 5;
    r[0] = (float)b;
    if (j < count[b]) {
        const float* d = det + (size_t)i * 6;
        r[1] = d[0]; r[2] = d[1]; r[3] = d[2]; r[4] = d[3];
        if (scores) scores[i] = d[4];
        if (idx) idx[i] = (long long)cand_anchor[(size_t)b * K + (int)slot[i]];
    } else {
        r[1] = r[2] = r[3] = r[4] = 0.0f;
        if (scores) scores[i] = 0.0f;
        if (idx) idx[i] = -1;
    }
}

extern "C" HD_API int hd_rpn_merge_levels(const float* boxes, const float* scores, const int64_t* const* sel /*host array of device pointers*/,
                                          const int32_t* k /*host*/, const int32_t* level_off /*host*/, int n_levels, int B, int N,
                                          float min_size, float score_thresh, float* cand_box, float* cand_score, int32_t* cand_lvl,
                                          int32_t* cand_anchor, int32_t* cand_count, void* stream) {
    HD_CHECK_ARG(n_levels >= 1 && n_levels <= HD_MAX_LEVELS && sel && k && level_off, "n_levels must be in [1,%d] and the level tables non-NULL", HD_MAX_LEVELS);
    HD_CHECK_ARG(B >= 0 && N >= 0, "bad shape B=%d N=%d", B, N);
    if (B == 0) return HD_OK;
    HD_CHECK_ARG(boxes && scores && cand_box && cand_score && cand_lvl && cand_anchor && cand_count, "null pointer");
    RpnMergeParams p;
    memset(&p, 0, sizeof(p));
    p.boxes = (const float4*)boxes; p.scores = scores; p.n_levels = n_levels; p.B = B; p.N = N;
    for (int l = 0; l < n_levels; ++l) {
        HD_CHECK_ARG(k[l] >= 0 && (k[l] == 0 || sel[l] != nullptr), "level %d: bad k or NULL selection", l);
        p.sel[l] = (const long long*)sel[l]; p.k[l] = k[l]; p.level_off[l] = level_off[l]; p.K += k[l];
    }
    p.min_size = min_size; p.score_thresh = score_thresh;
    p.cand_box = (float4*)cand_box; p.cand_score = cand_score; p.cand_lvl = cand_lvl; p.cand_anchor = cand_anchor; p.cand_count = cand_count;
    rpn_merge_levels_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(p);
    HD_CUDA_LAUNCH_CHECK("rpn_merge_levels_kernel");
    return HD_OK;
}

extern "C" HD_API int hd_rpn_finish_levels(const float* det, const int64_t* slot, const int32_t* count, const int32_t* cand_anchor, int B, int K,
                                           int n_post, float* rois, float* scores, int64_t* idx, void* stream) {
    HD_CHECK_ARG(B >= 0 && K >= 0 && n_post > 0, "bad shape");
    if (B == 0) return HD_OK;
    HD_CHECK_ARG(det && slot && count && cand_anchor && rois, "null pointer");
    const int n = B * n_post;
    rpn_finish_levels_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(det, (const long long*)slot, count, cand_anchor, B, K, n_post, rois,
                                                                            scores, (long long*)idx);
    HD_CUDA_LAUNCH_CHECK("rpn_finish_levels_kernel");
    return HD_OK;
}

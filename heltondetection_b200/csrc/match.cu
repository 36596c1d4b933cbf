// IoU-based label assignment (the training-side user of box_iou): anchor / proposal <-> ground-truth
// matching without materialising the [G, N] IoU matrix.  Restates torchvision det_utils.Matcher
// (models/detection/_utils.py:318-400: max over GTs, below-low -> -1, between -> -2, allow_low_quality_matches restores
// every prediction that attains a GT's best IoU) on top of box_iou (boxes.py:308-370); the lineage AnchorTargetCreator /
// ProposalTargetCreator label rule (bubbliiiing frcnn utils_fit) is the same matching read through anchor_labels().
#include "hd_common.cuh"

#define MATCH_GT_TILE 256

__device__ __forceinline__ float match_iou(const float4 g, float ag, const float4 a, float aa) {
    // torchvision box_iou: lt = max, rb = min, wh = (rb - lt).clamp(min=0), inter / (area1 + area2 - inter)   [GT is boxes1]
    const float w = fmaxf(__fsub_rn(fminf(g.z, a.z), fmaxf(g.x, a.x)), 0.0f);
    const float h = fmaxf(__fsub_rn(fminf(g.w, a.w), fmaxf(g.y, a.y)), 0.0f);
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ag, aa), inter));
}

// pass 1: per prediction the best GT (first index on ties, as torch.max) and per GT the best IoU over all predictions
__global__ void __launch_bounds__(256) match_pass1_kernel(const float4* __restrict__ gt, const int* __restrict__ gt_count, int Gmax,
                                                          const float4* __restrict__ pred, long long pred_stride, int N,
                                                          float* __restrict__ best_iou, int* __restrict__ best_gt, int* __restrict__ gt_best_bits) {
    __shared__ float4 sg[MATCH_GT_TILE];
    __shared__ float sa[MATCH_GT_TILE];
    __shared__ int smax[MATCH_GT_TILE];
    const int b = blockIdx.y;
    const int G = gt_count ? min(gt_count[b], Gmax) : Gmax;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < N) a = pred[(size_t)b * pred_stride + i];
    const float aa = hd_area(a);
    float best = -INFINITY;
    int bj = 0;   // torch.max over an all-NaN column returns index 0 too
    for (int g0 = 0; g0 < G; g0 += MATCH_GT_TILE) {
        const int m = min(MATCH_GT_TILE, G - g0);
        __syncthreads();
        if (threadIdx.x < m) {
            const float4 g = gt[(size_t)b * Gmax + g0 + threadIdx.x];
            sg[threadIdx.x] = g; sa[threadIdx.x] = hd_area(g); smax[threadIdx.x] = 0;
        }
        __syncthreads();
        if (i < N) {
            for (int j = 0; j < m; ++j) {
                const float v = match_iou(sg[j], sa[j], a, aa);
                if (v > best || (v != v && best == best)) { best = v; bj = g0 + j; }   // NaN propagates like torch.max
                // IoU >= 0 (or NaN, whose bits compare above every finite value as in torch.max): order by the bit pattern
                const int bits = __float_as_int(v);
                if (bits > smax[j]) atomicMax(&smax[j], bits);
            }
        }
        __syncthreads();
        if (threadIdx.x < m && smax[threadIdx.x] != 0) atomicMax(&gt_best_bits[(size_t)b * Gmax + g0 + threadIdx.x], smax[threadIdx.x]);
    }
    if (i < N) { best_iou[(size_t)b * N + i] = best; best_gt[(size_t)b * N + i] = bj; }
}

// pass 2: thresholds + (optionally) low-quality restore
__global__ void __launch_bounds__(256) match_pass2_kernel(const float4* __restrict__ gt, const int* __restrict__ gt_count, int Gmax,
                                                          const float4* __restrict__ pred, long long pred_stride, int N,
                                                          const float* __restrict__ best_iou, const int* __restrict__ best_gt,
                                                          const int* __restrict__ gt_best_bits, float high, float low, int allow_low,
                                                          long long* __restrict__ matches) {
    __shared__ float4 sg[MATCH_GT_TILE];
    __shared__ float sa[MATCH_GT_TILE];
    __shared__ int smax[MATCH_GT_TILE];
    const int b = blockIdx.y;
    const int G = gt_count ? min(gt_count[b], Gmax) : Gmax;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    float v0 = 0.f; int j0 = 0;
    if (i < N) { a = pred[(size_t)b * pred_stride + i]; v0 = best_iou[(size_t)b * N + i]; j0 = best_gt[(size_t)b * N + i]; }
    const float aa = hd_area(a);
    long long out = j0;
    if (v0 < low) out = -1;                          // BELOW_LOW_THRESHOLD
    else if (v0 >= low && v0 < high) out = -2;       // BETWEEN_THRESHOLDS
    // low-quality restore: a prediction that attains some GT's best IoU keeps its arg-max match (Matcher.set_low_quality_matches_)
    const bool want = allow_low && i < N && out < 0;
    bool restore = false;
    if (allow_low) {
        for (int g0 = 0; g0 < G; g0 += MATCH_GT_TILE) {        // block-uniform loop
            if (!__syncthreads_or(want && !restore)) break;
            const int m = min(MATCH_GT_TILE, G - g0);
            if (threadIdx.x < m) {
                const float4 g = gt[(size_t)b * Gmax + g0 + threadIdx.x];
                sg[threadIdx.x] = g; sa[threadIdx.x] = hd_area(g); smax[threadIdx.x] = gt_best_bits[(size_t)b * Gmax + g0 + threadIdx.x];
            }
            __syncthreads();
            if (want && !restore)
                for (int j = 0; j < m; ++j) {
                    const float v = match_iou(sg[j], sa[j], a, aa);
                    if (v == v && __float_as_int(v) == smax[j]) { restore = true; break; }
                }
        }
    }
    if (i < N) matches[(size_t)b * N + i] = restore ? (long long)j0 : out;
}

extern "C" HD_API size_t hd_match_workspace_size(int B, int Gmax, int N) {
    if (B < 0 || Gmax < 0 || N < 0) return 0;
    return hd_align_up((size_t)B * N * 4, 256) * 2 + hd_align_up((size_t)B * (Gmax > 0 ? Gmax : 1) * 4, 256) + 256;
}

extern "C" HD_API int hd_match(const float* gt_boxes, const int32_t* gt_count, int B, int Gmax, const float* pred_boxes, int pred_per_image, int N,
                               double high_threshold, double low_threshold, int allow_low_quality, int64_t* matches, float* matched_iou,
                               void* workspace, size_t workspace_bytes, void* stream) {
    HD_CHECK_ARG(B >= 0 && Gmax >= 0 && N >= 0, "bad shape B=%d Gmax=%d N=%d", B, Gmax, N);
    HD_CHECK_ARG(low_threshold <= high_threshold, "low_threshold must be <= high_threshold");
    if (B == 0 || N == 0) return HD_OK;
    HD_CHECK_ARG(Gmax > 0, "no ground-truth boxes: the reference Matcher raises for an empty target set, handle it in the caller");
    HD_CHECK_ARG(gt_boxes && pred_boxes && matches, "null pointer");
    HD_CHECK_ARG(B <= 65535, "B too large");
    uintptr_t w0 = hd_align_up((uintptr_t)workspace, 256);
    const size_t nb = hd_align_up((size_t)B * N * 4, 256), gb = hd_align_up((size_t)B * Gmax * 4, 256);
    if (!workspace || w0 + 2 * nb + gb > (uintptr_t)workspace + workspace_bytes)
        HD_FAIL(HD_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", hd_match_workspace_size(B, Gmax, N), workspace_bytes);
    float* best_iou = (float*)w0; int* best_gt = (int*)(w0 + nb); int* gt_best = (int*)(w0 + 2 * nb);
    cudaStream_t st = (cudaStream_t)stream;
    HD_CUDA_CALL(cudaMemsetAsync(gt_best, 0, gb, st));
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)B);
    const long long ps = pred_per_image ? (long long)N : 0;
    match_pass1_kernel<<<grid, 256, 0, st>>>((const float4*)gt_boxes, gt_count, Gmax, (const float4*)pred_boxes, ps, N, best_iou, best_gt, gt_best);
    HD_CUDA_LAUNCH_CHECK("match_pass1_kernel");
    // torch compares the fp32 IoU tensor with the Python threshold cast to fp32
    match_pass2_kernel<<<grid, 256, 0, st>>>((const float4*)gt_boxes, gt_count, Gmax, (const float4*)pred_boxes, ps, N, best_iou, best_gt, gt_best,
                                             (float)high_threshold, (float)low_threshold, allow_low_quality ? 1 : 0, (long long*)matches);
    HD_CUDA_LAUNCH_CHECK("match_pass2_kernel");
    if (matched_iou) HD_CUDA_CALL(cudaMemcpyAsync(matched_iou, best_iou, (size_t)B * N * 4, cudaMemcpyDeviceToDevice, st));
    return HD_OK;
}

// ------------------------------------------------------------------------------------------------ regression targets
// BoxCoder.encode_single (torchvision models/detection/_utils.py:75-110) = lineage bbox2loc with weights: the targets that
// train the RPN / RoI heads once the matching is known.  targets[b,i] = encode(gt[b, max(matches[b,i], 0)], pred[(b,) i]).
__global__ void __launch_bounds__(256) box_encode_kernel(const float4* __restrict__ gt, int Gmax, const long long* __restrict__ matches,
                                                         const float4* __restrict__ pred, long long pred_stride, int N, int B, float wx, float wy,
                                                         float ww, float wh, float4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N) return;
    const int b = (int)(i / N), k = (int)(i - (long long)b * N);
    const float4 p = pred[(size_t)b * pred_stride + k];
    long long gi = matches ? matches[i] : k;
    if (gi < 0) gi = 0;                       // torchvision clamps the -1 / -2 codes before gathering
    const float4 g = gt[(size_t)b * Gmax + gi];
    const float ew = __fsub_rn(p.z, p.x), eh = __fsub_rn(p.w, p.y);
    const float ecx = __fadd_rn(p.x, __fmul_rn(0.5f, ew)), ecy = __fadd_rn(p.y, __fmul_rn(0.5f, eh));
    const float gw = __fsub_rn(g.z, g.x), gh = __fsub_rn(g.w, g.y);
    const float gcx = __fadd_rn(g.x, __fmul_rn(0.5f, gw)), gcy = __fadd_rn(g.y, __fmul_rn(0.5f, gh));
    float4 t;
    t.x = __fdiv_rn(__fmul_rn(wx, __fsub_rn(gcx, ecx)), ew);
    t.y = __fdiv_rn(__fmul_rn(wy, __fsub_rn(gcy, ecy)), eh);
    t.z = __fmul_rn(ww, logf(__fdiv_rn(gw, ew)));
    t.w = __fmul_rn(wh, logf(__fdiv_rn(gh, eh)));
    out[i] = t;
}

extern "C" HD_API int hd_box_encode(const float* gt_boxes, int B, int Gmax, const int64_t* matches, const float* pred_boxes, int pred_per_image,
                                    int N, const float* weights, float* targets, void* stream) {
    HD_CHECK_ARG(B >= 0 && Gmax >= 0 && N >= 0, "bad shape B=%d Gmax=%d N=%d", B, Gmax, N);
    if (B == 0 || N == 0) return HD_OK;
    HD_CHECK_ARG(Gmax > 0 && gt_boxes && pred_boxes && weights && targets, "null pointer or no ground-truth boxes");
    HD_CHECK_ARG(matches != nullptr || Gmax == N, "without matches the call is elementwise: Gmax must equal N");
    const long long n = (long long)B * N, blocks = (n + 255) / 256;
    HD_CHECK_ARG(blocks < (1ll << 31), "grid too large");
    box_encode_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)gt_boxes, Gmax, (const long long*)matches, (const float4*)pred_boxes,
                                                                          pred_per_image ? (long long)N : 0, N, B, weights[0], weights[1], weights[2],
                                                                          weights[3], (float4*)targets);
    HD_CUDA_LAUNCH_CHECK("box_encode_kernel");
    return HD_OK;
}

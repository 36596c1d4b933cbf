// TTA map-back (a10) and Weighted Boxes Fusion (a11).  Reference feature: README.md:19; semantics SURVEY.md A.6
// (ZFTurbo ensemble-boxes `weighted_boxes_fusion`, conf types 'avg', 'max', 'box_and_model_avg', 'absent_model_aware_avg').
//
// One CTA per image.  The whole CTA prefilters the V views into records, orders them with the CTA radix
// sort (three stable passes: position desc, weighted score desc, label first-appearance asc -- the order
// ensemble-boxes visits them in), then each warp takes one label segment and runs the inherently sequential
// "match best cluster / re-average" loop with the cluster table in global scratch (L1-resident): lanes
// evaluate the IoUs against all clusters in parallel, a warp arg-max picks the first best, lane 0 updates.
// Arithmetic mirrors the numpy original: float64 rows and IoU, float32 accumulator for the fused box.
#include "hd_sort.cuh"

#define WBF_NT 256

struct WbfParams {
    const float* boxes;   // [B,V,M,4]
    const float* scores;  // [B,V,M]
    const float* labels;  // [B,V,M] (float, integral values)
    const int* counts;    // [B,V]
    int B, V, M, num_labels;
    double weights[16];
    double wsum, wmax;
    double iou_thr, skip_thr;
    int conf_max, allow_overflow;
    int conf_type;       // HD_WBF_AVG / MAX / BOX_AND_MODEL_AVG / ABSENT_MODEL_AWARE_AVG
    int rescale_sum;     // 'avg' without overflow: min(cluster size, sum(weights)) (ensemble-boxes <= 1.0.4) instead of min(.., len(weights))
    float* out_boxes;   // [B, cap, 4]
    double* out_scores; // [B, cap]
    float* out_labels;  // [B, cap]
    int* out_count;     // [B]
    // workspace, per-image stride cap = V*M
    int* r_label; int* r_pos; double* r_ws; double* r_w; double* r_box;  // records
    uint64_t* k0; uint64_t* k1; uint32_t* v0; uint32_t* v1;
    int* first_pos;  // [B, num_labels]
    double* c_box; double* c_score; double* c_conf; double* c_w; double* c_max; float* c_acc; int* c_cnt; int* c_label;
    int* seg_start;
    int* c_models;       // bit t: view/model t contributed a box to the cluster
};

__device__ __forceinline__ uint64_t orderable64(double d) {
    d = d + 0.0;
    uint64_t u = (uint64_t)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(WBF_NT) wbf_kernel(const __grid_constant__ WbfParams p) {
    __shared__ HdSortSmem<WBF_NT> ssm;
    __shared__ int s_n, s_nseg, s_next, s_nclu;
    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.x;
    const int cap = p.V * p.M;
    const size_t off = (size_t)b * cap;
    int* r_label = p.r_label + off; int* r_pos = p.r_pos + off;
    double* r_ws = p.r_ws + off; double* r_w = p.r_w + off; double* r_box = p.r_box + off * 4;
    uint64_t* k0 = p.k0 + off; uint64_t* k1 = p.k1 + off; uint32_t* v0 = p.v0 + off; uint32_t* v1 = p.v1 + off;
    int* first_pos = p.first_pos + (size_t)b * p.num_labels;
    double* c_box = p.c_box + off * 4; double* c_score = p.c_score + off; double* c_conf = p.c_conf + off;
    double* c_w = p.c_w + off; double* c_max = p.c_max + off; float* c_acc = p.c_acc + off * 4;
    int* c_cnt = p.c_cnt + off; int* c_label = p.c_label + off; int* seg_start = p.seg_start + off;
    int* c_models = p.c_models + off;

    if (tid == 0) { s_n = 0; s_nseg = 0; s_next = 0; s_nclu = 0; }
    for (int i = tid; i < p.num_labels; i += WBF_NT) first_pos[i] = 0x7fffffff;
    __syncthreads();
    // ---- 1. prefilter (ensemble-boxes prefilter_boxes): pos = t*M + j is the original visiting order
    for (int q = tid; q < cap; q += WBF_NT) {
        const int t = q / p.M, j = q - t * p.M;
        if (j >= p.counts[b * p.V + t]) continue;
        const size_t g = ((size_t)b * p.V + t) * p.M + j;
        const double score = (double)p.scores[g];
        if (score < p.skip_thr) continue;
        int label = (int)p.labels[g];
        if (label < 0 || label >= p.num_labels) continue;
        double x1 = (double)p.boxes[g * 4], y1 = (double)p.boxes[g * 4 + 1], x2 = (double)p.boxes[g * 4 + 2], y2 = (double)p.boxes[g * 4 + 3];
        if (x2 < x1) { double s = x1; x1 = x2; x2 = s; }
        if (y2 < y1) { double s = y1; y1 = y2; y2 = s; }
        x1 = fmin(fmax(x1, 0.0), 1.0); y1 = fmin(fmax(y1, 0.0), 1.0); x2 = fmin(fmax(x2, 0.0), 1.0); y2 = fmin(fmax(y2, 0.0), 1.0);
        if (__dmul_rn(__dsub_rn(x2, x1), __dsub_rn(y2, y1)) == 0.0) continue;
        const int slot = atomicAdd(&s_n, 1);
        r_label[slot] = label; r_pos[slot] = q;
        r_ws[slot] = __dmul_rn(score, p.weights[t]); r_w[slot] = p.weights[t];
        r_box[slot * 4] = x1; r_box[slot * 4 + 1] = y1; r_box[slot * 4 + 2] = x2; r_box[slot * 4 + 3] = y2;
        atomicMin(&first_pos[label], q);
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) {
        if (tid == 0) p.out_count[b] = 0;
        return;
    }
    // ---- 2. order: label first-appearance asc, weighted score desc, position desc (argsort(stable)[::-1])
    for (int i = tid; i < n; i += WBF_NT) { k0[i] = (uint64_t)(uint32_t)(~(uint32_t)r_pos[i]); v0[i] = (uint32_t)i; }
    __syncthreads();
    int cur = hd_cta_radix_sort<WBF_NT>(k0, v0, k1, v1, n, ssm, 0, 3);
    {
        uint64_t* kk = cur ? k1 : k0; uint32_t* vv = cur ? v1 : v0;
        for (int i = tid; i < n; i += WBF_NT) kk[i] = ~orderable64(r_ws[vv[i]]);
        __syncthreads();
        cur ^= cur ? hd_cta_radix_sort<WBF_NT>(k1, v1, k0, v0, n, ssm) : hd_cta_radix_sort<WBF_NT>(k0, v0, k1, v1, n, ssm);
    }
    {
        uint64_t* kk = cur ? k1 : k0; uint32_t* vv = cur ? v1 : v0;
        for (int i = tid; i < n; i += WBF_NT) kk[i] = (uint64_t)(uint32_t)first_pos[r_label[vv[i]]];
        __syncthreads();
        cur ^= cur ? hd_cta_radix_sort<WBF_NT>(k1, v1, k0, v0, n, ssm, 0, 3) : hd_cta_radix_sort<WBF_NT>(k0, v0, k1, v1, n, ssm, 0, 3);
    }
    const uint32_t* order = cur ? v1 : v0;
    // segment starts (any order; segments are independent)
    for (int i = tid; i < n; i += WBF_NT) {
        c_cnt[i] = 0;
        if (i == 0 || r_label[order[i]] != r_label[order[i - 1]]) seg_start[atomicAdd(&s_nseg, 1)] = i;
    }
    __syncthreads();
    // ---- 3. clustering: one warp per label segment; cluster c of a segment lives at index seg0 + c
    const int nseg = s_nseg;
    for (;;) {
        int s = 0;
        if (lane == 0) s = atomicAdd(&s_next, 1);
        s = __shfl_sync(HD_FULL, s, 0);
        if (s >= nseg) break;
        const int seg0 = seg_start[s];
        const int label = r_label[order[seg0]];
        int nclu = 0;
        for (int i = seg0; i < n; ++i) {
            const int rec = (int)order[i];
            if (r_label[rec] != label) break;
            const double bx1 = r_box[rec * 4], by1 = r_box[rec * 4 + 1], bx2 = r_box[rec * 4 + 2], by2 = r_box[rec * 4 + 3];
            const double areaB = __dmul_rn(__dsub_rn(bx2, bx1), __dsub_rn(by2, by1));
            double best = -1.0;
            int bi = 0x7fffffff;
            for (int c = lane; c < nclu; c += 32) {
                const double* cb = c_box + (size_t)(seg0 + c) * 4;
                const double xA = fmax(cb[0], bx1), yA = fmax(cb[1], by1), xB = fmin(cb[2], bx2), yB = fmin(cb[3], by2);
                const double inter = __dmul_rn(fmax(__dsub_rn(xB, xA), 0.0), fmax(__dsub_rn(yB, yA), 0.0));
                const double areaA = __dmul_rn(__dsub_rn(cb[2], cb[0]), __dsub_rn(cb[3], cb[1]));
                const double iou = __ddiv_rn(inter, __dsub_rn(__dadd_rn(areaA, areaB), inter));
                if (iou > best) { best = iou; bi = c; }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const double ob = __shfl_xor_sync(HD_FULL, best, d);
                const int oi = __shfl_xor_sync(HD_FULL, bi, d);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            const bool match = (nclu > 0) && (best > p.iou_thr);
            if (lane == 0) {
                const double ws = r_ws[rec], w = r_w[rec];
                const int model_bit = 1 << (r_pos[rec] / p.M);
                const int c = match ? bi : nclu;
                const size_t ci = (size_t)(seg0 + c);
                float* acc = c_acc + ci * 4;
                double* cb = c_box + ci * 4;
                if (!match) {
                    // new cluster: weighted row is a copy of the box; accumulators start as get_weighted_box would
                    acc[0] = (float)__dmul_rn(ws, bx1); acc[1] = (float)__dmul_rn(ws, by1);
                    acc[2] = (float)__dmul_rn(ws, bx2); acc[3] = (float)__dmul_rn(ws, by2);
                    c_conf[ci] = ws; c_w[ci] = w; c_max[ci] = ws; c_cnt[ci] = 1; c_label[ci] = label; c_models[ci] = model_bit;
                    cb[0] = bx1; cb[1] = by1; cb[2] = bx2; cb[3] = by2;
                    c_score[ci] = ws;
                } else {
                    // box[4:] += b[1]*b[4:]  (f64 product and add, stored f32);  conf += b[1];  w += b[2]
                    acc[0] = (float)__dadd_rn((double)acc[0], __dmul_rn(ws, bx1)); acc[1] = (float)__dadd_rn((double)acc[1], __dmul_rn(ws, by1));
                    acc[2] = (float)__dadd_rn((double)acc[2], __dmul_rn(ws, bx2)); acc[3] = (float)__dadd_rn((double)acc[3], __dmul_rn(ws, by2));
                    const double conf = __dadd_rn(c_conf[ci], ws);
                    c_conf[ci] = conf; c_w[ci] = __dadd_rn(c_w[ci], w);
                    c_max[ci] = fmax(c_max[ci], ws);
                    c_models[ci] |= model_bit;
                    const int cnt = ++c_cnt[ci];
                    c_score[ci] = p.conf_max ? (double)(float)c_max[ci] : (double)(float)__ddiv_rn(conf, (double)cnt);
                    cb[0] = (double)(float)__ddiv_rn((double)acc[0], conf); cb[1] = (double)(float)__ddiv_rn((double)acc[1], conf);
                    cb[2] = (double)(float)__ddiv_rn((double)acc[2], conf); cb[3] = (double)(float)__ddiv_rn((double)acc[3], conf);
                }
            }
            if (!match) ++nclu;
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- 4. confidence rescale + final order: score desc, position in the concatenated list desc
    for (int i = tid; i < n; i += WBF_NT) {
        if (c_cnt[i] > 0) {
            double sc = c_score[i];
            // weighted_boxes[i, 2]: the exact weight of a single-box cluster (the row is a copy of the box), the float32 sum otherwise
            const double w_row = c_cnt[i] == 1 ? c_w[i] : (double)(float)c_w[i];
            if (p.conf_max) sc = __ddiv_rn(sc, p.wmax);
            else if (p.conf_type == HD_WBF_BOX_AND_MODEL_AVG) {
                double uniq = 0.0;   // weights of the distinct models in the cluster, ascending model index (np.unique order)
                for (int t = 0; t < p.V; ++t) if ((c_models[i] >> t) & 1) uniq = __dadd_rn(uniq, p.weights[t]);
                sc = __ddiv_rn(__dmul_rn(sc, (double)c_cnt[i]), w_row);
                sc = __ddiv_rn(__dmul_rn(sc, uniq), p.wsum);
            } else if (p.conf_type == HD_WBF_ABSENT_MODEL_AWARE_AVG) {
                double absent = 0.0;  // weights of the models that have no box in the cluster
                for (int t = 0; t < p.V; ++t) if (!((c_models[i] >> t) & 1)) absent = __dadd_rn(absent, p.weights[t]);
                sc = __ddiv_rn(__dmul_rn(sc, (double)c_cnt[i]), __dadd_rn(w_row, absent));
            } else if (!p.allow_overflow) sc = __ddiv_rn(__dmul_rn(sc, fmin((double)c_cnt[i], p.rescale_sum ? p.wsum : (double)p.V)), p.wsum);
            else sc = __ddiv_rn(__dmul_rn(sc, (double)c_cnt[i]), p.wsum);
            c_score[i] = sc;
            const int slot = atomicAdd(&s_nclu, 1);
            k0[slot] = (uint64_t)(uint32_t)(~(uint32_t)i);
            v0[slot] = (uint32_t)i;
        }
    }
    __syncthreads();
    const int m = s_nclu;
    cur = hd_cta_radix_sort<WBF_NT>(k0, v0, k1, v1, m, ssm, 0, 3);
    {
        uint64_t* kk = cur ? k1 : k0; uint32_t* vv = cur ? v1 : v0;
        for (int i = tid; i < m; i += WBF_NT) kk[i] = ~orderable64(c_score[vv[i]]);
        __syncthreads();
        cur ^= cur ? hd_cta_radix_sort<WBF_NT>(k1, v1, k0, v0, m, ssm) : hd_cta_radix_sort<WBF_NT>(k0, v0, k1, v1, m, ssm);
    }
    const uint32_t* fo = cur ? v1 : v0;
    for (int q = tid; q < m; q += WBF_NT) {
        const int ci = (int)fo[q];
        float* ob = p.out_boxes + (off + q) * 4;
        ob[0] = (float)c_box[ci * 4]; ob[1] = (float)c_box[ci * 4 + 1]; ob[2] = (float)c_box[ci * 4 + 2]; ob[3] = (float)c_box[ci * 4 + 3];
        p.out_scores[off + q] = c_score[ci];
        p.out_labels[off + q] = (float)c_label[ci];
    }
    if (tid == 0) p.out_count[b] = m;
}

// ------------------------------------------------------------------------------------------------ TTA map-back
__global__ void tta_map_back_kernel(const float* __restrict__ det, const int* __restrict__ count, int B, int max_det, float scale,
                                    int hflip, float view_w, float img_w, float img_h, float* __restrict__ boxes, float* __restrict__ scores,
                                    float* __restrict__ labels, int* __restrict__ counts, int V, int v, int M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * max_det) return;
    const int b = i / max_det, j = i - b * max_det;
    const int n = min(count[b], min(max_det, M));
    if (j == 0) counts[b * V + v] = n;
    if (j >= n) return;
    const float* d = det + (size_t)i * 6;
    float x1 = d[0], y1 = d[1], x2 = d[2], y2 = d[3];
    if (hflip) { const float a = __fsub_rn(view_w, x2), c = __fsub_rn(view_w, x1); x1 = a; x2 = c; }
    const size_t g = ((size_t)b * V + v) * M + j;
    boxes[g * 4] = __fdiv_rn(__fdiv_rn(x1, scale), img_w); boxes[g * 4 + 1] = __fdiv_rn(__fdiv_rn(y1, scale), img_h);
    boxes[g * 4 + 2] = __fdiv_rn(__fdiv_rn(x2, scale), img_w); boxes[g * 4 + 3] = __fdiv_rn(__fdiv_rn(y2, scale), img_h);
    scores[g] = d[4];
    labels[g] = d[5];
}

// ------------------------------------------------------------------------------------------------ host
struct WbfWs { size_t o[20]; size_t total; };
static WbfWs wbf_layout(int B, int cap, int num_labels) {
    WbfWs w; size_t n = (size_t)B * cap, o = 0; int i = 0;
    auto put = [&](size_t bytes) { w.o[i++] = o; o = hd_align_up(o + bytes, 256); };
    put(n * 4); put(n * 4); put(n * 8); put(n * 8); put(n * 32);          // r_label r_pos r_ws r_w r_box
    put(n * 8); put(n * 8); put(n * 4); put(n * 4);                        // k0 k1 v0 v1
    put((size_t)B * num_labels * 4);                                       // first_pos
    put(n * 32); put(n * 8); put(n * 8); put(n * 8); put(n * 8); put(n * 16); put(n * 4); put(n * 4);  // c_*
    put(n * 4);                                                            // seg_start
    put(n * 4);                                                            // c_models
    w.total = o;
    return w;
}

extern "C" HD_API size_t hd_wbf_workspace_size(int B, int V, int M, int num_labels) {
    if (B < 0 || V < 0 || M < 0 || num_labels < 0) return 0;
    return wbf_layout(B, V * M, num_labels).total + 256;
}

extern "C" HD_API int hd_wbf(const float* boxes, const float* scores, const float* labels, const int32_t* counts, int B, int V, int M,
                             int num_labels, const double* weights /*host, nullable*/, double iou_thr, double skip_box_thr, int conf_type,
                             int allows_overflow, float* out_boxes, double* out_scores, float* out_labels, int32_t* out_count,
                             void* workspace, size_t workspace_bytes, void* stream) {
    HD_CHECK_ARG(B >= 0 && V >= 1 && V <= 16 && M >= 1, "bad shape B=%d V=%d (max 16) M=%d", B, V, M);
    HD_CHECK_ARG(num_labels >= 1, "num_labels must be >= 1");
    const int ctype = conf_type & 255;
    HD_CHECK_ARG(ctype >= HD_WBF_AVG && ctype <= HD_WBF_ABSENT_MODEL_AWARE_AVG && (conf_type & ~(255 | HD_WBF_RESCALE_SUM_WEIGHTS)) == 0,
                 "conf_type must be HD_WBF_AVG, _MAX, _BOX_AND_MODEL_AVG or _ABSENT_MODEL_AWARE_AVG (| HD_WBF_RESCALE_SUM_WEIGHTS), got %d", conf_type);
    HD_CHECK_ARG((long long)V * M < (1ll << 24), "V*M too large");
    if (B == 0) return HD_OK;
    HD_CHECK_ARG(boxes && scores && labels && counts && out_boxes && out_scores && out_labels && out_count, "null pointer");
    WbfWs w = wbf_layout(B, V * M, num_labels);
    uintptr_t w0 = hd_align_up((uintptr_t)workspace, 256);
    if (!workspace || w0 + w.total > (uintptr_t)workspace + workspace_bytes)
        HD_FAIL(HD_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total + 256, workspace_bytes);
    WbfParams p;
    memset(&p, 0, sizeof(p));
    p.boxes = boxes; p.scores = scores; p.labels = labels; p.counts = counts; p.B = B; p.V = V; p.M = M; p.num_labels = num_labels;
    p.wsum = 0; p.wmax = -INFINITY;
    for (int v = 0; v < V; ++v) {
        p.weights[v] = weights ? weights[v] : 1.0;
        p.wsum += p.weights[v];
        if (p.weights[v] > p.wmax) p.wmax = p.weights[v];
    }
    p.iou_thr = iou_thr; p.skip_thr = skip_box_thr; p.conf_max = ctype == HD_WBF_MAX; p.allow_overflow = allows_overflow;
    p.conf_type = ctype; p.rescale_sum = (conf_type & HD_WBF_RESCALE_SUM_WEIGHTS) ? 1 : 0;
    p.out_boxes = out_boxes; p.out_scores = out_scores; p.out_labels = out_labels; p.out_count = out_count;
    int i = 0;
    p.r_label = (int*)(w0 + w.o[i++]); p.r_pos = (int*)(w0 + w.o[i++]); p.r_ws = (double*)(w0 + w.o[i++]); p.r_w = (double*)(w0 + w.o[i++]);
    p.r_box = (double*)(w0 + w.o[i++]);
    p.k0 = (uint64_t*)(w0 + w.o[i++]); p.k1 = (uint64_t*)(w0 + w.o[i++]); p.v0 = (uint32_t*)(w0 + w.o[i++]); p.v1 = (uint32_t*)(w0 + w.o[i++]);
    p.first_pos = (int*)(w0 + w.o[i++]);
    p.c_box = (double*)(w0 + w.o[i++]); p.c_score = (double*)(w0 + w.o[i++]); p.c_conf = (double*)(w0 + w.o[i++]); p.c_w = (double*)(w0 + w.o[i++]);
    p.c_max = (double*)(w0 + w.o[i++]); p.c_acc = (float*)(w0 + w.o[i++]); p.c_cnt = (int*)(w0 + w.o[i++]); p.c_label = (int*)(w0 + w.o[i++]);
    p.seg_start = (int*)(w0 + w.o[i++]); p.c_models = (int*)(w0 + w.o[i++]);
    wbf_kernel<<<B, WBF_NT, 0, (cudaStream_t)stream>>>(p);
    HD_CUDA_LAUNCH_CHECK("wbf_kernel");
    return HD_OK;
}

extern "C" HD_API int hd_tta_map_back(const float* det, const int32_t* count, int B, int max_det, float scale, int hflip, float view_w,
                                      float img_w, float img_h, float* boxes, float* scores, float* labels, int32_t* counts, int V, int v, int M,
                                      void* stream) {
    HD_CHECK_ARG(B >= 0 && max_det >= 1 && V >= 1 && v >= 0 && v < V && M >= 1, "bad shape");
    if (B == 0) return HD_OK;
    HD_CHECK_ARG(det && count && boxes && scores && labels && counts, "null pointer");
    int n = B * max_det;
    tta_map_back_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(det, count, B, max_det, scale, hflip, view_w, img_w, img_h, boxes, scores,
                                                                        labels, counts, V, v, M);
    HD_CUDA_LAUNCH_CHECK("tta_map_back_kernel");
    return HD_OK;
}

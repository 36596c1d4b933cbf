/* hd_oracle.c -- plain-C restatement of the torchvision CPU kernels on the hot path (TEST INFRASTRUCTURE ONLY).
 *
 * Restates, in scalar fp32 C compiled with -ffp-contract=off (no FMA, like the x86-64 torchvision build):
 *   hdo_nms        torchvision.ops.nms CPU kernel   (boxes.py:20-48; greedy, stable descending sort, IoU > thr strict,
 *                                                    fp32 IoU compared against the DOUBLE threshold)
 *   hdo_box_iou    torchvision.ops.box_iou          (boxes.py:308-370)
 *   hdo_roi_align  torchvision roi_align forward    (roi_align.py:115-200 / C++ kernel: samples with y<-1 or y>H give 0)
 *   hdo_roi_pool   torchvision roi_pool forward     (roi_pool.py:15-53)
 * The reference's own source is not in the mount (README only): parity is pinned to these torchvision ops by
 * tests/test_oracle.py (bit-exact for nms / roi_pool, 1e-6 for roi_align).  Nothing in heltondetection_b200/ links this. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float score; int idx; int isnan; } hdo_key;

static int hdo_cmp(const void* pa, const void* pb) {
    const hdo_key* a = (const hdo_key*)pa; const hdo_key* b = (const hdo_key*)pb;
    if (a->isnan != b->isnan) return b->isnan - a->isnan;          /* NaN scores first */
    if (!a->isnan) { if (a->score > b->score) return -1; if (a->score < b->score) return 1; }
    return (a->idx > b->idx) - (a->idx < b->idx);                   /* ties: lower index first (stable) */
}

int64_t hdo_nms(const float* boxes, const float* scores, int64_t n, double thr, int64_t max_keep, int64_t* keep) {
    if (n <= 0) return 0;
    hdo_key* order = (hdo_key*)malloc(sizeof(hdo_key) * (size_t)n);
    float* area = (float*)malloc(sizeof(float) * (size_t)n);
    unsigned char* sup = (unsigned char*)calloc((size_t)n, 1);
    for (int64_t i = 0; i < n; ++i) {
        order[i].score = scores[i]; order[i].idx = (int)i; order[i].isnan = scores[i] != scores[i];
        area[i] = (boxes[4 * i + 2] - boxes[4 * i]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    }
    qsort(order, (size_t)n, sizeof(hdo_key), hdo_cmp);
    int64_t k = 0;
    for (int64_t _i = 0; _i < n; ++_i) {
        const int i = order[_i].idx;
        if (sup[i]) continue;
        keep[k++] = i;
        if (max_keep > 0 && k == max_keep) break;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3], ia = area[i];
        for (int64_t _j = _i + 1; _j < n; ++_j) {
            const int j = order[_j].idx;
            if (sup[j]) continue;
            const float xx1 = (ix1 < boxes[4 * j]) ? boxes[4 * j] : ix1;            /* std::max(ix1, x1[j]) */
            const float yy1 = (iy1 < boxes[4 * j + 1]) ? boxes[4 * j + 1] : iy1;
            const float xx2 = (boxes[4 * j + 2] < ix2) ? boxes[4 * j + 2] : ix2;    /* std::min(ix2, x2[j]) */
            const float yy2 = (boxes[4 * j + 3] < iy2) ? boxes[4 * j + 3] : iy2;
            const float dw = xx2 - xx1, dh = yy2 - yy1;
            const float w = (0.0f < dw) ? dw : 0.0f, h = (0.0f < dh) ? dh : 0.0f;
            const float inter = w * h;
            const float ovr = inter / (ia + area[j] - inter);
            if ((double)ovr > thr) sup[j] = 1;
        }
    }
    free(order); free(area); free(sup);
    return k;
}

void hdo_box_iou(const float* b1, int64_t n, const float* b2, int64_t m, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        const float a1 = (b1[4 * i + 2] - b1[4 * i]) * (b1[4 * i + 3] - b1[4 * i + 1]);
        for (int64_t j = 0; j < m; ++j) {
            const float a2 = (b2[4 * j + 2] - b2[4 * j]) * (b2[4 * j + 3] - b2[4 * j + 1]);
            float w = fminf(b1[4 * i + 2], b2[4 * j + 2]) - fmaxf(b1[4 * i], b2[4 * j]);
            float h = fminf(b1[4 * i + 3], b2[4 * j + 3]) - fmaxf(b1[4 * i + 1], b2[4 * j + 1]);
            w = w > 0.0f ? w : 0.0f; h = h > 0.0f ? h : 0.0f;
            const float inter = w * h;
            out[i * m + j] = inter / (a1 + a2 - inter);
        }
    }
}

static float hdo_bilinear(const float* f, int H, int W, float y, float x) {
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return 0.0f;
    if (y <= 0.0f) y = 0.0f;
    if (x <= 0.0f) x = 0.0f;
    int yl = (int)y, xl = (int)x, yh, xh;
    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
    const float ly = y - (float)yl, lx = x - (float)xl, hy = 1.0f - ly, hx = 1.0f - lx;
    return hy * hx * f[yl * W + xl] + hy * lx * f[yl * W + xh] + ly * hx * f[yh * W + xl] + ly * lx * f[yh * W + xh];
}

/* input [B,C,H,W], rois [K,5], out [K,C,PH,PW] */
void hdo_roi_align(const float* in, int C, int H, int W, const float* rois, int64_t K, float scale, int PH, int PW, int sr,
                   int aligned, float* out) {
    const float off = aligned ? 0.5f : 0.0f;
    for (int64_t k = 0; k < K; ++k) {
        const float* r = rois + 5 * k;
        const int b = (int)r[0];
        const float sw = r[1] * scale - off, sh = r[2] * scale - off, ew = r[3] * scale - off, eh = r[4] * scale - off;
        float rw = ew - sw, rh = eh - sh;
        if (!aligned) { rw = rw > 1.0f ? rw : 1.0f; rh = rh > 1.0f ? rh : 1.0f; }
        const float bh = rh / (float)PH, bw = rw / (float)PW;
        const int gh = sr > 0 ? sr : (int)ceilf(rh / (float)PH), gw = sr > 0 ? sr : (int)ceilf(rw / (float)PW);
        const float count = (float)((gh * gw > 1) ? gh * gw : 1);
        for (int c = 0; c < C; ++c) {
            const float* f = in + ((size_t)b * C + c) * H * W;
            for (int ph = 0; ph < PH; ++ph)
                for (int pw = 0; pw < PW; ++pw) {
                    float acc = 0.0f;
                    for (int iy = 0; iy < gh; ++iy) {
                        const float y = sh + (float)ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
                        for (int ix = 0; ix < gw; ++ix) {
                            const float x = sw + (float)pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
                            acc += hdo_bilinear(f, H, W, y, x);
                        }
                    }
                    out[(((size_t)k * C + c) * PH + ph) * PW + pw] = acc / count;
                }
        }
    }
}

void hdo_roi_pool(const float* in, int C, int H, int W, const float* rois, int64_t K, float scale, int PH, int PW, float* out) {
    for (int64_t k = 0; k < K; ++k) {
        const float* r = rois + 5 * k;
        const int b = (int)r[0];
        const int x1 = (int)roundf(r[1] * scale), y1 = (int)roundf(r[2] * scale), x2 = (int)roundf(r[3] * scale), y2 = (int)roundf(r[4] * scale);
        const int rw = (x2 - x1 + 1 > 1) ? x2 - x1 + 1 : 1, rh = (y2 - y1 + 1 > 1) ? y2 - y1 + 1 : 1;
        const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
        for (int c = 0; c < C; ++c) {
            const float* f = in + ((size_t)b * C + c) * H * W;
            for (int ph = 0; ph < PH; ++ph) {
                int hs = (int)floorf((float)ph * bh) + y1, he = (int)ceilf((float)(ph + 1) * bh) + y1;
                hs = hs < 0 ? 0 : (hs > H ? H : hs); he = he < 0 ? 0 : (he > H ? H : he);
                for (int pw = 0; pw < PW; ++pw) {
                    int ws = (int)floorf((float)pw * bw) + x1, we = (int)ceilf((float)(pw + 1) * bw) + x1;
                    ws = ws < 0 ? 0 : (ws > W ? W : ws); we = we < 0 ? 0 : (we > W ? W : we);
                    float mx = (he <= hs || we <= ws) ? 0.0f : -INFINITY;
                    for (int h = hs; h < he; ++h)
                        for (int w = ws; w < we; ++w)
                            if (f[h * W + w] > mx) mx = f[h * W + w];
                    out[(((size_t)k * C + c) * PH + ph) * PW + pw] = mx;
                }
            }
        }
    }
}

/* ---- label assignment: torchvision det_utils.Matcher over box_iou (models/detection/_utils.py:318-400, boxes.py:308-370).
 * matches[i] = argmax_g iou(g,i) (first on ties), -1 if max < low, -2 if low <= max < high; with allow_low every prediction that
 * attains some GT's best IoU keeps its arg-max match.  Thresholds are compared after a cast to float, as torch does. */
void hdo_match(const float* gt, int64_t G, const float* pred, int64_t N, double high, double low, int allow_low, int64_t* matches) {
    float* iou = (float*)malloc(sizeof(float) * (size_t)G * (size_t)N);
    hdo_box_iou(gt, G, pred, N, iou);
    const float hf = (float)high, lf = (float)low;
    int64_t* all = (int64_t*)malloc(sizeof(int64_t) * (size_t)N);
    for (int64_t i = 0; i < N; ++i) {
        float best = iou[i]; int64_t bj = 0;
        for (int64_t g = 1; g < G; ++g) if (iou[g * N + i] > best) { best = iou[g * N + i]; bj = g; }
        all[i] = bj;
        matches[i] = (best < lf) ? -1 : ((best < hf) ? -2 : bj);
    }
    if (allow_low)
        for (int64_t g = 0; g < G; ++g) {
            float mx = iou[g * N];
            for (int64_t i = 1; i < N; ++i) if (iou[g * N + i] > mx) mx = iou[g * N + i];
            for (int64_t i = 0; i < N; ++i) if (iou[g * N + i] == mx) matches[i] = all[i];
        }
    free(iou); free(all);
}

/* ---- letterbox inverse + clip (ultralytics scale_coords / clip_coords), fp32 op order: (x - pad) / gain, clamp */
void hdo_scale_coords(const float* det, int64_t n, float pad_x, float pad_y, float gain, float w0, float h0, int xywh, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        const float* d = det + 6 * i; float* o = out + 6 * i;
        float x1 = (d[0] - pad_x) / gain, y1 = (d[1] - pad_y) / gain, x2 = (d[2] - pad_x) / gain, y2 = (d[3] - pad_y) / gain;
        x1 = fminf(fmaxf(x1, 0.0f), w0); x2 = fminf(fmaxf(x2, 0.0f), w0); y1 = fminf(fmaxf(y1, 0.0f), h0); y2 = fminf(fmaxf(y2, 0.0f), h0);
        o[0] = x1; o[1] = y1; o[2] = xywh ? x2 - x1 : x2; o[3] = xywh ? y2 - y1 : y2; o[4] = d[4]; o[5] = d[5];
    }
}

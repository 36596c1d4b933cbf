/* Pure-C consumer of libhd_b200.so: no torch, no Python -- only cudart + the C ABI of include/hd_b200.h.
 * Runs hd_box_iou, hd_sort_nms_batched (small, bitonic+pruned and class-aware paths), hd_roi_align/hd_roi_pool, hd_match and
 * hd_scale_detections on seeded data and compares them with the plain-C oracle (oracle/c/hd_oracle.c).  Exit code 0 = parity. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/hd_b200.h"

int64_t hdo_nms(const float*, const float*, int64_t, double, int64_t, int64_t*);
void hdo_box_iou(const float*, int64_t, const float*, int64_t, float*);
void hdo_roi_align(const float*, int, int, int, const float*, int64_t, float, int, int, int, int, float*);
void hdo_roi_pool(const float*, int, int, int, const float*, int64_t, float, int, int, float*);
void hdo_match(const float*, int64_t, const float*, int64_t, double, double, int, int64_t*);
void hdo_scale_coords(const float*, int64_t, float, float, float, float, float, int, float*);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %s\n", cudaGetErrorString(e), #x); return 2; } } while (0)
#define HD(x) do { int rc = (x); if (rc) { printf("hd error %d: %s at %s\n", rc, hd_last_error(), #x); return 3; } } while (0)

static uint32_t rng = 12345u;
static float frand(void) { rng = rng * 1664525u + 1013904223u; return (float)(rng >> 8) / 16777216.0f; }

static void make_boxes(float* b, float* s, int n) {
    int k = n / 8 + 1;
    float* c = (float*)malloc(sizeof(float) * 4 * k);
    for (int i = 0; i < k; ++i) { c[4 * i] = frand() * 600; c[4 * i + 1] = frand() * 600; c[4 * i + 2] = frand() * 100 + 8; c[4 * i + 3] = frand() * 100 + 8; }
    for (int i = 0; i < n; ++i) {
        int p = (int)(frand() * k) % k;
        float cx = c[4 * p] + (frand() - 0.5f) * 8, cy = c[4 * p + 1] + (frand() - 0.5f) * 8;
        float w = c[4 * p + 2] * (0.9f + 0.2f * frand()), h = c[4 * p + 3] * (0.9f + 0.2f * frand());
        b[4 * i] = cx - w / 2; b[4 * i + 1] = cy - h / 2; b[4 * i + 2] = cx + w / 2; b[4 * i + 3] = cy + h / 2;
        s[i] = floorf(frand() * 64) / 64;   /* ties on purpose */
    }
    free(c);
}

static int check_nms(int n, double thr) {
    float *b = (float*)malloc(16 * n), *s = (float*)malloc(4 * n);
    make_boxes(b, s, n);
    int64_t* ref = (int64_t*)malloc(8 * n);
    int64_t kref = hdo_nms(b, s, n, thr, -1, ref);
    float *db, *ds; int64_t* didx; int32_t* dcnt; void* ws;
    size_t wsb = hd_sort_nms_workspace_size(1, n);
    CK(cudaMalloc((void**)&db, 16 * n)); CK(cudaMalloc((void**)&ds, 4 * n)); CK(cudaMalloc((void**)&didx, 8 * n));
    CK(cudaMalloc((void**)&dcnt, 4)); CK(cudaMalloc(&ws, wsb));
    CK(cudaMemcpy(db, b, 16 * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(ds, s, 4 * n, cudaMemcpyHostToDevice));
    HD(hd_sort_nms_batched(db, ds, NULL, NULL, NULL, n, 1, n, thr, HD_NMS_AGNOSTIC, 0.0f, 0, n, NULL, didx, dcnt, ws, wsb, NULL));
    int32_t cnt; int64_t* got = (int64_t*)malloc(8 * n);
    CK(cudaMemcpy(&cnt, dcnt, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(got, didx, 8 * n, cudaMemcpyDeviceToHost));
    int bad = (cnt != kref) || memcmp(got, ref, 8 * (size_t)kref);
    printf("nms n=%d thr=%.2f: kept %d (oracle %lld) %s\n", n, thr, cnt, (long long)kref, bad ? "MISMATCH" : "ok");
    cudaFree(db); cudaFree(ds); cudaFree(didx); cudaFree(dcnt); cudaFree(ws); free(b); free(s); free(ref); free(got);
    return bad;
}

int main(void) {
    int bad = 0;
    printf("hd_version %d\n", hd_version());
    /* argument errors come back as codes, not crashes */
    if (hd_box_iou(NULL, -1, NULL, 1, NULL, NULL) != HD_ERR_INVALID) { printf("expected HD_ERR_INVALID\n"); return 4; }
    bad |= check_nms(300, 0.5);      /* small-image kernel */
    bad |= check_nms(3000, 0.45);    /* bitonic sort + pruned pass */
    bad |= check_nms(12000, 0.7);    /* radix sort + pruned pass */
    {   /* box_iou */
        int n = 257, m = 301;
        float *b1 = (float*)malloc(16 * n), *s1 = (float*)malloc(4 * n), *b2 = (float*)malloc(16 * m), *s2 = (float*)malloc(4 * m);
        make_boxes(b1, s1, n); make_boxes(b2, s2, m);
        float *ref = (float*)malloc(4 * n * m), *got = (float*)malloc(4 * n * m), *d1, *d2, *dout;
        hdo_box_iou(b1, n, b2, m, ref);
        CK(cudaMalloc((void**)&d1, 16 * n)); CK(cudaMalloc((void**)&d2, 16 * m)); CK(cudaMalloc((void**)&dout, 4 * n * m));
        CK(cudaMemcpy(d1, b1, 16 * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d2, b2, 16 * m, cudaMemcpyHostToDevice));
        HD(hd_box_iou(d1, n, d2, m, dout, NULL));
        CK(cudaMemcpy(got, dout, 4 * n * m, cudaMemcpyDeviceToHost));
        int nb = 0; for (int i = 0; i < n * m; ++i) nb += !(got[i] == ref[i] || (got[i] != got[i] && ref[i] != ref[i]));
        printf("box_iou %dx%d: %d differing elements %s\n", n, m, nb, nb ? "MISMATCH" : "ok (bit-exact)");
        bad |= nb != 0;
    }
    {   /* roi_align / roi_pool, NCHW and NHWC layouts */
        const int B = 2, C = 16, H = 20, W = 24, K = 40, PH = 7, PW = 7;
        size_t nf = (size_t)B * C * H * W, no = (size_t)K * C * PH * PW;
        float *x = (float*)malloc(4 * nf), *xn = (float*)malloc(4 * nf), *rois = (float*)malloc(4 * 5 * K);
        for (size_t i = 0; i < nf; ++i) x[i] = frand() * 2 - 1;
        for (int b = 0; b < B; ++b) for (int c = 0; c < C; ++c) for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w)
            xn[(((size_t)b * H + h) * W + w) * C + c] = x[(((size_t)b * C + c) * H + h) * W + w];
        for (int k = 0; k < K; ++k) {
            float cx = frand() * 190, cy = frand() * 160, w = frand() * 120 + 4, h = frand() * 100 + 4;
            rois[5 * k] = (float)(k % B); rois[5 * k + 1] = cx - w / 2; rois[5 * k + 2] = cy - h / 2; rois[5 * k + 3] = cx + w / 2; rois[5 * k + 4] = cy + h / 2;
        }
        float *ra = (float*)malloc(4 * no), *rp = (float*)malloc(4 * no), *got = (float*)malloc(4 * no);
        hdo_roi_align(x, C, H, W, rois, K, 0.125f, PH, PW, 2, 0, ra);
        hdo_roi_pool(x, C, H, W, rois, K, 0.125f, PH, PW, rp);
        float *dx, *dxn, *dr, *dout;
        CK(cudaMalloc((void**)&dx, 4 * nf)); CK(cudaMalloc((void**)&dxn, 4 * nf)); CK(cudaMalloc((void**)&dr, 4 * 5 * K)); CK(cudaMalloc((void**)&dout, 4 * no));
        CK(cudaMemcpy(dx, x, 4 * nf, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dxn, xn, 4 * nf, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dr, rois, 4 * 5 * K, cudaMemcpyHostToDevice));
        for (int layout = 0; layout < 2; ++layout) {
            hd_roi_level lv; lv.data = layout ? dxn : dx; lv.H = H; lv.W = W; lv.spatial_scale = 0.125f;
            HD(hd_roi_align(&lv, 1, layout, C, dr, NULL, K, PH, PW, 2, 0, dout, NULL));
            CK(cudaMemcpy(got, dout, 4 * no, cudaMemcpyDeviceToHost));
            double mx = 0; for (size_t i = 0; i < no; ++i) { double d = fabs((double)got[i] - ra[i]); if (d > mx) mx = d; }
            printf("roi_align layout=%d: max abs err %.3g %s\n", layout, mx, mx > 1e-5 ? "MISMATCH" : "ok");
            bad |= mx > 1e-5;
            HD(hd_roi_pool(&lv, 1, layout, C, dr, NULL, K, PH, PW, dout, NULL, NULL));
            CK(cudaMemcpy(got, dout, 4 * no, cudaMemcpyDeviceToHost));
            int nb = memcmp(got, rp, 4 * no) != 0;
            printf("roi_pool  layout=%d: %s\n", layout, nb ? "MISMATCH" : "ok (bit-exact)");
            bad |= nb;
        }
    }
    {   /* label assignment (8f-3) and letterbox inverse (8f-2) */
        const int G = 37, N = 3000;
        float *gt = (float*)malloc(16 * G), *gs = (float*)malloc(4 * G), *pr = (float*)malloc(16 * N), *ps = (float*)malloc(4 * N);
        make_boxes(gt, gs, G); make_boxes(pr, ps, N);
        for (int i = 0; i < G / 2; ++i) memcpy(pr + 4 * i, gt + 4 * i, 16);   /* exact hits */
        int64_t *ref = (int64_t*)malloc(8 * N), *got = (int64_t*)malloc(8 * N);
        float *dg, *dp; int64_t* dm; void* ws;
        size_t wsb = hd_match_workspace_size(1, G, N);
        CK(cudaMalloc((void**)&dg, 16 * G)); CK(cudaMalloc((void**)&dp, 16 * N)); CK(cudaMalloc((void**)&dm, 8 * N)); CK(cudaMalloc(&ws, wsb));
        CK(cudaMemcpy(dg, gt, 16 * G, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dp, pr, 16 * N, cudaMemcpyHostToDevice));
        for (int allow = 0; allow < 2; ++allow) {
            hdo_match(gt, G, pr, N, 0.7, 0.3, allow, ref);
            HD(hd_match(dg, NULL, 1, G, dp, 0, N, 0.7, 0.3, allow, dm, NULL, ws, wsb, NULL));
            CK(cudaMemcpy(got, dm, 8 * N, cudaMemcpyDeviceToHost));
            int nb = memcmp(got, ref, 8 * (size_t)N) != 0;
            printf("match allow_low=%d: %s\n", allow, nb ? "MISMATCH" : "ok (bit-exact)");
            bad |= nb;
        }
        const int M = 200;
        float *det = (float*)malloc(24 * M), *rs = (float*)malloc(24 * M), *gsc = (float*)malloc(24 * M), *dd, *dout, *dmeta;
        for (int i = 0; i < M; ++i) { det[6 * i] = frand() * 600; det[6 * i + 1] = frand() * 600; det[6 * i + 2] = det[6 * i] + frand() * 200; det[6 * i + 3] = det[6 * i + 1] + frand() * 200; det[6 * i + 4] = frand(); det[6 * i + 5] = (float)(i % 7); }
        const float gain = 640.0f / 1920.0f, meta[5] = {0.0f, (640.0f - 1080.0f * gain) / 2, gain, 1920.0f, 1080.0f};
        CK(cudaMalloc((void**)&dd, 24 * M)); CK(cudaMalloc((void**)&dout, 24 * M)); CK(cudaMalloc((void**)&dmeta, 20));
        CK(cudaMemcpy(dd, det, 24 * M, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dmeta, meta, 20, cudaMemcpyHostToDevice));
        for (int xywh = 0; xywh < 2; ++xywh) {
            hdo_scale_coords(det, M, meta[0], meta[1], meta[2], meta[3], meta[4], xywh, rs);
            HD(hd_scale_detections(dd, NULL, 1, M, dmeta, xywh ? HD_BOX_XYWH : 0, dout, NULL));
            CK(cudaMemcpy(gsc, dout, 24 * M, cudaMemcpyDeviceToHost));
            int nb = memcmp(gsc, rs, 24 * (size_t)M) != 0;
            printf("scale_detections xywh=%d: %s\n", xywh, nb ? "MISMATCH" : "ok (bit-exact)");
            bad |= nb;
        }
    }
    CK(cudaDeviceSynchronize());
    printf(bad ? "FAILED\n" : "C ABI parity OK\n");
    return bad ? 1 : 0;
}

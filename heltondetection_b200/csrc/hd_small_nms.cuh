// Per-image sort + class-aware greedy NMS for images with at most HD_SMALL_N candidates, entirely in shared
// memory, for a 256-thread CTA.  One global round trip brings the candidates in (all loads independent), an
// enumeration sort orders them by (score desc, tiebreak asc), the lazy chunked NMS runs on the sorted copy and
// the padded rows are written from shared memory.  Used (a) inline by the fused YOLO kernel for the image whose
// last tile a CTA just completed and (b) as the standalone small_nms_kernel (many images per SM at once).
#pragma once
#include "hd_nms_core.cuh"

#define HD_SMALL_N 512
#define HD_SMALL_NT 256

struct HdRep {   // device copy of hd_replicas
    int n;
    float* det[HD_MAX_REPLICAS];
    int* cnt[HD_MAX_REPLICAS];
};

struct HdNmsTail {
    float thr;  // hd_thr_floor(iou)
    int class_mode;
    float offset_scale;
    int max_nms, max_det;
    float* out_det;
    long long* out_idx;
    int* out_count;
    HdRep rep;
};

struct HdSmallSmem {
    unsigned long long key[HD_SMALL_N];
    float4 raw_box[HD_SMALL_N];
    float4 box[HD_SMALL_N];
    float score[HD_SMALL_N];
    int cls[HD_SMALL_N];
    int tie[HD_SMALL_N];
    unsigned short order[HD_SMALL_N];
    unsigned short keep[HD_SMALL_N];
    uint32_t removed[HD_SMALL_N / 32 + 4];
    HdNmsSmem nms;
};

// returns false if the image is not handled here (n > HD_SMALL_N); all 256 threads must call
__device__ __forceinline__ bool hd_small_nms_image(HdSmallSmem& sm, const HdNmsTail& q, int b, int cap, int n, const float4* cand_box,
                                                   const float* cand_score, const int* cand_cls, const int* cand_tie) {
    const int tid = threadIdx.x;
    if (n > HD_SMALL_N) return false;
    if (n <= 0) {
        if (tid == 0) q.out_count[b] = 0;
        if (tid < q.rep.n) q.rep.cnt[tid][b] = 0;
        return true;
    }
    const size_t off = (size_t)b * cap;
    for (int i = tid; i < n; i += HD_SMALL_NT) {  // one round trip: four independent L2 loads per candidate
        const float sc = __ldcg(cand_score + off + i);
        const int tb = cand_tie ? __ldcg(cand_tie + off + i) : i;
        const float4 bx = __ldcg(cand_box + off + i);
        const int c = cand_cls ? __ldcg(cand_cls + off + i) : 0;
        sm.key[i] = ((unsigned long long)(~hd_orderable(sc)) << 32) | (uint32_t)tb;
        sm.raw_box[i] = bx; sm.score[i] = sc; sm.cls[i] = c; sm.tie[i] = tb;
    }
    __syncthreads();
    // enumeration sort: keys are unique, rank = number of smaller keys
    for (int i = tid; i < n; i += HD_SMALL_NT) {
        const unsigned long long k = sm.key[i];
        int r = 0;
#pragma unroll 8
        for (int j = 0; j < n; ++j) r += (sm.key[j] < k);
        sm.order[r] = (unsigned short)i;
    }
    __syncthreads();
    const int n_use = (q.max_nms > 0) ? min(n, q.max_nms) : n;
    for (int r = tid; r < n_use; r += HD_SMALL_NT) {
        const int slot = sm.order[r];
        float4 bx = sm.raw_box[slot];
        if (q.class_mode == HD_NMS_CLASS_OFFSET) {
            const float o = __fmul_rn((float)sm.cls[slot], q.offset_scale);
            bx.x = __fadd_rn(bx.x, o); bx.y = __fadd_rn(bx.y, o); bx.z = __fadd_rn(bx.z, o); bx.w = __fadd_rn(bx.w, o);
        }
        sm.box[r] = bx;
        // class of the sorted rank, reusing the key array (sort is done)
        ((int*)sm.key)[r] = sm.cls[slot];
    }
    __syncthreads();
    const int max_det = q.max_det > 0 ? q.max_det : n_use;
    const int kc = hd_cta_greedy_nms<HD_SMALL_NT, unsigned short>(sm.box, (q.class_mode == HD_NMS_CLASS_EXACT) ? (const int*)sm.key : nullptr, n_use, max_det,
                                                  q.thr, sm.removed, sm.keep, sm.nms);
    for (int k = tid; k < kc; k += HD_SMALL_NT) {
        const int slot = sm.order[sm.keep[k]];
        if (q.out_det) {
            const float4 bx = sm.raw_box[slot];
            const float sc = sm.score[slot], cf = (float)sm.cls[slot];
            const size_t ro = ((size_t)b * q.max_det + k) * 6;
            float* o = q.out_det + ro;
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = sc; o[5] = cf;
            for (int r = 0; r < q.rep.n; ++r) {   // posted stores into the peers' gather buffers
                float* pr = q.rep.det[r] + ro;
                pr[0] = bx.x; pr[1] = bx.y; pr[2] = bx.z; pr[3] = bx.w; pr[4] = sc; pr[5] = cf;
            }
        }
        if (q.out_idx) q.out_idx[(size_t)b * q.max_det + k] = (long long)sm.tie[slot];
    }
    if (tid == 0) q.out_count[b] = kc;
    if (tid < q.rep.n) q.rep.cnt[tid][b] = kc;
    __syncthreads();
    return true;
}

// CTA-wide lazy greedy NMS over boxes already sorted by score (shared by the detection NMS and the RPN).
// Boxes are visited in chunks of 64 in score order; the chunk's 64x64 upper-triangular IoU bitmask is
// built with all threads, warp 0 resolves the chunk serially with bit operations, and only the boxes
// that were actually kept are tested against the still-alive tail.  `removed` (>= n_use/32 + 2 words of
// shared memory) holds the suppression bitmap of the whole segment.  Work is kept*n instead of n^2/2 and
// the loop stops as soon as max_det boxes are kept.  Returns the number of keeps; keep_r[q] = sorted rank.
#pragma once
#include "hd_common.cuh"

#define HD_NMS_CHUNK 64

struct HdNmsSmem {
    float4 cbox[HD_NMS_CHUNK];
    float carea[HD_NMS_CHUNK];
    int ccls[HD_NMS_CHUNK];
    unsigned long long cmask[HD_NMS_CHUNK];
    unsigned long long s_kept;
    int s_kc;
};

__device__ __forceinline__ bool hd_iou_gt(const float4& a, float area_a, const float4& b, float area_b, float thr) {
    float xx1 = hd_stdmax(a.x, b.x), yy1 = hd_stdmax(a.y, b.y);
    float xx2 = hd_stdmin(a.z, b.z), yy2 = hd_stdmin(a.w, b.w);
    float w = hd_stdmax(0.0f, __fsub_rn(xx2, xx1));
    float h = hd_stdmax(0.0f, __fsub_rn(yy2, yy1));
    if (thr >= 0.0f && !(w > 0.0f && h > 0.0f)) return false;  // inter == 0 -> iou is 0, -0 or NaN: never > thr
    float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter)) > thr;
}

// all threads of the CTA must call; blockDim.x == NT (a multiple of 64, <= 1024)
template <int NT>
__device__ int hd_cta_greedy_nms(const float4* sbox, const int* scls, int n_use, int max_det, float thr, uint32_t* removed,
                                 int* keep_r, HdNmsSmem& sm) {
    static_assert(NT % 64 == 0 && NT <= 1024 && 64 % (NT / 64) == 0, "mask build maps NT/64 threads to each of 64 rows");
    constexpr int ROWT = NT / 64;       // threads per mask row
    constexpr int COLS = 64 / ROWT;     // columns per thread
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < (n_use + 31) / 32 + 2; i += NT) removed[i] = 0;
    __syncthreads();
    int kc = 0;
    for (int base = 0; base < n_use; base += HD_NMS_CHUNK) {
        const int m = min(HD_NMS_CHUNK, n_use - base);
        if (tid < HD_NMS_CHUNK) {
            sm.cmask[tid] = 0ull;
            if (tid < m) {
                float4 bx = sbox[base + tid];
                sm.cbox[tid] = bx;
                sm.carea[tid] = hd_area(bx);
                sm.ccls[tid] = scls ? scls[base + tid] : 0;
            }
        }
        __syncthreads();
        {   // 64x64 upper-triangular mask: ROWT threads per row, COLS columns each
            const int i = tid / ROWT, j0 = (tid % ROWT) * COLS;
            if (i < m) {
                unsigned long long bits = 0ull;
                const float4 bi = sm.cbox[i];
                const float ai = sm.carea[i];
                const int ci = sm.ccls[i];
#pragma unroll 4
                for (int q = 0; q < COLS; ++q) {
                    const int jj = j0 + q;
                    if (jj > i && jj < m && sm.ccls[jj] == ci && hd_iou_gt(bi, ai, sm.cbox[jj], sm.carea[jj], thr)) bits |= 1ull << jj;
                }
                if (bits) atomicOr(&sm.cmask[i], bits);
            }
        }
        __syncthreads();
        if (wid == 0) {
            const unsigned long long rem = (unsigned long long)removed[base >> 5] | ((unsigned long long)removed[(base >> 5) + 1] << 32);
            unsigned long long alive = ~rem & ((m == 64) ? ~0ull : ((1ull << m) - 1ull));
            unsigned long long kept = 0ull;
            int room = max_det - kc;
            while (alive && room > 0) {
                const int i = __ffsll((long long)alive) - 1;
                kept |= 1ull << i;
                alive &= ~sm.cmask[i];
                alive &= ~(1ull << i);
                --room;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int bit = lane + 32 * h;
                if ((kept >> bit) & 1ull) keep_r[kc + __popcll(kept & ((1ull << bit) - 1ull))] = base + bit;
            }
            if (lane == 0) {
                sm.s_kept = kept;
                sm.s_kc = kc + __popcll(kept);
            }
        }
        __syncthreads();
        const unsigned long long kept = sm.s_kept;
        kc = sm.s_kc;
        const bool done = kc >= max_det;
        if (!done && kept) {
            for (int jr = base + HD_NMS_CHUNK + tid; jr < n_use; jr += NT) {
                if ((removed[jr >> 5] >> (jr & 31)) & 1u) continue;
                const float4 bj = sbox[jr];
                const float aj = hd_area(bj);
                const int cj = scls ? scls[jr] : 0;
                unsigned long long kk = kept;
                while (kk) {
                    const int i = __ffsll((long long)kk) - 1;
                    kk &= kk - 1ull;
                    if (sm.ccls[i] == cj && hd_iou_gt(sm.cbox[i], sm.carea[i], bj, aj, thr)) {
                        atomicOr(&removed[jr >> 5], 1u << (jr & 31));
                        break;
                    }
                }
            }
        }
        __syncthreads();
        if (done) break;
    }
    return kc;
}

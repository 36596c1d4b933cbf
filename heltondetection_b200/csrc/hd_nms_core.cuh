// CTA-wide lazy greedy NMS over boxes already sorted by score (shared by the detection NMS and the RPN).
// Boxes are visited in chunks of 64 in score order; the chunk's 64x64 upper-triangular IoU bitmask is
// built with all threads, warp 0 resolves the chunk serially with bit operations, and only the boxes
// that were actually kept are tested against the still-alive tail.  `removed` (>= n_use/32 + 2 words of
// shared memory) holds the suppression bitmap of the whole segment.  Work is kept*n instead of n^2/2 and
// the loop stops as soon as max_det boxes are kept.  Returns the number of keeps; keep_r[q] = sorted rank.
#pragma once
#include "hd_common.cuh"

#define HD_NMS_CHUNK 64
#ifndef HD_GRID_CELL
#define HD_GRID_CELL 0.5f
#endif
#define HD_GRID_MIN_N 768  // segments larger than this take the grid-pruned pass

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
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    // two-sided filter: when inter is not within 1e-5*union of thr*union the correctly rounded quotient cannot land on
    // the other side of thr (its relative error is < 1e-6), so the division is only paid for the borderline pairs
    const float d = __fsub_rn(inter, __fmul_rn(thr, uni));
    const float tol = 1.0e-5f * uni;
    if (uni > 0.0f && uni < 3.0e38f) {
        if (d > tol) return true;
        if (d < -tol) return false;
    }
    return __fdiv_rn(inter, uni) > thr;
}

// all threads of the CTA must call; blockDim.x == NT (a multiple of 64, <= 1024)
template <int NT, typename KeepT>
__device__ int hd_cta_greedy_nms(const float4* sbox, const int* scls, int n_use, int max_det, float thr, uint32_t* removed,
                                 KeepT* keep_r, HdNmsSmem& sm) {
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
                const float4 bi = sm.cbox[i];
                const float ai = sm.carea[i];
                const int ci = sm.ccls[i];
#pragma unroll 4
                for (int q = 0; q < COLS; ++q) {
                    const int jj = j0 + q;
                    // column mask: bit i of cmask[j] <=> box i (i<j) suppresses box j
                    if (jj > i && jj < m && sm.ccls[jj] == ci && hd_iou_gt(bi, ai, sm.cbox[jj], sm.carea[jj], thr)) atomicOr(&sm.cmask[jj], 1ull << i);
                }
            }
        }
        __syncthreads();
        if (wid == 0) {
            const unsigned long long rem = (unsigned long long)removed[base >> 5] | ((unsigned long long)removed[(base >> 5) + 1] << 32);
            const unsigned long long alive = ~rem & ((m == 64) ? ~0ull : ((1ull << m) - 1ull));
            // parallel resolve: kept[j] = alive[j] && !(col[j] & kept), iterated to its unique (= greedy) fixed point
            const unsigned long long c0 = sm.cmask[lane], c1 = sm.cmask[lane + 32];
            unsigned long long kept = alive;
            for (int it = 0; it < 64; ++it) {
                const bool b0 = ((alive >> lane) & 1ull) && !(c0 & kept);
                const bool b1 = ((alive >> (lane + 32)) & 1ull) && !(c1 & kept);
                const unsigned long long nk = (unsigned long long)__ballot_sync(HD_FULL, b0) | ((unsigned long long)__ballot_sync(HD_FULL, b1) << 32);
                if (nk == kept) break;
                kept = nk;
            }
            const int room = max_det - kc;
            while (__popcll(kept) > room) kept &= ~(1ull << (63 - __clzll((long long)kept)));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int bit = lane + 32 * h;
                if ((kept >> bit) & 1ull) keep_r[kc + __popcll(kept & ((1ull << bit) - 1ull))] = (KeepT)(base + bit);
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

// ------------------------------------------------------------------------------------------------------------
// Grid-pruned variant for large segments.  Same greedy result, but a kept box is only tested against the
// boxes whose centre can lie inside it:
//   iou(i,j) > t  =>  inter >= t*area_j  =>  the intersection spans >= t of j's width and height
//   =>  centre_j is inside box_i grown by max(0, 0.5 - t) * (w_j, h_j)      (t = thr - 1e-3: fp32 rounding slack)
// Proper boxes (finite, x2>x1, y2>y1) are hashed by centre cell into T buckets (counting sort in shared memory,
// item list in global scratch); improper boxes can neither suppress nor be suppressed (their IoU is 0 or NaN) and
// are left out.  Per chunk, the (kept box, covered cell) pairs are flattened over all threads.  The chunk itself
// is resolved in parallel: kept[j] = alive[j] && !(col[j] & kept) iterated to its (unique = greedy) fixed point.
// ------------------------------------------------------------------------------------------------------------
struct HdGridSmem {
    int qx1[HD_NMS_CHUNK], qy1[HD_NMS_CHUNK], qnx[HD_NMS_CHUNK], qpref[HD_NMS_CHUNK + 1];
    unsigned qmagic[HD_NMS_CHUNK];  // umulhi(local, qmagic) == local / qnx
    float red[32][6];
    float invS, dx, dy;
    int log2T;
};

__device__ __forceinline__ bool hd_box_proper(const float4& b) {
    return (b.z > b.x) && (b.w > b.y) && (fabsf(b.x) < 3.0e38f) && (fabsf(b.y) < 3.0e38f) && (fabsf(b.z) < 3.0e38f) && (fabsf(b.w) < 3.0e38f);
}
__device__ __forceinline__ int hd_cell(float v, float invS) { return __float2int_rd(v * invS); }  // saturating, monotone
__device__ __forceinline__ uint32_t hd_cell_hash(int gx, int gy, int log2T) {
    return (((uint32_t)gx * 0x9E3779B1u) ^ ((uint32_t)gy * 0x85EBCA77u)) * 0xC2B2AE3Du >> (32 - log2T);
}

template <int NT, typename KeepT>
__device__ int hd_cta_greedy_nms_grid(const float4* sbox, const int* scls, int n_use, int max_det, float thr, uint32_t* removed,
                                      KeepT* keep_r, HdNmsSmem& sm, HdGridSmem& gs, int* bucket, int log2T, float4* gitem) {
    constexpr int ROWT = NT / 64, COLS = 64 / ROWT;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int T = 1 << log2T;
    // ---- grid parameters: block reductions over the proper boxes
    float cnt = 0.f, sarea = 0.f, wmax = 0.f, hmax = 0.f, cmax = 0.f;
    for (int r = tid; r < n_use; r += NT) {
        const float4 b = sbox[r];
        if (hd_box_proper(b)) {
            const float w = b.z - b.x, h = b.w - b.y;
            cnt += 1.f; sarea += sqrtf(w * h);
            wmax = fmaxf(wmax, w); hmax = fmaxf(hmax, h);
            cmax = fmaxf(cmax, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cnt += __shfl_xor_sync(HD_FULL, cnt, d); sarea += __shfl_xor_sync(HD_FULL, sarea, d);
        wmax = fmaxf(wmax, __shfl_xor_sync(HD_FULL, wmax, d)); hmax = fmaxf(hmax, __shfl_xor_sync(HD_FULL, hmax, d));
        cmax = fmaxf(cmax, __shfl_xor_sync(HD_FULL, cmax, d));
    }
    if (lane == 0) { gs.red[wid][0] = cnt; gs.red[wid][1] = sarea; gs.red[wid][2] = wmax; gs.red[wid][3] = hmax; gs.red[wid][4] = cmax; }
    for (int i = tid; i < T + 1; i += NT) bucket[i] = 0;
    for (int i = tid; i < (n_use + 31) / 32 + 2; i += NT) removed[i] = 0;
    __syncthreads();
    if (tid == 0) {
        float c = 0.f, sa = 0.f, wm = 0.f, hm = 0.f, cm = 0.f;
        for (int w = 0; w < NT / 32; ++w) { c += gs.red[w][0]; sa += gs.red[w][1]; wm = fmaxf(wm, gs.red[w][2]); hm = fmaxf(hm, gs.red[w][3]); cm = fmaxf(cm, gs.red[w][4]); }
        float S = (c > 0.f) ? HD_GRID_CELL * sa / c : 1.0f;   // cell = HD_GRID_CELL x mean box side
        if (!(S > 0.f) || !(S < 3.0e38f)) S = 1.0f;
        const float f = fmaxf(0.0f, 0.5f - (thr - 1.0e-3f));
        gs.invS = 1.0f / S;
        gs.dx = f * wm + 5.0e-7f * cm;
        gs.dy = f * hm + 5.0e-7f * cm;
        gs.log2T = log2T;
    }
    __syncthreads();
    const float invS = gs.invS, qdx = gs.dx, qdy = gs.dy;
    // ---- counting sort of the proper boxes by centre cell
    for (int r = tid; r < n_use; r += NT) {
        const float4 b = sbox[r];
        if (hd_box_proper(b)) atomicAdd(&bucket[hd_cell_hash(hd_cell(0.5f * (b.x + b.z), invS), hd_cell(0.5f * (b.y + b.w), invS), log2T)], 1);
    }
    __syncthreads();
    {   // exclusive scan of T counts (T/NT consecutive buckets per thread)
        const int per = (T + NT - 1) / NT;
        int loc = 0;
        for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < T) loc += bucket[i]; }
        int incl = loc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(HD_FULL, incl, d); if (lane >= d) incl += y; }
        __shared__ int wsum[NT / 32];
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        int pre = incl - loc;
        for (int w = 0; w < wid; ++w) pre += wsum[w];
        for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < T) { const int c = bucket[i]; bucket[i] = pre; pre += c; } }
    }
    __syncthreads();
    for (int r = tid; r < n_use; r += NT) {
        const float4 b = sbox[r];
        if (hd_box_proper(b)) {
            const uint32_t h = hd_cell_hash(hd_cell(0.5f * (b.x + b.z), invS), hd_cell(0.5f * (b.y + b.w), invS), log2T);
            const int pos = atomicAdd(&bucket[h], 1);        // afterwards bucket[h] = end of h = start of h+1
            // {centre x, centre y, area, rank}: the pair pass streams these 16-byte records bucket by bucket
            gitem[pos] = make_float4(0.5f * (b.x + b.z), 0.5f * (b.y + b.w), hd_area(b), __int_as_float(r));
        }
    }
    __syncthreads();

    HD_PHASE(8);   // grid built
    // chunk_rank[s] = sorted rank of the s-th box of the current chunk: chunks are made of the next 64 boxes that
    // are still ALIVE, so already-suppressed boxes never cost a chunk (dense scenes: ~10x fewer chunks)
    int* chunk_rank = gs.qy1;          // reused after the gather (query rows are rebuilt every chunk)
    __shared__ int s_next, s_m, s_rank[HD_NMS_CHUNK];
    if (tid == 0) s_next = 0;
    __syncthreads();
    int kc = 0, nchunk = 0;
    for (;;) {
        // ---- warp 0 gathers the next (up to) 64 alive ranks starting at s_next
        if (wid == 0) {
            int m = 0, pos = s_next;
            while (m < HD_NMS_CHUNK && pos < n_use) {
                const int w = (pos >> 5) + lane;
                uint32_t bits = 0u;
                if (w * 32 < n_use) {
                    bits = ~removed[w];
                    if (w * 32 + 32 > n_use) bits &= (1u << (n_use - w * 32)) - 1u;
                    if (lane == 0) bits &= ~0u << (pos & 31);
                }
                const int c = __popc(bits);
                int incl = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(HD_FULL, incl, d); if (lane >= d) incl += y; }
                const int excl = incl - c, total = __shfl_sync(HD_FULL, incl, 31);
                const int room = HD_NMS_CHUNK - m;
                const int take = min(c, max(room - excl, 0));
                uint32_t bb = bits;
                int last = -1;
                for (int t = 0; t < take; ++t) { const int bp = __ffs(bb) - 1; bb &= bb - 1u; last = w * 32 + bp; s_rank[m + excl + t] = last; }
                if (total >= room) {
                    // the lane that placed the room-th box knows where the next gather starts
                    const unsigned holder = __ballot_sync(HD_FULL, take > 0 && excl + take == room);
                    const int src = __ffs(holder) - 1;
                    pos = __shfl_sync(HD_FULL, last, src) + 1;
                    m = HD_NMS_CHUNK;
                } else {
                    m += total;
                    pos = ((pos >> 5) + 32) << 5;
                }
            }
            if (lane == 0) { s_m = m; s_next = min(pos, n_use); }
        }
        __syncthreads();
        const int m = s_m;
        if (m == 0) break;
        if (nchunk++ == 16) HD_PHASE(9);
        const int last_rank = s_rank[m - 1];
        if (tid < HD_NMS_CHUNK) {
            sm.cmask[tid] = 0ull;   // column mask: bit i of cmask[j] <=> box i (i<j) suppresses box j
            if (tid < m) {
                const int r = s_rank[tid];
                const float4 bx = sbox[r];
                sm.cbox[tid] = bx; sm.carea[tid] = hd_area(bx); sm.ccls[tid] = scls ? scls[r] : 0;
            }
        }
        __syncthreads();
        {
            const int i = tid / ROWT, j0 = (tid % ROWT) * COLS;
            if (i < m) {
                const float4 bi = sm.cbox[i];
                const float ai = sm.carea[i];
                const int ci = sm.ccls[i];
#pragma unroll 4
                for (int q = 0; q < COLS; ++q) {
                    const int jj = j0 + q;
                    if (jj > i && jj < m && sm.ccls[jj] == ci && hd_iou_gt(bi, ai, sm.cbox[jj], sm.carea[jj], thr)) atomicOr(&sm.cmask[jj], 1ull << i);
                }
            }
        }
        __syncthreads();
        if (wid == 0) {
            const unsigned long long alive = (m == 64) ? ~0ull : ((1ull << m) - 1ull);
            const unsigned long long c0 = sm.cmask[lane], c1 = sm.cmask[lane + 32];
            unsigned long long kept = alive;
            for (int it = 0; it < 64; ++it) {
                const bool b0 = ((alive >> lane) & 1ull) && !(c0 & kept);
                const bool b1 = ((alive >> (lane + 32)) & 1ull) && !(c1 & kept);
                const unsigned long long nk = (unsigned long long)__ballot_sync(HD_FULL, b0) | ((unsigned long long)__ballot_sync(HD_FULL, b1) << 32);
                if (nk == kept) break;
                kept = nk;
            }
            const int room = max_det - kc;
            while (__popcll(kept) > room) kept &= ~(1ull << (63 - __clzll((long long)kept)));
            int ncell[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int bit = lane + 32 * h;
                ncell[h] = 0;
                if ((kept >> bit) & 1ull) {
                    keep_r[kc + __popcll(kept & ((1ull << bit) - 1ull))] = (KeepT)s_rank[bit];
                    const float4 b = sm.cbox[bit];
                    if (hd_box_proper(b)) {
                        const int x1 = hd_cell(b.x - qdx, invS), x2 = hd_cell(b.z + qdx, invS);
                        const int y1 = hd_cell(b.y - qdy, invS), y2 = hd_cell(b.w + qdy, invS);
                        const long long nx = (long long)x2 - x1 + 1, ny = (long long)y2 - y1 + 1;
                        if (nx * ny >= (long long)T) { gs.qnx[bit] = 0; ncell[h] = T; }      // huge box: walk every bucket
                        else { gs.qx1[bit] = x1; chunk_rank[bit] = y1; gs.qnx[bit] = (int)nx; gs.qmagic[bit] = (nx > 1) ? 0xffffffffu / (unsigned)nx + 1u : 0u; /* nx == 1: local / nx == local */ ncell[h] = (int)(nx * ny); }
                    }
                }
            }
            int i0 = ncell[0], i1 = ncell[1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y0 = __shfl_up_sync(HD_FULL, i0, d), y1 = __shfl_up_sync(HD_FULL, i1, d);
                if (lane >= d) { i0 += y0; i1 += y1; }
            }
            const int tot0 = __shfl_sync(HD_FULL, i0, 31);
            gs.qpref[lane] = i0 - ncell[0];
            gs.qpref[lane + 32] = tot0 + i1 - ncell[1];
            if (lane == 31) gs.qpref[64] = tot0 + i1;
            if (lane == 0) { sm.s_kept = kept; sm.s_kc = kc + __popcll(kept); }
        }
        __syncthreads();
        const unsigned long long kept = sm.s_kept;
        kc = sm.s_kc;
        const bool done = kc >= max_det;
        const int total = gs.qpref[64];
        if (!done && kept && last_rank + 1 < n_use) {
            int i = 0;   // slot of the current pair: t grows monotonically per thread, so the slot pointer only moves forward
            for (int t = tid; t < total; t += NT) {
                while (gs.qpref[i + 1] <= t) ++i;
                const int local = t - gs.qpref[i];
                uint32_t hb;
                if (gs.qnx[i] == 0) hb = (uint32_t)local;
                else { const int nx = gs.qnx[i]; const int gy = (nx > 1) ? (int)__umulhi((unsigned)local, gs.qmagic[i]) : local; hb = hd_cell_hash(gs.qx1[i] + (local - gy * nx), chunk_rank[i] + gy, log2T); }
                const int s0 = hb ? bucket[hb - 1] : 0, s1 = bucket[hb];
                if (s1 <= s0) continue;
                const float4 bi = sm.cbox[i];
                const float ai = sm.carea[i];
                const int ci = sm.ccls[i];
                const float lx = bi.x - qdx, hx = bi.z + qdx, ly = bi.y - qdy, hy = bi.w + qdy;
                const float tq = thr - 1.0e-3f;          // iou <= min(area)/max(area): size-mismatched pairs cannot pass
                const float amin = tq * ai, amax = (tq > 0.0f) ? ai / tq : 3.0e38f;
                auto test = [&](const float4 it) {
                    if (it.x < lx || it.x > hx || it.y < ly || it.y > hy) return;   // centre outside (also rejects hash collisions)
                    if (it.z < amin || it.z > amax) return;
                    const int jr = __float_as_int(it.w);
                    if (jr <= last_rank) return;      // earlier boxes, or members of this chunk (handled by the mask)
                    if ((removed[jr >> 5] >> (jr & 31)) & 1u) return;
                    if ((scls ? scls[jr] : 0) != ci) return;
                    const float4 bj = sbox[jr];
                    if (hd_iou_gt(bi, ai, bj, it.z, thr)) atomicOr(&removed[jr >> 5], 1u << (jr & 31));
                };
                int k = s0;
                for (; k + 4 <= s1; k += 4) {   // records are contiguous per bucket: four independent 16-byte loads in flight
                    const float4 c0 = gitem[k], c1 = gitem[k + 1], c2 = gitem[k + 2], c3 = gitem[k + 3];
                    test(c0); test(c1); test(c2); test(c3);
                }
                for (; k < s1; ++k) test(gitem[k]);
            }
        }
        __syncthreads();
        if (done) break;
    }
    return kc;
}

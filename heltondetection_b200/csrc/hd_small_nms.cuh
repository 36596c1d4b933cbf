// Per-image sort + class-aware greedy NMS for images with at most HD_SMALL_N candidates, one 256-thread CTA per image,
// everything in shared memory and registers.  Built for latency (the n ~ 150 candidates of a COCO-like image at
// conf 0.25 used to cost ~21 us per image; the sharded 32-image step of an 8-GPU run is ~45 us in total):
//   1. one global round trip: a thread owns candidates t and t+256 and keeps their payload in registers for the whole kernel;
//   2. enumeration sort on the unique 64-bit composite (score desc, tiebreak asc): rank = number of smaller keys;
//   3. the class-offset boxes are scattered to their rank; the full suppression bitmask (bit q of row r: box q < r suppresses
//      box r) is built one 32-bit word per warp step with ballots, the rows dealt round-robin to the 8 warps;
//   4. warp 0 resolves the greedy keep set 32 ranks at a time: rows are tested against the kept bits of the earlier rounds,
//      the 32x32 dependencies inside a round are iterated with ballots to their unique fixed point (= the greedy answer);
//   5. a thread whose candidate is kept writes its row from registers at position popcount(kept bits below its rank);
//      rows [count, max_det) are zero-filled (index -1) so the padded outputs are deterministic.
// Same composite, same IoU arithmetic (hd_iou_gt) and the same tie rules as the large-segment kernels: bit-identical keeps.
#pragma once
#include "hd_nms_core.cuh"

#define HD_SMALL_N 512
#define HD_SMALL_NT 256
#define HD_SMALL_W (HD_SMALL_N / 32)
#define HD_SMALL_PER (HD_SMALL_N / HD_SMALL_NT)

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
    union {   // the sort keys are dead (a barrier later) before the first mask word is written
        unsigned long long key[HD_SMALL_N];
        uint32_t mask[HD_SMALL_W][HD_SMALL_N];  // word-major: mask[w][r] = suppressors of rank r among ranks 32w..32w+31
    };
    float4 box[HD_SMALL_N];               // by rank, class offset applied
    float area[HD_SMALL_N];
    int cls[HD_SMALL_N];                  // by rank (HD_NMS_CLASS_EXACT only)
    uint32_t kept[HD_SMALL_W];
    int kc;
};

// zero rows [kc, max_det) of image b in the local outputs and in every replica (8-byte stores: a row is 24 bytes), index -1:
// the padded outputs are then a function of the inputs alone (callers reuse the buffers across calls); whole CTA calls
__device__ __forceinline__ void hd_zero_tail(float* out_det, long long* out_idx, const HdRep& rep, int max_det, int b, int kc) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lo = kc * 3, hi = max_det * 3;   // float2 units
    if (out_det) {
        float2* o = reinterpret_cast<float2*>(out_det + (size_t)b * max_det * 6);
        for (int i = lo + tid; i < hi; i += nt) o[i] = make_float2(0.f, 0.f);
        for (int r = 0; r < rep.n; ++r) {
            float2* pr = reinterpret_cast<float2*>(rep.det[r] + (size_t)b * max_det * 6);
            for (int i = lo + tid; i < hi; i += nt) pr[i] = make_float2(0.f, 0.f);
        }
    }
    if (out_idx)
        for (int i = kc + tid; i < max_det; i += nt) out_idx[(size_t)b * max_det + i] = -1;
}
__device__ __forceinline__ void hd_small_zero_tail(const HdNmsTail& q, int b, int kc) { hd_zero_tail(q.out_det, q.out_idx, q.rep, q.max_det, b, kc); }

static __device__ long long hd_dbg_small[16];   // phase clocks of block 0 (developer aid, hd_debug_phases(2, ..))
#define HD_SPHASE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) hd_dbg_small[i] = clock64(); } while (0)

// returns false if the image is not handled here (n > HD_SMALL_N); all 256 threads must call
__device__ __forceinline__ bool hd_small_nms_image(HdSmallSmem& sm, const HdNmsTail& q, int b, int cap, int n, const float4* cand_box,
                                                   const float* cand_score, const int* cand_cls, const int* cand_tie) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (n > HD_SMALL_N) return false;
    if (n <= 0) {
        if (tid == 0) q.out_count[b] = 0;
        if (tid < q.rep.n) q.rep.cnt[tid][b] = 0;
        hd_small_zero_tail(q, b, 0);
        return true;
    }
    const size_t off = (size_t)b * cap;
    HD_SPHASE(0);
    // ---- 1. load (all loads independent)
    float4 bx[HD_SMALL_PER]; float sc[HD_SMALL_PER]; int cl[HD_SMALL_PER], tb[HD_SMALL_PER];
    unsigned long long key[HD_SMALL_PER];
#pragma unroll
    for (int u = 0; u < HD_SMALL_PER; ++u) {
        const int i = tid + u * HD_SMALL_NT;
        if (i < n) {
            sc[u] = __ldcg(cand_score + off + i);
            tb[u] = cand_tie ? __ldcg(cand_tie + off + i) : i;
            bx[u] = __ldcg(cand_box + off + i);
            cl[u] = cand_cls ? __ldcg(cand_cls + off + i) : 0;
        } else { sc[u] = 0.f; tb[u] = 0; bx[u] = make_float4(0.f, 0.f, 0.f, 0.f); cl[u] = 0; }
    }
#pragma unroll
    for (int u = 0; u < HD_SMALL_PER; ++u) {
        const int i = tid + u * HD_SMALL_NT;
        key[u] = ((unsigned long long)(~hd_orderable(sc[u])) << 32) | (uint32_t)tb[u];
        if (i < n) sm.key[i] = key[u];
    }
    __syncthreads();
    HD_SPHASE(1);
    // ---- 2. enumeration sort: composites are unique, rank = number of smaller composites
    int rank[HD_SMALL_PER];
#pragma unroll
    for (int u = 0; u < HD_SMALL_PER; ++u) rank[u] = 0;
    const bool two = n > HD_SMALL_NT;   // block-uniform: does any thread own a second candidate
    if (!two) {
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0, j = 0;   // four independent counters: the adds are not one dependency chain
        for (; j + 4 <= n; j += 4) {
            c0 += (sm.key[j] < key[0]); c1 += (sm.key[j + 1] < key[0]); c2 += (sm.key[j + 2] < key[0]); c3 += (sm.key[j + 3] < key[0]);
        }
        for (; j < n; ++j) c0 += (sm.key[j] < key[0]);
        rank[0] = (c0 + c1) + (c2 + c3);
    } else {
        int r0 = 0, r1 = 0;
#pragma unroll 8
        for (int j = 0; j < n; ++j) { const unsigned long long kj = sm.key[j]; r0 += (kj < key[0]); r1 += (kj < key[1]); }
        rank[0] = r0; rank[1] = r1;
    }
    const int n_use = (q.max_nms > 0) ? min(n, q.max_nms) : n;
    HD_SPHASE(2);
    // ---- 3. scatter the class-offset boxes to their rank, then the bitmask row of every rank
    bool improper = false;
#pragma unroll
    for (int u = 0; u < HD_SMALL_PER; ++u) {
        const int i = tid + u * HD_SMALL_NT;
        float4 ob = bx[u];
        if (q.class_mode == HD_NMS_CLASS_OFFSET) {
            const float o = __fmul_rn((float)cl[u], q.offset_scale);
            ob.x = __fadd_rn(ob.x, o); ob.y = __fadd_rn(ob.y, o); ob.z = __fadd_rn(ob.z, o); ob.w = __fadd_rn(ob.w, o);
        }
        if (i < n && rank[u] < n_use) {
            sm.box[rank[u]] = ob; sm.area[rank[u]] = hd_area(ob);
            sm.cls[rank[u]] = (q.class_mode == HD_NMS_CLASS_EXACT) ? cl[u] : 0;
            improper |= !(fabsf(ob.x) < 3.0e38f && fabsf(ob.y) < 3.0e38f && fabsf(ob.z) < 3.0e38f && fabsf(ob.w) < 3.0e38f);   // NaN / inf
        }
    }
    // (barrier + vote) any NaN/inf coordinate in the image -> the generic IoU test, whose min/max follow the CPU kernel's NaN rules
    const bool generic = __syncthreads_or(improper) != 0;
    HD_SPHASE(3);
    // mask words.  Warp `wid` owns the rows r = wid (mod 8); for word w its lanes hold the 32 boxes of rank 32w..32w+31.  Per
    // (row, word): one broadcast read, a branch-free overlap test and a ballot; only words in which some pair overlaps go on to
    // the IoU test.  With finite coordinates fminf/fmaxf equal the CPU kernel's std::min/max, so the overlap extents ARE the
    // intersection sides and the IoU test is inter, union, inter - thr*union against +-1e-5*union (hd_iou_gt's two-sided
    // filter, same fp32 operations in the same order); the IEEE division only runs for borderline pairs.  Four rows per step.
    const bool exact = q.class_mode == HD_NMS_CLASS_EXACT;
    const bool pre = q.thr >= 0.0f;      // a negative threshold lets disjoint boxes (IoU 0) suppress: no pre-test then
    const int nwords = (n_use + 31) >> 5;
    for (int w = 0; w < nwords; ++w) {
        const int qr = w * 32 + lane;
        float4 qb = make_float4(0.f, 0.f, 0.f, 0.f); float qa = 0.f; int qc = 0;
        if (qr < n_use) { qb = sm.box[qr]; qa = sm.area[qr]; qc = sm.cls[qr]; }
        for (int r0 = w * 32 + wid; r0 < n_use; r0 += 4 * (HD_SMALL_NT / 32)) {
            float4 rb[4]; float ix[4], iy[4]; uint32_t m[4]; bool ov[4];
            int rr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                rr[u] = r0 + u * (HD_SMALL_NT / 32);
                rb[u] = sm.box[min(rr[u], n_use - 1)];   // clamped duplicate rows are not stored
            }
            uint32_t any = 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ix[u] = __fsub_rn(fminf(qb.z, rb[u].z), fmaxf(qb.x, rb[u].x));
                iy[u] = __fsub_rn(fminf(qb.w, rb[u].w), fmaxf(qb.y, rb[u].y));
                bool o = qr < rr[u] && rr[u] < n_use;
                if (pre && !generic) o = o && ix[u] > 0.0f && iy[u] > 0.0f;
                if (exact) o = o && (qc == sm.cls[min(rr[u], n_use - 1)]);
                ov[u] = o;
                m[u] = __ballot_sync(HD_FULL, o);
                any |= m[u];
            }
            if (any) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float ra = sm.area[min(rr[u], n_use - 1)];
                    bool hit;
                    if (generic) {
                        hit = ov[u] && hd_iou_gt(qb, qa, rb[u], ra, q.thr);
                    } else {
                        const float inter = __fmul_rn(fmaxf(ix[u], 0.0f), fmaxf(iy[u], 0.0f));
                        const float uni = __fsub_rn(__fadd_rn(qa, ra), inter);
                        const float d = __fsub_rn(inter, __fmul_rn(q.thr, uni));
                        const float tol = 1.0e-5f * uni;
                        const bool ranged = uni > 0.0f && uni < 3.0e38f;
                        hit = ov[u] && ranged && d > tol;
                        const bool border = ov[u] && !(ranged && (d > tol || d < -tol));
                        if (__any_sync(HD_FULL, border)) { if (border) hit = __fdiv_rn(inter, uni) > q.thr; }
                    }
                    m[u] = __ballot_sync(HD_FULL, hit);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (lane == u && rr[u] < n_use) sm.mask[w][rr[u]] = m[u];
        }
    }
    __syncthreads();
    HD_SPHASE(4);
    // ---- 4. warp 0: greedy keep set, 32 ranks per round
    const int max_det = q.max_det > 0 ? q.max_det : n_use;
    if (wid == 0) {
        const int rounds = (n_use + 31) >> 5;
        int total = 0;
        for (int rd = 0; rd < rounds; ++rd) {
            const int r = rd * 32 + lane;
            const bool valid = r < n_use;
            uint32_t hit = 0u;
            if (valid)
                for (int w = 0; w < rd; ++w) hit |= sm.mask[w][r] & sm.kept[w];   // kept words of the earlier rounds (broadcast reads)
            const uint32_t self = valid ? sm.mask[rd][r] : 0u;
            const bool alive = valid && hit == 0u;
            uint32_t k = __ballot_sync(HD_FULL, alive);
            for (int it = 0; it < 32; ++it) {
                const uint32_t nk = __ballot_sync(HD_FULL, alive && !(self & k));
                if (nk == k) break;
                k = nk;
            }
            if (lane == 0) sm.kept[rd] = k;
            __syncwarp();
            total += __popc(k);
            if (total >= max_det) {   // later ranks cannot be output: stop (their kept words read as zero)
                for (int w = rd + 1 + lane; w < HD_SMALL_W; w += 32) sm.kept[w] = 0u;
                break;
            }
        }
        if (lane == 0) sm.kc = min(total, max_det);
    }
    __syncthreads();
    HD_SPHASE(5);
    const int kc = sm.kc;
    // ---- 5. kept candidates write their row from registers
#pragma unroll
    for (int u = 0; u < HD_SMALL_PER; ++u) {
        const int i = tid + u * HD_SMALL_NT;
        if (i >= n || rank[u] >= n_use) continue;
        const int r = rank[u], rw = r >> 5;
        if (rw >= ((n_use + 31) >> 5)) continue;
        const uint32_t kwv = sm.kept[rw];
        if (!((kwv >> (r & 31)) & 1u)) continue;
        int pos = __popc(kwv & ((1u << (r & 31)) - 1u));
        for (int w = 0; w < rw; ++w) pos += __popc(sm.kept[w]);
        if (pos >= max_det) continue;
        if (q.out_det) {
            const float cf = (float)cl[u];
            const size_t ro = ((size_t)b * q.max_det + pos) * 6;
            float2* o = reinterpret_cast<float2*>(q.out_det + ro);
            o[0] = make_float2(bx[u].x, bx[u].y); o[1] = make_float2(bx[u].z, bx[u].w); o[2] = make_float2(sc[u], cf);
            for (int rr = 0; rr < q.rep.n; ++rr) {   // posted stores into the peers' gather buffers
                float2* pr = reinterpret_cast<float2*>(q.rep.det[rr] + ro);
                pr[0] = make_float2(bx[u].x, bx[u].y); pr[1] = make_float2(bx[u].z, bx[u].w); pr[2] = make_float2(sc[u], cf);
            }
        }
        if (q.out_idx) q.out_idx[(size_t)b * q.max_det + pos] = (long long)tb[u];
    }
    if (tid == 0) q.out_count[b] = kc;
    if (tid < q.rep.n) q.rep.cnt[tid][b] = kc;
    HD_SPHASE(6);
    hd_small_zero_tail(q, b, kc);
    HD_SPHASE(7);
    return true;
}

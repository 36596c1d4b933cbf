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
    // ---- 2. enumeration sort: composites are unique, rank = number of smaller composites
    int rank[HD_SMALL_PER];
#pragma unroll
    for (int u = 0; u < HD_SMALL_PER; ++u) rank[u] = 0;
    const bool two = n > HD_SMALL_NT;   // block-uniform: does any thread own a second candidate
    if (!two) {
        int r0 = 0;
#pragma unroll 8
        for (int j = 0; j < n; ++j) r0 += (sm.key[j] < key[0]);
        rank[0] = r0;
    } else {
        int r0 = 0, r1 = 0;
#pragma unroll 8
        for (int j = 0; j < n; ++j) { const unsigned long long kj = sm.key[j]; r0 += (kj < key[0]); r1 += (kj < key[1]); }
        rank[0] = r0; rank[1] = r1;
    }
    const int n_use = (q.max_nms > 0) ? min(n, q.max_nms) : n;
    // ---- 3. scatter the class-offset boxes to their rank, then the bitmask row of every rank
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
        }
    }
    __syncthreads();
    // mask words.  Warp `wid` owns the rows r = wid (mod 8); for word w its lanes hold the 32 boxes of rank 32w..32w+31 and
    // every row costs one broadcast read + a branch-free overlap pre-test + a ballot; the exact IoU test (hd_iou_gt) only
    // runs for the words in which some pair overlaps at all.  The pre-test is the kernel's own "inter == 0 -> not
    // suppressed" rule (same fp32 subtraction), so it only rejects pairs the exact test rejects (NaN boxes included).
    const bool exact = q.class_mode == HD_NMS_CLASS_EXACT;
    const bool pre = q.thr >= 0.0f;      // a negative threshold lets disjoint boxes (IoU 0) suppress: no pre-test then
    const int nwords = (n_use + 31) >> 5;
    for (int w = 0; w < nwords; ++w) {
        const int qr = w * 32 + lane;
        float4 qb = make_float4(0.f, 0.f, 0.f, 0.f); float qa = 0.f; int qc = 0;
        if (qr < n_use) { qb = sm.box[qr]; qa = sm.area[qr]; qc = sm.cls[qr]; }
        for (int r = w * 32 + ((wid - w * 32) & 7); r < n_use; r += HD_SMALL_NT / 32) {
            const float4 rb = sm.box[r];
            bool ov = qr < r;
            if (pre) ov = ov && (__fsub_rn(fminf(qb.z, rb.z), fmaxf(qb.x, rb.x)) > 0.0f) && (__fsub_rn(fminf(qb.w, rb.w), fmaxf(qb.y, rb.y)) > 0.0f);
            if (exact) ov = ov && (qc == sm.cls[r]);
            uint32_t m = __ballot_sync(HD_FULL, ov);
            if (m) m = __ballot_sync(HD_FULL, ov && hd_iou_gt(qb, qa, rb, sm.area[r], q.thr));
            if (lane == 0) sm.mask[w][r] = m;
        }
    }
    __syncthreads();
    // ---- 4. warp 0: greedy keep set, 32 ranks per round
    const int max_det = q.max_det > 0 ? q.max_det : n_use;
    if (wid == 0) {
        uint32_t kw[HD_SMALL_W];
#pragma unroll
        for (int w = 0; w < HD_SMALL_W; ++w) kw[w] = 0u;
        const int rounds = (n_use + 31) >> 5;
        int total = 0;
        for (int rd = 0; rd < rounds; ++rd) {
            const int r = rd * 32 + lane;
            const bool valid = r < n_use;
            bool dead = false;
#pragma unroll
            for (int w = 0; w < HD_SMALL_W; ++w)
                if (w < rd && valid) dead |= (sm.mask[w][r] & kw[w]) != 0u;
            const uint32_t self = valid ? sm.mask[rd][r] : 0u;
            const bool alive = valid && !dead;
            uint32_t k = __ballot_sync(HD_FULL, alive);
            for (int it = 0; it < 32; ++it) {
                const uint32_t nk = __ballot_sync(HD_FULL, alive && !(self & k));
                if (nk == k) break;
                k = nk;
            }
#pragma unroll
            for (int w = 0; w < HD_SMALL_W; ++w)
                if (w == rd) kw[w] = k;
            if (lane == 0) sm.kept[rd] = k;
            total += __popc(k);
            if (total >= max_det) {   // later ranks cannot be output: stop (their kept words read as zero)
                for (int w = rd + 1 + lane; w < HD_SMALL_W; w += 32) sm.kept[w] = 0u;
                break;
            }
        }
        if (lane == 0) sm.kc = min(total, max_det);
    }
    __syncthreads();
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
    hd_small_zero_tail(q, b, kc);
    return true;
}

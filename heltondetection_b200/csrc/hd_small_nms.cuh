// Per-image sort + class-aware greedy NMS for images with at most HD_SMALL_N candidates, one 512-thread CTA per image,
// everything in shared memory and registers.  Built for latency (the n ~ 150 candidates of a COCO-like image at
// conf 0.25 used to cost ~21 us per image; the sharded 32-image step of an 8-GPU run is ~45 us in total):
//   1. one global round trip: thread t owns candidate t and keeps its payload in registers for the whole kernel;
//   2. enumeration sort on the unique 64-bit composite (score desc, tiebreak asc): rank = number of smaller keys;
//   3. the class-offset boxes are scattered to their rank; the full suppression bitmask (bit q of row r: box q < r suppresses
//      box r) is built in 32x32 tiles (lane = row, the 32 lower-ranked boxes read as broadcasts, no votes, no branches),
//      the tiles dealt round-robin to the 16 warps -- 16 warps because the phase is issue/latency bound: with the 8 warps of
//      a 256-thread CTA the SM ran at IPC 0.3;
//   4. warp 0 resolves the greedy keep set 32 ranks at a time: rows are tested against the kept bits of the earlier rounds,
//      the 32x32 dependencies inside a round are iterated with ballots to their unique fixed point (= the greedy answer);
//   5. a thread whose candidate is kept writes its row from registers at position popcount(kept bits below its rank);
//      rows [count, max_det) are zero-filled (index -1) so the padded outputs are deterministic.
// Same composite, same IoU arithmetic (hd_iou_gt) and the same tie rules as the large-segment kernels: bit-identical keeps.
#pragma once
#include "hd_nms_core.cuh"

#define HD_SMALL_N 512
#define HD_SMALL_NT 512
#define HD_SMALL_W (HD_SMALL_N / 32)

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

// zero rows [kc, max_det) of image b in the local outputs (8-byte stores: a row is 24 bytes), index -1: the padded outputs are
// then a function of the inputs alone (callers reuse the buffers across calls).  The REPLICAS only receive the kept rows and the
// count -- their rows at and beyond the count are unspecified -- so that a step does not push 15x its payload over NVLink.
// Whole CTA calls.
__device__ __forceinline__ void hd_zero_tail(float* out_det, long long* out_idx, const HdRep& rep, int max_det, int b, int kc) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lo = kc * 3, hi = max_det * 3;   // float2 units
    if (out_det) {
        float2* o = reinterpret_cast<float2*>(out_det + (size_t)b * max_det * 6);
        for (int i = lo + tid; i < hi; i += nt) o[i] = make_float2(0.f, 0.f);
    }
    (void)rep;
    if (out_idx)
        for (int i = kc + tid; i < max_det; i += nt) out_idx[(size_t)b * max_det + i] = -1;
}
__device__ __forceinline__ void hd_small_zero_tail(const HdNmsTail& q, int b, int kc) { hd_zero_tail(q.out_det, q.out_idx, q.rep, q.max_det, b, kc); }

static __device__ long long hd_dbg_small[16];   // phase clocks of block 0 (developer aid, hd_debug_phases(2, ..))
#define HD_SPHASE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) hd_dbg_small[i] = clock64(); } while (0)

// returns false if the image is not handled here (n > HD_SMALL_N); all HD_SMALL_NT threads must call
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
    // ---- 1. load: thread t owns candidate t (all loads independent); the payload stays in registers until the output
    const bool mine = tid < n;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f); float sc = 0.f; int cl = 0, tb = 0;
    if (mine) {
        sc = __ldcg(cand_score + off + tid);
        tb = cand_tie ? __ldcg(cand_tie + off + tid) : tid;
        bx = __ldcg(cand_box + off + tid);
        cl = cand_cls ? __ldcg(cand_cls + off + tid) : 0;
    }
    const unsigned long long key = ((unsigned long long)(~hd_orderable(sc)) << 32) | (uint32_t)tb;
    if (mine) sm.key[tid] = key;
    __syncthreads();
    HD_SPHASE(1);
    // ---- 2. enumeration sort: composites are unique, rank = number of smaller composites
    int rank = 0;
    if (mine) {
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0, j = 0;   // four independent counters: the adds are not one dependency chain
        for (; j + 4 <= n; j += 4) {
            c0 += (sm.key[j] < key); c1 += (sm.key[j + 1] < key); c2 += (sm.key[j + 2] < key); c3 += (sm.key[j + 3] < key);
        }
        for (; j < n; ++j) c0 += (sm.key[j] < key);
        rank = (c0 + c1) + (c2 + c3);
    }
    const int n_use = (q.max_nms > 0) ? min(n, q.max_nms) : n;
    HD_SPHASE(2);
    // ---- 3. scatter the class-offset boxes to their rank
    bool improper = false;
    {
        float4 ob = bx;
        if (q.class_mode == HD_NMS_CLASS_OFFSET) {
            const float o = __fmul_rn((float)cl, q.offset_scale);
            ob.x = __fadd_rn(ob.x, o); ob.y = __fadd_rn(ob.y, o); ob.z = __fadd_rn(ob.z, o); ob.w = __fadd_rn(ob.w, o);
        }
        if (mine && rank < n_use) {
            sm.box[rank] = ob; sm.area[rank] = hd_area(ob);
            sm.cls[rank] = (q.class_mode == HD_NMS_CLASS_EXACT) ? cl : 0;
            improper = !(fabsf(ob.x) < 3.0e38f && fabsf(ob.y) < 3.0e38f && fabsf(ob.z) < 3.0e38f && fabsf(ob.w) < 3.0e38f);   // NaN / inf
        }
    }
    // (barrier + vote) any NaN/inf coordinate in the image -> the generic IoU test, whose min/max follow the CPU kernel's NaN rules
    const bool generic = __syncthreads_or(improper) != 0;
    HD_SPHASE(3);
    // ---- 4. the suppression bitmask, tile by tile: task (v, c), c <= v, gives the 32 rows 32v..32v+31 their mask word for the
    // ranks 32c..32c+31.  Lane = row; the loop walks the 32 lower-ranked boxes as shared-memory broadcasts and is branch free:
    // with finite coordinates fminf/fmaxf equal the CPU kernel's std::min/max, the overlap extents ARE the intersection sides,
    // and the test is hd_iou_gt's two-sided filter (inter - thr*union against +-1e-5*union: same fp32 operations, same order);
    // the IEEE division only runs for borderline pairs.  The G(G+1)/2 tiles are dealt round-robin to the 16 warps.
    {
        const bool exact = q.class_mode == HD_NMS_CLASS_EXACT;
        const int G = (n_use + 31) >> 5;
        int v = 0, c = 0;
        for (int t = 0; t < wid; ++t) { if (c == v) { ++v; c = 0; } else ++c; }     // task `wid` in (0,0),(1,0),(1,1),(2,0).. order
        while (v < G) {
            const int r = v * 32 + lane;
            const bool rv = r < n_use;
            const float4 me = sm.box[min(r, HD_SMALL_N - 1)];
            const float ma = sm.area[min(r, HD_SMALL_N - 1)];
            const int mc = sm.cls[min(r, HD_SMALL_N - 1)];
            uint32_t word = 0u;
            const int q0 = c * 32, q1 = min(q0 + 32, n_use);
            if (!generic) {
#pragma unroll 4
                for (int qq = q0; qq < q1; ++qq) {
                    const float4 hb = sm.box[qq];
                    const float ha = sm.area[qq];
                    const float ix = __fsub_rn(fminf(hb.z, me.z), fmaxf(hb.x, me.x));
                    const float iy = __fsub_rn(fminf(hb.w, me.w), fmaxf(hb.y, me.y));
                    const float inter = __fmul_rn(fmaxf(ix, 0.0f), fmaxf(iy, 0.0f));
                    const float uni = __fsub_rn(__fadd_rn(ha, ma), inter);
                    const float d = __fsub_rn(inter, __fmul_rn(q.thr, uni));
                    const float tol = 1.0e-5f * uni;
                    const bool ranged = uni > 0.0f && uni < 3.0e38f;
                    bool hit = ranged && d > tol;
                    if (!(ranged && (d > tol || d < -tol)))      // borderline (rare): the exact quotient, after hd_iou_gt's inter == 0 rule
                        hit = (q.thr >= 0.0f && !(ix > 0.0f && iy > 0.0f)) ? false : (__fdiv_rn(inter, uni) > q.thr);
                    hit = hit && qq < r && (!exact || sm.cls[qq] == mc);
                    word |= hit ? (1u << (qq - q0)) : 0u;
                }
            } else {
                for (int qq = q0; qq < q1; ++qq) {
                    const bool hit = qq < r && (!exact || sm.cls[qq] == mc) && hd_iou_gt(sm.box[qq], sm.area[qq], me, ma, q.thr);
                    word |= hit ? (1u << (qq - q0)) : 0u;
                }
            }
            if (rv) sm.mask[c][r] = word;
            for (int t = 0; t < HD_SMALL_NT / 32; ++t) { if (c == v) { ++v; c = 0; } else ++c; }
        }
    }
    __syncthreads();
    HD_SPHASE(4);
    // ---- 5. warp 0: greedy keep set, 32 ranks per round
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
    // ---- 6. a kept candidate writes its row from registers
    if (mine && rank < n_use) {
        const int rw = rank >> 5;
        const uint32_t kwv = sm.kept[rw];
        if ((kwv >> (rank & 31)) & 1u) {
            int pos = __popc(kwv & ((1u << (rank & 31)) - 1u));
            for (int w = 0; w < rw; ++w) pos += __popc(sm.kept[w]);
            if (pos < max_det) {
                if (q.out_det) {
                    const float cf = (float)cl;
                    const size_t ro = ((size_t)b * q.max_det + pos) * 6;
                    float2* o = reinterpret_cast<float2*>(q.out_det + ro);
                    o[0] = make_float2(bx.x, bx.y); o[1] = make_float2(bx.z, bx.w); o[2] = make_float2(sc, cf);
                    for (int rr = 0; rr < q.rep.n; ++rr) {   // posted stores into the peers' gather buffers
                        float2* pr = reinterpret_cast<float2*>(q.rep.det[rr] + ro);
                        pr[0] = make_float2(bx.x, bx.y); pr[1] = make_float2(bx.z, bx.w); pr[2] = make_float2(sc, cf);
                    }
                }
                if (q.out_idx) q.out_idx[(size_t)b * q.max_det + pos] = (long long)tb;
            }
        }
    }
    if (tid == 0) q.out_count[b] = kc;
    if (tid < q.rep.n) q.rep.cnt[tid][b] = kc;
    HD_SPHASE(6);
    hd_small_zero_tail(q, b, kc);
    HD_SPHASE(7);
    return true;
}

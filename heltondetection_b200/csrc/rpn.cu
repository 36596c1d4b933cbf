// RPN proposal creation (a6): fused anchor generation + delta decode + clip + min-size + score, then per image
// radix-select top-k -> sort -> greedy NMS -> first n_post.  Reference feature: README.md:8,63-65; semantics
// SURVEY.md A.3 (bubbliiiing ProposalCreator / loc2bbox lineage, README.md:158; cross-checked with torchvision
// models/detection/rpn.py:231-297, _utils.py:183-224).
#include "hd_sort.cuh"
#include "hd_nms_core.cuh"
#include "hd_cluster_nms.cuh"

#define RPN_NT 1024
#define RPN_U 8   // keys in flight per thread in the latency-bound scans over the N proposals of an image

struct RpnParams {
    const float* obj[HD_MAX_LEVELS];
    const float* dlt[HD_MAX_LEVELS];
    int H[HD_MAX_LEVELS], W[HD_MAX_LEVELS];
    float stride[HD_MAX_LEVELS];
    float base[HD_MAX_LEVELS][4 * HD_MAX_ANCHORS];
    int cell_start[HD_MAX_LEVELS + 1];  // cumulative cells
    int level_off[HD_MAX_LEVELS + 1];   // cumulative anchors
    int n_levels, B, A, softmax;
    float img_h, img_w, min_size, clamp_dwh;
    int use_clamp;
    int key_logit;          // sort key from the raw objectness logit (torchvision takes its per-level top-k on the logits)
    int exact;              // HD_RPN_EXACT_MATH: transcendentals in fp64, rounded once to fp32
    long long total_cells;  // B * sum(HW)
    int N;                  // anchors per image
};

// ------------------------------------------------------------------------------------------------
// decode: one thread per (image, level, cell); it walks the A anchors of its cell, so every head plane is
// read coalesced along W and the A decoded boxes of a cell leave as one contiguous A*16-byte run
// (flat proposal index = level_off + cell*A + a, the permute(0,2,3,1) order of the reference).
//   key  = orderable(score) (0 is reserved for boxes dropped by the min-size test)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rpn_decode_kernel(const __grid_constant__ RpnParams p, float4* __restrict__ boxes,
                                                         float* __restrict__ scores, uint32_t* __restrict__ keys) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.total_cells) return;
    const int cells_per_img = p.cell_start[p.n_levels];
    const int b = (int)(g / cells_per_img);
    int r = (int)(g - (long long)b * cells_per_img);
    int l = 0;
#pragma unroll
    for (int q = 1; q < HD_MAX_LEVELS; ++q)
        if (q < p.n_levels && r >= p.cell_start[q]) l = q;
    const int cell = r - p.cell_start[l];
    const int HW = p.H[l] * p.W[l], W = p.W[l];
    const int gi = cell / W, gj = cell - gi * W;
    const float sx = __fmul_rn((float)gj, p.stride[l]), sy = __fmul_rn((float)gi, p.stride[l]);
    const int oc = p.softmax ? 2 * p.A : p.A;
    const float* __restrict__ ob = p.obj[l] + (size_t)b * oc * HW + cell;
    const float* __restrict__ dl = p.dlt[l] + (size_t)b * 4 * p.A * HW + cell;
    const size_t out0 = (size_t)b * p.N + p.level_off[l] + (size_t)cell * p.A;
    for (int a = 0; a < p.A; ++a) {
        float score, logit = 0.0f;
        if (p.softmax) {  // F.softmax over (bg, fg): exp(x - max) / sum
            const float s0 = hd_ldg_stream(ob + (size_t)(2 * a) * HW), s1 = hd_ldg_stream(ob + (size_t)(2 * a + 1) * HW);
            const float m = fmaxf(s0, s1);
            if (p.exact) {
                const double e0 = exp((double)s0 - (double)m), e1 = exp((double)s1 - (double)m);
                score = (float)(e1 / (e0 + e1));
            } else {
                const float e0 = expf(__fsub_rn(s0, m)), e1 = expf(__fsub_rn(s1, m));
                score = __fdiv_rn(e1, __fadd_rn(e0, e1));
            }
        } else {
            logit = hd_ldg_stream(ob + (size_t)a * HW);
            score = p.exact ? (float)(1.0 / (1.0 + exp(-(double)logit))) : hd_sigmoid(logit);
        }
        const float dx = hd_ldg_stream(dl + (size_t)(4 * a) * HW), dy = hd_ldg_stream(dl + (size_t)(4 * a + 1) * HW);
        float dw = hd_ldg_stream(dl + (size_t)(4 * a + 2) * HW), dh = hd_ldg_stream(dl + (size_t)(4 * a + 3) * HW);
        if (p.use_clamp) { dw = fminf(dw, p.clamp_dwh); dh = fminf(dh, p.clamp_dwh); }
        // anchor = base + shift (fp32 add, as enumerate_shifted_anchor)
        const float ax1 = __fadd_rn(p.base[l][4 * a], sx), ay1 = __fadd_rn(p.base[l][4 * a + 1], sy);
        const float ax2 = __fadd_rn(p.base[l][4 * a + 2], sx), ay2 = __fadd_rn(p.base[l][4 * a + 3], sy);
        // loc2bbox
        const float wa = __fsub_rn(ax2, ax1), ha = __fsub_rn(ay2, ay1);
        const float cxa = __fadd_rn(ax1, __fmul_rn(0.5f, wa)), cya = __fadd_rn(ay1, __fmul_rn(0.5f, ha));
        const float cx = __fadd_rn(__fmul_rn(dx, wa), cxa), cy = __fadd_rn(__fmul_rn(dy, ha), cya);
        const float ew = p.exact ? (float)exp((double)dw) : expf(dw), eh = p.exact ? (float)exp((double)dh) : expf(dh);
        const float w = __fmul_rn(ew, wa), h = __fmul_rn(eh, ha);
        float x1 = __fsub_rn(cx, __fmul_rn(0.5f, w)), y1 = __fsub_rn(cy, __fmul_rn(0.5f, h));
        float x2 = __fadd_rn(cx, __fmul_rn(0.5f, w)), y2 = __fadd_rn(cy, __fmul_rn(0.5f, h));
        // clip to the image (torch.clamp(min=0, max=size))
        x1 = fminf(fmaxf(x1, 0.0f), p.img_w); x2 = fminf(fmaxf(x2, 0.0f), p.img_w);
        y1 = fminf(fmaxf(y1, 0.0f), p.img_h); y2 = fminf(fmaxf(y2, 0.0f), p.img_h);
        const bool ok = (__fsub_rn(x2, x1) >= p.min_size) && (__fsub_rn(y2, y1) >= p.min_size);
        boxes[out0 + a] = make_float4(x1, y1, x2, y2);
        scores[out0 + a] = score;
        uint32_t k = hd_orderable((p.key_logit && !p.softmax) ? logit : score);
        keys[out0 + a] = ok ? (k == 0u ? 1u : k) : 0u;
    }
}

// ------------------------------------------------------------------------------------------------
// select + sort + NMS: one CTA per image.
//   1. radix select (MSB first, 8-bit digits, warp-aggregated shared histograms) of the k-th largest
//      64-bit composite (key << 32 | ~index): composites are unique, so "composite >= T" selects exactly k
//      elements with ties going to the lower index, like a stable descending sort.
//   2. the k selected elements are compacted and radix-sorted on (score desc, index asc);
//   3. lazy chunked greedy NMS (hd_nms_core.cuh) until n_post boxes are kept.
// ------------------------------------------------------------------------------------------------
struct RpnSelParams {
    const float4* boxes; const float* scores; const uint32_t* keys;
    int B, N, n_pre, n_post;
    long long img_stride;   // elements between consecutive images in boxes / scores / keys (N, or more for a level slice of a wider array)
    float thr;
    float* out_rois;   // [B, n_post, 5]
    float* out_scores; // [B, n_post] nullable
    long long* out_idx;  // [B, n_post] nullable
    int* out_count;
    int cap;           // per-image workspace stride (>= min(n_pre, N))
    uint64_t* k0; uint64_t* k1; uint32_t* v0; uint32_t* v1; float4* sbox; int* keep_r; float4* gitem;
    int sort_off, bitonic_cap;  // dynamic smem: word offset and capacity (keys) of the in-smem sort area
    const int* only;            // nullable: run only the images whose flag is set (fallback pass behind the cluster kernel)
};

__device__ __forceinline__ uint64_t rpn_composite(uint32_t key, int idx) { return ((uint64_t)key << 32) | (uint32_t)(~(uint32_t)idx); }

__global__ void __launch_bounds__(RPN_NT, 1) rpn_select_nms_kernel(const __grid_constant__ RpnSelParams p) {
    extern __shared__ uint32_t removed[];
    __shared__ HdSortSmem<RPN_NT> ssm;
    __shared__ HdNmsSmem nsm;
    __shared__ HdGridSmem gsm;
    __shared__ int s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need, s_valid, s_count, s_cand;

    const int tid = threadIdx.x, lane = tid & 31;
    const int b = blockIdx.x;
    if (p.only && p.only[b] == 0) return;
    const uint32_t* __restrict__ keys = p.keys + (size_t)b * p.img_stride;
    const size_t off = (size_t)b * p.cap;
    uint64_t* k0 = p.k0 + off; uint64_t* k1 = p.k1 + off;
    uint32_t* v0 = p.v0 + off; uint32_t* v1 = p.v1 + off;
    float4* sbox = p.sbox + off;
    int* keep_r = p.keep_r + off;

    // ---- radix select of the k-th largest composite.  Full scans of the keys are latency bound, so every scan
    // keeps four independent loads in flight per thread, and after two digit passes (16 score bits) the few
    // candidates that still match the prefix are pulled into shared memory, where the remaining digits are resolved.
    unsigned long long* cand = reinterpret_cast<unsigned long long*>(&ssm.warp_cnt[0][0]);  // sort scratch, free until the sort
    constexpr int CAND_CAP = (RPN_NT / 32) * 256 * 4 / 8;                                    // 4096 composites
    auto scan_hist = [&](int sh, unsigned long long prefix, unsigned long long himask) {
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        for (int i0 = 0; i0 < p.N; i0 += RPN_U * RPN_NT) {
            uint32_t kk[RPN_U];
#pragma unroll
            for (int u = 0; u < RPN_U; ++u) { const int i = i0 + u * RPN_NT + tid; kk[u] = (i < p.N) ? keys[i] : 0u; }
#pragma unroll
            for (int u = 0; u < RPN_U; ++u) {
                const int i = i0 + u * RPN_NT + tid;
                int dg = 256 + lane;
                if (kk[u] != 0u) {
                    const unsigned long long comp = rpn_composite(kk[u], i);
                    if (((comp ^ prefix) & himask) == 0ull) dg = (int)((comp >> sh) & 255);
                }
                // scores cluster in a few digits: lanes agreeing with the first active lane are counted with one
                // ballot + one atomic, the rest add individually (cheaper than MATCH.ANY on every key)
                const unsigned act = __ballot_sync(HD_FULL, dg < 256);
                if (act) {
                    const int d0 = __shfl_sync(HD_FULL, dg, __ffs(act) - 1);
                    const unsigned same = __ballot_sync(HD_FULL, dg == d0);
                    if (dg == d0) { if ((same & hd_lanemask_lt()) == 0u) atomicAdd(&s_hist[d0], __popc(same)); }
                    else if (dg < 256) atomicAdd(&s_hist[dg], 1);
                }
            }
        }
        __syncthreads();
    };
    auto pick_digit = [&](int sh) {  // thread 0: digit holding the s_need-th largest, walking from the top
        if (tid == 0) {
            int need = s_need, d = 255;
            for (; d > 0; --d) {
                if (s_hist[d] >= need) break;
                need -= s_hist[d];
            }
            s_need = need;
            s_prefix |= ((unsigned long long)d << sh);
        }
        __syncthreads();
    };
    HD_PHASE(0);
    if (tid == 0) { s_count = 0; s_prefix = 0ull; }
    scan_hist(56, 0ull, 0ull);  // byte 7 over every valid key: also yields the number of valid proposals
    if (tid == 0) {
        int v = 0;
        for (int d = 0; d < 256; ++d) v += s_hist[d];
        s_valid = v;
    }
    __syncthreads();
    const int k = (p.n_pre > 0) ? min(p.n_pre, s_valid) : s_valid;
    if (k == 0) {
        for (int q = tid; q < p.n_post; q += RPN_NT) {
            float* o = p.out_rois + ((size_t)b * p.n_post + q) * 5;
            o[0] = (float)b; o[1] = o[2] = o[3] = o[4] = 0.0f;
            if (p.out_scores) p.out_scores[(size_t)b * p.n_post + q] = 0.0f;
            if (p.out_idx) p.out_idx[(size_t)b * p.n_post + q] = -1;
        }
        if (tid == 0) p.out_count[b] = 0;
        return;
    }
    uint64_t T = 0;
    if (k < s_valid) {
        if (tid == 0) s_need = k;
        __syncthreads();
        pick_digit(56);
        scan_hist(48, s_prefix, ~0ull << 56);
        pick_digit(48);
        // candidates sharing the 16-bit prefix -> shared memory
        if (tid == 0) s_cand = 0;
        __syncthreads();
        {
            const unsigned long long prefix = s_prefix;
            for (int i0 = 0; i0 < p.N; i0 += RPN_U * RPN_NT) {
                uint32_t kk[RPN_U];
#pragma unroll
                for (int u = 0; u < RPN_U; ++u) { const int i = i0 + u * RPN_NT + tid; kk[u] = (i < p.N) ? keys[i] : 0u; }
#pragma unroll
                for (int u = 0; u < RPN_U; ++u) {
                    const int i = i0 + u * RPN_NT + tid;
                    const unsigned long long comp = rpn_composite(kk[u], i);
                    const bool hit = kk[u] != 0u && ((comp ^ prefix) >> 48) == 0ull;
                    const unsigned m = __ballot_sync(HD_FULL, hit);
                    if (m) {
                        int basep = 0;
                        if (lane == 0) basep = atomicAdd(&s_cand, __popc(m));
                        basep = __shfl_sync(HD_FULL, basep, 0);
                        const int slot = basep + __popc(m & hd_lanemask_lt());
                        if (hit && slot < CAND_CAP) cand[slot] = comp;
                    }
                }
            }
        }
        __syncthreads();
        const int nc = s_cand;
        for (int byte = 5; byte >= 0; --byte) {
            if (byte == 3) continue;  // index < 2^24: byte 3 of ~index is 0xff for every element
            const int sh = byte * 8;
            const unsigned long long himask = ~0ull << (sh + 8);
            const unsigned long long pf = s_prefix | (byte <= 2 ? 0xff000000ull : 0ull);
            if (nc <= CAND_CAP) {
                if (tid < 256) s_hist[tid] = 0;
                __syncthreads();
                for (int i = tid; i < nc; i += RPN_NT) {
                    const unsigned long long comp = cand[i];
                    if (((comp ^ pf) & himask) == 0ull) atomicAdd(&s_hist[(int)((comp >> sh) & 255)], 1);
                }
                __syncthreads();
            } else {
                scan_hist(sh, pf, himask);  // pathological tie mass: keep scanning the full key array
            }
            pick_digit(sh);
        }
        T = s_prefix | 0xff000000ull;
    }
    __syncthreads();
    HD_PHASE(1);
    // ---- ordered compaction of the selected set: index order is kept, so a STABLE sort on the 32-bit score key alone
    // reproduces (score desc, index asc) -- half the radix passes of a 64-bit (score, index) key
    uint32_t* kk0 = reinterpret_cast<uint32_t*>(k0);
    uint32_t* kk1 = reinterpret_cast<uint32_t*>(k1);
    __shared__ int s_wsum[RPN_NT / 32];
    __shared__ int s_base;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < p.N; i0 += RPN_U * RPN_NT) {
        const int ib = i0 + RPN_U * tid;        // RPN_U consecutive proposals per thread
        uint32_t kk[RPN_U];
        bool sel[RPN_U];
        int c = 0;
#pragma unroll
        for (int u = 0; u < RPN_U; ++u) {
            const int i = ib + u;
            kk[u] = (i < p.N) ? keys[i] : 0u;
            sel[u] = kk[u] != 0u && rpn_composite(kk[u], i) >= T;
            c += sel[u];
        }
        int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(HD_FULL, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) s_wsum[tid >> 5] = incl;
        __syncthreads();
        int pre = s_base + incl - c;
        for (int w = 0; w < (tid >> 5); ++w) pre += s_wsum[w];
        int tot = 0;
        if (tid == RPN_NT - 1) tot = pre + c;   // last thread knows the new running total
#pragma unroll
        for (int u = 0; u < RPN_U; ++u) {
            if (sel[u]) {
                if (pre < p.cap) { kk0[pre] = ~kk[u]; v0[pre] = (uint32_t)(ib + u); }
                ++pre;
            }
        }
        __syncthreads();
        if (tid == RPN_NT - 1) s_base = tot;
        __syncthreads();
    }
    const int n = min(s_base, p.cap);
    HD_PHASE(2);
    int res = 0;
    if (n <= p.bitonic_cap) {
        // fits in shared memory: bitonic sort of the unique composite (~score << 32 | index), then unpack the order
        unsigned long long* skey = reinterpret_cast<unsigned long long*>(removed + p.sort_off);
        const int Np = hd_bitonic_padded(n);
        for (int i = tid; i < Np; i += RPN_NT) skey[i] = (i < n) ? (((unsigned long long)kk0[i] << 32) | v0[i]) : ~0ull;
        __syncthreads();
        if (Np == 2048) hd_cta_bitonic_reg<2, false>(skey, nullptr);
        else if (Np == 4096) hd_cta_bitonic_reg<4, false>(skey, nullptr);
        else hd_cta_bitonic_reg<8, false>(skey, nullptr);
        for (int i = tid; i < n; i += RPN_NT) v0[i] = (uint32_t)skey[i];
        __syncthreads();
    } else {
        res = hd_cta_radix_sort<RPN_NT, uint32_t>(kk0, v0, kk1, v1, n, ssm);
    }
    HD_PHASE(3);
    const uint32_t* order = res ? v1 : v0;
    const float4* __restrict__ boxes = p.boxes + (size_t)b * p.img_stride;
    for (int r = tid; r < n; r += RPN_NT) sbox[r] = boxes[order[r]];
    __syncthreads();
    HD_PHASE(4);
    int kc;
    if (n > HD_GRID_MIN_N && p.thr > 0.05f)
        kc = hd_cta_greedy_nms_grid<RPN_NT, int>(sbox, nullptr, n, p.n_post, p.thr, removed, keep_r, nsm, gsm, &ssm.warp_cnt[0][0], 12,
                                                 p.gitem + off);
    else
        kc = hd_cta_greedy_nms<RPN_NT, int>(sbox, nullptr, n, p.n_post, p.thr, removed, keep_r, nsm);
    HD_PHASE(5);
    for (int q = tid; q < p.n_post; q += RPN_NT) {
        float* o = p.out_rois + ((size_t)b * p.n_post + q) * 5;
        o[0] = (float)b;
        if (q < kc) {
            const int r = keep_r[q];
            const float4 bx = sbox[r];
            o[1] = bx.x; o[2] = bx.y; o[3] = bx.z; o[4] = bx.w;
            if (p.out_scores) p.out_scores[(size_t)b * p.n_post + q] = p.scores[(size_t)b * p.img_stride + order[r]];
            if (p.out_idx) p.out_idx[(size_t)b * p.n_post + q] = (long long)order[r];
        } else {
            o[1] = o[2] = o[3] = o[4] = 0.0f;
            if (p.out_scores) p.out_scores[(size_t)b * p.n_post + q] = 0.0f;
            if (p.out_idx) p.out_idx[(size_t)b * p.n_post + q] = -1;
        }
    }
    if (tid == 0) p.out_count[b] = kc;
    HD_PHASE(6);
}

// ------------------------------------------------------------------------------------------------
// select + sort + NMS spread over a thread-block CLUSTER: CL = 1, 2, 4 or 8 CTAs (SMs) per image, chosen per launch from the
// batch size and the number of clusters the device holds at once (15 of 8, 33 of 4, 74 of 2 on a B200).
//   S. radix select of the k-th largest 32-bit score key: every CTA histograms its 1/CL slice of the keys (held in
//      shared memory after one global read), the histograms are summed through distributed shared memory (DSMEM), the
//      digit is picked by a parallel suffix scan; ties on the threshold key go to the lowest indices (stable-sort rule) by
//      giving each CTA a quota of them.
//   C. ordered compaction of the selected set (index order kept) as unique composites (~key << 32 | index).
//   M. each CTA bitonic-sorts an even 1/CL slice of the composites in registers/shared memory; the final rank of an
//      element is the sum of its lower bounds in all sorted slices (binary searches in a shared-memory copy).
//   N. greedy NMS as a DAG problem (hd_cluster_nms.cuh): a size-stratified spatial hash is built cooperatively (per-CTA
//      bucket counts combined over DSMEM), every CTA lists, for its boxes j, the higher-ranked boxes i with IoU(i,j) > thr
//      (<= RPNC_ADJ of them, else the image is flagged for the single-CTA kernel), and CTA 0 resolves
//      kept[j] = !any(kept[i], i in adj[j])  1024 ranks at a time by Jacobi iteration to the unique fixed point (= the
//      greedy answer); ranks are taken in doubling batches and the loop stops at n_post keeps.
// Bit-identical outputs to rpn_select_nms_kernel (same selection rule, same IoU predicate).
// ------------------------------------------------------------------------------------------------
struct RpnClParams {
    RpnSelParams s;
    uint64_t* comp;      // [B,cap] compacted composites (index order)
    uint64_t* sorted;    // [B,cap] 8 sorted slices
    uint32_t* order;     // [B,cap] rank -> proposal index
    float4* gbox;        // [B,cap] boxes bucket by bucket
    int* grank;          // [B,cap] their ranks
    int* adj_cnt;        // [B,cap]
    unsigned short* adj; // [B,cap,RPNC_ADJ]
    int* fallback;       // [B] set when an adjacency list overflows
    int per;             // keys per CTA slice
    int key_cache;       // 1: the slice is held in shared memory
    HdClLayout lay;      // dynamic shared memory layout of the NMS phases
};

__global__ void __launch_bounds__(RPN_NT, 1) rpn_select_nms_cluster_kernel(const __grid_constant__ RpnClParams q) {
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ HdClSmem csm;
    int (&s_hist)[2][256] = csm.hist2;
    int (&s_tot)[256] = csm.tot;
    int (&s_wsum)[32] = csm.wsum;
    int& s_total = csm.total;
    __shared__ int s_cnt[2];
    __shared__ uint32_t s_prefix;
    __shared__ int s_need, s_base[2];

    cg::cluster_group cluster = cg::this_cluster();
    const RpnSelParams& p = q.s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int crank = (int)cluster.block_rank();
    const int CL = (int)cluster.num_blocks();
    const int b = blockIdx.x / CL;
    const size_t off = (size_t)b * p.cap;
    const uint32_t* __restrict__ gkeys = p.keys + (size_t)b * p.img_stride;
    const int lo = min(crank * q.per, p.N), cntl = min(p.N, lo + q.per) - lo;   // this CTA's slice of the keys
    uint32_t* kcache = reinterpret_cast<uint32_t*>(dsm);
    if (q.key_cache) {
        for (int i = tid; i < cntl; i += RPN_NT) kcache[i] = gkeys[lo + i];
        __syncthreads();
    }
    auto KEY = [&](int il) -> uint32_t { return q.key_cache ? kcache[il] : gkeys[lo + il]; };

    HD_PHASE(0);
    // ---- S: radix select on the 32-bit key, one byte per pass, histograms summed over the cluster
    auto scan_hist = [&](int pass, int sh, uint32_t prefix, uint32_t himask) {
        int* h = s_hist[pass & 1];
        if (tid < 256) h[tid] = 0;
        __syncthreads();
        for (int i0 = 0; i0 < cntl; i0 += RPN_NT) {
            const int il = i0 + tid;
            const uint32_t key = (il < cntl) ? KEY(il) : 0u;
            int dg = 256 + lane;
            if (key != 0u && ((key ^ prefix) & himask) == 0u) dg = (int)((key >> sh) & 255u);
            const unsigned act = __ballot_sync(HD_FULL, dg < 256);
            if (act) {
                const int d0 = __shfl_sync(HD_FULL, dg, __ffs(act) - 1);
                const unsigned same = __ballot_sync(HD_FULL, dg == d0);
                if (dg == d0) { if ((same & hd_lanemask_lt()) == 0u) atomicAdd(&h[d0], __popc(same)); }
                else if (dg < 256) atomicAdd(&h[dg], 1);
            }
        }
        __syncthreads();
        cluster.sync();
        if (tid < 256) {
            int t = 0;
            for (int c = 0; c < CL; ++c) t += cluster.map_shared_rank(h, c)[tid];
            s_tot[tid] = t;
        }
        __syncthreads();
    };
    scan_hist(0, 24, 0u, 0u);
    if (tid == 0) {
        int v = 0;
        for (int d = 0; d < 256; ++d) v += s_tot[d];
        s_total = v;
    }
    __syncthreads();
    const int valid = s_total;
    const int k = min((p.n_pre > 0) ? min(p.n_pre, valid) : valid, p.cap);
    if (k == 0) {
        if (crank == 0) {
            for (int r = tid; r < p.n_post; r += RPN_NT) {
                float* o = p.out_rois + ((size_t)b * p.n_post + r) * 5;
                o[0] = (float)b; o[1] = o[2] = o[3] = o[4] = 0.0f;
                if (p.out_scores) p.out_scores[(size_t)b * p.n_post + r] = 0.0f;
                if (p.out_idx) p.out_idx[(size_t)b * p.n_post + r] = -1;
            }
            if (tid == 0) p.out_count[b] = 0;
        }
        cluster.sync();   // nobody leaves while its histogram may still be read
        return;
    }
    uint32_t Tkey = 0u;   // selected: key > Tkey, plus the `need` lowest-index elements with key == Tkey
    int need = 0;
    if (k < valid) {
        if (tid == 0) { s_need = k; s_prefix = 0u; }
        __syncthreads();
        for (int byte = 3; byte >= 0; --byte) {
            const int sh = byte * 8;
            if (byte < 3) scan_hist(3 - byte, sh, s_prefix, ~0u << (sh + 8));
            // digit holding the s_need-th largest key: suffix sums of the 256 totals (8 warps), then the one thread whose
            // suffix crosses s_need publishes the digit
            int suf = 0, mine = 0;
            if (tid < 256) {
                mine = s_tot[tid];
                suf = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_down_sync(HD_FULL, suf, d); if (lane + d < 32) suf += y; }
                if (lane == 0) s_wsum[wid] = suf;
            }
            __syncthreads();
            const int nd = s_need;
            __syncthreads();
            if (tid < 256) {
                for (int w = wid + 1; w < 8; ++w) suf += s_wsum[w];      // suf = sum of totals of digits >= tid
                const int above = suf - mine;                            // digits > tid
                if (suf >= nd && above < nd) { s_need = nd - above; s_prefix |= ((uint32_t)tid << sh); }
            }
            __syncthreads();
        }
        Tkey = s_prefix;
        need = s_need;
    }
    HD_PHASE(1);
    // ---- C: ordered compaction.  Per-CTA counts -> every CTA derives its output base and its quota of threshold ties
    {
        int cg_ = 0, ce = 0;
        for (int il = tid; il < cntl; il += RPN_NT) {
            const uint32_t key = KEY(il);
            cg_ += (key > Tkey);
            ce += (key == Tkey && key != 0u);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { cg_ += __shfl_xor_sync(HD_FULL, cg_, d); ce += __shfl_xor_sync(HD_FULL, ce, d); }
        if (lane == 0) { s_wsum[wid] = cg_; s_hist[0][wid] = ce; }   // (the select histograms are dead by now)
        __syncthreads();
        if (tid == 0) {
            int g = 0, e = 0;
            for (int w = 0; w < RPN_NT / 32; ++w) { g += s_wsum[w]; e += s_hist[0][w]; }
            s_cnt[0] = g; s_cnt[1] = e;
        }
    }
    cluster.sync();
    int out_base = 0, quota = 0;
    {
        int eq_before = 0;
        for (int c = 0; c < CL; ++c) {
            const int* rc = cluster.map_shared_rank(s_cnt, c);
            const int g = rc[0], e = rc[1];
            const int qc = min(max(need - eq_before, 0), e);
            if (c < crank) out_base += g + qc;
            if (c == crank) quota = qc;
            eq_before += e;
        }
    }
    uint64_t* comp = q.comp + off;
    if (tid == 0) { s_base[0] = 0; s_base[1] = 0; }
    __syncthreads();
    for (int i0 = 0; i0 < cntl; i0 += RPN_U * RPN_NT) {
        const int ib = i0 + RPN_U * tid;   // RPN_U consecutive keys per thread
        uint32_t kk[RPN_U];
        int cgt = 0, ceq = 0;
#pragma unroll
        for (int u = 0; u < RPN_U; ++u) {
            kk[u] = (ib + u < cntl) ? KEY(ib + u) : 0u;
            cgt += (kk[u] > Tkey);
            ceq += (kk[u] == Tkey && kk[u] != 0u);
        }
        const int mine = cgt | (ceq << 16);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(HD_FULL, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) s_wsum[wid] = incl;
        __syncthreads();
        int pre = incl - mine;
        for (int w = 0; w < wid; ++w) pre += s_wsum[w];
        int gb = s_base[0] + (pre & 0xffff), eb = s_base[1] + (pre >> 16);
        const int tot = pre + mine;
#pragma unroll
        for (int u = 0; u < RPN_U; ++u) {
            const bool gt = kk[u] > Tkey, eq = (kk[u] == Tkey && kk[u] != 0u);
            if (gt || (eq && eb < quota)) {
                const int pos = out_base + gb + min(eb, quota);
                if (pos < p.cap) comp[pos] = ((unsigned long long)(~kk[u]) << 32) | (uint32_t)(lo + ib + u);
            }
            gb += gt; eb += eq;
        }
        __syncthreads();
        if (tid == RPN_NT - 1) { s_base[0] += tot & 0xffff; s_base[1] += tot >> 16; }
        __syncthreads();
    }
    cluster.sync();
    HD_PHASE(2);
    // ---- M: slice sort + merge ranks
    const int n = k;
    const int m = (n + CL - 1) / CL;                           // <= 8192 (host check)
    const int slo = min(crank * m, n), slen = min(n, slo + m) - slo;
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(dsm);
    const int Np = hd_bitonic_padded(m);
    for (int i = tid; i < Np; i += RPN_NT) skey[i] = (i < slen) ? comp[slo + i] : ~0ull;
    __syncthreads();
    if (Np == 2048) hd_cta_bitonic_reg<2, false>(skey, nullptr);
    else if (Np == 4096) hd_cta_bitonic_reg<4, false>(skey, nullptr);
    else hd_cta_bitonic_reg<8, false>(skey, nullptr);
    uint64_t* sorted = q.sorted + off;
    for (int i = tid; i < slen; i += RPN_NT) sorted[slo + i] = skey[i];
    cluster.sync();
    for (int i = tid; i < n; i += RPN_NT) skey[i] = sorted[i];
    __syncthreads();
    uint32_t* order = q.order + off;
    float4* sbox = p.sbox + off;
    const float4* __restrict__ boxes = p.boxes + (size_t)b * p.img_stride;
    for (int i = tid; i < slen; i += RPN_NT) {
        const unsigned long long v = skey[slo + i];
        int rank = i;                                          // lower bound in the own slice
        for (int c = 0; c < CL; ++c) {
            if (c == crank) continue;
            int a = min(c * m, n), z = min(n, a + m);          // lower bound of v in slice c
            const int a0 = a;
            while (a < z) { const int mid = (a + z) >> 1; if (skey[mid] < v) a = mid + 1; else z = mid; }
            rank += a - a0;
        }
        const uint32_t idx = (uint32_t)v;
        order[rank] = idx;
        sbox[rank] = boxes[idx];
    }
    cluster.sync();
    HD_PHASE(3);
    // ---- N: greedy NMS over the cluster (hd_cluster_nms.cuh): size-stratified spatial hash -> adjacency lists -> Jacobi resolve,
    // in rank batches with early stop
    HdClWs w;
    w.sbox = sbox; w.scls = nullptr; w.gbox = q.gbox + off; w.grank = q.grank + off; w.adj_cnt = q.adj_cnt + off;
    w.adj = q.adj + off * RPNC_ADJ; w.fallback = q.fallback + b; w.keep_r = p.keep_r + off;
    const int kc = hd_cluster_greedy_nms<RPN_NT, false>(cluster, dsm, q.lay, csm, w, n, p.n_post, p.thr);
    if (crank != 0 || kc < 0) return;
    const int* keep_r = w.keep_r;
    HD_PHASE(6);
    for (int r = tid; r < p.n_post; r += RPN_NT) {
        float* o = p.out_rois + ((size_t)b * p.n_post + r) * 5;
        o[0] = (float)b;
        if (r < kc) {
            const int rr = keep_r[r];
            const float4 bx = sbox[rr];
            o[1] = bx.x; o[2] = bx.y; o[3] = bx.z; o[4] = bx.w;
            if (p.out_scores) p.out_scores[(size_t)b * p.n_post + r] = p.scores[(size_t)b * p.img_stride + order[rr]];
            if (p.out_idx) p.out_idx[(size_t)b * p.n_post + r] = (long long)order[rr];
        } else {
            o[1] = o[2] = o[3] = o[4] = 0.0f;
            if (p.out_scores) p.out_scores[(size_t)b * p.n_post + r] = 0.0f;
            if (p.out_idx) p.out_idx[(size_t)b * p.n_post + r] = -1;
        }
    }
    if (tid == 0) p.out_count[b] = kc;
    HD_PHASE(7);
}

// ------------------------------------------------------------------------------------------------ host
static int rpn_fill(RpnParams& p, const hd_rpn_level* levels, int n_levels, int B, int A, int flags, float img_h, float img_w,
                    float min_size, float clamp_dwh) {
    HD_CHECK_ARG(levels && n_levels >= 1 && n_levels <= HD_MAX_LEVELS, "n_levels must be in [1,%d], got %d", HD_MAX_LEVELS, n_levels);
    HD_CHECK_ARG(A >= 1 && A <= HD_MAX_ANCHORS, "A must be in [1,%d], got %d", HD_MAX_ANCHORS, A);
    HD_CHECK_ARG(B >= 0, "B must be >= 0");
    memset(&p, 0, sizeof(p));
    int cells = 0, anchors = 0;
    for (int l = 0; l < n_levels; ++l) {
        HD_CHECK_ARG(levels[l].H > 0 && levels[l].W > 0, "level %d has empty spatial size", l);
        HD_CHECK_ARG(B == 0 || (levels[l].objectness && levels[l].deltas), "level %d has a NULL head", l);
        p.obj[l] = levels[l].objectness; p.dlt[l] = levels[l].deltas;
        p.H[l] = levels[l].H; p.W[l] = levels[l].W; p.stride[l] = levels[l].stride;
        for (int q = 0; q < 4 * A; ++q) p.base[l][q] = levels[l].anchor_base[q];
        p.cell_start[l] = cells; p.level_off[l] = anchors;
        cells += levels[l].H * levels[l].W;
        anchors += A * levels[l].H * levels[l].W;
    }
    HD_CHECK_ARG(anchors < (1 << 24), "more than 2^24 anchors per image");
    p.cell_start[n_levels] = cells; p.level_off[n_levels] = anchors;
    p.n_levels = n_levels; p.B = B; p.A = A; p.softmax = (flags & HD_RPN_SOFTMAX) ? 1 : 0;
    p.img_h = img_h; p.img_w = img_w; p.min_size = min_size;
    p.use_clamp = (flags & HD_RPN_CLAMP_DWH) ? 1 : 0; p.clamp_dwh = clamp_dwh;
    p.key_logit = (flags & HD_RPN_KEY_LOGIT) ? 1 : 0;
    p.exact = (flags & HD_RPN_EXACT_MATH) ? 1 : 0;
    p.total_cells = (long long)B * cells; p.N = anchors;
    return HD_OK;
}

extern "C" HD_API int hd_rpn_num_anchors(const hd_rpn_level* levels, int n_levels, int A) {
    if (!levels || n_levels < 1 || n_levels > HD_MAX_LEVELS) return HD_ERR_INVALID;
    long long n = 0;
    for (int l = 0; l < n_levels; ++l) n += (long long)A * levels[l].H * levels[l].W;
    return (int)n;
}

extern "C" HD_API int hd_rpn_decode(const hd_rpn_level* levels, int n_levels, int B, int A, int flags, float img_h, float img_w,
                                    float min_size, float clamp_dwh, float* boxes, float* scores, uint32_t* keys, void* stream) {
    RpnParams p;
    int rc = rpn_fill(p, levels, n_levels, B, A, flags, img_h, img_w, min_size, clamp_dwh);
    if (rc) return rc;
    if (B == 0) return HD_OK;
    HD_CHECK_ARG(boxes && scores && keys, "null output");
    long long blocks = (p.total_cells + 255) / 256;
    HD_CHECK_ARG(blocks < (1ll << 31), "grid too large");
    rpn_decode_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, (float4*)boxes, scores, keys);
    HD_CUDA_LAUNCH_CHECK("rpn_decode_kernel");
    return HD_OK;
}

#define RPN_WS_PARTS 11
static bool rpn_cluster_ok(int cap) { return cap <= RPNC_MAXN; }
static void rpn_ws_layout(int B, int cap, size_t* offs, size_t* total) {
    size_t n = (size_t)B * cap, o = 0;
    offs[0] = o; o = hd_align_up(o + n * 8, 256);
    offs[1] = o; o = hd_align_up(o + n * 8, 256);
    offs[2] = o; o = hd_align_up(o + n * 4, 256);
    offs[3] = o; o = hd_align_up(o + n * 4, 256);
    offs[4] = o; o = hd_align_up(o + n * 16, 256);
    offs[5] = o; o = hd_align_up(o + n * 4, 256);
    offs[6] = o; o = hd_align_up(o + n * 16, 256);
    // cluster kernel only: bucket ranks, adjacency counts + lists, per-image fallback flags
    const size_t nc = rpn_cluster_ok(cap) ? n : 0;
    offs[7] = o; o = hd_align_up(o + nc * 4, 256);
    offs[8] = o; o = hd_align_up(o + nc * 4, 256);
    offs[9] = o; o = hd_align_up(o + nc * RPNC_ADJ * 2, 256);
    offs[10] = o; o = hd_align_up(o + (size_t)B * 4, 256);
    *total = o;
}
static int rpn_cluster_capacity(int CL) {
    // how many CL-CTA clusters of the stage-2 kernel the device holds at once (1 CTA per SM: 1024 threads x 64 registers)
    static int cached_dev[HD_MAX_DEVICES][RPNC_MAXCL + 1] = {{0}};   // per device; racing writers store the same value
    const int dev = hd_current_device();
    int* cached = cached_dev[(dev >= 0 && dev < HD_MAX_DEVICES) ? dev : 0];
    if (cached[CL]) return cached[CL];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL * 64); cfg.blockDim = dim3(RPN_NT); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int nc = 0;
    cudaFuncSetAttribute(rpn_select_nms_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (cudaOccupancyMaxActiveClusters(&nc, rpn_select_nms_cluster_kernel, &cfg) != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = hd_num_sms() / CL; }
    return cached[CL] = nc;
}
extern "C" HD_API int hd_rpn_cluster_capacity(int cluster_size) {
    if (cluster_size != 1 && cluster_size != 2 && cluster_size != 4 && cluster_size != 8) return -1;
    return rpn_cluster_capacity(cluster_size);
}
// developer/test knobs (process-wide, atomic): they select between kernels that return identical results
static int g_rpn_cluster = 0;   // 0: chosen per launch
extern "C" HD_API int hd_rpn_set_cluster_size(int cl) { return __atomic_exchange_n(&g_rpn_cluster, (cl == 1 || cl == 2 || cl == 4 || cl == 8) ? cl : 0, __ATOMIC_ACQ_REL); }
static int g_rpn_mode = 0;   // 0 auto, 1 single CTA per image, 2 cluster (when eligible)
extern "C" HD_API int hd_rpn_set_mode(int mode) { return __atomic_exchange_n(&g_rpn_mode, mode, __ATOMIC_ACQ_REL); }
static int rpn_cap(int N, int n_pre) { return (n_pre > 0 && n_pre < N) ? n_pre : N; }

extern "C" HD_API size_t hd_rpn_select_nms_workspace_size(int B, int N, int n_pre) {
    size_t offs[RPN_WS_PARTS], total;
    rpn_ws_layout(B < 0 ? 0 : B, rpn_cap(N < 0 ? 0 : N, n_pre), offs, &total);
    return total + 256;
}

extern "C" HD_API int hd_rpn_select_nms(const float* boxes, const float* scores, const uint32_t* keys, int B, int N, int n_pre, int n_post,
                                        double nms_iou, float* out_rois, float* out_scores, int64_t* out_idx, int32_t* out_count,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    return hd_rpn_select_nms_strided(boxes, scores, keys, B, N, (int64_t)N, n_pre, n_post, nms_iou, out_rois, out_scores, out_idx, out_count,
                                     workspace, workspace_bytes, stream);
}

extern "C" HD_API int hd_rpn_select_nms_strided(const float* boxes, const float* scores, const uint32_t* keys, int B, int N, int64_t image_stride,
                                                int n_pre, int n_post, double nms_iou, float* out_rois, float* out_scores, int64_t* out_idx,
                                                int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
    HD_CHECK_ARG(image_stride >= N, "image_stride %lld < N %d", (long long)image_stride, N);
    HD_CHECK_ARG(B >= 0 && N >= 0 && n_post > 0, "bad shape B=%d N=%d n_post=%d", B, N, n_post);
    HD_CHECK_ARG(N < (1 << 24), "more than 2^24 proposals per image");
    if (B == 0) return HD_OK;
    HD_CHECK_ARG(out_rois && out_count, "null output");
    HD_CHECK_ARG(N == 0 || (boxes && scores && keys), "null input");
    const int cap = rpn_cap(N, n_pre) > 0 ? rpn_cap(N, n_pre) : 1;
    size_t offs[RPN_WS_PARTS], total;
    rpn_ws_layout(B, cap, offs, &total);
    uintptr_t w0 = hd_align_up((uintptr_t)workspace, 256);
    if (!workspace || w0 + total > (uintptr_t)workspace + workspace_bytes)
        HD_FAIL(HD_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", total + 256, workspace_bytes);
    RpnSelParams p;
    p.only = nullptr;
    p.boxes = (const float4*)boxes; p.scores = scores; p.keys = keys; p.B = B; p.N = N; p.n_pre = n_pre; p.n_post = n_post;
    p.img_stride = image_stride;
    p.thr = hd_thr_floor(nms_iou);
    p.out_rois = out_rois; p.out_scores = out_scores; p.out_idx = (long long*)out_idx; p.out_count = out_count; p.cap = cap;
    p.k0 = (uint64_t*)(w0 + offs[0]); p.k1 = (uint64_t*)(w0 + offs[1]); p.v0 = (uint32_t*)(w0 + offs[2]); p.v1 = (uint32_t*)(w0 + offs[3]);
    p.sbox = (float4*)(w0 + offs[4]); p.keep_r = (int*)(w0 + offs[5]); p.gitem = (float4*)(w0 + offs[6]);
    size_t words = ((size_t)(cap + 31) / 32 + 4 + 1) & ~(size_t)1;
    HD_CHECK_ARG(words * 4 <= 56 * 1024, "n_pre too large for the shared-memory bitmap");
    p.sort_off = (int)words;
    // selections of up to 8192 proposals sort in shared memory; larger ones (e.g. 12 000) use the 32-bit stable radix
    // passes and keep the shared memory for L1 (the pruned NMS streams its records through it)
    p.bitonic_cap = (cap <= 8192) ? 8192 : 0;
    size_t smem = words * 4 + (size_t)p.bitonic_cap * 8;
    HD_ENSURE_SMEM(rpn_select_nms_kernel, 186 * 1024);
    HD_ENSURE_SMEM(rpn_select_nms_cluster_kernel, 200 * 1024);
    const int rpn_mode = __atomic_load_n(&g_rpn_mode, __ATOMIC_ACQUIRE);
    if (rpn_mode != 1 && rpn_cluster_ok(cap) && N > 0) {
        // CL SMs per image; an image whose adjacency lists overflow is redone by the single-CTA kernel below.
        // CL minimises waves * (serial + parallel / CL) with phase times measured on B200 (profiles/r1_results.md).
        int CL = __atomic_load_n(&g_rpn_cluster, __ATOMIC_ACQUIRE);
        if (!CL) {
            double best = 1e30;
            for (int c = RPNC_MAXCL; c >= 1; c >>= 1) {
                if ((cap + c - 1) / c > 8192) continue;
                const int capn = rpn_cluster_capacity(c);
                const double t = (double)((B + capn - 1) / capn) * (90.0 + 480.0 / c);
                if (t < best) { best = t; CL = c; }
            }
        }
        while ((cap + CL - 1) / CL > 8192) CL <<= 1;
        RpnClParams q;
        q.s = p;
        q.comp = p.k0; q.sorted = p.k1; q.order = p.v1; q.gbox = p.gitem;
        q.grank = (int*)(w0 + offs[7]); q.adj_cnt = (int*)(w0 + offs[8]); q.adj = (unsigned short*)(w0 + offs[9]);
        q.fallback = (int*)(w0 + offs[10]);
        q.per = (int)(((size_t)(N + CL - 1) / CL + 3) & ~(size_t)3);
        const size_t budget = 200 * 1024;
        q.key_cache = ((size_t)q.per * 4 <= budget) ? 1 : 0;
        const int mslice = (cap + CL - 1) / CL;
        const size_t np = mslice <= 2048 ? 2048 : (mslice <= 4096 ? 4096 : 8192);
        size_t sm_c = np * 8;                                           // slice sort
        if (sm_c < (size_t)cap * 8) sm_c = (size_t)cap * 8;             // merge: all composites
        if (q.key_cache && sm_c < (size_t)q.per * 4) sm_c = (size_t)q.per * 4;
        hd_cluster_layout(cap, budget, &q.lay, &sm_c);
        HD_CUDA_CALL(cudaMemsetAsync(q.fallback, 0, (size_t)B * 4, (cudaStream_t)stream));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(B * CL)); cfg.blockDim = dim3(RPN_NT); cfg.dynamicSmemBytes = sm_c; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        HD_CUDA_CALL(cudaLaunchKernelEx(&cfg, rpn_select_nms_cluster_kernel, q));
        hd_count_launch();
        p.only = q.fallback;
    }
    rpn_select_nms_kernel<<<B, RPN_NT, smem, (cudaStream_t)stream>>>(p);
    HD_CUDA_LAUNCH_CHECK("rpn_select_nms_kernel");
    return HD_OK;
}

extern "C" HD_API size_t hd_rpn_proposals_workspace_size(int B, int N, int n_pre) {
    size_t n = (size_t)(B < 0 ? 0 : B) * (N < 0 ? 0 : N);
    return hd_align_up(n * 16, 256) + hd_align_up(n * 4, 256) * 2 + hd_rpn_select_nms_workspace_size(B, N, n_pre) + 256;
}

extern "C" HD_API int hd_rpn_proposals(const hd_rpn_level* levels, int n_levels, int B, int A, int flags, float img_h, float img_w,
                                       float min_size, float clamp_dwh, int n_pre, int n_post, double nms_iou, float* out_rois,
                                       float* out_scores, int64_t* out_idx, int32_t* out_count, void* workspace, size_t workspace_bytes,
                                       void* stream) {
    int N = hd_rpn_num_anchors(levels, n_levels, A);
    HD_CHECK_ARG(N >= 0, "bad levels");
    if (B <= 0) return B == 0 ? HD_OK : HD_ERR_INVALID;
    size_t need = hd_rpn_proposals_workspace_size(B, N, n_pre);
    if (!workspace || workspace_bytes < need) HD_FAIL(HD_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    size_t n = (size_t)B * N;
    uintptr_t w = hd_align_up((uintptr_t)workspace, 256);
    float* boxes = (float*)w; w += hd_align_up(n * 16, 256);
    float* scores = (float*)w; w += hd_align_up(n * 4, 256);
    uint32_t* keys = (uint32_t*)w; w += hd_align_up(n * 4, 256);
    int rc = hd_rpn_decode(levels, n_levels, B, A, flags, img_h, img_w, min_size, clamp_dwh, boxes, scores, keys, stream);
    if (rc) return rc;
    return hd_rpn_select_nms(boxes, scores, keys, B, N, n_pre, n_post, nms_iou, out_rois, out_scores, out_idx, out_count, (void*)w,
                             (size_t)((uintptr_t)workspace + workspace_bytes - w), stream);
}

HD_DEFINE_PHASE_READER(hd_phase_reader_rpn)

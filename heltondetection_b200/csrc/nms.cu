// Batched per-image sort + greedy NMS (a3/a4) and pairwise IoU (a5).
// Replaces torchvision.ops.nms / batched_nms / box_iou (boxes.py:20-120, 308-370).
//
// One CTA per image.  Candidates are ordered by a CTA radix sort on (score desc, tiebreak asc), the
// sorted boxes are materialised once, then suppression runs *lazily*: boxes are visited in chunks of
// 64 in score order; the chunk's 64x64 upper-triangular IoU bitmask is built with all threads,
// one warp resolves the chunk serially with bit operations, and only the boxes that were actually
// kept are tested against the still-alive tail.  The `removed` bitmap lives in shared memory for the
// whole image.  Work is kept*n instead of n^2/2 and the loop stops as soon as max_det boxes are kept,
// which is what the detection callers (max_det=300, RPN post_nms_top_n) need.
#include "hd_sort.cuh"
#include "hd_small_nms.cuh"
#include "hd_cluster_nms.cuh"

#define NMS_NT 1024

struct NmsParams {
    const float4* boxes;
    const float* scores;
    const int* cls;
    const int* tiebreak;
    const int* counts;
    int n_fixed, B, cap, min_n;  // images with n <= min_n are skipped (already handled by the fused small-n path)
    float thr;  // hd_thr_floor(iou_thres)
    int class_mode;
    float offset_scale;
    int max_nms, max_det;
    float* out_det;
    long long* out_idx;
    int* out_count;
    // workspace (per image stride = cap)
    uint64_t* k0; uint64_t* k1; uint32_t* v0; uint32_t* v1;
    float4* sbox; int* scls; int* keep_r; float4* gitem;
    int sort_off, bitonic_cap;  // dynamic smem: word offset of the in-smem sort area and its capacity (0 = disabled)
    HdRep rep;
    const int* only;            // nullable: run only the images whose flag is set (images the cluster kernel hands back)
    unsigned* hint;             // nullable: mapped host word that receives `call_id` when some image takes this (large) path
    unsigned call_id;
    int bucket_sort;            // 1: dense images (n > NMS_BUCKET_MIN_N) use the bucket sort instead of the bitonic network
};

// NT = 1024: one whole SM per image (64 K registers, ~190 KB shared memory) -- the fast configuration for dense scenes.
// NT = 256 ("light"): the same algorithm in a CTA that fits beside the decode kernel's CTAs.  Launching the 1024-thread variant
// for a batch that has no large image still needs B empty SMs, i.e. it drains whatever else is running -- fatal for the
// software-pipelined small-batch steps -- so the host launches the light variant while recent calls saw no large image
// (see nms_large_recent below); either variant returns the same bits.
#define NMS_BUCKET_MIN_N 2048
template <int NT>
__global__ void __launch_bounds__(NT, 1) sort_nms_kernel(const __grid_constant__ NmsParams p) {
    extern __shared__ uint32_t removed[];  // ceil(cap/32)+4 words, (+pad)
    __shared__ HdSortSmem<NT> ssm;
    __shared__ HdNmsSmem nsm;
    __shared__ HdGridSmem gsm;

    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    int n = p.counts ? min(p.counts[b], p.cap) : p.n_fixed;
    if (p.min_n >= 0 && n <= p.min_n) return;
    if (p.only && p.only[b] == 0) return;
    if (n <= 0) {
        if (tid == 0) p.out_count[b] = 0;
        if (tid < p.rep.n) p.rep.cnt[tid][b] = 0;
        hd_zero_tail(p.out_det, p.out_idx, p.rep, p.max_det, b, 0);
        return;
    }
    if (p.hint && tid == 0) *reinterpret_cast<volatile unsigned*>(p.hint) = p.call_id;   // "this call had a large image"
    const size_t off = (size_t)b * p.cap;
    uint64_t* k0 = p.k0 + off; uint64_t* k1 = p.k1 + off;
    uint32_t* v0 = p.v0 + off; uint32_t* v1 = p.v1 + off;
    float4* sbox = p.sbox + off;
    int* scls = p.scls + off;
    int* keep_r = p.keep_r + off;

    HD_PHASE(0);
    const uint32_t* order;
    if (n <= p.bitonic_cap) {
        // segment fits in shared memory: bitonic sort of (key, slot) right there
        unsigned long long* skey = reinterpret_cast<unsigned long long*>(removed + p.sort_off);
        uint32_t* sval = reinterpret_cast<uint32_t*>(skey + p.bitonic_cap);
        const int N = hd_bitonic_padded(n);
        bool sorted = false;
        if (NT == 1024 && n > NMS_BUCKET_MIN_N && p.bucket_sort) {
            // dense image: composites straight into registers, bucket sort on the score word (hd_sort.cuh)
            unsigned long long kreg[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = tid + e * NT;
                kreg[e] = ~0ull;
                if (i < n) {
                    const uint32_t tb = p.tiebreak ? (uint32_t)p.tiebreak[off + i] : (uint32_t)i;
                    kreg[e] = ((uint64_t)(~hd_orderable(p.scores[off + i])) << 32) | tb;
                }
            }
            HD_PHASE(2);
            sorted = hd_cta_bucket_sort<NT, 8>(kreg, n, skey, sval, &ssm.warp_cnt[0][0], ssm.hist);
        }
        if (!sorted) {
            for (int i = tid; i < N; i += NT) {
                if (i < n) {
                    const uint32_t tb = p.tiebreak ? (uint32_t)p.tiebreak[off + i] : (uint32_t)i;
                    skey[i] = ((uint64_t)(~hd_orderable(p.scores[off + i])) << 32) | tb;
                    sval[i] = (uint32_t)i;
                } else {
                    skey[i] = ~0ull;
                    sval[i] = 0u;
                }
            }
            __syncthreads();
            HD_PHASE(2);
            if (NT == 1024) {   // (bitonic_cap is 0 for the light variant: the register-blocked network needs 1024 threads)
                if (N == 2048) hd_cta_bitonic_reg<2, true>(skey, sval);
                else if (N == 4096) hd_cta_bitonic_reg<4, true>(skey, sval);
                else hd_cta_bitonic_reg<8, true>(skey, sval);
            }
        }
        order = sval;
    } else {
        for (int i = tid; i < n; i += NT) {
            uint32_t tb = p.tiebreak ? (uint32_t)p.tiebreak[off + i] : (uint32_t)i;
            k0[i] = ((uint64_t)(~hd_orderable(p.scores[off + i])) << 32) | tb;
            v0[i] = (uint32_t)i;
        }
        __syncthreads();
        HD_PHASE(2);
        const int res = hd_cta_radix_sort<NT>(k0, v0, k1, v1, n, ssm);
        order = res ? v1 : v0;
    }
    HD_PHASE(3);

    const int n_use = (p.max_nms > 0) ? min(n, p.max_nms) : n;
    const int max_det = (p.max_det > 0) ? p.max_det : n_use;
    for (int r = tid; r < n_use; r += NT) {
        const uint32_t slot = order[r];
        float4 bx = p.boxes[off + slot];
        int c = p.cls ? p.cls[off + slot] : 0;
        if (p.class_mode == HD_NMS_CLASS_OFFSET) {
            const float o = __fmul_rn((float)c, p.offset_scale);
            bx.x = __fadd_rn(bx.x, o); bx.y = __fadd_rn(bx.y, o); bx.z = __fadd_rn(bx.z, o); bx.w = __fadd_rn(bx.w, o);
        }
        sbox[r] = bx;
        scls[r] = (p.class_mode == HD_NMS_CLASS_EXACT) ? c : 0;
    }
    __syncthreads();
    const int* clsp = (p.class_mode == HD_NMS_CLASS_EXACT) ? scls : nullptr;
    HD_PHASE(4);
    int kc;
    if (n_use > HD_GRID_MIN_N && p.thr > 0.05f) {
        // big segment: spatially pruned pass; buckets alias the (finished) sort scratch, items use the free key buffer
        // buckets alias the sort scratch: (NT/32)*256 ints -> 4096 buckets for the 1024-thread CTA, 1024 for the light one
        kc = hd_cta_greedy_nms_grid<NT, int>(sbox, clsp, n_use, max_det, p.thr, removed, keep_r, nsm, gsm, &ssm.warp_cnt[0][0],
                                                 NT == 1024 ? 12 : 10, p.gitem + off);
    } else {
        kc = hd_cta_greedy_nms<NT, int>(sbox, clsp, n_use, max_det, p.thr, removed, keep_r, nsm);
    }

    HD_PHASE(5);
    for (int q = tid; q < kc; q += NT) {
        const int r = keep_r[q];
        const uint32_t slot = order[r];
        if (p.out_det) {
            const float4 bx = p.boxes[off + slot];
            const float sc = p.scores[off + slot], cf = p.cls ? (float)p.cls[off + slot] : 0.0f;
            const size_t ro = ((size_t)b * p.max_det + q) * 6;
            float* o = p.out_det + ro;
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = sc; o[5] = cf;
            for (int r = 0; r < p.rep.n; ++r) {   // posted stores into the peers' gather buffers
                float* pr = p.rep.det[r] + ro;
                pr[0] = bx.x; pr[1] = bx.y; pr[2] = bx.z; pr[3] = bx.w; pr[4] = sc; pr[5] = cf;
            }
        }
        if (p.out_idx) p.out_idx[(size_t)b * p.max_det + q] = p.tiebreak ? (long long)p.tiebreak[off + slot] : (long long)slot;
    }
    if (tid == 0) p.out_count[b] = kc;
    if (tid < p.rep.n) p.rep.cnt[tid][b] = kc;
    hd_zero_tail(p.out_det, p.out_idx, p.rep, p.max_det, b, kc);
}

// ------------------------------------------------------------------------------------------------------------
// Cluster variant for batches that leave SMs idle (B < #SMs / 2): 2, 4 or 8 CTAs per image.  Each CTA bitonic-sorts an
// even slice of the (score desc, tiebreak asc) composites with the slot as payload, the final rank of an element is the
// sum of its lower bounds in all sorted slices (shared-memory copy), and the greedy NMS is hd_cluster_greedy_nms
// (hd_cluster_nms.cuh).  Images with more than RPNC_MAXN candidates, or whose adjacency lists overflow, are flagged and
// redone by sort_nms_kernel; images with <= min_n candidates belong to small_nms_kernel.  Bit-identical outputs.
// ------------------------------------------------------------------------------------------------------------
struct NmsClParams {
    NmsParams s;
    uint64_t* skeys; uint32_t* svals;   // [B,cap] sorted slices
    uint32_t* order;                    // [B,cap] rank -> slot
    float4* gbox; int* grank; int* adj_cnt; unsigned short* adj; int* fallback;
    HdClLayout lay;
    int val_off;                        // byte offset of the payload array in dynamic shared memory
    int capn;                           // per-image stride of the adjacency lists: min(cap, RPNC_MAXN)
};

__global__ void __launch_bounds__(NMS_NT, 1) sort_nms_cluster_kernel(const __grid_constant__ NmsClParams q) {
    if (q.s.hint && threadIdx.x == 0 && blockIdx.x == 0) {   // any image beyond the small kernel's range -> the call was "large"
        for (int b = 0; b < q.s.B; ++b) {
            const int nb = q.s.counts ? min(q.s.counts[b], q.s.cap) : q.s.n_fixed;
            if (nb > HD_SMALL_N) { *reinterpret_cast<volatile unsigned*>(q.s.hint) = q.s.call_id; break; }
        }
    }
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ HdClSmem csm;
    cg::cluster_group cluster = cg::this_cluster();
    const NmsParams& p = q.s;
    const int tid = threadIdx.x;
    const int crank = (int)cluster.block_rank(), CL = (int)cluster.num_blocks();
    const int b = blockIdx.x / CL;
    const int n = p.counts ? min(p.counts[b], p.cap) : p.n_fixed;
    if (p.min_n >= 0 && n <= p.min_n) return;          // small_nms_kernel's image (uniform over the cluster, no DSMEM touched yet)
    if (n > RPNC_MAXN || n <= 0) {                      // sort_nms_kernel's image
        if (crank == 0 && tid == 0) q.fallback[b] = 1;
        return;
    }
    const size_t off = (size_t)b * p.cap;
    // ---- slice sort (payload = slot) + merge ranks
    const int m = (n + CL - 1) / CL;
    const int slo = min(crank * m, n), slen = min(n, slo + m) - slo;
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(dsm);
    uint32_t* sval = reinterpret_cast<uint32_t*>(dsm + q.val_off);
    const int Np = hd_bitonic_padded(m);
    for (int i = tid; i < Np; i += NMS_NT) {
        if (i < slen) {
            const int slot = slo + i;
            const uint32_t tb = p.tiebreak ? (uint32_t)p.tiebreak[off + slot] : (uint32_t)slot;
            skey[i] = ((uint64_t)(~hd_orderable(p.scores[off + slot])) << 32) | tb;
            sval[i] = (uint32_t)slot;
        } else { skey[i] = ~0ull; sval[i] = 0u; }
    }
    __syncthreads();
    if (Np == 2048) hd_cta_bitonic_reg<2, true>(skey, sval);
    else if (Np == 4096) hd_cta_bitonic_reg<4, true>(skey, sval);
    else hd_cta_bitonic_reg<8, true>(skey, sval);
    uint64_t* sorted = q.skeys + off;
    for (int i = tid; i < slen; i += NMS_NT) sorted[slo + i] = skey[i];
    cluster.sync();
    for (int i = tid; i < n; i += NMS_NT) skey[i] = sorted[i];
    __syncthreads();
    const int n_use = (p.max_nms > 0) ? min(n, p.max_nms) : n;
    const int max_det = (p.max_det > 0) ? p.max_det : n_use;
    uint32_t* order = q.order + off;
    float4* sbox = p.sbox + off;
    int* scls = p.scls + off;
    for (int i = tid; i < slen; i += NMS_NT) {
        const unsigned long long v = skey[slo + i];
        int rank = i;
        for (int c = 0; c < CL; ++c) {
            if (c == crank) continue;
            int a = min(c * m, n), z = min(n, a + m);
            const int a0 = a;
            while (a < z) { const int mid = (a + z) >> 1; if (skey[mid] < v) a = mid + 1; else z = mid; }
            rank += a - a0;
        }
        const uint32_t slot = sval[i];
        order[rank] = slot;
        if (rank < n_use) {
            float4 bx = p.boxes[off + slot];
            const int c = p.cls ? p.cls[off + slot] : 0;
            if (p.class_mode == HD_NMS_CLASS_OFFSET) {
                const float o = __fmul_rn((float)c, p.offset_scale);
                bx.x = __fadd_rn(bx.x, o); bx.y = __fadd_rn(bx.y, o); bx.z = __fadd_rn(bx.z, o); bx.w = __fadd_rn(bx.w, o);
            }
            sbox[rank] = bx;
            scls[rank] = (p.class_mode == HD_NMS_CLASS_EXACT) ? c : 0;
        }
    }
    cluster.sync();
    HdClWs w;
    w.sbox = sbox; w.scls = scls; w.gbox = q.gbox + off; w.grank = q.grank + off; w.adj_cnt = q.adj_cnt + off;
    w.adj = q.adj + (size_t)b * q.capn * RPNC_ADJ; w.fallback = q.fallback + b; w.keep_r = p.keep_r + off;
    const int kc = (p.class_mode == HD_NMS_CLASS_EXACT) ? hd_cluster_greedy_nms<NMS_NT, true>(cluster, dsm, q.lay, csm, w, n_use, max_det, p.thr)
                                                        : hd_cluster_greedy_nms<NMS_NT, false>(cluster, dsm, q.lay, csm, w, n_use, max_det, p.thr);
    if (crank != 0 || kc < 0) return;
    const int* keep_r = w.keep_r;
    for (int qi = tid; qi < kc; qi += NMS_NT) {
        const int r = keep_r[qi];
        const uint32_t slot = order[r];
        if (p.out_det) {
            const float4 bx = p.boxes[off + slot];
            const float sc = p.scores[off + slot], cf = p.cls ? (float)p.cls[off + slot] : 0.0f;
            const size_t ro = ((size_t)b * p.max_det + qi) * 6;
            float* o = p.out_det + ro;
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = sc; o[5] = cf;
            for (int rr = 0; rr < p.rep.n; ++rr) {   // posted stores into the peers' gather buffers
                float* pr = p.rep.det[rr] + ro;
                pr[0] = bx.x; pr[1] = bx.y; pr[2] = bx.z; pr[3] = bx.w; pr[4] = sc; pr[5] = cf;
            }
        }
        if (p.out_idx) p.out_idx[(size_t)b * p.max_det + qi] = p.tiebreak ? (long long)p.tiebreak[off + slot] : (long long)slot;
    }
    if (tid == 0) p.out_count[b] = kc;
    if (tid < p.rep.n) p.rep.cnt[tid][b] = kc;
    hd_zero_tail(p.out_det, p.out_idx, p.rep, p.max_det, b, kc);
}

// small-image variant: 256 threads, everything in shared memory, several images per SM at once
__global__ void __launch_bounds__(HD_SMALL_NT, 2) small_nms_kernel(const __grid_constant__ NmsParams p, const __grid_constant__ HdNmsTail q) {
    __shared__ HdSmallSmem ssm;
    const int b = blockIdx.x;
    const int n = p.counts ? min(__ldcg(p.counts + b), p.cap) : p.n_fixed;
    hd_small_nms_image(ssm, q, b, p.cap, n, p.boxes, p.scores, p.cls, p.tiebreak);
}

// ------------------------------------------------------------------------------------------------ box_iou
__global__ void __launch_bounds__(256) box_iou_kernel(const float4* __restrict__ b1, long long N, const float4* __restrict__ b2,
                                                      long long M, float* __restrict__ iou) {
    // block = 8 rows x 256 columns tile; thread handles one column for 8 rows (column-contiguous stores)
    __shared__ float4 rows[8];
    __shared__ float rarea[8];
    const long long r0 = (long long)blockIdx.y * 8;
    const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
    if (threadIdx.x < 8 && r0 + threadIdx.x < N) {
        float4 a = b1[r0 + threadIdx.x];
        rows[threadIdx.x] = a;
        rarea[threadIdx.x] = hd_area(a);
    }
    __syncthreads();
    if (c >= M) return;
    const float4 bb = b2[c];
    const float ab = hd_area(bb);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (r0 + k < N) {
            const float4 a = rows[k];
            // torch: lt = max(b1[:, None, :2], b2[:, :2]); wh = (rb - lt).clamp(min=0)
            float w = fmaxf(__fsub_rn(fminf(a.z, bb.z), fmaxf(a.x, bb.x)), 0.0f);
            float h = fmaxf(__fsub_rn(fminf(a.w, bb.w), fmaxf(a.y, bb.y)), 0.0f);
            float inter = __fmul_rn(w, h);
            iou[(r0 + k) * M + c] = __fdiv_rn(inter, __fsub_rn(__fadd_rn(rarea[k], ab), inter));
        }
    }
}

// ------------------------------------------------------------------------------------------------ host
// ---- "did a recent call have a large image?"  One mapped host word per device: the large-path kernels
// store the call id there (a posted write over PCIe), the host reads it at the next call without synchronising.  The answer only
// selects between kernels that return identical bits, so a stale value costs time, never correctness.
#include <mutex>
#define NMS_HINT_SLOTS 256
static unsigned* g_hint_host[HD_MAX_DEVICES];
static unsigned* g_hint_dev[HD_MAX_DEVICES];
static unsigned g_hint_calls[HD_MAX_DEVICES][NMS_HINT_SLOTS];
static std::mutex g_hint_mutex;
static bool nms_hint_slot(const void* workspace, cudaStream_t st, unsigned** dev_word, unsigned* call_id, bool* large_recent) {
    const int d = hd_current_device();
    if (d < 0 || d >= HD_MAX_DEVICES) return false;
    std::lock_guard<std::mutex> lock(g_hint_mutex);
    if (!g_hint_host[d]) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return false; }
        void* h = nullptr; void* dv = nullptr;
        if (cudaHostAlloc(&h, NMS_HINT_SLOTS * sizeof(unsigned), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
            cudaHostGetDevicePointer(&dv, h, 0) != cudaSuccess) { cudaGetLastError(); return false; }
        memset(h, 0, NMS_HINT_SLOTS * sizeof(unsigned));
        for (int i = 0; i < NMS_HINT_SLOTS; ++i) g_hint_calls[d][i] = 1000u;
        g_hint_dev[d] = (unsigned*)dv;
        g_hint_host[d] = (unsigned*)h;
    }
    // one word per device: keying it by the workspace address looked finer-grained but missed exactly when it mattered -- the
    // warm-up calls before a CUDA-graph capture and the captured call do not always see the same workspace address (a first call
    // that allocates, a capture-time memory pool), so a dense workload could freeze the slow 256-thread kernel into its graph
    (void)workspace;
    const unsigned slot = 0u;
    const unsigned last_big = *reinterpret_cast<volatile unsigned*>(g_hint_host[d] + slot);
    const unsigned prev = g_hint_calls[d][slot];
    *large_recent = (prev - last_big) < 64u;         // one of the last 64 calls on this device reported a large image
    *call_id = ++g_hint_calls[d][slot];
    *dev_word = g_hint_dev[d] + slot;
    return true;
}

// developer/test knob (process-wide, atomic): 0 auto, 1 one CTA per image only, 2 cluster kernel whenever the batch allows,
// 3 light (256-thread) large-image kernel always, 4 heavy (1024-thread, no clusters unless the batch allows) always,
// 5 one CTA per image with the bitonic network for every size (A/B of the bucket sort)
static int g_nms_mode = 0;
extern "C" HD_API int hd_nms_set_mode(int mode) { return __atomic_exchange_n(&g_nms_mode, mode, __ATOMIC_ACQ_REL); }
static int nms_cluster_capacity(int CL) {
    static int cached_dev[HD_MAX_DEVICES][RPNC_MAXCL + 1] = {{0}};   // per device; racing writers store the same value
    const int dev = hd_current_device();
    int* cached = cached_dev[(dev >= 0 && dev < HD_MAX_DEVICES) ? dev : 0];
    if (cached[CL]) return cached[CL];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL * 64); cfg.blockDim = dim3(NMS_NT); cfg.dynamicSmemBytes = 180 * 1024;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int nc = 0;
    cudaFuncSetAttribute(sort_nms_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (cudaOccupancyMaxActiveClusters(&nc, sort_nms_cluster_kernel, &cfg) != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = hd_num_sms() / CL; }
    return cached[CL] = nc;
}
// CTAs per image for a batch of B images: the largest cluster (8 or 4) that still runs the whole batch in one wave; 0 = the
// batch is large enough for one CTA per image (or the mode forbids clusters)
static int nms_cluster_size(int B, bool for_layout = false) {
    const int mode = __atomic_load_n(&g_nms_mode, __ATOMIC_ACQUIRE);
    if (((mode == 1 || mode == 5) && !for_layout) || B <= 0) return 0;
    for (int c = RPNC_MAXCL; c >= 4; c >>= 1)     // clusters of 2 measured no faster than one CTA per image (cfg4, B=64)
        if (B <= nms_cluster_capacity(c)) return c;
    return 0;
}
static void nms_ws_layout(int B, int cap, size_t* offs, size_t* total) {
    size_t n = (size_t)B * cap, o = 0;
    offs[0] = o; o = hd_align_up(o + n * 8, 256);   // k0
    offs[1] = o; o = hd_align_up(o + n * 8, 256);   // k1
    offs[2] = o; o = hd_align_up(o + n * 4, 256);   // v0
    offs[3] = o; o = hd_align_up(o + n * 4, 256);   // v1
    offs[4] = o; o = hd_align_up(o + n * 16, 256);  // sbox
    offs[5] = o; o = hd_align_up(o + n * 4, 256);   // scls
    offs[6] = o; o = hd_align_up(o + n * 4, 256);   // keep_r
    offs[7] = o; o = hd_align_up(o + n * 16, 256);  // grid records
    // cluster kernel only (batches small enough to leave SMs idle)
    const size_t nc = nms_cluster_size(B, true) ? n : 0;   // (independent of hd_nms_set_mode: the size query and the call must agree)
    offs[8] = o; o = hd_align_up(o + nc * 4, 256);   // bucket ranks
    offs[9] = o; o = hd_align_up(o + nc * 4, 256);   // adjacency counts
    offs[10] = o; o = hd_align_up(o + (nc ? (size_t)B * (size_t)(cap < RPNC_MAXN ? cap : RPNC_MAXN) * RPNC_ADJ * 2 : 0), 256);   // adjacency lists, stride min(cap, RPNC_MAXN)
    offs[11] = o; o = hd_align_up(o + (size_t)B * 4, 256);   // hand-back flags
    *total = o;
}

extern "C" HD_API size_t hd_sort_nms_workspace_size(int B, int cap) {
    size_t offs[12], total;
    nms_ws_layout(B < 0 ? 0 : B, cap < 0 ? 0 : cap, offs, &total);
    return total + 256;
}

int hd_sort_nms_batched_min(const float* boxes, const float* scores, const int32_t* cls, const int32_t* tiebreak,
                            const int32_t* counts, int n_fixed, int B, int cap, double iou_thres, int class_mode,
                            float offset_scale, int max_nms, int max_det, float* out_det, int64_t* out_idx,
                            int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream, int min_n,
                            const hd_replicas* replicas);

extern "C" HD_API int hd_sort_nms_batched_replicated(const float* boxes, const float* scores, const int32_t* cls, const int32_t* tiebreak,
                                                     const int32_t* counts, int n_fixed, int B, int cap, double iou_thres, int class_mode,
                                                     float offset_scale, int max_nms, int max_det, float* out_det, int64_t* out_idx,
                                                     int32_t* out_count, const hd_replicas* replicas, void* workspace,
                                                     size_t workspace_bytes, void* stream) {
    return hd_sort_nms_batched_min(boxes, scores, cls, tiebreak, counts, n_fixed, B, cap, iou_thres, class_mode, offset_scale, max_nms,
                                   max_det, out_det, out_idx, out_count, workspace, workspace_bytes, stream, -1, replicas);
}

extern "C" HD_API int hd_sort_nms_batched(const float* boxes, const float* scores, const int32_t* cls, const int32_t* tiebreak,
                                          const int32_t* counts, int n_fixed, int B, int cap, double iou_thres, int class_mode,
                                          float offset_scale, int max_nms, int max_det, float* out_det, int64_t* out_idx,
                                          int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
    return hd_sort_nms_batched_min(boxes, scores, cls, tiebreak, counts, n_fixed, B, cap, iou_thres, class_mode, offset_scale, max_nms,
                                   max_det, out_det, out_idx, out_count, workspace, workspace_bytes, stream, -1, nullptr);
}

int hd_sort_nms_batched_min(const float* boxes, const float* scores, const int32_t* cls, const int32_t* tiebreak,
                            const int32_t* counts, int n_fixed, int B, int cap, double iou_thres, int class_mode,
                            float offset_scale, int max_nms, int max_det, float* out_det, int64_t* out_idx,
                            int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream, int min_n,
                            const hd_replicas* replicas) {
    HD_CHECK_ARG(replicas == nullptr || (replicas->n >= 0 && replicas->n <= HD_MAX_REPLICAS), "replicas->n out of [0,%d]", HD_MAX_REPLICAS);
    HD_CHECK_ARG(B >= 0 && cap >= 0, "bad shape B=%d cap=%d", B, cap);
    if (B == 0) return HD_OK;
    HD_CHECK_ARG(out_count != nullptr, "out_count is NULL");
    HD_CHECK_ARG(class_mode >= 0 && class_mode <= 2, "class_mode must be 0,1,2, got %d", class_mode);
    HD_CHECK_ARG(class_mode == HD_NMS_AGNOSTIC || cls != nullptr, "cls is NULL for a class-aware mode");
    HD_CHECK_ARG(counts != nullptr || (n_fixed >= 0 && n_fixed <= cap), "n_fixed=%d out of [0,cap=%d]", n_fixed, cap);
    HD_CHECK_ARG(max_det > 0 || (out_det == nullptr && out_idx == nullptr) || true, "max_det");
    cudaStream_t st = (cudaStream_t)stream;
    if (cap == 0) {
        HD_CUDA_CALL(cudaMemsetAsync(out_count, 0, sizeof(int) * (size_t)B, st));
        return HD_OK;
    }
    HD_CHECK_ARG(boxes && scores, "boxes/scores is NULL");
    HD_CHECK_ARG(max_det > 0, "max_det must be > 0 (row stride of out_det/out_idx), got %d", max_det);
    size_t offs[12], total;
    nms_ws_layout(B, cap, offs, &total);
    uintptr_t w0 = hd_align_up((uintptr_t)workspace, 256);
    if (!workspace || w0 + total > (uintptr_t)workspace + workspace_bytes)
        HD_FAIL(HD_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", total + 256, workspace_bytes);
    NmsParams p;
    p.boxes = (const float4*)boxes; p.scores = scores; p.cls = (class_mode == HD_NMS_AGNOSTIC) ? cls : cls; p.tiebreak = tiebreak;
    p.counts = counts; p.n_fixed = n_fixed; p.B = B; p.cap = cap; p.min_n = min_n;
    p.rep.n = 0;
    if (replicas) {
        p.rep.n = replicas->n;
        for (int r = 0; r < replicas->n; ++r) {
            HD_CHECK_ARG(replicas->det[r] && replicas->count[r], "replica %d has a NULL pointer", r);
            p.rep.det[r] = replicas->det[r]; p.rep.cnt[r] = replicas->count[r];
        }
    }
    p.thr = hd_thr_floor(iou_thres);
    p.class_mode = class_mode; p.offset_scale = offset_scale; p.max_nms = max_nms; p.max_det = max_det;
    p.out_det = out_det; p.out_idx = (long long*)out_idx; p.out_count = out_count;
    p.k0 = (uint64_t*)(w0 + offs[0]); p.k1 = (uint64_t*)(w0 + offs[1]);
    p.v0 = (uint32_t*)(w0 + offs[2]); p.v1 = (uint32_t*)(w0 + offs[3]);
    p.sbox = (float4*)(w0 + offs[4]); p.scls = (int*)(w0 + offs[5]); p.keep_r = (int*)(w0 + offs[6]); p.gitem = (float4*)(w0 + offs[7]);
    size_t words = ((size_t)(cap + 31) / 32 + 4 + 1) & ~(size_t)1;   // 8-byte aligned end
    HD_CHECK_ARG(words * 4 <= 64 * 1024, "cap=%d too large for the shared-memory removed bitmap", cap);
    // shared-memory sort area for segments of up to 8192 candidates: (u64 key, u32 slot)
    p.sort_off = (int)words;
    p.bitonic_cap = 8192;
    size_t smem = words * 4 + (size_t)p.bitonic_cap * 12;
    HD_ENSURE_SMEM(sort_nms_kernel<1024>, 186 * 1024);
    // heavy or light large-image kernels?  (auto: light until a recent call through this workspace saw a large image)
    const int mode = __atomic_load_n(&g_nms_mode, __ATOMIC_ACQUIRE);
    bool heavy = true;
    p.hint = nullptr; p.call_id = 0;
    if (counts != nullptr && min_n < 0) {   // variable-length segments: worth remembering what the data looks like
        bool recent = true;
        if (nms_hint_slot(workspace, st, &p.hint, &p.call_id, &recent)) heavy = recent;
    }
    if (mode == 3) heavy = false;
    else if (mode != 0) heavy = true;
    p.bucket_sort = (mode != 5);   // mode 5: heavy kernel with the bitonic network for every size (A/B of the bucket sort)
    if (min_n < 0 && (counts != nullptr || n_fixed <= HD_SMALL_N)) {
        // images with <= HD_SMALL_N candidates: shared-memory kernel; the radix-sort kernel below then only
        // works on the larger ones (it returns at once for the rest)
        HdNmsTail q;
        q.thr = p.thr; q.class_mode = class_mode; q.offset_scale = offset_scale; q.max_nms = max_nms; q.max_det = max_det;
        q.out_det = out_det; q.out_idx = (long long*)out_idx; q.out_count = out_count;
        q.rep = p.rep;
        small_nms_kernel<<<B, HD_SMALL_NT, 0, st>>>(p, q);
        HD_CUDA_LAUNCH_CHECK("small_nms_kernel");
        p.min_n = HD_SMALL_N;
        if (counts == nullptr) return HD_OK;
    }
    p.only = nullptr;
    if (!heavy) {
        // light variant: 256 threads, radix sort through the global scratch, 1024-bucket pruning grid; co-resident with other kernels
        p.bitonic_cap = 0;
        HD_ENSURE_SMEM(sort_nms_kernel<256>, 80 * 1024);   // the removed-bitmap of a large cap exceeds the 48 KB default
        sort_nms_kernel<256><<<B, 256, words * 4, st>>>(p);
        HD_CUDA_LAUNCH_CHECK("sort_nms_kernel<256>");
        return HD_OK;
    }
    const int CL = nms_cluster_size(B);
    if (CL) {
        NmsClParams q;
        q.s = p;
        q.skeys = p.k1; q.svals = p.v1; q.order = p.v0; q.gbox = p.gitem;
        q.grank = (int*)(w0 + offs[8]); q.adj_cnt = (int*)(w0 + offs[9]); q.adj = (unsigned short*)(w0 + offs[10]); q.fallback = (int*)(w0 + offs[11]);
        const int capn = cap < RPNC_MAXN ? cap : RPNC_MAXN;
        q.capn = capn;
        const int mslice = (capn + CL - 1) / CL;
        const size_t np = mslice <= 2048 ? 2048 : (mslice <= 4096 ? 4096 : 8192);
        size_t keys_b = (np > (size_t)capn ? np : (size_t)capn) * 8;
        q.val_off = (int)keys_b;
        size_t sm_c = keys_b + np * 4;
        hd_cluster_layout(capn, 180 * 1024, &q.lay, &sm_c);
        HD_CUDA_CALL(cudaMemsetAsync(q.fallback, 0, (size_t)B * 4, st));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(B * CL)); cfg.blockDim = dim3(NMS_NT); cfg.dynamicSmemBytes = sm_c; cfg.stream = st;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        HD_CUDA_CALL(cudaLaunchKernelEx(&cfg, sort_nms_cluster_kernel, q));
        hd_count_launch();
        p.only = q.fallback;
    }
    sort_nms_kernel<1024><<<B, NMS_NT, smem, st>>>(p);
    HD_CUDA_LAUNCH_CHECK("sort_nms_kernel");
    return HD_OK;
}

extern "C" HD_API int hd_box_iou(const float* boxes1, int64_t N, const float* boxes2, int64_t M, float* iou, void* stream) {
    HD_CHECK_ARG(N >= 0 && M >= 0, "bad shape N=%lld M=%lld", (long long)N, (long long)M);
    if (N == 0 || M == 0) return HD_OK;
    HD_CHECK_ARG(boxes1 && boxes2 && iou, "null pointer");
    long long gx = (M + 255) / 256, gy = (N + 7) / 8;
    HD_CHECK_ARG(gy <= 65535 * 1ll || true, "N too large");
    // grid.y is limited to 65535: loop in slabs
    for (long long y0 = 0; y0 < gy; y0 += 65535) {
        long long ny = (gy - y0 < 65535) ? (gy - y0) : 65535;
        dim3 grid((unsigned)gx, (unsigned)ny);
        box_iou_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)boxes1 + y0 * 8, N - y0 * 8, (const float4*)boxes2, M,
                                                              iou + y0 * 8 * M);
    }
    HD_CUDA_LAUNCH_CHECK("box_iou_kernel");
    return HD_OK;
}

HD_DEFINE_PHASE_READER(hd_phase_reader_nms)
int hd_phase_reader_small(long long* out16) { HD_CUDA_CALL(cudaDeviceSynchronize()); HD_CUDA_CALL(cudaMemcpyFromSymbol(out16, hd_dbg_small, sizeof(long long) * 16)); return HD_OK; }

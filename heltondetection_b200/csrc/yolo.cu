// YOLOv5 head kernels: dense decode (a1), fused decode+filter+compaction (a1+a2), filter on decoded pred (a2).
// Reference feature: README.md:9; semantics SURVEY.md A.1/A.2 (ultralytics / bubbliiiing lineage, README.md:158-162).
#include "hd_common.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

// head element types: fp32 (default) or, with HD_FLAG_IN_F16 / HD_FLAG_IN_BF16, 16-bit heads that are widened to fp32 on
// load (exact) -- half the HBM bytes of the dominant read; all arithmetic stays fp32
struct HdF16 { unsigned short v; };
struct HdBF16 { unsigned short v; };
__device__ __forceinline__ float hd_widen(float x) { return x; }
__device__ __forceinline__ float hd_widen(HdF16 x) { return __half2float(__ushort_as_half(x.v)); }
__device__ __forceinline__ float hd_widen(HdBF16 x) { return __uint_as_float((unsigned)x.v << 16); }
template <typename T> __device__ __forceinline__ float hd_load1(const T* q) { return hd_widen(*q); }
template <> __device__ __forceinline__ float hd_load1<float>(const float* q) { return hd_ldg_stream(q); }
// four consecutive elements with one streaming load (128-bit for fp32, 64-bit for the 16-bit types)
template <typename T> __device__ __forceinline__ void hd_load4(const T* q, float* v) {
    unsigned lo, hi;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(q));
    T e;
    e.v = (unsigned short)(lo & 0xffffu); v[0] = hd_widen(e); e.v = (unsigned short)(lo >> 16); v[1] = hd_widen(e);
    e.v = (unsigned short)(hi & 0xffffu); v[2] = hd_widen(e); e.v = (unsigned short)(hi >> 16); v[3] = hd_widen(e);
}
template <> __device__ __forceinline__ void hd_load4<float>(const float* q, float* v) {
    const float4 w = hd_ldg_stream4(q);
    v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
}
// the same four elements kept packed until they are consumed: 16-bit heads then hold twice the planes in flight per register
template <typename T> struct HdRaw4 { unsigned lo, hi; };
template <> struct HdRaw4<float> { float4 w; };
template <typename T> __device__ __forceinline__ HdRaw4<T> hd_load_raw4(const T* q) {
    HdRaw4<T> r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.lo), "=r"(r.hi) : "l"(q));
    return r;
}
template <> __device__ __forceinline__ HdRaw4<float> hd_load_raw4<float>(const float* q) { HdRaw4<float> r; r.w = hd_ldg_stream4(q); return r; }
template <typename T> __device__ __forceinline__ void hd_unpack4(const HdRaw4<T>& r, float* v) {
    T e;
    e.v = (unsigned short)(r.lo & 0xffffu); v[0] = hd_widen(e); e.v = (unsigned short)(r.lo >> 16); v[1] = hd_widen(e);
    e.v = (unsigned short)(r.hi & 0xffffu); v[2] = hd_widen(e); e.v = (unsigned short)(r.hi >> 16); v[3] = hd_widen(e);
}
template <> __device__ __forceinline__ void hd_unpack4<float>(const HdRaw4<float>& r, float* v) { v[0] = r.w.x; v[1] = r.w.y; v[2] = r.w.z; v[3] = r.w.w; }

struct YoloParams {
    const void* data[HD_MAX_LEVELS];
    int HW[HD_MAX_LEVELS];
    int W[HD_MAX_LEVELS];
    float stride[HD_MAX_LEVELS];
    float anchor[HD_MAX_LEVELS][2 * HD_MAX_ANCHORS];
    int tile_start[HD_MAX_LEVELS + 1];  // cumulative 128-cell tiles per (image, anchor)
    int level_off[HD_MAX_LEVELS + 1];   // flat anchor index offset of each level
    int n_levels, B, A, nc, no;
    float thr;   // (float) conf_thres: torch compares the fp32 tensor with the scalar cast to fp32
    float gate;  // conservative logit-domain pre-test for sigmoid(obj) > thr
    float gate2; // ... and for sigmoid(obj)*sigmoid(cls) > thr: min(obj,0) + min(cls,0) > ln(thr) - margin
    int ge, dense, cap;
    int multi;   // HD_FLAG_MULTI_LABEL: every (anchor, class) pair with obj*cls > thr is a candidate (ultralytics multi_label)
    int items_per_image;
    long long total_items;
};

// ------------------------------------------------------------------------------------------------
// Fused decode + filter + compaction.  One warp owns a tile of 128 consecutive cells of one
// (image, level, anchor): lane k holds cells 4k..4k+3, so every plane is read with one coalesced
// 512-byte request (128-bit per lane).  Channels live in planes H*W apart (the head's native NCHW),
// which is why the warp walks planes instead of rows.
//   - objectness plane first: if no cell of the tile can pass sigmoid(obj) > thr the 84 remaining
//     planes are skipped (exact, since cls <= 1 => obj*cls <= obj) unless HD_FLAG_DENSE_READ;
//   - class max is taken on the logits (sigmoid and x*obj are monotone), so only survivors pay for
//     expf; the one case where that could differ from torch.max over the products -- an earlier class
//     whose product rounds to the same float -- is detected through the runner-up L and resolved by
//     an exact rescan;
//   - survivors are compacted with one atomicAdd per warp.
// ------------------------------------------------------------------------------------------------
template <bool VEC, typename T>
__device__ __forceinline__ void yolo_decode_item_packed(const YoloParams& p, const long long item, const int lane, float* stage,
                                                 float4* __restrict__ cand_box, float* __restrict__ cand_score,
                                                 int* __restrict__ cand_cls, int* __restrict__ cand_anchor,
                                                 int* __restrict__ cand_count) {
    const int b = (int)(item / p.items_per_image);
    int r = (int)(item - (long long)b * p.items_per_image);
    int l = 0;
#pragma unroll
    for (int q = 1; q < HD_MAX_LEVELS; ++q)
        if (q < p.n_levels && r >= p.A * p.tile_start[q]) l = q;
    r -= p.A * p.tile_start[l];
    const int tiles_l = p.tile_start[l + 1] - p.tile_start[l];
    const int a = r / tiles_l;
    const int t = r - a * tiles_l;
    const int HW = p.HW[l];
    const int cell0 = t * 128 + lane * 4;
    const T* __restrict__ base = reinterpret_cast<const T*>(p.data[l]) + ((size_t)(b * p.A + a) * p.no) * HW + cell0;

    float o[4], bx[4][4], m[4], L[4];
    int j[4];
    bool valid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) valid[k] = (cell0 + k) < HW;

    bool need = true;  // lane fetches the non-objectness planes (narrowed below in sparse mode)
    auto load4 = [&](int plane, float* v) {
        const T* q = base + (size_t)plane * HW;
        if (VEC) {
            if (valid[0] && need) hd_load4<T>(q, v);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (valid[k] && need) v[k] = hd_load1<T>(q + k);
        }
    };

#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) bx[c][k] = 0.0f;
    }
    // U independent class-plane loads in flight per lane (same bytes in flight for 32-bit and 16-bit heads)
    constexpr int U = (sizeof(T) == 4) ? 8 : 16;
    constexpr int HEAD = 3;
    // dense read: nothing depends on the objectness test, so the objectness plane, the four box planes and the first nc % U class
    // planes (when there are at most HEAD of them) leave as ONE batch of independent loads -- a 15-plane head (nc = 10) costs two
    // memory round trips per tile instead of five
    const int t_head = VEC ? p.nc % U : 0;
    const bool merged = VEC && p.dense && t_head <= HEAD;
    HdRaw4<T> head_raw[HEAD];
    load4(4, o);
    if (merged) {
#pragma unroll
        for (int c = 0; c < 4; ++c) load4(c, bx[c]);
#pragma unroll
        for (int u = 0; u < HEAD; ++u)
            if (valid[0] && u < t_head) head_raw[u] = hd_load_raw4<T>(base + (size_t)(5 + u) * HW);
    }
    bool gate_any = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) gate_any |= valid[k] && (o[k] > p.gate);
    if (!p.dense && !__any_sync(HD_FULL, gate_any)) return;
    // sparse mode: only lanes that own a possible survivor touch the other 84 planes, so the traffic of a
    // surviving tile shrinks from 85 x 512 B to 85 x (one 32-byte sector per surviving lane)
    need = p.dense || gate_any;

    if (!merged) {
#pragma unroll
        for (int c = 0; c < 4; ++c) load4(c, bx[c]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { m[k] = -INFINITY; L[k] = -INFINITY; j[k] = 0; }

    auto upd = [&](const float* v, int c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bool g = v[k] > m[k];
            L[k] = g ? m[k] : L[k];
            j[k] = g ? c : j[k];
            m[k] = g ? v[k] : m[k];
        }
    };
    int c = 0;
    if (VEC) {
        const bool ld = valid[0] && need;
        if (merged) {
#pragma unroll
            for (int u = 0; u < HEAD; ++u)
                if (u < t_head) {
                    float v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    if (ld) hd_unpack4<T>(head_raw[u], v);
                    upd(v, u);
                }
            c = t_head;
        }
        for (; c + U <= p.nc; c += U) {
            HdRaw4<T> raw[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ld) raw[u] = hd_load_raw4<T>(base + (size_t)(5 + c + u) * HW);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                if (ld) hd_unpack4<T>(raw[u], v);
                upd(v, c + u);
            }
        }
        if (c < p.nc) {   // the last nc % U planes, again as one batch
            HdRaw4<T> raw[U];
#pragma unroll
            for (int u = 0; u < U - 1; ++u)
                if (ld && c + u < p.nc) raw[u] = hd_load_raw4<T>(base + (size_t)(5 + c + u) * HW);
#pragma unroll
            for (int u = 0; u < U - 1; ++u)
                if (c + u < p.nc) {
                    float v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    if (ld) hd_unpack4<T>(raw[u], v);
                    upd(v, c + u);
                }
            c = p.nc;
        }
    } else {
        for (; c + 8 <= p.nc; c += 8) {
            float v[8][4];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[u][k] = -INFINITY;
                load4(5 + c + u, v[u]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) upd(v[u], c + u);
        }
        for (; c < p.nc; ++c) {
            float v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            load4(5 + c, v);
            upd(v, c);
        }
    }

    // ---- survivors.  A dense scene at an evaluation threshold (conf 0.001) has ~9 candidates among the 128 cells of a tile: walking
    // the four cells of every lane would run the sigmoids with 7 % of the lanes active.  Instead a necessary condition in the logit
    // domain picks the cells that may pass -- sigmoid(x) < min(1, e^x), hence sigmoid(o)*sigmoid(m) > thr needs
    // min(o,0) + min(m,0) > ln(thr) -- and those cells are dealt out one per lane through a warp-private staging area, so the exact
    // fp32 test, the box decode and the (now coalesced) stores run with the lanes packed.
    unsigned pm[4];
    int before = 0, mine[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool pre = valid[k] && o[k] > p.gate && __fadd_rn(fminf(o[k], 0.0f), fminf(m[k], 0.0f)) > p.gate2;
        pm[k] = __ballot_sync(HD_FULL, pre);
        mine[k] = pre ? before + __popc(pm[k] & hd_lanemask_lt()) : -1;
        before += __popc(pm[k]);
    }
    const int n_pre = before;
    if (n_pre == 0) return;
    const float s = p.stride[l];
    const float aw = p.anchor[l][2 * a], ah = p.anchor[l][2 * a + 1];
    const int W = p.W[l];
    for (int r0 = 0; r0 < n_pre; r0 += 32) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int q = mine[k] - r0;
            if (q >= 0 && q < 32) {
                stage[0 * 32 + q] = o[k]; stage[1 * 32 + q] = m[k]; stage[2 * 32 + q] = L[k];
                stage[3 * 32 + q] = __int_as_float(j[k] | ((lane * 4 + k) << 16));
                stage[4 * 32 + q] = bx[0][k]; stage[5 * 32 + q] = bx[1][k]; stage[6 * 32 + q] = bx[2][k]; stage[7 * 32 + q] = bx[3][k];
            }
        }
        __syncwarp();
        bool ok = false;
        float cf = 0.0f, b0 = 0.0f, b1 = 0.0f, b2 = 0.0f, b3 = 0.0f;
        int jj = 0, cell_in = 0;
        if (r0 + lane < n_pre) {
            const float oo = stage[0 * 32 + lane], mm = stage[1 * 32 + lane], LL = stage[2 * 32 + lane];
            const int pk = __float_as_int(stage[3 * 32 + lane]);
            jj = pk & 0xffff; cell_in = pk >> 16;
            b0 = stage[4 * 32 + lane]; b1 = stage[5 * 32 + lane]; b2 = stage[6 * 32 + lane]; b3 = stage[7 * 32 + lane];
            const float po = hd_sigmoid(oo);
            cf = __fmul_rn(hd_sigmoid(mm), po);
            ok = p.ge ? (po >= p.thr && cf >= p.thr) : (po > p.thr && cf > p.thr);
            if (ok && LL > -INFINITY && __fmul_rn(hd_sigmoid(LL), po) == cf) {
                // an earlier class ties after rounding: torch.max returns the first maximal product
                const T* q = base - lane * 4 + cell_in;
                for (int cc = 0; cc < jj; ++cc) {
                    float lg = hd_load1<T>(q + (size_t)(5 + cc) * HW);
                    if (__fmul_rn(hd_sigmoid(lg), po) == cf) { jj = cc; break; }
                }
            }
        }
        // warp-aggregated slot claim
        const unsigned okm = __ballot_sync(HD_FULL, ok);
        if (okm == 0u) continue;
        int slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(cand_count + b, __popc(okm));
        slot0 = __shfl_sync(HD_FULL, slot0, 0);
        const int slot = slot0 + __popc(okm & hd_lanemask_lt());
        if (ok && slot < p.cap) {
            const int cell = t * 128 + cell_in;
            const int gi = cell / W, gj = cell - gi * W;
            // (p*2 - 0.5 + grid) * s ; (p*2)^2 * anchor : fp32, op order of the reference decode
            float px = __fmul_rn(hd_sigmoid(b0), 2.0f), py = __fmul_rn(hd_sigmoid(b1), 2.0f);
            float pw = __fmul_rn(hd_sigmoid(b2), 2.0f), ph = __fmul_rn(hd_sigmoid(b3), 2.0f);
            float cx = __fmul_rn(__fadd_rn(__fsub_rn(px, 0.5f), (float)gj), s);
            float cy = __fmul_rn(__fadd_rn(__fsub_rn(py, 0.5f), (float)gi), s);
            float w = __fmul_rn(__fmul_rn(pw, pw), aw), h = __fmul_rn(__fmul_rn(ph, ph), ah);
            float hw2 = __fmul_rn(w, 0.5f), hh2 = __fmul_rn(h, 0.5f);  // w/2 exact
            size_t g = (size_t)b * p.cap + slot;
            cand_box[g] = make_float4(__fsub_rn(cx, hw2), __fsub_rn(cy, hh2), __fadd_rn(cx, hw2), __fadd_rn(cy, hh2));
            cand_score[g] = cf;
            cand_cls[g] = jj;
            cand_anchor[g] = p.level_off[l] + a * HW + cell;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Lean variant of the tile walk above for deployment thresholds (conf >= 0.05 and nc >= 16): survivors are rare, so each lane
// simply walks its own four cells, and nothing but the plane loop holds registers -- on the 85-plane cfg2 head this form runs
// at the copy bandwidth (0.311 ms per 256 images), 6 % faster than the packed variant, whose staging bookkeeping costs six spilled
// registers inside the plane loop.  Same candidates (the order inside an image is arbitrary in both).
// ------------------------------------------------------------------------------------------------
template <bool VEC, typename T>
__device__ __forceinline__ void yolo_decode_item(const YoloParams& p, const long long item, const int lane,
                                                 float4* __restrict__ cand_box, float* __restrict__ cand_score,
                                                 int* __restrict__ cand_cls, int* __restrict__ cand_anchor,
                                                 int* __restrict__ cand_count) {
    const int b = (int)(item / p.items_per_image);
    int r = (int)(item - (long long)b * p.items_per_image);
    int l = 0;
#pragma unroll
    for (int q = 1; q < HD_MAX_LEVELS; ++q)
        if (q < p.n_levels && r >= p.A * p.tile_start[q]) l = q;
    r -= p.A * p.tile_start[l];
    const int tiles_l = p.tile_start[l + 1] - p.tile_start[l];
    const int a = r / tiles_l;
    const int t = r - a * tiles_l;
    const int HW = p.HW[l];
    const int cell0 = t * 128 + lane * 4;
    const T* __restrict__ base = reinterpret_cast<const T*>(p.data[l]) + ((size_t)(b * p.A + a) * p.no) * HW + cell0;

    float o[4], bx[4][4], m[4], L[4];
    int j[4];
    bool valid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) valid[k] = (cell0 + k) < HW;

    bool need = true;  // lane fetches the non-objectness planes (narrowed below in sparse mode)
    auto load4 = [&](int plane, float* v) {
        const T* q = base + (size_t)plane * HW;
        if (VEC) {
            if (valid[0] && need) hd_load4<T>(q, v);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (valid[k] && need) v[k] = hd_load1<T>(q + k);
        }
    };

#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = -INFINITY;
    load4(4, o);
    bool gate_any = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) gate_any |= valid[k] && (o[k] > p.gate);
    if (!p.dense && !__any_sync(HD_FULL, gate_any)) return;
    // sparse mode: only lanes that own a possible survivor touch the other 84 planes, so the traffic of a
    // surviving tile shrinks from 85 x 512 B to 85 x (one 32-byte sector per surviving lane)
    need = p.dense || gate_any;

#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) bx[c][k] = 0.0f;
        load4(c, bx[c]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { m[k] = -INFINITY; L[k] = -INFINITY; j[k] = 0; }

    auto upd = [&](const float* v, int c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bool g = v[k] > m[k];
            L[k] = g ? m[k] : L[k];
            j[k] = g ? c : j[k];
            m[k] = g ? v[k] : m[k];
        }
    };
    int c = 0;
    if (VEC) {
        // U independent plane loads in flight per lane (same bytes in flight for 32-bit and 16-bit heads)
        constexpr int U = (sizeof(T) == 4) ? 8 : 16;
        const bool ld = valid[0] && need;
        for (; c + U <= p.nc; c += U) {
            HdRaw4<T> raw[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ld) raw[u] = hd_load_raw4<T>(base + (size_t)(5 + c + u) * HW);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                if (ld) hd_unpack4<T>(raw[u], v);
                upd(v, c + u);
            }
        }
    } else {
        constexpr int U = 8;
        for (; c + U <= p.nc; c += U) {
            float v[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[u][k] = -INFINITY;
                load4(5 + c + u, v[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) upd(v[u], c + u);
        }
    }
    for (; c < p.nc; ++c) {
        float v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        load4(5 + c, v);
        upd(v, c);
    }

    // survivors
    float conf[4];
    int npass = 0;
    bool pass[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        pass[k] = false;
        if (valid[k] && o[k] > p.gate) {
            float po = hd_sigmoid(o[k]);
            float cf = __fmul_rn(hd_sigmoid(m[k]), po);
            bool ok = p.ge ? (po >= p.thr && cf >= p.thr) : (po > p.thr && cf > p.thr);
            if (ok) {
                if (L[k] > -INFINITY && __fmul_rn(hd_sigmoid(L[k]), po) == cf) {
                    // an earlier class ties after rounding: torch.max returns the first maximal product
                    const T* q = base + k;
                    for (int cc = 0; cc < j[k]; ++cc) {
                        float lg = hd_load1<T>(q + (size_t)(5 + cc) * HW);
                        if (__fmul_rn(hd_sigmoid(lg), po) == cf) { j[k] = cc; break; }
                    }
                }
                pass[k] = true;
                conf[k] = cf;
                ++npass;
            }
        }
    }
    // warp-aggregated slot claim
    int incl = npass;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(HD_FULL, incl, d);
        if (lane >= d) incl += y;
    }
    int total = __shfl_sync(HD_FULL, incl, 31);
    if (total == 0) return;
    int slot0 = 0;
    if (lane == 31) slot0 = atomicAdd(cand_count + b, total);
    slot0 = __shfl_sync(HD_FULL, slot0, 31);
    int slot = slot0 + incl - npass;
    const float s = p.stride[l];
    const float aw = p.anchor[l][2 * a], ah = p.anchor[l][2 * a + 1];
    const int W = p.W[l];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!pass[k]) continue;
        if (slot < p.cap) {
            const int cell = cell0 + k;
            const int gi = cell / W, gj = cell - gi * W;
            // (p*2 - 0.5 + grid) * s ; (p*2)^2 * anchor : fp32, op order of the reference decode
            float px = __fmul_rn(hd_sigmoid(bx[0][k]), 2.0f), py = __fmul_rn(hd_sigmoid(bx[1][k]), 2.0f);
            float pw = __fmul_rn(hd_sigmoid(bx[2][k]), 2.0f), ph = __fmul_rn(hd_sigmoid(bx[3][k]), 2.0f);
            float cx = __fmul_rn(__fadd_rn(__fsub_rn(px, 0.5f), (float)gj), s);
            float cy = __fmul_rn(__fadd_rn(__fsub_rn(py, 0.5f), (float)gi), s);
            float w = __fmul_rn(__fmul_rn(pw, pw), aw), h = __fmul_rn(__fmul_rn(ph, ph), ah);
            float hw2 = __fmul_rn(w, 0.5f), hh2 = __fmul_rn(h, 0.5f);  // w/2 exact
            size_t g = (size_t)b * p.cap + slot;
            cand_box[g] = make_float4(__fsub_rn(cx, hw2), __fsub_rn(cy, hh2), __fadd_rn(cx, hw2), __fadd_rn(cy, hh2));
            cand_score[g] = conf[k];
            cand_cls[g] = j[k];
            cand_anchor[g] = p.level_off[l] + a * HW + cell;
        }
        ++slot;
    }
}

// ------------------------------------------------------------------------------------------------
// multi_label variant (ultralytics non_max_suppression with multi_label=True, the evaluation setting conf 0.001 / iou 0.6 when
// nc > 1): after `x[:, 5:] *= x[:, 4:5]` EVERY class with obj*cls > conf_thres yields a candidate (box, obj*cls, class), in
// (anchor, class) order.  Same tile walk as above; per group of class planes the lanes count their hits (logit gate first, then the
// exact fp32 product), a warp scan claims the slots with one atomicAdd, and each hit decodes its box.  The candidate's tie-break /
// reported index is anchor * nc + class.
// ------------------------------------------------------------------------------------------------
template <bool VEC, typename T>
__device__ __forceinline__ void yolo_decode_item_multi(const YoloParams& p, const long long item, const int lane,
                                                       float4* __restrict__ cand_box, float* __restrict__ cand_score,
                                                       int* __restrict__ cand_cls, int* __restrict__ cand_anchor,
                                                       int* __restrict__ cand_count) {
    const int b = (int)(item / p.items_per_image);
    int r = (int)(item - (long long)b * p.items_per_image);
    int l = 0;
#pragma unroll
    for (int q = 1; q < HD_MAX_LEVELS; ++q)
        if (q < p.n_levels && r >= p.A * p.tile_start[q]) l = q;
    r -= p.A * p.tile_start[l];
    const int tiles_l = p.tile_start[l + 1] - p.tile_start[l];
    const int a = r / tiles_l;
    const int t = r - a * tiles_l;
    const int HW = p.HW[l];
    const int cell0 = t * 128 + lane * 4;
    const T* __restrict__ base = reinterpret_cast<const T*>(p.data[l]) + ((size_t)(b * p.A + a) * p.no) * HW + cell0;
    bool valid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) valid[k] = (cell0 + k) < HW;
    auto load4 = [&](int plane, float* v, bool need) {
        const T* q = base + (size_t)plane * HW;
        if (VEC) {
            if (valid[0] && need) hd_load4<T>(q, v);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (valid[k] && need) v[k] = hd_load1<T>(q + k);
        }
    };
    float o[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    load4(4, o, true);
    bool live[4];
    float po[4];
    bool any_live = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        live[k] = false; po[k] = 0.0f;
        if (valid[k] && o[k] > p.gate) {
            po[k] = hd_sigmoid(o[k]);
            live[k] = p.ge ? (po[k] >= p.thr) : (po[k] > p.thr);
        }
        any_live |= live[k];
    }
    if (!p.dense && !__any_sync(HD_FULL, any_live)) return;
    const bool need = p.dense || any_live;
    float bx[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int k = 0; k < 4; ++k) bx[c][k] = 0.0f;
        load4(c, bx[c], need);
    }
    const float s = p.stride[l];
    const float aw = p.anchor[l][2 * a], ah = p.anchor[l][2 * a + 1];
    const int W = p.W[l];
    constexpr int U = 8;
    for (int c0 = 0; c0 < p.nc; c0 += U) {
        float v[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[u][k] = -INFINITY;
            if (c0 + u < p.nc) load4(5 + c0 + u, v[u], need);
        }
        // hits of this lane in the group: bit (k * U + u); conf recomputed at emission (hits are rare next to the planes streamed)
        unsigned hits = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (live[k] && v[u][k] > p.gate) {
                    const float cf = __fmul_rn(hd_sigmoid(v[u][k]), po[k]);
                    if (p.ge ? (cf >= p.thr) : (cf > p.thr)) hits |= 1u << (k * U + u);
                }
        if (!__any_sync(HD_FULL, hits != 0u)) continue;
        const int nh = __popc(hits);
        int incl = nh;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(HD_FULL, incl, d); if (lane >= d) incl += y; }
        const int total = __shfl_sync(HD_FULL, incl, 31);
        int slot0 = 0;
        if (lane == 31) slot0 = atomicAdd(cand_count + b, total);
        slot0 = __shfl_sync(HD_FULL, slot0, 31);
        int slot = slot0 + incl - nh;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!((hits >> (k * U)) & ((1u << U) - 1u))) continue;
            const int cell = cell0 + k;
            const int gi = cell / W, gj = cell - gi * W;
            const float px = __fmul_rn(hd_sigmoid(bx[0][k]), 2.0f), py = __fmul_rn(hd_sigmoid(bx[1][k]), 2.0f);
            const float pw = __fmul_rn(hd_sigmoid(bx[2][k]), 2.0f), ph = __fmul_rn(hd_sigmoid(bx[3][k]), 2.0f);
            const float cx = __fmul_rn(__fadd_rn(__fsub_rn(px, 0.5f), (float)gj), s);
            const float cy = __fmul_rn(__fadd_rn(__fsub_rn(py, 0.5f), (float)gi), s);
            const float w = __fmul_rn(__fmul_rn(pw, pw), aw), h = __fmul_rn(__fmul_rn(ph, ph), ah);
            const float hw2 = __fmul_rn(w, 0.5f), hh2 = __fmul_rn(h, 0.5f);
            const float4 box = make_float4(__fsub_rn(cx, hw2), __fsub_rn(cy, hh2), __fadd_rn(cx, hw2), __fadd_rn(cy, hh2));
            const int anchor = p.level_off[l] + a * HW + cell;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!((hits >> (k * U + u)) & 1u)) continue;
                if (slot < p.cap) {
                    const size_t g = (size_t)b * p.cap + slot;
                    cand_box[g] = box;
                    cand_score[g] = __fmul_rn(hd_sigmoid(v[u][k]), po[k]);
                    cand_cls[g] = c0 + u;
                    cand_anchor[g] = anchor * p.nc + c0 + u;
                }
                ++slot;
            }
        }
    }
}

template <bool VEC, typename T = float>
__global__ void __launch_bounds__(256, 4) yolo_decode_filter_kernel(const __grid_constant__ YoloParams p,
                                                                 float4* __restrict__ cand_box,
                                                                 float* __restrict__ cand_score,
                                                                 int* __restrict__ cand_cls,
                                                                 int* __restrict__ cand_anchor,
                                                                 int* __restrict__ cand_count) {
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= p.total_items) return;
    yolo_decode_item<VEC, T>(p, item, threadIdx.x & 31, cand_box, cand_score, cand_cls, cand_anchor, cand_count);
}

// packed-survivor variant (evaluation thresholds / small heads), see yolo_decode_item_packed.  80 registers, 3 CTAs per SM: this
// variant is bound by issue slots and load latency, not by warps in flight, and at 64 registers it spilled inside the plane loop
// (cfg4: 211 -> 192 us one step at a time, 133 -> 122 us pipelined)
template <bool VEC, typename T = float>
__global__ void __launch_bounds__(256, 3) yolo_decode_filter_packed_kernel(const __grid_constant__ YoloParams p,
                                                                        float4* __restrict__ cand_box,
                                                                        float* __restrict__ cand_score,
                                                                        int* __restrict__ cand_cls,
                                                                        int* __restrict__ cand_anchor,
                                                                        int* __restrict__ cand_count) {
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    __shared__ float stage[8][8 * 32];   // per warp: 32 dealt-out cells x (obj, max logit, runner-up, class|cell, 4 box logits)
    if (item >= p.total_items) return;
    yolo_decode_item_packed<VEC, T>(p, item, threadIdx.x & 31, stage[threadIdx.x >> 5], cand_box, cand_score, cand_cls, cand_anchor, cand_count);
}

template <bool VEC, typename T = float>
__global__ void __launch_bounds__(256, 2) yolo_decode_filter_multi_kernel(const __grid_constant__ YoloParams p, float4* __restrict__ cand_box,
                                                                          float* __restrict__ cand_score, int* __restrict__ cand_cls,
                                                                          int* __restrict__ cand_anchor, int* __restrict__ cand_count) {
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= p.total_items) return;
    yolo_decode_item_multi<VEC, T>(p, item, threadIdx.x & 31, cand_box, cand_score, cand_cls, cand_anchor, cand_count);
}

// ------------------------------------------------------------------------------------------------
// NHWC heads (HD_FLAG_IN_NHWC; SURVEY.md 8f-4: the layout the final 1x1 conv produces in channels_last).  The A*(5+nc)
// values of a cell are contiguous, so a (cell, anchor) pair owns 5+nc consecutive floats.  A CTA takes a tile of
// NHWC_TC consecutive cells of one (image, level):
//   dense  -- the tile (TC * A * no floats, contiguous in memory) is streamed into shared memory with 128-bit loads,
//             then each thread scans the values of one (cell, anchor) pair (lane stride = no words: conflict free for odd no);
//   sparse -- each thread first reads only its pair's objectness; pairs that pass the gate read their own no floats
//             straight from global memory (rare: ~0.5 % of the pairs at conf 0.25).
// The per-pair arithmetic is the NCHW kernel's, value for value, so both layouts give identical candidates.
// ------------------------------------------------------------------------------------------------
#define NHWC_TC 64
__global__ void __launch_bounds__(256) yolo_decode_filter_nhwc_kernel(const __grid_constant__ YoloParams p, float4* __restrict__ cand_box,
                                                                      float* __restrict__ cand_score, int* __restrict__ cand_cls,
                                                                      int* __restrict__ cand_anchor, int* __restrict__ cand_count) {
    extern __shared__ __align__(16) float tile[];   // dense mode: [TC][A*no]
    const int tid = threadIdx.x, lane = tid & 31;
    // item -> (image, level, tile); tile_start[] counts NHWC_TC-cell tiles per image here (items_per_image = tiles, not A*tiles)
    const long long item = blockIdx.x;
    const int b = (int)(item / p.items_per_image);
    int r = (int)(item - (long long)b * p.items_per_image);
    int l = 0;
#pragma unroll
    for (int q = 1; q < HD_MAX_LEVELS; ++q)
        if (q < p.n_levels && r >= p.tile_start[q]) l = q;
    r -= p.tile_start[l];
    const int HW = p.HW[l], W = p.W[l], no = p.no, row = p.A * no;
    const int cell0 = r * NHWC_TC, ncell = min(NHWC_TC, HW - cell0);
    const float* __restrict__ gbase = reinterpret_cast<const float*>(p.data[l]) + ((size_t)b * HW + cell0) * row;
    const int nval = ncell * row;
    if (p.dense) {
        if ((((uintptr_t)gbase) & 15) == 0) {
            for (int i = tid; i < (nval >> 2); i += 256) reinterpret_cast<float4*>(tile)[i] = hd_ldg_stream4(gbase + 4 * i);
            for (int i = (nval & ~3) + tid; i < nval; i += 256) tile[i] = hd_ldg_stream(gbase + i);
        } else {
            for (int i = tid; i < nval; i += 256) tile[i] = hd_ldg_stream(gbase + i);
        }
        __syncthreads();
    }
    const int npair = ncell * p.A;
    for (int t0 = 0; t0 < npair; t0 += 256) {           // block-uniform trip count (warp ballots inside)
        const int t = t0 + tid;
        bool pass = false;
        float conf = 0.f; int j = 0, cell = 0, a = 0;
        const float* v = nullptr;
        if (t < npair) {
            const int lc = t / p.A;
            a = t - lc * p.A; cell = cell0 + lc;
            v = p.dense ? (tile + (size_t)lc * row + a * no) : (gbase + (size_t)lc * row + a * no);
            const float o = v[4];
            if (o > p.gate) {
                float m = -INFINITY, L = -INFINITY;
                for (int c = 0; c < p.nc; ++c) {
                    const float x = v[5 + c];
                    const bool g = x > m;
                    L = g ? m : L; j = g ? c : j; m = g ? x : m;
                }
                const float po = hd_sigmoid(o);
                const float cf = __fmul_rn(hd_sigmoid(m), po);
                const bool ok = p.ge ? (po >= p.thr && cf >= p.thr) : (po > p.thr && cf > p.thr);
                if (ok) {
                    if (L > -INFINITY && __fmul_rn(hd_sigmoid(L), po) == cf) {
                        // an earlier class ties after rounding: torch.max returns the first maximal product
                        for (int cc = 0; cc < j; ++cc)
                            if (__fmul_rn(hd_sigmoid(v[5 + cc]), po) == cf) { j = cc; break; }
                    }
                    pass = true; conf = cf;
                }
            }
        }
        const unsigned mk = __ballot_sync(HD_FULL, pass);
        if (mk) {
            int base = 0;
            if (lane == 0) base = atomicAdd(cand_count + b, __popc(mk));
            base = __shfl_sync(HD_FULL, base, 0);
            const int slot = base + __popc(mk & hd_lanemask_lt());
            if (pass && slot < p.cap) {
                const float s = p.stride[l];
                const float aw = p.anchor[l][2 * a], ah = p.anchor[l][2 * a + 1];
                const int gi = cell / W, gj = cell - gi * W;
                const float px = __fmul_rn(hd_sigmoid(v[0]), 2.0f), py = __fmul_rn(hd_sigmoid(v[1]), 2.0f);
                const float pw = __fmul_rn(hd_sigmoid(v[2]), 2.0f), ph = __fmul_rn(hd_sigmoid(v[3]), 2.0f);
                const float cx = __fmul_rn(__fadd_rn(__fsub_rn(px, 0.5f), (float)gj), s);
                const float cy = __fmul_rn(__fadd_rn(__fsub_rn(py, 0.5f), (float)gi), s);
                const float w = __fmul_rn(__fmul_rn(pw, pw), aw), h = __fmul_rn(__fmul_rn(ph, ph), ah);
                const float hw2 = __fmul_rn(w, 0.5f), hh2 = __fmul_rn(h, 0.5f);
                const size_t g = (size_t)b * p.cap + slot;
                cand_box[g] = make_float4(__fsub_rn(cx, hw2), __fsub_rn(cy, hh2), __fadd_rn(cx, hw2), __fadd_rn(cy, hh2));
                cand_score[g] = conf;
                cand_cls[g] = j;
                cand_anchor[g] = p.level_off[l] + a * HW + cell;
            }
        }
    }
}

// Dense NHWC variant with the tile moved by the TMA engine: a persistent CTA walks its tiles (32 cells = 32.6 KB at 255
// channels, contiguous in memory) with two shared-memory buffers; one thread issues cp.async.bulk for tile k+1 (completion on
// an mbarrier) while all threads scan tile k, so no issue slots or registers are spent on the loads and a load is always in
// flight per CTA.  Per-pair arithmetic identical to the kernels above.
#define NHWC_TMA_TC 32
__device__ __forceinline__ void yl_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void yl_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void yl_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void yl_bulk_load(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

__global__ void __launch_bounds__(128) yolo_decode_filter_nhwc_tma_kernel(const __grid_constant__ YoloParams p, float4* __restrict__ cand_box,
                                                                          float* __restrict__ cand_score, int* __restrict__ cand_cls,
                                                                          int* __restrict__ cand_anchor, int* __restrict__ cand_count, int buf_floats) {
    extern __shared__ __align__(128) float tiles[];   // [2][buf_floats]
    __shared__ unsigned long long bar[2];
    const int tid = threadIdx.x, lane = tid & 31;
    const int no = p.no, row = p.A * no;
    auto locate = [&](long long item, int& b, int& l, int& cell0, int& ncell, const float*& g) {
        b = (int)(item / p.items_per_image);
        int r = (int)(item - (long long)b * p.items_per_image);
        l = 0;
#pragma unroll
        for (int q = 1; q < HD_MAX_LEVELS; ++q)
            if (q < p.n_levels && r >= p.tile_start[q]) l = q;
        r -= p.tile_start[l];
        cell0 = r * NHWC_TMA_TC; ncell = min(NHWC_TMA_TC, p.HW[l] - cell0);
        g = reinterpret_cast<const float*>(p.data[l]) + ((size_t)b * p.HW[l] + cell0) * row;
    };
    if (tid == 0) {
        yl_mbar_init(&bar[0], 1); yl_mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long item = blockIdx.x;
    if (item >= p.total_items) return;
    if (tid == 0) {
        int b, l, c0, nc_; const float* g;
        locate(item, b, l, c0, nc_, g);
        const unsigned bytes = (unsigned)(nc_ * row * 4);
        yl_mbar_expect_tx(&bar[0], bytes);
        yl_bulk_load(tiles, g, bytes, &bar[0]);
    }
    for (int k = 0; item < p.total_items; item += gridDim.x, ++k) {
        const int cur = k & 1;
        const long long nxt = item + gridDim.x;
        if (tid == 0 && nxt < p.total_items) {   // buffer cur^1 was released by the __syncthreads that ended iteration k-1
            int b, l, c0, nc_; const float* g;
            locate(nxt, b, l, c0, nc_, g);
            const unsigned bytes = (unsigned)(nc_ * row * 4);
            yl_mbar_expect_tx(&bar[cur ^ 1], bytes);
            yl_bulk_load(tiles + (size_t)(cur ^ 1) * buf_floats, g, bytes, &bar[cur ^ 1]);
        }
        int b, l, cell0, ncell; const float* gsrc;
        locate(item, b, l, cell0, ncell, gsrc);
        yl_mbar_wait(&bar[cur], (unsigned)((k >> 1) & 1));
        const float* tile = tiles + (size_t)cur * buf_floats;
        const int HW = p.HW[l], W = p.W[l];
        const int npair = ncell * p.A;
        for (int t0 = 0; t0 < npair; t0 += 128) {
            const int t = t0 + tid;
            bool pass = false;
            float conf = 0.f; int j = 0, cell = 0, a = 0;
            const float* v = nullptr;
            if (t < npair) {
                const int lc = t / p.A;
                a = t - lc * p.A; cell = cell0 + lc;
                v = tile + (size_t)lc * row + a * no;
                const float o = v[4];
                if (o > p.gate) {
                    float m = -INFINITY, L = -INFINITY;
                    for (int c = 0; c < p.nc; ++c) {
                        const float x = v[5 + c];
                        const bool g = x > m;
                        L = g ? m : L; j = g ? c : j; m = g ? x : m;
                    }
                    const float po = hd_sigmoid(o);
                    const float cf = __fmul_rn(hd_sigmoid(m), po);
                    const bool ok = p.ge ? (po >= p.thr && cf >= p.thr) : (po > p.thr && cf > p.thr);
                    if (ok) {
                        if (L > -INFINITY && __fmul_rn(hd_sigmoid(L), po) == cf) {
                            for (int cc = 0; cc < j; ++cc)
                                if (__fmul_rn(hd_sigmoid(v[5 + cc]), po) == cf) { j = cc; break; }
                        }
                        pass = true; conf = cf;
                    }
                }
            }
            const unsigned mk = __ballot_sync(HD_FULL, pass);
            if (mk) {
                int base = 0;
                if (lane == 0) base = atomicAdd(cand_count + b, __popc(mk));
                base = __shfl_sync(HD_FULL, base, 0);
                const int slot = base + __popc(mk & hd_lanemask_lt());
                if (pass && slot < p.cap) {
                    const float s = p.stride[l];
                    const float aw = p.anchor[l][2 * a], ah = p.anchor[l][2 * a + 1];
                    const int gi = cell / W, gj = cell - gi * W;
                    const float px = __fmul_rn(hd_sigmoid(v[0]), 2.0f), py = __fmul_rn(hd_sigmoid(v[1]), 2.0f);
                    const float pw = __fmul_rn(hd_sigmoid(v[2]), 2.0f), ph = __fmul_rn(hd_sigmoid(v[3]), 2.0f);
                    const float cx = __fmul_rn(__fadd_rn(__fsub_rn(px, 0.5f), (float)gj), s);
                    const float cy = __fmul_rn(__fadd_rn(__fsub_rn(py, 0.5f), (float)gi), s);
                    const float w = __fmul_rn(__fmul_rn(pw, pw), aw), h = __fmul_rn(__fmul_rn(ph, ph), ah);
                    const float hw2 = __fmul_rn(w, 0.5f), hh2 = __fmul_rn(h, 0.5f);
                    const size_t g = (size_t)b * p.cap + slot;
                    cand_box[g] = make_float4(__fsub_rn(cx, hw2), __fsub_rn(cy, hh2), __fadd_rn(cx, hw2), __fadd_rn(cy, hh2));
                    cand_score[g] = conf;
                    cand_cls[g] = j;
                    cand_anchor[g] = p.level_off[l] + a * HW + cell;
                }
            }
        }
        __syncthreads();   // everyone is done with buffer `cur`: it may be refilled two iterations from now
    }
}

// ------------------------------------------------------------------------------------------------
// Dense decode to pred[B, N, 5+nc] (drop-in for decode_box).  A warp reads 32 consecutive cells of
// every plane (coalesced), transposes through padded shared memory and writes the 32 output rows,
// which are contiguous in pred, with fully coalesced stores.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) yolo_decode_kernel(const __grid_constant__ YoloParams p, float* __restrict__ pred,
                                                          int total_anchors) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* tile = smem + (size_t)wid * p.no * 33;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + wid;  // 32-cell tiles here
    if (item >= p.total_items) return;
    const int b = (int)(item / p.items_per_image);
    int r = (int)(item - (long long)b * p.items_per_image);
    int l = 0;
#pragma unroll
    for (int q = 1; q < HD_MAX_LEVELS; ++q)
        if (q < p.n_levels && r >= p.A * p.tile_start[q]) l = q;
    r -= p.A * p.tile_start[l];
    const int tiles_l = p.tile_start[l + 1] - p.tile_start[l];
    const int a = r / tiles_l, t = r - a * tiles_l;
    const int HW = p.HW[l], W = p.W[l];
    const int cell = t * 32 + lane;
    const bool valid = cell < HW;
    const float* __restrict__ base = reinterpret_cast<const float*>(p.data[l]) + ((size_t)(b * p.A + a) * p.no) * HW + cell;
    const float s = p.stride[l];
    const int gi = cell / W, gj = cell - gi * W;
    for (int c0 = 0; c0 < p.no; c0 += 32) {  // 32 independent plane loads in flight per lane: with 20 warps per SM (the
                                             // transpose tiles bound the occupancy) eight left only ~3 MB in flight chip-wide, i.e. ~3 TB/s (1.55 ms; 16: 1.39 ms)
        float raw[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) raw[u] = (valid && c0 + u < p.no) ? hd_ldg_stream(base + (size_t)(c0 + u) * HW) : 0.f;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            const int c = c0 + u;
            if (c >= p.no) break;
            float v = 0.f;
            if (valid) {
                const float pr = hd_sigmoid(raw[u]);
                if (c == 0) v = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(pr, 2.0f), 0.5f), (float)gj), s);
                else if (c == 1) v = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(pr, 2.0f), 0.5f), (float)gi), s);
                else if (c == 2 || c == 3) {
                    const float q2 = __fmul_rn(pr, 2.0f);
                    v = __fmul_rn(__fmul_rn(q2, q2), p.anchor[l][2 * a + (c - 2)]);
                } else v = pr;
            }
            tile[c * 33 + lane] = v;
        }
    }
    __syncwarp();
    const int rows = min(32, HW - t * 32);
    float* out = pred + ((size_t)b * total_anchors + p.level_off[l] + a * HW + t * 32) * p.no;
    {   // k = row*no + c walks the contiguous output; (row, c) advance incrementally (no >= 6, stride 32)
        int row = lane / p.no, c = lane - row * p.no;
        const int row_step = 32 / p.no, c_step = 32 - row_step * p.no;
        for (int k = lane; k < rows * p.no; k += 32) {
            out[k] = tile[c * 33 + row];
            c += c_step; row += row_step;
            if (c >= p.no) { c -= p.no; ++row; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Filter on a decoded prediction [B, N, 5+nc]: warp per 32 rows; objectness gathered first, rows of
// survivors are then read cooperatively (3 coalesced requests for nc=80) with a warp arg-max.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) yolo_filter_pred_kernel(const float* __restrict__ pred, int B, int N, int nc,
                                                               float thr, int ge, float4* __restrict__ cand_box,
                                                               float* __restrict__ cand_score, int* __restrict__ cand_cls,
                                                               int* __restrict__ cand_anchor, int* __restrict__ cand_count,
                                                               int cap) {
    const int lane = threadIdx.x & 31;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int groups = (N + 31) / 32;
    if (wg >= (long long)B * groups) return;
    const int b = (int)(wg / groups);
    const int row0 = (int)(wg - (long long)b * groups) * 32;
    const int no = 5 + nc;
    const float* __restrict__ img = pred + (size_t)b * N * no;
    const int myrow = row0 + lane;
    float obj = (myrow < N) ? __ldg(img + (size_t)myrow * no + 4) : -INFINITY;
    bool s = ge ? (obj >= thr) : (obj > thr);
    unsigned surv = __ballot_sync(HD_FULL, s);
    while (surv) {
        int src = __ffs(surv) - 1;
        surv &= surv - 1;
        const int row = row0 + src;
        const float po = __shfl_sync(HD_FULL, obj, src);
        const float* __restrict__ q = img + (size_t)row * no;
        float best = -INFINITY;
        int bj = 0x7fffffff;
        for (int c = lane; c < nc; c += 32) {
            float v = __fmul_rn(__ldg(q + 5 + c), po);
            if (bj == 0x7fffffff || v > best) { best = v; bj = c; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            float ob = __shfl_xor_sync(HD_FULL, best, d);
            int oj = __shfl_xor_sync(HD_FULL, bj, d);
            bool take = (oj != 0x7fffffff) && (bj == 0x7fffffff || ob > best || (ob == best && oj < bj));
            if (take) { best = ob; bj = oj; }
        }
        if (lane == 0 && bj != 0x7fffffff) {
            bool ok = ge ? (best >= thr) : (best > thr);
            if (ok) {
                int slot = atomicAdd(cand_count + b, 1);
                if (slot < cap) {
                    float cx = __ldg(q), cy = __ldg(q + 1), w = __ldg(q + 2), h = __ldg(q + 3);
                    float hw2 = __fmul_rn(w, 0.5f), hh2 = __fmul_rn(h, 0.5f);
                    size_t g = (size_t)b * cap + slot;
                    cand_box[g] = make_float4(__fsub_rn(cx, hw2), __fsub_rn(cy, hh2), __fadd_rn(cx, hw2), __fadd_rn(cy, hh2));
                    cand_score[g] = best;
                    cand_cls[g] = bj;
                    cand_anchor[g] = row;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ host
static int fill_params(YoloParams& p, const hd_yolo_level* levels, int n_levels, int B, int A, int nc, int tile_cells) {
    HD_CHECK_ARG(levels != nullptr, "levels is NULL");
    HD_CHECK_ARG(n_levels >= 1 && n_levels <= HD_MAX_LEVELS, "n_levels must be in [1,%d], got %d", HD_MAX_LEVELS, n_levels);
    HD_CHECK_ARG(A >= 1 && A <= HD_MAX_ANCHORS, "A must be in [1,%d], got %d", HD_MAX_ANCHORS, A);
    HD_CHECK_ARG(B >= 0 && nc >= 1 && nc < 65536, "B must be >= 0 and nc in [1, 65535], got B=%d nc=%d", B, nc);
    memset(&p, 0, sizeof(p));
    p.n_levels = n_levels; p.B = B; p.A = A; p.nc = nc; p.no = 5 + nc;
    int tiles = 0, off = 0;
    for (int l = 0; l < n_levels; ++l) {
        HD_CHECK_ARG(levels[l].H > 0 && levels[l].W > 0, "level %d has empty spatial size", l);
        HD_CHECK_ARG(B == 0 || levels[l].data != nullptr, "level %d data is NULL", l);
        p.data[l] = levels[l].data;
        p.HW[l] = levels[l].H * levels[l].W;
        p.W[l] = levels[l].W;
        p.stride[l] = levels[l].stride;
        for (int k = 0; k < 2 * A; ++k) p.anchor[l][k] = levels[l].anchor_wh[k];
        p.tile_start[l] = tiles;
        p.level_off[l] = off;
        tiles += (p.HW[l] + tile_cells - 1) / tile_cells;
        off += A * p.HW[l];
    }
    p.tile_start[n_levels] = tiles;
    p.level_off[n_levels] = off;
    p.items_per_image = A * tiles;
    p.total_items = (long long)B * p.items_per_image;
    return HD_OK;
}

static float conf_gate(double thr) {
    if (!(thr > 0.0)) return -INFINITY;
    if (thr >= 1.0) return 15.0f;
    return (float)(log(thr / (1.0 - thr)) - 0.01);
}

extern "C" HD_API int hd_yolo_decode_filter(const hd_yolo_level* levels, int n_levels, int B, int A, int nc, double conf_thres,
                                     int flags, float* cand_box, float* cand_score, int32_t* cand_cls,
                                     int32_t* cand_anchor, int32_t* cand_count, int cap, void* stream) {
    YoloParams p;
    int rc = fill_params(p, levels, n_levels, B, A, nc, 128);
    if (rc) return rc;
    HD_CHECK_ARG(cap > 0 && cand_box && cand_score && cand_cls && cand_anchor && cand_count, "null output or cap <= 0");
    p.thr = (float)conf_thres;
    p.gate = conf_gate(conf_thres);
    p.gate2 = (conf_thres > 0.0) ? (float)(log(conf_thres < 1.0 ? conf_thres : 1.0) - 1e-3) : -INFINITY;
    p.ge = (flags & HD_FLAG_CONF_GE) ? 1 : 0;
    p.dense = (flags & HD_FLAG_DENSE_READ) ? 1 : 0;
    p.multi = (flags & HD_FLAG_MULTI_LABEL) ? 1 : 0;
    p.cap = cap;
    if (p.multi) {
        HD_CHECK_ARG(!(flags & HD_FLAG_IN_NHWC), "HD_FLAG_MULTI_LABEL needs NCHW heads");
        long long tot = 0;
        for (int l = 0; l < n_levels; ++l) tot += (long long)A * p.HW[l];
        HD_CHECK_ARG(tot * nc < (1ll << 31), "anchors * classes must be < 2^31 for the multi_label candidate index");
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (B == 0) return HD_OK;
    HD_CUDA_CALL(cudaMemsetAsync(cand_count, 0, sizeof(int) * (size_t)B, st));
    if (flags & HD_FLAG_IN_NHWC) {
        HD_CHECK_ARG(!(flags & (HD_FLAG_IN_F16 | HD_FLAG_IN_BF16)), "NHWC heads are fp32 only");
        YoloParams pn;
        rc = fill_params(pn, levels, n_levels, B, A, nc, NHWC_TC);
        if (rc) return rc;
        pn.thr = p.thr; pn.gate = p.gate; pn.ge = p.ge; pn.dense = p.dense; pn.cap = p.cap;
        pn.items_per_image = pn.tile_start[n_levels];          // tiles of NHWC_TC cells; a CTA handles all A anchors of its cells
        pn.total_items = (long long)B * pn.items_per_image;
        HD_CHECK_ARG(pn.total_items < (1ll << 31), "grid too large");
        // dense + 16-byte aligned rows: TMA double-buffered persistent kernel
        bool tma_ok = pn.dense != 0;
        for (int l = 0; l < n_levels; ++l) tma_ok = tma_ok && (((uintptr_t)pn.data[l] & 15) == 0) && (((size_t)pn.HW[l] * A * (5 + nc) * 4) % 16 == 0);
        tma_ok = tma_ok && (((size_t)NHWC_TMA_TC * A * (5 + nc) * 4) % 16 == 0);
        if (tma_ok) {
            YoloParams pt;
            rc = fill_params(pt, levels, n_levels, B, A, nc, NHWC_TMA_TC);
            if (rc) return rc;
            pt.thr = p.thr; pt.gate = p.gate; pt.ge = p.ge; pt.dense = 1; pt.cap = p.cap;
            pt.items_per_image = pt.tile_start[n_levels];
            pt.total_items = (long long)B * pt.items_per_image;
            const int buf_floats = (int)hd_align_up((size_t)NHWC_TMA_TC * A * (5 + nc), 32);
            const size_t smt = 2 * (size_t)buf_floats * 4;
            if (smt <= 200 * 1024) {
                HD_ENSURE_SMEM(yolo_decode_filter_nhwc_tma_kernel, 200 * 1024);
                int per_sm = (int)((220 * 1024) / (smt + 1024));
                if (per_sm < 1) per_sm = 1;
                if (per_sm > 8) per_sm = 8;
                long long grid = (long long)hd_num_sms() * per_sm;
                if (grid > pt.total_items) grid = pt.total_items;
                yolo_decode_filter_nhwc_tma_kernel<<<(unsigned)grid, 128, smt, st>>>(pt, (float4*)cand_box, cand_score, cand_cls, cand_anchor, cand_count, buf_floats);
                HD_CUDA_LAUNCH_CHECK("yolo_decode_filter_nhwc_tma_kernel");
                return HD_OK;
            }
        }
        const size_t sm = pn.dense ? (size_t)NHWC_TC * A * (5 + nc) * 4 : 0;
        HD_CHECK_ARG(sm <= 200 * 1024, "A*(5+nc)=%d too large for the NHWC tile", A * (5 + nc));
        HD_ENSURE_SMEM(yolo_decode_filter_nhwc_kernel, 200 * 1024);
        yolo_decode_filter_nhwc_kernel<<<(unsigned)pn.total_items, 256, sm, st>>>(pn, (float4*)cand_box, cand_score, cand_cls, cand_anchor, cand_count);
        HD_CUDA_LAUNCH_CHECK("yolo_decode_filter_nhwc_kernel");
        return HD_OK;
    }
    const int dt = (flags & HD_FLAG_IN_F16) ? 1 : ((flags & HD_FLAG_IN_BF16) ? 2 : 0);
    HD_CHECK_ARG(!((flags & HD_FLAG_IN_F16) && (flags & HD_FLAG_IN_BF16)), "HD_FLAG_IN_F16 and HD_FLAG_IN_BF16 are exclusive");
    bool vec = true;
    for (int l = 0; l < n_levels; ++l) vec = vec && (p.HW[l] % 4 == 0) && (((uintptr_t)p.data[l] & (dt ? 7 : 15)) == 0);
    const int warps = 8;
    long long blocks = (p.total_items + warps - 1) / warps;
    HD_CHECK_ARG(blocks < (1ll << 31), "grid too large");
    // many survivors per tile (evaluation thresholds) or few planes per tile: the packed variant; else the lean one
    const bool packed = conf_thres < 0.05 || nc < 16;
#define HD_YOLO_LAUNCH(V, T)                                                                                                                     \
    do {                                                                                                                                         \
        if (p.multi) yolo_decode_filter_multi_kernel<V, T><<<(unsigned)blocks, warps * 32, 0, st>>>(p, (float4*)cand_box, cand_score, cand_cls, cand_anchor, cand_count); \
        else if (packed) yolo_decode_filter_packed_kernel<V, T><<<(unsigned)blocks, warps * 32, 0, st>>>(p, (float4*)cand_box, cand_score, cand_cls, cand_anchor, cand_count); \
        else yolo_decode_filter_kernel<V, T><<<(unsigned)blocks, warps * 32, 0, st>>>(p, (float4*)cand_box, cand_score, cand_cls, cand_anchor, cand_count);                \
    } while (0)
    if (dt == 0) { if (vec) HD_YOLO_LAUNCH(true, float); else HD_YOLO_LAUNCH(false, float); }
    else if (dt == 1) { if (vec) HD_YOLO_LAUNCH(true, HdF16); else HD_YOLO_LAUNCH(false, HdF16); }
    else { if (vec) HD_YOLO_LAUNCH(true, HdBF16); else HD_YOLO_LAUNCH(false, HdBF16); }
#undef HD_YOLO_LAUNCH
    HD_CUDA_LAUNCH_CHECK("yolo_decode_filter_kernel");
    return HD_OK;
}

extern "C" HD_API int hd_yolo_decode(const hd_yolo_level* levels, int n_levels, int B, int A, int nc, float* pred, void* stream) {
    YoloParams p;
    int rc = fill_params(p, levels, n_levels, B, A, nc, 32);
    if (rc) return rc;
    if (B == 0) return HD_OK;
    HD_CHECK_ARG(pred != nullptr, "pred is NULL");
    const int warps = 4;
    size_t smem = (size_t)warps * p.no * 33 * sizeof(float);
    HD_CHECK_ARG(smem <= 200 * 1024, "nc=%d too large for the decode transpose tile", nc);
    HD_ENSURE_SMEM(yolo_decode_kernel, 200 * 1024);
    long long blocks = (p.total_items + warps - 1) / warps;
    HD_CHECK_ARG(blocks < (1ll << 31), "grid too large");
    yolo_decode_kernel<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(p, pred, p.level_off[n_levels]);
    HD_CUDA_LAUNCH_CHECK("yolo_decode_kernel");
    return HD_OK;
}

extern "C" HD_API int hd_yolo_filter_pred(const float* pred, int B, int N, int nc, double conf_thres, int flags, float* cand_box,
                                   float* cand_score, int32_t* cand_cls, int32_t* cand_anchor, int32_t* cand_count,
                                   int cap, void* stream) {
    HD_CHECK_ARG(B >= 0 && N >= 0 && nc >= 1, "bad shape B=%d N=%d nc=%d", B, N, nc);
    HD_CHECK_ARG(cap > 0 && cand_box && cand_score && cand_cls && cand_anchor && cand_count, "null output or cap <= 0");
    if (B == 0) return HD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    HD_CUDA_CALL(cudaMemsetAsync(cand_count, 0, sizeof(int) * (size_t)B, st));
    if (N == 0) return HD_OK;
    HD_CHECK_ARG(pred != nullptr, "pred is NULL");
    long long groups = (long long)B * ((N + 31) / 32);
    long long blocks = (groups + 7) / 8;
    HD_CHECK_ARG(blocks < (1ll << 31), "grid too large");
    yolo_filter_pred_kernel<<<(unsigned)blocks, 256, 0, st>>>(pred, B, N, nc, (float)conf_thres, (flags & HD_FLAG_CONF_GE) ? 1 : 0,
                                                             (float4*)cand_box, cand_score, cand_cls, cand_anchor, cand_count, cap);
    HD_CUDA_LAUNCH_CHECK("yolo_filter_pred_kernel");
    return HD_OK;
}

// ------------------------------------------------------------------------------------------------ one-call entry
static void post_ws_layout(int B, int cap, size_t* offs, size_t* total) {
    size_t n = (size_t)B * cap, o = 0;
    offs[0] = o; o = hd_align_up(o + n * 16, 256);           // cand_box
    offs[1] = o; o = hd_align_up(o + n * 4, 256);            // cand_score
    offs[2] = o; o = hd_align_up(o + n * 4, 256);            // cand_cls
    offs[3] = o; o = hd_align_up(o + n * 4, 256);            // cand_anchor
    offs[4] = o; o = hd_align_up(o + (size_t)B * 4, 256);    // cand_count
    offs[5] = o; o += hd_sort_nms_workspace_size(B, cap);    // sort/NMS scratch of the large images
    *total = o;
}

extern "C" HD_API size_t hd_yolo_postprocess_workspace_size(int B, int total_anchors) {
    size_t offs[6], total;
    post_ws_layout(B < 0 ? 0 : B, total_anchors < 0 ? 0 : total_anchors, offs, &total);
    return total + 256;
}

extern "C" HD_API int hd_yolo_postprocess(const hd_yolo_level* levels, int n_levels, int B, int A, int nc, double conf_thres, double iou_thres,
                                          int flags, int class_mode, float offset_scale, int max_nms, int max_det, float* out_det,
                                          int64_t* out_idx, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
    return hd_yolo_postprocess_replicated(levels, n_levels, B, A, nc, conf_thres, iou_thres, flags, class_mode, offset_scale, max_nms,
                                          max_det, out_det, out_idx, out_count, nullptr, workspace, workspace_bytes, stream);
}

extern "C" HD_API int hd_yolo_postprocess_replicated(const hd_yolo_level* levels, int n_levels, int B, int A, int nc, double conf_thres,
                                                     double iou_thres, int flags, int class_mode, float offset_scale, int max_nms,
                                                     int max_det, float* out_det, int64_t* out_idx, int32_t* out_count,
                                                     const hd_replicas* replicas, void* workspace, size_t workspace_bytes, void* stream) {
    HD_CHECK_ARG(levels != nullptr && n_levels >= 1 && n_levels <= HD_MAX_LEVELS, "n_levels must be in [1,%d], got %d", HD_MAX_LEVELS, n_levels);
    if (B == 0) return HD_OK;
    long long cap_ll = 0;
    for (int l = 0; l < n_levels; ++l) cap_ll += (long long)A * levels[l].H * levels[l].W;
    HD_CHECK_ARG(cap_ll > 0 && cap_ll < (1ll << 31), "bad total anchor count");
    const int cap = (int)cap_ll;
    size_t offs[6], total;
    post_ws_layout(B, cap, offs, &total);
    uintptr_t w0 = hd_align_up((uintptr_t)workspace, 256);
    if (!workspace || w0 + total > (uintptr_t)workspace + workspace_bytes)
        HD_FAIL(HD_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", total + 256, workspace_bytes);
    float* cand_box = (float*)(w0 + offs[0]);
    float* cand_score = (float*)(w0 + offs[1]);
    int32_t* cand_cls = (int32_t*)(w0 + offs[2]);
    int32_t* cand_anchor = (int32_t*)(w0 + offs[3]);
    int32_t* cand_count = (int32_t*)(w0 + offs[4]);
    int rc = hd_yolo_decode_filter(levels, n_levels, B, A, nc, conf_thres, flags, cand_box, cand_score, cand_cls, cand_anchor, cand_count, cap, stream);
    if (rc) return rc;
    return hd_sort_nms_batched_replicated(cand_box, cand_score, cand_cls, cand_anchor, cand_count, 0, B, cap, iou_thres, class_mode,
                                          offset_scale, max_nms, max_det, out_det, out_idx, out_count, replicas, (void*)(w0 + offs[5]),
                                          hd_sort_nms_workspace_size(B, cap), stream);
}

// RoIAlign, streamed ("strip") kernel for many RoIs on channel-contiguous (NHWC) feature maps -- the cfg3 hot kernel.
// Replaces torchvision::roi_align for the MultiScaleRoIAlign call site (roi_align.py:204-260, poolers.py:147-227; README.md:65).
//
// Why: the per-RoI gather kernel (roi.cu) re-reads every feature cell once per RoI that touches it; with 2000 proposals per
// image the footprints cover the stride-4 map ~5x over, and the kernel sits on the L2->SM limit (12 GB moved for a 0.94 GB
// pyramid).  Here the RoIs are bucketed by (image, level, x-strip) and sorted by their first feature row, and one CTA per
// (bucket, 32-channel slice) streams the rows of its strip ONCE, top to bottom, through a shared-memory ring filled by TMA
// tensor loads (cp.async.bulk.tensor.4d, one row x strip width x 32 channels per request, completion on an mbarrier):
//   producer (1 thread)  re-arms a ring slot as soon as no unfinished RoI needs its row, and issues the next row;
//   8 consumer warps     take the RoIs in row order; lane = channel, so every shared-memory read is one conflict-free 128-byte
//                        row of the ring; per feature row a lane forms the PW x-interpolated values (separable merged-weight
//                        tables, identical to roi.cu) and adds them into the bins that row contributes to;
//   output               the [32, PH*PW] slice of the result tile is assembled in shared memory and leaves as 128-bit stores.
// Tall RoIs are split into two bin-row ranges so that the ring stays short; RoIs the strip cannot take (footprint wider than
// the strip halo, adaptive sampling grids, inverted boxes) go to the gather kernel through a device-side list.
// Same arithmetic as the gather kernel (same tables, same summation order), so both give the same values.
#include "hd_roi_axis.cuh"
#include "hd_roi_internal.cuh"
#include <cuda.h>

#define RS_NCW 11                      // consumer warps (13 warps x 152 registers fill the register file)
#define RS_THREADS ((RS_NCW + 1) * 32)
#define RS_CS 32                       // channels per slice (= lanes)
#define RS_P 7                         // PH, PW <= 7
#define RS_E 4                         // merged table entries per bin (sampling_ratio <= 2)
#define RS_MAXROWS 32                  // feature rows of one work item
#define RS_HALO 28                     // a strip's box is SW + RS_HALO cells wide: footprints up to RS_HALO + 1 cells
#define RS_MAXH 1024                   // feature-map height bound of the per-bucket counting sort

struct RsLevel { int H, W, SW, S, RW, unit_base; float scale; };
struct RsItem { int k, pr, y0, y1; };  // RoI index, bin rows [pr & 255, (pr >> 8) & 255) and bucket << 16, first / last feature row

struct RsParams {
    RsLevel lv[HD_MAX_LEVELS];
    int n_levels, C, PH, PW, sr, aligned, B, NR, units_per_img, n_units;
    const float* rois; const int* level_ids; long long K; float* out;
    int* unit_count; int* unit_start; int* unit_cursor; int* unit_ymax;
    RsItem* tmp; int* tmp_unit; RsItem* scat; RsItem* items; int* tables;
    int* n_tmp; int* fb_count; int* fb_list;
    int dbg;                    // developer switches (hd_roi_set_mode bits 4..): 1 no prefetch, 2 no row walk, 4 no tile store
    unsigned long long* prof;   // nullable developer counters: [0] table cycles [1] wait cycles [2] row-loop cycles [3] output cycles [4] items [5] producer wait
};
struct RsMaps { CUtensorMap m[HD_MAX_LEVELS]; };

struct RsGeom { float sw, sh, bw, bh, count; int g, lvl, img; bool ok; };

__device__ __forceinline__ RsGeom rs_geom(const RsParams& p, long long k) {
    RsGeom G;
    const float* roi = p.rois + k * 5;
    G.lvl = p.level_ids ? p.level_ids[k] : 0;
    G.ok = G.lvl >= 0 && G.lvl < p.n_levels;
    const int lvl = G.ok ? G.lvl : 0;
    const float sc = p.lv[lvl].scale;
    const float fb = roi[0];
    G.img = (int)fb;
    G.ok = G.ok && fb >= 0.0f && G.img < p.B;
    const float off = p.aligned ? 0.5f : 0.0f;
    G.sw = __fsub_rn(__fmul_rn(roi[1], sc), off); G.sh = __fsub_rn(__fmul_rn(roi[2], sc), off);
    const float ew = __fsub_rn(__fmul_rn(roi[3], sc), off), eh = __fsub_rn(__fmul_rn(roi[4], sc), off);
    float rw = __fsub_rn(ew, G.sw), rh = __fsub_rn(eh, G.sh);
    if (!p.aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
    G.bh = __fdiv_rn(rh, (float)p.PH); G.bw = __fdiv_rn(rw, (float)p.PW);
    G.g = p.sr;
    G.count = (float)max(p.sr * p.sr, 1);
    // sample positions must not decrease along an axis (inverted aligned boxes walk backwards), and everything must be finite
    G.ok = G.ok && rw >= 0.0f && rh >= 0.0f && fabsf(G.sw) < 1.0e8f && fabsf(G.sh) < 1.0e8f && rw < 1.0e8f && rh < 1.0e8f;
    return G;
}

// first / last cell touched by the samples of bins [b0, b1) of one axis (build_axis' sample arithmetic); false if no sample is valid
__device__ __forceinline__ bool rs_axis_range(float start, float bin, int g, int extent, int b0, int b1, int* lo_min, int* hi_max) {
    int lo_m = 0x7fffffff, hi_m = -1;
    for (int b = b0; b < b1; ++b)
        for (int i = 0; i < g; ++i) {
            float y = __fadd_rn(__fadd_rn(start, __fmul_rn((float)b, bin)), __fdiv_rn(__fmul_rn(__fadd_rn((float)i, 0.5f), bin), (float)g));
            if (y < -1.0f || y > (float)extent) continue;
            if (y <= 0.0f) y = 0.0f;
            int lo = (int)y, hi;
            if (lo >= extent - 1) hi = lo = extent - 1; else hi = lo + 1;
            lo_m = min(lo_m, lo); hi_m = max(hi_m, hi);
        }
    *lo_min = lo_m; *hi_max = hi_m;
    return hi_m >= 0;
}

// ------------------------------------------------------------------------------------------------ bucketing
__global__ void __launch_bounds__(256) rs_prep_kernel(const __grid_constant__ RsParams p) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p.K) return;
    const RsGeom G = rs_geom(p, k);
    bool strip = G.ok;
    int x0 = 0, x1 = -1, ya = 0, yb = -1, yc = 0, yd = -1, unit = 0, split = 0;
    if (strip) {
        const RsLevel& L = p.lv[G.lvl];
        strip = rs_axis_range(G.sw, G.bw, G.g, L.W, 0, p.PW, &x0, &x1) && rs_axis_range(G.sh, G.bh, G.g, L.H, 0, p.PH, &ya, &yb);
        strip = strip && (L.S == 1 || x1 - x0 <= RS_HALO) && (p.NR <= RS_MAXROWS || yb - ya + 1 <= RS_MAXROWS);   // the item table holds <= 32 rows
        if (strip && yb - ya + 1 > p.NR - 3) {   // tall: two bin-row ranges, each with its own (shorter) row span
            split = (p.PH + 1) / 2;
            const bool a = rs_axis_range(G.sh, G.bh, G.g, L.H, 0, split, &ya, &yb);
            const bool c = rs_axis_range(G.sh, G.bh, G.g, L.H, split, p.PH, &yc, &yd);
            strip = a && c && (yb - ya + 1 <= p.NR - 3) && (yd - yc + 1 <= p.NR - 3);
        }
        if (strip) unit = G.img * p.units_per_img + L.unit_base + min(x0 / L.SW, L.S - 1);
    }
    if (!strip) {
        p.fb_list[atomicAdd(p.fb_count, 1)] = (int)k;
        return;
    }
    const int n = split ? 2 : 1;
    // a few hundred buckets take all K updates: aggregate per warp (lanes of one bucket elect a leader) before touching memory
    const unsigned peers = __match_any_sync(__activemask(), unit);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    int nsum = 0, ymx = -1, rank = 0;
    for (unsigned m = peers; m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        const int on = __shfl_sync(peers, n, src), oy = __shfl_sync(peers, split ? max(yb, yd) : yb, src);
        if (src < lane) rank += on;
        nsum += on; ymx = max(ymx, oy);
    }
    int base = 0;
    if (lane == leader) {
        base = atomicAdd(p.n_tmp, nsum);
        atomicAdd(p.unit_count + unit, nsum);
        atomicMax(p.unit_ymax + unit, ymx);
    }
    const int at = __shfl_sync(peers, base, leader) + rank;
    RsItem it;
    it.k = (int)k;
    it.pr = (split ? (split << 8) : (p.PH << 8)) | (unit << 16); it.y0 = ya; it.y1 = yb;
    p.tmp[at] = it; p.tmp_unit[at] = unit;
    if (split) {
        it.pr = split | (p.PH << 8) | (unit << 16); it.y0 = yc; it.y1 = yd;
        p.tmp[at + 1] = it; p.tmp_unit[at + 1] = unit;
    }
}

__global__ void __launch_bounds__(1024) rs_scan_kernel(const __grid_constant__ RsParams p) {   // exclusive scan of the bucket sizes (one CTA)
    __shared__ int wsum[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < p.n_units; base += 1024) {
        const int i = base + tid;
        const int v = i < p.n_units ? p.unit_count[i] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(HD_FULL, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        int pre = carry;
        for (int w = 0; w < wid; ++w) pre += wsum[w];
        if (i < p.n_units) p.unit_start[i] = pre + incl - v;
        __syncthreads();
        if (tid == 1023) carry = pre + incl;
        __syncthreads();
    }
    if (tid == 0) p.unit_start[p.n_units] = carry;
}

__global__ void __launch_bounds__(256) rs_scatter_kernel(const __grid_constant__ RsParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *p.n_tmp) return;
    const int u = p.tmp_unit[i];
    p.scat[p.unit_start[u] + atomicAdd(p.unit_cursor + u, 1)] = p.tmp[i];
}

// one CTA per bucket: counting sort of its items by first row (any order inside a row: every RoI's arithmetic is its own)
__global__ void __launch_bounds__(256) rs_sort_kernel(const __grid_constant__ RsParams p) {
    __shared__ int hist[RS_MAXH + 1];
    const int u = blockIdx.x, tid = threadIdx.x;
    const int s = p.unit_start[u], n = p.unit_start[u + 1] - s;
    if (n <= 0) return;
    for (int i = tid; i <= RS_MAXH; i += 256) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) atomicAdd(&hist[min(max(p.scat[s + i].y0, 0), RS_MAXH - 1) + 1], 1);
    __syncthreads();
    if (tid == 0) for (int i = 1; i <= RS_MAXH; ++i) hist[i] += hist[i - 1];   // 1 K serial adds: negligible next to the main kernel
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const RsItem it = p.scat[s + i];
        p.items[s + atomicAdd(&hist[min(max(it.y0, 0), RS_MAXH - 1)], 1)] = it;
    }
}

// ------------------------------------------------------------------------------------------------ main kernel
__device__ __forceinline__ void rs_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rs_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// bounded spin: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void rs_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void rs_tma_row(void* sdst, const CUtensorMap* map, int c0, int x, int y, int n, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(map), "r"(c0), "r"(x), "r"(y), "r"(n),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// Per-item table (RS_TABW words, built once per item by rs_table_kernel and shared by the C/32 slice CTAs):
//   [0,28)   xo   ring-row float offset of x entry (pw*4 + e): (cell - strip start) * 32        (0 for unused entries)
//   [28,56)  wx   merged weight of that entry                                                    (0 for unused entries)
//   [56,63)  nx   entries of bin column pw;   [63] number of (row, bin-row) pairs
//   [64,92)  code (row - y0) << 8 | ph, sorted by row;   [92,120) weight of the pair
//   [120,124) k, p0 | p1 << 8, y0, y1
#define RS_TABW 128
#define RS_NTILE 6                     // output tiles shared by the consumer warps (taken with a shared-memory lock)

// one warp per item: the separable tables of roi.cu (build_axis) turned into the layout above
__global__ void __launch_bounds__(256) rs_table_kernel(const __grid_constant__ RsParams p) {
    __shared__ AxisEntry xt[8][RS_P * RS_E], yt[8][RS_P * RS_E];
    __shared__ int xc[8][8], yc[8][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + w;
    if (i >= *p.n_tmp) return;
    const RsItem it = p.items[i];
    const int p0 = it.pr & 255, p1 = (it.pr >> 8) & 255, unit = it.pr >> 16;
    const RsGeom G = rs_geom(p, it.k);
    const RsLevel& L = p.lv[G.lvl];
    const int xs = (unit % p.units_per_img - L.unit_base) * L.SW;
    if (lane < p.PH) build_axis(yt[w], yc[w], lane, RS_E, G.sh, G.bh, G.g, L.H, 1, 0);
    else if (lane >= 8 && lane < 8 + p.PW) build_axis(xt[w], xc[w], lane - 8, RS_E, G.sw, G.bw, G.g, L.W, 1, 0);
    __syncwarp();
    int* tab = p.tables + (size_t)i * RS_TABW;
    if (lane < RS_P * RS_E) {
        const int pw = lane / RS_E, e = lane - pw * RS_E;
        const bool on = pw < p.PW && e < xc[w][pw];
        tab[lane] = on ? (xt[w][lane].off - xs) * RS_CS : 0;
        tab[28 + lane] = on ? __float_as_int(xt[w][lane].w) : 0;
    }
    if (lane < RS_P) tab[56 + lane] = lane < p.PW ? xc[w][lane] : 0;
    // lane rr gathers the pairs of feature row y0 + rr; a warp scan puts them in row order
    int cnt = 0, code[RS_P]; float wt[RS_P];
    const int nrows = it.y1 - it.y0 + 1;
    if (lane < nrows) {
        const int row = it.y0 + lane;
        for (int ph = p0; ph < p1; ++ph)
            for (int a = 0; a < yc[w][ph]; ++a)
                if (yt[w][ph * RS_E + a].off == row && cnt < RS_P) { code[cnt] = (lane << 8) | ph; wt[cnt] = yt[w][ph * RS_E + a].w; ++cnt; }
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(HD_FULL, incl, d); if (lane >= d) incl += y; }
    const int total = __shfl_sync(HD_FULL, incl, 31);
    int at = incl - cnt;
#pragma unroll
    for (int j = 0; j < RS_P; ++j)
        if (j < cnt && at + j < RS_P * RS_E) { tab[64 + at + j] = code[j]; tab[92 + at + j] = __float_as_int(wt[j]); }
    if (lane == 0) { tab[63] = min(total, RS_P * RS_E); tab[120] = it.k; tab[121] = p0 | (p1 << 8); tab[122] = it.y0; tab[123] = it.y1; }
}

__global__ void __launch_bounds__(RS_THREADS, 1) roi_align_strip_kernel(const __grid_constant__ RsParams p, const __grid_constant__ RsMaps maps,
                                                                        int ring_off, int tab_off) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long full[RS_MAXROWS];
    __shared__ int prog[RS_NCW];           // first row of the item each consumer warp is working on (INT_MAX: finished)
    __shared__ int issued;                 // rows the producer has issued so far (a barrier's 1-bit phase says nothing before that)
    __shared__ int tile_lock[RS_NTILE];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int u = blockIdx.y, c0 = blockIdx.x * RS_CS;
    const int s0 = p.unit_start[u], n = p.unit_start[u + 1] - s0;
    if (n <= 0) return;
    // bucket -> (image, level, strip)
    const int img = u / p.units_per_img, rem = u - img * p.units_per_img;
    int lvl = 0;
#pragma unroll
    for (int q = 1; q < HD_MAX_LEVELS; ++q)
        if (q < p.n_levels && rem >= p.lv[q].unit_base) lvl = q;
    const RsLevel L = p.lv[lvl];
    const int xs = (rem - L.unit_base) * L.SW;
    const RsItem* items = p.items + s0;
    const int* tables = p.tables + (size_t)s0 * RS_TABW;
    const int ya = max(items[0].y0, 0), yb = p.unit_ymax[u];
    const int NR = p.NR, row_floats = L.RW * RS_CS;
    const int nb = p.PH * p.PW, tstride = nb | 1;                 // odd tile stride: lane-per-channel writes hit 32 banks
    float* ring = reinterpret_cast<float*>(smem + ring_off);
    int* T = reinterpret_cast<int*>(smem + tab_off) + (size_t)(wid < RS_NCW ? wid : 0) * RS_TABW;
    if (tid == 0) {
        for (int s = 0; s < NR; ++s) rs_mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        issued = 0;
    }
    if (tid < RS_NTILE) tile_lock[tid] = 0;
    if (wid < RS_NCW && lane == 0) prog[wid] = wid < n ? items[wid].y0 : 0x7fffffff;
    __syncthreads();

    if (wid == RS_NCW) {
        // ================================================================ producer: one thread streams the rows ya..yb
        if (lane == 0) {
            const CUtensorMap* map = &maps.m[lvl];
            const unsigned row_bytes = (unsigned)row_floats * 4u;
            const int total = yb - ya + 1;
            for (int idx = 0; idx < total; ++idx) {
                const int slot = idx % NR, r = ya + idx;
                if (idx >= NR) {
                    const long long tp0 = clock64();
                    // the slot holds row r - NR: free once every unfinished item starts below it ...
                    for (unsigned spin = 0;; ++spin) {
                        int m = 0x7fffffff;
#pragma unroll
                        for (int w = 0; w < RS_NCW; ++w) m = min(m, *reinterpret_cast<volatile int*>(&prog[w]));
                        if (m > r - NR) break;
                        __nanosleep(128);
                        if (spin > (1u << 23)) __trap();
                    }
                    // ... and its previous load must have landed before the barrier is re-armed (rows nobody waited for)
                    rs_mbar_wait(&full[slot], (unsigned)(((idx / NR) - 1) & 1));
                    if (p.prof) atomicAdd(p.prof + 5, (unsigned long long)(clock64() - tp0));
                }
                rs_mbar_expect_tx(&full[slot], row_bytes);
                rs_tma_row(ring + (size_t)slot * row_floats, map, c0, xs, r, img, &full[slot]);
                __threadfence_block();
                *reinterpret_cast<volatile int*>(&issued) = idx + 1;
            }
            // every issued row must have landed before the CTA (and its shared memory) may go away
            for (int idx = max(total - NR, 0); idx < total; ++idx) rs_mbar_wait(&full[idx % NR], (unsigned)((idx / NR) & 1));
        }
        return;
    }

    // ==================================================================== consumers: lane = channel c0 + lane
    const float inv = 1.0f / (float)max(p.sr * p.sr, 1);
    const bool pow2 = ((p.sr * p.sr) & (p.sr * p.sr - 1)) == 0;
    int4 nxt = make_int4(0, 0, 0, 0);
    if (wid < n) nxt = __ldg(reinterpret_cast<const int4*>(tables + (size_t)wid * RS_TABW) + lane);   // table of the first item
    for (int i = wid; i < n; i += RS_NCW) {
        const unsigned tc0 = (unsigned)clock();
        __syncwarp();
        reinterpret_cast<int4*>(T)[lane] = nxt;                                   // this item's table -> shared memory
        if (i + RS_NCW < n) nxt = __ldg(reinterpret_cast<const int4*>(tables + (size_t)(i + RS_NCW) * RS_TABW) + lane);   // prefetch the next
        __syncwarp();
        const int k = T[120], pr = T[121], y0 = T[122], y1 = T[123], np = T[63];
        if (lane == 0) *reinterpret_cast<volatile int*>(&prog[wid]) = y0;         // rows below y0 are no longer needed by this warp
        const int p0 = pr & 255, p1 = pr >> 8;
        // Lane = (channel quad q, bin-column pair g): 8 quads x 4 pairs.  A quarter warp reads one cell's 32 channels as 128
        // contiguous bytes (conflict free), and every instruction of the warp serves all 32 channels of the item -- the
        // one-channel-per-lane form of this loop executed 2 900 instructions per item and was issue bound.
        const int q = lane & 7, pwA = 2 * (lane >> 3), pwB = pwA + 1;
        int xoA[RS_E], xoB[RS_E]; float wA[RS_E], wB[RS_E];
        const int nA = T[56 + pwA], nB = pwB < RS_P ? T[56 + pwB] : 0;
        {
            const int4 o4 = reinterpret_cast<const int4*>(T)[pwA];
            const float4 w4 = reinterpret_cast<const float4*>(T + 28)[pwA];
            xoA[0] = (o4.x >> 2) + q; xoA[1] = (o4.y >> 2) + q; xoA[2] = (o4.z >> 2) + q; xoA[3] = (o4.w >> 2) + q;   // float4 units
            wA[0] = w4.x; wA[1] = w4.y; wA[2] = w4.z; wA[3] = w4.w;
            const int4 p4 = reinterpret_cast<const int4*>(T)[pwB < RS_P ? pwB : pwA];
            const float4 v4 = reinterpret_cast<const float4*>(T + 28)[pwB < RS_P ? pwB : pwA];
            xoB[0] = (p4.x >> 2) + q; xoB[1] = (p4.y >> 2) + q; xoB[2] = (p4.z >> 2) + q; xoB[3] = (p4.w >> 2) + q;
            wB[0] = v4.x; wB[1] = v4.y; wB[2] = v4.z; wB[3] = v4.w;
        }
        const unsigned tc1 = (unsigned)clock();
        // The rows of this item must have landed.  A consumer can be many ring revolutions ahead of the producer (buckets with few,
        // scattered RoIs), where the 1-bit phase parity of a slot's barrier would alias: first wait until the last row has been
        // ISSUED -- from then on each of the item's slots is in, or one past, exactly the phase of its row.
        for (unsigned spin = 0; *reinterpret_cast<volatile int*>(&issued) <= y1 - ya; ++spin) {
            __nanosleep(256);                                                      // do not take issue slots from the working warps
            if (spin > (1u << 22)) __trap();
        }
        __threadfence_block();
        for (int r = y0; r <= y1; ++r) {
            const int idx = r - ya;
            rs_mbar_wait(&full[idx % NR], (unsigned)((idx / NR) & 1));
        }
        const unsigned tc2 = (unsigned)clock();
        float4 acc[RS_P][2];
#pragma unroll
        for (int a = 0; a < RS_P; ++a) { acc[a][0] = make_float4(0.f, 0.f, 0.f, 0.f); acc[a][1] = make_float4(0.f, 0.f, 0.f, 0.f); }
        // pairs (row, bin row, weight) in row order: the x-interpolated values of a row are formed once, when the row changes
        const int base_slot = (y0 - ya) % NR;
        int cur = -1;
        float4 tA = make_float4(0.f, 0.f, 0.f, 0.f), tB = tA;
        int code = np > 0 ? T[64] : 0;
        float wy = np > 0 ? __int_as_float(T[92]) : 0.0f;
        for (int j = 0; j < np; ++j) {
            const int ncode = j + 1 < np ? T[64 + j + 1] : 0;                     // next pair: in flight while this one is used
            const float nwy = j + 1 < np ? __int_as_float(T[92 + j + 1]) : 0.0f;
            const int rr = code >> 8;
            if (rr != cur) {
                cur = rr;
                int slot = base_slot + rr;
                if (slot >= NR) slot -= NR;
                const float4* __restrict__ row = reinterpret_cast<const float4*>(ring + (size_t)slot * row_floats);
                tA = make_float4(0.f, 0.f, 0.f, 0.f); tB = tA;
#pragma unroll
                for (int e = 0; e < RS_E; ++e) {
                    if (e < nA) { const float4 v = row[xoA[e]]; tA.x = fmaf(wA[e], v.x, tA.x); tA.y = fmaf(wA[e], v.y, tA.y); tA.z = fmaf(wA[e], v.z, tA.z); tA.w = fmaf(wA[e], v.w, tA.w); }
                    if (e < nB) { const float4 v = row[xoB[e]]; tB.x = fmaf(wB[e], v.x, tB.x); tB.y = fmaf(wB[e], v.y, tB.y); tB.z = fmaf(wB[e], v.z, tB.z); tB.w = fmaf(wB[e], v.w, tB.w); }
                }
            }
            switch (code & 255) {   // warp-uniform: the accumulators stay in statically indexed registers
#define RS_ROW(k_) case k_: acc[k_][0].x = fmaf(wy, tA.x, acc[k_][0].x); acc[k_][0].y = fmaf(wy, tA.y, acc[k_][0].y); acc[k_][0].z = fmaf(wy, tA.z, acc[k_][0].z); acc[k_][0].w = fmaf(wy, tA.w, acc[k_][0].w); \
                       acc[k_][1].x = fmaf(wy, tB.x, acc[k_][1].x); acc[k_][1].y = fmaf(wy, tB.y, acc[k_][1].y); acc[k_][1].z = fmaf(wy, tB.z, acc[k_][1].z); acc[k_][1].w = fmaf(wy, tB.w, acc[k_][1].w); break;
                RS_ROW(0) RS_ROW(1) RS_ROW(2) RS_ROW(3) RS_ROW(4) RS_ROW(5) RS_ROW(6)
#undef RS_ROW
                default: break;
            }
            code = ncode; wy = nwy;
        }
        const unsigned tc3 = (unsigned)clock();
        // scale, then out: full items through a shared tile (contiguous [32, PH*PW] block, 128-bit stores), split items directly
        float* __restrict__ dst = p.out + ((size_t)k * p.C + c0) * nb;
        const bool whole = p0 == 0 && p1 == p.PH;
        float* tile = nullptr;
        int ts = 0;
        if (whole) {   // take one of the RS_NTILE tiles (the output phase is ~10% of an item, so the tiles are rarely all busy)
            if (lane == 0) {
                ts = wid % RS_NTILE;
                for (unsigned spin = 0; atomicCAS(&tile_lock[ts], 0, 1) != 0; ++spin) {
                    ts = ts + 1 == RS_NTILE ? 0 : ts + 1;
                    if (spin > (1u << 26)) __trap();
                }
            }
            ts = __shfl_sync(HD_FULL, ts, 0);
            tile = reinterpret_cast<float*>(smem) + (size_t)ts * RS_CS * tstride;
        }
#pragma unroll
        for (int ph = 0; ph < RS_P; ++ph) {
            if (ph < p0 || ph >= p1) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int pw = pwA + h;
                if (pw >= p.PW) continue;
                const float4 a4 = acc[ph][h];
                const float v[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float o = pow2 ? v[c] * inv : __fdiv_rn(v[c], (float)(p.sr * p.sr));
                    if (whole) tile[(4 * q + c) * tstride + ph * p.PW + pw] = o;
                    else dst[(size_t)(4 * q + c) * nb + ph * p.PW + pw] = o;
                }
            }
        }
        if (whole) {
            __syncwarp();
            const int total = RS_CS * nb;
            if (tstride == nb) {
                for (int e = lane * 4; e < total; e += 128) *reinterpret_cast<float4*>(dst + e) = *reinterpret_cast<const float4*>(tile + e);
            } else {
                for (int e = lane * 4; e < total; e += 128) {
                    float v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { const int c = (e + q) / nb; v[q] = tile[c * tstride + (e + q) - c * nb]; }
                    *reinterpret_cast<float4*>(dst + e) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
            __syncwarp();
            if (lane == 0) { __threadfence_block(); atomicExch(&tile_lock[ts], 0); }
        }
        if (p.prof && lane == 0) {
            const unsigned tc4 = (unsigned)clock();
            atomicAdd(p.prof + 0, (unsigned long long)(tc1 - tc0)); atomicAdd(p.prof + 1, (unsigned long long)(tc2 - tc1));
            atomicAdd(p.prof + 2, (unsigned long long)(tc3 - tc2)); atomicAdd(p.prof + 3, (unsigned long long)(tc4 - tc3));
            atomicAdd(p.prof + 4, 1ull);
        }
    }
    if (lane == 0) *reinterpret_cast<volatile int*>(&prog[wid]) = 0x7fffffff;
}

// ------------------------------------------------------------------------------------------------ row-walk kernel
// One CTA per RoI (in the bucketed order, so neighbouring CTAs work on neighbouring rows of one feature map), thread = (channel quad,
// bin-column pair).  The RoI's feature rows are walked ONCE: per row a thread forms the x-interpolated value of its two bin columns
// from at most 2x4 128-bit loads, then adds it into the bin rows that row feeds (the (row, bin row, weight) pairs of the item table).
// The per-RoI gather kernel of roi.cu visits every row once per bin row it feeds (~1.8x the loads, ~4x the instructions); the sums
// are formed in the same order, so the results are identical.  Output through the [C, PH*PW] shared tile + one TMA bulk store.
__global__ void __launch_bounds__(512, 2) roi_align_rowwalk_kernel(const __grid_constant__ RsParams p, const __grid_constant__ RoiParams g) {
    extern __shared__ __align__(128) float tile[];           // [C][PH*PW]
    __shared__ __align__(16) int T[RS_TABW];
    const int i = blockIdx.x, tid = threadIdx.x;
    if (i >= *p.n_tmp) return;
    if (tid < 32) reinterpret_cast<int4*>(T)[tid] = __ldg(reinterpret_cast<const int4*>(p.tables + (size_t)i * RS_TABW) + tid);
    __syncthreads();
    const int k = T[120], y0 = T[122], np = T[63];
    const int unit = p.items[i].pr >> 16;
    const int img = unit / p.units_per_img, rem = unit - img * p.units_per_img;
    int lvl = 0;
#pragma unroll
    for (int q = 1; q < HD_MAX_LEVELS; ++q)
        if (q < p.n_levels && rem >= p.lv[q].unit_base) lvl = q;
    const int W = p.lv[lvl].W, nq = p.C >> 2, nb = p.PH * p.PW;
    const float4* __restrict__ f = reinterpret_cast<const float4*>(g.data[lvl]) + (size_t)img * p.lv[lvl].H * W * nq;
    const float inv = 1.0f / (float)max(p.sr * p.sr, 1);
    const bool pow2 = ((p.sr * p.sr) & (p.sr * p.sr - 1)) == 0;
    // thread = (channel quad, bin column): 64 quads x 8 column slots (PW <= 7 of them used); 2 CTAs = 32 warps per SM, because the
    // walk is a chain of dependent row loads and only many warps hide its latency
    const int groups = 64;
    const int pw = tid / groups;
    const bool on = pw < p.PW;
    int xo[RS_E]; float wx[RS_E];
    const int nx = on ? T[56 + pw] : 0;
    {   // table x offsets are (cell - strip start) * 32 with strip start 0 here: cell = off / 32
        const int4 o4 = reinterpret_cast<const int4*>(T)[on ? pw : 0];
        const float4 w4 = reinterpret_cast<const float4*>(T + 28)[on ? pw : 0];
        xo[0] = (o4.x >> 5) * nq; xo[1] = (o4.y >> 5) * nq; xo[2] = (o4.z >> 5) * nq; xo[3] = (o4.w >> 5) * nq;
        wx[0] = w4.x; wx[1] = w4.y; wx[2] = w4.z; wx[3] = w4.w;
    }
    // The walk below is a chain of dependent row loads, and on a pyramid larger than L2 every link would pay a full DRAM round
    // trip (~1 us loaded: measured 15 us per RoI).  So every thread first asks L2 for all the lines it is going to read.
    {
        const int nrows = (p.dbg & 1) ? 0 : T[123] - y0 + 1;
        for (int q = tid % groups; q < nq && on; q += groups)
            for (int rr = 0; rr < nrows; ++rr) {
                const float4* row = f + (size_t)(y0 + rr) * W * nq + q;
#pragma unroll
                for (int e = 0; e < RS_E; ++e)
                    if (e < nx) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + xo[e]));
            }
    }
    for (int q = tid % groups; q < nq && on; q += groups) {
        float4 acc[RS_P];
#pragma unroll
        for (int a = 0; a < RS_P; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
        int cur = -1;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < ((p.dbg & 2) ? 0 : np); ++j) {
            const int code = T[64 + j];
            const float wy = __int_as_float(T[92 + j]);
            const int rr = code >> 8;
            if (rr != cur) {
                cur = rr;
                const float4* __restrict__ row = f + (size_t)(y0 + rr) * W * nq + q;
                float4 v[RS_E];
#pragma unroll
                for (int e = 0; e < RS_E; ++e) v[e] = e < nx ? __ldg(row + xo[e]) : make_float4(0.f, 0.f, 0.f, 0.f);   // up to 4 requests in flight
                t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int e = 0; e < RS_E; ++e)
                    if (e < nx) { t.x = fmaf(wx[e], v[e].x, t.x); t.y = fmaf(wx[e], v[e].y, t.y); t.z = fmaf(wx[e], v[e].z, t.z); t.w = fmaf(wx[e], v[e].w, t.w); }
            }
            switch (code & 255) {   // CTA-uniform: the accumulators stay in statically indexed registers
#define RS_ROW(k_) case k_: acc[k_].x = fmaf(wy, t.x, acc[k_].x); acc[k_].y = fmaf(wy, t.y, acc[k_].y); acc[k_].z = fmaf(wy, t.z, acc[k_].z); acc[k_].w = fmaf(wy, t.w, acc[k_].w); break;
                RS_ROW(0) RS_ROW(1) RS_ROW(2) RS_ROW(3) RS_ROW(4) RS_ROW(5) RS_ROW(6)
#undef RS_ROW
                default: break;
            }
        }
        const int rot = (tid & 31) >> 3;                      // lane-rotated store order: the 32 lanes of a step hit 32 banks
#pragma unroll
        for (int ph = 0; ph < RS_P; ++ph) {
            if (ph >= p.PH) continue;
            const float4 a4 = acc[ph];
            float v[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = pow2 ? v[c] * inv : __fdiv_rn(v[c], (float)(p.sr * p.sr));
            float* tq = tile + (size_t)(4 * q) * nb + ph * p.PW + pw;
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {
                const int c = (s2 + rot) & 3;
                tq[c * nb] = c == 0 ? v[0] : (c == 1 ? v[1] : (c == 2 ? v[2] : v[3]));
            }
        }
    }
    // tile -> out[k]: one TMA bulk store
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0 && !(p.dbg & 4)) {
        const unsigned saddr = (unsigned)__cvta_generic_to_shared(tile);
        float* gdst = p.out + (size_t)k * p.C * nb;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(p.C * nb * 4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*RsEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static RsEncodeFn rs_encode_fn() {
    static RsEncodeFn fn = nullptr;   // process-wide driver entry point (not per device); racing initialisers store the same value
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (RsEncodeFn)f;
        else cudaGetLastError();
    }
    return fn;
}

static unsigned long long* g_rs_prof = nullptr;   // developer aid (hd_debug_roi_profile): device counters of the streamed kernel
extern "C" HD_API int hd_debug_roi_profile(int enable, unsigned long long* out8 /*host, nullable*/) {
    if (enable && !g_rs_prof) { HD_CUDA_CALL(cudaMalloc(&g_rs_prof, 64)); HD_CUDA_CALL(cudaMemset(g_rs_prof, 0, 64)); }
    if (out8 && g_rs_prof) { HD_CUDA_CALL(cudaDeviceSynchronize()); HD_CUDA_CALL(cudaMemcpy(out8, g_rs_prof, 64, cudaMemcpyDeviceToHost)); HD_CUDA_CALL(cudaMemset(g_rs_prof, 0, 64)); }
    if (!enable && g_rs_prof) { cudaFree(g_rs_prof); g_rs_prof = nullptr; }
    return HD_OK;
}

static const int RS_NR = 24;
static void rs_ws_layout(int n_units, long long K, size_t* offs, size_t* total) {
    size_t o = 0;
    const size_t nu = (size_t)n_units, k2 = (size_t)K * 2;
    offs[0] = o; o = hd_align_up(o + (nu * 3 + 4) * 4, 256);      // unit_count | unit_cursor | unit_ymax | n_tmp, fb_count  (zeroed every call)
    offs[1] = o; o = hd_align_up(o + (nu + 1) * 4, 256);          // unit_start
    offs[2] = o; o = hd_align_up(o + k2 * sizeof(RsItem), 256);   // tmp
    offs[3] = o; o = hd_align_up(o + k2 * 4, 256);                // tmp_unit
    offs[4] = o; o = hd_align_up(o + k2 * sizeof(RsItem), 256);   // scat
    offs[5] = o; o = hd_align_up(o + k2 * sizeof(RsItem), 256);   // items
    offs[6] = o; o = hd_align_up(o + (size_t)K * 4, 256);         // fb_list
    offs[7] = o; o = hd_align_up(o + k2 * RS_TABW * 4, 256);      // per-item tables
    *total = o;
}

// levels -> strips; returns false when the streamed kernel does not apply to this call
static bool rs_plan(RsParams& p, const hd_roi_level* levels, int n_levels, int C, int batch, int PH, int PW, int sr, int64_t K, size_t* smem_out,
                    int* ring_off, int* tab_off, bool streamed) {
    if (C % RS_CS != 0 || PH < 1 || PW < 1 || PH > RS_P || PW > RS_P || sr < 1 || sr > 2 || batch < 1 || K < 512 || K >= (1ll << 30)) return false;
    const int nb = PH * PW, tstride = nb | 1;
    const size_t tiles = hd_align_up((size_t)RS_NTILE * RS_CS * tstride * 4, 128);
    const size_t tabs = hd_align_up((size_t)RS_NCW * RS_TABW * 4, 128);
    const size_t budget = 224 * 1024;   // dynamic shared memory; the 227 KB of an SM also hold this kernel's static barriers
    if (tiles + tabs + (size_t)RS_NR * (RS_HALO + 8) * RS_CS * 4 > budget) return false;
    const int rw_cap = streamed ? (int)((budget - tiles - tabs) / ((size_t)RS_NR * RS_CS * 4)) : 65535;   // row-walk: no strips
    int units = 0, rw_max = 0;
    for (int l = 0; l < n_levels; ++l) {
        RsLevel& L = p.lv[l];
        L.H = levels[l].H; L.W = levels[l].W; L.scale = levels[l].spatial_scale;
        if (L.H > RS_MAXH || L.W > 65535 || (((uintptr_t)levels[l].data) & 15) != 0) return false;
        if (L.W <= rw_cap && (L.W <= 256 || !streamed)) { L.S = 1; L.SW = L.W; L.RW = L.W; }
        else {
            const int rw = rw_cap < 256 ? rw_cap : 256;
            if (rw - RS_HALO < 8) return false;
            L.SW = rw - RS_HALO; L.S = (L.W + L.SW - 1) / L.SW; L.RW = rw;
        }
        L.unit_base = units; units += L.S;
        if (L.RW > rw_max) rw_max = L.RW;
    }
    if ((long long)units * batch > 65535) return false;
    p.n_levels = n_levels; p.C = C; p.PH = PH; p.PW = PW; p.sr = sr; p.B = batch; p.NR = streamed ? RS_NR : (1 << 20);   // row-walk: never split
    p.units_per_img = units; p.n_units = units * batch; p.K = K;
    *ring_off = (int)tiles; *tab_off = (int)(tiles + hd_align_up((size_t)RS_NR * rw_max * RS_CS * 4, 128));
    *smem_out = (size_t)*tab_off + tabs;
    if (!streamed) { *smem_out = (size_t)C * nb * 4; return C % 4 == 0 && (C * nb * 4) % 16 == 0 && *smem_out <= 200 * 1024; }
    return *smem_out <= 226 * 1024;
}

extern "C" HD_API size_t hd_roi_align_workspace_size(const hd_roi_level* levels, int n_levels, int C, int batch, int64_t K, int pooled_h,
                                                     int pooled_w, int sampling_ratio) {
    RsParams p;
    memset(&p, 0, sizeof(p));
    size_t smem; int ro, to;
    if (!levels || n_levels < 1 || n_levels > HD_MAX_LEVELS || K < 0) return 0;
    const int mode = hd_roi_mode() & 15;
    if (mode != 2 && mode != 3) return 256;   // the default per-RoI gather kernels need no workspace
    if (!rs_plan(p, levels, n_levels, C, batch, pooled_h, pooled_w, sampling_ratio, K, &smem, &ro, &to, mode == 2)) return 256;
    size_t offs[8], total;
    rs_ws_layout(p.n_units, K, offs, &total);
    return total + 256;
}

extern "C" HD_API int hd_roi_align_ws(const hd_roi_level* levels, int n_levels, int layout, int C, int batch, const float* rois,
                                      const int32_t* level_ids, int64_t K, int pooled_h, int pooled_w, int sampling_ratio, int aligned,
                                      float* out, void* workspace, size_t workspace_bytes, void* stream) {
    HD_CHECK_ARG(levels != nullptr && n_levels >= 1 && n_levels <= HD_MAX_LEVELS, "n_levels must be in [1,%d], got %d", HD_MAX_LEVELS, n_levels);
    RsParams p;
    memset(&p, 0, sizeof(p));
    size_t smem = 0; int ring_off = 0, tab_off = 0;
    RsEncodeFn enc = rs_encode_fn();
    // hd_roi_set_mode: 0 / 1 per-RoI gather kernels (the fastest measured on B200: 1.25 ms on cfg3), 2 streamed (TMA ring) kernel
    // (2.46 ms), 3 row-walk kernel over the bucketed RoIs (1.80 ms).  2 and 3 are bit-identical to the gather kernels and kept as
    // tested, opt-in experiments; DESIGN.md section 3 has the measurements and why they lose.
    const int mode = hd_roi_mode() & 15;
    const bool streamed = mode == 2;
    bool ok = layout == HD_LAYOUT_NHWC && (mode == 2 || mode == 3) && (enc != nullptr || !streamed) && workspace != nullptr &&
              (n_levels == 1 || level_ids != nullptr) && (((uintptr_t)out) & 15) == 0 &&
              rs_plan(p, levels, n_levels, C, batch, pooled_h, pooled_w, sampling_ratio, K, &smem, &ring_off, &tab_off, streamed);
    size_t offs[8], total = 0;
    uintptr_t w0 = hd_align_up((uintptr_t)workspace, 256);
    if (ok) {
        rs_ws_layout(p.n_units, K, offs, &total);
        ok = w0 + total <= (uintptr_t)workspace + workspace_bytes;
    }
    RsMaps maps;
    if (ok && streamed) {
        memset(&maps, 0, sizeof(maps));
        for (int l = 0; l < n_levels && ok; ++l) {
            const RsLevel& L = p.lv[l];
            cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)L.W, (cuuint64_t)L.H, (cuuint64_t)batch};
            cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)L.W * C * 4, (cuuint64_t)L.H * L.W * C * 4};
            cuuint32_t box[4] = {RS_CS, (cuuint32_t)L.RW, 1, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            ok = enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)levels[l].data, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
    }
    if (!ok)   // the streamed kernel does not apply (layout, sampling grid, few RoIs, no workspace ...): per-RoI gather kernels
        return hd_roi_align(levels, n_levels, layout, C, rois, level_ids, K, pooled_h, pooled_w, sampling_ratio, aligned, out, stream);
    HD_CHECK_ARG(rois && out, "rois/out is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    p.aligned = aligned ? 1 : 0; p.rois = rois; p.level_ids = level_ids; p.out = out;
    const size_t nu = (size_t)p.n_units;
    int* zero = (int*)(w0 + offs[0]);
    p.unit_count = zero; p.unit_cursor = zero + nu; p.unit_ymax = zero + 2 * nu; p.n_tmp = zero + 3 * nu; p.fb_count = zero + 3 * nu + 1;
    p.unit_start = (int*)(w0 + offs[1]);
    p.tmp = (RsItem*)(w0 + offs[2]); p.tmp_unit = (int*)(w0 + offs[3]); p.scat = (RsItem*)(w0 + offs[4]); p.items = (RsItem*)(w0 + offs[5]);
    p.fb_list = (int*)(w0 + offs[6]); p.tables = (int*)(w0 + offs[7]);
    p.prof = g_rs_prof; p.dbg = hd_roi_mode() >> 4;
    HD_CUDA_CALL(cudaMemsetAsync(zero, 0, (nu * 3 + 4) * 4, st));
    rs_prep_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(p);
    HD_CUDA_LAUNCH_CHECK("rs_prep_kernel");
    rs_scan_kernel<<<1, 1024, 0, st>>>(p);
    HD_CUDA_LAUNCH_CHECK("rs_scan_kernel");
    rs_scatter_kernel<<<(unsigned)((2 * K + 255) / 256), 256, 0, st>>>(p);
    HD_CUDA_LAUNCH_CHECK("rs_scatter_kernel");
    rs_sort_kernel<<<(unsigned)p.n_units, 256, 0, st>>>(p);
    HD_CUDA_LAUNCH_CHECK("rs_sort_kernel");
    rs_table_kernel<<<(unsigned)((2 * K + 7) / 8), 256, 0, st>>>(p);
    HD_CUDA_LAUNCH_CHECK("rs_table_kernel");
    RoiParams g;
    memset(&g, 0, sizeof(g));
    for (int l = 0; l < n_levels; ++l) { g.data[l] = levels[l].data; g.H[l] = levels[l].H; g.W[l] = levels[l].W; g.scale[l] = levels[l].spatial_scale; }
    g.n_levels = n_levels; g.C = C; g.PH = pooled_h; g.PW = pooled_w; g.sampling_ratio = sampling_ratio; g.aligned = p.aligned;
    g.rois = rois; g.level_ids = level_ids; g.K = K; g.out = out;
    if (streamed) {
        HD_ENSURE_SMEM(roi_align_strip_kernel, 226 * 1024);
        roi_align_strip_kernel<<<dim3((unsigned)(C / RS_CS), (unsigned)p.n_units), RS_THREADS, smem, st>>>(p, maps, ring_off, tab_off);
        HD_CUDA_LAUNCH_CHECK("roi_align_strip_kernel");
    } else {
        HD_ENSURE_SMEM(roi_align_rowwalk_kernel, 200 * 1024);
        roi_align_rowwalk_kernel<<<(unsigned)K, 512, smem, st>>>(p, g);   // items <= K (never split); CTAs beyond the item count exit
        HD_CUDA_LAUNCH_CHECK("roi_align_rowwalk_kernel");
    }
    // hand-back list (device-side count): gather kernel, grid-stride over the list
    long long ctas = K < 4 * (long long)hd_num_sms() ? K : 4 * (long long)hd_num_sms();
    return hd_roi_align_launch_list(g, p.fb_list, p.fb_count, (int)ctas, st);
}

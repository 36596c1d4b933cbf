// RoIAlign / RoIPool / FPN level assignment (a7-a9).  Replaces torchvision::roi_align / roi_pool
// (roi_align.py:204-260, roi_pool.py:15-53) and the LevelMapper / MultiScaleRoIAlign gather-scatter
// (poolers.py:73-84, 147-227).  Reference feature: README.md:65,73-78.
//
// Fast layout is channel-contiguous (NHWC / torch channels_last): one CTA per RoI, one thread per
// channel, so every corner fetch of a warp is one coalesced 128-byte line.  Bilinear sampling on the
// regular per-bin sample grid is separable, so each RoI first builds per-axis tables of
// (cell, merged weight) in shared memory; a bin then costs ny*nx (typically 4-9) loads instead of the
// 16*grid^2/4 corner reads of the sample-by-sample form.  The [C,PH,PW] result tile is assembled in
// shared memory in exactly the output layout and leaves with ONE TMA bulk store (cp.async.bulk).
#include "hd_common.cuh"
#include "hd_roi_axis.cuh"
#include "hd_roi_internal.cuh"

// one bin row with compile-time entry counts: all E*E loads of a bin are independent and issued together
template <int E>
__device__ __forceinline__ void roi_align_rows_fixed(const float* __restrict__ fc, const AxisEntry* ytab, const AxisEntry* xtab, int PH,
                                                     int PW, float count, float* tile_c) {
    for (int ph = 0; ph < PH; ++ph) {
        int yo[E]; float wy[E];
#pragma unroll
        for (int a = 0; a < E; ++a) { yo[a] = ytab[ph * E + a].off; wy[a] = ytab[ph * E + a].w; }
        for (int pw = 0; pw < PW; ++pw) {
            int xo[E]; float wx[E];
#pragma unroll
            for (int b = 0; b < E; ++b) { xo[b] = xtab[pw * E + b].off; wx[b] = xtab[pw * E + b].w; }
            float v[E][E];
#pragma unroll
            for (int a = 0; a < E; ++a)
#pragma unroll
                for (int b = 0; b < E; ++b) v[a][b] = __ldg(fc + yo[a] + xo[b]);
            float acc = 0.0f;
#pragma unroll
            for (int a = 0; a < E; ++a) {
                float r = 0.0f;
#pragma unroll
                for (int b = 0; b < E; ++b) r = fmaf(wx[b], v[a][b], r);
                acc = fmaf(wy[a], r, acc);
            }
            tile_c[ph * PW + pw] = __fdiv_rn(acc, count);
        }
    }
}

__device__ __forceinline__ void tile_store(float* __restrict__ gdst, const float* tile, int n_floats, bool use_tma) {
    if (use_tma) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned saddr = (unsigned)__cvta_generic_to_shared(tile);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(n_floats * 4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        __syncthreads();
        if ((n_floats & 3) == 0 && (((uintptr_t)gdst) & 15) == 0) {
            for (int i = threadIdx.x; i < (n_floats >> 2); i += blockDim.x) reinterpret_cast<float4*>(gdst)[i] = reinterpret_cast<const float4*>(tile)[i];
        } else {
            for (int i = threadIdx.x; i < n_floats; i += blockDim.x) gdst[i] = tile[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------ RoIAlign NHWC
__global__ void __launch_bounds__(256) roi_align_nhwc_kernel(const __grid_constant__ RoiParams p, int use_tma) {
    extern __shared__ __align__(128) float smem_f[];
    float* tile = smem_f;                                  // [C][PH*PW]
    AxisEntry* ytab = (AxisEntry*)(tile + (size_t)p.C * p.PH * p.PW);
    AxisEntry* xtab = ytab + ROI_TAB;
    __shared__ int ycnt[64], xcnt[64];

    const long long k = blockIdx.x;
    const float* roi = p.rois + k * 5;
    const int lvl = p.level_ids ? p.level_ids[k] : 0;
    const int H = p.H[lvl], W = p.W[lvl];
    const float sc = p.scale[lvl];
    const int bidx = (int)roi[0];
    const float off = p.aligned ? 0.5f : 0.0f;
    const float sw = __fsub_rn(__fmul_rn(roi[1], sc), off), sh = __fsub_rn(__fmul_rn(roi[2], sc), off);
    const float ew = __fsub_rn(__fmul_rn(roi[3], sc), off), eh = __fsub_rn(__fmul_rn(roi[4], sc), off);
    float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
    if (!p.aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
    const float bh = __fdiv_rn(rh, (float)p.PH), bw = __fdiv_rn(rw, (float)p.PW);
    const int gh = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)p.PH));
    const int gw = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)p.PW));
    const float count = (float)max(gh * gw, 1);
    // entries per bin and axis are <= 2*grid; the common cases get fixed widths (4: grid<=2, 8: grid<=4)
    const int need = 2 * max(max(gh, gw), 0);
    const int E = need <= 4 ? 4 : (need <= 8 ? 8 : 0);
    const int stride_y = E ? E : max(2 * max(gh, 0), 1), stride_x = E ? E : max(2 * max(gw, 0), 1);
    const bool fits = (long long)stride_y * p.PH <= ROI_TAB && (long long)stride_x * p.PW <= ROI_TAB;
    if (fits) {
        if (threadIdx.x < p.PH) build_axis(ytab, ycnt, threadIdx.x, stride_y, sh, bh, gh, H, W * p.C, E);
        else if (threadIdx.x >= 64 && threadIdx.x < 64 + p.PW) build_axis(xtab, xcnt, threadIdx.x - 64, stride_x, sw, bw, gw, W, p.C, E);
    }
    __syncthreads();
    const float* __restrict__ f = p.data[lvl] + (size_t)bidx * H * W * p.C;
    const int nb = p.PH * p.PW;
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
        const float* __restrict__ fc = f + c;
        if (fits && E == 4) { roi_align_rows_fixed<4>(fc, ytab, xtab, p.PH, p.PW, count, tile + c * nb); continue; }
        if (fits && E == 8) { roi_align_rows_fixed<8>(fc, ytab, xtab, p.PH, p.PW, count, tile + c * nb); continue; }
        for (int ph = 0; ph < p.PH; ++ph) {
            for (int pw = 0; pw < p.PW; ++pw) {
                float acc = 0.0f;
                if (fits) {
                    const AxisEntry* yt = ytab + ph * stride_y;
                    const AxisEntry* xt = xtab + pw * stride_x;
                    const int ny = ycnt[ph], nx = xcnt[pw];
                    for (int a = 0; a < ny; ++a) {
                        const float wy = yt[a].w;
                        const float* __restrict__ row = fc + yt[a].off;
                        float r = 0.0f;
                        int b = 0;
                        for (; b + 4 <= nx; b += 4) {
                            const float v0 = __ldg(row + xt[b].off), v1 = __ldg(row + xt[b + 1].off);
                            const float v2 = __ldg(row + xt[b + 2].off), v3 = __ldg(row + xt[b + 3].off);
                            r = fmaf(xt[b].w, v0, r); r = fmaf(xt[b + 1].w, v1, r); r = fmaf(xt[b + 2].w, v2, r); r = fmaf(xt[b + 3].w, v3, r);
                        }
                        for (; b < nx; ++b) r = fmaf(xt[b].w, __ldg(row + xt[b].off), r);
                        acc = fmaf(wy, r, acc);
                    }
                } else {
                    // huge adaptive grids: sample by sample, no tables
                    for (int iy = 0; iy < gh; ++iy) {
                        float y = __fadd_rn(__fadd_rn(sh, __fmul_rn((float)ph, bh)), __fdiv_rn(__fmul_rn(__fadd_rn((float)iy, 0.5f), bh), (float)gh));
                        if (y < -1.0f || y > (float)H) continue;
                        if (y <= 0.0f) y = 0.0f;
                        int yl = (int)y, yh;
                        if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
                        const float ly = y - (float)yl, hy = 1.0f - ly;
                        for (int ix = 0; ix < gw; ++ix) {
                            float x = __fadd_rn(__fadd_rn(sw, __fmul_rn((float)pw, bw)), __fdiv_rn(__fmul_rn(__fadd_rn((float)ix, 0.5f), bw), (float)gw));
                            if (x < -1.0f || x > (float)W) continue;
                            if (x <= 0.0f) x = 0.0f;
                            int xl = (int)x, xh;
                            if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
                            const float lx = x - (float)xl, hx = 1.0f - lx;
                            acc += hy * hx * __ldg(fc + ((size_t)yl * W + xl) * p.C) + hy * lx * __ldg(fc + ((size_t)yl * W + xh) * p.C) +
                                   ly * hx * __ldg(fc + ((size_t)yh * W + xl) * p.C) + ly * lx * __ldg(fc + ((size_t)yh * W + xh) * p.C);
                        }
                    }
                }
                tile[c * nb + ph * p.PW + pw] = __fdiv_rn(acc, count);
            }
        }
    }
    tile_store(p.out + (size_t)k * p.C * nb, tile, p.C * nb, use_tma != 0);
}

// ------------------------------------------------------------------------------------------------ RoIAlign NHWC, float4
// Channel-quad variant (C % 4 == 0): a thread owns 4 consecutive channels (one 128-bit load per corner), the 256
// threads form 256/QT bin groups that walk the PH*PW bins side by side.  ~5x fewer instructions per channel than
// the scalar kernel; zero-weight table padding is predicated off so it costs no L1 bandwidth; the four results
// of a thread are written to the [C][PH*PW] tile in a lane-rotated order that is free of bank conflicts.
// one table row (fixed y cell) of a bin: NX real x entries, all loads independent
template <int NX>
__device__ __forceinline__ float4 roi_row_quad(const float4* __restrict__ row, const int* xo, const float* wx) {
    float4 v[NX];
#pragma unroll
    for (int b = 0; b < NX; ++b) v[b] = __ldg(row + xo[b]);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int b = 0; b < NX; ++b) { r.x = fmaf(wx[b], v[b].x, r.x); r.y = fmaf(wx[b], v[b].y, r.y); r.z = fmaf(wx[b], v[b].z, r.z); r.w = fmaf(wx[b], v[b].w, r.w); }
    return r;
}

// R table rows (fixed y cells) of a bin with their NX x entries each: R*NX independent loads in flight, then the same fma order
// as the row-at-a-time form (r = sum_b wx[b] v[b]; acc = fma(wy, r, acc), rows ascending)
template <int R, int NX>
__device__ __forceinline__ void roi_rows_quad(const float* __restrict__ f, int q, const AxisEntry* yt, const int* xo, const float* wx, float4& acc, float4& last) {
    float4 v[R][NX];
#pragma unroll
    for (int a = 0; a < R; ++a) {
        const float4* __restrict__ row = reinterpret_cast<const float4*>(f + yt[a].off) + q;
#pragma unroll
        for (int b = 0; b < NX; ++b) v[a][b] = __ldg(row + xo[b]);
    }
#pragma unroll
    for (int a = 0; a < R; ++a) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int b = 0; b < NX; ++b) { r.x = fmaf(wx[b], v[a][b].x, r.x); r.y = fmaf(wx[b], v[a][b].y, r.y); r.z = fmaf(wx[b], v[a][b].z, r.z); r.w = fmaf(wx[b], v[a][b].w, r.w); }
        const float wy = yt[a].w;
        acc.x = fmaf(wy, r.x, acc.x); acc.y = fmaf(wy, r.y, acc.y); acc.z = fmaf(wy, r.z, acc.z); acc.w = fmaf(wy, r.w, acc.w);
        if (a == R - 1) last = r;
    }
}
template <int R, int NX>
__device__ __forceinline__ void roi_rows_quad(const float* __restrict__ f, int q, const AxisEntry* yt, const int* xo, const float* wx, float4& acc) {
    float4 v[R][NX];
#pragma unroll
    for (int a = 0; a < R; ++a) {
        const float4* __restrict__ row = reinterpret_cast<const float4*>(f + yt[a].off) + q;
#pragma unroll
        for (int b = 0; b < NX; ++b) v[a][b] = __ldg(row + xo[b]);
    }
#pragma unroll
    for (int a = 0; a < R; ++a) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int b = 0; b < NX; ++b) { r.x = fmaf(wx[b], v[a][b].x, r.x); r.y = fmaf(wx[b], v[a][b].y, r.y); r.z = fmaf(wx[b], v[a][b].z, r.z); r.w = fmaf(wx[b], v[a][b].w, r.w); }
        const float wy = yt[a].w;
        acc.x = fmaf(wy, r.x, acc.x); acc.y = fmaf(wy, r.y, acc.y); acc.z = fmaf(wy, r.z, acc.z); acc.w = fmaf(wy, r.w, acc.w);
    }
}
template <int NX, int MLP>
__device__ __forceinline__ void roi_bin_rows(const float* __restrict__ f, int q, const AxisEntry* yt, int ny, const int* xo, const float* wx, float4& acc) {
    int a = 0;
    for (; a + 2 <= ny; a += 2) roi_rows_quad<2, NX>(f, q, yt + a, xo, wx, acc);
    if (a < ny) roi_rows_quad<1, NX>(f, q, yt + a, xo, wx, acc);
}

template <int E, int MLP = 1>
__device__ __forceinline__ void roi_align_bins_quad(const float* __restrict__ f, int q, int g, int groups, const AxisEntry* ytab,
                                                    const AxisEntry* xtab, const int* ycnt, const int* xcnt, int PH, int PW, float count,
                                                    float* tile, int rot) {
    const int nb = PH * PW;
    const unsigned magic = 0xffffffffu / (unsigned)PW + 1u;            // bin / PW == umulhi(bin, magic) for bin < 2^16 (PW > 1)
    const int icount = (int)count;
    const bool pow2 = (icount & (icount - 1)) == 0;                   // x / 2^k == x * 2^-k exactly
    const float inv = 1.0f / count;
    for (int bin = g; bin < nb; bin += groups) {
        const int ph = (PW > 1) ? (int)__umulhi((unsigned)bin, magic) : bin, pw = bin - ph * PW;
        int xo[E]; float wx[E];
#pragma unroll
        for (int b = 0; b < E; ++b) { xo[b] = xtab[pw * E + b].off >> 2; wx[b] = xtab[pw * E + b].w; }
        const int ny = ycnt[ph], nx = xcnt[pw];   // real (merged) entries; the rest of the table is zero-weight padding
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (E == 4 && MLP > 1) {
            const AxisEntry* yt = ytab + ph * E;
            if (nx == 2) roi_bin_rows<2, MLP>(f, q, yt, ny, xo, wx, acc);
            else if (nx == 3) roi_bin_rows<3, MLP>(f, q, yt, ny, xo, wx, acc);
            else if (nx == 4) roi_bin_rows<4, MLP>(f, q, yt, ny, xo, wx, acc);
            else if (nx == 1) roi_bin_rows<1, MLP>(f, q, yt, ny, xo, wx, acc);
        } else if (E == 4) {
            for (int a = 0; a < ny; ++a) {
                const float wy = ytab[ph * E + a].w;
                const float4* __restrict__ row = reinterpret_cast<const float4*>(f + ytab[ph * E + a].off) + q;
                float4 r;   // warp-uniform switch: only the real cells are fetched
                if (nx == 2) r = roi_row_quad<2>(row, xo, wx);
                else if (nx == 3) r = roi_row_quad<3>(row, xo, wx);
                else if (nx == 4) r = roi_row_quad<4>(row, xo, wx);
                else if (nx == 1) r = roi_row_quad<1>(row, xo, wx);
                else continue;
                acc.x = fmaf(wy, r.x, acc.x); acc.y = fmaf(wy, r.y, acc.y); acc.z = fmaf(wy, r.z, acc.z); acc.w = fmaf(wy, r.w, acc.w);
            }
        } else {
#pragma unroll
            for (int a = 0; a < E; ++a) {   // wide tables: fully unrolled over the padded table (zero-weight rows skipped)
                const float wy = ytab[ph * E + a].w;
                if (wy == 0.0f) continue;
                const float4* __restrict__ row = reinterpret_cast<const float4*>(f + ytab[ph * E + a].off) + q;
                const float4 r = roi_row_quad<E>(row, xo, wx);
                acc.x = fmaf(wy, r.x, acc.x); acc.y = fmaf(wy, r.y, acc.y); acc.z = fmaf(wy, r.z, acc.z); acc.w = fmaf(wy, r.w, acc.w);
            }
        }
        if (pow2) { acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv; }
        else { acc.x = __fdiv_rn(acc.x, count); acc.y = __fdiv_rn(acc.y, count); acc.z = __fdiv_rn(acc.z, count); acc.w = __fdiv_rn(acc.w, count); }
        // lane-rotated store order (rot = lane/8): at every step the 32 lanes hit 32 different banks
        float* t = tile + (size_t)(4 * q) * nb + bin;
        const float a0 = (rot & 1) ? acc.y : acc.x, a1 = (rot & 1) ? acc.z : acc.y, a2 = (rot & 1) ? acc.w : acc.z, a3 = (rot & 1) ? acc.x : acc.w;
        const float b0 = (rot & 2) ? a2 : a0, b1 = (rot & 2) ? a3 : a1, b2 = (rot & 2) ? a0 : a2, b3 = (rot & 2) ? a1 : a3;  // b_s = acc[(s+rot)&3]
        t[((0 + rot) & 3) * nb] = b0; t[((1 + rot) & 3) * nb] = b1; t[((2 + rot) & 3) * nb] = b2; t[((3 + rot) & 3) * nb] = b3;
    }
}

// EFIX = table width fixed by the host from sampling_ratio (4: sr<=2, 8: sr<=4, 16: sr<=8; small tables, the sr<=2 variant
// runs 4 CTAs/SM); EFIX = 0: adaptive sampling, width chosen per RoI (3 CTAs/SM, more registers)
template <int EFIX, int MLP = 1>
__global__ void __launch_bounds__(256, (EFIX == 4 && MLP == 1) ? 4 : 3) roi_align_nhwc_quad_kernel(const __grid_constant__ RoiParams p, int use_tma, int QT, int tab) {
    extern __shared__ __align__(128) float smem_f[];
    float* tile = smem_f;                                  // [C][PH*PW]
    AxisEntry* ytab = (AxisEntry*)(tile + (size_t)p.C * p.PH * p.PW);
    AxisEntry* xtab = ytab + tab;                           // tab entries per axis (host: small when sampling_ratio is fixed)
    __shared__ int ycnt[64], xcnt[64];

    // one RoI per CTA, or (p.list) a grid-stride walk over a device-side list of RoI indices
    const long long nk = p.list ? (long long)*p.list_count : p.K;
    for (long long kk = blockIdx.x; kk < nk; kk += gridDim.x) {
    const long long k = p.list ? (long long)p.list[kk] : kk;
    const float* roi = p.rois + k * 5;
    const int lvl = p.level_ids ? p.level_ids[k] : 0;
    const int H = p.H[lvl], W = p.W[lvl];
    const float sc = p.scale[lvl];
    const int bidx = (int)roi[0];
    const float off = p.aligned ? 0.5f : 0.0f;
    const float sw = __fsub_rn(__fmul_rn(roi[1], sc), off), sh = __fsub_rn(__fmul_rn(roi[2], sc), off);
    const float ew = __fsub_rn(__fmul_rn(roi[3], sc), off), eh = __fsub_rn(__fmul_rn(roi[4], sc), off);
    float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
    if (!p.aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
    const float bh = __fdiv_rn(rh, (float)p.PH), bw = __fdiv_rn(rw, (float)p.PW);
    const int gh = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)p.PH));
    const int gw = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)p.PW));
    const float count = (float)max(gh * gw, 1);
    const int need = 2 * max(max(gh, gw), 0);
    const int E = EFIX ? EFIX : (need <= 4 ? 4 : (need <= 8 ? 8 : (need <= 16 ? 16 : 0)));
    const int nb = p.PH * p.PW;
    const float* __restrict__ f = p.data[lvl] + (size_t)bidx * H * W * p.C;
    if (E && E * p.PH <= tab && E * p.PW <= tab) {
        if (threadIdx.x < p.PH) build_axis(ytab, ycnt, threadIdx.x, E, sh, bh, gh, H, W * p.C, E);
        else if (threadIdx.x >= 64 && threadIdx.x < 64 + p.PW) build_axis(xtab, xcnt, threadIdx.x - 64, E, sw, bw, gw, W, p.C, E);
        __syncthreads();
        const int nq = p.C >> 2, groups = 256 / QT, g = threadIdx.x / QT, rot = (threadIdx.x & 31) >> 3;
        for (int q = threadIdx.x % QT; q < nq; q += QT) {
            if (EFIX) roi_align_bins_quad<(EFIX ? EFIX : 4), MLP>(f, q, g, groups, ytab, xtab, ycnt, xcnt, p.PH, p.PW, count, tile, rot);
            else if (E == 4) roi_align_bins_quad<4>(f, q, g, groups, ytab, xtab, ycnt, xcnt, p.PH, p.PW, count, tile, rot);
            else if (E == 8) roi_align_bins_quad<8>(f, q, g, groups, ytab, xtab, ycnt, xcnt, p.PH, p.PW, count, tile, rot);
            else roi_align_bins_quad<16>(f, q, g, groups, ytab, xtab, ycnt, xcnt, p.PH, p.PW, count, tile, rot);
        }
    } else if ((long long)2 * max(gh, 1) * p.PH <= tab && (long long)2 * max(gw, 1) * p.PW <= tab) {
        // large adaptive grids: merged tables with run-time entry counts (<= bin size + 1 cells per axis)
        const int stride_y = 2 * max(gh, 1), stride_x = 2 * max(gw, 1);
        if (threadIdx.x < p.PH) build_axis(ytab, ycnt, threadIdx.x, stride_y, sh, bh, gh, H, W * p.C, 0);
        else if (threadIdx.x >= 64 && threadIdx.x < 64 + p.PW) build_axis(xtab, xcnt, threadIdx.x - 64, stride_x, sw, bw, gw, W, p.C, 0);
        __syncthreads();
        const int nq = p.C >> 2, groups = 256 / QT, g = threadIdx.x / QT, rot = (threadIdx.x & 31) >> 3;
        for (int q = threadIdx.x % QT; q < nq; q += QT) {
            for (int bin = g; bin < nb; bin += groups) {
                const int ph = bin / p.PW, pw = bin - ph * p.PW;
                const AxisEntry* yt = ytab + ph * stride_y;
                const AxisEntry* xt = xtab + pw * stride_x;
                const int ny = ycnt[ph], nx = xcnt[pw];
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int a = 0; a < ny; ++a) {
                    const float4* __restrict__ row = reinterpret_cast<const float4*>(f + yt[a].off) + q;
                    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                    int b = 0;
                    for (; b + 4 <= nx; b += 4) {
                        const float4 v0 = __ldg(row + (xt[b].off >> 2)), v1 = __ldg(row + (xt[b + 1].off >> 2));
                        const float4 v2 = __ldg(row + (xt[b + 2].off >> 2)), v3 = __ldg(row + (xt[b + 3].off >> 2));
                        const float w0 = xt[b].w, w1 = xt[b + 1].w, w2 = xt[b + 2].w, w3 = xt[b + 3].w;
                        r.x = fmaf(w0, v0.x, r.x); r.y = fmaf(w0, v0.y, r.y); r.z = fmaf(w0, v0.z, r.z); r.w = fmaf(w0, v0.w, r.w);
                        r.x = fmaf(w1, v1.x, r.x); r.y = fmaf(w1, v1.y, r.y); r.z = fmaf(w1, v1.z, r.z); r.w = fmaf(w1, v1.w, r.w);
                        r.x = fmaf(w2, v2.x, r.x); r.y = fmaf(w2, v2.y, r.y); r.z = fmaf(w2, v2.z, r.z); r.w = fmaf(w2, v2.w, r.w);
                        r.x = fmaf(w3, v3.x, r.x); r.y = fmaf(w3, v3.y, r.y); r.z = fmaf(w3, v3.z, r.z); r.w = fmaf(w3, v3.w, r.w);
                    }
                    for (; b < nx; ++b) {
                        const float4 v0 = __ldg(row + (xt[b].off >> 2));
                        const float w0 = xt[b].w;
                        r.x = fmaf(w0, v0.x, r.x); r.y = fmaf(w0, v0.y, r.y); r.z = fmaf(w0, v0.z, r.z); r.w = fmaf(w0, v0.w, r.w);
                    }
                    const float wy = yt[a].w;
                    acc.x = fmaf(wy, r.x, acc.x); acc.y = fmaf(wy, r.y, acc.y); acc.z = fmaf(wy, r.z, acc.z); acc.w = fmaf(wy, r.w, acc.w);
                }
                acc.x = __fdiv_rn(acc.x, count); acc.y = __fdiv_rn(acc.y, count); acc.z = __fdiv_rn(acc.z, count); acc.w = __fdiv_rn(acc.w, count);
                float* t = tile + (size_t)(4 * q) * nb + bin;
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2) {
                    const int k2 = (s2 + rot) & 3;
                    t[k2 * nb] = (k2 == 0) ? acc.x : (k2 == 1) ? acc.y : (k2 == 2) ? acc.z : acc.w;
                }
            }
        }
    } else {
        // grids beyond the table capacity: sample by sample (reference order), one channel per thread
        for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
            const float* __restrict__ fc = f + c;
            for (int ph = 0; ph < p.PH; ++ph)
                for (int pw = 0; pw < p.PW; ++pw) {
                    float acc = 0.0f;
                    for (int iy = 0; iy < gh; ++iy) {
                        float y = __fadd_rn(__fadd_rn(sh, __fmul_rn((float)ph, bh)), __fdiv_rn(__fmul_rn(__fadd_rn((float)iy, 0.5f), bh), (float)gh));
                        if (y < -1.0f || y > (float)H) continue;
                        if (y <= 0.0f) y = 0.0f;
                        int yl = (int)y, yh;
                        if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
                        const float ly = y - (float)yl, hy = 1.0f - ly;
                        for (int ix = 0; ix < gw; ++ix) {
                            float x = __fadd_rn(__fadd_rn(sw, __fmul_rn((float)pw, bw)), __fdiv_rn(__fmul_rn(__fadd_rn((float)ix, 0.5f), bw), (float)gw));
                            if (x < -1.0f || x > (float)W) continue;
                            if (x <= 0.0f) x = 0.0f;
                            int xl = (int)x, xh;
                            if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
                            const float lx = x - (float)xl, hx = 1.0f - lx;
                            acc += hy * hx * __ldg(fc + ((size_t)yl * W + xl) * p.C) + hy * lx * __ldg(fc + ((size_t)yl * W + xh) * p.C) +
                                   ly * hx * __ldg(fc + ((size_t)yh * W + xl) * p.C) + ly * lx * __ldg(fc + ((size_t)yh * W + xh) * p.C);
                        }
                    }
                    tile[c * nb + ph * p.PW + pw] = __fdiv_rn(acc, count);
                }
        }
    }
    tile_store(p.out + (size_t)k * p.C * nb, tile, p.C * nb, use_tma != 0);
    __syncthreads();   // the tile and the tables are rewritten by the next RoI of the walk
    }
}

// ------------------------------------------------------------------------------------------------ RoIAlign NHWC, column owners
// Same tables, loads and fma order as the channel-quad kernel above (bit-identical results), different loop nest: warp = output
// column pw, lane = channel quad.  The x entries of a column (offsets, weights, their count) are fixed for the whole RoI, so they sit
// in registers and the code path is chosen once per warp instead of once per bin row; the bin loop is ph = 0..PH-1 with nothing
// but the y entries (broadcast shared-memory reads) changing.  ~2.3x fewer instructions per bin than the bin-group walk, which was
// half issue-bound (ncu: sm 53 %, 210 instructions per bin and thread of which ~80 are loads and fmas).
template <int NX>
__device__ __forceinline__ void roi_align_col_quad(const float* __restrict__ f, int lane, int nq, int pw, const AxisEntry* ytab, const AxisEntry* xt,
                                                   const int* ycnt, int PH, int PW, float count, float* tile) {
    int xo[NX]; float wx[NX];
#pragma unroll
    for (int b = 0; b < NX; ++b) { xo[b] = xt[b].off >> 2; wx[b] = xt[b].w; }
    const int nb = PH * PW;
    const int icount = (int)count;
    const bool pow2 = (icount & (icount - 1)) == 0;                   // x / 2^k == x * 2^-k exactly
    const float inv = 1.0f / count;
    const int rot = lane >> 3;
    for (int q = lane; q < nq; q += 32) {
        float* tq = tile + (size_t)(4 * q) * nb + pw;
        // Consecutive bins of a column usually share their boundary cell row (bin height ~1.8 cells at sampling_ratio 2: cells
        // {0,1,2} then {2,3,4}).  Its x-interpolated value r is the same number for both bins (same cells, same x weights), so it is
        // kept in registers: ~1/3 fewer loads and fmas, and most bins need ONE round of loads (two new rows) instead of two.
        float4 rl = make_float4(0.f, 0.f, 0.f, 0.f);
        int offl = -1;
        for (int ph = 0; ph < PH; ++ph) {
            const AxisEntry* yt = ytab + ph * 4;
            const int ny = ycnt[ph];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int a = 0;
            if (ny > 0 && yt[0].off == offl) {
                const float wy = yt[0].w;
                acc.x = fmaf(wy, rl.x, acc.x); acc.y = fmaf(wy, rl.y, acc.y); acc.z = fmaf(wy, rl.z, acc.z); acc.w = fmaf(wy, rl.w, acc.w);
                a = 1;
            }
            for (; a + 2 <= ny; a += 2) roi_rows_quad<2, NX>(f, q, yt + a, xo, wx, acc, rl);
            if (a < ny) roi_rows_quad<1, NX>(f, q, yt + a, xo, wx, acc, rl);
            if (ny > 0) offl = yt[ny - 1].off;
            if (pow2) { acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv; }
            else { acc.x = __fdiv_rn(acc.x, count); acc.y = __fdiv_rn(acc.y, count); acc.z = __fdiv_rn(acc.z, count); acc.w = __fdiv_rn(acc.w, count); }
            // lane-rotated store order (rot = lane/8): at every step the 32 lanes hit 32 different banks
            float* t = tq + ph * PW;
            const float a0 = (rot & 1) ? acc.y : acc.x, a1 = (rot & 1) ? acc.z : acc.y, a2 = (rot & 1) ? acc.w : acc.z, a3 = (rot & 1) ? acc.x : acc.w;
            const float b0 = (rot & 2) ? a2 : a0, b1 = (rot & 2) ? a3 : a1, b2 = (rot & 2) ? a0 : a2, b3 = (rot & 2) ? a1 : a3;  // b_s = acc[(s+rot)&3]
            t[((0 + rot) & 3) * nb] = b0; t[((1 + rot) & 3) * nb] = b1; t[((2 + rot) & 3) * nb] = b2; t[((3 + rot) & 3) * nb] = b3;
        }
    }
}

// blockDim.x = 32 * PW (3 <= PW <= 7: 224 threads, 71 registers, 4 CTAs per SM beside their 50 KB tiles), sampling_ratio 1..2, C % 4 == 0
__global__ void __launch_bounds__(224, 4) roi_align_nhwc_col_kernel(const __grid_constant__ RoiParams p, int use_tma, int tab, int dbg) {
    extern __shared__ __align__(128) float smem_f[];
    float* tile = smem_f;                                  // [C][PH*PW]
    AxisEntry* ytab = (AxisEntry*)(tile + (size_t)p.C * p.PH * p.PW);
    AxisEntry* xtab = ytab + tab;
    __shared__ int ycnt[64], xcnt[64];
    const long long nk = p.list ? (long long)*p.list_count : p.K;
    for (long long kk = blockIdx.x; kk < nk; kk += gridDim.x) {
        const long long k = p.list ? (long long)p.list[kk] : kk;
        const float* roi = p.rois + k * 5;
        const int lvl = p.level_ids ? p.level_ids[k] : 0;
        const int H = p.H[lvl], W = p.W[lvl];
        const float sc = p.scale[lvl];
        const int bidx = (int)roi[0];
        const float off = p.aligned ? 0.5f : 0.0f;
        const float sw = __fsub_rn(__fmul_rn(roi[1], sc), off), sh = __fsub_rn(__fmul_rn(roi[2], sc), off);
        const float ew = __fsub_rn(__fmul_rn(roi[3], sc), off), eh = __fsub_rn(__fmul_rn(roi[4], sc), off);
        float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
        if (!p.aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
        const float bh = __fdiv_rn(rh, (float)p.PH), bw = __fdiv_rn(rw, (float)p.PW);
        const int gh = p.sampling_ratio, gw = p.sampling_ratio;
        const float count = (float)max(gh * gw, 1);
        const float* __restrict__ f = p.data[lvl] + (size_t)bidx * H * W * p.C;
        if (threadIdx.x < p.PH) build_axis(ytab, ycnt, threadIdx.x, 4, sh, bh, gh, H, W * p.C, 4);
        else if (threadIdx.x >= 64 && threadIdx.x < 64 + p.PW) build_axis(xtab, xcnt, threadIdx.x - 64, 4, sw, bw, gw, W, p.C, 4);
        __syncthreads();
        const int pw = threadIdx.x >> 5, lane = threadIdx.x & 31, nq = p.C >> 2, nx = (dbg & 2) ? 0 : xcnt[pw];   // (dbg 2: no loads)
        const AxisEntry* xt = xtab + pw * 4;
        if (nx == 2) roi_align_col_quad<2>(f, lane, nq, pw, ytab, xt, ycnt, p.PH, p.PW, count, tile);
        else if (nx == 3) roi_align_col_quad<3>(f, lane, nq, pw, ytab, xt, ycnt, p.PH, p.PW, count, tile);
        else if (nx == 4) roi_align_col_quad<4>(f, lane, nq, pw, ytab, xt, ycnt, p.PH, p.PW, count, tile);
        else if (nx == 1) roi_align_col_quad<1>(f, lane, nq, pw, ytab, xt, ycnt, p.PH, p.PW, count, tile);
        else {   // every sample of this column lies outside the map: zeros
            const int nb = p.PH * p.PW;
            for (int q = lane; q < nq; q += 32)
                for (int ph = 0; ph < p.PH; ++ph)
#pragma unroll
                    for (int c = 0; c < 4; ++c) tile[(size_t)(4 * q + c) * nb + ph * p.PW + pw] = 0.0f;
        }
        if (!(dbg & 1)) tile_store(p.out + (size_t)k * p.C * (p.PH * p.PW), tile, p.C * (p.PH * p.PW), use_tma != 0 && !(dbg & 4));   // (dbg 1: no output, 4: thread stores)
        __syncthreads();   // the tile and the tables are rewritten by the next RoI of the walk
    }
}

// ------------------------------------------------------------------------------------------------ RoIAlign NHWC, channel slices
// Locality variant of the channel-quad kernel for sampling_ratio <= 2: gridDim.y channel slices per RoI (a 64-channel slice keeps a
// 12.5 KB tile, so four resident CTAs leave most of the SM's 256 KB to L1) and every CTA walks a CONTIGUOUS chunk of a RoI list
// that the caller sorted by position -- consecutive RoIs of an image overlap (2 000 proposals cover the stride-4 map ~7x), so the
// cells of RoI i+1 are mostly L1 hits instead of L2 round trips.  Same tables and arithmetic as the kernel above.
__global__ void __launch_bounds__(256, 4) roi_align_sliced_kernel(const __grid_constant__ RoiParams p, int QT, int tab, int chunk) {
    extern __shared__ __align__(128) float smem_f[];
    const int slices = gridDim.y, sl = blockIdx.y;
    const int Cs = p.C / slices, nb = p.PH * p.PW;
    float* tile = smem_f;                                  // [Cs][PH*PW]
    AxisEntry* ytab = (AxisEntry*)(tile + (size_t)Cs * nb);
    AxisEntry* xtab = ytab + tab;
    __shared__ int ycnt[64], xcnt[64];
    const long long nk = p.list ? (long long)*p.list_count : p.K;
    const long long k0 = (long long)blockIdx.x * chunk, k1 = (k0 + chunk < nk) ? k0 + chunk : nk;
    for (long long kk = k0; kk < k1; ++kk) {
        const long long k = p.list ? (long long)p.list[kk] : kk;
        const float* roi = p.rois + k * 5;
        const int lvl = p.level_ids ? p.level_ids[k] : 0;
        const int H = p.H[lvl], W = p.W[lvl];
        const float sc = p.scale[lvl];
        const int bidx = (int)roi[0];
        const float off = p.aligned ? 0.5f : 0.0f;
        const float sw = __fsub_rn(__fmul_rn(roi[1], sc), off), sh = __fsub_rn(__fmul_rn(roi[2], sc), off);
        const float ew = __fsub_rn(__fmul_rn(roi[3], sc), off), eh = __fsub_rn(__fmul_rn(roi[4], sc), off);
        float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
        if (!p.aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
        const float bh = __fdiv_rn(rh, (float)p.PH), bw = __fdiv_rn(rw, (float)p.PW);
        const int gh = p.sampling_ratio, gw = p.sampling_ratio;
        const float count = (float)max(gh * gw, 1);
        const float* __restrict__ f = p.data[lvl] + (size_t)bidx * H * W * p.C;
        if (threadIdx.x < p.PH) build_axis(ytab, ycnt, threadIdx.x, 4, sh, bh, gh, H, W * p.C, 4);
        else if (threadIdx.x >= 64 && threadIdx.x < 64 + p.PW) build_axis(xtab, xcnt, threadIdx.x - 64, 4, sw, bw, gw, W, p.C, 4);
        __syncthreads();
        const int nqs = Cs >> 2, q0 = sl * nqs, groups = 256 / QT, g = threadIdx.x / QT, rot = (threadIdx.x & 31) >> 3;
        for (int q = q0 + threadIdx.x % QT; q < q0 + nqs; q += QT)
            roi_align_bins_quad<4>(f, q, g, groups, ytab, xtab, ycnt, xcnt, p.PH, p.PW, count, tile - (size_t)(4 * q0) * nb, rot);
        tile_store(p.out + ((size_t)k * p.C + (size_t)sl * Cs) * nb, tile, Cs * nb, true);
        __syncthreads();   // the tile and the tables are rewritten by the next RoI of the walk
    }
}

// ------------------------------------------------------------------------------------------------ RoIAlign NHWC, staged rows
// Persistent, warp-specialised variant for sampling_ratio <= 2 (the detection-head configuration): one CTA per SM walks
// its RoIs; a PRODUCER warp derives each RoI's geometry (the separable tables, the distinct feature rows they touch and
// the x range) and streams exactly those row segments -- span * C contiguous floats in NHWC -- into a shared-memory
// ring with TMA bulk copies (cp.async.bulk + mbarrier complete_tx); 8 CONSUMER warps (warp = output column pw, lane =
// channel quads) compute the bins from shared memory and assemble the [C,PH,PW] tile, which leaves with one TMA bulk
// store.  The gather kernel above re-reads every cell ~2.5x through L1/L2 (overlapping bilinear footprints of
// neighbouring samples and bins); here every needed cell crosses L2->SM once, and the loads of RoI i+1 are in flight
// while RoI i is computed.  RoIs whose x range exceeds RR_SPAN cells (sparse sampling: no reuse to win) are computed by
// the same consumers straight from global memory.  Same arithmetic as the gather kernels (identical tables and order).
// STATUS: opt-in (hd_roi_set_mode(2)).  Measured on B200 (cfg3, 32 000 RoIs): 3.0 ms vs 1.2 ms for the gather kernel; streaming the
// rows alone (no compute, no store) already takes 1.0 ms.  With a 50 KB output tile only ~10 row slots (130 KB) fit beside it,
// of which a bin-row pins 3-4, so too few bytes are in flight per SM to cover the L2 latency, and 8 consumer warps cannot hide
// the shared-memory latency of the bin arithmetic.  Kept because it is bit-identical, tested, and the starting point for a
// tensor-map variant with channel-sliced tiles (several small CTAs per SM).
#define RR_NCW 8                      // consumer warps
#define RR_THREADS ((RR_NCW + 1) * 32)
#define RR_MAXSLOTS 16
#define RR_SPAN 16                    // widest staged x range (cells)
#define RR_MAXP 8                     // PH, PW <= 8
#define RR_NG 4                       // geometry records in flight

struct RrGeom {
    AxisEntry ytab[RR_MAXP * 4], xtab[RR_MAXP * 4];   // off = raw cell index
    int ycnt[RR_MAXP], xcnt[RR_MAXP];
    int yrow[RR_MAXP * 4];        // row sequence index of every y entry
    int rows[RR_MAXP * 4];        // distinct rows in load order
    int first_row[RR_MAXP + 1];   // lowest row index still needed from bin-row ph on
    int last_row[RR_MAXP];        // highest row index bin-row ph needs (-1: none)
    int nrows, xmin, span, staged, lvl, bidx;
    float count;
};

__device__ __forceinline__ void rr_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rr_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void rr_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rr_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void rr_bulk_load(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// one output bin of one channel quad: NY x NX separable taps (rows already resolved to pointers)
template <int NX, bool STAGED>
__device__ __forceinline__ float4 rr_bin(const float4* const* rowp, const float* wy, int ny, const int* xo, const float* wx, int q) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        if (a < ny) {
            const float4* row = rowp[a] + q;
            float4 v[NX];
#pragma unroll
            for (int b = 0; b < NX; ++b) v[b] = STAGED ? row[xo[b]] : __ldg(row + xo[b]);
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int b = 0; b < NX; ++b) { r.x = fmaf(wx[b], v[b].x, r.x); r.y = fmaf(wx[b], v[b].y, r.y); r.z = fmaf(wx[b], v[b].z, r.z); r.w = fmaf(wx[b], v[b].w, r.w); }
            acc.x = fmaf(wy[a], r.x, acc.x); acc.y = fmaf(wy[a], r.y, acc.y); acc.z = fmaf(wy[a], r.z, acc.z); acc.w = fmaf(wy[a], r.w, acc.w);
        }
    }
    return acc;
}

struct RrSmem {
    RrGeom geom[RR_NG];
    unsigned long long full[RR_MAXSLOTS], empty[RR_MAXSLOTS], gfull[RR_NG], gempty[RR_NG];
};

__global__ void __launch_bounds__(RR_THREADS, 1) roi_align_ring_kernel(const __grid_constant__ RoiParams p, int n_slots, int slot_bytes, int ring_off, int meta_off, int dbg) {
    extern __shared__ __align__(128) unsigned char smem_b[];
    float* tile = reinterpret_cast<float*>(smem_b);                       // [C][PH*PW]
    unsigned char* ring = smem_b + ring_off;
    RrSmem& sm = *reinterpret_cast<RrSmem*>(smem_b + meta_off);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nb = p.PH * p.PW, nq = p.C >> 2;
    if (tid == 0) {
        for (int s = 0; s < n_slots; ++s) { rr_mbar_init(&sm.full[s], 1); rr_mbar_init(&sm.empty[s], RR_NCW); }
        for (int g = 0; g < RR_NG; ++g) { rr_mbar_init(&sm.gfull[g], 1); rr_mbar_init(&sm.gempty[g], RR_NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long k0 = blockIdx.x, kstep = gridDim.x;
    // ring position of the next staged row: slot index and use count of that slot; producer and consumers advance identically
    int rslot = 0; unsigned ruse = 0;
    auto advance = [&](int rows) { rslot += rows; while (rslot >= n_slots) { rslot -= n_slots; ++ruse; } };
    if (wid == RR_NCW) {
        // ================================================================ producer warp
        int it = 0;
        for (long long k = k0; k < p.K; k += kstep, ++it) {
            RrGeom& G = sm.geom[it % RR_NG];
            rr_mbar_wait(&sm.gempty[it % RR_NG], ((it / RR_NG) & 1) ^ 1);
            const float* roi = p.rois + k * 5;
            const int lvl = p.level_ids ? p.level_ids[k] : 0;
            const int H = p.H[lvl], W = p.W[lvl];
            const float sc = p.scale[lvl];
            const float off = p.aligned ? 0.5f : 0.0f;
            const float sw = __fsub_rn(__fmul_rn(roi[1], sc), off), sh = __fsub_rn(__fmul_rn(roi[2], sc), off);
            const float ew = __fsub_rn(__fmul_rn(roi[3], sc), off), eh = __fsub_rn(__fmul_rn(roi[4], sc), off);
            float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
            if (!p.aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
            const float bh = __fdiv_rn(rh, (float)p.PH), bw = __fdiv_rn(rw, (float)p.PW);
            const int g = p.sampling_ratio;   // 1 or 2 (host check)
            if (lane < p.PH) build_axis(G.ytab, G.ycnt, lane, 4, sh, bh, g, H, 1, 4);
            else if (lane >= 16 && lane < 16 + p.PW) build_axis(G.xtab, G.xcnt, lane - 16, 4, sw, bw, g, W, 1, 4);
            __syncwarp();
            // lane e = (ph, a) owns one y entry.  The distinct rows, ascending, are the load order (sample positions grow with
            // ph, so a row is new only if it is above every earlier one): row index = number of distinct smaller rows.
            const int eph = lane >> 2, ea = lane & 3;
            const bool yvalid = eph < p.PH && ea < G.ycnt[eph];
            const int ycell = yvalid ? G.ytab[lane].off : 0x7fffffff - lane;     // distinct sentinels
            const unsigned peers = __match_any_sync(HD_FULL, ycell);
            const bool leader = yvalid && (peers & hd_lanemask_lt()) == 0u;
            int idx = 0;
#pragma unroll 8
            for (int l = 0; l < 32; ++l) {
                const int yl = __shfl_sync(HD_FULL, ycell, l);
                const int ll = __shfl_sync(HD_FULL, (int)leader, l);
                idx += (ll && yl < ycell) ? 1 : 0;
            }
            const int nrows = __popc(__ballot_sync(HD_FULL, leader));
            if (yvalid) G.yrow[lane] = idx;
            if (leader) G.rows[idx] = ycell;
            // x range
            const bool xvalid = eph < p.PW && ea < G.xcnt[eph];
            int xmin = xvalid ? G.xtab[lane].off : 0x7fffffff, xmax = xvalid ? G.xtab[lane].off : -1;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) { xmin = min(xmin, __shfl_xor_sync(HD_FULL, xmin, d)); xmax = max(xmax, __shfl_xor_sync(HD_FULL, xmax, d)); }
            const int span = (xmax >= 0) ? xmax - xmin + 1 : 0;
            if (xmax < 0) xmin = 0;
            const int staged = (nrows > 0 && span >= 1 && span <= RR_SPAN && bh >= 0.0f && bw >= 0.0f && !(dbg & 8)) ? 1 : 0;   // (inverted RoIs walk the rows downwards)
            __syncwarp();
            if (lane < p.PH) {   // per bin-row: highest row needed, and lowest row any bin-row >= ph still needs
                const int c = G.ycnt[lane];
                G.last_row[lane] = c > 0 ? G.yrow[lane * 4 + c - 1] : -1;
                int fr = nrows;
                for (int ph = p.PH - 1; ph >= lane; --ph) if (G.ycnt[ph] > 0) fr = G.yrow[ph * 4];
                G.first_row[lane] = fr;
            }
            if (lane == 0) {
                G.first_row[p.PH] = nrows;
                G.nrows = nrows; G.xmin = xmin; G.span = span; G.staged = staged; G.lvl = lvl; G.bidx = (int)roi[0];
                G.count = (float)max(g * g, 1);
            }
            __syncwarp();
            if (lane == 0) rr_mbar_arrive(&sm.gfull[it % RR_NG]);
            if (staged) {
                const float* f = p.data[lvl] + (size_t)(int)roi[0] * H * W * p.C;
                const unsigned bytes = (unsigned)span * (unsigned)p.C * 4u;
                // lane r issues row r, one wave of n_slots rows at a time: inside a wave every lane waits on a different
                // slot, so the 1-bit phase parity of the empty barriers cannot alias (a waiter is never two phases ahead)
                for (int base = 0; base < nrows; base += n_slots) {
                    if (lane >= base && lane < min(nrows, base + n_slots)) {
                        int slot = rslot + lane; unsigned use = ruse;
                        while (slot >= n_slots) { slot -= n_slots; ++use; }
                        rr_mbar_wait(&sm.empty[slot], (use & 1u) ^ 1u);
                        rr_mbar_expect_tx(&sm.full[slot], bytes);
                        rr_bulk_load(ring + (size_t)slot * slot_bytes, f + ((size_t)G.rows[lane] * W + xmin) * p.C, bytes, &sm.full[slot]);
                    }
                    __syncwarp();
                }
                advance(nrows);
            }
        }
    } else {
        // ================================================================ consumer warps
        const int rot = lane >> 3;
        int it = 0;
        for (long long k = k0; k < p.K; k += kstep, ++it) {
            const RrGeom& G = sm.geom[it % RR_NG];
            rr_mbar_wait(&sm.gfull[it % RR_NG], (it / RR_NG) & 1);
            const int lvl = G.lvl, W = p.W[lvl], H = p.H[lvl];
            const float* __restrict__ f = p.data[lvl] + (size_t)G.bidx * H * W * p.C;
            const float count = G.count;
            const int icount = (int)count;
            const bool pow2 = (icount & (icount - 1)) == 0;
            const float inv = 1.0f / count;
            const bool staged = G.staged != 0;
            const int xmin = G.xmin, nrows = G.nrows;
            // the previous tile must have left shared memory before it is overwritten
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(RR_NCW * 32) : "memory");
            int waited = 0, released = 0;
            for (int ph = 0; ph < p.PH; ++ph) {
                const int ny = G.ycnt[ph];
                const float4* rowp[4]; float wy[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    wy[a] = G.ytab[ph * 4 + a].w;
                    if (staged) {
                        int slot = rslot + ((a < ny) ? G.yrow[ph * 4 + a] : 0);
                        while (slot >= n_slots) slot -= n_slots;
                        rowp[a] = reinterpret_cast<const float4*>(ring + (size_t)slot * slot_bytes);
                    } else {
                        rowp[a] = reinterpret_cast<const float4*>(f + (size_t)G.ytab[ph * 4 + a].off * W * p.C);
                    }
                }
                if (staged) {
                    const int need = G.last_row[ph];
                    for (; waited <= need; ++waited) {
                        int slot = rslot + waited; unsigned use = ruse;
                        while (slot >= n_slots) { slot -= n_slots; ++use; }
                        rr_mbar_wait(&sm.full[slot], use & 1u);
                    }
                }
                for (int pw = wid; pw < p.PW && !(dbg & 2); pw += RR_NCW) {
                    const int nx = G.xcnt[pw];
                    int xo[4]; float wx[4];
#pragma unroll
                    for (int bq = 0; bq < 4; ++bq) {
                        const int xc = G.xtab[pw * 4 + bq].off;
                        xo[bq] = (staged ? (xc - xmin) : xc) * nq;        // float4 units
                        wx[bq] = G.xtab[pw * 4 + bq].w;
                    }
                    const int bin = ph * p.PW + pw;
                    for (int q = lane; q < nq; q += 32) {
                        float4 acc;
                        if (staged) {
                            if (nx == 2) acc = rr_bin<2, true>(rowp, wy, ny, xo, wx, q);
                            else if (nx == 3) acc = rr_bin<3, true>(rowp, wy, ny, xo, wx, q);
                            else if (nx == 4) acc = rr_bin<4, true>(rowp, wy, ny, xo, wx, q);
                            else if (nx == 1) acc = rr_bin<1, true>(rowp, wy, ny, xo, wx, q);
                            else acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        } else {
                            if (nx == 2) acc = rr_bin<2, false>(rowp, wy, ny, xo, wx, q);
                            else if (nx == 3) acc = rr_bin<3, false>(rowp, wy, ny, xo, wx, q);
                            else if (nx == 4) acc = rr_bin<4, false>(rowp, wy, ny, xo, wx, q);
                            else if (nx == 1) acc = rr_bin<1, false>(rowp, wy, ny, xo, wx, q);
                            else acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        if (pow2) { acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv; }
                        else { acc.x = __fdiv_rn(acc.x, count); acc.y = __fdiv_rn(acc.y, count); acc.z = __fdiv_rn(acc.z, count); acc.w = __fdiv_rn(acc.w, count); }
                        // lane-rotated store order (rot = lane/8): at every step the 32 lanes hit 32 different banks
                        float* t = tile + (size_t)(4 * q) * nb + bin;
                        const float a0 = (rot & 1) ? acc.y : acc.x, a1 = (rot & 1) ? acc.z : acc.y, a2 = (rot & 1) ? acc.w : acc.z, a3 = (rot & 1) ? acc.x : acc.w;
                        const float b0 = (rot & 2) ? a2 : a0, b1 = (rot & 2) ? a3 : a1, b2 = (rot & 2) ? a0 : a2, b3 = (rot & 2) ? a1 : a3;  // b_s = acc[(s+rot)&3]
                        t[((0 + rot) & 3) * nb] = b0; t[((1 + rot) & 3) * nb] = b1; t[((2 + rot) & 3) * nb] = b2; t[((3 + rot) & 3) * nb] = b3;
                    }
                }
                if (staged) {   // rows no later bin-row needs go back to the producer
                    __syncwarp();
                    const int upto = G.first_row[ph + 1];
                    if (lane == 0)
                        for (int r = released; r < upto; ++r) {
                            int slot = rslot + r;
                            while (slot >= n_slots) slot -= n_slots;
                            rr_mbar_arrive(&sm.empty[slot]);
                        }
                    released = max(released, upto);
                }
            }
            if (staged) advance(nrows);
            __syncwarp();
            if (lane == 0) rr_mbar_arrive(&sm.gempty[it % RR_NG]);
            // tile complete -> one bulk store
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(RR_NCW * 32) : "memory");
            if (tid == 0 && !(dbg & 1)) {
                unsigned saddr = (unsigned)__cvta_generic_to_shared(tile);
                float* gdst = p.out + (size_t)k * p.C * nb;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(p.C * nb * 4) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// ------------------------------------------------------------------------------------------------ RoIPool NHWC
__global__ void __launch_bounds__(256) roi_pool_nhwc_kernel(const __grid_constant__ RoiParams p, int use_tma) {
    extern __shared__ __align__(128) float smem_f[];
    float* tile = smem_f;
    const long long k = blockIdx.x;
    const float* roi = p.rois + k * 5;
    const int lvl = p.level_ids ? p.level_ids[k] : 0;
    const int H = p.H[lvl], W = p.W[lvl];
    const float sc = p.scale[lvl];
    const int bidx = (int)roi[0];
    const int x1 = (int)roundf(__fmul_rn(roi[1], sc)), y1 = (int)roundf(__fmul_rn(roi[2], sc));
    const int x2 = (int)roundf(__fmul_rn(roi[3], sc)), y2 = (int)roundf(__fmul_rn(roi[4], sc));
    const int rw = max(x2 - x1 + 1, 1), rh = max(y2 - y1 + 1, 1);
    const float bh = __fdiv_rn((float)rh, (float)p.PH), bw = __fdiv_rn((float)rw, (float)p.PW);
    const float* __restrict__ f = p.data[lvl] + (size_t)bidx * H * W * p.C;
    const int nb = p.PH * p.PW;
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
        for (int ph = 0; ph < p.PH; ++ph) {
            const int hs = min(max((int)floorf(__fmul_rn((float)ph, bh)) + y1, 0), H);
            const int he = min(max((int)ceilf(__fmul_rn((float)(ph + 1), bh)) + y1, 0), H);
            for (int pw = 0; pw < p.PW; ++pw) {
                const int ws = min(max((int)floorf(__fmul_rn((float)pw, bw)) + x1, 0), W);
                const int we = min(max((int)ceilf(__fmul_rn((float)(pw + 1), bw)) + x1, 0), W);
                const bool empty = (he <= hs) || (we <= ws);
                float mx = empty ? 0.0f : -INFINITY;
                int mi = -1;
                for (int h = hs; h < he; ++h) {
                    const float* __restrict__ row = f + (size_t)h * W * p.C + c;
                    int w = ws;
                    for (; w + 4 <= we; w += 4) {  // four independent loads in flight
                        const float v0 = __ldg(row + (size_t)w * p.C), v1 = __ldg(row + (size_t)(w + 1) * p.C);
                        const float v2 = __ldg(row + (size_t)(w + 2) * p.C), v3 = __ldg(row + (size_t)(w + 3) * p.C);
                        if (v0 > mx) { mx = v0; mi = h * W + w; }
                        if (v1 > mx) { mx = v1; mi = h * W + w + 1; }
                        if (v2 > mx) { mx = v2; mi = h * W + w + 2; }
                        if (v3 > mx) { mx = v3; mi = h * W + w + 3; }
                    }
                    for (; w < we; ++w) {
                        const float v = __ldg(row + (size_t)w * p.C);
                        if (v > mx) { mx = v; mi = h * W + w; }
                    }
                }
                tile[c * nb + ph * p.PW + pw] = mx;
                if (p.argmax) p.argmax[((size_t)k * p.C + c) * nb + ph * p.PW + pw] = mi;
            }
        }
    }
    tile_store(p.out + (size_t)k * p.C * nb, tile, p.C * nb, use_tma != 0);
}

// ------------------------------------------------------------------------------------------------ RoIPool NHWC, float4
__global__ void __launch_bounds__(256) roi_pool_nhwc_quad_kernel(const __grid_constant__ RoiParams p, int use_tma, int QT) {
    extern __shared__ __align__(128) float smem_f[];
    float* tile = smem_f;
    const long long k = blockIdx.x;
    const float* roi = p.rois + k * 5;
    const int lvl = p.level_ids ? p.level_ids[k] : 0;
    const int H = p.H[lvl], W = p.W[lvl];
    const float sc = p.scale[lvl];
    const int bidx = (int)roi[0];
    const int x1 = (int)roundf(__fmul_rn(roi[1], sc)), y1 = (int)roundf(__fmul_rn(roi[2], sc));
    const int x2 = (int)roundf(__fmul_rn(roi[3], sc)), y2 = (int)roundf(__fmul_rn(roi[4], sc));
    const int rw = max(x2 - x1 + 1, 1), rh = max(y2 - y1 + 1, 1);
    const float bh = __fdiv_rn((float)rh, (float)p.PH), bw = __fdiv_rn((float)rw, (float)p.PW);
    const float* __restrict__ f = p.data[lvl] + (size_t)bidx * H * W * p.C;
    const int nb = p.PH * p.PW, nq = p.C >> 2, groups = 256 / QT, g = threadIdx.x / QT, rot = (threadIdx.x & 31) >> 3;
    for (int q = threadIdx.x % QT; q < nq; q += QT) {
        const float4* __restrict__ fq = reinterpret_cast<const float4*>(f) + q;
        for (int bin = g; bin < nb; bin += groups) {
            const int ph = bin / p.PW, pw = bin - ph * p.PW;
            const int hs = min(max((int)floorf(__fmul_rn((float)ph, bh)) + y1, 0), H);
            const int he = min(max((int)ceilf(__fmul_rn((float)(ph + 1), bh)) + y1, 0), H);
            const int ws = min(max((int)floorf(__fmul_rn((float)pw, bw)) + x1, 0), W);
            const int we = min(max((int)ceilf(__fmul_rn((float)(pw + 1), bw)) + x1, 0), W);
            const bool empty = (he <= hs) || (we <= ws);
            const float init = empty ? 0.0f : -INFINITY;
            float4 mx = make_float4(init, init, init, init);
            int4 mi = make_int4(-1, -1, -1, -1);
            auto upd = [&](const float4& v, int idx) {
                if (v.x > mx.x) { mx.x = v.x; mi.x = idx; }
                if (v.y > mx.y) { mx.y = v.y; mi.y = idx; }
                if (v.z > mx.z) { mx.z = v.z; mi.z = idx; }
                if (v.w > mx.w) { mx.w = v.w; mi.w = idx; }
            };
            for (int h = hs; h < he; ++h) {
                const float4* __restrict__ row = fq + (size_t)h * W * nq;
                int w = ws;
                for (; w + 4 <= we; w += 4) {
                    const float4 v0 = __ldg(row + (size_t)w * nq), v1 = __ldg(row + (size_t)(w + 1) * nq);
                    const float4 v2 = __ldg(row + (size_t)(w + 2) * nq), v3 = __ldg(row + (size_t)(w + 3) * nq);
                    upd(v0, h * W + w); upd(v1, h * W + w + 1); upd(v2, h * W + w + 2); upd(v3, h * W + w + 3);
                }
                for (; w < we; ++w) upd(__ldg(row + (size_t)w * nq), h * W + w);
            }
            float* t = tile + (size_t)(4 * q) * nb + bin;
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {
                const int k2 = (s2 + rot) & 3;
                t[k2 * nb] = (k2 == 0) ? mx.x : (k2 == 1) ? mx.y : (k2 == 2) ? mx.z : mx.w;
            }
            if (p.argmax) {
                int* am = p.argmax + ((size_t)k * p.C + 4 * q) * nb + bin;
                am[0] = mi.x; am[nb] = mi.y; am[2 * nb] = mi.z; am[3 * nb] = mi.w;
            }
        }
    }
    tile_store(p.out + (size_t)k * p.C * nb, tile, p.C * nb, use_tma != 0);
}

// ------------------------------------------------------------------------------------------------ direct NCHW
// warp per (roi, channel), lane per bin: reference-order sample-by-sample arithmetic.  Used for few RoIs
// (where a layout pass over the whole feature map would dominate) and as the in-library cross-check.
template <bool POOL>
__global__ void __launch_bounds__(256) roi_nchw_kernel(const __grid_constant__ RoiParams p) {
    const int lane = threadIdx.x & 31;
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wg >= p.K * p.C) return;
    const long long k = wg / p.C;
    const int c = (int)(wg - k * p.C);
    const float* roi = p.rois + k * 5;
    const int lvl = p.level_ids ? p.level_ids[k] : 0;
    const int H = p.H[lvl], W = p.W[lvl];
    const float sc = p.scale[lvl];
    const int bidx = (int)roi[0];
    const float* __restrict__ f = p.data[lvl] + ((size_t)bidx * p.C + c) * H * W;
    const int nb = p.PH * p.PW;
    float* o = p.out + ((size_t)k * p.C + c) * nb;
    if (POOL) {
        const int x1 = (int)roundf(__fmul_rn(roi[1], sc)), y1 = (int)roundf(__fmul_rn(roi[2], sc));
        const int x2 = (int)roundf(__fmul_rn(roi[3], sc)), y2 = (int)roundf(__fmul_rn(roi[4], sc));
        const int rw = max(x2 - x1 + 1, 1), rh = max(y2 - y1 + 1, 1);
        const float bh = __fdiv_rn((float)rh, (float)p.PH), bw = __fdiv_rn((float)rw, (float)p.PW);
        for (int bin = lane; bin < nb; bin += 32) {
            const int ph = bin / p.PW, pw = bin - ph * p.PW;
            const int hs = min(max((int)floorf(__fmul_rn((float)ph, bh)) + y1, 0), H);
            const int he = min(max((int)ceilf(__fmul_rn((float)(ph + 1), bh)) + y1, 0), H);
            const int ws = min(max((int)floorf(__fmul_rn((float)pw, bw)) + x1, 0), W);
            const int we = min(max((int)ceilf(__fmul_rn((float)(pw + 1), bw)) + x1, 0), W);
            const bool empty = (he <= hs) || (we <= ws);
            float mx = empty ? 0.0f : -INFINITY;
            int mi = -1;
            for (int h = hs; h < he; ++h)
                for (int w = ws; w < we; ++w) {
                    float v = __ldg(f + (size_t)h * W + w);
                    if (v > mx) { mx = v; mi = h * W + w; }
                }
            o[bin] = mx;
            if (p.argmax) p.argmax[((size_t)k * p.C + c) * nb + bin] = mi;
        }
    } else {
        const float off = p.aligned ? 0.5f : 0.0f;
        const float sw = __fsub_rn(__fmul_rn(roi[1], sc), off), sh = __fsub_rn(__fmul_rn(roi[2], sc), off);
        const float ew = __fsub_rn(__fmul_rn(roi[3], sc), off), eh = __fsub_rn(__fmul_rn(roi[4], sc), off);
        float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
        if (!p.aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
        const float bh = __fdiv_rn(rh, (float)p.PH), bw = __fdiv_rn(rw, (float)p.PW);
        const int gh = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)p.PH));
        const int gw = p.sampling_ratio > 0 ? p.sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)p.PW));
        const float count = (float)max(gh * gw, 1);
        for (int bin = lane; bin < nb; bin += 32) {
            const int ph = bin / p.PW, pw = bin - ph * p.PW;
            float acc = 0.0f;
            for (int iy = 0; iy < gh; ++iy) {
                float y = __fadd_rn(__fadd_rn(sh, __fmul_rn((float)ph, bh)), __fdiv_rn(__fmul_rn(__fadd_rn((float)iy, 0.5f), bh), (float)gh));
                if (y < -1.0f || y > (float)H) continue;
                if (y <= 0.0f) y = 0.0f;
                int yl = (int)y, yh;
                if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
                const float ly = __fsub_rn(y, (float)yl), hy = __fsub_rn(1.0f, ly);
                for (int ix = 0; ix < gw; ++ix) {
                    float x = __fadd_rn(__fadd_rn(sw, __fmul_rn((float)pw, bw)), __fdiv_rn(__fmul_rn(__fadd_rn((float)ix, 0.5f), bw), (float)gw));
                    if (x < -1.0f || x > (float)W) continue;
                    if (x <= 0.0f) x = 0.0f;
                    int xl = (int)x, xh;
                    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
                    const float lx = __fsub_rn(x, (float)xl), hx = __fsub_rn(1.0f, lx);
                    // w1*v1 + w2*v2 + w3*v3 + w4*v4, left to right, then added to the running sum
                    float s = __fmul_rn(__fmul_rn(hy, hx), __ldg(f + (size_t)yl * W + xl));
                    s = __fadd_rn(s, __fmul_rn(__fmul_rn(hy, lx), __ldg(f + (size_t)yl * W + xh)));
                    s = __fadd_rn(s, __fmul_rn(__fmul_rn(ly, hx), __ldg(f + (size_t)yh * W + xl)));
                    s = __fadd_rn(s, __fmul_rn(__fmul_rn(ly, lx), __ldg(f + (size_t)yh * W + xh)));
                    acc = __fadd_rn(acc, s);
                }
            }
            o[bin] = __fdiv_rn(acc, count);
        }
    }
}

// ------------------------------------------------------------------------------------------------ layout pass
// [B, C, HW] -> [B, HW, C] through a 32x33 shared tile (both sides coalesced).
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW) {
    __shared__ float t[32][33];
    const int b = blockIdx.z;
    const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const float* src = in + (size_t)b * C * HW;
    float* dst = out + (size_t)b * C * HW;
#pragma unroll
    for (int r = ty; r < 32; r += 8)
        if (c0 + r < C && hw0 + tx < HW) t[r][tx] = hd_ldg_stream(src + (size_t)(c0 + r) * HW + hw0 + tx);
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8)
        if (hw0 + r < HW && c0 + tx < C) dst[(size_t)(hw0 + r) * C + c0 + tx] = t[tx][r];
}

// ------------------------------------------------------------------------------------------------ level map
__global__ void roi_level_map_kernel(const float* __restrict__ rois, int stride, int box_off, long long K, int style, int k_min,
                                     int k_max, float s0, float lvl0, float eps, int* __restrict__ out32, long long* __restrict__ out64) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const float* r = rois + i * stride + box_off;
    const float area = __fmul_rn(__fsub_rn(r[2], r[0]), __fsub_rn(r[3], r[1]));
    const float s = __fsqrt_rn(area);
    int lv;
    if (style == 1) {  // mmdet: floor(log2(s / finest_scale + eps)), clamp [0, L-1]
        float v = floorf(log2f(__fadd_rn(__fdiv_rn(s, s0), eps)));
        v = fminf(fmaxf(v, 0.0f), (float)(k_max - k_min));
        lv = (v != v) ? 0 : (int)v;
    } else {           // torchvision: floor(lvl0 + log2(s / s0) + eps), clamp [k_min, k_max], - k_min
        float v = floorf(__fadd_rn(__fadd_rn(lvl0, log2f(__fdiv_rn(s, s0))), eps));
        v = fminf(fmaxf(v, (float)k_min), (float)k_max);
        lv = ((v != v) ? k_min : (int)v) - k_min;
    }
    if (out32) out32[i] = lv;
    if (out64) out64[i] = lv;
}

// ------------------------------------------------------------------------------------------------ host
static int fill_roi(RoiParams& p, const hd_roi_level* levels, int n_levels, int C, const float* rois, const int32_t* level_ids,
                    int64_t K, int PH, int PW, int sampling_ratio, int aligned, float* out, int32_t* argmax) {
    HD_CHECK_ARG(levels != nullptr && n_levels >= 1 && n_levels <= ROI_MAX_LEVELS, "n_levels must be in [1,%d], got %d", ROI_MAX_LEVELS, n_levels);
    HD_CHECK_ARG(C >= 1 && PH >= 1 && PW >= 1 && PH <= 64 && PW <= 64, "bad C=%d or output size %dx%d (max 64x64)", C, PH, PW);
    HD_CHECK_ARG(K >= 0, "K must be >= 0");
    HD_CHECK_ARG(n_levels == 1 || level_ids != nullptr || K == 0, "level_ids is NULL for a multi-level call");
    memset(&p, 0, sizeof(p));
    for (int l = 0; l < n_levels; ++l) {
        HD_CHECK_ARG(levels[l].H > 0 && levels[l].W > 0, "level %d has an empty feature map", l);
        HD_CHECK_ARG(K == 0 || levels[l].data != nullptr, "level %d data is NULL", l);
        p.data[l] = levels[l].data; p.H[l] = levels[l].H; p.W[l] = levels[l].W; p.scale[l] = levels[l].spatial_scale;
    }
    p.n_levels = n_levels; p.C = C; p.PH = PH; p.PW = PW; p.sampling_ratio = sampling_ratio; p.aligned = aligned;
    p.rois = rois; p.level_ids = level_ids; p.K = K; p.out = out; p.argmax = argmax;
    return HD_OK;
}

// 0 auto (= gather kernels: the staged-row kernel measured 2.5x slower on B200, see profiles/r1_results.md), 1 gather kernels
// only, 2 staged-row kernel whenever eligible; bits 4..7 are debug switches of the staged-row kernel (1: no tile store,
// 2: no compute, 8: nothing staged); bit 8: gather kernel with one table row in flight per thread (round-1 form, A/B)
static int g_roi_mode_ = 0;   // developer/test knob (process-wide, atomic)
extern "C" HD_API int hd_roi_set_mode(int mode) { return __atomic_exchange_n(&g_roi_mode_, mode, __ATOMIC_ACQ_REL); }

int hd_roi_mode(void) { return __atomic_load_n(&g_roi_mode_, __ATOMIC_ACQUIRE); }

static int launch_roi(const RoiParams& p, int layout, bool pool, cudaStream_t st) {
    if (p.K == 0) return HD_OK;
    const int g_roi_mode = __atomic_load_n(&g_roi_mode_, __ATOMIC_ACQUIRE);
    HD_CHECK_ARG(p.rois && p.out, "rois/out is NULL");
    if (layout == HD_LAYOUT_NHWC) {
        size_t tile = (size_t)p.C * p.PH * p.PW * 4;
        // fixed sampling_ratio <= 8: at most 16 entries per bin -> small tables leave more of the SM's 228 KB to L1
        int mx = p.PH > p.PW ? p.PH : p.PW;
        int tab = (!pool && p.sampling_ratio > 0 && p.sampling_ratio <= 8) ? 16 * mx : ROI_TAB;
        if (tab > ROI_TAB) tab = ROI_TAB;
        size_t smem = tile + (pool ? 0 : 2 * (size_t)ROI_TAB * sizeof(AxisEntry));
        size_t smem_quad = tile + 2 * (size_t)tab * sizeof(AxisEntry);
        HD_CHECK_ARG(smem <= 220 * 1024, "C*PH*PW=%d floats exceed the shared-memory tile", p.C * p.PH * p.PW);
        HD_CHECK_ARG(p.K < (1ll << 31), "too many RoIs");
        int use_tma = (tile % 16 == 0) && (((uintptr_t)p.out & 15) == 0);
        HD_ENSURE_SMEM(roi_align_nhwc_kernel, 220 * 1024);
        HD_ENSURE_SMEM(roi_pool_nhwc_kernel, 220 * 1024);
        HD_ENSURE_SMEM(roi_align_nhwc_quad_kernel<4>, 220 * 1024);
        HD_ENSURE_SMEM(roi_align_nhwc_quad_kernel<8>, 220 * 1024);
        HD_ENSURE_SMEM(roi_align_nhwc_quad_kernel<16>, 220 * 1024);
        HD_ENSURE_SMEM(roi_align_nhwc_quad_kernel<0>, 220 * 1024);
        HD_ENSURE_SMEM(roi_pool_nhwc_quad_kernel, 220 * 1024);
        int threads = p.C >= 256 ? 256 : ((p.C + 31) / 32 * 32 < 128 ? 128 : (p.C + 31) / 32 * 32);
        for (int l = 0; l < p.n_levels; ++l) HD_CHECK_ARG((long long)p.H[l] * p.W[l] * p.C < (1ll << 31), "level %d: H*W*C must be < 2^31", l);
        bool quad = (p.C % 4 == 0);
        for (int l = 0; l < p.n_levels; ++l) quad = quad && (((uintptr_t)p.data[l] & 15) == 0);
        int nq = p.C / 4, QT = 8;               // threads per bin group: power of two >= channel quads, in [8,256]
        while (QT < nq && QT < 256) QT <<= 1;
        if (!pool && quad && (g_roi_mode & 15) != 1 && p.sampling_ratio >= 1 && p.sampling_ratio <= 2 && p.PH <= RR_MAXP && p.PW <= RR_MAXP &&
            (g_roi_mode & 15) == 2) {
            // staged-row persistent kernel: tile | ring of row slots | geometry records + mbarriers
            const size_t tile_b = hd_align_up(tile, 128), slot_b = hd_align_up((size_t)RR_SPAN * p.C * 4, 128), meta_b = hd_align_up(sizeof(RrSmem), 128);
            const size_t budget = 227 * 1024;
            int n_slots = tile_b + meta_b < budget ? (int)((budget - tile_b - meta_b) / slot_b) : 0;
            if (n_slots > RR_MAXSLOTS) n_slots = RR_MAXSLOTS;
            if (n_slots >= 6 && use_tma) {
                HD_ENSURE_SMEM(roi_align_ring_kernel, budget);
                const size_t smem_ring = tile_b + (size_t)n_slots * slot_b + meta_b;
                const unsigned grid = (unsigned)(p.K < hd_num_sms() ? p.K : hd_num_sms());
                roi_align_ring_kernel<<<grid, RR_THREADS, smem_ring, st>>>(p, n_slots, (int)slot_b, (int)tile_b, (int)(tile_b + (size_t)n_slots * slot_b), g_roi_mode >> 4);
                HD_CUDA_LAUNCH_CHECK("roi_align_ring_kernel");
                return HD_OK;
            }
        }
        if (pool && quad) roi_pool_nhwc_quad_kernel<<<(unsigned)p.K, 256, smem, st>>>(p, use_tma, QT);
        else if (pool) roi_pool_nhwc_kernel<<<(unsigned)p.K, threads, smem, st>>>(p, use_tma);
        else if (quad && tab < ROI_TAB && p.sampling_ratio <= 2 && ((g_roi_mode >> 8) & 3) == 1)   // A/B: one table row in flight, 4 CTAs/SM
            roi_align_nhwc_quad_kernel<4><<<(unsigned)p.K, 256, smem_quad, st>>>(p, use_tma, QT, tab);
        else if (quad && tab < ROI_TAB && p.sampling_ratio <= 2 && p.PW >= 3 && p.PW <= 7 && ((g_roi_mode >> 8) & 3) != 2) {
            HD_ENSURE_SMEM(roi_align_nhwc_col_kernel, 220 * 1024);
            roi_align_nhwc_col_kernel<<<(unsigned)p.K, 32 * p.PW, smem_quad, st>>>(p, use_tma, tab, (g_roi_mode >> 4) & 15);
        }
        else if (quad && tab < ROI_TAB && p.sampling_ratio <= 2) {
            // two table rows (up to 8 independent 128-bit loads) in flight per thread, 3 CTAs/SM: the kernel is bound by the chain of
            // dependent load steps per thread, not by bandwidth -- measured 1.24 -> 1.15 ms on cfg3 (whole bin in flight, 2 CTAs/SM: 1.35 ms)
            HD_ENSURE_SMEM((roi_align_nhwc_quad_kernel<4, 2>), 220 * 1024);
            roi_align_nhwc_quad_kernel<4, 2><<<(unsigned)p.K, 256, smem_quad, st>>>(p, use_tma, QT, tab);
        }
        else if (quad && tab < ROI_TAB && p.sampling_ratio <= 4) roi_align_nhwc_quad_kernel<8><<<(unsigned)p.K, 256, smem_quad, st>>>(p, use_tma, QT, tab);
        else if (quad && tab < ROI_TAB) roi_align_nhwc_quad_kernel<16><<<(unsigned)p.K, 256, smem_quad, st>>>(p, use_tma, QT, tab);
        else if (quad) roi_align_nhwc_quad_kernel<0><<<(unsigned)p.K, 256, smem_quad, st>>>(p, use_tma, QT, tab);
        else roi_align_nhwc_kernel<<<(unsigned)p.K, threads, smem, st>>>(p, use_tma);
        HD_CUDA_LAUNCH_CHECK("roi_nhwc_kernel");
    } else if (layout == HD_LAYOUT_NCHW) {
        long long warps = p.K * p.C;
        long long blocks = (warps + 7) / 8;
        HD_CHECK_ARG(blocks < (1ll << 31), "grid too large");
        if (pool) roi_nchw_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(p);
        else roi_nchw_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(p);
        HD_CUDA_LAUNCH_CHECK("roi_nchw_kernel");
    } else {
        HD_FAIL(HD_ERR_INVALID, "layout must be HD_LAYOUT_NCHW or HD_LAYOUT_NHWC, got %d", layout);
    }
    return HD_OK;
}

int hd_roi_align_launch_list(const RoiParams& p0, const int* list, const int* list_count, int ctas, cudaStream_t st) {
    RoiParams p = p0;
    p.list = list; p.list_count = list_count;
    const size_t tile = (size_t)p.C * p.PH * p.PW * 4;
    const int mx = p.PH > p.PW ? p.PH : p.PW;
    int tab = (p.sampling_ratio > 0 && p.sampling_ratio <= 8) ? 16 * mx : ROI_TAB;
    if (tab > ROI_TAB) tab = ROI_TAB;
    const size_t smem_quad = tile + 2 * (size_t)tab * sizeof(AxisEntry);
    HD_CHECK_ARG(p.C % 4 == 0 && smem_quad <= 220 * 1024, "list launch needs C %% 4 == 0 and a tile that fits shared memory");
    const int use_tma = (tile % 16 == 0) && (((uintptr_t)p.out & 15) == 0);
    int nq = p.C / 4, QT = 8;
    while (QT < nq && QT < 256) QT <<= 1;
    HD_ENSURE_SMEM(roi_align_nhwc_quad_kernel<4>, 220 * 1024);
    HD_ENSURE_SMEM(roi_align_nhwc_quad_kernel<0>, 220 * 1024);
    if (tab < ROI_TAB && p.sampling_ratio >= 1 && p.sampling_ratio <= 2) roi_align_nhwc_quad_kernel<4><<<(unsigned)ctas, 256, smem_quad, st>>>(p, use_tma, QT, tab);
    else roi_align_nhwc_quad_kernel<0><<<(unsigned)ctas, 256, tile + 2 * (size_t)ROI_TAB * sizeof(AxisEntry), st>>>(p, use_tma, QT, ROI_TAB);
    HD_CUDA_LAUNCH_CHECK("roi_align_nhwc_quad_kernel (list)");
    return HD_OK;
}

// developer entry point: sliced kernel over a caller-provided (position-sorted) device list; see tools/roi_slice_probe.py
extern "C" HD_API int hd_debug_roi_align_sliced(const hd_roi_level* levels, int n_levels, int C, const float* rois, const int32_t* level_ids,
                                                int64_t K, int pooled_h, int pooled_w, int sampling_ratio, int aligned, float* out,
                                                const int32_t* list, const int32_t* list_count, int slices, int chunk, void* stream) {
    RoiParams p;
    int rc = fill_roi(p, levels, n_levels, C, rois, level_ids, K, pooled_h, pooled_w, sampling_ratio, aligned, out, nullptr);
    if (rc) return rc;
    HD_CHECK_ARG(sampling_ratio >= 1 && sampling_ratio <= 2 && slices >= 1 && chunk >= 1 && C % (4 * slices) == 0, "sliced kernel: sr in [1,2], C %% (4*slices) == 0");
    HD_CHECK_ARG(pooled_h <= 64 && pooled_w <= 64, "pooled size");
    if (K == 0) return HD_OK;
    p.list = list; p.list_count = list_count;
    const int Cs = C / slices;
    const size_t tile = (size_t)Cs * pooled_h * pooled_w * 4;
    HD_CHECK_ARG(tile % 16 == 0 && (((uintptr_t)out & 15) == 0), "tile alignment");
    const int mx = pooled_h > pooled_w ? pooled_h : pooled_w;
    const int tab = 16 * mx;
    const size_t smem = tile + 2 * (size_t)tab * sizeof(AxisEntry);
    HD_CHECK_ARG(smem <= 220 * 1024, "tile too large");
    int nq = Cs / 4, QT = 8;
    while (QT < nq && QT < 256) QT <<= 1;
    HD_ENSURE_SMEM(roi_align_sliced_kernel, 220 * 1024);
    const long long ctas = (K + chunk - 1) / chunk;
    roi_align_sliced_kernel<<<dim3((unsigned)ctas, (unsigned)slices), 256, smem, (cudaStream_t)stream>>>(p, QT, tab, chunk);
    HD_CUDA_LAUNCH_CHECK("roi_align_sliced_kernel");
    return HD_OK;
}

extern "C" HD_API int hd_roi_align(const hd_roi_level* levels, int n_levels, int layout, int C, const float* rois, const int32_t* level_ids,
                                   int64_t K, int pooled_h, int pooled_w, int sampling_ratio, int aligned, float* out, void* stream) {
    RoiParams p;
    int rc = fill_roi(p, levels, n_levels, C, rois, level_ids, K, pooled_h, pooled_w, sampling_ratio, aligned, out, nullptr);
    if (rc) return rc;
    return launch_roi(p, layout, false, (cudaStream_t)stream);
}

extern "C" HD_API int hd_roi_pool(const hd_roi_level* levels, int n_levels, int layout, int C, const float* rois, const int32_t* level_ids,
                                  int64_t K, int pooled_h, int pooled_w, float* out, int32_t* argmax, void* stream) {
    RoiParams p;
    int rc = fill_roi(p, levels, n_levels, C, rois, level_ids, K, pooled_h, pooled_w, 0, 0, out, argmax);
    if (rc) return rc;
    return launch_roi(p, layout, true, (cudaStream_t)stream);
}

extern "C" HD_API int hd_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, void* stream) {
    HD_CHECK_ARG(B >= 0 && C >= 1 && H >= 1 && W >= 1, "bad shape");
    if (B == 0) return HD_OK;
    HD_CHECK_ARG(in && out, "null pointer");
    HD_CHECK_ARG(B <= 65535 && (C + 31) / 32 <= 65535, "B or C too large");
    int HW = H * W;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
    nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, C, HW);
    HD_CUDA_LAUNCH_CHECK("nchw_to_nhwc_kernel");
    return HD_OK;
}

extern "C" HD_API int hd_roi_level_map(const float* rois, int roi_stride, int box_offset, int64_t K, int style, int k_min, int k_max,
                                       float canonical_scale, float canonical_level, float eps, int32_t* levels32, int64_t* levels64,
                                       void* stream) {
    HD_CHECK_ARG(K >= 0 && roi_stride >= 4 && box_offset >= 0 && box_offset + 4 <= roi_stride, "bad roi layout");
    HD_CHECK_ARG(k_max >= k_min, "k_max < k_min");
    if (K == 0) return HD_OK;
    HD_CHECK_ARG(rois && (levels32 || levels64), "null pointer");
    float s0 = canonical_scale;
    if (style == 1) s0 = canonical_scale / exp2f(canonical_level - (float)k_min);  // mmdet finest_scale (56 for 224/4/2)
    roi_level_map_kernel<<<(unsigned)((K + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rois, roi_stride, box_offset, K, style, k_min, k_max, s0,
                                                                                      canonical_level, eps, levels32, (long long*)levels64);
    HD_CUDA_LAUNCH_CHECK("roi_level_map_kernel");
    return HD_OK;
}

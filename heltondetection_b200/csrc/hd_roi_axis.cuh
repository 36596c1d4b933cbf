// Separable bilinear sampling tables of one RoI axis (shared by the RoIAlign kernels in roi.cu and roi_strip.cu).
#pragma once
#include "hd_common.cuh"

struct AxisEntry { int off; float w; };  // off = cell index premultiplied by the element stride of the axis

// Per-axis table of one RoI: bin b owns entries [b*stride, b*stride + cnt[b]).  Sample positions grow
// monotonically inside a bin, so a cell that was already emitted is one of the last two entries.
__device__ __forceinline__ void build_axis(AxisEntry* tab, int* cnt, int b, int stride, float start, float bin_size, int grid,
                                           int extent, int elem_stride, int pad_to) {
    int n = 0;
    AxisEntry* t = tab + b * stride;
    auto add = [&](int cell, float w) {
        const int off = cell * elem_stride;
        if (n >= 1 && t[n - 1].off == off) t[n - 1].w = __fadd_rn(t[n - 1].w, w);
        else if (n >= 2 && t[n - 2].off == off) t[n - 2].w = __fadd_rn(t[n - 2].w, w);
        else { t[n].off = off; t[n].w = w; ++n; }
    };
    for (int i = 0; i < grid; ++i) {
        // y = roi_start + ph*bin_size + (iy + .5f) * bin_size / grid   (fp32, left to right)
        float y = __fadd_rn(__fadd_rn(start, __fmul_rn((float)b, bin_size)),
                            __fdiv_rn(__fmul_rn(__fadd_rn((float)i, 0.5f), bin_size), (float)grid));
        if (y < -1.0f || y > (float)extent) continue;  // sample contributes 0 (C++/CUDA kernel rule)
        if (y <= 0.0f) y = 0.0f;
        int lo = (int)y, hi;
        if (lo >= extent - 1) { hi = lo = extent - 1; y = (float)lo; } else hi = lo + 1;
        const float l = __fsub_rn(y, (float)lo), h = __fsub_rn(1.0f, l);
        add(lo, h);
        add(hi, l);
    }
    cnt[b] = n;
    // pad with zero-weight duplicates of a real cell so the consumer can run fixed-trip-count loops
    const int dup = n > 0 ? t[n - 1].off : 0;
    for (int i = n; i < pad_to; ++i) { t[i].off = dup; t[i].w = 0.0f; }
}


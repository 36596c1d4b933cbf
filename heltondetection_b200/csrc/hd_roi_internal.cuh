// RoI kernel parameter block + the internal entry points shared by roi.cu and roi_strip.cu.
#pragma once
#include "hd_common.cuh"

#define ROI_MAX_LEVELS HD_MAX_LEVELS
#define ROI_TAB 1024  // table entries per axis

struct RoiParams {
    const float* data[ROI_MAX_LEVELS];
    int H[ROI_MAX_LEVELS], W[ROI_MAX_LEVELS];
    float scale[ROI_MAX_LEVELS];
    int n_levels, C, PH, PW, sampling_ratio, aligned;
    const float* rois;       // [K,5]
    const int* level_ids;    // [K] or NULL (level 0)
    long long K;
    float* out;              // [K,C,PH,PW]
    int* argmax;             // roi_pool only, nullable
    const int* list;         // nullable: the CTAs walk list[0 .. *list_count) instead of one RoI per CTA (roi_strip.cu hand-back)
    const int* list_count;
};

int hd_roi_mode(void);   // hd_roi_set_mode knob (roi.cu): 1 = per-RoI gather kernels only

// gather RoIAlign kernel (NHWC, channel quads) over the RoIs named by a device-side list; grid of `ctas` CTAs
int hd_roi_align_launch_list(const RoiParams& p, const int* list, const int* list_count, int ctas, cudaStream_t st);


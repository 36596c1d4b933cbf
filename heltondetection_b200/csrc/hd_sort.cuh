// CTA-wide LSD radix sort of (uint64 key, uint32 value) pairs, ascending, stable.
// One CTA sorts one segment (an image's candidates, an RPN top-k set, a WBF label group); the data
// lives in caller-provided ping-pong buffers (global workspace that stays L1/L2 resident, or shared
// memory).  8-bit digits; a pass whose digit is identical for every key (the usual case for the
// exponent bytes of scores in (0,1) and the high bytes of the tiebreak id) is skipped after its
// histogram.  Requires blockDim.x == NT, NT a multiple of 256 and <= 1024.
#pragma once
#include "hd_common.cuh"

template <int NT>
struct HdSortSmem {
    int hist[256];
    int bin_off[256];
    int wsum[8];
    int warp_cnt[NT / 32][256];
};

// returns 0 if the sorted data ended in (k0,v0), 1 if in (k1,v1)
template <int NT, typename KeyT = uint64_t>
__device__ int hd_cta_radix_sort(KeyT* k0, uint32_t* v0, KeyT* k1, uint32_t* v1, int n, HdSortSmem<NT>& sm,
                                 int first_byte = 0, int last_byte = (int)sizeof(KeyT) - 1) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < (NT / 32) * 256; i += NT) (&sm.warp_cnt[0][0])[i] = 0;
    int cur = 0;
    for (int byte = first_byte; byte <= last_byte; ++byte) {
        KeyT* kin = cur ? k1 : k0;
        uint32_t* vin = cur ? v1 : v0;
        KeyT* kout = cur ? k0 : k1;
        uint32_t* vout = cur ? v0 : v1;
        const int sh = byte * 8;
        if (tid < 256) sm.hist[tid] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += NT) atomicAdd(&sm.hist[(int)((kin[i] >> sh) & 255)], 1);
        __syncthreads();
        const bool uniform = sm.hist[(int)((kin[0] >> sh) & 255)] == n;
        if (uniform) {
            __syncthreads();
            continue;
        }
        // exclusive scan of the 256 bins
        int h = 0, incl = 0;
        if (tid < 256) {
            h = sm.hist[tid];
            incl = h;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int y = __shfl_up_sync(HD_FULL, incl, d);
                if (lane >= d) incl += y;
            }
            if (lane == 31) sm.wsum[wid] = incl;
        }
        __syncthreads();
        if (tid < 256) {
            int pre = 0;
            for (int w = 0; w < wid; ++w) pre += sm.wsum[w];
            sm.bin_off[tid] = pre + incl - h;
        }
        __syncthreads();
        // stable scatter, one tile of NT keys at a time
        for (int tb = 0; tb < n; tb += NT) {
            const int i = tb + tid;
            const bool act = i < n;
            KeyT key = 0;
            uint32_t val = 0;
            int dg = 256 + lane;  // unique sentinel for idle lanes
            if (act) {
                key = kin[i];
                val = vin[i];
                dg = (int)((key >> sh) & 255);
            }
            const unsigned peers = __match_any_sync(HD_FULL, dg);
            const int rank = __popc(peers & hd_lanemask_lt());
            if (act && rank == 0) sm.warp_cnt[wid][dg] = __popc(peers);
            __syncthreads();
            if (tid < 256) {
                int run = sm.bin_off[tid];
#pragma unroll 8
                for (int w = 0; w < NT / 32; ++w) {
                    int c = sm.warp_cnt[w][tid];
                    if (c) sm.warp_cnt[w][tid] = run;
                    run += c;
                }
                sm.bin_off[tid] = run;
            }
            __syncthreads();
            if (act) {
                const int pos = sm.warp_cnt[wid][dg] + rank;
                kout[pos] = key;
                vout[pos] = val;
            }
            __syncwarp();
            if (act && rank == 0) sm.warp_cnt[wid][dg] = 0;
            __syncthreads();
        }
        cur ^= 1;
    }
    __syncthreads();
    return cur;
}

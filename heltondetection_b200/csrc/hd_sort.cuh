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

// ------------------------------------------------------------------------------------------------------------
// Register-blocked CTA bitonic sort: N = 1024*E unique uint64 keys (+ optional uint32 payload), ascending.
// blockDim.x must be 1024.  Thread t owns elements [t*E, t*E+E) in registers, so compare-exchange distances
//   j <  E      stay inside a thread            (register min/max, no memory traffic),
//   j < 32*E    stay inside a warp              (one __shfl_xor per register),
//   j >= 32*E   cross warps                      (one round trip through shared memory; 15 of the 91 stages at N=8192).
// Keys start and end in skey[0..N) (sval[0..N)); pad with ~0ull.  ~10x faster than a naive all-shared-memory
// network, which is shared-memory-bandwidth bound; the LSD radix sort through global scratch stays as the path for
// segments that do not fit.
// ------------------------------------------------------------------------------------------------------------
template <int E, bool HAS_VAL>
__device__ void hd_cta_bitonic_reg(unsigned long long* skey, uint32_t* sval) {
    constexpr int N = 1024 * E;
    const int tid = threadIdx.x, lane = tid & 31;
    const int i0 = tid * E;
    unsigned long long k[E];
    uint32_t v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { k[e] = skey[i0 + e]; v[e] = HAS_VAL ? sval[i0 + e] : 0u; }
    auto cx_local = [&](int e, int f, bool up) {  // e < f, both in this thread
        if ((k[e] > k[f]) == up) {
            const unsigned long long tk = k[e]; k[e] = k[f]; k[f] = tk;
            if (HAS_VAL) { const uint32_t tv = v[e]; v[e] = v[f]; v[f] = tv; }
        }
    };
    // phase 1: kk <= E, everything inside the thread (directions fold at compile time except kk == E)
#pragma unroll
    for (int kk = 2; kk <= E; kk <<= 1) {
#pragma unroll
        for (int j = kk >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int e = 0; e < E; ++e)
                if ((e & j) == 0) cx_local(e, e | j, ((i0 + e) & kk) == 0);
        }
    }
    // phase 2: kk > E; all E elements of a thread share the direction
    for (int kk = 2 * E; kk <= N; kk <<= 1) {
        const bool up = (i0 & kk) == 0;
        for (int j = kk >> 1; j >= E; j >>= 1) {
            if (j >= 32 * E) {
                __syncthreads();   // partners of the previous shared-memory stage are done reading
                // staging layout [e][tid] (the buffer is free while the data lives in registers): consecutive lanes hit
                // consecutive words, and the partner thread tid ^ (j/E) differs only in its warp bits -> no bank conflicts
#pragma unroll
                for (int e = 0; e < E; ++e) { skey[e * 1024 + tid] = k[e]; if (HAS_VAL) sval[e * 1024 + tid] = v[e]; }
                __syncthreads();
                const bool lower = (i0 & j) == 0;
                const bool take_min = (lower == up);
                const int pt = tid ^ (j / E);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned long long pk = skey[e * 1024 + pt];
                    if (take_min ? (pk < k[e]) : (pk > k[e])) { k[e] = pk; if (HAS_VAL) v[e] = sval[e * 1024 + pt]; }
                }
            } else {
                const int lx = j / E;
                const bool lower = (lane & lx) == 0;
                const bool take_min = (lower == up);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const unsigned long long pk = __shfl_xor_sync(HD_FULL, k[e], lx);
                    const uint32_t pv = HAS_VAL ? __shfl_xor_sync(HD_FULL, v[e], lx) : 0u;
                    if (take_min ? (pk < k[e]) : (pk > k[e])) { k[e] = pk; if (HAS_VAL) v[e] = pv; }
                }
            }
        }
#pragma unroll
        for (int j = E >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int e = 0; e < E; ++e)
                if ((e & j) == 0) cx_local(e, e | j, up);
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e) { skey[i0 + e] = k[e]; if (HAS_VAL) sval[i0 + e] = v[e]; }
    __syncthreads();
}

// picks the smallest register-blocked network that holds n keys; returns N (keys beyond n must be padded with ~0ull up to N)
__device__ __forceinline__ int hd_bitonic_padded(int n) { return n <= 2048 ? 2048 : (n <= 4096 ? 4096 : 8192); }

// ------------------------------------------------------------------------------------------------------------
// Bucket sort of n <= NT*E UNIQUE uint64 keys whose high word is a score bit pattern (piecewise-log in the score, so a
// dense image's candidates spread over thousands of distinct values).  One counting pass over HD_BUCKET_BINS bins of the high
// word scaled to the image's [min, max]; shared-memory atomics hand out arbitrary places inside a bin; then every element
// ranks itself among the members of its bin by enumeration (lanes of a warp sit in the same bin: broadcast reads).
// O(n + sum bin^2 / NT) instead of the bitonic network's O(n log^2 n): 6.8 k candidates sort in ~10 us instead of 79 us.
// Element e of thread t is key k[e] with payload t + e*NT (its index), valid if that index < n.  On success sval[r] = index of
// the rank-r element and the function returns true; skey is scratch.  Returns false (nothing usable written) when one bin
// would hold more than HD_BUCKET_MAX elements (degenerate score distribution) -- the caller then runs the bitonic network.
// `bins` = 2*HD_BUCKET_BINS ints of shared scratch, `red` = 64 ints.
// ------------------------------------------------------------------------------------------------------------
#define HD_BUCKET_BINS 4096
#define HD_BUCKET_MAX 512
template <int NT, int E>
__device__ bool hd_cta_bucket_sort(const unsigned long long (&k)[E], int n, unsigned long long* skey, uint32_t* sval,
                                   int* bins, int* red) {
    static_assert(HD_BUCKET_BINS % NT == 0, "bins per thread");
    constexpr int BPT = HD_BUCKET_BINS / NT;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int* start = bins;
    int* cursor = bins + HD_BUCKET_BINS;
    // 1. range of the high word
    uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (tid + e * NT < n) { const uint32_t h = (uint32_t)(k[e] >> 32); lo = min(lo, h); hi = max(hi, h); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(HD_FULL, lo, d));
        hi = max(hi, __shfl_xor_sync(HD_FULL, hi, d));
    }
    if (lane == 0) { red[wid] = (int)lo; red[32 + wid] = (int)hi; }
#pragma unroll
    for (int q = 0; q < BPT; ++q) start[tid * BPT + q] = 0;
    __syncthreads();
    lo = (uint32_t)red[lane % (NT / 32)];
    hi = (uint32_t)red[32 + lane % (NT / 32)];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = min(lo, __shfl_xor_sync(HD_FULL, lo, d));
        hi = max(hi, __shfl_xor_sync(HD_FULL, hi, d));
    }
    const uint32_t range = hi - lo;
    const int shift = range < HD_BUCKET_BINS ? 0 : (32 - __clz(range)) - 12;   // (range >> shift) < 4096
    // 2. histogram
    int bin[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        bin[e] = -1;
        if (tid + e * NT < n) { bin[e] = (int)(((uint32_t)(k[e] >> 32) - lo) >> shift); atomicAdd(&start[bin[e]], 1); }
    }
    __syncthreads();
    // 3. exclusive scan (thread t owns bins t*BPT ..), bin-size check
    int c[BPT], s = 0, cmax = 0;
#pragma unroll
    for (int q = 0; q < BPT; ++q) { c[q] = start[tid * BPT + q]; s += c[q]; cmax = max(cmax, c[q]); }
    int incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(HD_FULL, incl, d);
        if (lane >= d) incl += y;
    }
    const bool too_big = __syncthreads_or(cmax > HD_BUCKET_MAX) != 0;   // (also orders the reads of red[] above before the writes below)
    if (too_big) return false;
    if (lane == 31) red[wid] = incl;
    __syncthreads();
    int wtot = (lane < NT / 32) ? red[lane] : 0, winc = wtot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(HD_FULL, winc, d);
        if (lane >= d) winc += y;
    }
    int run = __shfl_sync(HD_FULL, winc - wtot, wid) + incl - s;
#pragma unroll
    for (int q = 0; q < BPT; ++q) { start[tid * BPT + q] = run; cursor[tid * BPT + q] = run; run += c[q]; }
    __syncthreads();
    // 4. place (arbitrary order inside a bin)
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (bin[e] >= 0) {
            const int pos = atomicAdd(&cursor[bin[e]], 1);
            skey[pos] = k[e];
            sval[pos] = (uint32_t)(tid + e * NT);
        }
    __syncthreads();
    // 5. rank inside the bin; cursor[b] is now the end of bin b
    int dest[E];
    uint32_t pv[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int q = tid + e * NT;
        dest[e] = -1;
        if (q < n) {
            const unsigned long long kq = skey[q];
            pv[e] = sval[q];
            const int b = (int)(((uint32_t)(kq >> 32) - lo) >> shift);
            const int s0 = start[b], s1 = cursor[b];
            int r = s0;
            for (int j = s0; j < s1; ++j) r += (skey[j] < kq) ? 1 : 0;
            dest[e] = r;
        }
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (dest[e] >= 0) sval[dest[e]] = pv[e];
    __syncthreads();
    return true;
}

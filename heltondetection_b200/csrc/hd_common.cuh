// Shared device/host helpers for the hd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "../../include/hd_b200.h"

#define HD_MAX_DEVICES 64   // per-device caches below are indexed by the CUDA device ordinal

// ---------------------------------------------------------------- per-device state (api.cu)
// Function attributes (cudaFuncSetAttribute) and occupancy figures are properties of a (function, device) pair, and the
// library may drive several GPUs from one process and from several host threads: every cache is per device and atomic.
int hd_current_device(void);                 // cudaGetDevice, -1 on error
int hd_num_sms(void);                        // multiProcessorCount of the current device (cached per device; 148 on a B200)
struct HdDeviceOnce { unsigned long long done; };   // one bit per device; zero-initialised static
// raise MaxDynamicSharedMemorySize of `func` on the current device once per device; returns an hd error code
int hd_ensure_max_smem(HdDeviceOnce* once, const void* func, int bytes, const char* name);
#define HD_ENSURE_SMEM(kernel, bytes)                                                                      \
    do {                                                                                                   \
        static HdDeviceOnce once__ = {0ull};                                                               \
        int rc__ = hd_ensure_max_smem(&once__, (const void*)(kernel), (int)(bytes), #kernel);              \
        if (rc__) return rc__;                                                                             \
    } while (0)

// ---------------------------------------------------------------- host error plumbing
void hd_set_error(const char* fmt, ...);
#define HD_FAIL(code, ...)        \
    do {                          \
        hd_set_error(__VA_ARGS__); \
        return (code);            \
    } while (0)
#define HD_CHECK_ARG(cond, ...)                        \
    do {                                               \
        if (!(cond)) HD_FAIL(HD_ERR_INVALID, __VA_ARGS__); \
    } while (0)
void hd_count_launch(void);   // every kernel launch of the library is counted (hd_debug_launch_count)
#define HD_CUDA_LAUNCH_CHECK(name)                                                        \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) HD_FAIL(HD_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
        hd_count_launch();                                                                \
    } while (0)
#define HD_CUDA_CALL(x)                                                                       \
    do {                                                                                      \
        cudaError_t e__ = (x);                                                                \
        if (e__ != cudaSuccess) HD_FAIL(HD_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e__)); \
    } while (0)

// largest float <= (double) thr: (double)iou > thr  <=>  iou > hd_thr_floor(thr) for float iou.
// torchvision's CPU nms compares the fp32 IoU against the double threshold.
static inline float hd_thr_floor(double thr) {
    float f = (float)thr;
    if ((double)f > thr) f = nextafterf(f, -INFINITY);
    return f;
}
// smallest float >= (double) thr: (double)x >= thr <=> x >= hd_thr_ceil(thr)
static inline float hd_thr_ceil(double thr) {
    float f = (float)thr;
    if ((double)f < thr) f = nextafterf(f, INFINITY);
    return f;
}
static inline size_t hd_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
#define HD_FULL 0xffffffffu
// phase clocks of block 0 of the per-image kernels (developer aid, read back with hd_debug_phases)
static __device__ long long hd_dbg_phase[16];  // one copy per translation unit
#define HD_DEFINE_PHASE_READER(fn) int fn(long long* out16) { HD_CUDA_CALL(cudaDeviceSynchronize()); HD_CUDA_CALL(cudaMemcpyFromSymbol(out16, hd_dbg_phase, sizeof(long long) * 16)); return HD_OK; }
#define HD_PHASE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) hd_dbg_phase[i] = clock64(); } while (0)

// streaming 128-bit load: read-only path, do not allocate in L1 (each byte is touched once)
__device__ __forceinline__ float4 hd_ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float hd_ldg_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// sigmoid exactly as torch: 1 / (1 + exp(-x)), IEEE add/div, accurate expf (no fast-math)
__device__ __forceinline__ float hd_sigmoid(float x) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

// float -> uint32 whose unsigned order equals torch.sort(descending=False) order:
// -0 == +0, NaN greatest.  Descending keys are the bitwise complement.
__device__ __forceinline__ uint32_t hd_orderable(float f) {
    if (f != f) return 0xffffffffu;
    f = f + 0.0f;  // -0 -> +0
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float hd_from_orderable(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// c++ std::max / std::min semantics of the torchvision CPU kernel: (a < b) ? b : a
__device__ __forceinline__ float hd_stdmax(float a, float b) { return (a < b) ? b : a; }
__device__ __forceinline__ float hd_stdmin(float a, float b) { return (b < a) ? b : a; }

__device__ __forceinline__ float hd_area(float4 b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}
// IoU with the rounding of the CPU oracle: inter / ((areaA + areaB) - inter), no FMA contraction.
// `a` is the higher-scoring (kept) box, as in the CPU loop.
__device__ __forceinline__ float hd_iou(float4 a, float area_a, float4 b, float area_b) {
    float xx1 = hd_stdmax(a.x, b.x), yy1 = hd_stdmax(a.y, b.y);
    float xx2 = hd_stdmin(a.z, b.z), yy2 = hd_stdmin(a.w, b.w);
    float w = hd_stdmax(0.0f, __fsub_rn(xx2, xx1));
    float h = hd_stdmax(0.0f, __fsub_rn(yy2, yy1));
    float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

__device__ __forceinline__ uint32_t hd_lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
#endif  // __CUDACC__

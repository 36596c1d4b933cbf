// Error plumbing + version of the C ABI (include/hd_b200.h).
#include "hd_common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void hd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0ull;
void hd_count_launch(void) { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
// kernels launched (or captured into a graph) by this library since load, all threads and devices: callers difference it
extern "C" HD_API unsigned long long hd_debug_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int hd_current_device(void) {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; }
    return d;
}

int hd_num_sms(void) {
    static int cached[HD_MAX_DEVICES] = {0};   // benign race: every writer stores the same value
    const int d = hd_current_device();
    if (d >= 0 && d < HD_MAX_DEVICES && cached[d] > 0) return cached[d];
    int n = 0;
    if (d < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
    if (d >= 0 && d < HD_MAX_DEVICES) cached[d] = n;
    return n;
}

int hd_ensure_max_smem(HdDeviceOnce* once, const void* func, int bytes, const char* name) {
    const int d = hd_current_device();
    const unsigned long long bit = (d >= 0 && d < HD_MAX_DEVICES) ? (1ull << d) : 0ull;
    if (bit && (__atomic_load_n(&once->done, __ATOMIC_ACQUIRE) & bit)) return HD_OK;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) HD_FAIL(HD_ERR_CUDA, "cudaFuncSetAttribute(%s, %d B): %s", name, bytes, cudaGetErrorString(e));
    if (bit) __atomic_fetch_or(&once->done, bit, __ATOMIC_RELEASE);
    return HD_OK;
}

extern "C" HD_API const char* hd_last_error(void) { return g_err; }
extern "C" HD_API int hd_version(void) { return 100; }

int hd_phase_reader_nms(long long* out16);
int hd_phase_reader_rpn(long long* out16);
int hd_phase_reader_small(long long* out16);
// developer aid: phase clocks of block 0 of the last sort_nms_kernel (which=0) / rpn_select_nms_kernel (which=1); synchronises
extern "C" HD_API int hd_debug_phases(int which, long long* out16) {
    HD_CHECK_ARG(out16 != nullptr, "out16 is NULL");
    return which == 1 ? hd_phase_reader_rpn(out16) : (which == 2 ? hd_phase_reader_small(out16) : hd_phase_reader_nms(out16));
}

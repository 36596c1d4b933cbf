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

extern "C" HD_API const char* hd_last_error(void) { return g_err; }
extern "C" HD_API int hd_version(void) { return 100; }

int hd_phase_reader_nms(long long* out16);
int hd_phase_reader_rpn(long long* out16);
// developer aid: phase clocks of block 0 of the last sort_nms_kernel (which=0) / rpn_select_nms_kernel (which=1); synchronises
extern "C" HD_API int hd_debug_phases(int which, long long* out16) {
    HD_CHECK_ARG(out16 != nullptr, "out16 is NULL");
    return which == 1 ? hd_phase_reader_rpn(out16) : hd_phase_reader_nms(out16);
}

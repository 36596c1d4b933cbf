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

// api.cu — error reporting and library identification for the C ABI (include/pcl_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace pcl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace pcl

extern "C" const char *pcl_last_error(void) { return pcl::g_err; }
extern "C" int pcl_version(void) { return 100; }
extern "C" int pcl_compiled_arch(void) { return 100; }

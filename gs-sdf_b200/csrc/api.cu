// Error string + version of the C ABI (include/gssdf_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace gssdf {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace gssdf

extern "C" const char *gssdf_last_error(void) { return gssdf::g_err; }
extern "C" const char *gssdf_version(void) { return "gssdf_b200 0.1 sm_100a"; }
extern "C" int32_t gssdf_abi_revision(void) { return GSSDF_ABI_REVISION; }

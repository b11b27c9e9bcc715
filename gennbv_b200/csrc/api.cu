// api.cu -- error reporting + ABI version of libgennbv_b200.
#include "common.cuh"

namespace gnbv {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace gnbv

extern "C" int gnbv_abi_version(void) { return GNBV_ABI_VERSION; }
extern "C" const char* gnbv_last_error(void) { return gnbv::g_err; }

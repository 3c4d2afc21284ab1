// Thread-local error text behind spvd_last_error() and the ABI version query.
#include <stdarg.h>

#include "common.cuh"

namespace spvd {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace spvd

extern "C" const char* spvd_last_error(void) { return spvd::g_err; }
extern "C" int spvd_abi_version(void) { return SPVD_ABI_VERSION; }

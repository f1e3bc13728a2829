// api.cu -- error plumbing and version for the C ABI (include/mdl_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mdl {
static thread_local char t_err[512] = {0};
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}
}  // namespace mdl

extern "C" int mdl_version(void) { return 100; }

extern "C" int mdl_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return MDL_ERR_ARG;
  strncpy(buf, mdl::t_err, n - 1);
  buf[n - 1] = 0;
  return MDL_OK;
}

extern "C" int64_t mdl_launch_count(void) { return mdl::g_launches.load(); }

// Library-wide entry points: version and thread-local error string.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "stx_common.cuh"

namespace stx {
namespace {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = !(getenv("STX_PDL") && atoi(getenv("STX_PDL")) == 0);
  return on;
}
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace stx

extern "C" int stx_version(void) { return STX_VERSION; }
extern "C" const char* stx_last_error_string(void) { return stx::g_err; }
extern "C" unsigned long long stx_launch_count(void) { return stx::g_launches.load(); }

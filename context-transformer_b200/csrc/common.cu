// common.cu — error plumbing and library-level entry points
#include "common.cuh"

#include <atomic>
#include <stdarg.h>

namespace ctx {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace ctx

extern "C" int ctx_version(void) { return 100; }
extern "C" const char* ctx_last_error(void) { return ctx::g_err; }
extern "C" unsigned long long ctx_launch_count(void) { return ctx::g_launches.load(); }

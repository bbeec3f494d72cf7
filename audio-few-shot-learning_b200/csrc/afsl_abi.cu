// Library-level pieces of the C ABI: version, thread-local error string, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "afsl_common.cuh"

namespace afsl {
namespace {
thread_local char g_error[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace afsl

extern "C" int afsl_version(void) { return AFSL_ABI_VERSION; }
extern "C" const char* afsl_last_error(void) { return afsl::g_error; }
extern "C" long long afsl_launch_count(void) { return afsl::g_launches.load(std::memory_order_relaxed); }

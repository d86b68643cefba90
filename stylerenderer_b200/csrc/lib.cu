// Library-level plumbing of the C ABI: version, thread-local error text, launch counter.
#include "common.cuh"
#include <atomic>

namespace sr {
namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace sr

extern "C" int sr_abi_version(void) { return 4; }
extern "C" const char *sr_last_error(void) { return sr::g_err; }
extern "C" int64_t sr_launch_count(void) { return sr::g_launches.load(std::memory_order_relaxed); }

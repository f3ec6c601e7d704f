// Error state, version and launch accounting shared by every entry point of the C ABI.
#include <atomic>

#include "common.cuh"

namespace aicam {

static thread_local std::string g_error;
static std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { g_error = msg; }
int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace aicam

extern "C" {

int aicam_version(void) { return 100; }
const char* aicam_last_error(void) { return aicam::g_error.c_str(); }
uint64_t aicam_launch_count(void) { return aicam::g_launches.load(); }

}  // extern "C"

// Library-level plumbing of the C ABI: error string, version, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/dm_b200.h"
#include "dm_common.h"

namespace dm {

std::atomic<long long> g_launch_count{0};

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code == 0 ? -1 : code;
}

}  // namespace dm

extern "C" const char* dm_last_error(void) { return dm::last_error_buf(); }
extern "C" int dm_version(void) { return 100; }
extern "C" long long dm_launch_count(void) { return dm::g_launch_count.load(std::memory_order_relaxed); }

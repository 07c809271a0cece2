// Library-wide state: ABI version, last-error string, launch counter.
#include "common.cuh"
#include <string.h>

namespace nextou {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace nextou

extern "C" {
int nextou_abi_version(void) { return NEXTOU_ABI_VERSION; }
const char* nextou_last_error(void) { return nextou::g_err; }
long long nextou_launch_count(void) { return nextou::g_launches.load(); }
}
